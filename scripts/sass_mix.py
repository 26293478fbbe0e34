"""Instruction mix of the hot kernels from the shipped object (cuobjdump -sass build/kernels.o): opcodes per kernel, grouped by
the pipe they issue to, and the share of multiplier-pipe (FMA-heavy) cycles that is not a multiply.  Usage:
    python scripts/sass_mix.py [kernels.o] > profiles/r2_sass_mix.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fhe_precompiles_b200/csrc/build/kernels.o")
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
HOT = ["k_ext_conv_d", "k_ext_ntt_d", "k_tensor_intt_d", "k_floor_sk_d", "k_digit_ntt_ksd", "k_ks_intt_ksd", "k_ks_finish_ksd", "k_ext_conv", "k_ext_ntt2", "k_tensor_intt", "k_floor_sk", "k_digit_ntt",
       "k_ks_finish", "k_ks_intt", "k_ntt"]
# FMA-heavy pipe cycles per warp instruction (scripts/pipe_probe.cu: IMAD.WIDE / IMAD.HI quarter rate, other IMAD half rate)
def fma_cycles(op):
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"):
        return 4
    if op.startswith("IMAD"):
        return 2
    return 0
kern = None
mix = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        kern = next((h for h in HOT if re.search(r"\d+" + h + r"(I|E)", name)), None)
        if kern and kern in mix:  # template instances (k_ntt<false/true>): keep them apart
            kern = kern + "'"
        if kern:
            mix[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if kern and m:
        mix[kern][m.group(1)] += 1
print("# SASS instruction mix of the hot kernels (static counts, `scripts/sass_mix.py`)\n")
print("Multiplier-pipe cycles: IMAD.WIDE / IMAD.HI = 4 per warp instruction, every other IMAD* = 2 (measured rates, DESIGN.md section 4).")
print("`not a multiply` = IMAD.MOV / IMAD.SHL / IMAD.X / IMAD.IADD: moves, shifts and carry adds that ptxas places on the multiplier pipe.\n")
print("| kernel | SASS instructions | IMAD.WIDE* | IMAD (mul) | IMAD.MOV | IMAD.SHL | IMAD.X | IMAD.IADD | ALU (IADD3/LOP3/SHF/SEL/ISETP/...) | LDS/STS | LDG/STG | FMA-pipe cycles | not a multiply |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, c in mix.items():
    tot = sum(c.values())
    wide = sum(v for o, v in c.items() if o.startswith("IMAD.WIDE") or o.startswith("IMAD.HI"))
    mov = sum(v for o, v in c.items() if o.startswith("IMAD.MOV"))
    shl = sum(v for o, v in c.items() if o.startswith("IMAD.SHL"))
    xx = sum(v for o, v in c.items() if o.startswith("IMAD.X"))
    iadd = sum(v for o, v in c.items() if o.startswith("IMAD.IADD"))
    imad = sum(v for o, v in c.items() if o.startswith("IMAD")) - wide - mov - shl - xx - iadd
    alu = sum(v for o, v in c.items() if re.match(r"(IADD3|LOP3|SHF|SEL|ISETP|LEA|PRMT|MOV|IABS|IMNMX|VIADD|UIADD|ULOP|USHF|UMOV|PLOP|P2R|R2P)", o))
    lds = sum(v for o, v in c.items() if o.startswith("LDS") or o.startswith("STS"))
    ldg = sum(v for o, v in c.items() if re.match(r"(LDG|STG|LDC|ULDC|LD\.|ST\.)", o))
    cyc = sum(fma_cycles(o) * v for o, v in c.items())
    waste = 2 * (mov + shl + xx + iadd)
    print(f"| `{k}` | {tot} | {wide} | {imad} | {mov} | {shl} | {xx} | {iadd} | {alu} | {lds} | {ldg} | {cyc} | {waste} ({100.0 * waste / max(cyc, 1):.1f} %) |")
print("\nThe NTT kernels contain one straight-line transform per modulus (switch over the limb), so the counts are sums over 3-6 bodies;")
print("ratios are what matters.  Per butterfly (one 4096-point transform = 24,576 butterflies over 512 threads = 48 per thread):\n")
