"""Turns a gpurun_out/<tag>_prof.ncu-rep + <tag>_launches.csv pair into the committed evidence under profiles/:
  profiles/<tag>_kernels.md       per-kernel table (duration, DRAM traffic, pipe utilisation, stalls)
  profiles/<tag>_launches.csv     the launch list of the bench command (name, duration)
  profiles/ncu_traffic.json       DRAM bytes per op per kernel (read by bench.py for roofline.traffic)
usage: python scripts/summarize_ncu.py <tag> <ops per launch>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, ops = sys.argv[1], int(sys.argv[2])
rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def f(r, name):
    v = r[idx[name]].replace(",", "") if name in idx else ""
    try:
        return float(v)
    except ValueError:
        return float("nan")


def to_bytes(r, name):
    v, u = f(r, name), units[idx[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


stall_names = [h for h in hdr if "issue_stalled" in h and h.endswith(".ratio")]
seen, lines, traffic, heavy = {}, [], {}, {}
for r in rows[2:]:
    k = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("fheb::", "")
    if k in seen:
        continue
    seen[k] = 1
    dur_us = f(r, "gpu__time_duration.sum") * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(units[idx["gpu__time_duration.sum"]], 1)
    dram = to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    traffic[k] = dram / ops
    heavy[k] = f(r, 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed')
    stalls = sorted(((f(r, n), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for n in stall_names), reverse=True)[:4]
    lines.append(
        f"| {k} | {r[idx['Grid Size']]} | {dur_us:.1f} | {dur_us / ops:.3f} | {dram / 1e6:.1f} | {dram / ops / 1e3:.0f} | "
        f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
        f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
        f"{f(r, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.1f} | "
        f"{f(r, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
        f"{f(r, 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
        f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {int(f(r, 'launch__registers_per_thread'))} | "
        + ", ".join(f"{n} {v:.2f}" for v, n in stalls)
        + " |"
    )
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_kernels.md"), "w") as out:
    out.write(f"# ncu --set full, tag {tag}: one launch of each kernel, {ops} ct x ct ops per launch (B200, clocks not locked)\n\n")
    out.write("Source: `scripts/gpu_profile.sh` (bench.py under ncu after the same command exited 0 without it).\n"
              "issue% = smsp__issue_active; ALU% = sm__inst_executed_pipe_alu (peak 0.5 warp-inst/clk/SMSP); FMA% = sm__pipe_fma_cycles_active\n"
              "(heavy + lite halves averaged); FMA-heavy% = sm__pipe_fmaheavy_cycles_active -- IMAD / IMAD.WIDE issue only to the heavy half, so\n"
              "this is the integer-multiply pipe's utilisation and the limiter of every multiply kernel;\n"
              "stalls = warps stalled per issue (top 4).\n\n")
    out.write("| kernel | grid | us/launch | us/op | DRAM MB/launch | DRAM KB/op | DRAM % | issue % | ALU % | FMA cyc % | FMA-heavy % | warps % | regs | top stalls |\n")
    out.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    out.write("\n".join(lines) + "\n")
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as out:
    json.dump({"tag": tag, "ops_per_launch": ops, "dram_bytes_per_op": traffic, "fmaheavy_pct": heavy,
               "dram_bytes_per_op_total": sum(traffic.values())}, out, indent=1)
# launch list
src = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
if os.path.exists(src):
    text = [l for l in open(src) if not l.startswith("==")]
    rd = list(csv.reader(text))
    h = rd[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot = {}
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as out:
        out.write("launch,kernel,gpu__time_duration_us\n")
        for n, r in enumerate(rd[1:]):
            us = float(r[vi].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(r[ui], 1)
            name = r[ki].split("(")[0].replace("void ", "").replace("fheb::", "")
            out.write(f"{n},{name},{us:.2f}\n")
            tot[name] = tot.get(name, 0) + us
    ours = {k: v for k, v in tot.items() if k.startswith("k_") and "_peak" not in k}
    s = sum(ours.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_kernels.md"), "a") as out:
        out.write("\n## Share of the step from the ncu launch list (cold-cache, serialised; compare shares, not absolutes)\n\n| kernel | total us | share |\n|---|---|---|\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
            out.write(f"| {k} | {v:.0f} | {v / s:.3f} |\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_kernels.md")).read())
