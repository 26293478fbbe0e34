"""A/B of the last (intra-warp) NTT exchange: shared memory + __syncwarp (default) against warp shuffles (FHE_B200_NTT_SHFL=1),
on the standalone limb-NTT kernel (BASELINE config 2 shape).  The knob is read once per process, so each arm is a child
process; both arms must produce identical bits (and equal the oracle on a sample).
    python scripts/shuffle_ab.py            # parent: runs both arms, prints one JSON object
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = 4096
MODULI = (0xFFFFEE001, 0xFFFFC4001, 0x1FFFFE0001, 0x1FFFFFFFFFFA4001, 0x1FFFFFFFFFF92001, 0x1FFFFFFFFFFDE001)


def child() -> None:
    import numpy as np
    import torch

    from fhe_precompiles_b200 import device as fdev
    from oracle import bfv

    fdev.init(0)
    res = {"shfl": os.environ.get("FHE_B200_NTT_SHFL", "0")}
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    for name, mods in (("q0_q1_P", [0, 1, 2]), ("b0_b1_msk", [3, 4, 5])):
        batch = 16384
        x = torch.empty((batch, 3, N), dtype=torch.int64, device="cuda")
        for l, m in enumerate(mods):
            x[:, l, :] = torch.randint(0, MODULI[m], (batch, N), generator=g, device="cuda", dtype=torch.int64)
        x0 = x[:2].clone()

        def timeit(inverse):
            best = 1e9
            for _ in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fdev.ntt_(x, mods, inverse=inverse)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best

        f = timeit(False)
        i = timeit(True)
        res[name] = {"fwd_Mlimb_per_s": 3 * batch / f / 1e3, "inv_Mlimb_per_s": 3 * batch / i / 1e3}
        y = x0.clone()
        fdev.ntt_(y, mods)
        yn = y.cpu().numpy().view(np.uint64)
        ok = all(np.array_equal(yn[b, l], bfv.ntt_fwd(x0[b, l].cpu().numpy().view(np.uint64), m)) for b in range(2) for l, m in enumerate(mods))
        fdev.ntt_(y, mods, inverse=True)
        res[name]["forward_equals_oracle"] = bool(ok)
        res[name]["roundtrip"] = bool(torch.equal(y, x0))
        res[name]["sha"] = hashlib.sha256(yn.tobytes()).hexdigest()[:16]
    print(json.dumps(res), flush=True)


def main() -> None:
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        return
    out = []
    for shfl in ("0", "1"):
        e = dict(os.environ, FHE_B200_NTT_SHFL=shfl)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=e, capture_output=True, text=True, timeout=600)
        out.append(json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"shfl": shfl, "error": r.stderr[-500:]})
    print(json.dumps({"arms": out, "identical_bits": all(out[0].get(k, {}).get("sha") == out[1].get(k, {}).get("sha") for k in ("q0_q1_P", "b0_b1_msk"))}))


if __name__ == "__main__":
    main()
