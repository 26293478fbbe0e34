"""A/B of the persistent TMA-staged k_digit_ntt (FHE_B200_TMA=1: 3 CTAs per SM, =2: 2 CTAs per SM) against the default kernel:
per-kernel CUDA-event time of 4,096 multiply + relinearise ops and bit-equality of the results.  One child process per arm."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child() -> None:
    import torch

    import bench
    from fhe_precompiles_b200 import device as fdev

    fdev.init(0)
    dev = torch.device("cuda", 0)
    n = 4096
    a, b = bench.synth_ciphertexts(torch, n, 2, dev), bench.synth_ciphertexts(torch, n, 3, dev)
    _, rk_h = fdev.parse_public_key(open(os.path.join(ROOT, "fhe_precompiles_b200/data/network.pub"), "rb").read())
    rk = rk_h.to(dev)
    out = torch.empty_like(a)
    for _ in range(3):
        fdev.mul_relin(a, b, rk, out=out)
    torch.cuda.synchronize()
    fdev.set_kernel_timing(True)
    for _ in range(10):
        fdev.mul_relin(a, b, rk, out=out)
    kt = fdev.kernel_timing_report(0)
    fdev.set_kernel_timing(False)
    print(json.dumps({"tma": os.environ.get("FHE_B200_TMA", "0"), "us_per_op": {k: v[0] * 1e3 / (10 * n) for k, v in kt.items()},
                      "sha": hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for mode in ("0", "1", "2"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, FHE_B200_TMA=mode), capture_output=True,
                               text=True, timeout=600)
            print(r.stdout.strip().splitlines()[-1] if r.returncode == 0 else json.dumps({"tma": mode, "error": r.stderr[-400:]}), flush=True)
