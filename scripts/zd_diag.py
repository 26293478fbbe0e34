"""Device zstd inflate timing: batch-oriented pipeline (FHE_B200_ZSTD_TWO_PHASE=2, default), two-phase (1), one warp per frame (0)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import zstd_device_check as Z
from fhe_precompiles_b200 import _lib
from oracle import formats as F
L = _lib.lib(); z = F.zstd(); rng = np.random.default_rng(1)
q = (0xFFFFEE001, 0xFFFFC4001)
def ct(): return bytes(97) + np.stack([rng.integers(0, q[l], 4096, dtype=np.uint64) for _ in range(2) for l in range(2)]).tobytes()
def run(tag, frames, payloads):
    got, st, ms = Z.inflate(L, frames)
    ok = all((s != 1) or g == p for g, p, s in zip(got, payloads, st))
    print(tag, "ok", sum(1 for s in st if s == 1), "/", len(st), "kernel_ms", round(ms, 2), "frames/s", round(len(frames) / (ms * 1e-3)), "match", ok, flush=True)
p = ct()
cts = [ct() for _ in range(16)]
frs = [z.compress(c, 3) for c in cts]
t = bytes(rng.choice(list(b"abcdefgh \n"), size=Z.PAYLOAD).astype(np.uint8))
for mode in (sys.argv[1:] or ("3", "2", "1", "0")):
    os.environ["FHE_B200_ZSTD_TWO_PHASE"] = mode
    print("two_phase =", mode)
    run("warm", [frs[0]], [cts[0]])
    run("ct_l3 x1", [frs[0]], [cts[0]])
    run("text_l3 x1", [z.compress(t, 3)], [t])
    run("structured x1", [F.zstd_structured_frame(cts[0])], [cts[0]])
    for n in (32, 256, 1024, 4096, 8192) if mode != "0" else (32, 1024):
        run(f"ct_l3 x{n}", [frs[i % 16] for i in range(n)], [cts[i % 16] for i in range(n)])
