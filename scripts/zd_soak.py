"""Differential soak of the device zstd decoder (fhe_b200_zstd_inflate, default generation) against the payloads the frames were
made from: uniform ciphertext payloads, and adversarial ones (runs of equal residues, small residues, sparse polynomials, repeated
blocks - long and overlapping matches, RLE / Huffman literals), at several libzstd levels.  A frame the device accepts must
reproduce its payload byte for byte; the share it hands back is reported.  Needs a GPU.
    python scripts/zd_soak.py [frames=8192] [seed=1]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import zstd_device_check as Z
from fhe_precompiles_b200 import _lib
from oracle import formats as F  # checker only: libzstd writes the frames

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = _lib.lib(); z = F.zstd(); rng = np.random.default_rng(seed)
q = (0xFFFFEE001, 0xFFFFC4001)

def payload(kind: int) -> bytes:
    w = np.stack([rng.integers(0, q[l], 4096, dtype=np.uint64) for _ in range(2) for l in range(2)]).reshape(-1)
    if kind == 1:    # runs of equal residues
        w = np.repeat(w[:: int(rng.integers(2, 64))], 64)[: 16384]
    elif kind == 2:  # small residues (short words: other literal / match split)
        w = w >> np.uint64(int(rng.integers(8, 34)))
    elif kind == 3:  # sparse
        w = w * (rng.integers(0, int(rng.integers(2, 16)), 16384) == 0).astype(np.uint64)
    elif kind == 4:  # repeated blocks
        b = int(rng.integers(3, 700))
        w = np.tile(w[:b], 16384 // b + 1)[:16384]
    elif kind == 5:  # top nibbles constant (one long chain)
        w = (w & np.uint64(0xFFFFFFFF)) | np.uint64(int(rng.integers(0, 16)) << 32)
    return bytes(rng.integers(0, 256, 97, dtype=np.uint8)) + np.ascontiguousarray(w, dtype=np.uint64).tobytes()

t0 = time.time()
ok = back = wrong = 0
by_kind = {}
batch = 2048
done = 0
while done < n:
    m = min(batch, n - done)
    kinds = [0 if i % 3 else int(rng.integers(1, 6)) for i in range(m)]
    pays = [payload(k) for k in kinds]
    lvls = [(3, 3, 3, 1, 7, -1)[int(rng.integers(0, 6))] for _ in range(m)]
    frames = [z.compress(p, l) for p, l in zip(pays, lvls)]
    got, st, ms = Z.inflate(L, frames)
    for g, p, s, k in zip(got, pays, st, kinds):
        c = by_kind.setdefault(k, [0, 0, 0])
        if s == 1 and g == p: ok += 1; c[0] += 1
        elif s == 1: wrong += 1; c[2] += 1
        else: back += 1; c[1] += 1
    done += m
    print("frames", done, "ok", ok, "handed back", back, "WRONG", wrong, "kernel_ms", round(ms, 2), flush=True)
print(json.dumps({"frames": n, "seed": seed, "accepted_and_identical": ok, "handed_back": back, "wrong": wrong,
                  "by_kind_ok_back_wrong": by_kind, "seconds": round(time.time() - t0, 1)}))
sys.exit(1 if wrong else 0)
