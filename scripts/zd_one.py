"""one level-3 ciphertext frame through the device inflate kernel (profiling target)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import zstd_device_check as Z
from fhe_precompiles_b200 import _lib
from oracle import formats as F
L = _lib.lib(); z = F.zstd(); rng = np.random.default_rng(1)
q = (0xFFFFEE001, 0xFFFFC4001)
p = bytes(97) + np.stack([rng.integers(0, q[l], 4096, dtype=np.uint64) for _ in range(2) for l in range(2)]).tobytes()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
got, st, ms = Z.inflate(L, [z.compress(p, 3)] * n)
print(st[:2], ms, got[0] == p)
