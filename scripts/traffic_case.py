"""One multiply+relinearise pass over 1184 ops (after a warm-up pass), for DRAM-traffic measurement under
`ncu --cache-control none` (L2 state carried from producer to consumer kernels as in a real run)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhe_precompiles_b200 import device as d
N = 4096
d.init(0)
n = 1184
a = torch.randint(0, 0xFFFFC4001, (n, 2, 2, N), dtype=torch.int64, device="cuda")
b = torch.randint(0, 0xFFFFC4001, (n, 2, 2, N), dtype=torch.int64, device="cuda")
_, rk = d.parse_public_key(open(os.path.join(ROOT, "fhe_precompiles_b200/data/network.pub"), "rb").read())
rk = rk.cuda()
out = torch.empty_like(a)
for _ in range(2):
    d.mul_relin(a, b, rk, out=out)
torch.cuda.synchronize()
print("done")
