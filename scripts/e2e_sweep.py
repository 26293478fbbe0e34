import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhe_precompiles_b200 import device as fdev
N=4096; n=4096
fdev.init(0)
a=torch.randint(0, 0xFFFFC4001, (n,2,2,N), dtype=torch.int64).pin_memory()
b=torch.randint(0, 0xFFFFC4001, (n,2,2,N), dtype=torch.int64).pin_memory()
out=torch.empty_like(a).pin_memory()
_, rk = fdev.parse_public_key(open(os.path.join(ROOT,"fhe_precompiles_b200/data/network.pub"),"rb").read())
for _ in range(2): fdev.mul_relin_host(a,b,rk,out)
t0=time.perf_counter()
for _ in range(5): fdev.mul_relin_host(a,b,rk,out)
dt=(time.perf_counter()-t0)/5
print(os.environ.get("FHE_B200_PIPE_CHUNK_OPS"), f"{n/dt:.0f} ops/s  H2D {2*n*131072/dt/1e9:.1f} GB/s")
