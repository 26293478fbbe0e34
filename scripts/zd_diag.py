import sys, os, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scripts')
import zstd_device_check as Z
from fhe_precompiles_b200 import _lib
from oracle import formats as F
L=_lib.lib(); z=F.zstd(); rng=np.random.default_rng(1)
q=(0xFFFFEE001,0xFFFFC4001)
def ct(): return bytes(97)+np.stack([rng.integers(0,q[l],4096,dtype=np.uint64) for _ in range(2) for l in range(2)]).tobytes()
def run(tag, frames, payloads):
    t0=time.time(); got,st,ms=Z.inflate(L,frames); dt=time.time()-t0
    ok=all((s!=1) or g==p for g,p,s in zip(got,payloads,st))
    print(tag, "status",st[:4], "kernel_ms",round(ms,2),"wall_s",round(dt,2),"match",ok, flush=True)
p=ct()
run("warm", [z.compress(p,3)],[p])
run("ct_l3 x1", [z.compress(p,3)],[p])
run("ct_l3 x32", [z.compress(p,3)]*32,[p]*32)
run("structured", [F.zstd_structured_frame(p)],[p])
for lvl in (-3,1,7,19):
    run(f"ct_l{lvl}", [z.compress(p,lvl)],[p])
t=bytes(rng.choice(list(b"abcdefgh \n"), size=Z.PAYLOAD).astype(np.uint8))
run("text_l3", [z.compress(t,3)],[t])
zz=bytes(Z.PAYLOAD)
run("zeros_l3", [z.compress(zz,3)],[zz])
r=rng.integers(0,256,Z.PAYLOAD,dtype=np.uint8).tobytes()
run("rand_l3", [z.compress(r,3)],[r])
run("ct_l3 x512", [z.compress(p,3)]*512,[p]*512)
