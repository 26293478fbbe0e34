"""Three threads run 2,304-call batches (big tiles + device zstd decoder) at once; results must equal the same batches run one after
the other and nothing may hang (nested host-pool loops, lanes growing concurrently).  Needs a GPU."""
import sys, os, threading, ctypes, numpy as np, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from helpers import KeySet, MODULI, N
from fhe_precompiles_b200 import FHE, _lib, pack
keys = KeySet.load(); L = _lib.lib(); rng = np.random.default_rng(5)
dt = b"sunscreen::types::bfv::signed::Signed,0.8.1,true"
def blob():
    w = np.stack([rng.integers(0, MODULI[l], N, dtype=np.uint64) for _ in range(2) for l in range(2)]).reshape(-1)
    out, ln = ctypes.c_void_p(), ctypes.c_int64()
    assert L.fhe_b200_write_ciphertext(w.ctypes.data, dt, ctypes.byref(out), ctypes.byref(ln)) == 0
    b = ctypes.string_at(out.value, ln.value); L.fhe_free(out); return b
cts = [blob() for _ in range(48)]
def mk(seed):
    r = np.random.default_rng(seed)
    return [(("mul","add","sub")[i % 3] + "_cipheri64_cipheri64", pack.pack_binary_operation(keys.pub_bytes, cts[int(r.integers(48))], cts[int(r.integers(48))])) for i in range(2304)]
batches = [mk(s) for s in (1, 2, 3)]
want = [FHE.run_batch(b) for b in batches]
got = [None] * 3
def run(i): got[i] = FHE.run_batch(batches[i])
t0 = time.time()
ts = [threading.Thread(target=run, args=(i,)) for i in range(3)]
[t.start() for t in ts]; [t.join(timeout=300) for t in ts]
assert all(not t.is_alive() for t in ts), "concurrent batches hung"
assert got == want
print("3 concurrent 2,304-call batches ok in", round(time.time() - t0, 2), "s")
