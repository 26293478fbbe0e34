"""p50 latency of the threshold-network byte surface (encrypt / decrypt / reencrypt, fhe.rs:594-779) through the C ABI."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhe_precompiles_b200 import FHE, pack


def p50(fn, n=300):
    for _ in range(10):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return round(ts[n // 2] * 1e3, 4)


pub = FHE.public_key_bytes()
res = {}
for kind, val in (("i64", pack.serialize_i64(-12345)), ("u64", pack.serialize_u64(77)), ("u256", pack.serialize_u256(1 << 200)), ("frac64", pack.serialize_frac64(3.25))):
    enc_in = pack.pack_two_arguments(val, bytes([1, 2, 3]))
    ct = getattr(FHE, f"encrypt_{kind}")(enc_in)
    res[f"encrypt_{kind}_ms"] = p50(lambda: getattr(FHE, f"encrypt_{kind}")(enc_in))
    dec_in = pack.pack_one_argument(ct)
    assert getattr(FHE, f"decrypt_{kind}")(dec_in) == val
    res[f"decrypt_{kind}_ms"] = p50(lambda: getattr(FHE, f"decrypt_{kind}")(dec_in))
    re_in = pack.pack_binary_operation(pub, ct, bytes([1, 2, 3]))
    res[f"reencrypt_{kind}_ms"] = p50(lambda: getattr(FHE, f"reencrypt_{kind}")(re_in))
print(json.dumps(res))
