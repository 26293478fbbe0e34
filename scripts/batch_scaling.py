"""Byte-surface batch throughput (fhe_b200_batch, mul_cipheri64_cipheri64) against host threads.
Run per tile size: FHE_B200_TILE_OPS=1 python scripts/batch_scaling.py   (the library reads the knob at load time)."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhe_precompiles_b200 import FHE, _lib, pack  # noqa: E402


def main() -> None:
    L = _lib.lib()
    rng = np.random.default_rng(3)
    q = (0xFFFFEE001, 0xFFFFC4001)
    dt = b"sunscreen::types::bfv::signed::Signed,0.8.1,true"

    def blob(structured: bool) -> bytes:
        w = np.stack([rng.integers(0, q[l], 4096, dtype=np.uint64) for _ in range(2) for l in range(2)]).reshape(-1)
        prev = L.fhe_b200_set_zstd_writer(1 if structured else 0)
        out, n = ctypes.c_void_p(), ctypes.c_int64()
        assert L.fhe_b200_write_ciphertext(w.ctypes.data, dt, ctypes.byref(out), ctypes.byref(n)) == 0
        b = ctypes.string_at(out.value, n.value)
        L.fhe_free(out)
        L.fhe_b200_set_zstd_writer(prev)
        return b

    net_pub = FHE.public_key_bytes(b"")
    n = int(os.environ.get("BATCH_CALLS", "2048"))
    res = {"tile_ops": os.environ.get("FHE_B200_TILE_OPS", "default"), "calls": n, "cores": os.cpu_count()}
    for name, structured in (("seal_frames", False), ("structured_frames", True)):
        packed = pack.pack_binary_operation(net_pub, blob(structured), blob(structured))
        calls = [("mul_cipheri64_cipheri64", packed)] * n
        FHE.run_batch(calls, host_threads=0)
        for th in (1, 2, 4, 8, 16, 32, 64):
            FHE.run_batch(calls[: max(64, n // 4)], host_threads=th)
            t0 = time.perf_counter()
            r = FHE.run_batch(calls, host_threads=th)
            dtm = time.perf_counter() - t0
            assert all(st == 0 for st, _ in r)
            res[f"{name}_threads{th}"] = round(n / dtm)
        # small batches: latency of the C call alone (array prepared before, outputs freed after), default threads
        buf = (ctypes.c_char * len(packed)).from_buffer_copy(packed)
        op = L.fhe_b200_op_index(b"mul_cipheri64_cipheri64")
        for small in (1, 4, 16, 64, 256):
            arr = (_lib.BatchCall * small)()
            ts = []
            for _ in range(40):
                for i in range(small):
                    arr[i].op, arr[i].bytes, arr[i].bytes_length = op, ctypes.cast(buf, ctypes.c_void_p), len(packed)
                t0 = time.perf_counter()
                failed = L.fhe_b200_batch(arr, small, 0)
                ts.append(time.perf_counter() - t0)
                assert failed == 0
                for i in range(small):
                    L.fhe_free(arr[i].output)
            res[f"{name}_batch{small}_p50_ms"] = round(sorted(ts)[20] * 1e3, 3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
