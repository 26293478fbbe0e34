"""Byte-surface throughput of a LARGE batch (fhe_b200_batch, mul_cipheri64_cipheri64, operands as SEAL writes them: libzstd level-3
frames; every call carries its own copy of the public key) under the engine's staging knobs.  The library reads the knobs at
load time, so every configuration runs in its own process:
    python scripts/byte_surface_bench.py                      # sweep (parent)
    python scripts/byte_surface_bench.py child <calls>        # one configuration, knobs from the environment
Checks the first and last results against the single-call symbol."""
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n: int, distinct: int = int(os.environ.get("BSB_DISTINCT", "256"))) -> None:
    from fhe_precompiles_b200 import FHE, _lib, pack

    L = _lib.lib()
    rng = np.random.default_rng(3)
    q = (0xFFFFEE001, 0xFFFFC4001)
    dt = b"sunscreen::types::bfv::signed::Signed,0.8.1,true"

    def blob() -> bytes:
        w = np.stack([rng.integers(0, q[l], 4096, dtype=np.uint64) for _ in range(2) for l in range(2)]).reshape(-1)
        out, ln = ctypes.c_void_p(), ctypes.c_int64()
        assert L.fhe_b200_write_ciphertext(w.ctypes.data, dt, ctypes.byref(out), ctypes.byref(ln)) == 0
        b = ctypes.string_at(out.value, ln.value)
        L.fhe_free(out)
        return b

    net_pub = FHE.public_key_bytes(b"")
    prev = L.fhe_b200_set_zstd_writer(0)
    packed = [pack.pack_binary_operation(net_pub, blob(), blob()) for _ in range(distinct)]
    L.fhe_b200_set_zstd_writer(prev)
    bufs = [(ctypes.c_char * len(p)).from_buffer_copy(p) for p in packed]
    op = L.fhe_b200_op_index(b"mul_cipheri64_cipheri64")
    arr = (_lib.BatchCall * n)()
    for i in range(n):
        arr[i].op, arr[i].bytes, arr[i].bytes_length = op, ctypes.cast(bufs[i % distinct], ctypes.c_void_p), len(packed[i % distinct])
    want = {k: FHE.mul_cipheri64_cipheri64(packed[k % distinct]) for k in (0, n - 1)}
    rates = []
    for rep in range(5):
        t0 = time.perf_counter()
        failed = L.fhe_b200_batch(arr, n, 0)
        dt_ = time.perf_counter() - t0
        assert failed == 0, failed
        for k, w in want.items():
            assert ctypes.string_at(arr[k].output, arr[k].output_length) == w, "batch result differs from the single call"
        for i in range(n):
            L.fhe_free(arr[i].output)
        rates.append(n / dt_)
    knobs = {k: os.environ.get(k, "default") for k in ("FHE_B200_DEVICE_ZSTD", "FHE_B200_HOST_INFLATE_PCT", "FHE_B200_BIG_TILE_OPS", "FHE_B200_ZSTD_WRITER")}
    print(json.dumps({"calls": n, "cores": os.cpu_count(), **knobs, "calls_per_s_by_rep": [round(r) for r in rates], "best": round(max(rates[1:]))}),
          flush=True)


def main() -> None:
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
        return
    sweeps = [
        (4096, {}),                                                                       # default: 16-call tiles, host inflate
        (4096, {"FHE_B200_BIG_TILE_OPS": "512"}),                                          # big tiles, host inflate only
        (4096, {"FHE_B200_BIG_TILE_OPS": "512", "FHE_B200_DEVICE_ZSTD": "2", "FHE_B200_HOST_INFLATE_PCT": "0"}),   # device inflate only
        (4096, {"FHE_B200_BIG_TILE_OPS": "512", "FHE_B200_DEVICE_ZSTD": "2"}),             # hybrid, 35 % on the host
        (4096, {"FHE_B200_BIG_TILE_OPS": "1024", "FHE_B200_DEVICE_ZSTD": "2", "FHE_B200_HOST_INFLATE_PCT": "50"}),
        (8192, {"FHE_B200_BIG_TILE_OPS": "1024", "FHE_B200_DEVICE_ZSTD": "2", "FHE_B200_HOST_INFLATE_PCT": "25"}),
    ]
    if len(sys.argv) > 1 and sys.argv[1] == "sweep3":  # the fourth-generation device decoder (zstd_plan3.cuh)
        sweeps = [(4096, {}), (4096, {"FHE_B200_ZSTD_WRITER": "1"})]
        for big in ("128", "256", "512"):
            for pct in ("0", "25", "50"):
                sweeps.append((8192, {"FHE_B200_BIG_TILE_OPS": big, "FHE_B200_DEVICE_ZSTD": "2", "FHE_B200_HOST_INFLATE_PCT": pct,
                                      "FHE_B200_DEVICE_ZSTD_MIN_FRAMES": "64", "FHE_B200_ZSTD_WRITER": "1"}))
    for n, env in sweeps:
        e = dict(os.environ)
        e.update(env)
        e.setdefault("FHE_B200_QUIET", "1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(n)], env=e, capture_output=True, text=True, timeout=900)
        print(r.stdout.strip() or ("FAILED " + json.dumps(env) + " " + r.stderr[-600:]), flush=True)


if __name__ == "__main__":
    main()
