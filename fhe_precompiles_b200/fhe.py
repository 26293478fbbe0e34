"""`FheApp`: host-side mirror of the reference's precompile object (/root/reference/src/fhe.rs:56-780).

Every method has the reference's name and contract -- packed bytes in (pack.rs framing), serialised result
out, `FheError` with the lib.rs code on failure -- and is a thin call through the C ABI
(include/fhe_precompiles_b200.h) into the CUDA engine.  `FHE` is the process-wide instance, like
`testnet::one::FHE` (testnet.rs:25).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Sequence, Tuple

from . import _lib
from .pack import FheError


def _call(name: str, data: bytes) -> bytes:
    L = _lib.lib()
    out = ctypes.c_void_p()
    out_len = ctypes.c_int64()
    # `bytes` is passed by reference (ctypes hands the C side a pointer into the object): no copy of the ~590 KB input
    rc = getattr(L, "c_fhe_" + name)(data if data else None, len(data), ctypes.byref(out), ctypes.byref(out_len))
    if rc != 0:
        raise FheError(rc, _lib.last_error() if rc == 7 else "")
    try:
        return ctypes.string_at(out.value, out_len.value)
    finally:
        L.fhe_free(out)


def set_zstd_writer(structured: bool) -> bool:
    """Extension: how zstd-mode results are written -- True (default) = structure-aware standard zstd frames (82 KB, ~30 us),
    False = libzstd level 3 as SEAL's default does (~88.5 KB, ~1 ms). Returns the previous setting."""
    return bool(_lib.lib().fhe_b200_set_zstd_writer(1 if structured else 0))


def call_breakdown(enable: bool = True) -> dict:
    """Extension: phase timing of this thread's last binary precompile call: microseconds spent on framing + key
    lookup, operand parse + inflate, H2D, kernels, D2H, result encode, and in total. `enable` switches the recording on or off
    for subsequent calls (it makes single calls launch kernel by kernel instead of replaying a CUDA graph)."""
    L = _lib.lib()
    us = (ctypes.c_double * 7)()
    L.fhe_b200_last_call_breakdown(us)
    L.fhe_b200_set_call_timing(1 if enable else 0)
    keys = ("unpack_key_us", "parse_inflate_us", "h2d_us", "kernels_us", "d2h_us", "encode_us", "total_us")
    return dict(zip(keys, us))


class FheApp:
    """The 49 precompiles of fhe.rs:161-779 as methods; see `_lib.PRECOMPILES` for the list."""

    def public_key_bytes(self, _input: bytes = b"") -> bytes:
        return _call("public_key_bytes", b"")

    def run_batch(self, calls: Sequence[Tuple[str, bytes]], host_threads: int = 0) -> List[Tuple[int, bytes]]:
        """Extension: runs independent precompile calls as one batch (fhe_b200_batch).

        `calls` = [(precompile name, packed input)]; returns [(status, output bytes)] in order.
        """
        L = _lib.lib()
        n = len(calls)
        if n == 0:
            return []
        arr = (_lib.BatchCall * n)()
        keep = {}
        for i, (name, data) in enumerate(calls):
            idx = L.fhe_b200_op_index(name.encode())
            if idx < 0:
                raise KeyError(name)
            buf = keep.get(id(data))  # the same bytes object may back many calls
            if buf is None:
                buf = (ctypes.c_char * max(len(data), 1)).from_buffer_copy(data or b"\0")
                keep[id(data)] = buf
            arr[i].op = idx
            arr[i].bytes = ctypes.cast(buf, ctypes.c_void_p)
            arr[i].bytes_length = len(data)
        L.fhe_b200_batch(arr, n, host_threads)
        res = []
        for i in range(n):
            if arr[i].status == 0:
                res.append((0, ctypes.string_at(arr[i].output, arr[i].output_length)))
                L.fhe_free(arr[i].output)
            else:
                res.append((arr[i].status, b""))
        return res


def _make(name: str):
    def method(self, input: bytes) -> bytes:  # noqa: A002 - the reference's parameter name
        return _call(name, bytes(input))

    method.__name__ = name
    method.__doc__ = f"Precompile `{name}` (reference: FheApp::{name}, fhe.rs:161-779; C symbol c_fhe_{name})."
    return method


for _name in _lib.PRECOMPILES:
    if _name != "public_key_bytes":
        setattr(FheApp, _name, _make(_name))

FHE = FheApp()
