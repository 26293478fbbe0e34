"""Outer framing of precompile inputs: host-side mirror of /root/reference/src/pack.rs:119-266.

Same function names and argument meaning as the reference; arguments are already-serialised byte strings
(`fhe_serialize()` outputs: bincode for Ciphertext / PublicKey, big-endian scalars per pack.rs:47-104).
"""
from __future__ import annotations

import struct
from typing import Tuple


class FheError(Exception):
    """Mirror of lib.rs:3-44: `.code` is the i32 the C ABI returns."""

    MESSAGES = {
        1: "Unexpected end of file",
        2: "Platform architecture invalid",
        3: "Invalid encoding",
        4: "Overflow in FHE program",
        5: "Invalid decryption",
        6: "Invalid encryption",
        7: "Base sunscreen error",
    }

    def __init__(self, code: int, detail: str = "") -> None:
        self.code = code
        msg = self.MESSAGES.get(code, "Unknown error")
        super().__init__(f"{msg} (code {code})" + (f": {detail}" if detail else ""))


INDEX_SIZE = 4  # pack.rs:11 `type Index = u32`


def pack_one_argument(a: bytes) -> bytes:  # pack.rs:119-124
    return bytes(a)


def unpack_one_argument(data: bytes) -> bytes:  # pack.rs:126-131
    return bytes(data)


def pack_two_arguments(a: bytes, b: bytes) -> bytes:  # pack.rs:133-151
    return struct.pack(">I", len(a) + INDEX_SIZE) + bytes(a) + bytes(b)


def unpack_two_arguments(data: bytes) -> Tuple[bytes, bytes]:  # pack.rs:153-175
    if len(data) < INDEX_SIZE:
        raise FheError(1)
    (ix1,) = struct.unpack(">I", data[:INDEX_SIZE])
    if ix1 < INDEX_SIZE or ix1 > len(data):
        raise FheError(1)  # the reference panics on an out-of-range offset
    return data[INDEX_SIZE:ix1], data[ix1:]


def pack_nullary_operation(public_key: bytes) -> bytes:  # pack.rs:185-187
    return bytes(public_key)


def unpack_nullary_operation(data: bytes) -> bytes:  # pack.rs:197-199
    return bytes(data)


def pack_binary_operation(public_key: bytes, a: bytes, b: bytes) -> bytes:  # pack.rs:208-231
    ix1 = len(public_key) + 2 * INDEX_SIZE
    ix2 = ix1 + len(a)
    return struct.pack(">II", ix1, ix2) + bytes(public_key) + bytes(a) + bytes(b)


def unpack_binary_operation(data: bytes) -> Tuple[bytes, bytes, bytes]:  # pack.rs:238-266
    if len(data) < 2 * INDEX_SIZE:
        raise FheError(1)
    ix1, ix2 = struct.unpack(">II", data[: 2 * INDEX_SIZE])
    if ix1 < 2 * INDEX_SIZE or ix2 < ix1 or ix2 > len(data):
        raise FheError(1)
    return data[2 * INDEX_SIZE : ix1], data[ix1:ix2], data[ix2:]


# scalar FHESerialize impls, pack.rs:47-104
def serialize_u64(v: int) -> bytes:
    return struct.pack(">Q", v)


def serialize_u256(v: int) -> bytes:
    return int(v).to_bytes(32, "big")


def serialize_i64(v: int) -> bytes:
    return struct.pack(">q", v)


def serialize_frac64(v: float) -> bytes:
    return struct.pack(">d", v)


SERIALIZE = {"u64": serialize_u64, "u256": serialize_u256, "i64": serialize_i64, "frac64": serialize_frac64}


def deserialize_scalar(kind: str, b: bytes):
    try:
        if kind == "u64":
            return struct.unpack(">Q", b)[0]
        if kind == "i64":
            return struct.unpack(">q", b)[0]
        if kind == "frac64":
            return struct.unpack(">d", b)[0]
        if kind == "u256":
            if len(b) != 32:
                raise struct.error
            return int.from_bytes(b, "big")
    except struct.error:
        raise FheError(3) from None
    raise KeyError(kind)
