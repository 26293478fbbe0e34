"""TEST INFRASTRUCTURE ONLY -- wire-format oracle (pure Python + ctypes libzstd).

Independent restatement of the byte formats on the precompile surface, used by
tests/ (and __graft_entry__.smoke / bench.py's cpu_baseline leg) to build inputs and
to check what the CUDA-backed C-ABI library emits.  The product never imports it.

What it follows
---------------
* outer framing ............ /root/reference/src/pack.rs:119-266
  (pack_one_argument, pack_two_arguments, pack_binary_operation and inverses)
* scalar operands .......... /root/reference/src/pack.rs:47-104 (big-endian bytes)
* inner blobs .............. bincode 1.3.3 default config (Cargo.toml:10) over
  sunscreen 0.8.1 serde types (Cargo.toml:16; call sites pack.rs:23,30,36,43 and
  fhe.rs:29,121-122) wrapping Microsoft SEAL 4.0 `save()` streams compressed with
  zstd.  sunscreen / SEAL are NOT vendored in the reference; the layout below was
  established from the reference's four key fixtures (every byte of each fixture is
  consumed by this grammar -- see tests/test_formats.py) and SEAL 4.0's published
  serialization format.

Pinning: key-file layout, SEAL headers, parms_id rule and zstd level are pinned by the
fixtures (src/data/network.{pub,pri}, tests/data/{public,private}_key.bin).  The
`Ciphertext.data_type` string has no fixture in the reference ("parity unpinned" for
that one field); it is treated as an opaque length-prefixed string everywhere.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import hashlib
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------- params
# /root/reference/src/testnet.rs:8-14
N = 4096
Q0 = 0xFFFFEE001
Q1 = 0xFFFFC4001
P_SPECIAL = 0x1FFFFE0001
COEFF_MODULUS = (Q0, Q1, P_SPECIAL)
T_PLAIN = 4096
SCHEME_BFV_SERDE = 0  # sunscreen SchemeType::Bfv as bincode u32
SECURITY_TC128_SERDE = 0
SEAL_SCHEME_BFV = 1  # SEAL scheme_type::bfv, used in the parms_id hash

SEAL_MAGIC = 0xA15E
SEAL_HEADER_SIZE = 16
COMPR_NONE, COMPR_ZLIB, COMPR_ZSTD = 0, 1, 2

# Type-name strings sunscreen puts in `Ciphertext.data_type` (recollection, unpinned).
TYPE_NAMES = {
    "i64": "sunscreen::types::bfv::signed::Signed",
    "u64": "sunscreen::types::bfv::unsigned::Unsigned<1>",
    "u256": "sunscreen::types::bfv::unsigned::Unsigned<4>",
    "frac64": "sunscreen::types::bfv::fractional::Fractional<64>",
}
SUNSCREEN_VERSION = "0.8.1"


def data_type_string(kind: str) -> str:
    return f"{TYPE_NAMES[kind]},{SUNSCREEN_VERSION},true"


# --------------------------------------------------------------------------- zstd
class _Zstd:
    def __init__(self) -> None:
        lib = None
        for name in ("libzstd.so.1", ctypes.util.find_library("zstd")):
            if not name:
                continue
            try:
                lib = ctypes.CDLL(name)
                break
            except OSError:
                continue
        if lib is None:
            raise RuntimeError("libzstd.so.1 not found")
        lib.ZSTD_compressBound.restype = ctypes.c_size_t
        lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        lib.ZSTD_compress.restype = ctypes.c_size_t
        lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.ZSTD_decompress.restype = ctypes.c_size_t
        lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        lib.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        lib.ZSTD_isError.restype = ctypes.c_uint
        lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
        self.lib = lib

    def compress(self, data: bytes, level: int = 3) -> bytes:
        bound = self.lib.ZSTD_compressBound(len(data))
        out = ctypes.create_string_buffer(bound)
        n = self.lib.ZSTD_compress(out, bound, data, len(data), level)
        if self.lib.ZSTD_isError(n):
            raise ValueError("zstd compress failed")
        return out.raw[:n]

    def decompress(self, data: bytes) -> bytes:
        size = self.lib.ZSTD_getFrameContentSize(data, len(data))
        if size >= (1 << 62):
            raise ValueError("zstd frame without content size")
        out = ctypes.create_string_buffer(max(int(size), 1))
        n = self.lib.ZSTD_decompress(out, int(size), data, len(data))
        if self.lib.ZSTD_isError(n) or n != size:
            raise ValueError("zstd decompress failed")
        return out.raw[: int(size)]


_zstd: Optional[_Zstd] = None


def zstd() -> _Zstd:
    global _zstd
    if _zstd is None:
        _zstd = _Zstd()
    return _zstd


# --------------------------------------------------------------------------- readers
class Reader:
    def __init__(self, buf: bytes, pos: int = 0) -> None:
        self.buf = buf
        self.pos = pos

    def take(self, n: int) -> bytes:
        if n < 0 or self.pos + n > len(self.buf):
            raise ValueError("unexpected end of buffer")
        b = self.buf[self.pos : self.pos + n]
        self.pos += n
        return b

    def u8(self) -> int:
        return self.take(1)[0]

    def u16(self) -> int:
        return struct.unpack("<H", self.take(2))[0]

    def u32(self) -> int:
        return struct.unpack("<I", self.take(4))[0]

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]

    def f64(self) -> float:
        return struct.unpack("<d", self.take(8))[0]

    def done(self) -> bool:
        return self.pos == len(self.buf)


# --------------------------------------------------------------------------- Params
@dataclass
class Params:
    lattice_dimension: int = N
    coeff_modulus: Tuple[int, ...] = COEFF_MODULUS
    plain_modulus: int = T_PLAIN
    scheme: int = SCHEME_BFV_SERDE
    security: int = SECURITY_TC128_SERDE

    def to_bytes(self) -> bytes:
        out = struct.pack("<QQ", self.lattice_dimension, len(self.coeff_modulus))
        for q in self.coeff_modulus:
            out += struct.pack("<Q", q)
        out += struct.pack("<QII", self.plain_modulus, self.scheme, self.security)
        return out

    @staticmethod
    def read(r: Reader) -> "Params":
        n = r.u64()
        k = r.u64()
        if k > 64:
            raise ValueError("bad coeff modulus count")
        qs = tuple(r.u64() for _ in range(k))
        t = r.u64()
        scheme = r.u32()
        sec = r.u32()
        return Params(n, qs, t, scheme, sec)


def parms_id(moduli: Tuple[int, ...], n: int = N, t: int = T_PLAIN) -> Tuple[int, int, int, int]:
    """SEAL 4.0 parms_id = BLAKE2b-256 over LE u64 [scheme, N, q_0.., t]."""
    words = [SEAL_SCHEME_BFV, n, *moduli, t]
    h = hashlib.blake2b(struct.pack(f"<{len(words)}Q", *words), digest_size=32).digest()
    return struct.unpack("<4Q", h)


PARMS_ID_KEY = parms_id(COEFF_MODULUS)
PARMS_ID_DATA = parms_id(COEFF_MODULUS[:2])


# --------------------------------------------------------------------------- SEAL blobs
def seal_header(compr: int, total_size: int) -> bytes:
    return struct.pack("<HBBBBHQ", SEAL_MAGIC, SEAL_HEADER_SIZE, 4, 0, compr, 0, total_size)


def seal_unwrap(blob: bytes) -> Tuple[bytes, int]:
    """SEAL blob (16-byte header + body) -> (decompressed payload, compr_mode)."""
    if len(blob) < SEAL_HEADER_SIZE:
        raise ValueError("SEAL blob too short")
    magic, hsz, vmaj, vmin, compr, rsv, size = struct.unpack("<HBBBBHQ", blob[:16])
    if magic != SEAL_MAGIC or hsz != SEAL_HEADER_SIZE or vmaj != 4 or size != len(blob):
        raise ValueError("bad SEAL header")
    body = blob[16:]
    if compr == COMPR_NONE:
        return body, compr
    if compr == COMPR_ZSTD:
        return zstd().decompress(body), compr
    raise ValueError("unsupported compr_mode %d" % compr)


def seal_wrap(payload: bytes, compr: int = COMPR_ZSTD, level: int = 3) -> bytes:
    body = payload if compr == COMPR_NONE else zstd().compress(payload, level)
    return seal_header(compr, SEAL_HEADER_SIZE + len(body)) + body


def dynarray_payload(words: np.ndarray) -> bytes:
    """SEAL DynArray<u64>::save with compr none: header(16) + count + data."""
    data = np.ascontiguousarray(words, dtype="<u8").tobytes()
    return seal_header(COMPR_NONE, SEAL_HEADER_SIZE + 8 + len(data)) + struct.pack("<Q", len(data) // 8) + data


def read_dynarray(r: Reader) -> np.ndarray:
    hdr = r.take(16)
    magic, hsz, vmaj, vmin, compr, rsv, size = struct.unpack("<HBBBBHQ", hdr)
    if magic != SEAL_MAGIC or compr != COMPR_NONE:
        raise ValueError("bad inner DynArray header")
    count = r.u64()
    if size != 24 + 8 * count:
        raise ValueError("bad inner DynArray size")
    return np.frombuffer(r.take(8 * count), dtype="<u8").copy()


@dataclass
class SealCiphertext:
    """SEAL Ciphertext::save_members payload (also the payload of a PublicKey)."""

    parms_id: Tuple[int, int, int, int]
    is_ntt_form: int
    size: int
    poly_modulus_degree: int
    coeff_modulus_size: int
    scale: float
    correction_factor: int
    data: np.ndarray  # u64, [size][coeff_modulus_size][N]

    def polys(self) -> np.ndarray:
        return self.data.reshape(self.size, self.coeff_modulus_size, self.poly_modulus_degree)

    def payload(self) -> bytes:
        return (
            struct.pack("<4Q", *self.parms_id)
            + struct.pack("<B", self.is_ntt_form)
            + struct.pack("<QQQ", self.size, self.poly_modulus_degree, self.coeff_modulus_size)
            + struct.pack("<d", self.scale)
            + struct.pack("<Q", self.correction_factor)
            + dynarray_payload(self.data)
        )

    @staticmethod
    def read(r: Reader) -> "SealCiphertext":
        pid = struct.unpack("<4Q", r.take(32))
        ntt = r.u8()
        size, n, k = r.u64(), r.u64(), r.u64()
        scale = r.f64()
        corr = r.u64()
        data = read_dynarray(r)
        if len(data) != size * n * k:
            raise ValueError("ciphertext data size mismatch")
        return SealCiphertext(pid, ntt, size, n, k, scale, corr, data)

    @staticmethod
    def from_payload(payload: bytes) -> "SealCiphertext":
        r = Reader(payload)
        ct = SealCiphertext.read(r)
        if not r.done():
            raise ValueError("trailing bytes in ciphertext payload")
        return ct


def fresh_data_ciphertext(polys: np.ndarray) -> SealCiphertext:
    """Wrap [size][2][N] residues as a data-level, coefficient-form BFV ciphertext."""
    size = polys.shape[0]
    return SealCiphertext(PARMS_ID_DATA, 0, size, N, 2, 1.0, 1, np.ascontiguousarray(polys, dtype=np.uint64).reshape(-1))


# --------------------------------------------------------------------------- keys
@dataclass
class RelinKeys:
    parms_id: Tuple[int, int, int, int]
    keys: List[List[SealCiphertext]]  # [dim1][dim2]

    def payload(self) -> bytes:
        out = struct.pack("<4Q", *self.parms_id) + struct.pack("<Q", len(self.keys))
        for row in self.keys:
            out += struct.pack("<Q", len(row))
            for k in row:
                p = k.payload()
                out += seal_header(COMPR_NONE, SEAL_HEADER_SIZE + len(p)) + p
        return out

    @staticmethod
    def from_payload(payload: bytes) -> "RelinKeys":
        r = Reader(payload)
        pid = struct.unpack("<4Q", r.take(32))
        dim1 = r.u64()
        keys = []
        for _ in range(dim1):
            dim2 = r.u64()
            row = []
            for _ in range(dim2):
                hdr = r.take(16)
                magic, hsz, vmaj, vmin, compr, rsv, size = struct.unpack("<HBBBBHQ", hdr)
                if magic != SEAL_MAGIC or compr != COMPR_NONE:
                    raise ValueError("bad kswitch key header")
                row.append(SealCiphertext.from_payload(r.take(size - 16)))
            keys.append(row)
        if not r.done():
            raise ValueError("trailing bytes in relin keys payload")
        return RelinKeys(pid, keys)


@dataclass
class SecretKey:
    parms_id: Tuple[int, int, int, int]
    data: np.ndarray  # [3][N] NTT form

    @staticmethod
    def from_payload(payload: bytes) -> "SecretKey":
        # SEAL SecretKey::save = Plaintext::save_members: parms_id, coeff_count, scale, DynArray
        r = Reader(payload)
        pid = struct.unpack("<4Q", r.take(32))
        coeff_count = r.u64()
        scale = r.f64()
        data = read_dynarray(r)
        if len(data) != coeff_count or not r.done():
            raise ValueError("bad secret key payload")
        return SecretKey(pid, data.reshape(-1, N))


@dataclass
class WithContext:
    params: Params
    blob: bytes  # raw SEAL blob (header + possibly-compressed body)

    def to_bytes(self) -> bytes:
        return self.params.to_bytes() + struct.pack("<Q", len(self.blob)) + self.blob

    @staticmethod
    def read(r: Reader) -> "WithContext":
        p = Params.read(r)
        n = r.u64()
        return WithContext(p, r.take(n))


@dataclass
class PublicKey:
    public_key: WithContext
    galois_key: Optional[WithContext]
    relin_key: Optional[WithContext]

    def to_bytes(self) -> bytes:
        out = self.public_key.to_bytes()
        for opt in (self.galois_key, self.relin_key):
            out += b"\x00" if opt is None else b"\x01" + opt.to_bytes()
        return out

    @staticmethod
    def from_bytes(buf: bytes) -> "PublicKey":
        r = Reader(buf)
        pk = WithContext.read(r)
        opts = []
        for _ in range(2):
            tag = r.u8()
            if tag == 0:
                opts.append(None)
            elif tag == 1:
                opts.append(WithContext.read(r))
            else:
                raise ValueError("bad Option tag")
        if not r.done():
            raise ValueError("trailing bytes after PublicKey")
        return PublicKey(pk, opts[0], opts[1])

    def pk_polys(self) -> np.ndarray:
        payload, _ = seal_unwrap(self.public_key.blob)
        return SealCiphertext.from_payload(payload).polys()

    def relin(self) -> RelinKeys:
        if self.relin_key is None:
            raise ValueError("no relin keys")
        payload, _ = seal_unwrap(self.relin_key.blob)
        return RelinKeys.from_payload(payload)


def read_private_key(buf: bytes) -> SecretKey:
    r = Reader(buf)
    wc = WithContext.read(r)
    if not r.done():
        raise ValueError("trailing bytes after PrivateKey")
    payload, _ = seal_unwrap(wc.blob)
    return SecretKey.from_payload(payload)


# --------------------------------------------------------------------------- Ciphertext
@dataclass
class Ciphertext:
    """sunscreen::Ciphertext { data_type, inner: Seal(Vec<WithContext<SealCiphertext>>) }."""

    data_type: str
    parts: List[Tuple[Params, SealCiphertext]] = field(default_factory=list)

    def to_bytes(self, compr: int = COMPR_ZSTD) -> bytes:
        dt = self.data_type.encode()
        out = struct.pack("<Q", len(dt)) + dt + struct.pack("<I", 0) + struct.pack("<Q", len(self.parts))
        for params, ct in self.parts:
            out += WithContext(params, seal_wrap(ct.payload(), compr)).to_bytes()
        return out

    @staticmethod
    def from_bytes(buf: bytes) -> "Ciphertext":
        r = Reader(buf)
        n = r.u64()
        dt = r.take(n).decode()
        if r.u32() != 0:
            raise ValueError("unknown InnerCiphertext variant")
        cnt = r.u64()
        if cnt > 16:
            raise ValueError("too many inner ciphertexts")
        parts = []
        for _ in range(cnt):
            wc = WithContext.read(r)
            payload, _ = seal_unwrap(wc.blob)
            parts.append((wc.params, SealCiphertext.from_payload(payload)))
        if not r.done():
            raise ValueError("trailing bytes after Ciphertext")
        return Ciphertext(dt, parts)

    def polys(self) -> np.ndarray:
        return self.parts[0][1].polys()


def make_ciphertext(kind: str, polys: np.ndarray) -> Ciphertext:
    return Ciphertext(data_type_string(kind), [(Params(), fresh_data_ciphertext(polys))])


# --------------------------------------------------------------------------- pack.rs framing
def pack_one_argument(a: bytes) -> bytes:
    return a


def pack_two_arguments(a: bytes, b: bytes) -> bytes:
    return struct.pack(">I", 4 + len(a)) + a + b


def unpack_two_arguments(buf: bytes) -> Tuple[bytes, bytes]:
    if len(buf) < 4:
        raise EOFError
    ix1 = struct.unpack(">I", buf[:4])[0]
    if ix1 < 4 or ix1 > len(buf):
        raise EOFError
    return buf[4:ix1], buf[ix1:]


def pack_binary_operation(pk: bytes, a: bytes, b: bytes) -> bytes:
    ix1 = 8 + len(pk)
    ix2 = ix1 + len(a)
    return struct.pack(">II", ix1, ix2) + pk + a + b


def unpack_binary_operation(buf: bytes) -> Tuple[bytes, bytes, bytes]:
    if len(buf) < 8:
        raise EOFError
    ix1, ix2 = struct.unpack(">II", buf[:8])
    if ix1 < 8 or ix2 < ix1 or ix2 > len(buf):
        raise EOFError
    return buf[8:ix1], buf[ix1:ix2], buf[ix2:]


# scalar operands, pack.rs:47-104
def ser_u64(v: int) -> bytes:
    return struct.pack(">Q", v & (2**64 - 1))


def ser_i64(v: int) -> bytes:
    return struct.pack(">q", v)


def ser_u256(v: int) -> bytes:
    return (v & (2**256 - 1)).to_bytes(32, "big")


def ser_f64(v: float) -> bytes:
    return struct.pack(">d", v)


SCALAR_SER = {"u64": ser_u64, "i64": ser_i64, "u256": ser_u256, "frac64": ser_f64}


def deser_scalar(kind: str, b: bytes):
    if kind == "u64":
        return struct.unpack(">Q", b)[0]
    if kind == "i64":
        return struct.unpack(">q", b)[0]
    if kind == "u256":
        if len(b) != 32:
            raise ValueError
        return int.from_bytes(b, "big")
    if kind == "frac64":
        return struct.unpack(">d", b)[0]
    raise KeyError(kind)
