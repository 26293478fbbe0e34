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
fixtures (src/data/network.{pub,pri}, tests/data/{public,private}_key.bin).  The whole
`Ciphertext` serialisation -- data_type string, bincode layout, SEAL ciphertext payload and
its libzstd level-3 compression -- is pinned by the reference's SHA-512 known answers of
`encrypt` / `reencrypt` (fhe.rs:2101-2244, tests/test_oracle_kat.py).
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

# Type-name strings sunscreen puts in `Ciphertext.data_type`: `#[derive(TypeName)]` = module_path!() + "::" + the struct's
# identifier WITHOUT generic arguments, so Unsigned<1> (Unsigned64) and Unsigned<4> (Unsigned256) share one name.  The
# Unsigned spelling and the version are pinned by the reference's SHA-512 known answers (tests/test_oracle_kat.py); Signed
# and Fractional follow the same derive.
TYPE_NAMES = {
    "i64": "sunscreen::types::bfv::signed::Signed",
    "u64": "sunscreen::types::bfv::unsigned::Unsigned",
    "u256": "sunscreen::types::bfv::unsigned::Unsigned",
    "frac64": "sunscreen::types::bfv::fractional::Fractional",
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

    def decompress_stream(self, data: bytes, in_chunk: int = 1000, out_chunk: int = 4096) -> bytes:
        """ZSTD_decompressStream fed in small pieces -- the API SEAL's Serialization::Load uses."""

        class Buf(ctypes.Structure):
            _fields_ = [("p", ctypes.c_void_p), ("size", ctypes.c_size_t), ("pos", ctypes.c_size_t)]

        L = self.lib
        L.ZSTD_createDStream.restype = ctypes.c_void_p
        L.ZSTD_freeDStream.argtypes = [ctypes.c_void_p]
        L.ZSTD_decompressStream.restype = ctypes.c_size_t
        L.ZSTD_decompressStream.argtypes = [ctypes.c_void_p, ctypes.POINTER(Buf), ctypes.POINTER(Buf)]
        ds = L.ZSTD_createDStream()
        out = bytearray()
        obuf = ctypes.create_string_buffer(out_chunk)
        try:
            pos, ret = 0, 1
            while pos < len(data):
                piece = ctypes.create_string_buffer(data[pos : pos + in_chunk], min(in_chunk, len(data) - pos))
                ib = Buf(ctypes.cast(piece, ctypes.c_void_p), len(piece), 0)
                while ib.pos < ib.size:
                    ob = Buf(ctypes.cast(obuf, ctypes.c_void_p), out_chunk, 0)
                    ret = L.ZSTD_decompressStream(ds, ctypes.byref(ob), ctypes.byref(ib))
                    if L.ZSTD_isError(ret):
                        raise ValueError("zstd stream decompress failed")
                    out += obuf.raw[: ob.pos]
                pos += len(piece)
            while ret != 0:  # flush what the decoder still holds
                ib = Buf(None, 0, 0)
                ob = Buf(ctypes.cast(obuf, ctypes.c_void_p), out_chunk, 0)
                ret = L.ZSTD_decompressStream(ds, ctypes.byref(ob), ctypes.byref(ib))
                if L.ZSTD_isError(ret):
                    raise ValueError("zstd stream decompress failed")
                out += obuf.raw[: ob.pos]
                if ob.pos == 0:
                    break
            if ret != 0:
                raise ValueError("zstd stream ended inside a frame")
        finally:
            L.ZSTD_freeDStream(ds)
        return bytes(out)

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
    if compr == COMPR_ZLIB:
        import zlib

        return zlib.decompress(body), compr
    raise ValueError("unsupported compr_mode %d" % compr)


def seal_wrap(payload: bytes, compr: int = COMPR_ZSTD, level: int = 3, structured: bool = False) -> bytes:
    if compr == COMPR_NONE:
        body = payload
    elif compr == COMPR_ZLIB:
        import zlib

        body = zlib.compress(payload)  # SEAL's zlib mode: zlib container, default level
    else:
        body = zstd_structured_frame(payload) if structured else None
        if body is None:
            body = zstd().compress(payload, level)
    return seal_header(compr, SEAL_HEADER_SIZE + len(body)) + body


# --------------------------------------------------------------------------- structured zstd frames
# The product writes zstd-mode ciphertext payloads as hand-laid-out RFC 8878 frames (csrc/codec.cpp, zstd_pack40).
# This is the independent restatement of that layout, straight from the RFC, used to check the bytes; the mini decoder
# below reads such frames WITHOUT libzstd (raw literals + RLE-mode sequences only), so the layout is checked twice:
# by the reference decoder (libzstd) and by the RFC's rules.
CT_PREFIX = 97  # SEAL ciphertext payload bytes before the coefficient words
_LL_BASE = list(range(16)) + [16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]
_LL_BITS = [0] * 16 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
_ML_BASE = list(range(3, 35)) + [35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539]
_ML_BITS = [0] * 32 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]


def _raw_literals_header(n: int) -> bytes:
    if n < 32:
        return bytes([n << 3])
    if n < 4096:
        return struct.pack("<H", (1 << 2) | (n << 4))
    return struct.pack("<I", (3 << 2) | (n << 4))[:3]


def _sequence_count(n: int) -> bytes:
    if n < 128:
        return bytes([n])
    if n < 0x7F00:
        return bytes([(n >> 8) + 0x80, n & 0xFF])
    return b"\xff" + struct.pack("<H", n - 0x7F00)


def _block(content: bytes, last: bool) -> bytes:
    return struct.pack("<I", int(last) | (2 << 1) | (len(content) << 3))[:3] + content


def zstd_structured_frame(payload: bytes, block_words: int = 16384) -> Optional[bytes]:
    """prefix(97) + 64-bit words below 2^40 -> zstd frame: block A (raw literals: prefix, word 0, 5 bytes of word 1; one
    sequence LL=110 ML=3 offset=8), then blocks of 5-byte literals with one (LL=5, ML=3, repeat-offset-1) sequence per word.
    None when the payload does not have that shape (the caller then uses libzstd)."""
    n = len(payload)
    if n < CT_PREFIX + 16 or (n - CT_PREFIX) % 8:
        return None
    words = np.frombuffer(payload, dtype="<u8", offset=CT_PREFIX)
    if (words >> np.uint64(40)).any() or (len(words) >= 16 and (words[:16] == words[0]).all()):
        return None
    w = len(words)
    out = b"\x28\xb5\x2f\xfd\xa0" + struct.pack("<I", n)
    lits_a = CT_PREFIX + 13
    bits = (lits_a - 64) | (3 << 6) | (1 << 9)  # from the top: end marker, offset extra (11 - 8), literal-length extra
    out += _block(_raw_literals_header(lits_a) + payload[:lits_a] + bytes([1, 0x54, 25, 3, 0]) + struct.pack("<H", bits), w == 2)
    low5 = np.frombuffer(payload, dtype=np.uint8, offset=CT_PREFIX).reshape(w, 8)[:, :5]
    done = 2
    while done < w:
        m = min(block_words, w - done)
        lits = low5[done : done + m].tobytes()
        done += m
        out += _block(_raw_literals_header(5 * m) + lits + _sequence_count(m) + bytes([0x54, 5, 0, 0, 1]), done == w)
    return out


def zstd_mini_decode(frame: bytes) -> bytes:
    """RFC 8878 decoder for the subset the structured writer uses: single-segment frames, compressed blocks with raw
    literals and all three sequence tables in RLE mode. Raises ValueError on anything else."""
    if frame[:4] != b"\x28\xb5\x2f\xfd":
        raise ValueError("not a zstd frame")
    fhd = frame[4]
    if fhd & 0x20 == 0 or fhd & 0x0F:
        raise ValueError("only single-segment frames without checksum / dictionary")
    fcs_len = {0: 1, 1: 2, 2: 4, 3: 8}[fhd >> 6]
    fcs = int.from_bytes(frame[5 : 5 + fcs_len], "little") + (256 if fcs_len == 2 else 0)
    pos = 5 + fcs_len
    out = bytearray()
    rep = [1, 4, 8]
    last = False
    while not last:
        bh = int.from_bytes(frame[pos : pos + 3], "little")
        last, btype, bsize = bool(bh & 1), (bh >> 1) & 3, bh >> 3
        pos += 3
        blk = frame[pos : pos + bsize]
        if len(blk) != bsize or btype != 2:
            raise ValueError("only complete compressed blocks")
        pos += bsize
        # literals section
        ltype, sf = blk[0] & 3, (blk[0] >> 2) & 3
        if ltype != 0:
            raise ValueError("only raw literals")
        if sf in (0, 2):
            nlit, p = blk[0] >> 3, 1
        elif sf == 1:
            nlit, p = int.from_bytes(blk[:2], "little") >> 4, 2
        else:
            nlit, p = int.from_bytes(blk[:3], "little") >> 4, 3
        lits = blk[p : p + nlit]
        p += nlit
        # sequences section
        b0 = blk[p]
        if b0 < 128:
            nseq, p = b0, p + 1
        elif b0 < 255:
            nseq, p = ((b0 - 128) << 8) + blk[p + 1], p + 2
        else:
            nseq, p = blk[p + 1] + (blk[p + 2] << 8) + 0x7F00, p + 3
        lp = 0
        if nseq:
            modes = blk[p]
            if modes != 0x54:
                raise ValueError("only RLE-mode sequence tables")
            ll_code, of_code, ml_code = blk[p + 1], blk[p + 2], blk[p + 3]
            stream = blk[p + 4 :]
            if not stream or stream[-1] == 0:
                raise ValueError("bad bitstream end")
            bits = int.from_bytes(stream, "little")
            top = bits.bit_length() - 1  # position of the end marker; fields are read downward from it

            def read(nb: int) -> int:
                nonlocal top
                if nb > top:
                    raise ValueError("bitstream underflow")
                top -= nb
                return (bits >> top) & ((1 << nb) - 1)

            for _ in range(nseq):
                of_value = (1 << of_code) + read(of_code)
                ml = _ML_BASE[ml_code] + read(_ML_BITS[ml_code])
                ll = _LL_BASE[ll_code] + read(_LL_BITS[ll_code])
                if of_value > 3:
                    offset = of_value - 3
                    rep = [offset, rep[0], rep[1]]
                else:
                    idx = of_value - 1 + (1 if ll == 0 else 0)
                    if idx == 0:
                        offset = rep[0]
                    elif idx == 3:
                        offset = rep[0] - 1
                        rep = [offset, rep[0], rep[1]]
                    else:
                        offset = rep[idx]
                        rep = [offset] + [r for i, r in enumerate(rep) if i != idx][:2]
                out += lits[lp : lp + ll]
                lp += ll
                if offset > len(out) or offset == 0:
                    raise ValueError("offset beyond window")
                for _ in range(ml):
                    out.append(out[-offset])
            if top != 0:
                raise ValueError("bitstream not fully consumed")
        out += lits[lp:]
    if pos != len(frame) or len(out) != fcs:
        raise ValueError("frame size mismatch")
    return bytes(out)


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

    def to_bytes(self, compr: int = COMPR_ZSTD, structured: bool = False) -> bytes:
        """structured=True: zstd payloads laid out as the product's default writer does (zstd_structured_frame)."""
        dt = self.data_type.encode()
        out = struct.pack("<Q", len(dt)) + dt + struct.pack("<I", 0) + struct.pack("<Q", len(self.parts))
        for params, ct in self.parts:
            out += WithContext(params, seal_wrap(ct.payload(), compr, structured=structured)).to_bytes()
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
