"""TEST INFRASTRUCTURE ONLY -- independent pure-Python restatement of the deterministic public-key encryption behind
`FheApp::encrypt` / `reencrypt` (/root/reference/src/fhe.rs:594-657): sunscreen 0.8.1 `encrypt_deterministic` ->
Sunscreen's SEAL 4.0 fork.  PINNED: `encrypt_deterministic()` below + oracle/formats.py reproduce all three SHA-512 known
answers the reference holds (fhe.rs:2101-2121, 2165-2185, 2224-2244), the Linux ones with the libstdc++ distributions and the
macOS ones with the libc++ distributions (tests/test_oracle_kat.py).  The C oracle (bfv_oracle.c: bfvo_seal_encrypt) is a
second restatement of the libstdc++ variant and is compared with this one.

sunscreen / seal_fhe / SEAL are not vendored, so every piece is a restatement of a published algorithm:
  * BLAKE2Xb (BLAKE2X paper, reference blake2xb.c) as SEAL's Blake2xbPRNG uses it: 4096-byte buffers,
    buffer c = blake2xb(out 4096, in = c as LE u64, key = 64-byte seed)
  * SEAL util::sample_poly_ternary: std::uniform_int_distribution<uint64_t>(0, 2) over a 32-bit engine
    (libstdc++ >= 11: Lemire's method; libc++: independent-bits engine, 2 bits with rejection)
  * SEAL util::sample_poly_normal (SEAL_USE_GAUSSIAN_NOISE): ClippedNormalDistribution(0, 3.2, 19.2) over
    std::normal_distribution<double> (both libraries: Marsaglia polar; they differ in which variate is returned first)
  * SEAL util::encrypt_zero_asymmetric at the DATA level (the fork's component-exporting encrypt disables the special
    modulus: no divide_and_round_q_last) + multiply_add_plain_with_scaling_variant
`scripts/kat_search2.py` is the search that found the combination (the stock-SEAL shape with modulus switching, kept below as
`encrypt_seeded`, does not match).
"""
from __future__ import annotations

import math
import struct
from typing import Callable, List

import numpy as np

from . import bfv
from . import formats as F

MASK64 = (1 << 64) - 1
IV = [0x6A09E667F3BCC908, 0xBB67AE8584CAA73B, 0x3C6EF372FE94F82B, 0xA54FF53A5F1D36F1,
      0x510E527FADE682D1, 0x9B05688C2B3E6C1F, 0x1F83D9ABFB41BD6B, 0x5BE0CD19137E2179]
SIGMA = [
    [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
    [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4], [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
    [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13], [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
    [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11], [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
    [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5], [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0],
    [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
]


def _rotr(x: int, n: int) -> int:
    return ((x >> n) | (x << (64 - n))) & MASK64


def _compress(h: List[int], block: bytes, t: int, last: bool) -> None:
    m = struct.unpack("<16Q", block)
    v = h[:] + IV[:]
    v[12] ^= t & MASK64
    v[13] ^= (t >> 64) & MASK64
    if last:
        v[14] ^= MASK64
    for r in range(12):
        s = SIGMA[r]
        for i, (a, b, c, d) in enumerate(((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15),
                                          (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14))):
            x, y = m[s[2 * i]], m[s[2 * i + 1]]
            v[a] = (v[a] + v[b] + x) & MASK64
            v[d] = _rotr(v[d] ^ v[a], 32)
            v[c] = (v[c] + v[d]) & MASK64
            v[b] = _rotr(v[b] ^ v[c], 24)
            v[a] = (v[a] + v[b] + y) & MASK64
            v[d] = _rotr(v[d] ^ v[a], 16)
            v[c] = (v[c] + v[d]) & MASK64
            v[b] = _rotr(v[b] ^ v[c], 63)
    for i in range(8):
        h[i] ^= v[i] ^ v[i + 8]


def blake2b_param(data: bytes, param: bytes, key: bytes = b"", outlen: int = 64) -> bytes:
    """BLAKE2b with an explicit 64-byte parameter block (RFC 7693 section 2.5 layout)."""
    assert len(param) == 64
    h = [iv ^ p for iv, p in zip(IV, struct.unpack("<8Q", param))]
    buf = (key.ljust(128, b"\0") if key else b"") + data
    t = 0
    while len(buf) > 128:
        t += 128
        _compress(h, buf[:128], t, False)
        buf = buf[128:]
    t += len(buf)
    _compress(h, buf.ljust(128, b"\0"), t, True)
    return struct.pack("<8Q", *h)[:outlen]


def _param(digest_length, key_length, fanout, depth, leaf_length, node_offset, xof_length, node_depth, inner_length) -> bytes:
    return struct.pack("<BBBBIIIBB", digest_length, key_length, fanout, depth, leaf_length, node_offset, xof_length,
                       node_depth, inner_length) + b"\0" * 14 + b"\0" * 16 + b"\0" * 16


def blake2xb(outlen: int, data: bytes, key: bytes) -> bytes:
    """BLAKE2Xb as in the BLAKE2 reference blake2xb.c (which SEAL bundles)."""
    root = blake2b_param(data, _param(64, len(key), 1, 1, 0, 0, outlen, 0, 0), key, 64)
    out = b""
    i = 0
    while len(out) < outlen:
        bs = min(64, outlen - len(out))
        out += blake2b_param(root, _param(bs, 0, 0, 0, 64, i, outlen, 0, 64), b"", bs)
        i += 1
    return out


_IVN = np.array(IV, dtype=np.uint64)


def blake2xb_fast(outlen: int, data: bytes, key: bytes) -> bytes:
    """blake2xb() with hashlib for the root hash and one numpy pass for all expansion nodes (hashlib refuses depth = 0)."""
    import hashlib

    root = hashlib.blake2b(data, digest_size=64, key=key, fanout=1, depth=1, node_offset=outlen << 32).digest()
    assert outlen % 64 == 0
    nb = outlen // 64
    p = np.stack([np.frombuffer(_param(64, 0, 0, 0, 64, i, outlen, 0, 64), dtype="<u8") for i in range(nb)])
    h = (_IVN[None, :] ^ p).T.copy()  # [8][nb]
    m = np.frombuffer(root.ljust(128, b"\0"), dtype="<u8")
    v = [h[i].copy() for i in range(8)] + [np.full(nb, _IVN[i], dtype=np.uint64) for i in range(8)]
    v[12] = v[12] ^ np.uint64(64)
    v[14] = v[14] ^ np.uint64(MASK64)
    rotr = lambda x, n: (x >> np.uint64(n)) | (x << np.uint64(64 - n))
    with np.errstate(over="ignore"):
        for r in range(12):
            s = SIGMA[r]
            for i, (a, b, c, d) in enumerate(((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15),
                                              (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14))):
                x, y = m[s[2 * i]], m[s[2 * i + 1]]
                v[a] = v[a] + v[b] + x
                v[d] = rotr(v[d] ^ v[a], 32)
                v[c] = v[c] + v[d]
                v[b] = rotr(v[b] ^ v[c], 24)
                v[a] = v[a] + v[b] + y
                v[d] = rotr(v[d] ^ v[a], 16)
                v[c] = v[c] + v[d]
                v[b] = rotr(v[b] ^ v[c], 63)
    return np.stack([h[i] ^ v[i] ^ v[i + 8] for i in range(8)], axis=1).astype("<u8").tobytes()


class Blake2xbPRNG:
    """SEAL Blake2xbPRNG: 4096-byte buffer refilled with blake2xb(counter as LE u64, key = 64-byte seed)."""

    BUF = 4096

    def __init__(self, seed_words, fast: bool = True) -> None:
        """seed_words: 8 u64 words (the SHA-512 digest read as little-endian words) or the 64 digest bytes."""
        self.seed = bytes(seed_words) if isinstance(seed_words, (bytes, bytearray)) else struct.pack("<8Q", *seed_words)
        self.fast = fast
        self.counter = 0
        self.buf = b""
        self.pos = 0
        self._refill()

    def _refill(self) -> None:
        data = struct.pack("<Q", self.counter)
        self.buf = blake2xb_fast(self.BUF, data, self.seed) if self.fast else blake2xb(self.BUF, data, self.seed)
        self.counter += 1
        self.pos = 0

    def generate(self, n: int) -> bytes:
        out = b""
        while n:
            take = min(n, self.BUF - self.pos)
            out += self.buf[self.pos : self.pos + take]
            self.pos += take
            n -= take
            if self.pos == self.BUF:
                self._refill()
        return out

    def u32(self) -> int:
        return struct.unpack("<I", self.generate(4))[0]


# ---- std::uniform_int_distribution<uint64_t>(0, 2) over a 32-bit URNG, libstdc++
def uniform3_lemire(prng: Blake2xbPRNG) -> int:
    """libstdc++ >= 11: _S_nd (Lemire's nearly divisionless) because the engine range is exactly 2^32."""
    rng = 3
    product = prng.u32() * rng
    low = product & 0xFFFFFFFF
    if low < rng:
        threshold = ((1 << 32) - rng) % rng
        while low < threshold:
            product = prng.u32() * rng
            low = product & 0xFFFFFFFF
    return product >> 32


def uniform3_downscale(prng: Blake2xbPRNG) -> int:
    """libstdc++ <= 10: classic downscaling with rejection."""
    uerange = 3
    scaling = 0xFFFFFFFF // uerange
    past = uerange * scaling
    while True:
        r = prng.u32()
        if r < past:
            return r // scaling


def sample_ternary(prng: Blake2xbPRNG, uniform3: Callable[[Blake2xbPRNG], int], n: int = F.N) -> np.ndarray:
    return np.array([uniform3(prng) - 1 for _ in range(n)], dtype=np.int64)


def sample_cbd(prng: Blake2xbPRNG, n: int = F.N) -> np.ndarray:
    raw = np.frombuffer(prng.generate(6 * n), dtype=np.uint8).reshape(n, 6).copy()
    raw[:, 2] &= 0x1F
    raw[:, 5] &= 0x1F
    pop = np.unpackbits(raw, axis=1).reshape(n, 6, 8).sum(axis=2).astype(np.int64)
    return pop[:, 0] + pop[:, 1] + pop[:, 2] - pop[:, 3] - pop[:, 4] - pop[:, 5]


def _canonical53(prng: Blake2xbPRNG) -> float:
    """libstdc++ std::generate_canonical<double, 53> over a 32-bit engine: two draws, low word first."""
    s = float(prng.u32())
    s += float(prng.u32()) * 4294967296.0
    r = s / 18446744073709551616.0
    return r if r < 1.0 else math.nextafter(1.0, 0.0)


def sample_clipped_normal(prng: Blake2xbPRNG, n: int = F.N, sigma: float = 3.2, max_dev: float = 19.2) -> np.ndarray:
    """SEAL util::sample_poly_normal (SEAL_USE_GAUSSIAN_NOISE): ClippedNormalDistribution over libstdc++'s
    std::normal_distribution<double> (Marsaglia polar, second variate cached), truncated toward zero to int64.
    The key fixtures' error statistics (P(0) = 0.245, variance 7.9) identify this sampler."""
    out = np.empty(n, dtype=np.int64)
    saved = None
    for i in range(n):
        while True:
            if saved is not None:
                ret, saved = saved, None
            else:
                while True:
                    x = 2.0 * _canonical53(prng) - 1.0
                    y = 2.0 * _canonical53(prng) - 1.0
                    r2 = x * x + y * y
                    if not (r2 > 1.0 or r2 == 0.0):
                        break
                mult = math.sqrt(-2.0 * math.log(r2) / r2)
                saved = x * mult
                ret = y * mult
            value = ret * sigma + 0.0
            if abs(value) <= max_dev:
                break
        out[i] = int(value)  # static_cast<int64_t>: truncation toward zero
    return out


def encrypt_seeded(pk: np.ndarray, plain: np.ndarray, seed_words: List[int], uniform3=uniform3_lemire,
                   noise=sample_cbd) -> np.ndarray:
    """Stock SEAL Encryptor::encrypt (BFV, asymmetric, WITH modulus switching) on a seeded Blake2xb PRNG -- the shape of
    the reference's randomised `runtime.encrypt`; NOT what `encrypt_deterministic` does (see encrypt_deterministic below).
    pk: [2][3][N] NTT form; returns the data-level ciphertext [2][2][N]."""
    mods = bfv.moduli()[:3]
    prng = Blake2xbPRNG(seed_words)
    u = sample_ternary(prng, uniform3)
    c = np.zeros((2, 3, F.N), dtype=np.uint64)
    for J in range(3):
        q = mods[J]
        uj = np.where(u < 0, q - 1, u).astype(np.uint64)
        un = bfv.ntt_fwd(uj, J)
        for j in range(2):
            prod = (un.astype(object) * pk[j, J].astype(object)) % q
            c[j, J] = bfv.ntt_inv(np.array(prod, dtype=np.uint64), J)
    for j in range(2):
        e = noise(prng)
        for J in range(3):
            q = mods[J]
            c[j, J] = ((c[j, J].astype(object) + np.where(e < 0, q + e, e).astype(object)) % q).astype(np.uint64)
    # RNSTool::divide_and_round_q_last_inplace per polynomial
    P = mods[2]
    half = P >> 1
    out = np.zeros((2, 2, F.N), dtype=np.uint64)
    for j in range(2):
        last = (c[j, 2].astype(object) + half) % P
        for l in range(2):
            q = mods[l]
            t = (last % q - half % q) % q
            out[j, l] = (((c[j, l].astype(object) - t) % q) * pow(P, -1, q) % q).astype(np.uint64)
    return bfv.add_plain(out, plain)


# ---- libc++ (macOS) variants of the two distributions
def uniform3_libcxx(prng: Blake2xbPRNG) -> int:
    """libc++ uniform_int_distribution<uint64_t>(0, 2): __independent_bits_engine with w = 2 bits, reject 3."""
    while True:
        v = prng.u32() & 3
        if v < 3:
            return v


def sample_clipped_normal_libcxx(prng: Blake2xbPRNG, n: int = F.N, sigma: float = 3.2, max_dev: float = 19.2) -> np.ndarray:
    """libc++ std::normal_distribution (polar on uniform_real(-1, 1); returns u*F first and keeps v*F) under SEAL's
    ClippedNormalDistribution; generate_canonical is the same two-draw sum without the >= 1 clamp."""
    out = np.empty(n, dtype=np.int64)
    saved = None
    canon = lambda: (float(prng.u32()) + float(prng.u32()) * 4294967296.0) / 18446744073709551616.0
    for i in range(n):
        while True:
            if saved is not None:
                up, saved = saved, None
            else:
                while True:
                    u = 2.0 * canon() + -1.0
                    v = 2.0 * canon() + -1.0
                    s = u * u + v * v
                    if not (s > 1.0 or s == 0.0):
                        break
                fp = math.sqrt(-2.0 * math.log(s) / s)
                saved = v * fp
                up = u * fp
            value = up * sigma + 0.0
            if abs(value) <= max_dev:
                break
        out[i] = int(value)
    return out


STDLIBS = {"libstdc++": (uniform3_lemire, sample_clipped_normal), "libc++": (uniform3_libcxx, sample_clipped_normal_libcxx)}


def sample_deterministic(seed, stdlib: str = "libstdc++"):
    """(u, e0, e1, 32-bit draws consumed) in SEAL's order: ternary u, then the two error polynomials, one PRNG."""
    u3, noise = STDLIBS[stdlib]
    prng = Blake2xbPRNG(seed)
    u = sample_ternary(prng, u3)
    e0 = noise(prng)
    e1 = noise(prng)
    return u, e0, e1, (prng.counter - 1) * (prng.BUF // 4) + prng.pos // 4


def encrypt_deterministic(pk: np.ndarray, plain: np.ndarray, seed, stdlib: str = "libstdc++") -> np.ndarray:
    """sunscreen 0.8.1 `encrypt_deterministic` (PINNED by the reference's known answers): c_j = pk_j[:2] * u + e_j at the data
    level, no modulus switching; c_0 += scaled plaintext.  pk: [2][3][N] NTT form; returns [2][2][N]."""
    mods = bfv.moduli()[:2]
    u, e0, e1, _ = sample_deterministic(seed, stdlib)
    c = np.zeros((2, 2, F.N), dtype=np.uint64)
    for J, q in enumerate(mods):
        un = bfv.ntt_fwd(np.where(u < 0, q - 1, u).astype(np.uint64), J)
        for j, e in enumerate((e0, e1)):
            prod = (un.astype(object) * pk[j, J].astype(object)) % q
            cj = bfv.ntt_inv(np.array(prod, dtype=np.uint64), J)
            c[j, J] = ((cj.astype(object) + np.where(e < 0, q + e, e).astype(object)) % q).astype(np.uint64)
    return bfv.add_plain(c, plain)
