"""Second, time-boxed search for the reference's SHA-512 known answers of `FheApp::encrypt`
(/root/reference/src/fhe.rs:2083-2121: one hash for Linux, one for macOS).  Pure CPU, oracle only.

New against scripts/kat_search.py:
  * encryption WITHOUT the special modulus (the Sunscreen SEAL fork's component-exporting encrypt has a
    `disable_special_modulus` switch: c = pk[:, :2] u + e at the data level, no divide_and_round_q_last);
  * libc++ variants of std::uniform_int_distribution / std::normal_distribution against the macOS hash
    (the two hashes differ only through the C++ standard library's distributions);
  * BLAKE2Xb through hashlib (node_offset carries xof_length in its upper 32 bits): ~100x faster PRNG;
  * more data_type spellings, seed word orders, and draw orders (noise before u).
A hit prints MATCH and the combination."""
import hashlib
import itertools
import math
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bfv, formats as F, seal_encrypt as S  # noqa: E402

KAT = {
    "linux": bytes([190, 214, 153, 167, 205, 130, 61, 102, 188, 80, 220, 159, 38, 110, 126, 216, 148, 46, 220, 80, 18, 189, 177, 187,
                    108, 99, 32, 72, 250, 225, 2, 166, 33, 155, 22, 86, 221, 82, 4, 174, 144, 196, 45, 28, 190, 100, 194, 192, 37, 81,
                    203, 227, 46, 179, 59, 153, 20, 118, 191, 69, 244, 113, 180, 123]),
    "macos": bytes([195, 187, 246, 29, 229, 222, 20, 246, 218, 16, 114, 27, 129, 99, 163, 244, 92, 32, 26, 147, 244, 249, 195, 53, 242,
                    255, 161, 187, 61, 209, 68, 3, 64, 1, 253, 115, 134, 15, 254, 196, 206, 149, 60, 174, 228, 18, 210, 5, 80, 214, 31,
                    131, 22, 81, 220, 190, 246, 192, 62, 177, 213, 218, 109, 67]),
}
SECRET = bytes([15, 17, 225, 5, 30, 1, 237, 218, 130, 19, 37, 95, 222, 218, 244, 172, 214, 175, 175, 110, 173, 103, 172, 60, 43, 76, 40, 150,
                215, 96, 23, 78, 22, 39, 30, 177, 107, 130, 124, 109, 27, 96, 206, 125, 104, 241, 10, 40, 88, 238, 117, 118, 79, 113, 213, 110,
                148, 179, 53, 19, 227, 154, 151, 122])


_IV = np.array(S.IV, dtype=np.uint64)


def _expand(root: bytes, outlen: int) -> bytes:
    """The BLAKE2Xb expansion nodes (depth 0, which hashlib refuses), all output blocks at once in numpy."""
    nb = outlen // 64
    p = np.zeros((nb, 8), dtype=np.uint64)
    for i in range(nb):
        p[i] = np.frombuffer(S._param(64, 0, 0, 0, 64, i, outlen, 0, 64), dtype="<u8")
    h = (_IV[None, :] ^ p).T.copy()  # [8][nb]
    m = np.frombuffer(root.ljust(128, b"\0"), dtype="<u8")
    v = [h[i].copy() for i in range(8)] + [np.full(nb, _IV[i], dtype=np.uint64) for i in range(8)]
    v[12] = v[12] ^ np.uint64(64)
    v[14] = v[14] ^ np.uint64(0xFFFFFFFFFFFFFFFF)

    def rotr(x, n):
        return (x >> np.uint64(n)) | (x << np.uint64(64 - n))

    with np.errstate(over="ignore"):
        for r in range(12):
            s = S.SIGMA[r]
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
    out = np.stack([h[i] ^ v[i] ^ v[i + 8] for i in range(8)], axis=1)  # [nb][8]
    return out.astype("<u8").tobytes()


class FastPRNG:
    """SEAL Blake2xbPRNG with hashlib doing the compressions (checked against oracle/seal_encrypt.py in main())."""

    def __init__(self, seed_bytes: bytes, buf: int = 4096):
        self.seed, self.counter, self.n, self.data, self.pos = seed_bytes, 0, buf, b"", 0
        self._refill()

    def _refill(self):
        root = hashlib.blake2b(struct.pack("<Q", self.counter), digest_size=64, key=self.seed, fanout=1, depth=1,
                               node_offset=self.n << 32).digest()
        self.data = _expand(root, self.n)
        self.counter += 1
        self.pos = 0

    def generate(self, n: int) -> bytes:
        out = b""
        while n:
            take = min(n, self.n - self.pos)
            out += self.data[self.pos:self.pos + take]
            self.pos += take
            n -= take
            if self.pos == self.n:
                self._refill()
        return out

    def u32(self) -> int:
        if self.pos + 4 <= self.n:
            v = int.from_bytes(self.data[self.pos:self.pos + 4], "little")
            self.pos += 4
            if self.pos == self.n:
                self._refill()
            return v
        return int.from_bytes(self.generate(4), "little")


def u3_libcxx(prng):
    """libc++ uniform_int_distribution<uint64_t>(0, 2): __independent_bits_engine with w = 2, reject 3."""
    while True:
        v = prng.u32() & 3
        if v < 3:
            return v


def _canon_libcxx(prng):
    s = float(prng.u32()) + float(prng.u32()) * 4294967296.0
    return s / 18446744073709551616.0


def normal_libcxx(prng, n=F.N, sigma=3.2, max_dev=19.2):
    """libc++ std::normal_distribution (polar; returns u*F first, keeps v*F) under SEAL's ClippedNormalDistribution."""
    out = np.empty(n, dtype=np.int64)
    saved = None
    for i in range(n):
        while True:
            if saved is not None:
                up, saved = saved, None
            else:
                while True:
                    u = 2.0 * _canon_libcxx(prng) + -1.0
                    v = 2.0 * _canon_libcxx(prng) + -1.0
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


def cbd(prng, n=F.N):
    return S.sample_cbd(prng, n)


def lift(e, q):
    return np.where(e < 0, q + e, e).astype(np.uint64)


def encrypt(pk, plain, prng, u3, noise, special, order="u_first"):
    mods = bfv.moduli()[:3]
    L = 3 if special else 2
    if order == "u_first":
        u = S.sample_ternary(prng, u3)
        es = None
    else:
        es = [noise(prng), noise(prng)]
        u = S.sample_ternary(prng, u3)
    c = np.zeros((2, L, F.N), dtype=np.uint64)
    for J in range(L):
        q = mods[J]
        un = bfv.ntt_fwd(lift(u, q), J)
        for j in range(2):
            prod = (un.astype(object) * pk[j, J].astype(object)) % q
            c[j, J] = bfv.ntt_inv(np.array(prod, dtype=np.uint64), J)
    for j in range(2):
        e = es[j] if es else noise(prng)
        for J in range(L):
            q = mods[J]
            c[j, J] = ((c[j, J].astype(object) + lift(e, q).astype(object)) % q).astype(np.uint64)
    if not special:
        return bfv.add_plain(c, plain)
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


def outer_variants(ct):
    bases = ["sunscreen::types::bfv::unsigned::Unsigned", "sunscreen::types::bfv::Unsigned", "sunscreen::types::Unsigned",
             "sunscreen_runtime::types::bfv::unsigned::Unsigned", "Unsigned", "sunscreen::types::bfv::unsigned::Unsigned256",
             "sunscreen::types::bfv::Unsigned256", "sunscreen::types::Unsigned256", "Unsigned256",
             "sunscreen::types::bfv::unsigned256::Unsigned256"]
    suffixes = ["", "<4>", "256", "<256>", "4", "<4usize>", "<4_usize>", "<{4}>", "<LIMBS>"]
    versions = ["0.8.1", "0.8.0"]
    params = F.Params()
    sealct = F.fresh_data_ciphertext(ct)
    for compr in (F.COMPR_ZSTD, F.COMPR_NONE):
        blob = F.seal_wrap(sealct.payload(), compr)
        wc = F.WithContext(params, blob).to_bytes()
        for base, suf, ver in itertools.product(bases, suffixes, versions):
            name = base + suf
            s = f"{name},{ver},true".encode()
            yield (compr, name, ver), struct.pack("<Q", len(s)) + s + struct.pack("<I", 0) + struct.pack("<Q", 1) + wc


def main():
    pub = open(os.path.join(ROOT, "fhe_precompiles_b200/data/network.pub"), "rb").read()
    pri = open(os.path.join(ROOT, "fhe_precompiles_b200/data/network.pri"), "rb").read()
    pk = F.PublicKey.from_bytes(pub).pk_polys()
    sk = F.read_private_key(pri).data
    digest = hashlib.sha512(bytes([1, 2, 3]) + SECRET + (12).to_bytes(32, "big")).digest()
    words = struct.unpack("<8Q", digest)
    # hashlib-backed PRNG == the restated one
    a, b = FastPRNG(digest), S.Blake2xbPRNG(list(words))
    assert a.generate(9000) == b.generate(9000)
    seeds = {"le": digest, "be_words": struct.pack("<8Q", *struct.unpack(">8Q", digest)), "rev": struct.pack("<8Q", *words[::-1])}
    plain = bfv.encode("u256", 12)
    u3s = {"lemire": S.uniform3_lemire, "downscale": S.uniform3_downscale, "libcxx": u3_libcxx}
    noises = {"normal_gnu": S.sample_clipped_normal, "normal_libcxx": normal_libcxx, "cbd": cbd}
    tried = 0
    for (sname, seed), (uname, u3), (nname, noise), special, order in itertools.product(
            seeds.items(), u3s.items(), noises.items(), (True, False), ("u_first", "e_first")):
        if (uname == "libcxx") != (nname == "normal_libcxx") and nname != "cbd":
            continue
        ct = encrypt(pk, plain, FastPRNG(seed), u3, noise, special, order)
        p, budget = bfv.decrypt(ct, sk)
        assert bfv.decode("u256", p) == 12
        tag = (sname, uname, nname, "special" if special else "nospecial", order, budget)
        for key, body in outer_variants(ct):
            tried += 1
            h = hashlib.sha512(body).digest()
            for osname, kat in KAT.items():
                if h == kat:
                    print("MATCH", osname, tag, key, flush=True)
                    return
        print("tried", tag, tried, flush=True)
    print("no match in", tried, "combinations")


if __name__ == "__main__":
    main()
