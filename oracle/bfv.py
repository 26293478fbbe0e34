"""TEST INFRASTRUCTURE ONLY: ctypes binding of liboracle.so (see bfv_oracle.h)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Tuple

import numpy as np

N = 4096
T = 4096
MOD_NAMES = ("q0", "q1", "P", "b0", "b1", "msk")
_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

_u64p = ctypes.POINTER(ctypes.c_uint64)


def build() -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return os.path.join(_HERE, "liboracle.so")


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.bfvo_init.restype = ctypes.c_int
        L.bfvo_constants.restype = ctypes.c_size_t
        for name in ("bfvo_encode_i64", "bfvo_encode_u64", "bfvo_encode_u256", "bfvo_encode_f64"):
            getattr(L, name).restype = ctypes.c_size_t
        L.bfvo_encode_i64.argtypes = [ctypes.c_int64, ctypes.c_void_p]
        L.bfvo_encode_u64.argtypes = [ctypes.c_uint64, ctypes.c_void_p]
        L.bfvo_encode_u256.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.bfvo_encode_f64.argtypes = [ctypes.c_double, ctypes.c_void_p]
        L.bfvo_decode_i64.restype = ctypes.c_int64
        L.bfvo_decode_i64.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.bfvo_decode_u64.restype = ctypes.c_uint64
        L.bfvo_decode_u64.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.bfvo_decode_u256.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.bfvo_decode_f64.restype = ctypes.c_double
        L.bfvo_decode_f64.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.bfvo_decrypt.restype = ctypes.c_int
        L.bfvo_decrypt.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        L.bfvo_encrypt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_void_p]
        L.bfvo_encrypt_samples.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t] + [ctypes.c_void_p] * 4
        L.bfvo_seal_prng.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.bfvo_seal_sample.restype = ctypes.c_size_t
        L.bfvo_seal_sample.argtypes = [ctypes.c_void_p] * 4
        L.bfvo_seal_sample_stream.restype = ctypes.c_size_t
        L.bfvo_seal_sample_stream.argtypes = [ctypes.c_void_p, ctypes.c_size_t] + [ctypes.c_void_p] * 3
        L.bfvo_encrypt_samples_data_level.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t] + [ctypes.c_void_p] * 4
        L.bfvo_seal_encrypt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        L.bfvo_batch_mul_relin.restype = ctypes.c_double
        L.bfvo_batch_mul_relin.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_int]
        L.bfvo_batch_ntt.restype = ctypes.c_double
        L.bfvo_batch_ntt.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        for name in ("bfvo_add", "bfvo_sub"):
            getattr(L, name).argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_size_t]
        L.bfvo_negate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.bfvo_add_plain.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.bfvo_sub_plain.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.bfvo_multiply_plain.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.bfvo_multiply.argtypes = [ctypes.c_void_p] * 3
        L.bfvo_relinearize.argtypes = [ctypes.c_void_p] * 3
        L.bfvo_mul_relin.argtypes = [ctypes.c_void_p] * 4
        L.bfvo_behz_extend.argtypes = [ctypes.c_void_p] * 3
        L.bfvo_behz_tensor.argtypes = [ctypes.c_void_p] * 2
        L.bfvo_behz_floor_sk.argtypes = [ctypes.c_void_p] * 2
        L.bfvo_ntt_fwd.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.bfvo_ntt_inv.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.bfvo_constants.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.bfvo_init()
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def constants() -> dict:
    buf = np.zeros(128, dtype=np.uint64)
    n = lib().bfvo_constants(_p(buf), 128)
    v = [int(x) for x in buf[:n]]
    it = iter(v)
    take = lambda k: [next(it) for _ in range(k)]
    d = {}
    d["moduli"] = take(6)
    d["roots"] = take(6)
    d["gamma"] = take(1)[0]
    d["ninv"] = take(6)
    d["delta_mod_q"] = take(2)
    d["inv_P_mod_q"] = take(2)
    d["half_P"] = take(1)[0]
    d["mtilde_mod_q"] = take(2)
    d["inv_punct_q"] = take(2)
    d["punct_q_mod_bsk"] = [take(3), take(3)]
    d["punct_q_mod_mtilde"] = take(2)
    d["neg_inv_q_mod_mtilde"] = take(1)[0]
    d["q_mod_bsk"] = take(3)
    d["inv_mtilde_mod_bsk"] = take(3)
    d["inv_q_mod_bsk"] = take(3)
    d["inv_punct_B"] = take(2)
    d["punct_B_mod_q"] = [take(2), take(2)]
    d["punct_B_mod_msk"] = take(2)
    d["inv_B_mod_msk"] = take(1)[0]
    d["B_mod_q"] = take(2)
    return d


def moduli() -> Tuple[int, ...]:
    return tuple(constants()["moduli"])


def ntt_fwd(a: np.ndarray, mod: int) -> np.ndarray:
    out = _c(a).copy()
    flat = out.reshape(-1, N)
    for row in flat:
        lib().bfvo_ntt_fwd(_p(row), mod)
    return out


def ntt_inv(a: np.ndarray, mod: int) -> np.ndarray:
    out = _c(a).copy()
    flat = out.reshape(-1, N)
    for row in flat:
        lib().bfvo_ntt_inv(_p(row), mod)
    return out


def add(a, b):
    a, b = _c(a), _c(b)
    out = np.empty_like(a)
    lib().bfvo_add(_p(a), _p(b), _p(out), a.size // (2 * N))
    return out


def sub(a, b):
    a, b = _c(a), _c(b)
    out = np.empty_like(a)
    lib().bfvo_sub(_p(a), _p(b), _p(out), a.size // (2 * N))
    return out


def negate(a):
    a = _c(a)
    out = np.empty_like(a)
    lib().bfvo_negate(_p(a), _p(out), a.size // (2 * N))
    return out


def add_plain(ct, plain):
    out = _c(ct).copy()
    pl = _c(plain)
    lib().bfvo_add_plain(_p(out), _p(pl), pl.size)
    return out


def sub_plain(ct, plain):
    out = _c(ct).copy()
    pl = _c(plain)
    lib().bfvo_sub_plain(_p(out), _p(pl), pl.size)
    return out


def multiply_plain(ct, plain):
    out = _c(ct).copy()
    pl = _c(plain)
    lib().bfvo_multiply_plain(_p(out), out.size // (2 * N), _p(pl), pl.size)
    return out


def multiply(a, b):
    a, b = _c(a), _c(b)
    out = np.empty((3, 2, N), dtype=np.uint64)
    lib().bfvo_multiply(_p(a), _p(b), _p(out))
    return out


def relinearize(ct3, rk):
    ct3, rk = _c(ct3), _c(rk)
    out = np.empty((2, 2, N), dtype=np.uint64)
    lib().bfvo_relinearize(_p(ct3), _p(rk), _p(out))
    return out


def mul_relin(a, b, rk):
    a, b, rk = _c(a), _c(b), _c(rk)
    out = np.empty((2, 2, N), dtype=np.uint64)
    lib().bfvo_mul_relin(_p(a), _p(b), _p(rk), _p(out))
    return out


def behz_extend(a, b):
    a, b = _c(a), _c(b)
    out = np.empty((4, 5, N), dtype=np.uint64)
    lib().bfvo_behz_extend(_p(a), _p(b), _p(out))
    return out


def behz_tensor(ext):
    ext = _c(ext)
    out = np.empty((3, 5, N), dtype=np.uint64)
    lib().bfvo_behz_tensor(_p(ext), _p(out))
    return out


def behz_floor_sk(tens):
    tens = _c(tens)
    out = np.empty((3, 2, N), dtype=np.uint64)
    lib().bfvo_behz_floor_sk(_p(tens), _p(out))
    return out


def encode(kind: str, value) -> np.ndarray:
    buf = np.zeros(N, dtype=np.uint64)
    L = lib()
    if kind == "i64":
        n = L.bfvo_encode_i64(int(value), _p(buf))
    elif kind == "u64":
        n = L.bfvo_encode_u64(int(value) & (2**64 - 1), _p(buf))
    elif kind == "u256":
        v = int(value) & (2**256 - 1)
        limbs = np.array([(v >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)
        n = L.bfvo_encode_u256(_p(limbs), _p(buf))
    elif kind == "frac64":
        n = L.bfvo_encode_f64(float(value), _p(buf))
        if n == 0:
            raise ValueError("frac64 out of range")
    else:
        raise KeyError(kind)
    return buf[:n].copy()


def decode(kind: str, plain: np.ndarray):
    pl = _c(plain)
    L = lib()
    if kind == "i64":
        return int(L.bfvo_decode_i64(_p(pl), pl.size))
    if kind == "u64":
        return int(L.bfvo_decode_u64(_p(pl), pl.size))
    if kind == "u256":
        limbs = np.zeros(4, dtype=np.uint64)
        L.bfvo_decode_u256(_p(pl), pl.size, _p(limbs))
        return sum(int(limbs[i]) << (64 * i) for i in range(4))
    if kind == "frac64":
        return float(L.bfvo_decode_f64(_p(pl), pl.size))
    raise KeyError(kind)


def encrypt(pk: np.ndarray, plain: np.ndarray, seed: int) -> np.ndarray:
    pk, pl = _c(pk), _c(plain)
    out = np.empty((2, 2, N), dtype=np.uint64)
    lib().bfvo_encrypt(_p(pk), _p(pl), pl.size, seed, _p(out))
    return out


# ---------------------------------------------------------------- SEAL-exact deterministic encryption
# What FheApp::encrypt / reencrypt (/root/reference/src/fhe.rs:594-657) obtain from sunscreen's `encrypt_deterministic`:
# Blake2xb PRNG keyed by the SHA-512 digest, libstdc++ ternary + clipped-normal samplers, encryption at the data level
# (no special modulus).  C restatement in bfv_oracle.c, pinned by the reference's SHA-512 known answers
# (tests/test_oracle_kat.py); oracle/seal_encrypt.py is an independent pure-Python restatement of the same stack.
def _seed8(seed8) -> np.ndarray:
    if isinstance(seed8, (bytes, bytearray)):
        assert len(seed8) == 64
        return np.frombuffer(bytes(seed8), dtype="<u8").astype(np.uint64)
    a = np.ascontiguousarray(seed8, dtype=np.uint64)
    assert a.size == 8
    return a


def seal_prng(seed8, nbytes: int) -> bytes:
    """First nbytes of SEAL's Blake2xbPRNG stream for the 512-bit seed (8 LE u64 words or the 64 digest bytes)."""
    out = np.empty(nbytes, dtype=np.uint8)
    lib().bfvo_seal_prng(_p(_seed8(seed8)), out.ctypes.data_as(ctypes.c_void_p), nbytes)
    return out.tobytes()


def seal_sample(seed8) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """(u, e0, e1, number of 32-bit draws consumed): SEAL's sample_poly_ternary, then sample_poly_normal twice."""
    smp = [np.empty(N, dtype=np.int8) for _ in range(3)]
    vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    drawn = lib().bfvo_seal_sample(_p(_seed8(seed8)), vp(smp[0]), vp(smp[1]), vp(smp[2]))
    return smp[0], smp[1], smp[2], int(drawn)


def seal_sample_stream(words: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """the same samplers on a caller-supplied stream of 32-bit draws (rare-path tests); drawn = 0: the stream ran out"""
    w = np.ascontiguousarray(words, dtype=np.uint32)
    smp = [np.empty(N, dtype=np.int8) for _ in range(3)]
    vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    drawn = lib().bfvo_seal_sample_stream(vp(w), w.size, vp(smp[0]), vp(smp[1]), vp(smp[2]))
    return smp[0], smp[1], smp[2], int(drawn)


def encrypt_samples_data_level(pk: np.ndarray, plain: np.ndarray, u, e0, e1) -> np.ndarray:
    pk, pl = _c(pk), _c(plain)
    out = np.empty((2, 2, N), dtype=np.uint64)
    vp = lambda x: np.ascontiguousarray(x, dtype=np.int8).ctypes.data_as(ctypes.c_void_p)
    u, e0, e1 = (np.ascontiguousarray(x, dtype=np.int8) for x in (u, e0, e1))
    lib().bfvo_encrypt_samples_data_level(_p(pk), _p(pl), pl.size, vp(u), vp(e0), vp(e1), _p(out))
    return out


def encrypt_seeded(pk: np.ndarray, plain: np.ndarray, seed8) -> np.ndarray:
    """sunscreen `encrypt_deterministic(plain, pk, seed)`: what c_fhe_encrypt_* / fhe_b200_encrypt must produce, bit for bit."""
    pk, pl = _c(pk), _c(plain)
    out = np.empty((2, 2, N), dtype=np.uint64)
    lib().bfvo_seal_encrypt(_p(pk), _p(pl), pl.size, _p(_seed8(seed8)), _p(out))
    return out


def decrypt(ct: np.ndarray, sk: np.ndarray):
    ct, sk = _c(ct), _c(sk)
    out = np.empty(N, dtype=np.uint64)
    budget = lib().bfvo_decrypt(_p(ct), ct.size // (2 * N), _p(sk), _p(out))
    return out, budget


def batch_mul_relin(a, b, rk, threads: int):
    a, b, rk = _c(a), _c(b), _c(rk)
    n = a.size // (4 * N)
    out = np.empty((n, 2, 2, N), dtype=np.uint64)
    secs = lib().bfvo_batch_mul_relin(_p(a), _p(b), _p(rk), _p(out), n, threads)
    return out, secs


def batch_ntt(limbs, mod: int, inverse: bool, threads: int):
    out = _c(limbs).copy()
    secs = lib().bfvo_batch_ntt(_p(out), out.size // N, mod, int(inverse), threads)
    return out, secs


def rk_array(relin_keys) -> np.ndarray:
    """formats.RelinKeys -> [digit][poly][limb][N] u64 (the layout bfvo_relinearize takes)."""
    row = relin_keys.keys[0]
    return np.stack([k.polys() for k in row]).astype(np.uint64)
