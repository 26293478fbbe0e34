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


# ---------------------------------------------------------------- restatement of the GPU encryptor's sampler
# (fhe_precompiles_b200/csrc/kernels.cu: chacha12_block, sample_ternary, sample_noise).  The reference hands SEAL the
# 512-bit SHA-512 digest as PRNG seed (/root/reference/src/fhe.rs:611-616); SEAL's own Blake2xb stream is not reproduced
# (SURVEY 8f-1), so the product expands the same seed with ChaCha12 and this is its bit-exact CPU twin.
NOISE_CDF = np.array([
    0x1f67485e1414e200, 0x3be85f5582810200, 0x53644e2dedd21400, 0x64f422f09cf1bc00, 0x70dfcc250f890800,
    0x7837f1b047d3fc00, 0x7c535c45b5071400, 0x7e690b1eb1011400, 0x7f5eeb470d610c00, 0x7fc5bca5a5143c00,
    0x7fecc2f990af3800, 0x7ffa349ee365e800, 0x7ffe68c004b14800, 0x7fff9a26cfa95400, 0x7fffe8d1193b7400,
    0x7ffffb3514071000, 0x7fffff1c06e24c00, 0x7fffffdc665b1800, 0x7ffffffe05c3f800], dtype=np.uint64)


def chacha_core(x: np.ndarray, double_rounds: int) -> np.ndarray:
    """ChaCha block function (RFC 7539 2.3) on columns of x (uint32 [16, n]) with 2 * double_rounds rounds."""
    x = x.astype(np.uint32)
    w = x.copy()
    rotl = lambda v, n: (v << np.uint32(n)) | (v >> np.uint32(32 - n))

    def qr(a, b, c, d):
        w[a] += w[b]; w[d] = rotl(w[d] ^ w[a], 16)
        w[c] += w[d]; w[b] = rotl(w[b] ^ w[c], 12)
        w[a] += w[b]; w[d] = rotl(w[d] ^ w[a], 8)
        w[c] += w[d]; w[b] = rotl(w[b] ^ w[c], 7)

    with np.errstate(over="ignore"):
        for _ in range(double_rounds):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        w += x
    return w


def _chacha12_blocks(seed8: np.ndarray, stream: int) -> np.ndarray:
    """[512 counters][8] u64 output words under the 512-bit seed (8 u64 words): key = words 0..3 xor 4..7, state words
    12 / 13 = (counter, stream), 14 / 15 = seed word 4; 12 rounds."""
    seed8 = np.asarray(seed8, dtype=np.uint64)
    key = seed8[:4] ^ seed8[4:]
    x = np.zeros((16, 512), dtype=np.uint32)
    x[0], x[1], x[2], x[3] = 0x61707865, 0x3320646E, 0x79622D32, 0x6B206574
    for i in range(4):
        x[4 + 2 * i] = np.uint32(int(key[i]) & 0xFFFFFFFF)
        x[5 + 2 * i] = np.uint32(int(key[i]) >> 32)
    x[12] = np.arange(512, dtype=np.uint32)
    x[13] = stream
    x[14] = np.uint32(int(seed8[4]) & 0xFFFFFFFF)
    x[15] = np.uint32(int(seed8[4]) >> 32)
    w = chacha_core(x, 6)
    out = w[0::2].astype(np.uint64) | (w[1::2].astype(np.uint64) << np.uint64(32))  # [8][512]
    return out.T.copy()


def gpu_sampler(seed8) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(u, e0, e1) int8 [N] exactly as k_encrypt_core draws them: word r of thread t's block -> coefficient r*512 + t."""
    res = []
    for stream in range(3):
        words = _chacha12_blocks(seed8, stream).T.reshape(-1)  # index r*512 + t
        if stream == 0:
            smp = np.zeros(N, dtype=np.int8)
            done = np.zeros(N, dtype=bool)
            for k in range(32):
                d = ((words >> np.uint64(2 * k)) & np.uint64(3)).astype(np.int8)
                take = ~done & (d != 3)
                smp[take] = d[take] - 1
                done |= take
        else:
            v = words >> np.uint64(1)
            mag = (v[:, None] >= NOISE_CDF[None, :]).sum(axis=1).astype(np.int8)
            smp = np.where((words & np.uint64(1)) == 1, -mag, mag).astype(np.int8)
        res.append(smp)
    return tuple(res)


def encrypt_seeded(pk: np.ndarray, plain: np.ndarray, seed8) -> np.ndarray:
    """What fhe_b200_encrypt / c_fhe_encrypt_* compute for the 512-bit seed (8 u64 words)."""
    pk, pl = _c(pk), _c(plain)
    u, e0, e1 = (np.ascontiguousarray(x) for x in gpu_sampler(seed8))
    out = np.empty((2, 2, N), dtype=np.uint64)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib().bfvo_encrypt_samples(_p(pk), _p(pl), pl.size, vp(u), vp(e0), vp(e1), _p(out))
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
