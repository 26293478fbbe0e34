"""The integer-domain form of the BEHZ base conversions (k_ext_conv / k_floor_sk) equals SEAL's step-by-step form
(fastbconv_m_tilde + sm_mrq, fast_floor + fastbconv_sk) -- CPU check of the algebra the kernels rely on."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("check_integer_domain", os.path.join(ROOT, "scripts", "check_integer_domain.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_base_extension_integer_domain():
    _load().check_ext(20000)


def test_fast_floor_sk_integer_domain():
    _load().check_floor(20000)


def _load_dual():
    spec = importlib.util.spec_from_file_location("check_dual_base", os.path.join(ROOT, "scripts", "check_dual_base.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_dual_base_is_large_enough():
    """the six primes of params.h (kDualPrime) are the largest NTT primes below 2^30 and carry every integer the kernels put
    on them: |t D| < 2^166 and |U_k| < 2^161 against S / 2 ~ 2^179, |f| < 2^96 against S4 / 2 ~ 2^119; every prime is within
    2^-12 of 2^30 (what lets the kernels take y / s from a shift)"""
    import math

    primes = _load_dual().dual_primes()
    assert primes == [0x3FFF4001, 0x3FFEE001, 0x3FFEA001, 0x3FFE8001, 0x3FFD6001, 0x3FFC0001]
    q = 0xFFFFEE001 * 0xFFFFC4001
    Q = q * 0x1FFFFE0001
    tD = 4096 * 4096 * (q // 2 + (q >> 29)) ** 2
    U = 2 * 4096 * (1 << 36) * 3 * Q
    assert math.prod(primes) > 2 ** 12 * 2 * tD and math.prod(primes) > 2 ** 12 * 2 * U
    assert math.prod(primes[:4]) > 2 ** 12 * 2 * (tD // q + 2)
    assert all((2 ** 30 - p) / 2 ** 30 < 2 ** -12 for p in primes) and all(4 * p < 2 ** 32 for p in primes)


def test_dual_base_multiply_and_key_switch_equal_seals(keys):
    """one real ciphertext pair through the integer formulation on the dual base (scripts/check_dual_base.py): the size-3 product
    and the relinearised ciphertext equal the oracle's (SEAL's form) bit for bit"""
    import numpy as np

    from helpers import encrypt_value

    m = _load_dual()
    a, b = encrypt_value(keys, "i64", 5, 1), encrypt_value(keys, "i64", -9, 2)
    assert m.check_multiply(a, b) == 0
    assert m.check_key_switch(np.concatenate([a, b[:1]]), keys.rk) == 0
