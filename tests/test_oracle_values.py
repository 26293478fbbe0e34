"""The reference's 48 arithmetic tests (fhe.rs:1038-2076: 16 op 4 -> 20 / 12 / 64 for every type and shape),
re-stated on the oracle, plus an independent exact big-integer BFV multiply that bounds BEHZ's error.  CPU only."""
import numpy as np
import pytest

from helpers import KINDS, MODULI, N, REF_A, REF_B, REF_EXPECT, decrypt_value, encrypt_value, oracle_binary, value_of
from oracle import bfv

Q = MODULI[0] * MODULI[1]
T = 4096


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("op", ("add", "sub", "mul"))
@pytest.mark.parametrize("shape", ("ctct", "ctpt", "ptct"))
def test_reference_values(keys, kind, op, shape):
    a, b = value_of(kind, REF_A), value_of(kind, REF_B)
    ca, cb = encrypt_value(keys, kind, a, 1), encrypt_value(keys, kind, b, 2)
    args = {"ctct": (ca, cb), "ctpt": (ca, b), "ptct": (a, cb)}[shape]
    out = oracle_binary(op, shape, kind, *args, keys.rk)
    assert decrypt_value(keys, kind, out) == value_of(kind, REF_EXPECT[op])


@pytest.mark.parametrize("kind,a,b", [("i64", -16, 4), ("i64", -(2**31), 2**31 - 1), ("u64", 2**40 + 3, 2**20 + 1),
                                      ("u256", 2**128 + 7, 2**100 + 9), ("frac64", 1.5, -2.25), ("frac64", 0.0, 3.0)])
def test_more_values(keys, kind, a, b):
    ca, cb = encrypt_value(keys, kind, a, 3), encrypt_value(keys, kind, b, 4)
    wrap = {"i64": lambda v: (v + 2**63) % 2**64 - 2**63, "u64": lambda v: v % 2**64, "u256": lambda v: v % 2**256, "frac64": float}[kind]
    assert decrypt_value(keys, kind, bfv.add(ca, cb)) == wrap(a + b)
    assert decrypt_value(keys, kind, bfv.sub(ca, cb)) == wrap(a - b)
    assert decrypt_value(keys, kind, bfv.mul_relin(ca, cb, keys.rk)) == wrap(a * b)
    assert decrypt_value(keys, kind, bfv.multiply_plain(ca, bfv.encode(kind, b))) == wrap(a * b)
    assert decrypt_value(keys, kind, bfv.negate(bfv.sub_plain(cb, bfv.encode(kind, a)))) == wrap(a - b)


def encoders_roundtrip_cases():
    return [("i64", v) for v in (0, 1, -1, 12, -(2**63) + 1, 2**63 - 1)] + [("u64", v) for v in (0, 1, 2**64 - 1)] + [
        ("u256", v) for v in (0, 2**255 + 12345, 2**256 - 1)
    ] + [("frac64", v) for v in (0.0, 12.0, -0.375, 3.141592653589793, -1e15, 2.0**-40)]


@pytest.mark.parametrize("kind,v", encoders_roundtrip_cases())
def test_encoders_roundtrip(kind, v):
    p = bfv.encode(kind, v)
    assert (p < T).all()
    assert bfv.decode(kind, p) == v


def negacyclic_mul_exact(a, b):
    """exact negacyclic product of two integer polynomials via Kronecker substitution (python big ints)"""
    bits = 2 * 74 + 14  # |coeff| < 2^73 each, N terms
    base = 1 << bits
    half = base >> 1

    def pack(p):
        acc = 0
        for c in reversed(p):
            acc = acc * base + c
        return acc

    prod = pack(a) * pack(b)
    coeffs = []
    for _ in range(2 * N - 1):
        c = prod % base
        if c >= half:
            c -= base
        prod = (prod - c) // base
        coeffs.append(c)
    coeffs.append(0)
    return [coeffs[i] - coeffs[i + N] for i in range(N)]


def test_behz_multiply_is_exact_scaled_tensor_up_to_small_error(keys):
    """round(t/q * (a (x) b)) mod q computed exactly vs the oracle's BEHZ multiply: BEHZ's fast base conversions may
    add a small integer error per coefficient (it is approximate by design) but nothing more."""
    ca, cb = encrypt_value(keys, "i64", 1234, 5), encrypt_value(keys, "i64", -77, 6)
    got = bfv.multiply(ca, cb)

    def lift(ct):  # centred CRT lift of each polynomial
        polys = []
        for p in range(2):
            x0, x1 = [int(v) for v in ct[p, 0]], [int(v) for v in ct[p, 1]]
            i0, i1 = pow(MODULI[1], -1, MODULI[0]), pow(MODULI[0], -1, MODULI[1])
            out = []
            for u, v in zip(x0, x1):
                X = (u * i0 % MODULI[0] * MODULI[1] + v * i1 % MODULI[1] * MODULI[0]) % Q
                out.append(X - Q if X > Q // 2 else X)
            polys.append(out)
        return polys

    A, B = lift(ca), lift(cb)
    d0 = negacyclic_mul_exact(A[0], B[0])
    d1 = [x + y for x, y in zip(negacyclic_mul_exact(A[0], B[1]), negacyclic_mul_exact(A[1], B[0]))]
    d2 = negacyclic_mul_exact(A[1], B[1])
    worst = 0
    for k, d in enumerate((d0, d1, d2)):
        for i in range(0, N, 7):
            exact = (T * d[i] + Q // 2) // Q  # round(t*d/q)
            g = (int(got[k, 0, i]) * pow(MODULI[1], -1, MODULI[0]) % MODULI[0] * MODULI[1]
                 + int(got[k, 1, i]) * pow(MODULI[0], -1, MODULI[1]) % MODULI[1] * MODULI[0]) % Q
            diff = (g - exact) % Q
            diff = diff - Q if diff > Q // 2 else diff
            worst = max(worst, abs(diff))
    assert worst <= 8, worst
