"""Pins the CPU oracle against every fixture the reference holds for the hot path:
the four key files.  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import MODULI, N, ROOT
from oracle import bfv
from oracle import formats as F

P = MODULI[2]


def centred(x: np.ndarray, q: int) -> np.ndarray:
    x = x.astype(np.int64)
    return np.where(x > q // 2, x - q, x)


def test_constants_match_seal_rules():
    c = bfv.constants()
    # testnet.rs:8-14 and SEAL's get_primes(2N, 61, 4): m_sk, gamma, then B
    assert c["moduli"] == list(MODULI)
    assert c["gamma"] == 0x1FFFFFFFFFFCE001
    for q in c["moduli"] + [c["gamma"]]:
        assert q % (2 * N) == 1
    # minimal primitive 2N-th roots
    assert c["roots"][:3] == [0x1720701, 0x1BAA271, 0x839EB4]
    for q, r in zip(c["moduli"], c["roots"]):
        assert pow(r, N, q) == q - 1  # primitive 2N-th root
        cof = (q - 1) // (2 * N)
        # minimal among all primitive roots: no smaller x with x^N == -1
        assert not any(pow(x, N, q) == q - 1 for x in range(2, min(r, 3000)))
    q = MODULI[0] * MODULI[1]
    assert c["delta_mod_q"] == [(q // 4096) % MODULI[0], (q // 4096) % MODULI[1]]
    assert c["neg_inv_q_mod_mtilde"] == (-pow(q, -1, 2**32)) % 2**32 == 0x73FB1FFF
    bsk = [MODULI[3], MODULI[4], MODULI[5]]
    assert c["inv_mtilde_mod_bsk"] == [pow(2**32, -1, p) for p in bsk]
    assert c["inv_q_mod_bsk"] == [pow(q, -1, p) for p in bsk]
    B = MODULI[3] * MODULI[4]
    assert c["inv_B_mod_msk"] == pow(B, -1, MODULI[5]) == 0x10DEFF4D9522A795
    assert c["B_mod_q"] == [B % MODULI[0], B % MODULI[1]]
    assert c["inv_P_mod_q"] == [pow(P, -1, MODULI[0]), pow(P, -1, MODULI[1])]


def test_parms_id_rule_matches_fixtures(keys):
    pk = F.PublicKey.from_bytes(keys.pub_bytes)
    payload, _ = F.seal_unwrap(pk.public_key.blob)
    assert F.SealCiphertext.from_payload(payload).parms_id == F.PARMS_ID_KEY
    assert [hex(x) for x in F.PARMS_ID_KEY] == ["0xee4a95b2988fe663", "0xd83c0c7f1720cce7", "0xda77d21b876b2345", "0x9f975d66df8e0564"]


@pytest.mark.parametrize("mod", range(6))
def test_ntt_is_evaluation_at_odd_powers_in_bit_reversed_order(mod):
    """SEAL's convention: out[k] = a(psi^(2*bitrev(k)+1)); checked against direct evaluation with python ints."""
    q = MODULI[mod]
    psi = bfv.constants()["roots"][mod]
    rng = np.random.default_rng(mod)
    a = rng.integers(0, q, size=N, dtype=np.uint64)
    out = bfv.ntt_fwd(a, mod)
    ai = [int(x) for x in a]
    for k in (0, 1, 2, 5, 2048, 4095):
        e = 2 * int(format(k, "012b")[::-1], 2) + 1
        x = pow(psi, e, q)
        acc = 0
        for coeff in reversed(ai):
            acc = (acc * x + coeff) % q
        assert acc == int(out[k])
    assert np.array_equal(bfv.ntt_inv(out, mod), a)


@pytest.mark.parametrize("which", ["tests", "network"])
def test_key_fixtures_are_consistent_under_the_oracle(keys, which):
    """sk is ternary; pk0 + pk1*s and rk[j].c0 + rk[j].c1*s - [l=j] P s^2 are small: fixes primes, roots, NTT
    ordering, key layout and the key-switch digit convention."""
    pk, rk, sk = (keys.pk, keys.rk, keys.sk) if which == "tests" else (keys.net_pk, keys.net_rk, keys.net_sk)
    s_coeff = [centred(bfv.ntt_inv(sk[l], l), MODULI[l]) for l in range(3)]
    assert set(np.unique(s_coeff[0])) == {-1, 0, 1}
    assert np.array_equal(s_coeff[0], s_coeff[1]) and np.array_equal(s_coeff[0], s_coeff[2])
    for l in range(3):
        q = MODULI[l]
        e = (pk[0, l].astype(object) + pk[1, l].astype(object) * sk[l].astype(object)) % q
        e = centred(bfv.ntt_inv(np.array(e, dtype=np.uint64), l), q)
        assert np.abs(e).max() <= 21
        s2 = (sk[l].astype(object) * sk[l].astype(object)) % q
        for j in range(2):
            v = rk[j, 0, l].astype(object) + rk[j, 1, l].astype(object) * sk[l].astype(object)
            if l == j:
                v = v - (P % q) * s2
            e = centred(bfv.ntt_inv(np.array(v % q, dtype=np.uint64), l), q)
            assert np.abs(e).max() <= 21, (j, l)


def test_golden_digests_are_stable(keys):
    """the committed digests (tests/golden/make_golden.py) still describe what the oracle computes"""
    from golden import make_golden  # noqa: F401  (import check only)

    want = json.load(open(os.path.join(ROOT, "tests/golden/digests.json")))
    dig = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()
    from helpers import encrypt_value, random_ct

    for m in range(6):
        x = np.random.default_rng(100 + m).integers(0, MODULI[m], size=(2, N), dtype=np.uint64)
        assert dig(bfv.ntt_fwd(x, m)) == want[f"ntt_fwd_mod{m}_seed{100 + m}"]
    a, b = random_ct(np.random.default_rng(1), 1)[0], random_ct(np.random.default_rng(2), 1)[0]
    assert dig(bfv.multiply(a, b)) == want["multiply_seed1_2"]
    assert dig(bfv.mul_relin(a, b, keys.net_rk)) == want["mul_relin_seed1_2_netkey"]
    ca, cb = encrypt_value(keys, "i64", 16, 11), encrypt_value(keys, "i64", 4, 12)
    assert dig(bfv.mul_relin(ca, cb, keys.rk)) == want["mul_relin_16_4_testkey"]
    fixture = open(os.path.join(ROOT, "tests/golden/ct_i64_mul_16_4.bin"), "rb").read()
    assert F.make_ciphertext("i64", bfv.mul_relin(ca, cb, keys.rk)).to_bytes() == fixture
