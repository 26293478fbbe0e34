"""Shared test helpers: key fixtures, oracle-built inputs, expected values of the reference's tests."""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from oracle import bfv
from oracle import formats as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 4096
KINDS = ("u256", "u64", "i64", "frac64")
MODULI = (0xFFFFEE001, 0xFFFFC4001, 0x1FFFFE0001, 0x1FFFFFFFFFFA4001, 0x1FFFFFFFFFF92001, 0x1FFFFFFFFFFDE001)


@dataclass
class KeySet:
    pub_bytes: bytes
    pri_bytes: bytes
    pk: np.ndarray  # [2][3][N]
    rk: np.ndarray  # [2][2][3][N]
    sk: np.ndarray  # [3][N]
    net_pub_bytes: bytes
    net_pri_bytes: bytes
    net_pk: np.ndarray
    net_rk: np.ndarray
    net_sk: np.ndarray

    @staticmethod
    def load() -> "KeySet":
        rd = lambda p: open(os.path.join(ROOT, p), "rb").read()
        pub, pri = rd("tests/data/public_key.bin"), rd("tests/data/private_key.bin")
        npub, npri = rd("fhe_precompiles_b200/data/network.pub"), rd("fhe_precompiles_b200/data/network.pri")
        pkf, npkf = F.PublicKey.from_bytes(pub), F.PublicKey.from_bytes(npub)
        return KeySet(
            pub, pri, pkf.pk_polys(), bfv.rk_array(pkf.relin()), F.read_private_key(pri).data,
            npub, npri, npkf.pk_polys(), bfv.rk_array(npkf.relin()), F.read_private_key(npri).data,
        )


def encrypt_value(keys: KeySet, kind: str, value, seed: int, network: bool = False) -> np.ndarray:
    pk = keys.net_pk if network else keys.pk
    return bfv.encrypt(pk, bfv.encode(kind, value), seed)


def decrypt_value(keys: KeySet, kind: str, ct: np.ndarray, network: bool = False):
    sk = keys.net_sk if network else keys.sk
    plain, budget = bfv.decrypt(ct, sk)
    assert budget > 0, "noise budget exhausted"
    return bfv.decode(kind, plain)


def plain_u16(kind: str, value) -> np.ndarray:
    out = np.zeros(N, dtype=np.uint16)
    p = bfv.encode(kind, value)
    out[: len(p)] = p.astype(np.uint16)
    return out


def value_of(kind: str, v):
    return float(v) if kind == "frac64" else int(v)


# the values of the reference's 48 arithmetic tests (fhe.rs:1038-2076): a = 16, b = 4
REF_A, REF_B = 16, 4
REF_EXPECT = {"add": 20, "sub": 12, "mul": 64}


def oracle_binary(op: str, shape: str, kind: str, a, b, rk):
    """What the reference computes for precompile `op` on (a, b); ct operands are [2][2][N] arrays,
    plaintext operands are python scalars. Follows the per-op SEAL call map of SURVEY 3.1."""
    if shape == "ctct":
        if op == "add":
            return bfv.add(a, b)
        if op == "sub":
            return bfv.sub(a, b)
        return bfv.mul_relin(a, b, rk)
    ct, pt = (a, b) if shape == "ctpt" else (b, a)
    plain = bfv.encode(kind, pt)
    if op == "add":
        return bfv.add_plain(ct, plain)
    if op == "mul":
        return bfv.multiply_plain(ct, plain)
    r = bfv.sub_plain(ct, plain)
    return r if shape == "ctpt" else bfv.negate(r)


def precompile_name(op: str, shape: str, kind: str) -> str:
    if shape == "ctct":
        return f"{op}_cipher{kind}_cipher{kind}"
    if shape == "ctpt":
        return f"{op}_cipher{kind}_{kind}"
    return f"{op}_{kind}_cipher{kind}"


def random_ct(rng: np.random.Generator, n: int) -> np.ndarray:
    """[n][2][2][N] uniform residues (not a valid encryption; exercises full-range arithmetic)."""
    out = np.empty((n, 2, 2, N), dtype=np.uint64)
    for l in range(2):
        out[:, :, l, :] = rng.integers(0, MODULI[l], size=(n, 2, N), dtype=np.uint64)
    return out
