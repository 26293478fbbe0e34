"""Generates the committed golden fixtures from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  Inputs are seeded; outputs are stored as SHA-256 digests plus two
full byte fixtures, so GPU parity tests and format tests have fixed references that do not depend on
rebuilding the oracle.  No golden ciphertext exists in the reference itself; these pin OUR
restatement against regressions, not SEAL's bytes."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import MODULI, N, KeySet, encrypt_value, random_ct  # noqa: E402
from oracle import bfv  # noqa: E402
from oracle import formats as F  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def main() -> None:
    keys = KeySet.load()
    out = {"note": "sha256 of little-endian u64 arrays; inputs are seeded, see make_golden.py"}
    for m in range(6):
        x = np.random.default_rng(100 + m).integers(0, MODULI[m], size=(2, N), dtype=np.uint64)
        out[f"ntt_fwd_mod{m}_seed{100 + m}"] = digest(bfv.ntt_fwd(x, m))
        out[f"ntt_inv_mod{m}_seed{100 + m}"] = digest(bfv.ntt_inv(x, m))
    a, b = random_ct(np.random.default_rng(1), 1)[0], random_ct(np.random.default_rng(2), 1)[0]
    plain = np.random.default_rng(3).integers(0, 4096, size=N, dtype=np.uint64)
    out["add_seed1_2"] = digest(bfv.add(a, b))
    out["sub_seed1_2"] = digest(bfv.sub(a, b))
    out["negate_seed1"] = digest(bfv.negate(a))
    out["add_plain_seed1_3"] = digest(bfv.add_plain(a, plain))
    out["sub_plain_seed1_3"] = digest(bfv.sub_plain(a, plain))
    out["multiply_plain_seed1_3"] = digest(bfv.multiply_plain(a, plain))
    out["behz_extend_seed1_2"] = digest(bfv.behz_extend(a, b))
    c3 = bfv.multiply(a, b)
    out["multiply_seed1_2"] = digest(c3)
    out["relinearize_seed1_2_testkey"] = digest(bfv.relinearize(c3, keys.rk))
    out["mul_relin_seed1_2_netkey"] = digest(bfv.mul_relin(a, b, keys.net_rk))
    # reference test values (fhe.rs:1721-1731): Signed 16 and 4 under tests/data/public_key.bin
    ca, cb = encrypt_value(keys, "i64", 16, 11), encrypt_value(keys, "i64", 4, 12)
    out["enc_i64_16_seed11"] = digest(ca)
    out["enc_i64_4_seed12"] = digest(cb)
    prod = bfv.mul_relin(ca, cb, keys.rk)
    out["mul_relin_16_4_testkey"] = digest(prod)
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "ct_i64_16_seed11.bin"), "wb") as f:
        f.write(F.make_ciphertext("i64", ca).to_bytes())
    with open(os.path.join(here, "ct_i64_mul_16_4.bin"), "wb") as f:
        f.write(F.make_ciphertext("i64", prod).to_bytes())
    with open(os.path.join(here, "digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", len(out) - 1, "digests and 2 byte fixtures")


if __name__ == "__main__":
    main()
