"""The oracle's restatement of the product's encryption sampler (oracle/bfv.py: chacha_core, gpu_sampler, encrypt_seeded;
product: csrc/kernels.cu chacha12_block / sample_ternary / sample_noise).  CPU only; the GPU side is bit-compared with this
in tests/test_gpu_parity.py::test_device_encrypt_matches_oracle."""
import numpy as np

from helpers import decrypt_value
from oracle import bfv


def test_chacha_core_is_rfc7539():
    """RFC 7539 section 2.3.2 test vector (20 rounds): the 12-round generator is the same function with fewer rounds."""
    x = np.zeros((16, 1), dtype=np.uint32)
    x[:4, 0] = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574]
    x[4:12, 0] = np.frombuffer(bytes(range(32)), dtype="<u4")
    x[12, 0] = 1
    x[13:16, 0] = np.frombuffer(bytes.fromhex("000000090000004a00000000"), dtype="<u4")
    want = ("e4e7f110 15593bd1 1fdd0f50 c47120a3 c7f4d1c7 0368c033 9aaa2204 4e6cd4c3 "
            "466482d2 09aa9f07 05d7c214 a2028bd9 d19c12b5 b94e16de e883d0cb 4e3c50a2")
    assert " ".join("%08x" % v for v in bfv.chacha_core(x, 10)[:, 0]) == want


def test_sampler_distributions_match_the_key_fixtures():
    """u uniform ternary; errors Gaussian sigma 3.2 clipped at 6 sigma, truncated toward zero: P(0) = 0.245, variance 8.0 --
    the statistics of the errors inside the reference's four key files."""
    rng = np.random.default_rng(1)
    us, es = [], []
    for _ in range(24):
        u, e0, e1 = bfv.gpu_sampler(rng.integers(0, 2**63, size=8, dtype=np.uint64))
        us.append(u)
        es += [e0, e1]
    u, e = np.concatenate(us), np.concatenate(es).astype(np.int64)
    assert set(np.unique(u)) == {-1, 0, 1}
    assert all(abs((u == v).mean() - 1 / 3) < 0.01 for v in (-1, 0, 1))
    assert abs((e == 0).mean() - 0.2453) < 0.005
    assert abs(e.var() - 8.0) < 0.15 and abs(e.mean()) < 0.03 and np.abs(e).max() <= 19
    assert len(bfv.NOISE_CDF) == 19 and (np.diff(bfv.NOISE_CDF.astype(np.float64)) > 0).all()


def test_every_seed_word_matters_and_streams_differ():
    seed = np.arange(1, 9, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    base = bfv.gpu_sampler(seed)
    assert not np.array_equal(base[1], base[2])
    for w in range(8):
        s2 = seed.copy()
        s2[w] ^= np.uint64(1)
        other = bfv.gpu_sampler(s2)
        assert all(not np.array_equal(a, b) for a, b in zip(base, other)), w


def test_encrypt_seeded_is_a_valid_fresh_encryption(keys):
    seed = np.frombuffer(bytes(range(64)), dtype="<u8")
    for kind, v in (("i64", -99), ("u256", 2**255 + 1), ("frac64", 3.5)):
        ct = bfv.encrypt_seeded(keys.net_pk, bfv.encode(kind, v), seed)
        assert decrypt_value(keys, kind, ct, network=True) == v
        assert bfv.decrypt(ct, keys.net_sk)[1] >= 52  # a fresh SEAL encryption has 53 bits of budget at these parameters
