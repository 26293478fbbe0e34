"""Wire formats: fixtures are fully consumed by the grammar, re-serialise byte-for-byte, and the C-ABI library's
host codec agrees with the Python format oracle.  CPU only (no compute calls)."""
import ctypes
import os
import struct

import numpy as np
import pytest

from helpers import N, ROOT, encrypt_value
from oracle import bfv
from oracle import formats as F


@pytest.mark.parametrize("path", ["tests/data/public_key.bin", "fhe_precompiles_b200/data/network.pub"])
def test_public_key_fixture_roundtrip(path):
    buf = open(os.path.join(ROOT, path), "rb").read()
    pk = F.PublicKey.from_bytes(buf)
    assert pk.galois_key is None and pk.relin_key is not None
    assert pk.to_bytes() == buf
    payload, compr = F.seal_unwrap(pk.public_key.blob)
    assert compr == F.COMPR_ZSTD and len(payload) == 196705
    ct = F.SealCiphertext.from_payload(payload)
    assert (ct.is_ntt_form, ct.size, ct.coeff_modulus_size, ct.scale, ct.correction_factor) == (1, 2, 3, 1.0, 1)
    assert ct.payload() == payload
    # zstd level 3 of the system library reproduces the public-key blob byte for byte
    assert F.seal_wrap(payload) == pk.public_key.blob
    rk_payload, _ = F.seal_unwrap(pk.relin_key.blob)
    rk = F.RelinKeys.from_payload(rk_payload)
    assert len(rk.keys) == 1 and len(rk.keys[0]) == 2 and rk.payload() == rk_payload


@pytest.mark.parametrize("path", ["tests/data/private_key.bin", "fhe_precompiles_b200/data/network.pri"])
def test_private_key_fixture(path):
    sk = F.read_private_key(open(os.path.join(ROOT, path), "rb").read())
    assert sk.data.shape == (3, N) and sk.parms_id == F.PARMS_ID_KEY


def test_pack_framing_roundtrip_and_errors():
    from fhe_precompiles_b200 import pack

    a, b, k = b"\x01" * 5, b"\x02" * 9, b"\x03" * 17
    packed = pack.pack_binary_operation(k, a, b)
    assert packed == F.pack_binary_operation(k, a, b)
    assert packed[:8] == struct.pack(">II", 8 + 17, 8 + 17 + 5)
    assert pack.unpack_binary_operation(packed) == (k, a, b)
    assert pack.pack_binary_operation(*pack.unpack_binary_operation(packed)) == packed  # pack.rs round-trip tests
    two = pack.pack_two_arguments(a, b)
    assert two[:4] == struct.pack(">I", 9) and pack.unpack_two_arguments(two) == (a, b)
    assert pack.pack_two_arguments(b"", b"") == struct.pack(">I", 4)  # empty operands
    for bad in (b"", b"\x00\x00\x00"):
        with pytest.raises(pack.FheError) as e:
            pack.unpack_two_arguments(bad)
        assert e.value.code == 1
    with pytest.raises(pack.FheError) as e:
        pack.unpack_binary_operation(b"\x00" * 7)
    assert e.value.code == 1
    assert pack.deserialize_scalar("i64", pack.serialize_i64(-5)) == -5
    assert pack.deserialize_scalar("u256", pack.serialize_u256(2**255 + 1)) == 2**255 + 1
    assert pack.deserialize_scalar("frac64", pack.serialize_frac64(-0.5)) == -0.5
    with pytest.raises(pack.FheError) as e:
        pack.deserialize_scalar("u64", b"\x00" * 7)
    assert e.value.code == 3


def test_golden_ciphertext_fixture_parses(keys):
    buf = open(os.path.join(ROOT, "tests/golden/ct_i64_16_seed11.bin"), "rb").read()
    ct = F.Ciphertext.from_bytes(buf)
    assert ct.data_type == "sunscreen::types::bfv::signed::Signed,0.8.1,true"
    assert ct.parts[0][1].parms_id == F.PARMS_ID_DATA
    assert np.array_equal(ct.polys(), encrypt_value(keys, "i64", 16, 11))
    assert ct.to_bytes() == buf
    assert bfv.decode("i64", bfv.decrypt(ct.polys(), keys.sk)[0]) == 16


# ---------------------------------------------------------------- the library's host codec (no GPU needed)
@pytest.fixture(scope="module")
def lib():
    from fhe_precompiles_b200 import _lib

    return _lib.lib()


def test_library_codec_agrees_with_format_oracle(lib, keys):
    pk = np.zeros(2 * 3 * N, dtype=np.uint64)
    rk = np.zeros(2 * 2 * 3 * N, dtype=np.uint64)
    assert lib.fhe_b200_parse_public_key(keys.pub_bytes, len(keys.pub_bytes), pk.ctypes.data, rk.ctypes.data) == 0
    assert np.array_equal(pk.reshape(2, 3, N), keys.pk) and np.array_equal(rk.reshape(2, 2, 3, N), keys.rk)
    sk = np.zeros(3 * N, dtype=np.uint64)
    assert lib.fhe_b200_parse_private_key(keys.pri_bytes, len(keys.pri_bytes), sk.ctypes.data) == 0
    assert np.array_equal(sk.reshape(3, N), keys.sk)
    pid = (ctypes.c_uint64 * 4)()
    lib.fhe_b200_parms_id(0, pid)
    assert tuple(pid) == F.PARMS_ID_KEY
    lib.fhe_b200_parms_id(1, pid)
    assert tuple(pid) == F.PARMS_ID_DATA

    buf = open(os.path.join(ROOT, "tests/golden/ct_i64_mul_16_4.bin"), "rb").read()
    words = np.zeros(4 * N, dtype=np.uint64)
    dt = ctypes.create_string_buffer(256)
    assert lib.fhe_b200_parse_ciphertext(buf, len(buf), words.ctypes.data, dt, 256) == 0
    assert np.array_equal(words.reshape(2, 2, N), F.Ciphertext.from_bytes(buf).polys())
    out, n = ctypes.c_void_p(), ctypes.c_int64()
    assert lib.fhe_b200_write_ciphertext(words.ctypes.data, dt.value, ctypes.byref(out), ctypes.byref(n)) == 0
    assert ctypes.string_at(out.value, n.value) == buf  # byte-identical re-serialisation (zstd level 3)
    lib.fhe_free(out)


def test_library_codec_rejects_malformed_inputs(lib, keys):
    good = open(os.path.join(ROOT, "tests/golden/ct_i64_16_seed11.bin"), "rb").read()
    words = np.zeros(4 * N, dtype=np.uint64)
    dt = ctypes.create_string_buffer(256)
    parse = lambda b: lib.fhe_b200_parse_ciphertext(b, len(b), words.ctypes.data, dt, 256)
    assert parse(good) == 0
    assert parse(b"") == 3 and parse(good[:-1]) == 3 and parse(good + b"\x00") == 3
    assert parse(good[:60] + b"\xff" + good[61:]) in (3, 7)  # corrupt Params
    corrupt = bytearray(good)
    corrupt[-20] ^= 0xFF  # inside the zstd frame
    assert parse(bytes(corrupt)) == 3
    # coefficient >= q must be rejected like SEAL's checked load
    ct = F.Ciphertext.from_bytes(good)
    ct.parts[0][1].data[5] = 0xFFFFEE001
    assert parse(ct.to_bytes()) == 3
    # uncompressed SEAL blobs (compr_mode none) are accepted too
    ct = F.Ciphertext.from_bytes(good)
    assert parse(ct.to_bytes(compr=F.COMPR_NONE)) == 0
    # key-level parms_id on a ciphertext is invalid
    ct.parts[0][1].parms_id = F.PARMS_ID_KEY
    assert parse(ct.to_bytes()) == 3
    pk = np.zeros(2 * 3 * N, dtype=np.uint64)
    assert lib.fhe_b200_parse_public_key(keys.pub_bytes[:-3], len(keys.pub_bytes) - 3, pk.ctypes.data, None) == 3
    assert lib.fhe_b200_parse_public_key(keys.pri_bytes, len(keys.pri_bytes), pk.ctypes.data, None) == 3


def test_codec_survives_mutated_inputs(lib, keys):
    """boundary hardening: random truncations, bit flips and length-field edits of valid ciphertext / key bytes must come
    back as error codes (or parse cleanly), never crash or read out of bounds (the reference panics on some of these)."""
    rng = np.random.default_rng(1234)
    good_ct = open(os.path.join(ROOT, "tests/golden/ct_i64_16_seed11.bin"), "rb").read()
    words = np.zeros(4 * N, dtype=np.uint64)
    pk_words = np.zeros(2 * 3 * N, dtype=np.uint64)
    rk_words = np.zeros(2 * 2 * 3 * N, dtype=np.uint64)
    dt = ctypes.create_string_buffer(64)  # deliberately small: long type strings must be truncated, not overflow

    def mutate(buf: bytes) -> bytes:
        b = bytearray(buf)
        kind = rng.integers(0, 5)
        if kind == 0:
            return bytes(b[: rng.integers(0, len(b))])
        if kind == 1:
            for _ in range(rng.integers(1, 8)):
                b[rng.integers(0, len(b))] ^= 1 << rng.integers(0, 8)
            return bytes(b)
        if kind == 2:  # clobber one of the early length / count fields
            off = int(rng.integers(0, 200))
            b[off : off + 8] = int(rng.integers(0, 2**63)).to_bytes(8, "little")
            return bytes(b)
        if kind == 3:
            return bytes(b) + bytes(rng.integers(0, 256, size=rng.integers(1, 64), dtype=np.uint8))
        i = rng.integers(0, len(b) - 16)
        b[i : i + 16] = bytes(rng.integers(0, 256, size=16, dtype=np.uint8))
        return bytes(b)

    codes = set()
    for _ in range(150):
        m = mutate(good_ct)
        codes.add(lib.fhe_b200_parse_ciphertext(m, len(m), words.ctypes.data, dt, 64))
    for _ in range(40):
        m = mutate(keys.pub_bytes)
        codes.add(lib.fhe_b200_parse_public_key(m, len(m), pk_words.ctypes.data, rk_words.ctypes.data))
        m = mutate(keys.pri_bytes)
        codes.add(lib.fhe_b200_parse_private_key(m, len(m), pk_words.ctypes.data))
    assert codes <= {0, 3, 7}, codes
    assert 3 in codes
