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
    prev = lib.fhe_b200_set_zstd_writer(0)
    try:
        out, n = ctypes.c_void_p(), ctypes.c_int64()
        assert lib.fhe_b200_write_ciphertext(words.ctypes.data, dt.value, ctypes.byref(out), ctypes.byref(n)) == 0
        assert ctypes.string_at(out.value, n.value) == buf  # libzstd writer: byte-identical re-serialisation (level 3)
        lib.fhe_free(out)
    finally:
        lib.fhe_b200_set_zstd_writer(prev)


def _frame_of(ct_bytes: bytes) -> bytes:
    """the zstd frame inside a serialised sunscreen Ciphertext"""
    r = F.Reader(ct_bytes)
    r.take(r.u64())
    r.u32(), r.u64()
    blob = F.WithContext.read(r).blob
    assert blob[5] == F.COMPR_ZSTD
    return blob[16:]


@pytest.fixture
def structured_writer(lib):
    """switch the library to its structured-frame writer for one test (the default is libzstd level 3 = SEAL's bytes)"""
    prev = lib.fhe_b200_set_zstd_writer(1)
    assert prev == 0, "libzstd level 3 (SEAL's bytes) is the default writer"
    yield
    lib.fhe_b200_set_zstd_writer(prev)


def test_structured_zstd_frames(lib, keys, structured_writer):
    """The structured writer lays ciphertext payloads out as RFC 8878 frames by hand (codec.cpp zstd_pack40). The bytes must
    equal the format oracle's independent restatement, decode to the same payload with libzstd (one-shot and streaming,
    which is what SEAL uses) and with the oracle's RFC-only mini decoder, and both of the library's readers must agree."""
    rng = np.random.default_rng(77)
    dt = F.data_type_string("i64").encode()
    words = np.zeros(4 * N, dtype=np.uint64)
    name = ctypes.create_string_buffer(256)

    def write(polys):
        w = np.ascontiguousarray(polys, dtype=np.uint64).reshape(-1)
        out, n = ctypes.c_void_p(), ctypes.c_int64()
        assert lib.fhe_b200_write_ciphertext(w.ctypes.data, dt, ctypes.byref(out), ctypes.byref(n)) == 0
        b = ctypes.string_at(out.value, n.value)
        lib.fhe_free(out)
        return b

    q = bfv.moduli()[:2]
    cases = [np.stack([[rng.integers(0, q[l], N, dtype=np.uint64) for l in range(2)] for _ in range(2)]) for _ in range(3)]
    cases.append(np.stack([[np.full(N, q[l] - 1, dtype=np.uint64) for l in range(2)] for _ in range(2)]))  # constant: libzstd
    cases.append(np.zeros((2, 2, N), dtype=np.uint64))  # transparent: libzstd path
    edge = cases[0].copy()
    edge[0, 0, :2] = 0  # block A's two words
    edge[1, 1, -1] = q[1] - 1
    cases.append(edge)
    sizes = []
    for polys in cases:
        got = write(polys)
        want = F.make_ciphertext("i64", polys)
        assert got == want.to_bytes(structured=True)
        frame, payload = _frame_of(got), want.parts[0][1].payload()
        assert F.zstd().decompress(frame) == payload
        assert F.zstd().decompress_stream(frame, in_chunk=777, out_chunk=1000) == payload
        structured = frame[4] == 0xA0 and len(frame) > 5 * 4 * N
        if structured:
            assert F.zstd_mini_decode(frame) == payload
            assert len(frame) == 9 + (3 + 2 + 110 + 7) + (3 + 3 + 5 * (4 * N - 2) + 2 + 5)
        sizes.append((len(frame), len(F.zstd().compress(payload, 3))))
        for blob in (got, want.to_bytes()):  # the recogniser and the libzstd reader
            assert lib.fhe_b200_parse_ciphertext(blob, len(blob), words.ctypes.data, name, 256) == 0
            assert np.array_equal(words.reshape(2, 2, N), polys)
    assert sizes[0][0] < sizes[0][1], "structured frames are smaller than libzstd level 3 on real residues"
    assert sizes[4][0] < 1000, "constant data still goes to libzstd"


def test_structured_frame_reader_is_strict(lib, keys):
    """Every single-byte edit of the header bytes of a structured frame must either be rejected or decode (through the
    libzstd fallback) to exactly what libzstd itself says; literal edits change only the word they belong to."""
    rng = np.random.default_rng(78)
    q = bfv.moduli()[:2]
    polys = np.stack([[rng.integers(0, q[l], N, dtype=np.uint64) for l in range(2)] for _ in range(2)])
    good = F.make_ciphertext("i64", polys).to_bytes(structured=True)
    frame = _frame_of(good)
    start = good.index(frame)
    words = np.zeros(4 * N, dtype=np.uint64)
    name = ctypes.create_string_buffer(256)
    header_positions = list(range(0, 9 + 3 + 2)) + list(range(9 + 5 + 110, 9 + 5 + 110 + 7 + 6)) + list(range(len(frame) - 8, len(frame)))
    for pos in header_positions:
        for flip in (0x01, 0x80):
            bad = bytearray(good)
            bad[start + pos] ^= flip
            bad = bytes(bad)
            rc = lib.fhe_b200_parse_ciphertext(bad, len(bad), words.ctypes.data, name, 256)
            try:
                ref = F.Ciphertext.from_bytes(bad).polys()
            except Exception:  # noqa: BLE001
                ref = None
            if ref is None or (ref >= np.array(q, dtype=np.uint64)[None, :, None]).any():
                assert rc == 3, (pos, flip, rc)
            else:
                assert rc == 0 and np.array_equal(words.reshape(2, 2, N), ref), (pos, flip)
    # a literal byte belongs to exactly one coefficient
    bad = bytearray(good)
    bad[start + 9 + 122 + 6 + 5 * 100] ^= 0x04
    assert lib.fhe_b200_parse_ciphertext(bytes(bad), len(bad), words.ctypes.data, name, 256) == 0
    assert np.array_equal(words.reshape(2, 2, N), F.Ciphertext.from_bytes(bytes(bad)).polys())
    assert (words.reshape(2, 2, N) != polys).sum() == 1


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
    # ... and zlib ones (compr_mode 1)
    assert parse(ct.to_bytes(compr=F.COMPR_ZLIB)) == 0
    assert np.array_equal(words.reshape(2, 2, N), ct.polys())
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
