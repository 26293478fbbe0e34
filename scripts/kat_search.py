"""Searches the unpinned degrees of freedom of the reference's deterministic encryption for a combination that
reproduces its SHA-512 known answer (fhe_encrypt_test, /root/reference/src/fhe.rs:2083-2121, non-macOS branch).
Pure CPU (oracle only). A hit would pin data_type strings, encoders, PRNG, samplers, NTT roots, modulus switching,
serialisation and zstd against real SEAL output."""
import hashlib
import itertools
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bfv, formats as F, seal_encrypt as S  # noqa: E402

KAT_ENCRYPT = bytes([190, 214, 153, 167, 205, 130, 61, 102, 188, 80, 220, 159, 38, 110, 126, 216, 148, 46, 220, 80, 18, 189, 177, 187,
                     108, 99, 32, 72, 250, 225, 2, 166, 33, 155, 22, 86, 221, 82, 4, 174, 144, 196, 45, 28, 190, 100, 194, 192, 37, 81,
                     203, 227, 46, 179, 59, 153, 20, 118, 191, 69, 244, 113, 180, 123])
SECRET = bytes([15, 17, 225, 5, 30, 1, 237, 218, 130, 19, 37, 95, 222, 218, 244, 172, 214, 175, 175, 110, 173, 103, 172, 60, 43, 76, 40, 150,
                215, 96, 23, 78, 22, 39, 30, 177, 107, 130, 124, 109, 27, 96, 206, 125, 104, 241, 10, 40, 88, 238, 117, 118, 79, 113, 213, 110,
                148, 179, 53, 19, 227, 154, 151, 122])


def main():
    pub = open(os.path.join(ROOT, "fhe_precompiles_b200/data/network.pub"), "rb").read()
    pri = open(os.path.join(ROOT, "fhe_precompiles_b200/data/network.pri"), "rb").read()
    pk = F.PublicKey.from_bytes(pub).pk_polys()
    sk = F.read_private_key(pri).data
    value_bytes = (12).to_bytes(32, "big")
    public_data = bytes([1, 2, 3])
    seed = struct.unpack("<8Q", hashlib.sha512(public_data + SECRET + value_bytes).digest())
    plain = bfv.encode("u256", 12)
    cts = {}
    for uname, ufn in (("lemire", S.uniform3_lemire), ("downscale", S.uniform3_downscale)):
        for nname, nfn in (("normal", S.sample_clipped_normal), ("cbd", S.sample_cbd)):
            ct = S.encrypt_seeded(pk, plain, list(seed), ufn, nfn)
            p, budget = bfv.decrypt(ct, sk)
            print(uname, nname, "decrypts to", bfv.decode("u256", p), "budget", budget)
            cts[uname + "+" + nname] = ct
    bases = ["sunscreen::types::bfv::unsigned::Unsigned", "sunscreen::types::bfv::Unsigned", "sunscreen::types::Unsigned",
             "sunscreen_runtime::types::bfv::unsigned::Unsigned", "Unsigned"]
    suffixes = ["", "<4>", "256", "<256>", "4", "<4usize>", "<4_usize>"]
    versions = ["0.8.1", "0.8.0", "0.8.2"]
    params = F.Params()
    tried = 0
    for uname, ct in cts.items():
        sealct = F.fresh_data_ciphertext(ct)
        for compr in (F.COMPR_ZSTD, F.COMPR_NONE):
            blob = F.seal_wrap(sealct.payload(), compr)
            wc = F.WithContext(params, blob).to_bytes()
            for base, suf, ver in itertools.product(bases, suffixes, versions):
                name = base + suf
                layouts = {
                    "string": struct.pack("<Q", len(f"{name},{ver},true")) + f"{name},{ver},true".encode(),
                    "struct": struct.pack("<Q", len(name)) + name.encode() + struct.pack("<Q", len(ver)) + ver.encode() + b"\x01",
                    "struct_semver": struct.pack("<Q", len(name)) + name.encode() + struct.pack("<QQQ", *map(int, ver.split("."))) + b"\x01",
                }
                for lname, head in layouts.items():
                    for tag in (struct.pack("<I", 0), b""):
                        body = head + tag + struct.pack("<Q", 1) + wc
                        tried += 1
                        if hashlib.sha512(body).digest() == KAT_ENCRYPT:
                            print("MATCH:", uname, compr, name, ver, lname, len(tag))
                            return
    print("no match in", tried, "combinations")


if __name__ == "__main__":
    main()
