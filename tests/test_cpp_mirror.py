"""include/fhe_precompiles_b200.hpp -- the C++ mirror of the reference crate's typed surface (FheApp, pack, FheError,
testnet::one::FHE; the reference is compiled Rust and this image has no Rust toolchain) -- driven by tests/fheapp_test.cpp, a
restatement of the reference's own tests (fhe.rs:1038-2303) against that mirror.  CPU: it compiles warning-free, links against
the shared library and passes its host-only part.  GPU: the 36 precompiles give 20 / 12 / 64 for every type and argument
shape, the threshold API round-trips, and the reference's two API-reachable SHA-512 known answers (fhe_encrypt_test,
fhe_reencrypt_test) are reproduced -- all through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "fheapp_test"
    lib_dir = os.path.join(ROOT, "fhe_precompiles_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "fheapp_test.cpp"), "-L" + lib_dir, "-lfhe_precompiles_b200",
                    "-Wl,-rpath," + lib_dir, "-o", str(out)], check=True)
    return str(out)


def test_cpp_mirror_builds_and_host_part_passes(exe):
    r = subprocess.run([exe, "--cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok cpu"), r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_tests_through_the_cpp_mirror(exe):
    env = dict(os.environ, FHE_B200_ZSTD_WRITER="lib")  # the known answers hash SEAL's bytes (the default writer)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "data")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.startswith("ok gpu"), r.stdout + r.stderr
