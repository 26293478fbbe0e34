// Host-side parameter context: primes, roots, twiddle tables, BEHZ / key-switch constants and the
// per-GPU copies of them.  Replaces what SEALContext + RNSTool::initialize + NTTTables build inside
// `Runtime::new_fhe(&PARAMS)` (/root/reference/src/fhe.rs:116, testnet.rs:8-17).
#pragma once
#include <cstdint>
#include <vector>

#include "devconsts.h"

namespace fheb {

struct HostContext {
    u64 root[kNumMod];                   // minimal primitive 2N-th roots
    std::vector<ulonglong2> twf[kNumTab];  // forward twiddles (w, Shoup); dual limbs: both 32-bit lanes packed in each word
    std::vector<ulonglong2> twi[kNumTab];  // inverse twiddles
    DevConsts dc;
    uint64_t parms_id_key[4];   // BLAKE2b-256 over [scheme, N, q0, q1, P, t]
    uint64_t parms_id_data[4];  // BLAKE2b-256 over [scheme, N, q0, q1, t]
    // CRT helpers for decryption on the host
    u64 inv_q1_mod_q0, inv_q0_mod_q1;

    // Built once; throws std::runtime_error if the derived aux primes differ from params.h.
    static const HostContext &get();
};

struct DeviceContext {
    int device = -1;
    DevTables tabs{};
    void *table_mem = nullptr;
};

// Lazily creates the context of `device` (tables in HBM, constants in __constant__ memory, kernel
// attributes) and makes it the current device.  Throws std::runtime_error on any CUDA error.
DeviceContext &device_context(int device);
int device_count();

// host modular helpers (also used by the host-side encoders / decryptor)
u64 h_mulmod(u64 a, u64 b, u64 q);
u64 h_powmod(u64 b, u64 e, u64 q);
u64 h_invmod(u64 a, u64 q);  // q prime
bool is_prime_u64(u64 n);

// SHA-512 (FIPS 180-4), used for the deterministic-encryption seeds of fhe.rs:600-611, 646-649
void sha512(const void *in, size_t inlen, uint8_t out[64]);           // libcrypto's when present, else the portable one
void sha512_portable(const void *in, size_t inlen, uint8_t out[64]);  // own FIPS 180-4 implementation

// BLAKE2b with variable digest length (RFC 7693), used for SEAL parms_id
void blake2b(const void *in, size_t inlen, void *out, size_t outlen);

}  // namespace fheb
