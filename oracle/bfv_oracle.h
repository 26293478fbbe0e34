/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the BFV arithmetic that
 * the reference's precompiles execute.  Nothing in the product links or calls this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs load liboracle.so.
 *
 * The arithmetic of Sunscreen-tech/fhe_precompiles lives in an un-vendored
 * dependency: sunscreen 0.8.1 (reference Cargo.toml:16) -> seal_fhe -> Microsoft
 * SEAL 4.0 (C++).  Neither is on disk, so this is a restatement of SEAL 4.0's
 * published algorithms anchored on the reference's call sites:
 *   FheApp::run              /root/reference/src/fhe.rs:138-152
 *   36 #[fhe_program] bodies /root/reference/src/fhe.rs:814-1022 (a+b, a-b, a*b)
 *   PARAMS                   /root/reference/src/testnet.rs:8-14
 *   encrypt / decrypt        /root/reference/src/fhe.rs:594-618, 688-699
 *
 * Pinning status: PINNED to the reference's byte-level known answers -- the three
 * SHA-512 digests the reference asserts on the output of FheApp::encrypt / reencrypt
 * (fhe.rs:2101-2121, 2165-2185, 2224-2244) are reproduced by bfvo_seal_encrypt +
 * bfvo_decrypt + oracle/formats.py (tests/test_oracle_kat.py).  That fixes PRNG,
 * samplers, primes, roots, NTT order, key layout, encoders, scaling, decryption,
 * the serialisation and its libzstd level-3 compression against real SEAL output.
 * For add/sub/mul the reference holds no golden ciphertext; those
 * routines share all of the conventions above and are pinned in addition by:
 *   - the four key files: sk is ternary under this NTT convention, pk0+pk1*s and
 *     rk[j].c0+rk[j].c1*s-P*s^2[limb j] are small -> fixes primes, roots, NTT order,
 *     key layout and the key-switch digit convention (tests/test_oracle_fixtures.py);
 *   - the 48 decrypted-value tests of fhe.rs:1038-2076 (16 op 4 -> 20/12/64 for every
 *     type x shape), re-stated in tests/test_oracle_values.py;
 *   - an exact big-integer BFV multiply (independent Python) bounds BEHZ's error.
 */
#ifndef BFV_ORACLE_H
#define BFV_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFVO_N 4096
#define BFVO_LOGN 12
#define BFVO_T 4096ull
/* modulus indices */
enum { BFVO_Q0 = 0, BFVO_Q1 = 1, BFVO_P = 2, BFVO_B0 = 3, BFVO_B1 = 4, BFVO_MSK = 5, BFVO_NMOD = 6 };

/* Builds primes, roots, twiddles and BEHZ constants.  Idempotent, thread-safe after
 * the first call returns.  Returns 0 on success. */
int bfvo_init(void);

/* out[0..5] moduli, out[6..11] minimal 2N-th roots, out[12] gamma (decrypt aux prime),
 * out[13..] BEHZ constants in the order documented in bfv_oracle.c. Returns count. */
size_t bfvo_constants(uint64_t *out, size_t cap);

/* In-place negacyclic NTT of one limb (SEAL ntt_negacyclic_harvey convention:
 * natural order in, bit-reversed order out; canonical residues). */
void bfvo_ntt_fwd(uint64_t *a, int mod);
void bfvo_ntt_inv(uint64_t *a, int mod);

/* Data-level ciphertext layout everywhere: [poly][limb(q0,q1)][N] u64, coefficient form. */
void bfvo_add(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t npolys);
void bfvo_sub(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t npolys);
void bfvo_negate(const uint64_t *a, uint64_t *out, size_t npolys);
/* SEAL multiply_{add,sub}_plain_with_scaling_variant on c0; plain has plain_len coeffs < t */
void bfvo_add_plain(uint64_t *ct, const uint64_t *plain, size_t plain_len);
void bfvo_sub_plain(uint64_t *ct, const uint64_t *plain, size_t plain_len);
/* SEAL multiply_plain_normal on a size-`npolys` ciphertext */
void bfvo_multiply_plain(uint64_t *ct, size_t npolys, const uint64_t *plain, size_t plain_len);
/* SEAL bfv_multiply (BEHZ) size2 x size2 -> size3 */
void bfvo_multiply(const uint64_t *a, const uint64_t *b, uint64_t *out3);
/* SEAL relinearize (switch_key_inplace) size3 -> size2.
 * rk layout: [digit 0..1][poly 0..1][limb q0,q1,P][N], NTT form (as stored in the key file). */
void bfvo_relinearize(const uint64_t *ct3, const uint64_t *rk, uint64_t *out2);
void bfvo_mul_relin(const uint64_t *a, const uint64_t *b, const uint64_t *rk, uint64_t *out2);

/* Stage taps for per-kernel parity tests.  ext: [4 polys a0,a1,b0,b1][5 limbs q0,q1,b0,b1,msk][N]
 * = output of fastbconv_m_tilde + sm_mrq (coefficient form, canonical). */
void bfvo_behz_extend(const uint64_t *a, const uint64_t *b, uint64_t *ext);
/* tens: [3 polys][5 limbs][N] = INTT(tensor) * t, canonical, coefficient form */
void bfvo_behz_tensor(const uint64_t *ext, uint64_t *tens);
/* fast_floor + fastbconv_sk: tens -> out3 [3][2][N] */
void bfvo_behz_floor_sk(const uint64_t *tens, uint64_t *out3);

/* sunscreen plaintext encoders. Return plaintext length; plain must hold N. */
size_t bfvo_encode_i64(int64_t v, uint64_t *plain);
size_t bfvo_encode_u64(uint64_t v, uint64_t *plain);
size_t bfvo_encode_u256(const uint64_t limbs_le[4], uint64_t *plain);
/* returns 0 on failure (NaN/inf/out of range) */
size_t bfvo_encode_f64(double v, uint64_t *plain);
int64_t bfvo_decode_i64(const uint64_t *plain, size_t len);
uint64_t bfvo_decode_u64(const uint64_t *plain, size_t len);
void bfvo_decode_u256(const uint64_t *plain, size_t len, uint64_t limbs_le[4]);
double bfvo_decode_f64(const uint64_t *plain, size_t len);

/* Public-key encryption in stock SEAL's shape (key level + modulus switching) with a throw-away PRNG: makes the
 * operands of the add/sub/mul tests, as the reference's randomised runtime.encrypt does.
 * pk: [2][3][N] NTT form at key level. ct_out: [2][2][N]. */
void bfvo_encrypt(const uint64_t *pk, const uint64_t *plain, size_t plain_len, uint64_t seed, uint64_t *ct_out);
/* the same computation with caller-supplied samples (u ternary, e0 / e1 errors, N int8 each) */
void bfvo_encrypt_samples(const uint64_t *pk, const uint64_t *plain, size_t plain_len, const int8_t *u, const int8_t *e0,
                          const int8_t *e1, uint64_t *ct_out);
/* SEAL-exact deterministic encryption (what FheApp::encrypt / reencrypt produce; pinned by the reference's SHA-512
 * known answers, see bfv_oracle.c).  seed = the SHA-512 digest as 8 little-endian words.
 * bfvo_seal_prng: the first nbytes of Blake2xbPRNG(seed)'s stream.  bfvo_seal_sample: (u, e0, e1), returns the number of
 * 32-bit draws consumed.  bfvo_seal_encrypt: pk [2][3][N] NTT form -> ct [2][2][N] (no special modulus). */
void bfvo_seal_prng(const uint64_t seed[8], uint8_t *out, size_t nbytes);
size_t bfvo_seal_sample(const uint64_t seed[8], int8_t *u, int8_t *e0, int8_t *e1);
size_t bfvo_seal_sample_stream(const uint32_t *words, size_t nwords, int8_t *u, int8_t *e0, int8_t *e1);
void bfvo_encrypt_samples_data_level(const uint64_t *pk, const uint64_t *plain, size_t plain_len, const int8_t *u, const int8_t *e0,
                                     const int8_t *e1, uint64_t *ct_out);
void bfvo_seal_encrypt(const uint64_t *pk, const uint64_t *plain, size_t plain_len, const uint64_t seed[8], uint64_t *ct_out);
/* Decrypt size-2 or size-3 data-level ciphertext with sk [3][N] (NTT form).
 * plain_out gets N coefficients < t.  Returns the invariant noise budget in bits (<=0: failed). */
int bfvo_decrypt(const uint64_t *ct, size_t npolys, const uint64_t *sk, uint64_t *plain_out);

/* Batched helpers for the CPU baseline: n independent ops, `threads` host threads. Seconds. */
double bfvo_batch_mul_relin(const uint64_t *a, const uint64_t *b, const uint64_t *rk, uint64_t *out, size_t n, int threads);
double bfvo_batch_ntt(uint64_t *limbs, size_t n_limbs, int mod, int inverse, int threads);

#ifdef __cplusplus
}
#endif
#endif
