// See context.h.  Derivation rules restated from SEAL 4.0 (un-vendored dependency of the reference,
// Cargo.toml:16): util::get_primes, util::try_minimal_primitive_root, NTTTables::initialize,
// SEALContext::validate (coeff_div_plain_modulus, upper-half constants) and RNSTool::initialize.
#include "context.h"

#include <cmath>

#include <dlfcn.h>

#include <cstdlib>
#include <initializer_list>

#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>

#include "kernels.h"

namespace fheb {

typedef unsigned __int128 u128;

u64 h_mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
u64 h_powmod(u64 b, u64 e, u64 q) {
    u64 r = 1 % q;
    b %= q;
    for (; e; e >>= 1) {
        if (e & 1) r = h_mulmod(r, b, q);
        b = h_mulmod(b, b, q);
    }
    return r;
}
u64 h_invmod(u64 a, u64 q) { return h_powmod(a % q, q - 2, q); }
static u64 shoup_of(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
static Shoup mk_shoup(u64 w, u64 q) { return Shoup{w, shoup_of(w, q)}; }

// ---------------------------------------------------------------- BLAKE2b (RFC 7693)
namespace {
const uint64_t kIV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                         0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
const uint8_t kSigma[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
void compress(uint64_t h[8], const uint8_t block[128], u128 t, bool last) {
    uint64_t m[16], v[16];
    memcpy(m, block, 128);
    for (int i = 0; i < 8; i++) {
        v[i] = h[i];
        v[i + 8] = kIV[i];
    }
    v[12] ^= (uint64_t)t;
    v[13] ^= (uint64_t)(t >> 64);
    if (last) v[14] = ~v[14];
    for (int r = 0; r < 12; r++) {
        const uint8_t *s = kSigma[r];
#define G(a, b, c, d, x, y)                 \
    v[a] = v[a] + v[b] + x;                 \
    v[d] = rotr(v[d] ^ v[a], 32);           \
    v[c] = v[c] + v[d];                     \
    v[b] = rotr(v[b] ^ v[c], 24);           \
    v[a] = v[a] + v[b] + y;                 \
    v[d] = rotr(v[d] ^ v[a], 16);           \
    v[c] = v[c] + v[d];                     \
    v[b] = rotr(v[b] ^ v[c], 63);
        G(0, 4, 8, 12, m[s[0]], m[s[1]]);
        G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        G(2, 6, 10, 14, m[s[4]], m[s[5]]);
        G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        G(0, 5, 10, 15, m[s[8]], m[s[9]]);
        G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        G(2, 7, 8, 13, m[s[12]], m[s[13]]);
        G(3, 4, 9, 14, m[s[14]], m[s[15]]);
#undef G
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
}  // namespace

void blake2b(const void *in, size_t inlen, void *out, size_t outlen) {
    uint64_t h[8];
    for (int i = 0; i < 8; i++) h[i] = kIV[i];
    h[0] ^= 0x01010000ULL ^ (uint64_t)outlen;
    const uint8_t *p = (const uint8_t *)in;
    u128 t = 0;
    while (inlen > 128) {
        t += 128;
        compress(h, p, t, false);
        p += 128;
        inlen -= 128;
    }
    uint8_t block[128] = {0};
    memcpy(block, p, inlen);
    t += inlen;
    compress(h, block, t, true);
    memcpy(out, h, outlen);
}

// ---------------------------------------------------------------- SHA-512 (FIPS 180-4)
namespace {
const uint64_t kSha512K[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL,
    0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL, 0xd807aa98a3030242ULL, 0x12835b0145706fbeULL,
    0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL, 0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL,
    0xc19bf174cf692694ULL, 0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL, 0x983e5152ee66dfabULL,
    0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL,
    0x06ca6351e003826fULL, 0x142929670a0e6e70ULL, 0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL,
    0x53380d139d95b3dfULL, 0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL,
    0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL, 0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL,
    0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL,
    0x682e6ff3d6b2b8a3ULL, 0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL, 0xca273eceea26619cULL,
    0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL,
    0x113f9804bef90daeULL, 0x1b710b35131c471bULL, 0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL,
    0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL};
inline uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
// FIPS 180-4 compression function; rounds unrolled by eight so that the working variables never move
#define SHA512_RND(a, b, c, d, e, f, g, h, i)                                                                   \
    {                                                                                                           \
        const uint64_t t1 = h + (ror64(e, 14) ^ ror64(e, 18) ^ ror64(e, 41)) + (g ^ (e & (f ^ g))) + kSha512K[i] + w[i]; \
        const uint64_t t2 = (ror64(a, 28) ^ ror64(a, 34) ^ ror64(a, 39)) + (((a | b) & c) | (a & b));           \
        d += t1;                                                                                                \
        h = t1 + t2;                                                                                            \
    }
void sha512_block(uint64_t h[8], const uint8_t *p) {
    uint64_t w[80];
    for (int i = 0; i < 16; i++) {
        uint64_t v;
        memcpy(&v, p + 8 * i, 8);
        w[i] = __builtin_bswap64(v);
    }
    for (int i = 16; i < 80; i++) {
        const uint64_t s0 = ror64(w[i - 15], 1) ^ ror64(w[i - 15], 8) ^ (w[i - 15] >> 7);
        const uint64_t s1 = ror64(w[i - 2], 19) ^ ror64(w[i - 2], 61) ^ (w[i - 2] >> 6);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 80; i += 8) {
        SHA512_RND(a, b, c, d, e, f, g, hh, i)
        SHA512_RND(hh, a, b, c, d, e, f, g, i + 1)
        SHA512_RND(g, hh, a, b, c, d, e, f, i + 2)
        SHA512_RND(f, g, hh, a, b, c, d, e, i + 3)
        SHA512_RND(e, f, g, hh, a, b, c, d, i + 4)
        SHA512_RND(d, e, f, g, hh, a, b, c, i + 5)
        SHA512_RND(c, d, e, f, g, hh, a, b, i + 6)
        SHA512_RND(b, c, d, e, f, g, hh, a, i + 7)
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
#undef SHA512_RND

// OpenSSL's SHA512 (assembly, AVX2) when libcrypto is on the machine: reencrypt hashes ~500 KB per call (fhe.rs:632-678)
typedef unsigned char *(*Sha512Fn)(const unsigned char *, size_t, unsigned char *);
Sha512Fn libcrypto_sha512() {
    static Sha512Fn fn = [] {
        if (getenv("FHE_B200_NO_LIBCRYPTO")) return (Sha512Fn) nullptr;
        for (const char *name : {"libcrypto.so.3", "libcrypto.so.1.1", "libcrypto.so"}) {
            if (void *h = dlopen(name, RTLD_NOW | RTLD_LOCAL)) {
                if (void *sym = dlsym(h, "SHA512")) return (Sha512Fn)sym;
            }
        }
        return (Sha512Fn) nullptr;
    }();
    return fn;
}
}  // namespace

void sha512_portable(const void *in, size_t inlen, uint8_t out[64]) {
    uint64_t h[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                     0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    const uint8_t *p = (const uint8_t *)in;
    size_t n = inlen;
    while (n >= 128) {
        sha512_block(h, p);
        p += 128;
        n -= 128;
    }
    uint8_t tail[256] = {0};
    memcpy(tail, p, n);
    tail[n] = 0x80;
    size_t tl = (n + 17 <= 128) ? 128 : 256;
    unsigned __int128 bits = (unsigned __int128)inlen * 8;
    for (int k = 0; k < 16; k++) tail[tl - 1 - k] = (uint8_t)(bits >> (8 * k));
    sha512_block(h, tail);
    if (tl == 256) sha512_block(h, tail + 128);
    for (int i = 0; i < 8; i++)
        for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(h[i] >> (56 - 8 * k));
}

void sha512(const void *in, size_t inlen, uint8_t out[64]) {
    if (Sha512Fn f = libcrypto_sha512()) {
        f((const unsigned char *)in, inlen, out);
        return;
    }
    sha512_portable(in, inlen, out);
}

// ---------------------------------------------------------------- number theory
bool is_prime_u64(u64 n) {
    static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return false;
    for (u64 b : bases)
        if (n % b == 0) return n == b;
    u64 d = n - 1;
    int s = 0;
    while (!(d & 1)) d >>= 1, s++;
    for (u64 b : bases) {
        u64 x = h_powmod(b, d, n);
        if (x == 1 || x == n - 1) continue;
        bool composite = true;
        for (int r = 1; r < s && composite; r++) {
            x = h_mulmod(x, x, n);
            if (x == n - 1) composite = false;
        }
        if (composite) return false;
    }
    return true;
}
namespace {
// SEAL util::get_primes: descending primes congruent to 1 mod factor with exactly `bits` bits
std::vector<u64> seal_get_primes(u64 factor, int bits, size_t count) {
    std::vector<u64> out;
    u64 v = (((u64)1 << bits) - 1) / factor * factor + 1;
    const u64 lower = (u64)1 << (bits - 1);
    for (; out.size() < count && v > lower; v -= factor)
        if (is_prime_u64(v)) out.push_back(v);
    if (out.size() != count) throw std::runtime_error("fhe_b200: not enough auxiliary primes");
    return out;
}
// SEAL util::try_minimal_primitive_root: smallest primitive degree-th root of unity (degree = 2N)
u64 minimal_root(u64 degree, u64 q) {
    const u64 cof = (q - 1) / degree;
    u64 root = 0;
    for (u64 g = 2; g < q && !root; g++) {
        u64 c = h_powmod(g, cof, q);
        if (h_powmod(c, degree / 2, q) == q - 1) root = c;
    }
    const u64 sq = h_mulmod(root, root, q);
    u64 best = root, cur = root;
    for (u64 i = 0; i < degree; i += 2) {
        if (cur < best) best = cur;
        cur = h_mulmod(cur, sq, q);
    }
    return best;
}
u32 bitrev12(u32 x) {
    u32 r = 0;
    for (int i = 0; i < kLogN; i++) r = (r << 1) | ((x >> i) & 1);
    return r;
}

void build(HostContext &H) {
    // --- aux primes by SEAL's rule must equal the compile-time constants the kernels use
    std::vector<u64> aux = seal_get_primes(2 * kN, 61, 4);
    if (aux[0] != kModulus[MSK] || aux[1] != kGamma || aux[2] != kModulus[MB0] || aux[3] != kModulus[MB1])
        throw std::runtime_error("fhe_b200: BEHZ auxiliary primes differ from compiled constants");
    for (int i = 0; i < kNumMod; i++)
        if (!is_prime_u64(kModulus[i]) || kModulus[i] % (2 * kN) != 1)
            throw std::runtime_error("fhe_b200: modulus is not an NTT prime");

    // --- twiddles
    for (int mi = 0; mi < kNumMod; mi++) {
        const u64 q = kModulus[mi];
        const u64 psi = minimal_root(2 * kN, q);
        H.root[mi] = psi;
        H.twf[mi].resize(kN);
        H.twi[mi].resize(kN);
        u64 pw = 1;
        for (u32 i = 0; i < (u32)kN; i++) {
            u32 k = bitrev12(i);
            H.twf[mi][k] = make_ulonglong2(pw, shoup_of(pw, q));
            pw = h_mulmod(pw, psi, q);
        }
        for (u32 k = 0; k < (u32)kN; k++) {
            u64 inv = h_invmod(H.twf[mi][k].x, q);
            H.twi[mi][k] = make_ulonglong2(inv, shoup_of(inv, q));
        }
    }

    // --- the dual base: primes, packed twiddles (32-bit Shoup quotients floor(w 2^32 / s))
    {
        u64 v = (((u64)1 << 30) - 1) / (2 * kN) * (2 * kN) + 1;
        for (int i = 0; i < 2 * kNumDual; v -= 2 * kN)
            if (is_prime_u64(v)) {
                if (v != kDualPrime[i]) throw std::runtime_error("fhe_b200: dual-base primes differ from compiled constants");
                i++;
            }
    }
    for (int d = 0; d < kNumDual; d++) {
        const int ti = kNumMod + d;
        H.twf[ti].assign(kN, make_ulonglong2(0, 0));
        H.twi[ti].assign(kN, make_ulonglong2(0, 0));
        for (int lane = 0; lane < 2; lane++) {
            const u64 s = kDualPrime[2 * d + lane];
            const u64 psi = minimal_root(2 * kN, s);
            const int sh = 32 * lane;
            u64 pw = 1;
            for (u32 i = 0; i < (u32)kN; i++) {
                const u32 k = bitrev12(i);
                const u64 inv = h_invmod(pw, s);
                H.twf[ti][k].x |= pw << sh;
                H.twf[ti][k].y |= ((pw << 32) / s) << sh;
                H.twi[ti][k].x |= inv << sh;
                H.twi[ti][k].y |= ((inv << 32) / s) << sh;
                pw = h_mulmod(pw, psi, s);
            }
        }
    }

    // pass-9 copies (params.h: kTwPass9)
    for (int ti = 0; ti < kNumTab; ti++)
        for (auto *tab : {&H.twf[ti], &H.twi[ti]}) {
            tab->resize(kTwEntries);
            for (int t = 0; t < 512; t++) {
                (*tab)[kN + 0 * 512 + t] = (*tab)[512 + t];
                for (int h = 0; h < 2; h++) (*tab)[kN + (1 + h) * 512 + t] = (*tab)[1024 + 2 * t + h];
                for (int h = 0; h < 4; h++) (*tab)[kN + (3 + h) * 512 + t] = (*tab)[2048 + 4 * t + h];
            }
        }

    DevConsts &c = H.dc;
    memset(&c, 0, sizeof(c));
    const u64 q0 = kModulus[MQ0], q1 = kModulus[MQ1], P = kModulus[MP];
    const u64 b0 = kModulus[MB0], b1 = kModulus[MB1], msk = kModulus[MSK];
    const u64 qs[2] = {q0, q1};
    const u64 bsk[3] = {b0, b1, msk};
    const u128 q = (u128)q0 * q1;

    for (int mi = 0; mi < kNumMod; mi++) {
        const u64 m = kModulus[mi];
        u64 ninv = h_invmod(kN, m);
        c.ninv[mi] = mk_shoup(ninv, m);
        c.ninv_t[mi] = mk_shoup(h_mulmod(ninv, kT % m, m), m);
        const u64 w_last = H.twi[mi][1].x;  // inverse twiddle of the last Gentleman-Sande stage
        c.ninv_w[mi] = mk_shoup(h_mulmod(ninv, w_last, m), m);
        c.ninv_t_w[mi] = mk_shoup(h_mulmod(c.ninv_t[mi].w, w_last, m), m);
    }
    H.inv_q1_mod_q0 = h_invmod(q1 % q0, q0);
    H.inv_q0_mod_q1 = h_invmod(q0 % q1, q1);
    const u64 inv_punct_q[2] = {H.inv_q1_mod_q0, H.inv_q0_mod_q1};

    // base extension via m_tilde (integer domain: k_ext_conv needs q as words and -q mod p_k)
    for (int l = 0; l < 2; l++) {
        c.ext_in[l] = mk_shoup(h_mulmod(kMTilde % qs[l], inv_punct_q[l], qs[l]), qs[l]);
        c.inv_punct_q[l] = mk_shoup(inv_punct_q[l], qs[l]);
    }
    {
        u32 ql = (u32)(u64)q, x = ql;  // Newton iteration for q^-1 mod 2^32
        for (int i = 0; i < 5; i++) x *= 2u - ql * x;
        c.neg_inv_q_mod_mtilde = 0u - x;
    }
    c.q_w[0] = (u32)(u64)q;
    c.q_w[1] = (u32)((u64)q >> 32);
    c.q_w[2] = (u32)(u64)(q >> 64);
    u64 inv_q_mod[3];  // q^-1 mod p_k  (fast_floor)
    for (int k = 0; k < 3; k++) {
        const u64 p = bsk[k];
        const u64 q_mod_p = (u64)(q % p);
        inv_q_mod[k] = h_invmod(q_mod_p, p);
        c.extNeg[k] = (p - q_mod_p) % p;
    }
    // Shenoy-Kumaresan, merged with fast_floor's last step (devconsts.h)
    const u64 punct_B[2] = {b1, b0};
    const u64 inv_punct_B[2] = {h_invmod(b1 % b0, b0), h_invmod(b0 % b1, b1)};  // (B/b_j)^-1 mod b_j
    const u128 B = (u128)b0 * b1;
    const u64 inv_B_mod_msk = h_invmod((u64)(B % msk), msk);
    for (int j = 0; j < 2; j++) {
        c.skV[j] = h_mulmod(inv_q_mod[j], inv_punct_B[j], bsk[j]);
        c.skVs[j] = shoup_of(c.skV[j], bsk[j]);
        c.skD[j] = (u32)(msk - bsk[j]);
        for (int l = 0; l < 2; l++) c.pBq[j][l] = mk_shoup(punct_B[j] % qs[l], qs[l]);
    }
    c.alK2 = (msk - h_mulmod(inv_q_mod[2], inv_B_mod_msk, msk)) % msk;
    c.alK2s = shoup_of(c.alK2, msk);
    c.nib = mk_shoup((msk - inv_B_mod_msk) % msk, msk);
    for (int l = 0; l < 2; l++) {
        const u64 bq = (u64)(B % qs[l]);
        c.Bq[l] = mk_shoup(bq, qs[l]);
        c.nBq[l] = mk_shoup(qs[l] - bq, qs[l]);
    }
    // exact recovery of the tensor's q-limbs from its Bsk limbs (devconsts.h)
    for (int i = 0; i < 3; i++) {
        const u64 p = bsk[i];
        const u64 o1 = bsk[(i + 1) % 3], o2 = bsk[(i + 2) % 3];  // Bsk / p_i = o1 * o2
        c.crt3[i] = mk_shoup(h_invmod(h_mulmod(o1 % p, o2 % p, p), p), p);
        for (int l = 0; l < 2; l++)
            c.crtK[i][l] = mk_shoup(h_mulmod(h_mulmod(o1 % qs[l], o2 % qs[l], qs[l]), inv_punct_q[l], qs[l]), qs[l]);
    }
    for (int l = 0; l < 2; l++) {
        const u64 b3 = h_mulmod(h_mulmod(b0 % qs[l], b1 % qs[l], qs[l]), msk % qs[l], qs[l]);
        c.crtNB[l] = mk_shoup(h_mulmod((qs[l] - b3) % qs[l], inv_punct_q[l], qs[l]), qs[l]);
    }
    // BEHZ on the dual base (devconsts.h)
    {
        typedef unsigned __int128 U;
        auto mulmod_many = [](const u64 *f, int n, int skip, u64 m) {  // prod_{i != skip} f_i mod m
            u64 r = 1 % m;
            for (int i = 0; i < n; i++)
                if (i != skip) r = h_mulmod(r, f[i] % m, m);
            return r;
        };
        u64 sp[6];
        for (int i = 0; i < 6; i++) sp[i] = kDualPrime[i];
        {   // S must exceed 2 |t D|: |t D| <= t N (q/2)^2 (1 + 2^-29)
            long double lg = 0;
            for (int i = 0; i < 6; i++) lg += log2l((long double)sp[i]);
            const long double need = 1 + kLogT + kLogN + 2 * (log2l((long double)q0) + log2l((long double)q1) - 1) + 0.01L;
            if (lg < need + 8) throw std::runtime_error("fhe_b200: dual base too small");
        }
        for (int i = 0; i < 6; i++) {
            const u64 s = sp[i];
            c.d_R32[i] = (u32)(((u64)1 << 32) % s);
            c.d_R61[i] = (u32)(((u64)1 << 61) % s);
            c.d_NQ[i] = (u32)((s - (u64)(q % s)) % s);
            c.d_mu61[i] = (u32)(((u64)1 << 61) / s);
            const u64 ci = h_invmod(mulmod_many(sp, 6, i, s), s);
            c.d_C[i] = (u32)ci;
            c.d_Cs[i] = (u32)((ci << 32) / s);
            c.d_R48[i] = (u32)(((u64)1 << 48) / s);
            for (int l = 0; l < 2; l++) c.d_K[i][l] = mk_shoup(h_mulmod(mulmod_many(sp, 6, i, qs[l]), inv_punct_q[l], qs[l]), qs[l]);
        }
        for (int l = 0; l < 2; l++) {
            const u64 sq = mulmod_many(sp, 6, -1, qs[l]);
            c.d_KN[l] = h_mulmod((qs[l] - sq) % qs[l], inv_punct_q[l], qs[l]);
            c.d_NS4[l] = (qs[l] - mulmod_many(sp, 4, -1, qs[l])) % qs[l];
        }
        for (int i = 0; i < 4; i++) {
            const u64 s = sp[i];
            c.d_R58[i] = (u32)(((u64)1 << 58) % s);
            const u64 w = h_mulmod(h_invmod((u64)(q % s), s), h_invmod(mulmod_many(sp, 4, i, s), s), s);
            c.d_W[i] = (u32)w;
            c.d_Ws[i] = (u32)((w << 32) / s);
            for (int l = 0; l < 2; l++) c.d_P[i][l] = mk_shoup(mulmod_many(sp, 4, i, qs[l]), qs[l]);
        }
        for (int d = 0; d < kNumDual; d++) {
            u64 w[2] = {0, 0}, ws[2] = {0, 0}, ww[2] = {0, 0}, wws[2] = {0, 0};
            for (int lane = 0; lane < 2; lane++) {
                const u64 s = sp[2 * d + lane];
                const u64 nt = h_mulmod(h_invmod(kN, s), kT % s, s);
                const u64 w_last = (H.twi[kNumMod + d][1].x >> (32 * lane)) & 0xffffffffull;
                const u64 ntw = h_mulmod(nt, w_last, s);
                w[lane] = nt, ws[lane] = (nt << 32) / s, ww[lane] = ntw, wws[lane] = (ntw << 32) / s;
            }
            c.d_ninv_t[d].w = w[0] | (w[1] << 32), c.d_ninv_t[d].ws = ws[0] | (ws[1] << 32);
            c.d_ninv_t_w[d].w = ww[0] | (ww[1] << 32), c.d_ninv_t_w[d].ws = wws[0] | (wws[1] << 32);
        }
        // key switch on the dual base
        const u64 km[3] = {q0, q1, P};
        for (int d = 0; d < kNumDual; d++) {
            u64 w[2], ws[2], ww[2], wws[2];
            for (int lane = 0; lane < 2; lane++) {
                const u64 s = sp[2 * d + lane];
                const u64 ni = h_invmod(kN, s);
                const u64 w_last = (H.twi[kNumMod + d][1].x >> (32 * lane)) & 0xffffffffull;
                const u64 niw = h_mulmod(ni, w_last, s);
                w[lane] = ni, ws[lane] = (ni << 32) / s, ww[lane] = niw, wws[lane] = (niw << 32) / s;
            }
            c.d_ninv[d].w = w[0] | (w[1] << 32), c.d_ninv[d].ws = ws[0] | (ws[1] << 32);
            c.d_ninv_w[d].w = ww[0] | (ww[1] << 32), c.d_ninv_w[d].ws = wws[0] | (wws[1] << 32);
        }
        for (int m = 0; m < 3; m++) {
            const u64 o1 = km[(m + 1) % 3], o2 = km[(m + 2) % 3];  // Q / m = o1 o2
            c.lk_C[m] = mk_shoup(h_invmod(h_mulmod(o1 % km[m], o2 % km[m], km[m]), km[m]), km[m]);
            for (int i = 0; i < 6; i++) {
                const u64 r = h_mulmod(o1 % sp[i], o2 % sp[i], sp[i]);
                c.lk_R[m][i] = (u32)r;
                c.lk_Rs[m][i] = (u32)((r << 32) / sp[i]);
            }
            c.ksKN[m] = (km[m] - mulmod_many(sp, 6, -1, km[m])) % km[m];
            for (int i = 0; i < 6; i++) c.ksK[i][m] = mk_shoup(mulmod_many(sp, 6, i, km[m]), km[m]);
        }
        for (int l = 0; l < 2; l++) {  // the division by P folded into the q-limb sums (k_ks_finish_ksd)
            const u64 ql = km[l], ip = h_invmod(P % ql, ql);
            for (int i = 0; i < 6; i++) c.ksKd[i][l] = h_mulmod(c.ksK[i][l].w, ip, ql);
            c.ksKNd[l] = h_mulmod(c.ksKN[l], ip, ql);
            c.ksNd[l] = (ql - ip) % ql;
            c.ksNd30[l] = h_mulmod(c.ksNd[l], (1ull << 30) % ql, ql);
            c.ksHd[l] = h_mulmod((P >> 1) % ql, ip, ql);
        }
        (void)sizeof(U);
    }
    // key switching
    c.half_P = P >> 1;
    for (int l = 0; l < 2; l++) {
        c.inv_P_mod_q[l] = mk_shoup(h_invmod(P % qs[l], qs[l]), qs[l]);
        c.half_P_mod_q[l] = c.half_P % qs[l];
    }
    for (int l = 0; l < 2; l++) c.crt_inv[l] = mk_shoup(inv_punct_q[l], qs[l]);
    c.q_lo = (u64)q;
    c.q_hi = (u64)(q >> 64);
    c.qhalf_lo = (u64)((q - 1) / 2);
    c.qhalf_hi = (u64)(((q - 1) / 2) >> 64);
    // plain ops
    const u128 delta = q / kT;
    for (int l = 0; l < 2; l++) {
        c.delta_mod_q[l] = (u64)(delta % qs[l]);
        c.upper_half_incr[l] = qs[l] - kT;
    }
    c.q_mod_t = (u64)(q % kT);
    c.upper_half_threshold = (kT + 1) >> 1;

    // parms_id (SEAL 4.0): BLAKE2b-256 over LE u64 words [scheme=BFV(1), N, q_i..., t]
    {
        uint64_t w_key[6] = {1, (uint64_t)kN, q0, q1, P, kT};
        uint64_t w_data[5] = {1, (uint64_t)kN, q0, q1, kT};
        blake2b(w_key, sizeof(w_key), H.parms_id_key, 32);
        blake2b(w_data, sizeof(w_data), H.parms_id_data, 32);
    }
}
}  // namespace

const HostContext &HostContext::get() {
    static HostContext *inst = nullptr;
    static std::once_flag flag;
    static std::string err;
    std::call_once(flag, [] {
        HostContext *h = new HostContext();
        try {
            build(*h);
            inst = h;
        } catch (const std::exception &e) {
            err = e.what();
            delete h;
        }
    });
    if (!inst) throw std::runtime_error(err.empty() ? "fhe_b200: context build failed" : err);
    return *inst;
}

// ---------------------------------------------------------------- per-device context
namespace {
constexpr int kMaxDevices = 64;
DeviceContext g_dev[kMaxDevices];
std::mutex g_dev_mu[kMaxDevices];

void cuda_check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("fhe_b200: CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
}  // namespace

int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

DeviceContext &device_context(int device) {
    if (device < 0 || device >= kMaxDevices) throw std::runtime_error("fhe_b200: bad device index");
    std::lock_guard<std::mutex> lk(g_dev_mu[device]);
    DeviceContext &d = g_dev[device];
    cuda_check(cudaSetDevice(device), "cudaSetDevice (no CUDA device? this library has no CPU fallback)");
    if (d.device == device) return d;
    const HostContext &H = HostContext::get();
    const size_t per = (size_t)kTwEntries * sizeof(ulonglong2);
    void *mem = nullptr;
    cuda_check(cudaMalloc(&mem, per * kNumTab * 2), "cudaMalloc(twiddles)");
    for (int mi = 0; mi < kNumTab; mi++) {
        char *f = (char *)mem + per * (size_t)(2 * mi);
        char *i = f + per;
        cuda_check(cudaMemcpy(f, H.twf[mi].data(), per, cudaMemcpyHostToDevice), "upload twf");
        cuda_check(cudaMemcpy(i, H.twi[mi].data(), per, cudaMemcpyHostToDevice), "upload twi");
        d.tabs.twf[mi] = (const ulonglong2 *)f;
        d.tabs.twi[mi] = (const ulonglong2 *)i;
    }
    {
        std::vector<DevTwLow> lo(1);
        for (int mi = 0; mi < kNumTab; mi++)
            for (int k = 0; k < 64; k++) {
                lo[0].f[mi][k] = H.twf[mi][(size_t)k];
                lo[0].i[mi][k] = H.twi[mi][(size_t)k];
            }
        cuda_check(upload_constants(H.dc, d.tabs, lo[0]), "upload constants");
    }
    cuda_check(kernels_configure(), "kernel attributes");
    d.table_mem = mem;
    d.device = device;
    return d;
}

}  // namespace fheb
