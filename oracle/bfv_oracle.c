/*
 * TEST INFRASTRUCTURE ONLY -- see bfv_oracle.h for scope, provenance and pinning status.
 *
 * CPU restatement of the SEAL 4.0 BFV routines that the reference's add/sub/mul
 * precompiles execute (reference call site: FheApp::run, /root/reference/src/fhe.rs:138-152;
 * op bodies fhe.rs:785-806 and 814-1022; parameters testnet.rs:8-14).
 * SEAL itself is an un-vendored dependency (sunscreen 0.8.1 -> seal_fhe -> SEAL 4.0), so each
 * function names the SEAL 4.0 routine whose published algorithm it restates.
 */
#define _GNU_SOURCE
#include "bfv_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

#define N BFVO_N
#define LOGN BFVO_LOGN
#define T BFVO_T

/* ------------------------------------------------------------------ modulus */
typedef struct {
    u64 q;
    u64 r0, r1; /* floor(2^128 / q) -- SEAL Modulus::const_ratio */
} Mod;

static void mod_init(Mod *m, u64 q) {
    m->q = q;
    /* floor(2^128/q): 2^128 = (2^128-1) + 1 */
    u128 all = ~(u128)0;
    u128 quo = all / q;
    u128 rem = all % q;
    if (rem + 1 == q) quo += 1;
    m->r0 = (u64)quo;
    m->r1 = (u64)(quo >> 64);
}

/* SEAL util::barrett_reduce_128 */
static inline u64 red128(u128 z, const Mod *m) {
    u64 z0 = (u64)z, z1 = (u64)(z >> 64);
    u64 carry = (u64)(((u128)z0 * m->r0) >> 64);
    u128 t2 = (u128)z0 * m->r1;
    u128 t1 = (u128)(u64)t2 + carry;
    u64 tmp3 = (u64)(t2 >> 64) + (u64)(t1 >> 64);
    t2 = (u128)z1 * m->r0;
    u128 t1b = (u128)(u64)t1 + (u64)t2;
    carry = (u64)(t2 >> 64) + (u64)(t1b >> 64);
    u64 quo = z1 * m->r1 + tmp3 + carry;
    u64 r = z0 - quo * m->q;
    while (r >= m->q) r -= m->q;
    return r;
}
/* SEAL util::barrett_reduce_64 */
static inline u64 red64(u64 x, const Mod *m) {
    u64 quo = (u64)(((u128)x * m->r1) >> 64);
    u64 r = x - quo * m->q;
    while (r >= m->q) r -= m->q;
    return r;
}
static inline u64 mulmod(u64 a, u64 b, const Mod *m) { return red128((u128)a * b, m); }
static inline u64 addmod(u64 a, u64 b, const Mod *m) {
    u64 s = a + b;
    return s >= m->q ? s - m->q : s;
}
static inline u64 submod(u64 a, u64 b, const Mod *m) { return a >= b ? a - b : a + m->q - b; }
static inline u64 negmod(u64 a, const Mod *m) { return a ? m->q - a : 0; }
static u64 powmod(u64 b, u64 e, const Mod *m) {
    u64 r = 1;
    b = red64(b, m);
    while (e) {
        if (e & 1) r = mulmod(r, b, m);
        b = mulmod(b, b, m);
        e >>= 1;
    }
    return r;
}
static u64 invmod_prime(u64 a, const Mod *m) { return powmod(a, m->q - 2, m); }

/* SEAL MultiplyUIntModOperand (Shoup): quotient = floor(operand * 2^64 / q) */
static inline u64 shoup_quot(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
/* lazy: result in [0, 2q) */
static inline u64 shoup_lazy(u64 x, u64 w, u64 wq, u64 q) {
    u64 hi = (u64)(((u128)x * wq) >> 64);
    return x * w - hi * q;
}

/* ------------------------------------------------------------------ primes and roots */
static u64 mulmod_slow(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
static u64 powmod_slow(u64 b, u64 e, u64 q) {
    u64 r = 1;
    b %= q;
    while (e) {
        if (e & 1) r = mulmod_slow(r, b, q);
        b = mulmod_slow(b, b, q);
        e >>= 1;
    }
    return r;
}
/* deterministic Miller-Rabin for 64-bit */
static int is_prime(u64 n) {
    static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (size_t i = 0; i < 12; i++) {
        if (n % bases[i] == 0) return n == bases[i];
    }
    u64 d = n - 1;
    int s = 0;
    while (!(d & 1)) {
        d >>= 1;
        s++;
    }
    for (size_t i = 0; i < 12; i++) {
        u64 x = powmod_slow(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int r = 1; r < s; r++) {
            x = mulmod_slow(x, x, n);
            if (x == n - 1) {
                comp = 0;
                break;
            }
        }
        if (comp) return 0;
    }
    return 1;
}
/* SEAL util::get_primes(factor = 2N, bit_size = 61, count): descending primes = 1 mod 2N */
static void get_primes(u64 factor, int bits, size_t count, u64 *out) {
    u64 v = ((((u64)1) << bits) - 1) / factor * factor + 1;
    u64 lower = ((u64)1) << (bits - 1);
    size_t k = 0;
    while (k < count && v > lower) {
        if (is_prime(v)) out[k++] = v;
        v -= factor;
    }
}
/* SEAL util::try_minimal_primitive_root: the numerically smallest primitive 2N-th root */
static u64 minimal_primitive_root(u64 degree, const Mod *m) {
    u64 q = m->q;
    u64 cofactor = (q - 1) / degree;
    u64 root = 0;
    for (u64 g = 2; g < q; g++) {
        u64 c = powmod(g, cofactor, m);
        /* primitive iff c^(degree/2) == -1 */
        if (powmod(c, degree / 2, m) == q - 1) {
            root = c;
            break;
        }
    }
    u64 gsq = mulmod(root, root, m), cur = root, best = root;
    for (u64 i = 0; i < degree; i += 2) {
        if (cur < best) best = cur;
        cur = mulmod(cur, gsq, m);
    }
    return best;
}
static inline uint32_t bitrev(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

/* ------------------------------------------------------------------ context */
typedef struct {
    Mod mod[BFVO_NMOD];
    u64 root[BFVO_NMOD];
    u64 *rp[BFVO_NMOD], *rps[BFVO_NMOD];   /* rp[bitrev(i)] = psi^i, Shoup quotients */
    u64 *irp[BFVO_NMOD], *irps[BFVO_NMOD]; /* irp[k] = rp[k]^-1 */
    u64 ninv[BFVO_NMOD], ninvs[BFVO_NMOD];
    u64 gamma;
    Mod mtilde; /* 2^32 */
    /* data level */
    u64 delta_mod_q[2]; /* floor(q/t) mod q_l */
    u64 q_mod_t;
    u64 upper_half_threshold;    /* (t+1)>>1 */
    u64 upper_half_increment[2]; /* q_l - t */
    u64 inv_P_mod_q[2];
    u64 half_P;
    /* BEHZ (SEAL RNSTool), Bsk index order: b0, b1, msk */
    u64 mtilde_mod_q[2];
    u64 inv_punct_q[2];        /* (q/q_l)^-1 mod q_l */
    u64 punct_q_mod_bsk[2][3]; /* (q/q_l) mod p_k */
    u64 punct_q_mod_mtilde[2];
    u64 neg_inv_q_mod_mtilde;
    u64 q_mod_bsk[3];
    u64 inv_mtilde_mod_bsk[3];
    u64 inv_q_mod_bsk[3];
    u64 inv_punct_B[2];       /* (B/b_j)^-1 mod b_j */
    u64 punct_B_mod_q[2][2];  /* [j][l] (B/b_j) mod q_l */
    u64 punct_B_mod_msk[2];
    u64 inv_B_mod_msk;
    u64 B_mod_q[2];
    /* CRT for decrypt */
    u64 inv_q1_mod_q0, inv_q0_mod_q1;
} Ctx;

static Ctx C;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static const int BSK[3] = {BFVO_B0, BFVO_B1, BFVO_MSK};

static void build_tables(int mi) {
    const Mod *m = &C.mod[mi];
    u64 psi = minimal_primitive_root(2 * N, m);
    C.root[mi] = psi;
    C.rp[mi] = (u64 *)malloc(N * 8);
    C.rps[mi] = (u64 *)malloc(N * 8);
    C.irp[mi] = (u64 *)malloc(N * 8);
    C.irps[mi] = (u64 *)malloc(N * 8);
    u64 pw = 1;
    for (uint32_t i = 0; i < N; i++) {
        uint32_t k = bitrev(i, LOGN);
        C.rp[mi][k] = pw;
        pw = mulmod(pw, psi, m);
    }
    for (uint32_t k = 0; k < N; k++) {
        C.rps[mi][k] = shoup_quot(C.rp[mi][k], m->q);
        C.irp[mi][k] = invmod_prime(C.rp[mi][k], m);
        C.irps[mi][k] = shoup_quot(C.irp[mi][k], m->q);
    }
    C.ninv[mi] = invmod_prime(N, m);
    C.ninvs[mi] = shoup_quot(C.ninv[mi], m->q);
}

static void ctx_build(void) {
    /* testnet.rs:8-14 */
    mod_init(&C.mod[BFVO_Q0], 0xffffee001ull);
    mod_init(&C.mod[BFVO_Q1], 0xffffc4001ull);
    mod_init(&C.mod[BFVO_P], 0x1ffffe0001ull);
    /* SEAL RNSTool::initialize: |B| = |q| = 2 (32 + bits(t) + bits(q) = 32+13+72 < 61*2+61),
     * primes = get_primes(2N, 61, |Bsk|+1): first m_sk, second gamma, then B. m_tilde = 2^32 */
    u64 pr[4];
    get_primes(2 * N, 61, 4, pr);
    mod_init(&C.mod[BFVO_MSK], pr[0]);
    C.gamma = pr[1];
    mod_init(&C.mod[BFVO_B0], pr[2]);
    mod_init(&C.mod[BFVO_B1], pr[3]);
    mod_init(&C.mtilde, ((u64)1) << 32);
    for (int i = 0; i < BFVO_NMOD; i++) build_tables(i);

    const u64 q0 = C.mod[0].q, q1 = C.mod[1].q, P = C.mod[2].q;
    u128 q = (u128)q0 * q1;
    u128 delta = q / T;
    for (int l = 0; l < 2; l++) {
        C.delta_mod_q[l] = (u64)(delta % C.mod[l].q);
        C.upper_half_increment[l] = C.mod[l].q - T;
        C.inv_P_mod_q[l] = invmod_prime(P % C.mod[l].q, &C.mod[l]);
        C.mtilde_mod_q[l] = C.mtilde.q % C.mod[l].q;
    }
    C.q_mod_t = (u64)(q % T);
    C.upper_half_threshold = (T + 1) >> 1;
    C.half_P = P >> 1;
    C.inv_q1_mod_q0 = invmod_prime(q1 % q0, &C.mod[0]);
    C.inv_q0_mod_q1 = invmod_prime(q0 % q1, &C.mod[1]);
    C.inv_punct_q[0] = C.inv_q1_mod_q0; /* q/q0 = q1 */
    C.inv_punct_q[1] = C.inv_q0_mod_q1;
    u64 punct_q[2] = {q1, q0};
    for (int l = 0; l < 2; l++) {
        for (int k = 0; k < 3; k++) C.punct_q_mod_bsk[l][k] = punct_q[l] % C.mod[BSK[k]].q;
        C.punct_q_mod_mtilde[l] = punct_q[l] & 0xffffffffull;
    }
    /* -q^-1 mod 2^32 (Newton on the odd number q mod 2^32) */
    {
        uint32_t ql = (uint32_t)(u64)q;
        uint32_t x = ql;
        for (int i = 0; i < 5; i++) x *= 2u - ql * x;
        C.neg_inv_q_mod_mtilde = (u64)(uint32_t)(0u - x);
    }
    for (int k = 0; k < 3; k++) {
        const Mod *m = &C.mod[BSK[k]];
        C.q_mod_bsk[k] = (u64)(q % m->q);
        C.inv_mtilde_mod_bsk[k] = invmod_prime(C.mtilde.q % m->q, m);
        C.inv_q_mod_bsk[k] = invmod_prime(C.q_mod_bsk[k], m);
    }
    const u64 b0 = C.mod[BFVO_B0].q, b1 = C.mod[BFVO_B1].q;
    const Mod *msk = &C.mod[BFVO_MSK];
    C.inv_punct_B[0] = invmod_prime(b1 % b0, &C.mod[BFVO_B0]);
    C.inv_punct_B[1] = invmod_prime(b0 % b1, &C.mod[BFVO_B1]);
    u64 punct_B[2] = {b1, b0};
    for (int j = 0; j < 2; j++) {
        for (int l = 0; l < 2; l++) C.punct_B_mod_q[j][l] = punct_B[j] % C.mod[l].q;
        C.punct_B_mod_msk[j] = punct_B[j] % msk->q;
    }
    u128 B = (u128)b0 * b1;
    C.inv_B_mod_msk = invmod_prime((u64)(B % msk->q), msk);
    for (int l = 0; l < 2; l++) C.B_mod_q[l] = (u64)(B % C.mod[l].q);
}

int bfvo_init(void) {
    pthread_once(&g_once, ctx_build);
    return 0;
}

size_t bfvo_constants(uint64_t *out, size_t cap) {
    bfvo_init();
    u64 tmp[128];
    size_t n = 0;
    for (int i = 0; i < 6; i++) tmp[n++] = C.mod[i].q;
    for (int i = 0; i < 6; i++) tmp[n++] = C.root[i];
    tmp[n++] = C.gamma;
    for (int i = 0; i < 6; i++) tmp[n++] = C.ninv[i];
    for (int l = 0; l < 2; l++) tmp[n++] = C.delta_mod_q[l];
    for (int l = 0; l < 2; l++) tmp[n++] = C.inv_P_mod_q[l];
    tmp[n++] = C.half_P;
    for (int l = 0; l < 2; l++) tmp[n++] = C.mtilde_mod_q[l];
    for (int l = 0; l < 2; l++) tmp[n++] = C.inv_punct_q[l];
    for (int l = 0; l < 2; l++)
        for (int k = 0; k < 3; k++) tmp[n++] = C.punct_q_mod_bsk[l][k];
    for (int l = 0; l < 2; l++) tmp[n++] = C.punct_q_mod_mtilde[l];
    tmp[n++] = C.neg_inv_q_mod_mtilde;
    for (int k = 0; k < 3; k++) tmp[n++] = C.q_mod_bsk[k];
    for (int k = 0; k < 3; k++) tmp[n++] = C.inv_mtilde_mod_bsk[k];
    for (int k = 0; k < 3; k++) tmp[n++] = C.inv_q_mod_bsk[k];
    for (int j = 0; j < 2; j++) tmp[n++] = C.inv_punct_B[j];
    for (int j = 0; j < 2; j++)
        for (int l = 0; l < 2; l++) tmp[n++] = C.punct_B_mod_q[j][l];
    for (int j = 0; j < 2; j++) tmp[n++] = C.punct_B_mod_msk[j];
    tmp[n++] = C.inv_B_mod_msk;
    for (int l = 0; l < 2; l++) tmp[n++] = C.B_mod_q[l];
    size_t k = n < cap ? n : cap;
    memcpy(out, tmp, k * 8);
    return n;
}

/* ------------------------------------------------------------------ NTT */
/* SEAL ntt_negacyclic_harvey: Cooley-Tukey, natural in -> bit-reversed out, Harvey lazy
 * butterflies, final correction to [0,q). */
static void ntt_fwd(u64 *a, int mi) {
    const u64 q = C.mod[mi].q, two_q = 2 * q;
    const u64 *rp = C.rp[mi], *rps = C.rps[mi];
    for (size_t m = 1, gap = N >> 1; m < N; m <<= 1, gap >>= 1) {
        for (size_t i = 0; i < m; i++) {
            u64 w = rp[m + i], ws = rps[m + i];
            u64 *x = a + 2 * i * gap, *y = x + gap;
            for (size_t j = 0; j < gap; j++) {
                u64 X = x[j];
                if (X >= two_q) X -= two_q;
                u64 Tt = shoup_lazy(y[j], w, ws, q);
                x[j] = X + Tt;
                y[j] = X + two_q - Tt;
            }
        }
    }
    for (size_t j = 0; j < N; j++) {
        u64 v = a[j];
        if (v >= two_q) v -= two_q;
        if (v >= q) v -= q;
        a[j] = v;
    }
}
/* SEAL inverse_ntt_negacyclic_harvey: Gentleman-Sande, bit-reversed in -> natural out, x N^-1 */
static void ntt_inv(u64 *a, int mi) {
    const u64 q = C.mod[mi].q, two_q = 2 * q;
    const u64 *irp = C.irp[mi], *irps = C.irps[mi];
    for (size_t m = N >> 1, gap = 1; m >= 1; m >>= 1, gap <<= 1) {
        for (size_t i = 0; i < m; i++) {
            u64 w = irp[m + i], ws = irps[m + i];
            u64 *x = a + 2 * i * gap, *y = x + gap;
            for (size_t j = 0; j < gap; j++) {
                u64 X = x[j], Y = y[j];
                u64 S = X + Y;
                if (S >= two_q) S -= two_q;
                x[j] = S;
                y[j] = shoup_lazy(X + two_q - Y, w, ws, q);
            }
        }
    }
    const u64 ni = C.ninv[mi], nis = C.ninvs[mi];
    for (size_t j = 0; j < N; j++) {
        u64 v = shoup_lazy(a[j], ni, nis, q);
        if (v >= q) v -= q;
        a[j] = v;
    }
}
void bfvo_ntt_fwd(uint64_t *a, int mod) {
    bfvo_init();
    ntt_fwd(a, mod);
}
void bfvo_ntt_inv(uint64_t *a, int mod) {
    bfvo_init();
    ntt_inv(a, mod);
}

/* ------------------------------------------------------------------ add / sub / negate */
/* SEAL Evaluator::add_inplace / sub_inplace / negate_inplace (add_poly_coeffmod etc.) */
void bfvo_add(const u64 *a, const u64 *b, u64 *out, size_t npolys) {
    bfvo_init();
    for (size_t p = 0; p < npolys; p++)
        for (int l = 0; l < 2; l++) {
            size_t o = (p * 2 + l) * N;
            for (size_t i = 0; i < N; i++) out[o + i] = addmod(a[o + i], b[o + i], &C.mod[l]);
        }
}
void bfvo_sub(const u64 *a, const u64 *b, u64 *out, size_t npolys) {
    bfvo_init();
    for (size_t p = 0; p < npolys; p++)
        for (int l = 0; l < 2; l++) {
            size_t o = (p * 2 + l) * N;
            for (size_t i = 0; i < N; i++) out[o + i] = submod(a[o + i], b[o + i], &C.mod[l]);
        }
}
void bfvo_negate(const u64 *a, u64 *out, size_t npolys) {
    bfvo_init();
    for (size_t p = 0; p < npolys; p++)
        for (int l = 0; l < 2; l++) {
            size_t o = (p * 2 + l) * N;
            for (size_t i = 0; i < N; i++) out[o + i] = negmod(a[o + i], &C.mod[l]);
        }
}

/* SEAL util::multiply_add_plain_with_scaling_variant / multiply_sub_plain_with_scaling_variant */
static void plain_scaled(u64 *ct, const u64 *plain, size_t len, int subtract) {
    for (size_t i = 0; i < len; i++) {
        u128 numer = (u128)plain[i] * C.q_mod_t + C.upper_half_threshold;
        u64 fix = (u64)(numer / T);
        for (int l = 0; l < 2; l++) {
            const Mod *m = &C.mod[l];
            u64 scaled = addmod(mulmod(plain[i], C.delta_mod_q[l], m), red64(fix, m), m);
            u64 *c = ct + (size_t)l * N + i; /* poly 0 */
            *c = subtract ? submod(*c, scaled, m) : addmod(*c, scaled, m);
        }
    }
}
void bfvo_add_plain(u64 *ct, const u64 *plain, size_t len) {
    bfvo_init();
    plain_scaled(ct, plain, len, 0);
}
void bfvo_sub_plain(u64 *ct, const u64 *plain, size_t len) {
    bfvo_init();
    plain_scaled(ct, plain, len, 1);
}

/* SEAL Evaluator::multiply_plain_normal (generic branch; the monomial branch is the same map) */
void bfvo_multiply_plain(u64 *ct, size_t npolys, const u64 *plain, size_t len) {
    bfvo_init();
    u64 *tmp = (u64 *)calloc(2 * N, 8);
    for (int l = 0; l < 2; l++) {
        for (size_t i = 0; i < len; i++)
            tmp[l * N + i] = plain[i] + (plain[i] >= C.upper_half_threshold ? C.upper_half_increment[l] : 0);
        ntt_fwd(tmp + l * N, l);
    }
    for (size_t p = 0; p < npolys; p++)
        for (int l = 0; l < 2; l++) {
            u64 *c = ct + (p * 2 + l) * N;
            ntt_fwd(c, l);
            for (size_t i = 0; i < N; i++) c[i] = mulmod(c[i], tmp[l * N + i], &C.mod[l]);
            ntt_inv(c, l);
        }
    free(tmp);
}

/* ------------------------------------------------------------------ BEHZ multiply */
/* limb order in the 5-limb extended base: q0, q1, b0, b1, msk */
static const int EXT[5] = {BFVO_Q0, BFVO_Q1, BFVO_B0, BFVO_B1, BFVO_MSK};

/* SEAL RNSTool::fastbconv_m_tilde + RNSTool::sm_mrq for one polynomial.
 * in: [2][N] base q;  out: [5][N] (q-part copied, Bsk-part computed) */
static void behz_extend_poly(const u64 *in, u64 *out) {
    memcpy(out, in, 2 * N * 8);
    for (size_t i = 0; i < N; i++) {
        /* fastbconv_m_tilde: temp = in * m_tilde mod q_l; BaseConverter::fast_convert_array */
        u64 tmp[2];
        for (int l = 0; l < 2; l++) {
            u64 v = mulmod(in[l * N + i], C.mtilde_mod_q[l], &C.mod[l]);
            tmp[l] = mulmod(v, C.inv_punct_q[l], &C.mod[l]);
        }
        u64 y[3];
        for (int k = 0; k < 3; k++) {
            const Mod *m = &C.mod[BSK[k]];
            y[k] = red128((u128)tmp[0] * C.punct_q_mod_bsk[0][k] + (u128)tmp[1] * C.punct_q_mod_bsk[1][k], m);
        }
        u64 ymt = (u64)(((u128)tmp[0] * C.punct_q_mod_mtilde[0] + (u128)tmp[1] * C.punct_q_mod_mtilde[1]) & 0xffffffffull);
        /* sm_mrq */
        u64 r = (ymt * C.neg_inv_q_mod_mtilde) & 0xffffffffull;
        for (int k = 0; k < 3; k++) {
            const Mod *m = &C.mod[BSK[k]];
            u64 rr = r;
            if (rr >= (C.mtilde.q >> 1)) rr += m->q - C.mtilde.q;
            u64 v = addmod(mulmod(rr, C.q_mod_bsk[k], m), y[k], m);
            out[(2 + k) * N + i] = mulmod(v, C.inv_mtilde_mod_bsk[k], m);
        }
    }
}
void bfvo_behz_extend(const u64 *a, const u64 *b, u64 *ext) {
    bfvo_init();
    for (int p = 0; p < 2; p++) behz_extend_poly(a + (size_t)p * 2 * N, ext + (size_t)p * 5 * N);
    for (int p = 0; p < 2; p++) behz_extend_poly(b + (size_t)p * 2 * N, ext + (size_t)(2 + p) * 5 * N);
}
/* steps (3)-(6) of Evaluator::bfv_multiply: NTT, dyadic tensor, INTT, multiply by t */
void bfvo_behz_tensor(const u64 *ext, u64 *tens) {
    bfvo_init();
    u64 *w = (u64 *)malloc(4 * N * 8);
    for (int e = 0; e < 5; e++) {
        int mi = EXT[e];
        const Mod *m = &C.mod[mi];
        for (int p = 0; p < 4; p++) {
            memcpy(w + (size_t)p * N, ext + ((size_t)p * 5 + e) * N, N * 8);
            ntt_fwd(w + (size_t)p * N, mi);
        }
        const u64 *a0 = w, *a1 = w + N, *b0 = w + 2 * N, *b1 = w + 3 * N;
        u64 *d0 = tens + ((size_t)0 * 5 + e) * N, *d1 = tens + ((size_t)1 * 5 + e) * N, *d2 = tens + ((size_t)2 * 5 + e) * N;
        for (size_t i = 0; i < N; i++) {
            d0[i] = mulmod(a0[i], b0[i], m);
            d1[i] = addmod(mulmod(a0[i], b1[i], m), mulmod(a1[i], b0[i], m), m);
            d2[i] = mulmod(a1[i], b1[i], m);
        }
        u64 *d[3] = {d0, d1, d2};
        for (int p = 0; p < 3; p++) {
            ntt_inv(d[p], mi);
            for (size_t i = 0; i < N; i++) d[p][i] = mulmod(d[p][i], T, m);
        }
    }
    free(w);
}
/* SEAL RNSTool::fast_floor + RNSTool::fastbconv_sk, per coefficient */
void bfvo_behz_floor_sk(const u64 *tens, u64 *out3) {
    bfvo_init();
    const Mod *msk = &C.mod[BFVO_MSK];
    for (int p = 0; p < 3; p++) {
        const u64 *in = tens + (size_t)p * 5 * N;
        u64 *out = out3 + (size_t)p * 2 * N;
        for (size_t i = 0; i < N; i++) {
            /* fast_floor: q-part -> Bsk, then (in_Bsk - conv) * q^-1 */
            u64 tmp[2];
            for (int l = 0; l < 2; l++) tmp[l] = mulmod(in[l * N + i], C.inv_punct_q[l], &C.mod[l]);
            u64 f[3];
            for (int k = 0; k < 3; k++) {
                const Mod *m = &C.mod[BSK[k]];
                u64 c = red128((u128)tmp[0] * C.punct_q_mod_bsk[0][k] + (u128)tmp[1] * C.punct_q_mod_bsk[1][k], m);
                f[k] = mulmod(in[(2 + k) * N + i] + (m->q - c), C.inv_q_mod_bsk[k], m);
            }
            /* fastbconv_sk: B -> q and B -> m_sk */
            u64 tb[2];
            for (int j = 0; j < 2; j++) tb[j] = mulmod(f[j], C.inv_punct_B[j], &C.mod[BSK[j]]);
            u64 h = red128((u128)tb[0] * C.punct_B_mod_msk[0] + (u128)tb[1] * C.punct_B_mod_msk[1], msk);
            u64 alpha = mulmod(h + (msk->q - f[2]), C.inv_B_mod_msk, msk);
            for (int l = 0; l < 2; l++) {
                const Mod *m = &C.mod[l];
                u64 g = red128((u128)tb[0] * C.punct_B_mod_q[0][l] + (u128)tb[1] * C.punct_B_mod_q[1][l], m);
                u64 v;
                if (alpha > (msk->q >> 1))
                    v = addmod(mulmod(red64(msk->q - alpha, m), C.B_mod_q[l], m), g, m);
                else
                    v = addmod(mulmod(red64(alpha, m), m->q - C.B_mod_q[l], m), g, m);
                out[l * N + i] = v;
            }
        }
    }
}
/* SEAL Evaluator::bfv_multiply */
void bfvo_multiply(const u64 *a, const u64 *b, u64 *out3) {
    u64 *ext = (u64 *)malloc((size_t)4 * 5 * N * 8);
    u64 *tens = (u64 *)malloc((size_t)3 * 5 * N * 8);
    bfvo_behz_extend(a, b, ext);
    bfvo_behz_tensor(ext, tens);
    bfvo_behz_floor_sk(tens, out3);
    free(ext);
    free(tens);
}

/* ------------------------------------------------------------------ relinearize */
/* SEAL Evaluator::relinearize_internal -> switch_key_inplace (BFV branch), size 3 -> 2.
 * rk: [digit][poly][limb q0,q1,P][N] in NTT form. */
void bfvo_relinearize(const u64 *ct3, const u64 *rk, u64 *out2) {
    bfvo_init();
    const u64 *c2 = ct3 + (size_t)2 * 2 * N;
    u64 *acc = (u64 *)malloc((size_t)2 * 3 * N * 8); /* [poly][limb J][N] */
    u64 *dig = (u64 *)malloc(N * 8);
    u128 *lazy = (u128 *)malloc((size_t)2 * N * sizeof(u128));
    for (int J = 0; J < 3; J++) {
        const Mod *m = &C.mod[J];
        memset(lazy, 0, (size_t)2 * N * sizeof(u128));
        for (int I = 0; I < 2; I++) {
            for (size_t i = 0; i < N; i++) {
                u64 v = c2[(size_t)I * N + i];
                dig[i] = (C.mod[I].q <= m->q) ? v : red64(v, m);
            }
            ntt_fwd(dig, J);
            for (int k = 0; k < 2; k++) {
                const u64 *key = rk + (((size_t)I * 2 + k) * 3 + J) * N;
                for (size_t i = 0; i < N; i++) lazy[(size_t)k * N + i] += (u128)dig[i] * key[i];
            }
        }
        for (int k = 0; k < 2; k++)
            for (size_t i = 0; i < N; i++) acc[((size_t)k * 3 + J) * N + i] = red128(lazy[(size_t)k * N + i], m);
    }
    const Mod *mp = &C.mod[BFVO_P];
    for (int k = 0; k < 2; k++) {
        u64 *last = acc + ((size_t)k * 3 + 2) * N;
        ntt_inv(last, BFVO_P);
        for (size_t i = 0; i < N; i++) last[i] = red64(last[i] + C.half_P, mp);
        for (int l = 0; l < 2; l++) {
            const Mod *m = &C.mod[l];
            u64 *x = acc + ((size_t)k * 3 + l) * N;
            ntt_inv(x, l);
            u64 half_mod = red64(C.half_P, m);
            const u64 *cin = ct3 + ((size_t)k * 2 + l) * N;
            u64 *o = out2 + ((size_t)k * 2 + l) * N;
            for (size_t i = 0; i < N; i++) {
                u64 tl = submod(red64(last[i], m), half_mod, m);
                u64 v = mulmod(submod(x[i], tl, m), C.inv_P_mod_q[l], m);
                o[i] = addmod(v, cin[i], m);
            }
        }
    }
    free(acc);
    free(dig);
    free(lazy);
}
void bfvo_mul_relin(const u64 *a, const u64 *b, const u64 *rk, u64 *out2) {
    u64 *c3 = (u64 *)malloc((size_t)3 * 2 * N * 8);
    bfvo_multiply(a, b, c3);
    bfvo_relinearize(c3, rk, out2);
    free(c3);
}

/* ------------------------------------------------------------------ encoders (sunscreen types) */
static size_t sigbits64(u64 v) { return v ? 64 - (size_t)__builtin_clzll(v) : 0; }
size_t bfvo_encode_i64(int64_t v, u64 *plain) {
    u64 mag = v < 0 ? (u64)0 - (u64)v : (u64)v;
    size_t n = sigbits64(mag);
    for (size_t i = 0; i < n; i++) {
        u64 bit = (mag >> i) & 1;
        plain[i] = (v < 0) ? bit * (T - bit) : bit;
    }
    return n;
}
size_t bfvo_encode_u64(u64 v, u64 *plain) {
    size_t n = sigbits64(v);
    for (size_t i = 0; i < n; i++) plain[i] = (v >> i) & 1;
    return n;
}
size_t bfvo_encode_u256(const u64 limbs[4], u64 *plain) {
    size_t n = 0;
    for (int w = 3; w >= 0; w--)
        if (limbs[w]) {
            n = (size_t)w * 64 + sigbits64(limbs[w]);
            break;
        }
    for (size_t i = 0; i < n; i++) plain[i] = (limbs[i / 64] >> (i % 64)) & 1;
    return n;
}
size_t bfvo_encode_f64(double v, u64 *plain) {
    if (isnan(v) || isinf(v)) return 0;
    memset(plain, 0, N * 8);
    if (v == 0.0 || fpclassify(v) == FP_SUBNORMAL) return N;
    u64 bits;
    memcpy(&bits, &v, 8);
    u64 mant = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
    int64_t power = (int64_t)((bits >> 52) & 0x7ff) - 1023;
    u64 sign = bits >> 63;
    if (power + 1 > 64) return 0;
    for (int i = 0; i < 53; i++) {
        u64 bit = (mant >> i) & 1;
        int64_t bp = power - (53 - i - 1);
        size_t idx = bp >= 0 ? (size_t)bp : (size_t)((int64_t)N + bp);
        u64 s = bp >= 0 ? sign : (sign ^ 1);
        plain[idx] = (s == 0) ? bit : (bit ? T - bit : 0);
    }
    return N;
}
int64_t bfvo_decode_i64(const u64 *plain, size_t len) {
    size_t bits = len < 64 ? len : 64;
    u64 cutoff = (T + 1) / 2;
    u64 val = 0; /* wrapping, as Rust release arithmetic */
    for (size_t i = 0; i < bits; i++) {
        u64 c = plain[i];
        if (c < cutoff)
            val += ((u64)1 << i) * c;
        else
            val -= ((u64)1 << i) * (T - c);
    }
    return (int64_t)val;
}
uint64_t bfvo_decode_u64(const u64 *plain, size_t len) { return (u64)bfvo_decode_i64(plain, len); }
void bfvo_decode_u256(const u64 *plain, size_t len, u64 limbs[4]) {
    size_t bits = len < 256 ? len : 256;
    u64 cutoff = (T + 1) / 2;
    u64 acc[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < bits; i++) {
        u64 c = plain[i];
        int neg = c >= cutoff;
        u64 mag = neg ? T - c : c;
        /* term = mag << i as 256-bit */
        u64 term[4] = {0, 0, 0, 0};
        size_t w = i / 64, s = i % 64;
        term[w] = mag << s;
        if (s && w + 1 < 4) term[w + 1] = mag >> (64 - s);
        unsigned char carry = 0;
        for (int k = 0; k < 4; k++) {
            u128 r;
            if (!neg) {
                r = (u128)acc[k] + term[k] + carry;
                acc[k] = (u64)r;
                carry = (unsigned char)(r >> 64);
            } else {
                u128 sub = (u128)term[k] + carry;
                carry = (u128)acc[k] < sub;
                acc[k] = (u64)((u128)acc[k] - sub);
            }
        }
    }
    memcpy(limbs, acc, 32);
}
double bfvo_decode_f64(const u64 *plain, size_t len) {
    double val = 0.0;
    u64 cutoff = (T + 1) / 2;
    size_t n = len < N ? len : N;
    for (size_t i = 0; i < n; i++) {
        int64_t power = i < 64 ? (int64_t)i : (int64_t)i - (int64_t)N;
        double sign = power >= 0 ? 1.0 : -1.0;
        u64 c = plain[i];
        if (c < cutoff)
            val += sign * (double)c * exp2((double)power);
        else
            val -= sign * (double)(T - c) * exp2((double)power);
    }
    return val;
}

/* ------------------------------------------------------------------ encrypt / decrypt */
static inline u64 splitmix(u64 *s) {
    u64 z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* Shape of SEAL encrypt_zero_asymmetric at key level + divide_and_round_q_last_inplace +
 * multiply_add_plain_with_scaling_variant; randomness is our own. */
void bfvo_encrypt(const u64 *pk, const u64 *plain, size_t len, u64 seed, u64 *ct) {
    u64 st = seed * 0xD1342543DE82EF95ull + 0x1234567;
    int8_t *smp = (int8_t *)malloc((size_t)3 * N);
    for (size_t i = 0; i < N; i++) {
        u64 r;
        do r = splitmix(&st) & 3;
        while (r == 3);
        smp[i] = (int8_t)r - 1;
    }
    for (int j = 0; j < 2; j++)
        for (size_t i = 0; i < N; i++) { /* centred binomial, 21 bits each side (sigma ~3.24) */
            u64 r = splitmix(&st);
            smp[(size_t)(1 + j) * N + i] = (int8_t)(__builtin_popcountll(r & 0x1fffff) - __builtin_popcountll((r >> 21) & 0x1fffff));
        }
    bfvo_encrypt_samples(pk, plain, len, smp, smp + N, smp + 2 * N, ct);
    free(smp);
}
/* The same with the samples supplied by the caller: u ternary, e0 / e1 small errors (int8).  Used to check the GPU
 * encryptor, whose sampler (ChaCha12 + inverse-CDF, kernels.cu) is restated in oracle/bfv.py. */
void bfvo_encrypt_samples(const u64 *pk, const u64 *plain, size_t len, const int8_t *u, const int8_t *e0, const int8_t *e1, u64 *ct) {
    bfvo_init();
    u64 *un = (u64 *)malloc((size_t)3 * N * 8);
    for (int J = 0; J < 3; J++) {
        for (size_t i = 0; i < N; i++) un[(size_t)J * N + i] = u[i] < 0 ? C.mod[J].q - 1 : (u64)u[i];
        ntt_fwd(un + (size_t)J * N, J);
    }
    u64 *c = (u64 *)malloc((size_t)3 * N * 8);
    for (int j = 0; j < 2; j++) {
        const int8_t *e = j == 0 ? e0 : e1;
        for (int J = 0; J < 3; J++) {
            const Mod *m = &C.mod[J];
            u64 *x = c + (size_t)J * N;
            const u64 *k = pk + ((size_t)j * 3 + J) * N;
            for (size_t i = 0; i < N; i++) x[i] = mulmod(k[i], un[(size_t)J * N + i], m);
            ntt_inv(x, J);
            for (size_t i = 0; i < N; i++) x[i] = e[i] < 0 ? submod(x[i], (u64)(-e[i]), m) : addmod(x[i], (u64)e[i], m);
        }
        /* RNSTool::divide_and_round_q_last_inplace */
        u64 *last = c + (size_t)2 * N;
        for (size_t i = 0; i < N; i++) last[i] = red64(last[i] + C.half_P, &C.mod[BFVO_P]);
        for (int l = 0; l < 2; l++) {
            const Mod *m = &C.mod[l];
            u64 half_mod = red64(C.half_P, m);
            u64 *o = ct + ((size_t)j * 2 + l) * N;
            for (size_t i = 0; i < N; i++) {
                u64 tl = submod(red64(last[i], m), half_mod, m);
                o[i] = mulmod(submod(c[(size_t)l * N + i], tl, m), C.inv_P_mod_q[l], m);
            }
        }
    }
    plain_scaled(ct, plain, len, 0);
    free(un);
    free(c);
}

/* ------------------------------------------------------------------ SEAL-exact deterministic encryption
 * What `FheApp::encrypt` / `reencrypt` (/root/reference/src/fhe.rs:594-657) hand to sunscreen 0.8.1's
 * `encrypt_deterministic`: the Sunscreen SEAL fork's component-exporting encrypt, seeded with the SHA-512 digest,
 * WITHOUT the special modulus (the ciphertext is formed directly at the data level from the first two limbs of the
 * public key), built with SEAL_USE_GAUSSIAN_NOISE.  Pinned by the reference's own SHA-512 known answers
 * (fhe.rs:2101-2121, 2165-2185, 2224-2244; tests/test_oracle_kat.py reproduces all of them):
 *   prng   = Blake2xbPRNG(seed): 4096-byte buffers, buffer c = BLAKE2Xb(out 4096, in = c as LE u64, key = 64-byte seed)
 *   u      = sample_poly_ternary: std::uniform_int_distribution<uint64_t>(0, 2) on 32-bit draws (libstdc++ >= 11: Lemire)
 *   e0, e1 = sample_poly_normal: ClippedNormalDistribution(0, 3.2, 19.2) over libstdc++ std::normal_distribution<double>
 *            (Marsaglia polar on generate_canonical<double, 53> = two 32-bit draws, low word first; the second variate
 *            of a pair is kept for the next call; one distribution object per polynomial), truncated toward zero
 *   c_j    = INTT(NTT(u) * pk_j) + e_j  mod (q0, q1);  c_0 += round-scaled plaintext
 * BLAKE2b per RFC 7693; BLAKE2X per the BLAKE2X paper / reference blake2xb.c (node_offset = block index, xof_length in
 * the upper half of the node-offset field, expansion nodes with fanout = depth = 0, leaf_length = inner_length = 64). */
static const u64 B2B_IV[8] = {0x6A09E667F3BCC908ull, 0xBB67AE8584CAA73Bull, 0x3C6EF372FE94F82Bull, 0xA54FF53A5F1D36F1ull,
                              0x510E527FADE682D1ull, 0x9B05688C2B3E6C1Full, 0x1F83D9ABFB41BD6Bull, 0x5BE0CD19137E2179ull};
static const uint8_t B2B_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static inline u64 rotr64(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
/* one compression: h <- F(h, m, t, last) (RFC 7693 3.2; t < 2^64 here) */
static void b2b_compress(u64 h[8], const u64 m[16], u64 t, int last) {
    u64 v[16];
    for (int i = 0; i < 8; i++) v[i] = h[i], v[i + 8] = B2B_IV[i];
    v[12] ^= t;
    if (last) v[14] = ~v[14];
#define B2B_G(a, b, c, d, x, y)                                                                                  \
    v[a] = v[a] + v[b] + (x), v[d] = rotr64(v[d] ^ v[a], 32), v[c] = v[c] + v[d], v[b] = rotr64(v[b] ^ v[c], 24), \
    v[a] = v[a] + v[b] + (y), v[d] = rotr64(v[d] ^ v[a], 16), v[c] = v[c] + v[d], v[b] = rotr64(v[b] ^ v[c], 63)
    for (int r = 0; r < 12; r++) {
        const uint8_t *s = B2B_SIGMA[r];
        B2B_G(0, 4, 8, 12, m[s[0]], m[s[1]]);
        B2B_G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        B2B_G(2, 6, 10, 14, m[s[4]], m[s[5]]);
        B2B_G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        B2B_G(0, 5, 10, 15, m[s[8]], m[s[9]]);
        B2B_G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        B2B_G(2, 7, 8, 13, m[s[12]], m[s[13]]);
        B2B_G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
#undef B2B_G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
/* parameter block words 0..2 (the rest is zero): digest_length | key_length<<8 | fanout<<16 | depth<<24 | leaf_length<<32,
 * node_offset | xof_length<<32, node_depth | inner_length<<8 */
static void b2b_init(u64 h[8], unsigned digest, unsigned keylen, unsigned fanout, unsigned depth, u64 leaf, u64 node_offset,
                     u64 xof, unsigned inner) {
    for (int i = 0; i < 8; i++) h[i] = B2B_IV[i];
    h[0] ^= (u64)digest | ((u64)keylen << 8) | ((u64)fanout << 16) | ((u64)depth << 24) | (leaf << 32);
    h[1] ^= node_offset | (xof << 32);
    h[2] ^= (u64)inner << 8;
}
#define SEAL_PRNG_BUF 4096
/* buffer `counter` of the PRNG stream (4096 bytes, as 512 LE words) */
static void seal_prng_buffer(const u64 seed[8], u64 counter, u64 out[SEAL_PRNG_BUF / 8]) {
    u64 h[8], m[16] = {0};
    b2b_init(h, 64, 64, 1, 1, 0, 0, SEAL_PRNG_BUF, 0);
    for (int i = 0; i < 8; i++) m[i] = seed[i]; /* the key, zero padded to one block */
    b2b_compress(h, m, 128, 0);
    for (int i = 0; i < 16; i++) m[i] = 0;
    m[0] = counter;
    b2b_compress(h, m, 136, 1);
    u64 root[16] = {0};
    for (int i = 0; i < 8; i++) root[i] = h[i];
    for (u64 node = 0; node < SEAL_PRNG_BUF / 64; node++) {
        u64 g[8];
        b2b_init(g, 64, 0, 0, 0, 64, node, SEAL_PRNG_BUF, 64);
        b2b_compress(g, root, 64, 1);
        for (int i = 0; i < 8; i++) out[node * 8 + i] = g[i];
    }
}
void bfvo_seal_prng(const uint64_t seed[8], uint8_t *out, size_t nbytes) {
    u64 buf[SEAL_PRNG_BUF / 8];
    for (size_t off = 0, c = 0; off < nbytes; off += SEAL_PRNG_BUF, c++) {
        seal_prng_buffer(seed, c, buf);
        size_t k = nbytes - off < SEAL_PRNG_BUF ? nbytes - off : SEAL_PRNG_BUF;
        memcpy(out + off, buf, k); /* little-endian host */
    }
}
typedef struct {
    u64 seed[8], counter;
    u64 buf[SEAL_PRNG_BUF / 8];
    size_t pos; /* in 32-bit words */
    size_t drawn;
    const uint32_t *ext; /* test hook: draw from a caller-supplied stream instead (bfvo_seal_sample_stream) */
    size_t ext_n;
    int exhausted;
} SealPrng;
static void sprng_init(SealPrng *p, const u64 seed[8]) {
    memcpy(p->seed, seed, 64);
    p->counter = 0, p->drawn = 0, p->ext = NULL, p->ext_n = 0, p->exhausted = 0;
    seal_prng_buffer(p->seed, p->counter++, p->buf);
    p->pos = 0;
}
static inline uint32_t sprng_u32(SealPrng *p) {
    if (p->ext) {
        if (p->drawn >= p->ext_n) {
            p->exhausted = 1;
            return 0x9E3779B9u; /* any accepted value: lets the samplers terminate */
        }
        return p->ext[p->drawn++];
    }
    uint32_t v = (uint32_t)(p->buf[p->pos >> 1] >> (32 * (p->pos & 1)));
    p->pos++, p->drawn++;
    if (p->pos == SEAL_PRNG_BUF / 4) {
        seal_prng_buffer(p->seed, p->counter++, p->buf);
        p->pos = 0;
    }
    return v;
}
/* libstdc++ (GCC >= 11) uniform_int_distribution<uint64_t>(0, 2) over a URNG with range exactly 2^32: _S_nd */
static inline uint32_t uniform3(SealPrng *p) {
    u64 product = (u64)sprng_u32(p) * 3u;
    uint32_t low = (uint32_t)product;
    if (low < 3u) {
        const uint32_t threshold = (uint32_t)(0u - 3u) % 3u;
        while (low < threshold) {
            product = (u64)sprng_u32(p) * 3u;
            low = (uint32_t)product;
        }
    }
    return (uint32_t)(product >> 32);
}
/* libstdc++ generate_canonical<double, 53> over a 32-bit URNG: two draws, low word first */
static inline double canonical53(SealPrng *p) {
    double sum = (double)sprng_u32(p);
    sum += (double)sprng_u32(p) * 4294967296.0;
    double r = sum / 18446744073709551616.0;
    return r >= 1.0 ? nextafter(1.0, 0.0) : r;
}
static void sample_normal_poly(SealPrng *p, int8_t *out) {
    int have = 0;
    double saved = 0.0;
    for (size_t i = 0; i < N; i++) {
        double value;
        for (;;) {
            double ret;
            if (have) {
                have = 0, ret = saved;
            } else {
                double x, y, r2;
                do {
                    x = 2.0 * canonical53(p) - 1.0;
                    y = 2.0 * canonical53(p) - 1.0;
                    r2 = x * x + y * y;
                } while (r2 > 1.0 || r2 == 0.0);
                const double mult = sqrt(-2 * log(r2) / r2);
                saved = x * mult, have = 1;
                ret = y * mult;
            }
            value = ret * 3.2 + 0.0;
            if (fabs(value - 0.0) <= 19.2) break;
        }
        out[i] = (int8_t)(int64_t)value;
    }
}
size_t bfvo_seal_sample(const uint64_t seed[8], int8_t *u, int8_t *e0, int8_t *e1) {
    SealPrng *p = (SealPrng *)malloc(sizeof(SealPrng));
    sprng_init(p, seed);
    for (size_t i = 0; i < N; i++) u[i] = (int8_t)uniform3(p) - 1;
    sample_normal_poly(p, e0);
    sample_normal_poly(p, e1);
    size_t drawn = p->drawn;
    free(p);
    return drawn;
}
/* the same samplers on a caller-supplied stream of 32-bit draws (to test the rare paths: a zero draw in u, a clipped
 * variate); returns the draws consumed, 0 if the stream ran out */
size_t bfvo_seal_sample_stream(const uint32_t *words, size_t nwords, int8_t *u, int8_t *e0, int8_t *e1) {
    SealPrng *p = (SealPrng *)calloc(1, sizeof(SealPrng));
    p->ext = words, p->ext_n = nwords;
    for (size_t i = 0; i < N; i++) u[i] = (int8_t)uniform3(p) - 1;
    sample_normal_poly(p, e0);
    sample_normal_poly(p, e1);
    size_t drawn = p->exhausted ? 0 : p->drawn;
    free(p);
    return drawn;
}
/* encrypt_zero_asymmetric at the DATA level (first two limbs of each public-key polynomial; no modulus switching) +
 * multiply_add_plain_with_scaling_variant */
void bfvo_encrypt_samples_data_level(const u64 *pk, const u64 *plain, size_t len, const int8_t *u, const int8_t *e0, const int8_t *e1,
                                     u64 *ct) {
    bfvo_init();
    u64 *un = (u64 *)malloc((size_t)2 * N * 8);
    for (int J = 0; J < 2; J++) {
        for (size_t i = 0; i < N; i++) un[(size_t)J * N + i] = u[i] < 0 ? C.mod[J].q - 1 : (u64)u[i];
        ntt_fwd(un + (size_t)J * N, J);
    }
    for (int j = 0; j < 2; j++) {
        const int8_t *e = j == 0 ? e0 : e1;
        for (int J = 0; J < 2; J++) {
            const Mod *m = &C.mod[J];
            u64 *x = ct + ((size_t)j * 2 + J) * N;
            const u64 *k = pk + ((size_t)j * 3 + J) * N;
            for (size_t i = 0; i < N; i++) x[i] = mulmod(k[i], un[(size_t)J * N + i], m);
            ntt_inv(x, J);
            for (size_t i = 0; i < N; i++) x[i] = e[i] < 0 ? submod(x[i], (u64)(-e[i]), m) : addmod(x[i], (u64)e[i], m);
        }
    }
    plain_scaled(ct, plain, len, 0);
    free(un);
}
void bfvo_seal_encrypt(const u64 *pk, const u64 *plain, size_t len, const uint64_t seed[8], u64 *ct) {
    int8_t *smp = (int8_t *)malloc((size_t)3 * N);
    bfvo_seal_sample(seed, smp, smp + N, smp + 2 * N);
    bfvo_encrypt_samples_data_level(pk, plain, len, smp, smp + N, smp + 2 * N, ct);
    free(smp);
}

/* SEAL Decryptor::bfv_decrypt: dot product with powers of s, then exact round(t*x/q) mod t
 * (SEAL's decrypt_scale_and_round yields the same value whenever the noise budget is > 0). */
int bfvo_decrypt(const u64 *ct, size_t npolys, const u64 *sk, u64 *plain_out) {
    bfvo_init();
    u64 *x = (u64 *)malloc((size_t)2 * N * 8);
    u64 *t = (u64 *)malloc(N * 8);
    for (int l = 0; l < 2; l++) {
        const Mod *m = &C.mod[l];
        const u64 *s = sk + (size_t)l * N;
        u64 *acc = x + (size_t)l * N;
        memset(acc, 0, N * 8);
        /* Horner in NTT domain: ((c_{k-1} * s + c_{k-2}) * s + ...) */
        for (size_t p = npolys; p-- > 0;) {
            memcpy(t, ct + (p * 2 + l) * N, N * 8);
            ntt_fwd(t, l);
            for (size_t i = 0; i < N; i++) acc[i] = addmod(mulmod(acc[i], s[i], m), t[i], m);
        }
        ntt_inv(acc, l);
    }
    const u64 q0 = C.mod[0].q, q1 = C.mod[1].q;
    const u128 q = (u128)q0 * q1;
    u128 max_noise = 0;
    for (size_t i = 0; i < N; i++) {
        u64 v0 = mulmod(x[i], C.inv_q1_mod_q0, &C.mod[0]);
        u64 v1 = mulmod(x[N + i], C.inv_q0_mod_q1, &C.mod[1]);
        u128 X = (u128)v0 * q1 + (u128)v1 * q0;
        if (X >= q) X -= q;
        u128 tx = X * T;
        u128 mm = (tx + (q >> 1)) / q;
        plain_out[i] = (u64)(mm % T);
        u128 noise = tx % q;
        if (noise > (q >> 1)) noise = q - noise;
        if (noise > max_noise) max_noise = noise;
    }
    free(x);
    free(t);
    int nb = 0;
    while (max_noise) {
        nb++;
        max_noise >>= 1;
    }
    /* SEAL invariant_noise_budget = bit_count(q) - bit_count(noise) - 1 */
    return 72 - nb - 1;
}

/* ------------------------------------------------------------------ batched timing helpers */
typedef struct {
    const u64 *a, *b, *rk;
    u64 *out;
    size_t lo, hi;
    int mod, inverse;
    u64 *limbs;
} Job;
static void *job_mul(void *p) {
    Job *j = (Job *)p;
    for (size_t i = j->lo; i < j->hi; i++)
        bfvo_mul_relin(j->a + i * 4 * N, j->b + i * 4 * N, j->rk, j->out + i * 4 * N);
    return NULL;
}
static void *job_ntt(void *p) {
    Job *j = (Job *)p;
    for (size_t i = j->lo; i < j->hi; i++) {
        if (j->inverse)
            ntt_inv(j->limbs + i * N, j->mod);
        else
            ntt_fwd(j->limbs + i * N, j->mod);
    }
    return NULL;
}
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static double run_jobs(void *(*fn)(void *), Job *proto, size_t n, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = (int)(n ? n : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    Job *jobs = (Job *)malloc(sizeof(Job) * (size_t)threads);
    double t0 = now_s();
    for (int t = 0; t < threads; t++) {
        jobs[t] = *proto;
        jobs[t].lo = n * (size_t)t / (size_t)threads;
        jobs[t].hi = n * (size_t)(t + 1) / (size_t)threads;
        pthread_create(&th[t], NULL, fn, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    double t1 = now_s();
    free(th);
    free(jobs);
    return t1 - t0;
}
double bfvo_batch_mul_relin(const u64 *a, const u64 *b, const u64 *rk, u64 *out, size_t n, int threads) {
    bfvo_init();
    Job j;
    memset(&j, 0, sizeof j);
    j.a = a;
    j.b = b;
    j.rk = rk;
    j.out = out;
    return run_jobs(job_mul, &j, n, threads);
}
double bfvo_batch_ntt(u64 *limbs, size_t n_limbs, int mod, int inverse, int threads) {
    bfvo_init();
    Job j;
    memset(&j, 0, sizeof j);
    j.limbs = limbs;
    j.mod = mod;
    j.inverse = inverse;
    return run_jobs(job_ntt, &j, n_limbs, threads);
}
