// Constants shared by the host context (fills them, context.cpp) and the kernels (read them from
// __constant__ memory).  Everything here is derived from testnet.rs:8-14 by the rules of SEAL 4.0's
// SEALContext / RNSTool::initialize (restated in context.cpp).  "S" suffix = Shoup quotient
// floor(x * 2^64 / modulus) of the constant next to it.
#pragma once
#include <cuda_runtime.h>

#include "params.h"

namespace fheb {

struct Shoup {
    u64 w, ws;
};

struct DevConsts {
    // inverse-NTT output scalings per modulus
    Shoup ninv[kNumMod];    // N^-1
    Shoup ninv_t[kNumMod];  // N^-1 * t   (BEHZ step 6 folded into the inverse transform)
    // the same scalings pre-multiplied by the last inverse stage's twiddle (table index 1), so that the scaling
    // costs one multiplication per butterfly of the last stage instead of one per coefficient
    Shoup ninv_w[kNumMod];
    Shoup ninv_t_w[kNumMod];

    // ---- BEHZ base extension q -> Bsk via m_tilde (RNSTool::fastbconv_m_tilde + sm_mrq), integer domain (k_ext_conv)
    Shoup ext_in[2];            // m_tilde * (q/q_l)^-1 mod q_l
    u32 neg_inv_q_mod_mtilde;   // -q^-1 mod 2^32
    u32 q_w[3];                 // q = q0*q1 as three 32-bit words (q_w[2] < 2^8)
    u64 extNeg[3];              // -q mod p_k: added when the centred m~ residue is negative

    // ---- fast_floor + fastbconv_sk (input already multiplied by t), constants of SEAL's steps merged (k_floor_sk):
    //   t_l   = v_l (q/q_l)^-1 mod q_l,   y0 = t0 q1 + t1 q0 (integer)
    //   tb_j  = [(v_bj - y0) skV[j]]_{b_j},            skV[j] = q^-1 (B/b_j)^-1 mod b_j
    //   alpha = [w nib + (v_msk - y0) alK2]_{m_sk},    w = tb0 skD[1] + tb1 skD[0],  nib = -(B^-1),  alK2 = -(q^-1 B^-1)
    //   out_l = [tb0 pBq[0][l] + tb1 pBq[1][l] -/+ |alpha| Bq[l]]_{q_l}
    Shoup inv_punct_q[2];  // (q/q_l)^-1 mod q_l
    u64 skV[2], skVs[2];   // value and Shoup quotient
    u64 alK2, alK2s;
    u32 skD[2];            // m_sk - b_j > 0: with every Bsk prime 2^61 - c, b_j == -skD[j] (mod m_sk)
    Shoup nib;
    Shoup pBq[2][2];       // (B/b_j) mod q_l   [j][l]
    Shoup Bq[2], nBq[2];   // prod(B) mod q_l and its negation
    // ---- the q-limbs of the tensor product recovered from its Bsk limbs (k_floor_sk): the product of
    // the extended operands is an INTEGER polynomial with |t D| < 2^166 < Bsk/2 (Bsk = b0 b1 m_sk ~ 2^183), so its residues
    // mod q_l -- which SEAL obtains from 14 more transforms on the q-limbs -- follow exactly from the three Bsk residues:
    //   y_i = [r_i (Bsk/p_i)^-1]_{p_i},  v = round(sum y_i / p_i),  t D = sum y_i (Bsk/p_i) - v Bsk
    //   t_l = [t D (q/q_l)^-1]_{q_l} = [sum y_i crtK[i][l] + v crtNB[l]]_{q_l}        (inv_punct_q merged in)
    Shoup crt3[3];         // (Bsk/p_i)^-1 mod p_i
    Shoup crtK[3][2];      // (Bsk/p_i) (q/q_l)^-1 mod q_l
    Shoup crtNB[2];        // -Bsk (q/q_l)^-1 mod q_l

    // ---- BEHZ multiply on the dual base s_0..s_5 (params.h: kDualPrime; kernels k_*_d).  S = prod s_i, S4 = s_0 s_1 s_2 s_3.
    //   extension   : a' = base + m 2^61 (- q)  ->  [base_lo + base_hi R32 + m R61 (+ NQ)]_{s_i}, Barrett from below 2^61
    //   after INTT  : r_i = t D mod s_i.   y_i = [r_i C_i]_{s_i},  v = round(sum y_i / s_i)  (|t D| / S < 2^-13)
    //   t_l         = [sum y_i K[i][l] + v KN[l]]_{q_l}            (= t D (q/q_l)^-1 mod q_l, the quantity fast_floor starts from)
    //   y0          = t0 q1 + t1 q0;   tb_i = [(r_i - y0) W_i]_{s_i}, i < 4,   v' = round(sum tb_i / s_i)  (|f| / S4 < 2^-24)
    //   out_l       = [sum tb_i P[i][l] + v' NS4[l]]_{q_l}         (= f mod q_l,  f = (t D - y0) / q)
    u32 d_R32[6], d_R61[6], d_NQ[6], d_mu61[6];  // 2^32 mod s, 2^61 mod s, -q mod s, floor(2^61 / s)
    Shoup d_ninv_t[3], d_ninv_t_w[3];            // N^-1 t and N^-1 t w_last per dual limb: both lanes packed; ws = 32-bit Shoup quotients
    u32 d_C[6], d_Cs[6];                         // (S/s_i)^-1 mod s_i and floor(. 2^32 / s_i)
    u32 d_R48[6];                                // floor(2^48 / s_i): y_i / s_i to 16 bits
    Shoup d_K[6][2];                             // (S/s_i) (q/q_l)^-1 mod q_l
    u64 d_KN[2];                                 // -S (q/q_l)^-1 mod q_l
    u32 d_R58[4];                                // 2^58 mod s_i
    u32 d_W[4], d_Ws[4];                         // q^-1 (S4/s_i)^-1 mod s_i
    Shoup d_P[4][2];                             // (S4/s_i) mod q_l
    u64 d_NS4[2];                                // -S4 mod q_l

    // ---- key switch on the dual base (k_*_ksd kernels): U_k = sum_j d_j * RK_jk over Z, with the key lifted to integers
    //   RK = sum_m [rk_m (Q/m)^-1]_m (Q/m) (Q = q0 q1 P, RK < 3 Q, == rk mod every m), |U_k| < 2^161 << S / 2.
    //   lift: y_m = [coef_m lk_C[m]]_m,  RK mod s_i = sum_m (y_m mod s_i) lk_R[m][i]
    //   finish: y_i = [r_i C_i]_{s_i}, v = round(sum y_i / s_i), U mod m = [sum y_i ksK[i][m] + v ksKN[m]]_m for m = q0, q1, P
    u32 opaque_zero;               // 0, but unknown to ptxas: a third addend that keeps two-input adds off the multiplier pipe (IMAD.IADD)
    Shoup d_ninv[3], d_ninv_w[3];  // N^-1 and N^-1 w_last per dual limb (lanes packed, 32-bit Shoup quotients)
    Shoup lk_C[3];                 // (Q/m)^-1 mod m
    u32 lk_R[3][6], lk_Rs[3][6];   // (Q/m) mod s_i and its 32-bit Shoup quotient
    Shoup ksK[6][3];               // (S/s_i) mod m,  m = q0, q1, P
    u64 ksKN[3];                   // -S mod m
    // the rounded division by P folded into the sums for q0, q1: out_l = [c_l + sum y_i ksKd[i][l] + v ksKNd[l] + last_lo ksNd[l]
    //   + last_hi ksNd30[l] + ksHd[l]]_{q_l} with last = (U mod P + P/2) mod P = last_lo + 2^30 last_hi
    u64 ksKd[6][2];                // (S/s_i) P^-1 mod q_l
    u64 ksKNd[2];                  // -S P^-1 mod q_l
    u64 ksNd[2], ksNd30[2];        // -P^-1 mod q_l and -2^30 P^-1 mod q_l
    u64 ksHd[2];                   // (P/2 mod q_l) P^-1 mod q_l

    // ---- key switching (switch_key_inplace, BFV branch)
    Shoup inv_P_mod_q[2];
    u64 half_P;
    u64 half_P_mod_q[2];

    // ---- decryption (Decryptor::bfv_decrypt): CRT recombination and exact round(t*x/q)
    Shoup crt_inv[2];     // (q/q_l)^-1 mod q_l  (= inv_punct_q)
    u64 q_lo, q_hi;       // q = q0*q1 as a 128-bit integer
    u64 qhalf_lo, qhalf_hi;  // (q-1)/2

    // ---- plain ops (multiply_add_plain_with_scaling_variant, multiply_plain_normal)
    u64 delta_mod_q[2];       // floor(q/t) mod q_l
    u64 q_mod_t;              // q mod t
    u64 upper_half_threshold; // (t+1)>>1
    u64 upper_half_incr[2];   // q_l - t
};

// twiddles of the first six stages (table indices 1..63), read through the constant cache: they are
// CTA-uniform (first pass) or warp-uniform (second pass)
struct DevTwLow {
    ulonglong2 f[kNumTab][64];
    ulonglong2 i[kNumTab][64];
};

struct DevTables {
    const ulonglong2 *twf[kNumTab];  // [k] = (rp[k], rp[k] Shoup), rp[bitrev(i)] = psi^i
    const ulonglong2 *twi[kNumTab];  // [k] = (rp[k]^-1, Shoup)
};

}  // namespace fheb
