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

    // ---- BEHZ base extension q -> Bsk via m_tilde (RNSTool::fastbconv_m_tilde + sm_mrq)
    Shoup ext_in[2];              // m_tilde * (q/q_l)^-1 mod q_l
    u32 punct_q_mod_mtilde[2];    // (q/q_l) mod 2^32
    u32 neg_inv_q_mod_mtilde;     // -q^-1 mod 2^32
    u64 extA[3], extB[3], extC[3];  // per Bsk prime k: (q/q_0)*m~^-1, (q/q_1)*m~^-1, q*m~^-1 (mod p_k)

    // ---- fast_floor (input already multiplied by t)
    Shoup inv_punct_q[2];        // (q/q_l)^-1 mod q_l
    u64 flV[3], flA[3], flB[3];  // f_k = v_k*flV + tmp0*flA + tmp1*flB mod p_k
                                 //   flV = q^-1, flA = -(q/q_0) q^-1, flB = -(q/q_1) q^-1

    // merged forms used by k_floor_sk (one reduction per output instead of one per SEAL step):
    //   tb_j  = [f_j * (B/b_j)^-1]_{b_j}      = v_bj*skV[j] + tmp0*skA[j] + tmp1*skB[j]            (j = 0,1)
    //   alpha = [(h - f_msk) * B^-1]_{m_sk}    = tb0*alK[0] + tb1*alK[1] + v_msk*alK[2] + tmp0*alK[3] + tmp1*alK[4]
    u64 skV[2], skA[2], skB[2];
    u64 alK[5];

    // Shoup quotients of the merged constants (k_ext_conv / k_floor_sk evaluate each output as one ShoupSum)
    u64 extAs[3], extBs[3], extCs[3];
    u64 extNeg[3];  // (p_k - m~) * extC[k] mod p_k: the correction term when the m~ residue is negative
    u64 skVs[2], skAs[2], skBs[2];
    u64 alKs[5];
    u32 q_w[3];           // q = q0*q1 as three 32-bit words (q_w[2] < 2^8)
    u32 skD[2];           // m_sk - b_j > 0: with every Bsk prime 2^61 - c, b_{j} == -skD[j] (mod m_sk)
    Shoup nib;            // -(B^-1) mod m_sk
    Shoup pBq[2][2];      // punct_B_mod_q [j][l]
    Shoup Bq[2], nBq[2];  // B_mod_q, neg_B_mod_q

    // ---- fastbconv_sk
    Shoup inv_punct_B[2];      // (B/b_j)^-1 mod b_j
    u64 punct_B_mod_q[2][2];   // [j][l]
    u64 punct_B_mod_msk[2];
    Shoup inv_B_mod_msk;
    u64 B_mod_q[2];      // prod(B) mod q_l
    u64 neg_B_mod_q[2];  // q_l - prod(B) mod q_l

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
    ulonglong2 f[kNumMod][64];
    ulonglong2 i[kNumMod][64];
};

struct DevTables {
    const ulonglong2 *twf[kNumMod];  // [k] = (rp[k], rp[k] Shoup), rp[bitrev(i)] = psi^i
    const ulonglong2 *twi[kNumMod];  // [k] = (rp[k]^-1, Shoup)
};

}  // namespace fheb
