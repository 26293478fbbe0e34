// 64-bit modular arithmetic on the sm_100a integer pipe.
// Replaces SEAL util/uintarithsmallmod.h (barrett_reduce_64/128, MultiplyUIntModOperand)
// on the reference's hot path (FheApp::run, /root/reference/src/fhe.rs:138-152).
//
// Two prime classes, selected at compile time by Mod<MI>::kSmall:
//  * 36/37-bit data and key-switch primes (q0, q1, P): 27 bits of headroom in a 64-bit word, so
//    butterflies never conditionally subtract; values drift up to ~2^50 and are reduced once.
//  * 61-bit BEHZ primes (b0, b1, m_sk): Harvey lazy ranges [0,4q) forward / [0,2q) inverse.
#pragma once
#include "params.h"

namespace fheb {

template <int MI>
struct Mod {
    static constexpr int kIndex = MI;
    static constexpr u64 q = kModulus[MI];
    static constexpr u64 two_q = 2 * kModulus[MI];
    static constexpr u64 r1 = barrett_ratio(kModulus[MI]).hi;  // floor(2^128/q) >> 64 == floor(2^64/q)
    static constexpr u64 r0 = barrett_ratio(kModulus[MI]).lo;
    static constexpr bool kSmall = kModulus[MI] < (1ull << 40);
};

__device__ __forceinline__ u64 mulhi64(u64 a, u64 b) { return __umul64hi(a, b); }

// (hi:lo) += a * b
__device__ __forceinline__ void mac128(u64 &lo, u64 &hi, u64 a, u64 b) {
    asm("mad.lo.cc.u64 %0, %2, %3, %0;\n\t"
        "madc.hi.u64 %1, %2, %3, %1;"
        : "+l"(lo), "+l"(hi)
        : "l"(a), "l"(b));
}

template <class M>
__device__ __forceinline__ u64 csub(u64 x, u64 m) {
    return x >= m ? x - m : x;
}

// Shoup multiplication by a precomputed (w, ws = floor(w * 2^64 / q)); any 64-bit x; result in [0, 2q)
template <class M>
__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 w, u64 ws) {
    return x * w - mulhi64(x, ws) * M::q;
}
template <class M>
__device__ __forceinline__ u64 shoup(u64 x, u64 w, u64 ws) {
    return csub<M>(shoup_lazy<M>(x, w, ws), M::q);
}

// x mod q for any 64-bit x (SEAL barrett_reduce_64)
template <class M>
__device__ __forceinline__ u64 reduce64(u64 x) {
    u64 r = x - mulhi64(x, M::r1) * M::q;
    return csub<M>(r, M::q);
}

// (hi:lo) mod q, requires (hi:lo) < q * 2^64 (SEAL barrett_reduce_128)
template <class M>
__device__ __forceinline__ u64 reduce128(u64 hi, u64 lo) {
    // floor(z * ratio / 2^128), exact nested floors
    u64 carry = mulhi64(lo, M::r0);
    u64 t_lo = lo * M::r1;
    u64 t_hi = mulhi64(lo, M::r1);
    u64 s = t_lo + carry;
    u64 tmp3 = t_hi + (s < carry);
    u64 u_lo = hi * M::r0;
    u64 u_hi = mulhi64(hi, M::r0);
    u64 s2 = s + u_lo;
    u64 c2 = u_hi + (s2 < u_lo);
    u64 quo = hi * M::r1 + tmp3 + c2;
    u64 r = lo - quo * M::q;
    return csub<M>(r, M::q);
}

template <class M>
__device__ __forceinline__ u64 mulmod(u64 a, u64 b) {
    return reduce128<M>(mulhi64(a, b), a * b);
}

template <class M>
__device__ __forceinline__ u64 addmod(u64 a, u64 b) {
    return csub<M>(a + b, M::q);
}
template <class M>
__device__ __forceinline__ u64 submod(u64 a, u64 b) {
    return a >= b ? a - b : a + M::q - b;
}

}  // namespace fheb
