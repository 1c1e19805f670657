// modarith.cuh -- 64-bit modular arithmetic for RNS residues (q < 2^61), sm_100a.
//
// B200 has no 64-bit integer multiplier: mul.hi.u64 / mul.lo.u64 lower to chains of 32-bit
// IMAD.WIDE on the fma pipe, so every routine below is written to minimise the number of
// 64x64 products:  Shoup (1 mulhi + 2 mullo) wherever one operand is a constant (twiddles,
// q_last^{-1}, N^{-1}), shift-Barrett (2 mulhi + 2 mullo) for data x data products.
//
// All functions return the same canonical value SEAL's uintarithsmallmod.h routines return;
// how the value is reached is free.
#pragma once
#include <stdint.h>

#ifdef B200HE_EMU
#include "emu/cuda_shim.h"
#else
#include <cuda_runtime.h>
#endif

namespace b200he {

typedef uint64_t u64;
typedef uint32_t u32;

// Per-modulus constants, resident in device memory (one entry per prime: key-level chain,
// then BEHZ auxiliary primes, then the plain modulus where needed).
struct Mod {
    u64 q;        // the prime
    u64 two_q;    // 2q
    u64 nq;       // 2^64 - q   ("+ k*nq" subtracts k*q in 64-bit wrap-around arithmetic)
    u64 mu;       // floor(2^(62+bits) / q)  -- shift-Barrett constant, < 2^63
    u64 r64;      // floor(2^64 / q)         -- single-word Barrett (SEAL const_ratio[1])
    u64 ninv;     // N^{-1} mod q
    u64 ninv_s;   // Shoup quotient of ninv
    u32 bits;     // bit length of q
    u32 sh;       // bits - 2
    // Lazy-reduction schedule of the CTA-local transforms (built on the host from the pass schedule and
    // floor(2^64/q), see build_tables): bit P = reduce all coefficients to [0,2q) before pass P,
    // bit 8+P = reduce again after the first K/2 stages of pass P.
    u32 fwd_mask, inv_mask;
    u32 inv_c[4];   // inverse: bound multiplier C (values < C*q) at the start of pass P
    u32 acc_period; // key-switch inner product: reduce the lazy accumulators every acc_period digits
    u32 pad_;
};

__host__ __device__ __forceinline__ u64 mulhi64(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

// x - q if x >= q   (x < 2q)
__host__ __device__ __forceinline__ u64 csub(u64 x, u64 q) { return x >= q ? x - q : x; }

__host__ __device__ __forceinline__ u64 add_mod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
__host__ __device__ __forceinline__ u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }

// Shoup product w*y mod q in [0,2q), valid for ANY 64-bit y (w < q, ws = floor(w 2^64 / q)).
__host__ __device__ __forceinline__ u64 shoup_lazy(u64 y, u64 w, u64 ws, u64 q)
{
    return y * w - mulhi64(y, ws) * q;
}
__host__ __device__ __forceinline__ u64 shoup(u64 y, u64 w, u64 ws, u64 q) { return csub(shoup_lazy(y, w, ws, q), q); }

// w*y (+ x) with the Shoup quotient: x + (w*y mod q) + {0,q}, all in 64-bit wrap-around arithmetic.
// 10 integer multiply-adds: 4 for the high product, 6 for the two low products (the subtraction of
// qhat*q is folded into the multiply-add chain through nq = 2^64 - q).
__host__ __device__ __forceinline__ u64 shoup_mad(u64 y, u64 w, u64 ws, u64 nq, u64 x)
{
    return y * w + mulhi64(y, ws) * nq + x;
}
// any 64-bit x -> x mod q + {0,q}  (in [0,2q)).  NARROW: floor(2^64/q) fits 32 bits (q > 2^32), then
// the quotient costs 2 multiply-adds instead of 4.  Callers branch once per CTA on Mod::bits.
template <bool NARROW> __host__ __device__ __forceinline__ u64 reduce_lazy_t(u64 x, const Mod &m)
{
    if (NARROW) {
        const u32 r = (u32)m.r64, x0 = (u32)x, x1 = (u32)(x >> 32);
        const u64 t = (u64)x0 * r;
        const u64 u = (u64)x1 * r + (t >> 32);
        return x + (u >> 32) * m.nq;
    }
    return x + mulhi64(x, m.r64) * m.nq;
}
__host__ __device__ __forceinline__ u64 reduce_lazy(u64 x, const Mod &m) { return m.bits > 32 ? reduce_lazy_t<true>(x, m) : reduce_lazy_t<false>(x, m); }
// any 64-bit x -> canonical x mod q
__host__ __device__ __forceinline__ u64 reduce_full(u64 x, const Mod &m) { return csub(reduce_lazy(x, m), m.q); }

// x mod q for any 64-bit x (SEAL barrett_reduce_64)
__host__ __device__ __forceinline__ u64 reduce64(u64 x, const Mod &m)
{
    u64 r = x - mulhi64(x, m.r64) * m.q;
    return csub(r, m.q);
}

// (hi:lo) mod q for hi:lo < 2^(2*bits): shift-Barrett.  qe >= floor(x/q) - 2.
__host__ __device__ __forceinline__ u64 reduce128(u64 hi, u64 lo, const Mod &m)
{
    u64 xh = (hi << (64 - m.sh)) | (lo >> m.sh);
    u64 r = lo - mulhi64(xh, m.mu) * m.q;
    r = csub(r, m.two_q);
    return csub(r, m.q);
}

// a*b mod q, a,b < q
__host__ __device__ __forceinline__ u64 mul_mod(u64 a, u64 b, const Mod &m)
{
    return reduce128(mulhi64(a, b), a * b, m);
}
// (a*b + c) mod q, a,b,c < q
__host__ __device__ __forceinline__ u64 mad_mod(u64 a, u64 b, u64 c, const Mod &m)
{
    u64 lo = a * b, hi = mulhi64(a, b);
    lo += c;
    hi += (lo < c);
    return reduce128(hi, lo, m);
}

// Harvey lazy Cooley-Tukey butterfly: x,y in [0,4q) -> [0,4q)
__host__ __device__ __forceinline__ void ct_bfly(u64 &x, u64 &y, u64 w, u64 ws, u64 q, u64 two_q)
{
    u64 u = csub(x, two_q);
    u64 v = shoup_lazy(y, w, ws, q);
    x = u + v;
    y = u - v + two_q;
}
// Gentleman-Sande lazy butterfly: x,y in [0,2q) -> [0,2q)
__host__ __device__ __forceinline__ void gs_bfly(u64 &x, u64 &y, u64 w, u64 ws, u64 q, u64 two_q)
{
    u64 u = x, v = y;
    x = csub(u + v, two_q);
    y = shoup_lazy(u - v + two_q, w, ws, q);
}

}   // namespace b200he
