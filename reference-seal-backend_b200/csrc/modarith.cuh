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
    u32 dp;         // 1: the transforms of this modulus run in the FP64 domain (q < 2^DP_MAX_BITS, see below)
    u32 pm_c;       // q = 2^bits - pm_c with pm_c small enough for reduce_pm (0: not of that form)
    u32 pad_;
    double dq, dnq, dqinv, dqinv_up;   // q, -q, RN(1/q), 1/q rounded up
    double dninv, dninv_q;             // N^{-1} mod q and RN(N^{-1}/q)
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
// q = 2^b - c (every prime SEAL's CoeffModulus::Create picks is the largest few below a power of two, c ~ 2^14..2^22):
// x = xh 2^b + xl = xl + xh c (mod q), one shift, one mask and one IMAD.WIDE instead of the 3 wide products of the
// Barrett quotient.  The host sets Mod::pm_c only when (2^(64-b) - 1) c + 2^b - 1 < 2q, so the result is in [0, 2q).
__host__ __device__ __forceinline__ u64 reduce_pm(u64 x, u32 b, u32 c)
{
    return (x & ((u64(1) << b) - 1)) + (u64)(u32)(x >> b) * c;
}
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
__host__ __device__ __forceinline__ u64 reduce_lazy(u64 x, const Mod &m)
{
    if (m.pm_c) return reduce_pm(x, m.bits, m.pm_c);
    return m.bits > 32 ? reduce_lazy_t<true>(x, m) : reduce_lazy_t<false>(x, m);
}
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

// ------------------------------------------------------------------------------------ FP64 domain
// B200 (sm_100a) keeps a full-rate FP64 unit: DFMA issues at 64 lanes/clk/SM, twice the rate of IMAD.WIDE, and one
// DFMA multiplies 53 x 53 bits where IMAD.WIDE multiplies 32 x 32 (tools/fp64_peak.cu, profiles/r1i_fp64_peak.json:
// a complete butterfly costs 19 cycles per warp in FP64 against 37 on the integer pipe).  For primes below
// 2^DP_MAX_BITS the transforms therefore compute on exact integers held in doubles, in a symmetric lazy
// representation (any integer congruent to the residue, magnitude < 2^51):
//   * w*y: (h, l) = two-product by FMA (h + l = w*y exactly), quotient k = rint(y * RN(w/q)) through the 1.5*2^52
//     rounding constant, result (h - k q) + l -- both steps exact -- an integer of magnitude <= 0.75 q;
//   * sums and differences of such values are exact as long as they stay below 2^53;
//   * leaving the domain: floor division by q with a directed-rounding FMA (1/q rounded up, operand made positive),
//     which is exact, then the integer is read from the mantissa.
// Every operation is exact integer arithmetic, so results are bit-identical to the integer path.
#define B200HE_DP_MAX_BITS 46
#define B200HE_DP_MAGIC 6755399441055744.0   /* 1.5 * 2^52 */
#define B200HE_DP_TWO52 4503599627370496.0
__device__ __forceinline__ double as_d(u64 v) { return __longlong_as_double((long long)v); }
__device__ __forceinline__ u64 as_u(double d) { return (u64)__double_as_longlong(d); }
// integer < 2^52 -> double (exact)
__device__ __forceinline__ double dp_from(u64 x) { return __dadd_rn(as_d(x | 0x4330000000000000ull), -B200HE_DP_TWO52); }
// w*y mod q + e q, |result| <= 0.75 q, for |y| < 2^51; wq = RN(w/q), nq = -q
__device__ __forceinline__ double dp_mul(double y, double w, double wq, double nq)
{
    const double h = __dmul_rn(y, w);
    const double l = __fma_rn(y, w, -h);
    const double k = __dadd_rn(__fma_rn(y, wq, B200HE_DP_MAGIC), -B200HE_DP_MAGIC);
    return __dadd_rn(__fma_rn(k, nq, h), l);
}
// x - rint(x/q) q: magnitude <= q/2 + 1, for |x| < 2^51
__device__ __forceinline__ double dp_reduce(double x, double qinv, double nq)
{
    return __fma_rn(__dadd_rn(__fma_rn(x, qinv, B200HE_DP_MAGIC), -B200HE_DP_MAGIC), nq, x);
}
// data x data product a*b mod q + e q for |a|, |b| < 2^46 (no precomputed quotient): |result| <= 0.52 q
__device__ __forceinline__ double dp_mul_dd(double a, double b, double qinv, double nq)
{
    const double h = __dmul_rn(a, b);
    const double l = __fma_rn(a, b, -h);
    const double k = __dadd_rn(__fma_rn(h, qinv, B200HE_DP_MAGIC), -B200HE_DP_MAGIC);
    return __dadd_rn(__fma_rn(k, nq, h), l);
}
// lazy value (|x| <= 16 q) -> canonical residue in [0, q) as an integer
__device__ __forceinline__ u64 dp_canon(double x, const Mod &m)
{
    const double a = __fma_rn(16.0, m.dq, x);                                           // > 0, exact
    const double k = __dadd_rn(__fma_rd(a, m.dqinv_up, B200HE_DP_MAGIC), -B200HE_DP_MAGIC);   // floor(a / q)
    return as_u(__fma_rn(k, m.dnq, __dadd_rn(a, B200HE_DP_TWO52))) & 0x000fffffffffffffull;
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
