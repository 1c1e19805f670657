// ntt_core.cuh -- CTA-level negacyclic NTT over one RNS limb held in shared memory.
//
// Replaces SEAL util/ntt.cpp ntt_negacyclic_harvey / inverse_ntt_negacyclic_harvey
// (SURVEY.md §2.2 K1/K2, Appendix A.3-A.4): forward = Cooley-Tukey, natural -> bit-reversed,
// twiddle table tw[k] = psi^{brv(k)}; inverse = Gentleman-Sande, bit-reversed -> natural,
// itw[k] = tw[k]^{-1}, N^{-1} folded into the final stage.
//
// B200 mapping.  One CTA of T = N/16 threads owns one limb (N <= 16384; 8 B/coeff =
// 64-128 KiB of the SM's 227 KiB shared memory).  Every thread keeps 16 coefficients in
// registers and runs a radix-2^k pass (k in {2,3,4}) on them; between passes the limb is
// exchanged through shared memory.  log2(N) stages therefore cost only 2-3 shared-memory
// round trips, and the first pass reads / the last pass writes global memory directly,
// so a limb crosses HBM exactly once in each direction.
//   * strided passes with k<=3 hold C = 16>>k adjacent coefficient columns per thread, so
//     every global and shared access is a 128-bit vector and a warp's accesses are
//     contiguous (fully coalesced LDG.128/STG.128);
//   * the index swizzle  i ^ (((i>>4)&7)<<1)  (the TMA SWIZZLE_128B pattern on 8-byte
//     elements) makes every pass bank-conflict free: 16-coefficient-contiguous passes use
//     128-bit accesses whose 16 B chunk index is XORed with the row, strided passes touch
//     one aligned 128 B row per quarter/half warp;
//   * twiddles are (w, floor(w 2^64/q)) pairs read with one 128-bit ld.global.nc each and
//     shared by the C columns of a thread; the table is shared by the whole batch and
//     lives in L2.
// Warp shuffles are deliberately NOT used for the inner stages: a 64-bit exchange costs
// 2 SHFL per coefficient per stage, while the shared-memory round trip costs one
// LDS.128 + one STS.128 per coefficient PAIR per 3-4 stages.
#pragma once
#include "modarith.cuh"

namespace b200he {

// Forward pass schedule per log2(N): stages per pass.  Every strided pass keeps stride >= 16
// coefficients; the last pass is the contiguous 16-coefficient (4-stage) pass.
template <int LOGN> struct Sched;
template <> struct Sched<10> { static constexpr int NP = 3; static constexpr int K[4] = { 3, 3, 4, 0 }; };
template <> struct Sched<11> { static constexpr int NP = 3; static constexpr int K[4] = { 3, 4, 4, 0 }; };
template <> struct Sched<12> { static constexpr int NP = 4; static constexpr int K[4] = { 2, 3, 3, 4 }; };
template <> struct Sched<13> { static constexpr int NP = 4; static constexpr int K[4] = { 3, 3, 3, 4 }; };
template <> struct Sched<14> { static constexpr int NP = 4; static constexpr int K[4] = { 3, 3, 4, 4 }; };

template <int LOGN> __host__ __device__ constexpr int sched_start(int p)
{
    int s = 0;
    for (int i = 0; i < p; i++) s += Sched<LOGN>::K[i];
    return s;
}

template <int LOGN> struct NttCfg {
    static constexpr int N = 1 << LOGN;
    static constexpr int THREADS = N / 16;
    static constexpr int SMEM_BYTES = N * 8;
};

__host__ __device__ __forceinline__ int swz(int i) { return i ^ (((i >> 4) & 7) << 1); }

// Geometry of pass P: K stages starting at stage S, stride G, C adjacent columns per thread.
template <int LOGN, int P> struct Pass {
    static constexpr int N = 1 << LOGN;
    static constexpr int S = sched_start<LOGN>(P);
    static constexpr int K = Sched<LOGN>::K[P];
    static constexpr int C = 16 >> K;
    static constexpr int G = N >> (S + K);            // coefficient stride between a thread's rows
    static constexpr int IPB = (G / C) > 0 ? (G / C) : 1;   // thread items per butterfly block
    static constexpr int BLK = N >> S;                // coefficients per block at stage S
    __host__ __device__ static __forceinline__ int hi(int tid) { return tid / IPB; }
    __host__ __device__ static __forceinline__ int base(int tid) { return (tid / IPB) * BLK + (tid % IPB) * C; }
    // coefficient index of register x[j*C + c]
    __host__ __device__ static __forceinline__ int elem(int tid, int j, int c) { return base(tid) + j * G + c; }
};

// ---- register <-> shared-memory exchange for pass P (in place) ----
template <int LOGN, int P, bool STORE> __device__ __forceinline__ void smem_xfer(u64 (&x)[16], u64 *sm, int tid)
{
    typedef Pass<LOGN, P> G;
    const int b = G::base(tid);
    if constexpr (G::C >= 2) {
#pragma unroll
        for (int j = 0; j < (1 << G::K); j++)
#pragma unroll
            for (int c = 0; c < G::C; c += 2) {
                ulonglong2 *p = reinterpret_cast<ulonglong2 *>(sm + swz(b + j * G::G + c));
                if constexpr (STORE) *p = make_ulonglong2(x[j * G::C + c], x[j * G::C + c + 1]);
                else { ulonglong2 v = *p; x[j * G::C + c] = v.x; x[j * G::C + c + 1] = v.y; }
            }
    } else if constexpr (G::G == 1) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            ulonglong2 *p = reinterpret_cast<ulonglong2 *>(sm + swz(b + j));
            if constexpr (STORE) *p = make_ulonglong2(x[j], x[j + 1]);
            else { ulonglong2 v = *p; x[j] = v.x; x[j + 1] = v.y; }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if constexpr (STORE) sm[swz(b + j * G::G)] = x[j];
            else x[j] = sm[swz(b + j * G::G)];
        }
    }
}

__device__ __forceinline__ ulonglong2 ld_tw(const ulonglong2 *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// ---- butterflies of pass P on the 16 registers ----
// `pre` = 2^c + r selects local chunk r of a limb split 2^c ways (pre = 1: unsplit): local stage
// S+u of chunk r is global stage c+S+u, whose twiddle index is (pre << (S+u)) + local block index.
template <int LOGN, int P> __device__ __forceinline__ void bfly_fwd(u64 (&x)[16], const ulonglong2 *__restrict__ tw, u64 q, u64 two_q, int tid, int pre)
{
    typedef Pass<LOGN, P> G;
    const int hi = G::hi(tid);
#pragma unroll
    for (int u = 0; u < G::K; u++) {
        const int half = 1 << (G::K - 1 - u);
#pragma unroll
        for (int blk = 0; blk < (1 << u); blk++) {
            const ulonglong2 w = ld_tw(tw + (pre << (G::S + u)) + (hi << u) + blk);
#pragma unroll
            for (int jj = 0; jj < half; jj++) {
                const int j0 = blk * 2 * half + jj, j1 = j0 + half;
#pragma unroll
                for (int c = 0; c < G::C; c++) ct_bfly(x[j0 * G::C + c], x[j1 * G::C + c], w.x, w.y, q, two_q);
            }
        }
    }
}

// FINAL: this pass contains global stage 0, into which N^{-1} is folded
// (itw[0] holds (w1^{-1} * N^{-1}, shoup) for that purpose).
template <int LOGN, int P, bool FINAL>
__device__ __forceinline__ void bfly_inv(u64 (&x)[16], const ulonglong2 *__restrict__ itw, const Mod &m, int tid, int pre)
{
    typedef Pass<LOGN, P> G;
    const int hi = G::hi(tid);
    const u64 q = m.q, two_q = m.two_q;
    constexpr bool FOLD = FINAL && P == 0;   // this pass ends with global stage 0
#pragma unroll
    for (int u = G::K - 1; u >= (FOLD ? 1 : 0); u--) {
        const int half = 1 << (G::K - 1 - u);
#pragma unroll
        for (int blk = 0; blk < (1 << u); blk++) {
            const ulonglong2 w = ld_tw(itw + (pre << (G::S + u)) + (hi << u) + blk);
#pragma unroll
            for (int jj = 0; jj < half; jj++) {
                const int j0 = blk * 2 * half + jj, j1 = j0 + half;
#pragma unroll
                for (int c = 0; c < G::C; c++) gs_bfly(x[j0 * G::C + c], x[j1 * G::C + c], w.x, w.y, q, two_q);
            }
        }
    }
    if constexpr (FOLD) {
        const ulonglong2 w = ld_tw(itw);
        constexpr int half = 1 << (G::K - 1);
#pragma unroll
        for (int jj = 0; jj < half; jj++)
#pragma unroll
            for (int c = 0; c < G::C; c++) {
                u64 &a = x[jj * G::C + c], &b = x[(jj + half) * G::C + c];
                const u64 s = a + b, d = a - b + two_q;
                a = shoup_lazy(s, m.ninv, m.ninv_s, q);
                b = shoup_lazy(d, w.x, w.y, q);
            }
    }
}

// Forward transform of local chunk r (of 2^c).  In: x in pass-0 layout (Pass<LOGN,0>::elem),
// values < 4q.  Out: x in the contiguous layout (local coefficient 16*tid + j in x[j]), [0,4q).
// Uses sm (2^LOGN words).  Caller must __syncthreads() before reusing sm for another transform.
template <int LOGN, int P = 0>
__device__ __forceinline__ void ntt_fwd_regs_split(u64 (&x)[16], u64 *sm, const ulonglong2 *__restrict__ tw, u64 q, u64 two_q, int tid, int c, int r)
{
    if constexpr (P > 0) {
        __syncthreads();
        smem_xfer<LOGN, P, false>(x, sm, tid);
    }
    bfly_fwd<LOGN, P>(x, tw, q, two_q, tid, (1 << c) + r);
    if constexpr (P + 1 < Sched<LOGN>::NP) {
        smem_xfer<LOGN, P, true>(x, sm, tid);
        ntt_fwd_regs_split<LOGN, P + 1>(x, sm, tw, q, two_q, tid, c, r);
    }
}

// Inverse transform of local chunk r.  In: x in the contiguous layout, values < 2q.
// Out: x in pass-0 layout, [0,2q); FINAL (unsplit limb): already multiplied by N^{-1}.
template <int LOGN, bool FINAL, int P = Sched<LOGN>::NP - 1>
__device__ __forceinline__ void ntt_inv_regs_split(u64 (&x)[16], u64 *sm, const ulonglong2 *__restrict__ itw, const Mod &m, int tid, int c, int r)
{
    if constexpr (P + 1 < Sched<LOGN>::NP) {
        __syncthreads();
        smem_xfer<LOGN, P, false>(x, sm, tid);
    }
    bfly_inv<LOGN, P, FINAL>(x, itw, m, tid, (1 << c) + r);
    if constexpr (P > 0) {
        smem_xfer<LOGN, P, true>(x, sm, tid);
        ntt_inv_regs_split<LOGN, FINAL, P - 1>(x, sm, itw, m, tid, c, r);
    }
}

// ---- global-memory access in the two register layouts (128-bit, pairs) ----
// pass-0 ("strided") layout: register pair (j, c..c+1) <-> coefficients elem(tid,j,c), +1
template <int LOGN, class F> __device__ __forceinline__ void for_pairs_strided(int tid, F &&f)
{
    typedef Pass<LOGN, 0> G;
    static_assert(G::C >= 2, "first pass must hold >= 2 adjacent columns");
#pragma unroll
    for (int j = 0; j < (1 << G::K); j++)
#pragma unroll
        for (int c = 0; c < G::C; c += 2) f(j * G::C + c, G::elem(tid, j, c));
}
// contiguous layout: register pair (2p, 2p+1) <-> coefficients 16*tid + 2p, +1
template <class F> __device__ __forceinline__ void for_pairs_contig(int tid, F &&f)
{
#pragma unroll
    for (int p = 0; p < 16; p += 2) f(p, 16 * tid + p);
}

__device__ __forceinline__ ulonglong2 ldg2(const u64 *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const ulonglong2 *>(p));
#else
    return *reinterpret_cast<const ulonglong2 *>(p);
#endif
}
__device__ __forceinline__ ulonglong2 ld2(const u64 *p) { return *reinterpret_cast<const ulonglong2 *>(p); }
__device__ __forceinline__ void st2(u64 *p, u64 a, u64 b) { *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(a, b); }

}   // namespace b200he
