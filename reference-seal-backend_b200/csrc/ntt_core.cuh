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

// ---- barrier for the exchange between pass P and pass P+1 (either direction) ----
// After stage S_P the transform has split into independent blocks of BLK_P = N >> S_P coefficients,
// so the registers a pass-(P+1) thread picks up were written by the BLK_P/16 threads that own the
// enclosing pass-P block: only those need to meet.  For N = 8192 that is one CTA-wide barrier, one
// 64-thread named barrier and one __syncwarp per transform instead of three CTA-wide barriers.
template <int LOGN, int P> __device__ __forceinline__ void xchg_sync(int tid)
{
    constexpr int GROUP = (Pass<LOGN, P>::BLK) / 16;
#ifdef B200HE_EMU
    (void)tid;
    __syncthreads();
#else
    if constexpr (GROUP >= NttCfg<LOGN>::THREADS) __syncthreads();
    else if constexpr (GROUP <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + tid / GROUP), "r"(GROUP) : "memory");
#endif
}

// ---- twiddles of one pass, fetched into registers ahead of the exchange that precedes the pass ----
// `pre` = 2^c + r selects local chunk r of a limb split 2^c ways (pre = 1: unsplit): local stage
// S+u of chunk r is global stage c+S+u, whose twiddle index is (pre << (S+u)) + local block index.
template <int LOGN, int P> struct TwRegs { ulonglong2 w[(1 << Pass<LOGN, P>::K) - 1]; };

// Stages [U0, U1) of pass P.  A 4-stage pass needs 15 twiddles (60 registers): the stages executed first
// are fetched ahead of the exchange (load_tw_early), the rest once the first stage has retired
// (load_tw_late) so that the transform fits the 128-register budget of a 512-thread CTA.
template <int LOGN, int P, int U0, int U1>
__device__ __forceinline__ void load_tw_range(TwRegs<LOGN, P> &t, const ulonglong2 *__restrict__ tw, int tid, int pre)
{
    typedef Pass<LOGN, P> G;
    const int hi = G::hi(tid);
#pragma unroll
    for (int u = U0; u < U1; u++)
#pragma unroll
        for (int blk = 0; blk < (1 << u); blk++) t.w[(1 << u) - 1 + blk] = ld_tw(tw + (pre << (G::S + u)) + (hi << u) + blk);
}
template <int K> struct TwSplit {
    static constexpr int FWD_EARLY_END = K <= 3 ? K : 3;      // forward runs u = 0..K-1
    static constexpr int INV_EARLY_BEGIN = K <= 3 ? 0 : K - 1;   // inverse runs u = K-1..0
};
template <int LOGN, int P, bool INVERSE> __device__ __forceinline__ void load_tw_early(TwRegs<LOGN, P> &t, const ulonglong2 *__restrict__ tw, int tid, int pre)
{
    constexpr int K = Pass<LOGN, P>::K;
    if constexpr (INVERSE) load_tw_range<LOGN, P, TwSplit<K>::INV_EARLY_BEGIN, K>(t, tw, tid, pre);
    else load_tw_range<LOGN, P, 0, TwSplit<K>::FWD_EARLY_END>(t, tw, tid, pre);
}
template <int LOGN, int P, bool INVERSE> __device__ __forceinline__ void load_tw_late(TwRegs<LOGN, P> &t, const ulonglong2 *__restrict__ tw, int tid, int pre)
{
    constexpr int K = Pass<LOGN, P>::K;
    if constexpr (INVERSE) load_tw_range<LOGN, P, 0, TwSplit<K>::INV_EARLY_BEGIN>(t, tw, tid, pre);
    else load_tw_range<LOGN, P, TwSplit<K>::FWD_EARLY_END, K>(t, tw, tid, pre);
}

// integers (< 2^52) -> FP64 domain, in place (bit patterns of doubles in the same registers)
__device__ __forceinline__ void to_dp_all(u64 (&x)[16])
{
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = as_u(dp_from(x[i]));
}
__device__ __forceinline__ void reduce_all(u64 (&x)[16], const Mod &m)
{
    if (m.dp) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = as_u(dp_reduce(as_d(x[i]), m.dqinv, m.dnq));
        return;
    }
    if (m.pm_c) {
        const u32 b = m.bits, c = m.pm_c;
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = reduce_pm(x[i], b, c);
        return;
    }
    if (m.bits > 32) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = reduce_lazy_t<true>(x[i], m);
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = reduce_lazy_t<false>(x[i], m);
    }
}
// all 16 registers (lazy output of a forward transform, either domain) -> canonical residues as integers
__device__ __forceinline__ void canon_all(u64 (&x)[16], const Mod &m)
{
    if (m.dp) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = dp_canon(as_d(x[i]), m);
        return;
    }
    reduce_all(x, m);
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = csub(x[i], m.q);
}

// ---- butterflies of pass P on the 16 registers ----
// Lazy Cooley-Tukey: X = x + w*y, Y = x - w*y + 2q with w*y in [0,2q) from the Shoup product, which
// accepts ANY 64-bit y.  Nothing is reduced per butterfly: the bound of every value grows by 2q per
// stage and the host-built schedule (Mod::fwd_mask) says where a pass must first pull the registers
// back to [0,2q) so that nothing reaches 2^64.
// FP64 domain (Mod::dp): the same pass on exact integers in doubles.  X = x + v, Y = x - v with |v| <= 0.75 q from
// dp_mul; magnitudes grow by 0.75 q per stage (< 14 q after 15 stages from inputs < 2q), so nothing is ever reduced.
template <int LOGN, int P>
__device__ __forceinline__ void bfly_fwd_dp(u64 (&x)[16], TwRegs<LOGN, P> &t, const Mod &m, const ulonglong2 *__restrict__ tw, int tid, int pre)
{
    typedef Pass<LOGN, P> G;
    const double nq = m.dnq;
#pragma unroll
    for (int u = 0; u < G::K; u++) {
        if (u == 1) load_tw_late<LOGN, P, false>(t, tw, tid, pre);
        const int half = 1 << (G::K - 1 - u);
#pragma unroll
        for (int blk = 0; blk < (1 << u); blk++) {
            const ulonglong2 w = t.w[(1 << u) - 1 + blk];
#pragma unroll
            for (int jj = 0; jj < half; jj++) {
                const int j0 = blk * 2 * half + jj, j1 = j0 + half;
#pragma unroll
                for (int c = 0; c < G::C; c++) {
                    u64 &a = x[j0 * G::C + c], &b = x[j1 * G::C + c];
                    const double av = as_d(a), v = dp_mul(as_d(b), as_d(w.x), as_d(w.y), nq);
                    b = as_u(__dadd_rn(av, -v));
                    a = as_u(__dadd_rn(av, v));
                }
            }
        }
    }
}

template <int LOGN, int P>
__device__ __forceinline__ void bfly_fwd(u64 (&x)[16], TwRegs<LOGN, P> &t, const Mod &m, const ulonglong2 *__restrict__ tw, int tid, int pre)
{
    typedef Pass<LOGN, P> G;
    if (m.dp) {
        bfly_fwd_dp<LOGN, P>(x, t, m, tw, tid, pre);
        return;
    }
    if ((m.fwd_mask >> P) & 1) reduce_all(x, m);
    const bool mid = (m.fwd_mask >> (8 + P)) & 1;
    const u64 nq = m.nq, two_q = m.two_q;
#pragma unroll
    for (int u = 0; u < G::K; u++) {
        if (u == 1) load_tw_late<LOGN, P, false>(t, tw, tid, pre);
        if (u == G::K / 2 && mid) reduce_all(x, m);
        const int half = 1 << (G::K - 1 - u);
#pragma unroll
        for (int blk = 0; blk < (1 << u); blk++) {
            const ulonglong2 w = t.w[(1 << u) - 1 + blk];
#pragma unroll
            for (int jj = 0; jj < half; jj++) {
                const int j0 = blk * 2 * half + jj, j1 = j0 + half;
#pragma unroll
                for (int c = 0; c < G::C; c++) {
                    u64 &a = x[j0 * G::C + c], &b = x[j1 * G::C + c];
                    const u64 v = shoup_mad(b, w.x, w.y, nq, 0);
                    b = a + two_q - v;
                    a = a + v;
                }
            }
        }
    }
}

// Lazy Gentleman-Sande: X = x + y, Y = w*(x - y + C q) where C q bounds y; the Shoup product brings Y
// back to [0,2q) while the X path doubles its bound per stage (schedule: Mod::inv_mask / inv_c).
// FINAL: this pass ends with global stage 0, into which N^{-1} is folded
// (itw[0] holds (w1^{-1} * N^{-1}, shoup) for that purpose); outputs are then in [0,2q).
// FP64 domain: X = x + y doubles its magnitude per stage, Y = w (x - y) comes back to 0.75 q.  Every pass but the
// first (canonical inputs) starts with dp_reduce (|x| <= q/2), so a 4-stage pass peaks at 16 q < 2^50.
template <int LOGN, int P, bool FINAL>
__device__ __forceinline__ void bfly_inv_dp(u64 (&x)[16], TwRegs<LOGN, P> &t, const Mod &m, ulonglong2 wfold, const ulonglong2 *__restrict__ itw, int tid, int pre)
{
    typedef Pass<LOGN, P> G;
    constexpr bool FOLD = FINAL && P == 0;
    if (P != Sched<LOGN>::NP - 1) reduce_all(x, m);
    const double nq = m.dnq;
#pragma unroll
    for (int u = G::K - 1; u >= (FOLD ? 1 : 0); u--) {
        if (u == G::K - 2) load_tw_late<LOGN, P, true>(t, itw, tid, pre);
        const int half = 1 << (G::K - 1 - u);
#pragma unroll
        for (int blk = 0; blk < (1 << u); blk++) {
            const ulonglong2 w = t.w[(1 << u) - 1 + blk];
#pragma unroll
            for (int jj = 0; jj < half; jj++) {
                const int j0 = blk * 2 * half + jj, j1 = j0 + half;
#pragma unroll
                for (int c = 0; c < G::C; c++) {
                    u64 &a = x[j0 * G::C + c], &b = x[j1 * G::C + c];
                    const double av = as_d(a), bv = as_d(b);
                    a = as_u(__dadd_rn(av, bv));
                    b = as_u(dp_mul(__dadd_rn(av, -bv), as_d(w.x), as_d(w.y), nq));
                }
            }
        }
    }
    if constexpr (FOLD) {
        constexpr int half = 1 << (G::K - 1);
#pragma unroll
        for (int jj = 0; jj < half; jj++)
#pragma unroll
            for (int c = 0; c < G::C; c++) {
                u64 &a = x[jj * G::C + c], &b = x[(jj + half) * G::C + c];
                const double av = as_d(a), bv = as_d(b);
                a = as_u(dp_mul(__dadd_rn(av, bv), m.dninv, m.dninv_q, nq));
                b = as_u(dp_mul(__dadd_rn(av, -bv), as_d(wfold.x), as_d(wfold.y), nq));
            }
    }
}

template <int LOGN, int P, bool FINAL>
__device__ __forceinline__ void bfly_inv(u64 (&x)[16], TwRegs<LOGN, P> &t, const Mod &m, ulonglong2 wfold, const ulonglong2 *__restrict__ itw, int tid, int pre)
{
    typedef Pass<LOGN, P> G;
    constexpr bool FOLD = FINAL && P == 0;
    if (m.dp) {
        bfly_inv_dp<LOGN, P, FINAL>(x, t, m, wfold, itw, tid, pre);
        return;
    }
    if ((m.inv_mask >> P) & 1) reduce_all(x, m);
    const bool mid = (m.inv_mask >> (8 + P)) & 1;
    const u64 nq = m.nq;
    u64 cq = m.q * m.inv_c[P];
#pragma unroll
    for (int u = G::K - 1; u >= (FOLD ? 1 : 0); u--) {
        if (u == G::K - 2) load_tw_late<LOGN, P, true>(t, itw, tid, pre);
        if (G::K - 1 - u == G::K / 2 && mid) {
            reduce_all(x, m);
            cq = m.two_q;
        }
        const int half = 1 << (G::K - 1 - u);
#pragma unroll
        for (int blk = 0; blk < (1 << u); blk++) {
            const ulonglong2 w = t.w[(1 << u) - 1 + blk];
#pragma unroll
            for (int jj = 0; jj < half; jj++) {
                const int j0 = blk * 2 * half + jj, j1 = j0 + half;
#pragma unroll
                for (int c = 0; c < G::C; c++) {
                    u64 &a = x[j0 * G::C + c], &b = x[j1 * G::C + c];
                    const u64 d = a + cq - b;
                    a = a + b;
                    b = shoup_mad(d, w.x, w.y, nq, 0);
                }
            }
        }
        cq += cq;
    }
    if constexpr (FOLD) {
        if (G::K - 1 == G::K / 2 && mid) {
            reduce_all(x, m);
            cq = m.two_q;
        }
        constexpr int half = 1 << (G::K - 1);
#pragma unroll
        for (int jj = 0; jj < half; jj++)
#pragma unroll
            for (int c = 0; c < G::C; c++) {
                u64 &a = x[jj * G::C + c], &b = x[(jj + half) * G::C + c];
                const u64 s = a + b, d = a + cq - b;
                a = shoup_mad(s, m.ninv, m.ninv_s, nq, 0);
                b = shoup_mad(d, wfold.x, wfold.y, nq, 0);
            }
    }
}

// Forward transform of local chunk r (of 2^c).  In: x in pass-0 layout (Pass<LOGN,0>::elem), values
// < 2q (+2q per split pre-stage) -- for Mod::dp moduli already converted with to_dp_all (load_fwd_split does) --;
// t = the pass-0 twiddles (load_tw_early<LOGN,0,false>, issued by the caller next to its data loads).
// Out: x in the contiguous layout (local coefficient 16*tid + j in x[j]), lazy: any value < 2^64 congruent to the
// result, or a lazy FP64-domain value for Mod::dp moduli; canon_all finishes either.
// Uses sm (2^LOGN words).  A CTA that runs several transforms through the same buffer passes REUSE = true: the
// CTA-wide barrier that protects the buffer then sits right before the first store into it -- after the global loads
// and the first pass of butterflies -- so warps leave the previous transform's elementwise tail (and enter their loads)
// at their own pace instead of in lockstep.
// `hook` runs right before the butterflies of the last pass: the place from which a caller starts loads it needs
// after the transform (k_ks_inner: the key tile), so that their latency hides behind that pass.
struct NoHook { __device__ __forceinline__ void operator()() const {} };
template <int LOGN, bool REUSE = false, int P = 0, class Hook = NoHook>
__device__ __forceinline__ void ntt_fwd_regs_split(u64 (&x)[16], u64 *sm, const ulonglong2 *__restrict__ tw, const Mod &m, int tid, int c, int r,
                                                   TwRegs<LOGN, P> &t, Hook &&hook = Hook())
{
    if constexpr (P + 1 == Sched<LOGN>::NP) hook();
    bfly_fwd<LOGN, P>(x, t, m, tw, tid, (1 << c) + r);
    if constexpr (P + 1 < Sched<LOGN>::NP) {
        TwRegs<LOGN, P + 1> tn;
        load_tw_early<LOGN, P + 1, false>(tn, tw, tid, (1 << c) + r);
        if constexpr (REUSE && P == 0) __syncthreads();
        smem_xfer<LOGN, P, true>(x, sm, tid);
        xchg_sync<LOGN, P>(tid);
        smem_xfer<LOGN, P + 1, false>(x, sm, tid);
        ntt_fwd_regs_split<LOGN, REUSE, P + 1>(x, sm, tw, m, tid, c, r, tn, hook);
    }
}

// Inverse transform of local chunk r.  In: x in the contiguous layout, canonical values (< q);
// t = twiddles of the last pass (load_tw_early<LOGN,NP-1,true> on the inverse table).
// Out: x in pass-0 layout; FINAL (unsplit limb): multiplied by N^{-1}, in [0,2q); otherwise lazy (Mod::dp moduli:
// FP64-domain values, to be passed through reduce_all and cross_inv).
// IN_DP / OUT_DP (Mod::dp moduli only): the caller hands over / takes back FP64-domain values (in: magnitude <= 1.75 q,
// out: <= 0.75 q) instead of canonical integers, because its own pre- / post-processing runs in the domain too.
template <int LOGN, bool FINAL, bool REUSE = false, int P = Sched<LOGN>::NP - 1, bool IN_DP = false, bool OUT_DP = false>
__device__ __forceinline__ void ntt_inv_regs_split(u64 (&x)[16], u64 *sm, const ulonglong2 *__restrict__ itw, const Mod &m, int tid, int c, int r,
                                                   TwRegs<LOGN, P> &t)
{
    ulonglong2 wfold = make_ulonglong2(0, 0);
    if constexpr (FINAL && P == 0) wfold = ld_tw(itw);
    if constexpr (P == Sched<LOGN>::NP - 1 && !IN_DP) {
        if (m.dp) to_dp_all(x);
    }
    bfly_inv<LOGN, P, FINAL>(x, t, m, wfold, itw, tid, (1 << c) + r);
    if constexpr (FINAL && P == 0 && !OUT_DP) {
        if (m.dp) canon_all(x, m);   // finished values leave the FP64 domain as canonical integers
    }
    if constexpr (P > 0) {
        TwRegs<LOGN, P - 1> tn;
        load_tw_early<LOGN, P - 1, true>(tn, itw, tid, (1 << c) + r);
        if constexpr (REUSE && P == Sched<LOGN>::NP - 1) __syncthreads();
        smem_xfer<LOGN, P, true>(x, sm, tid);
        xchg_sync<LOGN, P - 1>(tid);
        smem_xfer<LOGN, P - 1, false>(x, sm, tid);
        ntt_inv_regs_split<LOGN, FINAL, REUSE, P - 1, IN_DP, OUT_DP>(x, sm, itw, m, tid, c, r, tn);
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

// ---- contiguous layout <-> coalesced global access, through the warp's own slice of the transform buffer ----
// In the contiguous layout a thread owns 16 adjacent coefficients (128 B), so a warp-wide 128-bit access touches
// 32 different 128-byte lines and uses 16 bytes of each: 2x sector over-fetch on loads, partial-line writes, and
// 8 dependent store instructions through one register quad.  A warp's 32 threads together own one contiguous
// 4 KiB range (512 coefficients), which is exactly that warp's private slice of the swizzled transform buffer
// (the last forward / first inverse pass touches no other slice), so an intra-warp transpose through it turns
// the layout into the "lane-interleaved" one -- register pair i <-> coefficients co_elem(tid, i), +1 -- in which
// every warp-wide access covers 512 contiguous bytes.  Both directions are bank-conflict free under swz().
__device__ __forceinline__ void warp_sync()
{
#ifdef B200HE_EMU
    __syncthreads();   // fibers: every thread of the block reaches this point (uniform control flow)
#else
    __syncwarp();
#endif
}
__host__ __device__ __forceinline__ int co_elem(int tid, int i) { return ((tid >> 5) << 9) + (i << 6) + ((tid & 31) << 1); }
// contiguous -> lane-interleaved.  The slice must be free (no pending reads of it by this warp's later code).
__device__ __forceinline__ void contig_to_co(u64 (&x)[16], u64 *sm, int tid)
{
#pragma unroll
    for (int p = 0; p < 8; p++) *reinterpret_cast<ulonglong2 *>(sm + swz(16 * tid + 2 * p)) = make_ulonglong2(x[2 * p], x[2 * p + 1]);
    warp_sync();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(sm + swz(co_elem(tid, i)));
        x[2 * i] = v.x;
        x[2 * i + 1] = v.y;
    }
}
// lane-interleaved -> contiguous; ends with a warp_sync so the slice may be overwritten right away
__device__ __forceinline__ void co_to_contig(u64 (&x)[16], u64 *sm, int tid)
{
#pragma unroll
    for (int i = 0; i < 8; i++) *reinterpret_cast<ulonglong2 *>(sm + swz(co_elem(tid, i))) = make_ulonglong2(x[2 * i], x[2 * i + 1]);
    warp_sync();
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(sm + swz(16 * tid + 2 * p));
        x[2 * p] = v.x;
        x[2 * p + 1] = v.y;
    }
    warp_sync();
}
// lane-interleaved layout: register pair (2i, 2i+1) <-> coefficients co_elem(tid, i), +1
template <class F> __device__ __forceinline__ void for_pairs_co(int tid, F &&f)
{
#pragma unroll
    for (int i = 0; i < 8; i++) f(2 * i, co_elem(tid, i));
}

// ---- TMA bulk copies (cp.async.bulk) of contiguous limb tiles into shared memory, completion on an mbarrier ----
// One elected thread arms the barrier with the byte count and issues the copies; the copy engine moves the tile while
// the CTA computes; every thread waits on the barrier's phase right before it reads the tile.  Used for operands that
// are needed only after a transform (the epilogue operands of k_moddown): their HBM latency disappears behind the NTT.
__device__ __forceinline__ void tma_bar_init(u64 *bar)   // one thread; other threads may touch the barrier only after a CTA-wide barrier
{
#ifdef B200HE_EMU
    *bar = 0;
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((u32)__cvta_generic_to_shared(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void tma_bar_expect(u64 *bar, u32 bytes)   // the issuing thread, before its copies
{
#ifdef B200HE_EMU
    (void)bar; (void)bytes;
#else
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, u32 bytes, u64 *bar)   // bytes % 16 == 0, 16-byte aligned
{
#ifdef B200HE_EMU
    (void)bar;
    memcpy(smem_dst, gsrc, bytes);
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (u32)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((u32)__cvta_generic_to_shared(bar))
                 : "memory");
#endif
}
__device__ __forceinline__ void tma_bar_wait(u64 *bar, u32 phase)
{
#ifdef B200HE_EMU
    (void)bar; (void)phase;
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"((u32)__cvta_generic_to_shared(bar)),
        "r"(phase)
        : "memory");
#endif
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
// ---- Galois permutation applied on load (K8 fused into the key switch) ----
// NTT-form automorphism: g(x)[i] = x[table[i]] (SEAL GaloisTool::apply_galois_ntt).  The table maps every aligned block of
// 2^k consecutive indices onto an aligned block of 2^k indices (consecutive i differ in the top bits of brv(i)), so a
// warp that gathers the 64 coefficients of one lane-interleaved access reads one contiguous 512-byte range of the
// source, lanes permuted: the gather costs no extra sectors.
struct TabPair { u32 a, b; };
__device__ __forceinline__ TabPair ld_tab2(const u32 *p)   // table entries p[0], p[1] (p 8-byte aligned)
{
#if defined(__CUDA_ARCH__)
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    return TabPair{ t.x, t.y };
#else
    return TabPair{ p[0], p[1] };
#endif
}
__device__ __forceinline__ u64 ldg1(const u64 *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// x <- limb[table[.]] in the lane-interleaved layout of chunk `off`: all table loads are issued before the first data load
__device__ __forceinline__ void gather_pairs_co(u64 (&x)[16], const u64 *__restrict__ limb, const u32 *__restrict__ tab_chunk, int tid)
{
    TabPair t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = ld_tab2(tab_chunk + co_elem(tid, i));
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[2 * i] = ldg1(limb + t[i].a);
        x[2 * i + 1] = ldg1(limb + t[i].b);
    }
}
__device__ __forceinline__ void st2(u64 *p, u64 a, u64 b) { *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(a, b); }

}   // namespace b200he
