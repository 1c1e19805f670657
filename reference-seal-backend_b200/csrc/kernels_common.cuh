// kernels_common.cuh -- shared pieces of the NTT-bearing kernel families of the ciphertext-evaluation hot path.
//
// Kernel <-> SEAL routine map and where each lives:
//   k_ntt_fwd / k_ntt_inv   K1 / K2  ntt_negacyclic_harvey and its inverse (+ the "+q/2" rounding and the fused
//                           relinearize+rescale correction of the last limb)                      kernels_ntt.cuh
//   k_ks_inner              K6 step 2: lift digit, NTT, inner product with the key (digits never leave the SM);
//                           K8: the Galois permutation of the target applied on load              kernels_ks.cuh
//   k_moddown*              K6 step 3 and K9: NTT of the rounded last limb fused with subtract, *q_last^{-1}, add
//                                                                                                 kernels_moddown.cuh
//   k_ew / k_tensor* / k_galois_* / k_copy_limbs / k_batch_sum / k_moddown_coeff   K3/K4/K8/K10/K11 and the BFV
//                           coefficient-form variants (not templates, one translation unit)       kernels_ew.cuh
//
// CTA shape.  Every NTT-bearing kernel runs CTAs of N_loc/16 threads (512 for N >= 8192) that own a local transform of
// N_loc = min(N, 8192) coefficients: 64 KiB of shared memory, 16 coefficients per thread in registers.  For N = 16384 /
// 32768 a limb belongs to a group of 2 / 4 CTAs; CTA r owns chunk r.  INVERSE transforms exchange the one or two stages that
// span chunks through distributed shared memory inside a cluster of 2 / 4 (cross_inv: one radix-2/4 butterfly per offset, no
// butterfly computed twice, no second kernel).  In the FORWARD direction a stage that pairs two chunks is cheaper to compute
// from global memory than to exchange: a group of 2 reads its sibling's chunk (L2) and exchanges nothing; a group of 4 does
// that for its first cross stage and exchanges the second inside clusters of TWO (load_fwd_split, cross_fwd_pair) -- clusters
// of four keep only 132 of the 148 SMs busy.  k_ks_inner's special-prime units, which also run inverse transforms, still use
// the four-CTA exchange of cross_fwd.  (Measured)
#pragma once
#include "ntt_core.cuh"
#ifndef B200HE_EMU
#include <cooperative_groups.h>
#endif

namespace b200he {

struct Tables {
    const Mod *mods;          // [M]
    const ulonglong2 *tw;     // [M][N]  forward twiddles (w, shoup)
    const ulonglong2 *itw;    // [M][N]  inverse twiddles; itw[0] = (w1^{-1} N^{-1}, shoup)
    const ulonglong2 *qinv;   // [M][M]  qinv[x*M + j] = (q_x^{-1} mod q_j, shoup)
    const u64 *halfmod;       // [M][M]  halfmod[x*M + j] = (q_x >> 1) mod q_j
    int N, M;
};

__device__ __forceinline__ u64 *dyn_smem()
{
#ifdef B200HE_EMU
    return reinterpret_cast<u64 *>(emu::block_smem());
#else
    extern __shared__ __align__(16) unsigned char b200he_smem[];
    return reinterpret_cast<u64 *>(b200he_smem);
#endif
}

// Input transforms fused into the first-pass load: pair(v, idx) maps the coefficient pair at limb index idx, idx+1.
// (a functor whose pair() already returns FP64-domain values declares gives_dp: load_fwd_split then skips its own conversion)
struct PreNone {
    static constexpr bool gives_dp = false;
    __device__ __forceinline__ ulonglong2 pair(ulonglong2 v, size_t) const { return v; }
};
// Lift of a residue of a wide modulus (up to 61 bits: more than a double's mantissa) into an FP64-domain modulus,
// without an integer multiply: v = vh 2^30 + vl, lifted value = (2^30 vh mod q) + vl with the product in the FP64 domain
// (|result| <= 0.75 q + 2^30).  Residues of moduli of at most 48 bits need no lift at all: the transform accepts any
// integer of magnitude below 2^48 as a lazy value (its outputs then stay below 2^48 + 14 q < 2^50), so they use PreNone.
struct PreLiftDp {
    static constexpr bool gives_dp = true;
    double wq, nq;   // RN(2^30 / q), -q
    __device__ __forceinline__ u64 one(u64 v) const
    {
        const double hi = dp_from(v >> 30), lo = dp_from(v & 0x3fffffffull);
        return as_u(__dadd_rn(dp_mul(hi, 1073741824.0, wq, nq), lo));
    }
    __device__ __forceinline__ ulonglong2 pair(ulonglong2 v, size_t) const { return make_ulonglong2(one(v.x), one(v.y)); }
};
// (the functors carry q and floor(2^64/q) only: a by-value copy of the whole Mod lands in local memory)
__device__ __forceinline__ u64 reduce64_qr(u64 x, u64 q, u64 r64) { return csub(x - mulhi64(x, r64) * q, q); }
// FP64-domain versions of the mod-down input transforms (k_moddown, FP64 instance).  W* = the source modulus is wider
// than 48 bits (split lift); narrower residues are lazy values as they are.
template <bool W> __device__ __forceinline__ double lift_dp(u64 v, double wq30, double nq)
{
    if (!W) return dp_from(v);
    return __dadd_rn(dp_mul(dp_from(v >> 30), 1073741824.0, wq30, nq), dp_from(v & 0x3fffffffull));
}
template <bool W> struct PreReduceFixDp {   // lift(v) + fix
    static constexpr bool gives_dp = true;
    double wq30, nq, fix;
    __device__ __forceinline__ ulonglong2 pair(ulonglong2 v, size_t) const
    {
        return make_ulonglong2(as_u(__dadd_rn(lift_dp<W>(v.x, wq30, nq), fix)), as_u(__dadd_rn(lift_dp<W>(v.y, wq30, nq), fix)));
    }
};
template <bool W1, bool W2> struct PreTwoDp {   // ((lift(u1) + fix1) s + lift(u2) + fix2) r, see PreTwo
    static constexpr bool gives_dp = true;
    double wq30, nq, fix1, fix2, s, sq, r, rq;
    const u64 *rp2;
    __device__ __forceinline__ u64 one(u64 v1, u64 v2) const
    {
        const double a = dp_mul(__dadd_rn(lift_dp<W1>(v1, wq30, nq), fix1), s, sq, nq);
        return as_u(dp_mul(__dadd_rn(__dadd_rn(a, lift_dp<W2>(v2, wq30, nq)), fix2), r, rq, nq));
    }
    __device__ __forceinline__ ulonglong2 pair(ulonglong2 v, size_t idx) const
    {
        const ulonglong2 w = ldg2(rp2 + idx);
        return make_ulonglong2(one(v.x, w.x), one(v.y, w.y));
    }
};
// Lifting a residue of modulus q_x (a value below q_x) into modulus q needs a Barrett reduction only when q_x >= 2q;
// for q_x < 2q -- two 60-bit primes, two 45-bit primes -- one conditional subtraction does it.  WIDE is a template
// parameter: callers branch once on lift_wide() and instantiate both (a CTA-uniform flag inside the functor cost more in
// code shape than the multiplies it saved).
__device__ __forceinline__ bool lift_wide(u64 qx, u64 q) { return qx >= 2 * q; }
template <bool WIDE> __device__ __forceinline__ u64 lift(u64 v, u64 q, u64 r64) { return WIDE ? reduce64_qr(v, q, r64) : csub(v, q); }
template <bool WIDE> struct PreReduce {   // v mod q (lift of a digit into another modulus, SEAL modulo_poly_coeffs)
    static constexpr bool gives_dp = false;
    u64 q, r64;
    __device__ __forceinline__ ulonglong2 pair(ulonglong2 v, size_t) const { return make_ulonglong2(lift<WIDE>(v.x, q, r64), lift<WIDE>(v.y, q, r64)); }
};
template <bool WIDE> struct PreReduceFix {   // (v mod q) + fix   (mod-down / rescale: fix = q - (q_last/2 mod q))
    static constexpr bool gives_dp = false;
    u64 q, r64, fix;
    __device__ __forceinline__ ulonglong2 pair(ulonglong2 v, size_t) const { return make_ulonglong2(lift<WIDE>(v.x, q, r64) + fix, lift<WIDE>(v.y, q, r64) + fix); }
};
// Fused relinearize + rescale (k_moddown with two rounded limbs): the two mod-down corrections of output limb j,
// NTT(u1) * s * r (key switch, s = q_sp^{-1}) and NTT(u2) * r (rescale, r = q_last^{-1}), are one transform of
// (u1 * s + u2) * r because the transform is linear over Z_q.  Result in [0, 2q).
template <bool WIDE> struct PreTwo {
    static constexpr bool gives_dp = false;
    u64 q, r64;
    u64 fix1, fix2;
    ulonglong2 s, r;
    const u64 *rp2;
    __device__ __forceinline__ u64 one(u64 v1, u64 v2) const
    {
        const u64 a = shoup_lazy(lift<WIDE>(v1, q, r64) + fix1, s.x, s.y, q);
        return shoup_lazy(a + lift<WIDE>(v2, q, r64) + fix2, r.x, r.y, q);
    }
    __device__ __forceinline__ ulonglong2 pair(ulonglong2 v, size_t idx) const
    {
        const ulonglong2 w = ldg2(rp2 + idx);
        return make_ulonglong2(one(v.x, w.x), one(v.y, w.y));
    }
};

// lazy Cooley-Tukey butterfly used by the split pre-stages (bound of both outputs: bound(a) + 2q)
__device__ __forceinline__ void ct_lazy(u64 &a, u64 &b, ulonglong2 w, const Mod &m)
{
    if (m.dp) {   // FP64 domain (modarith.cuh): magnitudes grow by 0.75 q
        const double av = as_d(a), v = dp_mul(as_d(b), as_d(w.x), as_d(w.y), m.dnq);
        b = as_u(__dadd_rn(av, -v));
        a = as_u(__dadd_rn(av, v));
        return;
    }
    const u64 v = shoup_mad(b, w.x, w.y, m.nq, 0);
    b = a + m.two_q - v;
    a = a + v;
}

// ---- thread-block clusters: a limb of N = 2^c * NL coefficients belongs to a cluster of 2^c CTAs ----
// CTA r of the cluster owns chunk r (coefficients [r NL, (r+1) NL)) in its registers / shared memory.  The c
// transform stages that span chunks (the first c Cooley-Tukey stages, the last c Gentleman-Sande stages) pair the
// SAME offset e of different chunks, so they run as one radix-2^c butterfly per offset on values exchanged through
// distributed shared memory: every thread owns 16 >> c of its 16 offsets (register pairs [own r, own (r+1))), pulls
// the other chunks' values at those offsets from the peers' transform buffers, computes all 2^c outputs once, keeps
// its own and pushes the others back into the slots it just read.  No butterfly is computed twice and a limb
// crosses HBM exactly once per direction for every N (the earlier design recomputed the cross stages per CTA from
// global memory and finished split inverses in a second kernel).
#ifdef B200HE_EMU
__device__ __forceinline__ void cluster_sync() { emu::cluster_sync(); }
__device__ __forceinline__ u64 *cluster_peer(u64 *sm, int rank)
{
    return reinterpret_cast<u64 *>(emu::cluster_smem((unsigned)rank) + (reinterpret_cast<unsigned char *>(sm) - emu::block_smem()));
}
#else
__device__ __forceinline__ void cluster_sync() { cooperative_groups::this_cluster().sync(); }
__device__ __forceinline__ u64 *cluster_peer(u64 *sm, int rank) { return cooperative_groups::this_cluster().map_shared_rank(sm, (unsigned)rank); }
#endif

// 4-way select / scatter by the (CTA-uniform) chunk rank, with constant register indices on every path
__device__ __forceinline__ u64 sel4(int r, u64 a, u64 b, u64 c, u64 d) { return r == 0 ? a : r == 1 ? b : r == 2 ? c : d; }
__device__ __forceinline__ void put4(int r, u64 v0, u64 v1, u64 v2, u64 v3, u64 &a, u64 &b, u64 &c, u64 &d)
{
    if (r == 0) a = v0;
    else if (r == 1) b = v1;
    else if (r == 2) c = v2;
    else d = v3;
}
// publish the register pairs owned by other CTAs / fetch them back after the owners have pushed the results
template <int LOGN> __device__ __forceinline__ void cross_publish(const u64 (&x)[16], u64 *sm, int c, int r, int tid)
{
    const int own = 16 >> c;
    for_pairs_strided<LOGN>(tid, [&](int reg, int e) {
        if (reg / own != r) st2(sm + swz(e), x[reg], x[reg + 1]);
    });
}
template <int LOGN> __device__ __forceinline__ void cross_collect(u64 (&x)[16], const u64 *sm, int c, int r, int tid)
{
    const int own = 16 >> c;
    for_pairs_strided<LOGN>(tid, [&](int reg, int e) {
        if (reg / own != r) {
            const ulonglong2 v = ld2(sm + swz(e));
            x[reg] = v.x;
            x[reg + 1] = v.y;
        }
    });
}

// First c Cooley-Tukey stages across the chunks of a cluster.  In: x = pass-0 layout of chunk r, values < 2q.
// Out: the same registers after global stages 0..c-1, values < (2 + 2c) q.  Twiddles: stage 0 tw[1]; stage 1 tw[2]
// (chunks 0,1) and tw[3] (chunks 2,3).
// `hook` runs right before the second cluster barrier: the place from which a caller starts loads it needs after the cross
// stages (k_ks_inner: the twiddles of the first local pass), so that they are not live across the radix-4 butterflies and
// their latency hides behind the barrier.
template <int LOGN, class Hook = NoHook>
__device__ __forceinline__ void cross_fwd(u64 (&x)[16], u64 *sm, int c, int r, int tid, const ulonglong2 *__restrict__ tw, const Mod &m,
                                          Hook &&hook = Hook())
{
    const int own = 16 >> c;
    cross_publish<LOGN>(x, sm, c, r, tid);
    cluster_sync();
    {   // c == 2 (clusters of two never come here: load_fwd_split computes their single cross stage from global memory)
        u64 *peer[4];
#pragma unroll
        for (int q = 0; q < 4; q++) peer[q] = cluster_peer(sm, q);
        // CTA r owns register pairs (4r, 4r+1) and (4r+2, 4r+3) = rows 2r and 2r+1 of the pass-0 layout.  All six peer
        // values are requested before the first is used (one DSMEM latency, not two), the own operands are selected from
        // the four candidate register groups so that no register index depends on r.
        typedef Pass<LOGN, 0> G0;
        const ulonglong2 w1 = ld_tw(tw + 1), w2 = ld_tw(tw + 2), w3 = ld_tw(tw + 3);
        int es[2];
        ulonglong2 pv[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            es[i] = swz(G0::elem(tid, 2 * r + i, 0));
            const ulonglong2 own2 = make_ulonglong2(sel4(r, x[2 * i], x[4 + 2 * i], x[8 + 2 * i], x[12 + 2 * i]),
                                                    sel4(r, x[2 * i + 1], x[5 + 2 * i], x[9 + 2 * i], x[13 + 2 * i]));
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (q == r) pv[i][q] = own2;
                else pv[i][q] = ld2(peer[q] + es[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            u64 o[4][2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                u64 a0 = h ? pv[i][0].y : pv[i][0].x, a1 = h ? pv[i][1].y : pv[i][1].x, a2 = h ? pv[i][2].y : pv[i][2].x, a3 = h ? pv[i][3].y : pv[i][3].x;
                ct_lazy(a0, a2, w1, m);
                ct_lazy(a1, a3, w1, m);
                ct_lazy(a0, a1, w2, m);
                ct_lazy(a2, a3, w3, m);
                o[0][h] = a0; o[1][h] = a1; o[2][h] = a2; o[3][h] = a3;
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (q != r) st2(peer[q] + es[i], o[q][0], o[q][1]);
            put4(r, o[0][0], o[1][0], o[2][0], o[3][0], x[2 * i], x[4 + 2 * i], x[8 + 2 * i], x[12 + 2 * i]);
            put4(r, o[0][1], o[1][1], o[2][1], o[3][1], x[2 * i + 1], x[5 + 2 * i], x[9 + 2 * i], x[13 + 2 * i]);
        }
    }
    hook();
    cluster_sync();
    cross_collect<LOGN>(x, sm, c, r, tid);
}

// ONE Cooley-Tukey cross stage exchanged inside a cluster of TWO: `lo` = which chunk of the pair this CTA holds (its cluster
// rank), w = the stage's twiddle.  Used by the forward-only kernels for limbs of four chunks (load_fwd_split, PAIRS): the
// stage that pairs chunk r with r ^ 2 is computed from global memory, this one pairs r with r ^ 1.  Clusters of two fill
// all 148 SMs; clusters of four only 132 (tools/cluster_occupancy.cu).
// CTA `lo` owns register pairs [4 lo, 4 lo + 4): all four peer values are requested before the first is used, the own
// operands are selected from the two candidate register groups so that no register index depends on lo.
template <int LOGN, class Hook = NoHook>
__device__ __forceinline__ void cross_fwd_pair(u64 (&x)[16], u64 *sm, int lo, int tid, ulonglong2 w, const Mod &m, Hook &&hook = Hook())
{
    typedef Pass<LOGN, 0> G0;
    cross_publish<LOGN>(x, sm, 1, lo, tid);
    cluster_sync();
    {
        u64 *peer = cluster_peer(sm, lo ^ 1);
        ulonglong2 pv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) pv[i] = ld2(peer + swz(G0::elem(tid, 4 * lo + i, 0)));
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const u64 o0 = lo ? x[8 + 2 * i] : x[2 * i], o1 = lo ? x[9 + 2 * i] : x[2 * i + 1];
            u64 a0 = lo ? pv[i].x : o0, a1 = lo ? pv[i].y : o1;   // lower chunk of the pair
            u64 b0 = lo ? o0 : pv[i].x, b1 = lo ? o1 : pv[i].y;   // upper chunk
            ct_lazy(a0, b0, w, m);
            ct_lazy(a1, b1, w, m);
            st2(peer + swz(G0::elem(tid, 4 * lo + i, 0)), lo ? a0 : b0, lo ? a1 : b1);
            if (lo) { x[8 + 2 * i] = b0; x[9 + 2 * i] = b1; }
            else { x[2 * i] = a0; x[2 * i + 1] = a1; }
        }
    }
    hook();
    cluster_sync();
    cross_collect<LOGN>(x, sm, 1, lo, tid);
}

// final cross-chunk stage of the inverse: (a + b) N^{-1} and (a - b) wn, results in [0, 2q) as integers (either domain)
__device__ __forceinline__ void inv_last(u64 a, u64 b, ulonglong2 wn, const Mod &m, u64 &s, u64 &d)
{
    if (m.dp) {
        const double av = as_d(a), bv = as_d(b);
        s = dp_canon(dp_mul(__dadd_rn(av, bv), m.dninv, m.dninv_q, m.dnq), m);
        d = dp_canon(dp_mul(__dadd_rn(av, -bv), as_d(wn.x), as_d(wn.y), m.dnq), m);
        return;
    }
    s = shoup_lazy(a + b, m.ninv, m.ninv_s, m.q);
    d = shoup_lazy(a - b + m.two_q, wn.x, wn.y, m.q);
}
// Gentleman-Sande butterfly of the cross-chunk stages: inputs reduced (reduce_all), either domain
__device__ __forceinline__ void gs_cross(u64 &x, u64 &y, ulonglong2 w, const Mod &m)
{
    if (m.dp) {
        const double av = as_d(x), bv = as_d(y);
        x = as_u(__dadd_rn(av, bv));
        y = as_u(dp_mul(__dadd_rn(av, -bv), as_d(w.x), as_d(w.y), m.dnq));
        return;
    }
    gs_bfly(x, y, w.x, w.y, m.q, m.two_q);
}

// Last c Gentleman-Sande stages across the chunks of a cluster, with N^{-1} folded into the final one.
// In: x = pass-0 layout of chunk r after the local stages, reduced (reduce_all: [0, 2q), or |x| <= q/2 in the FP64
// domain).  Out: finished values in [0, 2q), integers in either domain.
template <int LOGN>
__device__ __forceinline__ void cross_inv(u64 (&x)[16], u64 *sm, int c, int r, int tid, const ulonglong2 *__restrict__ itw, const Mod &m)
{
    const int own = 16 >> c;
    cross_publish<LOGN>(x, sm, c, r, tid);
    cluster_sync();
    const ulonglong2 wn = ld_tw(itw);
    if (c == 1) {
        typedef Pass<LOGN, 0> G0;
        // (limbs are split only into chunks of 8192 coefficients -- KERNEL_DISPATCH -- whose first pass holds 8 rows x 2 columns)
        u64 *peer = cluster_peer(sm, r ^ 1);
        ulonglong2 pv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) pv[i] = ld2(peer + swz(G0::elem(tid, 4 * r + i, 0)));
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const u64 o0 = r ? x[8 + 2 * i] : x[2 * i], o1 = r ? x[9 + 2 * i] : x[2 * i + 1];
            const u64 a0 = r ? pv[i].x : o0, a1 = r ? pv[i].y : o1, b0 = r ? o0 : pv[i].x, b1 = r ? o1 : pv[i].y;
            u64 s0, s1, d0, d1;
            inv_last(a0, b0, wn, m, s0, d0);
            inv_last(a1, b1, wn, m, s1, d1);
            st2(peer + swz(G0::elem(tid, 4 * r + i, 0)), r ? s0 : d0, r ? s1 : d1);
            if (r) { x[8 + 2 * i] = d0; x[9 + 2 * i] = d1; }
            else { x[2 * i] = s0; x[2 * i + 1] = s1; }
        }
    } else {
        u64 *peer[4];
#pragma unroll
        for (int q = 0; q < 4; q++) peer[q] = cluster_peer(sm, q);
        // (same structure as cross_fwd: all six peer values requested first, own operands selected by r)
        typedef Pass<LOGN, 0> G0;
        const ulonglong2 w2 = ld_tw(itw + 2), w3 = ld_tw(itw + 3);
        int es[2];
        ulonglong2 pv[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            es[i] = swz(G0::elem(tid, 2 * r + i, 0));
            const ulonglong2 own2 = make_ulonglong2(sel4(r, x[2 * i], x[4 + 2 * i], x[8 + 2 * i], x[12 + 2 * i]),
                                                    sel4(r, x[2 * i + 1], x[5 + 2 * i], x[9 + 2 * i], x[13 + 2 * i]));
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (q == r) pv[i][q] = own2;
                else pv[i][q] = ld2(peer[q] + es[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            u64 o[4][2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                u64 a0 = h ? pv[i][0].y : pv[i][0].x, a1 = h ? pv[i][1].y : pv[i][1].x, a2 = h ? pv[i][2].y : pv[i][2].x, a3 = h ? pv[i][3].y : pv[i][3].x;
                gs_cross(a0, a1, w2, m);
                gs_cross(a2, a3, w3, m);
                inv_last(a0, a2, wn, m, o[0][h], o[2][h]);
                inv_last(a1, a3, wn, m, o[1][h], o[3][h]);
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (q != r) st2(peer[q] + es[i], o[q][0], o[q][1]);
            put4(r, o[0][0], o[1][0], o[2][0], o[3][0], x[2 * i], x[4 + 2 * i], x[8 + 2 * i], x[12 + 2 * i]);
            put4(r, o[0][1], o[1][1], o[2][1], o[3][1], x[2 * i + 1], x[5 + 2 * i], x[9 + 2 * i], x[13 + 2 * i]);
        }
    }
    cluster_sync();
    cross_collect<LOGN>(x, sm, c, r, tid);
}

// Load the pass-0 register layout of chunk r of a limb (src points at the limb, N = NL << c coefficients), apply the
// input transform, and run the c cross-chunk stages.  pre.pair() must return values < 2q; the result is < (2 + 2c) q
// (Mod::dp moduli: converted to the FP64 domain right after the load).
// REUSE: the CTA has used the transform buffer before (see ntt_fwd_regs_split).
// PAIRS: the kernel was launched with clusters of TWO although the limb has four chunks (forward-only kernels): the first
// cross stage comes from global memory as for c == 1, the second is exchanged inside the pair (cross_fwd_pair).
template <int LOGN, bool REUSE = false, bool PAIRS = false, class Pre, class Hook = NoHook>
__device__ __forceinline__ void load_fwd_split(u64 (&x)[16], const u64 *__restrict__ src, int c, int r, int tid,
                                               const ulonglong2 *__restrict__ tw, const Mod &m, Pre pre, u64 *sm, Hook &&hook = Hook())
{
    constexpr int NL = 1 << LOGN;
    const size_t off = (size_t)r * NL;
    for_pairs_strided<LOGN>(tid, [&](int reg, int e) {
        const ulonglong2 v = pre.pair(ldg2(src + off + e), off + e);
        x[reg] = v.x;
        x[reg + 1] = v.y;
    });
    if constexpr (!Pre::gives_dp) {
        if (m.dp) to_dp_all(x);
    }
    if (c == 1) {
        // Clusters of two, forward direction: the one cross-chunk stage pairs offset e of chunk 0 with offset e of chunk 1,
        // and both come straight from global memory.  Each CTA therefore loads the sibling chunk as well (served from L2:
        // the sibling reads the same lines at the same time) and computes the stage for its own half -- one duplicated
        // multiply per coefficient (half a stage of fourteen) against two cluster barriers and a DSMEM round trip.
        // Measured on the B200 (C3 / C4 Val through the plugin): k_ks_inner -4 % / -10 %, k_moddown -8 % / -9 %.  (For
        // clusters of four the same trade costs three multiplies per coefficient instead of one and four times the loads:
        // they keep the DSMEM exchange of cross_fwd.)
        const size_t offp = (size_t)(r ^ 1) * NL;
        const ulonglong2 w = ld_tw(tw + 1);
        for_pairs_strided<LOGN>(tid, [&](int reg, int e) {
            ulonglong2 p = pre.pair(ldg2(src + offp + e), offp + e);
            if constexpr (!Pre::gives_dp) {
                if (m.dp) { p.x = as_u(dp_from(p.x)); p.y = as_u(dp_from(p.y)); }
            }
            u64 a0 = r ? p.x : x[reg], a1 = r ? p.y : x[reg + 1];   // chunk 0
            u64 b0 = r ? x[reg] : p.x, b1 = r ? x[reg + 1] : p.y;   // chunk 1
            ct_lazy(a0, b0, w, m);
            ct_lazy(a1, b1, w, m);
            x[reg] = r ? b0 : a0;
            x[reg + 1] = r ? b1 : a1;
        });
        hook();
        return;
    }
    if (PAIRS && c == 2) {
        const int hi = r >> 1;   // this chunk is the upper one of the stage that pairs r with r ^ 2
        const size_t offp = (size_t)(r ^ 2) * NL;
        const ulonglong2 w = ld_tw(tw + 1);
        for_pairs_strided<LOGN>(tid, [&](int reg, int e) {
            ulonglong2 p = pre.pair(ldg2(src + offp + e), offp + e);
            if constexpr (!Pre::gives_dp) {
                if (m.dp) { p.x = as_u(dp_from(p.x)); p.y = as_u(dp_from(p.y)); }
            }
            u64 a0 = hi ? p.x : x[reg], a1 = hi ? p.y : x[reg + 1];
            u64 b0 = hi ? x[reg] : p.x, b1 = hi ? x[reg + 1] : p.y;
            ct_lazy(a0, b0, w, m);
            ct_lazy(a1, b1, w, m);
            x[reg] = hi ? b0 : a0;
            x[reg + 1] = hi ? b1 : a1;
        });
        if (REUSE) __syncthreads();
        cross_fwd_pair<LOGN>(x, sm, r & 1, tid, ld_tw(tw + 2 + hi), m, hook);
        return;
    }
    if (c > 0) {
        if (REUSE) __syncthreads();
        cross_fwd<LOGN>(x, sm, c, r, tid, tw, m, hook);
    }
}

}   // namespace b200he
