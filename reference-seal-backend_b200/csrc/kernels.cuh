// kernels.cuh -- sm_100a kernels of the ciphertext-evaluation hot path.
//
// Kernel <-> SEAL routine map:
//   k_ntt_fwd            K1  ntt_negacyclic_harvey
//   k_ntt_inv (+tail)    K2  inverse_ntt_negacyclic_harvey (+ "+q/2" of mod-down/rescale fused)
//   k_ew / k_tensor      K3/K4/K11  add/sub/dyadic product, ckks_multiply, multiply_plain, add_plain
//   k_ks_inner           K6 step 2: lift digit, NTT, inner product with the key (digits never leave the SM)
//   k_moddown            K6 step 3 and K9: NTT of the rounded last limb fused with subtract, *q_last^{-1}, add
//   k_galois_*           K8 apply_galois_ntt / apply_galois
//
// CTA shape.  Every NTT-bearing kernel runs 512-thread CTAs (N <= 8192: N/16 threads) that
// own a local transform of n_loc = min(N, 8192) coefficients: 64 KiB of shared memory, 16
// coefficients per thread in registers.  For N = 16384 / 32768 a limb is split over 2 / 4
// independent CTAs: after the first c = log2(N/n_loc) Cooley-Tukey stages the transform
// decomposes into 2^c independent sub-transforms on contiguous output ranges, so CTA r
// recomputes those c stages for its own range straight from global memory (c+1... inputs per
// output, second reader hits L2) and never talks to its siblings.  The inverse direction
// runs the local stages per 8192-chunk and finishes the last c stages in an elementwise
// tail kernel.
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
template <int LOGN>
__device__ __forceinline__ void cross_fwd(u64 (&x)[16], u64 *sm, int c, int r, int tid, const ulonglong2 *__restrict__ tw, const Mod &m)
{
    const int own = 16 >> c;
    cross_publish<LOGN>(x, sm, c, r, tid);
    cluster_sync();
    if (c == 1) {
        // CTA r owns register pairs [4r, 4r + 4): pair i of those sits at offset base + (4r + i) G.  All four peer
        // values are requested before the first is used (DSMEM latency paid once); the own operands are selected
        // from the two candidate register groups so that no register index depends on r.
        typedef Pass<LOGN, 0> G0;
        // (limbs are split only into chunks of 8192 coefficients -- KERNEL_DISPATCH -- whose first pass holds 8 rows x 2 columns)
        u64 *peer = cluster_peer(sm, r ^ 1);
        const ulonglong2 w = ld_tw(tw + 1);
        ulonglong2 pv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) pv[i] = ld2(peer + swz(G0::elem(tid, 4 * r + i, 0)));
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const u64 o0 = r ? x[8 + 2 * i] : x[2 * i], o1 = r ? x[9 + 2 * i] : x[2 * i + 1];
            u64 a0 = r ? pv[i].x : o0, a1 = r ? pv[i].y : o1;   // chunk 0
            u64 b0 = r ? o0 : pv[i].x, b1 = r ? o1 : pv[i].y;   // chunk 1
            ct_lazy(a0, b0, w, m);
            ct_lazy(a1, b1, w, m);
            st2(peer + swz(G0::elem(tid, 4 * r + i, 0)), r ? a0 : b0, r ? a1 : b1);
            if (r) { x[8 + 2 * i] = b0; x[9 + 2 * i] = b1; }
            else { x[2 * i] = a0; x[2 * i + 1] = a1; }
        }
    } else {
        u64 *peer[4];
#pragma unroll
        for (int q = 0; q < 4; q++) peer[q] = cluster_peer(sm, q);
        const ulonglong2 w1 = ld_tw(tw + 1), w2 = ld_tw(tw + 2), w3 = ld_tw(tw + 3);
        for_pairs_strided<LOGN>(tid, [&](int reg, int e) {
            if (reg / own != r) return;
            ulonglong2 v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = (q == r) ? make_ulonglong2(x[reg], x[reg + 1]) : ld2(peer[q] + swz(e));
            u64 o[4][2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                u64 a0 = h ? v[0].y : v[0].x, a1 = h ? v[1].y : v[1].x, a2 = h ? v[2].y : v[2].x, a3 = h ? v[3].y : v[3].x;
                ct_lazy(a0, a2, w1, m);
                ct_lazy(a1, a3, w1, m);
                ct_lazy(a0, a1, w2, m);
                ct_lazy(a2, a3, w3, m);
                o[0][h] = a0; o[1][h] = a1; o[2][h] = a2; o[3][h] = a3;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (q == r) { x[reg] = o[q][0]; x[reg + 1] = o[q][1]; }
                else st2(peer[q] + swz(e), o[q][0], o[q][1]);
            }
        });
    }
    cluster_sync();
    cross_collect<LOGN>(x, sm, c, r, tid);
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
        const ulonglong2 w2 = ld_tw(itw + 2), w3 = ld_tw(itw + 3);
        for_pairs_strided<LOGN>(tid, [&](int reg, int e) {
            if (reg / own != r) return;
            ulonglong2 v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = (q == r) ? make_ulonglong2(x[reg], x[reg + 1]) : ld2(peer[q] + swz(e));
            u64 o[4][2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                u64 a0 = h ? v[0].y : v[0].x, a1 = h ? v[1].y : v[1].x, a2 = h ? v[2].y : v[2].x, a3 = h ? v[3].y : v[3].x;
                gs_cross(a0, a1, w2, m);
                gs_cross(a2, a3, w3, m);
                inv_last(a0, a2, wn, m, o[0][h], o[2][h]);
                inv_last(a1, a3, wn, m, o[1][h], o[3][h]);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (q == r) { x[reg] = o[q][0]; x[reg + 1] = o[q][1]; }
                else st2(peer[q] + swz(e), o[q][0], o[q][1]);
            }
        });
    }
    cluster_sync();
    cross_collect<LOGN>(x, sm, c, r, tid);
}

// Load the pass-0 register layout of chunk r of a limb (src points at the limb, N = NL << c coefficients), apply the
// input transform, and run the c cross-chunk stages.  pre.pair() must return values < 2q; the result is < (2 + 2c) q
// (Mod::dp moduli: converted to the FP64 domain right after the load).
// REUSE: the CTA has used the transform buffer before (see ntt_fwd_regs_split).
template <int LOGN, bool REUSE = false, class Pre>
__device__ __forceinline__ void load_fwd_split(u64 (&x)[16], const u64 *__restrict__ src, int c, int r, int tid,
                                               const ulonglong2 *__restrict__ tw, const Mod &m, Pre pre, u64 *sm)
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
    if (c > 0) {
        if (REUSE) __syncthreads();
        cross_fwd<LOGN>(x, sm, c, r, tid, tw, m);
    }
}

// ------------------------------------------------------------------------------------ K1
// dst[w] = NTT(src[w]) for w < nlimbs; modulus id = mod_base + (w % L).  grid = nlimbs << c.
// Limb w lives at base + (w / L) * outer + (w % L) * N  (outer = L*N for a contiguous batch; a larger
// outer stride addresses one limb per polynomial, e.g. the special-prime limb of the key-switch accumulator).
// C = log2 of the cluster size (CTAs per limb), a compile-time constant so that the unsplit case carries none of
// the cluster code.
template <int LOGN, int C>
__global__ void __launch_bounds__(NttCfg<LOGN>::THREADS) k_ntt_fwd(Tables T, const u64 *__restrict__ src, u64 *__restrict__ dst,
                                                                   size_t src_outer, size_t dst_outer, int L, int mod_base)
{
    constexpr int c = C;
    constexpr int NL = 1 << LOGN;
    u64 *sm = dyn_smem();
    const int tid = threadIdx.x;
    const int w = blockIdx.x >> c, r = blockIdx.x & ((1 << c) - 1);
    const int mid = mod_base + (w % L);
    const Mod m = T.mods[mid];
    const ulonglong2 *tw = T.tw + (size_t)mid * T.N;
    u64 x[16];
    TwRegs<LOGN, 0> t0;
    load_tw_early<LOGN, 0, false>(t0, tw, tid, (1 << c) + r);
    load_fwd_split<LOGN>(x, src + (size_t)(w / L) * src_outer + (size_t)(w % L) * T.N, c, r, tid, tw, m, PreNone(), sm);
    ntt_fwd_regs_split<LOGN>(x, sm, tw, m, tid, c, r, t0);
    u64 *out = dst + (size_t)(w / L) * dst_outer + (size_t)(w % L) * T.N + (size_t)r * NL;
    canon_all(x, m);
    contig_to_co(x, sm, tid);
    for_pairs_co(tid, [&](int reg, int e) { st2(out + e, x[reg], x[reg + 1]); });
}

// Persistent variant for unsplit limbs: one CTA per SM walks the limbs w = blockIdx.x, blockIdx.x + gridDim.x, ...
// With one 512-thread CTA per SM nothing overlaps a CTA's first global loads; here the next limb's 8 NL bytes arrive by
// a TMA bulk copy in a second shared-memory buffer while the current limb is transformed, the first pass reads them with
// conflict-free 128-bit shared loads, and the stores of a limb drain behind the next limb's arithmetic.
template <int LOGN> struct NttFwdPCfg {
    static constexpr int SMEM_BYTES = 2 * NttCfg<LOGN>::SMEM_BYTES + 16;   // transform buffer | landing buffer | mbarrier
};
template <int LOGN>
__global__ void __launch_bounds__(NttCfg<LOGN>::THREADS) k_ntt_fwd_p(Tables T, const u64 *__restrict__ src, u64 *__restrict__ dst,
                                                                     size_t src_outer, size_t dst_outer, int L, int mod_base, int nlimbs)
{
    constexpr int NL = 1 << LOGN;
    u64 *sm = dyn_smem();
    u64 *land = sm + NL, *bar = sm + 2 * NL;
    const int tid = threadIdx.x;
    auto limb_src = [&](int w) { return src + (size_t)(w / L) * src_outer + (size_t)(w % L) * T.N; };
    // round `it` gives CTA b limb it G + (b + it) mod G: consecutive rounds of a CTA land on different moduli, so the
    // slow (60-bit) and fast (FP64-domain) limbs spread evenly over the SMs whatever G mod L is
    const int G = gridDim.x, b = blockIdx.x;
    auto limb_of = [&](int it) { return it * G + (b + it) % G; };
    if (tid == 0) {
        tma_bar_init(bar);
        tma_bar_expect(bar, NL * 8);
        tma_load_1d(land, limb_src(limb_of(0)), NL * 8, bar);
    }
    __syncthreads();   // the barrier is initialised before anyone waits on it
    u32 phase = 0;
    for (int it = 0; limb_of(it) < nlimbs; it++, phase ^= 1) {
        const int w = limb_of(it);
        const int mid = mod_base + (w % L);
        const Mod m = T.mods[mid];
        const ulonglong2 *tw = T.tw + (size_t)mid * T.N;
        u64 x[16];
        TwRegs<LOGN, 0> t0;
        load_tw_early<LOGN, 0, false>(t0, tw, tid, 1);
        tma_bar_wait(bar, phase);
        for_pairs_strided<LOGN>(tid, [&](int reg, int e) {
            const ulonglong2 v = ld2(land + e);
            x[reg] = v.x;
            x[reg + 1] = v.y;
        });
        if (m.dp) to_dp_all(x);
        // every thread has read the landing buffer (and finished with the transform buffer of the previous limb): the
        // next limb's copy may start
        __syncthreads();
        if (tid == 0 && limb_of(it + 1) < nlimbs) {
            tma_bar_expect(bar, NL * 8);
            tma_load_1d(land, limb_src(limb_of(it + 1)), NL * 8, bar);
        }
        ntt_fwd_regs_split<LOGN>(x, sm, tw, m, tid, 0, 0, t0);
        u64 *out = dst + (size_t)(w / L) * dst_outer + (size_t)(w % L) * T.N;
        canon_all(x, m);
        contig_to_co(x, sm, tid);
        for_pairs_co(tid, [&](int reg, int e) { st2(out + e, x[reg], x[reg + 1]); });
        // (the warp's slice of the transform buffer is read only by this warp; its next use is the first-pass store of the
        //  next limb, which sits behind the CTA-wide barrier above)
    }
}

// ------------------------------------------------------------------------------------ K2
enum { INV_PLAIN = 0, INV_ADDHALF = 1 };
__device__ __forceinline__ u64 inv_finish(u64 v /* [0,2q) */, const Mod &m, int mode)
{
    v = csub(v, m.q);
    if (mode == INV_ADDHALF) v = csub(v + (m.q >> 1), m.q);
    return v;
}
// Optional fusion around the inverse transform of limb w = (b, p) (b = w / P, p = w % P), used by the fused
// relinearize + rescale for the last data limb j:
//   input   y = src * s + add[b][p]          (s = q_x^{-1} mod q_j: the key-switch accumulator scaled and added to the
//                                             input ciphertext, still in NTT form)
//   output  iNTT(y) - ((sub[w] mod q_j) + fix) * s   (the mod-down correction applied in coefficient form: the
//                                             transform is linear, so NTT(u) never has to be computed for this limb)
// followed by the usual finish (INV_ADDHALF: + q_j / 2, the rounding of the rescale that follows).
struct InvFuse {
    const u64 *add;        // nullptr: plain transform
    size_t add_ct_stride, add_poly_stride;
    const u64 *sub;        // [nlimbs][N] coefficient form (the rounded special-prime limb)
    int P, x;              // polys per ciphertext; x = modulus id of the prime dropped by the key switch
};
__device__ __forceinline__ u64 inv_post(u64 v /* finished, canonical */, u64 subv, const Mod &m, ulonglong2 s, u64 fix)
{
    return sub_mod(v, shoup(reduce64(subv, m) + fix, s.x, s.y, m.q), m.q);
}
// dst = iNTT(src), finished (c > 0: the cluster's CTAs exchange the cross-chunk stages through DSMEM).
// KIND: 0 = the launch holds limbs of both kinds (Mod::dp decides per CTA); 1 = integer-pipe moduli only; 2 = FP64-domain
// moduli only.  With the kind known at compile time the other instance is not in the kernel, and its register demands
// with it (see k_moddown).
enum { KIND_BOTH = 0, KIND_INT = 1, KIND_DP = 2 };
template <int KIND> __device__ __forceinline__ Mod load_mod(const Tables &T, int mid)
{
    Mod m = T.mods[mid];
    if (KIND == KIND_INT) m.dp = 0;
    if (KIND == KIND_DP) m.dp = 1;
    return m;
}
template <int LOGN, int C, int KIND>
__global__ void __launch_bounds__(NttCfg<LOGN>::THREADS) k_ntt_inv(Tables T, const u64 *__restrict__ src, u64 *__restrict__ dst,
                                                                   size_t src_outer, size_t dst_outer, int L, int mod_base, int mode, InvFuse F)
{
    constexpr int c = C;
    constexpr int NL = 1 << LOGN;
    u64 *sm = dyn_smem();
    const int tid = threadIdx.x;
    const int w = blockIdx.x >> c, r = blockIdx.x & ((1 << c) - 1);
    const int mid = mod_base + (w % L);
    const Mod m = load_mod<KIND>(T, mid);
    const ulonglong2 *itw = T.itw + (size_t)mid * T.N;
    const u64 *in = src + (size_t)(w / L) * src_outer + (size_t)(w % L) * T.N + (size_t)r * NL;
    u64 x[16];
    TwRegs<LOGN, Sched<LOGN>::NP - 1> tl;
    load_tw_early<LOGN, Sched<LOGN>::NP - 1, true>(tl, itw, tid, (1 << c) + r);
    ulonglong2 fs = make_ulonglong2(0, 0);
    u64 ffix = 0;
    if constexpr (C == 0 && KIND != KIND_INT) {
        if (F.add && m.dp) {
            // FP64-domain modulus: the fused pre- and post-processing (see InvFuse) stay in the domain -- scaling of the
            // accumulator, the split lift of the rounded special-prime limb, its scaling and the subtraction -- and
            // the result leaves it once, in the final store
            constexpr int NPL = Sched<LOGN>::NP - 1;
            const double nq = m.dnq, sd = dp_from(T.qinv[(size_t)F.x * T.M + mid].x), sq = __dmul_rn(sd, m.dqinv);
            const double fixd = dp_from(m.q - T.halfmod[(size_t)F.x * T.M + mid]), wq30 = 1073741824.0 * m.dqinv;
            const double half = mode == INV_ADDHALF ? dp_from(m.q >> 1) : 0.0;
            const u64 *ad = F.add + (size_t)(w / F.P) * F.add_ct_stride + (size_t)(w % F.P) * F.add_poly_stride;
            for_pairs_co(tid, [&](int reg, int e) {
                const ulonglong2 v = ldg2(in + e), a = ldg2(ad + e);
                x[reg] = as_u(__dadd_rn(dp_mul(dp_from(v.x), sd, sq, nq), dp_from(a.x)));
                x[reg + 1] = as_u(__dadd_rn(dp_mul(dp_from(v.y), sd, sq, nq), dp_from(a.y)));
            });
            co_to_contig(x, sm, tid);
            ntt_inv_regs_split<LOGN, true, false, NPL, true, true>(x, sm, itw, m, tid, 0, 0, tl);
            u64 *outp = dst + (size_t)(w / L) * dst_outer + (size_t)(w % L) * T.N;
            const u64 *sb = F.sub + (size_t)w * T.N;
            for_pairs_strided<LOGN>(tid, [&](int reg, int e) {
                const ulonglong2 u = ldg2(sb + e);
                const double t0 = dp_mul(__dadd_rn(lift_dp<true>(u.x, wq30, nq), fixd), sd, sq, nq);
                const double t1 = dp_mul(__dadd_rn(lift_dp<true>(u.y, wq30, nq), fixd), sd, sq, nq);
                st2(outp + e, dp_canon(__dadd_rn(__dadd_rn(as_d(x[reg]), -t0), half), m), dp_canon(__dadd_rn(__dadd_rn(as_d(x[reg + 1]), -t1), half), m));
            });
            return;
        }
    }
    if (F.add) {
        fs = T.qinv[(size_t)F.x * T.M + mid];
        ffix = m.q - T.halfmod[(size_t)F.x * T.M + mid];
        const u64 *ad = F.add + (size_t)(w / F.P) * F.add_ct_stride + (size_t)(w % F.P) * F.add_poly_stride + (size_t)r * NL;
        for_pairs_co(tid, [&](int reg, int e) {
            const ulonglong2 v = ldg2(in + e), a = ldg2(ad + e);
            x[reg] = add_mod(shoup(v.x, fs.x, fs.y, m.q), a.x, m.q);
            x[reg + 1] = add_mod(shoup(v.y, fs.x, fs.y, m.q), a.y, m.q);
        });
    } else {
        for_pairs_co(tid, [&](int reg, int e) {
            ulonglong2 v = ldg2(in + e);
            x[reg] = v.x;
            x[reg + 1] = v.y;
        });
    }
    co_to_contig(x, sm, tid);
    u64 *out = dst + (size_t)(w / L) * dst_outer + (size_t)(w % L) * T.N + (size_t)r * NL;
    if (c == 0)
        ntt_inv_regs_split<LOGN, true>(x, sm, itw, m, tid, 0, 0, tl);
    else {
        ntt_inv_regs_split<LOGN, false>(x, sm, itw, m, tid, c, r, tl);
        reduce_all(x, m);
        cross_inv<LOGN>(x, sm, c, r, tid, itw, m);
    }
    // x: finished values in [0, 2q), pass-0 layout of chunk r
    if (F.add) {
        const u64 *sb = F.sub + (size_t)w * T.N + (size_t)r * NL;
        for_pairs_strided<LOGN>(tid, [&](int reg, int e) {
            const ulonglong2 u = ldg2(sb + e);
            const u64 v0 = inv_post(csub(x[reg], m.q), u.x, m, fs, ffix), v1 = inv_post(csub(x[reg + 1], m.q), u.y, m, fs, ffix);
            st2(out + e, mode == INV_ADDHALF ? csub(v0 + (m.q >> 1), m.q) : v0, mode == INV_ADDHALF ? csub(v1 + (m.q >> 1), m.q) : v1);
        });
        return;
    }
    for_pairs_strided<LOGN>(tid, [&](int reg, int e) { st2(out + e, inv_finish(x[reg], m, mode), inv_finish(x[reg + 1], m, mode)); });
}
// ------------------------------------------------------------------------------------ K6 step 2
// acc[b][k][I] = sum_J NTT_{q_I}(t[b][J] mod q_I) (.) key[J][k][I]      (I == L  <->  special prime)
// CKKS: the I == J term reuses the NTT-form target.  grid = B * (L+1) << c.
//
// One CTA owns (ciphertext b, output modulus I[, chunk r]) and loops over the L digits: the digit is
// lifted and transformed in registers, then multiplied into both key components.  The keys carry
// their Shoup quotients (interleaved at upload by k_shoup_quotients), so a multiply-
// accumulate is one shoup_mad -- 10 integer multiply-adds, valid for ANY 64-bit digit value (the
// transform output needs no reduction) -- and the accumulators stay lazy (Mod::acc_period).  The two
// accumulator limbs live in shared memory between digits ([p][tid] pairs, conflict-free 128-bit
// accesses) so the transform has the whole register file.
// The CTA (cluster) of the special prime finishes with the inverse transform and the "+ q_sp/2" rounding of
// its two accumulators and writes them straight to rp (the input of k_moddown); that limb never goes to HBM
// in NTT form.
struct KsInnerArgs {
    const u64 *tcoef;      // target in coefficient form: tcoef + b*tcoef_stride + J*N
    size_t tcoef_stride;
    const u64 *target;     // NTT-form target (CKKS) or nullptr (BFV): target + b*target_stride + J*N
    size_t target_stride;
    const u64 *key;        // [Ltop][2][K] limbs of 2N words: keys interleaved with their Shoup quotients (k_shoup_quotients)
    u64 *acc;              // [B][2][L+1][N]
    u64 *rp;               // [B][2][N]: rounded special-prime limb in coefficient form
    int L, K, B;           // B = ciphertexts in this launch
};
// lazy accumulator of the inner product (either domain) -> canonical residue
__device__ __forceinline__ u64 acc_finish(u64 a, const Mod &m) { return m.dp ? dp_canon(as_d(a), m) : reduce_full(a, m); }
template <int LOGN> struct KsCfg {
    static constexpr int SMEM_BYTES = 3 * NttCfg<LOGN>::SMEM_BYTES;   // transform buffer + 2 accumulator limbs
};
// DP = Mod::dp of the CTA's modulus as a compile-time constant: the kernel branches once, at the top, into one of two
// complete instances of the body, so the integer and the FP64-domain code never share live ranges (with the branch
// inside the multiply-accumulate loop the register allocator spilled in both).
template <int LOGN, int C, bool DP>
__device__ __forceinline__ void ks_inner_body(const Tables &T, const KsInnerArgs &A, Mod m, int b, int I, int ki, int r)
{
    constexpr int c = C;
    constexpr int NL = 1 << LOGN, TH = NttCfg<LOGN>::THREADS;
    m.dp = DP;
    u64 *sm = dyn_smem();
    u64 *acc_sm[2] = { sm + NL, sm + 2 * NL };
    const int tid = threadIdx.x;
    const int L = A.L;
    const ulonglong2 *tw = T.tw + (size_t)ki * T.N;
    const size_t N = T.N, off = (size_t)r * NL;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        st2(acc_sm[0] + (p * TH + tid) * 2, 0, 0);
        st2(acc_sm[1] + (p * TH + tid) * 2, 0, 0);
    }
    for (int J = 0; J < L; J++) {
        u64 x[16];
        // device key layout (key_word_index): per limb and chunk [p][tid][k(e), k(e+1), k'(e), k'(e+1)], e = 16 tid + 2p,
        // k' = Shoup quotient; FP64-domain limbs [p][tid][k(e), k(e+1)] as doubles (first half of the limb's 2N-word
        // slot): a warp's loads cover 1 KiB / 512 B of contiguous memory per p
        const u64 *kp0 = A.key + 2 * ((((size_t)J * 2 + 0) * A.K + ki) * N + off) + (DP ? 2 : 4) * tid;
        const u64 *kp1 = A.key + 2 * ((((size_t)J * 2 + 1) * A.K + ki) * N + off) + (DP ? 2 : 4) * tid;
        ulonglong2 kd0[8];   // DP: key component 0, requested before the last pass of the transform
        auto prefetch_key0 = [&]() {
            if constexpr (DP) {
#pragma unroll
                for (int p = 0; p < 8; p++) kd0[p] = ldg2(kp0 + (size_t)p * 2 * TH);
            }
        };
        if (A.target && I == J) {
            const u64 *tp = A.target + (size_t)b * A.target_stride + (size_t)J * N + off;
            for_pairs_co(tid, [&](int reg, int e) {
                ulonglong2 v = ldg2(tp + e);
                x[reg] = v.x;
                x[reg + 1] = v.y;
            });
            prefetch_key0();
            __syncthreads();   // the transform buffer may still be read by the previous digit's transform
            co_to_contig(x, sm, tid);
            if (m.dp) to_dp_all(x);
        } else {
            const u64 *tp = A.tcoef + (size_t)b * A.tcoef_stride + (size_t)J * N;
            TwRegs<LOGN, 0> t0;
            load_tw_early<LOGN, 0, false>(t0, tw, tid, (1 << c) + r);
            if constexpr (DP) {   // FP64 domain: digits of moduli up to 48 bits are lazy values as they are (2^48 + 14 q < 2^50)
                if (T.mods[J].bits > 48)
                    load_fwd_split<LOGN, true>(x, tp, c, r, tid, tw, m, PreLiftDp{ 1073741824.0 * m.dqinv, m.dnq }, sm);
                else
                    load_fwd_split<LOGN, true>(x, tp, c, r, tid, tw, m, PreNone(), sm);
            } else if (lift_wide(T.mods[J].q, m.q))
                load_fwd_split<LOGN, true>(x, tp, c, r, tid, tw, m, PreReduce<true>{ m.q, m.r64 }, sm);
            else if (T.mods[J].q > m.q)
                load_fwd_split<LOGN, true>(x, tp, c, r, tid, tw, m, PreReduce<false>{ m.q, m.r64 }, sm);
            else
                load_fwd_split<LOGN, true>(x, tp, c, r, tid, tw, m, PreNone(), sm);
            ntt_fwd_regs_split<LOGN, true>(x, sm, tw, m, tid, c, r, t0, prefetch_key0);
        }
        // x: the digit in NTT form, lazy (any 64-bit value congruent to it, or an FP64-domain value of magnitude < 14 q)
        const bool fold = ((J + 1) % (int)m.acc_period) == 0;
        if constexpr (DP) {
            // FP64 domain: the key is one double per coefficient; the quotient estimate RN(k / q) that dp_mul wants is
            // replaced by RN(k * RN(1/q)) computed here (relative error 2^-52 instead of 2^-53: the result stays below
            // 0.75 q for |x| < 2^50), which halves the key bytes streamed from L2.  Accumulators grow by 0.75 q per digit.
            ulonglong2 kd1[8];
#pragma unroll
            for (int p = 0; p < 8; p++) kd1[p] = ldg2(kp1 + (size_t)p * 2 * TH);   // in flight during component 0
#pragma unroll
            for (int k = 0; k < 2; k++) {
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    const ulonglong2 kv = k ? kd1[p] : kd0[p];
                    ulonglong2 a = ld2(acc_sm[k] + (p * TH + tid) * 2);
                    const double k0 = as_d(kv.x), k1 = as_d(kv.y);
                    double a0 = __dadd_rn(as_d(a.x), dp_mul(as_d(x[2 * p]), k0, __dmul_rn(k0, m.dqinv), m.dnq));
                    double a1 = __dadd_rn(as_d(a.y), dp_mul(as_d(x[2 * p + 1]), k1, __dmul_rn(k1, m.dqinv), m.dnq));
                    if (fold) {
                        a0 = dp_reduce(a0, m.dqinv, m.dnq);
                        a1 = dp_reduce(a1, m.dqinv, m.dnq);
                    }
                    st2(acc_sm[k] + (p * TH + tid) * 2, as_u(a0), as_u(a1));
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const u64 *kp = k ? kp1 : kp0;
                ulonglong2 kv[8], ks[8], a[8];
#pragma unroll
                for (int p = 0; p < 8; p++) {   // all 16 key loads in flight before the first use
                    kv[p] = ldg2(kp + (size_t)p * 4 * TH);
                    ks[p] = ldg2(kp + (size_t)p * 4 * TH + 2);
                }
#pragma unroll
                for (int p = 0; p < 8; p++) a[p] = ld2(acc_sm[k] + (p * TH + tid) * 2);
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    a[p].x = shoup_mad(x[2 * p], kv[p].x, ks[p].x, m.nq, a[p].x);
                    a[p].y = shoup_mad(x[2 * p + 1], kv[p].y, ks[p].y, m.nq, a[p].y);
                }
                if (fold) {
#pragma unroll
                    for (int p = 0; p < 8; p++) {
                        a[p].x = reduce_lazy(a[p].x, m);
                        a[p].y = reduce_lazy(a[p].y, m);
                    }
                }
#pragma unroll
                for (int p = 0; p < 8; p++) st2(acc_sm[k] + (p * TH + tid) * 2, a[p].x, a[p].y);
            }
        }
        // no barrier here: the accumulator slots are thread-private, and the transform buffer is protected by the
        // barrier in front of its next first store (REUSE)
    }
    if (I == L) {
        const ulonglong2 *itw = T.itw + (size_t)ki * T.N;
#pragma unroll 1
        for (int k = 0; k < 2; k++) {
            u64 x[16];
            TwRegs<LOGN, Sched<LOGN>::NP - 1> tl;
            load_tw_early<LOGN, Sched<LOGN>::NP - 1, true>(tl, itw, tid, (1 << c) + r);
            for_pairs_contig(tid, [&](int reg, int e) {
                const ulonglong2 a = ld2(acc_sm[k] + ((reg >> 1) * TH + tid) * 2);
                x[reg] = acc_finish(a.x, m);
                x[reg + 1] = acc_finish(a.y, m);
            });
            if (c == 0)
                ntt_inv_regs_split<LOGN, true, true>(x, sm, itw, m, tid, 0, 0, tl);
            else {
                ntt_inv_regs_split<LOGN, false, true>(x, sm, itw, m, tid, c, r, tl);
                reduce_all(x, m);
                cross_inv<LOGN>(x, sm, c, r, tid, itw, m);
            }
            u64 *out = A.rp + ((size_t)b * 2 + k) * N + off;
            for_pairs_strided<LOGN>(tid, [&](int reg, int e) { st2(out + e, inv_finish(x[reg], m, INV_ADDHALF), inv_finish(x[reg + 1], m, INV_ADDHALF)); });
        }
        return;
    }
    __syncthreads();   // the epilogue stages through the transform buffer
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        u64 *o = A.acc + (((size_t)b * 2 + k) * (L + 1) + I) * N + off;
        u64 x[16];
        for_pairs_contig(tid, [&](int reg, int) {
            const ulonglong2 a = ld2(acc_sm[k] + ((reg >> 1) * TH + tid) * 2);
            x[reg] = acc_finish(a.x, m);
            x[reg + 1] = acc_finish(a.y, m);
        });
        contig_to_co(x, sm, tid);
        for_pairs_co(tid, [&](int reg, int e) { st2(o + e, x[reg], x[reg + 1]); });
        warp_sync();   // the slice is rewritten by the next component
    }
}
template <int LOGN, int C>
__global__ void __launch_bounds__(NttCfg<LOGN>::THREADS, 1) k_ks_inner(Tables T, KsInnerArgs A)
{
    const int r = blockIdx.x & ((1 << C) - 1);
    const int unit = blockIdx.x >> C;
    // the special-prime units (longest: they also run the fused inverse transforms) are scheduled first
    const int L = A.L;
    const int b = unit < A.B ? unit : (unit - A.B) / L, I = unit < A.B ? L : (unit - A.B) % L;
    const int ki = (I == L) ? A.K - 1 : I;
    const Mod m = T.mods[ki];
    if (m.dp) ks_inner_body<LOGN, C, true>(T, A, m, b, I, ki, r);
    else ks_inner_body<LOGN, C, false>(T, A, m, b, I, ki, r);
}

// Device form of a key-switching key, built once per upload from SEAL's [Ltop][2][K][N] array: every limb becomes
// 2N words holding the key residues interleaved with their Shoup quotients floor(k * 2^64 / q) (FP64-domain moduli:
// the residues as doubles, N words) in the order
// k_ks_inner consumes them -- chunk r of NL = 16*TH coefficients, then [p][tid][k(e), k(e+1), k'(e), k'(e+1)] with
// e = 16 tid + 2p -- so the inner product reads the key with fully coalesced 128-bit loads.
// (restoring division, 64 steps: k < q < 2^61 so the running remainder never overflows).
__host__ __device__ __forceinline__ size_t key_word_index(size_t limb, size_t e, size_t N, int lognl)
{
    const size_t NL = (size_t)1 << lognl, TH = NL / 16;
    const size_t r = e >> lognl, el = e & (NL - 1), tid = el >> 4, p = (el & 15) >> 1, h = el & 1;
    return 2 * (limb * N + r * NL) + (p * TH + tid) * 4 + h;
}
__global__ void __launch_bounds__(256) k_shoup_quotients(Tables T, const u64 *__restrict__ key, u64 *__restrict__ dkey, int K, size_t words, int lognl)
{
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= words) return;
    const Mod &m = T.mods[(gid / T.N) % K];
    const u64 q = m.q;
    const size_t o = key_word_index(gid / T.N, gid % T.N, T.N, lognl);
    u64 rem = key[gid], quot = 0;
    if (m.dp) {   // FP64-domain modulus: the residue as a double, [p][tid][2] in the first half of the limb's slot
        const size_t e = gid % T.N, NL = (size_t)1 << lognl, TH = NL / 16, el = e & (NL - 1);
        dkey[2 * ((gid / T.N) * T.N + (e >> lognl) * NL) + (((el & 15) >> 1) * TH + (el >> 4)) * 2 + (el & 1)] = as_u(dp_from(rem));
        return;
    }
    dkey[o] = rem;
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        rem <<= 1;
        quot <<= 1;
        if (rem >= q) {
            rem -= q;
            quot |= 1;
        }
    }
    dkey[o + 2] = quot;
}

// ------------------------------------------------------------------------------------ K6 step 3 / K9
// out[b][p][j] = (addend[b][p][j] +) (base[b][p][j] - NTT_{q_j}(rp[b][p] mod q_j + fix)) * q_x^{-1} mod q_j
// rp = rounded last limb in coefficient form ((iNTT(last) + q_x/2) mod q_x), x = modulus id of the dropped prime.
// Key switch:  base = acc (stride over L+1 limbs), addend = input ct component (or none), x = K-1.
// Rescale:     base = input ct, addend = none, x = L-1.       grid = B * P * nJ << c.
__device__ __forceinline__ int nth_set_bit(unsigned mask, int n)   // index of the n-th (0-based) set bit
{
#if defined(__CUDA_ARCH__)
    return (int)__fns(mask, 0, n + 1);
#else
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) {
            if (n == 0) return i;
            n--;
        }
    return 0;
#endif
}
struct ModDownArgs {
    const u64 *rp;         // [B][P][N]  (k_moddown_coeff: may be nullptr, then rp_raw is used)
    const u64 *rp_raw;     // un-rounded last limb, coefficient form: rp_raw + (b*P + p)*rp_raw_stride
    size_t rp_raw_stride;
    const u64 *base;       // base + b*base_ct_stride + p*base_poly_stride + j*N
    size_t base_ct_stride, base_poly_stride;
    const u64 *addend[2];  // per poly (P <= 2 when addend used) or nullptr; addend[p] + b*add_ct_stride + j*N
    size_t add_ct_stride;
    u64 *out;              // out + b*out_ct_stride + p*out_poly_stride + j*N
    size_t out_ct_stride, out_poly_stride;
    int P, nJ, x;          // x = modulus id of the dropped prime
    // limbs handled by this launch: bit j of jmask, nJsub = popcount(jmask) (the host launches the integer kernel for
    // the limbs of wide moduli and the FP64 kernel for the others; jmask = 2^nJ - 1 when there is only one kind)
    unsigned jmask;
    int nJsub;
    // fused relinearize + rescale: rp2 = rounded last data limb (coefficient form, [B][P][N]) of the key-switched
    // ciphertext, x2 = its modulus id.  out = (base * s + addend) * r - NTT((u1 * s + u2) * r), s = q_x^{-1}, r = q_x2^{-1}
    const u64 *rp2;
    int x2;
};
// Shared memory: transform buffer | TMA landing zone of the accumulator tile | of the addend tile | mbarrier.
// The two epilogue operands of a CTA are contiguous 8 NL-byte tiles; thread 0 starts their bulk copies before the
// transform and the epilogue reads them from shared memory.
template <int LOGN> struct ModDownCfg {
    static constexpr int SMEM_BYTES = 3 * NttCfg<LOGN>::SMEM_BYTES + 16;
};
// DP = Mod::dp of the output limb's modulus, a compile-time constant (one branch at the top of the kernel, two complete
// instances, as in k_ks_inner).  In the FP64 instance the lifted input, the transform, and the epilogue's two constant
// multiplies stay in the FP64 domain; the result leaves it once, in the final store.
template <int LOGN, int C, bool DP>
__device__ __forceinline__ void moddown_body(const Tables &T, const ModDownArgs &A)
{
    constexpr int c = C;
    constexpr int NL = 1 << LOGN;
    u64 *sm = dyn_smem();
    u64 *stage_b = sm + NL, *stage_a = sm + 2 * NL, *bar = sm + 3 * NL;
    const int tid = threadIdx.x;
    const int r = blockIdx.x & ((1 << c) - 1);
    int unit = blockIdx.x >> c;
    const int j = nth_set_bit(A.jmask, unit % A.nJsub);
    unit /= A.nJsub;
    const int p = unit % A.P, b = unit / A.P;
    Mod m = T.mods[j];
    m.dp = DP;
    const ulonglong2 *tw = T.tw + (size_t)j * T.N;
    const size_t N = T.N, off = (size_t)r * NL;
    const u64 fix = m.q - T.halfmod[(size_t)A.x * T.M + j];
    const ulonglong2 qi = T.qinv[(size_t)A.x * T.M + j];
    u64 x[16];
    TwRegs<LOGN, 0> t0;
    load_tw_early<LOGN, 0, false>(t0, tw, tid, (1 << c) + r);
    const u64 *bp = A.base + (size_t)b * A.base_ct_stride + (size_t)p * A.base_poly_stride + (size_t)j * N + off;
    const u64 *ap = A.addend[p] ? A.addend[p] + (size_t)b * A.add_ct_stride + (size_t)j * N + off : nullptr;
    u64 *op = A.out + (size_t)b * A.out_ct_stride + (size_t)p * A.out_poly_stride + (size_t)j * N + off;
    // epilogue operands: bulk copies (TMA) into shared memory, in flight during the transform.  out may alias the
    // addend element for element (b200he_apply_galois): this CTA is the only one that touches its tile, and it has
    // read the whole tile before it writes.
    // (thread 0 initialises, arms and uses the barrier; everyone else first touches it after the CTA-wide barriers of
    // the transform, which order the initialisation before their wait)
    if (tid == 0) {
        tma_bar_init(bar);
        tma_bar_expect(bar, (ap ? 2u : 1u) * NL * 8);
        tma_load_1d(stage_b, bp, NL * 8, bar);
        if (ap) tma_load_1d(stage_a, ap, NL * 8, bar);
    }
    if constexpr (DP) {
        const double wq30 = 1073741824.0 * m.dqinv, nq = m.dnq;
        const double sd = dp_from(qi.x), sq = __dmul_rn(sd, m.dqinv);
        const u64 *rp = A.rp + ((size_t)b * A.P + p) * N;
        const bool w1 = T.mods[A.x].bits > 48;
        if (A.rp2) {
            const double rd = dp_from(T.qinv[(size_t)A.x2 * T.M + j].x), rq = __dmul_rn(rd, m.dqinv);
            const double fix2 = dp_from(m.q - T.halfmod[(size_t)A.x2 * T.M + j]);
            const u64 *rp2 = A.rp2 + ((size_t)b * A.P + p) * N;
            if (w1 && T.mods[A.x2].bits <= 48)   // the usual case: special prime wide, last data prime narrow
                load_fwd_split<LOGN>(x, rp, c, r, tid, tw, m, PreTwoDp<true, false>{ wq30, nq, dp_from(fix), fix2, sd, sq, rd, rq, rp2 }, sm);
            else
                load_fwd_split<LOGN>(x, rp, c, r, tid, tw, m, PreTwoDp<true, true>{ wq30, nq, dp_from(fix), fix2, sd, sq, rd, rq, rp2 }, sm);
            ntt_fwd_regs_split<LOGN>(x, sm, tw, m, tid, c, r, t0);
            contig_to_co(x, sm, tid);   // FP64-domain values, |x| < 14 q
            tma_bar_wait(bar, 0);
            for_pairs_co(tid, [&](int reg, int e) {
                const ulonglong2 bv = ld2(stage_b + e), av = ld2(stage_a + e);
                const double g0 = __dadd_rn(dp_mul(dp_from(bv.x), sd, sq, nq), dp_from(av.x)), g1 = __dadd_rn(dp_mul(dp_from(bv.y), sd, sq, nq), dp_from(av.y));
                const double h0 = dp_mul(g0, rd, rq, nq), h1 = dp_mul(g1, rd, rq, nq);
                st2(op + e, dp_canon(__dadd_rn(h0, -as_d(x[reg])), m), dp_canon(__dadd_rn(h1, -as_d(x[reg + 1])), m));
            });
            return;
        }
        if (w1) load_fwd_split<LOGN>(x, rp, c, r, tid, tw, m, PreReduceFixDp<true>{ wq30, nq, dp_from(fix) }, sm);
        else load_fwd_split<LOGN>(x, rp, c, r, tid, tw, m, PreReduceFixDp<false>{ wq30, nq, dp_from(fix) }, sm);
        ntt_fwd_regs_split<LOGN>(x, sm, tw, m, tid, c, r, t0);
        contig_to_co(x, sm, tid);
        tma_bar_wait(bar, 0);
        for_pairs_co(tid, [&](int reg, int e) {
            const ulonglong2 bv = ld2(stage_b + e);
            double v0 = dp_mul(__dadd_rn(dp_from(bv.x), -as_d(x[reg])), sd, sq, nq), v1 = dp_mul(__dadd_rn(dp_from(bv.y), -as_d(x[reg + 1])), sd, sq, nq);
            if (ap) {
                const ulonglong2 av = ld2(stage_a + e);
                v0 = __dadd_rn(v0, dp_from(av.x));
                v1 = __dadd_rn(v1, dp_from(av.y));
            }
            st2(op + e, dp_canon(v0, m), dp_canon(v1, m));
        });
        return;
    }
    if (A.rp2) {
        const ulonglong2 ri = T.qinv[(size_t)A.x2 * T.M + j];
        const u64 fix2 = m.q - T.halfmod[(size_t)A.x2 * T.M + j];
        if (lift_wide(T.mods[A.x].q, m.q) || lift_wide(T.mods[A.x2].q, m.q))
            load_fwd_split<LOGN>(x, A.rp + ((size_t)b * A.P + p) * N, c, r, tid, tw, m,
                                 PreTwo<true>{ m.q, m.r64, fix, fix2, qi, ri, A.rp2 + ((size_t)b * A.P + p) * N }, sm);
        else
            load_fwd_split<LOGN>(x, A.rp + ((size_t)b * A.P + p) * N, c, r, tid, tw, m,
                                 PreTwo<false>{ m.q, m.r64, fix, fix2, qi, ri, A.rp2 + ((size_t)b * A.P + p) * N }, sm);
        ntt_fwd_regs_split<LOGN>(x, sm, tw, m, tid, c, r, t0);
        canon_all(x, m);
        contig_to_co(x, sm, tid);
        tma_bar_wait(bar, 0);
        for_pairs_co(tid, [&](int reg, int e) {
            const ulonglong2 bv = ld2(stage_b + e), av = ld2(stage_a + e);
            const u64 g0 = shoup_lazy(bv.x, qi.x, qi.y, m.q) + av.x, g1 = shoup_lazy(bv.y, qi.x, qi.y, m.q) + av.y;   // < 3q
            const u64 h0 = shoup(g0, ri.x, ri.y, m.q), h1 = shoup(g1, ri.x, ri.y, m.q);
            st2(op + e, sub_mod(h0, x[reg], m.q), sub_mod(h1, x[reg + 1], m.q));
        });
        return;
    }
    if (lift_wide(T.mods[A.x].q, m.q))
        load_fwd_split<LOGN>(x, A.rp + ((size_t)b * A.P + p) * N, c, r, tid, tw, m, PreReduceFix<true>{ m.q, m.r64, fix }, sm);
    else
        load_fwd_split<LOGN>(x, A.rp + ((size_t)b * A.P + p) * N, c, r, tid, tw, m, PreReduceFix<false>{ m.q, m.r64, fix }, sm);
    ntt_fwd_regs_split<LOGN>(x, sm, tw, m, tid, c, r, t0);
    canon_all(x, m);
    contig_to_co(x, sm, tid);
    tma_bar_wait(bar, 0);
    for_pairs_co(tid, [&](int reg, int e) {
        ulonglong2 bv = ld2(stage_b + e);
        u64 u0 = x[reg], u1 = x[reg + 1];
        u64 v0 = shoup(sub_mod(bv.x, u0, m.q), qi.x, qi.y, m.q);
        u64 v1 = shoup(sub_mod(bv.y, u1, m.q), qi.x, qi.y, m.q);
        if (ap) {
            ulonglong2 av = ld2(stage_a + e);
            v0 = add_mod(v0, av.x, m.q);
            v1 = add_mod(v1, av.y, m.q);
        }
        st2(op + e, v0, v1);
    });
}

// One kernel per instance, and a third with both for launches that mix limbs of the two kinds: inlined into one kernel,
// the FP64 instance costs the integer instance registers (spills 24 -> 92 bytes, +12 % on the all-integer mod-down of
// the C2 step), so launches of a single kind use a kernel that holds only their instance.
template <int LOGN, int C>
__global__ void __launch_bounds__(NttCfg<LOGN>::THREADS) k_moddown(Tables T, ModDownArgs A)
{
    moddown_body<LOGN, C, false>(T, A);
}
template <int LOGN, int C>
__global__ void __launch_bounds__(NttCfg<LOGN>::THREADS) k_moddown_dp(Tables T, ModDownArgs A)
{
    moddown_body<LOGN, C, true>(T, A);
}
// limbs of both kinds in one launch (splitting such a launch in two costs more in tails and launch gaps than the
// shared register allocation costs the integer instance)
template <int LOGN, int C>
__global__ void __launch_bounds__(NttCfg<LOGN>::THREADS) k_moddown_mix(Tables T, ModDownArgs A)
{
    if (T.mods[nth_set_bit(A.jmask, ((int)blockIdx.x >> C) % A.nJsub)].dp) moddown_body<LOGN, C, true>(T, A);
    else moddown_body<LOGN, C, false>(T, A);
}

// Coefficient-form variant (BFV key switch / BFV mod-switch): no transform, one thread per coefficient pair.
// base must already be in coefficient form.  grid covers B * P * nJ * N / 2 threads.
__global__ void __launch_bounds__(256) k_moddown_coeff(Tables T, ModDownArgs A, size_t B)
{
    const size_t N = T.N, per_ct = (size_t)A.P * A.nJ * N / 2;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * per_ct) return;
    const size_t b = gid / per_ct, rem = (gid % per_ct) * 2;
    const size_t limb = rem / N, e = rem % N;
    const int p = (int)(limb / A.nJ), j = (int)(limb % A.nJ);
    const Mod m = T.mods[j];
    const u64 fix = m.q - T.halfmod[(size_t)A.x * T.M + j];
    const ulonglong2 qi = T.qinv[(size_t)A.x * T.M + j];
    ulonglong2 r;
    if (A.rp) r = ld2(A.rp + (b * A.P + p) * N + e);
    else {
        const Mod mx = T.mods[A.x];
        r = ld2(A.rp_raw + (b * A.P + p) * A.rp_raw_stride + e);
        r.x = csub(r.x + (mx.q >> 1), mx.q);
        r.y = csub(r.y + (mx.q >> 1), mx.q);
    }
    const u64 u0 = csub(reduce64(r.x, m) + fix, m.q), u1 = csub(reduce64(r.y, m) + fix, m.q);
    const ulonglong2 bv = ld2(A.base + b * A.base_ct_stride + (size_t)p * A.base_poly_stride + (size_t)j * N + e);
    u64 v0 = shoup(sub_mod(bv.x, u0, m.q), qi.x, qi.y, m.q);
    u64 v1 = shoup(sub_mod(bv.y, u1, m.q), qi.x, qi.y, m.q);
    if (A.addend[p]) {
        const ulonglong2 av = ld2(A.addend[p] + b * A.add_ct_stride + (size_t)j * N + e);
        v0 = add_mod(v0, av.x, m.q);
        v1 = add_mod(v1, av.y, m.q);
    }
    st2(A.out + b * A.out_ct_stride + (size_t)p * A.out_poly_stride + (size_t)j * N + e, v0, v1);
}

// ------------------------------------------------------------------------------------ K3 / K4 / K11
// Elementwise kernels: one thread per coefficient pair; ciphertext i of the output pairs
// a[ai[i]] with b[bi[i]] (index maps express the reference's b0 x b1 result grid without copies).
enum { EW_ADD = 0, EW_SUB = 1, EW_MUL = 2 };
struct EwArgs {
    const u64 *a, *b;
    u64 *out;
    const u32 *ai, *bi;         // nullable
    size_t a_stride, b_stride, out_stride;   // words per ciphertext
    int polys, b_polys;         // polys in out/a; polys in b (1 = plaintext broadcast over polys / only c0 for add)
    int L, mod_base;
    size_t n;                   // ciphertexts
};
template <int OP> __global__ void __launch_bounds__(256) k_ew(Tables T, EwArgs A)
{
    const size_t N = T.N, per_ct = (size_t)A.polys * A.L * N / 2;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= A.n * per_ct) return;
    const size_t i = gid / per_ct, rem = (gid % per_ct) * 2;
    const size_t limb_idx = rem / N, e = rem % N;
    const int p = (int)(limb_idx / A.L), l = (int)(limb_idx % A.L);
    const Mod m = T.mods[A.mod_base + l];
    const size_t ia = A.ai ? A.ai[i] : i, ib = A.bi ? A.bi[i] : i;
    ulonglong2 va = ld2(A.a + ia * A.a_stride + rem);
    u64 *o = A.out + i * A.out_stride + rem;
    if (A.b_polys == 1 && p > 0 && OP != EW_MUL) {   // add_plain / sub_plain touch c0 only
        st2(o, va.x, va.y);
        return;
    }
    const size_t boff = (A.b_polys == 1 ? 0 : (size_t)p * A.L * N) + (size_t)l * N + e;
    ulonglong2 vb = ld2(A.b + ib * A.b_stride + boff);
    if (OP == EW_ADD) st2(o, add_mod(va.x, vb.x, m.q), add_mod(va.y, vb.y, m.q));
    else if (OP == EW_SUB) st2(o, sub_mod(va.x, vb.x, m.q), sub_mod(va.y, vb.y, m.q));
    else if (m.dp)
        st2(o, dp_canon(dp_mul_dd(dp_from(va.x), dp_from(vb.x), m.dqinv, m.dnq), m), dp_canon(dp_mul_dd(dp_from(va.y), dp_from(vb.y), m.dqinv, m.dnq), m));
    else st2(o, mul_mod(va.x, vb.x, m), mul_mod(va.y, vb.y, m));
}

// CKKS / NTT-domain tensor product (2 x 2 -> 3): reads 4 polys, writes 3, one pass.
__global__ void __launch_bounds__(256) k_tensor(Tables T, EwArgs A)
{
    const size_t N = T.N, LN = (size_t)A.L * N, per_ct = LN / 2;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= A.n * per_ct) return;
    const size_t i = gid / per_ct, rem = (gid % per_ct) * 2;
    const int l = (int)(rem / N);
    const Mod m = T.mods[A.mod_base + l];
    const size_t ia = A.ai ? A.ai[i] : i, ib = A.bi ? A.bi[i] : i;
    const u64 *pa = A.a + ia * A.a_stride + rem, *pb = A.b + ib * A.b_stride + rem;
    ulonglong2 a0 = ld2(pa), a1 = ld2(pa + LN), b0 = ld2(pb), b1 = ld2(pb + LN);
    u64 *o = A.out + i * A.out_stride + rem;
    if (m.dp) {   // FP64 domain (warp-uniform: a warp's coefficient pairs belong to one limb): 4 exact products per coefficient
        const double qi = m.dqinv, nq = m.dnq;
        const double x0 = dp_from(a0.x), x1 = dp_from(a1.x), y0 = dp_from(b0.x), y1 = dp_from(b1.x);
        const double u0 = dp_from(a0.y), u1 = dp_from(a1.y), v0 = dp_from(b0.y), v1 = dp_from(b1.y);
        st2(o, dp_canon(dp_mul_dd(x0, y0, qi, nq), m), dp_canon(dp_mul_dd(u0, v0, qi, nq), m));
        st2(o + LN, dp_canon(__dadd_rn(dp_mul_dd(x0, y1, qi, nq), dp_mul_dd(x1, y0, qi, nq)), m),
            dp_canon(__dadd_rn(dp_mul_dd(u0, v1, qi, nq), dp_mul_dd(u1, v0, qi, nq)), m));
        st2(o + 2 * LN, dp_canon(dp_mul_dd(x1, y1, qi, nq), m), dp_canon(dp_mul_dd(u1, v1, qi, nq), m));
        return;
    }
    st2(o, mul_mod(a0.x, b0.x, m), mul_mod(a0.y, b0.y, m));
    st2(o + LN, mad_mod(a0.x, b1.x, mul_mod(a1.x, b0.x, m), m), mad_mod(a0.y, b1.y, mul_mod(a1.y, b0.y, m), m));
    st2(o + 2 * LN, mul_mod(a1.x, b1.x, m), mul_mod(a1.y, b1.y, m));
}

// ------------------------------------------------------------------------------------ K8
// NTT-form Galois automorphism of a size-2 ciphertext batch: out0 = g(c0) -> dst ct poly 0,
// g(c1) -> target buffer [B][L][N]; dst poly 1 is produced by the key switch that follows.
struct GaloisArgs {
    const u64 *src;       // [B][2][L][N]
    u64 *dst0;            // g(c0): dst0 + b*dst_stride + l*N
    u64 *dst1;            // g(c1): dst1 + b*L*N + l*N
    size_t src_stride, dst_stride;
    const u32 *table;     // [N] NTT-form permutation (CKKS)
    u32 elt;              // Galois element (BFV coefficient form)
    int L, logn;
    size_t B;
    int add_input;        // dst0 = g(c0) + c0 (rotate-and-add of accumulate: the key switch then adds c1 as its second addend)
};
__global__ void __launch_bounds__(256) k_galois_ntt(Tables T, GaloisArgs A)
{
    const size_t N = T.N, per_ct = 2 * (size_t)A.L * N;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= A.B * per_ct) return;
    const size_t b = gid / per_ct, rem = gid % per_ct;
    const size_t p = rem / ((size_t)A.L * N), le = rem % ((size_t)A.L * N), l = le / N, e = le % N;
    const u64 v = A.src[b * A.src_stride + p * A.L * N + l * N + A.table[e]];
    if (p == 0) A.dst0[b * A.dst_stride + le] = A.add_input ? add_mod(v, A.src[b * A.src_stride + le], T.mods[l].q) : v;
    else A.dst1[b * A.L * N + le] = v;
}
// Coefficient-form automorphism (BFV): coefficient i moves to i*elt mod N, negated when floor(i*elt/N) is odd.
__global__ void __launch_bounds__(256) k_galois_coeff(Tables T, GaloisArgs A)
{
    const size_t N = T.N, per_ct = 2 * (size_t)A.L * N;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= A.B * per_ct) return;
    const size_t b = gid / per_ct, rem = gid % per_ct;
    const size_t p = rem / ((size_t)A.L * N), le = rem % ((size_t)A.L * N), l = le / N, i = le % N;
    const u64 q = T.mods[l].q;
    u64 v = A.src[b * A.src_stride + rem];
    const u64 raw = (u64)i * A.elt;
    const size_t idx = raw & (N - 1);
    if ((raw >> A.logn) & 1) v = v ? q - v : 0;
    if (p == 0) A.dst0[b * A.dst_stride + l * N + idx] = A.add_input ? add_mod(v, A.src[b * A.src_stride + l * N + idx], q) : v;
    else A.dst1[b * A.L * N + l * N + idx] = v;
}

// strided copy of polys/limbs (mod_switch_drop_to_next, ciphertext gather): out[i][p][l] = in[idx[i]][p][l], l < L_out
struct CopyArgs {
    const u64 *src;
    u64 *dst;
    const u32 *idx;       // gather map (source ciphertext of output i), nullable
    const u32 *dst_idx;   // scatter map (destination ciphertext of item i), nullable
    size_t src_stride, dst_stride;
    int polys, L_in, L_out;
    size_t n;
};
__global__ void __launch_bounds__(256) k_copy_limbs(Tables T, CopyArgs A)
{
    const size_t N = T.N, per_ct = (size_t)A.polys * A.L_out * N / 2;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= A.n * per_ct) return;
    const size_t i = gid / per_ct, rem = (gid % per_ct) * 2;
    const size_t limb_idx = rem / N, e = rem % N, p = limb_idx / A.L_out, l = limb_idx % A.L_out;
    const size_t is = A.idx ? A.idx[i] : i, id = A.dst_idx ? A.dst_idx[i] : i;
    ulonglong2 v = ld2(A.src + is * A.src_stride + (p * A.L_in + l) * N + e);
    st2(A.dst + id * A.dst_stride + rem, v.x, v.y);
}

// out[0] = sum_i in[i] over a batch (collapse of per-sample ciphertexts, R/src/engine/seal_context.cpp:397-400):
// one thread per coefficient pair walks the batch; modular adds commute, so any order gives the reference's bits.
__global__ void __launch_bounds__(256) k_batch_sum(Tables T, const u64 *__restrict__ src, u64 *__restrict__ dst, size_t n, int polys, int L)
{
    const size_t N = T.N, per_ct = (size_t)polys * L * N / 2;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_ct) return;
    const size_t rem = gid * 2;
    const u64 q = T.mods[(rem / N) % L].q;
    u64 s0 = 0, s1 = 0;
    for (size_t i = 0; i < n; i++) {
        const ulonglong2 v = ld2(src + i * 2 * per_ct + rem);
        s0 = add_mod(s0, v.x, q);
        s1 = add_mod(s1, v.y, q);
    }
    st2(dst + rem, s0, s1);
}

}   // namespace b200he
