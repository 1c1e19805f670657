// kernels_ntt.cuh -- K1 / K2: the batched transform kernels k_ntt_fwd, k_ntt_fwd_p, k_ntt_inv (instantiated in tu_ntt.cu).
#pragma once
#include "kernels_common.cuh"

namespace b200he {

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
    load_fwd_split<LOGN, false, true>(x, src + (size_t)(w / L) * src_outer + (size_t)(w % L) * T.N, c, r, tid, tw, m, PreNone(), sm);
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
// Optional fusion around the inverse transform of limb w, used by the fused relinearize + rescale for the last data
// limb j.  The input is y = acc * s + c (NTT form; k_ks_inner writes it, KsInnerArgs::fuse_add); this
// kernel computes
//   iNTT(y) - ((sub[w] mod q_j) + fix) * s      s = q_x^{-1} mod q_j, fix = q_j - (q_x / 2 mod q_j)
// -- the mod-down correction applied in coefficient form: the transform is linear, so NTT(u) never has to be computed
// for this limb -- followed by the usual finish (INV_ADDHALF: + q_j / 2, the rounding of the rescale that follows).
struct InvFuse {
    const u64 *sub;        // [nlimbs][N] coefficient form (the rounded special-prime limb); nullptr: plain transform
    int x;                 // modulus id of the prime dropped by the key switch
    // Galois automorphism applied on load (plain transform only): dst = iNTT(g(src)), g(src)[i] = src[gal[i]] -- the
    // target of a rotation's key switch enters the key switch without a permutation pass
    const u32 *gal;        // [N] or nullptr
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
// (Measured on the B200 and dropped, C2 step, both k_ntt_inv launches together 0.358 ms with the plain loads below:
//  the epilogue operand staged by a TMA bulk copy issued before the transform, as in k_moddown: 0.372 ms; a persistent
//  variant with both input tiles of the next limb prefetched by TMA: 0.371 ms.  The warps of this kernel already overlap
//  one another's load and compute phases; the extra CTA-wide synchronisation costs more than the hidden latency.)
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
        if (F.sub && m.dp) {
            // FP64-domain modulus: the fused post-processing (see InvFuse) stays in the domain -- the split lift of the
            // rounded special-prime limb, its scaling and the subtraction -- and the result leaves it once, in the final store
            constexpr int NPL = Sched<LOGN>::NP - 1;
            const double nq = m.dnq, sd = dp_from(T.qinv[(size_t)F.x * T.M + mid].x), sq = __dmul_rn(sd, m.dqinv);
            const double fixd = dp_from(m.q - T.halfmod[(size_t)F.x * T.M + mid]), wq30 = 1073741824.0 * m.dqinv;
            const double half = mode == INV_ADDHALF ? dp_from(m.q >> 1) : 0.0;
            for_pairs_co(tid, [&](int reg, int e) {
                const ulonglong2 v = ldg2(in + e);
                x[reg] = as_u(dp_from(v.x));
                x[reg + 1] = as_u(dp_from(v.y));
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
    if (F.sub) {
        fs = T.qinv[(size_t)F.x * T.M + mid];
        ffix = m.q - T.halfmod[(size_t)F.x * T.M + mid];
    }
    if (F.gal) {
        gather_pairs_co(x, in - (size_t)r * NL, F.gal + (size_t)r * NL, tid);
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
    if (F.sub) {
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

}   // namespace b200he
