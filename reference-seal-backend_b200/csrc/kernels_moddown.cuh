// kernels_moddown.cuh -- K6 step 3 / K9: the NTT-form mod-down kernels (instantiated in tu_moddown.cu).
#pragma once
#include "kernels_common.cuh"

namespace b200he {

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
    const u64 *addend[3];  // per poly (a ciphertext has at most 3) or nullptr; addend[p] + b*add_ct_stride + j*N
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
    // Galois automorphism fused into the addend of polynomial 0 (K8): out[b][0][j] += g(gal_src[b][j]),
    // g(x)[i] = x[gal[i]] -- the g(c0) term of apply_galois_inplace, gathered from the INPUT ciphertext instead of
    // being written by a permutation pass.  gal_chunk[r] = the 2^LOGN-coefficient chunk of the source limb that chunk r
    // of the output gathers from (the table maps aligned chunks onto aligned chunks).
    const u32 *gal;        // [N] or nullptr
    const u64 *gal_src;    // gal_src + b*gal_ct_stride + j*N
    size_t gal_ct_stride;
    unsigned char gal_chunk[4];
};
// Shared memory: transform buffer | TMA landing zone of the accumulator tile | of the addend tile | mbarrier.
// The two epilogue operands of a CTA are contiguous 8 NL-byte tiles; thread 0 starts their bulk copies before the
// transform and the epilogue reads them from shared memory.
template <int LOGN> struct ModDownCfg {
    static constexpr int SMEM_BYTES = 3 * NttCfg<LOGN>::SMEM_BYTES + 16;
    static constexpr int SMEM_BYTES_GAL = SMEM_BYTES + 4 * (1 << LOGN);   // + this chunk of the Galois table (u32)
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
    // twiddles of the first local pass: next to the data loads for unsplit limbs; for clusters from inside the cross stages
    // (before their second barrier), so that they are not live across the radix-2/4 butterflies (as in k_ks_inner)
    TwRegs<LOGN, 0> t0;
    auto load_t0 = [&]() { load_tw_early<LOGN, 0, false>(t0, tw, tid, (1 << c) + r); };
    if constexpr (C == 0) load_t0();
    auto split = [&](const u64 *src, auto pre) {
        if constexpr (C == 0) load_fwd_split<LOGN>(x, src, c, r, tid, tw, m, pre, sm);   // (limbs of four chunks: clusters of two, PAIRS)
        else load_fwd_split<LOGN, false, true>(x, src, c, r, tid, tw, m, pre, sm, load_t0);
    };
    const u64 *bp = A.base + (size_t)b * A.base_ct_stride + (size_t)p * A.base_poly_stride + (size_t)j * N + off;
    const u64 *ap = A.addend[p] ? A.addend[p] + (size_t)b * A.add_ct_stride + (size_t)j * N + off : nullptr;
    u64 *op = A.out + (size_t)b * A.out_ct_stride + (size_t)p * A.out_poly_stride + (size_t)j * N + off;
    // epilogue operands: bulk copies (TMA) into shared memory, in flight during the transform.  out may alias the
    // addend element for element (b200he_apply_galois): this CTA is the only one that touches its tile, and it has
    // read the whole tile before it writes.
    // (thread 0 initialises, arms and uses the barrier; everyone else first touches it after the CTA-wide barriers of
    // the transform, which order the initialisation before their wait)
    // Galois-fused addend (polynomial 0 only): the chunk of the input limb that this output chunk gathers from lands in
    // stage_a, this chunk of the table behind the barrier word.  A contiguous addend of the same launch (rotate-and-add)
    // shares stage_a when it is the same tile (unsplit limbs) and is read straight from global memory otherwise.
    const bool gal = A.gal && p == 0;
    const u32 *tab_sm = reinterpret_cast<const u32 *>(bar + 2);
    const u64 *gp = gal ? A.gal_src + (size_t)b * A.gal_ct_stride + (size_t)j * N + (size_t)A.gal_chunk[r] * NL : nullptr;
    const bool ap_staged = ap && (!gal || gp == ap), ap_direct = ap && !ap_staged;
    if (tid == 0) {
        tma_bar_init(bar);
        tma_bar_expect(bar, ((ap_staged || gal) ? 2u : 1u) * NL * 8 + (gal ? NL * 4u : 0u));
        tma_load_1d(stage_b, bp, NL * 8, bar);
        if (ap_staged) tma_load_1d(stage_a, ap, NL * 8, bar);
        else if (gal) tma_load_1d(stage_a, gp, NL * 8, bar);
        if (gal) tma_load_1d(const_cast<u32 *>(tab_sm), A.gal + off, NL * 4, bar);
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
                split(rp, PreTwoDp<true, false>{ wq30, nq, dp_from(fix), fix2, sd, sq, rd, rq, rp2 });
            else
                split(rp, PreTwoDp<true, true>{ wq30, nq, dp_from(fix), fix2, sd, sq, rd, rq, rp2 });
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
        if (w1) split(rp, PreReduceFixDp<true>{ wq30, nq, dp_from(fix) });
        else split(rp, PreReduceFixDp<false>{ wq30, nq, dp_from(fix) });
        ntt_fwd_regs_split<LOGN>(x, sm, tw, m, tid, c, r, t0);
        contig_to_co(x, sm, tid);
        tma_bar_wait(bar, 0);
        for_pairs_co(tid, [&](int reg, int e) {
            const ulonglong2 bv = ld2(stage_b + e);
            double v0 = dp_mul(__dadd_rn(dp_from(bv.x), -as_d(x[reg])), sd, sq, nq), v1 = dp_mul(__dadd_rn(dp_from(bv.y), -as_d(x[reg + 1])), sd, sq, nq);
            if (ap) {
                const ulonglong2 av = ap_direct ? ldg2(ap + e) : ld2(stage_a + e);
                v0 = __dadd_rn(v0, dp_from(av.x));
                v1 = __dadd_rn(v1, dp_from(av.y));
            }
            if (gal) {   // + g(c0): gathered from the staged source chunk
                const u32 t0 = tab_sm[e] & (NL - 1), t1 = tab_sm[e + 1] & (NL - 1);
                v0 = __dadd_rn(v0, dp_from(stage_a[t0]));
                v1 = __dadd_rn(v1, dp_from(stage_a[t1]));
            }
            st2(op + e, dp_canon(v0, m), dp_canon(v1, m));
        });
        return;
    }
    if (A.rp2) {
        const ulonglong2 ri = T.qinv[(size_t)A.x2 * T.M + j];
        const u64 fix2 = m.q - T.halfmod[(size_t)A.x2 * T.M + j];
        if (lift_wide(T.mods[A.x].q, m.q) || lift_wide(T.mods[A.x2].q, m.q))
            split(A.rp + ((size_t)b * A.P + p) * N, PreTwo<true>{ m.q, m.r64, fix, fix2, qi, ri, A.rp2 + ((size_t)b * A.P + p) * N });
        else
            split(A.rp + ((size_t)b * A.P + p) * N, PreTwo<false>{ m.q, m.r64, fix, fix2, qi, ri, A.rp2 + ((size_t)b * A.P + p) * N });
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
        split(A.rp + ((size_t)b * A.P + p) * N, PreReduceFix<true>{ m.q, m.r64, fix });
    else
        split(A.rp + ((size_t)b * A.P + p) * N, PreReduceFix<false>{ m.q, m.r64, fix });
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
            const ulonglong2 av = ap_direct ? ldg2(ap + e) : ld2(stage_a + e);
            v0 = add_mod(v0, av.x, m.q);
            v1 = add_mod(v1, av.y, m.q);
        }
        if (gal) {   // + g(c0): gathered from the staged source chunk
            const u32 t0 = tab_sm[e] & (NL - 1), t1 = tab_sm[e + 1] & (NL - 1);
            v0 = add_mod(v0, stage_a[t0], m.q);
            v1 = add_mod(v1, stage_a[t1], m.q);
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

}   // namespace b200he
