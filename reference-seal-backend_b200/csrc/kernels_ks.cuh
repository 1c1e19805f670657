// kernels_ks.cuh -- K6 step 2: the key-switch inner product k_ks_inner (instantiated in tu_ks.cu).
#pragma once
#include "kernels_ntt.cuh"

namespace b200he {

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
    const u32 *gal;        // Galois permutation of the NTT-form target, applied on load ([N], nullptr: none): see InvFuse::gal
    // Fused relinearize + rescale: the accumulator limb of the LAST data modulus leaves this kernel as
    //   y = acc * s + c        (s = q_sp^{-1} mod q_{L-1}, c = limb L-1 of the input ciphertext's component k, NTT form)
    // -- the value whose inverse transform the rescale rounds -- so that the inverse transform that follows reads one
    // tile instead of two and owns no pre-processing.  fuse_add + b * ct_stride + k * poly_stride + (L-1) N, or nullptr.
    const u64 *fuse_add;
    size_t fuse_ct_stride, fuse_poly_stride;
    int unit0;             // first unit of this launch (limbs of four chunks run as two launches: the special-prime units, then the rest)
};
// lazy accumulator of the inner product (either domain) -> canonical residue
__device__ __forceinline__ u64 acc_finish(u64 a, const Mod &m) { return m.dp ? dp_canon(as_d(a), m) : reduce_full(a, m); }
template <int LOGN> struct KsCfg {
    static constexpr int SMEM_BYTES = 3 * NttCfg<LOGN>::SMEM_BYTES;   // transform buffer + 2 accumulator limbs
};
// DP = Mod::dp of the CTA's modulus as a compile-time constant: the kernel branches once, at the top, into one of two
// complete instances of the body, so the integer and the FP64-domain code never share live ranges (with the branch
// inside the multiply-accumulate loop the register allocator spilled in both).
template <int LOGN, int C, bool DP, bool PAIRS>
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
            if (A.gal) gather_pairs_co(x, tp - off, A.gal + off, tid);
            else
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
            // twiddles of the first local pass: next to the data loads for unsplit limbs; for clusters from inside the cross
            // stages (before their second barrier), so that 28 registers are not live across the radix-2/4 butterflies
            TwRegs<LOGN, 0> t0;
            auto load_t0 = [&]() { load_tw_early<LOGN, 0, false>(t0, tw, tid, (1 << c) + r); };
            if constexpr (C == 0) load_t0();
            auto split = [&](auto pre) {
                if constexpr (C == 0) load_fwd_split<LOGN, true, false>(x, tp, c, r, tid, tw, m, pre, sm);
                else load_fwd_split<LOGN, true, PAIRS>(x, tp, c, r, tid, tw, m, pre, sm, load_t0);
            };
            if constexpr (DP) {   // FP64 domain: digits of moduli up to 48 bits are lazy values as they are (2^48 + 14 q < 2^50)
                if (T.mods[J].bits > 48) split(PreLiftDp{ 1073741824.0 * m.dqinv, m.dnq });
                else split(PreNone());
            } else if (lift_wide(T.mods[J].q, m.q))
                split(PreReduce<true>{ m.q, m.r64 });
            else if (T.mods[J].q > m.q)
                split(PreReduce<false>{ m.q, m.r64 });
            else
                split(PreNone());
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
        if (A.fuse_add && I == L - 1) {   // y = acc * s + c, canonical (CTA-uniform branch)
            const u64 *cp = A.fuse_add + (size_t)b * A.fuse_ct_stride + (size_t)k * A.fuse_poly_stride + (size_t)I * N + off;
            const ulonglong2 s = T.qinv[(size_t)(A.K - 1) * T.M + ki];
            if constexpr (DP) {
                const double sd = dp_from(s.x), sq = __dmul_rn(sd, m.dqinv);
                for_pairs_co(tid, [&](int reg, int e) {
                    const ulonglong2 cv = ldg2(cp + e);
                    st2(o + e, dp_canon(__dadd_rn(dp_mul(dp_from(x[reg]), sd, sq, m.dnq), dp_from(cv.x)), m),
                        dp_canon(__dadd_rn(dp_mul(dp_from(x[reg + 1]), sd, sq, m.dnq), dp_from(cv.y)), m));
                });
            } else {
                for_pairs_co(tid, [&](int reg, int e) {
                    const ulonglong2 cv = ldg2(cp + e);
                    st2(o + e, add_mod(shoup(x[reg], s.x, s.y, m.q), cv.x, m.q), add_mod(shoup(x[reg + 1], s.x, s.y, m.q), cv.y, m.q));
                });
            }
        } else
            for_pairs_co(tid, [&](int reg, int e) { st2(o + e, x[reg], x[reg + 1]); });
        warp_sync();   // the slice is rewritten by the next component
    }
}
// PAIRS (limbs of four chunks only): the launch holds no special-prime unit, hence no inverse transform, and runs in clusters
// of TWO -- the first cross stage from global memory, the second exchanged inside the pair (load_fwd_split) -- which fill all
// 148 SMs where clusters of four fill 132.  The special-prime units keep their clusters of four in a launch of their own.
template <int LOGN, int C, bool PAIRS = false>
__global__ void __launch_bounds__(NttCfg<LOGN>::THREADS, 1) k_ks_inner(Tables T, KsInnerArgs A)
{
    const int r = blockIdx.x & ((1 << C) - 1);
    const int unit = (blockIdx.x >> C) + A.unit0;
    // the special-prime units (longest: they also run the fused inverse transforms) are scheduled first
    const int L = A.L;
    const int b = unit < A.B ? unit : (unit - A.B) / L, I = unit < A.B ? L : (unit - A.B) % L;
    const int ki = (I == L) ? A.K - 1 : I;
    const Mod m = T.mods[ki];
    if (m.dp) ks_inner_body<LOGN, C, true, PAIRS>(T, A, m, b, I, ki, r);
    else ks_inner_body<LOGN, C, false, PAIRS>(T, A, m, b, I, ki, r);
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


}   // namespace b200he
