// behz.cuh -- BFV ciphertext multiplication in RNS (BEHZ), SEAL Evaluator::bfv_multiply + util/rns.cpp
// RNSTool::{fastbconv_m_tilde, sm_mrq, fast_floor, fastbconv_sk}.
// Reached from R/src/benchmarks/bfv/seal_bfv_element_wise_benchmark.cpp:326 and the BFV dot-product /
// matmul workloads.
//
// The base conversions are per-coefficient small matrix-vector products over the RNS limbs, so they
// are elementwise kernels (one thread per coefficient pair, 128-bit accesses); the transforms between
// them reuse k_ntt_fwd / k_ntt_inv on the extended base q u Bsk.
//   k_behz_extend   x (base q, coeff form) -> [x | SmMRq(FastBConv_{q -> Bsk u m~}(m~ x))]   (steps 1-2)
//   k_behz_tensor   (a0,a1) x (b0,b1) -> (d0,d1,d2) in q u Bsk, NTT domain                     (step 3)
//   k_behz_floor_sk t*d -> fast floor in Bsk -> Shenoy-Kumaresan back to q                      (steps 4-6)
#pragma once
#include "kernels.cuh"

namespace b200he {

constexpr int BEHZ_MAXL = 8;    // data primes
constexpr int BEHZ_MAXB = 10;   // |Bsk| = |B| + 1

struct BehzConst {
    int L, nB, nbsk, K;
    u64 mt_invp_q[BEHZ_MAXL];            // m~ * (q/q_l)^{-1} mod q_l
    u64 t_invp_q[BEHZ_MAXL];             // t  * (q/q_l)^{-1} mod q_l
    u64 q2bsk[BEHZ_MAXB][BEHZ_MAXL];     // (q/q_l) mod p_i
    u32 q2mt[BEHZ_MAXL];                 // (q/q_l) mod 2^32
    u32 neg_inv_q_mt;                    // -q^{-1} mod 2^32
    u64 prod_q_bsk[BEHZ_MAXB];           // q mod p_i
    u64 inv_prod_q_bsk[BEHZ_MAXB];       // q^{-1} mod p_i
    u64 inv_mt_bsk[BEHZ_MAXB];           // m~^{-1} mod p_i
    u64 t_bsk[BEHZ_MAXB];                // t mod p_i
    u64 invp_B[BEHZ_MAXB];               // (B/B_j)^{-1} mod B_j
    u64 B2q[BEHZ_MAXL][BEHZ_MAXB];       // (B/B_j) mod q_l
    u64 B2msk[BEHZ_MAXB];                // (B/B_j) mod m_sk
    u64 inv_prod_B_msk;                  // B^{-1} mod m_sk
    u64 prod_B_q[BEHZ_MAXL];             // B mod q_l
};

struct Behz {
    BehzConst h;              // host copy
    void *d_blob = nullptr;   // device copy
    const BehzConst *d() const { return (const BehzConst *)d_blob; }
};

struct BehzExtArgs {
    const u64 *a, *b;          // operand batches [.][2][L][N], coefficient form
    const u32 *ai, *bi;        // nullable index maps
    size_t a_stride, b_stride;
    u64 *ext;                  // [n][4][L+nbsk][N]: polys a0,a1,b0,b1; limbs q then Bsk
    size_t n;
};

// steps 1-2 for all four input polynomials of every product
__global__ void __launch_bounds__(256) k_behz_extend(Tables T, const BehzConst *__restrict__ Cp, BehzExtArgs A)
{
    const BehzConst &C = *Cp;
    const size_t N = T.N, per = 4 * N / 2;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= A.n * per) return;
    const size_t i = gid / per, rem = gid % per;
    const int poly = (int)(rem / (N / 2));
    const size_t e = (rem % (N / 2)) * 2;
    const int L = C.L, nb = C.nbsk, W = L + nb;
    const u64 *src = (poly < 2) ? A.a + (A.ai ? A.ai[i] : i) * A.a_stride + (size_t)poly * L * N
                                : A.b + (A.bi ? A.bi[i] : i) * A.b_stride + (size_t)(poly - 2) * L * N;
    u64 *dst = A.ext + (i * 4 + poly) * (size_t)W * N;
    u64 tmp[BEHZ_MAXL][2];
    u32 r32[2] = { 0, 0 };
    for (int l = 0; l < L; l++) {
        const ulonglong2 v = ld2(src + (size_t)l * N + e);
        st2(dst + (size_t)l * N + e, v.x, v.y);
        const Mod m = T.mods[l];
        tmp[l][0] = mul_mod(v.x, C.mt_invp_q[l], m);
        tmp[l][1] = mul_mod(v.y, C.mt_invp_q[l], m);
        r32[0] += (u32)tmp[l][0] * C.q2mt[l];
        r32[1] += (u32)tmp[l][1] * C.q2mt[l];
    }
    r32[0] *= C.neg_inv_q_mt;
    r32[1] *= C.neg_inv_q_mt;
    for (int k = 0; k < nb; k++) {
        const Mod m = T.mods[C.K + k];
        u64 o[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            u64 acc = 0;
            for (int l = 0; l < L; l++) acc = mad_mod(tmp[l][h], C.q2bsk[k][l], acc, m);
            u64 r = r32[h];
            if (r >= (u64(1) << 31)) r += m.q - (u64(1) << 32);
            o[h] = mul_mod(mad_mod(r, C.prod_q_bsk[k], acc, m), C.inv_mt_bsk[k], m);
        }
        st2(dst + (size_t)(L + k) * N + e, o[0], o[1]);
    }
}

// step 3: dyadic tensor product of the transformed operands in every limb of q u Bsk
__global__ void __launch_bounds__(256) k_behz_tensor(Tables T, const BehzConst *__restrict__ Cp, const u64 *__restrict__ ext, u64 *__restrict__ prod, size_t n)
{
    const BehzConst &C = *Cp;
    const size_t N = T.N;
    const int W = C.L + C.nbsk;
    const size_t WN = (size_t)W * N, per = WN / 2;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * per) return;
    const size_t i = gid / per, rem = (gid % per) * 2;
    const int l = (int)(rem / N);
    const Mod m = T.mods[l < C.L ? l : C.K + (l - C.L)];
    const u64 *p = ext + i * 4 * WN + rem;
    const ulonglong2 a0 = ld2(p), a1 = ld2(p + WN), b0 = ld2(p + 2 * WN), b1 = ld2(p + 3 * WN);
    u64 *o = prod + i * 3 * WN + rem;
    st2(o, mul_mod(a0.x, b0.x, m), mul_mod(a0.y, b0.y, m));
    st2(o + WN, mad_mod(a0.x, b1.x, mul_mod(a1.x, b0.x, m), m), mad_mod(a0.y, b1.y, mul_mod(a1.y, b0.y, m), m));
    st2(o + 2 * WN, mul_mod(a1.x, b1.x, m), mul_mod(a1.y, b1.y, m));
}

// steps 4-6 on the coefficient-form products: out[i][p][l] for p < 3, l < L
__global__ void __launch_bounds__(256) k_behz_floor_sk(Tables T, const BehzConst *__restrict__ Cp, const u64 *__restrict__ prod, u64 *__restrict__ out, size_t n)
{
    const BehzConst &C = *Cp;
    const size_t N = T.N, per = 3 * N / 2;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * per) return;
    const size_t i = gid / per, rem = gid % per;
    const int poly = (int)(rem / (N / 2));
    const size_t e = (rem % (N / 2)) * 2;
    const int L = C.L, nb = C.nbsk, nB = C.nB, W = L + nb;
    const u64 *src = prod + (i * 3 + poly) * (size_t)W * N;
    u64 *dst = out + (i * 3 + poly) * (size_t)L * N;
    u64 tq[BEHZ_MAXL][2], fl[BEHZ_MAXB][2];
    for (int l = 0; l < L; l++) {
        const ulonglong2 v = ld2(src + (size_t)l * N + e);
        const Mod m = T.mods[l];
        tq[l][0] = mul_mod(v.x, C.t_invp_q[l], m);
        tq[l][1] = mul_mod(v.y, C.t_invp_q[l], m);
    }
    // fast floor: (t*y_p - FastBConv_{q->p}(t*y_q)) * q^{-1} mod p
    for (int k = 0; k < nb; k++) {
        const Mod m = T.mods[C.K + k];
        const ulonglong2 v = ld2(src + (size_t)(L + k) * N + e);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            u64 conv = 0;
            for (int l = 0; l < L; l++) conv = mad_mod(tq[l][h], C.q2bsk[k][l], conv, m);
            const u64 tb = mul_mod(h ? v.y : v.x, C.t_bsk[k], m);
            fl[k][h] = mul_mod(sub_mod(tb, conv, m.q), C.inv_prod_q_bsk[k], m);
        }
    }
    // Shenoy-Kumaresan: alpha from m_sk, then B -> q
    const Mod msk = T.mods[C.K + nB];
    u64 alpha[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        u64 s = 0;
        for (int j = 0; j < nB; j++) {
            const Mod mj = T.mods[C.K + j];
            fl[j][h] = mul_mod(fl[j][h], C.invp_B[j], mj);   // now holds fl_j * (B/B_j)^{-1} mod B_j
            s = mad_mod(fl[j][h], C.B2msk[j], s, msk);
        }
        alpha[h] = mul_mod(sub_mod(s, fl[nB][h], msk.q), C.inv_prod_B_msk, msk);
    }
    for (int l = 0; l < L; l++) {
        const Mod m = T.mods[l];
        u64 o[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            u64 acc = 0;
            for (int j = 0; j < nB; j++) acc = mad_mod(fl[j][h], C.B2q[l][j], acc, m);
            if (alpha[h] > (msk.q >> 1)) o[h] = mad_mod(reduce64(msk.q - alpha[h], m), C.prod_B_q[l], acc, m);
            else o[h] = mad_mod(reduce64(alpha[h], m), m.q - C.prod_B_q[l], acc, m);
        }
        st2(dst + (size_t)l * N + e, o[0], o[1]);
    }
}

}   // namespace b200he
