// kernels_ew.cuh -- the non-template kernels of the hot path: elementwise / tensor / Galois / copy kernels, the key
// upload transform and the coefficient-form mod-down.  Included by exactly ONE translation unit (b200he.cu): plain
// __global__ functions have external linkage.  The NTT-bearing template kernels live in kernels.cuh and are instantiated
// in their own translation units (tu_ntt.cu, tu_ks.cu, tu_moddown.cu) so that the kernel families compile in parallel.
#pragma once
#include "kernels.cuh"

namespace b200he {

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

// Matrix of ciphertexts times matrix of ciphertexts, the inner loop of MatMult CipherBatchAxis
// (R/src/benchmarks/ckks/seal_ckks_matmult_cipherbatchaxis_benchmark.cpp:385-422): for every output cell (i, j)
//     out[i][j] = sum_k a[i][k] (x) b[k][j]          ((x) = the 2 x 2 -> 3 tensor product of k_tensor, size-3 result)
// (b is held by columns: bt[j][k] = b[k][j], so both operands of a cell are contiguous runs of `inner` ciphertexts)
// which the reference evaluates as `inner` multiply() calls and inner - 1 add_inplace() calls per cell.  Modular
// arithmetic is exact, so summing the 128-bit products and reducing once gives the same canonical residues.
//
// One thread owns ONE coefficient of a TI x TJ tile of cells: the three accumulators of every cell are 128-bit integers
// in registers (a term adds less than 2 q^2, MacArgs::reduce_every bounds the run), a step of the k loop loads
// 2 TI + 2 TJ words and feeds 4 TI TJ wide multiply-accumulates, and nothing but the final residues is written: 
// (2 TI + 2 TJ) / (TI TJ) loaded words per product against 7 loaded + 9 stored for the multiply()/add_inplace() pair.
// Blocks are ordered tile-fastest: the CTAs resident at any moment work on the same coefficient range of many tiles, in
// step through k, so the slices of a[i][k] and b[k][j] they share are served from L2 and cross HBM about once.
struct MacArgs {
    const u64 *a, *b;      // a: ciphertexts [rows][inner] (index i * inner + k), b: by columns, [cols][inner]; size 2, NTT form
    u64 *out;              // [rows * cols] size-3 ciphertexts
    size_t a_stride, b_stride, out_stride;   // words per ciphertext
    int L;
    u32 rows, inner, cols, tiles_j, ntiles;
    u32 reduce_every;      // fold the accumulators before a run passes this many terms (host: floor(2^127 / max q^2), at least 1)
    u32 flush_every;       // integer-pipe limbs: steps of the k loop between flushes of MacAcc's odd part (host, from max q; at least 1)
};
// 128-bit accumulator of 64 x 64-bit products for the integer-pipe limbs of k_tensor_mac, laid out so that a product
// costs exactly four IMAD.WIDE.U32 and no register moves: with x = x1 2^32 + x0,
//   even part (w3..w0, 128 bits)  += x0 y0 + 2^64 x1 y1     two multiply-adds chained by the carry flag
//   odd part  (o1:o0, weight 2^32) += x0 y1 + x1 y0          two multiply-adds, no carry word: the caller adds the odd
//                                                            part into the even one (flush) before it can overflow --
//                                                            every MacArgs::flush_every steps of the k loop
// Every multiply-add reads and writes an aligned register pair.  (The obvious forms -- mad.lo.cc.u64 / madc.hi.u64, or
// one schoolbook chain over four words -- cost ~20 and ~13 instructions per product in SASS: seven multiplies and
// compare-and-select carries, or a register move per multiply because the middle words straddle two pairs.)
struct MacAcc {
#if defined(__CUDA_ARCH__)
    u64 elo, ehi, odd;   // 64-bit values: the register allocator keeps each in an aligned pair, which is what IMAD.WIDE wants
    __device__ __forceinline__ void zero() { elo = ehi = odd = 0; }
    __device__ __forceinline__ void mac(u64 x, u64 y)
    {
        const u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
        asm("{\n\t"
            ".reg .u64 p, q;\n\t"
            "mul.wide.u32 p, %3, %5;\n\t"
            "mul.wide.u32 q, %4, %6;\n\t"
            "add.cc.u64 %0, %0, p;\n\t"
            "addc.u64 %1, %1, q;\n\t"
            "mad.wide.u32 %2, %3, %6, %2;\n\t"
            "mad.wide.u32 %2, %4, %5, %2;\n\t"
            "}"
            : "+l"(elo), "+l"(ehi), "+l"(odd)
            : "r"(x0), "r"(x1), "r"(y0), "r"(y1));
    }
    __device__ __forceinline__ void flush()
    {
        // (ehi:elo) += odd * 2^32
        asm("{\n\t"
            ".reg .u64 a, b;\n\t"
            "shl.b64 a, %2, 32;\n\t"
            "shr.u64 b, %2, 32;\n\t"
            "add.cc.u64 %0, %0, a;\n\t"
            "addc.u64 %1, %1, b;\n\t"
            "}"
            : "+l"(elo), "+l"(ehi)
            : "l"(odd));
        odd = 0;
    }
    __device__ __forceinline__ u64 lo() const { return elo; }
    __device__ __forceinline__ u64 hi() const { return ehi; }
    __device__ __forceinline__ void set(u64 l) { elo = l; ehi = 0; }
#else   // the emulation build keeps the same two parts (and the same overflow behaviour of the odd one)
    unsigned __int128 even;
    u64 odd;
    void zero() { even = 0; odd = 0; }
    void mac(u64 x, u64 y)
    {
        const u64 x0 = (u32)x, x1 = x >> 32, y0 = (u32)y, y1 = y >> 32;
        even += (unsigned __int128)(x0 * y0) + ((unsigned __int128)(x1 * y1) << 64);
        odd += x0 * y1 + x1 * y0;
    }
    void flush() { even += (unsigned __int128)odd << 32; odd = 0; }
    u64 lo() const { return (u64)even; }
    u64 hi() const { return (u64)(even >> 64); }
    void set(u64 l) { even = l; }
#endif
};
// any 128-bit value -> canonical residue: (hi mod q) * (2^64 mod q) + (lo mod q)
__device__ __forceinline__ u64 fold128(u64 lo, u64 hi, const Mod &m, u64 r64q) { return mad_mod(reduce64(hi, m), r64q, reduce64(lo, m), m); }
// FP64-domain instance (moduli below 2^46, modarith.cuh): a residue splits exactly into two halves below 2^23 held in
// doubles, a 46-bit product becomes four partial products below 2^46, and every accumulator is three doubles (weights
// 1, 2^23, 2^46) fed by DFMA -- exact while the sums stay below 2^53, i.e. for runs of MAC_DP_RUN terms, after which each
// double is reduced modulo q in place (x - rint(x/q) q, exact).  4 DFMA per product against ~5 IMAD.WIDE + carries.
#define B200HE_MAC_DP_RUN 16   /* the middle accumulator of c1 gains 4 partial products < 2^46 per term: 16 * 2^48 + q < 2^53 */
struct DpHalves { double lo, hi; };
__device__ __forceinline__ DpHalves dp_split23(u64 x) { return DpHalves{ dp_from(x & 0x7fffffull), dp_from(x >> 23) }; }
struct DpAcc {
    double ll, mid, hh;
    __device__ __forceinline__ void mac(const DpHalves &a, const DpHalves &b)
    {
        ll = __fma_rn(a.lo, b.lo, ll);
        mid = __fma_rn(a.lo, b.hi, mid);
        mid = __fma_rn(a.hi, b.lo, mid);
        hh = __fma_rn(a.hi, b.hi, hh);
    }
    __device__ __forceinline__ void reduce(const Mod &m)
    {
        ll = dp_reduce(ll, m.dqinv, m.dnq);
        mid = dp_reduce(mid, m.dqinv, m.dnq);
        hh = dp_reduce(hh, m.dqinv, m.dnq);
    }
    // ll + 2^23 mid + 2^46 hh mod q, canonical
    __device__ __forceinline__ u64 finish(const Mod &m)
    {
        reduce(m);
        const double w = 8388608.0, wq = 8388608.0 * m.dqinv;   // 2^23 RN(1/q) = RN(2^23 / q) exactly
        const double t = __dadd_rn(dp_mul(hh, w, wq, m.dnq), mid);
        return dp_canon(__dadd_rn(dp_mul(t, w, wq, m.dnq), ll), m);
    }
};
template <int TI, int TJ> __device__ __forceinline__ void tensor_mac_dp(const Tables &T, const MacArgs &A, const Mod &m, size_t ce, u32 i0, u32 j0)
{
    const size_t LN = (size_t)A.L * T.N;
    DpAcc acc[TI][TJ][3];
#pragma unroll
    for (int ti = 0; ti < TI; ti++)
#pragma unroll
        for (int tj = 0; tj < TJ; tj++)
#pragma unroll
            for (int c = 0; c < 3; c++) acc[ti][tj][c] = DpAcc{ 0.0, 0.0, 0.0 };
    const u64 *pa[TI], *pb[TJ];   // walk the inner dimension: one stride per step
#pragma unroll
    for (int ti = 0; ti < TI; ti++) pa[ti] = A.a + (size_t)(i0 + ti < A.rows ? i0 + ti : A.rows - 1) * A.inner * A.a_stride + ce;
#pragma unroll
    for (int tj = 0; tj < TJ; tj++) pb[tj] = A.b + (size_t)(j0 + tj < A.cols ? j0 + tj : A.cols - 1) * A.inner * A.b_stride + ce;
    // (requesting the operands of step k + 1 before the multiply-accumulates of step k was measured: 28.0 -> 30.9 ms on the
    // 48 x 48 x 48 probe -- at 128 registers the second operand set costs more in moves and spills than the latency it hides)
    u32 run = 0;
#pragma unroll 1
    for (u32 k = 0; k < A.inner; k++) {
        DpHalves a0[TI], a1[TI];
#pragma unroll
        for (int ti = 0; ti < TI; ti++) {
            a0[ti] = dp_split23(ldg1(pa[ti]));
            a1[ti] = dp_split23(ldg1(pa[ti] + LN));
            pa[ti] += A.a_stride;
        }
#pragma unroll
        for (int tj = 0; tj < TJ; tj++) {
            const DpHalves b0 = dp_split23(ldg1(pb[tj])), b1 = dp_split23(ldg1(pb[tj] + LN));
            pb[tj] += A.b_stride;
#pragma unroll
            for (int ti = 0; ti < TI; ti++) {
                acc[ti][tj][0].mac(a0[ti], b0);
                acc[ti][tj][1].mac(a0[ti], b1);
                acc[ti][tj][1].mac(a1[ti], b0);
                acc[ti][tj][2].mac(a1[ti], b1);
            }
        }
        if (++run == B200HE_MAC_DP_RUN && k + 1 < A.inner) {
            run = 0;
#pragma unroll
            for (int ti = 0; ti < TI; ti++)
#pragma unroll
                for (int tj = 0; tj < TJ; tj++)
#pragma unroll
                    for (int c = 0; c < 3; c++) acc[ti][tj][c].reduce(m);
        }
    }
#pragma unroll
    for (int ti = 0; ti < TI; ti++)
#pragma unroll
        for (int tj = 0; tj < TJ; tj++) {
            if (i0 + ti >= A.rows || j0 + tj >= A.cols) continue;
            u64 *o = A.out + ((size_t)(i0 + ti) * A.cols + (j0 + tj)) * A.out_stride + ce;
#pragma unroll
            for (int c = 0; c < 3; c++) o[(size_t)c * LN] = acc[ti][tj][c].finish(m);
        }
}
template <int TI, int TJ> __device__ __forceinline__ void tensor_mac_int(const Tables &T, const MacArgs &A, const Mod &m, size_t ce, u32 i0, u32 j0)
{
    const size_t LN = (size_t)A.L * T.N;
    const u64 r64q = reduce64(m.nq, m);   // 2^64 mod q
    MacAcc acc[TI][TJ][3];
#pragma unroll
    for (int ti = 0; ti < TI; ti++)
#pragma unroll
        for (int tj = 0; tj < TJ; tj++)
#pragma unroll
            for (int c = 0; c < 3; c++) acc[ti][tj][c].zero();
    // rows / columns beyond the matrix (edge tiles) repeat the last valid one; their results are not stored
    const u64 *pa[TI], *pb[TJ];   // walk the inner dimension: one stride per step
#pragma unroll
    for (int ti = 0; ti < TI; ti++) pa[ti] = A.a + (size_t)(i0 + ti < A.rows ? i0 + ti : A.rows - 1) * A.inner * A.a_stride + ce;
#pragma unroll
    for (int tj = 0; tj < TJ; tj++) pb[tj] = A.b + (size_t)(j0 + tj < A.cols ? j0 + tj : A.cols - 1) * A.inner * A.b_stride + ce;
    u32 run = 0;
    for (u32 k = 0; k < A.inner;) {
        const u32 kend = k + A.flush_every < A.inner ? k + A.flush_every : A.inner;
#pragma unroll 1
        for (; k < kend; k++) {
            u64 a0[TI], a1[TI], b0[TJ], b1[TJ];
#pragma unroll
            for (int ti = 0; ti < TI; ti++) {
                a0[ti] = ldg1(pa[ti]);
                a1[ti] = ldg1(pa[ti] + LN);
                pa[ti] += A.a_stride;
            }
#pragma unroll
            for (int tj = 0; tj < TJ; tj++) {
                b0[tj] = ldg1(pb[tj]);
                b1[tj] = ldg1(pb[tj] + LN);
                pb[tj] += A.b_stride;
            }
#pragma unroll
            for (int ti = 0; ti < TI; ti++)
#pragma unroll
                for (int tj = 0; tj < TJ; tj++) {
                    acc[ti][tj][0].mac(a0[ti], b0[tj]);
                    acc[ti][tj][1].mac(a0[ti], b1[tj]);
                    acc[ti][tj][1].mac(a1[ti], b0[tj]);
                    acc[ti][tj][2].mac(a1[ti], b1[tj]);
                }
        }
        run += A.flush_every;
        const bool fold = run + A.flush_every > A.reduce_every && k < A.inner;   // the next stretch would pass the 128-bit bound
#pragma unroll
        for (int ti = 0; ti < TI; ti++)
#pragma unroll
            for (int tj = 0; tj < TJ; tj++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    acc[ti][tj][c].flush();
                    if (fold) acc[ti][tj][c].set(fold128(acc[ti][tj][c].lo(), acc[ti][tj][c].hi(), m, r64q));
                }
        if (fold) run = 0;
    }
#pragma unroll
    for (int ti = 0; ti < TI; ti++)
#pragma unroll
        for (int tj = 0; tj < TJ; tj++) {
            if (i0 + ti >= A.rows || j0 + tj >= A.cols) continue;
            u64 *o = A.out + ((size_t)(i0 + ti) * A.cols + (j0 + tj)) * A.out_stride + ce;
#pragma unroll
            for (int c = 0; c < 3; c++) o[(size_t)c * LN] = fold128(acc[ti][tj][c].lo(), acc[ti][tj][c].hi(), m, r64q);
        }
}
// a block covers MAC_THREADS consecutive coefficients of one limb ([L][N], N a multiple of the block size): the modulus,
// and with it the arithmetic domain, is uniform over the block
#define B200HE_MAC_THREADS 128
template <int TI, int TJ> __global__ void __launch_bounds__(B200HE_MAC_THREADS, 4) k_tensor_mac(Tables T, MacArgs A)
{
    const size_t N = T.N, LN = (size_t)A.L * N;
    const u32 tile = blockIdx.x % A.ntiles, cb = blockIdx.x / A.ntiles;
    const size_t ce = (size_t)cb * blockDim.x + threadIdx.x;   // coefficient index within [L][N]
    if (ce >= LN) return;
    const Mod m = T.mods[ce / N];
    const u32 i0 = (tile / A.tiles_j) * TI, j0 = (tile % A.tiles_j) * TJ;
    if (m.dp) tensor_mac_dp<TI, TJ>(T, A, m, ce, i0, j0);
    else {
        // a 2 x 2 tile of MacAcc (12 x 6 words, next to 8 operands and 4 pointers) does not fit the 128 registers that
        // four resident blocks allow -- the allocator answered with ~70 register moves per step -- so the rows of the
        // tile are done one after the other; the second pass finds the b operands in L2
#pragma unroll 1
        for (int ti = 0; ti < TI; ti++)
            if (i0 + ti < A.rows) tensor_mac_int<1, TJ>(T, A, m, ce, i0 + ti, j0);
    }
}

// ------------------------------------------------------------------------------------ K8
// Coefficient-form (BFV) Galois automorphism of a size-2 ciphertext batch: out0 = g(c0) -> dst ct poly 0,
// g(c1) -> target buffer [B][L][N]; dst poly 1 is produced by the key switch that follows.  (The NTT-form automorphism
// of CKKS has no kernel of its own: it is a gather on load inside the key switch, kernels_ks.cuh / kernels_moddown.cuh.)
struct GaloisArgs {
    const u64 *src;       // [B][2][L][N]
    u64 *dst0;            // g(c0): dst0 + b*dst_stride + l*N
    u64 *dst1;            // g(c1): dst1 + b*L*N + l*N
    size_t src_stride, dst_stride;
    u32 elt;              // Galois element (BFV coefficient form)
    int L, logn;
    size_t B;
    int add_input;        // dst0 = g(c0) + c0 (rotate-and-add of accumulate: the key switch then adds c1 as its second addend)
};
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
