/*
 * he_oracle_workloads.cpp -- CPU restatement of the reference's workload bodies (operate() of
 * R/src/benchmarks/{ckks,bfv}/ sources and the composite helpers of R/src/engine/seal_context.cpp:255-458)
 * as compositions of the evaluator primitives of he_oracle.cpp.
 *
 * TEST INFRASTRUCTURE ONLY (see he_oracle.h).  These functions are the checker for the ciphertexts the
 * plugin's store() returns: the tests feed them the SAME input encryptions the plugin loaded and the SAME
 * fresh encryptions the reference draws inside its hot path (encrypt_zero at seal_context.cpp:360, encrypt
 * at :440 -- randomised in SEAL, injected here), and compare bit for bit.
 *
 * Every function follows the reference's op ORDER (where a rescale sits, which operand is switched down by
 * matchLevel, NAF order of rotations); modular additions commute, so the order of the mutex-guarded
 * add_inplace reductions (seal_context.cpp:397-400, ...matmult_row...:514-515) is free.
 *
 * Only the public API of he_oracle.h is used.  Scales are host metadata (double) and are not modelled.
 */
#include "he_oracle.h"

#include <cstring>
#include <vector>

typedef uint64_t u64;

namespace {
inline bool is_ckks(const orc_ctx *c) { return orc_ctx_scheme(c) == ORC_SCHEME_CKKS; }

// Evaluator::multiply for either scheme (size 2 x size 2 -> size 3)
void multiply(const orc_ctx *c, size_t L, const u64 *a, const u64 *b, u64 *out3)
{
    if (is_ckks(c)) orc_ckks_multiply(c, L, a, b, out3);
    else orc_bfv_multiply(c, a, b, out3);   // BFV multiply is defined at the top data level (L == K - 1)
}
// mod_switch_to_inplace on a CKKS ciphertext / plaintext: keep the first L_out limbs of each polynomial
// (R/src/engine/seal_context.cpp:260,262,388,451)
void drop_to(size_t N, size_t size, size_t L_in, size_t L_out, const u64 *in, u64 *out)
{
    for (size_t p = 0; p < size; p++) memmove(out + p * L_out * N, in + p * L_in * N, L_out * N * sizeof(u64));
}
}   // namespace

// ---------------------------------------------------------------------------------------------------------
// MatMultValBenchmark::doMatMultVal -- R/src/benchmarks/ckks/seal_ckks_matmultval_benchmark.cpp:235-270,
// R/src/benchmarks/bfv/seal_bfv_matmultval_benchmark.cpp:235-270.
//   out[i][j] = accumulate( [rescale]( relinearize( M0[i] * M1T[j] ) ), cols_M0 )      (rescale: CKKS only, :255)
// m0: [r0][2][L][N], m1t: [c1][2][L][N] (M1 transposed at encode, :213-226), out: [r0*c1][2][Lout][N],
// Lout = L - 1 (CKKS) / L (BFV).
extern "C" int orc_matmul_val(const orc_ctx *c, size_t L, size_t r0, size_t c0, size_t c1, const uint64_t *m0, const uint64_t *m1t,
                              const uint64_t *relin, const uint32_t *elts, const uint64_t *const *gal, size_t ngal, uint64_t *out, int threads)
{
    const size_t N = orc_ctx_N(c), ct = 2 * L * N;
    const bool ckks = is_ckks(c);
    const size_t Lout = ckks ? L - 1 : L;
    if (ckks && L < 2) return -3;
    if (threads <= 0) threads = orc_max_threads();
    int err = 0;
#pragma omp parallel for collapse(2) num_threads(threads)
    for (long i = 0; i < (long)r0; i++)
        for (long j = 0; j < (long)c1; j++) {
            std::vector<u64> t3(3 * L * N), t2(2 * L * N);
            multiply(c, L, m0 + i * ct, m1t + j * ct, t3.data());
            orc_relinearize(c, L, t3.data(), relin, t2.data());
            u64 *dst = out + (size_t)(i * (long)c1 + j) * 2 * Lout * N;
            if (ckks) orc_rescale(c, L, 2, t2.data(), dst);
            else memcpy(dst, t2.data(), ct * sizeof(u64));
            int rc = orc_accumulate(c, Lout, dst, c0, elts, gal, ngal);
            if (rc) err = rc;
        }
    return err;
}

// ---------------------------------------------------------------------------------------------------------
// MatMultRowLatencyBenchmark::matmultrow -- R/src/benchmarks/ckks/seal_ckks_matmult_row_benchmark.cpp:472-523,
// R/src/benchmarks/bfv/seal_bfv_matmult_row_benchmark.cpp:486-539.
//   base = relinearize(A[i] * B);  result[i] = base + sum_{j=1}^{dim2-1} rotate(base, j * spacers)
// (rotate_vector / rotate_rows: arbitrary steps, SEAL's NAF fallback).  A: [nA][2][L][N], B: [2][L][N].
extern "C" int orc_matmul_row(const orc_ctx *c, size_t L, size_t nA, size_t dim2, int spacers, const uint64_t *A, const uint64_t *B,
                              const uint64_t *relin, const uint32_t *elts, const uint64_t *const *gal, size_t ngal, uint64_t *out, int threads)
{
    const size_t N = orc_ctx_N(c), ct = 2 * L * N;
    if (threads <= 0) threads = orc_max_threads();
    std::vector<u64> base(nA * ct);
    for (size_t i = 0; i < nA; i++) {
        std::vector<u64> t3(3 * L * N);
        multiply(c, L, A + i * ct, B, t3.data());
        orc_relinearize(c, L, t3.data(), relin, base.data() + i * ct);
    }
    // the (row, j) rotations are independent; their sum per row is order-free
    const long units = (long)nA * (long)(dim2 > 0 ? dim2 - 1 : 0);
    std::vector<u64> rot((size_t)units * ct);
    int err = 0;
#pragma omp parallel for num_threads(threads)
    for (long u = 0; u < units; u++) {
        const size_t i = (size_t)u / (dim2 - 1), j = 1 + (size_t)u % (dim2 - 1);
        u64 *r = rot.data() + (size_t)u * ct;
        memcpy(r, base.data() + i * ct, ct * sizeof(u64));
        int rc = orc_rotate(c, L, r, (int)j * spacers, elts, gal, ngal);
        if (rc) err = rc;
    }
    if (err) return err;
    for (size_t i = 0; i < nA; i++) {
        u64 *dst = out + i * ct;
        memcpy(dst, base.data() + i * ct, ct * sizeof(u64));
        for (size_t j = 1; j < dim2; j++) orc_add(c, L, 2, dst, rot.data() + (i * (dim2 - 1) + (j - 1)) * ct, dst);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// MatMultCipherBatchAxisBenchmark::operate -- R/src/benchmarks/ckks/seal_ckks_matmult_cipherbatchaxis_benchmark.cpp:385-441:
//   out[i][j] = sum_k m0[i][k] * m1[k][j] kept as size-3 ciphertexts, then ONE relinearize + rescale per cell;
// BFV (R/src/benchmarks/bfv/seal_bfv_matmult_cipherbatchaxis_benchmark.cpp:395-408): every product is relinearized
// before it is added, no rescale.  m0: [r0*c0][2][L][N] row-major, m1: [c0*c1][2][L][N] row-major,
// out: [r0*c1][2][Lout][N] row-major.
extern "C" int orc_matmul_cba(const orc_ctx *c, size_t L, size_t r0, size_t c0, size_t c1, const uint64_t *m0, const uint64_t *m1,
                              const uint64_t *relin, uint64_t *out, int threads)
{
    const size_t N = orc_ctx_N(c), ct = 2 * L * N;
    const bool ckks = is_ckks(c);
    const size_t Lout = ckks ? L - 1 : L;
    if (ckks && L < 2) return -3;
    if (threads <= 0) threads = orc_max_threads();
#pragma omp parallel for collapse(2) num_threads(threads)
    for (long i = 0; i < (long)r0; i++)
        for (long j = 0; j < (long)c1; j++) {
            u64 *dst = out + (size_t)(i * (long)c1 + j) * 2 * Lout * N;
            if (ckks) {
                std::vector<u64> acc(3 * L * N), t3(3 * L * N), t2(2 * L * N);
                for (size_t k = 0; k < c0; k++) {
                    multiply(c, L, m0 + ((size_t)i * c0 + k) * ct, m1 + (k * c1 + (size_t)j) * ct, k == 0 ? acc.data() : t3.data());
                    if (k) orc_add(c, L, 3, acc.data(), t3.data(), acc.data());
                }
                orc_relinearize(c, L, acc.data(), relin, t2.data());
                orc_rescale(c, L, 2, t2.data(), dst);
            } else {
                std::vector<u64> t3(3 * L * N), t2(2 * L * N);
                for (size_t k = 0; k < c0; k++) {
                    multiply(c, L, m0 + ((size_t)i * c0 + k) * ct, m1 + (k * c1 + (size_t)j) * ct, t3.data());
                    orc_relinearize(c, L, t3.data(), relin, k == 0 ? dst : t2.data());
                    if (k) orc_add(c, L, 2, dst, t2.data(), dst);
                }
            }
        }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// SEALContextWrapper::collapseCKKS -- R/src/engine/seal_context.cpp:349-415.
//   retval = Enc(0)                                                           (:360, injected: zero_ct, top level)
//   per i:  tmp = i > 0 ? rotate_vector(ct_i, -i) : ct_i                      (:377-380; i = first_index + local index)
//           plain = encode(e_i, scale); mod_switch_to(plain, tmp.parms_id)    (:382-388; injected: masks, at level L)
//           multiply_plain; relinearize (size 2: no-op); rescale_to_next      (:389-391)
//           matchLevel(retval, tmp); add_inplace(retval, tmp)                 (:397-400)
// cts: [n][2][L][N]; masks: [n][L][N]; zero_ct: [2][Ltop][N] or NULL (a partial sum of a sharded batch);
// out: [2][L-1][N].
extern "C" int orc_collapse(const orc_ctx *c, size_t L, size_t n, const uint64_t *cts, size_t first_index, const uint64_t *masks,
                            const uint64_t *zero_ct, const uint32_t *elts, const uint64_t *const *gal, size_t ngal, uint64_t *out, int threads)
{
    const size_t N = orc_ctx_N(c), Ltop = orc_ctx_K(c) - 1, ct = 2 * L * N, oc = 2 * (L - 1) * N;
    if (!is_ckks(c) || L < 2) return -3;
    if (threads <= 0) threads = orc_max_threads();
    std::vector<u64> part(n * oc);
    int err = 0;
#pragma omp parallel for num_threads(threads)
    for (long i = 0; i < (long)n; i++) {
        std::vector<u64> tmp(cts + i * ct, cts + (i + 1) * ct), prod(ct);
        const size_t gi = first_index + (size_t)i;
        if (gi > 0) {
            int rc = orc_rotate(c, L, tmp.data(), -(int)gi, elts, gal, ngal);
            if (rc) err = rc;
        }
        orc_multiply_plain(c, L, 2, tmp.data(), masks + (size_t)i * L * N, prod.data());
        orc_rescale(c, L, 2, prod.data(), part.data() + i * oc);
    }
    if (err) return err;
    if (zero_ct) drop_to(N, 2, Ltop, L - 1, zero_ct, out);
    else memset(out, 0, oc * sizeof(u64));
    for (size_t i = 0; i < n; i++) orc_add(c, L - 1, 2, out, part.data() + i * oc, out);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// SEALContextWrapper::evaluatePolynomial (Horner) -- R/src/engine/seal_context.cpp:417-458.
//   retval = Enc(a_d)                                             (:440, injected: seed, top level)
//   for a_{d-1} .. a_0:  matchLevel(x, retval); retval *= x; relinearize; rescale;        (:444-448)
//                        mod_switch_to(plain a_k, retval.parms_id); add_plain             (:451-454)
// x: [2][Lx][N] (switched down in place as the loop descends, like the reference's cipher_input);
// seed: [2][Ltop][N]; coeffs: [ncoef][Ltop][N] = the plaintexts of a_{d-1}, ..., a_0 in that order (top level, NTT form);
// out: [2][Lout][N].  Returns Lout (> 0) or a negative error.
extern "C" int orc_horner(const orc_ctx *c, size_t Lx, const uint64_t *x, const uint64_t *seed, size_t ncoef, const uint64_t *coeffs,
                          const uint64_t *relin, uint64_t *out)
{
    const size_t N = orc_ctx_N(c), Ltop = orc_ctx_K(c) - 1;
    if (!is_ckks(c)) return -3;
    std::vector<u64> xv(x, x + 2 * Lx * N), r(seed, seed + 2 * Ltop * N), t3, t2, pl;
    size_t lx = Lx, lr = Ltop;
    for (size_t k = 0; k < ncoef; k++) {
        // matchLevel: the operand with more limbs is switched down (limbs dropped)
        if (lx > lr) { drop_to(N, 2, lx, lr, xv.data(), xv.data()); lx = lr; }
        else if (lr > lx) { drop_to(N, 2, lr, lx, r.data(), r.data()); lr = lx; }
        if (lr < 2) return -4;
        t3.assign(3 * lr * N, 0);
        t2.assign(2 * lr * N, 0);
        orc_ckks_multiply(c, lr, r.data(), xv.data(), t3.data());
        orc_relinearize(c, lr, t3.data(), relin, t2.data());
        orc_rescale(c, lr, 2, t2.data(), r.data());
        lr -= 1;
        pl.assign(coeffs + k * Ltop * N, coeffs + k * Ltop * N + lr * N);   // first lr limbs of the top-level plaintext
        orc_add_plain(c, lr, 2, r.data(), pl.data(), r.data());
    }
    memcpy(out, r.data(), 2 * lr * N * sizeof(u64));
    return (int)lr;
}

// ---------------------------------------------------------------------------------------------------------
// LogRegHornerBenchmark::operate -- R/src/benchmarks/ckks/seal_ckks_logreg_horner.cpp:388-481.
//   per sample:  multiply(W, X_i); relinearize; accumulateCKKS(n_features); rescale_to_next     (:425-444)
//   lr = collapseCKKS(dots, do_rotate = true)                                                     (:455)
//   matchLevel(b, lr); add_inplace(lr, b)                                                         (:461-465)
//   result = evaluatePolynomial(lr, coefficients)                                                 (:476)
// W, b: [2][Ltop][N]; X: [batch][2][Ltop][N]; masks: [batch][Ltop-1][N] (encode(e_i) switched to the dots' level);
// zero_ct, seed: injected fresh encryptions (top level); coeffs as in orc_horner.  out: [2][Lout][N]; returns Lout.
extern "C" int orc_logreg(const orc_ctx *c, size_t n_features, size_t batch, const uint64_t *W, const uint64_t *b, const uint64_t *X,
                          const uint64_t *masks, const uint64_t *zero_ct, const uint64_t *seed, size_t ncoef, const uint64_t *coeffs,
                          const uint64_t *relin, const uint32_t *elts, const uint64_t *const *gal, size_t ngal, uint64_t *out, int threads)
{
    const size_t N = orc_ctx_N(c), Ltop = orc_ctx_K(c) - 1, ct = 2 * Ltop * N;
    if (!is_ckks(c) || Ltop < 3) return -3;
    if (threads <= 0) threads = orc_max_threads();
    const size_t Ld = Ltop - 1, dct = 2 * Ld * N;
    std::vector<u64> dots(batch * dct);
    int err = 0;
#pragma omp parallel for num_threads(threads)
    for (long i = 0; i < (long)batch; i++) {
        std::vector<u64> t3(3 * Ltop * N), t2(ct);
        orc_ckks_multiply(c, Ltop, W, X + i * ct, t3.data());
        orc_relinearize(c, Ltop, t3.data(), relin, t2.data());
        int rc = orc_accumulate(c, Ltop, t2.data(), n_features, elts, gal, ngal);
        if (rc) err = rc;
        orc_rescale(c, Ltop, 2, t2.data(), dots.data() + i * dct);
    }
    if (err) return err;
    const size_t Lc = Ld - 1;
    std::vector<u64> lr(2 * Lc * N), bb(2 * Lc * N);
    int rc = orc_collapse(c, Ld, batch, dots.data(), 0, masks, zero_ct, elts, gal, ngal, lr.data(), threads);
    if (rc) return rc;
    drop_to(N, 2, Ltop, Lc, b, bb.data());
    orc_add(c, Lc, 2, lr.data(), bb.data(), lr.data());
    return orc_horner(c, Lc, lr.data(), seed, ncoef, coeffs, relin, out);
}
