/*
 * he_oracle.h -- CPU oracle for the ciphertext-evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (reference-seal-backend_b200/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker or as
 * the reported CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic of the reference lives in Microsoft SEAL
 * (tag v3.7.2, /root/reference/cmake/third-party/SEAL.version:1), which is
 * fetched from the network at configure time
 * (/root/reference/cmake/third-party/SEAL.cmake:5-12) and is absent here.  The
 * reference ships no golden vectors.  This file restates SEAL's
 * published algorithms and is anchored on the reference's
 * call sites: every function cites the reference file:line that reaches it.
 *
 * Layout conventions (same as seal::Ciphertext):
 *   ciphertext  = uint64[size][L][N]   (poly-major, limb, coefficient)
 *   kswitch key = uint64[L_top][2][K][N]  (digit, component, key limb, coeff)
 * K = number of primes in the key-level chain, L_top = K-1, special prime = q[K-1].
 */
#ifndef HE_ORACLE_H
#define HE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

enum { ORC_SCHEME_BFV = 1, ORC_SCHEME_CKKS = 2 };

/* ---- number theory (SEAL util/numth.cpp, modulus.cpp restated) ---- */
int orc_is_prime(uint64_t v);
/* get_primes(factor, bits, count): descending primes = 1 mod factor, < 2^bits */
int orc_get_primes(uint64_t factor, int bits, size_t count, uint64_t *out);
/* CoeffModulus::Create(N, bit_sizes) -- R/src/engine/seal_context.cpp:89,117 */
int orc_coeff_modulus_create(size_t N, const int *bits, size_t n, uint64_t *out);
/* PlainModulus::Batching(N, bits) -- R/src/engine/seal_context.cpp:118 */
uint64_t orc_plain_modulus_batching(size_t N, int bits);
/* try_minimal_primitive_root(2N, q) */
int orc_minimal_primitive_root(uint64_t degree, uint64_t q, uint64_t *root);
/* naf(v): returns number of terms written */
int orc_naf(int value, int *out);
/* GaloisTool::get_elt_from_step */
uint32_t orc_galois_elt_from_step(int step, size_t N);
/* GaloisTool::get_elts_all: returns count (2*(log2N-1)+1) */
int orc_galois_elts_all(size_t N, uint32_t *out);
/* Galois NTT-form permutation table (generate_table_ntt) */
void orc_galois_table_ntt(size_t N, uint32_t elt, uint32_t *table);

/* ---- context ---- */
orc_ctx *orc_ctx_create(int scheme, size_t N, size_t K, const uint64_t *moduli, uint64_t plain_modulus);
void orc_ctx_destroy(orc_ctx *c);
uint64_t orc_ctx_psi(const orc_ctx *c, size_t limb);
size_t orc_ctx_N(const orc_ctx *c);
size_t orc_ctx_K(const orc_ctx *c);
int orc_ctx_scheme(const orc_ctx *c);
/* number of BEHZ auxiliary primes |Bsk| (BFV only) and their values (Bsk..., m_sk last) */
size_t orc_ctx_bsk_size(const orc_ctx *c);
void orc_ctx_bsk(const orc_ctx *c, uint64_t *out);

/* ---- K1/K2: single-limb transforms, in place; limb = index into the key-level chain ---- */
void orc_ntt_fwd(const orc_ctx *c, size_t limb, uint64_t *poly);
void orc_ntt_inv(const orc_ctx *c, size_t limb, uint64_t *poly);
/* O(N^2) definition: out[i] = sum_j a_j psi^{(2 brv(i)+1) j}; for self-tests */
void orc_ntt_fwd_direct(const orc_ctx *c, size_t limb, const uint64_t *in, uint64_t *out);

/* ---- K3: Evaluator::add / add_inplace (size = polys), sub, negate ---- */
void orc_add(const orc_ctx *c, size_t L, size_t size, const uint64_t *a, const uint64_t *b, uint64_t *out);
void orc_sub(const orc_ctx *c, size_t L, size_t size, const uint64_t *a, const uint64_t *b, uint64_t *out);
/* ---- K4: Evaluator::multiply (CKKS, 2x2 -> 3) ---- */
void orc_ckks_multiply(const orc_ctx *c, size_t L, const uint64_t *a, const uint64_t *b, uint64_t *out3);
/* ---- K5: Evaluator::multiply (BFV, BEHZ, 2x2 -> 3) at the top data level ---- */
void orc_bfv_multiply(const orc_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out3);
/* ---- K6: Evaluator::switch_key_inplace.  ct: [2][L][N] in/out, target: [L][N] ---- */
void orc_switch_key(const orc_ctx *c, size_t L, uint64_t *ct, const uint64_t *target, const uint64_t *key);
/* ---- K7: relinearize_inplace: ct3 [3][L][N] -> out2 [2][L][N] ---- */
void orc_relinearize(const orc_ctx *c, size_t L, const uint64_t *ct3, const uint64_t *relin_key, uint64_t *out2);
/* ---- K8: apply_galois_inplace on a size-2 ct (in place) ---- */
void orc_apply_galois(const orc_ctx *c, size_t L, uint64_t *ct, uint32_t elt, const uint64_t *galois_key);
/* rotate_vector / rotate_rows with NAF fallback.  keys given as parallel arrays.
 * returns 0 ok, -1 if a needed key is missing */
int orc_rotate(const orc_ctx *c, size_t L, uint64_t *ct, int step, const uint32_t *elts,
               const uint64_t *const *keys, size_t nkeys);
/* ---- K9: rescale_to_next / BFV mod_switch_to_next: [size][L][N] -> [size][L-1][N] ---- */
void orc_rescale(const orc_ctx *c, size_t L, size_t size, const uint64_t *in, uint64_t *out);
/* ---- K10: mod_switch_drop_to_next (CKKS ct / plain): drop the last limb ---- */
void orc_mod_drop(const orc_ctx *c, size_t L, size_t size, const uint64_t *in, uint64_t *out);
/* ---- K11: multiply_plain (NTT form) / add_plain ---- */
void orc_multiply_plain(const orc_ctx *c, size_t L, size_t size, const uint64_t *ct, const uint64_t *plain, uint64_t *out);
void orc_add_plain(const orc_ctx *c, size_t L, size_t size, const uint64_t *ct, const uint64_t *plain, uint64_t *out);

/* ---- composites that mirror the reference's operate() bodies, batched with OpenMP
 *      like the reference (R/src/benchmarks/ckks/seal_ckks_dot_product_benchmark.cpp:315) ---- */
/* multiply -> relinearize -> rescale (BASELINE.json configs[1]); a,b: [n][2][L][N], out: [n][2][L-1][N] */
void orc_batch_mul_relin_rescale(const orc_ctx *c, size_t L, size_t n, const uint64_t *a, const uint64_t *b,
                                 const uint64_t *relin_key, uint64_t *out, int threads);
/* accumulateCKKS / accumulateBFV-rows (R/src/engine/seal_context.cpp:321-347): in place */
int orc_accumulate(const orc_ctx *c, size_t L, uint64_t *ct, size_t count, const uint32_t *elts,
                   const uint64_t *const *keys, size_t nkeys);
/* dot product: multiply -> relinearize -> accumulate(n) per pair */
int orc_batch_dot(const orc_ctx *c, size_t L, size_t n, const uint64_t *a, const uint64_t *b, size_t count,
                  const uint64_t *relin_key, const uint32_t *elts, const uint64_t *const *keys, size_t nkeys,
                  uint64_t *out, int threads);
/* batched single-limb forward NTTs (for the NTT limb-ops/s baseline) */
void orc_batch_ntt(const orc_ctx *c, size_t limb, size_t n, uint64_t *polys, int inverse, int threads);
int orc_max_threads(void);
/* Barrett fast path of the hot loops vs the plain 128-bit `%` forms: number of mismatches over n random triples */
size_t orc_selftest_fastmod(uint64_t q, size_t n, uint64_t seed);

/* ---- workload bodies (he_oracle_workloads.cpp): the reference's operate() bodies and the composite helpers of
 *      R/src/engine/seal_context.cpp:255-458, fed the same input and injected fresh encryptions as the plugin; they
 *      check the ciphertexts store() returns, bit for bit.  Layouts: ciphertext arrays [count][size][L][N]. ---- */
/* MatMultVal (R/src/benchmarks/ckks/seal_ckks_matmultval_benchmark.cpp:235-270): out [r0*c1][2][Lout][N], Lout = L-1 CKKS / L BFV */
int orc_matmul_val(const orc_ctx *c, size_t L, size_t r0, size_t c0, size_t c1, const uint64_t *m0, const uint64_t *m1t,
                   const uint64_t *relin, const uint32_t *elts, const uint64_t *const *gal, size_t ngal, uint64_t *out, int threads);
/* MatMultRow (R/src/benchmarks/ckks/seal_ckks_matmult_row_benchmark.cpp:472-523): out [nA][2][L][N] */
int orc_matmul_row(const orc_ctx *c, size_t L, size_t nA, size_t dim2, int spacers, const uint64_t *A, const uint64_t *B,
                   const uint64_t *relin, const uint32_t *elts, const uint64_t *const *gal, size_t ngal, uint64_t *out, int threads);
/* MatMult CipherBatchAxis (R/src/benchmarks/ckks/seal_ckks_matmult_cipherbatchaxis_benchmark.cpp:385-441) */
int orc_matmul_cba(const orc_ctx *c, size_t L, size_t r0, size_t c0, size_t c1, const uint64_t *m0, const uint64_t *m1,
                   const uint64_t *relin, uint64_t *out, int threads);
/* collapseCKKS (R/src/engine/seal_context.cpp:349-415): out [2][L-1][N] */
int orc_collapse(const orc_ctx *c, size_t L, size_t n, const uint64_t *cts, size_t first_index, const uint64_t *masks,
                 const uint64_t *zero_ct, const uint32_t *elts, const uint64_t *const *gal, size_t ngal, uint64_t *out, int threads);
/* evaluatePolynomial (R/src/engine/seal_context.cpp:417-458): returns the level of out (> 0) or an error (< 0) */
int orc_horner(const orc_ctx *c, size_t Lx, const uint64_t *x, const uint64_t *seed, size_t ncoef, const uint64_t *coeffs,
               const uint64_t *relin, uint64_t *out);
/* LogRegHornerBenchmark::operate (R/src/benchmarks/ckks/seal_ckks_logreg_horner.cpp:388-481): returns the level of out */
int orc_logreg(const orc_ctx *c, size_t n_features, size_t batch, const uint64_t *W, const uint64_t *b, const uint64_t *X,
               const uint64_t *masks, const uint64_t *zero_ct, const uint64_t *seed, size_t ncoef, const uint64_t *coeffs,
               const uint64_t *relin, const uint32_t *elts, const uint64_t *const *gal, size_t ngal, uint64_t *out, int threads);

#ifdef __cplusplus
}
#endif
#endif
