/*
 * b200he.h -- C ABI of the B200 ciphertext-evaluation library (libb200he.so).
 *
 * This is the drop-in boundary for the hot path of hebench/reference-seal-backend: the
 * reference's src/engine + src/benchmarks call seal::Evaluator on host seal::Ciphertext
 * objects; a backend built on this library calls the entry points below on device-resident
 * batches instead.  Every evaluator entry cites the reference call site(s) it replaces
 * (R/ = /root/reference/).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns 0 on success, non-zero on error
 *    (b200he_last_error() returns the message for the calling thread);
 *  - one context = one GPU + one CUDA stream; all evaluator calls are asynchronous on that
 *    stream, b200he_ctx_sync / b200he_batch_download block until the results exist;
 *  - memory layouts are SEAL's: ciphertext = uint64[size][L][N], key-switch key =
 *    uint64[L_top][2][K][N], plaintext (CKKS, NTT form) = uint64[L][N];
 *  - a batch is `count` ciphertexts (or plaintexts, size = 1) sharing size, level L, NTT flag
 *    and scale, stored contiguously in HBM as uint64[count][size][L][N];
 *  - index maps (`ai`, `bi`: host arrays of n uint32, or NULL for the identity) select which
 *    input ciphertext feeds output i, so the reference's b0 x b1 result grids
 *    (R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:325-345) need no copies;
 *  - outputs are written into an existing batch handle which is reshaped as needed; an output
 *    may alias an input only where stated.
 *  - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef B200HE_H
#define B200HE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200he_ctx b200he_ctx;
typedef struct b200he_batch b200he_batch;

enum { B200HE_BFV = 1, B200HE_CKKS = 2 };

const char *b200he_last_error(void);
/* library build id, e.g. "b200he sm_100a <date>" */
const char *b200he_version(void);

/* ---- context: replaces seal::SEALContext + seal::Evaluator construction
 *      (R/src/engine/seal_context.cpp:90,119 and :56).  moduli = key-level chain (special
 *      prime last), psi = minimal primitive 2N-th roots per prime (as held by SEAL's
 *      NTTTables), plain_modulus = t for BFV (0 for CKKS). ---- */
int b200he_ctx_create(int scheme, uint32_t N, uint32_t K, const uint64_t *moduli, const uint64_t *psi,
                      uint64_t plain_modulus, int device, b200he_ctx **out);
void b200he_ctx_destroy(b200he_ctx *ctx);
/* run on an existing CUDA stream (cudaStream_t passed as void*), e.g. torch's current stream.  Every call is ordered on that
 * stream; contexts with N = 32768 also own a side stream for one pair of key-switch launches, forked from and joined back into
 * the caller's stream with events inside the call, so the ordering the caller sees does not change. */
int b200he_ctx_set_stream(b200he_ctx *ctx, void *cuda_stream);
int b200he_ctx_sync(b200he_ctx *ctx);
/* scratch budget per evaluator call in bytes (large batches are processed in chunks) */
int b200he_ctx_set_workspace(b200he_ctx *ctx, uint64_t bytes);

/* ---- keys: replaces seal::RelinKeys / seal::GaloisKeys held by SEALContextWrapper
 *      (R/include/engine/seal_context.h:101-102, generated at R/src/engine/seal_context.cpp:53,69).
 *      Uploaded once, HBM-resident. ---- */
int b200he_set_relin_key(b200he_ctx *ctx, const uint64_t *key);
int b200he_set_galois_key(b200he_ctx *ctx, uint32_t galois_elt, const uint64_t *key);
int b200he_has_galois_key(const b200he_ctx *ctx, uint32_t galois_elt);

/* ---- batches: device-resident std::vector<seal::Ciphertext> / <seal::Plaintext>.
 *      upload/download are the H2D/D2D halves of BaseBenchmark::load / store
 *      (R/src/benchmarks/ckks/seal_ckks_dot_product_benchmark.cpp:267-291). ---- */
int b200he_batch_create(b200he_ctx *ctx, b200he_batch **out);
void b200he_batch_destroy(b200he_batch *b);
int b200he_batch_resize(b200he_batch *b, uint64_t count, int size, int L, int ntt_form, double scale);
int b200he_batch_upload(b200he_batch *b, uint64_t first, uint64_t n, const uint64_t *host);
int b200he_batch_download(const b200he_batch *b, uint64_t first, uint64_t n, uint64_t *host);
/* same without the wait: `host` (pinned memory) is valid after the next b200he_ctx_sync */
int b200he_batch_download_async(const b200he_batch *b, uint64_t first, uint64_t n, uint64_t *host);
/* The whole of BaseBenchmark::load / store for a parameter: n separately allocated pageable host ciphertexts
 * (host[i] = seal::Ciphertext::data() of ciphertext first+i, ct_words each; R/src/benchmarks/ckks/
 * seal_ckks_dot_product_benchmark.cpp:267-291) <-> ciphertexts [first, first+n) of the batch.  Staged through the
 * context's double-buffered pinned memory: host threads gather/scatter one chunk while the copy engine moves the
 * other.  upload returns once every host[i] has been read (the last H2D copy may still be in flight on the context's
 * stream); download returns when every host[i] is complete. */
int b200he_batch_upload_scattered(b200he_batch *b, uint64_t first, uint64_t n, const uint64_t *const *host);
int b200he_batch_download_scattered(const b200he_batch *b, uint64_t first, uint64_t n, uint64_t *const *host);
/* dst (a batch of any context, any GPU) = a copy of src (another context / another GPU): device to device over NVLink
 * (cudaMemcpyPeerAsync), stream-ordered on both contexts, no host round trip and no wait.  The one exchange step of the
 * path: the per-GPU partial sums of logistic regression's collapse meet on GPU 0
 * (R/src/engine/seal_context.cpp:397-400 is the reference's mutex-guarded add_inplace). */
int b200he_batch_copy_from(b200he_batch *dst, const b200he_batch *src);
uint64_t b200he_batch_count(const b200he_batch *b);
int b200he_batch_size(const b200he_batch *b);
int b200he_batch_level(const b200he_batch *b);      /* L = number of RNS limbs */
int b200he_batch_ntt_form(const b200he_batch *b);
double b200he_batch_scale(const b200he_batch *b);
int b200he_batch_set_scale(b200he_batch *b, double scale);   /* R/src/engine/seal_context.cpp:395,399,452 */
void *b200he_batch_device_ptr(b200he_batch *b);

/* ---- evaluator ---- */
/* Evaluator::add / add_inplace: out[i] = a[ai[i]] + b[bi[i]]  (out may alias a)
 * R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:340, R/src/engine/seal_context.cpp:303,338,400 */
int b200he_add(b200he_ctx *ctx, const b200he_batch *a, const uint32_t *ai, const b200he_batch *b, const uint32_t *bi,
               uint64_t n, b200he_batch *out);
int b200he_sub(b200he_ctx *ctx, const b200he_batch *a, const uint32_t *ai, const b200he_batch *b, const uint32_t *bi,
               uint64_t n, b200he_batch *out);
/* Evaluator::multiply (size 2 x size 2 -> size 3; CKKS dyadic tensor product, BFV BEHZ)
 * R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:343, R/src/benchmarks/bfv/seal_bfv_element_wise_benchmark.cpp:326,
 * R/src/benchmarks/ckks/seal_ckks_dot_product_benchmark.cpp:325, R/src/engine/seal_context.cpp:446 */
int b200he_multiply(b200he_ctx *ctx, const b200he_batch *a, const uint32_t *ai, const b200he_batch *b, const uint32_t *bi,
                    uint64_t n, b200he_batch *out);
/* The inner loop of MatMult CipherBatchAxis: a = `rows x inner` matrix of ciphertexts (row-major), bt = the right-hand
 * `inner x cols` matrix stored by COLUMNS (`cols x inner`: a row of a and a column of b are both contiguous runs);
 * out[i * cols + j] = sum_k a[i * inner + k] * bt[j * inner + k] as size-3 ciphertexts -- `inner` Evaluator::multiply calls
 * and inner - 1 Evaluator::add_inplace calls per cell in the reference, one pass with register accumulators here; same
 * bits (sums of residues are exact).  CKKS, NTT form, out must not alias an input.
 * R/src/benchmarks/ckks/seal_ckks_matmult_cipherbatchaxis_benchmark.cpp:385-422 */
int b200he_matmul_accumulate(b200he_ctx *ctx, const b200he_batch *a, const b200he_batch *bt, uint64_t rows, uint64_t inner,
                             uint64_t cols, b200he_batch *out);
/* Evaluator::relinearize_inplace (size 3 -> 2; size 2 is a no-op copy)  (out may alias in)
 * R/src/benchmarks/ckks/seal_ckks_dot_product_benchmark.cpp:329, R/src/engine/seal_context.cpp:390,447 */
int b200he_relinearize(b200he_ctx *ctx, const b200he_batch *in, b200he_batch *out);
/* Evaluator::rotate_vector / rotate_rows with SEAL's NAF fallback when no key exists for the step
 * R/src/engine/seal_context.cpp:302,337,378, R/src/benchmarks/ckks/seal_ckks_matmult_row_benchmark.cpp:507  (out may alias in) */
int b200he_rotate(b200he_ctx *ctx, const b200he_batch *in, int step, b200he_batch *out);
/* per-ciphertext rotation steps: out[i] = rotate_vector(in[i], steps[i]) with SEAL's NAF decomposition per
 * ciphertext; ciphertexts sharing a power-of-two term are key-switched together.
 * R/src/engine/seal_context.cpp:378 (collapseCKKS rotates sample i by -i)  (out may alias in) */
int b200he_rotate_each(b200he_ctx *ctx, const b200he_batch *in, const int32_t *steps, b200he_batch *out);
/* out (one ciphertext) = sum over the batch: the mutex-guarded add_inplace reduction of
 * R/src/engine/seal_context.cpp:397-400 */
int b200he_sum(b200he_ctx *ctx, const b200he_batch *in, b200he_batch *out);
/* Evaluator::rotate_columns_inplace (BFV) / complex_conjugate (CKKS): Galois element 2N-1
 * R/src/engine/seal_context.cpp:308 */
int b200he_rotate_columns(b200he_ctx *ctx, const b200he_batch *in, b200he_batch *out);
/* Evaluator::apply_galois_inplace with an explicit Galois element */
int b200he_apply_galois(b200he_ctx *ctx, const b200he_batch *in, uint32_t galois_elt, b200he_batch *out);
/* Evaluator::rescale_to_next_inplace (CKKS) / mod_switch_to_next_inplace (BFV ciphertext): drop the
 * last prime with rounding.  R/src/engine/seal_context.cpp:391,448, R/src/benchmarks/ckks/seal_ckks_matmultval_benchmark.cpp:255 */
int b200he_rescale_to_next(b200he_ctx *ctx, const b200he_batch *in, b200he_batch *out);
/* Evaluator::relinearize_inplace immediately followed by Evaluator::rescale_to_next_inplace -- the pair every CKKS
 * multiply of the matmul / logistic-regression workloads issues (R/src/benchmarks/ckks/seal_ckks_matmultval_benchmark.cpp:252-255,
 * R/src/engine/seal_context.cpp:390-391,447-448).  Bit-identical to the two calls; the transform being linear over Z_q,
 * the fused form needs 25 % (L = 2) to 18 % (L = 6) fewer NTTs.  BFV / size-2 inputs take the two plain calls. */
int b200he_relinearize_rescale(b200he_ctx *ctx, const b200he_batch *in, b200he_batch *out);
/* Evaluator::mod_switch_to_inplace on CKKS ciphertexts / plaintexts: drop limbs down to L_target
 * R/src/engine/seal_context.cpp:260,262,388,451 */
int b200he_mod_drop(b200he_ctx *ctx, const b200he_batch *in, int L_target, b200he_batch *out);
/* Evaluator::multiply_plain_inplace (NTT form): out[i] = ct[i] (.) plain[pi[i]]   R/src/engine/seal_context.cpp:389 */
int b200he_multiply_plain(b200he_ctx *ctx, const b200he_batch *ct, const b200he_batch *plain, const uint32_t *pi,
                          b200he_batch *out);
/* Evaluator::add_plain_inplace (CKKS): out[i].c0 = ct[i].c0 + plain[pi[i]]   R/src/engine/seal_context.cpp:454 */
int b200he_add_plain(b200he_ctx *ctx, const b200he_batch *ct, const b200he_batch *plain, const uint32_t *pi,
                     b200he_batch *out);
/* SEALContextWrapper::accumulateCKKS / accumulateBFV (R/src/engine/seal_context.cpp:289-347):
 * ceil(log2 count) x (rotate by 2^k, add), plus the column swap for BFV when count > N/2.  In place. */
int b200he_accumulate(b200he_ctx *ctx, b200he_batch *inout, uint64_t count);
/* gather: out[i] = in[idx[i]] (device-side copy used by store() and by workload drivers) */
int b200he_gather(b200he_ctx *ctx, const b200he_batch *in, const uint32_t *idx, uint64_t n, b200he_batch *out);
/* raw per-limb transforms of a whole batch (K1/K2), used by tests and the NTT limb-ops/s metric */
int b200he_ntt_forward(b200he_ctx *ctx, const b200he_batch *in, b200he_batch *out);
int b200he_ntt_inverse(b200he_ctx *ctx, const b200he_batch *in, b200he_batch *out);

/* ---- measurement hooks ---- */
enum {
    B200HE_KERN_NTT_FWD = 0, B200HE_KERN_NTT_INV, B200HE_KERN_TENSOR_MAC, B200HE_KERN_KS_INNER, B200HE_KERN_MODDOWN,
    B200HE_KERN_ELEMENTWISE, B200HE_KERN_TENSOR, B200HE_KERN_GALOIS, B200HE_KERN_COPY, B200HE_KERN_BEHZ, B200HE_KERN_COUNT
};
/* kernels launched by this context since creation */
uint64_t b200he_launch_count(const b200he_ctx *ctx);
/* per-kernel-class CUDA-event timing: begin enables event recording around every launch; end syncs and
 * returns accumulated milliseconds and launch counts per class (arrays of B200HE_KERN_COUNT) */
int b200he_profile_begin(b200he_ctx *ctx);
int b200he_profile_end(b200he_ctx *ctx, double *ms, uint64_t *launches);
/* algorithmic work of the launches recorded between profile_begin and profile_end, per kernel class (arrays of
 * B200HE_KERN_COUNT; valid until the next profile_begin): modular-butterfly equivalents executed on the integer pipe and on
 * the FP64 pipe (the unit the pipes' peaks are measured in), and the bytes that have to cross HBM.  The
 * roofline fractions of ANY workload follow from these and the measured times, without per-workload formulas. */
int b200he_profile_work(const b200he_ctx *ctx, double *bfly_int, double *bfly_fp64, double *bytes);
const char *b200he_kernel_name(int kern_class);

#ifdef __cplusplus
}
#endif
#endif
