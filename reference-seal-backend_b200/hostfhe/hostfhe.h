/*
 * hostfhe.h -- host-side FHE stand-in for the parts of Microsoft SEAL that stay on the
 * host in this design: parameter chain, keygen, encode, encrypt, decrypt, decode.
 *
 * In the reference these are seal::{KeyGenerator,Encryptor,Decryptor,CKKSEncoder,
 * BatchEncoder} owned by SEALContextWrapper (R/src/engine/seal_context.cpp:46-70,
 * R/include/engine/seal_context.h:52-62); they are untimed by HEBench and OUT of the
 * accelerated hot path.  SEAL is not available offline, so this
 * module provides format-compatible equivalents: same RNS prime chain, same ciphertext /
 * key-switch-key memory layout (uint64[size][L][N], uint64[L_top][2][K][N]), same slot
 * ordering (generator 3 index map) so that Galois element 3^k rotates slots left by k.
 * It is NOT the ciphertext-evaluation path: no Evaluator operation lives here.
 * Randomness: a ChaCha20 keystream keyed with 256 bits from the operating system (seed 0, the
 * default; SEAL's is a Blake2/Shake XOF seeded from the OS and the reference never fixes it,
 * R/src/engine/seal_context.cpp:87-90), so bit-equality with SEAL's *encryptions* is neither
 * possible nor required.  A non-zero seed makes keys and encryptions reproducible: that is for
 * tests and benchmarks only.  Secret: uniform ternary; noise: centred binomial (21 coin pairs),
 * SEAL's defaults.  This module is a BENCHMARK STAND-IN for host SEAL, not a vetted cryptographic
 * library: production deployments bind the real SEAL objects here.
 */
#ifndef HOSTFHE_H
#define HOSTFHE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hfhe_ctx hfhe_ctx;
enum { HFHE_BFV = 1, HFHE_CKKS = 2 };

/* coeff_modulus = {60, bits x (depth-1), 60}  (R/src/engine/seal_context.cpp:79-82,107-110);
 * scale_or_plain_bits: CKKS scale exponent, or BFV plain-modulus bits (PlainModulus::Batching).
 * seed: 0 = key material from OS entropy (getrandom); non-zero = reproducible (tests / benchmarks only).
 * Returns NULL for unsupported parameters or when no entropy source is available. */
hfhe_ctx *hfhe_create(int scheme, size_t N, size_t depth, int coeff_bits, int scale_or_plain_bits, uint64_t seed);
void hfhe_destroy(hfhe_ctx *c);
size_t hfhe_N(const hfhe_ctx *c);
size_t hfhe_K(const hfhe_ctx *c);                 /* primes at key level (= depth + 1) */
const uint64_t *hfhe_moduli(const hfhe_ctx *c);   /* K primes, special prime last */
const uint64_t *hfhe_psi(const hfhe_ctx *c);      /* minimal primitive 2N-th roots, per prime */
uint64_t hfhe_plain_modulus(const hfhe_ctx *c);
double hfhe_scale(const hfhe_ctx *c);

/* keys, SEAL KSwitchKeys layout [L_top][2][K][N], NTT form */
const uint64_t *hfhe_relin_key(hfhe_ctx *c);
/* Galois elements of KeyGenerator::create_galois_keys() (all power-of-two steps + 2N-1) */
size_t hfhe_galois_count(hfhe_ctx *c);
uint32_t hfhe_galois_elt(hfhe_ctx *c, size_t i);
const uint64_t *hfhe_galois_key(hfhe_ctx *c, uint32_t elt);   /* generated lazily, cached; NULL if elt invalid */
size_t hfhe_kswitch_key_words(const hfhe_ctx *c);             /* L_top*2*K*N */

/* CKKS: encode n doubles (rest zero) at the top data level, NTT form: out [L_top][N] */
void hfhe_ckks_encode(hfhe_ctx *c, const double *vals, size_t n, double scale, uint64_t *plain);
/* decode a plaintext at level L (NTT form) with the given scale into N/2 doubles (real parts) */
void hfhe_ckks_decode(hfhe_ctx *c, const uint64_t *plain, size_t L, double scale, double *out);
/* BFV: batch-encode n int64 (rest zero) -> plaintext coefficients mod t: out [N] */
void hfhe_bfv_encode(hfhe_ctx *c, const int64_t *vals, size_t n, uint64_t *plain);
void hfhe_bfv_decode(hfhe_ctx *c, const uint64_t *plain, int64_t *out /*N*/);

/* public-key encryption at the top data level: ct [2][L_top][N] (CKKS: NTT form; BFV: coeff form).
 * Thread safe: every encryption draws from its own keystream.  hfhe_encrypt takes the next free stream;
 * hfhe_reserve_encryptions(n) hands out n consecutive stream indices (returns the first) for callers that
 * encrypt a vector on several threads with hfhe_encrypt_at and still want reproducible bits under a fixed seed. */
void hfhe_encrypt(hfhe_ctx *c, const uint64_t *plain, uint64_t *ct);
uint64_t hfhe_reserve_encryptions(hfhe_ctx *c, uint64_t n);
void hfhe_encrypt_at(hfhe_ctx *c, const uint64_t *plain, uint64_t *ct, uint64_t index);
/* decrypt a size-`size` ciphertext at level L: CKKS -> plain [L][N] NTT form; BFV -> plain [N] mod t */
void hfhe_decrypt(hfhe_ctx *c, const uint64_t *ct, size_t size, size_t L, uint64_t *plain);

#ifdef __cplusplus
}
#endif
#endif
