// b200_context.h -- SEALContextWrapper of the B200 backend: same role and method names as
// R/include/engine/seal_context.h:13-111, but the Evaluator half lives in HBM.
//   * host side (untimed by HEBench): parameter chain, keygen, encode, encrypt, decrypt -- the host FHE
//     stand-in (reference-seal-backend_b200/hostfhe) in place of seal::KeyGenerator / Encryptor /
//     Decryptor / CKKSEncoder / BatchEncoder (SEAL is not available offline);
//   * device side (the timed operate() path): one b200he context per GPU with the relinearization and
//     Galois keys resident in HBM; matchLevel / accumulate / collapse / evaluatePolynomial act on
//     device batches through the C ABI of include/b200he.h.
#pragma once
#include <map>
#include <mutex>
#include <string>

#include "hebench/api_bridge/cpp/hebench.hpp"

#include "engine/b200_types.h"

struct hfhe_ctx;

namespace sbe {

class SEALContextWrapper
{
public:
    HEBERROR_DECLARE_CLASS_NAME(SEALContextWrapper)
    SEALContextWrapper(const SEALContextWrapper &) = delete;
    SEALContextWrapper &operator=(const SEALContextWrapper &) = delete;
    typedef std::shared_ptr<SEALContextWrapper> Ptr;

    // R/include/engine/seal_context.h:32-50: coeff_modulus = {60, bits x (num_coeff_moduli - 1), 60}
    static Ptr createCKKSContext(std::size_t poly_modulus_degree, std::size_t num_coeff_moduli, int coeff_moduli_bits, int scale_bits);
    static Ptr createBFVContext(std::size_t poly_modulus_degree, std::size_t num_coeff_moduli, int coeff_moduli_bits, int plaintext_modulus_bits = 20);
    ~SEALContextWrapper();

    // ---- host side
    bool isCKKS() const { return m_ckks; }
    std::size_t polyModulusDegree() const { return m_N; }
    std::size_t slotCount() const { return m_ckks ? m_N / 2 : m_N; }
    std::size_t topLevel() const { return m_K - 1; }
    double scale() const { return m_scale; }
    int scaleBits() const { return m_scale_bits; }
    std::uint64_t plainModulus() const { return m_t; }
    std::vector<int> coeffModulusBits() const;
    Plaintext encodeVector(const std::vector<double> &values);
    Plaintext encodeVector(const std::vector<double> &values, double scale);
    Plaintext encodeVector(const std::vector<std::int64_t> &values);
    std::vector<double> decodeCKKS(const Plaintext &plain);
    std::vector<std::int64_t> decodeBFV(const Plaintext &plain);
    Ciphertext encrypt(const Plaintext &plain);
    std::vector<Ciphertext> encrypt(const std::vector<Plaintext> &plain);
    Plaintext decrypt(const Ciphertext &cipher);
    std::vector<Plaintext> decrypt(const std::vector<Ciphertext> &cipher);

    // ---- device side
    int gpuCount() const { return (int)m_dev.size(); }
    b200he_ctx *device(int g) const { return m_dev[g]; }
    DeviceBatchPtr newBatch(int g) const { return std::make_shared<DeviceBatch>(m_dev[g]); }
    // H2D: the ciphertexts [first, first+n) of `src` become a batch on GPU g (all must share size/level/scale)
    DeviceBatchPtr upload(int g, const std::vector<Ciphertext> &src, std::size_t first, std::size_t n) const;
    DeviceBatchPtr upload(int g, const Ciphertext &src) const;
    DeviceBatchPtr uploadPlain(int g, const std::vector<Plaintext> &src) const;
    // D2H (blocks until the batch's stream has produced the data)
    std::vector<Ciphertext> download(const DeviceBatch &b) const;
    void syncAll() const;
    // split [0, n) into gpuCount() contiguous blocks
    std::vector<std::uint64_t> partition(std::uint64_t n) const;
    // every call into libb200he goes through check(): a non-zero return becomes HEBenchError(HEBSEAL_ECODE_SEAL_ERROR)
    void check(int rc, const char *what) const;
    // Diagnostic trace (off unless HEB_B200_TRACE_DIR names a directory): the ciphertexts that cross load() / store() and
    // the fresh encryptions drawn inside operate() (R/src/engine/seal_context.cpp:360,440) are written to <dir>/<tag>.bin
    // (header: 8 x u64 = magic "B200TRC1", items written, size, L, N, ntt form, scale as IEEE bits, items in the vector; then the
    // uint64 data; HEB_B200_TRACE_PICK_<tag>="i,j,..." restricts a tag to those items).
    // tests/test_workload_parity.py replays them through the CPU oracle's workload bodies and compares bit for bit.
    bool tracing() const { return !m_trace_dir.empty(); }
    void trace(const std::string &tag, const std::vector<Ciphertext> &v) const;
    void trace(const std::string &tag, const Ciphertext &c) const { trace(tag, std::vector<Ciphertext>(1, c)); }

    // composite operations of R/src/engine/seal_context.cpp:255-458, on device batches
    void matchLevel(DeviceBatch &a, DeviceBatch &b) const;
    void accumulateBFV(DeviceBatch &cipher, std::size_t count) const;
    void accumulateCKKS(DeviceBatch &cipher, std::size_t count) const;
    // sum_i mask_i (.) rotate(ciphers[i], -(first_index + i)); masks encoded for `total` samples.  One ciphertext out.
    DeviceBatchPtr collapseCKKS(DeviceBatch &ciphers, std::size_t first_index, std::size_t total, bool add_encrypted_zero);
    DeviceBatchPtr evaluatePolynomial(DeviceBatch &cipher_input, const std::vector<Plaintext> &plain_coefficients);

private:
    SEALContextWrapper() {}
    void init(bool ckks, std::size_t N, std::size_t depth, int coeff_bits, int scale_or_plain_bits);
    DeviceBatchPtr maskBatch(int g, std::size_t first_index, std::size_t n, std::size_t total, int level);

    bool m_ckks = true;
    std::size_t m_N = 0, m_K = 0;
    int m_scale_bits = 0;
    double m_scale   = 1.0;
    std::uint64_t m_t = 0;
    hfhe_ctx *m_host  = nullptr;
    std::vector<b200he_ctx *> m_dev;
    std::vector<int> m_dev_of_ctx;
    std::string m_trace_dir;
    std::map<std::string, DeviceBatchPtr> m_mask_cache;   // collapse masks per (gpu, first, n, total, level)
};

}   // namespace sbe
