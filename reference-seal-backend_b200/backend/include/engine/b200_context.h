// b200_context.h -- SEALContextWrapper of the B200 backend: same role and method names as
// R/include/engine/seal_context.h:13-111, but the Evaluator half lives in HBM.
//   * host side (untimed by HEBench): parameter chain, keygen, encode, encrypt, decrypt -- the host FHE
//     stand-in (reference-seal-backend_b200/hostfhe) in place of seal::KeyGenerator / Encryptor /
//     Decryptor / CKKSEncoder / BatchEncoder (SEAL is not available offline);
//   * device side (the timed operate() path): one b200he context per GPU with the relinearization and
//     Galois keys resident in HBM; matchLevel / accumulate / collapse / evaluatePolynomial act on
//     device batches through the C ABI of include/b200he.h.
#pragma once
#include <functional>
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
    DeviceBatchPtr upload(int g, const std::vector<Ciphertext> &src, const std::vector<std::size_t> &items) const;   // src[items[.]]
    DeviceBatchPtr upload(int g, const Ciphertext &src) const;
    DeviceBatchPtr uploadPlain(int g, const std::vector<Plaintext> &src) const;
    // D2H (blocks until the batch's stream has produced the data)
    std::vector<Ciphertext> download(const DeviceBatch &b, std::shared_ptr<HostSlab> slab = nullptr) const;
    // called by operate() once a result's kernels are enqueued and before it waits for them: host memory for store()
    void prepareStore(ShardedCiphertexts &res) const;
    void syncAll() const;
    // split [0, n) into gpuCount() contiguous blocks
    std::vector<std::uint64_t> partition(std::uint64_t n) const;
    // fn(g) for every GPU, on one host thread per GPU (every GPU's copies and launches are issued concurrently); the first
    // exception is rethrown on the caller's thread once all have finished
    void forEachGpu(const std::function<void(int)> &fn) const;
    // operand placement of a result grid: items of src0 / src1 come in units of unit0 / unit1 consecutive ciphertexts
    // (a matrix row of CipherBatchAxis is cols_M0 ciphertexts); items1 (optional) lists src1's ciphertexts per item when
    // they are not consecutive (a matrix column)
    GridOperands loadGrid(const std::vector<Ciphertext> &src0, std::size_t unit0, const std::vector<Ciphertext> &src1, std::size_t unit1,
                          const std::function<std::vector<std::size_t>(std::size_t)> &items1 = nullptr) const;
    // GPU g's share of the grid [v0[0], v0[0] + b[0]) x [v0[1], v0[1] + b[1]) (result r = i_rel * b[1] + j_rel)
    GridShare gridShare(const GridOperands &in, int g, const std::uint64_t v0[2], const std::uint64_t b[2]) const;
    // D2H of a sharded result vector, one host thread per GPU, items placed by global index
    std::vector<Ciphertext> gather(const ShardedCiphertexts &src) const;
    // every call into libb200he goes through check(): a non-zero return becomes HEBenchError(HEBSEAL_ECODE_SEAL_ERROR)
    void check(int rc, const char *what) const;
    // Diagnostic trace (off unless HEB_B200_TRACE_DIR names a directory): the ciphertexts that cross load() / store() and
    // the fresh encryptions drawn inside operate() (R/src/engine/seal_context.cpp:360,440) are written to <dir>/<tag>.bin
    // (header: 8 x u64 = magic "B200TRC1", items written, size, L, N, ntt form, scale as IEEE bits, items in the vector; then the
    // uint64 data; HEB_B200_TRACE_PICK_<tag>="i,j,..." restricts a tag to those items).
    // tests/test_workload_parity.py replays them through the CPU oracle's workload bodies and compares bit for bit.
    // operate() brackets.  endOperate waits for every GPU (the harness times operate() by wall clock).  With
    // HEB_B200_PROFILE_JSON=<file> the operate() calls after the first HEB_B200_PROFILE_SKIP ones run with per-kernel CUDA
    // events (b200he_profile_*) and append one JSON line each: per kernel class the milliseconds, launches and algorithmic
    // work summed over the GPUs -- bench.py derives the per-workload roofline fractions from it.
    void beginOperate();
    void endOperate(std::uint64_t results);
    bool tracing() const { return !m_trace_dir.empty(); }
    void trace(const std::string &tag, const std::vector<Ciphertext> &v) const;
    void trace(const std::string &tag, const Ciphertext &c) const { trace(tag, std::vector<Ciphertext>(1, c)); }

    // composite operations of R/src/engine/seal_context.cpp:255-458, on device batches
    void matchLevel(DeviceBatch &a, DeviceBatch &b) const;
    void accumulateBFV(DeviceBatch &cipher, std::size_t count) const;
    void accumulateCKKS(DeviceBatch &cipher, std::size_t count) const;
    // sum_i mask_i (.) rotate(ciphers[i], -(first_index + i)); masks encoded for `total` samples.  One ciphertext out.
    // encrypted_zero: the fresh Enc(0) the reference draws at R/src/engine/seal_context.cpp:360 (nullptr: this GPU's shard
    // yields a partial sum only)
    DeviceBatchPtr collapseCKKS(DeviceBatch &ciphers, std::size_t first_index, std::size_t total, const Ciphertext *encrypted_zero);
    // masks of collapseCKKS for samples [first_index, first_index + n) of `total`, resident on GPU g (cached)
    DeviceBatchPtr maskBatch(int g, std::size_t first_index, std::size_t n, std::size_t total, int level);
    // Horner evaluation; seed = the fresh Enc(a_d) of R/src/engine/seal_context.cpp:440; coefficient plaintexts a_{d-1} .. a_0
    // are taken from coeffBatch() (device-resident, one per level, cached)
    DeviceBatchPtr evaluatePolynomial(DeviceBatch &cipher_input, const std::vector<Plaintext> &plain_coefficients, const Ciphertext &seed);
    DeviceBatchPtr coeffBatch(int g, const std::vector<Plaintext> &plain_coefficients, std::size_t index, int level);

private:
    SEALContextWrapper() {}
    void init(bool ckks, std::size_t N, std::size_t depth, int coeff_bits, int scale_or_plain_bits);
    bool m_ckks = true;
    std::size_t m_N = 0, m_K = 0;
    int m_scale_bits = 0;
    double m_scale   = 1.0;
    std::uint64_t m_t = 0;
    hfhe_ctx *m_host  = nullptr;
    std::vector<b200he_ctx *> m_dev;
    std::vector<int> m_dev_of_ctx;
    std::string m_trace_dir, m_profile_path;
    long m_profile_skip = 0, m_operate_calls = 0;
    bool m_profiling = false;
    std::map<std::string, DeviceBatchPtr> m_mask_cache;   // collapse masks per (gpu, first, n, total, level); Horner coefficients per (gpu, index, level)
    mutable std::mutex m_cache_mtx;
};

}   // namespace sbe
