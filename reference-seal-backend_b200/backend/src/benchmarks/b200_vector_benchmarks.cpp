// b200_vector_benchmarks.cpp -- shared description plumbing + element-wise add / multiply and dot product.
// Replaces R/src/benchmarks/{ckks,bfv}/seal_*_element_wise_benchmark.cpp and seal_*_dot_product_benchmark.cpp.
#include <cstring>
#include <sstream>

#include "benchmarks/b200_benchmarks.h"

namespace sbe {

using hebench::APIBridge::Category;
using hebench::APIBridge::DataPack;
using hebench::APIBridge::DataPackCollection;
using hebench::APIBridge::Handle;
using hebench::APIBridge::ParameterIndexer;
using hebench::APIBridge::Workload;
using hebench::cpp::HEBenchError;

// ------------------------------------------------------------------ description plumbing
void B200BenchmarkDescription::setup(bool ckks, Workload w, Category cat, std::int64_t other, const char *algo_name, const char *algo_desc,
                                     const std::vector<std::uint64_t> &workload_defaults, const std::vector<const char *> &workload_names,
                                     const EncryptionParams &d)
{
    m_ckks      = ckks;
    m_algo_name = algo_name;
    m_algo_desc = algo_desc;
    std::memset(&m_descriptor, 0, sizeof(m_descriptor));
    m_descriptor.workload          = w;
    m_descriptor.data_type         = ckks ? hebench::APIBridge::DataType::Float64 : hebench::APIBridge::DataType::Int64;
    m_descriptor.category          = cat;
    m_descriptor.cipher_param_mask = HEBENCH_HE_PARAM_FLAGS_ALL_CIPHER;
    m_descriptor.scheme            = ckks ? HEBENCH_HE_SCHEME_CKKS : HEBENCH_HE_SCHEME_BFV;
    m_descriptor.security          = HEBENCH_HE_SECURITY_128;
    m_descriptor.other             = other;
    if (cat == Category::Latency) m_descriptor.cat_params.latency.warmup_iterations_count = 1;
    // offline: data_count stays 0 = sample count chosen by the harness (R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:43-45)
    hebench::cpp::WorkloadParams::Common p;
    for (std::size_t i = 0; i < workload_defaults.size(); ++i) p.add<std::uint64_t>(workload_defaults[i], workload_names[i]);
    m_extra_start = workload_defaults.size();
    p.add<std::uint64_t>(d.poly_modulus_degree, "PolyModulusDegree");
    p.add<std::uint64_t>(d.multiplicative_depth, "MultiplicativeDepth");
    p.add<std::uint64_t>(d.coeff_modulus_bits, "CoefficientModulusBits");
    p.add<std::uint64_t>(d.scale_or_plain_bits, ckks ? "ScaleBits" : "PlainModulusBits");
    p.add<std::uint64_t>(d.num_threads, "NumThreads");
    addDefaultParameters(p);
}

EncryptionParams B200BenchmarkDescription::encryptionParams(const hebench::APIBridge::WorkloadParams &p) const
{
    if (p.count < m_extra_start + 5)
        throw HEBenchError(HEBERROR_MSG_CLASS("Invalid workload parameters: encryption parameters missing."), HEBENCH_ECODE_INVALID_ARGS);
    const hebench::APIBridge::WorkloadParam *w = p.params + m_extra_start;
    return EncryptionParams{ w[0].u_param, w[1].u_param, w[2].u_param, w[3].u_param, w[4].u_param };
}

void B200BenchmarkDescription::destroyBenchmark(hebench::cpp::BaseBenchmark *p_bench) { delete p_bench; }

// CSV text appended to the report header (R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:84-115)
std::string B200BenchmarkDescription::getBenchmarkDescription(const hebench::APIBridge::WorkloadParams *p_w_params) const
{
    std::stringstream ss;
    ss << ", Encryption parameters" << std::endl;
    if (p_w_params) {
        const EncryptionParams ep = encryptionParams(*p_w_params);
        ss << ", , HE Library, B200-native SEAL-compatible evaluator (libb200he)" << std::endl
           << ", , Poly modulus degree, " << ep.poly_modulus_degree << std::endl
           << ", , Coefficient Modulus, 60";
        for (std::uint64_t i = 1; i < ep.multiplicative_depth; ++i) ss << ", " << ep.coeff_modulus_bits;
        ss << ", 60" << std::endl;
        if (m_ckks) ss << ", , Scale, 2^" << ep.scale_or_plain_bits << std::endl;
        else ss << ", , Plain modulus bits, " << ep.scale_or_plain_bits << std::endl;
        const char *g = getenv("HEB_B200_GPUS");
        ss << ", Algorithm, " << m_algo_name << ", " << m_algo_desc << std::endl
           << ", GPUs, " << (g ? g : "1") << ", batch-sharded (NumThreads is not used by the GPU path)" << std::endl;
    }
    return ss.str();
}

SEALContextWrapper::Ptr makeContext(bool ckks, const EncryptionParams &ep)
{
    try {
        return ckks ? SEALContextWrapper::createCKKSContext(ep.poly_modulus_degree, ep.multiplicative_depth, (int)ep.coeff_modulus_bits,
                                                           (int)ep.scale_or_plain_bits)
                    : SEALContextWrapper::createBFVContext(ep.poly_modulus_degree, ep.multiplicative_depth, (int)ep.coeff_modulus_bits,
                                                          (int)ep.scale_or_plain_bits);
    } catch (HEBenchError &) {
        throw;
    } catch (std::exception &ex) {
        throw HEBenchError(ex.what(), HEBSEAL_ECODE_SEAL_ERROR);
    }
}

// ------------------------------------------------------------------ descriptions
static EncryptionParams vectorDefaults(bool ckks, VectorOp op)
{
    // R/include/benchmarks/{ckks,bfv}/seal_*_{element_wise,dot_product}_benchmark.h:23-29
    if (ckks) return EncryptionParams{ 8192, 2, op == VectorOp::Dot ? 40u : 45u, op == VectorOp::Dot ? 40u : 45u, 0 };
    return EncryptionParams{ 8192, 2, op == VectorOp::Dot ? 45u : 40u, 20, 0 };
}

template <bool CKKS> ElementWiseBenchmarkDescriptionT<CKKS>::ElementWiseBenchmarkDescriptionT(Category category, Workload op)
{
    if (op != Workload::EltwiseAdd && op != Workload::EltwiseMultiply)
        throw HEBenchError(HEBERROR_MSG_CLASS("Invalid workload. Only EltwiseAdd and EltwiseMultiply are supported."), HEBENCH_ECODE_INVALID_ARGS);
    setup(CKKS, op, category, 0, AlgorithmName, AlgorithmDescription, { 1000 }, { "n" },
          vectorDefaults(CKKS, op == Workload::EltwiseAdd ? VectorOp::Add : VectorOp::Multiply));
}
template <bool CKKS>
hebench::cpp::BaseBenchmark *ElementWiseBenchmarkDescriptionT<CKKS>::createBenchmark(hebench::cpp::BaseEngine &engine,
                                                                                    const hebench::APIBridge::WorkloadParams *p_params)
{
    if (!p_params) throw HEBenchError(HEBERROR_MSG_CLASS("Invalid empty workload parameters. This workload requires flexible parameters."), HEBENCH_ECODE_CRITICAL_ERROR);
    return new VectorBenchmarkT<CKKS>(engine, m_descriptor, *p_params, encryptionParams(*p_params),
                                      m_descriptor.workload == Workload::EltwiseAdd ? VectorOp::Add : VectorOp::Multiply);
}
template <bool CKKS> DotProductBenchmarkDescriptionT<CKKS>::DotProductBenchmarkDescriptionT(Category category)
{
    setup(CKKS, Workload::DotProduct, category, 0, AlgorithmName, AlgorithmDescription, { 100 }, { "n" }, vectorDefaults(CKKS, VectorOp::Dot));
}
template <bool CKKS>
hebench::cpp::BaseBenchmark *DotProductBenchmarkDescriptionT<CKKS>::createBenchmark(hebench::cpp::BaseEngine &engine,
                                                                                   const hebench::APIBridge::WorkloadParams *p_params)
{
    if (!p_params) throw HEBenchError(HEBERROR_MSG_CLASS("Invalid empty workload parameters. This workload requires flexible parameters."), HEBENCH_ECODE_CRITICAL_ERROR);
    return new VectorBenchmarkT<CKKS>(engine, m_descriptor, *p_params, encryptionParams(*p_params), VectorOp::Dot);
}

// ------------------------------------------------------------------ benchmark
namespace {
typedef std::vector<std::vector<Plaintext>> EncodedParams;     // [param][sample]
typedef std::vector<std::vector<Ciphertext>> EncryptedParams;
typedef GridOperands LoadedParams;                              // the longer parameter split over the GPUs, the other replicated
}   // namespace

template <bool CKKS>
VectorBenchmarkT<CKKS>::VectorBenchmarkT(hebench::cpp::BaseEngine &engine, const hebench::APIBridge::BenchmarkDescriptor &bench_desc,
                                         const hebench::APIBridge::WorkloadParams &bench_params, const EncryptionParams &ep, VectorOp op)
    : hebench::cpp::BaseBenchmark(engine, bench_desc, bench_params), m_w_params(bench_params), m_op(op)
{
    if (m_w_params.n() == 0 || m_w_params.n() > (CKKS ? ep.poly_modulus_degree / 2 : ep.poly_modulus_degree))
        throw HEBenchError(HEBERROR_MSG_CLASS("Invalid workload parameters. This workload only supports vectors of size up to the number of slots."),
                           HEBENCH_ECODE_INVALID_ARGS);
    m_p_ctx_wrapper = makeContext(CKKS, ep);
}

template <bool CKKS> Handle VectorBenchmarkT<CKKS>::encode(const DataPackCollection *p_parameters)
{
    if (p_parameters->pack_count != 2)
        throw HEBenchError(HEBERROR_MSG_CLASS("Invalid number of parameters detected in parameter pack. Expected 2."), HEBENCH_ECODE_INVALID_ARGS);
    EncodedParams params(2);
    for (std::uint64_t pos = 0; pos < 2; ++pos) {
        const DataPack &pack = findDataPack(*p_parameters, pos);
        params[pos].resize(pack.buffer_count);
        for (std::uint64_t s = 0; s < pack.buffer_count; ++s) {
            const hebench::APIBridge::NativeDataBuffer &buf = pack.p_buffers[s];
            if (!buf.p) throw HEBenchError(HEBERROR_MSG_CLASS("Unexpected empty input buffer."), HEBENCH_ECODE_INVALID_ARGS);
            if (CKKS) {
                const double *v = reinterpret_cast<const double *>(buf.p);
                params[pos][s]  = m_p_ctx_wrapper->encodeVector(std::vector<double>(v, v + buf.size / sizeof(double)));
            } else {
                const std::int64_t *v = reinterpret_cast<const std::int64_t *>(buf.p);
                params[pos][s]        = m_p_ctx_wrapper->encodeVector(std::vector<std::int64_t>(v, v + buf.size / sizeof(std::int64_t)));
            }
        }
    }
    return this->getEngine().template createHandle<EncodedParams>(sizeof(EncodedParams), 0, std::move(params));
}

template <bool CKKS> void VectorBenchmarkT<CKKS>::decode(Handle encoded_data, DataPackCollection *p_native)
{
    const std::vector<Plaintext> &encoded = this->getEngine().template retrieveFromHandle<std::vector<Plaintext>>(encoded_data);
    if (p_native->pack_count == 0) return;
    DataPack &pack                = p_native->p_data_packs[findDataPackIndex(*p_native, 0)];
    const std::uint64_t n_results = std::min<std::uint64_t>(pack.buffer_count, encoded.size());
    const std::uint64_t n_values  = m_op == VectorOp::Dot ? 1 : m_w_params.n();
    for (std::uint64_t r = 0; r < n_results; ++r) {
        hebench::APIBridge::NativeDataBuffer &buf = pack.p_buffers[r];
        if (!buf.p) continue;
        if (CKKS) {
            std::vector<double> v = m_p_ctx_wrapper->decodeCKKS(encoded[r]);
            const std::uint64_t n = std::min<std::uint64_t>(n_values, buf.size / sizeof(double));
            for (std::uint64_t i = 0; i < n; ++i) reinterpret_cast<double *>(buf.p)[i] = flushTiny(v[i]);
        } else {
            std::vector<std::int64_t> v = m_p_ctx_wrapper->decodeBFV(encoded[r]);
            const std::uint64_t n       = std::min<std::uint64_t>(n_values, buf.size / sizeof(std::int64_t));
            for (std::uint64_t i = 0; i < n; ++i) reinterpret_cast<std::int64_t *>(buf.p)[i] = v[i];
        }
    }
}

template <bool CKKS> Handle VectorBenchmarkT<CKKS>::encrypt(Handle encoded_data)
{
    const EncodedParams &encoded = this->getEngine().template retrieveFromHandle<EncodedParams>(encoded_data);
    EncryptedParams encrypted(encoded.size());
    for (std::size_t p = 0; p < encoded.size(); ++p) encrypted[p] = m_p_ctx_wrapper->encrypt(encoded[p]);
    return this->getEngine().template createHandle<EncryptedParams>(sizeof(EncryptedParams), 0, std::move(encrypted));
}

template <bool CKKS> Handle VectorBenchmarkT<CKKS>::decrypt(Handle encrypted_data)
{
    const std::vector<Ciphertext> &encrypted = this->getEngine().template retrieveFromHandle<std::vector<Ciphertext>>(encrypted_data);
    std::vector<Plaintext> plain             = m_p_ctx_wrapper->decrypt(encrypted);
    return this->getEngine().template createHandle<std::vector<Plaintext>>(sizeof(plain), 0, std::move(plain));
}

// load = host -> HBM.  The RESULT grid b0 x b1 is what gets partitioned: the parameter with more samples is
// split into one contiguous block per GPU, the other goes to every GPU, so a sample crosses PCIe once unless every GPU
// needs it (R/src/benchmarks/ckks/seal_ckks_dot_product_benchmark.cpp:315-318 is the loop being partitioned).
template <bool CKKS> Handle VectorBenchmarkT<CKKS>::load(const Handle *p_local_data, std::uint64_t count)
{
    if (count != 1) throw HEBenchError(HEBERROR_MSG_CLASS("Invalid number of handles. Expected 1."), HEBENCH_ECODE_INVALID_ARGS);
    const EncryptedParams &enc = this->getEngine().template retrieveFromHandle<EncryptedParams>(p_local_data[0]);
    if (enc.size() != 2) throw HEBenchError(HEBERROR_MSG_CLASS("Expected 2 operation parameters."), HEBENCH_ECODE_INVALID_ARGS);
    for (int p = 0; p < 2; ++p) m_p_ctx_wrapper->trace(p ? "in1" : "in0", enc[p]);
    LoadedParams loaded = m_p_ctx_wrapper->loadGrid(enc[0], 1, enc[1], 1);
    return this->getEngine().template createHandle<LoadedParams>(sizeof(LoadedParams), 0, std::move(loaded));
}

// store = HBM -> host; unused output handles are zeroed (R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:290-300)
template <bool CKKS> void VectorBenchmarkT<CKKS>::store(Handle remote_data, Handle *p_local_data, std::uint64_t count)
{
    if (count > 0) {
        std::memset(p_local_data, 0, sizeof(Handle) * count);
        const ShardedCiphertexts &res = this->getEngine().template retrieveFromHandle<ShardedCiphertexts>(remote_data);
        std::vector<Ciphertext> host  = m_p_ctx_wrapper->gather(res);
        m_p_ctx_wrapper->trace("out", host);
        p_local_data[0]               = this->getEngine().template createHandle<std::vector<Ciphertext>>(sizeof(host), 0, std::move(host));
    }
}

template <bool CKKS> Handle VectorBenchmarkT<CKKS>::operate(Handle h_remote_packed, const ParameterIndexer *p_param_indexers, std::uint64_t indexers_count)
{
    if (indexers_count < 2) throw HEBenchError(HEBERROR_MSG_CLASS("Invalid number of indexers. Expected 2."), HEBENCH_ECODE_INVALID_ARGS);
    const LoadedParams &in = this->getEngine().template retrieveFromHandle<LoadedParams>(h_remote_packed);
    std::uint64_t b[2], v0[2];
    for (int p = 0; p < 2; ++p) {
        v0[p] = p_param_indexers[p].value_index;
        b[p]  = p_param_indexers[p].batch_size;
        if (v0[p] + b[p] > in.p[p].total()) {
            std::stringstream ss;
            ss << "Invalid parameter indexer for operation parameter " << p << ". Expected index in range [0, " << in.p[p].total()
               << "), but " << v0[p] << " + " << b[p] << " received.";
            throw HEBenchError(HEBERROR_MSG_CLASS(ss.str()), HEBENCH_ECODE_INVALID_ARGS);
        }
    }
    SEALContextWrapper &cw = *m_p_ctx_wrapper;
    cw.beginOperate();
    ShardedCiphertexts out;
    out.n_total = b[0] * b[1];
    out.shard.resize(cw.gpuCount());
    out.ids.resize(cw.gpuCount());
    out.first.assign(cw.gpuCount() + 1, 0);
    cw.forEachGpu([&](int g) {
        // this GPU's share of the b0 x b1 result grid (r = i*b1 + j): the samples of the split parameter it holds
        GridShare sh     = cw.gridShare(in, g, v0, b);
        const std::uint64_t n = sh.result.size();
        b200he_ctx *c    = cw.device(g);
        DeviceBatchPtr r = cw.newBatch(g);
        b200he_batch *A = in.p[0].shard[g]->get(), *B = in.p[1].shard[g]->get();
        if (m_op == VectorOp::Add)
            cw.check(b200he_add(c, A, sh.ai.data(), B, sh.bi.data(), n, r->get()), "b200he_add");
        else {
            cw.check(b200he_multiply(c, A, sh.ai.data(), B, sh.bi.data(), n, r->get()), "b200he_multiply");
            if (m_op == VectorOp::Dot && n > 0) {
                cw.check(b200he_relinearize(c, r->get(), r->get()), "b200he_relinearize");
                if (CKKS) cw.accumulateCKKS(*r, m_w_params.n());
                else cw.accumulateBFV(*r, m_w_params.n());
            }
        }
        out.shard[g] = r;
        out.ids[g]   = std::move(sh.result);
    });
    cw.prepareStore(out);         // host memory for store(), touched while the GPUs work
    cw.endOperate(out.n_total);   // waits for every GPU: the harness times this call by wall clock
    return this->getEngine().template createHandle<ShardedCiphertexts>(sizeof(ShardedCiphertexts), 0, std::move(out));
}

template class ElementWiseBenchmarkDescriptionT<true>;
template class ElementWiseBenchmarkDescriptionT<false>;
template class DotProductBenchmarkDescriptionT<true>;
template class DotProductBenchmarkDescriptionT<false>;
template class VectorBenchmarkT<true>;
template class VectorBenchmarkT<false>;

}   // namespace sbe
