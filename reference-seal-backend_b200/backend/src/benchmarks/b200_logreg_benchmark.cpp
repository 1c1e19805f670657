// b200_logreg_benchmark.cpp -- CKKS logistic-regression inference with a degree-3 (reference) / 5 / 7 sigmoid (Horner).
// Replaces R/src/benchmarks/ckks/seal_ckks_logreg_horner.cpp.  operate() keeps the reference's op sequence:
//   per sample:  multiply(W, X_i) -> relinearize -> accumulateCKKS(n) -> rescale        (…logreg_horner.cpp:425-444)
//   collapse:    rotate sample i by -i, mask, rescale, sum                               (R/src/engine/seal_context.cpp:349-415)
//   + bias, Horner evaluation of the sigmoid polynomial                                  (…logreg_horner.cpp:461-476)
// The samples are the independent units: they are sharded over the GPUs, each GPU collapses its shard to one
// partial ciphertext, the partials are summed on GPU 0 (the path's only exchange step).
#include <cstring>
#include <sstream>

#include "benchmarks/b200_benchmarks.h"

namespace sbe {
namespace ckks {

using hebench::APIBridge::Category;
using hebench::APIBridge::DataPack;
using hebench::APIBridge::DataPackCollection;
using hebench::APIBridge::Handle;
using hebench::APIBridge::ParameterIndexer;
using hebench::APIBridge::Workload;
using hebench::cpp::HEBenchError;

constexpr double LogRegHornerBenchmarkDescription::SigmoidPolyCoeff[];
constexpr double LogRegHornerBenchmarkDescription::SigmoidPolyCoeffD5[];
constexpr double LogRegHornerBenchmarkDescription::SigmoidPolyCoeffD7[];

std::size_t LogRegHornerBenchmarkDescription::polynomialDegree(Workload w)
{
    switch (w) {
    case Workload::LogisticRegression_PolyD3: return 3;
    case Workload::LogisticRegression_PolyD5: return 5;
    case Workload::LogisticRegression_PolyD7: return 7;
    default: return 0;
    }
}
std::vector<double> LogRegHornerBenchmarkDescription::sigmoidCoefficients(Workload w)
{
    switch (polynomialDegree(w)) {
    case 3: return std::vector<double>(SigmoidPolyCoeff, SigmoidPolyCoeff + 4);
    case 5: return std::vector<double>(SigmoidPolyCoeffD5, SigmoidPolyCoeffD5 + 6);
    case 7: return std::vector<double>(SigmoidPolyCoeffD7, SigmoidPolyCoeffD7 + 8);
    default: return {};
    }
}

LogRegHornerBenchmarkDescription::LogRegHornerBenchmarkDescription(Category category, std::size_t batch_size, Workload workload)
{
    // R/include/benchmarks/ckks/seal_ckks_logreg_horner.h:57-61: N = 16384, {60, 45 x 5, 60} for degree 3.
    // One level for W.X, one for the collapse mask, one per Horner step, one to spare: 6 / 8 / 10; the deeper chains
    // ({60, 45 x 9, 60} = 525 bits for degree 7) need N = 32768 at 128-bit security.
    const std::size_t degree = polynomialDegree(workload);
    setup(true, workload, category, 1 /* LogRegOtherID */, AlgorithmName, AlgorithmDescription, { 16 }, { "n" },
          EncryptionParams{ degree > 3 ? 32768u : 16384u, (std::uint64_t)(degree + 3), 45, 45, 0 });
    if (category == Category::Offline) {   // W and b: one sample each; X: the batch (0 = chosen by the harness)
        m_descriptor.cat_params.offline.data_count[Index_W] = 1;
        m_descriptor.cat_params.offline.data_count[Index_b] = 1;
        m_descriptor.cat_params.offline.data_count[Index_X] = batch_size;
    }
}

hebench::cpp::BaseBenchmark *LogRegHornerBenchmarkDescription::createBenchmark(hebench::cpp::BaseEngine &engine,
                                                                               const hebench::APIBridge::WorkloadParams *p_params)
{
    if (!p_params) throw HEBenchError(HEBERROR_MSG_CLASS("Invalid empty workload parameters. This workload requires flexible parameters."), HEBENCH_ECODE_CRITICAL_ERROR);
    return new LogRegHornerBenchmark(engine, m_descriptor, *p_params, encryptionParams(*p_params));
}

LogRegHornerBenchmark::LogRegHornerBenchmark(hebench::cpp::BaseEngine &engine, const hebench::APIBridge::BenchmarkDescriptor &bench_desc,
                                             const hebench::APIBridge::WorkloadParams &bench_params, const EncryptionParams &ep)
    : hebench::cpp::BaseBenchmark(engine, bench_desc, bench_params), m_w_params(bench_params)
{
    const hebench::APIBridge::BenchmarkDescriptor &d = getDescriptor();
    const std::size_t degree = LogRegHornerBenchmarkDescription::polynomialDegree(d.workload);
    if (degree == 0 || d.data_type != hebench::APIBridge::DataType::Float64
        || (d.cipher_param_mask & 0x03) != 0x03 || d.scheme != HEBENCH_HE_SCHEME_CKKS || d.security != HEBENCH_HE_SECURITY_128)
        throw HEBenchError(HEBERROR_MSG_CLASS("Benchmark descriptor received is not supported."), HEBENCH_ECODE_INVALID_ARGS);
    if (d.category == Category::Offline && (d.cat_params.offline.data_count[0] > 1 || d.cat_params.offline.data_count[1] > 1))
        throw HEBenchError(HEBERROR_MSG_CLASS("Benchmark descriptor received is not supported."), HEBENCH_ECODE_INVALID_ARGS);
    if (ep.coeff_modulus_bits < 1) throw HEBenchError(HEBERROR_MSG_CLASS("Multiplicative depth must be greater than 0."), HEBENCH_ECODE_INVALID_ARGS);
    // one level for W.X, one for the collapse mask, one per Horner step (degree 3: 6 as in the reference, :110)
    if (ep.multiplicative_depth < degree + 3)
        throw HEBenchError(HEBERROR_MSG_CLASS("Multiplicative depth must be at least " + std::to_string(degree + 3) + " for this workload."), HEBENCH_ECODE_INVALID_ARGS);
    m_p_ctx_wrapper = makeContext(true, ep);
    if (m_w_params.n() > m_p_ctx_wrapper->slotCount())
        throw HEBenchError(HEBERROR_MSG_CLASS("Invalid workload parameter 'n'. Number of features must be under " + std::to_string(m_p_ctx_wrapper->slotCount()) + "."),
                           HEBENCH_ECODE_INVALID_ARGS);
    // sigmoid coefficients, each broadcast to every slot
    for (double coeff : LogRegHornerBenchmarkDescription::sigmoidCoefficients(d.workload))
        m_plain_coeff.push_back(m_p_ctx_wrapper->encodeVector(std::vector<double>(m_p_ctx_wrapper->slotCount(), coeff)));
}

LogRegHornerBenchmark::~LogRegHornerBenchmark()
{
    if (m_fresh.valid()) m_fresh.wait();
}

void LogRegHornerBenchmark::drawAhead()
{
    SEALContextWrapper *cw                  = m_p_ctx_wrapper.get();
    const std::vector<Plaintext> *coeff     = &m_plain_coeff;
    m_fresh = std::async(std::launch::async, [cw, coeff]() {
        FreshEncryptions f;
        f.zero = cw->encrypt(cw->encodeVector(std::vector<double>(1, 0.0), cw->scale()));   // encrypt_zero, scale() forced (:360-361)
        f.seed = cw->encrypt(coeff->back());                                                 // Enc(a_d) (:440)
        return f;
    });
}

LogRegHornerBenchmark::FreshEncryptions LogRegHornerBenchmark::takeFresh()
{
    if (!m_fresh.valid()) drawAhead();   // operate() without a preceding load() of this object
    FreshEncryptions f = m_fresh.get();
    drawAhead();                         // the next call's pair is drawn while the GPUs run this one
    return f;
}

Handle LogRegHornerBenchmark::encode(const DataPackCollection *p_parameters)
{
    if (p_parameters->pack_count != LogRegHornerBenchmarkDescription::NumOpParams)
        throw HEBenchError(HEBERROR_MSG_CLASS("Invalid number of operation parameters detected in parameter pack. Expected 3."), HEBENCH_ECODE_INVALID_ARGS);
    SEALContextWrapper &cw = *m_p_ctx_wrapper;
    const DataPack &pw = findDataPack(*p_parameters, LogRegHornerBenchmarkDescription::Index_W);
    const DataPack &pb = findDataPack(*p_parameters, LogRegHornerBenchmarkDescription::Index_b);
    const DataPack &px = findDataPack(*p_parameters, LogRegHornerBenchmarkDescription::Index_X);
    if (pw.buffer_count < 1 || !pw.p_buffers || !pw.p_buffers[0].p) throw HEBenchError(HEBERROR_MSG_CLASS("Unexpected empty DataPack for 'W'."), HEBENCH_ECODE_INVALID_ARGS);
    if (pw.p_buffers[0].size / sizeof(double) < m_w_params.n()) {
        std::stringstream ss;
        ss << "Insufficient features for 'W'. Expected " << m_w_params.n() << ", but " << pw.p_buffers[0].size / sizeof(double) << " received.";
        throw HEBenchError(HEBERROR_MSG_CLASS(ss.str()), HEBENCH_ECODE_INVALID_ARGS);
    }
    if (pb.buffer_count < 1 || !pb.p_buffers || !pb.p_buffers[0].p || pb.p_buffers[0].size < sizeof(double))
        throw HEBenchError(HEBERROR_MSG_CLASS("Unexpected empty DataPack for 'b'."), HEBENCH_ECODE_INVALID_ARGS);
    const std::uint64_t batch = getDescriptor().category == Category::Offline && getDescriptor().cat_params.offline.data_count[LogRegHornerBenchmarkDescription::Index_X] > 0
                                    ? getDescriptor().cat_params.offline.data_count[LogRegHornerBenchmarkDescription::Index_X]
                                    : (getDescriptor().category == Category::Offline ? px.buffer_count : 1);
    if (!px.p_buffers || px.buffer_count < batch) {
        std::stringstream ss;
        ss << "Unexpected batch size for inputs. Expected, at least, " << batch << ", but " << px.buffer_count << " received.";
        throw HEBenchError(HEBERROR_MSG_CLASS(ss.str()), HEBENCH_ECODE_INVALID_ARGS);
    }
    if (batch > cw.slotCount()) throw HEBenchError(HEBERROR_MSG_CLASS("Batch size exceeds the number of slots of the result ciphertext."), HEBENCH_ECODE_INVALID_ARGS);
    const double *w = reinterpret_cast<const double *>(pw.p_buffers[0].p);
    EncodedOpParams enc;
    std::get<0>(enc) = cw.encodeVector(std::vector<double>(w, w + m_w_params.n()));
    std::get<1>(enc) = cw.encodeVector(std::vector<double>(cw.slotCount(), *reinterpret_cast<const double *>(pb.p_buffers[0].p)));
    for (std::uint64_t i = 0; i < batch; ++i) {
        if (!px.p_buffers[i].p || px.p_buffers[i].size / sizeof(double) < m_w_params.n())
            throw HEBenchError(HEBERROR_MSG_CLASS("Invalid input sample " + std::to_string(i) + "."), HEBENCH_ECODE_INVALID_ARGS);
        const double *x = reinterpret_cast<const double *>(px.p_buffers[i].p);
        std::get<2>(enc).push_back(cw.encodeVector(std::vector<double>(x, x + m_w_params.n())));
    }
    return this->getEngine().createHandle<EncodedOpParams>(sizeof(EncodedOpParams), EncodedOpParamsTag, std::move(enc));
}

Handle LogRegHornerBenchmark::encrypt(Handle encoded_data)
{
    const EncodedOpParams &enc = this->getEngine().retrieveFromHandle<EncodedOpParams>(encoded_data, EncodedOpParamsTag);
    EncryptedOpParams out;
    std::get<0>(out) = m_p_ctx_wrapper->encrypt(std::get<0>(enc));
    std::get<1>(out) = m_p_ctx_wrapper->encrypt(std::get<1>(enc));
    std::get<2>(out) = m_p_ctx_wrapper->encrypt(std::get<2>(enc));
    return this->getEngine().createHandle<EncryptedOpParams>(sizeof(EncryptedOpParams), EncryptedOpParamsTag, std::move(out));
}

Handle LogRegHornerBenchmark::load(const Handle *p_local_data, std::uint64_t count)
{
    if (count != 1) throw HEBenchError(HEBERROR_MSG_CLASS("Invalid number of handles. Expected 1."), HEBENCH_ECODE_INVALID_ARGS);
    const EncryptedOpParams &enc = this->getEngine().retrieveFromHandle<EncryptedOpParams>(p_local_data[0], EncryptedOpParamsTag);
    SEALContextWrapper &cw       = *m_p_ctx_wrapper;
    LoadedOpParams loaded;
    const std::vector<Ciphertext> &X = std::get<2>(enc);
    cw.trace("W", std::get<0>(enc));
    cw.trace("b", std::get<1>(enc));
    cw.trace("X", X);
    loaded.X.n_total = X.size();
    loaded.X.first   = cw.partition(X.size());
    loaded.X.shard.resize(cw.gpuCount());
    loaded.W.resize(cw.gpuCount());
    loaded.b.resize(cw.gpuCount());
    const int top = (int)cw.topLevel();
    cw.forEachGpu([&](int g) {
        const std::uint64_t first = loaded.X.first[g], n = loaded.X.first[g + 1] - first;
        loaded.W[g]       = cw.upload(g, std::get<0>(enc));
        loaded.b[g]       = cw.upload(g, std::get<1>(enc));
        loaded.X.shard[g] = cw.upload(g, X, first, n);
        // everything deterministic that operate() needs is staged here, outside the timed call: the collapse masks of this
        // GPU's samples (level of the rescaled dot products) and, on GPU 0, the sigmoid coefficients at the levels of the
        // Horner steps (R/src/engine/seal_context.cpp:382-388,451)
        if (n > 0) cw.maskBatch(g, first, n, X.size(), top - 1);
        if (g == 0)
            for (std::size_t k = 0; k + 1 < m_plain_coeff.size(); ++k) {
                const int level = top - 3 - (int)(m_plain_coeff.size() - 2 - k);   // a_{d-1} is added at level top - 3
                if (level >= 1) cw.coeffBatch(0, m_plain_coeff, k, level);
            }
    });
    drawAhead();   // the fresh encryptions of the first operate() call
    return this->getEngine().createHandle<LoadedOpParams>(sizeof(LoadedOpParams), EncryptedOpParamsTag, std::move(loaded));
}

void LogRegHornerBenchmark::store(Handle remote_data, Handle *p_local_data, std::uint64_t count)
{
    if (count > 0) {
        std::memset(p_local_data, 0, sizeof(Handle) * count);
        const DeviceBatchPtr &res = this->getEngine().retrieveFromHandle<DeviceBatchPtr>(remote_data, EncryptedResultTag);
        std::vector<Ciphertext> h = m_p_ctx_wrapper->download(*res);
        m_p_ctx_wrapper->trace("out", h);
        p_local_data[0]           = this->getEngine().createHandle<Ciphertext>(sizeof(Ciphertext), EncryptedResultTag, std::move(h.at(0)));
    }
}

Handle LogRegHornerBenchmark::decrypt(Handle encrypted_data)
{
    const Ciphertext &c = this->getEngine().retrieveFromHandle<Ciphertext>(encrypted_data, EncryptedResultTag);
    Plaintext p         = m_p_ctx_wrapper->decrypt(c);
    return this->getEngine().createHandle<Plaintext>(sizeof(Plaintext), EncodedResultTag, std::move(p));
}

// one double per sample: buffer i <- slot i of the single result (…logreg_horner.cpp:295-329)
void LogRegHornerBenchmark::decode(Handle encoded_data, DataPackCollection *p_native)
{
    if (p_native->pack_count == 0) return;
    const Plaintext &p   = this->getEngine().retrieveFromHandle<Plaintext>(encoded_data, EncodedResultTag);
    std::vector<double> v = m_p_ctx_wrapper->decodeCKKS(p);
    DataPack &pack        = p_native->p_data_packs[findDataPackIndex(*p_native, 0)];
    for (std::uint64_t i = 0; i < pack.buffer_count && i < v.size(); ++i)
        if (pack.p_buffers[i].p && pack.p_buffers[i].size >= sizeof(double)) *reinterpret_cast<double *>(pack.p_buffers[i].p) = flushTiny(v[i]);
}

Handle LogRegHornerBenchmark::operate(Handle h_remote_packed, const ParameterIndexer *p_param_indexers, std::uint64_t indexers_count)
{
    if (indexers_count < LogRegHornerBenchmarkDescription::NumOpParams) {
        std::stringstream ss;
        ss << "Invalid number of indexers. Expected " << LogRegHornerBenchmarkDescription::NumOpParams << ", but " << indexers_count << " received.";
        throw HEBenchError(HEBERROR_MSG_CLASS(ss.str()), HEBENCH_ECODE_INVALID_ARGS);
    }
    const LoadedOpParams &in = this->getEngine().retrieveFromHandle<LoadedOpParams>(h_remote_packed, EncryptedOpParamsTag);
    const std::uint64_t batch = in.X.total();
    const ParameterIndexer &ix = p_param_indexers[LogRegHornerBenchmarkDescription::Index_X];
    if (ix.value_index != 0 || (getDescriptor().category == Category::Offline && ix.batch_size != batch)
        || (getDescriptor().category == Category::Latency && ix.batch_size != 1))
        throw HEBenchError(HEBERROR_MSG_CLASS("Invalid indexer range for parameter " + std::to_string(LogRegHornerBenchmarkDescription::Index_X) + " detected."),
                           HEBENCH_ECODE_INVALID_ARGS);
    SEALContextWrapper &cw = *m_p_ctx_wrapper;
    const FreshEncryptions fresh = takeFresh();
    cw.beginOperate();
    // linear part + collapse, per GPU on its shard of the samples, one issuing host thread per GPU.  Everything only
    // enqueues work on the GPUs' streams (staged uploads of the two fresh encryptions included).
    std::vector<DeviceBatchPtr> partial(cw.gpuCount());
    cw.forEachGpu([&](int g) {
        const std::uint64_t first = in.X.first[g], n = in.X.first[g + 1] - first;
        b200he_ctx *c       = cw.device(g);
        DeviceBatchPtr dots = cw.newBatch(g);
        if (n > 0) {
            std::vector<uint32_t> wi(n, 0);
            cw.check(b200he_multiply(c, in.W[g]->get(), wi.data(), in.X.shard[g]->get(), nullptr, n, dots->get()), "b200he_multiply");
            cw.check(b200he_relinearize(c, dots->get(), dots->get()), "b200he_relinearize");
            cw.accumulateCKKS(*dots, m_w_params.n());
            cw.check(b200he_rescale_to_next(c, dots->get(), dots->get()), "b200he_rescale_to_next");
        }
        if (n > 0 || g == 0) partial[g] = cw.collapseCKKS(*dots, first, batch, g == 0 ? &fresh.zero : nullptr);
    });
    // the path's one exchange step: the other GPUs' partial sums travel device to device (NVLink) and are added on GPU 0
    DeviceBatchPtr lr = partial[0];
    for (int g = 1; g < cw.gpuCount(); ++g) {
        if (!partial[g]) continue;
        DeviceBatchPtr p0 = cw.newBatch(0);
        cw.check(b200he_batch_copy_from(p0->get(), partial[g]->get()), "b200he_batch_copy_from");
        cw.check(b200he_add(cw.device(0), lr->get(), nullptr, p0->get(), nullptr, 1, lr->get()), "b200he_add");
    }
    // bias: level-matched copy, scales forced (…logreg_horner.cpp:461-465)
    b200he_ctx *c0  = cw.device(0);
    DeviceBatchPtr b = cw.newBatch(0);
    cw.check(b200he_gather(c0, in.b[0]->get(), nullptr, 1, b->get()), "b200he_gather");
    cw.matchLevel(*b, *lr);
    cw.check(b200he_batch_set_scale(b->get(), cw.scale()), "b200he_batch_set_scale");
    cw.check(b200he_batch_set_scale(lr->get(), cw.scale()), "b200he_batch_set_scale");
    cw.check(b200he_add(c0, lr->get(), nullptr, b->get(), nullptr, 1, lr->get()), "b200he_add");
    // sigmoid
    DeviceBatchPtr result = cw.evaluatePolynomial(*lr, m_plain_coeff, fresh.seed);
    cw.endOperate(batch);
    return this->getEngine().createHandle<DeviceBatchPtr>(sizeof(DeviceBatchPtr), EncryptedResultTag, std::move(result));
}

}   // namespace ckks
}   // namespace sbe
