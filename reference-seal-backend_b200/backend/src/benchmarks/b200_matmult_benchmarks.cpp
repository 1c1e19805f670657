// b200_matmult_benchmarks.cpp -- the three matrix-multiplication algorithms of the reference (latency category):
//   Val              R/src/benchmarks/{ckks,bfv}/seal_*_matmultval_benchmark.cpp         (descriptor.other = 0)
//   CipherBatchAxis  R/src/benchmarks/{ckks,bfv}/seal_*_matmult_cipherbatchaxis_benchmark.cpp   (other = 1)
//   Row              R/src/benchmarks/{ckks,bfv}/seal_*_matmult_row_benchmark.cpp        (other = 2)
// Packing, op sequences and decode conventions follow the reference; the output cells / rows are the
// independent units that are sharded across GPUs and batched into single kernel launches.
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

namespace {
typedef std::array<std::vector<Plaintext>, 2> EncodedMats;
typedef std::array<std::vector<Ciphertext>, 2> EncryptedMats;
typedef GridOperands LoadedMats;   // result cells / rows partitioned: row block of M0 + all of M1, or all of M0 + column block of M1
constexpr std::int64_t ResultCipherTag = 0x10, ResultPlainTag = 0x20;   // R/src/benchmarks/ckks/seal_ckks_matmultval_benchmark.cpp:401-403,463-469

const char *algoName(MatMultAlgo a) { return a == MatMultAlgo::Val ? "MatMultVal" : a == MatMultAlgo::Row ? "MatMulRow" : "CipherBatchAxis"; }
const char *algoDesc(MatMultAlgo a)
{
    return a == MatMultAlgo::Val ? "One matrix row per ciphertext"
           : a == MatMultAlgo::Row ? "Row-replicated M0, M1 packed in one ciphertext"
                                   : "One matrix element per ciphertext";
}
EncryptionParams matDefaults(bool ckks, MatMultAlgo a)
{
    // R/include/benchmarks/{ckks,bfv}/seal_*_matmult*_benchmark.h: depth 2 for Val, 3 for Row / CipherBatchAxis
    const std::uint64_t depth = a == MatMultAlgo::Val ? 2 : 3;
    return ckks ? EncryptionParams{ 8192, depth, 45, 45, 0 } : EncryptionParams{ 8192, depth, 40, 20, 0 };
}
}   // namespace

template <bool CKKS> MatMultBenchmarkDescriptionT<CKKS>::MatMultBenchmarkDescriptionT(MatMultAlgo algo) : m_algo(algo)
{
    setup(CKKS, Workload::MatrixMultiply, Category::Latency, (std::int64_t)algo, algoName(algo), algoDesc(algo), { 10, 9, 8 },
          { "rows_M0", "cols_M0", "cols_M1" }, matDefaults(CKKS, algo));
}
template <bool CKKS>
hebench::cpp::BaseBenchmark *MatMultBenchmarkDescriptionT<CKKS>::createBenchmark(hebench::cpp::BaseEngine &engine,
                                                                                const hebench::APIBridge::WorkloadParams *p_params)
{
    if (!p_params) throw HEBenchError(HEBERROR_MSG_CLASS("Invalid empty workload parameters. Matrix Multiplication requires parameters."), HEBENCH_ECODE_CRITICAL_ERROR);
    return new MatMultBenchmarkT<CKKS>(engine, m_descriptor, *p_params, encryptionParams(*p_params), m_algo);
}

template <bool CKKS>
MatMultBenchmarkT<CKKS>::MatMultBenchmarkT(hebench::cpp::BaseEngine &engine, const hebench::APIBridge::BenchmarkDescriptor &bench_desc,
                                           const hebench::APIBridge::WorkloadParams &bench_params, const EncryptionParams &ep, MatMultAlgo algo)
    : hebench::cpp::BaseBenchmark(engine, bench_desc, bench_params), m_w_params(bench_params), m_algo(algo)
{
    const std::uint64_t r0 = m_w_params.rows_M0(), c0 = m_w_params.cols_M0(), c1 = m_w_params.cols_M1();
    if (r0 == 0 || c0 == 0 || c1 == 0) throw HEBenchError(HEBERROR_MSG_CLASS("Matrix dimensions must be greater than 0."), HEBENCH_ECODE_INVALID_ARGS);
    const std::uint64_t slots = CKKS ? ep.poly_modulus_degree / 2 : ep.poly_modulus_degree;
    if (algo == MatMultAlgo::Val && c0 > slots)
        throw HEBenchError(HEBERROR_MSG_CLASS("Invalid workload parameters. Number of columns of M0 exceeds the slots of a ciphertext."), HEBENCH_ECODE_INVALID_ARGS);
    // Row: cols_M0 * cols_M1 must fit one rotation row (R/src/benchmarks/ckks/seal_ckks_matmult_row_benchmark.cpp:142)
    if (algo == MatMultAlgo::Row && c0 * c1 > ep.poly_modulus_degree / 2)
        throw HEBenchError(HEBERROR_MSG_CLASS("Invalid workload parameters. cols_M0 * cols_M1 exceeds the available slots."), HEBENCH_ECODE_INVALID_ARGS);
    m_p_ctx_wrapper = makeContext(CKKS, ep);
}

// ------------------------------------------------------------------ packing
template <bool CKKS> std::vector<Plaintext> MatMultBenchmarkT<CKKS>::encodeM0(const Scalar *m) const
{
    const std::size_t r0 = m_w_params.rows_M0(), c0 = m_w_params.cols_M0(), c1 = m_w_params.cols_M1();
    SEALContextWrapper &cw = *m_p_ctx_wrapper;
    std::vector<Plaintext> out;
    if (m_algo == MatMultAlgo::Val) {   // one row per plaintext
        for (std::size_t i = 0; i < r0; ++i) out.push_back(cw.encodeVector(std::vector<Scalar>(m + i * c0, m + (i + 1) * c0)));
    } else if (m_algo == MatMultAlgo::CipherBatchAxis) {   // one element per plaintext, broadcast to every slot; row-major
        out.resize(r0 * c0);
#pragma omp parallel for schedule(dynamic, 4)
        for (long i = 0; i < (long)(r0 * c0); ++i) out[i] = cw.encodeVector(std::vector<Scalar>(cw.slotCount(), m[i]));
    } else if (CKKS) {   // Row: slot[spacers*j + k] = M0[i][j] for k < cols_M1 (…ckks_matmult_row…:234-244)
        const std::size_t spacers = cw.slotCount() / c0;
        for (std::size_t i = 0; i < r0; ++i) {
            std::vector<Scalar> v(cw.slotCount(), 0);
            for (std::size_t j = 0; j < c0; ++j)
                for (std::size_t k = 0; k < c1; ++k) v[spacers * j + k] = m[i * c0 + j];
            out.push_back(cw.encodeVector(v));
        }
    } else {   // BFV Row: two matrix rows per plaintext, one per batching row (…bfv_matmult_row…:226-255)
        const std::size_t row_size = cw.slotCount() / 2, spacers = row_size / c0;
        for (std::size_t i = 0; i < r0; i += 2) {
            std::vector<Scalar> v(cw.slotCount(), 0);
            for (std::size_t j = 0; j < c0; ++j)
                for (std::size_t k = 0; k < c1; ++k) {
                    v[spacers * j + k] = m[i * c0 + j];
                    if (i + 1 < r0) v[row_size + spacers * j + k] = m[(i + 1) * c0 + j];
                }
            out.push_back(cw.encodeVector(v));
        }
    }
    return out;
}

template <bool CKKS> std::vector<Plaintext> MatMultBenchmarkT<CKKS>::encodeM1(const Scalar *m) const
{
    const std::size_t c0 = m_w_params.cols_M0(), c1 = m_w_params.cols_M1();
    SEALContextWrapper &cw = *m_p_ctx_wrapper;
    std::vector<Plaintext> out;
    if (m_algo == MatMultAlgo::Val) {   // transposed: one COLUMN of M1 per plaintext (…matmultval…:213-226)
        for (std::size_t j = 0; j < c1; ++j) {
            std::vector<Scalar> col(c0);
            for (std::size_t k = 0; k < c0; ++k) col[k] = m[k * c1 + j];
            out.push_back(cw.encodeVector(col));
        }
    } else if (m_algo == MatMultAlgo::CipherBatchAxis) {
        out.resize(c0 * c1);
#pragma omp parallel for schedule(dynamic, 4)
        for (long i = 0; i < (long)(c0 * c1); ++i) out[i] = cw.encodeVector(std::vector<Scalar>(cw.slotCount(), m[i]));
    } else {   // Row: slot[spacers*j + k] = M1[j][k]; BFV repeats it in the second batching row
        const std::size_t row_size = CKKS ? cw.slotCount() : cw.slotCount() / 2, spacers = row_size / c0;
        std::vector<Scalar> v(cw.slotCount(), 0);
        for (std::size_t j = 0; j < c0; ++j)
            for (std::size_t k = 0; k < c1; ++k) {
                v[spacers * j + k] = m[j * c1 + k];
                if (!CKKS) v[row_size + spacers * j + k] = m[j * c1 + k];
            }
        out.push_back(cw.encodeVector(v));
    }
    return out;
}

template <bool CKKS> Handle MatMultBenchmarkT<CKKS>::encode(const DataPackCollection *p_parameters)
{
    if (p_parameters->pack_count != 2)
        throw HEBenchError(HEBERROR_MSG_CLASS("Invalid number of parameters detected in parameter pack. Expected 2."), HEBENCH_ECODE_INVALID_ARGS);
    const std::uint64_t want[2] = { m_w_params.rows_M0() * m_w_params.cols_M0(), m_w_params.cols_M0() * m_w_params.cols_M1() };
    const Scalar *mat[2];
    for (std::uint64_t pos = 0; pos < 2; ++pos) {
        const DataPack &pack = findDataPack(*p_parameters, pos);
        if (pack.buffer_count < 1 || !pack.p_buffers[0].p || pack.p_buffers[0].size < want[pos] * sizeof(Scalar))
            throw HEBenchError(HEBERROR_MSG_CLASS("Unexpected empty or undersized buffer for matrix " + std::to_string(pos) + "."), HEBENCH_ECODE_INVALID_ARGS);
        mat[pos] = reinterpret_cast<const Scalar *>(pack.p_buffers[0].p);
    }
    EncodedMats enc = { encodeM0(mat[0]), encodeM1(mat[1]) };
    return this->getEngine().template createHandle<EncodedMats>(sizeof(EncodedMats), 0, std::move(enc));
}

template <bool CKKS> Handle MatMultBenchmarkT<CKKS>::encrypt(Handle encoded_data)
{
    const EncodedMats &enc = this->getEngine().template retrieveFromHandle<EncodedMats>(encoded_data);
    EncryptedMats out      = { m_p_ctx_wrapper->encrypt(enc[0]), m_p_ctx_wrapper->encrypt(enc[1]) };
    return this->getEngine().template createHandle<EncryptedMats>(sizeof(EncryptedMats), 0, std::move(out));
}

template <bool CKKS> Handle MatMultBenchmarkT<CKKS>::load(const Handle *p_local_data, std::uint64_t count)
{
    if (count != 1) throw HEBenchError(HEBERROR_MSG_CLASS("Expected only 1 local handle to load."), HEBENCH_ECODE_INVALID_ARGS);
    const EncryptedMats &enc = this->getEngine().template retrieveFromHandle<EncryptedMats>(p_local_data[0]);
    m_p_ctx_wrapper->trace("in0", enc[0]);
    m_p_ctx_wrapper->trace("in1", enc[1]);
    // the result cells are the independent units: the matrix with more rows (M0) / columns (M1) is cut into
    // one block per GPU, the other goes to every GPU.  CipherBatchAxis holds one ciphertext per element: an item of M0 is a
    // row of cols_M0 consecutive ciphertexts, an item of M1 a column (ciphertexts k * cols_M1 + j).
    const std::size_t c0 = m_w_params.cols_M0(), c1 = m_w_params.cols_M1();
    LoadedMats loaded;
    if (m_algo == MatMultAlgo::CipherBatchAxis)
        loaded = m_p_ctx_wrapper->loadGrid(enc[0], c0, enc[1], c0, [c0, c1](std::size_t j) {
            std::vector<std::size_t> col(c0);
            for (std::size_t k = 0; k < c0; ++k) col[k] = k * c1 + j;
            return col;
        });
    else
        loaded = m_p_ctx_wrapper->loadGrid(enc[0], 1, enc[1], 1);
    return this->getEngine().template createHandle<LoadedMats>(sizeof(LoadedMats), 0, std::move(loaded));
}

template <bool CKKS> void MatMultBenchmarkT<CKKS>::store(Handle remote_data, Handle *p_local_data, std::uint64_t count)
{
    if (count > 0) {
        std::memset(p_local_data, 0, sizeof(Handle) * count);
        const ShardedCiphertexts &res = this->getEngine().template retrieveFromHandle<ShardedCiphertexts>(remote_data, ResultCipherTag);
        std::vector<Ciphertext> host  = m_p_ctx_wrapper->gather(res);
        m_p_ctx_wrapper->trace("out", host);
        p_local_data[0] = this->getEngine().template createHandle<std::vector<Ciphertext>>(sizeof(host), ResultCipherTag, std::move(host));
    }
}

template <bool CKKS> Handle MatMultBenchmarkT<CKKS>::decrypt(Handle encrypted_data)
{
    const std::vector<Ciphertext> &enc = this->getEngine().template retrieveFromHandle<std::vector<Ciphertext>>(encrypted_data, ResultCipherTag);
    std::vector<Plaintext> plain       = m_p_ctx_wrapper->decrypt(enc);
    return this->getEngine().template createHandle<std::vector<Plaintext>>(sizeof(plain), ResultPlainTag, std::move(plain));
}

// one buffer rows_M0 x cols_M1, row-major; "copy as much as fits" (…ckks_matmult_row…:317-326)
template <bool CKKS> void MatMultBenchmarkT<CKKS>::decode(Handle encoded_data, DataPackCollection *p_native)
{
    const std::vector<Plaintext> &plain = this->getEngine().template retrieveFromHandle<std::vector<Plaintext>>(encoded_data, ResultPlainTag);
    if (p_native->pack_count == 0) return;
    DataPack &pack = p_native->p_data_packs[findDataPackIndex(*p_native, 0)];
    if (pack.buffer_count == 0 || !pack.p_buffers[0].p) return;
    Scalar *out                = reinterpret_cast<Scalar *>(pack.p_buffers[0].p);
    const std::size_t capacity = pack.p_buffers[0].size / sizeof(Scalar);
    const std::size_t r0 = m_w_params.rows_M0(), c1 = m_w_params.cols_M1();
    const std::size_t row_size = m_p_ctx_wrapper->slotCount() / 2;
    auto slots = [&](const Plaintext &p) {
        std::vector<Scalar> v;
        if constexpr (CKKS) v = m_p_ctx_wrapper->decodeCKKS(p);
        else v = m_p_ctx_wrapper->decodeBFV(p);
        return v;
    };
    auto put = [&](std::size_t idx, Scalar v) {
        if (idx < capacity) {
            if constexpr (CKKS) out[idx] = flushTiny(v);
            else out[idx] = v;
        }
    };
    if (m_algo == MatMultAlgo::Row) {
        for (std::size_t ct = 0; ct < plain.size(); ++ct) {
            const std::vector<Scalar> v = slots(plain[ct]);
            if (CKKS) {
                for (std::size_t k = 0; k < c1; ++k) put(ct * c1 + k, v[k]);
            } else {
                for (std::size_t k = 0; k < c1; ++k) {
                    put(2 * ct * c1 + k, v[k]);
                    if (2 * ct + 1 < r0) put((2 * ct + 1) * c1 + k, v[row_size + k]);
                }
            }
        }
    } else {   // Val / CipherBatchAxis: one ciphertext per output cell, value in slot 0
        for (std::size_t cell = 0; cell < plain.size() && cell < r0 * c1; ++cell) put(cell, slots(plain[cell])[0]);
    }
}

// ------------------------------------------------------------------ operate
template <bool CKKS>
Handle MatMultBenchmarkT<CKKS>::operate(Handle h_remote_packed, const ParameterIndexer *p_param_indexers, std::uint64_t indexers_count)
{
    if (indexers_count < 2) throw HEBenchError(HEBERROR_MSG_CLASS("Invalid number of indexers. Expected 2."), HEBENCH_ECODE_INVALID_ARGS);
    for (int i = 0; i < 2; ++i) {   // no sub-range indexing (R/src/benchmarks/ckks/seal_ckks_matmultval_benchmark.cpp:451-461)
        if (p_param_indexers[i].value_index > 0) throw HEBenchError(HEBERROR_MSG_CLASS("Unexpected index in parameter indexer."), HEBENCH_ECODE_INVALID_ARGS);
        if (p_param_indexers[i].batch_size > 1) throw HEBenchError(HEBERROR_MSG_CLASS("Batch size must be 1 for latency test."), HEBENCH_ECODE_INVALID_ARGS);
    }
    const LoadedMats &in = this->getEngine().template retrieveFromHandle<LoadedMats>(h_remote_packed);
    m_p_ctx_wrapper->beginOperate();
    ShardedCiphertexts out = m_algo == MatMultAlgo::Val ? operateVal(in) : m_algo == MatMultAlgo::Row ? operateRow(in) : operateCipherBatchAxis(in);
    m_p_ctx_wrapper->prepareStore(out);
    m_p_ctx_wrapper->endOperate(1);
    return this->getEngine().template createHandle<ShardedCiphertexts>(sizeof(ShardedCiphertexts), ResultCipherTag, std::move(out));
}

// out[i][j] = accumulate(rescale(relin(M0_i * M1T_j)), cols_M0)   (…matmultval…:235-270; BFV: no rescale)
template <bool CKKS> ShardedCiphertexts MatMultBenchmarkT<CKKS>::operateVal(const LoadedMats &in)
{
    SEALContextWrapper &cw = *m_p_ctx_wrapper;
    const std::uint64_t r0 = m_w_params.rows_M0(), c0 = m_w_params.cols_M0(), c1 = m_w_params.cols_M1();
    const std::uint64_t v0[2] = { 0, 0 }, dims[2] = { r0, c1 };
    ShardedCiphertexts out;
    out.n_total = r0 * c1;
    out.shard.resize(cw.gpuCount());
    out.ids.resize(cw.gpuCount());
    out.first.assign(cw.gpuCount() + 1, 0);
    cw.forEachGpu([&](int g) {
        GridShare sh          = cw.gridShare(in, g, v0, dims);
        const std::uint64_t n = sh.result.size();
        b200he_ctx *c    = cw.device(g);
        DeviceBatchPtr r = cw.newBatch(g);
        cw.check(b200he_multiply(c, in.p[0].shard[g]->get(), sh.ai.data(), in.p[1].shard[g]->get(), sh.bi.data(), n, r->get()), "b200he_multiply");
        if (n > 0) {
            if (CKKS) {
                // relinearize_inplace + rescale_to_next_inplace (R/src/benchmarks/ckks/seal_ckks_matmultval_benchmark.cpp:252-255): fused, same bits
                cw.check(b200he_relinearize_rescale(c, r->get(), r->get()), "b200he_relinearize_rescale");
                cw.accumulateCKKS(*r, c0);
            } else {
                cw.check(b200he_relinearize(c, r->get(), r->get()), "b200he_relinearize");
                cw.accumulateBFV(*r, c0);
            }
        }
        out.shard[g] = r;
        out.ids[g]   = std::move(sh.result);
    });
    return out;
}

// per ciphertext of M0: base = relin(A * B); result = base + sum_{j=1}^{cols_M0-1} rotate(base, j*spacers)   (…matmult_row…:472-523)
template <bool CKKS> ShardedCiphertexts MatMultBenchmarkT<CKKS>::operateRow(const LoadedMats &in)
{
    SEALContextWrapper &cw = *m_p_ctx_wrapper;
    const std::uint64_t c0 = m_w_params.cols_M0();
    const int spacers      = (int)((cw.polyModulusDegree() / 2) / c0);
    if (in.split != 0 || in.p[1].total() != 1) throw HEBenchError(HEBERROR_MSG_CLASS("MatMultRow expects M1 packed in one ciphertext."), HEBENCH_ECODE_INVALID_ARGS);
    ShardedCiphertexts out;
    out.n_total = in.p[0].total();
    out.first   = in.p[0].first;   // one result ciphertext per ciphertext of M0, same blocks
    out.shard.resize(cw.gpuCount());
    cw.forEachGpu([&](int g) {
        const std::uint64_t n = in.p[0].first[g + 1] - in.p[0].first[g];
        std::vector<uint32_t> bi(n, 0);
        b200he_ctx *c = cw.device(g);
        DeviceBatchPtr base = cw.newBatch(g), result = cw.newBatch(g), rotated = cw.newBatch(g);
        cw.check(b200he_multiply(c, in.p[0].shard[g]->get(), nullptr, in.p[1].shard[g]->get(), bi.data(), n, base->get()), "b200he_multiply");
        if (n > 0) {
            cw.check(b200he_relinearize(c, base->get(), base->get()), "b200he_relinearize");
            cw.check(b200he_gather(c, base->get(), nullptr, n, result->get()), "b200he_gather");
            for (std::uint64_t j = 1; j < c0; ++j) {
                cw.check(b200he_rotate(c, base->get(), (int)j * spacers, rotated->get()), "b200he_rotate");
                cw.check(b200he_add(c, result->get(), nullptr, rotated->get(), nullptr, n, result->get()), "b200he_add");
            }
        } else
            cw.check(b200he_gather(c, base->get(), nullptr, 0, result->get()), "b200he_gather");
        out.shard[g] = result;
    });
    return out;
}

// out[i][j] = sum_k m0[i][k] * m1[k][j]; CKKS keeps the size-3 products and relinearizes + rescales once per cell,
// BFV relinearizes every product (…ckks…cipherbatchaxis…:385-441, …bfv…cipherbatchaxis…:400-408)
template <bool CKKS> ShardedCiphertexts MatMultBenchmarkT<CKKS>::operateCipherBatchAxis(const LoadedMats &in)
{
    SEALContextWrapper &cw = *m_p_ctx_wrapper;
    const std::uint64_t r0 = m_w_params.rows_M0(), c0 = m_w_params.cols_M0(), c1 = m_w_params.cols_M1();
    ShardedCiphertexts out;
    out.n_total = r0 * c1;
    out.shard.resize(cw.gpuCount());
    out.ids.resize(cw.gpuCount());
    out.first.assign(cw.gpuCount() + 1, 0);
    cw.forEachGpu([&](int g) {
        // this GPU's block: rows [i0, i1) x columns [j0, j1) of the result (one of the two ranges is the full one)
        const std::uint64_t i0 = in.split == 0 ? in.p[0].first[g] : 0, i1 = in.split == 0 ? in.p[0].first[g + 1] : r0;
        const std::uint64_t j0 = in.split == 1 ? in.p[1].first[g] : 0, j1 = in.split == 1 ? in.p[1].first[g + 1] : c1;
        const std::uint64_t nr = i1 - i0, nc = j1 - j0, n = nr * nc;
        b200he_ctx *c      = cw.device(g);
        DeviceBatchPtr acc = cw.newBatch(g);
        b200he_batch *A = in.p[0].shard[g]->get(), *B = in.p[1].shard[g]->get();   // A: [nr][c0] row-major, B: columns, [nc][c0]
        out.ids[g].resize(n);
        for (std::uint64_t cell = 0; cell < n; ++cell) out.ids[g][cell] = (i0 + cell / nc) * c1 + (j0 + cell % nc);
        if (CKKS) {
            // sum_k multiply(m0[i][k], m1[k][j]) in one pass with register accumulators (same bits as the reference's
            // multiply / add_inplace chain), then ONE relinearize + rescale per cell
            if (n > 0) {
                cw.check(b200he_matmul_accumulate(c, A, B, nr, c0, nc, acc->get()), "b200he_matmul_accumulate");
                cw.check(b200he_relinearize_rescale(c, acc->get(), acc->get()), "b200he_relinearize_rescale");
            } else
                cw.check(b200he_batch_resize(acc->get(), 0, 2, (int)cw.topLevel() - 1, 1, cw.scale()), "b200he_batch_resize");
        } else {
            DeviceBatchPtr prod = cw.newBatch(g);
            std::vector<uint32_t> ai(n), bi(n);
            for (std::uint64_t k = 0; k < c0; ++k) {
                for (std::uint64_t cell = 0; cell < n; ++cell) {
                    ai[cell] = (uint32_t)((cell / nc) * c0 + k);
                    bi[cell] = (uint32_t)((cell % nc) * c0 + k);
                }
                b200he_batch *dst = k == 0 ? acc->get() : prod->get();
                cw.check(b200he_multiply(c, A, ai.data(), B, bi.data(), n, dst), "b200he_multiply");
                if (n > 0) cw.check(b200he_relinearize(c, dst, dst), "b200he_relinearize");
                if (k > 0) cw.check(b200he_add(c, acc->get(), nullptr, prod->get(), nullptr, n, acc->get()), "b200he_add");
            }
        }
        out.shard[g] = acc;
    });
    return out;
}

template class MatMultBenchmarkDescriptionT<true>;
template class MatMultBenchmarkDescriptionT<false>;
template class MatMultBenchmarkT<true>;
template class MatMultBenchmarkT<false>;

}   // namespace sbe
