// hebench_cpp.cpp -- the exported HEBench API Bridge C ABI (api.h) implemented on the C++ base classes
// of hebench.hpp: every entry point converts exceptions into error codes and keeps the last error
// text on the engine, as upstream's hebench_cpp static library does for the reference backend.
#include "hebench.hpp"

#include <algorithm>

namespace hebench {
namespace cpp {

BaseEngine::BaseEngine()
{
    m_errors[HEBENCH_ECODE_SUCCESS]        = "Success";
    m_errors[HEBENCH_ECODE_CRITICAL_ERROR] = "Critical error";
    m_errors[HEBENCH_ECODE_INVALID_ARGS]   = "Invalid arguments";
}
static const std::string s_unknown = "Unknown";
const std::string &BaseEngine::getSchemeName(APIBridge::Scheme s) const
{
    auto it = m_schemes.find(s);
    return it == m_schemes.end() ? s_unknown : it->second;
}
const std::string &BaseEngine::getSecurityName(APIBridge::Security s) const
{
    auto it = m_securities.find(s);
    return it == m_securities.end() ? s_unknown : it->second;
}
const std::string &BaseEngine::getErrorDescription(APIBridge::ErrorCode c) const
{
    auto it = m_errors.find(c);
    return it == m_errors.end() ? s_unknown : it->second;
}

namespace {

struct BenchDescRef {
    std::shared_ptr<BenchmarkDescription> p;
};
struct BenchmarkRef {
    BaseEngine *engine;
    std::shared_ptr<BenchmarkDescription> desc;
    BaseBenchmark *bench;
    ~BenchmarkRef()
    {
        if (bench) desc->destroyBenchmark(bench);
    }
};
struct EngineRef {
    BaseEngine *engine;
    ~EngineRef() { destroyEngine(engine); }
};

thread_local std::string t_last_error = "";

BaseEngine &engineFrom(APIBridge::Handle h)
{
    EngineObject *o = reinterpret_cast<EngineObject *>(h.p);
    if (!o || (h.tag & EngineObject::tag) != EngineObject::tag) throw HEBenchError("invalid engine handle", HEBENCH_ECODE_CRITICAL_ERROR);
    return *reinterpret_cast<EngineRef *>(o->obj.get())->engine;
}
BenchmarkRef &benchFrom(APIBridge::Handle h)
{
    EngineObject *o = reinterpret_cast<EngineObject *>(h.p);
    if (!o || (h.tag & EngineObject::tagBenchmark) != EngineObject::tagBenchmark) throw HEBenchError("invalid benchmark handle", HEBENCH_ECODE_CRITICAL_ERROR);
    return *reinterpret_cast<BenchmarkRef *>(o->obj.get());
}
BenchmarkDescription &descFrom(APIBridge::Handle h)
{
    EngineObject *o = reinterpret_cast<EngineObject *>(h.p);
    if (!o || (h.tag & EngineObject::tagBenchDesc) != EngineObject::tagBenchDesc) throw HEBenchError("invalid benchmark description handle", HEBENCH_ECODE_CRITICAL_ERROR);
    return *reinterpret_cast<BenchDescRef *>(o->obj.get())->p;
}
template <class T> APIBridge::Handle wrap(std::shared_ptr<T> p, std::int64_t tag)
{
    EngineObject *o = new EngineObject();
    o->obj          = p;
    o->obj_tag      = tag;
    return APIBridge::Handle{ o, 0, tag };
}
std::uint64_t copyText(const std::string &s, char *p, std::uint64_t size)
{
    const std::uint64_t need = s.size() + 1;
    if (p && size > 0) {
        const std::uint64_t n = std::min<std::uint64_t>(need, size);
        std::memcpy(p, s.c_str(), n - 1);
        p[n - 1] = '\0';
    }
    return need;
}

template <class F> APIBridge::ErrorCode guarded(BaseEngine *engine, F &&f)
{
    APIBridge::ErrorCode rc = HEBENCH_ECODE_SUCCESS;
    std::string text;
    try {
        f();
    } catch (HEBenchError &e) {
        rc   = e.getErrorCode();
        text = e.what();
    } catch (std::exception &e) {
        rc   = HEBENCH_ECODE_CRITICAL_ERROR;
        text = e.what();
    } catch (...) {
        rc   = HEBENCH_ECODE_CRITICAL_ERROR;
        text = "unexpected error";
    }
    if (rc != HEBENCH_ECODE_SUCCESS) {
        t_last_error = text;
        if (engine) engine->setLastError(rc, text);
    }
    return rc;
}

}   // namespace
}   // namespace cpp

namespace APIBridge {
using namespace hebench::cpp;

extern "C" {

ErrorCode initEngine(Handle *h_engine, const int8_t *p_buffer, uint64_t size)
{
    return guarded(nullptr, [&]() {
        if (!h_engine) throw HEBenchError("initEngine(): null handle pointer", HEBENCH_ECODE_INVALID_ARGS);
        BaseEngine *e = createEngine(p_buffer, size);
        *h_engine     = wrap(std::shared_ptr<EngineRef>(new EngineRef{ e }), EngineObject::tag);
    });
}

ErrorCode destroyHandle(Handle h)
{
    return guarded(nullptr, [&]() { delete reinterpret_cast<EngineObject *>(h.p); });
}

ErrorCode subscribeBenchmarksCount(Handle h_engine, uint64_t *p_count)
{
    BaseEngine *e = nullptr;
    return guarded(e, [&]() {
        e = &engineFrom(h_engine);
        if (!p_count) throw HEBenchError("subscribeBenchmarksCount(): null pointer", HEBENCH_ECODE_INVALID_ARGS);
        *p_count = e->benchmarks().size();
    });
}

ErrorCode subscribeBenchmarks(Handle h_engine, Handle *p_h_bench_descs, uint64_t count)
{
    return guarded(nullptr, [&]() {
        BaseEngine &e = engineFrom(h_engine);
        if (!p_h_bench_descs) throw HEBenchError("subscribeBenchmarks(): null pointer", HEBENCH_ECODE_INVALID_ARGS);
        const uint64_t n = std::min<uint64_t>(count, e.benchmarks().size());
        for (uint64_t i = 0; i < n; ++i)
            p_h_bench_descs[i] = wrap(std::shared_ptr<BenchDescRef>(new BenchDescRef{ e.benchmarks()[i] }), EngineObject::tagBenchDesc);
    });
}

ErrorCode getWorkloadParamsDetails(Handle h_engine, Handle h_bench_desc, uint64_t *p_param_count, uint64_t *p_default_count)
{
    return guarded(nullptr, [&]() {
        engineFrom(h_engine);
        BenchmarkDescription &d = descFrom(h_bench_desc);
        if (p_param_count) *p_param_count = d.getWorkloadParameterCount();
        if (p_default_count) *p_default_count = d.getWorkloadParameters().size();
    });
}

ErrorCode describeBenchmark(Handle h_engine, Handle h_bench_desc, BenchmarkDescriptor *p_bench_desc, WorkloadParams *p_default_params,
                            uint64_t default_count)
{
    return guarded(nullptr, [&]() {
        engineFrom(h_engine);
        BenchmarkDescription &d = descFrom(h_bench_desc);
        if (p_bench_desc) *p_bench_desc = d.getBenchmarkDescriptor();
        if (p_default_params) {
            const auto &sets = d.getWorkloadParameters();
            for (uint64_t i = 0; i < default_count && i < sets.size(); ++i) {
                const uint64_t n = std::min<uint64_t>(p_default_params[i].count, sets[i].size());
                for (uint64_t k = 0; k < n; ++k) p_default_params[i].params[k] = sets[i][k];
            }
        }
    });
}

ErrorCode createBenchmark(Handle h_engine, Handle h_bench_desc, const WorkloadParams *p_params, Handle *h_benchmark)
{
    BaseEngine *e = nullptr;
    try { e = &engineFrom(h_engine); } catch (...) {}
    return guarded(e, [&]() {
        BaseEngine &engine = engineFrom(h_engine);
        if (!h_benchmark) throw HEBenchError("createBenchmark(): null handle pointer", HEBENCH_ECODE_INVALID_ARGS);
        EngineObject *o = reinterpret_cast<EngineObject *>(h_bench_desc.p);
        descFrom(h_bench_desc);
        std::shared_ptr<BenchmarkDescription> d = reinterpret_cast<BenchDescRef *>(o->obj.get())->p;
        if (d->getWorkloadParameterCount() > 0 && (!p_params || p_params->count < d->getWorkloadParameterCount()))
            throw HEBenchError("createBenchmark(): workload parameters missing", HEBENCH_ECODE_INVALID_ARGS);
        BaseBenchmark *b = d->createBenchmark(engine, p_params);
        if (!b) throw HEBenchError("createBenchmark(): backend returned null", HEBENCH_ECODE_CRITICAL_ERROR);
        *h_benchmark = wrap(std::shared_ptr<BenchmarkRef>(new BenchmarkRef{ &engine, d, b }), EngineObject::tagBenchmark);
    });
}

ErrorCode initBenchmark(Handle h_benchmark, const BenchmarkDescriptor *p_concrete_desc)
{
    BaseEngine *e = nullptr;
    try { e = benchFrom(h_benchmark).engine; } catch (...) {}
    return guarded(e, [&]() {
        if (!p_concrete_desc) throw HEBenchError("initBenchmark(): null descriptor", HEBENCH_ECODE_INVALID_ARGS);
        benchFrom(h_benchmark).bench->initialize(*p_concrete_desc);
    });
}

#define BENCH_CALL(h_benchmark, BODY)                                              \
    BaseEngine *e_ = nullptr;                                                      \
    try { e_ = benchFrom(h_benchmark).engine; } catch (...) {}                     \
    return guarded(e_, [&]() { BaseBenchmark &b = *benchFrom(h_benchmark).bench; BODY; })

ErrorCode encode(Handle h_benchmark, const DataPackCollection *p_parameters, Handle *h_plaintext)
{
    BENCH_CALL(h_benchmark, { if (!h_plaintext) throw HEBenchError("encode(): null output", HEBENCH_ECODE_INVALID_ARGS); *h_plaintext = b.encode(p_parameters); });
}
ErrorCode decode(Handle h_benchmark, Handle h_plaintext, DataPackCollection *p_native) { BENCH_CALL(h_benchmark, b.decode(h_plaintext, p_native)); }
ErrorCode encrypt(Handle h_benchmark, Handle h_plaintext, Handle *h_ciphertext)
{
    BENCH_CALL(h_benchmark, { if (!h_ciphertext) throw HEBenchError("encrypt(): null output", HEBENCH_ECODE_INVALID_ARGS); *h_ciphertext = b.encrypt(h_plaintext); });
}
ErrorCode decrypt(Handle h_benchmark, Handle h_ciphertext, Handle *h_plaintext)
{
    BENCH_CALL(h_benchmark, { if (!h_plaintext) throw HEBenchError("decrypt(): null output", HEBENCH_ECODE_INVALID_ARGS); *h_plaintext = b.decrypt(h_ciphertext); });
}
ErrorCode load(Handle h_benchmark, const Handle *h_local_packed_params, uint64_t local_count, Handle *h_remote)
{
    BENCH_CALL(h_benchmark, { if (!h_remote) throw HEBenchError("load(): null output", HEBENCH_ECODE_INVALID_ARGS); *h_remote = b.load(h_local_packed_params, local_count); });
}
ErrorCode store(Handle h_benchmark, Handle h_remote, Handle *h_local_packed_params, uint64_t local_count)
{
    BENCH_CALL(h_benchmark, b.store(h_remote, h_local_packed_params, local_count));
}
ErrorCode operate(Handle h_benchmark, Handle h_remote_packed_params, const ParameterIndexer *p_param_indexers, uint64_t indexers_count,
                  Handle *h_remote_output)
{
    BENCH_CALL(h_benchmark, { if (!h_remote_output) throw HEBenchError("operate(): null output", HEBENCH_ECODE_INVALID_ARGS); *h_remote_output = b.operate(h_remote_packed_params, p_param_indexers, indexers_count); });
}

uint64_t getSchemeName(Handle h_engine, Scheme s, char *p_name, uint64_t size)
{
    try { return copyText(engineFrom(h_engine).getSchemeName(s), p_name, size); } catch (...) { return 0; }
}
uint64_t getSchemeSecurityName(Handle h_engine, Scheme s, Security sec, char *p_name, uint64_t size)
{
    (void)s;
    try { return copyText(engineFrom(h_engine).getSecurityName(sec), p_name, size); } catch (...) { return 0; }
}
uint64_t getBenchmarkDescriptionEx(Handle h_engine, Handle h_bench_desc, const WorkloadParams *p_w_params, char *p_description, uint64_t size)
{
    try {
        engineFrom(h_engine);
        return copyText(descFrom(h_bench_desc).getBenchmarkDescription(p_w_params), p_description, size);
    } catch (...) { return 0; }
}
uint64_t getErrorDescription(Handle h_engine, ErrorCode code, char *p_description, uint64_t size)
{
    try { return copyText(engineFrom(h_engine).getErrorDescription(code), p_description, size); } catch (...) { return 0; }
}
uint64_t getLastErrorDescription(Handle h_engine, char *p_description, uint64_t size)
{
    try {
        if (h_engine.p) return copyText(engineFrom(h_engine).getLastErrorDescription(), p_description, size);
    } catch (...) {}
    return copyText(t_last_error, p_description, size);
}

}   // extern "C"
}   // namespace APIBridge
}   // namespace hebench
