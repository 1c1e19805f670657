// hebench.hpp -- header-compatible SUBSET of the HEBench C++ wrapper (hebench_cpp, upstream
// hebench/frontend v0.9.0-beta) written for this repository: the base classes the reference backend
// derives from (BaseEngine, BenchmarkDescription, BaseBenchmark), the error type, the workload
// parameter helpers and the tagged-handle plumbing.  The upstream library is fetched from the
// network by the reference's CMake (R/cmake/third-party/API_BRIDGE.cmake:6-13) and is absent
// offline; the census of what the reference uses is in SURVEY.md §8b.  hebench_cpp.cpp implements
// the exported C ABI (api.h) on top of these classes, like upstream's static library does.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../api.h"
#include "../types.h"

#define HEBERROR_DECLARE_CLASS_NAME(class_name) static constexpr const char *m_private_class_name = #class_name;
#define HEBERROR_MSG(message) (std::string(__func__) + "(): " + std::string(message))
#define HEBERROR_MSG_CLASS(message) (std::string(m_private_class_name) + "::" + std::string(__func__) + "(): " + std::string(message))

namespace hebench {
namespace cpp {

class HEBenchError : public std::runtime_error
{
public:
    HEBenchError(const std::string &msg, APIBridge::ErrorCode err_code) : std::runtime_error(msg), m_err_code(err_code) {}
    APIBridge::ErrorCode getErrorCode() const { return m_err_code; }

private:
    APIBridge::ErrorCode m_err_code;
};

class ITaggedObject
{
public:
    virtual ~ITaggedObject() {}
    virtual std::int64_t classTag() const = 0;
};

// What a Handle points at: a tagged, reference-counted, type-erased object.
struct EngineObject {
    static constexpr std::int64_t tag        = 0x4000000000000000;   // the engine itself
    static constexpr std::int64_t tagBenchDesc = 0x2000000000000000;
    static constexpr std::int64_t tagBenchmark = 0x1000000000000000;
    std::shared_ptr<void> obj;
    std::int64_t obj_tag = 0;
    std::uint64_t size   = 0;
};

namespace WorkloadParams {

class Common
{
public:
    Common() {}
    explicit Common(const APIBridge::WorkloadParams &p) : m_w_params(p.params, p.params + p.count) {}
    virtual ~Common() {}
    template <class T> void add(const T &value, const std::string &name = std::string());
    template <class T> void set(std::size_t index, const T &value, const std::string &name = std::string());
    template <class T> T get(std::size_t index) const;
    const std::vector<APIBridge::WorkloadParam> &getParams() const { return m_w_params; }
    std::size_t size() const { return m_w_params.size(); }

protected:
    std::vector<APIBridge::WorkloadParam> m_w_params;
    void requireMin(std::size_t n, const char *who) const
    {
        if (m_w_params.size() < n) throw HEBenchError(std::string(who) + ": insufficient workload parameters", HEBENCH_ECODE_INVALID_ARGS);
    }
};
template <> inline void Common::set<std::uint64_t>(std::size_t i, const std::uint64_t &v, const std::string &name)
{
    m_w_params.at(i).data_type = APIBridge::WorkloadParamType::UInt64_wp;
    m_w_params[i].u_param      = v;
    std::strncpy(m_w_params[i].name, name.c_str(), HEBENCH_MAX_BUFFER_SIZE - 1);
}
template <> inline void Common::set<std::int64_t>(std::size_t i, const std::int64_t &v, const std::string &name)
{
    m_w_params.at(i).data_type = APIBridge::WorkloadParamType::Int64_wp;
    m_w_params[i].i_param      = v;
    std::strncpy(m_w_params[i].name, name.c_str(), HEBENCH_MAX_BUFFER_SIZE - 1);
}
template <> inline void Common::set<double>(std::size_t i, const double &v, const std::string &name)
{
    m_w_params.at(i).data_type = APIBridge::WorkloadParamType::Float64_wp;
    m_w_params[i].f_param      = v;
    std::strncpy(m_w_params[i].name, name.c_str(), HEBENCH_MAX_BUFFER_SIZE - 1);
}
template <class T> inline void Common::add(const T &value, const std::string &name)
{
    APIBridge::WorkloadParam p;
    std::memset(&p, 0, sizeof(p));
    m_w_params.push_back(p);
    set<T>(m_w_params.size() - 1, value, name);
}
template <> inline std::uint64_t Common::get<std::uint64_t>(std::size_t i) const { return m_w_params.at(i).u_param; }
template <> inline std::int64_t Common::get<std::int64_t>(std::size_t i) const { return m_w_params.at(i).i_param; }
template <> inline double Common::get<double>(std::size_t i) const { return m_w_params.at(i).f_param; }

class VectorSize : public Common
{
public:
    VectorSize(std::uint64_t n = 0) { add<std::uint64_t>(n, "n"); }
    VectorSize(const APIBridge::WorkloadParams &p) : Common(p) { requireMin(1, "VectorSize"); }
    std::uint64_t n() const { return get<std::uint64_t>(0); }
};
typedef VectorSize EltwiseAdd;
typedef VectorSize EltwiseMultiply;
typedef VectorSize DotProduct;
typedef VectorSize LogisticRegression;

class MatrixMultiply : public Common
{
public:
    MatrixMultiply(std::uint64_t r0 = 0, std::uint64_t c0 = 0, std::uint64_t c1 = 0)
    {
        add<std::uint64_t>(r0, "rows_M0");
        add<std::uint64_t>(c0, "cols_M0");
        add<std::uint64_t>(c1, "cols_M1");
    }
    MatrixMultiply(const APIBridge::WorkloadParams &p) : Common(p) { requireMin(3, "MatrixMultiply"); }
    std::uint64_t rows_M0() const { return get<std::uint64_t>(0); }
    std::uint64_t cols_M0() const { return get<std::uint64_t>(1); }
    std::uint64_t cols_M1() const { return get<std::uint64_t>(2); }
};

}   // namespace WorkloadParams

class BaseEngine;
class BaseBenchmark;

class BenchmarkDescription
{
public:
    BenchmarkDescription() { std::memset(&m_descriptor, 0, sizeof(m_descriptor)); }
    virtual ~BenchmarkDescription() {}
    const APIBridge::BenchmarkDescriptor &getBenchmarkDescriptor() const { return m_descriptor; }
    const std::vector<std::vector<APIBridge::WorkloadParam>> &getWorkloadParameters() const { return m_default_params; }
    std::size_t getWorkloadParameterCount() const { return m_default_params.empty() ? 0 : m_default_params.front().size(); }
    virtual BaseBenchmark *createBenchmark(BaseEngine &engine, const APIBridge::WorkloadParams *p_params) = 0;
    virtual void destroyBenchmark(BaseBenchmark *p_bench)                                                = 0;
    // text appended to the benchmark's report header (CSV)
    virtual std::string getBenchmarkDescription(const APIBridge::WorkloadParams *p_w_params) const
    {
        (void)p_w_params;
        return std::string();
    }

protected:
    APIBridge::BenchmarkDescriptor m_descriptor;
    void addDefaultParameters(const WorkloadParams::Common &default_params_set) { m_default_params.push_back(default_params_set.getParams()); }

private:
    std::vector<std::vector<APIBridge::WorkloadParam>> m_default_params;
};

class BaseBenchmark : public ITaggedObject
{
public:
    static constexpr std::int64_t tag = EngineObject::tagBenchmark;
    ~BaseBenchmark() override {}
    virtual APIBridge::Handle encode(const APIBridge::DataPackCollection *p_parameters)                              = 0;
    virtual void decode(APIBridge::Handle encoded_data, APIBridge::DataPackCollection *p_native)                     = 0;
    virtual APIBridge::Handle encrypt(APIBridge::Handle encoded_data)                                                = 0;
    virtual APIBridge::Handle decrypt(APIBridge::Handle encrypted_data)                                              = 0;
    virtual APIBridge::Handle load(const APIBridge::Handle *p_local_data, std::uint64_t count)                       = 0;
    virtual void store(APIBridge::Handle remote_data, APIBridge::Handle *p_local_data, std::uint64_t count)          = 0;
    virtual APIBridge::Handle operate(APIBridge::Handle h_remote_packed, const APIBridge::ParameterIndexer *p_param_indexers,
                                      std::uint64_t indexers_count)                                                  = 0;
    virtual void initialize(const APIBridge::BenchmarkDescriptor &concrete_desc) { m_descriptor = concrete_desc; }
    std::int64_t classTag() const override { return BaseBenchmark::tag; }
    BaseEngine &getEngine() const { return m_engine; }
    const APIBridge::BenchmarkDescriptor &getDescriptor() const { return m_descriptor; }

    // data pack for operation parameter `param_position`; throws when absent
    static const APIBridge::DataPack &findDataPack(const APIBridge::DataPackCollection &c, std::uint64_t param_position)
    {
        return c.p_data_packs[findDataPackIndex(c, param_position)];
    }
    static std::uint64_t findDataPackIndex(const APIBridge::DataPackCollection &c, std::uint64_t param_position)
    {
        for (std::uint64_t i = 0; i < c.pack_count; ++i)
            if (c.p_data_packs[i].param_position == param_position) return i;
        throw HEBenchError("BaseBenchmark::findDataPackIndex(): no data pack for parameter " + std::to_string(param_position), HEBENCH_ECODE_INVALID_ARGS);
    }

protected:
    BaseBenchmark(BaseEngine &engine, const APIBridge::BenchmarkDescriptor &bench_desc, const APIBridge::WorkloadParams &bench_params)
        : m_engine(engine), m_descriptor(bench_desc)
    {
        (void)bench_params;
    }

private:
    BaseEngine &m_engine;
    APIBridge::BenchmarkDescriptor m_descriptor;
};

class BaseEngine : public ITaggedObject
{
public:
    static constexpr std::int64_t tag = EngineObject::tag;
    ~BaseEngine() override {}
    std::int64_t classTag() const override { return BaseEngine::tag; }

    // ---- handle plumbing used by the benchmarks
    template <class T> APIBridge::Handle createHandle(std::uint64_t size, std::int64_t extra_tags, T &&value) const
    {
        typedef typename std::decay<T>::type V;
        EngineObject *p = new EngineObject();
        p->obj          = std::make_shared<V>(std::forward<T>(value));
        p->obj_tag      = extra_tags;
        p->size         = size;
        return APIBridge::Handle{ p, size, extra_tags };
    }
    template <class T> T &retrieveFromHandle(APIBridge::Handle h, std::int64_t extra_tags = 0) const
    {
        EngineObject *p = reinterpret_cast<EngineObject *>(h.p);
        if (!p || !p->obj) throw HEBenchError("BaseEngine::retrieveFromHandle(): invalid null handle", HEBENCH_ECODE_INVALID_ARGS);
        if ((h.tag & extra_tags) != extra_tags || (p->obj_tag & extra_tags) != extra_tags)
            throw HEBenchError("BaseEngine::retrieveFromHandle(): handle has the wrong tag", HEBENCH_ECODE_INVALID_ARGS);
        return *reinterpret_cast<T *>(p->obj.get());
    }
    APIBridge::Handle duplicateHandle(APIBridge::Handle h, std::int64_t extra_tags = 0) const
    {
        EngineObject *src = reinterpret_cast<EngineObject *>(h.p);
        if (!src || !src->obj) throw HEBenchError("BaseEngine::duplicateHandle(): invalid null handle", HEBENCH_ECODE_INVALID_ARGS);
        if ((h.tag & extra_tags) != extra_tags) throw HEBenchError("BaseEngine::duplicateHandle(): handle has the wrong tag", HEBENCH_ECODE_INVALID_ARGS);
        EngineObject *p = new EngineObject(*src);   // shares the payload
        return APIBridge::Handle{ p, h.size, h.tag };
    }

    // ---- registry read by the C ABI
    const std::vector<std::shared_ptr<BenchmarkDescription>> &benchmarks() const { return m_descriptions; }
    const std::string &getSchemeName(APIBridge::Scheme s) const;
    const std::string &getSecurityName(APIBridge::Security s) const;
    const std::string &getErrorDescription(APIBridge::ErrorCode c) const;
    void setLastError(APIBridge::ErrorCode c, const std::string &text)
    {
        m_last_error      = c;
        m_last_error_text = text;
    }
    const std::string &getLastErrorDescription() const { return m_last_error_text; }
    APIBridge::ErrorCode getLastError() const { return m_last_error; }

protected:
    BaseEngine();
    virtual void init() = 0;
    void addBenchmarkDescription(std::shared_ptr<BenchmarkDescription> p) { m_descriptions.push_back(p); }
    void addSchemeName(APIBridge::Scheme s, const std::string &name) { m_schemes[s] = name; }
    void addSecurityName(APIBridge::Security s, const std::string &name) { m_securities[s] = name; }
    void addErrorCode(APIBridge::ErrorCode c, const std::string &description) { m_errors[c] = description; }

private:
    std::vector<std::shared_ptr<BenchmarkDescription>> m_descriptions;
    std::unordered_map<APIBridge::Scheme, std::string> m_schemes;
    std::unordered_map<APIBridge::Security, std::string> m_securities;
    std::unordered_map<APIBridge::ErrorCode, std::string> m_errors;
    APIBridge::ErrorCode m_last_error = HEBENCH_ECODE_SUCCESS;
    std::string m_last_error_text;
};

// supplied by the backend (R/src/engine/seal_engine.cpp:36,59)
BaseEngine *createEngine(const std::int8_t *p_buffer, std::uint64_t size);
void destroyEngine(BaseEngine *p);

}   // namespace cpp
}   // namespace hebench
