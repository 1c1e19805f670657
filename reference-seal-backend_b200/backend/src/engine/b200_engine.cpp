// b200_engine.cpp -- engine creation and the 20 benchmark descriptors, in the reference's order
// (R/src/engine/seal_engine.cpp:36-152).  An error code for failures below the HEBench layer
// (CUDA library) takes the place of the reference's "SEAL error".
#include "engine/b200_engine.h"

#include <sstream>

#include "benchmarks/b200_benchmarks.h"
#include "engine/b200_types.h"

namespace hebench {
namespace cpp {

BaseEngine *createEngine(const std::int8_t *p_buffer, std::uint64_t size)
{
    (void)p_buffer;   // the backend needs no extra initialisation data
    (void)size;
    if (HEBENCH_API_VERSION_MAJOR != HEBENCH_API_VERSION_NEEDED_MAJOR || HEBENCH_API_VERSION_MINOR != HEBENCH_API_VERSION_NEEDED_MINOR
        || HEBENCH_API_VERSION_REVISION < HEBENCH_API_VERSION_NEEDED_REVISION) {
        std::stringstream ss;
        ss << "Critical: Invalid HEBench API version detected. Required: " << HEBENCH_API_VERSION_NEEDED_MAJOR << "."
           << HEBENCH_API_VERSION_NEEDED_MINOR << "." << HEBENCH_API_VERSION_NEEDED_REVISION << ", but " << HEBENCH_API_VERSION_MAJOR << "."
           << HEBENCH_API_VERSION_MINOR << "." << HEBENCH_API_VERSION_REVISION << " received.";
        throw HEBenchError(HEBERROR_MSG(ss.str()), HEBENCH_ECODE_CRITICAL_ERROR);
    }
    return SEALEngine::create();
}

void destroyEngine(BaseEngine *p) { SEALEngine::destroy(dynamic_cast<SEALEngine *>(p)); }

}   // namespace cpp
}   // namespace hebench

SEALEngine *SEALEngine::create()
{
    SEALEngine *p = new SEALEngine();
    p->init();
    return p;
}
void SEALEngine::destroy(SEALEngine *p) { delete p; }
SEALEngine::SEALEngine() {}
SEALEngine::~SEALEngine() {}

void SEALEngine::init()
{
    using hebench::APIBridge::Category;
    using hebench::APIBridge::Workload;
    addErrorCode(HEBSEAL_ECODE_SEAL_ERROR, "SEAL error");   // same code and text as the reference: evaluator-level failure
    addSchemeName(HEBENCH_HE_SCHEME_CKKS, "CKKS");
    addSchemeName(HEBENCH_HE_SCHEME_BFV, "BFV");
    addSecurityName(HEBENCH_HE_SECURITY_128, "128 bits");

    for (Workload op : { Workload::EltwiseAdd, Workload::EltwiseMultiply })
        for (Category cat : { Category::Latency, Category::Offline }) {
            addBenchmarkDescription(std::make_shared<sbe::bfv::ElementWiseBenchmarkDescription>(cat, op));
            addBenchmarkDescription(std::make_shared<sbe::ckks::ElementWiseBenchmarkDescription>(cat, op));
        }
    for (Category cat : { Category::Latency, Category::Offline }) {
        addBenchmarkDescription(std::make_shared<sbe::bfv::DotProductBenchmarkDescription>(cat));
        addBenchmarkDescription(std::make_shared<sbe::ckks::DotProductBenchmarkDescription>(cat));
    }
    addBenchmarkDescription(std::make_shared<sbe::bfv::MatMultCipherBatchAxisBenchmarkDescription>());
    addBenchmarkDescription(std::make_shared<sbe::ckks::MatMultCipherBatchAxisBenchmarkDescription>());
    addBenchmarkDescription(std::make_shared<sbe::bfv::MatMultValBenchmarkDescription>());
    addBenchmarkDescription(std::make_shared<sbe::ckks::MatMultValBenchmarkDescription>());
    addBenchmarkDescription(std::make_shared<sbe::bfv::MatMultRowBenchmarkDescription>());
    addBenchmarkDescription(std::make_shared<sbe::ckks::MatMultRowBenchmarkDescription>());
    addBenchmarkDescription(std::make_shared<sbe::ckks::LogRegHornerBenchmarkDescription>(Category::Latency));
    addBenchmarkDescription(std::make_shared<sbe::ckks::LogRegHornerBenchmarkDescription>(Category::Offline, 0));
    // beyond the reference's 20 (kept last so the reference's indices are unchanged): the api-bridge's degree-5 and
    // degree-7 logistic-regression workloads (BASELINE.json configs[4]), same Horner benchmark class
    for (hebench::APIBridge::Workload w : { hebench::APIBridge::Workload::LogisticRegression_PolyD5, hebench::APIBridge::Workload::LogisticRegression_PolyD7 }) {
        addBenchmarkDescription(std::make_shared<sbe::ckks::LogRegHornerBenchmarkDescription>(Category::Latency, 0, w));
        addBenchmarkDescription(std::make_shared<sbe::ckks::LogRegHornerBenchmarkDescription>(Category::Offline, 0, w));
    }
}
