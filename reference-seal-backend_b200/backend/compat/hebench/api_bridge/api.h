// api.h -- the C ABI a HEBench backend exports and test_harness binds with dlsym
// (subset, see types.h).  Every function returns an ErrorCode (0 = success) unless noted.
#ifndef HEBENCH_COMPAT_API_H
#define HEBENCH_COMPAT_API_H
#include "types.h"

namespace hebench {
namespace APIBridge {
extern "C" {

ErrorCode initEngine(Handle *h_engine, const int8_t *p_buffer, uint64_t size);
ErrorCode destroyHandle(Handle h);
ErrorCode subscribeBenchmarksCount(Handle h_engine, uint64_t *p_count);
ErrorCode subscribeBenchmarks(Handle h_engine, Handle *p_h_bench_descs, uint64_t count);
ErrorCode getWorkloadParamsDetails(Handle h_engine, Handle h_bench_desc, uint64_t *p_param_count, uint64_t *p_default_count);
ErrorCode describeBenchmark(Handle h_engine, Handle h_bench_desc, BenchmarkDescriptor *p_bench_desc,
                            WorkloadParams *p_default_params, uint64_t default_count);
ErrorCode createBenchmark(Handle h_engine, Handle h_bench_desc, const WorkloadParams *p_params, Handle *h_benchmark);
ErrorCode initBenchmark(Handle h_benchmark, const BenchmarkDescriptor *p_concrete_desc);
ErrorCode encode(Handle h_benchmark, const DataPackCollection *p_parameters, Handle *h_plaintext);
ErrorCode decode(Handle h_benchmark, Handle h_plaintext, DataPackCollection *p_native);
ErrorCode encrypt(Handle h_benchmark, Handle h_plaintext, Handle *h_ciphertext);
ErrorCode decrypt(Handle h_benchmark, Handle h_ciphertext, Handle *h_plaintext);
ErrorCode load(Handle h_benchmark, const Handle *h_local_packed_params, uint64_t local_count, Handle *h_remote);
ErrorCode store(Handle h_benchmark, Handle h_remote, Handle *h_local_packed_params, uint64_t local_count);
ErrorCode operate(Handle h_benchmark, Handle h_remote_packed_params, const ParameterIndexer *p_param_indexers,
                  uint64_t indexers_count, Handle *h_remote_output);
// the following return the number of bytes (including the terminator) the text needs; text is
// copied into p_buffer when it is not NULL and size is large enough
uint64_t getSchemeName(Handle h_engine, Scheme s, char *p_name, uint64_t size);
uint64_t getSchemeSecurityName(Handle h_engine, Scheme s, Security sec, char *p_name, uint64_t size);
uint64_t getBenchmarkDescriptionEx(Handle h_engine, Handle h_bench_desc, const WorkloadParams *p_w_params, char *p_description, uint64_t size);
uint64_t getErrorDescription(Handle h_engine, ErrorCode code, char *p_description, uint64_t size);
uint64_t getLastErrorDescription(Handle h_engine, char *p_description, uint64_t size);
}
}   // namespace APIBridge
}   // namespace hebench
#endif
