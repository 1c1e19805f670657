// types.h -- header-compatible SUBSET of the HEBench API Bridge (hebench/frontend v0.9.0-beta,
// api-bridge v0.8.0-beta; /root/reference/cmake/third-party/API_BRIDGE.version) written for this
// repository because the upstream headers are fetched from the network by the reference's CMake and
// are not available offline.  Only what the reference backend uses is declared
// (census in SURVEY.md §8b).  When the real api-bridge is installed, build with
// -DAPI_BRIDGE_INSTALL_DIR=... and this directory drops out of the include path.
#ifndef HEBENCH_COMPAT_TYPES_H
#define HEBENCH_COMPAT_TYPES_H
#include <stdint.h>

#define HEBENCH_API_VERSION_MAJOR 0
#define HEBENCH_API_VERSION_MINOR 8
#define HEBENCH_API_VERSION_REVISION 0
#define HEBENCH_API_VERSION_BUILD "beta-compat"

#define HEBENCH_MAX_OP_PARAMS 32
#define HEBENCH_MAX_BUFFER_SIZE 256
#define HEBENCH_MAX_CATEGORY_PARAMS (HEBENCH_MAX_OP_PARAMS * 2)

#define HEBENCH_ECODE_SUCCESS 0
#define HEBENCH_ECODE_CRITICAL_ERROR 0x7fffffff
#define HEBENCH_ECODE_INVALID_ARGS 0x7ffffffe

#define HEBENCH_HE_SCHEME_PLAIN 0
#define HEBENCH_HE_SCHEME_CKKS 1
#define HEBENCH_HE_SCHEME_BFV 2
#define HEBENCH_HE_SCHEME_BGV 3
#define HEBENCH_HE_PARAM_FLAGS_ALL_PLAIN 0x0
#define HEBENCH_HE_PARAM_FLAGS_ALL_CIPHER 0xffffffff

namespace hebench {
namespace APIBridge {

typedef int32_t ErrorCode;
typedef int32_t Scheme;
typedef int32_t Security;

struct _FlexibleData {
    void *p;
    uint64_t size;
    int64_t tag;
};
typedef _FlexibleData Handle;
typedef _FlexibleData DataBuffer;
typedef _FlexibleData NativeDataBuffer;

struct DataPack {
    DataBuffer *p_buffers;      // one buffer per sample
    uint64_t buffer_count;
    uint64_t param_position;    // which operation parameter these samples feed
};
struct DataPackCollection {
    DataPack *p_data_packs;
    uint64_t pack_count;
};
struct ParameterIndexer {
    uint64_t value_index;
    uint64_t batch_size;
};

enum Workload {
    MatrixMultiply = 1,
    EltwiseMultiply,
    EltwiseAdd,
    DotProduct,
    LogisticRegression,
    LogisticRegression_PolyD3,
    LogisticRegression_PolyD5,
    LogisticRegression_PolyD7,
    SimpleSetIntersection,
    Generic
};
enum DataType { Int32 = 1, Int64, Float32, Float64 };
enum Category { Latency = 1, Offline };
enum WorkloadParamType { Int64_wp = 1, UInt64_wp, Float64_wp };

struct WorkloadParam {
    WorkloadParamType data_type;
    char name[HEBENCH_MAX_BUFFER_SIZE];
    union {
        int64_t i_param;
        uint64_t u_param;
        double f_param;
    };
};
struct WorkloadParams {
    WorkloadParam *params;
    uint64_t count;
};

struct CategoryParams {
    uint64_t min_test_time_ms;
    union {
        uint64_t reserved[HEBENCH_MAX_CATEGORY_PARAMS];
        struct {
            uint64_t warmup_iterations_count;
        } latency;
        struct {
            uint64_t data_count[HEBENCH_MAX_OP_PARAMS];
        } offline;
    };
};

struct BenchmarkDescriptor {
    Workload workload;
    DataType data_type;
    Category category;
    CategoryParams cat_params;
    uint32_t cipher_param_mask;
    Scheme scheme;
    Security security;
    int64_t other;
};

// The layout the plugin's descriptors are exchanged in.  These are the upstream sizes AS RECALLED (the upstream header
// cannot be fetched offline): one 64-bit minimum test time, then a union of 2 * HEBENCH_MAX_OP_PARAMS reserved words.  A
// build against the real api-bridge (CMakeLists.txt, -DAPI_BRIDGE_INSTALL_DIR) does not include this file at all; these
// asserts make a silent drift of the restated structs impossible.
static_assert(sizeof(_FlexibleData) == 24, "Handle / DataBuffer: pointer, size, tag");
static_assert(sizeof(CategoryParams) == 8 + 8 * HEBENCH_MAX_CATEGORY_PARAMS, "CategoryParams: min_test_time_ms + reserved[]");
static_assert(sizeof(WorkloadParam) == 8 + HEBENCH_MAX_BUFFER_SIZE + 8, "WorkloadParam: type (padded), name, value");
static_assert(sizeof(BenchmarkDescriptor) == 16 + sizeof(CategoryParams) + 24, "BenchmarkDescriptor layout");

}   // namespace APIBridge
}   // namespace hebench
#endif
