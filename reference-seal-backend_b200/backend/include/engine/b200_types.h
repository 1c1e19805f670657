// b200_types.h -- host/device data types of the B200 HEBench backend.
// Mirrors the role of R/include/engine/seal_types.h (error code, security id) and replaces the
// seal::Plaintext / seal::Ciphertext payloads that the reference moves between API calls.
#pragma once
#include <cstdint>
#include <memory>
#include <new>
#include <type_traits>
#include <utility>
#include <vector>

#include "b200he.h"

#define HEBSEAL_ECODE_SEAL_ERROR 2        // R/include/engine/seal_types.h:13 (any error below the HEBench layer)
#define HEBENCH_HE_SECURITY_128 0         // R/include/engine/seal_types.h:9

namespace sbe {

// host image of seal::Plaintext: CKKS = NTT form [L][N] with a scale; BFV = coefficients mod t [N]
struct Plaintext {
    std::vector<std::uint64_t> data;
    int L        = 0;      // RNS limbs (CKKS); 0 for BFV
    double scale = 1.0;
};
// allocator whose resize() leaves new words uninitialised: a result ciphertext is overwritten in full by store(), and
// zero-filling gigabytes first costs as much as the copy itself
template <class T> struct DefaultInitAllocator : std::allocator<T> {
    template <class U> struct rebind { typedef DefaultInitAllocator<U> other; };
    using std::allocator<T>::allocator;
    template <class U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void *>(p)) U; }
    template <class U, class... Args> void construct(U *p, Args &&...args) { ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...); }
};
// host image of seal::Ciphertext: uint64[size][L][N]
struct Ciphertext {
    std::vector<std::uint64_t, DefaultInitAllocator<std::uint64_t>> data;
    int size     = 0;
    int L        = 0;
    bool ntt     = false;
    double scale = 1.0;
};

class SEALContextWrapper;

// device-resident std::vector<seal::Ciphertext> on one GPU (RAII over b200he_batch)
class DeviceBatch
{
public:
    DeviceBatch(b200he_ctx *ctx);
    ~DeviceBatch();
    DeviceBatch(const DeviceBatch &) = delete;
    DeviceBatch &operator=(const DeviceBatch &) = delete;
    b200he_batch *get() const { return m_b; }
    b200he_ctx *ctx() const { return m_ctx; }
    std::uint64_t count() const { return b200he_batch_count(m_b); }
    int size() const { return b200he_batch_size(m_b); }
    int level() const { return b200he_batch_level(m_b); }
    double scale() const { return b200he_batch_scale(m_b); }

private:
    b200he_ctx *m_ctx;
    b200he_batch *m_b;
};
typedef std::shared_ptr<DeviceBatch> DeviceBatchPtr;

// A logical vector of ciphertexts placed on the GPUs of the engine: either every GPU holds a full replica,
// or GPU g holds the contiguous block [first[g], first[g+1]).  A RESULT vector whose shards are not contiguous blocks
// (a result grid split by columns) carries `ids`: the global index of every item of every shard.
struct ShardedCiphertexts {
    std::vector<DeviceBatchPtr> shard;                  // one per GPU (may hold 0 items)
    std::vector<std::uint64_t> first;                   // size n_gpus + 1
    std::vector<std::vector<std::uint64_t>> ids;        // optional, per GPU
    bool replicated = false;
    std::uint64_t n_total = 0;
    std::uint64_t total() const { return n_total; }
};

// The two operands of a result grid rows x cols (element-wise / dot-product sample grids, matrix-product cells): the
// result space is what gets partitioned, so the operand with more items is split into contiguous blocks -- one per GPU
// -- and the other is replicated ("slice of parameter 0 + all of parameter 1"; "row block of M0 + all of M1").
struct GridOperands {
    ShardedCiphertexts p[2];
    int split = 0;   // which operand is split
};
// this GPU's share of a grid operation: operand indices local to the GPU's batches, and the global result index of each unit
struct GridShare {
    std::vector<std::uint32_t> ai, bi;
    std::vector<std::uint64_t> result;
};

}   // namespace sbe
