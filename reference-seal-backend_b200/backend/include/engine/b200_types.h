// b200_types.h -- host/device data types of the B200 HEBench backend.
// Mirrors the role of R/include/engine/seal_types.h (error code, security id) and replaces the
// seal::Plaintext / seal::Ciphertext payloads that the reference moves between API calls.
#pragma once
#include <atomic>
#include <cstddef>
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
// Arena behind the result ciphertexts of one store(): a single anonymous mapping, 2 MB aligned and advised for
// transparent huge pages, carved into ciphertexts by bump allocation.  Measured reason (tools/host_path_probe.py,
// profiles/r2o_host_path.json): store() into 10^4 separately malloc'ed 512 KB vectors is bound by first-touch page
// faults, not by PCIe.  The mapping is returned to the system when the last ciphertext carved from it is destroyed.
class HostSlab
{
public:
    // An arena of at least `bytes`: a recycled one when the pool holds a fitting mapping (already resident, no system call),
    // a new mapping otherwise; nullptr when the mapping cannot be made.  The last owner hands the arena back to the pool
    // instead of unmapping it: HEBench calls operate() again and again on the same shapes, and unmapping gigabytes (or
    // touching them for the first time) next to a running operate() costs that call tens of milliseconds of address-space
    // lock contention (measured at 2 GPUs: C3 operate() 127 -> 163 ms before the pool existed).
    static std::shared_ptr<HostSlab> create(std::size_t bytes);
    // Make the arena's pages resident from a few background threads without changing their contents (so it may run
    // while store() is already copying into the arena); continues where an earlier call stopped.  The threads stop when
    // the last owner lets go of the arena.
    static void populate(const std::shared_ptr<HostSlab> &slab, int threads);
    static void trim();   // unmap everything the pool holds (benchmark teardown)
    std::size_t bytes() const { return m_bytes; }
    ~HostSlab();
    HostSlab(const HostSlab &) = delete;
    HostSlab &operator=(const HostSlab &) = delete;
    void *take(std::size_t bytes)   // 64-byte aligned; nullptr when the arena is exhausted
    {
        bytes                  = (bytes + 63) & ~std::size_t(63);
        const std::size_t at   = m_used.fetch_add(bytes, std::memory_order_relaxed);
        return at + bytes <= m_bytes ? static_cast<char *>(m_base) + at : nullptr;
    }
    bool owns(const void *p) const { return p >= m_base && p < static_cast<const char *>(m_base) + m_bytes; }

private:
    HostSlab(void *map, std::size_t map_bytes, void *base, std::size_t bytes) : m_map(map), m_map_bytes(map_bytes), m_base(base), m_bytes(bytes) {}
    static void recycle(HostSlab *slab);
    void *m_map;
    std::size_t m_map_bytes;
    void *m_base;
    std::size_t m_bytes;
    std::atomic<std::size_t> m_used{ 0 };
    std::atomic<std::size_t> m_resident{ 0 };   // populate()'s progress: the next stripe to touch
};
// Allocator of Ciphertext::data.  (1) resize() leaves new words uninitialised: a result ciphertext is overwritten in
// full by store(), and zero-filling gigabytes first costs as much as the copy itself.  (2) Optionally bound to a
// HostSlab: allocations come from the arena while it has room, from the heap otherwise; copies of a ciphertext own heap
// memory, moves keep the arena alive.
template <class T> struct HostAllocator {
    typedef T value_type;
    typedef std::true_type propagate_on_container_move_assignment;
    typedef std::true_type propagate_on_container_swap;
    typedef std::false_type propagate_on_container_copy_assignment;
    template <class U> struct rebind { typedef HostAllocator<U> other; };
    std::shared_ptr<HostSlab> slab;
    HostAllocator() noexcept {}
    explicit HostAllocator(std::shared_ptr<HostSlab> s) noexcept : slab(std::move(s)) {}
    template <class U> HostAllocator(const HostAllocator<U> &o) noexcept : slab(o.slab) {}
    HostAllocator select_on_container_copy_construction() const { return HostAllocator(); }
    T *allocate(std::size_t n)
    {
        if (slab)
            if (void *p = slab->take(n * sizeof(T))) return static_cast<T *>(p);
        return static_cast<T *>(::operator new(n * sizeof(T)));
    }
    void deallocate(T *p, std::size_t) noexcept
    {
        if (slab && slab->owns(p)) return;   // the arena goes as a whole
        ::operator delete(p);
    }
    template <class U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void *>(p)) U; }
    template <class U, class... Args> void construct(U *p, Args &&...args) { ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...); }
    template <class U> bool operator==(const HostAllocator<U> &o) const noexcept { return slab == o.slab; }
    template <class U> bool operator!=(const HostAllocator<U> &o) const noexcept { return slab != o.slab; }
};
// host image of seal::Ciphertext: uint64[size][L][N]
struct Ciphertext {
    typedef std::vector<std::uint64_t, HostAllocator<std::uint64_t>> Words;
    Words data;
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
    // results only, optional: per GPU, the host arena store() will copy the shard into -- made, and its pages touched in
    // the background, while the GPUs are still computing (SEALContextWrapper::prepareStore)
    std::vector<std::shared_ptr<HostSlab>> host;
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
