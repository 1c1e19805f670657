// b200_context.cpp -- SEALContextWrapper of the B200 backend (see b200_context.h).
// Replaces R/src/engine/seal_context.cpp: key generation / encode / encrypt / decrypt stay on the host,
// every Evaluator call becomes a call on device batches through include/b200he.h.
#include "engine/b200_context.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <sstream>
#include <thread>

#include <sys/mman.h>

#include "../../hostfhe/hostfhe.h"

namespace sbe {

// ------------------------------------------------------------------ DeviceBatch
DeviceBatch::DeviceBatch(b200he_ctx *ctx) : m_ctx(ctx), m_b(nullptr)
{
    if (b200he_batch_create(ctx, &m_b))
        throw hebench::cpp::HEBenchError(std::string("DeviceBatch: ") + b200he_last_error(), HEBSEAL_ECODE_SEAL_ERROR);
}
DeviceBatch::~DeviceBatch() { b200he_batch_destroy(m_b); }

// ------------------------------------------------------------------ construction
SEALContextWrapper::Ptr SEALContextWrapper::createCKKSContext(std::size_t poly_modulus_degree, std::size_t num_coeff_moduli,
                                                              int coeff_moduli_bits, int scale_bits)
{
    Ptr p(new SEALContextWrapper());
    p->init(true, poly_modulus_degree, num_coeff_moduli, coeff_moduli_bits, scale_bits);
    return p;
}
SEALContextWrapper::Ptr SEALContextWrapper::createBFVContext(std::size_t poly_modulus_degree, std::size_t num_coeff_moduli,
                                                             int coeff_moduli_bits, int plaintext_modulus_bits)
{
    Ptr p(new SEALContextWrapper());
    p->init(false, poly_modulus_degree, num_coeff_moduli, coeff_moduli_bits, plaintext_modulus_bits);
    return p;
}

void SEALContextWrapper::check(int rc, const char *what) const
{
    if (rc) throw hebench::cpp::HEBenchError(std::string(what) + ": " + b200he_last_error(), HEBSEAL_ECODE_SEAL_ERROR);
}

void SEALContextWrapper::init(bool ckks, std::size_t N, std::size_t depth, int coeff_bits, int scale_or_plain_bits)
{
    if (depth < 1) throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("Multiplicative depth must be greater than 0."), HEBENCH_ECODE_INVALID_ARGS);
    if (N < 1024 || N > 32768 || (N & (N - 1)))
        throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("Polynomial modulus degree must be a power of 2 in [1024, 32768]."), HEBENCH_ECODE_INVALID_ARGS);
    m_ckks       = ckks;
    m_N          = N;
    m_K          = depth + 1;
    m_scale_bits = scale_or_plain_bits;
    if (const char *e = getenv("HEB_B200_TRACE_DIR")) m_trace_dir = e;
    if (const char *e = getenv("HEB_B200_PROFILE_JSON")) m_profile_path = e;
    if (const char *e = getenv("HEB_B200_PROFILE_SKIP")) m_profile_skip = atol(e);
    // key material and encryption randomness come from OS entropy (seed 0) unless HEB_B200_SEED fixes them: reproducible
    // keys and encryptions are for the parity tests and benchmarks only
    std::uint64_t seed = 0;
    if (const char *e = getenv("HEB_B200_SEED")) seed = strtoull(e, nullptr, 0);
    m_host = hfhe_create(ckks ? HFHE_CKKS : HFHE_BFV, N, depth, coeff_bits, scale_or_plain_bits, seed);
    if (!m_host) throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("Invalid encryption parameters."), HEBSEAL_ECODE_SEAL_ERROR);
    m_scale = ckks ? hfhe_scale(m_host) : 1.0;
    m_t     = ckks ? 0 : hfhe_plain_modulus(m_host);

    // one device context per GPU; keys replicated
    int n_gpus = 1;
    if (const char *e = getenv("HEB_B200_GPUS")) n_gpus = std::max(1, atoi(e));
    const std::uint64_t *keyr = hfhe_relin_key(m_host);
    const std::size_t n_gal   = hfhe_galois_count(m_host);
    for (int g = 0; g < n_gpus; ++g) {
        b200he_ctx *c = nullptr;
        check(b200he_ctx_create(ckks ? B200HE_CKKS : B200HE_BFV, (uint32_t)N, (uint32_t)m_K, hfhe_moduli(m_host), hfhe_psi(m_host), m_t, g, &c),
              "b200he_ctx_create");
        m_dev.push_back(c);
        check(b200he_set_relin_key(c, keyr), "b200he_set_relin_key");
        for (std::size_t i = 0; i < n_gal; ++i) {
            const uint32_t elt = hfhe_galois_elt(m_host, i);
            check(b200he_set_galois_key(c, elt, hfhe_galois_key(m_host, elt)), "b200he_set_galois_key");
        }
    }
}

SEALContextWrapper::~SEALContextWrapper()
{
    HostSlab::trim();
    m_mask_cache.clear();
    for (b200he_ctx *c : m_dev) b200he_ctx_destroy(c);
    if (m_host) hfhe_destroy(m_host);
}

std::vector<int> SEALContextWrapper::coeffModulusBits() const
{
    std::vector<int> bits;
    const std::uint64_t *q = hfhe_moduli(m_host);
    for (std::size_t i = 0; i < m_K; ++i) {
        int b = 0;
        for (std::uint64_t v = q[i]; v; v >>= 1) ++b;
        bits.push_back(b);
    }
    return bits;
}

// ------------------------------------------------------------------ host side
Plaintext SEALContextWrapper::encodeVector(const std::vector<double> &values) { return encodeVector(values, m_scale); }
Plaintext SEALContextWrapper::encodeVector(const std::vector<double> &values, double scale)
{
    if (values.size() > slotCount())
        throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("Not enough slots available to create packed plaintext"), HEBENCH_ECODE_INVALID_ARGS);
    Plaintext p;
    p.L     = (int)topLevel();
    p.scale = scale;
    p.data.resize(topLevel() * m_N);
    hfhe_ckks_encode(m_host, values.data(), values.size(), scale, p.data.data());
    return p;
}
Plaintext SEALContextWrapper::encodeVector(const std::vector<std::int64_t> &values)
{
    if (values.size() > slotCount())
        throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("Not enough slots available to create packed plaintext"), HEBENCH_ECODE_INVALID_ARGS);
    Plaintext p;
    p.data.resize(m_N);
    hfhe_bfv_encode(m_host, values.data(), values.size(), p.data.data());
    return p;
}
std::vector<double> SEALContextWrapper::decodeCKKS(const Plaintext &plain)
{
    std::vector<double> out(m_N / 2);
    hfhe_ckks_decode(m_host, plain.data.data(), (std::size_t)plain.L, plain.scale, out.data());
    return out;
}
std::vector<std::int64_t> SEALContextWrapper::decodeBFV(const Plaintext &plain)
{
    std::vector<std::int64_t> out(m_N);
    hfhe_bfv_decode(m_host, plain.data.data(), out.data());
    return out;
}
Ciphertext SEALContextWrapper::encrypt(const Plaintext &plain)
{
    Ciphertext c;
    c.size  = 2;
    c.L     = (int)topLevel();
    c.ntt   = m_ckks;
    c.scale = m_ckks ? plain.scale : 1.0;
    c.data.resize(2 * topLevel() * m_N);
    hfhe_encrypt(m_host, plain.data.data(), c.data.data());   // thread safe: one keystream per encryption
    return c;
}
// every encryption of the vector draws from its own keystream (index reserved up front, so a fixed seed gives the
// same bits whatever the thread schedule): the vector is encrypted on all host threads
std::vector<Ciphertext> SEALContextWrapper::encrypt(const std::vector<Plaintext> &plain)
{
    std::vector<Ciphertext> out(plain.size());
    const std::uint64_t first = hfhe_reserve_encryptions(m_host, plain.size());
#pragma omp parallel for schedule(dynamic, 1)
    for (long i = 0; i < (long)plain.size(); ++i) {
        Ciphertext &c = out[i];
        c.size        = 2;
        c.L           = (int)topLevel();
        c.ntt         = m_ckks;
        c.scale       = m_ckks ? plain[i].scale : 1.0;
        c.data.resize(2 * topLevel() * m_N);
        hfhe_encrypt_at(m_host, plain[i].data.data(), c.data.data(), first + (std::uint64_t)i);
    }
    return out;
}
Plaintext SEALContextWrapper::decrypt(const Ciphertext &cipher)
{
    Plaintext p;
    p.L     = m_ckks ? cipher.L : 0;
    p.scale = cipher.scale;
    p.data.resize(m_ckks ? (std::size_t)cipher.L * m_N : m_N);
    hfhe_decrypt(m_host, cipher.data.data(), (std::size_t)cipher.size, (std::size_t)cipher.L, p.data.data());
    return p;
}
std::vector<Plaintext> SEALContextWrapper::decrypt(const std::vector<Ciphertext> &cipher)
{
    std::vector<Plaintext> out(cipher.size());
#pragma omp parallel for schedule(dynamic, 1)
    for (long i = 0; i < (long)cipher.size(); ++i) out[i] = decrypt(cipher[i]);   // read-only on the host context
    return out;
}

void SEALContextWrapper::trace(const std::string &tag, const std::vector<Ciphertext> &v) const
{
    if (m_trace_dir.empty()) return;
    // HEB_B200_TRACE_PICK_<tag> = "i,j,k,...": write only those items (full-shape runs trace a few result cells and the
    // operands they depend on instead of gigabytes)
    std::vector<std::size_t> pick;
    if (const char *e = getenv(("HEB_B200_TRACE_PICK_" + tag).c_str())) {
        for (const char *p = e; *p;) {
            char *end            = nullptr;
            const std::size_t id = strtoull(p, &end, 10);
            if (end == p) break;
            if (id < v.size()) pick.push_back(id);
            p = *end ? end + 1 : end;
        }
    } else
        for (std::size_t i = 0; i < v.size(); ++i) pick.push_back(i);
    const std::string path = m_trace_dir + "/" + tag + ".bin";
    FILE *f                = fopen(path.c_str(), "wb");
    if (!f) throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("HEB_B200_TRACE_DIR: cannot write " + path), HEBENCH_ECODE_CRITICAL_ERROR);
    std::uint64_t hdr[8] = { 0x3143525430303242ull /* "B200TRC1" */, pick.size(), 0, 0, m_N, 0, 0, v.size() };
    if (!pick.empty()) {
        const Ciphertext &c0 = v[pick[0]];
        hdr[2]               = (std::uint64_t)c0.size;
        hdr[3]               = (std::uint64_t)c0.L;
        hdr[5]               = c0.ntt ? 1 : 0;
        std::memcpy(&hdr[6], &c0.scale, sizeof(double));
    }
    bool ok = fwrite(hdr, sizeof hdr, 1, f) == 1;
    for (std::size_t id : pick) {
        const Ciphertext &c = v[id];
        if (c.size != v[pick[0]].size || c.L != v[pick[0]].L) ok = false;   // one shape per file
        ok = ok && fwrite(c.data.data(), sizeof(std::uint64_t), c.data.size(), f) == c.data.size();
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("HEB_B200_TRACE_DIR: short write to " + path), HEBENCH_ECODE_CRITICAL_ERROR);
}

// ------------------------------------------------------------------ device side
std::vector<std::uint64_t> SEALContextWrapper::partition(std::uint64_t n) const
{
    const std::uint64_t g = m_dev.size();
    std::vector<std::uint64_t> first(g + 1, 0);
    for (std::uint64_t i = 0; i < g; ++i) first[i + 1] = first[i] + n / g + (i < n % g ? 1 : 0);
    return first;
}

DeviceBatchPtr SEALContextWrapper::upload(int g, const std::vector<Ciphertext> &src, std::size_t first, std::size_t n) const
{
    if (first + n > src.size()) throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("upload range out of bounds"), HEBENCH_ECODE_INVALID_ARGS);
    std::vector<std::size_t> items(n);
    std::iota(items.begin(), items.end(), first);
    return upload(g, src, items);
}
DeviceBatchPtr SEALContextWrapper::upload(int g, const std::vector<Ciphertext> &src, const std::vector<std::size_t> &items) const
{
    DeviceBatchPtr b    = newBatch(g);
    const std::size_t n = items.size();
    if (n == 0) {
        check(b200he_batch_resize(b->get(), 0, 2, (int)topLevel(), m_ckks, m_scale), "b200he_batch_resize");
        return b;
    }
    for (std::size_t i : items)
        if (i >= src.size()) throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("upload item out of bounds"), HEBENCH_ECODE_INVALID_ARGS);
    const Ciphertext &c0 = src[items[0]];
    check(b200he_batch_resize(b->get(), n, c0.size, c0.L, c0.ntt, c0.scale), "b200he_batch_resize");
    // one call for the whole list: the library gathers the separately allocated ciphertexts into pinned staging
    // buffers on a few host threads while the previous chunk crosses PCIe, and returns once every source has been read
    std::vector<const std::uint64_t *> ptrs(n);
    for (std::size_t i = 0; i < n; ++i) {
        const Ciphertext &c = src[items[i]];
        if (c.size != c0.size || c.L != c0.L || c.ntt != c0.ntt)
            throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("ciphertexts of one batch must share size, level and form"), HEBENCH_ECODE_INVALID_ARGS);
        ptrs[i] = c.data.data();
    }
    check(b200he_batch_upload_scattered(b->get(), 0, n, ptrs.data()), "b200he_batch_upload_scattered");
    return b;
}
DeviceBatchPtr SEALContextWrapper::upload(int g, const Ciphertext &src) const
{
    std::vector<Ciphertext> v(1, src);
    return upload(g, v, 0, 1);
}
DeviceBatchPtr SEALContextWrapper::uploadPlain(int g, const std::vector<Plaintext> &src) const
{
    DeviceBatchPtr b = newBatch(g);
    if (src.empty()) throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("empty plaintext batch"), HEBENCH_ECODE_INVALID_ARGS);
    check(b200he_batch_resize(b->get(), src.size(), 1, src[0].L, 1, src[0].scale), "b200he_batch_resize");
    std::vector<const std::uint64_t *> ptrs(src.size());
    for (std::size_t i = 0; i < src.size(); ++i) {
        if (src[i].L != src[0].L) throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("plaintexts of one batch must share the level"), HEBENCH_ECODE_INVALID_ARGS);
        ptrs[i] = src[i].data.data();
    }
    check(b200he_batch_upload_scattered(b->get(), 0, src.size(), ptrs.data()), "b200he_batch_upload_scattered");
    return b;
}
HostSlab::~HostSlab()
{
    if (m_map) munmap(m_map, m_map_bytes);
}
namespace {
// never destroyed: a background populate thread may hand an arena back while the process is already running its static
// destructors
std::mutex &g_slab_mtx                = *new std::mutex;
std::vector<HostSlab *> &g_slab_pool = *new std::vector<HostSlab *>;
std::size_t g_slab_pool_bytes         = 0;
std::size_t slabPoolCap()
{
    std::size_t mb = 32768;   // arenas kept for reuse, in total
    if (const char *e = std::getenv("HEB_B200_HOST_POOL_MB")) mb = (std::size_t)std::atoll(e);
    return mb << 20;
}
}   // namespace
void HostSlab::recycle(HostSlab *slab)
{
    {
        std::lock_guard<std::mutex> lock(g_slab_mtx);
        if (g_slab_pool_bytes + slab->m_bytes <= slabPoolCap()) {
            g_slab_pool.push_back(slab);
            g_slab_pool_bytes += slab->m_bytes;
            return;
        }
    }
    delete slab;
}
void HostSlab::trim()
{
    std::vector<HostSlab *> all;
    {
        std::lock_guard<std::mutex> lock(g_slab_mtx);
        all.swap(g_slab_pool);
        g_slab_pool_bytes = 0;
    }
    for (HostSlab *s : all) delete s;
}
std::shared_ptr<HostSlab> HostSlab::create(std::size_t bytes)
{
    const std::size_t huge = std::size_t(2) << 20;
    if (!bytes) return nullptr;
    if (const char *e = std::getenv("HEB_B200_HOST_SLAB"))
        if (std::atoi(e) == 0) return nullptr;
    {   // smallest pooled arena that fits without wasting more than half of itself
        std::lock_guard<std::mutex> lock(g_slab_mtx);
        std::size_t best = g_slab_pool.size();
        for (std::size_t i = 0; i < g_slab_pool.size(); ++i) {
            const std::size_t have = g_slab_pool[i]->m_bytes;
            if (have >= bytes && have / 2 <= bytes && (best == g_slab_pool.size() || have < g_slab_pool[best]->m_bytes)) best = i;
        }
        if (best != g_slab_pool.size()) {
            HostSlab *slab = g_slab_pool[best];
            g_slab_pool.erase(g_slab_pool.begin() + (std::ptrdiff_t)best);
            g_slab_pool_bytes -= slab->m_bytes;
            slab->m_used.store(0, std::memory_order_relaxed);
            return std::shared_ptr<HostSlab>(slab, &HostSlab::recycle);
        }
    }
    const std::size_t span = ((bytes + huge - 1) & ~(huge - 1)) + huge;
    void *map              = mmap(nullptr, span, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (map == MAP_FAILED) return nullptr;
    void *base = reinterpret_cast<void *>((reinterpret_cast<std::uintptr_t>(map) + huge - 1) & ~std::uintptr_t(huge - 1));
    madvise(base, span - huge, MADV_HUGEPAGE);   // advisory: 4 KB pages where transparent huge pages are off
    return std::shared_ptr<HostSlab>(new HostSlab(map, span, base, span - huge), &HostSlab::recycle);
}

void HostSlab::populate(const std::shared_ptr<HostSlab> &slab, int threads)
{
    if (!slab || threads < 1 || slab->m_resident.load(std::memory_order_relaxed) >= slab->m_bytes) return;
    // Pages are made resident by touching them (an atomic OR of 0: the byte keeps its value, so store() may already be
    // copying into the arena).  MADV_POPULATE_WRITE does the same in one call per stripe but holds the address space's lock
    // shared for the whole stripe, and everything that needs it exclusively -- thread creation, the driver's and malloc's
    // mmap calls inside operate() -- queues behind it: measured at 2 GPUs, operate() calls overlapping such a populate ran
    // 168 / 135 ms instead of 120.5 ms; page faults take the per-VMA lock and leave them alone (120.6 ms).
    const std::size_t stripe = std::size_t(4) << 20;
    for (int t = 0; t < threads; ++t) {
        std::weak_ptr<HostSlab> weak = slab;
        std::thread([weak, stripe]() {
            while (std::shared_ptr<HostSlab> s = weak.lock()) {   // held for one stripe at a time
                const std::size_t at = s->m_resident.fetch_add(stripe, std::memory_order_relaxed);
                if (at >= s->m_bytes) return;
                char *p               = static_cast<char *>(s->m_base) + at;
                const std::size_t len = std::min(stripe, s->m_bytes - at);
                for (std::size_t o = 0; o < len; o += 4096) __atomic_fetch_or(p + o, 0, __ATOMIC_RELAXED);
            }
        }).detach();
    }
}

// Only where it cannot cost the timed call anything measurable: results of at least HEB_B200_POPULATE_MIN_MB (default
// 1024) -- an operate() that produces a gigabyte of ciphertexts runs for tens of milliseconds at least, against the
// ~0.1 ms of starting the threads.  Measured on C3 (5.2 GB of results): operate 239.6 ms either way, store 291 -> 139 ms.
void SEALContextWrapper::prepareStore(ShardedCiphertexts &res) const
{
    if (res.replicated) return;
    int threads = 4;
    std::size_t min_mb = 1024;
    if (const char *e = std::getenv("HEB_B200_POPULATE_THREADS")) threads = std::atoi(e);
    if (const char *e = std::getenv("HEB_B200_POPULATE_MIN_MB")) min_mb = (std::size_t)std::atoll(e);
    if (threads < 1) return;
    std::vector<std::size_t> bytes(res.shard.size(), 0);
    std::size_t shards = 0, total = 0;
    for (std::size_t g = 0; g < res.shard.size(); ++g) {
        const DeviceBatchPtr &b = res.shard[g];
        if (!b || b->count() <= 1) continue;
        bytes[g] = b->count() * (((std::size_t)b->size() * b->level() * m_N * 8 + 63) & ~std::size_t(63));
        total += bytes[g];
        ++shards;
    }
    if (total < (min_mb << 20)) return;
    res.host.assign(res.shard.size(), nullptr);
    for (std::size_t g = 0; g < res.shard.size(); ++g) {
        if (!bytes[g]) continue;
        res.host[g] = HostSlab::create(bytes[g]);
        HostSlab::populate(res.host[g], std::max<int>(1, threads / (int)shards));
    }
}

std::vector<Ciphertext> SEALContextWrapper::download(const DeviceBatch &b, std::shared_ptr<HostSlab> slab) const
{
    std::vector<Ciphertext> out(b.count());
    const std::size_t words = (std::size_t)b.size() * b.level() * m_N;
    std::vector<std::uint64_t *> ptrs(out.size());
    const bool ntt = b200he_batch_ntt_form(b.get()) != 0;
    // all ciphertexts of the batch are carved from one huge-page arena (nothing is touched here: the first touch is the
    // staged download's copy, on its host threads)
    // (made by prepareStore while the GPUs were busy, or here)
    if (!slab && out.size() > 1) slab = HostSlab::create(out.size() * ((words * 8 + 63) & ~std::size_t(63)));
    for (std::size_t i = 0; i < out.size(); ++i) {
        Ciphertext &c = out[i];
        c.size        = b.size();
        c.L           = b.level();
        c.ntt         = ntt;
        c.scale       = b.scale();
        c.data        = Ciphertext::Words(HostAllocator<std::uint64_t>(slab));
        c.data.resize(words);
        ptrs[i] = c.data.data();
    }
    check(b200he_batch_download_scattered(b.get(), 0, out.size(), ptrs.data()), "b200he_batch_download_scattered");
    return out;
}
void SEALContextWrapper::syncAll() const
{
    for (b200he_ctx *c : m_dev) check(b200he_ctx_sync(c), "b200he_ctx_sync");
}

void SEALContextWrapper::beginOperate()
{
    m_profiling = !m_profile_path.empty() && m_operate_calls >= m_profile_skip;
    ++m_operate_calls;
    if (m_profiling)
        for (b200he_ctx *c : m_dev) check(b200he_profile_begin(c), "b200he_profile_begin");
}
void SEALContextWrapper::endOperate(std::uint64_t results)
{
    syncAll();
    if (!m_profiling) return;
    double ms[B200HE_KERN_COUNT] = {}, wi[B200HE_KERN_COUNT] = {}, wd[B200HE_KERN_COUNT] = {}, wb[B200HE_KERN_COUNT] = {}, ms_max[B200HE_KERN_COUNT] = {};
    std::uint64_t launches[B200HE_KERN_COUNT] = {};
    for (b200he_ctx *c : m_dev) {
        double m1[B200HE_KERN_COUNT], a[B200HE_KERN_COUNT], b[B200HE_KERN_COUNT], d[B200HE_KERN_COUNT];
        std::uint64_t l1[B200HE_KERN_COUNT];
        check(b200he_profile_end(c, m1, l1), "b200he_profile_end");
        check(b200he_profile_work(c, a, b, d), "b200he_profile_work");
        for (int i = 0; i < B200HE_KERN_COUNT; ++i) {
            ms[i] += m1[i];
            ms_max[i] = std::max(ms_max[i], m1[i]);
            launches[i] += l1[i];
            wi[i] += a[i];
            wd[i] += b[i];
            wb[i] += d[i];
        }
    }
    m_profiling = false;
    FILE *f = fopen(m_profile_path.c_str(), "a");
    if (!f) throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("HEB_B200_PROFILE_JSON: cannot write " + m_profile_path), HEBENCH_ECODE_CRITICAL_ERROR);
    fprintf(f, "{\"operate_call\": %ld, \"gpus\": %d, \"results\": %llu, \"kernels\": {", m_operate_calls - 1, gpuCount(), (unsigned long long)results);
    bool first = true;
    for (int i = 0; i < B200HE_KERN_COUNT; ++i) {
        if (!launches[i]) continue;
        fprintf(f, "%s\"%s\": {\"ms_sum_over_gpus\": %.6f, \"ms_max_over_gpus\": %.6f, \"launches\": %llu, \"bfly_int\": %.6e, \"bfly_fp64\": %.6e, \"bytes\": %.6e}",
                first ? "" : ", ", b200he_kernel_name(i), ms[i], ms_max[i], (unsigned long long)launches[i], wi[i], wd[i], wb[i]);
        first = false;
    }
    fprintf(f, "}}\n");
    fclose(f);
}

void SEALContextWrapper::forEachGpu(const std::function<void(int)> &fn) const
{
    const int n = gpuCount();
    if (n == 1) {
        fn(0);
        return;
    }
    std::vector<std::exception_ptr> err(n);
    std::vector<std::thread> th;
    for (int g = 0; g < n; ++g)
        th.emplace_back([&, g]() {
            try {
                fn(g);
            } catch (...) {
                err[g] = std::current_exception();
            }
        });
    for (std::thread &t : th) t.join();
    for (const std::exception_ptr &e : err)
        if (e) std::rethrow_exception(e);
}

GridOperands SEALContextWrapper::loadGrid(const std::vector<Ciphertext> &src0, std::size_t unit0, const std::vector<Ciphertext> &src1, std::size_t unit1,
                                          const std::function<std::vector<std::size_t>(std::size_t)> &items1) const
{
    GridOperands out;
    const std::vector<Ciphertext> *src[2] = { &src0, &src1 };
    const std::size_t unit[2] = { unit0 ? unit0 : 1, unit1 ? unit1 : 1 };
    const std::size_t n_items[2] = { src0.size() / unit[0], src1.size() / unit[1] };
    out.split = n_items[1] > n_items[0] ? 1 : 0;
    const int G = gpuCount();
    for (int p = 0; p < 2; ++p) {
        ShardedCiphertexts &s = out.p[p];
        s.n_total    = n_items[p];
        s.replicated = p != out.split;
        s.first      = s.replicated ? std::vector<std::uint64_t>(G + 1, n_items[p]) : partition(n_items[p]);
        if (s.replicated) s.first[0] = 0;
        s.shard.resize(G);
    }
    forEachGpu([&](int g) {
        for (int p = 0; p < 2; ++p) {
            ShardedCiphertexts &s     = out.p[p];
            const std::uint64_t first = s.replicated ? 0 : s.first[g], last = s.replicated ? n_items[p] : s.first[g + 1];
            std::vector<std::size_t> items;
            for (std::uint64_t it = first; it < last; ++it) {
                if (p == 1 && items1) {
                    const std::vector<std::size_t> v = items1(it);
                    items.insert(items.end(), v.begin(), v.end());
                } else
                    for (std::size_t u = 0; u < unit[p]; ++u) items.push_back(it * unit[p] + u);
            }
            s.shard[g] = upload(g, *src[p], items);
        }
    });
    return out;
}

GridShare SEALContextWrapper::gridShare(const GridOperands &in, int g, const std::uint64_t v0[2], const std::uint64_t b[2]) const
{
    GridShare sh;
    const int s = in.split;
    // the requested range of the split operand that lives on this GPU
    const std::uint64_t f = in.p[s].first[g], l = in.p[s].first[g + 1];
    const std::uint64_t lo = std::max<std::uint64_t>(f, v0[s]), hi = std::min<std::uint64_t>(l, v0[s] + b[s]);
    if (hi <= lo) return sh;
    const std::uint64_t nb = hi - lo, n = nb * b[1 - s];
    sh.ai.resize(n);
    sh.bi.resize(n);
    sh.result.resize(n);
    for (std::uint64_t k = 0; k < n; ++k) {
        std::uint64_t i, j;   // global item indices
        if (s == 0) {
            i = lo + k / b[1];
            j = v0[1] + k % b[1];
        } else {
            i = v0[0] + k / nb;
            j = lo + k % nb;
        }
        sh.ai[k]     = (std::uint32_t)(s == 0 ? i - f : i);
        sh.bi[k]     = (std::uint32_t)(s == 1 ? j - f : j);
        sh.result[k] = (i - v0[0]) * b[1] + (j - v0[1]);
    }
    return sh;
}

std::vector<Ciphertext> SEALContextWrapper::gather(const ShardedCiphertexts &src) const
{
    std::vector<Ciphertext> out(src.total());
    std::vector<std::vector<Ciphertext>> part(src.shard.size());
    if (src.replicated) {
        if (!src.shard.empty()) out = download(*src.shard[0]);
        return out;
    }
    forEachGpu([&](int g) {
        if ((std::size_t)g < src.shard.size() && src.shard[g] && src.shard[g]->count() > 0) part[g] = download(*src.shard[g], (std::size_t)g < src.host.size() ? src.host[g] : nullptr);
    });
    for (std::size_t g = 0; g < part.size(); ++g)
        for (std::size_t k = 0; k < part[g].size(); ++k) {
            const std::uint64_t id = src.ids.empty() ? src.first[g] + k : src.ids[g][k];
            if (id >= out.size()) throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("result index out of range"), HEBENCH_ECODE_CRITICAL_ERROR);
            out[id] = std::move(part[g][k]);
        }
    return out;
}

// R/src/engine/seal_context.cpp:255-263: the ciphertext with more limbs is switched down (CKKS: limbs dropped)
void SEALContextWrapper::matchLevel(DeviceBatch &a, DeviceBatch &b) const
{
    if (a.level() > b.level()) check(b200he_mod_drop(a.ctx(), a.get(), b.level(), a.get()), "b200he_mod_drop");
    else if (a.level() < b.level()) check(b200he_mod_drop(b.ctx(), b.get(), a.level(), b.get()), "b200he_mod_drop");
}
// R/src/engine/seal_context.cpp:289-347.  count == 0 (fresh encryption of zero in the reference) is rejected.
void SEALContextWrapper::accumulateBFV(DeviceBatch &cipher, std::size_t count) const
{
    check(b200he_accumulate(cipher.ctx(), cipher.get(), count), "b200he_accumulate");
}
void SEALContextWrapper::accumulateCKKS(DeviceBatch &cipher, std::size_t count) const
{
    if (count > slotCount()) count = slotCount();
    check(b200he_accumulate(cipher.ctx(), cipher.get(), count), "b200he_accumulate");
}

// masks e_i = encode(unit vector i of length `total`, scale()) switched down to `level`
// (R/src/engine/seal_context.cpp:382-388); deterministic, so they are cached in HBM per shape.
DeviceBatchPtr SEALContextWrapper::maskBatch(int g, std::size_t first_index, std::size_t n, std::size_t total, int level)
{
    std::ostringstream key;
    key << "mask:" << g << ':' << first_index << ':' << n << ':' << total << ':' << level;
    {
        std::lock_guard<std::mutex> lock(m_cache_mtx);
        auto it = m_mask_cache.find(key.str());
        if (it != m_mask_cache.end()) return it->second;
    }
    std::vector<Plaintext> masks(n);
#pragma omp parallel for
    for (long i = 0; i < (long)n; ++i) {
        std::vector<double> identity(total, 0.0);
        identity[first_index + i] = 1.0;
        Plaintext p = encodeVector(identity, m_scale);
        p.data.resize((std::size_t)level * m_N);   // mod_switch_to_inplace(plain): drop the trailing limbs
        p.L = level;
        masks[i] = std::move(p);
    }
    DeviceBatchPtr b = uploadPlain(g, masks);
    std::lock_guard<std::mutex> lock(m_cache_mtx);
    m_mask_cache[key.str()] = b;
    return b;
}

// coefficient plaintext `index` of a Horner evaluation switched down to `level` (mod_switch_to_inplace(plain): the
// trailing limbs dropped, R/src/engine/seal_context.cpp:451), resident on GPU g; deterministic, so cached
DeviceBatchPtr SEALContextWrapper::coeffBatch(int g, const std::vector<Plaintext> &plain_coefficients, std::size_t index, int level)
{
    std::ostringstream key;
    key << "coeff:" << g << ':' << index << ':' << level << ':' << (const void *)plain_coefficients.data();
    {
        std::lock_guard<std::mutex> lock(m_cache_mtx);
        auto it = m_mask_cache.find(key.str());
        if (it != m_mask_cache.end()) return it->second;
    }
    Plaintext p = plain_coefficients.at(index);
    p.data.resize((std::size_t)level * m_N);
    p.L              = level;
    DeviceBatchPtr b = uploadPlain(g, std::vector<Plaintext>(1, p));
    std::lock_guard<std::mutex> lock(m_cache_mtx);
    m_mask_cache[key.str()] = b;
    return b;
}

// R/src/engine/seal_context.cpp:349-415 on one GPU's shard of the samples
DeviceBatchPtr SEALContextWrapper::collapseCKKS(DeviceBatch &ciphers, std::size_t first_index, std::size_t total, const Ciphertext *encrypted_zero)
{
    b200he_ctx *c = ciphers.ctx();
    int g = 0;
    for (std::size_t i = 0; i < m_dev.size(); ++i)
        if (m_dev[i] == c) g = (int)i;
    const std::size_t n = ciphers.count();
    DeviceBatchPtr result = newBatch(g);
    if (n > 0) {
        std::vector<int32_t> steps(n);
        for (std::size_t i = 0; i < n; ++i) steps[i] = -(int32_t)(first_index + i);
        check(b200he_rotate_each(c, ciphers.get(), steps.data(), ciphers.get()), "b200he_rotate_each");
        DeviceBatchPtr masks = maskBatch(g, first_index, n, total, ciphers.level());
        check(b200he_multiply_plain(c, ciphers.get(), masks->get(), nullptr, ciphers.get()), "b200he_multiply_plain");
        // relinearize_inplace (size 2: no-op, as in the reference) + rescale_to_next_inplace
        check(b200he_relinearize_rescale(c, ciphers.get(), ciphers.get()), "b200he_relinearize_rescale");
        check(b200he_batch_set_scale(ciphers.get(), m_scale), "b200he_batch_set_scale");
        check(b200he_sum(c, ciphers.get(), result->get()), "b200he_sum");
    }
    if (encrypted_zero) {
        // retval = Enc(0) at the top level, switched down to the summands' level, scale forced (:360-361, :397-400).  The
        // encryption itself was drawn ahead of this call (LogRegHornerBenchmark::freshEncryptions); the upload only stages.
        trace("collapse_zero", *encrypted_zero);
        DeviceBatchPtr z = upload(g, *encrypted_zero);
        if (n > 0) {
            check(b200he_mod_drop(c, z->get(), result->level(), z->get()), "b200he_mod_drop");
            check(b200he_batch_set_scale(z->get(), result->scale()), "b200he_batch_set_scale");
            check(b200he_add(c, z->get(), nullptr, result->get(), nullptr, 1, result->get()), "b200he_add");
        } else
            result = z;
    }
    return result;
}

// R/src/engine/seal_context.cpp:417-458 (Horner), cipher_input and the result are single-ciphertext batches
DeviceBatchPtr SEALContextWrapper::evaluatePolynomial(DeviceBatch &cipher_input, const std::vector<Plaintext> &plain_coefficients, const Ciphertext &seed)
{
    if (plain_coefficients.empty())
        throw hebench::cpp::HEBenchError(HEBERROR_MSG_CLASS("Polynomial must have, at least, 1 coefficient."), HEBENCH_ECODE_INVALID_ARGS);
    b200he_ctx *c = cipher_input.ctx();
    int g = 0;
    for (std::size_t i = 0; i < m_dev.size(); ++i)
        if (m_dev[i] == c) g = (int)i;
    trace("horner_seed", seed);
    DeviceBatchPtr retval = upload(g, seed);   // Enc(a_d), drawn ahead of the call
    for (std::size_t k = plain_coefficients.size() - 1; k-- > 0;) {
        matchLevel(cipher_input, *retval);
        check(b200he_multiply(c, retval->get(), nullptr, cipher_input.get(), nullptr, 1, retval->get()), "b200he_multiply");
        check(b200he_relinearize_rescale(c, retval->get(), retval->get()), "b200he_relinearize_rescale");
        // mod_switch_to_inplace(plain, retval.parms_id()); scale forced (:451-452)
        DeviceBatchPtr dp = coeffBatch(g, plain_coefficients, k, retval->level());
        check(b200he_batch_set_scale(retval->get(), plain_coefficients[k].scale), "b200he_batch_set_scale");
        check(b200he_add_plain(c, retval->get(), dp->get(), nullptr, retval->get()), "b200he_add_plain");
    }
    return retval;
}

}   // namespace sbe
