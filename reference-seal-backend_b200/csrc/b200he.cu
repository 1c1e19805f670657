// b200he.cu -- host side of libb200he.so: the C ABI of include/b200he.h on top of the sm_100a
// kernels in kernels.cuh.  One context = one GPU + one stream; every evaluator entry enqueues
// kernels and returns, b200he_ctx_sync / b200he_batch_download wait.
//
// There is no CPU path in this file: without a CUDA device b200he_ctx_create fails.  (The same
// source is compiled as host C++ with -DB200HE_EMU by tests/emu/ only, to check the kernels' index
// arithmetic against the oracle in a GPU-less container; that build is test infrastructure and the
// Python binding of the product refuses to load it.)
#include "../../include/b200he.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "hostmath.h"
#include "launch.h"
#include "kernels_ew.cuh"
#include "behz.cuh"

using namespace b200he;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return -1;
}
#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return fail("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define TRY(call)            \
    do {                     \
        int rc_ = (call);    \
        if (rc_) return rc_; \
    } while (0)

extern "C" const char *b200he_last_error(void) { return g_err.c_str(); }
extern "C" const char *b200he_version(void)
{
#ifdef B200HE_EMU
    return "b200he EMU (test infrastructure, host C++)";
#else
    return "b200he sm_100a " __DATE__;
#endif
}

// ------------------------------------------------------------------------------------ device pool
// Stream-ordered reuse: every allocation is used on the context's single stream only, so a block
// handed back to the pool may be given out again immediately.
struct DevPool {
    std::multimap<size_t, void *> free_blocks;
    std::map<void *, size_t> sizes;
    size_t total = 0;
    void *get(size_t bytes)
    {
        bytes = (bytes + 511) / 512 * 512;
        if (!bytes) bytes = 512;
        auto it = free_blocks.lower_bound(bytes);
        if (it != free_blocks.end() && it->first <= bytes + bytes / 4) {
            void *p = it->second;
            free_blocks.erase(it);
            return p;
        }
        void *p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            release();
            if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
        }
        sizes[p] = bytes;
        total += bytes;
        return p;
    }
    void put(void *p)
    {
        if (p) free_blocks.emplace(sizes[p], p);
    }
    void release()
    {
        for (auto &kv : free_blocks) {
            total -= kv.first;
            sizes.erase(kv.second);
            cudaFree(kv.second);
        }
        free_blocks.clear();
    }
};

// ------------------------------------------------------------------------------------ context
struct ProfRec {
    int cls;
    cudaEvent_t a, b;
};

struct b200he_ctx {
    int scheme = 0, device = 0;
    u32 N = 0, K = 0;
    int logn = 0, lognl = 0, c = 0;   // N = 2^logn; CTA-local transform 2^lognl; limb split 2^c ways
    u64 t = 0;
    int M = 0;                        // moduli in the tables: K chain primes (+ BEHZ auxiliary primes)
    int n_sm = 148;                   // multiprocessors of the device (grid of the persistent kernels)
    std::vector<Mod> mods;
    Tables T{};
    void *d_tables = nullptr;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    DevPool pool;
    u64 workspace = u64(2) << 30;
    u64 *relin = nullptr;
    std::map<u32, u64 *> gal;
    struct GalTab {
        u32 *d = nullptr;                           // device copy of the NTT-form permutation table [N]
        unsigned char chunk[4] = { 0, 0, 0, 0 };    // source chunk (2^lognl coefficients) of every output chunk
    };
    std::map<u32, GalTab> galtab;
    std::set<b200he_batch *> batches;   // live batches: orphaned (ctx = NULL, no storage) when the context goes first
    std::mutex batches_mtx;
    uint64_t launches = 0;
    bool prof = false;
    std::vector<ProfRec> recs;
    // algorithmic work of the launches recorded while profiling (b200he_profile_work): butterfly-equivalents on the
    // integer pipe / on the FP64 pipe, and bytes that have to cross HBM, per kernel class
    double work_int[B200HE_KERN_COUNT] = {}, work_dp[B200HE_KERN_COUNT] = {}, work_bytes[B200HE_KERN_COUNT] = {};
    Behz behz{};   // BFV only
    int nBsk = 0;
    // pinned staging of b200he_batch_{upload,download}_scattered: two buffers, each guarded by the event of the last
    // copy that used it
    void *stage_buf[2] = { nullptr, nullptr };
    cudaEvent_t stage_ev[2] = { nullptr, nullptr };
    cudaStream_t aux = nullptr;                     // side stream of launch pairs that may overlap (launch.h Geo)
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    size_t stage_bytes = 0;
};

struct b200he_batch {
    b200he_ctx *ctx = nullptr;
    u64 *d = nullptr;
    size_t cap_words = 0;
    uint64_t count = 0;
    int size = 0, L = 0, ntt = 0;
    double scale = 1.0;
    u32 n_poly = 0;   // N of the owning context (kept so that the shape queries survive the context)
    size_t ct_words() const { return (size_t)size * L * n_poly; }
};
// every entry that reaches a context THROUGH a batch goes through this: a batch that outlived its context is an error
#define LIVE(b, what)                                                                              \
    do {                                                                                           \
        if (!(b)->ctx) return fail("%s: the batch's context has been destroyed", what);            \
    } while (0)

static inline void prof_pre(b200he_ctx *c, int cls)
{
    c->launches++;
    if (c->prof) {
        ProfRec r{ cls, nullptr, nullptr };
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, c->stream);
        c->recs.push_back(r);
    }
}
// Work accounting (only while profiling).  Units: one "butterfly-equivalent" = one 64-bit modular multiply by a constant
// with its add/sub (a Shoup butterfly: the unit tools/imad_peak.cu and tools/fp64_peak.cu measure the pipes in); a
// 64 x 64 -> 128 multiply-accumulate of k_tensor_mac counts 0.5, a Barrett data x data product 1.5.
static inline double bfly_per_limb(const b200he_ctx *c) { return 0.5 * c->N * c->logn; }
static inline void work_add(b200he_ctx *c, int cls, int mod_id, double bf, double bytes)
{
    if (!c->prof) return;
    (c->mods[mod_id].dp ? c->work_dp : c->work_int)[cls] += bf;
    c->work_bytes[cls] += bytes;
}
static inline void prof_post(b200he_ctx *c)
{
    if (c->prof) cudaEventRecord(c->recs.back().b, c->stream);
}
#define LAUNCH(ctx, cls, kernel, grid, block, smem, ...)                          \
    do {                                                                          \
        prof_pre(ctx, cls);                                                       \
        B200HE_LAUNCH(kernel, grid, block, smem, (ctx)->stream, __VA_ARGS__);     \
        prof_post(ctx);                                                           \
    } while (0)
#define LAUNCH_CHECK() CK(cudaGetLastError())
// the NTT-bearing kernel families live in their own translation units (launch.h); PROF wraps a launcher call with the
// per-class launch counter / event pair
#define PROF(ctx, cls, CALL) \
    do {                     \
        prof_pre(ctx, cls);  \
        CALL;                \
        prof_post(ctx);      \
    } while (0)
static inline Geo geo(const b200he_ctx *c) { return Geo{ c->lognl, c->c, c->stream, c->aux, c->aux_fork, c->aux_join }; }

static inline unsigned blocks_for(size_t threads, unsigned block = 256) { return (unsigned)((threads + block - 1) / block); }

// ------------------------------------------------------------------------------------ tables
static Mod make_mod(u64 q, u64 N)
{
    Mod m{};
    m.q = q;
    m.two_q = 2 * q;
    m.nq = 0 - q;
    m.bits = (u32)hm::bitlen(q);
    m.sh = m.bits - 2;
    m.mu = (u64)((((hm::u128)1) << (62 + m.bits)) / q);
    m.r64 = (u64)((((hm::u128)1) << 64) / q);
    m.ninv = hm::invmod(N % q, q);
    m.ninv_s = hm::shoup(m.ninv, q);
    // pseudo-Mersenne lazy reduction (modarith.cuh reduce_pm): q = 2^bits - c with (2^(64-bits) - 1) c + 2^bits - 1 < 2q
    {
        const hm::u128 c = ((hm::u128)1 << m.bits) - q;
        static const bool no_pm = getenv("B200HE_NO_PM") && atoi(getenv("B200HE_NO_PM"));
        if (!no_pm && m.bits > 32 && c < ((hm::u128)1 << 32) &&
            (((hm::u128)1 << (64 - m.bits)) - 1) * c + (((hm::u128)1 << m.bits) - 1) < (hm::u128)2 * q)
            m.pm_c = (u32)c;
    }
    // FP64 domain for small primes (modarith.cuh); B200HE_NO_DP=1 keeps every modulus on the integer pipe (A/B timing)
    static const bool no_dp = getenv("B200HE_NO_DP") && atoi(getenv("B200HE_NO_DP"));
    m.dp = (m.bits <= B200HE_DP_MAX_BITS && !no_dp) ? 1 : 0;
    m.dq = (double)q;
    m.dnq = -m.dq;
    m.dqinv = 1.0 / m.dq;
    m.dqinv_up = m.dqinv;
    if (fma(m.dqinv_up, m.dq, -1.0) < 0) m.dqinv_up = nextafter(m.dqinv_up, 2.0);   // sign of qinv*q - 1 is exact under one rounding
    m.dninv = (double)m.ninv;
    m.dninv_q = m.dninv / m.dq;
    return m;
}
// twiddle in the form the modulus' transforms consume: (w, floor(w 2^64 / q)), or (w, RN(w / q)) as doubles
static ulonglong2 make_tw(u64 w, const Mod &m)
{
    if (!m.dp) return make_ulonglong2(w, hm::shoup(w, m.q));
    const double d = (double)w, dq = d / m.dq;
    ulonglong2 r;
    memcpy(&r.x, &d, 8);
    memcpy(&r.y, &dq, 8);
    return r;
}

// Lazy-reduction schedule of the CTA-local transforms for one modulus (see ntt_core.cuh bfly_fwd /
// bfly_inv).  Values may grow up to Bmax*q <= 2^64 between reductions, Bmax = floor((2^64-1)/q).
template <int LG> static void lazy_schedule(Mod &m, int c, int L_top)
{
    const u64 bmax = ~u64(0) / m.q;
    const int NP = Sched<LG>::NP;
    // forward: inputs < 2q, each split pre-stage and each stage adds 2q
    u64 b = 2 + 2 * (u64)c;
    m.fwd_mask = 0;
    for (int P = 0; P < NP; P++) {
        const u64 K = Sched<LG>::K[P];
        if (b + 2 * K > bmax) { m.fwd_mask |= 1u << P; b = 2; }
        if (b + 2 * K > bmax) { m.fwd_mask |= 1u << (8 + P); b = 2 + 2 * (K - K / 2); }
        else b += 2 * K;
    }
    // inverse: canonical inputs (< q); the sum path doubles its bound per stage
    b = 1;
    m.inv_mask = 0;
    for (int P = NP - 1; P >= 0; P--) {
        const int K = Sched<LG>::K[P];
        if ((b << K) > bmax) { m.inv_mask |= 1u << P; b = 2; }
        m.inv_c[P] = (u32)b;
        if ((b << K) > bmax) { m.inv_mask |= 1u << (8 + P); b = u64(2) << (K - K / 2); }
        else b <<= K;
    }
    // key-switch inner product: every digit adds a Shoup product < 2q to the accumulators
    const u64 period = bmax >= 4 ? (bmax - 2) / 2 : 1;
    m.acc_period = (u32)(period > (u64)L_top + 1 ? (u64)L_top + 1 : period);
    if (m.dp) m.acc_period = 16;   // FP64 domain: 0.75 q per digit, dp_canon accepts magnitudes up to 16 q
}

static int build_tables(b200he_ctx *c, const std::vector<u64> &moduli, const std::vector<u64> &psi)
{
    const size_t N = c->N, M = moduli.size();
    c->M = (int)M;
    c->mods.resize(M);
    std::vector<ulonglong2> tw(M * N), itw(M * N), qinv(M * M);
    std::vector<u64> halfmod(M * M);
    for (size_t i = 0; i < M; i++) {
        const u64 q = moduli[i];
        c->mods[i] = make_mod(q, N);
        NTT_DISPATCH(*c, lazy_schedule<LG>(c->mods[i], c->c, (int)c->K));
        const u64 ipsi = hm::invmod(psi[i], q);
        u64 p = 1, ip = 1, ipsi_half = 0;
        for (size_t k = 0; k < N; k++) {
            const size_t r = hm::brv((uint32_t)k, c->logn);
            tw[i * N + r] = make_tw(p, c->mods[i]);
            itw[i * N + r] = make_tw(ip, c->mods[i]);
            if (r == 1) ipsi_half = ip;
            p = hm::mulmod(p, psi[i], q);
            ip = hm::mulmod(ip, ipsi, q);
        }
        // slot 0 (the unused psi^0) carries the last inverse stage's twiddle with N^{-1} folded in
        itw[i * N] = make_tw(hm::mulmod(ipsi_half, c->mods[i].ninv, q), c->mods[i]);
        for (size_t x = 0; x < M; x++) {
            const u64 qx = moduli[x];
            u64 inv = (x == i) ? 0 : hm::invmod(qx % q, q);
            qinv[x * M + i] = make_ulonglong2(inv, hm::shoup(inv, q));
            halfmod[x * M + i] = (qx >> 1) % q;
        }
    }
    const size_t b_mods = M * sizeof(Mod), b_tw = M * N * sizeof(ulonglong2), b_qinv = M * M * sizeof(ulonglong2),
                 b_half = M * M * sizeof(u64);
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t total = al(b_mods) + 2 * al(b_tw) + al(b_qinv) + al(b_half);
    CK(cudaMalloc(&c->d_tables, total));
    unsigned char *p = (unsigned char *)c->d_tables;
    CK(cudaMemcpy(p, c->mods.data(), b_mods, cudaMemcpyHostToDevice));
    c->T.mods = (const Mod *)p;
    p += al(b_mods);
    CK(cudaMemcpy(p, tw.data(), b_tw, cudaMemcpyHostToDevice));
    c->T.tw = (const ulonglong2 *)p;
    p += al(b_tw);
    CK(cudaMemcpy(p, itw.data(), b_tw, cudaMemcpyHostToDevice));
    c->T.itw = (const ulonglong2 *)p;
    p += al(b_tw);
    CK(cudaMemcpy(p, qinv.data(), b_qinv, cudaMemcpyHostToDevice));
    c->T.qinv = (const ulonglong2 *)p;
    p += al(b_qinv);
    CK(cudaMemcpy(p, halfmod.data(), b_half, cudaMemcpyHostToDevice));
    c->T.halfmod = (const u64 *)p;
    c->T.N = (int)N;
    c->T.M = (int)M;
    return 0;
}

static int set_smem_attrs(const b200he_ctx *c)
{
    const Geo g = geo(c);
    CK((cudaError_t)smem_attrs_ntt(g));
    CK((cudaError_t)smem_attrs_ks(g));
    CK((cudaError_t)smem_attrs_moddown(g));
    return 0;
}

static int init_behz(b200he_ctx *c, std::vector<u64> &moduli, std::vector<u64> &psi);
struct b200he_ctx;
static int stage_init(b200he_ctx *c, size_t ct_bytes);
static int upload_behz(b200he_ctx *c);

extern "C" int b200he_ctx_create(int scheme, uint32_t N, uint32_t K, const uint64_t *moduli, const uint64_t *psi,
                                 uint64_t plain_modulus, int device, b200he_ctx **out)
{
    if (!out) return fail("ctx_create: out is NULL");
    *out = nullptr;
    if (scheme != B200HE_BFV && scheme != B200HE_CKKS) return fail("ctx_create: unknown scheme %d", scheme);
    int logn = 0;
    while ((1u << logn) < N) logn++;
    if ((1u << logn) != N || logn < 10 || logn > 15) return fail("ctx_create: N=%u must be a power of two in [1024, 32768]", N);
    if (K < 1 || K > 32) return fail("ctx_create: K=%u out of range", K);
    if (!moduli || !psi) return fail("ctx_create: moduli/psi NULL");
    std::vector<u64> mv(moduli, moduli + K), pv(psi, psi + K);
    for (u32 i = 0; i < K; i++) {
        if (mv[i] >> 61 || mv[i] < 3 || (mv[i] - 1) % (2 * (u64)N)) return fail("ctx_create: modulus %u (%llu) must be < 2^61 and = 1 mod 2N", i, (unsigned long long)mv[i]);
        if (hm::powmod(pv[i], N, mv[i]) != mv[i] - 1) return fail("ctx_create: psi[%u] is not a primitive 2N-th root", i);
        for (u32 j = 0; j < i; j++)
            if (mv[j] == mv[i]) return fail("ctx_create: repeated modulus");
    }
    if (scheme == B200HE_BFV && (plain_modulus < 2 || K < 2)) return fail("ctx_create: BFV needs plain_modulus >= 2 and K >= 2");
#ifndef B200HE_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("ctx_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail("ctx_create: device %d out of range (%d present)", device, ndev);
#endif
    CK(cudaSetDevice(device));
    b200he_ctx *c = new b200he_ctx;
#ifndef B200HE_EMU
    cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device);
#else
    c->n_sm = 3;   // emulation: a few CTAs, several limbs each
#endif
    c->scheme = scheme;
    c->device = device;
    c->N = N;
    c->K = K;
    c->logn = logn;
    c->lognl = logn > 13 ? 13 : logn;
    if (const char *e = getenv("B200HE_LOGNL")) {   // tuning knob: CTA-local transform size (limb split 2^(logn-lognl) ways, <= 4)
        const int v = atoi(e);
        if (v >= 10 && v <= 13 && v == logn) c->lognl = v;   // limbs are split only into chunks of 8192
    }
    c->c = logn - c->lognl;
    c->t = plain_modulus;
    // every failure below releases what has been set up so far through b200he_ctx_destroy (tables, BEHZ constants, stream,
    // pinned staging); the error text of the failing step is kept
    auto bail = [&](int rc) {
        const std::string keep = g_err;
        b200he_ctx_destroy(c);
        g_err = keep;
        return rc ? rc : -1;
    };
    if (scheme == B200HE_BFV && init_behz(c, mv, pv)) return bail(-1);
    if (build_tables(c, mv, pv)) return bail(-1);
    if (scheme == B200HE_BFV && upload_behz(c)) return bail(-1);
#ifndef B200HE_EMU
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail("ctx_create: stream"));
    c->own_stream = true;
    {
        const char *e = getenv("B200HE_KS_OVERLAP");
        if ((!e || atoi(e) != 0) && c->c == 2) {   // only limbs of four chunks have a launch pair to overlap
            if (cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&c->aux_fork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&c->aux_join, cudaEventDisableTiming) != cudaSuccess)
                return bail(fail("ctx_create: side stream"));
        }
    }
#endif
    if (int rc = set_smem_attrs(c)) return bail(rc);
    // pinned staging for load()/store() (page-locking 64 MB takes tens of milliseconds: not inside the first load)
    if (stage_init(c, 0)) return bail(-1);
    // the tables were uploaded with synchronous copies from pageable memory (NULL stream): make sure they have landed
    // before the first kernel on the context's non-blocking stream can run
    if (cudaDeviceSynchronize() != cudaSuccess) return bail(fail("ctx_create: device synchronisation failed"));
    *out = c;
    return 0;
}

extern "C" void b200he_ctx_destroy(b200he_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    {   // batches that outlive the context keep their handle but lose their storage: later calls on them fail cleanly
        std::lock_guard<std::mutex> lock(c->batches_mtx);
        for (b200he_batch *b : c->batches) {
            b->ctx = nullptr;
            b->d = nullptr;
            b->cap_words = 0;
            b->count = 0;
        }
        c->batches.clear();
    }
    for (auto &r : c->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (c->relin) cudaFree(c->relin);
    for (auto &kv : c->gal) cudaFree(kv.second);
    for (auto &kv : c->galtab) cudaFree(kv.second.d);
    c->pool.release();
    for (auto &kv : c->pool.sizes) cudaFree(kv.first);   // blocks still held by undestroyed batches
    if (c->d_tables) cudaFree(c->d_tables);
    if (c->behz.d_blob) cudaFree(c->behz.d_blob);
    for (int i = 0; i < 2; i++) {
        if (c->stage_buf[i]) cudaFreeHost(c->stage_buf[i]);
        if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
    }
#ifndef B200HE_EMU
    if (c->own_stream) cudaStreamDestroy(c->stream);
    if (c->aux) cudaStreamDestroy(c->aux);
    if (c->aux_fork) cudaEventDestroy(c->aux_fork);
    if (c->aux_join) cudaEventDestroy(c->aux_join);
#endif
    delete c;
}

extern "C" int b200he_ctx_set_stream(b200he_ctx *c, void *s)
{
    if (!c) return fail("set_stream: ctx NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
#ifndef B200HE_EMU
    if (c->own_stream) cudaStreamDestroy(c->stream);
#endif
    c->own_stream = false;
    c->stream = (cudaStream_t)s;
    return 0;
}
extern "C" int b200he_ctx_sync(b200he_ctx *c)
{
    if (!c) return fail("sync: ctx NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return 0;
}
extern "C" int b200he_ctx_set_workspace(b200he_ctx *c, uint64_t bytes)
{
    if (!c || bytes < (u64(1) << 20)) return fail("set_workspace: need >= 1 MiB");
    c->workspace = bytes;
    return 0;
}

// ------------------------------------------------------------------------------------ keys
static size_t key_words(const b200he_ctx *c) { return (size_t)(c->K - 1) * 2 * c->K * c->N; }

// device layout of a key: 2 * key_words, residues interleaved with their Shoup quotients (kernels.cuh key_word_index)
static int upload_key(b200he_ctx *c, u64 *d, const uint64_t *key)
{
    const size_t words = key_words(c);
    for (size_t J = 0; J + 1 < c->K; J++)   // residues must be canonical: the Shoup form relies on k < q
        for (size_t k = 0; k < 2; k++)
            for (size_t l = 0; l < c->K; l++) {
                const uint64_t *p = key + ((J * 2 + k) * c->K + l) * c->N, q = c->mods[l].q;
                for (size_t n = 0; n < c->N; n++)
                    if (p[n] >= q) return fail("set key: residue out of range (digit %zu, component %zu, limb %zu)", J, k, l);
            }
    u64 *stage = (u64 *)c->pool.get(words * 8);
    if (!stage) return fail("set key: out of device memory");
    cudaError_t e = cudaMemcpyAsync(stage, key, words * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        LAUNCH(c, B200HE_KERN_COPY, k_shoup_quotients, blocks_for(words), 256, 0, c->T, stage, d, (int)c->K, words, c->lognl);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    c->pool.put(stage);
    if (e != cudaSuccess) return fail("set key: %s", cudaGetErrorString(e));
    return 0;
}

static int galois_table(b200he_ctx *c, u32 elt, const b200he_ctx::GalTab **out);

extern "C" int b200he_set_relin_key(b200he_ctx *c, const uint64_t *key)
{
    if (!c || !key) return fail("set_relin_key: NULL argument");
    if (c->K < 2) return fail("set_relin_key: context has no special prime (K < 2)");
    CK(cudaSetDevice(c->device));
    // the key is published only once it is valid and resident: a rejected upload leaves the previous state untouched
    u64 *d = nullptr;
    CK(cudaMalloc((void **)&d, 2 * key_words(c) * 8));
    if (int rc = upload_key(c, d, key)) { cudaFree(d); return rc; }
    if (c->relin) cudaFree(c->relin);   // (upload_key has synchronised the stream: nothing in flight reads the old key)
    c->relin = d;
    return 0;
}
extern "C" int b200he_set_galois_key(b200he_ctx *c, uint32_t elt, const uint64_t *key)
{
    if (!c || !key) return fail("set_galois_key: NULL argument");
    if (c->K < 2) return fail("set_galois_key: context has no special prime (K < 2)");
    if (!(elt & 1) || elt >= 2 * c->N) return fail("set_galois_key: invalid Galois element %u", elt);
    CK(cudaSetDevice(c->device));
    u64 *d = nullptr;
    CK(cudaMalloc((void **)&d, 2 * key_words(c) * 8));
    if (int rc = upload_key(c, d, key)) { cudaFree(d); return rc; }
    if (c->scheme == B200HE_CKKS) {   // the element's permutation table goes up with its key, not inside the first rotation
        const b200he_ctx::GalTab *t = nullptr;
        if (int rc = galois_table(c, elt, &t)) { cudaFree(d); return rc; }
    }
    auto it = c->gal.find(elt);
    if (it != c->gal.end()) { cudaFree(it->second); it->second = d; }
    else c->gal.emplace(elt, d);
    return 0;
}
extern "C" int b200he_has_galois_key(const b200he_ctx *c, uint32_t elt) { return c && c->gal.count(elt) ? 1 : 0; }

// NTT-form Galois permutation (SEAL GaloisTool::generate_table_ntt): out[i] = in[table[i]]
static int galois_table(b200he_ctx *c, u32 elt, const b200he_ctx::GalTab **out)
{
    auto it = c->galtab.find(elt);
    if (it == c->galtab.end()) {
        const u32 N = c->N;
        std::vector<u32> tab(N);
        for (u32 i = 0; i < N; i++) {
            const u32 odd = 2 * hm::brv(i, c->logn) + 1;
            const u32 idx = (u32)(((u64)elt * odd) & (2 * (u64)N - 1));
            tab[i] = hm::brv((idx - 1) >> 1, c->logn);
        }
        b200he_ctx::GalTab t;
        // the table maps aligned blocks of 2^k indices onto aligned blocks of 2^k indices; the kernels rely on it at
        // chunk granularity (the chunk an output chunk gathers from) -- checked here rather than assumed
        const u32 NL = 1u << c->lognl;
        for (u32 r = 0; r < (N >> c->lognl); r++) {
            t.chunk[r] = (unsigned char)(tab[r * NL] >> c->lognl);
            for (u32 i = 0; i < NL; i++)
                if ((tab[r * NL + i] >> c->lognl) != t.chunk[r]) return fail("galois table of element %u does not map chunks onto chunks", elt);
        }
        CK(cudaMalloc((void **)&t.d, N * sizeof(u32)));
        // on the context's stream: a plain cudaMemcpy from pageable memory may return before its DMA (ordered on the
        // NULL stream, which this non-blocking stream does not wait for) has landed
        CK(cudaMemcpyAsync(t.d, tab.data(), N * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        it = c->galtab.emplace(elt, t).first;
    }
    *out = &it->second;
    return 0;
}

// ------------------------------------------------------------------------------------ batches
extern "C" int b200he_batch_create(b200he_ctx *c, b200he_batch **out)
{
    if (!c || !out) return fail("batch_create: NULL argument");
    b200he_batch *b = new b200he_batch;
    b->ctx = c;
    b->n_poly = c->N;
    {
        std::lock_guard<std::mutex> lock(c->batches_mtx);
        c->batches.insert(b);
    }
    *out = b;
    return 0;
}
extern "C" void b200he_batch_destroy(b200he_batch *b)
{
    if (!b) return;
    if (b200he_ctx *c = b->ctx) {   // (an orphaned batch owns nothing: its block went with the context's pool)
        c->pool.put(b->d);
        std::lock_guard<std::mutex> lock(c->batches_mtx);
        c->batches.erase(b);
    }
    delete b;
}
static int batch_shape(b200he_batch *b, uint64_t count, int size, int L, int ntt, double scale)
{
    LIVE(b, "batch");
    b200he_ctx *c = b->ctx;
    if (size < 1 || size > 3) return fail("batch: size %d out of range", size);
    const int Lmax = c->K > 1 ? (int)c->K - 1 : 1;
    if (L < 1 || L > Lmax) return fail("batch: level L=%d out of range [1,%d]", L, Lmax);
    const size_t words = (size_t)count * size * L * c->N;
    if (words > b->cap_words) {
        CK(cudaSetDevice(c->device));
        c->pool.put(b->d);
        b->d = (u64 *)c->pool.get(words * 8);
        if (!b->d) { b->cap_words = 0; return fail("batch: out of device memory (%zu bytes)", words * 8); }
        b->cap_words = c->pool.sizes[b->d] / 8;
    }
    b->count = count;
    b->size = size;
    b->L = L;
    b->ntt = ntt;
    b->scale = scale;
    return 0;
}
extern "C" int b200he_batch_resize(b200he_batch *b, uint64_t count, int size, int L, int ntt_form, double scale)
{
    if (!b) return fail("batch_resize: NULL batch");
    return batch_shape(b, count, size, L, ntt_form, scale);
}
extern "C" int b200he_batch_upload(b200he_batch *b, uint64_t first, uint64_t n, const uint64_t *host)
{
    if (!b || !host) return fail("batch_upload: NULL argument");
    LIVE(b, "batch_upload");
    if (first + n > b->count) return fail("batch_upload: range [%llu,%llu) exceeds count %llu", (unsigned long long)first, (unsigned long long)(first + n), (unsigned long long)b->count);
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaMemcpyAsync(b->d + first * b->ct_words(), host, n * b->ct_words() * 8, cudaMemcpyHostToDevice, b->ctx->stream));
    return 0;
}
extern "C" int b200he_batch_download(const b200he_batch *b, uint64_t first, uint64_t n, uint64_t *host)
{
    if (!b || !host) return fail("batch_download: NULL argument");
    LIVE(b, "batch_download");
    if (first + n > b->count) return fail("batch_download: range exceeds count");
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaMemcpyAsync(host, b->d + first * b->ct_words(), n * b->ct_words() * 8, cudaMemcpyDeviceToHost, b->ctx->stream));
    CK(cudaStreamSynchronize(b->ctx->stream));
    CK(cudaGetLastError());
    return 0;
}
extern "C" int b200he_batch_download_async(const b200he_batch *b, uint64_t first, uint64_t n, uint64_t *host)
{
    if (!b || !host) return fail("batch_download_async: NULL argument");
    LIVE(b, "batch_download_async");
    if (first + n > b->count) return fail("batch_download_async: range exceeds count");
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaMemcpyAsync(host, b->d + first * b->ct_words(), n * b->ct_words() * 8, cudaMemcpyDeviceToHost, b->ctx->stream));
    return 0;
}
// ---- load / store of separately allocated host ciphertexts through pinned staging ----
static int stage_init(b200he_ctx *c, size_t ct_bytes)
{
    size_t want = size_t(32) << 20;
    if (const char *e = getenv("B200HE_STAGE_MB")) want = (size_t)atoi(e) << 20;
    if (want < ct_bytes) want = ct_bytes;
    if (c->stage_bytes >= want) return 0;
    for (int i = 0; i < 2; i++) {
        if (c->stage_ev[i]) CK(cudaEventSynchronize(c->stage_ev[i]));
        if (c->stage_buf[i]) CK(cudaFreeHost(c->stage_buf[i]));
        c->stage_buf[i] = nullptr;
        CK(cudaMallocHost(&c->stage_buf[i], want));
        if (!c->stage_ev[i]) CK(cudaEventCreate(&c->stage_ev[i]));
    }
    c->stage_bytes = want;
    return 0;
}
// Host worker threads of the staged copies, started once per process and kept: a 32 MB staging chunk is copied in about
// a millisecond, the same order as creating and joining 16 threads for it (measured, tools/host_path_probe.py).
// The pool is never destroyed (its threads are detached; a static destructor joining threads at exit of a dlopen'ed
// library is a known way to hang the host process).
namespace {
struct HostPool {
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    size_t threads = 0;
    explicit HostPool(size_t T) : threads(T)
    {
        for (size_t t = 0; t < T; t++)
            std::thread([this]() {
                for (;;) {
                    std::function<void()> job;
                    {
                        std::unique_lock<std::mutex> lock(m);
                        cv.wait(lock, [this]() { return !q.empty(); });
                        job = std::move(q.front());
                        q.pop_front();
                    }
                    job();
                }
            }).detach();
    }
    void submit(std::function<void()> job)
    {
        {
            std::lock_guard<std::mutex> lock(m);
            q.push_back(std::move(job));
        }
        cv.notify_one();
    }
};
HostPool &host_pool()
{
    static HostPool *pool = []() {
        size_t T = std::thread::hardware_concurrency();
        if (T > 16) T = 16;   // the copies are memory-bound; into fresh pageable destinations, page-fault-bound
        if (const char *e = getenv("B200HE_HOST_THREADS")) T = (size_t)atoi(e);
        return new HostPool(T > 1 ? T - 1 : 0);   // the calling thread works too
    }();
    return *pool;
}
}   // namespace
// staging -> pageable destination with streaming stores where the platform has them: the destination is not read again
// by these threads, and a regular store would first pull every line into the cache (a third more memory traffic)
#if defined(__x86_64__) && !defined(B200HE_EMU)
#include <emmintrin.h>
static void copy_out(void *dst, const void *src, size_t len)
{
    static const bool nt = []() { const char *e = getenv("B200HE_NT_COPY"); return !e || atoi(e) != 0; }();
    if (!nt || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15) || (len & 63)) {
        memcpy(dst, src, len);
        return;
    }
    __m128i *d = (__m128i *)dst;
    const __m128i *s = (const __m128i *)src;
    for (size_t i = 0; i < len / 16; i += 4) {
        const __m128i a = _mm_load_si128(s + i), b = _mm_load_si128(s + i + 1), c = _mm_load_si128(s + i + 2), e = _mm_load_si128(s + i + 3);
        _mm_stream_si128(d + i, a);
        _mm_stream_si128(d + i + 1, b);
        _mm_stream_si128(d + i + 2, c);
        _mm_stream_si128(d + i + 3, e);
    }
    _mm_sfence();
}
#else
static void copy_out(void *dst, const void *src, size_t len) { memcpy(dst, src, len); }
#endif
// run fn(i) for i in [0, n) on the calling thread and the pool; items are handed out one at a time
template <class F> static void host_parallel(size_t n, F fn)
{
    if (n == 0) return;
    HostPool &pool = host_pool();
    size_t helpers = pool.threads < n - 1 ? pool.threads : n - 1;
    if (helpers == 0) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    struct Shared {
        std::atomic<size_t> next{ 0 };
        std::mutex m;
        std::condition_variable cv;
        size_t running;
    } sh;
    sh.running = helpers;
    auto work  = [&sh, &fn, n]() {
        for (size_t i; (i = sh.next.fetch_add(1, std::memory_order_relaxed)) < n;) fn(i);
    };
    for (size_t t = 0; t < helpers; t++)
        pool.submit([&sh, work]() {
            work();
            std::lock_guard<std::mutex> lock(sh.m);   // held while notifying: sh lives on the caller's stack
            if (--sh.running == 0) sh.cv.notify_one();
        });
    work();
    std::unique_lock<std::mutex> lock(sh.m);
    sh.cv.wait(lock, [&sh]() { return sh.running == 0; });
}
extern "C" int b200he_batch_upload_scattered(b200he_batch *b, uint64_t first, uint64_t n, const uint64_t *const *host)
{
    if (!b || (!host && n)) return fail("batch_upload_scattered: NULL argument");
    LIVE(b, "batch_upload_scattered");
    if (first + n > b->count) return fail("batch_upload_scattered: range exceeds count");
    if (!n) return 0;
    for (uint64_t i = 0; i < n; i++)
        if (!host[i]) return fail("batch_upload_scattered: host[%llu] is NULL", (unsigned long long)i);
    b200he_ctx *c = b->ctx;
    CK(cudaSetDevice(c->device));
    const size_t ctb = b->ct_words() * 8;
    TRY(stage_init(c, ctb));
    const uint64_t per = c->stage_bytes / ctb;
    const size_t piece = size_t(256) << 10, pieces = (ctb + piece - 1) / piece;   // unit of work of a host thread
    int k = 0;
    for (uint64_t at = 0; at < n; at += per, k ^= 1) {
        const uint64_t m = n - at < per ? n - at : per;
        CK(cudaEventSynchronize(c->stage_ev[k]));   // the copy that last used this buffer has finished
        unsigned char *buf = (unsigned char *)c->stage_buf[k];
        host_parallel(m * pieces, [&](size_t w) {
            const size_t i = w / pieces, off = (w % pieces) * piece, len = off + piece < ctb ? piece : ctb - off;
            memcpy(buf + i * ctb + off, (const unsigned char *)host[at + i] + off, len);
        });
        CK(cudaMemcpyAsync(b->d + (first + at) * b->ct_words(), buf, m * ctb, cudaMemcpyHostToDevice, c->stream));
        CK(cudaEventRecord(c->stage_ev[k], c->stream));
    }
    return 0;
}
extern "C" int b200he_batch_download_scattered(const b200he_batch *b, uint64_t first, uint64_t n, uint64_t *const *host)
{
    if (!b || (!host && n)) return fail("batch_download_scattered: NULL argument");
    LIVE(b, "batch_download_scattered");
    if (first + n > b->count) return fail("batch_download_scattered: range exceeds count");
    if (!n) return 0;
    for (uint64_t i = 0; i < n; i++)
        if (!host[i]) return fail("batch_download_scattered: host[%llu] is NULL", (unsigned long long)i);
    b200he_ctx *c = b->ctx;
    CK(cudaSetDevice(c->device));
    const size_t ctb = b->ct_words() * 8;
    TRY(stage_init(c, ctb));
    const uint64_t per = c->stage_bytes / ctb;
    const size_t piece = size_t(256) << 10, pieces = (ctb + piece - 1) / piece;   // unit of work of a host thread
    // chunk j+1 moves over PCIe while the host threads scatter chunk j
    auto issue = [&](uint64_t at, int k) -> int {
        const uint64_t m = n - at < per ? n - at : per;
        CK(cudaEventSynchronize(c->stage_ev[k]));
        CK(cudaMemcpyAsync(c->stage_buf[k], b->d + (first + at) * b->ct_words(), m * ctb, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaEventRecord(c->stage_ev[k], c->stream));
        return 0;
    };
    TRY(issue(0, 0));
    int k = 0;
    for (uint64_t at = 0; at < n; at += per, k ^= 1) {
        const uint64_t m = n - at < per ? n - at : per;
        if (at + per < n) TRY(issue(at + per, k ^ 1));
        CK(cudaEventSynchronize(c->stage_ev[k]));
        const unsigned char *buf = (const unsigned char *)c->stage_buf[k];
        host_parallel(m * pieces, [&](size_t w) {
            const size_t i = w / pieces, off = (w % pieces) * piece, len = off + piece < ctb ? piece : ctb - off;
            copy_out((unsigned char *)host[at + i] + off, buf + i * ctb + off, len);
        });
    }
    CK(cudaGetLastError());
    return 0;
}
extern "C" int b200he_batch_copy_from(b200he_batch *dst, const b200he_batch *src)
{
    if (!dst || !src) return fail("batch_copy_from: NULL argument");
    if (dst == src) return 0;
    LIVE(dst, "batch_copy_from");
    LIVE(src, "batch_copy_from");
    b200he_ctx *cd = dst->ctx, *cs = src->ctx;
    if (cd->N != cs->N || cd->K != cs->K || cd->scheme != cs->scheme) return fail("batch_copy_from: contexts differ in parameters");
    TRY(batch_shape(dst, src->count, src->size, src->L, src->ntt, src->scale));
    const size_t bytes = (size_t)src->count * src->ct_words() * 8;
    if (!bytes) return 0;
    // src's stream has produced the data -> dst's stream copies -> src's stream may reuse the block afterwards
    cudaEvent_t ready = nullptr, done = nullptr;
    CK(cudaSetDevice(cs->device));
    CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    CK(cudaEventRecord(ready, cs->stream));
    CK(cudaSetDevice(cd->device));
    CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    CK(cudaStreamWaitEvent(cd->stream, ready, 0));
    CK(cudaMemcpyPeerAsync(dst->d, cd->device, src->d, cs->device, bytes, cd->stream));
    CK(cudaEventRecord(done, cd->stream));
    CK(cudaSetDevice(cs->device));
    CK(cudaStreamWaitEvent(cs->stream, done, 0));
    CK(cudaEventDestroy(ready));   // destruction is deferred by the runtime until the event has completed
    CK(cudaEventDestroy(done));
    return 0;
}
extern "C" uint64_t b200he_batch_count(const b200he_batch *b) { return b ? b->count : 0; }
extern "C" int b200he_batch_size(const b200he_batch *b) { return b ? b->size : 0; }
extern "C" int b200he_batch_level(const b200he_batch *b) { return b ? b->L : 0; }
extern "C" int b200he_batch_ntt_form(const b200he_batch *b) { return b ? b->ntt : 0; }
extern "C" double b200he_batch_scale(const b200he_batch *b) { return b ? b->scale : 0.0; }
extern "C" int b200he_batch_set_scale(b200he_batch *b, double s)
{
    if (!b) return fail("batch_set_scale: NULL batch");
    b->scale = s;
    return 0;
}
extern "C" void *b200he_batch_device_ptr(b200he_batch *b) { return b ? b->d : nullptr; }

// Output staging: ops write into a fresh block when the output handle aliases an input, then swap.
struct OutBuf {
    b200he_batch *out;
    b200he_batch tmp;
    bool aliased;
    OutBuf(b200he_batch *o, const b200he_batch *a, const b200he_batch *b = nullptr) : out(o), aliased(o == a || o == b)
    {
        tmp.ctx = o->ctx;   // (a scratch handle on the stack: not registered with the context, never outlives the call)
        tmp.n_poly = o->n_poly;
    }
    int shape(uint64_t count, int size, int L, int ntt, double scale)
    {
        return batch_shape(aliased ? &tmp : out, count, size, L, ntt, scale);
    }
    u64 *ptr() { return aliased ? tmp.d : out->d; }
    void commit()
    {
        if (!aliased) return;
        out->ctx->pool.put(out->d);
        out->d = tmp.d;
        out->cap_words = tmp.cap_words;
        out->count = tmp.count;
        out->size = tmp.size;
        out->L = tmp.L;
        out->ntt = tmp.ntt;
        out->scale = tmp.scale;
        tmp.d = nullptr;
    }
    ~OutBuf()
    {
        if (tmp.d) out->ctx->pool.put(tmp.d);
    }
};

// device copy of a host index map (NULL stays NULL = identity); bounds-checked on the host
struct DevIdx {
    b200he_ctx *c;
    u32 *d = nullptr;
    DevIdx(b200he_ctx *c_) : c(c_) {}
    int set(const uint32_t *host, uint64_t n, uint64_t limit, const char *what)
    {
        if (!host) return 0;
        for (uint64_t i = 0; i < n; i++)
            if (host[i] >= limit) return fail("%s: index %u at position %llu out of range (count %llu)", what, host[i], (unsigned long long)i, (unsigned long long)limit);
        d = (u32 *)c->pool.get(n * sizeof(u32));
        if (!d) return fail("%s: out of device memory", what);
        CK(cudaMemcpyAsync(d, host, n * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
        return 0;
    }
    ~DevIdx() { c->pool.put(d); }
};

static bool same_scale(double a, double b) { return fabs(a - b) <= 1e-9 * fmax(fabs(a), fabs(b)) || a == b; }

// ------------------------------------------------------------------------------------ transforms
// forward NTT of nlimbs limbs (grouped L per outer stride).  In place only when the limb is unsplit.
// k_moddown over D.nJ output limbs of `polys` polynomials: the kernel that holds only the integer instance, only the
// FP64 instance, or both, by the kinds of moduli among those limbs
static void launch_moddown(b200he_ctx *c, ModDownArgs D, size_t polys)
{
    unsigned dpmask = 0;
    for (int j = 0; j < D.nJ; j++)
        if (c->mods[j].dp) dpmask |= 1u << j;
    const unsigned all = D.nJ >= 32 ? ~0u : ((1u << D.nJ) - 1);
    D.jmask = all;
    D.nJsub = D.nJ;
    const int kind = dpmask == 0 ? KIND_INT : dpmask == all ? KIND_DP : KIND_BOTH;
    PROF(c, B200HE_KERN_MODDOWN, launch_moddown_kind(geo(c), kind, c->T, D, polys * D.nJ));
    if (c->prof) {
        // per output limb: one transform + the constant multiplies of the epilogue (1, or 3 with the fused rescale);
        // traffic: the accumulator tile, the output, the addend where there is one, the rounded limb(s) once per polynomial
        const double N = c->N, mults = D.rp2 ? 3.0 : 1.0, n_add = (D.addend[0] ? 0.5 : 0.0) + (D.addend[1] ? 0.5 : 0.0) + (D.gal ? 0.5 : 0.0);
        for (int j = 0; j < D.nJ; j++) work_add(c, B200HE_KERN_MODDOWN, j, polys * (bfly_per_limb(c) + mults * N), polys * (2.0 + n_add) * N * 8);
        work_add(c, B200HE_KERN_MODDOWN, 0, 0, polys * (D.rp2 ? 2.0 : 1.0) * N * 8);
    }
}
static int ntt_fwd(b200he_ctx *c, const u64 *src, u64 *dst, size_t nlimbs, size_t src_outer, size_t dst_outer, int L, int mod_base)
{
    if (!nlimbs) return 0;
    static const bool persistent = !(getenv("B200HE_NO_PERSISTENT") && atoi(getenv("B200HE_NO_PERSISTENT")));
    // unsplit limbs, more than one wave: persistent CTAs with TMA prefetch
    const unsigned pctas = (c->c == 0 && persistent && nlimbs > (size_t)c->n_sm) ? (unsigned)c->n_sm : 0u;
    PROF(c, B200HE_KERN_NTT_FWD, launch_ntt_fwd(geo(c), c->T, src, dst, src_outer, dst_outer, L, mod_base, nlimbs, pctas));
    if (c->prof)
        for (int l = 0; l < L; l++) work_add(c, B200HE_KERN_NTT_FWD, mod_base + l, (double)nlimbs / L * bfly_per_limb(c), (double)nlimbs / L * 2.0 * c->N * 8);
    LAUNCH_CHECK();
    return 0;
}
static int ntt_inv(b200he_ctx *c, const u64 *src, u64 *dst, size_t nlimbs, size_t src_outer, size_t dst_outer, int L, int mod_base, int mode,
                   const InvFuse *fuse = nullptr)
{
    if (!nlimbs) return 0;
    InvFuse F{};
    if (fuse) F = *fuse;
    // kinds of moduli among the launch's limbs (modulus ids mod_base .. mod_base + L - 1)
    int n_dp = 0;
    for (int l = 0; l < L; l++) n_dp += c->mods[mod_base + l].dp ? 1 : 0;
    const int kind = n_dp == 0 ? KIND_INT : n_dp == L ? KIND_DP : KIND_BOTH;
    PROF(c, B200HE_KERN_NTT_INV, launch_ntt_inv(geo(c), kind, c->T, src, dst, src_outer, dst_outer, L, mod_base, mode, F, nlimbs));
    if (c->prof)   // fused relinearize+rescale last limb: one more operand tile and one constant multiply per coefficient
        for (int l = 0; l < L; l++)
            work_add(c, B200HE_KERN_NTT_INV, mod_base + l, (double)nlimbs / L * (bfly_per_limb(c) + (F.sub ? 1.0 * c->N : 0.0)),
                     (double)nlimbs / L * (F.sub ? 3.0 : 2.0) * c->N * 8);
    LAUNCH_CHECK();
    return 0;
}

static int batch_transform(b200he_ctx *c, const b200he_batch *in, b200he_batch *out, bool inverse)
{
    if (!c || !in || !out) return fail("ntt: NULL argument");
    if (in->ctx != c || out->ctx != c) return fail("ntt: batch belongs to another context");
    if ((in->ntt != 0) != inverse) return fail("ntt: input is %s NTT form", in->ntt ? "already in" : "not in");
    CK(cudaSetDevice(c->device));
    OutBuf ob(out, (c->c > 0 || true) ? in : nullptr);
    TRY(ob.shape(in->count, in->size, in->L, inverse ? 0 : 1, in->scale));
    const size_t nl = (size_t)in->count * in->size * in->L, outer = (size_t)in->L * c->N;
    if (inverse) TRY(ntt_inv(c, in->d, ob.ptr(), nl, outer, outer, in->L, 0, INV_PLAIN));
    else TRY(ntt_fwd(c, in->d, ob.ptr(), nl, outer, outer, in->L, 0));
    ob.commit();
    return 0;
}
extern "C" int b200he_ntt_forward(b200he_ctx *c, const b200he_batch *in, b200he_batch *out) { return batch_transform(c, in, out, false); }
extern "C" int b200he_ntt_inverse(b200he_ctx *c, const b200he_batch *in, b200he_batch *out) { return batch_transform(c, in, out, true); }

// ------------------------------------------------------------------------------------ elementwise
static int check_pair(b200he_ctx *c, const b200he_batch *a, const b200he_batch *b, b200he_batch *out, const char *what)
{
    if (!c || !a || !b || !out) return fail("%s: NULL argument", what);
    if (a->ctx != c || b->ctx != c || out->ctx != c) return fail("%s: batch belongs to another context", what);
    if (a->L != b->L) return fail("%s: operands are at different levels (%d vs %d)", what, a->L, b->L);
    if (a->ntt != b->ntt) return fail("%s: operands differ in NTT form", what);
    return 0;
}

static int add_sub(b200he_ctx *c, const b200he_batch *a, const uint32_t *ai, const b200he_batch *b, const uint32_t *bi, uint64_t n,
                   b200he_batch *out, bool sub)
{
    const char *what = sub ? "sub" : "add";
    TRY(check_pair(c, a, b, out, what));
    if (c->scheme == B200HE_CKKS && !same_scale(a->scale, b->scale)) return fail("%s: scale mismatch", what);
    if ((!ai && n > a->count) || (!bi && n > b->count)) return fail("%s: n exceeds operand count", what);
    CK(cudaSetDevice(c->device));
    DevIdx da(c), db(c);
    TRY(da.set(ai, n, a->count, what));
    TRY(db.set(bi, n, b->count, what));
    const int smin = a->size < b->size ? a->size : b->size, smax = a->size < b->size ? b->size : a->size;
    if (sub && b->size > a->size) return fail("sub: size(b) > size(a) is not supported");   // before out is reshaped or anything is enqueued
    OutBuf ob(out, a, b);
    TRY(ob.shape(n, smax, a->L, a->ntt, a->scale));
    if (!n) { ob.commit(); return 0; }
    const size_t LN = (size_t)a->L * c->N;
    EwArgs A{};
    A.a = a->d; A.b = b->d; A.out = ob.ptr(); A.ai = da.d; A.bi = db.d;
    A.a_stride = a->ct_words(); A.b_stride = b->ct_words(); A.out_stride = smax * LN;
    A.polys = smin; A.b_polys = smin; A.L = a->L; A.mod_base = 0; A.n = n;
    const unsigned grid = blocks_for(n * smin * LN / 2);
    if (sub) LAUNCH(c, B200HE_KERN_ELEMENTWISE, k_ew<EW_SUB>, grid, 256, 0, c->T, A);
    else LAUNCH(c, B200HE_KERN_ELEMENTWISE, k_ew<EW_ADD>, grid, 256, 0, c->T, A);
    LAUNCH_CHECK();
    if (c->prof) c->work_bytes[B200HE_KERN_ELEMENTWISE] += (double)n * smin * 3.0 * LN * 8;
    if (smax > smin) {   // the longer operand's extra polynomials pass through (negated for b in a - b)
        const b200he_batch *big = a->size > b->size ? a : b;
        CopyArgs C{};
        C.src = big->d + smin * LN; C.dst = ob.ptr() + smin * LN; C.idx = (big == a) ? da.d : db.d;
        C.src_stride = big->ct_words(); C.dst_stride = smax * LN; C.polys = smax - smin; C.L_in = a->L; C.L_out = a->L; C.n = n;
        LAUNCH(c, B200HE_KERN_COPY, k_copy_limbs, blocks_for(n * (smax - smin) * LN / 2), 256, 0, c->T, C);
        LAUNCH_CHECK();
    }
    ob.commit();
    return 0;
}
extern "C" int b200he_add(b200he_ctx *c, const b200he_batch *a, const uint32_t *ai, const b200he_batch *b, const uint32_t *bi, uint64_t n, b200he_batch *out)
{
    return add_sub(c, a, ai, b, bi, n, out, false);
}
extern "C" int b200he_sub(b200he_ctx *c, const b200he_batch *a, const uint32_t *ai, const b200he_batch *b, const uint32_t *bi, uint64_t n, b200he_batch *out)
{
    return add_sub(c, a, ai, b, bi, n, out, true);
}

static int bfv_multiply(b200he_ctx *c, const b200he_batch *a, const u32 *dai, const b200he_batch *b, const u32 *dbi, uint64_t n, u64 *out);

extern "C" int b200he_multiply(b200he_ctx *c, const b200he_batch *a, const uint32_t *ai, const b200he_batch *b, const uint32_t *bi, uint64_t n, b200he_batch *out)
{
    TRY(check_pair(c, a, b, out, "multiply"));
    if (a->size != 2 || b->size != 2) return fail("multiply: operands must have size 2 (got %d, %d)", a->size, b->size);
    if ((!ai && n > a->count) || (!bi && n > b->count)) return fail("multiply: n exceeds operand count");
    CK(cudaSetDevice(c->device));
    DevIdx da(c), db(c);
    TRY(da.set(ai, n, a->count, "multiply"));
    TRY(db.set(bi, n, b->count, "multiply"));
    OutBuf ob(out, a, b);
    if (c->scheme == B200HE_CKKS) {
        if (!a->ntt) return fail("multiply: CKKS operands must be in NTT form");
        TRY(ob.shape(n, 3, a->L, 1, a->scale * b->scale));
        if (n) {
            EwArgs A{};
            A.a = a->d; A.b = b->d; A.out = ob.ptr(); A.ai = da.d; A.bi = db.d;
            A.a_stride = a->ct_words(); A.b_stride = b->ct_words(); A.out_stride = 3 * (size_t)a->L * c->N;
            A.polys = 2; A.b_polys = 2; A.L = a->L; A.mod_base = 0; A.n = n;
            LAUNCH(c, B200HE_KERN_TENSOR, k_tensor, blocks_for(n * (size_t)a->L * c->N / 2), 256, 0, c->T, A);
            LAUNCH_CHECK();
            for (int l = 0; l < a->L && c->prof; l++) work_add(c, B200HE_KERN_TENSOR, l, n * 4.0 * 1.5 * c->N, n * 7.0 * c->N * 8);
        }
    } else {
        if (a->ntt) return fail("multiply: BFV operands must be in coefficient form");
        if (a->L != (int)c->K - 1) return fail("multiply: BFV multiply is implemented at the top data level only (L=%d)", (int)c->K - 1);
        TRY(ob.shape(n, 3, a->L, 0, 1.0));
        if (n) TRY(bfv_multiply(c, a, da.d, b, db.d, n, ob.ptr()));
    }
    ob.commit();
    return 0;
}

// MatMult CipherBatchAxis inner loop: out[i][j] = sum_k a[i][k] (x) bt[j][k], size-3 results (kernels_ew.cuh k_tensor_mac)
extern "C" int b200he_matmul_accumulate(b200he_ctx *c, const b200he_batch *a, const b200he_batch *b, uint64_t rows, uint64_t inner, uint64_t cols,
                                        b200he_batch *out)
{
    TRY(check_pair(c, a, b, out, "matmul_accumulate"));
    if (c->scheme != B200HE_CKKS || !a->ntt) return fail("matmul_accumulate: CKKS operands in NTT form only (the BFV workload relinearizes every product)");
    if (a->size != 2 || b->size != 2) return fail("matmul_accumulate: operands must have size 2 (got %d, %d)", a->size, b->size);
    if (out == a || out == b) return fail("matmul_accumulate: out must not alias an input");
    if (inner == 0) return fail("matmul_accumulate: inner dimension is 0");
    if (rows * inner != a->count || inner * cols != b->count)
        return fail("matmul_accumulate: a holds %llu ciphertexts (want %llu x %llu), b holds %llu (want %llu x %llu)", (unsigned long long)a->count,
                    (unsigned long long)rows, (unsigned long long)inner, (unsigned long long)b->count, (unsigned long long)inner, (unsigned long long)cols);
    if (rows * cols >= (uint64_t(1) << 31)) return fail("matmul_accumulate: too many result cells");
    CK(cudaSetDevice(c->device));
    TRY(batch_shape(out, rows * cols, 3, a->L, 1, a->scale * b->scale));
    if (!rows || !cols) return 0;
    const size_t LN = (size_t)a->L * c->N;
    MacArgs A{};
    A.a = a->d; A.b = b->d; A.out = out->d;
    A.a_stride = a->ct_words(); A.b_stride = b->ct_words(); A.out_stride = 3 * LN;
    A.L = a->L; A.rows = (u32)rows; A.inner = (u32)inner; A.cols = (u32)cols;
    // a term adds two products below q^2 to the middle accumulator: runs of floor(2^127 / q^2) terms fit 128 bits
    u64 run = ~u64(0);
    for (int l = 0; l < a->L; l++) {
        const hm::u128 q2 = (hm::u128)c->mods[l].q * c->mods[l].q;
        const u64 r = (u64)((((hm::u128)1) << 127) / q2);
        if (r < run) run = r;
    }
    A.reduce_every = (u32)(run < 1 ? 1 : run > 0x7fffffff ? 0x7fffffff : run);
    // MacAcc's odd part gains four products below 2^32 * ceil(q / 2^32) per step of the k loop (the middle polynomial) and holds 64 bits
    u64 flush = 1u << 20;
    for (int l = 0; l < a->L; l++) {
        if (c->mods[l].dp) continue;
        const hm::u128 per_step = (hm::u128)4 * ((hm::u128)1 << 32) * ((c->mods[l].q >> 32) + 1);
        const u64 f = (u64)((((hm::u128)1) << 64) / per_step);
        if (f < flush) flush = f;
    }
    if (flush < 1) return fail("matmul_accumulate: modulus too wide for the integer accumulators");
    A.flush_every = (u32)(flush < A.reduce_every ? flush : A.reduce_every);
    // 2 x 2 tiles.  Measured against 1 x 2 (80 registers, six resident blocks instead of four, half the operand reuse):
    // 28.0 vs 30.6 ms on the 48 x 48 x 48 probe -- the kernel is bound by FP64 / IMAD issue, not by latency.
    constexpr int TI = 2, TJ = 2;
    const u64 ti = (rows + TI - 1) / TI, tj = (cols + TJ - 1) / TJ, cblocks = (LN + B200HE_MAC_THREADS - 1) / B200HE_MAC_THREADS;
    if (ti * tj * cblocks >= (u64(1) << 31)) return fail("matmul_accumulate: grid too large");
    A.tiles_j = (u32)tj;
    A.ntiles = (u32)(ti * tj);
    LAUNCH(c, B200HE_KERN_TENSOR_MAC, (k_tensor_mac<TI, TJ>), (unsigned)(ti * tj * cblocks), B200HE_MAC_THREADS, 0, c->T, A);
    LAUNCH_CHECK();
    if (c->prof) {   // 4 multiply-accumulates per coefficient and term (0.5 butterfly-equivalents each) on the limb's pipe; every operand read once, results written once
        for (int l = 0; l < a->L; l++) (c->mods[l].dp ? c->work_dp : c->work_int)[B200HE_KERN_TENSOR_MAC] += (double)rows * cols * inner * 4.0 * 0.5 * c->N;
        c->work_bytes[B200HE_KERN_TENSOR_MAC] += ((double)(rows + cols) * inner * 2.0 + (double)rows * cols * 3.0) * LN * 8;
    }
    return 0;
}

static int plain_op(b200he_ctx *c, const b200he_batch *ct, const b200he_batch *pl, const uint32_t *pi, b200he_batch *out, bool mul)
{
    const char *what = mul ? "multiply_plain" : "add_plain";
    TRY(check_pair(c, ct, pl, out, what));
    if (pl->size != 1) return fail("%s: plaintext batch must have size 1", what);
    if (c->scheme != B200HE_CKKS || !ct->ntt) return fail("%s: implemented for CKKS (NTT form) only, as used by the reference", what);
    if (!mul && !same_scale(ct->scale, pl->scale)) return fail("add_plain: scale mismatch");
    const uint64_t n = ct->count;
    std::vector<u32> bcast;
    if (!pi && pl->count != n) {
        if (pl->count != 1) return fail("%s: plaintext count %llu does not match ciphertext count %llu", what, (unsigned long long)pl->count, (unsigned long long)n);
        bcast.assign(n, 0);
        pi = bcast.data();
    }
    CK(cudaSetDevice(c->device));
    DevIdx dp(c);
    TRY(dp.set(pi, n, pl->count, what));
    OutBuf ob(out, ct, pl);
    TRY(ob.shape(n, ct->size, ct->L, 1, mul ? ct->scale * pl->scale : ct->scale));
    if (n) {
        EwArgs A{};
        A.a = ct->d; A.b = pl->d; A.out = ob.ptr(); A.ai = nullptr; A.bi = dp.d;
        A.a_stride = ct->ct_words(); A.b_stride = pl->ct_words(); A.out_stride = ct->ct_words();
        A.polys = ct->size; A.b_polys = 1; A.L = ct->L; A.mod_base = 0; A.n = n;
        const unsigned grid = blocks_for(n * ct->ct_words() / 2);
        if (mul) LAUNCH(c, B200HE_KERN_ELEMENTWISE, k_ew<EW_MUL>, grid, 256, 0, c->T, A);
        else LAUNCH(c, B200HE_KERN_ELEMENTWISE, k_ew<EW_ADD>, grid, 256, 0, c->T, A);
        LAUNCH_CHECK();
        if (c->prof) c->work_bytes[B200HE_KERN_ELEMENTWISE] += (double)n * (mul ? 3.0 * ct->size : 2.0 * ct->size + 1.0) * ct->L * c->N * 8;
    }
    ob.commit();
    return 0;
}
extern "C" int b200he_multiply_plain(b200he_ctx *c, const b200he_batch *ct, const b200he_batch *pl, const uint32_t *pi, b200he_batch *out)
{
    return plain_op(c, ct, pl, pi, out, true);
}
extern "C" int b200he_add_plain(b200he_ctx *c, const b200he_batch *ct, const b200he_batch *pl, const uint32_t *pi, b200he_batch *out)
{
    return plain_op(c, ct, pl, pi, out, false);
}

static int copy_limbs(b200he_ctx *c, const u64 *src, size_t src_stride, u64 *dst, size_t dst_stride, const u32 *didx, uint64_t n, int polys,
                      int L_in, int L_out)
{
    if (!n) return 0;
    CopyArgs C{};
    C.src = src; C.dst = dst; C.idx = didx; C.src_stride = src_stride; C.dst_stride = dst_stride;
    C.polys = polys; C.L_in = L_in; C.L_out = L_out; C.n = n;
    LAUNCH(c, B200HE_KERN_COPY, k_copy_limbs, blocks_for(n * polys * (size_t)L_out * c->N / 2), 256, 0, c->T, C);
    LAUNCH_CHECK();
    if (c->prof) c->work_bytes[B200HE_KERN_COPY] += (double)n * polys * 2.0 * L_out * c->N * 8;
    return 0;
}

extern "C" int b200he_mod_drop(b200he_ctx *c, const b200he_batch *in, int L_target, b200he_batch *out)
{
    if (!c || !in || !out) return fail("mod_drop: NULL argument");
    if (in->ctx != c || out->ctx != c) return fail("mod_drop: batch belongs to another context");
    if (L_target < 1 || L_target > in->L) return fail("mod_drop: target level %d not in [1,%d]", L_target, in->L);
    if (c->scheme != B200HE_CKKS) return fail("mod_drop: CKKS only (BFV mod-switch rescales: use rescale_to_next)");
    CK(cudaSetDevice(c->device));
    OutBuf ob(out, in);
    TRY(ob.shape(in->count, in->size, L_target, in->ntt, in->scale));
    TRY(copy_limbs(c, in->d, in->ct_words(), ob.ptr(), (size_t)in->size * L_target * c->N, nullptr, in->count, in->size, in->L, L_target));
    ob.commit();
    return 0;
}

extern "C" int b200he_gather(b200he_ctx *c, const b200he_batch *in, const uint32_t *idx, uint64_t n, b200he_batch *out)
{
    if (!c || !in || !out) return fail("gather: NULL argument");
    if (in->ctx != c || out->ctx != c) return fail("gather: batch belongs to another context");
    if (!idx && n > in->count) return fail("gather: n exceeds count");
    CK(cudaSetDevice(c->device));
    DevIdx di(c);
    TRY(di.set(idx, n, in->count, "gather"));
    OutBuf ob(out, in);
    TRY(ob.shape(n, in->size, in->L, in->ntt, in->scale));
    TRY(copy_limbs(c, in->d, in->ct_words(), ob.ptr(), in->ct_words(), di.d, n, in->size, in->L, in->L));
    ob.commit();
    return 0;
}

// ------------------------------------------------------------------------------------ key switching (K6)
// out[b] = (add0[b], add1[b]) + KeySwitch(target[b])  for b < B, at level L.
//   target: [L][N] per ciphertext (stride t_stride), NTT form (CKKS) or coefficient form (BFV)
//   add0/add1: per-ciphertext addends (stride add_stride) or nullptr (zero)
//   out: [2][L][N] per ciphertext (stride out_stride); out may alias add0/add1 element-for-element
//
// rescale = true (CKKS, add0/add1 required, L >= 2): out[b] = rescale_to_next((add0[b], add1[b]) + KeySwitch(target[b])),
// [2][L-1][N] per ciphertext -- SEAL's relinearize_inplace followed by rescale_to_next_inplace, bit for bit, with
// 2 + 2(L-1) transforms where the two separate calls need 2L + 2 + 2(L-1): every step is exact arithmetic in Z_q and
// the transform is linear, so the last limb's mod-down correction is applied in coefficient form (InvFuse) and the two
// corrections of each remaining limb share one transform (PreTwo).
//
// gal (CKKS): the Galois automorphism g of apply_galois_inplace fused into the switch -- `target` is the
// un-permuted c1 and the kernels gather g(c1) on load; out[b] additionally receives (g(gal_src[b]), 0), gal_src = c0.
struct GalFuse {
    const b200he_ctx::GalTab *tab;
    const u64 *src;
    size_t src_stride;
};
static int key_switch(b200he_ctx *c, int L, size_t B, const u64 *target, size_t t_stride, const u64 *key, const u64 *add0, const u64 *add1,
                      size_t add_stride, u64 *out, size_t out_stride, bool rescale = false, const GalFuse *gal = nullptr)
{
    const size_t N = c->N, K = c->K;
    const bool ckks = c->scheme == B200HE_CKKS;
    if (rescale && (!ckks || !add0 || !add1 || L < 2)) return fail("key_switch: fused rescale needs CKKS, both addends and L >= 2");
    if (gal && (!ckks || rescale)) return fail("key_switch: the fused Galois permutation is the NTT-form one (CKKS), without rescale");
    // scratch per ciphertext: tcoef [L][N] (CKKS), acc [2][L+1][N], rp [2][N] (+ rp2 [2][N] for the fused rescale)
    const size_t w_t = ckks ? (size_t)L * N : 0, w_acc = 2 * (size_t)(L + 1) * N, w_rp = rescale ? 4 * N : 2 * N;
    const size_t per_ct = (w_t + w_acc + w_rp) * 8;
    size_t chunk = c->workspace / per_ct;
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    u64 *ws = (u64 *)c->pool.get(chunk * per_ct);
    if (!ws) return fail("key_switch: out of device memory for %zu bytes of workspace", chunk * per_ct);
    u64 *tcoef = ws, *acc = ws + chunk * w_t, *rp = acc + chunk * w_acc;
    int rc = 0;
    for (size_t b0 = 0; b0 < B && !rc; b0 += chunk) {
        const size_t nb = (B - b0 < chunk) ? B - b0 : chunk;
        const u64 *tg = target + b0 * t_stride;
        KsInnerArgs A{};
        if (ckks) {
            InvFuse G{};
            G.gal = gal ? gal->tab->d : nullptr;
            rc = ntt_inv(c, tg, tcoef, nb * L, t_stride, (size_t)L * N, L, 0, INV_PLAIN, gal ? &G : nullptr);
            if (rc) break;
            A.tcoef = tcoef; A.tcoef_stride = (size_t)L * N; A.target = tg; A.target_stride = t_stride;
        } else {
            A.tcoef = tg; A.tcoef_stride = t_stride; A.target = nullptr; A.target_stride = 0;
        }
        A.key = key; A.acc = acc; A.rp = rp; A.L = L; A.K = (int)K; A.B = (int)nb;
        A.gal = gal ? gal->tab->d : nullptr;
        if (rescale) {   // the last data limb leaves k_ks_inner as acc * s + c (the value the rescale's inverse transform rounds)
            A.fuse_add = add0 + b0 * add_stride;
            A.fuse_ct_stride = add_stride;
            A.fuse_poly_stride = (size_t)(add1 - add0);
        }
        // (the rounded special-prime limb comes out of k_ks_inner in coefficient form: rp)
        PROF(c, B200HE_KERN_KS_INNER, launch_ks_inner(geo(c), c->T, A, nb * (L + 1)));
        if (c->prof) {
            // output modulus I < L: L - 1 forward transforms (CKKS: the I == J digit is reused in NTT form; BFV: L) and
            // 2 L N multiply-accumulates; special prime: L forward + 2 inverse transforms and the same inner product.
            // Traffic: target in coefficient and NTT form in, 2 L accumulator limbs + 2 rounded limbs out; the key once
            // per launch (16 B per coefficient with its Shoup quotient, 8 B for FP64-domain limbs).
            const double bf = bfly_per_limb(c), Nd = (double)N;
            for (int I = 0; I <= L; I++) {
                const int ki = I == L ? (int)K - 1 : I;
                const double ntts = I == L ? L + 2.0 : (ckks ? L - 1.0 : (double)L);
                work_add(c, B200HE_KERN_KS_INNER, ki, nb * (ntts * bf + 2.0 * L * Nd), 2.0 * L * (c->mods[ki].dp ? 1.0 : 2.0) * Nd * 8);
            }
            work_add(c, B200HE_KERN_KS_INNER, 0, 0, nb * (2.0 * L + 2.0 * (L + 1)) * Nd * 8);
            if (rescale) work_add(c, B200HE_KERN_KS_INNER, L - 1, nb * 2.0 * Nd, nb * 2.0 * Nd * 8);   // y = acc * s + c on the last data limb
        }
        if (cudaGetLastError() != cudaSuccess) { rc = fail("key_switch: k_ks_inner launch failed"); break; }
        ModDownArgs D{};
        D.rp = rp; D.base = acc; D.base_ct_stride = w_acc; D.base_poly_stride = (size_t)(L + 1) * N;
        D.addend[0] = add0 ? add0 + b0 * add_stride : nullptr;
        D.addend[1] = add1 ? add1 + b0 * add_stride : nullptr;
        D.add_ct_stride = add_stride;
        D.out = out + b0 * out_stride; D.out_ct_stride = out_stride; D.out_poly_stride = (size_t)L * N;
        D.P = 2; D.nJ = L; D.x = (int)K - 1;
        if (gal) {
            D.gal = gal->tab->d; D.gal_src = gal->src + b0 * gal->src_stride; D.gal_ct_stride = gal->src_stride;
            memcpy(D.gal_chunk, gal->tab->chunk, sizeof D.gal_chunk);
        }
        if (rescale) {
            // last data limb: rp2 = iNTT(acc * s + addend) - u1 * s + q/2   (rounded last limb of the switched ciphertext;
            // acc * s + addend is what k_ks_inner left in the accumulator's limb L-1)
            u64 *rp2 = rp + nb * 2 * N;
            InvFuse F{};
            F.sub = rp; F.x = (int)K - 1;
            rc = ntt_inv(c, acc + (size_t)(L - 1) * N, rp2, nb * 2, (size_t)(L + 1) * N, N, 1, L - 1, INV_ADDHALF, &F);
            if (rc) break;
            D.rp2 = rp2; D.x2 = L - 1; D.nJ = L - 1; D.out_poly_stride = (size_t)(L - 1) * N;
            launch_moddown(c, D, nb * 2);
        } else if (ckks) {
            launch_moddown(c, D, nb * 2);
        } else {
            // BFV: accumulators back to coefficient form (in place, unsplit per-limb strides), then elementwise mod-down
            rc = ntt_inv(c, acc, acc, nb * 2 * L, (size_t)(L + 1) * N, (size_t)(L + 1) * N, L, 0, INV_PLAIN);
            if (rc) break;
            LAUNCH(c, B200HE_KERN_MODDOWN, k_moddown_coeff, blocks_for(nb * 2 * L * N / 2), 256, 0, c->T, D, nb);
        }
        if (cudaGetLastError() != cudaSuccess) { rc = fail("key_switch: mod-down launch failed"); break; }
    }
    c->pool.put(ws);
    return rc;
}

extern "C" int b200he_relinearize(b200he_ctx *c, const b200he_batch *in, b200he_batch *out)
{
    if (!c || !in || !out) return fail("relinearize: NULL argument");
    if (in->ctx != c || out->ctx != c) return fail("relinearize: batch belongs to another context");
    CK(cudaSetDevice(c->device));
    if (in->size == 2) {   // SEAL: nothing to do (R/src/engine/seal_context.cpp:390)
        if (out == in) return 0;
        TRY(batch_shape(out, in->count, 2, in->L, in->ntt, in->scale));
        return copy_limbs(c, in->d, in->ct_words(), out->d, in->ct_words(), nullptr, in->count, 2, in->L, in->L);
    }
    if (in->size != 3) return fail("relinearize: ciphertext size must be 2 or 3");
    if (!c->relin) return fail("relinearize: no relinearization key uploaded");
    if ((c->scheme == B200HE_CKKS) != (in->ntt != 0)) return fail("relinearize: wrong NTT form for scheme");
    OutBuf ob(out, in);
    TRY(ob.shape(in->count, 2, in->L, in->ntt, in->scale));
    const size_t LN = (size_t)in->L * c->N;
    TRY(key_switch(c, in->L, in->count, in->d + 2 * LN, 3 * LN, c->relin, in->d, in->d + LN, 3 * LN, ob.ptr(), 2 * LN));
    ob.commit();
    return 0;
}

// relinearize_inplace immediately followed by rescale_to_next_inplace, as every multiply of the CKKS matmul and
// logistic-regression workloads does (R/src/benchmarks/ckks/seal_ckks_matmultval_benchmark.cpp:252-255,
// R/src/engine/seal_context.cpp:390-391,447-448): same bits as the two calls, fewer transforms (see key_switch).
extern "C" int b200he_relinearize_rescale(b200he_ctx *c, const b200he_batch *in, b200he_batch *out)
{
    if (!c || !in || !out) return fail("relinearize_rescale: NULL argument");
    if (in->ctx != c || out->ctx != c) return fail("relinearize_rescale: batch belongs to another context");
    if (c->scheme != B200HE_CKKS || in->size != 3) {   // BFV, or nothing to relinearize: the two plain calls
        TRY(b200he_relinearize(c, in, out));
        return b200he_rescale_to_next(c, out, out);
    }
    if (in->L < 2) return fail("rescale: already at the last level");
    if (!c->relin) return fail("relinearize: no relinearization key uploaded");
    if (!in->ntt) return fail("relinearize: wrong NTT form for scheme");
    CK(cudaSetDevice(c->device));
    const int L = in->L;
    OutBuf ob(out, in);
    TRY(ob.shape(in->count, 2, L - 1, 1, in->scale / (double)c->mods[L - 1].q));
    const size_t LN = (size_t)L * c->N;
    if (in->count)
        TRY(key_switch(c, L, in->count, in->d + 2 * LN, 3 * LN, c->relin, in->d, in->d + LN, 3 * LN, ob.ptr(), 2 * (size_t)(L - 1) * c->N, true));
    ob.commit();
    return 0;
}

// ------------------------------------------------------------------------------------ Galois (K8)
// out = apply_galois(in, elt), or in + apply_galois(in, elt) when add_input (the "rotated = rotate(retval); retval += rotated"
// step of accumulateCKKS / accumulateBFV, R/src/engine/seal_context.cpp:302-303,337-338, without a separate add pass:
// g(c0) + c0 comes out of the permutation kernel and c1 rides along as the key switch's second addend)
static int apply_galois_impl(b200he_ctx *c, const b200he_batch *in, uint32_t elt, b200he_batch *out, bool add_input)
{
    if (!c || !in || !out) return fail("apply_galois: NULL argument");
    if (in->ctx != c || out->ctx != c) return fail("apply_galois: batch belongs to another context");
    if (in->size != 2) return fail("apply_galois: ciphertext size must be 2");
    if (!(elt & 1) || elt >= 2 * c->N) return fail("apply_galois: invalid Galois element %u", elt);
    auto kit = c->gal.find(elt);
    if (kit == c->gal.end()) return fail("apply_galois: no Galois key for element %u", elt);
    const bool ckks = c->scheme == B200HE_CKKS;
    if (ckks != (in->ntt != 0)) return fail("apply_galois: wrong NTT form for scheme");
    CK(cudaSetDevice(c->device));
    OutBuf ob(out, in);
    TRY(ob.shape(in->count, 2, in->L, in->ntt, in->scale));
    const size_t B = in->count, LN = (size_t)in->L * c->N;
    if (!B) { ob.commit(); return 0; }
    if (ckks) {
        // NTT form: no permutation pass.  The inverse transform of the target, the I == J term of the inner product and the
        // addend of the mod-down gather through the element's table; c0 / c1 of the input ride along as plain addends for
        // the rotate-and-add of accumulate (out = in + apply_galois(in)).
        const b200he_ctx::GalTab *tab = nullptr;
        TRY(galois_table(c, elt, &tab));
        GalFuse G{ tab, in->d, 2 * LN };
        TRY(key_switch(c, in->L, B, in->d + LN, 2 * LN, kit->second, add_input ? in->d : nullptr, add_input ? in->d + LN : nullptr, 2 * LN, ob.ptr(),
                       2 * LN, false, &G));
        ob.commit();
        return 0;
    }
    // BFV: coefficient-form automorphism (signed index map) into out / a scratch target, then the key switch
    u64 *g1 = (u64 *)c->pool.get(B * LN * 8);
    if (!g1) return fail("apply_galois: out of device memory");
    GaloisArgs G{};
    G.src = in->d; G.dst0 = ob.ptr(); G.dst1 = g1; G.src_stride = 2 * LN; G.dst_stride = 2 * LN;
    G.elt = elt; G.L = in->L; G.logn = c->logn; G.B = B; G.add_input = add_input ? 1 : 0;
    LAUNCH(c, B200HE_KERN_GALOIS, k_galois_coeff, blocks_for(B * 2 * LN), 256, 0, c->T, G);
    int rc = 0;
    if (cudaGetLastError() != cudaSuccess) rc = fail("apply_galois: launch failed");
    if (!rc) {
        if (add_input) rc = key_switch(c, in->L, B, g1, LN, kit->second, ob.ptr(), in->d + LN, 2 * LN, ob.ptr(), 2 * LN);
        else rc = key_switch(c, in->L, B, g1, LN, kit->second, ob.ptr(), nullptr, 2 * LN, ob.ptr(), 2 * LN);
    }
    c->pool.put(g1);
    if (rc) return rc;
    ob.commit();
    return 0;
}
extern "C" int b200he_apply_galois(b200he_ctx *c, const b200he_batch *in, uint32_t elt, b200he_batch *out)
{
    return apply_galois_impl(c, in, elt, out, false);
}

// SEAL GaloisTool::get_elt_from_step
static u32 elt_from_step(const b200he_ctx *c, int step)
{
    const u32 m = 2 * c->N, half = c->N / 2;
    if (step == 0) return m - 1;
    const u32 pos = (u32)(step < 0 ? -step : step);
    if (pos >= half) return 0;
    const u32 e = step < 0 ? half - pos : pos;
    u64 g = 1;
    for (u32 i = 0; i < e; i++) g = (g * 3) & (m - 1);
    return (u32)g;
}

static int rotate_rec(b200he_ctx *c, const b200he_batch *in, int step, b200he_batch *out)
{
    const u32 elt = elt_from_step(c, step);
    if (!elt) return fail("rotate: step %d out of range", step);
    if (c->gal.count(elt)) return b200he_apply_galois(c, in, elt, out);
    // SEAL Evaluator::rotate_internal: non-adjacent form of the step, least significant term first
    std::vector<int> terms;
    {
        int v = step < 0 ? -step : step;
        for (int i = 0; v; i++) {
            const int z = (v & 1) ? 2 - (v & 3) : 0;
            v = (v - z) >> 1;
            if (z) terms.push_back((step < 0 ? -z : z) * (1 << i));
        }
    }
    if (terms.size() == 1) return fail("rotate: Galois key for step %d missing", step);
    const b200he_batch *cur = in;
    for (int term : terms) {
        if ((u32)(term < 0 ? -term : term) == c->N / 2) continue;
        TRY(rotate_rec(c, cur, term, out));
        cur = out;
    }
    if (cur == in && out != in) return b200he_gather(c, in, nullptr, in->count, out);
    return 0;
}

// rotate ciphertext i by steps[i] (the collapse step of logistic regression rotates sample i by -i,
// R/src/engine/seal_context.cpp:378).  Every ciphertext goes through exactly the key switches SEAL's
// rotate_vector would apply to it -- the non-adjacent form of its own step, least significant term first --
// but ciphertexts that share a term are switched together: for each power of two and sign, gather the
// ciphertexts whose decomposition holds that term, apply the Galois element once, scatter back.
extern "C" int b200he_rotate_each(b200he_ctx *c, const b200he_batch *in, const int32_t *steps, b200he_batch *out)
{
    if (!c || !in || !out || !steps) return fail("rotate_each: NULL argument");
    if (in->ctx != c || out->ctx != c) return fail("rotate_each: batch belongs to another context");
    if (in->size != 2) return fail("rotate_each: ciphertext size must be 2");
    CK(cudaSetDevice(c->device));
    const uint64_t n = in->count;
    const int half = (int)(c->N / 2);
    // term lists: terms[k][sign] = ciphertexts whose step has +-2^k in its decomposition
    std::vector<std::vector<u32>> plus(32), minus(32);
    for (uint64_t i = 0; i < n; i++) {
        const int step = steps[i];
        if (step == 0) continue;
        if (step <= -half || step >= half) return fail("rotate_each: step %d of ciphertext %llu out of range", step, (unsigned long long)i);
        const u32 elt = elt_from_step(c, step);
        if (c->gal.count(elt)) {   // a key for the whole step: single switch, like SEAL
            int k = 0;
            const int a = step < 0 ? -step : step;
            if ((a & (a - 1)) == 0) {
                while ((1 << k) < a) k++;
                (step < 0 ? minus : plus)[k].push_back((u32)i);
                continue;
            }
            return fail("rotate_each: steps with a dedicated non-power-of-two key are not supported");
        }
        int v = step < 0 ? -step : step, nterms = 0;
        for (int k = 0; v; k++) {
            const int z = (v & 1) ? 2 - (v & 3) : 0;
            v = (v - z) >> 1;
            if (!z) continue;
            nterms++;
            if ((1 << k) == half) continue;
            const bool neg = (step < 0) != (z < 0);
            (neg ? minus : plus)[k].push_back((u32)i);
        }
        if (nterms == 1) return fail("rotate_each: Galois key for step %d missing", step);
    }
    if (out != in) TRY(b200he_gather(c, in, nullptr, n, out));
    b200he_batch *sub = nullptr, *rot = nullptr;
    TRY(b200he_batch_create(c, &sub));
    TRY(b200he_batch_create(c, &rot));
    int rc = 0;
    for (int k = 0; k < 32 && !rc; k++)
        for (int sgn = 0; sgn < 2 && !rc; sgn++) {
            const std::vector<u32> &ids = sgn ? minus[k] : plus[k];
            if (ids.empty()) continue;
            const int step = sgn ? -(1 << k) : (1 << k);
            const u32 elt = elt_from_step(c, step);
            if (!c->gal.count(elt)) { rc = fail("rotate_each: Galois key for step %d missing", step); break; }
            rc = b200he_gather(c, out, ids.data(), ids.size(), sub);
            if (!rc) rc = b200he_apply_galois(c, sub, elt, rot);
            if (!rc) {
                DevIdx di(c);
                rc = di.set(ids.data(), ids.size(), n, "rotate_each");
                if (!rc) {
                    CopyArgs C{};
                    C.src = rot->d; C.dst = out->d; C.idx = nullptr; C.dst_idx = di.d;
                    C.src_stride = rot->ct_words(); C.dst_stride = out->ct_words(); C.polys = 2; C.L_in = in->L; C.L_out = in->L; C.n = ids.size();
                    LAUNCH(c, B200HE_KERN_COPY, k_copy_limbs, blocks_for(ids.size() * out->ct_words() / 2), 256, 0, c->T, C);
                    if (cudaGetLastError() != cudaSuccess) rc = fail("rotate_each: scatter launch failed");
                }
            }
        }
    b200he_batch_destroy(sub);
    b200he_batch_destroy(rot);
    return rc;
}

// out (1 ciphertext) = sum of all ciphertexts of in
extern "C" int b200he_sum(b200he_ctx *c, const b200he_batch *in, b200he_batch *out)
{
    if (!c || !in || !out) return fail("sum: NULL argument");
    if (in->ctx != c || out->ctx != c) return fail("sum: batch belongs to another context");
    if (in->count == 0) return fail("sum: empty batch");
    CK(cudaSetDevice(c->device));
    OutBuf ob(out, in);
    TRY(ob.shape(1, in->size, in->L, in->ntt, in->scale));
    LAUNCH(c, B200HE_KERN_ELEMENTWISE, k_batch_sum, blocks_for(in->ct_words() / 2), 256, 0, c->T, in->d, ob.ptr(), (size_t)in->count, in->size, in->L);
    LAUNCH_CHECK();
    ob.commit();
    return 0;
}

extern "C" int b200he_rotate(b200he_ctx *c, const b200he_batch *in, int step, b200he_batch *out)
{
    if (!c || !in || !out) return fail("rotate: NULL argument");
    if (in->ctx != c || out->ctx != c) return fail("rotate: batch belongs to another context");
    if (in->size != 2) return fail("rotate: ciphertext size must be 2");
    if (step == 0) return out == in ? 0 : b200he_gather(c, in, nullptr, in->count, out);
    return rotate_rec(c, in, step, out);
}
extern "C" int b200he_rotate_columns(b200he_ctx *c, const b200he_batch *in, b200he_batch *out)
{
    if (!c) return fail("rotate_columns: NULL argument");
    return b200he_apply_galois(c, in, 2 * c->N - 1, out);
}

// SEALContextWrapper::accumulateCKKS / accumulateBFV (R/src/engine/seal_context.cpp:289-347)
extern "C" int b200he_accumulate(b200he_ctx *c, b200he_batch *io, uint64_t count)
{
    if (!c || !io) return fail("accumulate: NULL argument");
    if (io->ctx != c) return fail("accumulate: batch belongs to another context");
    const bool ckks = c->scheme == B200HE_CKKS;
    const uint64_t slots = c->N / 2;
    if (count == 0) return fail("accumulate: count == 0 (the reference substitutes a fresh encryption of zero; do that on the host)");
    uint64_t rows = count > slots ? slots : count;
    int rot = 0;
    while ((uint64_t(1) << rot) < rows) rot++;
    b200he_batch *tmp = nullptr;
    TRY(b200he_batch_create(c, &tmp));
    int rc = 0;
    if (io->size != 2) { b200he_batch_destroy(tmp); return fail("rotate: ciphertext size must be 2"); }
    for (int k = 0; k < rot && !rc; k++) {
        const u32 elt = elt_from_step(c, 1 << k);
        if (elt && c->gal.count(elt)) rc = apply_galois_impl(c, io, elt, io, true);   // io += rotate(io, 2^k) in one key switch
        else {   // no key for this power of two: SEAL's NAF fallback, then the add
            rc = b200he_rotate(c, io, 1 << k, tmp);
            if (!rc) rc = b200he_add(c, io, nullptr, tmp, nullptr, io->count, io);
        }
    }
    if (!rc && !ckks && count > slots) {
        const u32 elt = 2 * c->N - 1;
        if (c->gal.count(elt)) rc = apply_galois_impl(c, io, elt, io, true);
        else rc = fail("apply_galois: no Galois key for element %u", elt);
    }
    b200he_batch_destroy(tmp);
    return rc;
}

// ------------------------------------------------------------------------------------ rescale (K9)
extern "C" int b200he_rescale_to_next(b200he_ctx *c, const b200he_batch *in, b200he_batch *out)
{
    if (!c || !in || !out) return fail("rescale: NULL argument");
    if (in->ctx != c || out->ctx != c) return fail("rescale: batch belongs to another context");
    if (in->L < 2) return fail("rescale: already at the last level");
    const bool ckks = c->scheme == B200HE_CKKS;
    if (ckks != (in->ntt != 0)) return fail("rescale: wrong NTT form for scheme");
    CK(cudaSetDevice(c->device));
    const int L = in->L, P = in->size;
    const size_t N = c->N, B = in->count;
    OutBuf ob(out, in);
    TRY(ob.shape(B, P, L - 1, in->ntt, ckks ? in->scale / (double)c->mods[L - 1].q : in->scale));
    if (!B) { ob.commit(); return 0; }
    size_t chunk = c->workspace / (P * N * 8);
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    u64 *rp = (u64 *)c->pool.get(chunk * P * N * 8);
    if (!rp) return fail("rescale: out of device memory");
    int rc = 0;
    for (size_t b0 = 0; b0 < B && !rc; b0 += chunk) {
        const size_t nb = B - b0 < chunk ? B - b0 : chunk;
        const u64 *src = in->d + b0 * in->ct_words();
        ModDownArgs D{};
        D.base = src; D.base_ct_stride = in->ct_words(); D.base_poly_stride = (size_t)L * N;
        D.addend[0] = D.addend[1] = nullptr; D.add_ct_stride = 0;
        D.out = ob.ptr() + b0 * (size_t)P * (L - 1) * N; D.out_ct_stride = (size_t)P * (L - 1) * N; D.out_poly_stride = (size_t)(L - 1) * N;
        D.P = P; D.nJ = L - 1; D.x = L - 1;
        if (ckks) {
            rc = ntt_inv(c, src + (size_t)(L - 1) * N, rp, nb * P, (size_t)L * N, N, 1, L - 1, INV_ADDHALF);
            if (rc) break;
            D.rp = rp;
            launch_moddown(c, D, nb * P);
        } else {
            // coefficient form: rp = last + q_last/2 mod q_last, elementwise
            D.rp = nullptr; D.rp_raw = src + (size_t)(L - 1) * N; D.rp_raw_stride = (size_t)L * N;
            LAUNCH(c, B200HE_KERN_MODDOWN, k_moddown_coeff, blocks_for(nb * P * (L - 1) * N / 2), 256, 0, c->T, D, nb);
        }
        if (cudaGetLastError() != cudaSuccess) rc = fail("rescale: launch failed");
    }
    c->pool.put(rp);
    if (rc) return rc;
    ob.commit();
    return 0;
}

// ------------------------------------------------------------------------------------ BFV multiply (K5)
#include "behz_host.inl"

// ------------------------------------------------------------------------------------ measurement hooks
extern "C" uint64_t b200he_launch_count(const b200he_ctx *c) { return c ? c->launches : 0; }
extern "C" int b200he_profile_begin(b200he_ctx *c)
{
    if (!c) return fail("profile_begin: ctx NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (auto &r : c->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    c->recs.clear();
    for (int i = 0; i < B200HE_KERN_COUNT; i++) c->work_int[i] = c->work_dp[i] = c->work_bytes[i] = 0;
    c->prof = true;
    return 0;
}
extern "C" int b200he_profile_work(const b200he_ctx *c, double *bfly_int, double *bfly_fp64, double *bytes)
{
    if (!c || !bfly_int || !bfly_fp64 || !bytes) return fail("profile_work: NULL argument");
    for (int i = 0; i < B200HE_KERN_COUNT; i++) {
        bfly_int[i] = c->work_int[i];
        bfly_fp64[i] = c->work_dp[i];
        bytes[i] = c->work_bytes[i];
    }
    return 0;
}
extern "C" int b200he_profile_end(b200he_ctx *c, double *ms, uint64_t *launches)
{
    if (!c || !ms || !launches) return fail("profile_end: NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->prof = false;
    for (int i = 0; i < B200HE_KERN_COUNT; i++) { ms[i] = 0; launches[i] = 0; }
    for (auto &r : c->recs) {
        float t = 0;
        cudaEventElapsedTime(&t, r.a, r.b);
        ms[r.cls] += t;
        launches[r.cls]++;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    c->recs.clear();
    return 0;
}
extern "C" const char *b200he_kernel_name(int k)
{
    static const char *names[B200HE_KERN_COUNT] = { "k_ntt_fwd", "k_ntt_inv", "k_tensor_mac", "k_ks_inner", "k_moddown",
                                                    "k_ew", "k_tensor", "k_galois", "k_copy_limbs", "k_behz" };
    return (k >= 0 && k < B200HE_KERN_COUNT) ? names[k] : "?";
}
