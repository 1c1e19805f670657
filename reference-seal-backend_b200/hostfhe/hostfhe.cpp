// hostfhe.cpp -- host-side keygen / encode / encrypt / decrypt / decode (see hostfhe.h).
// Stand-in for the host-resident SEAL objects of SEALContextWrapper
// (R/src/engine/seal_context.cpp:46-70).  No Evaluator arithmetic lives here.
#include "hostfhe.h"

#include <sys/random.h>

#include <atomic>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

namespace {
typedef uint64_t u64;
typedef unsigned __int128 u128;

inline u64 mulm(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
inline u64 addm(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
inline u64 subm(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
u64 powm(u64 a, u64 e, u64 q)
{
    u64 r = 1;
    for (a %= q; e; e >>= 1, a = mulm(a, a, q))
        if (e & 1) r = mulm(r, a, q);
    return r;
}
inline u64 invm(u64 a, u64 q) { return powm(a, q - 2, q); }
uint32_t bitrev(uint32_t x, int bits)
{
    uint32_t r = 0;
    for (int i = 0; i < bits; i++, x >>= 1) r = (r << 1) | (x & 1);
    return r;
}
bool probable_prime(u64 n)
{
    if (n < 4) return n == 2 || n == 3;
    if (!(n & 1)) return false;
    u64 d = n - 1;
    int s = 0;
    for (; !(d & 1); d >>= 1) s++;
    for (u64 a : { 2ull, 325ull, 9375ull, 28178ull, 450775ull, 9780504ull, 1795265022ull }) {   // Jaeschke/Sinclair 64-bit set
        u64 x = powm(a % n, d, n);
        if (a % n == 0 || x == 1 || x == n - 1) continue;
        int i = 1;
        for (; i < s; i++) {
            x = mulm(x, x, n);
            if (x == n - 1) break;
        }
        if (i == s) return false;
    }
    return true;
}
// descending primes p = 1 (mod 2N) with exactly `bits` bits
std::vector<u64> ntt_primes(size_t N, int bits, size_t count)
{
    std::vector<u64> r;
    u64 step = 2 * (u64)N;
    for (u64 v = (u64(1) << bits) - step + 1; r.size() < count && v > (u64(1) << (bits - 1)); v -= step)
        if (probable_prime(v)) r.push_back(v);
    return r;
}
u64 min_root(size_t N, u64 q)
{
    u64 deg = 2 * (u64)N, e = (q - 1) / deg, r = 0;
    for (u64 g = 2; !r; g++) {
        u64 x = powm(g, e, q);
        if (powm(x, deg / 2, q) == q - 1) r = x;
    }
    u64 sq = mulm(r, r, q), best = r, cur = r;
    for (u64 i = 0; i < deg / 2; i++, cur = mulm(cur, sq, q))
        if (cur < best) best = cur;
    return best;
}

struct Ntt {
    u64 q = 0, psi = 0, ninv = 0;
    size_t N = 0;
    int logn = 0;
    std::vector<u64> w, ws, iw, iws;
    static u64 sq(u64 x, u64 q) { return (u64)(((u128)x << 64) / q); }
    void init(u64 q_, size_t N_)
    {
        q = q_; N = N_;
        for (logn = 0; (size_t(1) << logn) < N; logn++) {}
        psi = min_root(N, q);
        u64 ipsi = invm(psi, q), p = 1, ip = 1;
        w.resize(N); ws.resize(N); iw.resize(N); iws.resize(N);
        for (size_t i = 0; i < N; i++, p = mulm(p, psi, q), ip = mulm(ip, ipsi, q)) {
            size_t k = bitrev((uint32_t)i, logn);
            w[k] = p; ws[k] = sq(p, q);
            iw[k] = ip; iws[k] = sq(ip, q);
        }
        ninv = invm(N % q, q);
    }
    static inline u64 smul(u64 y, u64 w, u64 ws, u64 q)
    {
        u64 r = y * w - (u64)(((u128)y * ws) >> 64) * q;
        return r >= q ? r - q : r;
    }
    void fwd(u64 *x) const   // natural -> bit-reversed
    {
        for (size_t m = 1, gap = N >> 1; m < N; m <<= 1, gap >>= 1)
            for (size_t i = 0; i < m; i++) {
                u64 *a = x + 2 * i * gap, *b = a + gap;
                for (size_t j = 0; j < gap; j++) {
                    u64 t = smul(b[j], w[m + i], ws[m + i], q), u = a[j];
                    a[j] = addm(u, t, q);
                    b[j] = subm(u, t, q);
                }
            }
    }
    void inv(u64 *x) const   // bit-reversed -> natural
    {
        for (size_t m = N >> 1, gap = 1; m >= 1; m >>= 1, gap <<= 1)
            for (size_t i = 0; i < m; i++) {
                u64 *a = x + 2 * i * gap, *b = a + gap;
                for (size_t j = 0; j < gap; j++) {
                    u64 u = a[j], v = b[j];
                    a[j] = addm(u, v, q);
                    b[j] = smul(subm(u, v, q), iw[m + i], iws[m + i], q);
                }
            }
        for (size_t j = 0; j < N; j++) x[j] = mulm(x[j], ninv, q);
    }
};

// tiny unsigned big integer (little-endian words)
struct Big {
    std::vector<u64> w;
    explicit Big(u64 v = 0) : w(1, v) {}
    void trim() { while (w.size() > 1 && !w.back()) w.pop_back(); }
    void mul_small(u64 m)
    {
        u64 c = 0;
        for (auto &x : w) { u128 p = (u128)x * m + c; x = (u64)p; c = (u64)(p >> 64); }
        if (c) w.push_back(c);
    }
    void add_mul(const Big &b, u64 m)   // this += b*m
    {
        if (w.size() < b.w.size() + 1) w.resize(b.w.size() + 1, 0);
        u64 c = 0;
        size_t i = 0;
        for (; i < b.w.size(); i++) { u128 p = (u128)b.w[i] * m + w[i] + c; w[i] = (u64)p; c = (u64)(p >> 64); }
        for (; c; i++) {
            if (i == w.size()) w.push_back(0);
            u128 p = (u128)w[i] + c; w[i] = (u64)p; c = (u64)(p >> 64);
        }
        trim();
    }
    int cmp(const Big &b) const
    {
        size_t n = std::max(w.size(), b.w.size());
        for (size_t i = n; i-- > 0;) {
            u64 x = i < w.size() ? w[i] : 0, y = i < b.w.size() ? b.w[i] : 0;
            if (x != y) return x < y ? -1 : 1;
        }
        return 0;
    }
    void sub(const Big &b)   // this -= b (this >= b)
    {
        u64 br = 0;
        for (size_t i = 0; i < w.size(); i++) {
            u64 y = i < b.w.size() ? b.w[i] : 0;
            u128 d = (u128)w[i] - y - br;
            w[i] = (u64)d;
            br = (d >> 64) ? 1 : 0;
        }
        trim();
    }
    u64 divmod_small(u64 d)   // this /= d, returns remainder
    {
        u64 r = 0;
        for (size_t i = w.size(); i-- > 0;) { u128 cur = ((u128)r << 64) | w[i]; w[i] = (u64)(cur / d); r = (u64)(cur % d); }
        trim();
        return r;
    }
    u64 mod_small(u64 d) const
    {
        u64 r = 0;
        for (size_t i = w.size(); i-- > 0;) r = (u64)((((u128)r << 64) | w[i]) % d);
        return r;
    }
    void shr1()
    {
        for (size_t i = 0; i < w.size(); i++) w[i] = (w[i] >> 1) | (i + 1 < w.size() ? w[i + 1] << 63 : 0);
        trim();
    }
    long double to_ld() const
    {
        long double r = 0;
        for (size_t i = w.size(); i-- > 0;) r = r * 18446744073709551616.0L + (long double)w[i];
        return r;
    }
};

bool os_entropy(void *dst, size_t n)
{
    unsigned char *p = (unsigned char *)dst;
    size_t got = 0;
    while (got < n) {
        ssize_t r = getrandom(p + got, n - got, 0);
        if (r <= 0) break;
        got += (size_t)r;
    }
    if (got == n) return true;
    FILE *f = fopen("/dev/urandom", "rb");
    if (!f) return false;
    const bool ok = fread(p, 1, n, f) == n;
    fclose(f);
    return ok;
}

// ChaCha20 keystream generator (RFC 8439 block function; 256-bit key, 64-bit stream id, 64-bit block counter).
// Every purpose draws from its own stream of the context's master key (secret key, public key, relinearization key,
// one per Galois element, one per encryption), so what a stream yields does not depend on the order in which other
// streams are consumed, and encryptions can run on several threads.
struct Rng {
    uint32_t key[8];
    u64 stream, block = 0;
    uint32_t buf[16];
    int used = 16;
    Rng(const uint32_t (&k)[8], u64 stream_id) : stream(stream_id) { memcpy(key, k, sizeof key); }
    static uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
    static void qr(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d)
    {
        a += b; d ^= a; d = rotl32(d, 16);
        c += d; b ^= c; b = rotl32(b, 12);
        a += b; d ^= a; d = rotl32(d, 8);
        c += d; b ^= c; b = rotl32(b, 7);
    }
    void refill()
    {
        uint32_t in[16] = { 0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                            (uint32_t)block, (uint32_t)(block >> 32), (uint32_t)stream, (uint32_t)(stream >> 32) };
        uint32_t x[16];
        memcpy(x, in, sizeof x);
        for (int r = 0; r < 10; r++) {
            qr(x[0], x[4], x[8], x[12]); qr(x[1], x[5], x[9], x[13]); qr(x[2], x[6], x[10], x[14]); qr(x[3], x[7], x[11], x[15]);
            qr(x[0], x[5], x[10], x[15]); qr(x[1], x[6], x[11], x[12]); qr(x[2], x[7], x[8], x[13]); qr(x[3], x[4], x[9], x[14]);
        }
        for (int i = 0; i < 16; i++) buf[i] = x[i] + in[i];
        block++;
        used = 0;
    }
    u64 next()
    {
        if (used > 14) refill();
        u64 r = (u64)buf[used] | ((u64)buf[used + 1] << 32);
        used += 2;
        return r;
    }
    u64 below(u64 q)   // uniform in [0,q) by rejection
    {
        u64 lim = ~u64(0) - (~u64(0) % q + 1) % q;
        for (;;) { u64 v = next(); if (v <= lim) return v % q; }
    }
};
enum : u64 { STREAM_SK = 1, STREAM_PK = 2, STREAM_RELIN = 3, STREAM_GALOIS = u64(4) << 32, STREAM_ENCRYPT = u64(5) << 32 };
}   // namespace

struct hfhe_ctx {
    int scheme;
    size_t N, K, Ltop;
    int logn;
    std::vector<u64> q, psi;
    std::vector<Ntt> ntt;
    u64 t = 0;
    Ntt ntt_t;   // BFV batching
    double scale = 1.0;
    uint32_t master[8];                        // ChaCha20 key of every stream of this context
    std::atomic<u64> enc_streams{ 0 };         // encryption streams handed out so far
    std::vector<u64> sk;        // [K][N] NTT form
    std::vector<u64> pk;        // [2][K][N] NTT form
    std::vector<u64> relin;     // [Ltop][2][K][N]
    std::map<uint32_t, std::vector<u64>> galois;
    std::vector<uint32_t> galois_elts, slot_map;   // slot_map: matrix_reps_index_map
    std::vector<std::complex<double>> croots;      // zeta^{bitrev(k)}

    void to_rns_signed(const std::vector<int64_t> &v, size_t limbs, u64 *out) const
    {
        for (size_t l = 0; l < limbs; l++)
            for (size_t n = 0; n < N; n++) out[l * N + n] = v[n] >= 0 ? (u64)v[n] % q[l] : q[l] - ((u64)(-v[n]) % q[l]);
    }
    // uniform ternary secret / encryption randomness (SEAL sample_poly_ternary)
    static void sample_ternary(Rng &rng, std::vector<int64_t> &v) { for (auto &x : v) x = (int64_t)rng.below(3) - 1; }
    static void sample_error(Rng &rng, std::vector<int64_t> &v)
    {   // centred binomial, 21 coin pairs: sigma ~ 3.24 (SEAL's default noise since 3.6: sample_poly_cbd, same 21 pairs)
        for (auto &x : v) {
            u64 r = rng.next();
            x = (int64_t)__builtin_popcountll(r & 0x1fffff) - (int64_t)__builtin_popcountll((r >> 21) & 0x1fffff);
        }
    }
    // (b, a) with b = -(a s + e) over `limbs` key-level primes, NTT form; b,a: [limbs][N] strided by K in dst
    void zero_sym(Rng &rng, u64 *b, u64 *a)
    {
        std::vector<int64_t> e(N);
        sample_error(rng, e);
        std::vector<u64> er(K * N);
        to_rns_signed(e, K, er.data());
        for (size_t l = 0; l < K; l++) {
            ntt[l].fwd(er.data() + l * N);
            for (size_t n = 0; n < N; n++) {
                u64 av = rng.below(q[l]);
                a[l * N + n] = av;
                u64 v = addm(mulm(av, sk[l * N + n], q[l]), er[l * N + n], q[l]);
                b[l * N + n] = v ? q[l] - v : 0;
            }
        }
    }
    // SEAL KeyGenerator::generate_one_kswitch_key restated: new_key [K][N] NTT form
    void make_kswitch(Rng &rng, const u64 *new_key, std::vector<u64> &out)
    {
        out.assign(Ltop * 2 * K * N, 0);
        for (size_t j = 0; j < Ltop; j++) {
            u64 *b = out.data() + (j * 2 + 0) * K * N, *a = out.data() + (j * 2 + 1) * K * N;
            zero_sym(rng, b, a);
            u64 factor = q[K - 1] % q[j];
            for (size_t n = 0; n < N; n++) b[j * N + n] = addm(b[j * N + n], mulm(new_key[j * N + n], factor, q[j]), q[j]);
        }
    }
    void galois_table(uint32_t elt, std::vector<uint32_t> &tab) const
    {
        tab.resize(N);
        for (size_t i = 0; i < N; i++) {
            u64 e = (u64)elt * (2 * (u64)bitrev((uint32_t)i, logn) + 1);
            tab[i] = bitrev((uint32_t)((e >> 1) & (N - 1)), logn);
        }
    }
    void cfft_fwd(std::complex<double> *x) const
    {
        for (size_t m = 1, gap = N >> 1; m < N; m <<= 1, gap >>= 1)
            for (size_t i = 0; i < m; i++) {
                std::complex<double> w = croots[m + i];
                for (size_t j = 2 * i * gap; j < 2 * i * gap + gap; j++) {
                    std::complex<double> t = w * x[j + gap], u = x[j];
                    x[j] = u + t;
                    x[j + gap] = u - t;
                }
            }
    }
    void cfft_inv(std::complex<double> *x) const
    {
        for (size_t m = N >> 1, gap = 1; m >= 1; m >>= 1, gap <<= 1)
            for (size_t i = 0; i < m; i++) {
                std::complex<double> w = std::conj(croots[m + i]);
                for (size_t j = 2 * i * gap; j < 2 * i * gap + gap; j++) {
                    std::complex<double> u = x[j], v = x[j + gap];
                    x[j] = u + v;
                    x[j + gap] = (u - v) * w;
                }
            }
        for (size_t j = 0; j < N; j++) x[j] /= (double)N;
    }
};

extern "C" hfhe_ctx *hfhe_create(int scheme, size_t N, size_t depth, int coeff_bits, int sp_bits, uint64_t seed)
{
    hfhe_ctx *c = new hfhe_ctx;
    if (seed == 0) {   // the default: 256 bits from the operating system
        if (!os_entropy(c->master, sizeof c->master)) { delete c; return nullptr; }
    } else {           // reproducible runs (tests, bench.py, HEB_B200_SEED): the key is expanded from the seed.  NOT for real keys.
        for (int i = 0; i < 8; i += 2) {   // splitmix64
            u64 z = (seed += 0x9e3779b97f4a7c15ull);
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            z ^= z >> 31;
            c->master[i] = (uint32_t)z;
            c->master[i + 1] = (uint32_t)(z >> 32);
        }
    }
    c->scheme = scheme;
    c->N = N;
    c->K = depth + 1;
    c->Ltop = depth;
    for (c->logn = 0; (size_t(1) << c->logn) < N; c->logn++) {}
    // CoeffModulus::Create({60, bits x (depth-1), 60}): equal-size primes are handed out smallest-first,
    // so the first 60-bit entry is the 2nd largest 60-bit prime and the special prime the largest.
    size_t n60 = 2 + (coeff_bits == 60 ? depth - 1 : 0);
    std::vector<u64> p60 = ntt_primes(N, 60, n60), pb;
    if (coeff_bits != 60) pb = ntt_primes(N, coeff_bits, depth - 1);
    c->q.resize(c->K);
    c->q[0] = p60.back(); p60.pop_back();
    for (size_t i = 1; i < depth; i++) {
        std::vector<u64> &src = coeff_bits == 60 ? p60 : pb;
        c->q[i] = src.back(); src.pop_back();
    }
    c->q[c->K - 1] = p60.back();
    c->ntt.resize(c->K);
    c->psi.resize(c->K);
    for (size_t i = 0; i < c->K; i++) { c->ntt[i].init(c->q[i], N); c->psi[i] = c->ntt[i].psi; }
    if (scheme == HFHE_BFV) {
        c->t = ntt_primes(N, sp_bits, 1)[0];
        c->ntt_t.init(c->t, N);
    } else {
        c->scale = sp_bits > 0 ? std::pow(2.0, sp_bits) : 1.0;
        c->croots.resize(N);
        const long double pi = 3.14159265358979323846264338327950288L;
        for (size_t k = 0; k < N; k++) {
            long double ang = pi * (long double)bitrev((uint32_t)k, c->logn) / (long double)N;
            c->croots[k] = { (double)cosl(ang), (double)sinl(ang) };
        }
    }
    // slot index map (generator 3): slot i <-> evaluation point zeta^{3^i}; second half: conjugates / row 2
    c->slot_map.resize(N);
    {
        u64 m = 2 * N, pos = 1;
        for (size_t i = 0; i < N / 2; i++, pos = pos * 3 % m) {
            c->slot_map[i] = bitrev((uint32_t)((pos - 1) >> 1), c->logn);
            c->slot_map[N / 2 + i] = bitrev((uint32_t)((m - pos - 1) >> 1), c->logn);
        }
    }
    // secret key, public key, relin key
    std::vector<int64_t> s(N);
    {
        Rng r(c->master, STREAM_SK);
        hfhe_ctx::sample_ternary(r, s);
    }
    c->sk.resize(c->K * N);
    c->to_rns_signed(s, c->K, c->sk.data());
    for (size_t l = 0; l < c->K; l++) c->ntt[l].fwd(c->sk.data() + l * N);
    c->pk.resize(2 * c->K * N);
    {
        Rng r(c->master, STREAM_PK);
        c->zero_sym(r, c->pk.data(), c->pk.data() + c->K * N);
    }
    std::vector<u64> s2(c->K * N);
    for (size_t l = 0; l < c->K; l++)
        for (size_t n = 0; n < N; n++) s2[l * N + n] = mulm(c->sk[l * N + n], c->sk[l * N + n], c->q[l]);
    {
        Rng r(c->master, STREAM_RELIN);
        c->make_kswitch(r, s2.data(), c->relin);
    }
    // default Galois elements: 3^(2^k), 3^-(2^k), then 2N-1
    {
        u64 m = 2 * N, pos = 3, neg = 1;
        while (neg * 3 % m != 1) neg += 2;
        for (int i = 0; i < c->logn - 1; i++) {
            c->galois_elts.push_back((uint32_t)pos);
            c->galois_elts.push_back((uint32_t)neg);
            pos = pos * pos % m;
            neg = neg * neg % m;
        }
        c->galois_elts.push_back((uint32_t)(m - 1));
    }
    return c;
}
extern "C" void hfhe_destroy(hfhe_ctx *c) { delete c; }
extern "C" size_t hfhe_N(const hfhe_ctx *c) { return c->N; }
extern "C" size_t hfhe_K(const hfhe_ctx *c) { return c->K; }
extern "C" const uint64_t *hfhe_moduli(const hfhe_ctx *c) { return c->q.data(); }
extern "C" const uint64_t *hfhe_psi(const hfhe_ctx *c) { return c->psi.data(); }
extern "C" uint64_t hfhe_plain_modulus(const hfhe_ctx *c) { return c->t; }
extern "C" double hfhe_scale(const hfhe_ctx *c) { return c->scale; }
extern "C" const uint64_t *hfhe_relin_key(hfhe_ctx *c) { return c->relin.data(); }
extern "C" size_t hfhe_galois_count(hfhe_ctx *c) { return c->galois_elts.size(); }
extern "C" uint32_t hfhe_galois_elt(hfhe_ctx *c, size_t i) { return c->galois_elts[i]; }
extern "C" size_t hfhe_kswitch_key_words(const hfhe_ctx *c) { return c->Ltop * 2 * c->K * c->N; }
extern "C" const uint64_t *hfhe_galois_key(hfhe_ctx *c, uint32_t elt)
{
    if (!(elt & 1) || elt >= 2 * c->N) return nullptr;
    auto it = c->galois.find(elt);
    if (it != c->galois.end()) return it->second.data();
    std::vector<uint32_t> tab;
    c->galois_table(elt, tab);
    std::vector<u64> rs(c->K * c->N);
    for (size_t l = 0; l < c->K; l++)
        for (size_t n = 0; n < c->N; n++) rs[l * c->N + n] = c->sk[l * c->N + tab[n]];
    std::vector<u64> &dst = c->galois[elt];
    Rng r(c->master, STREAM_GALOIS | elt);
    c->make_kswitch(r, rs.data(), dst);
    return dst.data();
}

extern "C" void hfhe_ckks_encode(hfhe_ctx *c, const double *vals, size_t n, double scale, uint64_t *plain)
{
    const size_t N = c->N, slots = N / 2;
    std::vector<std::complex<double>> x(N, 0.0);
    for (size_t i = 0; i < n && i < slots; i++) {
        x[c->slot_map[i]] = vals[i];
        x[c->slot_map[slots + i]] = vals[i];   // conj of a real value
    }
    c->cfft_inv(x.data());
    std::vector<int64_t> coeff(N);
    for (size_t j = 0; j < N; j++) coeff[j] = (int64_t)std::llround(x[j].real() * scale);
    c->to_rns_signed(coeff, c->Ltop, plain);
    for (size_t l = 0; l < c->Ltop; l++) c->ntt[l].fwd(plain + l * N);
}

extern "C" void hfhe_ckks_decode(hfhe_ctx *c, const uint64_t *plain, size_t L, double scale, double *out)
{
    const size_t N = c->N;
    std::vector<u64> p(plain, plain + L * N);
    for (size_t l = 0; l < L; l++) c->ntt[l].inv(p.data() + l * N);
    Big Q(1);
    for (size_t l = 0; l < L; l++) Q.mul_small(c->q[l]);
    Big halfQ = Q;
    halfQ.shr1();
    std::vector<Big> punct(L);
    std::vector<u64> ipunct(L);
    for (size_t l = 0; l < L; l++) {
        Big P(1);
        for (size_t j = 0; j < L; j++)
            if (j != l) P.mul_small(c->q[j]);
        punct[l] = P;
        ipunct[l] = invm(P.mod_small(c->q[l]), c->q[l]);
    }
    std::vector<std::complex<double>> x(N);
    for (size_t n = 0; n < N; n++) {
        Big acc(0);
        for (size_t l = 0; l < L; l++) acc.add_mul(punct[l], mulm(p[l * N + n], ipunct[l], c->q[l]));
        while (acc.cmp(Q) >= 0) acc.sub(Q);
        long double v;
        if (acc.cmp(halfQ) > 0) {
            Big neg = Q;
            neg.sub(acc);
            v = -neg.to_ld();
        } else
            v = acc.to_ld();
        x[n] = (double)(v / (long double)scale);
    }
    c->cfft_fwd(x.data());
    for (size_t i = 0; i < N / 2; i++) out[i] = x[c->slot_map[i]].real();
}

extern "C" void hfhe_bfv_encode(hfhe_ctx *c, const int64_t *vals, size_t n, uint64_t *plain)
{
    const size_t N = c->N;
    const u64 t = c->t;
    std::vector<u64> x(N, 0);
    for (size_t i = 0; i < n && i < N; i++) {
        int64_t v = vals[i] % (int64_t)t;
        x[c->slot_map[i]] = v < 0 ? (u64)(v + (int64_t)t) : (u64)v;
    }
    c->ntt_t.inv(x.data());
    memcpy(plain, x.data(), N * sizeof(u64));
}
extern "C" void hfhe_bfv_decode(hfhe_ctx *c, const uint64_t *plain, int64_t *out)
{
    const size_t N = c->N;
    const u64 t = c->t;
    std::vector<u64> x(plain, plain + N);
    c->ntt_t.fwd(x.data());
    for (size_t i = 0; i < N; i++) {
        u64 v = x[c->slot_map[i]];
        out[i] = v > t / 2 ? (int64_t)v - (int64_t)t : (int64_t)v;
    }
}

extern "C" uint64_t hfhe_reserve_encryptions(hfhe_ctx *c, uint64_t n) { return c->enc_streams.fetch_add(n); }
extern "C" void hfhe_encrypt(hfhe_ctx *c, const uint64_t *plain, uint64_t *ct) { hfhe_encrypt_at(c, plain, ct, hfhe_reserve_encryptions(c, 1)); }
extern "C" void hfhe_encrypt_at(hfhe_ctx *c, const uint64_t *plain, uint64_t *ct, uint64_t index)
{
    const size_t N = c->N, L = c->Ltop, K = c->K;
    Rng rng(c->master, STREAM_ENCRYPT + index);
    std::vector<int64_t> u(N), e(N);
    hfhe_ctx::sample_ternary(rng, u);
    std::vector<u64> ur(L * N), er(L * N);
    c->to_rns_signed(u, L, ur.data());
    for (size_t l = 0; l < L; l++) c->ntt[l].fwd(ur.data() + l * N);
    for (size_t k = 0; k < 2; k++) {
        hfhe_ctx::sample_error(rng, e);
        c->to_rns_signed(e, L, er.data());
        for (size_t l = 0; l < L; l++) {
            c->ntt[l].fwd(er.data() + l * N);
            const u64 *pkp = c->pk.data() + (k * K + l) * N;
            u64 *dst = ct + (k * L + l) * N;
            for (size_t n = 0; n < N; n++) dst[n] = addm(mulm(pkp[n], ur[l * N + n], c->q[l]), er[l * N + n], c->q[l]);
        }
    }
    if (c->scheme == HFHE_CKKS) {
        for (size_t l = 0; l < L; l++)
            for (size_t n = 0; n < N; n++) ct[l * N + n] = addm(ct[l * N + n], plain[l * N + n], c->q[l]);
    } else {
        for (size_t k = 0; k < 2; k++)
            for (size_t l = 0; l < L; l++) c->ntt[l].inv(ct + (k * L + l) * N);
        // SEAL multiply_add_plain_with_scaling_variant restated: c0 += round(Q*m/t)
        Big Q(1);
        for (size_t l = 0; l < L; l++) Q.mul_small(c->q[l]);
        Big Qdiv = Q;
        u64 q_mod_t = Qdiv.divmod_small(c->t);
        std::vector<u64> qdiv_mod(L);
        for (size_t l = 0; l < L; l++) qdiv_mod[l] = Qdiv.mod_small(c->q[l]);
        for (size_t n = 0; n < N; n++) {
            u64 m = plain[n];
            u64 fix = (u64)(((u128)m * q_mod_t + (c->t + 1) / 2) / c->t);
            for (size_t l = 0; l < L; l++) {
                u64 v = (u64)(((u128)m * qdiv_mod[l] + fix) % c->q[l]);
                ct[l * N + n] = addm(ct[l * N + n], v, c->q[l]);
            }
        }
    }
}

extern "C" void hfhe_decrypt(hfhe_ctx *c, const uint64_t *ct, size_t size, size_t L, uint64_t *plain)
{
    const size_t N = c->N;
    const bool ckks = c->scheme == HFHE_CKKS;
    std::vector<u64> acc(L * N), tmp(N), spow(N);
    for (size_t l = 0; l < L; l++) {
        const u64 q = c->q[l];
        for (size_t n = 0; n < N; n++) spow[n] = 1;
        for (size_t k = 0; k < size; k++) {
            memcpy(tmp.data(), ct + (k * L + l) * N, N * sizeof(u64));
            if (!ckks) c->ntt[l].fwd(tmp.data());
            for (size_t n = 0; n < N; n++) {
                u64 v = mulm(tmp[n], spow[n], q);
                acc[l * N + n] = k ? addm(acc[l * N + n], v, q) : v;
                spow[n] = mulm(spow[n], c->sk[l * N + n], q);
            }
        }
        if (!ckks) c->ntt[l].inv(acc.data() + l * N);
    }
    if (ckks) {
        memcpy(plain, acc.data(), L * N * sizeof(u64));
        return;
    }
    // BFV: m = round(t*x/Q) mod t via the CRT fractional representation
    const u64 t = c->t;
    std::vector<u64> ipunct(L);
    for (size_t l = 0; l < L; l++) {
        u64 p = 1;
        for (size_t j = 0; j < L; j++)
            if (j != l) p = mulm(p, c->q[j] % c->q[l], c->q[l]);
        ipunct[l] = invm(p, c->q[l]);
    }
    for (size_t n = 0; n < N; n++) {
        u64 ipart = 0;
        long double frac = 0;
        for (size_t l = 0; l < L; l++) {
            u64 cl = mulm(acc[l * N + n], ipunct[l], c->q[l]);
            u128 tc = (u128)t * cl;
            ipart = (ipart + (u64)((tc / c->q[l]) % t)) % t;
            frac += (long double)(u64)(tc % c->q[l]) / (long double)c->q[l];
        }
        u64 r = (u64)floorl(frac + 0.5L);
        plain[n] = (ipart + r) % t;
    }
}
