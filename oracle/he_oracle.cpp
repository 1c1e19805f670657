/*
 * he_oracle.cpp -- CPU oracle (scalar C++17, unsigned __int128) for the
 * ciphertext-evaluation hot path of hebench/reference-seal-backend.
 *
 * TEST INFRASTRUCTURE ONLY (see he_oracle.h).  PARITY UNPINNED: restates
 * Microsoft SEAL v3.7.2's published algorithms (SEAL is not vendored under
 * /root/reference); anchored on the reference's
 * own call sites, cited per function as R/<path>:<line> = /root/reference/<path>:<line>.
 *
 * Every routine produces canonical residues in [0,q), which is what SEAL stores
 * in a Ciphertext after each Evaluator call; internal lazy ranges are free.
 */
#include "he_oracle.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef unsigned __int128 u128;

// ------------------------------------------------------------------ scalar arithmetic
static inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
static inline u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }   // a,b < q < 2^63
static inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
static u64 powmod(u64 a, u64 e, u64 q)
{
    u64 r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}
static inline u64 invmod_prime(u64 a, u64 q) { return powmod(a, q - 2, q); }
static inline u64 shoup_quot(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
// w*y mod q, result in [0,2q) for any 64-bit y (Harvey/Shoup lazy product)
static inline u64 shoup_mul_lazy(u64 y, u64 w, u64 ws, u64 q)
{
    u64 hi = (u64)(((u128)y * ws) >> 64);
    return y * w - hi * q;
}
// Barrett reduction with the two-word ratio floor(2^128 / q) (the form SEAL's Modulus::const_ratio() holds): the hot
// loops below use it; the plain `%` forms above remain for set-up code and as the checker of this one
// (orc_selftest_fastmod, tests/test_oracle.py).  Valid for q < 2^62.
struct FastMod {
    u64 q = 0, r0 = 0, r1 = 0;   // floor(2^128 / q) = r1 * 2^64 + r0
    void init(u64 q_)
    {
        q = q_;
        // 2^128 / q by long division of (2^128 - 1) / q: the quotients agree because q does not divide 2^128
        const u128 all = ~(u128)0;
        const u128 quo = all / q;
        r0 = (u64)quo;
        r1 = (u64)(quo >> 64);
    }
    // (hi:lo) mod q for any 128-bit input
    inline u64 red128(u64 hi, u64 lo) const
    {
        u64 carry = (u64)(((u128)lo * r0) >> 64);
        const u128 t2 = (u128)lo * r1;
        const u64 tmp1 = (u64)t2 + carry;
        const u64 tmp3 = (u64)(t2 >> 64) + (tmp1 < carry);
        const u128 t3 = (u128)hi * r0;
        const u64 tmp1b = tmp1 + (u64)t3;
        carry = (u64)(t3 >> 64) + (tmp1b < tmp1);
        const u64 qe = hi * r1 + tmp3 + carry;
        const u64 r = lo - qe * q;
        return r >= q ? r - q : r;
    }
    inline u64 mul(u64 a, u64 b) const
    {
        const u128 p = (u128)a * b;
        return red128((u64)(p >> 64), (u64)p);
    }
    // a * b + c mod q (a, b, c < 2^63)
    inline u64 mad(u64 a, u64 b, u64 c) const
    {
        const u128 p = (u128)a * b + c;
        return red128((u64)(p >> 64), (u64)p);
    }
    // x mod q for any 64-bit x
    inline u64 red64(u64 x) const
    {
        const u64 r = x - (u64)(((u128)x * r1) >> 64) * q;
        return r >= q ? r - q : r;
    }
};
static inline uint32_t brv(uint32_t x, int bits)
{
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

// self-check of FastMod against the plain `%` forms on n pseudo-random operand triples (and the extreme operands);
// returns the number of mismatches
extern "C" size_t orc_selftest_fastmod(uint64_t q, size_t n, uint64_t seed)
{
    FastMod fm;
    fm.init(q);
    size_t bad = 0;
    u64 s = seed | 1;
    auto next = [&]() {   // xorshift64*
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        return s * 0x2545F4914F6CDD1Dull;
    };
    auto check = [&](u64 a, u64 b, u64 c, u64 x) {
        if (fm.mul(a, b) != mulmod(a, b, q)) bad++;
        if (fm.mad(a, b, c) != (u64)(((u128)a * b + c) % q)) bad++;
        if (fm.red64(x) != x % q) bad++;
        if (fm.red128(x, a) != (u64)((((u128)x << 64) | a) % q)) bad++;
    };
    const u64 ext[] = { 0, 1, q - 1, q >> 1, (q >> 1) + 1 };
    for (u64 a : ext)
        for (u64 b : ext)
            for (u64 c : ext) check(a, b, c, ~u64(0) - a);
    for (size_t i = 0; i < n; i++) check(next() % q, next() % q, next() % q, next());
    return bad;
}

// ------------------------------------------------------------------ number theory
extern "C" int orc_is_prime(uint64_t n)
{
    if (n < 2) return 0;
    static const u64 small[] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37 };
    for (u64 p : small) {
        if (n == p) return 1;
        if (n % p == 0) return 0;
    }
    u64 d = n - 1;
    int r = 0;
    while ((d & 1) == 0) { d >>= 1; r++; }
    for (u64 a : small) {   // deterministic for all n < 2^64
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < r; i++) {
            x = mulmod(x, x, n);
            if (x == n - 1) { comp = false; break; }
        }
        if (comp) return 0;
    }
    return 1;
}

// SEAL util::get_primes: candidates 2^bits - k*factor + 1, descending, while > 2^(bits-1)
extern "C" int orc_get_primes(uint64_t factor, int bits, size_t count, uint64_t *out)
{
    u64 value = ((u64(1) << bits) - 1) / factor * factor + 1;
    u64 lower = u64(1) << (bits - 1);
    size_t got = 0;
    while (got < count && value > lower) {
        if (orc_is_prime(value)) out[got++] = value;
        value -= factor;
    }
    return got == count ? 0 : -1;
}

// SEAL CoeffModulus::Create: per distinct bit size fetch as many primes as needed, then
// walk bit_sizes in order taking .back() and pop_back().  Called from
// R/src/engine/seal_context.cpp:89 (CKKS) and :117 (BFV) with {60, b x (depth-1), 60}.
extern "C" int orc_coeff_modulus_create(size_t N, const int *bits, size_t n, uint64_t *out)
{
    std::vector<std::vector<u64>> table(64);
    int cnt[64] = { 0 };
    for (size_t i = 0; i < n; i++) {
        if (bits[i] < 2 || bits[i] > 60) return -1;
        cnt[bits[i]]++;
    }
    for (int b = 0; b < 64; b++)
        if (cnt[b]) {
            table[b].resize(cnt[b]);
            if (orc_get_primes(2 * (u64)N, b, cnt[b], table[b].data())) return -1;
        }
    for (size_t i = 0; i < n; i++) {
        out[i] = table[bits[i]].back();
        table[bits[i]].pop_back();
    }
    return 0;
}

extern "C" uint64_t orc_plain_modulus_batching(size_t N, int bits)
{
    u64 p = 0;
    if (orc_get_primes(2 * (u64)N, bits, 1, &p)) return 0;
    return p;
}

// SEAL try_minimal_primitive_root: the smallest integer among all primitive degree-th roots
extern "C" int orc_minimal_primitive_root(uint64_t degree, uint64_t q, uint64_t *root)
{
    if ((q - 1) % degree) return -1;
    u64 e = (q - 1) / degree, r = 0;
    for (u64 g = 2; g < 4096; g++) {
        u64 x = powmod(g, e, q);
        if (powmod(x, degree / 2, q) == q - 1) { r = x; break; }
    }
    if (!r) return -1;
    u64 sq = mulmod(r, r, q), cur = r, best = r;
    for (u64 i = 0; i < degree / 2; i++) {   // all odd powers of r
        if (cur < best) best = cur;
        cur = mulmod(cur, sq, q);
    }
    *root = best;
    return 0;
}

// SEAL util::naf (used by Evaluator::rotate_internal when a Galois key is missing;
// reached from R/src/engine/seal_context.cpp:378 and R/src/benchmarks/ckks/seal_ckks_matmult_row_benchmark.cpp:507)
extern "C" int orc_naf(int value, int *out)
{
    int n = 0;
    bool sign = value < 0;
    value = std::abs(value);
    for (int i = 0; value; i++) {
        int zi = (value & 1) ? 2 - (value & 3) : 0;
        value = (value - zi) >> 1;
        if (zi) out[n++] = (sign ? -zi : zi) * (1 << i);
    }
    return n;
}

extern "C" uint32_t orc_galois_elt_from_step(int step, size_t N)
{
    uint32_t m = (uint32_t)(2 * N);
    if (step == 0) return m - 1;
    bool sign = step < 0;
    uint32_t pos = (uint32_t)std::abs(step);
    if (pos >= (N >> 1)) return 0;   // invalid
    uint32_t s = sign ? (uint32_t)(N >> 1) - pos : pos;
    u64 elt = 1;
    for (uint32_t i = 0; i < s; i++) elt = (elt * 3) & (m - 1);
    return (uint32_t)elt;
}

// GaloisTool::get_elts_all: 3^(2^k), 3^(-2^k) for k < log2(N/2), then 2N-1
extern "C" int orc_galois_elts_all(size_t N, uint32_t *out)
{
    uint32_t m = (uint32_t)(2 * N);
    int logn = 0;
    while ((size_t(1) << logn) < N) logn++;
    int n = 0;
    u64 pos = 3, neg = 0;
    // inverse of 3 mod 2N
    for (u64 x = 1; x < m; x += 2)
        if ((x * 3) % m == 1) { neg = x; break; }
    for (int i = 0; i < logn - 1; i++) {
        out[n++] = (uint32_t)pos;
        pos = (pos * pos) & (m - 1);
        out[n++] = (uint32_t)neg;
        neg = (neg * neg) & (m - 1);
    }
    out[n++] = m - 1;
    return n;
}

extern "C" void orc_galois_table_ntt(size_t N, uint32_t elt, uint32_t *table)
{
    int logn = 0;
    while ((size_t(1) << logn) < N) logn++;
    for (size_t i = 0; i < N; i++) {
        uint32_t reversed = brv((uint32_t)(i + N), logn + 1);
        u64 raw = ((u64)elt * reversed) >> 1;
        raw &= (u64)(N - 1);
        table[i] = brv((uint32_t)raw, logn);
    }
}

// ------------------------------------------------------------------ context
struct Limb {
    u64 q = 0, psi = 0;
    std::vector<u64> rp, rps, irp, irps;   // rp[k] = psi^{brv(k)}, irp[k] = rp[k]^{-1}; *s = Shoup quotients
    u64 ninv = 0, ninvs = 0;
    FastMod fm;
    void init(u64 q_, size_t N, int logn)
    {
        q = q_;
        fm.init(q_);
        orc_minimal_primitive_root(2 * N, q, &psi);
        rp.assign(N, 0); rps.assign(N, 0); irp.assign(N, 0); irps.assign(N, 0);
        u64 ipsi = invmod_prime(psi, q);
        u64 p = 1, ip = 1;
        for (size_t i = 0; i < N; i++) {
            size_t k = brv((uint32_t)i, logn);
            rp[k] = p; rps[k] = shoup_quot(p, q);
            irp[k] = ip; irps[k] = shoup_quot(ip, q);
            p = mulmod(p, psi, q);
            ip = mulmod(ip, ipsi, q);
        }
        ninv = invmod_prime((u64)N % q, q);
        ninvs = shoup_quot(ninv, q);
    }
};

// SEAL util::BaseConverter restated (util/rns.cpp): fast (approximate) base conversion
struct Conv {
    std::vector<u64> ib, ob;
    std::vector<u64> inv_punct;              // (prod ib / ib_i)^{-1} mod ib_i
    std::vector<std::vector<u64>> mat;       // mat[o][i] = (prod ib / ib_i) mod ob_o
    std::vector<FastMod> fib, fob;
    void init(const std::vector<u64> &ibase, const std::vector<u64> &obase)
    {
        ib = ibase; ob = obase;
        size_t n = ib.size();
        fib.resize(n);
        for (size_t i = 0; i < n; i++) fib[i].init(ib[i]);
        fob.resize(ob.size());
        for (size_t o = 0; o < ob.size(); o++) fob[o].init(ob[o]);
        inv_punct.resize(n);
        for (size_t i = 0; i < n; i++) {
            u64 p = 1;
            for (size_t j = 0; j < n; j++)
                if (j != i) p = mulmod(p, ib[j] % ib[i], ib[i]);
            inv_punct[i] = invmod_prime(p, ib[i]);
        }
        mat.assign(ob.size(), std::vector<u64>(n));
        for (size_t o = 0; o < ob.size(); o++)
            for (size_t i = 0; i < n; i++) {
                u64 p = 1 % ob[o];
                for (size_t j = 0; j < n; j++)
                    if (j != i) p = mulmod(p, ib[j] % ob[o], ob[o]);
                mat[o][i] = p;
            }
    }
    // in: [n_ib][N] (canonical or any 64-bit), out: [n_ob][N]
    void convert(const u64 *in, u64 *out, size_t N) const
    {
        size_t n = ib.size();
        std::vector<u64> tmp(n * N);
        for (size_t i = 0; i < n; i++)
            for (size_t c = 0; c < N; c++) tmp[i * N + c] = fib[i].mul(fib[i].red64(in[i * N + c]), inv_punct[i]);
        for (size_t o = 0; o < ob.size(); o++)
            for (size_t c = 0; c < N; c++) {
                u64 acc = 0;   // reduce as we go
                for (size_t i = 0; i < n; i++) acc = fob[o].mad(tmp[i * N + c], mat[o][i], acc);
                out[o * N + c] = acc;
            }
    }
};

struct orc_ctx {
    int scheme = 0;
    size_t N = 0, K = 0;
    int logn = 0;
    std::vector<Limb> limbs;
    u64 t = 0;
    // BEHZ tool at the top data level (BFV only; SEAL util::RNSTool::initialize restated)
    size_t nB = 0;
    std::vector<Limb> bsk;   // B primes then m_sk
    u64 m_sk = 0, gamma = 0, m_tilde = u64(1) << 32;
    Conv q_to_bsk, q_to_mtilde, B_to_q, B_to_msk;
    std::vector<u64> prod_q_mod_bsk, inv_prod_q_mod_bsk, inv_mtilde_mod_bsk, prod_B_mod_q;
    u64 neg_inv_prod_q_mod_mtilde = 0, inv_prod_B_mod_msk = 0;
};

static size_t bigprod_bits(const std::vector<u64> &f)
{
    std::vector<u64> w(1, 1);
    for (u64 x : f) {
        u64 carry = 0;
        for (size_t i = 0; i < w.size(); i++) {
            u128 p = (u128)w[i] * x + carry;
            w[i] = (u64)p;
            carry = (u64)(p >> 64);
        }
        if (carry) w.push_back(carry);
    }
    size_t bits = (w.size() - 1) * 64;
    u64 top = w.back();
    while (top) { bits++; top >>= 1; }
    return bits;
}

static void init_behz(orc_ctx *c)
{
    size_t L = c->K - 1, N = c->N;
    std::vector<u64> q(L);
    for (size_t i = 0; i < L; i++) q[i] = c->limbs[i].q;
    int tbits = 0;
    for (u64 x = c->t; x; x >>= 1) tbits++;
    size_t nB = L;
    if (32 + (size_t)tbits + bigprod_bits(q) >= 61 * L + 61) nB++;
    c->nB = nB;
    std::vector<u64> pr(nB + 2);
    orc_get_primes(2 * N, 61, nB + 2, pr.data());
    c->m_sk = pr[0];
    c->gamma = pr[1];
    std::vector<u64> B(pr.begin() + 2, pr.end()), Bsk(B);
    Bsk.push_back(c->m_sk);
    c->bsk.resize(Bsk.size());
    for (size_t i = 0; i < Bsk.size(); i++) c->bsk[i].init(Bsk[i], N, c->logn);
    c->q_to_bsk.init(q, Bsk);
    c->B_to_q.init(B, q);
    c->B_to_msk.init(B, std::vector<u64>{ c->m_sk });
    // q -> {m_tilde}: inverse punctured products are mod q_i; matrix entries mod 2^32
    c->q_to_mtilde.ib = q;
    c->q_to_mtilde.ob = { c->m_tilde };
    c->q_to_mtilde.inv_punct = c->q_to_bsk.inv_punct;
    c->q_to_mtilde.fib = c->q_to_bsk.fib;
    c->q_to_mtilde.mat.assign(1, std::vector<u64>(L));
    for (size_t i = 0; i < L; i++) {
        u64 p = 1;
        for (size_t j = 0; j < L; j++)
            if (j != i) p = (p * (q[j] & 0xffffffffull)) & 0xffffffffull;
        c->q_to_mtilde.mat[0][i] = p;
    }
    c->prod_q_mod_bsk.resize(Bsk.size());
    c->inv_prod_q_mod_bsk.resize(Bsk.size());
    c->inv_mtilde_mod_bsk.resize(Bsk.size());
    for (size_t i = 0; i < Bsk.size(); i++) {
        u64 p = 1;
        for (u64 x : q) p = mulmod(p, x % Bsk[i], Bsk[i]);
        c->prod_q_mod_bsk[i] = p;
        c->inv_prod_q_mod_bsk[i] = invmod_prime(p, Bsk[i]);
        c->inv_mtilde_mod_bsk[i] = invmod_prime(c->m_tilde % Bsk[i], Bsk[i]);
    }
    // -q^{-1} mod 2^32 (Newton iteration on the odd residue)
    u64 qm = 1;
    for (u64 x : q) qm = (qm * (x & 0xffffffffull)) & 0xffffffffull;
    u64 inv = 1;
    for (int i = 0; i < 6; i++) inv = (inv * (2 - qm * inv)) & 0xffffffffull;
    c->neg_inv_prod_q_mod_mtilde = (c->m_tilde - inv) & 0xffffffffull;
    c->prod_B_mod_q.resize(L);
    for (size_t i = 0; i < L; i++) {
        u64 p = 1;
        for (u64 x : B) p = mulmod(p, x % q[i], q[i]);
        c->prod_B_mod_q[i] = p;
    }
    u64 pb = 1;
    for (u64 x : B) pb = mulmod(pb, x % c->m_sk, c->m_sk);
    c->inv_prod_B_mod_msk = invmod_prime(pb, c->m_sk);
}

extern "C" orc_ctx *orc_ctx_create(int scheme, size_t N, size_t K, const uint64_t *moduli, uint64_t plain_modulus)
{
    orc_ctx *c = new orc_ctx;
    c->scheme = scheme;
    c->N = N;
    c->K = K;
    while ((size_t(1) << c->logn) < N) c->logn++;
    c->limbs.resize(K);
    for (size_t i = 0; i < K; i++) c->limbs[i].init(moduli[i], N, c->logn);
    c->t = plain_modulus;
    if (scheme == ORC_SCHEME_BFV && K >= 2) init_behz(c);
    return c;
}
extern "C" void orc_ctx_destroy(orc_ctx *c) { delete c; }
extern "C" uint64_t orc_ctx_psi(const orc_ctx *c, size_t limb) { return c->limbs[limb].psi; }
extern "C" size_t orc_ctx_N(const orc_ctx *c) { return c->N; }
extern "C" size_t orc_ctx_K(const orc_ctx *c) { return c->K; }
extern "C" int orc_ctx_scheme(const orc_ctx *c) { return c->scheme; }
extern "C" size_t orc_ctx_bsk_size(const orc_ctx *c) { return c->bsk.size(); }
extern "C" void orc_ctx_bsk(const orc_ctx *c, uint64_t *out)
{
    for (size_t i = 0; i < c->bsk.size(); i++) out[i] = c->bsk[i].q;
}

// ------------------------------------------------------------------ K1 / K2
// Forward negacyclic NTT, Cooley-Tukey, natural -> bit-reversed, Harvey lazy butterflies
// (SEAL util/ntt.cpp ntt_negacyclic_harvey, util/dwthandler.h transform_to_rev), then
// canonicalised to [0,q).
static void ntt_fwd(const Limb &lm, size_t N, u64 *x)
{
    const u64 q = lm.q, two_q = 2 * q;
    size_t gap = N >> 1;
    for (size_t m = 1; m < N; m <<= 1, gap >>= 1) {
        for (size_t i = 0; i < m; i++) {
            const u64 w = lm.rp[m + i], ws = lm.rps[m + i];
            u64 *a = x + 2 * i * gap, *b = a + gap;
            for (size_t j = 0; j < gap; j++) {
                u64 u = a[j] >= two_q ? a[j] - two_q : a[j];
                u64 v = shoup_mul_lazy(b[j], w, ws, q);
                a[j] = u + v;
                b[j] = u - v + two_q;
            }
        }
    }
    for (size_t j = 0; j < N; j++) {
        u64 v = x[j];
        if (v >= two_q) v -= two_q;
        if (v >= q) v -= q;
        x[j] = v;
    }
}
// Inverse: Gentleman-Sande, bit-reversed -> natural, N^{-1} applied at the end
// (SEAL inverse_ntt_negacyclic_harvey folds it into the last stage; same values).
static void ntt_inv(const Limb &lm, size_t N, u64 *x)
{
    const u64 q = lm.q, two_q = 2 * q;
    size_t gap = 1;
    for (size_t m = N >> 1; m >= 1; m >>= 1, gap <<= 1) {
        for (size_t i = 0; i < m; i++) {
            const u64 w = lm.irp[m + i], ws = lm.irps[m + i];
            u64 *a = x + 2 * i * gap, *b = a + gap;
            for (size_t j = 0; j < gap; j++) {
                u64 u = a[j], v = b[j];   // both in [0,2q)
                u64 s = u + v;
                a[j] = s >= two_q ? s - two_q : s;
                b[j] = shoup_mul_lazy(u - v + two_q, w, ws, q);
            }
        }
    }
    for (size_t j = 0; j < N; j++) {
        u64 v = shoup_mul_lazy(x[j], lm.ninv, lm.ninvs, q);
        x[j] = v >= q ? v - q : v;
    }
}
extern "C" void orc_ntt_fwd(const orc_ctx *c, size_t limb, uint64_t *p) { ntt_fwd(c->limbs[limb], c->N, p); }
extern "C" void orc_ntt_inv(const orc_ctx *c, size_t limb, uint64_t *p) { ntt_inv(c->limbs[limb], c->N, p); }
extern "C" void orc_ntt_fwd_direct(const orc_ctx *c, size_t limb, const uint64_t *in, uint64_t *out)
{
    const Limb &lm = c->limbs[limb];
    size_t N = c->N;
    for (size_t i = 0; i < N; i++) {
        u64 e = 2 * (u64)brv((uint32_t)i, c->logn) + 1;
        u64 w = powmod(lm.psi, e, lm.q), p = 1, acc = 0;
        for (size_t j = 0; j < N; j++) {
            acc = addmod(acc, mulmod(in[j] % lm.q, p, lm.q), lm.q);
            p = mulmod(p, w, lm.q);
        }
        out[i] = acc;
    }
}

// ------------------------------------------------------------------ K3 / K4 / K10 / K11
// Evaluator::add(_inplace): R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:340,
// R/src/engine/seal_context.cpp:303,338,400
extern "C" void orc_add(const orc_ctx *c, size_t L, size_t size, const uint64_t *a, const uint64_t *b, uint64_t *out)
{
    size_t N = c->N;
    for (size_t p = 0; p < size; p++)
        for (size_t l = 0; l < L; l++) {
            u64 q = c->limbs[l].q;
            size_t o = (p * L + l) * N;
            for (size_t j = 0; j < N; j++) out[o + j] = addmod(a[o + j], b[o + j], q);
        }
}
extern "C" void orc_sub(const orc_ctx *c, size_t L, size_t size, const uint64_t *a, const uint64_t *b, uint64_t *out)
{
    size_t N = c->N;
    for (size_t p = 0; p < size; p++)
        for (size_t l = 0; l < L; l++) {
            u64 q = c->limbs[l].q;
            size_t o = (p * L + l) * N;
            for (size_t j = 0; j < N; j++) out[o + j] = submod(a[o + j], b[o + j], q);
        }
}
// Evaluator::multiply on CKKS (ckks_multiply): R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:343,
// R/src/benchmarks/ckks/seal_ckks_dot_product_benchmark.cpp:325
extern "C" void orc_ckks_multiply(const orc_ctx *c, size_t L, const uint64_t *a, const uint64_t *b, uint64_t *out)
{
    size_t N = c->N;
    for (size_t l = 0; l < L; l++) {
        const FastMod &fm = c->limbs[l].fm;
        const u64 *a0 = a + l * N, *a1 = a + (L + l) * N, *b0 = b + l * N, *b1 = b + (L + l) * N;
        u64 *c0 = out + l * N, *c1 = out + (L + l) * N, *c2 = out + (2 * L + l) * N;
        for (size_t j = 0; j < N; j++) {
            c0[j] = fm.mul(a0[j], b0[j]);
            c1[j] = fm.mad(a0[j], b1[j], fm.mul(a1[j], b0[j]));
            c2[j] = fm.mul(a1[j], b1[j]);
        }
    }
}
// mod_switch_to_inplace on CKKS ct / plain (drop limbs): R/src/engine/seal_context.cpp:260,262,388,451
extern "C" void orc_mod_drop(const orc_ctx *c, size_t L, size_t size, const uint64_t *in, uint64_t *out)
{
    size_t N = c->N;
    for (size_t p = 0; p < size; p++)
        memmove(out + p * (L - 1) * N, in + p * L * N, (L - 1) * N * sizeof(u64));
}
// multiply_plain_inplace (NTT form): R/src/engine/seal_context.cpp:389
extern "C" void orc_multiply_plain(const orc_ctx *c, size_t L, size_t size, const uint64_t *ct, const uint64_t *pl, uint64_t *out)
{
    size_t N = c->N;
    for (size_t p = 0; p < size; p++)
        for (size_t l = 0; l < L; l++) {
            const FastMod &fm = c->limbs[l].fm;
            size_t o = (p * L + l) * N;
            for (size_t j = 0; j < N; j++) out[o + j] = fm.mul(ct[o + j], pl[l * N + j]);
        }
}
// add_plain_inplace (CKKS): R/src/engine/seal_context.cpp:454
extern "C" void orc_add_plain(const orc_ctx *c, size_t L, size_t size, const uint64_t *ct, const uint64_t *pl, uint64_t *out)
{
    size_t N = c->N;
    if (out != ct) memcpy(out, ct, size * L * N * sizeof(u64));
    for (size_t l = 0; l < L; l++) {
        u64 q = c->limbs[l].q;
        for (size_t j = 0; j < N; j++) out[l * N + j] = addmod(ct[l * N + j], pl[l * N + j], q);
    }
}

// ------------------------------------------------------------------ K6: switch_key_inplace
// SEAL Evaluator::switch_key_inplace restated.  Reached from
// relinearize_inplace (R/src/benchmarks/ckks/seal_ckks_dot_product_benchmark.cpp:329) and from
// rotate_vector / rotate_rows (R/src/engine/seal_context.cpp:302,337,378).
extern "C" void orc_switch_key(const orc_ctx *c, size_t L, uint64_t *ct, const uint64_t *target, const uint64_t *key)
{
    const size_t N = c->N, K = c->K, sp = K - 1;
    const bool ckks = c->scheme == ORC_SCHEME_CKKS;
    std::vector<u64> t(target, target + L * N);
    if (ckks)
        for (size_t j = 0; j < L; j++) ntt_inv(c->limbs[j], N, t.data() + j * N);
    // acc[k][I] for I in {0..L-1, sp}
    std::vector<u64> acc(2 * (L + 1) * N, 0), d(N);
    for (size_t I = 0; I <= L; I++) {
        const size_t ki = (I == L) ? sp : I;
        const u64 qi = c->limbs[ki].q;
        for (size_t J = 0; J < L; J++) {
            const u64 *op;
            if (ckks && I == J) {
                op = target + J * N;
            } else {
                const u64 qj = c->limbs[J].q;
                const FastMod &fi = c->limbs[ki].fm;
                for (size_t n = 0; n < N; n++) d[n] = (qj > qi) ? fi.red64(t[J * N + n]) : t[J * N + n];
                ntt_fwd(c->limbs[ki], N, d.data());
                op = d.data();
            }
            for (size_t k = 0; k < 2; k++) {
                const u64 *kp = key + ((J * 2 + k) * K + ki) * N;
                u64 *ap = acc.data() + (k * (L + 1) + I) * N;
                const FastMod &fi = c->limbs[ki].fm;
                for (size_t n = 0; n < N; n++) ap[n] = fi.mad(op[n], kp[n], ap[n]);
            }
        }
    }
    // mod-down by the special prime with rounding, add into (c0, c1)
    const u64 qk = c->limbs[sp].q, half = qk >> 1;
    for (size_t k = 0; k < 2; k++) {
        u64 *last = acc.data() + (k * (L + 1) + L) * N;
        ntt_inv(c->limbs[sp], N, last);
        for (size_t n = 0; n < N; n++) last[n] = addmod(last[n], half, qk);
        for (size_t J = 0; J < L; J++) {
            const u64 qj = c->limbs[J].q;
            const u64 fix = qj - half % qj;
            const u64 inv = invmod_prime(qk % qj, qj);
            u64 *aj = acc.data() + (k * (L + 1) + J) * N;
            const FastMod &fj = c->limbs[J].fm;
            for (size_t n = 0; n < N; n++) d[n] = addmod(fj.red64(last[n]), fix % qj, qj);
            if (ckks)
                ntt_fwd(c->limbs[J], N, d.data());
            else
                ntt_inv(c->limbs[J], N, aj);
            u64 *dst = ct + (k * L + J) * N;
            for (size_t n = 0; n < N; n++) {
                u64 v = fj.mul(submod(aj[n], d[n], qj), inv);
                dst[n] = addmod(dst[n], v, qj);
            }
        }
    }
}

// relinearize_inplace (size 3 -> 2): R/src/benchmarks/ckks/seal_ckks_dot_product_benchmark.cpp:329
extern "C" void orc_relinearize(const orc_ctx *c, size_t L, const uint64_t *ct3, const uint64_t *relin_key, uint64_t *out2)
{
    size_t N = c->N;
    memcpy(out2, ct3, 2 * L * N * sizeof(u64));
    orc_switch_key(c, L, out2, ct3 + 2 * L * N, relin_key);
}

// ------------------------------------------------------------------ K8: Galois
static void galois_coeff(const orc_ctx *c, size_t L, const u64 *in, uint32_t elt, u64 *out)
{
    size_t N = c->N;
    for (size_t l = 0; l < L; l++) {
        u64 q = c->limbs[l].q;
        for (size_t i = 0; i < N; i++) {
            u64 raw = (u64)i * elt;
            size_t idx = raw & (N - 1);
            u64 v = in[l * N + i];
            if ((raw >> c->logn) & 1) v = v ? q - v : 0;
            out[l * N + idx] = v;
        }
    }
}
// Evaluator::apply_galois_inplace
extern "C" void orc_apply_galois(const orc_ctx *c, size_t L, uint64_t *ct, uint32_t elt, const uint64_t *gkey)
{
    size_t N = c->N;
    std::vector<u64> g0(L * N), g1(L * N);
    if (c->scheme == ORC_SCHEME_CKKS) {
        std::vector<uint32_t> tab(N);
        orc_galois_table_ntt(N, elt, tab.data());
        for (size_t l = 0; l < L; l++)
            for (size_t i = 0; i < N; i++) {
                g0[l * N + i] = ct[l * N + tab[i]];
                g1[l * N + i] = ct[(L + l) * N + tab[i]];
            }
    } else {
        galois_coeff(c, L, ct, elt, g0.data());
        galois_coeff(c, L, ct + L * N, elt, g1.data());
    }
    memcpy(ct, g0.data(), L * N * sizeof(u64));
    memset(ct + L * N, 0, L * N * sizeof(u64));
    orc_switch_key(c, L, ct, g1.data(), gkey);
}

static const u64 *find_key(uint32_t elt, const uint32_t *elts, const uint64_t *const *keys, size_t nkeys)
{
    for (size_t i = 0; i < nkeys; i++)
        if (elts[i] == elt) return keys[i];
    return nullptr;
}
// Evaluator::rotate_internal (rotate_vector / rotate_rows), NAF fallback in LSB-first order
extern "C" int orc_rotate(const orc_ctx *c, size_t L, uint64_t *ct, int step, const uint32_t *elts,
                          const uint64_t *const *keys, size_t nkeys)
{
    if (step == 0) return 0;
    uint32_t elt = orc_galois_elt_from_step(step, c->N);
    if (!elt) return -1;
    if (const u64 *k = find_key(elt, elts, keys, nkeys)) {
        orc_apply_galois(c, L, ct, elt, k);
        return 0;
    }
    int terms[40];
    int n = orc_naf(step, terms);
    if (n == 1) return -1;
    for (int i = 0; i < n; i++) {
        if ((size_t)std::abs(terms[i]) == (c->N >> 1)) continue;
        int rc = orc_rotate(c, L, ct, terms[i], elts, keys, nkeys);
        if (rc) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------ K9: rescale / BFV mod-switch
// rescale_to_next_inplace: R/src/engine/seal_context.cpp:391,448, R/src/benchmarks/ckks/seal_ckks_matmultval_benchmark.cpp:255
// (RNSTool::divide_and_round_q_last_ntt_inplace / _inplace restated)
extern "C" void orc_rescale(const orc_ctx *c, size_t L, size_t size, const uint64_t *in, uint64_t *out)
{
    const size_t N = c->N;
    const bool ckks = c->scheme == ORC_SCHEME_CKKS;
    const u64 ql = c->limbs[L - 1].q, half = ql >> 1;
    std::vector<u64> last(N), d(N);
    for (size_t p = 0; p < size; p++) {
        memcpy(last.data(), in + (p * L + L - 1) * N, N * sizeof(u64));
        if (ckks) ntt_inv(c->limbs[L - 1], N, last.data());
        for (size_t n = 0; n < N; n++) last[n] = addmod(last[n], half, ql);
        for (size_t i = 0; i + 1 < L; i++) {
            const u64 qi = c->limbs[i].q;
            const u64 fix = qi - half % qi;
            const u64 inv = invmod_prime(ql % qi, qi);
            const FastMod &fi = c->limbs[i].fm;
            for (size_t n = 0; n < N; n++) d[n] = addmod(fi.red64(last[n]), fix % qi, qi);
            if (ckks) ntt_fwd(c->limbs[i], N, d.data());
            const u64 *src = in + (p * L + i) * N;
            u64 *dst = out + (p * (L - 1) + i) * N;
            for (size_t n = 0; n < N; n++) dst[n] = fi.mul(submod(src[n], d[n], qi), inv);
        }
    }
}

// ------------------------------------------------------------------ K5: BFV multiply (BEHZ)
// Evaluator::multiply on BFV: R/src/benchmarks/bfv/seal_bfv_element_wise_benchmark.cpp:326,
// R/src/benchmarks/bfv/seal_bfv_dot_product_benchmark.cpp:312.  SURVEY.md A.12.
static void behz_extend(const orc_ctx *c, const u64 *x /*[L][N] coeff form*/, u64 *xq /*[L][N] NTT*/, u64 *xb /*[nBsk][N] NTT*/)
{
    const size_t N = c->N, L = c->K - 1, nb = c->bsk.size();
    memcpy(xq, x, L * N * sizeof(u64));
    for (size_t l = 0; l < L; l++) ntt_fwd(c->limbs[l], N, xq + l * N);
    // fastbconv_m_tilde: multiply by m_tilde mod q, convert to Bsk and to {m_tilde}
    std::vector<u64> tmp(L * N), tb((nb + 1) * N);
    for (size_t l = 0; l < L; l++) {
        u64 q = c->limbs[l].q, mt = c->m_tilde % q;
        const FastMod &fm = c->limbs[l].fm;
        for (size_t n = 0; n < N; n++) tmp[l * N + n] = fm.mul(x[l * N + n], mt);
    }
    c->q_to_bsk.convert(tmp.data(), tb.data(), N);
    {   // to m_tilde = 2^32
        const Conv &cv = c->q_to_mtilde;
        for (size_t n = 0; n < N; n++) {
            u64 acc = 0;
            for (size_t l = 0; l < L; l++) {
                u64 v = cv.fib[l].mul(tmp[l * N + n], cv.inv_punct[l]);
                acc += v * cv.mat[0][l];   // mod 2^64, then masked
            }
            tb[nb * N + n] = acc & 0xffffffffull;
        }
    }
    // sm_mrq: small Montgomery reduction mod q, result in Bsk
    const u64 mt = c->m_tilde, mt_half = mt >> 1;
    for (size_t i = 0; i < nb; i++) {
        const u64 p = c->bsk[i].q, pq = c->prod_q_mod_bsk[i], imt = c->inv_mtilde_mod_bsk[i];
        for (size_t n = 0; n < N; n++) {
            u64 r = (tb[nb * N + n] * c->neg_inv_prod_q_mod_mtilde) & 0xffffffffull;
            if (r >= mt_half) r += p - mt;
            u64 v = c->bsk[i].fm.mad(r, pq, tb[i * N + n]);
            xb[i * N + n] = c->bsk[i].fm.mul(v, imt);
        }
    }
    for (size_t i = 0; i < nb; i++) ntt_fwd(c->bsk[i], N, xb + i * N);
}

extern "C" void orc_bfv_multiply(const orc_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out)
{
    const size_t N = c->N, L = c->K - 1, nb = c->bsk.size(), nB = c->nB;
    std::vector<u64> aq(2 * L * N), ab(2 * nb * N), bq(2 * L * N), bb(2 * nb * N);
    for (size_t p = 0; p < 2; p++) {
        behz_extend(c, a + p * L * N, aq.data() + p * L * N, ab.data() + p * nb * N);
        behz_extend(c, b + p * L * N, bq.data() + p * L * N, bb.data() + p * nb * N);
    }
    std::vector<u64> dq(3 * L * N), db(3 * nb * N);
    auto tensor = [&](const std::vector<u64> &x, const std::vector<u64> &y, std::vector<u64> &d, size_t nl,
                      const std::vector<Limb> &base) {
        for (size_t l = 0; l < nl; l++) {
            const FastMod &fm = base[l].fm;
            const u64 *x0 = &x[l * N], *x1 = &x[(nl + l) * N], *y0 = &y[l * N], *y1 = &y[(nl + l) * N];
            u64 *d0 = &d[l * N], *d1 = &d[(nl + l) * N], *d2 = &d[(2 * nl + l) * N];
            for (size_t n = 0; n < N; n++) {
                d0[n] = fm.mul(x0[n], y0[n]);
                d1[n] = fm.mad(x0[n], y1[n], fm.mul(x1[n], y0[n]));
                d2[n] = fm.mul(x1[n], y1[n]);
            }
            ntt_inv(base[l], N, d0);
            ntt_inv(base[l], N, d1);
            ntt_inv(base[l], N, d2);
        }
    };
    tensor(aq, bq, dq, L, c->limbs);
    tensor(ab, bb, db, nb, c->bsk);
    std::vector<u64> tq(L * N), tb(nb * N), fl(nb * N), conv(nb * N), sk(N), outq(L * N);
    for (size_t p = 0; p < 3; p++) {
        // (6) multiply by t
        for (size_t l = 0; l < L; l++) {
            u64 q = c->limbs[l].q;
            for (size_t n = 0; n < N; n++) tq[l * N + n] = c->limbs[l].fm.mul(dq[(p * L + l) * N + n], c->t % q);
        }
        for (size_t i = 0; i < nb; i++) {
            u64 q = c->bsk[i].q;
            for (size_t n = 0; n < N; n++) tb[i * N + n] = c->bsk[i].fm.mul(db[(p * nb + i) * N + n], c->t % q);
        }
        // (7) fast_floor: (y_p - conv_{q->p}(y_q)) * q^{-1} mod p
        c->q_to_bsk.convert(tq.data(), conv.data(), N);
        for (size_t i = 0; i < nb; i++) {
            u64 q = c->bsk[i].q;
            for (size_t n = 0; n < N; n++)
                fl[i * N + n] = c->bsk[i].fm.mul(submod(tb[i * N + n], conv[i * N + n], q), c->inv_prod_q_mod_bsk[i]);
        }
        // (8) fastbconv_sk: Shenoy-Kumaresan Bsk -> q
        c->B_to_q.convert(fl.data(), outq.data(), N);
        c->B_to_msk.convert(fl.data(), sk.data(), N);
        const u64 msk = c->m_sk, msk_half = msk >> 1;
        for (size_t n = 0; n < N; n++)
            sk[n] = c->bsk[nB].fm.mul(submod(sk[n], fl[nB * N + n], msk), c->inv_prod_B_mod_msk);
        for (size_t l = 0; l < L; l++) {
            const u64 q = c->limbs[l].q, pB = c->prod_B_mod_q[l];
            u64 *dst = out + (p * L + l) * N;
            for (size_t n = 0; n < N; n++) {
                u64 al = sk[n];
                const FastMod &fm = c->limbs[l].fm;
                if (al > msk_half)
                    dst[n] = fm.mad(fm.red64(msk - al), pB, outq[l * N + n]);
                else
                    dst[n] = fm.mad(fm.red64(al), q - pB, outq[l * N + n]);
            }
        }
    }
}

// ------------------------------------------------------------------ composites (batched, OpenMP)
extern "C" int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

extern "C" void orc_batch_mul_relin_rescale(const orc_ctx *c, size_t L, size_t n, const uint64_t *a, const uint64_t *b,
                                            const uint64_t *relin_key, uint64_t *out, int threads)
{
    const size_t N = c->N, ct = 2 * L * N;
    if (threads <= 0) threads = orc_max_threads();
#pragma omp parallel for num_threads(threads)
    for (long i = 0; i < (long)n; i++) {
        std::vector<u64> t3(3 * L * N), t2(2 * L * N);
        orc_ckks_multiply(c, L, a + i * ct, b + i * ct, t3.data());
        orc_relinearize(c, L, t3.data(), relin_key, t2.data());
        orc_rescale(c, L, 2, t2.data(), out + i * 2 * (L - 1) * N);
    }
}

// accumulateCKKS (R/src/engine/seal_context.cpp:321-347) / accumulateBFV rows part (:289-304)
extern "C" int orc_accumulate(const orc_ctx *c, size_t L, uint64_t *ct, size_t count, const uint32_t *elts,
                              const uint64_t *const *keys, size_t nkeys)
{
    const size_t N = c->N, slots = N / 2;
    const bool ckks = c->scheme == ORC_SCHEME_CKKS;
    if (ckks && count > slots) count = slots;
    size_t row_count = (!ckks && count > slots) ? slots : count;
    if (count == 0) return -2;   // reference would encrypt_zero (randomised); not on the measured path
    int rot = 0;
    for (size_t v = row_count; v; v >>= 1) rot++;   // get_significant_bit_count
    if ((size_t(1) << (rot - 1)) == row_count) rot--;
    std::vector<u64> r(2 * L * N);
    for (int k = 0; k < rot; k++) {
        memcpy(r.data(), ct, 2 * L * N * sizeof(u64));
        int rc = orc_rotate(c, L, r.data(), 1 << k, elts, keys, nkeys);
        if (rc) return rc;
        orc_add(c, L, 2, ct, r.data(), ct);
    }
    if (!ckks && count > slots) {   // rotate_columns_inplace: R/src/engine/seal_context.cpp:305-310
        memcpy(r.data(), ct, 2 * L * N * sizeof(u64));
        uint32_t elt = (uint32_t)(2 * N - 1);
        const u64 *k = find_key(elt, elts, keys, nkeys);
        if (!k) return -1;
        orc_apply_galois(c, L, r.data(), elt, k);
        orc_add(c, L, 2, ct, r.data(), ct);
    }
    return 0;
}

extern "C" int orc_batch_dot(const orc_ctx *c, size_t L, size_t n, const uint64_t *a, const uint64_t *b, size_t count,
                             const uint64_t *relin_key, const uint32_t *elts, const uint64_t *const *keys, size_t nkeys,
                             uint64_t *out, int threads)
{
    const size_t N = c->N, ct = 2 * L * N;
    if (threads <= 0) threads = orc_max_threads();
    int err = 0;
#pragma omp parallel for num_threads(threads)
    for (long i = 0; i < (long)n; i++) {
        std::vector<u64> t3(3 * L * N);
        if (c->scheme == ORC_SCHEME_CKKS)
            orc_ckks_multiply(c, L, a + i * ct, b + i * ct, t3.data());
        else
            orc_bfv_multiply(c, a + i * ct, b + i * ct, t3.data());
        orc_relinearize(c, L, t3.data(), relin_key, out + i * ct);
        int rc = orc_accumulate(c, L, out + i * ct, count, elts, keys, nkeys);
        if (rc) err = rc;
    }
    return err;
}

extern "C" void orc_batch_ntt(const orc_ctx *c, size_t limb, size_t n, uint64_t *polys, int inverse, int threads)
{
    if (threads <= 0) threads = orc_max_threads();
#pragma omp parallel for num_threads(threads)
    for (long i = 0; i < (long)n; i++) {
        if (inverse)
            ntt_inv(c->limbs[limb], c->N, polys + i * c->N);
        else
            ntt_fwd(c->limbs[limb], c->N, polys + i * c->N);
    }
}
