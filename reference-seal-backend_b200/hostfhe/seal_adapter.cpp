// seal_adapter.cpp -- the hostfhe.h interface implemented over Microsoft SEAL.
//
// This is the production form of the host side: the reference keeps seal::{KeyGenerator, Encryptor, Decryptor,
// CKKSEncoder, BatchEncoder} for keygen / encode / encrypt / decrypt / decode (R/src/engine/seal_context.cpp:46-70,
// R/include/engine/seal_context.h:52-62), and so does the B200 backend when SEAL is installed: configure the backend
// with -DSEAL_INSTALL_DIR=<prefix> (backend/CMakeLists.txt, mirroring R/cmake/utils/import-library.cmake:54-66) and this
// file replaces hostfhe.cpp behind the same C interface -- b200_context.cpp does not change.  With real SEAL keys and
// real SEAL encryptions going through libb200he.so, `test_harness --backend_lib_path libhebench_seal_backend.so` is the
// literal run the north star describes, and comparing store()'s ciphertexts with seal::Evaluator's on the same inputs
// pins the oracle.
//
// NOT COMPILED IN THIS REPOSITORY'S OFFLINE BUILD: SEAL (v3.7.2, R/cmake/third-party/SEAL.version) is fetched from the
// network by the reference's CMake and is absent here, so this translation unit is guarded by B200HE_WITH_SEAL and has
// not been compiled or run; it is written against the public SEAL >= 3.6 API the reference itself uses
// (create_public_key(pk), SEALContext by value, lowercase scheme_type).
#ifdef B200HE_WITH_SEAL
#include "hostfhe.h"

#include <seal/seal.h>

#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

struct hfhe_ctx {
    int scheme = 0;
    size_t N = 0, K = 0;
    double scale = 1.0;
    std::unique_ptr<seal::SEALContext> context;
    seal::SecretKey sk;
    seal::PublicKey pk;
    seal::RelinKeys relin;
    seal::GaloisKeys galois;
    std::unique_ptr<seal::Encryptor> encryptor;
    std::unique_ptr<seal::Decryptor> decryptor;
    std::unique_ptr<seal::CKKSEncoder> ckks_encoder;
    std::unique_ptr<seal::BatchEncoder> batch_encoder;
    std::vector<uint64_t> moduli, psi, relin_packed;
    std::vector<uint32_t> galois_elts;
    std::map<uint32_t, std::vector<uint64_t>> galois_packed;
    std::mutex galois_mtx;
    std::atomic<uint64_t> enc_streams{ 0 };

    // seal::KSwitchKeys::data()[index] is std::vector<PublicKey>: one size-2 key-level ciphertext [2][K][N] per digit.
    // libb200he wants the L_top digits contiguously: uint64[L_top][2][K][N] (include/b200he.h).
    std::vector<uint64_t> pack(const std::vector<seal::PublicKey> &digits) const
    {
        std::vector<uint64_t> buf;
        buf.reserve(digits.size() * 2 * K * N);
        for (const seal::PublicKey &d : digits) buf.insert(buf.end(), d.data().data(), d.data().data() + 2 * K * N);
        return buf;
    }
    // parms_id of the data level with L primes
    seal::parms_id_type level_parms(size_t L) const
    {
        auto cd = context->first_context_data();
        while (cd && cd->parms().coeff_modulus().size() > L) cd = cd->next_context_data();
        if (!cd || cd->parms().coeff_modulus().size() != L) throw std::invalid_argument("no such level");
        return cd->parms_id();
    }
};

extern "C" hfhe_ctx *hfhe_create(int scheme, size_t N, size_t depth, int coeff_bits, int scale_or_plain_bits, uint64_t seed)
{
    try {
        std::unique_ptr<hfhe_ctx> c(new hfhe_ctx);
        c->scheme = scheme;
        c->N      = N;
        c->K      = depth + 1;
        // R/src/engine/seal_context.cpp:79-89 (CKKS) / :107-118 (BFV): {60, bits x (depth - 1), 60}
        std::vector<int> bits(depth + 1, coeff_bits);
        bits.front() = 60;
        bits.back()  = 60;
        seal::EncryptionParameters parms(scheme == HFHE_CKKS ? seal::scheme_type::ckks : seal::scheme_type::bfv);
        parms.set_poly_modulus_degree(N);
        parms.set_coeff_modulus(seal::CoeffModulus::Create(N, bits));
        if (scheme == HFHE_BFV) parms.set_plain_modulus(seal::PlainModulus::Batching(N, scale_or_plain_bits));
        if (seed != 0) {   // reproducible keys / encryptions (parity tests only); default: SEAL's own OS-seeded generator
            seal::prng_seed_type s{};
            for (auto &w : s) w = (seed += 0x9e3779b97f4a7c15ull);
            parms.set_random_generator(std::make_shared<seal::Blake2xbPRNGFactory>(s));
        }
        c->context.reset(new seal::SEALContext(parms, true, seal::sec_level_type::tc128));
        if (!c->context->parameters_set()) return nullptr;
        c->scale = scheme == HFHE_CKKS ? std::pow(2.0, scale_or_plain_bits) : 1.0;
        // R/src/engine/seal_context.cpp:46-70
        seal::KeyGenerator keygen(*c->context);
        c->sk = keygen.secret_key();
        keygen.create_public_key(c->pk);
        keygen.create_relin_keys(c->relin);
        keygen.create_galois_keys(c->galois);
        c->encryptor.reset(new seal::Encryptor(*c->context, c->pk));
        c->decryptor.reset(new seal::Decryptor(*c->context, c->sk));
        if (scheme == HFHE_CKKS) c->ckks_encoder.reset(new seal::CKKSEncoder(*c->context));
        else c->batch_encoder.reset(new seal::BatchEncoder(*c->context));
        const auto &kd = *c->context->key_context_data();
        for (size_t i = 0; i < c->K; ++i) {
            c->moduli.push_back(kd.parms().coeff_modulus()[i].value());
            c->psi.push_back(kd.small_ntt_tables()[i].get_root());   // SEAL's minimal primitive 2N-th root
        }
        c->relin_packed = c->pack(c->relin.data()[0]);
        for (uint32_t elt : kd.galois_tool()->get_elts_all())
            if (c->galois.has_key(elt)) c->galois_elts.push_back(elt);
        return c.release();
    } catch (...) {
        return nullptr;
    }
}
extern "C" void hfhe_destroy(hfhe_ctx *c) { delete c; }
extern "C" size_t hfhe_N(const hfhe_ctx *c) { return c->N; }
extern "C" size_t hfhe_K(const hfhe_ctx *c) { return c->K; }
extern "C" const uint64_t *hfhe_moduli(const hfhe_ctx *c) { return c->moduli.data(); }
extern "C" const uint64_t *hfhe_psi(const hfhe_ctx *c) { return c->psi.data(); }
extern "C" uint64_t hfhe_plain_modulus(const hfhe_ctx *c)
{
    return c->scheme == HFHE_BFV ? c->context->first_context_data()->parms().plain_modulus().value() : 0;
}
extern "C" double hfhe_scale(const hfhe_ctx *c) { return c->scale; }
extern "C" const uint64_t *hfhe_relin_key(hfhe_ctx *c) { return c->relin_packed.data(); }
extern "C" size_t hfhe_galois_count(hfhe_ctx *c) { return c->galois_elts.size(); }
extern "C" uint32_t hfhe_galois_elt(hfhe_ctx *c, size_t i) { return c->galois_elts[i]; }
extern "C" size_t hfhe_kswitch_key_words(const hfhe_ctx *c) { return (c->K - 1) * 2 * c->K * c->N; }
extern "C" const uint64_t *hfhe_galois_key(hfhe_ctx *c, uint32_t elt)
{
    if (!c->galois.has_key(elt)) return nullptr;
    std::lock_guard<std::mutex> lock(c->galois_mtx);
    auto it = c->galois_packed.find(elt);
    if (it == c->galois_packed.end()) it = c->galois_packed.emplace(elt, c->pack(c->galois.key(elt))).first;
    return it->second.data();
}

// CKKSEncoder::encode at the top data level: plain [L_top][N], NTT form (R/src/engine/seal_context.cpp:265-276)
extern "C" void hfhe_ckks_encode(hfhe_ctx *c, const double *vals, size_t n, double scale, uint64_t *plain)
{
    seal::Plaintext p;
    c->ckks_encoder->encode(std::vector<double>(vals, vals + n), scale, p);
    std::memcpy(plain, p.data(), (c->K - 1) * c->N * sizeof(uint64_t));
}
extern "C" void hfhe_ckks_decode(hfhe_ctx *c, const uint64_t *plain, size_t L, double scale, double *out)
{
    seal::Plaintext p;
    p.parms_id() = seal::parms_id_zero;
    p.resize(L * c->N);
    std::memcpy(p.data(), plain, L * c->N * sizeof(uint64_t));
    p.parms_id() = c->level_parms(L);
    p.scale()    = scale;
    std::vector<double> v;
    c->ckks_encoder->decode(p, v);
    std::memcpy(out, v.data(), (c->N / 2) * sizeof(double));
}
// BatchEncoder (R/src/engine/seal_context.cpp:278-287): plain = N coefficients mod t
extern "C" void hfhe_bfv_encode(hfhe_ctx *c, const int64_t *vals, size_t n, uint64_t *plain)
{
    seal::Plaintext p;
    c->batch_encoder->encode(std::vector<int64_t>(vals, vals + n), p);
    std::memset(plain, 0, c->N * sizeof(uint64_t));
    std::memcpy(plain, p.data(), std::min(p.coeff_count(), c->N) * sizeof(uint64_t));
}
extern "C" void hfhe_bfv_decode(hfhe_ctx *c, const uint64_t *plain, int64_t *out)
{
    seal::Plaintext p(c->N);
    std::memcpy(p.data(), plain, c->N * sizeof(uint64_t));
    std::vector<int64_t> v;
    c->batch_encoder->decode(p, v);
    std::memcpy(out, v.data(), c->N * sizeof(int64_t));
}

// Encryptor::encrypt at the top data level: ct [2][L_top][N] (CKKS: NTT form; BFV: coefficient form).
// SEAL draws its own randomness per call (thread safe); the stream index of the stand-in has no meaning here.
extern "C" uint64_t hfhe_reserve_encryptions(hfhe_ctx *c, uint64_t n) { return c->enc_streams.fetch_add(n); }
extern "C" void hfhe_encrypt_at(hfhe_ctx *c, const uint64_t *plain, uint64_t *ct, uint64_t) { hfhe_encrypt(c, plain, ct); }
extern "C" void hfhe_encrypt(hfhe_ctx *c, const uint64_t *plain, uint64_t *ct)
{
    const size_t L = c->K - 1, N = c->N;
    seal::Plaintext p;
    if (c->scheme == HFHE_CKKS) {
        p.parms_id() = seal::parms_id_zero;
        p.resize(L * N);
        std::memcpy(p.data(), plain, L * N * sizeof(uint64_t));
        p.parms_id() = c->context->first_parms_id();
        p.scale()    = c->scale;   // metadata only: the ciphertext bits do not depend on it
    } else {
        p.resize(N);
        std::memcpy(p.data(), plain, N * sizeof(uint64_t));
    }
    seal::Ciphertext out;
    c->encryptor->encrypt(p, out);
    std::memcpy(ct, out.data(), 2 * L * N * sizeof(uint64_t));   // seal::Ciphertext::data() is uint64[size][L][N]
}
// Decryptor::decrypt of a size-`size` ciphertext at level L
extern "C" void hfhe_decrypt(hfhe_ctx *c, const uint64_t *ct, size_t size, size_t L, uint64_t *plain)
{
    const size_t N = c->N;
    seal::Ciphertext in;
    in.resize(*c->context, c->level_parms(L), size);
    std::memcpy(in.data(), ct, size * L * N * sizeof(uint64_t));
    in.is_ntt_form() = c->scheme == HFHE_CKKS;
    in.scale()       = 1.0;   // metadata only; the caller's decode supplies the real scale
    seal::Plaintext p;
    c->decryptor->decrypt(in, p);
    if (c->scheme == HFHE_CKKS)
        std::memcpy(plain, p.data(), L * N * sizeof(uint64_t));
    else {
        std::memset(plain, 0, N * sizeof(uint64_t));
        std::memcpy(plain, p.data(), std::min(p.coeff_count(), N) * sizeof(uint64_t));
    }
}
#endif   // B200HE_WITH_SEAL
