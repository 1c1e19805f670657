// behz_host.inl -- host side of the BFV (BEHZ) multiply: auxiliary-base selection and constants
// (SEAL util/rns.cpp RNSTool::initialize restated) and the launch sequence.
// Included by b200he.cu.

static int init_behz(b200he_ctx *c, std::vector<u64> &moduli, std::vector<u64> &psi)
{
    using namespace hm;
    const size_t L = c->K - 1, N = c->N;
    if (L > (size_t)BEHZ_MAXL) return fail("ctx_create: BFV multiply supports at most %d data primes", BEHZ_MAXL);
    std::vector<u64> q(moduli.begin(), moduli.begin() + L);
    // |B| = L, one more when q*t*2^32 does not fit (SEAL RNSTool::initialize)
    size_t nB = L;
    if (32 + (size_t)bitlen(c->t) + product_bits(q) >= 61 * L + 61) nB++;
    const size_t nbsk = nB + 1;
    if (nbsk > (size_t)BEHZ_MAXB) return fail("ctx_create: BEHZ auxiliary base too large");
    std::vector<u64> pr = primes_below(2 * (u64)N, 61, nB + 2);
    if (pr.size() != nB + 2) return fail("ctx_create: not enough 61-bit NTT primes for the BEHZ base");
    const u64 m_sk = pr[0];   // pr[1] is gamma (decryption only)
    std::vector<u64> B(pr.begin() + 2, pr.end()), bsk(B);
    bsk.push_back(m_sk);
    for (u64 p : bsk) {
        moduli.push_back(p);
        const u64 r = any_primitive_root(2 * (u64)N, p);
        if (!r) return fail("ctx_create: no primitive root for auxiliary prime");
        psi.push_back(r);
    }
    BehzConst &h = c->behz.h;
    memset(&h, 0, sizeof h);
    h.L = (int)L; h.nB = (int)nB; h.nbsk = (int)nbsk; h.K = (int)c->K;
    const u64 mt = u64(1) << 32;
    u32 qmt = 1;
    for (size_t l = 0; l < L; l++) {
        u64 punct = 1;
        u32 pmt = 1;
        for (size_t j = 0; j < L; j++)
            if (j != l) { punct = mulmod(punct, q[j] % q[l], q[l]); pmt *= (u32)q[j]; }
        const u64 invp = invmod(punct, q[l]);
        h.mt_invp_q[l] = mulmod(mt % q[l], invp, q[l]);
        h.t_invp_q[l] = mulmod(c->t % q[l], invp, q[l]);
        h.q2mt[l] = pmt;
        qmt *= (u32)q[l];
        u64 pb = 1;
        for (u64 x : B) pb = mulmod(pb, x % q[l], q[l]);
        h.prod_B_q[l] = pb;
        for (size_t j = 0; j < nB; j++) {
            u64 p = 1;
            for (size_t k = 0; k < nB; k++)
                if (k != j) p = mulmod(p, B[k] % q[l], q[l]);
            h.B2q[l][j] = p;
        }
    }
    u32 inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - qmt * inv;   // Newton: q^{-1} mod 2^32
    h.neg_inv_q_mt = 0u - inv;
    for (size_t k = 0; k < nbsk; k++) {
        const u64 p = bsk[k];
        u64 pq = 1;
        for (u64 x : q) pq = mulmod(pq, x % p, p);
        h.prod_q_bsk[k] = pq;
        h.inv_prod_q_bsk[k] = invmod(pq, p);
        h.inv_mt_bsk[k] = invmod(mt % p, p);
        h.t_bsk[k] = c->t % p;
        for (size_t l = 0; l < L; l++) {
            u64 v = 1;
            for (size_t j = 0; j < L; j++)
                if (j != l) v = mulmod(v, q[j] % p, p);
            h.q2bsk[k][l] = v;
        }
    }
    u64 pbm = 1;
    for (u64 x : B) pbm = mulmod(pbm, x % m_sk, m_sk);
    h.inv_prod_B_msk = invmod(pbm, m_sk);
    for (size_t j = 0; j < nB; j++) {
        u64 punct = 1, pm = 1;
        for (size_t k = 0; k < nB; k++)
            if (k != j) { punct = mulmod(punct, B[k] % B[j], B[j]); pm = mulmod(pm, B[k] % m_sk, m_sk); }
        h.invp_B[j] = invmod(punct, B[j]);
        h.B2msk[j] = pm;
    }
    c->nBsk = (int)nbsk;
    return 0;
}

static int upload_behz(b200he_ctx *c)
{
    CK(cudaMalloc(&c->behz.d_blob, sizeof(BehzConst)));
    CK(cudaMemcpy(c->behz.d_blob, &c->behz.h, sizeof(BehzConst), cudaMemcpyHostToDevice));
    return 0;
}

// out[i] = a[ai[i]] * b[bi[i]]  (size 3, coefficient form), chunked to the workspace budget
static int bfv_multiply(b200he_ctx *c, const b200he_batch *a, const u32 *dai, const b200he_batch *b, const u32 *dbi, uint64_t n, u64 *out)
{
    const size_t N = c->N, L = c->K - 1, nb = c->nBsk, W = L + nb, K = c->K;
    const size_t per = (4 + 4) * W * N * 8;   // extended operands (coeff form), their transforms (reused for the products)
    size_t chunk = c->workspace / per;
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    u64 *ws = (u64 *)c->pool.get(chunk * per);
    if (!ws) return fail("multiply: out of device memory for BEHZ workspace");
    u64 *ext = ws, *tr = ws + chunk * 4 * W * N;
    int rc = 0;
    for (size_t i0 = 0; i0 < n && !rc; i0 += chunk) {
        const size_t m = n - i0 < chunk ? n - i0 : chunk;
        BehzExtArgs A{};
        A.a = a->d; A.b = b->d; A.ai = dai ? dai + i0 : nullptr; A.bi = dbi ? dbi + i0 : nullptr;
        if (!dai) A.a += i0 * a->ct_words();
        if (!dbi) A.b += i0 * b->ct_words();
        A.a_stride = a->ct_words(); A.b_stride = b->ct_words(); A.ext = ext; A.n = m;
        LAUNCH(c, B200HE_KERN_BEHZ, k_behz_extend, blocks_for(m * 4 * N / 2), 256, 0, c->T, c->behz.d(), A);
        if (cudaGetLastError() != cudaSuccess) { rc = fail("multiply: k_behz_extend launch failed"); break; }
        // transforms in q (mod ids 0..L-1) and in Bsk (mod ids K..K+nb-1)
        if ((rc = ntt_fwd(c, ext, tr, m * 4 * L, W * N, W * N, (int)L, 0))) break;
        if ((rc = ntt_fwd(c, ext + L * N, tr + L * N, m * 4 * nb, W * N, W * N, (int)nb, (int)K))) break;
        u64 *prod = ext;   // the coefficient-form operands are dead: reuse for the 3 products
        LAUNCH(c, B200HE_KERN_BEHZ, k_behz_tensor, blocks_for(m * W * N / 2), 256, 0, c->T, c->behz.d(), tr, prod, m);
        if (cudaGetLastError() != cudaSuccess) { rc = fail("multiply: k_behz_tensor launch failed"); break; }
        if ((rc = ntt_inv(c, prod, prod, m * 3 * L, W * N, W * N, (int)L, 0, INV_PLAIN))) break;
        if ((rc = ntt_inv(c, prod + L * N, prod + L * N, m * 3 * nb, W * N, W * N, (int)nb, (int)K, INV_PLAIN))) break;
        LAUNCH(c, B200HE_KERN_BEHZ, k_behz_floor_sk, blocks_for(m * 3 * N / 2), 256, 0, c->T, c->behz.d(), prod, out + i0 * 3 * L * N, m);
        if (cudaGetLastError() != cudaSuccess) { rc = fail("multiply: k_behz_floor_sk launch failed"); break; }
    }
    c->pool.put(ws);
    return rc;
}
