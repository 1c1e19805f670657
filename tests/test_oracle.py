"""Pins the CPU oracle (oracle/, the checker of every parity test) against everything that can be known about
the reference's arithmetic without running SEAL (the reference ships no golden vectors):
  * tests/golden/known_answers.json -- prime chains, plain moduli, minimal 2N-th roots, NAF vectors, Galois
    elements, recomputed by tests/golden/make_known_answers.py with plain Python integers;
  * the mathematical definitions, evaluated with Python big integers: the O(N^2) transform, schoolbook
    negacyclic products, CRT statements of rescale / BFV mod-switch, Galois automorphisms;
  * decrypt-level semantics under real keys from the host stand-in (multiply, relinearize, rescale, rotate,
    accumulate; BFV multiply) -- a wrong key-switch or BEHZ step decrypts to noise.
CPU only; runs in seconds."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from helpers import BFV, CKKS, Host, Oracle, oracle, p64, rand_residues, u32p

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "known_answers.json")) as f:
    KA = json.load(f)


def chain(N, bits):
    arr = (C.c_int * len(bits))(*bits)
    out = np.zeros(len(bits), dtype=np.uint64)
    assert oracle().orc_coeff_modulus_create(N, arr, len(bits), p64(out)) == 0
    return out


# ------------------------------------------------------------------------------- known answers
@pytest.mark.parametrize("name", sorted(KA["chains"]))
def test_prime_chains(name):
    e = KA["chains"][name]
    assert [hex(int(q)) for q in chain(e["N"], e["bits"])] == e["moduli"]


def test_plain_modulus():
    for N, t in KA["plain_modulus"].items():
        assert int(oracle().orc_plain_modulus_batching(int(N), 20)) == t


def test_minimal_roots():
    for key, psi in KA["psi"].items():
        N, q = key.split(":")
        root = C.c_uint64()
        assert oracle().orc_minimal_primitive_root(2 * int(N), int(q, 16), C.byref(root)) == 0
        assert root.value == psi, key


def test_naf():
    for v, terms in KA["naf"].items():
        out = (C.c_int * 64)()
        oracle().orc_naf.argtypes = [C.c_int, C.POINTER(C.c_int)]
        n = oracle().orc_naf(int(v), out)
        assert list(out[:n]) == terms, v


def test_galois_elements():
    for N, d in KA["galois_elt_from_step"].items():
        for s, e in d.items():
            assert oracle().orc_galois_elt_from_step(int(s), int(N)) == e, (N, s)
    for N, elts in KA["galois_elts_all"].items():
        out = np.zeros(64, dtype=np.uint32)
        n = oracle().orc_galois_elts_all(int(N), out.ctypes.data_as(u32p))
        assert sorted(out[:n].tolist()) == sorted(elts)


# ------------------------------------------------------------------------------- definitions
N_SMALL = 1024


@pytest.fixture(scope="module")
def small():
    moduli = chain(N_SMALL, [60, 45, 40, 60])
    return Oracle(CKKS, N_SMALL, moduli), [int(q) for q in moduli]


def brv(i, bits):
    return int(format(i, f"0{bits}b")[::-1], 2)


def test_ntt_is_the_definition(small):
    """out[i] = sum_j a_j psi^((2 brv(i) + 1) j)  (SURVEY Appendix A.4), and the inverse undoes it"""
    orc, moduli = small
    rng = np.random.default_rng(7)
    for l, q in enumerate(moduli):
        a = rng.integers(0, q, size=N_SMALL, dtype=np.uint64)
        got = orc.ntt(l, a)
        direct = np.empty(N_SMALL, dtype=np.uint64)
        orc.o.orc_ntt_fwd_direct(orc.c, l, p64(a), p64(direct))
        assert np.array_equal(got, direct)
        psi = int(orc.o.orc_ctx_psi(orc.c, l))
        for i in (0, 1, 5, N_SMALL - 1):   # spot-check the definition itself with Python integers
            w = pow(psi, 2 * brv(i, 10) + 1, q)
            acc, p = 0, 1
            for j in range(N_SMALL):
                acc = (acc + int(a[j]) * p) % q
                p = p * w % q
            assert int(got[i]) == acc
        assert np.array_equal(orc.ntt(l, got, inverse=True), a)


def negacyclic(a, b, q):
    n = len(a)
    out = [0] * n
    for i, ai in enumerate(a):
        if ai == 0:
            continue
        for j, bj in enumerate(b):
            k = i + j
            if k < n:
                out[k] = (out[k] + ai * bj) % q
            else:
                out[k - n] = (out[k - n] - ai * bj) % q
    return out


def test_dyadic_product_is_negacyclic_convolution(small):
    orc, moduli = small
    rng = np.random.default_rng(8)
    n_nz = 48   # sparse operand keeps the O(N^2) check fast
    for l, q in enumerate(moduli[:2]):
        a = np.zeros(N_SMALL, dtype=np.uint64)
        a[rng.choice(N_SMALL, n_nz, replace=False)] = rng.integers(1, q, size=n_nz, dtype=np.uint64)
        b = rng.integers(0, q, size=N_SMALL, dtype=np.uint64)
        fa, fb = orc.ntt(l, a), orc.ntt(l, b)
        prod = np.array([int(x) * int(y) % q for x, y in zip(fa, fb)], dtype=np.uint64)
        assert orc.ntt(l, prod, inverse=True).tolist() == negacyclic([int(v) for v in a], [int(v) for v in b], q)


def crt(residues, moduli):
    M = 1
    for q in moduli:
        M *= q
    x = 0
    for r, q in zip(residues, moduli):
        Mi = M // q
        x += int(r) * Mi * pow(Mi, -1, q)
    return x % M


def test_rescale_is_rounded_division(small):
    """divide_and_round_q_last: out = floor((x + floor(q_last/2)) / q_last) per coefficient"""
    orc, moduli = small
    rng = np.random.default_rng(9)
    L = 3
    ct = rand_residues(rng, np.array(moduli[:L], dtype=np.uint64), (2,), N_SMALL)   # coefficient form
    ntt = np.stack([[orc.ntt(l, ct[p, l]) for l in range(L)] for p in range(2)])
    out = orc.rescale(L, 2, np.ascontiguousarray(ntt).reshape(-1)).reshape(2, L - 1, N_SMALL)
    for p in range(2):
        coeff = [orc.ntt(l, out[p, l], inverse=True) for l in range(L - 1)]
        for k in range(0, N_SMALL, 37):
            x = crt([ct[p, l, k] for l in range(L)], moduli[:L])
            want = (x + moduli[L - 1] // 2) // moduli[L - 1]
            for l in range(L - 1):
                assert int(coeff[l][k]) == want % moduli[l]


def test_bfv_modswitch_is_rounded_division():
    N = 1024
    moduli = chain(N, [60, 40, 60])
    t = int(oracle().orc_plain_modulus_batching(N, 20))
    orc = Oracle(BFV, N, moduli, t)
    rng = np.random.default_rng(10)
    ct = rand_residues(rng, moduli[:2], (2,), N)
    out = orc.rescale(2, 2, ct.reshape(-1)).reshape(2, 1, N)
    q = [int(v) for v in moduli]
    for p in range(2):
        for k in range(0, N, 29):
            x = crt([ct[p, 0, k], ct[p, 1, k]], q[:2])
            assert int(out[p, 0, k]) == ((x + q[1] // 2) // q[1]) % q[0]


def test_galois_ntt_table_matches_coefficient_automorphism(small):
    """apply_galois_ntt(NTT(a)) == NTT(a(x^elt))  """
    orc, moduli = small
    rng = np.random.default_rng(11)
    q = moduli[1]
    a = rng.integers(0, q, size=N_SMALL, dtype=np.uint64)
    for elt in (3, 9, 2 * N_SMALL - 1, pow(3, N_SMALL // 2 - 1, 2 * N_SMALL)):
        table = np.zeros(N_SMALL, dtype=np.uint32)
        orc.o.orc_galois_table_ntt(N_SMALL, elt, table.ctypes.data_as(u32p))
        b = np.zeros(N_SMALL, dtype=np.uint64)
        for i in range(N_SMALL):
            raw = i * elt
            v = int(a[i])
            b[raw % N_SMALL] = (q - v) % q if (raw // N_SMALL) & 1 else v
        assert np.array_equal(orc.ntt(1, a)[table], orc.ntt(1, b))


# ------------------------------------------------------------------------------- decrypt-level semantics
@pytest.fixture(scope="module")
def ckks_host():
    host = Host(CKKS, 2048, 3, 40, 40, seed=5)
    return host, Oracle(CKKS, 2048, host.moduli)


def test_ckks_multiply_relinearize_rescale_decrypts_to_the_product(ckks_host):
    host, orc = ckks_host
    rng = np.random.default_rng(12)
    L = host.Ltop
    x, y = rng.uniform(-1, 1, 64), rng.uniform(-1, 1, 64)
    ct3 = orc.ckks_multiply(L, host.enc_vec(x), host.enc_vec(y))
    assert np.allclose(host.dec_vec(ct3, 3, L, host.scale ** 2)[:64], x * y, atol=1e-5)
    ct2 = orc.relinearize(L, ct3, host.relin_key())
    assert np.allclose(host.dec_vec(ct2, 2, L, host.scale ** 2)[:64], x * y, atol=1e-5)
    ct1 = orc.rescale(L, 2, ct2)
    assert np.allclose(host.dec_vec(ct1, 2, L - 1, host.scale ** 2 / float(host.moduli[L - 1]))[:64], x * y, atol=1e-5)
    # one level down: the key switch at L-1 uses the same keys (digits 0..L-2 and the special prime)
    sq = orc.relinearize(L - 1, orc.ckks_multiply(L - 1, ct1, ct1), host.relin_key())
    s1 = host.scale ** 2 / float(host.moduli[L - 1])
    assert np.allclose(host.dec_vec(sq, 2, L - 1, s1 * s1)[:64], (x * y) ** 2, atol=1e-4)


def test_ckks_rotation_and_accumulate_decrypt_correctly(ckks_host):
    host, orc = ckks_host
    rng = np.random.default_rng(13)
    L, slots = host.Ltop, host.N // 2
    x = rng.uniform(-1, 1, slots)
    ct = host.enc_vec(x)
    keys = {e: host.galois_key(e) for e in host.galois_elts()}
    for step in (1, -1, 4, 100, -3):   # 100 and -3 go through the NAF fallback
        got = host.dec_vec(orc.rotate(L, ct, step, keys), 2, L)
        assert np.allclose(got, np.roll(x, -step), atol=1e-5), step
    v = np.zeros(slots)
    v[:9] = rng.uniform(-1, 1, 9)
    acc = host.dec_vec(orc.accumulate(L, host.enc_vec(v), 9, keys), 2, L)
    assert abs(acc[0] - v.sum()) < 1e-4


def test_bfv_multiply_and_rotate_decrypt_correctly():
    host = Host(BFV, 2048, 2, 40, 20, seed=6)
    orc = Oracle(BFV, 2048, host.moduli, host.t)
    rng = np.random.default_rng(14)
    L, N = host.Ltop, host.N
    x, y = rng.integers(-10, 11, N), rng.integers(-10, 11, N)
    ct3 = orc.bfv_multiply(host.enc_vec(x), host.enc_vec(y))
    assert np.array_equal(host.dec_vec(ct3, 3, L), x * y)
    ct2 = orc.relinearize(L, ct3, host.relin_key())
    assert np.array_equal(host.dec_vec(ct2, 2, L), x * y)
    keys = {e: host.galois_key(e) for e in host.galois_elts()}
    rot = host.dec_vec(orc.rotate(L, ct2, 3, keys), 2, L)
    want = np.concatenate([np.roll((x * y)[:N // 2], -3), np.roll((x * y)[N // 2:], -3)])
    assert np.array_equal(rot, want)
    down = orc.rescale(L, 2, ct2)   # BFV mod_switch_to_next keeps the plaintext
    assert np.array_equal(host.dec_vec(down, 2, L - 1), x * y)


# ------------------------------------------------------------------------------- exact statements: key switch, BEHZ
N_TINY = 64   # big-integer polynomial arithmetic in pure Python


def negacyclic_full(a, b, q):
    n = len(a)
    out = [0] * n
    for i, ai in enumerate(a):
        for j, bj in enumerate(b):
            k = i + j
            if k < n:
                out[k] += ai * bj
            else:
                out[k - n] -= ai * bj
    return [v % q for v in out]


@pytest.mark.parametrize("scheme", [CKKS, BFV], ids=["ckks", "bfv"])
def test_switch_key_is_rounded_division_by_the_special_prime(scheme):
    """Evaluator::switch_key_inplace, as one statement over the integers: with d_J the NON-centred digits of
    the target (coefficient form, 0 <= d_J < q_J), K_{J,k} the key polynomials and P the special prime,
        A_k   = sum_J d_J (*) K_{J,k}  mod (q_0 ... q_{L-1} P)      (negacyclic products, canonical representative)
        out_k = c_k + floor((A_k + floor(P/2)) / P)   mod q_j
    -- a rounding or digit-lift convention that differs by one from this shows up in the low bits, which the decrypt-level
    tests cannot see."""
    N = N_TINY
    moduli = chain(N, [60, 45, 40, 60])
    q = [int(v) for v in moduli]
    K, P = len(q), q[-1]
    t = int(oracle().orc_plain_modulus_batching(N, 20)) if scheme == BFV else 0
    orc = Oracle(scheme, N, moduli, t)
    rng = np.random.default_rng(21)
    key = rand_residues(rng, moduli, (K - 1, 2), N)                      # [Ltop][2][K][N], NTT form
    key_coeff = [[[[int(v) for v in orc.ntt(i, key[J, k, i], inverse=True)] for i in range(K)] for k in range(2)] for J in range(K - 1)]
    for L in (K - 1, K - 2):
        ct = rand_residues(rng, moduli[:L], (2,), N)
        target = rand_residues(rng, moduli[:L], (), N)
        got = orc.switch_key(L, ct.reshape(-1), target.reshape(-1), key.reshape(-1)).reshape(2, L, N)
        if scheme == CKKS:   # ciphertext and target are in NTT form
            digits = [[int(v) for v in orc.ntt(J, target[J], inverse=True)] for J in range(L)]
            c_coeff = [[[int(v) for v in orc.ntt(j, ct[k, j], inverse=True)] for j in range(L)] for k in range(2)]
            got_coeff = [[[int(v) for v in orc.ntt(j, got[k, j], inverse=True)] for j in range(L)] for k in range(2)]
        else:
            digits = [[int(v) for v in target[J]] for J in range(L)]
            c_coeff = [[[int(v) for v in ct[k, j]] for j in range(L)] for k in range(2)]
            got_coeff = [[[int(v) for v in got[k, j]] for j in range(L)] for k in range(2)]
        out_mods = q[:L] + [P]
        for k in range(2):
            acc = []   # A_k modulo each of q_0..q_{L-1}, P
            for idx, qi in enumerate(out_mods):
                ki = idx if idx < L else K - 1
                s = [0] * N
                for J in range(L):
                    prod = negacyclic_full(digits[J], key_coeff[J][k][ki], qi)
                    s = [(x + y) % qi for x, y in zip(s, prod)]
                acc.append(s)
            for n in range(N):
                A = crt([acc[i][n] for i in range(L + 1)], out_mods)
                rounded = (A + P // 2) // P
                for j in range(L):
                    assert got_coeff[k][j][n] == (c_coeff[k][j][n] + rounded) % q[j], (scheme, L, k, j, n)


def is_prime(n):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def fastbconv(res, base, p):
    """Bajard et al. fast base conversion: sum_i [x_i (b/b_i)^-1]_{b_i} (b/b_i) mod p, over the integers"""
    b = 1
    for v in base:
        b *= v
    tot = 0
    for x, bi in zip(res, base):
        punct = b // bi
        tot += (x * pow(punct, -1, bi) % bi) * punct
    return tot % p


def test_bfv_multiply_is_behz_step_by_step():
    """Evaluator::bfv_multiply restated independently with Python integers -- base extension with the small
    Montgomery reduction, tensor product in q and Bsk, multiplication by t, fast floor, Shenoy-Kumaresan conversion --
    and compared with the oracle word for word.  (BEHZ is approximate by design: the statement is the algorithm, step by
    step, not round(t x y / q).)"""
    N = N_TINY
    moduli = chain(N, [60, 40, 60])
    q = [int(v) for v in moduli[:2]]
    L = len(q)
    t = int(oracle().orc_plain_modulus_batching(N, 20))
    orc = Oracle(BFV, N, moduli, t)
    Q = q[0] * q[1]
    nB = L + (1 if 32 + t.bit_length() + Q.bit_length() >= 61 * L + 61 else 0)
    pr, v = [], ((1 << 61) - 1) // (2 * N) * (2 * N) + 1     # descending candidates = 1 mod 2N below 2^61
    while len(pr) < nB + 2:
        if is_prime(v):
            pr.append(v)
        v -= 2 * N
    m_sk, B = pr[0], pr[2:]
    Bsk = B + [m_sk]
    bsk_orc = np.zeros(orc.o.orc_ctx_bsk_size(orc.c), dtype=np.uint64)
    orc.o.orc_ctx_bsk(orc.c, p64(bsk_orc))
    assert [int(x) for x in bsk_orc] == Bsk
    mt = 1 << 32
    Bprod = 1
    for x in B:
        Bprod *= x
    rng = np.random.default_rng(22)
    a = rand_residues(rng, moduli[:L], (2,), N)
    b = rand_residues(rng, moduli[:L], (2,), N)

    def extend(x):   # x: [L][N] residues -> residues in Bsk after the small Montgomery reduction
        out = [[0] * N for _ in Bsk]
        for n in range(N):
            xt = [int(x[i][n]) * mt % q[i] for i in range(L)]
            r = (-fastbconv(xt, q, mt) * pow(Q, -1, mt)) % mt
            if r >= mt // 2:
                r -= mt
            for pi, p in enumerate(Bsk):
                out[pi][n] = (fastbconv(xt, q, p) + Q * r) * pow(mt, -1, p) % p
        return out

    aq = [[[int(v) for v in a[p][i]] for i in range(L)] for p in range(2)]
    bq = [[[int(v) for v in b[p][i]] for i in range(L)] for p in range(2)]
    ab, bb = [extend(a[p]) for p in range(2)], [extend(b[p]) for p in range(2)]

    def tensor(x, y, base):   # 3 polynomials per base prime
        out = []
        for i, p in enumerate(base):
            d0 = negacyclic_full(x[0][i], y[0][i], p)
            d1 = [(u + w) % p for u, w in zip(negacyclic_full(x[0][i], y[1][i], p), negacyclic_full(x[1][i], y[0][i], p))]
            d2 = negacyclic_full(x[1][i], y[1][i], p)
            out.append((d0, d1, d2))
        return out

    dq, db = tensor(aq, bq, q), tensor(ab, bb, Bsk)
    got = orc.bfv_multiply(a.reshape(-1), b.reshape(-1)).reshape(3, L, N)
    for c in range(3):
        for n in range(N):
            yq = [t * dq[i][c][n] % q[i] for i in range(L)]
            z = [(t * db[pi][c][n] - fastbconv(yq, q, p)) * pow(Q, -1, p) % p for pi, p in enumerate(Bsk)]
            zB = z[:nB]
            alpha = (fastbconv(zB, B, m_sk) - z[nB]) * pow(Bprod, -1, m_sk) % m_sk
            if alpha > m_sk // 2:
                alpha -= m_sk   # centred
            for i in range(L):
                want = (fastbconv(zB, B, q[i]) - alpha * Bprod) % q[i]
                assert int(got[c, i, n]) == want, (c, i, n)
