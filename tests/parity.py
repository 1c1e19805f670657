"""Parity cases shared by the CPU-side emulation tests (tests/test_emu_parity.py, small N) and the
GPU tests (tests/test_gpu_parity.py, the BASELINE.json sizes): every case drives the C ABI of
include/b200he.h and compares the result bit-for-bit with the CPU oracle (oracle/, checker only).

Inputs are uniform random residues and uniform random key arrays: the evaluator arithmetic is
defined for any residues (SURVEY.md §8c "bit-exactness definition"), so this exercises every limb
value range, not only well-formed encryptions.
"""
import numpy as np

import pyb200he as hb
from helpers import BFV, CKKS, Oracle, oracle, rand_residues, p64


def chain(N, bits):
    o = oracle()
    import ctypes as C
    arr = (C.c_int * len(bits))(*bits)
    out = np.zeros(len(bits), dtype=np.uint64)
    assert o.orc_coeff_modulus_create(N, arr, len(bits), p64(out)) == 0
    return out


class Env:
    """a device context + the oracle context on the same chain, with random keys"""

    def __init__(self, lib, scheme, N, bits, seed=1234, plain_bits=20, galois_steps=(1, 2, 4, 8, -1, -2, -4, -8), columns=False):
        self.scheme, self.N = scheme, N
        self.moduli = chain(N, bits)
        self.K = len(bits)
        self.Ltop = self.K - 1
        self.t = int(oracle().orc_plain_modulus_batching(N, plain_bits)) if scheme == BFV else 0
        self.orc = Oracle(scheme, N, self.moduli, self.t)
        self.ctx = hb.Context(scheme, N, self.moduli, self.orc.psi(), self.t, lib=lib)
        self.rng = np.random.default_rng(seed)
        self.relin = self.rand_key()
        self.ctx.set_relin_key(self.relin)
        self.gkeys = {}
        o = oracle()
        elts = [o.orc_galois_elt_from_step(s, N) for s in galois_steps]
        if columns:
            elts.append(2 * N - 1)
        for e in elts:
            self.gkeys[e] = self.rand_key()
            self.ctx.set_galois_key(e, self.gkeys[e])

    def rand_key(self):
        return rand_residues(self.rng, self.moduli, (self.Ltop, 2), self.N).reshape(-1)

    def rand_ct(self, n, size=2, L=None):
        L = self.Ltop if L is None else L
        return rand_residues(self.rng, self.moduli[:L], (n, size), self.N)

    def batch(self, host, size=2, L=None, scale=1.0):
        L = self.Ltop if L is None else L
        return self.ctx.batch(host, size=size, L=L, scale=scale)

    def close(self):
        self.ctx.close()


def eq(a, b, what=""):
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.nonzero(a != b)[0]
        raise AssertionError(f"{what}: {len(bad)} of {a.size} words differ, first at {bad[0]}: {a[bad[0]]} vs {b[bad[0]]}")


# ---------------------------------------------------------------------------------------------
def case_ntt(env, n=3):
    """K1/K2: forward/inverse transform of every limb vs the oracle, and the round trip"""
    L = env.Ltop
    x = env.rand_ct(n, size=2)
    b = env.ctx.batch(x, size=2, L=L, ntt_form=False)
    f = env.ctx.ntt_forward(b)
    got = f.download()
    for i in range(n):
        for p in range(2):
            for l in range(L):
                eq(got[i, p, l], env.orc.ntt(l, x[i, p, l]), f"ntt_fwd ct{i} poly{p} limb{l}")
    back = env.ctx.ntt_inverse(f)
    eq(back.download(), x, "inverse(forward(x))")
    # inverse vs the oracle on arbitrary NTT-form input
    g = env.ctx.batch(x, size=2, L=L, ntt_form=True)
    got = env.ctx.ntt_inverse(g).download()
    for l in range(L):
        eq(got[0, 0, l], env.orc.ntt(l, x[0, 0, l], inverse=True), f"ntt_inv limb{l}")
    # in-place handles
    env.ctx.ntt_forward(b, out=b)
    eq(b.download(), f.download(), "in-place forward")


def case_scattered(env, n=7):
    """load()/store() of separately allocated host ciphertexts through the pinned staging buffers
    (b200he_batch_{upload,download}_scattered): round trip, sub-ranges, and agreement with the contiguous calls"""
    x = env.rand_ct(n, size=2)
    parts = [np.ascontiguousarray(x[i]) for i in range(n)]
    b = env.ctx.batch(np.zeros_like(x), size=2, L=env.Ltop)
    b.upload_scattered(parts)
    eq(b.download(), x, "scattered upload vs contiguous download")
    outs = b.download_scattered()
    for i in range(n):
        eq(outs[i], x[i], f"scattered download ct{i}")
    b.upload_scattered(parts[:2], first=n - 2)   # sub-range
    eq(b.download(n - 2, 2), x[:2], "scattered upload into a sub-range")
    outs = b.download_scattered(1, 3)
    for i in range(3):
        eq(outs[i], x[1 + i], f"scattered download sub-range ct{i}")
    b.upload_scattered([])
    assert b.download_scattered(0, 0) == []


def case_copy_between_contexts(env, n=3):
    """b200he_batch_copy_from: the cross-GPU exchange step of logistic regression (partial collapse sums meet on GPU 0),
    here between two contexts of the same device"""
    other = hb.Context(env.scheme, env.N, env.moduli, env.orc.psi(), env.t, lib=env.ctx.lib)
    x = env.rand_ct(n)
    src = other.batch(x, size=2, L=env.Ltop, scale=3.0)
    dst = hb.Batch(env.ctx).copy_from(src)
    assert (dst.count, dst.size, dst.L, dst.scale) == (n, 2, env.Ltop, 3.0)
    eq(dst.download(), x, "copy between contexts")
    eq(env.ctx.add(dst, dst).download()[1], env.orc.add(env.Ltop, 2, x[1].reshape(-1), x[1].reshape(-1)), "the copy is usable on the destination stream")
    other.close()


def case_batch_outlives_context(env):
    """a batch whose context was destroyed first: every call on it fails cleanly, destroying it is harmless"""
    import pytest
    other = hb.Context(env.scheme, env.N, env.moduli, env.orc.psi(), env.t, lib=env.ctx.lib)
    x = env.rand_ct(2)
    b = other.batch(x, size=2, L=env.Ltop)
    other.close()
    assert b.count == 0
    for call in (lambda: other._ck(b.lib.b200he_batch_resize(b.h, 1, 2, env.Ltop, 1, 1.0)),
                 lambda: other._ck(b.lib.b200he_batch_download(b.h, 0, 0, x.ctypes.data)),
                 lambda: hb.Batch(env.ctx).copy_from(b),
                 lambda: env.ctx.add(b, b)):
        with pytest.raises(hb.B200HEError):
            call()
    del b   # b200he_batch_destroy on the orphan


def case_extremes(env):
    """Worst-case magnitudes for the lazy-reduction schedules of both arithmetic domains (integer pipe, and the FP64
    domain of primes below 2^46): every residue q-1, every residue 0, alternating q-1 / 0 and q-1 / 1 patterns, with
    an all-(q-1) relinearization key, through the transforms, relinearize and the fused relinearize+rescale"""
    L, N = env.Ltop, env.N
    q = env.moduli[:L].astype(np.uint64)
    pats = []
    for kind in range(4):
        v = np.zeros((3, L, N), dtype=np.uint64)
        for l in range(L):
            if kind == 0: v[:, l, :] = q[l] - np.uint64(1)
            elif kind == 2: v[:, l, ::2] = q[l] - np.uint64(1)
            elif kind == 3:
                v[:, l, ::2] = q[l] - np.uint64(1)
                v[:, l, 1::2] = 1
        pats.append(v)
    x = np.stack(pats)   # [4][3][L][N]
    b = env.ctx.batch(np.ascontiguousarray(x[:, :2]), size=2, L=L, ntt_form=False)
    f = env.ctx.ntt_forward(b).download()
    g = env.ctx.ntt_inverse(env.ctx.batch(np.ascontiguousarray(x[:, :2]), size=2, L=L, ntt_form=True)).download()
    for i in range(4):
        for l in range(L):
            eq(f[i, 0, l], env.orc.ntt(l, x[i, 0, l]), f"extremes fwd pattern{i} limb{l}")
            eq(g[i, 1, l], env.orc.ntt(l, x[i, 1, l], inverse=True), f"extremes inv pattern{i} limb{l}")
    if env.scheme == CKKS:   # dyadic products of the extreme patterns, every pattern against every pattern
        X2 = env.batch(np.ascontiguousarray(x[:, :2]), size=2, L=L)
        ai, bi = np.repeat(np.arange(4), 4), np.tile(np.arange(4), 4)
        got = env.ctx.multiply(X2, X2, ai, bi).download()
        for r in range(16):
            eq(got[r], env.orc.ckks_multiply(L, x[ai[r], :2].reshape(-1), x[bi[r], :2].reshape(-1)), f"extremes multiply {ai[r]}x{bi[r]}")
    key = np.empty((L, 2, env.K, N), dtype=np.uint64)
    for k in range(env.K):
        key[:, :, k, :] = env.moduli[k] - np.uint64(1)
    key = key.reshape(-1)
    env.ctx.set_relin_key(key)
    try:
        X = env.batch(x, size=3, L=L, scale=2.0 ** 80)
        got = env.ctx.relinearize(X).download()
        for i in range(4):
            eq(got[i], env.orc.relinearize(L, x[i].reshape(-1), key), f"extremes relinearize pattern{i}")
        if env.scheme == CKKS and L >= 2:
            got = env.ctx.relinearize_rescale(X).download()
            for i in range(4):
                eq(got[i], env.orc.rescale(L, 2, env.orc.relinearize(L, x[i].reshape(-1), key)), f"extremes fused pattern{i}")
    finally:
        env.ctx.set_relin_key(env.relin)


def case_elementwise(env, n0=3, n1=2):
    """K3/K4: add, sub, CKKS multiply over the reference's b0 x b1 result grid (index maps)"""
    L = env.Ltop
    a, b = env.rand_ct(n0), env.rand_ct(n1)
    A, B = env.batch(a), env.batch(b)
    ai = np.repeat(np.arange(n0), n1)
    bi = np.tile(np.arange(n1), n0)
    got = env.ctx.add(A, B, ai, bi).download()
    for r in range(n0 * n1):
        eq(got[r], env.orc.add(L, 2, a[ai[r]].reshape(-1), b[bi[r]].reshape(-1)), f"add {r}")
    got = env.ctx.sub(A, B, ai, bi).download()
    for r in range(n0 * n1):
        want = np.empty(2 * L * env.N, dtype=np.uint64)
        env.orc.o.orc_sub(env.orc.c, L, 2, p64(a[ai[r]].reshape(-1).copy()), p64(b[bi[r]].reshape(-1).copy()), p64(want))
        eq(got[r], want, f"sub {r}")
    if env.scheme == CKKS:
        out = env.ctx.multiply(A, B, ai, bi)
        assert out.size == 3 and out.count == n0 * n1
        got = out.download()
        for r in range(n0 * n1):
            eq(got[r], env.orc.ckks_multiply(L, a[ai[r]].reshape(-1), b[bi[r]].reshape(-1)), f"multiply {r}")
    # identity maps, in-place output
    A2 = env.batch(a)
    env.ctx.add(A2, A2, out=A2)
    eq(A2.download()[0], env.orc.add(L, 2, a[0].reshape(-1), a[0].reshape(-1)), "add in place")
    # size 3 + size 3 (CipherBatchAxis accumulates unrelinearized products)
    c3, d3 = env.rand_ct(2, size=3), env.rand_ct(2, size=3)
    got = env.ctx.add(env.batch(c3, size=3), env.batch(d3, size=3)).download()
    eq(got[1], env.orc.add(L, 3, c3[1].reshape(-1), d3[1].reshape(-1)), "add size 3")


def case_matmul_accumulate(env, rows=3, inner=5, cols=3, L=None):
    """MatMult CipherBatchAxis inner loop (R/src/benchmarks/ckks/seal_ckks_matmult_cipherbatchaxis_benchmark.cpp:385-422):
    out[i][j] = sum_k a[i][k] (x) b[k][j] in one pass == the reference's multiply / add_inplace sequence (b is passed by
    columns: bt[j][k]); odd shapes exercise the edge tiles"""
    L = env.Ltop if L is None else L
    a, b = env.rand_ct(rows * inner, L=L), env.rand_ct(inner * cols, L=L)
    A, B = env.batch(a, L=L, scale=2.0 ** 30), env.batch(b, L=L, scale=2.0 ** 30)
    out = env.ctx.matmul_accumulate(A, B, rows, inner, cols)
    assert out.size == 3 and out.count == rows * cols and out.L == L and out.scale == 2.0 ** 60
    got = out.download()
    for i in range(rows):
        for j in range(cols):
            want = None
            for k in range(inner):
                t = env.orc.ckks_multiply(L, a[i * inner + k].reshape(-1), b[j * inner + k].reshape(-1))
                want = t if want is None else env.orc.add(L, 3, want, t)
            eq(got[i * cols + j], want, f"matmul_accumulate cell ({i},{j})")
    # worst-case magnitudes: every residue q - 1, more terms than one accumulator run holds for 60-bit primes
    n = 140
    ones = np.empty((1, 2, L, env.N), dtype=np.uint64)
    for l in range(L):
        ones[:, :, l, :] = env.moduli[l] - np.uint64(1)
    x = np.ascontiguousarray(np.repeat(ones, n, axis=0))
    got = env.ctx.matmul_accumulate(env.batch(x, L=L), env.batch(x, L=L), 1, n, 1).download()
    t = env.orc.ckks_multiply(L, ones.reshape(-1), ones.reshape(-1)).reshape(3, L, env.N)
    want = np.empty_like(t)
    for l in range(L):
        want[:, l, :] = (t[:, l, :].astype(object) * n % int(env.moduli[l])).astype(np.uint64)
    eq(got[0], want, "matmul_accumulate, all residues q - 1, 140 terms")


def case_relinearize(env, n=2, L=None):
    """K6/K7: relinearize (size 3 -> 2) at level L"""
    L = env.Ltop if L is None else L
    x = env.rand_ct(n, size=3, L=L)
    X = env.batch(x, size=3, L=L)
    out = env.ctx.relinearize(X)
    assert out.size == 2 and out.L == L
    got = out.download()
    for i in range(n):
        eq(got[i], env.orc.relinearize(L, x[i].reshape(-1), env.relin), f"relinearize L={L} ct{i}")
    env.ctx.relinearize(X, out=X)   # in place
    eq(X.download(), got, "relinearize in place")
    # size-2 input: SEAL leaves it untouched
    y = env.rand_ct(1, size=2, L=L)
    eq(env.ctx.relinearize(env.batch(y, L=L)).download(), y, "relinearize size 2 is a no-op")


def case_rotate(env, n=2, L=None, steps=(1, 4, -1, 3, 7, -3)):
    """K8: rotate_vector / rotate_rows, including SEAL's NAF decomposition for steps without a key"""
    L = env.Ltop if L is None else L
    x = env.rand_ct(n, L=L)
    X = env.batch(x, L=L)
    for s in steps:
        got = env.ctx.rotate(X, s).download()
        for i in range(n):
            eq(got[i], env.orc.rotate(L, x[i].reshape(-1), s, env.gkeys), f"rotate step {s} L={L} ct{i}")
    eq(env.ctx.rotate(X, 0).download(), x, "rotate by 0")
    Y = env.batch(x, L=L)
    env.ctx.rotate(Y, 1, out=Y)
    eq(Y.download()[0], env.orc.rotate(L, x[0].reshape(-1), 1, env.gkeys), "rotate in place")


def case_rescale(env, n=2, sizes=(2, 3)):
    """K9/K10: rescale_to_next (CKKS) / mod_switch_to_next (BFV), and CKKS limb dropping"""
    for L in range(env.Ltop, 1, -1):
        for size in sizes:
            x = env.rand_ct(n, size=size, L=L)
            out = env.ctx.rescale_to_next(env.batch(x, size=size, L=L, scale=2.0 ** 80))
            assert out.L == L - 1 and out.size == size
            if env.scheme == CKKS:
                assert out.scale == 2.0 ** 80 / float(env.moduli[L - 1])
            got = out.download()
            for i in range(n):
                eq(got[i], env.orc.rescale(L, size, x[i].reshape(-1)), f"rescale L={L} size={size} ct{i}")
    if env.scheme == CKKS and env.Ltop >= 2:
        L = env.Ltop
        x = env.rand_ct(n, L=L)
        got = env.ctx.mod_drop(env.batch(x, L=L), L - 1).download()
        eq(got[0], env.orc.mod_drop(L, 2, x[0].reshape(-1)), "mod_drop")
        eq(got, x[:, :, : L - 1, :], "mod_drop keeps the leading limbs")
        if L >= 3:
            eq(env.ctx.mod_drop(env.batch(x, L=L), L - 2).download(), x[:, :, : L - 2, :], "mod_drop two levels")


def case_plain(env, n=3):
    """K11: multiply_plain / add_plain (CKKS, NTT form) with a plaintext index map"""
    L = env.Ltop
    x = env.rand_ct(n, L=L)
    pl = env.rand_ct(2, size=1, L=L)
    X, P = env.batch(x, L=L, scale=2.0 ** 40), env.ctx.batch(pl, size=1, L=L, scale=2.0 ** 40)
    pi = np.array([1, 0, 1][:n], dtype=np.uint32)
    out = env.ctx.multiply_plain(X, P, pi)
    assert out.scale == 2.0 ** 80
    got = out.download()
    for i in range(n):
        eq(got[i], env.orc.multiply_plain(L, 2, x[i].reshape(-1), pl[pi[i]].reshape(-1)), f"multiply_plain {i}")
    got = env.ctx.add_plain(X, P, pi).download()
    for i in range(n):
        eq(got[i], env.orc.add_plain(L, 2, x[i].reshape(-1), pl[pi[i]].reshape(-1)), f"add_plain {i}")


def case_dot(env, n0=2, n1=2, count=5, L=None):
    """DotProduct operate(): multiply -> relinearize -> accumulate(count) over the b0 x b1 grid
    (R/src/benchmarks/ckks/seal_ckks_dot_product_benchmark.cpp:315-331)"""
    L = env.Ltop if L is None else L
    a, b = env.rand_ct(n0, L=L), env.rand_ct(n1, L=L)
    ai = np.repeat(np.arange(n0), n1)
    bi = np.tile(np.arange(n1), n0)
    r = env.ctx.multiply(env.batch(a, L=L), env.batch(b, L=L), ai, bi)
    env.ctx.relinearize(r, out=r)
    env.ctx.accumulate(r, count)
    got = r.download()
    ga = np.ascontiguousarray(a[ai]).reshape(-1)
    gb = np.ascontiguousarray(b[bi]).reshape(-1)
    want = env.orc.batch_dot(L, n0 * n1, ga, gb, count, env.relin, env.gkeys)
    eq(got, want, f"dot product count={count}")


def case_mul_relin_rescale(env, n=3):
    """BASELINE.json configs[1]: CKKS multiply + relinearize + rescale"""
    L = env.Ltop
    a, b = env.rand_ct(n), env.rand_ct(n)
    r = env.ctx.multiply(env.batch(a, scale=2.0 ** 40), env.batch(b, scale=2.0 ** 40))
    env.ctx.relinearize(r, out=r)
    env.ctx.rescale_to_next(r, out=r)
    assert r.L == L - 1 and r.size == 2
    want = env.orc.mul_relin_rescale(L, n, a.reshape(-1), b.reshape(-1), env.relin)
    eq(r.download(), want, "multiply+relinearize+rescale")


def case_relin_rescale_fused(env, n=3, L=None):
    """b200he_relinearize_rescale == relinearize_inplace then rescale_to_next_inplace, bit for bit, and == the two
    separate C-ABI calls"""
    L = env.Ltop if L is None else L
    ct3 = env.rand_ct(n, size=3, L=L)
    b = env.batch(ct3, size=3, L=L, scale=2.0 ** 80)
    fused = env.ctx.relinearize_rescale(b)
    assert fused.L == L - 1 and fused.size == 2 and fused.ntt_form
    got = fused.download()
    for i in range(n):
        want = env.orc.rescale(L, 2, env.orc.relinearize(L, ct3[i].reshape(-1), env.relin))
        eq(got[i], want, f"fused relinearize+rescale L={L} ct {i}")
    two = env.ctx.rescale_to_next(env.ctx.relinearize(b))
    eq(two.download(), got, "fused vs two calls")
    assert abs(fused.scale - two.scale) <= 1e-9 * two.scale
    env.ctx.relinearize_rescale(b, out=b)   # in place
    eq(b.download(), got, "fused in place")


def case_bfv_multiply(env, n0=2, n1=2):
    """K5: BFV (BEHZ) multiply over the result grid"""
    a, b = env.rand_ct(n0), env.rand_ct(n1)
    ai = np.repeat(np.arange(n0), n1)
    bi = np.tile(np.arange(n1), n0)
    out = env.ctx.multiply(env.batch(a), env.batch(b), ai, bi)
    assert out.size == 3 and not out.ntt_form
    got = out.download()
    for r in range(n0 * n1):
        eq(got[r], env.orc.bfv_multiply(a[ai[r]].reshape(-1), b[bi[r]].reshape(-1)), f"bfv multiply {r}")


def case_bfv_accumulate_columns(env):
    """accumulateBFV with count > N/2: row rotations plus the column swap (R/src/engine/seal_context.cpp:305-310)"""
    L = env.Ltop
    x = env.rand_ct(1)
    X = env.batch(x)
    count = env.N // 2 + 3
    env.ctx.accumulate(X, count)
    eq(X.download()[0], env.orc.accumulate(L, x[0].reshape(-1), count, env.gkeys), "accumulateBFV with column swap")


def case_errors(env):
    """argument checking mirrors SEAL's exceptions: mismatched levels, missing keys, bad sizes"""
    import pytest
    L = env.Ltop
    x = env.batch(env.rand_ct(1, L=L), L=L)
    with pytest.raises(hb.B200HEError):
        env.ctx.rotate(x, 512 if env.N > 1024 else env.N // 2)   # step out of range / no key and single NAF term
    with pytest.raises(hb.B200HEError):
        env.ctx.apply_galois(x, 5)   # no such key
    x3 = env.batch(env.rand_ct(1, size=3, L=L), size=3, L=L)
    with pytest.raises(hb.B200HEError):
        env.ctx.rotate(x3, 1)        # size must be 2
    with pytest.raises(hb.B200HEError):
        env.ctx.multiply(x3, x3)
    if L >= 2:
        y = env.batch(env.rand_ct(1, L=L - 1), L=L - 1)
        with pytest.raises(hb.B200HEError):
            env.ctx.add(x, y)        # different levels
    with pytest.raises(hb.B200HEError):
        env.ctx.add(x, x, ai=[3], bi=[0])   # index out of range
    out = env.batch(env.rand_ct(1, L=L), L=L)
    with pytest.raises(hb.B200HEError):
        env.ctx.sub(x, x3, out=out)         # size(b) > size(a): rejected before `out` is touched
    assert out.size == 2 and out.count == 1
    # the fused entry reports what the two calls would: last level, and size-2 input = plain rescale
    x1 = env.batch(env.rand_ct(1, size=3, L=1), size=3, L=1)
    with pytest.raises(hb.B200HEError):
        env.ctx.relinearize_rescale(x1)
    if L >= 2:
        x2 = env.rand_ct(2, L=L)
        eq(env.ctx.relinearize_rescale(env.batch(x2, L=L)).download(), env.ctx.rescale_to_next(env.batch(x2, L=L)).download(),
           "relinearize_rescale on size-2 input")
        assert env.ctx.relinearize_rescale(env.ctx.batch(np.zeros(0, dtype=np.uint64), size=3, L=L)).count == 0
    # empty batches are fine
    e = env.ctx.batch(np.zeros(0, dtype=np.uint64), L=L)
    assert env.ctx.add(e, e).count == 0
    assert env.ctx.relinearize(env.ctx.batch(np.zeros(0, dtype=np.uint64), size=3, L=L)).count == 0


def case_decrypt_level(lib, N, n=100):
    """value-level check with real keys (host FHE stand-in): the dot-product pipeline decrypts to the
    cleartext dot product, and multiply+relinearize+rescale to the slot-wise product"""
    from helpers import Host
    h = Host(CKKS, N, 2, 45, 45)
    ctx = hb.Context(CKKS, N, h.moduli, h.psi, 0, lib=lib)
    ctx.set_relin_key(h.relin_key())
    for e in h.galois_elts():
        ctx.set_galois_key(e, h.galois_key(e))
    rng = np.random.default_rng(7)
    u, v = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    A = ctx.batch(h.enc_vec(u), scale=h.scale)
    B = ctx.batch(h.enc_vec(v), scale=h.scale)
    r = ctx.multiply(A, B)
    ctx.relinearize(r, out=r)
    s = ctx.rescale_to_next(r)
    dec = h.dec_vec(s.download()[0].reshape(-1), 2, 1, scale=s.scale)
    assert np.max(np.abs(dec[:n] - u * v)) < 1e-4
    ctx.accumulate(r, n)
    dec = h.dec_vec(r.download()[0].reshape(-1), 2, 2, scale=r.scale)
    want = float(np.dot(u, v))
    assert abs(dec[0] - want) < 1e-4 * max(1.0, abs(want)), (dec[0], want)
    ctx.close()


def case_rotate_each_and_sum(env, n=7, L=None):
    """collapse building blocks: per-ciphertext rotation steps (sample i rotated by -i) and the batch sum"""
    L = env.Ltop if L is None else L
    x = env.rand_ct(n, L=L)
    X = env.batch(x, L=L)
    steps = [-i for i in range(n)]
    got = env.ctx.rotate_each(X, steps).download()
    for i in range(n):
        want = x[i].reshape(-1) if i == 0 else env.orc.rotate(L, x[i].reshape(-1), -i, env.gkeys)
        eq(got[i], want, f"rotate_each ct{i} by {-i}")
    s = env.ctx.sum(X).download()
    want = x[0].reshape(-1).copy()
    for i in range(1, n):
        want = env.orc.add(L, 2, want, x[i].reshape(-1))
    eq(s[0], want, "batch sum")
