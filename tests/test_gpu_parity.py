"""GPU parity tests: the CUDA path through the C ABI (libb200he.so) against the CPU oracle, bit-exact,
at the parameter sets of BASELINE.json's configs, plus size-independent properties at full batch."""
import numpy as np
import pytest

import parity
from helpers import BFV, CKKS

pytestmark = pytest.mark.gpu

# BASELINE.json configs: C1 BFV N=8192 {60,40,60}; C2 CKKS N=8192 {60,45,60}; C3 CKKS N=16384 {60,40,60};
# C4 CKKS N=16384 {60,45x5,60}; C5 CKKS N=32768 {60,45x5,60}
CKKS_CHAINS = {
    "C2-N8192-K3": (8192, [60, 45, 60]),
    "C3-N16384-K3": (16384, [60, 40, 60]),
    "C4-N16384-K7": (16384, [60, 45, 45, 45, 45, 45, 60]),
    "C5-N32768-K7": (32768, [60, 45, 45, 45, 45, 45, 60]),
    "small-N4096-K4": (4096, [60, 40, 40, 60]),
    # FP64-domain boundary: 46-bit primes run on the FP64 pipe, 47-bit ones on the integer pipe (modarith.cuh)
    "dp-edge-N16384-K5": (16384, [47, 46, 30, 46, 60]),
}


@pytest.fixture(scope="module", params=list(CKKS_CHAINS))
def ckks(request, gpu_lib):
    N, bits = CKKS_CHAINS[request.param]
    env = parity.Env(gpu_lib, CKKS, N, bits)
    yield env
    env.close()


@pytest.fixture(scope="module")
def bfv(gpu_lib):
    N = 8192
    env = parity.Env(gpu_lib, BFV, N, [60, 40, 60], columns=True,
                     galois_steps=tuple(1 << k for k in range(12)) + (-1, -4))
    yield env
    env.close()


def test_loaded_library_is_cuda(gpu_lib):
    assert b"sm_100a" in gpu_lib.b200he_version()


def test_ntt(ckks):
    parity.case_ntt(ckks, n=2)


def test_extremes(ckks):
    parity.case_extremes(ckks)


def test_scattered_load_store(ckks, monkeypatch):
    parity.case_scattered(ckks)


def test_elementwise(ckks):
    parity.case_elementwise(ckks)


def test_copy_between_contexts(ckks):
    parity.case_copy_between_contexts(ckks)


def test_batch_outlives_context(ckks):
    parity.case_batch_outlives_context(ckks)


def test_matmul_accumulate(ckks):
    parity.case_matmul_accumulate(ckks, rows=5, inner=9, cols=4)
    if ckks.Ltop > 2:
        parity.case_matmul_accumulate(ckks, rows=2, inner=3, cols=2, L=ckks.Ltop - 1)


def test_relinearize(ckks):
    for L in sorted({ckks.Ltop, max(1, ckks.Ltop - 1), 1}, reverse=True):
        parity.case_relinearize(ckks, L=L)


def test_rotate(ckks):
    parity.case_rotate(ckks)
    if ckks.Ltop > 2:
        parity.case_rotate(ckks, L=ckks.Ltop - 1, steps=(1, 3))


def test_rescale(ckks):
    parity.case_rescale(ckks)


def test_rotate_each_and_sum(ckks):
    parity.case_rotate_each_and_sum(ckks)


def test_plain(ckks):
    parity.case_plain(ckks)


def test_dot(ckks):
    parity.case_dot(ckks, count=9)


def test_mul_relin_rescale(ckks):
    parity.case_mul_relin_rescale(ckks)


def test_relin_rescale_fused(ckks):
    for L in sorted({ckks.Ltop, max(2, ckks.Ltop - 1), 2}, reverse=True):
        parity.case_relin_rescale_fused(ckks, L=L)


def test_errors(ckks):
    parity.case_errors(ckks)


def test_bfv_elementwise(bfv):
    parity.case_elementwise(bfv)


def test_bfv_multiply(bfv):
    parity.case_bfv_multiply(bfv, n0=3, n1=2)


def test_bfv_relinearize(bfv):
    parity.case_relinearize(bfv)


def test_bfv_rotate(bfv):
    parity.case_rotate(bfv)


def test_bfv_modswitch(bfv):
    parity.case_rescale(bfv)


def test_bfv_dot(bfv):
    parity.case_dot(bfv, count=100)
    parity.case_bfv_accumulate_columns(bfv)


def test_chunked_workspace_matches_unchunked(gpu_lib):
    """a small workspace forces the key switch to run in several chunks: same bits"""
    env = parity.Env(gpu_lib, CKKS, 8192, [60, 45, 60], galois_steps=(1,))
    x = env.rand_ct(37, size=3)
    want = env.ctx.relinearize(env.batch(x, size=3)).download()
    env.ctx.set_workspace(8 << 20)
    got = env.ctx.relinearize(env.batch(x, size=3)).download()
    parity.eq(got, want, "chunked relinearize")
    parity.eq(got[36], env.orc.relinearize(2, x[36].reshape(-1), env.relin), "last ciphertext vs oracle")
    env.close()


def test_full_batch(gpu_lib):
    """BASELINE.json configs[1] at full size (1000 ciphertexts).  The batch tiles 8 distinct ciphertext
    pairs, so every one of the 1000 results is pinned by 8 oracle results; plus exact size-independent
    properties over fully random batches: commutativity of multiply, (a+b)-b = a, inverse(forward) = id."""
    env = parity.Env(gpu_lib, CKKS, 8192, [60, 45, 60], galois_steps=(1,))
    n, L, period = 1000, 2, 8
    a8, b8 = env.rand_ct(period), env.rand_ct(period)
    reps = (n + period - 1) // period
    a = np.ascontiguousarray(np.tile(a8, (reps, 1, 1, 1))[:n])
    b = np.ascontiguousarray(np.tile(b8, (reps, 1, 1, 1))[:n])
    A, B = env.batch(a, scale=2.0 ** 45), env.batch(b, scale=2.0 ** 45)
    r = env.ctx.multiply(A, B)
    env.ctx.relinearize(r, out=r)
    env.ctx.rescale_to_next(r, out=r)
    got = r.download()
    want8 = env.orc.mul_relin_rescale(L, period, a8.reshape(-1), b8.reshape(-1), env.relin).reshape(period, -1)
    for i in range(n):
        parity.eq(got[i], want8[i % period], f"ciphertext {i} of {n}")
    # rotation over the full batch, same construction
    rot = env.ctx.rotate(A, 1).download()
    want8 = [env.orc.rotate(L, a8[i].reshape(-1), 1, env.gkeys) for i in range(period)]
    for i in range(n):
        parity.eq(rot[i], want8[i % period], f"rotated ciphertext {i} of {n}")
    # exact algebraic properties on fully random batches
    x, y = env.rand_ct(n), env.rand_ct(n)
    X, Y = env.batch(x), env.batch(y)
    parity.eq(env.ctx.multiply(X, Y).download(), env.ctx.multiply(Y, X).download(), "multiply commutes")
    P3 = env.ctx.multiply(X, Y)
    parity.eq(env.ctx.relinearize_rescale(P3).download(), env.ctx.rescale_to_next(env.ctx.relinearize(P3)).download(),
              "fused relinearize+rescale == the two calls over 1000 ciphertexts")
    parity.eq(env.ctx.sub(env.ctx.add(X, Y), Y).download(), x, "(x + y) - y = x")
    C = env.ctx.batch(x, ntt_form=False)
    parity.eq(env.ctx.ntt_inverse(env.ctx.ntt_forward(C)).download(), x, "inverse(forward) over 1000 ciphertexts")
    env.close()


def test_decrypt_level(gpu_lib):
    parity.case_decrypt_level(gpu_lib, 8192)
