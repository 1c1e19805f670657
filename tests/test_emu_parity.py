"""CPU-side check of the CUDA sources' index arithmetic: the kernels compiled as host C++ (tests/emu,
fibers for CUDA threads) against the oracle at small N.  The GPU run of the same cases at the
BASELINE.json sizes is tests/test_gpu_parity.py."""
import pytest

import parity
from helpers import BFV, CKKS

CKKS_CHAINS = {
    "N1024-3": (1024, [60, 40, 60]),
    "N2048-4": (2048, [60, 45, 45, 60]),
    # FP64-domain boundary: 46-bit primes run on the FP64 pipe, 47-bit ones on the integer pipe (modarith.cuh)
    "N1024-dp-edge": (1024, [47, 46, 30, 46, 60]),
}


@pytest.fixture(scope="module", params=list(CKKS_CHAINS))
def ckks(request, emu_lib):
    N, bits = CKKS_CHAINS[request.param]
    env = parity.Env(emu_lib, CKKS, N, bits)
    yield env
    env.close()


@pytest.fixture(scope="module")
def bfv(emu_lib):
    env = parity.Env(emu_lib, BFV, 1024, [60, 40, 60], columns=True,
                     galois_steps=tuple(1 << k for k in range(9)) + (-1, -4))
    yield env
    env.close()


def test_ntt(ckks):
    parity.case_ntt(ckks)


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
    parity.case_matmul_accumulate(ckks)


def test_relinearize(ckks):
    for L in range(ckks.Ltop, 0, -1):
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
    parity.case_dot(ckks)


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
    parity.case_bfv_multiply(bfv)


def test_bfv_relinearize(bfv):
    parity.case_relinearize(bfv)


def test_bfv_rotate(bfv):
    parity.case_rotate(bfv)


def test_bfv_modswitch(bfv):
    parity.case_rescale(bfv)


def test_bfv_dot(bfv):
    parity.case_dot(bfv)
    parity.case_bfv_accumulate_columns(bfv)


def test_split_limb(emu_lib):
    """N = 16384: a limb is split over two CTAs (c = 1); N = 32768: four (c = 2)"""
    for N in (16384, 32768):
        env = parity.Env(emu_lib, CKKS, N, [60, 40, 60], galois_steps=(1,))
        parity.case_ntt(env, n=1)
        parity.case_relinearize(env, n=1)
        parity.case_rescale(env, n=1, sizes=(2,))
        parity.case_relin_rescale_fused(env, n=1)
        parity.case_rotate(env, n=1, steps=(1,))      # Galois gather across chunks (key switch fused permutation)
        parity.case_dot(env, n0=1, n1=1, count=2)     # rotate-and-add: contiguous addend + gathered addend
        env.close()
    env = parity.Env(emu_lib, CKKS, 16384, [46, 47, 46, 60], galois_steps=(1,))
    parity.case_extremes(env)
    env.close()


def test_decrypt_level(emu_lib):
    parity.case_decrypt_level(emu_lib, 2048, n=20)
