"""Worker of tests/test_distributed.py: one rank of a world_size-N gloo job that runs its shard of a b0 x b1
multiply+relinearize+rescale result grid through the C ABI (the host-C++ emulation build of the CUDA sources --
test infrastructure, no GPU here), then the host-side gather; rank 0 checks the gathered grid against the oracle."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
import torch.distributed as dist  # noqa: E402

import parity  # noqa: E402
import pyb200he  # noqa: E402
from helpers import CKKS  # noqa: E402
from pyb200he.shard import Ranks, grid_shard  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rk = Ranks(dist)
    lib = pyb200he.declare(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libb200he_emu.so")))
    env = parity.Env(lib, CKKS, 1024, [60, 40, 60], seed=77, galois_steps=())   # same seed: keys replicated on every rank
    n0, n1 = 3, 2
    a, b = env.rand_ct(n0), env.rand_ct(n1)          # same inputs on every rank (load() uploads what a block needs)
    first, ai, bi = grid_shard(n0, n1, rk.rank, rk.world)
    r = env.ctx.multiply(env.batch(a, scale=2.0 ** 40), env.batch(b, scale=2.0 ** 40), ai, bi)
    env.ctx.relinearize(r, out=r)
    env.ctx.rescale_to_next(r, out=r)
    mine = r.download()
    rk.barrier()
    t = rk.max_over_ranks(1.0 + rk.rank)              # the timing reduction of bench.py
    units = rk.sum_over_ranks(len(ai))
    got = rk.gather_host(mine)
    if rk.rank == 0:
        assert t == float(rk.world), t
        assert units == n0 * n1, units
        L = env.Ltop
        assert got.shape[0] == n0 * n1
        for k in range(n0 * n1):
            want = env.orc.mul_relin_rescale(L, 1, a[k // n1].reshape(-1), b[k % n1].reshape(-1), env.relin)
            parity.eq(got[k], want, f"grid cell {k}")
        print(f"DIST_OK world={rk.world} cells={n0 * n1}")
    env.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
