"""N>1 path on CPU: world_size-2 (and 3, ragged shards) gloo jobs run the sharded result grid through the C ABI
(emulation build) with the host-side gather, and the partition helpers are checked directly."""
import os
import subprocess
import sys

import numpy as np
import pytest

from pyb200he.shard import block_partition, grid_shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_partition_covers_everything_once():
    for n in (0, 1, 5, 8, 1000, 10007):
        for world in (1, 2, 3, 8):
            parts = block_partition(n, world)
            assert len(parts) == world and parts[0][0] == 0
            assert sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_grid_shard_is_the_reference_pairing():
    n0, n1 = 5, 3
    cells = []
    for r in range(4):
        first, ai, bi = grid_shard(n0, n1, r, 4)
        cells += [(first + k, int(i), int(j)) for k, (i, j) in enumerate(zip(ai, bi))]
    assert cells == [(i * n1 + j, i, j) for i in range(n0) for j in range(n1)]
    assert grid_shard(2, 2, 3, 8)[1].size == 1 and grid_shard(2, 2, 4, 8)[1].size == 0   # more ranks than cells


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_grid_over_gloo(emu_lib, world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world + (os.getpid() % 200)), os.path.join(ROOT, "tests", "dist_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert f"DIST_OK world={world} cells=6" in p.stdout
