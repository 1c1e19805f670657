"""Multi-GPU plumbing of the hot path: independent result ciphertexts are sharded over ranks (one process per
GPU), keys and tables are replicated, and there is NO collective on the data path -- results are
gathered on the host side.  torch.distributed is used for the barrier, the max-over-ranks time and the host-side
gather only; the same code runs over NCCL on the GPU box (bench.py) and over gloo in the CPU tests
(tests/test_distributed.py).  Mirrors SEALContextWrapper::partition of the C++ backend
(reference-seal-backend_b200/backend/src/engine/b200_context.cpp)."""
import numpy as np


def block_partition(n, world):
    """split [0, n) into `world` contiguous blocks whose sizes differ by at most one: [(first, count), ...]"""
    base, extra = divmod(int(n), int(world))
    out, first = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def grid_shard(n0, n1, rank, world):
    """this rank's block of the reference's b0 x b1 result grid (row-major, result i*n1 + j pairs a[i] with b[j],
    R/src/benchmarks/ckks/seal_ckks_element_wise_benchmark.cpp:325-345): returns (first, ai, bi) index maps"""
    first, cnt = block_partition(n0 * n1, world)[rank]
    flat = np.arange(first, first + cnt, dtype=np.int64)
    return first, (flat // n1).astype(np.uint32), (flat % n1).astype(np.uint32)


class Ranks:
    """rank bookkeeping + the three non-data-path collectives"""

    def __init__(self, dist=None, device=None):
        self.dist = dist
        self.device = device
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, value):
        if self.dist is None:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device or "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, value):
        if self.dist is None:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device or "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_host(self, array):
        """host-side gather of per-rank result blocks (store()): rank 0 gets the concatenation, others None"""
        if self.dist is None:
            return array
        parts = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(array, parts, dst=0)
        return np.concatenate(parts) if self.rank == 0 else None
