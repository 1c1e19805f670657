#!/usr/bin/env python
"""k_tensor_mac at the C4 chain (N = 16384, {60, 45 x 5, 60}): a rows x inner x cols product of ciphertext matrices on
synthetic residues, for ncu captures and quick timings.    python tools/tensor_mac_probe.py [rows inner cols]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
import numpy as np

import pyb200he as hb
from pyb200he.hostfhe import CKKS, Host

rows, inner, cols = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (24, 24, 24)
N, depth = 16384, 6
host = Host(CKKS, N, depth, 45, 45)
ctx = hb.Context(CKKS, N, host.moduli, host.psi, 0)
rng = np.random.default_rng(1)


def rand(n):
    out = np.empty((n, 2, depth, N), dtype=np.uint64)
    for l in range(depth):
        out[:, :, l, :] = rng.integers(0, int(host.moduli[l]), size=(n, 2, N), dtype=np.uint64)
    return out


A, B = ctx.batch(rand(rows * inner), L=depth), ctx.batch(rand(cols * inner), L=depth)
out = hb.Batch(ctx)
for _ in range(2):
    ctx.matmul_accumulate(A, B, rows, inner, cols, out=out)
ctx.profile_begin()
for _ in range(3):
    ctx.matmul_accumulate(A, B, rows, inner, cols, out=out)
prof, work = ctx.profile_end(), ctx.profile_work()
ms, n = prof["k_tensor_mac"]
print(f"k_tensor_mac {rows}x{inner}x{cols}: {ms / n:.3f} ms per launch, {rows * cols * inner / (ms / n / 1e3) / 1e6:.2f} M ciphertext products/s")
