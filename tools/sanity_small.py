#!/usr/bin/env python
"""Small end-to-end pass over the C ABI for compute-sanitizer runs (memcheck / racecheck / synccheck): transforms
(persistent kernel, limbs split over 2 and 4 CTAs), multiply + fused relinearize+rescale, rotation, accumulate (rotate-and-add
with the Galois permutation fused into the key switch), the CipherBatchAxis product, scattered load/store, at N = 8192, 16384
and 32768.    compute-sanitizer --tool memcheck python tools/sanity_small.py [scale]   (scale divides the batch sizes;
where compute-sanitizer is not available -- it is closed on the round-2 GPU pool -- tests/emu/run_asan.sh is the bounds check)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
import numpy as np
import pyb200he as hb
from pyb200he.hostfhe import CKKS, Host

SCALE = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for N, depth, batch in ((8192, 2, max(300 // SCALE, 4)), (16384, 3, 9), (32768, 3, 3)):
    host = Host(CKKS, N, depth, 45, 45)
    ctx = hb.Context(CKKS, N, host.moduli, host.psi, 0)
    ctx.set_relin_key(host.relin_key())
    for e in host.galois_elts()[:3]:
        ctx.set_galois_key(e, host.galois_key(e))
    L = host.Ltop
    rng = np.random.default_rng(5)
    x = np.empty((batch, 2, L, N), dtype=np.uint64)
    for l in range(L):
        x[:, :, l, :] = rng.integers(0, int(host.moduli[l]), size=(batch, 2, N), dtype=np.uint64)
    A = ctx.batch(x, scale=host.scale)
    C = ctx.batch(x, scale=host.scale, ntt_form=False)
    F = ctx.ntt_forward(C)
    I = ctx.ntt_inverse(F)
    assert np.array_equal(I.download(), x)
    R = ctx.multiply(A, A)
    ctx.relinearize_rescale(R, out=R)
    T = ctx.rotate(A, 1)
    S = ctx.batch(x[:3], scale=host.scale)
    ctx.accumulate(S, 4)   # rotations by 1 and 2, each added to the running sum
    M = hb.Batch(ctx)
    six = np.concatenate([x[:3], x[:3]])   # 3 x 2 times 2 x 3
    ctx.matmul_accumulate(ctx.batch(six, scale=host.scale), ctx.batch(six[::-1], scale=host.scale), 3, 2, 3, out=M)
    parts = [np.ascontiguousarray(x[i]) for i in range(min(batch, 12))]
    A.upload_scattered(parts)
    outs = A.download_scattered(0, len(parts))
    assert all(np.array_equal(o, p) for o, p in zip(outs, parts))
    ctx.sync()
    print("ok", N, R.L, T.count, flush=True)
    ctx.close()
