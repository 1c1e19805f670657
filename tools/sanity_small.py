#!/usr/bin/env python
"""Small end-to-end pass over the C ABI for compute-sanitizer runs (memcheck): transforms (persistent and cluster
kernels), multiply + fused relinearize+rescale, rotation, scattered load/store, at N = 8192 and N = 16384.
Usage: compute-sanitizer --tool memcheck python tools/sanity_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
import numpy as np
import pyb200he as hb
from pyb200he.hostfhe import CKKS, Host

for N, depth, batch in ((8192, 2, 300), (16384, 3, 9)):
    host = Host(CKKS, N, depth, 45, 45)
    ctx = hb.Context(CKKS, N, host.moduli, host.psi, 0)
    ctx.set_relin_key(host.relin_key())
    e = host.galois_elts()[0]
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
    parts = [np.ascontiguousarray(x[i]) for i in range(min(batch, 12))]
    A.upload_scattered(parts)
    outs = A.download_scattered(0, len(parts))
    assert all(np.array_equal(o, p) for o, p in zip(outs, parts))
    ctx.sync()
    print("ok", N, R.L, T.count, flush=True)
    ctx.close()
