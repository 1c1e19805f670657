#!/usr/bin/env python
"""End-to-end pipeline probe for bench.py's e2e leg: streams x chunk sweep, and the copy-only floor."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
import numpy as np
import torch
import pyb200he as hb
from pyb200he.hostfhe import CKKS, Host

N, BATCH = 8192, 1000
host = Host(CKKS, N, 2, 45, 45, seed=1234)
L = host.Ltop
words_in, words_out = 2 * L * N, 2 * (L - 1) * N
rng = np.random.default_rng(0)
a_pin = torch.from_numpy(rng.integers(0, 1 << 40, size=BATCH * words_in, dtype=np.int64)).pin_memory()
b_pin = torch.from_numpy(rng.integers(0, 1 << 40, size=BATCH * words_in, dtype=np.int64)).pin_memory()
out_pin = torch.empty(BATCH * words_out, dtype=torch.int64).pin_memory()


def run(n_streams, chunk, compute=True, steps=5):
    e2e = []
    for i in range(n_streams):
        cx = hb.Context(CKKS, N, host.moduli, host.psi, 0)
        st = torch.cuda.Stream()
        cx.set_stream(st.cuda_stream)
        cx.set_relin_key(host.relin_key())
        bufs = [hb.Batch(cx), hb.Batch(cx), hb.Batch(cx)]
        bufs[0].resize(chunk, 2, L, True, host.scale)
        bufs[1].resize(chunk, 2, L, True, host.scale)
        bufs[2].resize(chunk, 2, L - 1, True, host.scale)
        e2e.append((cx, st, bufs))

    def step():
        t0 = time.perf_counter()
        for k, first in enumerate(range(0, BATCH, chunk)):
            cx, st, (ca, cb, cr) = e2e[k % n_streams]
            n = min(chunk, BATCH - first)
            ca.upload_from(a_pin.data_ptr() + first * words_in * 8, 0, n)
            cb.upload_from(b_pin.data_ptr() + first * words_in * 8, 0, n)
            if compute:
                cx.multiply(ca, cb, n=n, out=cr)
                cx.relinearize(cr, out=cr)
                cx.rescale_to_next(cr, out=cr)
            cr.download_to(out_pin.data_ptr() + first * words_out * 8, 0, n, wait=False)
        t1 = time.perf_counter()
        for cx, st, _ in e2e:
            cx.sync()
        return (t1 - t0) * 1e3

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    enq = 0.0
    for _ in range(steps):
        enq += step()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    for cx, _, _b in e2e:
        cx.close()
    return round(ms, 3), round(enq / steps, 3)


res = {}
for ns, ch in ((3, 125), (2, 125), (4, 125), (3, 50), (4, 50), (3, 250), (2, 500), (4, 25), (6, 25)):
    res[f"s{ns}_c{ch}"] = run(ns, ch)
res["copy_only_s3_c125"] = run(3, 125, compute=False)
print(json.dumps(res))
