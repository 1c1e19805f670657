#!/usr/bin/env python
"""load()/store() probe: scattered (separately allocated, pageable) host ciphertexts <-> a device batch through the
library's pinned staging (b200he_batch_{upload,download}_scattered) against per-ciphertext pageable copies.
Usage: tools/load_store_probe.py [N] [count]; env B200HE_HOST_THREADS, B200HE_STAGE_MB select the staging shape."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
import numpy as np
import pyb200he as hb
from pyb200he.hostfhe import CKKS, Host

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
host = Host(CKKS, N, 2, 40, 40)
ctx = hb.Context(CKKS, N, host.moduli, host.psi, 0)
L = host.Ltop
rng = np.random.default_rng(1)
parts = [rng.integers(0, 1 << 40, size=(2, L, N), dtype=np.uint64) for _ in range(count)]
b = ctx.batch(np.zeros((count, 2, L, N), dtype=np.uint64), size=2, L=L)
ctx.sync()
mb = count * 2 * L * N * 8 / 1e6
out = {"N": N, "count": count, "MB": mb, "threads": os.environ.get("B200HE_HOST_THREADS"), "stage_mb": os.environ.get("B200HE_STAGE_MB")}
for name, fn in (("upload_scattered", lambda: (b.upload_scattered(parts), ctx.sync())),
                 ("upload_per_ct_pageable", lambda: ([b.upload_from(p.ctypes.data, i, 1) for i, p in enumerate(parts)], ctx.sync())),
                 ("download_scattered", lambda: b.download_scattered()),
                 ("download_per_ct_pageable", lambda: [b.download(i, 1) for i in range(count)])):
    best, cold = 1e9, None
    for _ in range(3):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        cold = dt if cold is None else cold
        best = min(best, dt)
    out[name] = {"ms": round(best * 1e3, 2), "first_call_ms": round(cold * 1e3, 2), "GBps": round(mb / 1e3 / best, 2)}
print(json.dumps(out))
