#!/usr/bin/env python
"""NTT limb-ops/s (BASELINE.json metric, second half): batched forward / inverse negacyclic NTT through the
C ABI, timed per kernel class with CUDA events.  Usage: tools/ntt_bench.py [N] [bits,...] [n_ct] [iters]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
import numpy as np
import pyb200he as hb
from pyb200he.hostfhe import CKKS, Host

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
bits = int(sys.argv[3]) if len(sys.argv) > 3 else 45
n_ct = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
host = Host(CKKS, N, depth, bits, bits)
ctx = hb.Context(CKKS, N, host.moduli, host.psi, 0)
L = host.Ltop
rng = np.random.default_rng(1)
x = np.empty((n_ct, 2, L, N), dtype=np.uint64)
for l in range(L):
    x[:, :, l, :] = rng.integers(0, int(host.moduli[l]), size=(n_ct, 2, N), dtype=np.uint64)
A = ctx.batch(x, ntt_form=False)
F, B = hb.Batch(ctx), hb.Batch(ctx)
for _ in range(3):
    ctx.ntt_forward(A, out=F)
    ctx.ntt_inverse(F, out=B)
ctx.sync()
ctx.profile_begin()
for _ in range(iters):
    ctx.ntt_forward(A, out=F)
    ctx.ntt_inverse(F, out=B)
prof = ctx.profile_end()
assert np.array_equal(B.download(0, 2), x[:2])
limbs = n_ct * 2 * L
out = {"N": N, "moduli_bits": [int(m).bit_length() for m in host.moduli[:L]], "limbs_per_launch": limbs}
for k, (ms, n) in prof.items():
    per = ms / n
    out[k] = {"ms_per_launch": per, "launches": n}
fwd = prof["k_ntt_fwd"][0] / prof["k_ntt_fwd"][1]
inv = sum(prof[k][0] for k in prof if k.startswith("k_ntt_inv")) / prof["k_ntt_inv"][1]
out["fwd_limb_ntts_per_s"] = limbs / (fwd / 1e3)
out["inv_limb_ntts_per_s"] = limbs / (inv / 1e3)
out["fwd_GBps_algorithmic"] = limbs * 2 * N * 8 / (fwd / 1e3) / 1e9
out["inv_GBps_algorithmic"] = limbs * 2 * N * 8 / (inv / 1e3) / 1e9
out["fwd_butterflies_per_s"] = out["fwd_limb_ntts_per_s"] * (N // 2) * (N.bit_length() - 1)
print(json.dumps(out))
