#!/usr/bin/env python
"""Device-resident per-kernel timing of one workload step through the C ABI (no torch: starts in a second).
Usage: tools/step_bench.py [mrr|dot|rot|ntt] [N] [depth] [bits] [batch] [iters]
  mrr = CKKS multiply + relinearize + rescale (bench.py's step), dot = multiply + relinearize + accumulate(100),
  rot = one rotation, ntt = forward + inverse transforms.  Env B200HE_LOGNL selects the CTA-local transform size."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
import numpy as np
import pyb200he as hb
from pyb200he.hostfhe import CKKS, Host

wl = sys.argv[1] if len(sys.argv) > 1 else "mrr"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
bits = int(sys.argv[4]) if len(sys.argv) > 4 else 45
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 1000
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5

host = Host(CKKS, N, depth, bits, bits)
lib = hb.load_library(os.environ["B200HE_LIB"]) if os.environ.get("B200HE_LIB") else None   # experiment builds
ctx = hb.Context(CKKS, N, host.moduli, host.psi, 0, lib=lib)
ctx.set_relin_key(host.relin_key())
L = host.Ltop
rng = np.random.default_rng(1)


def synth(n):
    x = np.empty((n, 2, L, N), dtype=np.uint64)
    for l in range(L):
        x[:, :, l, :] = rng.integers(0, int(host.moduli[l]), size=(n, 2, N), dtype=np.uint64)
    return x


one = synth(min(batch, 64))
reps = (batch + len(one) - 1) // len(one)
A = ctx.batch(np.concatenate([one] * reps)[:batch], scale=host.scale, ntt_form=(wl != "ntt"))
B = ctx.batch(np.concatenate([one[::-1]] * reps)[:batch], scale=host.scale, ntt_form=(wl != "ntt"))
R, T = hb.Batch(ctx), hb.Batch(ctx)
if wl in ("dot", "rot"):
    for e in host.galois_elts()[:16 if wl == "dot" else 2]:
        ctx.set_galois_key(e, host.galois_key(e))


def step():
    if wl == "mrr":
        ctx.multiply(A, B, out=R)
        ctx.relinearize_rescale(R, out=R)
    elif wl == "mrr2":   # the two separate calls
        ctx.multiply(A, B, out=R)
        ctx.relinearize(R, out=R)
        ctx.rescale_to_next(R, out=R)
    elif wl == "dot":
        ctx.multiply(A, B, out=R)
        ctx.relinearize(R, out=R)
        ctx.accumulate(R, 100)
    elif wl == "rot":
        ctx.rotate(A, 1, out=R)
    else:
        ctx.ntt_forward(A, out=R)
        ctx.ntt_inverse(R, out=T)


for _ in range(2):
    step()
ctx.sync()
ctx.profile_begin()
t0 = time.perf_counter()
for _ in range(iters):
    step()
prof = ctx.profile_end()
wall = (time.perf_counter() - t0) * 1e3 / iters
tot = sum(v[0] for v in prof.values()) / iters
out = {"workload": wl, "N": N, "K": depth + 1, "bits": bits, "batch": batch, "lognl": os.environ.get("B200HE_LOGNL"),
       "kernel_ms_per_step": round(tot, 4), "wall_ms_per_step": round(wall, 4), "samples_per_s": round(batch / (tot / 1e3)),
       "kernels": {k: [round(v[0] / iters, 4), v[1] // iters] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
if wl == "ntt":
    limbs = batch * 2 * L
    out["fwd_limb_ntts_per_s"] = round(limbs / (prof["k_ntt_fwd"][0] / iters / 1e3))
    out["fwd_Tbfly_per_s"] = round(out["fwd_limb_ntts_per_s"] * (N // 2) * (N.bit_length() - 1) / 1e12, 4)
    inv = sum(v[0] for k, v in prof.items() if k.startswith("k_ntt_inv")) / iters
    out["inv_limb_ntts_per_s"] = round(limbs / (inv / 1e3))
print(json.dumps(out))
