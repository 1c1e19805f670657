#!/usr/bin/env python
"""Race hunt: every evaluator op repeated on the same inputs must give the same bits every time.
Usage: tools/stress_determinism.py [N] [depth] [bits] [reps] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
import numpy as np
import pyb200he as hb
from pyb200he.hostfhe import CKKS, Host

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 6
bits = int(sys.argv[3]) if len(sys.argv) > 3 else 45
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 40
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 7
host = Host(CKKS, N, depth, bits, bits)
ctx = hb.Context(CKKS, N, host.moduli, host.psi, 0)
ctx.set_relin_key(host.relin_key())
for e in host.galois_elts()[:10]:
    ctx.set_galois_key(e, host.galois_key(e))
L = host.Ltop
rng = np.random.default_rng(3)


def synth(n, size):
    x = np.empty((n, size, L, N), dtype=np.uint64)
    for l in range(L):
        x[:, :, l, :] = rng.integers(0, int(host.moduli[l]), size=(n, size, N), dtype=np.uint64)
    return x


X2 = ctx.batch(synth(batch, 2), size=2, scale=host.scale)
X3 = ctx.batch(synth(batch, 3), size=3, scale=host.scale)
C2 = ctx.batch(synth(batch, 2), size=2, ntt_form=False)
ops = {
    "sum": lambda: ctx.sum(X2),
    "ntt_forward": lambda: ctx.ntt_forward(C2),
    "ntt_inverse": lambda: ctx.ntt_inverse(X2),
    "relinearize": lambda: ctx.relinearize(X3),
    "rescale": lambda: ctx.rescale_to_next(X2),
    "relin_rescale": lambda: ctx.relinearize_rescale(X3),
    "rotate1": lambda: ctx.rotate(X2, 1),
    "rotate_each": lambda: ctx.rotate_each(X2, [-i for i in range(batch)]),
}
bad = 0
for name, f in ops.items():
    ref = f().download()
    nbad = 0
    for _ in range(reps):
        got = f().download()
        if not np.array_equal(got, ref):
            nbad += 1
            d = np.argwhere(got != ref)
            print(f"  {name}: {len(d)} words differ, first at {d[0].tolist()} last at {d[-1].tolist()}")
    print(f"{name}: {nbad} of {reps} runs differ")
    bad += nbad
sys.exit(1 if bad else 0)
