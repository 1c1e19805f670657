#!/usr/bin/env python
"""Host<->device copy rates on this box (pinned memory), to size the end-to-end pipeline of bench.py."""
import json
import torch

dev = torch.device("cuda", 0)
n = 512 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(n // 4, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        s1.synchronize(); s2.synchronize()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def h2d():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    h2d(); d2h()


def h2d_chunks(k):
    def f():
        c = n // k
        with torch.cuda.stream(s1):
            for i in range(k):
                d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
    return f


out = {"h2d_GBps": n / timed(h2d) / 1e6, "d2h_GBps": (n // 4) / timed(d2h) / 1e6}
t = timed(both)
out["bidir_ms_for_512MiB_h2d_plus_128MiB_d2h"] = t
out["bidir_h2d_equiv_GBps"] = n / t / 1e6
for k in (8, 32, 128):
    out[f"h2d_{k}_chunks_GBps"] = n / timed(h2d_chunks(k)) / 1e6
print(json.dumps(out))
