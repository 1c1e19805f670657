#!/usr/bin/env python
"""store() / load() host path: where the time of b200he_batch_{download,upload}_scattered goes on this box.

C3's result grid is 10^4 ciphertexts of 524 288 bytes.  The probe moves n such ciphertexts between HBM and
  fresh   n separate malloc() blocks never touched before (what a std::vector<seal::Ciphertext> is on its first store)
  warm    the same blocks again (pages resident)
  slab    one anonymous mapping for all n, 2 MB aligned, madvise(MADV_HUGEPAGE), fresh and then warm
  pinned  one cudaMallocHost block, plain cudaMemcpyAsync (the PCIe ceiling of the direction)
and prints one JSON line.    python tools/host_path_probe.py [n]"""
import ctypes as C
import json
import mmap
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))
import torch

import pyb200he as hb
from pyb200he.hostfhe import CKKS, Host

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N, depth = 16384, 2
host = Host(CKKS, N, depth, 40, 40)
ctx = hb.Context(CKKS, N, host.moduli, host.psi, 0)
b = hb.Batch(ctx).resize(n, 2, depth, True)
ctb = 2 * depth * N * 8
libc = C.CDLL(None, use_errno=True)
libc.malloc.restype = C.c_void_p
libc.malloc.argtypes = [C.c_size_t]
libc.free.argtypes = [C.c_void_p]
libc.mmap.restype = C.c_void_p
libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
libc.munmap.argtypes = [C.c_void_p, C.c_size_t]


def timed(fn):
    ctx.sync()
    t = time.perf_counter()
    fn()
    ctx.sync()
    return time.perf_counter() - t


def scattered(ptrs, down):
    arr = (C.c_void_p * n)(*ptrs)
    fn = ctx.lib.b200he_batch_download_scattered if down else ctx.lib.b200he_batch_upload_scattered
    return timed(lambda: ctx._ck(fn(b.h, 0, n, arr)))


def thp():
    try:
        return open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
    except OSError:
        return None


out = {"n": n, "ciphertext_bytes": ctb, "GB": n * ctb / 1e9, "transparent_hugepage": thp(), "host_threads": os.cpu_count()}
gbps = lambda s: round(n * ctb / s / 1e9, 2)
# warm the staging buffers of the context
small = [libc.malloc(ctb) for _ in range(8)]
arr = (C.c_void_p * 8)(*small)
ctx._ck(ctx.lib.b200he_batch_download_scattered(b.h, 0, 8, arr))

blocks = [libc.malloc(ctb) for _ in range(n)]
out["store_fresh_GBps"] = gbps(scattered(blocks, True))
out["store_warm_GBps"] = gbps(scattered(blocks, True))
out["load_warm_GBps"] = gbps(scattered(blocks, False))
for p in blocks:
    libc.free(p)

HUGE = 2 << 20
span = n * ctb + HUGE
base = libc.mmap(None, span, mmap.PROT_READ | mmap.PROT_WRITE, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS, -1, 0)
assert base and base != C.c_void_p(-1).value, C.get_errno()
al = (base + HUGE - 1) & ~(HUGE - 1)
out["madvise_hugepage_rc"] = libc.madvise(al, n * ctb, 14)
slab = [al + i * ctb for i in range(n)]
out["store_slab_fresh_GBps"] = gbps(scattered(slab, True))
out["store_slab_warm_GBps"] = gbps(scattered(slab, True))
out["load_slab_warm_GBps"] = gbps(scattered(slab, False))
libc.munmap(base, span)

# the same without huge pages: one mapping, 4 KB pages
base = libc.mmap(None, span, mmap.PROT_READ | mmap.PROT_WRITE, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS, -1, 0)
libc.madvise(base, span, 15)   # MADV_NOHUGEPAGE
slab = [base + i * ctb for i in range(n)]
out["store_slab_4k_fresh_GBps"] = gbps(scattered(slab, True))
libc.munmap(base, span)

pin = torch.empty(n * ctb // 8, dtype=torch.int64, pin_memory=True)
b.download_to(pin.data_ptr(), 0, n)
out["store_pinned_GBps"] = gbps(timed(lambda: b.download_to(pin.data_ptr(), 0, n)))
out["load_pinned_GBps"] = gbps(timed(lambda: b.upload_from(pin.data_ptr(), 0, n)))
for T in (4, 8, 16, 32):
    os.environ["B200HE_HOST_THREADS"] = str(T)
    blocks = [libc.malloc(ctb) for _ in range(n)]
    out[f"store_fresh_T{T}_GBps"] = gbps(scattered(blocks, True))
    for p in blocks:
        libc.free(p)
print(json.dumps(out))
