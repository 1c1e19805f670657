#!/usr/bin/env python
"""The box's CONCURRENT host<->device ceiling: every rank (one per GPU, under torchrun) moves the bench step's traffic
(512 MiB host->device + 128 MiB device->host, pinned memory, both directions at once) between the same two barriers; the
aggregate rate is the denominator of bench.py's end-to-end scaling curve.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_concurrent_probe.py"""
import json
import os

import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_in, n_out = 512 << 20, 128 << 20
h_in, h_out = torch.empty(n_in, dtype=torch.uint8).pin_memory(), torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n_in, dtype=torch.uint8, device="cuda"), torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
best = 1e9
for rep in range(6):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a)
    s2.wait_event(a)
    for _ in range(4):   # four steps' worth back to back
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rep > 0:
        best = min(best, float(t.item()))
if rank == 0:
    per_step_ms = best / 4
    print(json.dumps({"gpus": world, "ms_per_step_traffic": per_step_ms, "aggregate_GBps": world * (n_in + n_out) / per_step_ms / 1e6,
                      "per_gpu_GBps": (n_in + n_out) / per_step_ms / 1e6,
                      "bench_samples_per_s_ceiling": world * 1000 / (per_step_ms / 1e3) * (524288000 + 131072000) / (n_in + n_out),
                      "note": "512 MiB H2D + 128 MiB D2H per rank per step, both directions concurrently, max over ranks, best of 5"}))
if world > 1:
    dist.destroy_process_group()
