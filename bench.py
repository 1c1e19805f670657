#!/usr/bin/env python
"""bench.py -- HEBench samples/s of the ciphertext-evaluation hot path on B200.

Headline workload (BASELINE.json configs[1]): CKKS element-wise multiply + relinearize + rescale,
N = 8192, coefficient modulus {60,45,60} (3 limbs at key level, 2 data limbs), 1000 ciphertext
pairs per GPU per step.  One "sample" = one result ciphertext.  Inputs are synthetic: uniform random
residues (seed 1234), real key-switching keys from the host FHE stand-in.

    python bench.py [--gpus N] [--steps K] [--warmup W]            the B200 arm
    python bench.py --impl reference [...]                          the CPU arm (SEAL-restatement oracle)
    python bench.py --configs none|all|C1,C3,...                    which BASELINE configs run through the plugin (default all)

The JSON line carries, beside the contract's keys:
  * `configs`: EVERY BASELINE.json config at its stated shape through the HEBench plugin (libhebench_seal_backend.so driven
    by the mini harness in HEBench order: encode, encrypt, load, operate, store, decrypt, decode with pageable host
    ciphertexts, decoded results validated): samples/s of operate(), end-to-end samples/s over load + operate + store, the
    dominant kernel and its two roofline fractions (from the library's own work accounting, b200he_profile_work);
  * `sustained`: the headline step run back to back for >= 1 s with its clock / power record.

N > 1: launched by torchrun, one rank per GPU; the headline batch is sharded (weak scaling: 1000 pairs per GPU),
no data-path collective; torch.distributed is used for the barrier and the max-over-ranks time only.  The plugin's
own multi-GPU path (one process, HEB_B200_GPUS = N) is then driven by rank 0 for the strong-scaling configs (C3's
10^4-result grid, C5's batch: fixed work split over N GPUs) while the other ranks wait.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "reference-seal-backend_b200"))

N_POLY = 8192
COEFF_BITS = 45
DEPTH = 2                 # K = 3 primes {60,45,60}
BATCH = 1000              # ciphertext pairs per GPU per step
SEED = 1234
METRIC = "hebench_samples_per_s_ckks_eltwise_mul_relin_rescale_n8192"
UNIT = "samples/s"
CONFIG = {
    "workload": "CKKS eltwise multiply+relinearize+rescale, N=8192, coeff_modulus {60,45,60}, batch 1000 ciphertext pairs per GPU (BASELINE.json configs[1])",
    "poly_modulus_degree": N_POLY, "coeff_modulus_bits": [60, 45, 60], "batch_per_gpu": BATCH,
    "parallelism": "batch-sharded, no collective",
    "l2_policy": "inputs (2 x 262 MB) + intermediates per step exceed the 126 MB L2; no flush needed",
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def butterfly_peak():
    """register-resident modular butterflies per second of the whole chip, measured on this pool's B200: 64-bit Shoup
    butterflies on the integer (fma) pipe (tools/imad_peak.cu, profiles/r1d_imad_peak.json) and FP64-domain
    butterflies of primes below 2^46 on the FP64 pipe (tools/fp64_peak.cu, profiles/r1i_fp64_peak.json): the compute
    rooflines of the NTT-bearing kernels.  Returns (int_peak, dp_peak, source)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1d_imad_peak.json")) as f:
            ip = json.load(f)["bfly_exact_mulhi"]["Gbfly_per_s"] * 1e9
        with open(os.path.join(ROOT, "profiles", "r1i_fp64_peak.json")) as f:
            dp = json.load(f)["bfly_dp_45bit"]["Gbfly_per_s"] * 1e9
        return ip, dp, "measured (profiles/r1d_imad_peak.json, profiles/r1i_fp64_peak.json)"
    except Exception:
        return 1.0e12, 1.93e12, "fallback (DESIGN.md §3.1)"


DP_MAX_BITS = 46   # csrc/modarith.cuh B200HE_DP_MAX_BITS: moduli at or below run in the FP64 domain


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML; nvidia-smi as a fallback)"""

    def __init__(self, gpu, period=0.002):
        self.gpu, self.sm, self.mx, self.reasons, self.stop = gpu, [], None, set(), threading.Event()
        self.power = []
        self.period = period
        self.t = threading.Thread(target=self.run, daemon=True)
        self.n = 0

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                    "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while True:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
                self.n += 1
                if self.stop.wait(self.period):
                    break
        except Exception:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            while True:
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    r = [x.strip() for x in out.split(",")]
                    self.sm.append(int(r[0]))
                    self.mx = int(r[1])
                    self.reasons |= {names[i] for i in range(4) if r[2 + i].lower().startswith("active")}
                    self.n += 1
                except Exception:
                    pass
                if self.stop.wait(0.05):
                    break

    def __enter__(self):
        self.t.start()
        time.sleep(0.01)
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(self.sm)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "samples": self.n}
        if self.power:
            out["power_w_max"] = max(self.power)
        return out


def bind_to_gpu_numa_node(gpu):
    """Pin this rank to the CPU cores of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are allocated
    (first-touch places them on that node), so the per-step host<->device copies of several ranks do not all cross
    one socket's memory controller / the inter-socket link.  Best effort: returns a description or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(gpu)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]   # 0000:17:00.0
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} ({len(cpus)} cores)"
    except Exception:
        return None


def synth_inputs(moduli, n, L, N, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    out = np.empty((n, 2, L, N), dtype=np.uint64)
    for l in range(L):
        out[:, :, l, :] = rng.integers(0, int(moduli[l]), size=(n, 2, N), dtype=np.uint64)
    return out


def kernel_rooflines(prof, work, steps, peak_hbm, bf_int_peak, bf_dp_peak):
    """per kernel class: average launch time, HBM fraction (algorithmic bytes / time / measured copy bandwidth) and
    arithmetic-pipe fraction (time the launch's butterflies need at the measured register-resident rates of the integer
    and FP64 pipes / measured time).  prof: {class: (ms, launches)}, work: {class: (bfly_int, bfly_fp64, bytes)} from
    the library's own accounting (b200he_profile_work), both summed over `steps` steps."""
    out = {}
    for name, (ms, n) in prof.items():
        bi, bd, by = work.get(name, (0.0, 0.0, 0.0))
        sec = ms / 1e3
        floor_s = bi / bf_int_peak + bd / bf_dp_peak
        out[name] = {"ms_per_step": ms / steps, "launches_per_step": n / steps,
                     "hbm_GBps": by / sec / 1e9 if sec > 0 else None, "hbm_frac": by / sec / 1e9 / peak_hbm if sec > 0 else None,
                     "pipe_frac": floor_s / sec if sec > 0 and floor_s > 0 else None,
                     "butterflies_per_s": (bi + bd) / sec if sec > 0 and (bi + bd) > 0 else None,
                     "fp64_share_of_butterflies": bd / (bi + bd) if (bi + bd) > 0 else None,
                     "algorithmic_bytes_per_launch": by / n if n else None}
    return out


# ------------------------------------------------------------------------------------------- the BASELINE configs through the plugin
BACKEND = os.path.join(ROOT, "reference-seal-backend_b200", "backend")
HARNESS = os.path.join(BACKEND, "mini_harness")
PLUGIN = os.path.join(BACKEND, "libhebench_seal_backend.so")
# name -> (BASELINE.json config, harness arguments, timed operate() iterations); shapes are the stated ones
PLUGIN_CONFIGS = {
    "C1": ("configs[0]: BFV eltwise multiply ct x ct, N=8192, n=100, 10 x 10 samples", ["--filter", "EltwiseMultiply BFV Offline", "--n", "100", "--samples", "10,10"], 5),
    "C2add": ("configs[1]: CKKS eltwise add, N=8192 {60,45,60}, 100 x 10 = 1000 results", ["--filter", "EltwiseAdd CKKS Offline", "--n", "1000", "--samples", "100,10"], 5),
    "C2mul": ("configs[1]: CKKS eltwise multiply (the reference's descriptor: no relinearize / rescale), 1000 results",
              ["--filter", "EltwiseMultiply CKKS Offline", "--n", "1000", "--samples", "100,10"], 5),
    "C3": ("configs[2]: CKKS dot product n=100, N=16384 {60,40,60}, 100 x 100 = 10^4 results",
           ["--filter", "DotProduct CKKS Offline", "--n", "100", "--poly", "16384", "--samples", "100,100"], 2),
    "C5": ("configs[4]: CKKS logistic regression PolyD3, 16 features, N=32768 {60,45x5,60}, batch 1024",
           ["--filter", "LogisticRegression_PolyD3 CKKS Offline", "--poly", "32768", "--batch", "1024"], 2),
    "C4row": ("configs[3]: CKKS MatMult Row 100x100x100 (needs N=32768: cols_M0 * cols_M1 <= N/2), depth 3",
              ["--filter", "MatrixMultiply CKKS Latency other=2", "--dims", "100,100,100", "--poly", "32768", "--depth", "3"], 1),
    "C4val": ("configs[3]: CKKS MatMult Val 100x100x100, N=16384 {60,45x5,60}",
              ["--filter", "MatrixMultiply CKKS Latency other=0", "--dims", "100,100,100", "--poly", "16384", "--depth", "6"], 1),
    "C4cba": ("configs[3]: CKKS MatMult CipherBatchAxis 100x100x100, N=16384 {60,45x5,60} (2 x 10^4 input ciphertexts)",
              ["--filter", "MatrixMultiply CKKS Latency other=1", "--dims", "100,100,100", "--poly", "16384", "--depth", "6"], 1),
}
STRONG_SCALING = ["C3", "C5"]   # fixed work split over N GPUs by the plugin


def run_plugin_config(name, gpus, peak_hbm, bf_int_peak, bf_dp_peak, timeout=600):
    import tempfile
    desc, hargs, iters = PLUGIN_CONFIGS[name]
    with tempfile.TemporaryDirectory() as tmp:
        jpath, ppath = os.path.join(tmp, "harness.json"), os.path.join(tmp, "profile.json")
        env = dict(os.environ, HEB_B200_SEED=str(SEED), HEB_B200_GPUS=str(gpus), HEB_B200_PROFILE_JSON=ppath, HEB_B200_PROFILE_SKIP=str(1 + iters))
        env.pop("OMP_NUM_THREADS", None)   # torchrun pins it to 1; the host phases (encode / encrypt / decrypt) use every core
        t0 = time.time()
        try:
            p = subprocess.run([HARNESS, "--backend_lib_path", PLUGIN, "--iterations", str(iters), "--extra-operate", "1", "--json", jpath] + hargs,
                               capture_output=True, text=True, env=env, timeout=timeout)
        except subprocess.TimeoutExpired:
            return {"config": desc, "error": f"timeout after {timeout} s"}
        wall = time.time() - t0
        if "Failed: 0" not in p.stdout or not os.path.exists(jpath):
            return {"config": desc, "error": (p.stdout[-400:] + p.stderr[-400:]).strip()}
        with open(jpath) as f:
            h = json.loads(f.readline())
        prof = None
        if os.path.exists(ppath):
            with open(ppath) as f:
                prof = json.loads(f.readline())
    n = h["results_per_operate"]
    op_ms = h["operate_ms"]
    e2e_ms = h["load_ms"] + op_ms + h["store_ms"]
    out = {"config": desc, "gpus": gpus, "results_per_operate": n, "validated": h["validated"], "operate_ms": op_ms,
           "samples_per_s": n / (op_ms / 1e3), "load_ms": h["load_ms"], "store_ms": h["store_ms"],
           "e2e_samples_per_s": n / (e2e_ms / 1e3), "e2e_ms": e2e_ms,
           "e2e_path": "libhebench_seal_backend.so load -> operate -> store, pageable host ciphertexts",
           "host_ms": {k: h[k] for k in ("encode_ms", "encrypt_ms", "decrypt_ms", "decode_ms")}, "harness_wall_s": wall}
    if prof:
        ks = prof["kernels"]
        tot = sum(v["ms_sum_over_gpus"] for v in ks.values())
        top = max(ks, key=lambda k: ks[k]["ms_sum_over_gpus"])
        def frac(v):
            sec = v["ms_sum_over_gpus"] / 1e3
            floor_s = v["bfly_int"] / bf_int_peak + v["bfly_fp64"] / bf_dp_peak
            return {"ms": v["ms_sum_over_gpus"] / max(prof["gpus"], 1), "share": v["ms_sum_over_gpus"] / tot if tot else None,
                    "hbm_frac": v["bytes"] / sec / 1e9 / peak_hbm if sec > 0 else None, "pipe_frac": floor_s / sec if sec > 0 and floor_s > 0 else None}
        out["dominant_kernel"] = top
        out["dominant"] = frac(ks[top])
        out["kernels"] = {k: frac(v) for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["ms_sum_over_gpus"])}
        # whole operate(): the time its butterflies and its bytes need at the measured peaks / the profiled kernel time
        floor_pipe = sum(v["bfly_int"] / bf_int_peak + v["bfly_fp64"] / bf_dp_peak for v in ks.values())
        floor_hbm = sum(v["bytes"] for v in ks.values()) / (peak_hbm * 1e9)
        out["operate_roofline"] = {"kernel_ms": tot / max(prof["gpus"], 1), "pipe_floor_ms": floor_pipe * 1e3 / max(prof["gpus"], 1),
                                   "hbm_floor_ms": floor_hbm * 1e3 / max(prof["gpus"], 1),
                                   "frac_of_binding_floor": max(floor_pipe, floor_hbm) * 1e3 / tot if tot else None}
    return out


def run_plugin_configs(names, gpus, peak_hbm, bf_int_peak, bf_dp_peak, budget_s):
    out, t0 = {}, time.time()
    if not (os.path.exists(HARNESS) and os.path.exists(PLUGIN)):
        return {"error": "plugin / mini harness not built (python -c 'import __graft_entry__ as g; g.build()')"}
    for name in names:
        if time.time() - t0 > budget_s:
            out[name] = {"config": PLUGIN_CONFIGS[name][0], "skipped": f"time budget of {budget_s:.0f} s for the configs block used up"}
            continue
        out[name] = run_plugin_config(name, gpus, peak_hbm, bf_int_peak, bf_dp_peak)
    return out


def run_b200(args):
    import numpy as np
    import torch
    import pyb200he as hb
    from pyb200he.hostfhe import CKKS, Host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torchrun, one rank per GPU")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU path; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    from pyb200he.shard import Ranks
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # a CPU-side group for the one wait that must not occupy the GPUs: while rank 0 drives the plugin over all the GPUs,
        # the other ranks' processes must not keep an NCCL barrier kernel spinning on theirs (kernels of different
        # processes time-slice a GPU: the plugin would get half of it)
        cpu_group = dist.new_group(backend="gloo")
    ranks = Ranks(dist, device=torch.device("cuda", local))

    host = Host(CKKS, N_POLY, DEPTH, COEFF_BITS, COEFF_BITS, seed=SEED)
    ctx = hb.Context(CKKS, N_POLY, host.moduli, host.psi, 0, device=local)
    stream = torch.cuda.Stream(device=local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_relin_key(host.relin_key())
    L, K, N = host.Ltop, host.K, N_POLY
    scale = host.scale

    # this rank's shard of the global batch (weak scaling: BATCH pairs per GPU), pinned on the host
    a_np = synth_inputs(host.moduli, BATCH, L, N, SEED + 2 * rank)
    b_np = synth_inputs(host.moduli, BATCH, L, N, SEED + 2 * rank + 1)
    words_in, words_out = 2 * L * N, 2 * (L - 1) * N
    a_pin = torch.from_numpy(a_np.view(np.int64).reshape(-1)).pin_memory()
    b_pin = torch.from_numpy(b_np.view(np.int64).reshape(-1)).pin_memory()
    out_pin = torch.empty(BATCH * words_out, dtype=torch.int64).pin_memory()
    out_pin2 = torch.empty(BATCH * words_out, dtype=torch.int64).pin_memory()   # e2e results alternate between the two
    A, B, R = hb.Batch(ctx), hb.Batch(ctx), hb.Batch(ctx)
    A.resize(BATCH, 2, L, True, scale)
    B.resize(BATCH, 2, L, True, scale)

    def upload():
        A.upload_from(a_pin.data_ptr(), 0, BATCH)
        B.upload_from(b_pin.data_ptr(), 0, BATCH)

    def step():
        ctx.multiply(A, B, out=R)
        ctx.relinearize_rescale(R, out=R)   # relinearize_inplace + rescale_to_next_inplace, bit-identical, one fused entry

    # end-to-end leg: the batch is cut into chunks that travel through E2E_STREAMS contexts (one stream each), so
    # the H2D copy of one chunk, the kernels of another and the D2H copy of a third overlap on the PCIe/compute engines
    E2E_STREAMS, CHUNK = 3, 50
    e2e = []
    for i in range(E2E_STREAMS):
        cx = hb.Context(CKKS, N_POLY, host.moduli, host.psi, 0, device=local)
        st = torch.cuda.Stream(device=local)
        cx.set_stream(st.cuda_stream)
        cx.set_relin_key(host.relin_key())
        bufs = [hb.Batch(cx), hb.Batch(cx), hb.Batch(cx)]
        bufs[0].resize(CHUNK, 2, L, True, scale)
        bufs[1].resize(CHUNK, 2, L, True, scale)
        bufs[2].resize(CHUNK, 3, L, True, scale)   # sized for the largest intermediate (the size-3 product)
        e2e.append((cx, st, bufs))

    # Steps are pipelined one deep, as a streaming service would run them: step i's chunks are enqueued, then the host
    # waits for step i-1's results (complete in their own pinned buffer) while the copy engines already work on step i,
    # so the H2D engine never drains between steps.  Every step still moves its inputs host->device and its results
    # device->host inside the timed region.
    out_pins = [out_pin, out_pin2]
    step_done = [[torch.cuda.Event() for _ in range(E2E_STREAMS)] for _ in range(2)]

    def e2e_step(i):
        dst = out_pins[i % 2]
        for k, first in enumerate(range(0, BATCH, CHUNK)):
            cx, st, (ca, cb, cr) = e2e[k % E2E_STREAMS]
            n = min(CHUNK, BATCH - first)
            ca.upload_from(a_pin.data_ptr() + first * words_in * 8, 0, n)
            cb.upload_from(b_pin.data_ptr() + first * words_in * 8, 0, n)
            cx.multiply(ca, cb, n=n, out=cr)
            cx.relinearize_rescale(cr, out=cr)
            cr.download_to(dst.data_ptr() + first * words_out * 8, 0, n, wait=False)
        for s_, (cx, st, _) in enumerate(e2e):
            step_done[i % 2][s_].record(st)
        if i > 0:
            for ev in step_done[(i - 1) % 2]:
                ev.synchronize()   # the results of step i-1 are on the host

    def e2e_drain():
        for cx, st, _ in e2e:
            cx.sync()

    def barrier():
        ranks.barrier()
        torch.cuda.synchronize()

    max_over_ranks = ranks.max_over_ranks

    upload()
    ctx.sync()
    for _ in range(max(args.warmup, 3)):
        step()
    ctx.sync()

    # ---- device-resident timing: inputs already in HBM, CUDA events on the launching stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = ctx.launch_count()
    with ClockSampler(local) as clk:
        ctx.profile_begin()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        prof = ctx.profile_end()
        work = ctx.profile_work()
        barrier()
    launches = ctx.launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = clk.summary()

    # ---- sustained leg: the same step back to back for >= args.sustain seconds (clocks and power under a long load)
    sustained = None
    if args.sustain > 0:
        n_sus = max(int(args.sustain * 1.1 / (ms_total / args.steps / 1e3)), args.steps)
        barrier()
        with ClockSampler(local, period=0.01) as clk_s:
            ev0.record(stream)
            for _ in range(n_sus):
                step()
            ev1.record(stream)
            ctx.sync()
            barrier()
        sus_ms = max_over_ranks(ev0.elapsed_time(ev1))
        sustained = {"seconds": sus_ms / 1e3, "steps": n_sus, "ms_per_step": sus_ms / n_sus, "value": BATCH * world * n_sus / (sus_ms / 1e3),
                     "unit": UNIT, "clocks": clk_s.summary()}

    # ---- end to end through the C ABI with host buffers (pinned): H2D + compute + D2H per step.  Several streams
    # are involved, so the region is bracketed by events on the default stream that every stream joins.
    for i in range(2):
        e2e_step(i)
    e2e_drain()
    barrier()
    main = torch.cuda.current_stream()
    # the end-to-end region is timed too: its samples join the device-timed region's (a slower poll: the enqueue loop of
    # this region is Python and shares the interpreter with the sampling thread).  The sampler starts -- and sleeps its
    # start-up 10 ms -- before the region opens.
    with ClockSampler(local, period=0.02) as clk2:
        barrier()
        ev0.record(main)
        for _, st, _b in e2e:
            st.wait_event(ev0)
        for i in range(args.steps):
            e2e_step(i)
        for _, st, _b in e2e:
            main.wait_stream(st)   # the last step's results included
        ev1.record(main)
        barrier()
    clk.sm += clk2.sm
    clk.reasons |= clk2.reasons
    clk.n += clk2.n
    clk.mx = clk.mx or clk2.mx
    clocks = clk.summary()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
    e2e_launches = sum(cx.launch_count() for cx, _, _b in e2e)

    # ---- second half of BASELINE.json's metric: NTT limb-ops/s (batched forward / inverse transforms of the same
    # 1000 x 2 x L limbs through the C ABI), outside the timed regions above
    C0 = ctx.batch(a_np[:1], ntt_form=False, scale=scale)
    C0.resize(BATCH, 2, L, False, scale)
    C0.upload_from(a_pin.data_ptr(), 0, BATCH)
    F0, I0 = hb.Batch(ctx), hb.Batch(ctx)
    for _ in range(2):
        ctx.ntt_forward(C0, out=F0)
        ctx.ntt_inverse(F0, out=I0)
    ctx.profile_begin()
    for _ in range(5):
        ctx.ntt_forward(C0, out=F0)
        ctx.ntt_inverse(F0, out=I0)
    nprof = ctx.profile_end()
    assert np.array_equal(I0.download(0, 1), a_np[:1]), "inverse(forward) is not the identity"
    limbs = BATCH * 2 * L
    fwd_s = nprof["k_ntt_fwd"][0] / nprof["k_ntt_fwd"][1] / 1e3
    inv_s = nprof["k_ntt_inv"][0] / nprof["k_ntt_inv"][1] / 1e3
    bfly = (N // 2) * (N.bit_length() - 1)
    ntt_metric = {"N": N, "limbs_per_launch": limbs, "fwd_limb_ops_per_s": limbs / fwd_s, "inv_limb_ops_per_s": limbs / inv_s,
                  "fwd_butterflies_per_s": limbs * bfly / fwd_s, "fwd_GBps_algorithmic": limbs * 2 * N * 8 / fwd_s / 1e9}

    # sanity: the timed path produced the bits the step defines (first result vs a fresh single-ciphertext run)
    chk = ctx.multiply(ctx.batch(a_np[:1], scale=scale), ctx.batch(b_np[:1], scale=scale))
    ctx.relinearize(chk, out=chk)        # the two separate calls: the fused entry must give the same bits
    ctx.rescale_to_next(chk, out=chk)
    e2e_drain()
    got = out_pins[(args.steps - 1) % 2].numpy()[:words_out].view(np.uint64)
    assert np.array_equal(got, chk.download().reshape(-1)), "timed path result differs from a fresh evaluation"

    names = select_configs(args.configs, world)
    if rank != 0:
        if names:
            torch.cuda.synchronize()
            dist.barrier(group=cpu_group)   # rank 0 drives the plugin's own multi-GPU path over all the GPUs meanwhile
        if dist is not None:
            dist.destroy_process_group()
        return
    samples = BATCH * world * args.steps
    value = samples / (ms_total / 1e3)
    peak, peak_src = peaks()
    # dominant kernel class of the step
    top = max(prof.items(), key=lambda kv: kv[1][0])
    top_name, (top_ms, top_n) = top
    dp = [1 if int(q).bit_length() <= DP_MAX_BITS and not os.environ.get("B200HE_NO_DP") else 0 for q in host.moduli]
    bf_int_peak, bf_dp_peak, bf_src = butterfly_peak()
    # per kernel class, from the library's own work accounting of the timed launches (b200he_profile_work)
    kr = kernel_rooflines(prof, work, args.steps, peak, bf_int_peak, bf_dp_peak)
    top_r = kr[top_name]
    achieved = top_r["hbm_GBps"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(top_name)
    bf_rate = top_r["butterflies_per_s"]
    bf_peak = bf_rate / top_r["pipe_frac"] if bf_rate and top_r["pipe_frac"] else None
    # the whole step against its floors: the time its butterflies need at the measured pipe rates, its bytes at the measured copy rate
    step_pipe_floor = sum(w[0] / bf_int_peak + w[1] / bf_dp_peak for w in work.values()) / args.steps
    step_hbm_floor = sum(w[2] for w in work.values()) / args.steps / (peak * 1e9)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic (uniform random residues, seed 1234; real relinearization key)",
        "config": CONFIG, "clocks": clocks,
        "e2e": {"value": samples / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": 2 * BATCH * words_in * 8 * world,
                "d2h_bytes_per_step": BATCH * words_out * 8 * world, "ms_per_step": e2e_ms / args.steps,
                "pipeline": f"{E2E_STREAMS} streams x chunks of {CHUNK} ciphertext pairs, steps pipelined one deep (double-buffered results)", "host_binding": numa},
        "gpu_launches": launches * world,
        "roofline": {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                     "avg_launch_ms": top_ms / top_n, "share_of_step": top_ms / (ms_total_local(prof)),
                     "algorithmic_bytes_per_launch": top_r["algorithmic_bytes_per_launch"],
                     # the NTT-bearing kernels are bound by the arithmetic pipes, not HBM: fraction of the measured
                     # register-resident butterfly rate of the chip, 60-bit limbs on the integer pipe and limbs below
                     # 2^46 on the FP64 pipe, weighted by this launch's mix
                     "int_pipe": {"achieved": bf_rate, "peak": bf_peak, "unit": "butterflies/s",
                                  "frac": (bf_rate / bf_peak) if bf_rate else None, "peak_source": bf_src,
                                  "fp64_share_of_butterflies": top_r["fp64_share_of_butterflies"],
                                  "int_peak": bf_int_peak, "fp64_peak": bf_dp_peak},
                     "step": {"ms": ms_total / args.steps, "pipe_floor_ms": step_pipe_floor * 1e3, "hbm_floor_ms": step_hbm_floor * 1e3,
                              "frac_of_binding_floor": max(step_pipe_floor, step_hbm_floor) * 1e3 / (ms_total / args.steps)}},
        "kernels_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
        "kernels": {k: {kk: (round(vv, 4) if isinstance(vv, float) and kk.endswith(("frac", "ms_per_step")) else vv) for kk, vv in v.items()}
                    for k, v in sorted(kr.items(), key=lambda kv: -kv[1]["ms_per_step"])},
        "sustained": sustained,
    }
    n_dp = sum(dp[:L])
    ntt_peak = L / ((L - n_dp) / bf_int_peak + n_dp / bf_dp_peak)   # the transformed limbs cycle over the L data moduli
    ntt_metric["fwd_frac_of_butterfly_peak"] = ntt_metric["fwd_butterflies_per_s"] / ntt_peak
    ntt_metric["butterfly_peak"] = ntt_peak
    ntt_metric["fp64_limbs"] = f"{n_dp} of {L}"
    ntt_metric["fwd_frac_of_hbm_peak"] = ntt_metric["fwd_GBps_algorithmic"] / peak
    line["ntt_limb_ops"] = ntt_metric   # per GPU
    line["cpu_baseline"] = cpu_baseline(budget_s=12.0) if world == 1 else None
    # ---- every BASELINE config at its stated shape through the plugin (N > 1: the fixed-work ones, split over the N GPUs)
    if names:
        torch.cuda.empty_cache()   # (the harness processes allocate on the same GPUs: 180 GB leaves room for both)
        block = run_plugin_configs(names, world, peak, bf_int_peak, bf_dp_peak, budget_s=args.configs_budget)
        line["configs" if world == 1 else "strong_scaling"] = block
        if dist is not None:
            dist.barrier(group=cpu_group)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def select_configs(spec, world):
    """which plugin configs this run drives: every rank evaluates the same answer"""
    if spec == "none":
        return []
    if spec == "all":
        return list(PLUGIN_CONFIGS) if world == 1 else list(STRONG_SCALING)
    names = [x for x in spec.split(",") if x]
    for x in names:
        if x not in PLUGIN_CONFIGS:
            raise SystemExit(f"--configs: unknown config {x} (known: {', '.join(PLUGIN_CONFIGS)})")
    return names


def ms_total_local(prof):
    return sum(v[0] for v in prof.values())


# ------------------------------------------------------------------------------------------- CPU arm
def oracle_setup():
    """the SEAL-restatement oracle (oracle/, test infrastructure) -- used here only as the timed CPU baseline"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from helpers import CKKS, Oracle
    from pyb200he.hostfhe import Host
    host = Host(CKKS, N_POLY, DEPTH, COEFF_BITS, COEFF_BITS, seed=SEED)
    orc = Oracle(CKKS, N_POLY, host.moduli)
    return np, host, orc


def host_threads(orc):
    """all the host threads the CPU arm can use: the cores this process may run on -- not omp_get_max_threads(), which
    torchrun pins to 1 through OMP_NUM_THREADS for every rank it launches"""
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except AttributeError:
        return max(orc.o.orc_max_threads(), 1)


def cpu_sample(np, host, orc, n, threads):
    a = synth_inputs(host.moduli, n, host.Ltop, N_POLY, SEED).reshape(-1)
    b = synth_inputs(host.moduli, n, host.Ltop, N_POLY, SEED + 1).reshape(-1)
    key = host.relin_key()
    t = time.perf_counter()
    orc.mul_relin_rescale(host.Ltop, n, a, b, key, threads=threads)
    return time.perf_counter() - t


def cpu_baseline(budget_s=12.0):
    np, host, orc = oracle_setup()
    threads = host_threads(orc)
    probe = 4 * threads
    dt = cpu_sample(np, host, orc, probe, threads)
    n = int(max(probe, min(BATCH, probe * budget_s / max(dt, 1e-6))))
    cpu_sample(np, host, orc, n, threads)   # warm-up at the measured size (first touch of the result arrays, OpenMP team start):
    dt = cpu_sample(np, host, orc, n, threads)   # without it this figure came out at half of what `--impl reference` measures
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} of the {BATCH} ciphertext pairs of one step, {threads} OpenMP threads over the batch like the reference's operate(); "
                      "SEAL-restatement oracle (SEAL itself is not buildable offline)"}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    np, host, orc = oracle_setup()
    threads = host_threads(orc)
    probe = 4 * threads
    dt = cpu_sample(np, host, orc, probe, threads)
    total_steps = max(args.warmup, 0) + args.steps
    per_step_budget = min(20.0, 150.0 / max(total_steps, 1))
    n = int(max(threads, min(BATCH, probe * per_step_budget / max(dt, 1e-6))))
    for _ in range(args.warmup):
        cpu_sample(np, host, orc, n, threads)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_sample(np, host, orc, n, threads)
    value = n * args.steps / t
    sample = f"{n} of the {BATCH} ciphertext pairs per step, {threads} OpenMP threads (SEAL-restatement oracle; SEAL not buildable offline)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic (uniform random residues, seed 1234)", "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--configs", default="all", help="BASELINE configs driven through the plugin: all | none | C1,C2add,C2mul,C3,C4val,C4row,C4cba,C5")
    ap.add_argument("--configs-budget", type=float, default=200.0, help="wall-clock budget of the configs block in seconds (later configs are skipped)")
    ap.add_argument("--sustain", type=float, default=1.0, help="seconds of the sustained leg (0: skip)")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries that write to file descriptor 1 on their own (NCCL prints its
    # version there at communicator creation) are sent to stderr; the result line goes to the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
