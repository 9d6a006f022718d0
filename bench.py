#!/usr/bin/env python
"""Benchmark of the sliced-OT hot path (BASELINE.json metric: OT iterations/sec at conv4_1 of a 1024^2 image).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode cdf|sort|chol|pca|sym] [--impl ours|reference]

A "step" is one OT iteration  out = hist_match(P R, S R, mode) R^T  (optex.py:167-177) on one synthetic
feature block P, S = [1, 128, 128, 512] fp32 (N_p = N_s = 16384 pixels, C = 512 channels), rotation drawn
per step.  One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.

  value       device-resident inputs, CUDA-event timed, whole-job aggregate over all ranks (weak scaling:
              every rank transports its own independent feature block - no data-path collective)
  e2e         the same step through the host-buffer C-ABI call (optex_ot_step_host): pinned host memory in,
              host memory out, H2D + D2H inside the timed region
  roofline    the dominant kernel of the step: algorithmic bytes or flops / its event-timed duration
  cpu_baseline / --impl reference
              the CPU port of the reference (oracle/, pinned bit-exactly to the reference's own outputs)
              on the host cores of the same box, rotation draw included as the reference does (optex.py:168)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ot_iters_per_sec_conv4_1_1024"
UNIT = "it/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="cdf", choices=["cdf", "sort", "chol", "pca", "sym"])
    ap.add_argument("--gemm", default="auto", choices=["auto", "fp32", "tf32x3", "tf32"])
    ap.add_argument("--hw", type=int, default=128, help="feature-map side: conv4_1 of a 1024^2 image = 128")
    ap.add_argument("--channels", type=int, default=512)
    ap.add_argument("--sets", type=int, default=4, help="distinct input sets cycled so inputs exceed L2")
    ap.add_argument("--prewarm-s", type=float, default=1.0, dest="prewarm_s",
                    help="seconds of untimed steps before the warm-up (GPU clock ramp of a fresh box)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--all-modes", action="store_true", help="also time the other hist modes (extra keys)")
    ap.add_argument("--no-pdl", action="store_true", help="launch the kernels without programmatic dependent launch")
    ap.add_argument("--no-synthesis", action="store_true",
                    help="skip the end-to-end synthesis block (BASELINE metric ii: output pixels/sec of forward())")
    ap.add_argument("--synthesis-size", type=int, default=512, dest="synthesis_size")
    ap.add_argument("--synthesis-cpu", action="store_true", dest="synthesis_cpu",
                    help="also time the oracle's forward() of the same synthesis on the host cores (about a minute)")
    ap.add_argument("--sharded", action="store_true",
                    help="N > 1: ONE feature block, rotated channels sharded over the ranks + NCCL all-gather "
                         "(strong scaling) instead of one independent block per rank")
    return ap.parse_args()


LAYERS = {(64, 512): "conv5_1", (128, 512): "conv4_1", (256, 256): "conv3_1", (512, 128): "conv2_1", (1024, 64): "conv1_1"}


def workload_name(a):
    layer = LAYERS.get((a.hw, a.channels))
    where = f"{layer}@1024^2" if layer else "custom block"
    return (f"ot_step {where}: P,S=[1,{a.hw},{a.hw},{a.channels}] fp32 "
            f"(N_p=N_s={a.hw * a.hw}, C={a.channels}), hist_mode={a.mode}, rotation drawn per step")


def step_work(n_p, n_s, c):
    """algorithmic work of one OT iteration."""
    return {"flops": 2.0 * c * c * (2 * n_p + n_s), "bytes": 4.0 * c * (2 * n_p + n_s) + 4.0 * c * c}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe), read through NVML in-process.

    (Spawning `nvidia-smi -lms` next to the timed region cost the first bench of a fresh box 40 %: its start-up -
    NVML init, driver locks, CPU time - fell inside a 12 ms region.  NVML is initialised in __init__, i.e. before
    the pre-warm; a sample is three cheap queries every 2 ms.)"""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.stop_flag, self.thread, self.h, self.nv = [], threading.Event(), None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber: resolve through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                        self.h = h
                        break
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _one(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        except Exception:  # noqa: BLE001
            pw = float("nan")
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:  # noqa: BLE001
            mask = 0
        self.samples.append((sm, pw, mask))

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                self._one()
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"]}
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples"]}
        reasons = set()
        for _, _, mask in self.samples:
            for bit, name in self.REASONS.items():
                if mask & bit:
                    reasons.add(name)
        sm = [x[0] for x in self.samples]
        pw = [x[1] for x in self.samples if x[1] == x[1]]
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": self.max_sm,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(torch, a, seed, device):
    """SURVEY.md 8(d) synthetic input: P = relu(randn), S = relu(1.3 randn + 0.2), fp32 NHWC."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    shape = (1, a.hw, a.hw, a.channels)
    p = torch.relu(torch.randn(shape, generator=g))
    s = torch.relu(1.3 * torch.randn(shape, generator=g) + 0.2)
    return p.to(device), s.to(device)


# ------------------------------------------------------------------------------------------- reference arm
def cpu_reference_rate(a, steps, warmup, budget_s=None):
    """The reference's algorithm on the host cores: oracle port (bit-exact with the reference's outputs,
    tests/test_oracle_golden.py) including the per-iteration scipy-style rotation draw (optex.py:149,168)."""
    import numpy as np
    import torch

    from oracle import ot_oracle, rotation as rot_oracle, sort_oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p, s = make_inputs(torch, a, 0, "cpu")
    rng = np.random.RandomState(0)

    def one(p):
        r = torch.tensor(rot_oracle.haar_rotation_qr(a.channels, rng))          # float64 like optex.py:149
        if a.mode == "sort":
            return sort_oracle.ot_step_sort(p, s, r)
        return ot_oracle.ot_step(p, s, r, a.mode)

    for _ in range(warmup):
        p = one(p)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        p = one(p)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s and done >= 2:
            break
    dt = time.perf_counter() - t0
    return done / dt, dt / done * 1e3, done, cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, ms, done, cores = cpu_reference_rate(a, a.steps, a.warmup)
    sample = f"{done} full-size steps ({workload_name(a)}), oracle port incl. rotation draw, torch {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": done,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a)},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - optimaltextures_b200 has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    from optimaltextures_b200 import build as _build

    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    import optimaltextures_b200 as ob
    from optimaltextures_b200 import _lib
    from optimaltextures_b200._runtime import call, ptr, stream_ptr, workspace

    lib = _lib.lib()
    _lib.check(lib.optex_device_check())
    ob.set_gemm_mode(a.gemm)
    lib.optex_set_pdl(0 if a.no_pdl else 1)
    K, W = a.steps, a.warmup
    n = a.hw * a.hw
    c = a.channels
    mode = _lib.mode_id(a.mode)
    sets = [make_inputs(torch, a, 1000 * rank + i, device) for i in range(a.sets)]
    outs = [torch.empty_like(sets[0][0]) for _ in range(2)]
    ws = workspace(device, lib.optex_ot_workspace_bytes(n, n, c, mode))
    rot_ws = torch.empty(lib.optex_rotations_workspace_bytes(c, K), dtype=torch.uint8, device=device)
    rots = torch.empty(K, c, c, dtype=torch.float32, device=device)
    st = stream_ptr(device)

    def gen_rotations(count, first):
        call("optex_random_rotations", ptr(rots), c, count, 1234 + (0 if a.sharded else rank), first, None, ptr(rot_ws), rot_ws.numel(), st)

    sharded = a.sharded and world > 1
    if sharded:
        from optimaltextures_b200 import parallel

        sets = [make_inputs(torch, a, i, device) for i in range(a.sets)]      # replicated inputs
        shard_ops = parallel.cuda_ops()

    def step(i):
        p, s = sets[i % a.sets]
        if sharded:
            parallel.optimal_transport_sharded(p, s, a.mode, rots[i % K], ops=shard_ops)
            return
        call("optex_ot_step", ptr(p), ptr(s), ptr(rots[i % K]), ptr(outs[i % 2]), 1, n, 1, n, c, mode, 1.0, None,
             0.0, ptr(ws), ws.numel(), st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident inputs
    gen_rotations(K, 0)
    sampler = ClockSampler(local)      # NVML init happens here, outside the timed region
    # untimed pre-warm on top of the W warm-up steps: a fresh box idles at low clocks and the first ~100 ms of
    # work run up to 2x slow (measured: 486 vs 265 us/step for the first bench of a box) - W steps are only ~1.5 ms
    t_pre = time.perf_counter()
    n_pre = 0
    while time.perf_counter() - t_pre < a.prewarm_s:
        for i in range(20):
            step(n_pre + i)
        n_pre += 20
        torch.cuda.synchronize()
    for i in range(W):
        step(i)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.optex_launch_count()
    barrier()
    fence = torch.zeros(1, device=device)
    e0.record()
    gen_rotations(K, W)                       # the per-step rotation draw, batched (as optex_ot_loop does)
    for i in range(K):
        step(i)
    fence.add_(1)      # an ordinary kernel: it starts only after the last (PDL-launched) kernel has fully drained
    e1.record()
    barrier()
    launches = lib.optex_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / K
    value = (1 if sharded else world) * K / (ms_total * 1e-3)

    # ---- breakdown: the step's stages through the exported building blocks, event-timed in the same loop
    stages = {}
    if rank == 0:
        per_channel = a.mode in ("cdf", "sort")
        if per_channel:
            rp = torch.empty(c, n, dtype=torch.float32, device=device)
            rs = torch.empty(c, n, dtype=torch.float32, device=device)
            names = ["split_R", "rotate_forward_P", "rotate_forward_S", f"{a.mode}_match", "rotate_inverse"]
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(K)]
            pws = torch.empty(lib.optex_rotation_prepare_workspace_bytes(c), dtype=torch.uint8, device=device)
            mws = torch.empty(max(lib.optex_cdf_match_workspace_bytes(c, 256),
                                  lib.optex_sort_match_workspace_bytes(c, n, n), 256), dtype=torch.uint8, device=device)
            lib.optex_set_pdl(0)      # events between kernels need ordinary stream serialisation
            for i in range(K):
                p, s = sets[i % a.sets]
                r = rots[i % K]
                ev[i][0].record()
                # R -> tf32 hi / lo once for the three GEMMs, as optex_ot_step does (so each rotate_* stage below is
                # exactly one GEMM launch)
                call("optex_rotation_prepare", ptr(r), c, ptr(pws), pws.numel(), st)
                ev[i][1].record()
                call("optex_rotate_forward", ptr(p), ptr(r), ptr(rp), n, c, st)
                ev[i][2].record()
                call("optex_rotate_forward", ptr(s), ptr(r), ptr(rs), n, c, st)
                ev[i][3].record()
                if a.mode == "cdf":
                    call("optex_cdf_match", ptr(rp), ptr(rs), ptr(rp), c, n, n, 256, None, ptr(mws), mws.numel(), st)
                else:
                    call("optex_sort_match", ptr(rp), ptr(rs), ptr(rp), c, n, n, None, ptr(mws), mws.numel(), st)
                ev[i][4].record()
                call("optex_rotate_inverse", ptr(rp), ptr(r), ptr(outs[i % 2]), n, c, None, 0.0, st)
                ev[i][5].record()
            call("optex_rotation_prepare", None, 0, None, 0, None)
            torch.cuda.synchronize()
            lib.optex_set_pdl(0 if a.no_pdl else 1)
            for j, name in enumerate(names):
                stages[name] = statistics.mean(ev[i][j].elapsed_time(ev[i][j + 1]) for i in range(K))

    # ---- e2e: host buffers through optex_ot_step_host (H2D + step + D2H per call)
    e2e = None
    if not a.no_e2e:
        hp = [torch.empty(1, a.hw, a.hw, c).pin_memory() for _ in range(2)]
        hs = [torch.empty(1, a.hw, a.hw, c).pin_memory() for _ in range(2)]
        ho = [torch.empty(1, a.hw, a.hw, c).pin_memory() for _ in range(2)]
        for i in range(2):
            hp[i].copy_(sets[i % a.sets][0]); hs[i].copy_(sets[i % a.sets][1])
        Ke = max(4, min(K, 30))
        # (1) synchronous call: H2D -> step -> D2H -> sync, one step at a time
        for i in range(2):
            ob.optimal_transport_host(hp[i % 2], hs[i % 2], None, a.mode, out=ho[0], seed=99, counter=i)
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            ob.optimal_transport_host(hp[i % 2], hs[i % 2], None, a.mode, out=ho[0], seed=99, counter=2 + i)
        torch.cuda.synchronize()
        dt_sync = max_over_ranks(time.perf_counter() - t0)
        # (2) the double-buffered form of the same call (optex_ot_step_host_async, slots 0/1 on two streams):
        #     independent steps, so step i+1 uploads while step i computes and downloads
        streams = [torch.cuda.Stream(device=device) for _ in range(2)]
        for i in range(2):
            ob.optimal_transport_host(hp[i], hs[i], None, a.mode, out=ho[i], seed=99, counter=i, slot=i,
                                      stream=streams[i])
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            ob.optimal_transport_host(hp[i % 2], hs[i % 2], None, a.mode, out=ho[i % 2], seed=99, counter=2 + i,
                                      slot=i % 2, stream=streams[i % 2])
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * Ke / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * 4 * n * c,
               "d2h_bytes_per_step": 4 * n * c, "steps": Ke, "ms_per_step": dt / Ke * 1e3,
               "api": "optex_ot_step_host_async (C-ABI, pinned host buffers in and out, rotation drawn on device, "
                      "two slots double-buffered so uploads overlap compute + download)",
               "unpipelined": {"value": world * Ke / dt_sync, "ms_per_step": dt_sync / Ke * 1e3,
                               "api": "optex_ot_step_host (one synchronous call per step)"}}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    pk = peaks()
    work = step_work(n, n, c)
    roofline = None
    kernels = {}
    if stages:
        alg = {
            "rotate_forward_P": ("tensor", 2.0 * c * c * n), "rotate_forward_S": ("tensor", 2.0 * c * c * n),
            "rotate_inverse": ("tensor", 2.0 * c * c * n),
            f"{a.mode}_match": ("hbm", 4.0 * c * (n + n) + 4.0 * c * n),
            "split_R": ("hbm", 4.0 * c * c * 3),
        }
        for name, ms in stages.items():
            bound, amount = alg[name]
            if bound == "tensor":
                ach, peak, unit = amount / (ms * 1e-3) / 1e12, pk["bf16_tflops_sustained"], "TFLOP/s"
            else:
                ach, peak, unit = amount / (ms * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s"
            kernels[name] = {"ms": ms, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak}
        # dominant KERNEL: the matcher stage is several launches (its largest kernel is ~half of it, see
        # profiles/r01_launch_list_summary.csv), the rotation stages are one GEMM launch each
        single = {k_: v for k_, v in stages.items() if k_.startswith("rotate_")}
        top = max(single, key=single.get)
        for name in kernels:
            kernels[name]["launches"] = 1 if (name.startswith("rotate_") or name == "split_R") else (3 if a.mode == "cdf" else 2)
        k = kernels[top]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{top}:{a.mode}:{a.gemm}")
        roofline = {"kernel": top, "bound": k["bound"], "achieved": k["achieved"], "peak": k["peak"], "unit": k["unit"],
                    "frac": k["frac"], "traffic": traffic, "peak_source": pk["source"] +
                    (" bf16 sustained (tf32 tensor peak is nominally half of it)" if k["bound"] == "tensor" else " copy")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "gemm": a.gemm,
                   "prewarm": f"{a.prewarm_s:g} s of untimed steps before the {W} warm-up steps (clock ramp)",
                   "l2": f"inputs cycle over {a.sets} distinct (P,S) sets = {a.sets * 2 * 4 * n * c / 1e6:.0f} MB > 126 MB L2",
                   "sharding": ("one block, rotated channels sharded over ranks + NCCL all-gather" if sharded else
                                "independent feature blocks per rank, no data-path collective")},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "kernels": kernels,
        "step_roofline": {"hbm_frac": work["bytes"] / (ms_step * 1e-3) / 1e9 / pk["hbm_gbs"],
                          "tensor_frac": work["flops"] / (ms_step * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                          "flops": work["flops"], "bytes": work["bytes"]},
    }
    if world == 1 and not a.no_cpu_baseline:
        rate, ms, done, cores = cpu_reference_rate(a, steps=12, warmup=1, budget_s=20.0)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": ms,
                                "sample": f"{done} full-size steps of the same workload (oracle port incl. rotation "
                                          f"draw, torch {cores} threads)"}
    if a.all_modes and world == 1:
        extra = {}
        for m in ("cdf", "sort", "chol", "pca", "sym"):
            if m == a.mode:
                continue
            try:
                mid = _lib.mode_id(m)
                w2 = workspace(device, lib.optex_ot_workspace_bytes(n, n, c, mid))
                def st2(i):
                    p, s = sets[i % a.sets]
                    call("optex_ot_step", ptr(p), ptr(s), ptr(rots[i % K]), ptr(outs[i % 2]), 1, n, 1, n, c, mid,
                         1.0, None, 0.0, ptr(w2), w2.numel(), st)
                for i in range(3):
                    st2(i)
                torch.cuda.synchronize()
                e0.record()
                for i in range(K):
                    st2(i)
                e1.record()
                torch.cuda.synchronize()
                extra[m] = K / (e0.elapsed_time(e1) * 1e-3)
            except Exception as exc:  # noqa: BLE001
                extra[m] = f"error: {exc}"
        line["other_modes_it_s"] = extra
        gm = {}
        for g in ("tf32", "fp32"):       # rotation-GEMM arithmetic variants of the headline mode
            try:
                ob.set_gemm_mode(g)
                for i in range(3):
                    step(i)
                torch.cuda.synchronize()
                e0.record()
                for i in range(K):
                    step(i)
                e1.record()
                torch.cuda.synchronize()
                gm[g] = K / (e0.elapsed_time(e1) * 1e-3)
            except Exception as exc:  # noqa: BLE001
                gm[g] = f"error: {exc}"
            finally:
                ob.set_gemm_mode(a.gemm)
        line["other_gemm_modes_it_s"] = gm
    if world == 1 and not a.no_synthesis:
        try:
            line["synthesis"] = synthesis_block(a)
        except Exception as exc:  # noqa: BLE001  (the headline line must survive a failure of the extra block)
            line["synthesis"] = {"error": f"{type(exc).__name__}: {exc}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- end-to-end synthesis
def synthesis_block(a):
    """BASELINE.json metric (ii): output pixels / second of OptimalTexture.forward (optex.py:81-139, timed like the
    reference's own `time()` pair at optex.py:287-289 but with a device synchronisation on both sides) on
    configs[1]: texture synthesis, all five VGG layers, hist_mode pca, 5 passes, 500 iterations.  Synthetic style
    image of the bundled graffiti.jpg's shape at that size, random-init weights of the reference's architecture
    (its ./models/*.pth are not in this repository), rotations drawn on the device."""
    import torch

    import optimaltextures_b200 as ob
    from optimaltextures_b200 import texture, vgg

    size = a.synthesis_size
    kw = dict(size=size, iters=500, passes=5, hist_mode="pca")
    sd = vgg.random_state_dicts(0)
    g = torch.Generator().manual_seed(0)
    style = torch.rand(1, 3, round(size * 736 / 512 / 32) * 32, size, generator=g)
    pastiche = torch.rand(1, 3, size, size, generator=g)
    model = texture.OptimalTexture(state_dicts=sd, **kw)
    dev_style, dev_pastiche = style.cuda(), pastiche.cuda()
    lib = ob._lib.lib()
    out = {"workload": f"texture synthesis {size}^2 (configs[1]): 5 VGG layers, hist pca, 5 passes, 500 iters; style "
                       f"{tuple(style.shape)} synthetic, random-init weights", "unit": "px/s"}
    for rep in range(2):                                   # first run allocates workspaces: warm-up
        ob.manual_seed(0)
        model.ot_calls = 0
        model.profile = {} if rep == 1 else None
        l0 = lib.optex_launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = model.forward(dev_pastiche, [dev_style])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        launches = lib.optex_launch_count() - l0
    out.update({"value": res.shape[0] * res.shape[2] * res.shape[3] / dt, "seconds": dt, "ot_iters": model.ot_calls,
                "gpu_launches": int(launches), "pca_k_last_pass": model.last_pca_k,
                "pca_jacobi_sweeps_per_pass": model.pca_sweeps,
                "stage_ms": {k: round(v, 2) for k, v in model.stage_ms().items()},
                "finite": bool(torch.isfinite(res).all())})
    # additive option: component counts rounded up to multiples of 32 (tensor-core path for the C x C chains)
    model32 = texture.OptimalTexture(state_dicts=sd, pca_round_k=32, **kw)
    for rep in range(2):
        ob.manual_seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res32 = model32.forward(dev_pastiche, [dev_style])
        torch.cuda.synchronize()
        dt32 = time.perf_counter() - t0
    out["pca_round_k_32"] = {"value": res32.shape[0] * res32.shape[2] * res32.shape[3] / dt32, "seconds": dt32,
                             "pca_k_last_pass": model32.last_pca_k, "finite": bool(torch.isfinite(res32).all())}
    if a.synthesis_cpu:
        from oracle import texture_oracle

        torch.set_num_threads(os.cpu_count() or 1)
        rot_cache = {}

        def rot(c, index):                                 # the reference draws one scipy rotation per call
            from scipy.stats import special_ortho_group

            return torch.tensor(special_ortho_group.rvs(c)) if c > 1 else torch.ones(1, 1, dtype=torch.float64)

        cpu_model = texture_oracle.OptimalTexture(sd, rotation_fn=rot, **kw)
        t0 = time.perf_counter()
        with torch.inference_mode():
            ref = cpu_model.forward(pastiche, [style])
        dt_cpu = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": ref.shape[0] * ref.shape[2] * ref.shape[3] / dt_cpu, "unit": "px/s",
                               "seconds": dt_cpu, "cores": os.cpu_count(), "kind": "port",
                               "sample": "the whole synthesis once (oracle port, scipy rotation per OT call)"}
    return out


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
