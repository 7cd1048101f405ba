#!/usr/bin/env python
"""Benchmark of the geodesic curve-energy hot path (BASELINE.json metric: spline-steps/sec,
8778-pair 10-decoder eVAE energy optimisation).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision f16|f16x3|tf32|fp32]

A "step" is one Adam step of EVERY curve of the workload (= n_curves spline-steps): spline
evaluation, all-decoder forward + input-gradient backward, MC pair energy, Adam -- one pass of
the hot path over the whole pair list.  For N > 1 (torchrun, one rank per GPU) the pair list is
split into contiguous shards, no collective on the step path (strong scaling of the named job).

One JSON line on stdout (rank 0).  `value` = device-resident throughput, `e2e` = the same
metric through the public API with host buffers (H2D of curve state + D2H of results inside the
timed region), `roofline` = algorithmic tensor FLOP/s of the step kernel against the measured
dense bf16 peak, `cpu_baseline` = the oracle's PyTorch CPU port of the reference loop on this
box's host cores (bounded sample).  `--impl reference` prints the CPU arm alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_CURVES = 8778          # 133 classes -> 133*132/2 pairs (BASELINE config 3)
T_POINTS = 2000          # src/optimize.py:130
N_POLY = 4
K_DEC = 10
M_MC = 2
FLOP_PER_POINT_DECODER = 92160  # fwd + input-grad bwd of 2->128->128->50
METRIC = "spline-steps/sec, 8778-pair 10-decoder eVAE energy opt"
# arithmetic of the two 128-wide decoder layers (layer 1, the energy, the spline and Adam are fp32 everywhere)
PRECISION_NOTE = {
    "f16": "f16: tcgen05 kind::f16, fp16 operands (11-bit significand, as TF32), fp32 accumulate; <=1e-3 rel. on lengths",
    "f16x3": "f16x3: tcgen05 kind::f16, hi+lo fp16 operands, 3 MMAs per product, fp32 accumulate; fp32-grade (2e-6 rel. per-step energy)",
    "tf32": "tf32: tcgen05 kind::tf32, fp32 accumulate; <=1e-3 rel. on lengths",
    "fp32": "fp32: CUDA-core FFMA; <=1e-4 rel. per-step energy",
}


def synthetic_workload(n_curves, seed=0):
    """tasic-pca50-shaped synthetic job: the committed eVAE seed-12 decoder weights when the
    golden file is present (else default nn.Linear init), random end points in the latent box,
    near-straight initial splines (the 'euclidean' init is ~0)."""
    gold = ROOT / "tests" / "golden" / "evae_seed12_decoders.npz"
    if gold.exists():
        w = dict(np.load(gold))
        weights = "evae_seed12 checkpoint decoders"
    else:
        torch.manual_seed(seed)
        import torch.nn as nn
        nets = [nn.Sequential(nn.Linear(2, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 50))
                for _ in range(K_DEC)]
        w = {}
        for name, idx in (("1", 0), ("2", 2), ("3", 4)):
            w["W" + name] = torch.stack([n[idx].weight.detach() for n in nets]).numpy()
            w["b" + name] = torch.stack([n[idx].bias.detach() for n in nets]).numpy()
        weights = "random-init decoders"
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(n_curves, 2, generator=g) * 7 - 3.5
    b = torch.rand(n_curves, 2, generator=g) * 7 - 3.5
    omega = 0.05 * torch.randn(n_curves, N_POLY + 1, 2, generator=g)
    return w, a, b, omega, weights


def shard_range(n, rank, world):
    from vlg_b200.sharding import shard_range as sr
    return sr(n, rank, world)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (dense bf16; TF32 tensor rate is half)"
    return 1400.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle's PyTorch port of the reference loop (src/optimize.py:152-162)
# ----------------------------------------------------------------------------------------------

def cpu_reference_rate(w, a, b, omega, steps, warmup, budget_s):
    """spline-steps/s of the reference algorithm on this box's host cores, on a bounded sample
    of the workload (first `sample` curves; steps are homogeneous)."""
    from oracle import torch_port as TP
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    decs = [TP.make_decoder({k: w[k][i] for k in ("W1", "b1", "W2", "b2", "W3", "b3")}) for i in range(K_DEC)]
    import vlg_b200
    basis, _ = vlg_b200.construct_nullspace_basis(N_POLY)
    t = torch.linspace(0, 1, T_POINTS)

    def run(sample, nsteps):
        m = TP.SplineBatch(a[:sample].clone(), b[:sample].clone(), basis, omega[:sample].clone(), N_POLY)
        t0 = time.perf_counter()
        TP.run_steps(m, decs, t, nsteps, M=M_MC)
        return time.perf_counter() - t0

    probe = run(8, 1)                       # also warms the allocator / thread pool
    probe = min(probe, run(8, 1))
    per_curve_step = probe / 8
    total_steps = max(1, steps) + max(0, warmup)
    sample = int(max(4, min(200, budget_s / (per_curve_step * total_steps))))  # reference batch-size is 200
    if warmup:
        run(sample, warmup)
    dt = run(sample, max(1, steps))
    rate = sample * max(1, steps) / dt
    return rate, cores, f"{sample} curves x {max(1, steps)} steps (T=2000, K=10, M=2), torch {torch.__version__} CPU, {torch.get_num_threads()} threads"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, a, b, omega, weights = synthetic_workload(N_CURVES)
    rate, cores, sample = cpu_reference_rate(w, a, b, omega, args.steps, args.warmup, budget_s=90.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "spline-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * N_CURVES / rate,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(weights, args.gpus, "gpu"), precision="fp32 (PyTorch CPU)"),
        "cpu_baseline": {"value": rate, "unit": "spline-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "spline-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(weights, gpus, where):
    return {"workload": "BASELINE config 3: 8778 pairs (133 classes), 10-decoder eVAE, T=2000, n_poly=4, M=2",
            "n_curves": N_CURVES, "T": T_POINTS, "K": K_DEC, "M": M_MC, "n_poly": N_POLY, "weights": weights,
            "sharding": f"pair list split over {gpus} GPU(s), no collective on the step path",
            "cache": "L2 flushed (256 MiB write) between timed launches" if where == "gpu" else "n/a"}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def run_gpu_arm(args):
    import torch.distributed as dist
    import vlg_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local if world > 1 else 0)
    torch.cuda.set_device(dev)
    vlg_b200.build.build()

    w, a, b, omega, weights = synthetic_workload(N_CURVES)
    lo, hi = shard_range(N_CURVES, rank, world)
    n_local = hi - lo
    dec = vlg_b200.DecoderEnsemble.from_arrays(*[w[k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")], dev)
    basis, _ = vlg_b200.construct_nullspace_basis(N_POLY)
    basis = basis.to(dev)
    t = torch.linspace(0, 1, T_POINTS, device=dev)
    # pinned host copies of this rank's shard (e2e leg) and device-resident state (value leg)
    h_a, h_b, h_om = (x[lo:hi].contiguous().pin_memory() for x in (a, b, omega))
    model = vlg_b200.GeodesicSplineBatch(h_a.to(dev), h_b.to(dev), basis, h_om.to(dev), N_POLY)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    chunk = max(1, min(args.steps, args.chunk))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def launch(nsteps):
        return vlg_b200.optimize_splines(model, dec, t, nsteps, M=M_MC, seed=0, curve_id0=lo, precision=args.precision)

    # clocks are sampled (nvidia-smi, 100 ms period) from the warm-up on: the timed region of a sharded run can
    # be shorter than one sampling period, so the same load is kept up before and after it (see below)
    sampler = ClockSampler(dev.index)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        launch(1)
    barrier()

    # ---- value: K steps, state resident in HBM ----
    kernel_ms, launches = 0.0, 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    done = 0
    kernel_events = []
    while done < args.steps:
        ns = min(chunk, args.steps - done)
        flush.fill_(done & 0xFF)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        launch(ns)
        k1.record()
        kernel_events.append((k0, k1, ns))
        launches += 1
        done += ns
    ev1.record()
    barrier()
    # untimed tail under the same load until the sampler has seen the GPU busy at least three times
    t_tail = time.time()
    while len(sampler.rows) < 3 and time.time() - t_tail < 2.0:
        launch(1)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed region + same-load tail"
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = sum(x.elapsed_time(y) for x, y, _ in kernel_events)

    # ---- e2e: same steps through the public API with host buffers ----
    h_out_om = torch.empty_like(h_om).pin_memory()
    h_out_e = torch.empty(n_local, dtype=torch.float32).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_call(ns):
        m2 = vlg_b200.GeodesicSplineBatch(h_a.to(dev, non_blocking=True), h_b.to(dev, non_blocking=True), basis,
                                          h_om.to(dev, non_blocking=True), N_POLY)
        en = vlg_b200.optimize_splines(m2, dec, t, ns, M=M_MC, seed=0, curve_id0=lo, precision=args.precision)
        h_out_om.copy_(m2.omega, non_blocking=True)
        h_out_e.copy_(en, non_blocking=True)

    e2e_call(1)   # untimed warm-up of this path (first pinned H2D + allocator growth cost ~90 ms once)
    barrier()
    e0.record()
    done = 0
    while done < args.steps:
        ns = min(chunk, args.steps - done)
        e2e_call(ns)
        done += ns
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    h2d = (h_a.numel() + h_b.numel() + h_om.numel()) * 4
    d2h = (h_out_om.numel() + h_out_e.numel()) * 4

    # max over ranks
    times = torch.tensor([ms_total, ms_e2e, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, kernel_ms = (float(x) for x in times.cpu())

    if rank == 0:
        spline_steps = N_CURVES * args.steps
        value = spline_steps / (ms_total * 1e-3)
        peak, peak_src = measured_peak()
        # roofline of the step kernel on THIS rank: algorithmic FLOPs / kernel time
        flops_local = n_local * args.steps * T_POINTS * K_DEC * FLOP_PER_POINT_DECODER
        achieved = flops_local / (kernel_ms * 1e-3) / 1e12
        traffic = None
        tf = ROOT / "profiles" / "r01_traffic.json"
        if tf.exists() and args.precision in json.loads(tf.read_text()):
            # DRAM bytes per launch, scaled from the committed ncu capture of the same kernel
            per = json.loads(tf.read_text())[args.precision]["dram_bytes_per_spline_step"]
            traffic = per * n_local * args.steps / max(1, len(kernel_events))
        line = {
            "metric": METRIC, "value": value, "unit": "spline-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": {"f16": "f16", "f16x3": "f16x3", "tf32": "tf32", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": dict(workload_config(weights, world, "gpu"), precision=PRECISION_NOTE[args.precision],
                           steps_per_launch=chunk),
            "clocks": clocks,
            "e2e": {"value": spline_steps / (ms_e2e * 1e-3), "unit": "spline-steps/s",
                    "h2d_bytes_per_step": h2d * len(kernel_events) / args.steps,
                    "d2h_bytes_per_step": d2h * len(kernel_events) / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": {"f16": "tc_curve_kernel<true, FMT_F16>", "f16x3": "tc_curve_kernel<true, FMT_F16X3>",
                                    "tf32": "tc_curve_kernel<true, FMT_TF32>", "fp32": "simt_curve_kernel<true>"}[args.precision],
                         "flop_per_spline_step": T_POINTS * K_DEC * FLOP_PER_POINT_DECODER,
                         "kernel_ms_per_step": kernel_ms / args.steps},
        }
        if world == 1 and not args.no_cpu:
            rate, cores, sample = cpu_reference_rate(w, a, b, omega, steps=2, warmup=1, budget_s=20.0)
            line["cpu_baseline"] = {"value": rate, "unit": "spline-steps/s", "cores": cores, "kind": "port",
                                    "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vlg", choices=["vlg", "reference"])
    ap.add_argument("--precision", default="f16", choices=["f16", "f16x3", "tf32", "fp32"])
    ap.add_argument("--chunk", type=int, default=50, help="Adam steps per kernel launch")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
