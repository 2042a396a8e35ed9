#!/usr/bin/env python
"""bench.py — curve samples/s of the hot path on N B200s (contract in the task prompt, ④).

Headline workload (BASELINE.json configs[4], the one north_star's target is quoted on):
  C5  dynamic clamped equidistant cubic BSpline over [f32;4] colour stops (examples/gradient.rs
      shape, 5 stops), `take(2^32)`; the GLOBAL sample range is sharded contiguously over the N
      ranks (strong scaling: total work fixed; no collective on the data path).
A step = one pass of the hot path over a rank's whole shard (one kernel launch).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log2-samples 32] [--stops 5] [--extra]
  python bench.py --impl reference ...      # the CPU oracle (restatement of the reference) on host cores
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "curve samples/sec"
UNIT = "samples/s"
BYTES_PER_SAMPLE = 16  # [f32;4] store only: t is generated on the device


def u01_f32(n, seed):
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = idx + np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def gradient_builder(stops: int):
    from enterpolation_b200 import BSpline
    el = u01_f32(stops * 4, 0xE1).reshape(stops, 4)
    return BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).normalized().dynamic()


def workload_name(log2n, stops):
    return (f"C5 dynamic clamped equidistant cubic BSpline [f32;4], {stops} stops, take(2^{log2n}) "
            f"sharded over ranks by contiguous global sample range")


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(builder, n_total, first, count, threads):
    """Times the CPU oracle (restatement of enterpolation 0.3.0, see oracle/) on `count` samples."""
    from oracle import oracle
    ref = oracle.OracleCurve(builder.to_desc())
    t0 = time.perf_counter()
    ref.take(n_total, first, count, threads=threads)
    dt = time.perf_counter() - t0
    return count / dt


def run_reference(args):
    """--impl reference: the reference's CPU path (C++ restatement; the Rust crate cannot be built here:
    no cargo/rustc in the image) with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    b = gradient_builder(args.stops)
    n_total = 1 << args.log2_samples
    count = min(n_total, 1 << 24)
    first = (n_total - count) // 2
    from oracle import oracle
    ref = oracle.OracleCurve(b.to_desc())
    for _ in range(args.warmup):
        ref.take(n_total, first, count, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.take(n_total, first, count, threads=cores)
    dt = time.perf_counter() - t0
    value = count * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.log2_samples, args.stops),
                   "reference": "C++ restatement of enterpolation 0.3.0 eval (oracle/; Rust toolchain unavailable)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{count} consecutive samples from the middle of the 2^{args.log2_samples} take, per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from enterpolation_b200 import _abi, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    lib = _abi.load_library()
    builder = gradient_builder(args.stops)
    curve = builder.on_device(local).build()
    curve.upload(local)
    n_total = 1 << args.log2_samples
    lo, hi = shard_range(n_total, rank, world)
    count = hi - lo
    out = torch.empty((count, 4), dtype=torch.float32, device=f"cuda:{local}")
    stream = torch.cuda.current_stream(local)

    def step():
        curve.take_device(n_total, lo, count, out=out, device=local)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.enterp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.enterp_launch_count() - launches0
    clocks = sampler.finish()
    tms = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    tl = torch.tensor([launches], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
    ms_max = float(tms.item())
    value = n_total * args.steps / (ms_max * 1e-3)

    # ---- e2e: the user-facing call with HOST buffers (build + take(n).collect() of a window) -------
    e2e_count_total = min(n_total, 1 << args.log2_e2e)
    e_lo, e_hi = shard_range(e2e_count_total, rank, world)
    e_first = (n_total - e2e_count_total) // 2 + e_lo
    e_count = e_hi - e_lo
    host = torch.empty((e_count, 4), dtype=torch.float32, pin_memory=True)
    table_bytes = args.stops * 16 + (args.stops + 2) * 4

    def e2e_step():
        c = builder.on_device(local).build()  # builder validation + table upload inside the timed region
        st = lib.enterp_take_host(c._h, local, n_total, e_first, e_count, host.data_ptr())
        assert st == 0, st
        return float(host[-1, 0])  # the device->host result is read

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tdt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
    e2e_value = e2e_count_total * e2e_steps / float(tdt.item())

    if rank == 0:
        peaks, how = measured_peaks()
        per_launch_ms = ms / args.steps  # rank 0's kernel: one launch per step
        achieved = count * BYTES_PER_SAMPLE / (per_launch_ms * 1e-3) / 1e9
        peak = float(peaks.get("hbm_gbs", 6650.0))
        cores = os.cpu_count() or 1
        cpu_n1 = min(n_total, 1 << 22)
        cpu_nall = min(n_total, 1 << 25)
        mid = (n_total - cpu_nall) // 2
        cpu_1t = cpu_reference_rate(builder, n_total, mid, cpu_n1, 1)
        cpu_all = cpu_reference_rate(builder, n_total, mid, cpu_nall, cores)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.log2_samples, args.stops),
                       "samples_total": n_total, "samples_per_rank": count, "elements": f"[f32;4] x {args.stops}",
                       "l2_policy": "outputs (16 B/sample, %.1f GiB per rank) are larger than L2; nothing is re-read"
                                    % (count * 16 / 2 ** 30),
                       "parity": "bit-exact vs CPU oracle (tests/test_gpu_parity.py::test_config_shapes_small)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": table_bytes,
                    "d2h_bytes_per_step": e_count * 16,
                    "what": f"builder.build() + enterp_take_host over a {e2e_count_total}-sample window of the "
                            f"2^{args.log2_samples} take (pinned host buffer, chunked kernel/D2H pipeline), "
                            f"{e2e_steps} steps, wall clock max over ranks"},
            "gpu_launches": int(tl.item()),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "bspline_kernel<float,4,false,3,true>",
                         "algorithmic_bytes_per_sample": BYTES_PER_SAMPLE, "peak_source": f"MEASURED_PEAKS.json ({how}, burst copy)"},
            "cpu_baseline": {"value": cpu_all, "unit": UNIT, "cores": cores, "kind": "port",
                             "single_thread": cpu_1t,
                             "sample": f"{cpu_nall} consecutive samples (all cores) / {cpu_n1} (1 thread) from the middle of the take; "
                                       "C++ restatement of enterpolation 0.3.0 (oracle/), per-eval heap workspace as DynSpace"},
        }
        print(json.dumps(line), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-samples", type=int, default=32)
    ap.add_argument("--log2-e2e", type=int, default=28)
    ap.add_argument("--stops", type=int, default=5)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
