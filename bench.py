#!/usr/bin/env python
"""bench.py — curve samples/s of the hot path on N B200s (contract in the task prompt, ④).

Headline workload (BASELINE.json configs[4], the one north_star's target is quoted on):
  C5  dynamic clamped equidistant cubic BSpline over [f32;4] colour stops (examples/gradient.rs
      shape, 5 stops), `take(2^32)`; the GLOBAL sample range is sharded contiguously over the N
      ranks (strong scaling: total work fixed; no collective on the data path).
A step = one pass of the hot path over a rank's whole shard (one kernel launch).
The other BASELINE.json configs (C1-C4) are timed briefly as well and reported under "configs" in
the same JSON line (skip with --no-extra).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log2-samples 32] [--stops 5] [--no-extra]
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

# FP-pipe issue peaks, lane-operations per second, both stated: MEASURED on this pool's B200 with
# tools/pipebench.cu (profiles/r01_pipebench.json: dependency-free unfused FMUL/FADD chains, whole chip, clocks
# settle at ~1.82 GHz under that load) and THEORETICAL = 148 SMs x 128 (64) lanes x 1.965 GHz (max SM clock).
FP32_LANE_OPS = 3.45e13
FP64_LANE_OPS = 1.80e13
FP32_LANE_OPS_THEORETICAL = 148 * 128 * 1.965e9
FP64_LANE_OPS_THEORETICAL = 148 * 64 * 1.965e9
# ncu --set full of the stepper-run kernel bspline_rows_kernel<float,4,0,3,1,1,4> on a 2^30-sample launch
# (profiles/r02_ncu_c5_runs.txt): (dram__bytes_write.sum + dram__bytes_read.sum) of both launches of a step / samples
NCU_C5_DRAM_BYTES_PER_SAMPLE = (16.851780e9 + 92.9e3 + 0.208069e9 + 11.0e3) / (1 << 30)  # runs kernel + 2^24-sample head

# ---- deterministic synthetic inputs: splitmix64(seed + index) ----------------------
def splitmix64(idx, seed):
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def u01(n, seed, dtype, offset=0):
    z = splitmix64(np.arange(offset, offset + n, dtype=np.uint64), seed)
    if dtype == np.float32:
        return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def u01_torch(n, seed, dtype, device, offset=0):
    """Same generator on the GPU (int64 arithmetic wraps like uint64; logical shifts are masked)."""
    import torch

    def lsr(z, k):
        return (z >> k) & ((1 << (64 - k)) - 1)

    def c(v):  # uint64 constant as a wrapped int64
        return v - (1 << 64) if v >= (1 << 63) else v

    out = torch.empty(n, dtype=dtype, device=device)
    chunk = 1 << 26
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        z = torch.arange(offset + lo, offset + lo + m, dtype=torch.int64, device=device) + c(seed) + c(0x9E3779B97F4A7C15)
        z = (z ^ lsr(z, 30)) * c(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * c(0x94D049BB133111EB)
        z = z ^ lsr(z, 31)
        if dtype == torch.float32:
            out[lo:lo + m] = lsr(z, 40).to(torch.float32) * (2.0 ** -24)
        else:
            out[lo:lo + m] = lsr(z, 11).to(torch.float64) * (2.0 ** -53)
    return out


def best_rate(fn, n, reps=3):
    """samples/s of fn(): one untimed warm-up, then the best of `reps` timed runs."""
    fn()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return n / best


# ---- workloads ------------------------------------------------------------------------------------
class Workload:
    """One BASELINE.json config, sharded for (rank, world). Subclasses build device state in setup()."""
    key = ""
    name = ""
    dtype = "f32"
    kernel = ""
    bytes_per_sample = 0      # algorithmic HBM bytes per sample
    fp_ops_per_sample = 0     # exact-order FP operations per sample, div = 1
    fp_peak = FP32_LANE_OPS

    def __init__(self, rank=0, world=1, device=0):
        self.rank, self.world, self.device = rank, world, device
        self.samples = 0         # samples this rank evaluates per step
        self.samples_total = 0   # whole job

    def setup(self):
        raise NotImplementedError

    def step(self):
        raise NotImplementedError

    def cpu_rate(self, threads, n):
        """samples/s of the CPU oracle on a bounded sample of this workload."""
        raise NotImplementedError


def gradient_builder(stops: int):
    from enterpolation_b200 import BSpline
    el = gradient_elements(stops)
    return BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).normalized().dynamic()


class TakeWorkload(Workload):
    """take(n_total) of one curve, global sample range sharded contiguously (C1, C5)."""

    def __init__(self, builder, log2n, dim, torch_dtype, **kw):
        super().__init__(**kw)
        self.builder, self.log2n, self.dim, self.torch_dtype = builder, log2n, dim, torch_dtype

    def setup(self):
        import torch
        from enterpolation_b200 import shard_range
        self.n_total = 1 << self.log2n
        self.samples_total = self.n_total
        self.lo, hi = shard_range(self.n_total, self.rank, self.world)
        self.samples = hi - self.lo
        self.curve = self.builder.on_device(self.device).build()
        self.curve.upload(self.device)
        shape = (self.samples, self.dim) if self.dim > 1 else (self.samples,)
        self.out = torch.empty(shape, dtype=self.torch_dtype, device=f"cuda:{self.device}")
        return self

    def step(self):
        self.curve.take_device(self.n_total, self.lo, self.samples, out=self.out, device=self.device)

    def cpu_rate(self, threads, n):
        from oracle import oracle
        ref = oracle.OracleCurve(self.builder.to_desc())
        n = min(n, self.n_total)
        first = (self.n_total - n) // 2
        out = np.zeros((n, self.dim), dtype=ref.dtype)  # pre-touched: no first-touch page faults inside the timing
        return best_rate(lambda: ref.take(self.n_total, first, n, threads=threads, out=out), n)


class C5(TakeWorkload):
    key = "C5"
    # VARIANT 4 = stepper runs (rows.cuh RUNS): beyond index 2^24 an f32 take repeats its parameter in runs; every
    # distinct parameter is evaluated once and the samples are streamed out of shared memory. The 2^24-sample head
    # runs the per-sample variant 2 (5 stops) / 3 (more stops).
    kernel = "bspline_rows_kernel<float,4,false,3,true,POW2,4>"
    bytes_per_sample = 16
    fp_ops_per_sample = 125

    def __init__(self, log2n=32, stops=5, **kw):
        import torch
        super().__init__(gradient_builder(stops), log2n, 4, torch.float32, **kw)
        self.stops = stops
        self.name = c5_name(log2n, stops)


class C1(TakeWorkload):
    key = "C1"
    dtype = "f64"
    kernel = "bspline_rows_kernel<double,1,false,3,true,false>"
    bytes_per_sample = 8
    fp_ops_per_sample = 71
    fp_peak = FP64_LANE_OPS

    def __init__(self, log2n=24, **kw):
        import torch
        from enterpolation_b200 import BSpline
        b = (BSpline.builder().clamped().elements([0.0, 0.0, 0.0, 6.0, 0.0, 0.0, 0.0]).equidistant("f64").degree(3)
             .domain(-2.0, 2.0).constant(4))
        super().__init__(b, log2n, 1, torch.float64, **kw)
        self.name = f"C1 README cardinal cubic clamped BSpline, 7 f64 elements, equidistant knots on [-2,2], take(2^{log2n})"


class BatchWorkload(Workload):
    """sample(t) over a device-resident batch of parameters, batch sharded contiguously (C2, C4)."""

    def make(self):
        """-> (builder, n_total, dim, np dtype, torch dtype, seed_t)"""
        raise NotImplementedError

    def setup(self):
        import torch
        from enterpolation_b200 import shard_range
        self.builder, self.n_total, self.dim, npdt, tdt = self.make()
        self.samples_total = self.n_total
        lo, hi = shard_range(self.n_total, self.rank, self.world)
        self.lo, self.samples = lo, hi - lo
        self.curve = self.builder.on_device(self.device).build()
        self.curve.upload(self.device)
        d = self.curve.domain()
        dev = f"cuda:{self.device}"
        u = u01_torch(self.samples, 0xE4, tdt, dev, offset=lo)
        self.t = (float(d[0]) + u * (float(d[1]) - float(d[0]))).to(tdt).contiguous()
        self.out = torch.empty((self.samples, self.dim), dtype=tdt, device=dev)
        self.npdt = npdt
        return self

    def step(self):
        self.curve.sample_device(self.t, out=self.out)

    def cpu_rate(self, threads, n):
        from oracle import oracle
        ref = oracle.OracleCurve(self.builder.to_desc())
        n = min(n, self.samples)
        t = self.t[:n].cpu().numpy()
        out = np.zeros((n, self.dim), dtype=ref.dtype)
        return best_rate(lambda: ref.eval(t, threads=threads, out=out), n)


class C2(BatchWorkload):
    key = "C2"
    kernel = "linear_kernel<float,3,false,false>"
    bytes_per_sample = 16
    fp_ops_per_sample = 13

    def __init__(self, log2n=28, log2_knots=20, **kw):
        super().__init__(**kw)
        self.log2n, self.log2k = log2n, log2_knots
        self.name = (f"C2 Linear [f32;3], 2^{log2_knots} sorted non-uniform knots, 2^{log2n} random sample parameters "
                     f"(device resident)")

    def make(self):
        import torch
        from enterpolation_b200 import Linear
        k = 1 << self.log2k
        inc = 0.5 + u01(k - 1, 0xE2, np.float64)
        knots = np.concatenate([[0.0], np.cumsum(inc)]).astype(np.float32)
        el = (u01(k * 3, 0xE1, np.float32) * np.float32(2) - np.float32(1)).reshape(k, 3)
        return Linear.builder().elements(el).knots(knots), 1 << self.log2n, 3, np.float32, torch.float32


class C4(BatchWorkload):
    key = "C4"
    dtype = "f64"
    kernel = "bspline_kernel<double,4,true,3,false>"
    bytes_per_sample = 32
    fp_ops_per_sample = 111
    fp_peak = FP64_LANE_OPS

    def __init__(self, log2n=26, log2_ctrl=16, **kw):
        super().__init__(**kw)
        self.log2n, self.log2c = log2n, log2_ctrl
        self.name = (f"C4 NURBS (Weighted BSpline) degree 3 [f64;3] + weights, 2^{log2_ctrl} control points, "
                     f"non-uniform knots, 2^{log2n} random sample parameters (device resident)")

    def make(self):
        import torch
        from enterpolation_b200 import BSpline
        e = 1 << self.log2c
        pts = (u01(e * 3, 0xE1, np.float64) * 2.0 - 1.0).reshape(e, 3)
        w = 0.5 + u01(e, 0xE3, np.float64)
        inner = np.cumsum(0.5 + u01(e - 3, 0xE2, np.float64))  # e + 2 knots: 3 x 0, e - 4 interior, last x 3
        knots = np.concatenate([[0.0, 0.0, 0.0], inner[:-1], [inner[-1]] * 3])
        assert knots.shape[0] == e + 2
        b = BSpline.builder().elements_with_weights(pts, w).knots(knots).constant(4)
        return b, 1 << self.log2n, 3, np.float64, torch.float64


class C3(Workload):
    key = "C3"
    kernel = "bezier_many_kernel<float,3,16>"
    bytes_per_sample = 12.75
    fp_ops_per_sample = 1081

    def __init__(self, log2_curves=20, n_ctrl=16, samples=256, **kw):
        super().__init__(**kw)
        self.log2c, self.n_ctrl, self.per_curve = log2_curves, n_ctrl, samples
        self.name = (f"C3 Bezier degree {n_ctrl - 1} [f32;3], 2^{log2_curves} independent curves x take({samples}) "
                     f"(control points device resident)")

    def setup(self):
        import torch
        from enterpolation_b200 import shard_range
        n_curves = 1 << self.log2c
        lo, hi = shard_range(n_curves, self.rank, self.world)
        self.curves = hi - lo
        self.samples = self.curves * self.per_curve
        self.samples_total = n_curves * self.per_curve
        dev = f"cuda:{self.device}"
        per = self.n_ctrl * 3
        u = u01_torch(self.curves * per, 0xE1, torch.float32, dev, offset=lo * per)
        self.ctrl = (u * 2.0 - 1.0).reshape(self.curves, self.n_ctrl, 3).contiguous()
        self.out = torch.empty((self.curves, self.per_curve, 3), dtype=torch.float32, device=dev)
        return self

    def step(self):
        from enterpolation_b200 import bezier_many_take
        bezier_many_take(self.ctrl, self.per_curve, out=self.out)

    def cpu_rate(self, threads, n):
        from oracle import oracle
        curves = max(threads, min(self.curves, n // self.per_curve))
        ctrl = self.ctrl[:curves].cpu().numpy()
        return best_rate(lambda: oracle.bezier_many_take(ctrl, self.per_curve, threads=threads), curves * self.per_curve)


def make_workload(case, log2n=None, arg=None, rank=0, world=1, device=0):
    case = case.lower()
    kw = dict(rank=rank, world=world, device=device)
    if case == "c5":
        return C5(log2n or 32, arg or 5, **kw).setup()
    if case == "c1":
        return C1(log2n or 24, **kw).setup()
    if case == "c2":
        return C2(log2n or 28, arg or 20, **kw).setup()
    if case == "c3":
        return C3(log2n or 20, arg or 16, **kw).setup()
    if case == "c4":
        return C4(log2n or 26, arg or 16, **kw).setup()
    raise ValueError(case)


# ---- clocks -----------------------------------------------------------------------------------------
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


# ---- shared by both arms -----------------------------------------------------------------------------
def c5_name(log2n, stops):
    return (f"C5 dynamic clamped equidistant cubic BSpline [f32;4], {stops} stops, take(2^{log2n}) "
            f"sharded over ranks by contiguous global sample range")


def headline_config(args):
    """The `config` object of the JSON line: identical in both arms (--impl b200 / reference)."""
    return {"workload": c5_name(args.log2_samples, args.stops), "samples_total": 1 << args.log2_samples,
            "elements": f"[f32;4] x {args.stops}", "n_gpus": args.gpus,
            "l2_policy": "outputs (16 B/sample, 64 GiB / n_gpus per rank at 2^32) are larger than L2; nothing is re-read",
            "parity": "bit-exact vs CPU oracle (tests/test_gpu_parity.py, tests/test_gpu_fullsize.py)"}


def gradient_elements(stops):
    return u01(stops * 4, 0xE1, np.float32).reshape(stops, 4)


def oracle_gradient(stops):
    """The C5 curve built for the CPU oracle WITHOUT importing the product package (oracle/desc.py)."""
    from oracle import desc, oracle
    d = desc.bspline(gradient_elements(stops), desc.F32, mode=desc.CLAMPED,
                     equidistant=(desc.EQ_BY_DEGREE, 3, desc.EQ_NORMALIZED, 0.0, 0.0), space=(desc.SPACE_DYNAMIC, 0))
    return oracle.OracleCurve(d)


README_ELEMENTS = [
    943.0, 978.0, 579.0, 15.0, 608.0, 938.0, 669.0, 98.0, 720.0, 303.0, 345.0, 421.0, 767.0, 798.0, 379.0, 379.0, 794.0,
    774.0, 559.0, 315.0, 141.0, 465.0, 721.0, 398.0, 265.0, 921.0, 664.0, 701.0, 433.0, 494.0, 286.0, 331.0, 976.0, 493.0,
    296.0, 93.0, 963.0, 422.0, 398.0, 358.0, 906.0, 80.0, 272.0, 287.0, 867.0, 995.0, 109.0, 874.0, 840.0, 931.0, 554.0,
    405.0, 643.0, 820.0, 229.0, 736.0, 117.0, 317.0, 105.0, 168.0, 16.0, 267.0, 481.0, 547.0, 308.0, 463.0, 664.0, 948.0,
    655.0, 201.0, 439.0, 910.0, 431.0, 945.0, 441.0, 682.0, 272.0, 716.0, 851.0, 916.0, 98.0, 162.0, 536.0, 129.0, 369.0,
    753.0, 735.0, 511.0, 346.0, 137.0, 221.0, 715.0, 565.0, 305.0, 649.0, 913.0, 215.0, 713.0, 321.0, 95.0]


def readme_anchor():
    """The reference's only published timing (README.md:133, benches/benches.rs:5-27, 62-93): a degree-3 B-spline
    over 100 f64 elements, `take(200).collect()`, in its dynamic / constant / uniform flavours. Timed through the
    CPU oracle, single thread, per-iteration output allocation like criterion's `b.iter::<Vec<f64>, _>`; it anchors
    the CPU denominators of this bench to the crate's own numbers (other hardware, rustc instead of g++ -O2)."""
    from oracle import desc, oracle
    knots = [0.0, 0.0] + [float(i) for i in range(98)] + [97.0, 97.0]
    assert len(knots) == 102 and len(README_ELEMENTS) == 100
    curves = {
        "dynamic": desc.bspline(README_ELEMENTS, desc.F64, knots=knots, space=(desc.SPACE_DYNAMIC, 0)),
        "constant": desc.bspline(README_ELEMENTS, desc.F64, knots=knots, space=(desc.SPACE_CONSTANT, 4)),
        "uniform": desc.bspline(README_ELEMENTS, desc.F64, equidistant=(desc.EQ_BY_DEGREE, 3, desc.EQ_NORMALIZED, 0.0, 0.0),
                                space=(desc.SPACE_CONSTANT, 4)),
    }
    published = {"dynamic": 23.160e3 / 200, "constant": 6.6834e3 / 200, "uniform": 9.9186e3 / 200}
    out = {"shape": "BSpline degree 3, 100 f64 elements, take(200).collect(), single thread (benches/benches.rs:5-27, 62-93)",
           "unit": "ns/sample", "published_source": "README.md:133 (other hardware, rustc)"}
    for name, d in curves.items():
        ref = oracle.OracleCurve(d)
        ref.bench_take(200, 2000)
        best = min(ref.bench_take(200, 20000) for _ in range(3))
        out[name] = {"oracle_ns_per_sample": best / (200 * 20000) * 1e9, "published_ns_per_sample": published[name]}
    return out


def reference_subprocess(args, threads):
    """samples/s of `bench.py --impl reference --cpu-threads T` run as a child process (None on failure)."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
           "--log2-samples", str(args.log2_samples), "--stops", str(args.stops), "--cpu-threads", str(threads)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        return float(json.loads(r.stdout.strip().splitlines()[-1])["value"])
    except Exception:
        return None


# ---- the reference arm --------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU path (C++ restatement in oracle/; the Rust crate cannot be
    built here: no cargo/rustc in the image) with all host threads, bounded sample per step. Does not import the
    product package (the descriptor is built by oracle/desc.py), so no CUDA library is loaded."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = args.cpu_threads or cores
    ref = oracle_gradient(args.stops)
    n_total = 1 << args.log2_samples
    count = min(n_total, 1 << (24 if threads > 1 else 22))
    first = (n_total - count) // 2
    out = np.zeros((count, 4), dtype=np.float32)
    for _ in range(max(args.warmup, 1)):
        ref.take(n_total, first, count, threads=threads, out=out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.take(n_total, first, count, threads=threads, out=out)
    dt = time.perf_counter() - t0
    value = count * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "threads": threads, "kind": "port",
                         "what": "C++ restatement of enterpolation 0.3.0 eval (oracle/; Rust toolchain unavailable), "
                                 "std::thread over contiguous index ranges, rank 0 only",
                         "sample": f"{count} consecutive samples from the middle of the 2^{args.log2_samples} take, per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- our arm ------------------------------------------------------------------------------------------
def time_steps(w, steps, warmup, stream, barrier):
    import torch
    for _ in range(warmup):
        w.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        w.step()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1)


def bind_to_gpu_numa(index: int) -> bool:
    """One process per GPU: run on the cores NVML reports as local to that GPU, so that the pinned host buffers
    of the end-to-end leg are first touched (and therefore placed) on the GPU's own NUMA node. Without it every
    rank's device->host stream converges on whichever node the scheduler picked. Not applied at N = 1, where
    rank 0 also times the all-cores CPU baseline."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return True
    except Exception:
        return False


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from enterpolation_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    numa_bound = world > 1 and bind_to_gpu_numa(local)  # before any allocation: pinned pages land next to the GPU
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce_max(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    lib = _abi.load_library()
    stream = torch.cuda.current_stream(local)
    peaks, how = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    cores = os.cpu_count() or 1
    warmup = max(args.warmup, 3)

    # ---- headline: C5 ---------------------------------------------------------------------------
    w = C5(args.log2_samples, args.stops, rank=rank, world=world, device=local).setup()
    for _ in range(warmup):
        w.step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.enterp_launch_count()
    ms = time_steps(w, args.steps, 0, stream, barrier)
    launches = lib.enterp_launch_count() - launches0
    clocks = sampler.finish()
    ms_max = reduce_max(ms)
    total_launches = int(reduce_sum(launches))
    value = w.samples_total * args.steps / (ms_max * 1e-3)
    n_total = w.n_total
    per_launch_ms = ms / args.steps  # rank 0's step (one launch; rank 0 also launches the 2^24-sample head kernel)
    achieved = w.samples * w.bytes_per_sample / (per_launch_ms * 1e-3) / 1e9
    c5_samples = w.samples
    rows_kernel = w.kernel
    builder = w.builder

    # ---- the same take with 256 stops (SURVEY.md 8d "second run 256 stops": the general, non-power-of-two case) ----
    c5_256 = None
    if args.stops != 256:
        try:
            x = C5(args.log2_samples, 256, rank=rank, world=world, device=local)
            x.builder, x.curve = gradient_builder(256), None
            x.n_total, x.samples_total, x.lo, x.samples = w.n_total, w.samples_total, w.lo, w.samples
            x.curve = x.builder.on_device(local).build()
            x.curve.upload(local)
            x.out = w.out  # same output buffer (the 64 GiB take does not fit twice)
            xsteps = max(3, min(args.steps, 10))
            xms_own = time_steps(x, xsteps, 3, stream, barrier)
            xms = reduce_max(xms_own)
            own_gbs = x.samples * 16 / (xms_own / xsteps * 1e-3) / 1e9
            c5_256 = {"workload": c5_name(args.log2_samples, 256), "value": x.samples_total * xsteps / (xms * 1e-3),
                      "unit": UNIT, "ms_per_step": xms / xsteps, "dtype": "f32", "kernel": x.kernel,
                      "hbm": {"algorithmic_bytes_per_sample": 16, "achieved_gbs": own_gbs, "peak_gbs": hbm_peak,
                              "frac": own_gbs / hbm_peak}}
            del x
        except Exception as ex:
            c5_256 = {"error": repr(ex)}

    # ---- e2e: the user-facing call with HOST buffers (build + take(n).collect() of a window), per-rank window ----
    e_count = min(w.samples, 1 << args.log2_e2e)           # samples per RANK (weak in the window, like the D2H link)
    e2e_total = e_count * world
    e_first = w.lo + (w.samples - e_count) // 2            # the middle of this rank's shard of the take
    del w
    torch.cuda.empty_cache()
    host = torch.empty((e_count, 4), dtype=torch.float32, pin_memory=True)
    host.zero_()  # touch the pinned pages before timing
    table_bytes = args.stops * 16 + (args.stops + 2) * 4

    def e2e_step():
        c = builder.on_device(local).build()  # builder validation + table upload inside the timed region
        st = lib.enterp_take_host(c._h, local, n_total, e_first, e_count, host.data_ptr())
        assert st == 0, st
        return float(host[-1, 0])  # the device->host result is read

    for _ in range(3):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    import ctypes as C
    pcie0_in, pcie0_out = C.c_ulonglong(0), C.c_ulonglong(0)
    lib.enterp_host_transfer_bytes(C.byref(pcie0_in), C.byref(pcie0_out))
    # three blocks of e2e_steps steps, the median block counts: the host side of this path (memory bandwidth of a
    # shared host) is noisy from one second to the next, the device side is not
    block_dts = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        block_dts.append(reduce_max(time.perf_counter() - t0))
    e2e_dt = sorted(block_dts)[1]
    e2e_value = e2e_total * e2e_steps / e2e_dt
    pcie1_in, pcie1_out = C.c_ulonglong(0), C.c_ulonglong(0)
    lib.enterp_host_transfer_bytes(C.byref(pcie1_in), C.byref(pcie1_out))
    pcie_d2h = (pcie1_out.value - pcie0_out.value) // (3 * e2e_steps)  # bytes that actually crossed PCIe per step (rank 0)

    # The limiter, measured in the same run: plain pinned cudaMemcpyAsync D2H of the same bytes, all ranks at once.
    probe_src = torch.empty((e_count, 4), dtype=torch.float32, device=dev)
    probe_src.zero_()
    for _ in range(2):
        host.copy_(probe_src, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host.copy_(probe_src, non_blocking=True)
        torch.cuda.synchronize()
    probe_dt = reduce_max(time.perf_counter() - t0)
    d2h_limit_gbs = e2e_total * 16 * e2e_steps / probe_dt / 1e9
    del host, probe_src

    cpu_1t = cpu_all = anchor = None
    if rank == 0 and world == 1:  # the CPU baseline is an N = 1 leg (at N > 1 the ranks are bound to GPU-local cores)
        # the reference arm itself, in a fresh process (this one holds a CUDA context, torch's thread pools and a
        # clock sampler, which cost the all-cores oracle a third of its rate when it ran in-process): same code path,
        # same sample, so the in-line baseline and `--impl reference` agree by construction
        cpu_1t = reference_subprocess(args, 1)
        cpu_all = reference_subprocess(args, cores)
        try:
            anchor = readme_anchor()
        except Exception as ex:
            anchor = {"error": repr(ex)}
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configs, briefly ---------------------------------------------------
    extra = {}
    if c5_256 is not None:
        extra["C5_256"] = c5_256
    if not args.no_extra:
        for cls in (C1, C2, C3, C4):
            try:
                x = cls(rank=rank, world=world, device=local).setup()
                xsteps = 10
                xms_own = time_steps(x, xsteps, 3, stream, barrier)
                xms = reduce_max(xms_own)
                xval = x.samples_total * xsteps / (xms * 1e-3)
                own = x.samples / (xms_own / xsteps * 1e-3)  # this rank's kernel
                theo = FP64_LANE_OPS_THEORETICAL if x.fp_peak == FP64_LANE_OPS else FP32_LANE_OPS_THEORETICAL
                entry = {
                    "workload": x.name, "value": xval, "unit": UNIT, "ms_per_step": xms / xsteps, "dtype": x.dtype,
                    "kernel": x.kernel,
                    "hbm": {"algorithmic_bytes_per_sample": x.bytes_per_sample,
                            "achieved_gbs": own * x.bytes_per_sample / 1e9, "peak_gbs": hbm_peak,
                            "frac": own * x.bytes_per_sample / 1e9 / hbm_peak},
                    "fp_pipe": {"exact_order_fp_ops_per_sample": x.fp_ops_per_sample,
                                "achieved_lane_ops_per_s": own * x.fp_ops_per_sample,
                                "peak_lane_ops_per_s": x.fp_peak, "frac": own * x.fp_ops_per_sample / x.fp_peak,
                                "peak_source": "measured: tools/pipebench.cu unfused FMUL/FADD (DMUL/DADD) chains, "
                                               "profiles/r01_pipebench.json (~1.82 GHz under load)",
                                "peak_theoretical_lane_ops_per_s": theo, "frac_of_theoretical": own * x.fp_ops_per_sample / theo,
                                "peak_theoretical_source": "148 SMs x 128 (f64: 64) lanes x 1.965 GHz"},
                }
                if rank == 0 and world == 1:
                    entry["cpu_baseline"] = {"single_thread": x.cpu_rate(1, 1 << 20), "all_cores": x.cpu_rate(cores, 1 << 23),
                                             "cores": cores, "unit": UNIT, "kind": "port"}
                extra[x.key] = entry
                del x
                torch.cuda.empty_cache()
            except Exception as ex:  # an extra must never take the headline down
                extra[cls.key] = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config(args),
            "samples_per_rank": c5_samples,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": table_bytes,
                    "d2h_bytes_per_step": pcie_d2h, "host_output_bytes_per_step": e_count * 16,
                    "transfer": "run-length: an f32 take repeats its parameter in runs of 2^(floor(log2 i)-23) samples beyond "
                                "index 2^24 (256 in this window); only the distinct values are evaluated and copied over PCIe "
                                "(enterp_host_transfer_bytes), the library replicates them into the pinned output with "
                                "host threads (byte copies, no curve arithmetic on the CPU); ENTERP_NO_RLE=1 copies every sample",
                    "window": "weak: %d samples per rank from the middle of each rank's shard" % e_count,
                    "achieved_gbs": e2e_value * 16 / 1e9, "limit_gbs": d2h_limit_gbs,
                    "frac_of_limit": e2e_value * 16 / 1e9 / d2h_limit_gbs,
                    "limit_source": "plain pinned cudaMemcpyAsync D2H of the OUTPUT bytes (16 B/sample) on every rank at once, same run: "
                                    "what a transfer of every sample is bound by; frac_of_limit > 1 = the run-length transfer beats it",
                    "what": f"builder.build() + enterp_take_host over a {e_count}-sample window per rank of the "
                            f"2^{args.log2_samples} take (pinned host buffer, chunked kernel/D2H pipeline), "
                            f"3 blocks of {e2e_steps} steps, median block, wall clock max over ranks",
                    "block_values": [e2e_total * e2e_steps / d for d in block_dts]},
            "gpu_launches": total_launches, "numa_bound": bool(numa_bound),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         # ncu --set full of this kernel on a 2^30-sample launch: DRAM bytes per sample, scaled to this
                         # launch's sample count
                         "traffic": c5_samples * NCU_C5_DRAM_BYTES_PER_SAMPLE / 1e9, "traffic_unit": "GB per launch",
                         "traffic_source": "profiles/r02_ncu_c5_runs.txt (per-sample DRAM bytes of a 2^30-sample "
                                           "launch x this launch's samples)",
                         "kernel": rows_kernel,
                         "algorithmic_bytes_per_sample": 16, "peak_source": f"MEASURED_PEAKS.json ({how}, burst copy)",
                         # FP work: one de Boor evaluation per DISTINCT stepper value (run length 2^(floor(log2 i)-23) for
                         # i >= 2^24: 256 in the upper half of a take(2^32)); the per-sample kernels of round 1 needed 93
                         # FMA-pipe lane-ops per sample (5 stops), the reference's own order 125
                         "evaluations_per_sample": (float(1 << 24) + 8 * float(1 << 23)) / float(1 << 32)
                         if args.log2_samples == 32 else None,
                         "reference_order_fp_ops_per_sample": 125},
            "cpu_baseline": {"value": cpu_all, "unit": UNIT, "cores": cores, "kind": "port",
                             "single_thread": cpu_1t,
                             "sample": f"{1 << 24} consecutive samples (all cores) / {1 << 22} (1 thread) from the middle of the take; "
                                       "`bench.py --impl reference [--cpu-threads 1]` in a child process: pre-touched output, "
                                       "1 warm-up, mean of 3 steps; "
                                       "C++ restatement of enterpolation 0.3.0 (oracle/), per-eval heap workspace as DynSpace",
                             "readme_anchor": anchor}
            if world == 1 else None,
            "configs": extra,
        }
        emit(line)
    barrier()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    # libraries (NCCL prints its version banner) must not pollute the one-line stdout contract
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-samples", type=int, default=32)
    ap.add_argument("--log2-e2e", type=int, default=28)
    ap.add_argument("--stops", type=int, default=5)
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0, help="reference arm: host threads (0 = all cores)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
