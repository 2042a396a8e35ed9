"""Host-side mirror of the reference's `Signal`/`Curve` surface over the C ABI.

reference: src/base/signal.rs:15-19 (eval), 48-58 (extract), 166-173 (sample), 194 (domain),
221-228 (take); Take/Stepper 527-645.

Every evaluation goes through libenterp_b200.so (hand-written sm_100a kernels). There is no CPU
implementation in this package: without the built library or without a CUDA device the calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .errors import DeviceError, from_info


def _check(lib, status: int):
    if status == _abi.OK:
        return
    if status in (_abi.NO_DEVICE, _abi.CUDA_ERROR):
        raise DeviceError(f"{lib.enterp_status_str(status).decode()}: {lib.enterp_last_cuda_error().decode()}")
    if status == _abi.NOT_UPLOADED:
        raise DeviceError(lib.enterp_status_str(status).decode())
    info = _abi.ErrorInfo()
    info.status = status
    raise from_info(info)


def _stream_ptr(stream):
    if stream is None:
        return None
    return getattr(stream, "cuda_stream", stream)


class Curve:
    """A built curve: immutable, uploaded once per device, evaluated on the GPU.

    Mirrors `impl Signal<R> + Curve<R>` of Linear / Bezier / BSpline / Weighted<..>.
    """

    def __init__(self, desc_arrays: _abi.DescArrays | None, device: int = 0, _handle=None):
        self._lib = _abi.load_library()
        self._da = desc_arrays
        if _handle is not None:
            h = _handle
        else:
            h = C.c_void_p()
            info = _abi.ErrorInfo()
            st = self._lib.enterp_curve_create(C.byref(desc_arrays.desc), C.byref(h), C.byref(info))
            if st != _abi.OK:
                info.status = st
                raise from_info(info)
        self._h = h
        ci = _abi.CurveInfo()
        _check(self._lib, self._lib.enterp_curve_get_info(h, C.byref(ci)))
        self.info = ci
        self.scalar = ci.scalar
        self.dim = ci.dim
        self.dtype = _abi.NP_DTYPE[ci.scalar]
        self.device = device
        self._da = None  # the library copied everything it needs

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.enterp_curve_destroy(h)
            self._h = None

    # ---- Curve<R> ----------------------------------------------------------------------------
    def domain(self):
        """`Curve::domain()` -> [start, end] in R."""
        d = (C.c_double * 2)()
        _check(self._lib, self._lib.enterp_curve_domain(self._h, C.byref(d)))
        return [self.dtype(d[0]), self.dtype(d[1])]

    def resolved(self) -> _abi.CurveResolved:
        """Constants the reference's built types carry as fields (Equidistant step/offset, TransformInput
        addition/multiplication, Plateau min/max, workspace length) — see enterpolation_b200.serde."""
        r = _abi.CurveResolved()
        _check(self._lib, self._lib.enterp_curve_get_resolved(self._h, C.byref(r)))
        return r

    @property
    def degree(self) -> int:
        return int(self.info.degree)

    def clamp(self) -> "Curve":
        """`Curve::clamp()` (signal.rs:271-290 -> Clamp, adaptors.rs:14-44): parameters are clamped to the
        domain before evaluation. The new curve shares this one's device tables."""
        h = C.c_void_p()
        _check(self._lib, self._lib.enterp_curve_clamp(self._h, C.byref(h)))
        return Curve(None, device=self.device, _handle=h)

    def slice(self, start=None, end=None) -> "Curve":
        """`Curve::slice(start..end)` (signal.rs:241-262 -> Slice, adaptors.rs:55-106); None = unbounded."""
        h = C.c_void_p()
        _check(self._lib, self._lib.enterp_curve_slice(self._h, start is not None, float(start or 0.0),
                                                       end is not None, float(end or 0.0), C.byref(h)))
        return Curve(None, device=self.device, _handle=h)

    def take(self, samples: int) -> "Take":
        """`Curve::take(samples)`: `samples` equidistant parameters over the domain (signal.rs:221-228)."""
        return Take(self, int(samples))

    # ---- Signal<R> ---------------------------------------------------------------------------
    def eval(self, t):
        """`Signal::eval` for one parameter (a 1-sample batch through the GPU path)."""
        out = self.sample([t])
        return out[0]

    def sample(self, ts, device: int | None = None, devices=None):
        """`Signal::sample(iter).collect()`: host parameters in, host elements out (numpy). `devices=[..]` spreads
        the batch over several GPUs of the box with one call (enterp_sample_host_multi)."""
        dev = self.device if device is None else device
        t = np.ascontiguousarray(np.asarray(ts, dtype=self.dtype)).reshape(-1)
        out = np.empty((t.shape[0], self.dim), dtype=self.dtype)
        if t.shape[0]:
            if devices is not None:
                arr = (C.c_int * len(devices))(*[int(d) for d in devices])
                st = self._lib.enterp_sample_host_multi(self._h, arr, len(devices), t.ctypes.data, t.shape[0], out.ctypes.data)
            else:
                st = self._lib.enterp_sample_host(self._h, dev, t.ctypes.data, t.shape[0], out.ctypes.data)
            _check(self._lib, st)
        return out.reshape(-1) if self.dim == 1 else out

    extract = sample

    # ---- device-resident batches (torch tensors are only memory handles here) ------------------
    def _torch_dtype(self):
        import torch
        return torch.float32 if self.scalar == _abi.F32 else torch.float64

    def _new_out(self, n, device):
        import torch
        return torch.empty((n, self.dim) if self.dim > 1 else (n,), dtype=self._torch_dtype(), device=f"cuda:{device}")

    def upload(self, device: int | None = None):
        dev = self.device if device is None else device
        _check(self._lib, self._lib.enterp_curve_upload(self._h, dev))
        return self

    def sample_device(self, t, out=None, stream=None):
        """out[i] = eval(t[i]) for a CUDA tensor of parameters; returns a CUDA tensor [n][dim]."""
        import torch
        assert t.is_cuda and t.dtype == self._torch_dtype() and t.is_contiguous()
        dev = t.device.index
        n = t.numel()
        if out is None:
            out = self._new_out(n, dev)
        s = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        _check(self._lib, self._lib.enterp_eval_batch(self._h, dev, t.data_ptr(), n, out.data_ptr(), s))
        return out

    def take_device(self, n_total, first=0, count=None, out=None, device=None, stream=None):
        """Global samples [first, first+count) of take(n_total), generated and evaluated on the device."""
        import torch
        dev = self.device if device is None else device
        count = n_total - first if count is None else count
        if out is None:
            out = self._new_out(count, dev)
        s = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        _check(self._lib, self._lib.enterp_eval_take(self._h, dev, n_total, first, count, out.data_ptr(), s))
        return out

    def span_device(self, t, stream=None):
        """Knot-span decision only (parity/debug): (idx uint32 as int32 tensor, factor)."""
        import torch
        dev = t.device.index
        n = t.numel()
        idx = torch.empty(n, dtype=torch.int32, device=t.device)
        fac = torch.empty(n, dtype=self._torch_dtype(), device=t.device)
        s = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        _check(self._lib, self._lib.enterp_span_batch(self._h, dev, t.data_ptr(), n, idx.data_ptr(), fac.data_ptr(), s))
        return idx, fac

    def span_take_device(self, n_total, first=0, count=None, device=None, stream=None):
        import torch
        dev = self.device if device is None else device
        count = n_total - first if count is None else count
        idx = torch.empty(count, dtype=torch.int32, device=f"cuda:{dev}")
        fac = torch.empty(count, dtype=self._torch_dtype(), device=f"cuda:{dev}")
        s = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        _check(self._lib, self._lib.enterp_span_take(self._h, dev, n_total, first, count, idx.data_ptr(), fac.data_ptr(), s))
        return idx, fac

    def gen_with_tangent(self, t):
        """`Bezier::gen_with_tangent` (bezier/mod.rs:270-276) for every parameter of `t` (CUDA tensor or
        array-like): returns [n][2][dim] = (value, tangent)."""
        return self._derivatives(t, 2, True)

    def gen_with_derivatives(self, t, k: int):
        """`Bezier::gen_with_deriatives::<K>` (bezier/mod.rs:279-285): [n][k][dim]."""
        return self._derivatives(t, int(k), False)

    def _derivatives(self, t, k, tangent):
        import torch
        host = not (hasattr(t, "is_cuda") and t.is_cuda)
        tt = torch.as_tensor(np.ascontiguousarray(np.asarray(t, dtype=self.dtype)).reshape(-1)).cuda(self.device) if host else t
        dev = tt.device.index
        n = tt.numel()
        out = torch.empty((n, k, self.dim), dtype=self._torch_dtype(), device=tt.device)
        s = torch.cuda.current_stream(dev).cuda_stream
        if tangent:
            st = self._lib.enterp_bezier_tangent_batch(self._h, dev, tt.data_ptr(), n, out.data_ptr(), s)
        else:
            st = self._lib.enterp_bezier_derivatives_batch(self._h, dev, tt.data_ptr(), n, k, out.data_ptr(), s)
        _check(self._lib, st)
        return out.cpu().numpy() if host else out

    def invalid_count(self, device=None, stream=None) -> int:
        """How many samples so far would have made the reference panic (written as NaN). With `stream` only that
        stream is synchronised (enterp_invalid_count_stream); otherwise the whole device."""
        dev = self.device if device is None else device
        n = C.c_ulonglong(0)
        if stream is not None:
            _check(self._lib, self._lib.enterp_invalid_count_stream(self._h, dev, _stream_ptr(stream), C.byref(n)))
        else:
            _check(self._lib, self._lib.enterp_invalid_count(self._h, dev, C.byref(n)))
        return int(n.value)

    def invalid_reset(self, device=None, stream=None):
        """Zero the counter in stream order, so that reset -> eval -> invalid_count(stream=..) belongs to one call."""
        dev = self.device if device is None else device
        _check(self._lib, self._lib.enterp_invalid_reset(self._h, dev, _stream_ptr(stream)))


class Take:
    """`Take<C, R>` (signal.rs:527-576): an exact-size iterator over `samples` elements.

    `collect()` evaluates the requested global range on the GPU in one call; iteration yields rows.
    `shard(rank, world)` restricts to the contiguous global range a rank owns (the stepper still
    uses the GLOBAL index and the global sample count, so shards agree bit for bit).
    """

    def __init__(self, curve: Curve, samples: int, first: int = 0, count: int | None = None):
        if samples < 1:
            raise ValueError("take(0) underflows in the reference (signal.rs:218-220)")
        self.curve, self.samples = curve, samples
        self.first = first
        self.count = samples - first if count is None else count

    def __len__(self):
        return self.count

    def shard(self, rank: int, world: int) -> "Take":
        from .sharding import shard_range
        lo, hi = shard_range(self.samples, rank, world)
        return Take(self.curve, self.samples, lo, hi - lo)

    def collect(self, device: int | None = None, devices=None) -> np.ndarray:
        """`take(n).collect()`; `devices=[..]`: one call sharded over several GPUs (enterp_take_host_multi)."""
        c = self.curve
        dev = c.device if device is None else device
        out = np.empty((self.count, c.dim), dtype=c.dtype)
        if self.count:
            if devices is not None:
                arr = (C.c_int * len(devices))(*[int(d) for d in devices])
                st = c._lib.enterp_take_host_multi(c._h, arr, len(devices), self.samples, self.first, self.count, out.ctypes.data)
            else:
                st = c._lib.enterp_take_host(c._h, dev, self.samples, self.first, self.count, out.ctypes.data)
            _check(c._lib, st)
        return out.reshape(-1) if c.dim == 1 else out

    def collect_device(self, device: int | None = None, out=None, stream=None):
        return self.curve.take_device(self.samples, self.first, self.count, out=out, device=device, stream=stream)

    def __iter__(self):
        return iter(self.collect())


def stepper_device(scalar, start, end, n_total, first=0, count=None, device=0):
    """`Stepper::new(n_total, start, end)` values [first, first+count) as a CUDA tensor."""
    import torch
    lib = _abi.load_library()
    count = n_total - first if count is None else count
    out = torch.empty(count, dtype=torch.float32 if scalar == _abi.F32 else torch.float64, device=f"cuda:{device}")
    s = torch.cuda.current_stream(device).cuda_stream
    _check(lib, lib.enterp_stepper(scalar, device, float(start), float(end), n_total, first, count, out.data_ptr(), s))
    return out


def bezier_many_take(ctrl, samples: int, out=None, stream=None):
    """Many independent Bezier curves: ctrl CUDA tensor [n_curves][n_ctrl][dim] -> [n_curves][samples][dim]."""
    import torch
    lib = _abi.load_library()
    assert ctrl.is_cuda and ctrl.is_contiguous() and ctrl.dim() == 3
    scalar = _abi.F32 if ctrl.dtype == torch.float32 else _abi.F64
    n_curves, n_ctrl, dim = ctrl.shape
    dev = ctrl.device.index
    if out is None:
        out = torch.empty((n_curves, samples, dim), dtype=ctrl.dtype, device=ctrl.device)
    s = _stream_ptr(stream) if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    _check(lib, lib.enterp_bezier_many_take(scalar, dim, n_curves, n_ctrl, ctrl.data_ptr(), samples, dev, out.data_ptr(), s))
    return out
