"""TEST INFRASTRUCTURE — ctypes wrapper over oracle/libenterp_oracle.so (the CPU restatement of
enterpolation 0.3.0's eval path, oracle/enterp_oracle.hpp). Parity: PINNED by tests/test_oracle_kat.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import desc as _abi  # the checker's own view of the descriptor structs: the product package is not imported

_HERE = os.path.dirname(os.path.abspath(__file__))
# ENTERP_ORACLE_SO: another build of the same sources (tests/test_oracle_sanitized.py points it at the ASan/UBSan build)
_SO = os.environ.get("ENTERP_ORACLE_SO") or os.path.join(_HERE, "libenterp_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("enterp_oracle_capi.cpp", "enterp_oracle.hpp", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "enterp.h"))
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    vp, u64 = C.c_void_p, C.c_uint64
    L.ora_create.argtypes = [vp, C.POINTER(vp), vp]  # any ctypes view of enterp_curve_desc
    L.ora_create.restype = C.c_int
    L.ora_destroy.argtypes = [vp]
    L.ora_destroy.restype = None
    L.ora_bezier_derivatives.argtypes = [vp, vp, u64, u64, C.c_int, vp]
    L.ora_bezier_derivatives.restype = C.c_int64
    L.ora_clamp.argtypes = [vp]
    L.ora_clamp.restype = vp
    L.ora_slice.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_double]
    L.ora_slice.restype = vp
    L.ora_domain.argtypes = [vp, C.POINTER(C.c_double * 2)]
    L.ora_domain.restype = None
    L.ora_info.argtypes = [vp, vp]
    L.ora_bench_take.argtypes = [vp, u64, u64]
    L.ora_bench_take.restype = C.c_double
    L.ora_info.restype = None
    L.ora_eval.argtypes = [vp, vp, u64, vp]
    L.ora_eval.restype = u64
    L.ora_take.argtypes = [vp, u64, u64, u64, vp]
    L.ora_take.restype = u64
    L.ora_span.argtypes = [vp, vp, u64, vp, vp]
    L.ora_span.restype = u64
    L.ora_span_take.argtypes = [vp, u64, u64, u64, vp, vp]
    L.ora_span_take.restype = u64
    L.ora_take_mt.argtypes = [vp, u64, u64, u64, vp, C.c_int]
    L.ora_take_mt.restype = u64
    L.ora_eval_mt.argtypes = [vp, vp, u64, vp, C.c_int]
    L.ora_eval_mt.restype = u64
    L.ora_stepper.argtypes = [C.c_int, C.c_double, C.c_double, u64, u64, u64, vp]
    L.ora_stepper.restype = None
    L.ora_bezier_many_take.argtypes = [C.c_int, C.c_int, u64, C.c_uint32, vp, C.c_uint32, vp, C.c_int]
    L.ora_bezier_many_take.restype = C.c_int
    L.ora_chain_op.argtypes = [C.c_int, C.c_int, u64, C.POINTER(C.c_double), u64, C.c_double, C.c_double,
                               C.c_int, C.c_double, u64, u64, C.POINTER(u64), C.POINTER(C.c_double)]
    L.ora_chain_op.restype = C.c_int
    L.ora_equidistant_step.argtypes = [C.c_int, u64, C.c_double, C.c_double]
    L.ora_equidistant_step.restype = C.c_double
    _lib = L
    return L


class OracleError(Exception):
    def __init__(self, info: _abi.ErrorInfo):
        self.status = info.status
        self.found, self.necessary, self.index, self.mode = info.found, info.necessary, info.index, info.mode
        super().__init__(f"oracle builder error status={info.status} found={info.found} "
                         f"necessary={info.necessary} index={info.index} mode={info.mode}")


class OracleCurve:
    """A curve built by the oracle from the same descriptor the product consumes."""

    def __init__(self, da, _handle=None, _base=None):
        self._da = da
        self._base = _base  # adaptors keep the wrapped curve alive
        if _handle is not None:
            self._h = C.c_void_p(_handle)
            self.scalar, self.dim = _base.scalar, _base.dim
            self.dtype = _abi.NP_DTYPE[self.scalar]
            return
        h = C.c_void_p()
        info = _abi.ErrorInfo()
        st = lib().ora_create(C.addressof(da.desc), C.byref(h), C.byref(info))
        if st != 0:
            info.status = st
            raise OracleError(info)
        self._h = h
        self.scalar = da.desc.scalar
        self.dim = da.desc.dim
        self.dtype = _abi.NP_DTYPE[self.scalar]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.ora_destroy(h)
            self._h = None

    def clamp(self) -> "OracleCurve":
        """Curve::clamp() (signal.rs:271-290)."""
        return OracleCurve(None, _handle=lib().ora_clamp(self._h), _base=self)

    def slice(self, start=None, end=None) -> "OracleCurve":
        """Curve::slice(start..end) (signal.rs:241-262); None = unbounded."""
        h = lib().ora_slice(self._h, start is not None, float(start or 0.0), end is not None, float(end or 0.0))
        return OracleCurve(None, _handle=h, _base=self)

    def domain(self):
        d = (C.c_double * 2)()
        lib().ora_domain(self._h, C.byref(d))
        return [self.dtype(d[0]), self.dtype(d[1])]

    def info(self) -> _abi.CurveInfo:
        i = _abi.CurveInfo()
        lib().ora_info(self._h, C.byref(i))
        return i

    def _out(self, n):
        return np.empty((n, self.dim), dtype=self.dtype)

    def _shape(self, out):
        return out.reshape(-1) if self.dim == 1 else out

    def eval(self, t, threads: int = 1, return_invalid: bool = False, out=None):
        t = np.ascontiguousarray(np.asarray(t, dtype=self.dtype)).reshape(-1)
        out = self._out(t.shape[0]) if out is None else out
        inv = lib().ora_eval_mt(self._h, t.ctypes.data, t.shape[0], out.ctypes.data, threads)
        return (self._shape(out), inv) if return_invalid else self._shape(out)

    def take(self, n_total, first=0, count=None, threads: int = 1, return_invalid: bool = False, out=None):
        count = n_total - first if count is None else count
        out = self._out(count) if out is None else out
        inv = lib().ora_take_mt(self._h, n_total, first, count, out.ctypes.data, threads)
        return (self._shape(out), inv) if return_invalid else self._shape(out)

    def bench_take(self, n_total, reps) -> float:
        """Seconds for `reps` single-threaded `take(n_total).collect()` into fresh vectors (benches/benches.rs:62-69)."""
        return float(lib().ora_bench_take(self._h, n_total, reps))

    def gen_with_tangent(self, t):
        """Bezier::gen_with_tangent (bezier/mod.rs:270-276) for each t -> [n][2][dim]"""
        return self._deriv(t, 2, True)

    def gen_with_derivatives(self, t, k):
        """Bezier::gen_with_deriatives::<K> (bezier/mod.rs:279-285) for each t -> [n][k][dim]"""
        return self._deriv(t, k, False)

    def _deriv(self, t, k, tangent):
        t = np.ascontiguousarray(np.asarray(t, dtype=self.dtype)).reshape(-1)
        out = np.empty((t.shape[0], k, self.dim), dtype=self.dtype)
        rc = lib().ora_bezier_derivatives(self._h, t.ctypes.data, t.shape[0], k, int(tangent), out.ctypes.data)
        if rc < 0:
            raise TypeError("derivatives exist for plain Bezier curves only")
        return out

    def span(self, t):
        t = np.ascontiguousarray(np.asarray(t, dtype=self.dtype)).reshape(-1)
        idx = np.empty(t.shape[0], dtype=np.uint32)
        fac = np.empty(t.shape[0], dtype=self.dtype)
        lib().ora_span(self._h, t.ctypes.data, t.shape[0], idx.ctypes.data, fac.ctypes.data)
        return idx, fac

    def span_take(self, n_total, first=0, count=None):
        count = n_total - first if count is None else count
        idx = np.empty(count, dtype=np.uint32)
        fac = np.empty(count, dtype=self.dtype)
        lib().ora_span_take(self._h, n_total, first, count, idx.ctypes.data, fac.ctypes.data)
        return idx, fac


def stepper(scalar, start, end, n_total, first=0, count=None):
    count = n_total - first if count is None else count
    out = np.empty(count, dtype=_abi.NP_DTYPE[scalar])
    lib().ora_stepper(scalar, float(start), float(end), n_total, first, count, out.ctypes.data)
    return out


def bezier_many_take(ctrl: np.ndarray, samples: int, threads: int = 1) -> np.ndarray:
    """ctrl [n_curves][n_ctrl][dim] -> [n_curves][samples][dim]"""
    ctrl = np.ascontiguousarray(ctrl)
    scalar = _abi.F32 if ctrl.dtype == np.float32 else _abi.F64
    n_curves, n_ctrl, dim = ctrl.shape
    out = np.empty((n_curves, samples, dim), dtype=ctrl.dtype)
    st = lib().ora_bezier_many_take(scalar, dim, n_curves, n_ctrl, ctrl.ctypes.data, samples, out.ctypes.data, threads)
    assert st == 0
    return out


# knot-chain hooks for the list.rs / adaptors.rs KATs ------------------------------------------
SORTED, EQUIDISTANT, CONST_EQUIDISTANT = 0, 1, 2
NO_ADAPTOR, BORDER_BUFFER, BORDER_DELETION = 0, 1, 2


class Chain:
    def __init__(self, chain, values=None, length=0, step=0.0, offset=0.0, adaptor=NO_ADAPTOR, border_n=0):
        self.chain, self.adaptor, self.border_n = chain, adaptor, border_n
        self.values = None if values is None else np.ascontiguousarray(np.asarray(values, dtype=np.float64))
        self.length = length if values is None else len(self.values)
        self.step, self.offset = step, offset

    @staticmethod
    def sorted(values, **kw):
        return Chain(SORTED, values=values, **kw)

    @staticmethod
    def equidistant_normalized(n, **kw):
        return Chain(EQUIDISTANT, length=n, step=lib().ora_equidistant_step(_abi.EQ_NORMALIZED, n, 0, 0), offset=0.0, **kw)

    @staticmethod
    def equidistant(n, start, end, **kw):
        return Chain(EQUIDISTANT, length=n, step=lib().ora_equidistant_step(_abi.EQ_DOMAIN, n, start, end), offset=start, **kw)

    @staticmethod
    def const_equidistant(n, **kw):
        return Chain(CONST_EQUIDISTANT, length=n, **kw)

    def _op(self, op, x=0.0, lo=0, hi=0):
        idx = (C.c_uint64 * 2)()
        val = (C.c_double * 1)()
        vals = None if self.values is None else self.values.ctypes.data_as(C.POINTER(C.c_double))
        rc = lib().ora_chain_op(self.chain, self.adaptor, self.border_n, vals, self.length, self.step, self.offset,
                                op, float(x), lo, hi, idx, val)
        if rc != 0:
            raise RuntimeError("reference would panic")
        return idx, val

    def strict_upper_bound(self, x):
        return int(self._op(0, x)[0][0])

    def strict_upper_bound_clamped(self, x, lo, hi):
        return int(self._op(1, x, lo, hi)[0][0])

    def upper_border(self, x):
        idx, val = self._op(2, x)
        return int(idx[0]), int(idx[1]), float(val[0])

    def eval(self, i):
        return float(self._op(3, 0.0, i)[1][0])

    def __len__(self):
        return int(self._op(4)[0][0])
