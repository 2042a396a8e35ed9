"""ctypes view of include/enterp.h (the C ABI) — structures, enums and the library loader.

Nothing here computes: it only describes the boundary. The loader FAILS LOUDLY when the CUDA
extension has not been built (there is no CPU fallback in the product path).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

# enum mirrors of include/enterp.h ------------------------------------------------------------
OK = 0
TOO_FEW_ELEMENTS = 1
TOO_FEW_KNOTS = 2
KNOT_ELEMENT_INEQUALITY = 3
NOT_SORTED = 4
INCONGRUOUS_ELEMENTS_KNOTS = 5
INCONGRUOUS_ELEMENTS_DEGREE = 6
INVALID_DEGREE = 7
TOO_SMALL_WORKSPACE = 8
EMPTY = 9
INVALID_ARGUMENT = 10
UNSUPPORTED = 11
NO_DEVICE = 12
CUDA_ERROR = 13
NOT_UPLOADED = 14

LINEAR, BEZIER, BSPLINE = 0, 1, 2
F32, F64 = 0, 1
OPEN, CLAMPED, LEGACY = 0, 1, 2
KNOTS_SORTED, KNOTS_EQUIDISTANT, KNOTS_CONST_EQUIDISTANT = 0, 1, 2
EQ_BY_DEGREE, EQ_BY_QUANTITY = 0, 1
EQ_DOMAIN, EQ_NORMALIZED, EQ_DISTANCE = 0, 1, 2
SPACE_CONSTANT, SPACE_DYNAMIC, SPACE_WORKSPACE = 0, 1, 2
INPUT_NORMALIZED, INPUT_DOMAIN, INPUT_TRANSFORM = 0, 1, 2
EASE_IDENTITY, EASE_FLIP, EASE_SMOOTHSTART, EASE_SMOOTHEND, EASE_SMOOTHSTEP, EASE_SMOOTHERSTEP, EASE_PLATEAU = range(7)

NP_DTYPE = {F32: np.float32, F64: np.float64}


class CurveDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("scalar", C.c_int32), ("dim", C.c_int32), ("weighted", C.c_int32),
        ("n_elements", C.c_uint64), ("elements", C.c_void_p), ("weights", C.c_void_p),
        ("mode", C.c_int32), ("knot_base", C.c_int32),
        ("n_knots", C.c_uint64), ("knots", C.c_void_p),
        ("eq_by", C.c_int32), ("eq_form", C.c_int32),
        ("eq_count", C.c_uint64), ("eq_a", C.c_double), ("eq_b", C.c_double),
        ("space", C.c_int32), ("input_form", C.c_int32),
        ("workspace_n", C.c_uint64), ("in_a", C.c_double), ("in_b", C.c_double),
        ("easing", C.c_int32), ("easing_n", C.c_int32), ("easing_param", C.c_double),
    ]


class ErrorInfo(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("mode", C.c_int32),
        ("found", C.c_uint64), ("necessary", C.c_uint64), ("index", C.c_uint64),
    ]


class CurveInfo(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("scalar", C.c_int32), ("dim", C.c_int32), ("weighted", C.c_int32),
        ("knot_base", C.c_int32), ("adaptor", C.c_int32),
        ("n_elements", C.c_uint64), ("n_knots", C.c_uint64), ("degree", C.c_uint64),
        ("border", C.c_uint64), ("domain", C.c_double * 2),
    ]


class CurveResolved(C.Structure):
    _fields_ = [
        ("eq_len", C.c_uint64), ("eq_step", C.c_double), ("eq_offset", C.c_double),
        ("has_transform", C.c_int32), ("easing", C.c_int32),
        ("in_addition", C.c_double), ("in_multiplication", C.c_double),
        ("space_len", C.c_uint64), ("ease_min", C.c_double), ("ease_max", C.c_double),
    ]


# every symbol include/enterp.h declares (tests check the built library exports all of them)
EXPORTS = [
    "enterp_curve_validate", "enterp_curve_create", "enterp_curve_destroy", "enterp_curve_get_info", "enterp_curve_get_resolved", "enterp_curve_domain",
    "enterp_curve_clamp", "enterp_curve_slice", "enterp_curve_upload", "enterp_eval_batch", "enterp_eval_take", "enterp_span_batch",
    "enterp_span_take", "enterp_stepper", "enterp_bezier_many_take", "enterp_bezier_tangent_batch", "enterp_bezier_derivatives_batch", "enterp_sample_host",
    "enterp_take_host", "enterp_sample_host_multi", "enterp_take_host_multi", "enterp_host_scratch_release", "enterp_host_transfer_bytes",
    "enterp_host_alloc", "enterp_host_free", "enterp_invalid_count", "enterp_invalid_count_stream", "enterp_invalid_reset",
    "enterp_status_str", "enterp_last_cuda_error", "enterp_device_count", "enterp_launch_count",
    "enterp_version",
]

LIB_NAME = "libenterp_b200.so"
_lib = None


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


def load_library() -> C.CDLL:
    """Load the in-tree CUDA library. Raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(enterpolation_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    lib.enterp_curve_create.argtypes = [C.POINTER(CurveDesc), C.POINTER(vp), C.POINTER(ErrorInfo)]
    lib.enterp_curve_create.restype = i32
    lib.enterp_curve_validate.argtypes = [C.POINTER(CurveDesc), i32, C.POINTER(ErrorInfo)]
    lib.enterp_curve_validate.restype = i32
    lib.enterp_curve_destroy.argtypes = [vp]
    lib.enterp_curve_destroy.restype = None
    lib.enterp_curve_get_info.argtypes = [vp, C.POINTER(CurveInfo)]
    lib.enterp_curve_get_info.restype = i32
    lib.enterp_curve_get_resolved.argtypes = [vp, C.POINTER(CurveResolved)]
    lib.enterp_curve_get_resolved.restype = i32
    lib.enterp_curve_domain.argtypes = [vp, C.POINTER(C.c_double * 2)]
    lib.enterp_curve_domain.restype = i32
    lib.enterp_curve_clamp.argtypes = [vp, C.POINTER(vp)]
    lib.enterp_curve_clamp.restype = i32
    lib.enterp_curve_slice.argtypes = [vp, i32, C.c_double, i32, C.c_double, C.POINTER(vp)]
    lib.enterp_curve_slice.restype = i32
    lib.enterp_curve_upload.argtypes = [vp, i32]
    lib.enterp_curve_upload.restype = i32
    lib.enterp_eval_batch.argtypes = [vp, i32, vp, u64, vp, vp]
    lib.enterp_eval_batch.restype = i32
    lib.enterp_eval_take.argtypes = [vp, i32, u64, u64, u64, vp, vp]
    lib.enterp_eval_take.restype = i32
    lib.enterp_span_batch.argtypes = [vp, i32, vp, u64, vp, vp, vp]
    lib.enterp_span_batch.restype = i32
    lib.enterp_span_take.argtypes = [vp, i32, u64, u64, u64, vp, vp, vp]
    lib.enterp_span_take.restype = i32
    lib.enterp_stepper.argtypes = [i32, i32, C.c_double, C.c_double, u64, u64, u64, vp, vp]
    lib.enterp_stepper.restype = i32
    lib.enterp_bezier_many_take.argtypes = [i32, i32, u64, C.c_uint32, vp, C.c_uint32, i32, vp, vp]
    lib.enterp_bezier_many_take.restype = i32
    lib.enterp_bezier_tangent_batch.argtypes = [vp, i32, vp, u64, vp, vp]
    lib.enterp_bezier_tangent_batch.restype = i32
    lib.enterp_bezier_derivatives_batch.argtypes = [vp, i32, vp, u64, C.c_uint32, vp, vp]
    lib.enterp_bezier_derivatives_batch.restype = i32
    lib.enterp_sample_host.argtypes = [vp, i32, vp, u64, vp]
    lib.enterp_sample_host.restype = i32
    lib.enterp_take_host.argtypes = [vp, i32, u64, u64, u64, vp]
    lib.enterp_take_host.restype = i32
    lib.enterp_sample_host_multi.argtypes = [vp, C.POINTER(C.c_int), i32, vp, u64, vp]
    lib.enterp_sample_host_multi.restype = i32
    lib.enterp_take_host_multi.argtypes = [vp, C.POINTER(C.c_int), i32, u64, u64, u64, vp]
    lib.enterp_take_host_multi.restype = i32
    lib.enterp_host_transfer_bytes.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.enterp_host_transfer_bytes.restype = None
    lib.enterp_host_scratch_release.argtypes = [i32]
    lib.enterp_host_scratch_release.restype = i32
    lib.enterp_invalid_count_stream.argtypes = [vp, i32, vp, C.POINTER(C.c_ulonglong)]
    lib.enterp_invalid_count_stream.restype = i32
    lib.enterp_invalid_reset.argtypes = [vp, i32, vp]
    lib.enterp_invalid_reset.restype = i32
    lib.enterp_host_alloc.argtypes = [C.POINTER(vp), u64]
    lib.enterp_host_alloc.restype = i32
    lib.enterp_host_free.argtypes = [vp]
    lib.enterp_host_free.restype = None
    lib.enterp_invalid_count.argtypes = [vp, i32, C.POINTER(C.c_ulonglong)]
    lib.enterp_invalid_count.restype = i32
    lib.enterp_status_str.argtypes = [i32]
    lib.enterp_status_str.restype = C.c_char_p
    lib.enterp_last_cuda_error.argtypes = []
    lib.enterp_last_cuda_error.restype = C.c_char_p
    lib.enterp_device_count.argtypes = []
    lib.enterp_device_count.restype = i32
    lib.enterp_launch_count.argtypes = []
    lib.enterp_launch_count.restype = C.c_ulonglong
    lib.enterp_version.argtypes = []
    lib.enterp_version.restype = C.c_char_p
    _lib = lib
    return lib


class DescArrays:
    """Keeps the numpy arrays a CurveDesc points into alive."""

    def __init__(self, desc: CurveDesc, keep):
        self.desc = desc
        self.keep = keep


def make_desc(kind, scalar, elements, weights=None, *, mode=OPEN, knots=None, equidistant=None,
              space=(SPACE_DYNAMIC, 0), bezier_domain=None, easing=(EASE_IDENTITY, 0, 0.0),
              const_equidistant=False, bezier_transform=None) -> DescArrays:
    """Flatten one finished builder chain into the C descriptor.

    elements: array-like [n] or [n][dim]; weights: [n] or None; knots: [k] or None;
    equidistant: (eq_by, eq_count, eq_form, a, b) or None; space: (kind, n).
    """
    dt = NP_DTYPE[scalar]
    el = np.ascontiguousarray(np.asarray(elements, dtype=dt))
    if el.ndim == 1:
        el = el.reshape(-1, 1)
    if el.ndim != 2:
        raise ValueError("elements must be [n] or [n][dim]")
    n, dim = el.shape
    if n == 0:
        dim = max(dim, 1)
    d = CurveDesc()
    keep = [el]
    d.kind, d.scalar, d.dim = kind, scalar, dim
    d.n_elements = n
    d.elements = el.ctypes.data if n else None
    if weights is not None:
        w = np.ascontiguousarray(np.asarray(weights, dtype=dt)).reshape(-1)
        if w.shape[0] != n:
            raise ValueError("one weight per element")
        keep.append(w)
        d.weighted = 1
        d.weights = w.ctypes.data if n else None
    d.mode = mode
    if knots is not None:
        k = np.ascontiguousarray(np.asarray(knots, dtype=dt)).reshape(-1)
        keep.append(k)
        d.knot_base = KNOTS_SORTED
        d.n_knots = k.shape[0]
        d.knots = k.ctypes.data if k.shape[0] else None
    elif equidistant is not None:
        d.knot_base = KNOTS_EQUIDISTANT
        d.eq_by, d.eq_count, d.eq_form, a, b = equidistant
        d.eq_a, d.eq_b = float(dt(a)), float(dt(b))
    elif const_equidistant:
        d.knot_base = KNOTS_CONST_EQUIDISTANT
        if const_equidistant is not True:  # BSpline over ConstEquidistant<R, N>: N is given (Linear: N = n_elements)
            d.n_knots = int(const_equidistant)
    else:
        d.knot_base = KNOTS_SORTED if kind != BEZIER else 0
    d.space, d.workspace_n = space
    d.easing, d.easing_n, d.easing_param = int(easing[0]), int(easing[1]), float(dt(easing[2]))
    if bezier_domain is not None:
        d.input_form = INPUT_DOMAIN
        d.in_a, d.in_b = float(dt(bezier_domain[0])), float(dt(bezier_domain[1]))
    if bezier_transform is not None:  # stored TransformInput { addition, multiplication }
        d.input_form = INPUT_TRANSFORM
        d.in_a, d.in_b = float(dt(bezier_transform[0])), float(dt(bezier_transform[1]))
    return DescArrays(d, keep)
