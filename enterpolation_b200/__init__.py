"""enterpolation_b200 — B200-native batched curve evaluation behind enterpolation's builder API.

    from enterpolation_b200 import BSpline, Linear, Bezier
    spline = (BSpline.builder().clamped().elements([0, 0, 0, 6, 0, 0, 0])
              .equidistant("f64").degree(3).domain(-2, 2).constant(4).build())
    values = spline.take(1 << 24).collect()

The package is a thin host mirror of the reference's `Signal`/`Curve` surface; all arithmetic runs
in hand-written sm_100a kernels inside libenterp_b200.so (see include/enterp.h). No CPU fallback.
"""
from . import _abi, easing
from .builders import Bezier, BezierBuilder, BSpline, BSplineBuilder, Linear, LinearBuilder
from .curve import Curve, Take, bezier_many_take, stepper_device
from .errors import (BezierError, BSplineError, DeviceError, Empty, EnterpolationError,
                     IncongruousElementsDegree, IncongruousElementsKnots, InvalidArgument,
                     InvalidDegree, KnotElementInequality, LinearError, NotSorted, TooFewElements,
                     TooFewKnots, TooSmallWorkspace, Unsupported)
from .sharding import shard_range, shard_sizes
from . import serde  # noqa: E402  (imports the builders)

__all__ = [
    "BSpline", "Linear", "Bezier", "BSplineBuilder", "LinearBuilder", "BezierBuilder", "Curve", "Take",
    "bezier_many_take", "stepper_device", "serde", "easing", "shard_range", "shard_sizes", "EnterpolationError",
    "LinearError", "BezierError", "BSplineError", "TooFewElements", "TooFewKnots",
    "KnotElementInequality", "NotSorted", "IncongruousElementsKnots", "IncongruousElementsDegree",
    "InvalidDegree", "TooSmallWorkspace", "Empty", "InvalidArgument", "Unsupported", "DeviceError",
]
__version__ = "0.1.0"
