"""Error taxonomy of the reference's builders, as Python exceptions.

LinearError  src/linear/error.rs:13-20   BezierError  src/bezier/error.rs:12-17
BSplineError src/bspline/error.rs:16-31  leaf structs src/builder.rs:85-190, src/base/list.rs:325-334
"""
from __future__ import annotations

from . import _abi


class EnterpolationError(Exception):
    """Base of every builder error; carries the payload of the reference's error struct."""

    status = -1

    def __init__(self, message="", *, found=0, necessary=0, index=0, mode=0):
        super().__init__(message)
        self.found, self.necessary, self.index, self.mode = found, necessary, index, mode


class LinearError(EnterpolationError):
    pass


class BezierError(EnterpolationError):
    pass


class BSplineError(EnterpolationError):
    pass


class TooFewElements(LinearError, BSplineError):
    status = _abi.TOO_FEW_ELEMENTS


class TooFewKnots(BSplineError):
    status = _abi.TOO_FEW_KNOTS


class KnotElementInequality(LinearError):
    status = _abi.KNOT_ELEMENT_INEQUALITY


class NotSorted(LinearError, BSplineError):
    status = _abi.NOT_SORTED


class IncongruousElementsKnots(BSplineError):
    status = _abi.INCONGRUOUS_ELEMENTS_KNOTS


class IncongruousElementsDegree(BSplineError):
    status = _abi.INCONGRUOUS_ELEMENTS_DEGREE


class InvalidDegree(BSplineError):
    status = _abi.INVALID_DEGREE


class TooSmallWorkspace(BezierError, BSplineError):
    status = _abi.TOO_SMALL_WORKSPACE


class Empty(BezierError):
    status = _abi.EMPTY


class InvalidArgument(EnterpolationError, ValueError):
    status = _abi.INVALID_ARGUMENT


class Unsupported(EnterpolationError, NotImplementedError):
    status = _abi.UNSUPPORTED


class DeviceError(RuntimeError):
    """ENTERP_NO_DEVICE / ENTERP_CUDA_ERROR: raised by every eval entry point; there is no CPU path."""


_BY_STATUS = {c.status: c for c in (
    TooFewElements, TooFewKnots, KnotElementInequality, NotSorted, IncongruousElementsKnots,
    IncongruousElementsDegree, InvalidDegree, TooSmallWorkspace, Empty, InvalidArgument, Unsupported)}

_MESSAGES = {
    _abi.TOO_FEW_ELEMENTS: "To few elements given for the interpolation. {found} elements were given, but at least 2 are necessary.",
    _abi.TOO_FEW_KNOTS: "To few knots given for the interpolation. {found} knots were given, but at least 2 are necessary.",
    _abi.KNOT_ELEMENT_INEQUALITY: "There has to be as many knots as elements, however we found {found} elements and {necessary} knots.",
    _abi.NOT_SORTED: "Given knots are not sorted. From index {index} to {index1} we found decreasing values.",
    _abi.INCONGRUOUS_ELEMENTS_KNOTS: "Found {found} elements and {necessary} knots, which do not fit a bspline of build mode {mode}.",
    _abi.INCONGRUOUS_ELEMENTS_DEGREE: "Found {found} elements and a degree of {necessary}, which do not fit a bspline of build mode {mode}.",
    _abi.INVALID_DEGREE: "The degree of the resulting curve is {found} and such not valid.",
    _abi.TOO_SMALL_WORKSPACE: "The given workspace is too small with space for {found} elements, at least {necessary} have to fit.",
    _abi.EMPTY: "No elements given, an empty generator is not allowed.",
    _abi.INVALID_ARGUMENT: "malformed curve description",
    _abi.UNSUPPORTED: "the reference has no builder path for this combination",
}


def from_info(info: "_abi.ErrorInfo") -> EnterpolationError:
    cls = _BY_STATUS.get(info.status, EnterpolationError)
    msg = _MESSAGES.get(info.status, f"status {info.status}").format(
        found=info.found, necessary=info.necessary, index=info.index, index1=info.index + 1, mode=info.mode)
    return cls(msg, found=info.found, necessary=info.necessary, index=info.index, mode=info.mode)
