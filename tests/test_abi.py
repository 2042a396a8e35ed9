"""The C-ABI library loads on a CPU-only box and exports every symbol include/enterp.h declares."""
import ctypes as C
import os
import re

from enterpolation_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "enterp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"ENTERP_API[^;(]*?\b(enterp_\w+)\s*\(", text)))


def test_header_and_loader_agree():
    assert declared_symbols() == sorted(_abi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    for name in declared_symbols():
        assert getattr(lib, name) is not None, name


def test_desc_layout_matches_header():
    # field order/types are the ABI; sizes are what a C compiler produces for the header's struct
    assert C.sizeof(_abi.CurveDesc) == 144
    assert C.sizeof(_abi.ErrorInfo) == 32
    assert C.sizeof(_abi.CurveInfo) == 72


def test_no_device_is_loud():
    """Without a GPU every eval entry point fails with a status, never a CPU fallback."""
    import numpy as np
    import pytest
    from enterpolation_b200 import DeviceError, Linear
    lib = _abi.load_library()
    if lib.enterp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    lin = Linear.builder().elements([0.0, 3.0]).knots([0.0, 1.0]).build()
    assert lin.domain() == [0.0, 1.0]  # host-side metadata needs no device
    with pytest.raises(DeviceError):
        lin.sample(np.array([0.5]))
    with pytest.raises(DeviceError):
        lin.take(3).collect()
    assert lib.enterp_launch_count() == 0


def test_status_strings():
    lib = _abi.load_library()
    assert lib.enterp_status_str(_abi.OK) == b"ok"
    assert lib.enterp_status_str(_abi.NOT_SORTED) == b"NotSorted"
    assert b"sm_100a" in lib.enterp_version()


def test_oracle_descriptor_view_matches_the_products():
    """oracle/desc.py restates the descriptor structs so that the checker never imports the product package."""
    from oracle import desc
    for name in ("CurveDesc", "ErrorInfo", "CurveInfo"):
        a, b = getattr(desc, name), getattr(_abi, name)
        assert C.sizeof(a) == C.sizeof(b), name
        assert [(f[0], f[1]) for f in a._fields_] == [(f[0], f[1]) for f in b._fields_], name
    for const in ("LINEAR", "BEZIER", "BSPLINE", "F32", "F64", "OPEN", "CLAMPED", "LEGACY", "KNOTS_SORTED", "KNOTS_EQUIDISTANT",
                  "EQ_BY_DEGREE", "EQ_BY_QUANTITY", "EQ_DOMAIN", "EQ_NORMALIZED", "EQ_DISTANCE", "SPACE_CONSTANT", "SPACE_DYNAMIC"):
        assert getattr(desc, const) == getattr(_abi, const), const


def test_reference_arm_does_not_load_the_product():
    import subprocess
    import sys
    code = ("import sys; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','1','--log2-samples','20'];"
            "import bench; bench.main(); assert 'enterpolation_b200' not in sys.modules; "
            "assert not any('libenterp_b200' in l for l in open('/proc/self/maps'))")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert '"impl": "reference"' in r.stdout
