"""The Rust host crates under rust/ are source (no Rust toolchain in the image): keep the raw binding in lock-step
with the C ABI. rust/enterpolation-b200-sys/src/lib.rs must equal what tools/gen_rust_sys.py generates from
include/enterp.h, bind every exported symbol with the header's arity, and mirror every struct field for field."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import gen_rust_sys  # noqa: E402

from enterpolation_b200 import _abi  # noqa: E402

LIB_RS = os.path.join(ROOT, "rust", "enterpolation-b200-sys", "src", "lib.rs")


def test_committed_binding_is_the_generated_one():
    assert open(LIB_RS).read() == gen_rust_sys.generate(), "run `python tools/gen_rust_sys.py` after editing include/enterp.h"


def test_extern_block_binds_every_symbol_with_the_headers_arity():
    rs = open(LIB_RS).read()
    block = rs[rs.index('extern "C" {'):]
    bound = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (enterp_\w+)\(([^)]*)\)", block)}
    assert sorted(bound) == sorted(_abi.EXPORTS)
    header = gen_rust_sys.strip_comments(open(gen_rust_sys.HEADER).read())
    for name, args in re.findall(r"ENTERP_API\s+[^;(]*?\b(enterp_\w+)\s*\(([^)]*)\)\s*;", header, flags=re.S):
        n_c = 0 if args.strip() in ("", "void") else len(args.split(","))
        n_rs = 0 if not bound[name].strip() else len(bound[name].split(","))
        assert n_c == n_rs, name


def test_structs_mirror_the_ctypes_view_field_for_field():
    import ctypes as C
    rs = open(LIB_RS).read()
    to_rust = {C.c_int32: "i32", C.c_uint64: "u64", C.c_double: "f64", C.c_void_p: "*const c_void", C.c_double * 2: "[f64; 2]"}
    for rust_name, cls in (("enterp_curve_desc", _abi.CurveDesc), ("enterp_error_info", _abi.ErrorInfo),
                           ("enterp_curve_info", _abi.CurveInfo), ("enterp_curve_resolved", _abi.CurveResolved)):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % rust_name, rs, flags=re.S).group(1)
        fields = re.findall(r"pub (\w+): ([^,]+),", body)
        assert [f[0] for f in fields] == [f[0] for f in cls._fields_], rust_name
        for (fname, ftype), (_, ctype) in zip(fields, cls._fields_):
            assert ftype == to_rust[ctype], (rust_name, fname, ftype)


def test_status_codes_match():
    rs = open(LIB_RS).read()
    codes = dict(re.findall(r"pub const (ENTERP_\w+): enterp_status = (\d+);", rs))
    for name in ("OK", "TOO_FEW_ELEMENTS", "TOO_FEW_KNOTS", "KNOT_ELEMENT_INEQUALITY", "NOT_SORTED", "INCONGRUOUS_ELEMENTS_KNOTS",
                 "INCONGRUOUS_ELEMENTS_DEGREE", "INVALID_DEGREE", "TOO_SMALL_WORKSPACE", "EMPTY", "INVALID_ARGUMENT", "UNSUPPORTED",
                 "NO_DEVICE", "CUDA_ERROR", "NOT_UPLOADED"):
        assert int(codes["ENTERP_" + name]) == getattr(_abi, name), name


def test_safe_wrapper_only_calls_bound_symbols():
    """rust/enterpolation-b200/src/lib.rs: every `sys::enterp_*` it names exists in the -sys crate."""
    rs = open(LIB_RS).read()
    wrapper = open(os.path.join(ROOT, "rust", "enterpolation-b200", "src", "lib.rs")).read()
    for sym in set(re.findall(r"sys::(enterp_\w+)\b", wrapper)):
        assert re.search(r"\b%s\b" % sym, rs), sym
    for const in set(re.findall(r"sys::(ENTERP_\w+)\b", wrapper)):
        assert re.search(r"pub const %s\b" % const, rs), const
