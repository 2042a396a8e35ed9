"""The checker must itself be clean where the reference panics: the oracle's C++ restatement, rebuilt with
AddressSanitizer + UBSan (oracle/Makefile `san`), evaluates the fuzz generators' curves — descending and zero-width
knot vectors, infinite knots, non-finite parameters — in a child process; any report aborts it. (Round 2: an index
below the lower cut of the span search made the oracle read in front of its element array, and the CUDA path fault.)"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(HERE, "..", "oracle")


def _runtime(name):
    for cxx in (os.environ.get("SAN_CXX"), "/usr/bin/g++", "g++"):
        if not cxx:
            continue
        try:
            path = subprocess.run([cxx, f"-print-file-name={name}"], capture_output=True, text=True, check=True).stdout.strip()
        except (OSError, subprocess.CalledProcessError):
            continue
        if os.path.isabs(path) and os.path.exists(path):
            return os.path.realpath(path)
    return None


def test_oracle_is_clean_under_asan_and_ubsan():
    asan = _runtime("libasan.so")
    if asan is None:
        pytest.skip("no AddressSanitizer runtime in this image")
    made = subprocess.run(["make", "-C", ORACLE, "san"], capture_output=True, text=True)
    if made.returncode != 0:
        pytest.skip("sanitizer build of the oracle failed: " + made.stderr[-300:])
    env = dict(os.environ)
    env["ENTERP_ORACLE_SO"] = os.path.join(ORACLE, "_san", "libenterp_oracle_san.so")
    # libstdc++ next to it: ASan's __cxa_throw interceptor needs the real one resolvable when it initialises (the python
    # binary itself does not link libstdc++)
    env["LD_PRELOAD"] = asan + " " + (_runtime("libstdc++.so.6") or "")
    env["ASAN_OPTIONS"] = "detect_leaks=0:abort_on_error=1"
    env["UBSAN_OPTIONS"] = "halt_on_error=1:print_stacktrace=1"
    run = subprocess.run([sys.executable, os.path.join(HERE, "oracle_sanitized_driver.py"), "0", "30"], env=env, capture_output=True,
                         text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    assert "no sanitizer report" in run.stdout


def test_host_resolution_is_clean_under_asan_and_ubsan(tmp_path):
    """The product's host side (descriptor validation and table construction, csrc/resolve.cpp) is plain C++: the same
    generators, the same sanitizers, float-to-integer casts checked too (bucket and slot indices are computed from knots
    that may be infinite or NaN-adjacent)."""
    asan = _runtime("libasan.so")
    cxx = os.environ.get("SAN_CXX") or "/usr/bin/g++"
    if asan is None or not os.path.exists(cxx):
        pytest.skip("no AddressSanitizer runtime in this image")
    lib = str(tmp_path / "libresolve_san.so")
    csrc = os.path.join(HERE, "..", "enterpolation_b200", "csrc")
    made = subprocess.run([cxx, "-O1", "-g1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                           "-fsanitize=address,undefined,float-cast-overflow", "-fno-sanitize=float-divide-by-zero",
                           "-fno-sanitize-recover=all", "-o", lib, os.path.join(csrc, "resolve.cpp"),
                           os.path.join(HERE, "cpp", "resolve_san_harness.cpp")], capture_output=True, text=True)
    if made.returncode != 0:
        pytest.skip("sanitizer build of resolve.cpp failed: " + made.stderr[-300:])
    env = dict(os.environ)
    env["LD_PRELOAD"] = asan + " " + (_runtime("libstdc++.so.6") or "")
    env["ASAN_OPTIONS"] = "detect_leaks=0:abort_on_error=1"
    env["UBSAN_OPTIONS"] = "halt_on_error=1:print_stacktrace=1"
    run = subprocess.run([sys.executable, os.path.join(HERE, "resolve_sanitized_driver.py"), lib, "0", "60"], env=env,
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    assert "no sanitizer report" in run.stdout
