"""Feeds the fuzz generators' descriptors (tests/fuzzgen.py) to the sanitizer build of the product's host resolution;
started by tests/test_oracle_sanitized.py in a child process. usage: resolve_sanitized_driver.py LIB FIRST COUNT"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fuzzgen import _degenerate_curve, _random_const_equidistant, _random_curve  # noqa: E402


def main(lib, first, count):
    L = C.CDLL(lib)
    L.resolve_san.argtypes = [C.c_void_p]
    L.resolve_san.restype = C.c_int
    n = ok = 0
    for seed in range(first, first + count):
        rng = np.random.default_rng(0x7E50 + seed)
        for gen in (_random_curve, _degenerate_curve, _random_const_equidistant, _degenerate_curve):
            for _ in range(6):
                _, b = gen(rng)
                try:
                    da = b.to_desc()
                except Exception:
                    continue
                ok += L.resolve_san(C.addressof(da.desc)) == 0
                n += 1
    print(f"sanitized resolve: {n} descriptors, {ok} accepted, no sanitizer report")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
