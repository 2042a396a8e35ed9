"""Runs the fuzz generators (tests/fuzzgen.py) through the ORACLE ALONE; started by tests/test_oracle_sanitized.py in a
child process with the AddressSanitizer/UBSan build of the oracle preloaded. Any out-of-bounds access, signed overflow or
misaligned access inside the checker aborts the process (non-zero exit). usage: oracle_sanitized_driver.py FIRST COUNT"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fuzzgen import DT, _degenerate_curve, _degenerate_inputs, _random_const_equidistant, _random_curve  # noqa: E402
from oracle import oracle  # noqa: E402


def exercise(rng, scalar, r):
    d = r.domain()
    lo, hi = float(d[0]), float(d[1])
    t = _degenerate_inputs(rng, DT[scalar], lo, hi, int(rng.choice([7, 1000, 5000])))
    variants = [r]
    if rng.random() < 0.3:
        variants.append(r.clamp())
    if rng.random() < 0.3 and np.isfinite(lo) and np.isfinite(hi):
        variants.append(r.slice(lo + 0.25 * (hi - lo), hi + 0.1))
    n_inv = 0
    for rv in variants:
        _, inv = rv.eval(t, return_invalid=True)
        n_inv += inv
        rv.span(t)
        k = int(rng.integers(20, 33))
        n = int(rng.choice([1, 2, 33, 4099]))
        first = (1 << k) - int(rng.integers(0, n + 1))
        total = first + n + int(rng.choice([0, 1, 12345, 1 << k]))
        _, inv = rv.take(total, first, n, return_invalid=True)
        n_inv += inv
        rv.take(n)
    return n_inv


def main(first, count):
    curves = panics = 0
    for seed in range(first, first + count):
        rng = np.random.default_rng(0x5A17 + seed)
        for gen in (_random_curve, _degenerate_curve, _random_const_equidistant, _degenerate_curve):
            for _ in range(6):
                scalar, b = gen(rng)
                try:
                    r = oracle.OracleCurve(b.to_desc())
                except Exception:
                    continue  # a builder rule was violated
                panics += exercise(rng, scalar, r)
                curves += 1
    print(f"sanitized oracle: {curves} curves, {panics} samples where the reference panics, no sanitizer report")


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
