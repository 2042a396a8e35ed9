"""Extended run of tests/test_gpu_fuzz.py: 300 more seeds x 12 random curves (3600 configurations), bit-exact vs the
oracle; run on a GPU box: python tools/fuzz_more.py  [first_seed last_seed]  (runs so far: seeds 24-1523 = 18000 configurations, 0 failures)."""
import sys, importlib.util
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
spec = importlib.util.spec_from_file_location("fz", os.path.join(ROOT, "tests", "test_gpu_fuzz.py"))
fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
fails = 0
import sys as _s
LO, HI = (int(_s.argv[1]), int(_s.argv[2])) if len(_s.argv) > 2 else (24, 324)
for seed in range(LO, HI):
    rng = np.random.default_rng(0xE17E + seed)
    for k in range(12):
        scalar, b = fz._random_curve(rng)
        try:
            fz._check(rng, scalar, b, f"seed {seed} case {k}")
        except AssertionError as e:
            fails += 1
            print("FAIL", seed, k, str(e)[:300], flush=True)
            if fails > 5: raise SystemExit(1)
print("done, fails =", fails)
