"""SASS evidence for the headline kernels: counts of the mnemonics that prove TMA bulk copies (UBLKCP + SYNCS),
packed f32x2 arithmetic (FMUL2/FFMA2/FADD2), 128/256-bit global accesses and shared-memory traffic, per kernel, from the
shipped libenterp_b200.so. Runs without a GPU.  usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "enterpolation_b200", "libenterp_b200.so")

KERNELS = [
    ("C5 take(2^32), stepper runs (VARIANT 4)", r"bspline_rows_kernelIfLi4ELb0ELi3ELb1ELb1ELi4E"),
    ("C5 head < 2^24, 5 stops (VARIANT 2, TMA-staged rows)", r"bspline_rows_kernelIfLi4ELb0ELi3ELb1ELb1ELi2E"),
    ("C5 head < 2^24, 256 stops (VARIANT 3, exact-division rows)", r"bspline_rows_kernelIfLi4ELb0ELi3ELb1ELb0ELi3E"),
    ("C5 256 stops, stepper runs (VARIANT 4)", r"bspline_rows_kernelIfLi4ELb0ELi3ELb1ELb0ELi4E"),
    ("C5-like random batch (VARIANT 0)", r"bspline_rows_kernelIfLi4ELb0ELi3ELb1ELb1ELi0E"),
    ("C1 f64 take (VARIANT 3)", r"bspline_rows_kernelIdLi1ELb0ELi3ELb1ELb0ELi3E"),
    ("C2 Linear [f32;3] sorted batch, slot table (linear_slots_kernel)", r"linear_slots_kernelIfLi3ELb0E"),
    ("C2 fallback: Linear [f32;3] sorted batch, bucket index + records", r"linear_kernelIfLi3ELb0ELb0E"),
    ("cubic [f32;4] B-spline over sorted knots, knot-row slots", r"bspline_slots_kernelIfLi4ELb0ELi3E"),
    ("run-length host transfer: stepper values at representable indices", r"stepper_strided_kernelIfE"),
    ("C3 bezier_many [f32;3] x 16", r"bezier_many_kernelIfLi3ELi16E"),
    ("C4 NURBS f64 generic", r"bspline_kernelIdLi4ELb1ELi3ELb0E"),
]
PATTERNS = ["UBLKCP", "SYNCS", "FMUL2", "FFMA2", "FADD2", "FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD", "MUFU",
            r"LDG\.E\S*\.256", r"LDG\.E\S*\.128", r"STG\.E\S*\.256", r"STG\.E\S*\.128", "LDS", "STS", r"BAR\.SYNC", "ATOM|RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    by_name = {}
    for b in blocks[1:]:
        name, _, body = b.partition("\n")
        by_name[name.strip()] = body
    print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)} ({os.path.getsize(SO) / 1e6:.1f} MB), {len(by_name)} kernels")
    for label, pat in KERNELS:
        hits = [n for n in by_name if re.search(pat, n)]
        if not hits:
            print(f"\n## {label}: NOT FOUND ({pat})")
            continue
        name = hits[0]
        lines = [l for l in by_name[name].splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        ops = collections.Counter()
        for l in lines:
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ops[m.group(1)] += 1
        print(f"\n## {label}\n{name}\ninstructions: {len(lines)}")
        for p in PATTERNS:
            rx = re.compile(r"^(?:" + p + r")(?:\.|$)") if not ("\\" in p or "|" in p) else re.compile(p)
            n = sum(c for o, c in ops.items() if rx.search(o))
            if n:
                print(f"  {p:18s} {n}")


if __name__ == "__main__":
    sys.exit(main())
