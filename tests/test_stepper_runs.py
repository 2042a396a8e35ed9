"""Integer restatement of the index arithmetic of the f32 stepper-run kernel (rows.cuh, RUNS): which representable
index `Equidistant::eval`'s `R::from_usize(i)` (src/base/list.rs:416-418) rounds a sample index to, computed per
tile / sub-tile exactly as the kernel does, against the hardware u64 -> f32 conversion (round to nearest even)."""
import numpy as np
import pytest

NV = 256  # BLOCK


def kernel_value_index(i, lt):
    """For sample indices `i` (uint64 array, all >= 2^24): the index v whose float the kernel evaluates for them."""
    i = np.asarray(i, dtype=np.uint64)
    out = np.empty_like(i)
    for n, ii in enumerate(i.tolist()):
        tile = ii >> lt
        base = tile << lt
        ls = base.bit_length() - 1 - 23
        assert ls >= 1
        T = 1 << lt
        sub = T if (ls >= lt or (NV << ls) > T) else (NV << ls)
        sb = base + (ii - base) // sub * sub
        vb = (sb >> ls) << ls
        nvals = ((sb + sub - 1 - vb) >> ls) + 2
        assert nvals <= NV + 1
        dd = ii - vb
        hm1 = ((1 << ls) >> 1) - 1
        par = ((dd >> ls) + ((vb >> ls) & 1)) & 1
        m = (dd + hm1 + par) >> ls
        assert m < nvals
        out[n] = vb + (m << ls)
    return out


@pytest.mark.parametrize("lt", [12, 13, 16])
def test_run_index_matches_float_conversion(lt):
    rng = np.random.default_rng(lt)
    parts = []
    for k in range(24, 63):
        p = 1 << k
        parts.append(np.arange(p, p + 700, dtype=np.uint64))
        parts.append(np.arange(p - 700, p, dtype=np.uint64) if k > 24 else np.arange(p, p + 1, dtype=np.uint64))
        parts.append(rng.integers(p, 2 * p, 400, dtype=np.uint64))
        s = 1 << (k - 23)
        mid = p + (p >> 1)
        # around run boundaries (ties to even in both directions)
        for c in (mid, mid + s, mid + 3 * s, p + s, 2 * p - s):
            parts.append(np.arange(c - s // 2 - 3, c - s // 2 + 4, dtype=np.uint64))
            parts.append(np.arange(c + s // 2 - 3, c + s // 2 + 4, dtype=np.uint64))
    i = np.concatenate(parts)
    i = i[i >= (1 << 24)]
    want = i.astype(np.float32)          # from_usize: one rounding, nearest even
    got = kernel_value_index(i, lt)
    assert np.array_equal(got.astype(np.float32), want)
    assert np.array_equal(got.astype(np.float32).astype(np.uint64), got)  # the evaluated index is itself a float
