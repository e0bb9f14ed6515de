"""The chunk-culling rule of the downwash kernels, restated in NumPy float32 and checked as a property on
CPU: whenever the box test says "skip", EVERY pair between the two boxes fails the reference's predicate
(BaseAviary.py:800) or takes the pair term's underflow early-out -- i.e. it would have added exactly 0.0f.  (The GPU
tests check the same thing end to end: culled == all-pairs, bit for bit.)"""
import numpy as np

DW1, DW2, DW3 = np.float32(2267.18), np.float32(0.16), np.float32(-0.11)
f32 = np.float32


def pair_is_zero(dz, dxy2):
    """dw_pair()'s early-out conditions (csrc/quadsim.cu), float32 arithmetic."""
    beta = DW2 * dz + DW3
    b2 = beta * beta
    return ~((dz > 0) & (dxy2 < f32(100.0))) | (dxy2 > f32(220.0) * b2)


def box_skip(rlo, rhi, clo, chi):
    """The warp's test of a source chunk box [clo, chi] against the box of its rows [rlo, rhi] (float32)."""
    dzhi = chi[2] - rlo[2]
    dzlo = np.maximum(clo[2] - rhi[2], f32(0))
    gx = np.maximum(np.maximum(clo[0] - rhi[0], rlo[0] - chi[0]), f32(0))
    gy = np.maximum(np.maximum(clo[1] - rhi[1], rlo[1] - chi[1]), f32(0))
    g2 = gx * gx + gy * gy
    bm = np.maximum(np.abs(DW2 * dzlo + DW3), np.abs(DW2 * dzhi + DW3))
    act = (dzhi > 0) & ~(g2 > f32(100.001)) & ~(g2 > f32(220.0) * bm * bm)
    return not act


def test_box_skip_implies_every_pair_is_zero():
    rng = np.random.default_rng(0)
    skipped = kept = 0
    for trial in range(4000):
        scale = f32(rng.choice([0.3, 1.0, 4.0, 15.0]))
        zs = f32(rng.choice([0.1, 0.8, 5.0]))
        c0 = rng.uniform(-1, 1, 3).astype(np.float32) * np.array([scale, scale, zs], np.float32)
        rows = (c0 * f32(0) + rng.uniform(-1, 1, (32, 3)).astype(np.float32) * np.array([0.5, 0.5, zs], np.float32))
        src = (c0 + rng.uniform(-1, 1, (32, 3)).astype(np.float32) * np.array([0.5, 0.5, zs], np.float32))
        if box_skip(rows.min(0), rows.max(0), src.min(0), src.max(0)):
            skipped += 1
            dz = src[None, :, 2] - rows[:, None, 2]
            dx = src[None, :, 0] - rows[:, None, 0]
            dy = src[None, :, 1] - rows[:, None, 1]
            assert pair_is_zero(dz, dx * dx + dy * dy).all(), trial
        else:
            kept += 1
    assert skipped > 500 and kept > 500          # the property is exercised on both sides


def test_underflow_early_out_is_below_float32_range():
    """dxy^2 > 220 beta^2  =>  exp(-dxy^2 / (2 beta^2)) < 2^-158: far below the smallest float32 subnormal (2^-149)
    and a fortiori below ex2.approx.ftz's flush threshold (2^-126)."""
    u2 = np.float64(220.0)
    assert -0.5 * u2 * np.log2(np.e) < -158
