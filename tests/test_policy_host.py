"""Host side of the on-device policy (CPU): the fragment-ordered weight layout of include/quadsim.h (QsPolicy) and the two-term
float16 split it stores."""
import numpy as np
import torch

from gym_pybullet_drones_b200.policy import MlpPolicy


def _words(part):
    return part.contiguous().view(torch.int16).to(torch.int32) & 0xFFFF


def test_fragment_order_matches_the_header_definition():
    torch.manual_seed(0)
    K, Nc = 48, 16
    hi, lo = torch.randn(K, Nc).to(torch.float16), torch.randn(K, Nc).to(torch.float16)
    for first in (False, True):
        fr = MlpPolicy._fragment_order(hi, lo, first)
        assert fr.shape == (K // 16, Nc // 8, 32, 4) and fr.dtype == torch.int32
        for ks in range(K // 16):
            for n in range(Nc // 8):
                for lane in range(32):
                    g, t = lane // 4, lane % 4
                    ka, kb = (4 * t, 4 * t + 2) if first else (2 * t, 2 * t + 8)
                    for w, (part, k) in enumerate(((hi, ka), (hi, kb), (lo, ka), (lo, kb))):
                        u = _words(part)
                        want = int(u[16 * ks + k, 8 * n + g]) | (int(u[16 * ks + k + 1, 8 * n + g]) << 16)
                        assert (int(fr[ks, n, lane, w]) & 0xFFFFFFFF) == want, (first, ks, n, lane, w)


def test_two_term_split_reproduces_fp32_weights():
    """hi + lo' / 2048 equals the fp32 weight to ~2^-22 relative -- below the fp16 normal range (|w| < 6e-5) to 2e-11 absolute --
    which is what makes the tensor-core product fp32-accurate."""
    g = torch.Generator().manual_seed(1)
    w = (torch.rand(144, 64, generator=g) * 2 - 1) * torch.logspace(-4, 1, 64)
    hi = w.to(torch.float16)
    lo = ((w - hi.to(torch.float32)) * 2048.0).to(torch.float16)
    back = hi.to(torch.float64) + lo.to(torch.float64) / 2048.0
    err = (back - w.to(torch.float64)).abs()
    assert bool((err <= 2.0 ** -21 * w.abs().to(torch.float64) + 2e-11).all())


def test_prepare_pads_rows_to_16_and_last_columns_to_8():
    pol = MlpPolicy.random(27, 1, seed=2, critic=True, device="cpu")        # HoverAviary ONE_D_RPM: in 27, out 1
    (w1, b1), (w2, b2), (w3, b3) = pol._split["actor"]
    assert w1.shape == (2, 8, 32, 4) and w2.shape == (4, 8, 32, 4) and w3.shape == (4, 1, 32, 4)
    assert b1.shape == (64,) and b3.shape == (8,) and float(b3[1:].abs().max()) == 0.0
    # zero rows past in_dim: k-step 1 holds rows 16..31, rows 27..31 are padding -> layer-1 lanes t with 4t+j >= 11 in that k-step
    t = torch.arange(32) % 4
    pad = (4 * t + 2 + 16 >= 28)                                             # second pair (rows 16+4t+2, +3) entirely past row 27
    assert int(w1[1, :, pad, 1].abs().max()) == 0 and int(w1[1, :, pad, 3].abs().max()) == 0
    assert np.isfinite(pol.forward_torch(torch.zeros(3, 27))[0].numpy()).all()
