"""Host-side logic of formation.py that needs no GPU: Z-order permutation, shard alignment rule."""
import numpy as np
import pytest

from gym_pybullet_drones_b200.formation import FormationShard, morton_order
from gym_pybullet_drones_b200.utils.enums import Physics


def test_morton_order_is_a_permutation_with_compact_chunks():
    side = 64
    k = np.arange(side * side)
    xy = np.stack([0.15 * (k % side), 0.15 * (k // side)], axis=1)
    perm = morton_order(xy)
    assert sorted(perm.tolist()) == k.tolist()
    # 32 consecutive drones of the Z-order cover an 8 x 4 cell block; of the row-major order a 32 x 1 strip
    def mean_diag(order):
        p = xy[order].reshape(-1, 32, 2)
        return np.linalg.norm(p.max(1) - p.min(1), axis=1).mean()
    assert mean_diag(perm) < 0.3 * mean_diag(k)
    # degenerate input (all drones on one point / one line) still gives a permutation
    assert sorted(morton_order(np.zeros((10, 2))).tolist()) == list(range(10))
    assert sorted(morton_order(np.stack([np.arange(10.), np.zeros(10)], 1)).tolist()) == list(range(10))


def test_p2p_exchange_needs_chunk_aligned_shards():
    xyz = np.zeros((100, 3))
    with pytest.raises(ValueError, match="multiple of 32"):
        FormationShard(xyz, physics=Physics.PYB_DW, exchange="p2p", rank=0, world=2)
    with pytest.raises(ValueError):
        FormationShard(xyz, physics=Physics.PYB_DW, exchange="local", rank=0, world=2)
    with pytest.raises(ValueError):
        FormationShard(xyz[:, :2], physics=Physics.PYB_DW)
