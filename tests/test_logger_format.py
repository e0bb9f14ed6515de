"""Logger writes the reference's on-disk layout; checked against the reference's own Logger when it
is importable in this container (needs nothing GPU)."""
import os

import numpy as np
import pytest

from gym_pybullet_drones_b200.utils import Logger


def test_logger_layout_and_roundtrip(tmp_path):
    lg = Logger(logging_freq_hz=48, output_folder=str(tmp_path), num_drones=2)
    rng = np.random.default_rng(0)
    states = rng.normal(size=(5, 2, 20)); ctrls = rng.normal(size=(5, 2, 12))
    for t in range(5):
        if t % 2:
            lg.log_all(t / 48, states[t], ctrls[t])
        else:
            for j in range(2):
                lg.log(drone=j, timestamp=t / 48, state=states[t, j], control=ctrls[t, j])
    path = lg.save()
    d = np.load(path)
    assert d["timestamps"].shape == (2, 5) and d["states"].shape == (2, 16, 5) and d["controls"].shape == (2, 12, 5)
    want = np.concatenate([states[..., 0:3], states[..., 10:13], states[..., 7:10], states[..., 13:20]], axis=-1)   # Logger.py:117
    assert np.array_equal(d["states"], want.transpose(1, 2, 0)) and np.array_equal(d["controls"], ctrls.transpose(1, 2, 0))
    assert os.path.isdir(lg.save_as_csv("x"))
    with pytest.raises(ValueError):
        lg.log(drone=3, timestamp=0, state=states[0, 0])


def test_logger_matches_reference_logger(tmp_path):
    from oracle.ref_loader import reference_available
    if not reference_available():
        pytest.skip("reference tree not present (GPU box)")
    import sys
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "cycler"):          # plotting deps of the reference module; unused here
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["cycler"].cycler = lambda *a, **k: None
    from oracle.ref_loader import load_reference
    load_reference()
    from gym_pybullet_drones.utils.Logger import Logger as RefLogger
    a, b = RefLogger(logging_freq_hz=48, output_folder=str(tmp_path / "r"), num_drones=3), Logger(logging_freq_hz=48, output_folder=str(tmp_path / "m"), num_drones=3)
    rng = np.random.default_rng(1)
    for t in range(7):
        for j in range(3):
            s, c = rng.normal(size=20), rng.normal(size=12)
            a.log(drone=j, timestamp=t / 48, state=s, control=c)
            b.log(drone=j, timestamp=t / 48, state=s, control=c)
    assert np.array_equal(a.timestamps, b._trimmed()[0]) and np.array_equal(a.states, b._trimmed()[1]) and np.array_equal(a.controls, b._trimmed()[2])
