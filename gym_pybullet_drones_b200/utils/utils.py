"""Script helpers under the reference's names (gym_pybullet_drones/utils/utils.py): `sync` paces a loop to the wall
clock, `str2bool` parses argparse booleans."""
import argparse
import time

_TRUE = {"yes", "true", "t", "y", "1"}
_FALSE = {"no", "false", "f", "n", "0"}


def sync(i, start_time, timestep):
    """Sleep so that iteration `i` of a loop with period `timestep` does not run ahead of real time.  Like the
    reference, fast loops (period <= 40 ms) are only checked about 24 times per simulated second."""
    stride = max(1, int(1.0 / (24.0 * timestep))) if timestep <= 0.04 else 1
    if i % stride:
        return
    ahead = i * timestep - (time.time() - start_time)
    if ahead > 0:
        time.sleep(ahead)


def str2bool(val):
    if isinstance(val, bool):
        return val
    key = str(val).strip().lower()
    if key in _TRUE:
        return True
    if key in _FALSE:
        return False
    raise argparse.ArgumentTypeError("[ERROR] in str2bool(), a Boolean value is expected")
