"""Script helpers with the reference's names (gym_pybullet_drones/utils/utils.py:10-54)."""
import argparse
import time


def sync(i, start_time, timestep):
    """Paces a loop to the wall clock (utils.py:10-30)."""
    if timestep > .04 or i % (int(1 / (24 * timestep))) == 0:
        elapsed = time.time() - start_time
        if elapsed < (i * timestep):
            time.sleep(timestep * i - elapsed)


def str2bool(val):
    """argparse boolean (utils.py:33-54)."""
    if isinstance(val, bool):
        return val
    if val.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if val.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError("[ERROR] in str2bool(), a Boolean value is expected")
