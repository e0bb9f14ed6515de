"""Single-drone hover task on the GPU simulator (reference: gym_pybullet_drones/envs/HoverAviary.py)."""
import numpy as np

from .. import _native as N
from ..utils.enums import ActionType, DroneModel, ObservationType, Physics
from .BaseRLAviary import BaseRLAviary


class HoverAviary(BaseRLAviary):
    """Single agent RL problem: hover at position (HoverAviary.py:6).

    reward = max(0, 2 - |target - pos|^4); terminated = |target - pos| < 1e-4; truncated on |x|,|y| > 1.5,
    z > 2, |roll|,|pitch| > 0.4 or after EPISODE_LEN_SEC (HoverAviary.py:68-117) -- all evaluated inside qs_step."""

    def __init__(self,
                 drone_model: DroneModel = DroneModel.CF2X,
                 initial_xyzs=None,
                 initial_rpys=None,
                 physics: Physics = Physics.PYB,
                 pyb_freq: int = 240,
                 ctrl_freq: int = 30,
                 gui=False,
                 record=False,
                 obs: ObservationType = ObservationType.KIN,
                 act: ActionType = ActionType.RPM,
                 **vec_kwargs):
        self.TARGET_POS = np.array([0, 0, 1])            # HoverAviary.py:51
        self.EPISODE_LEN_SEC = 8                         # HoverAviary.py:52
        super().__init__(drone_model=drone_model, num_drones=1, initial_xyzs=initial_xyzs, initial_rpys=initial_rpys,
                         physics=physics, pyb_freq=pyb_freq, ctrl_freq=ctrl_freq, gui=gui, record=record,
                         obs=obs, act=act, **vec_kwargs)

    def _task(self):
        return N.TASK_HOVER

    def _task_params(self):
        return dict(xy_bound=1.5, z_bound=2.0, tilt_bound=0.4, term_dist=1e-4)

    def _target_table(self):
        return np.asarray(self.TARGET_POS, dtype=np.float64).reshape(1, 3)
