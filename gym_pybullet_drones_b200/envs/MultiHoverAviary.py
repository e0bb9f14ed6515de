"""Multi-drone hover task on the GPU simulator (reference: gym_pybullet_drones/envs/MultiHoverAviary.py)."""
import numpy as np

from .. import _native as N
from ..utils.enums import ActionType, DroneModel, ObservationType, Physics
from .BaseRLAviary import BaseRLAviary


class MultiHoverAviary(BaseRLAviary):
    """Multi-agent RL problem: leader-follower hover (MultiHoverAviary.py:6).

    One scalar reward per aviary = sum_i max(0, 2 - |target_i - pos_i|^4); terminated when the summed distance
    is < 1e-4; truncated when ANY drone leaves |x|,|y| <= 2, z <= 2, |roll|,|pitch| <= 0.4 or on time-out
    (MultiHoverAviary.py:75-130)."""

    def __init__(self,
                 drone_model: DroneModel = DroneModel.CF2X,
                 num_drones: int = 2,
                 neighbourhood_radius: float = np.inf,
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
        self.EPISODE_LEN_SEC = 8                         # MultiHoverAviary.py:57
        self._num_drones_for_target = num_drones
        super().__init__(drone_model=drone_model, num_drones=num_drones, neighbourhood_radius=neighbourhood_radius,
                         initial_xyzs=initial_xyzs, initial_rpys=initial_rpys, physics=physics,
                         pyb_freq=pyb_freq, ctrl_freq=ctrl_freq, gui=gui, record=record, obs=obs, act=act, **vec_kwargs)

    def _task(self):
        return N.TASK_HOVER

    def _task_params(self):
        return dict(xy_bound=2.0, z_bound=2.0, tilt_bound=0.4, term_dist=1e-4)

    def _target_table(self):
        # TARGET_POS = INIT_XYZS + [0, 0, 1/(i+1)]  (MultiHoverAviary.py:71)
        self.TARGET_POS = self.INIT_XYZS + np.array([[0, 0, 1 / (i + 1)] for i in range(self._num_drones_for_target)])
        return self.TARGET_POS
