"""Raw-RPM control aviary on the GPU simulator (reference: gym_pybullet_drones/envs/CtrlAviary.py)."""
import numpy as np

from .. import _native as N
from .._compat import spaces
from ..utils.enums import DroneModel, Physics
from .BaseAviary import BaseAviary


class CtrlAviary(BaseAviary):
    """Multi-drone environment class for control applications (CtrlAviary.py:7): action = RPMs clipped to
    [0, MAX_RPM] (CtrlAviary.py:121-140), observation = the 20-float state vector of every drone
    (CtrlAviary.py:106-117), reward -1, never terminated/truncated (CtrlAviary.py:144-185)."""

    def __init__(self,
                 drone_model: DroneModel = DroneModel.CF2X,
                 num_drones: int = 1,
                 neighbourhood_radius: float = np.inf,
                 initial_xyzs=None,
                 initial_rpys=None,
                 physics: Physics = Physics.PYB,
                 pyb_freq: int = 240,
                 ctrl_freq: int = 240,
                 gui=False,
                 record=False,
                 obstacles=False,
                 user_debug_gui=True,
                 output_folder='results',
                 **vec_kwargs):
        super().__init__(drone_model=drone_model, num_drones=num_drones, neighbourhood_radius=neighbourhood_radius,
                         initial_xyzs=initial_xyzs, initial_rpys=initial_rpys, physics=physics,
                         pyb_freq=pyb_freq, ctrl_freq=ctrl_freq, gui=gui, record=record, obstacles=obstacles,
                         user_debug_gui=user_debug_gui, output_folder=output_folder, **vec_kwargs)

    def _act_type(self):
        return N.ACT_RAW_RPM

    def _actionSpace(self):
        lo = np.array([[0., 0., 0., 0.] for i in range(self.NUM_DRONES)])
        hi = np.array([[self.MAX_RPM] * 4 for i in range(self.NUM_DRONES)])
        return spaces.Box(low=lo, high=hi, dtype=np.float32)

    def _observationSpace(self):
        m = self.MAX_RPM
        lo = np.array([[-np.inf, -np.inf, 0., -1., -1., -1., -1., -np.pi, -np.pi, -np.pi, -np.inf, -np.inf, -np.inf, -np.inf, -np.inf, -np.inf, 0., 0., 0., 0.] for i in range(self.NUM_DRONES)])
        hi = np.array([[np.inf, np.inf, np.inf, 1., 1., 1., 1., np.pi, np.pi, np.pi, np.inf, np.inf, np.inf, np.inf, np.inf, np.inf, m, m, m, m] for i in range(self.NUM_DRONES)])
        return spaces.Box(low=lo, high=hi, dtype=np.float32)

    def _computeObs(self):
        obs = self._obs_buf[self._cur]
        return self._shape_obs(obs) if self.VECTORIZED else self._obs_to_host_single(obs)

    def _launch(self, action_dev):
        obs = super()._launch(action_dev)
        # dummy task (CtrlAviary.py:144-185)
        self._reward.fill_(-1.0)
        return obs
