"""RL plumbing of the GPU aviaries: action types, action buffer, KIN observations.

Mirrors `BaseRLAviary` (gym_pybullet_drones/envs/BaseRLAviary.py:13-322).  The per-drone
Python loops of `_preprocessAction` / `_computeObs` live inside the fused CUDA step
(qs_step): this class only declares the spaces and the kernel configuration.
"""
import numpy as np

from .. import _native as N
from .._compat import spaces
from ..utils.enums import ActionType, DroneModel, ObservationType, Physics
from .BaseAviary import BaseAviary

_ACT = {ActionType.RPM: (N.ACT_RPM, 4), ActionType.VEL: (N.ACT_VEL, 4), ActionType.PID: (N.ACT_PID, 3),
        ActionType.ONE_D_RPM: (N.ACT_ONE_D_RPM, 1), ActionType.ONE_D_PID: (N.ACT_ONE_D_PID, 1)}


class BaseRLAviary(BaseAviary):
    """Base single and multi-agent environment class for reinforcement learning (BaseRLAviary.py:10)."""

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
                 obs: ObservationType = ObservationType.KIN,
                 act: ActionType = ActionType.RPM,
                 **vec_kwargs):
        #### Create a buffer for the last .5 sec of actions (BaseRLAviary.py:66) ####
        self.ACTION_BUFFER_SIZE = int(ctrl_freq // 2)
        if obs != ObservationType.KIN:
            raise NotImplementedError("ObservationType.RGB needs PyBullet's renderer; the GPU simulator provides KIN")
        if act not in _ACT:
            print("[ERROR] in BaseRLAviary._actionSpace()")
            raise ValueError("unknown ActionType %r" % (act,))
        self.OBS_TYPE = obs
        self.ACT_TYPE = act
        if act in [ActionType.PID, ActionType.VEL, ActionType.ONE_D_PID] and drone_model not in [DroneModel.CF2X, DroneModel.CF2P]:
            raise ValueError("[ERROR] in BaseRLAviary.__init()__, no controller is available for the specified drone_model")   # :77-78
        super().__init__(drone_model=drone_model, num_drones=num_drones, neighbourhood_radius=neighbourhood_radius,
                         initial_xyzs=initial_xyzs, initial_rpys=initial_rpys, physics=physics,
                         pyb_freq=pyb_freq, ctrl_freq=ctrl_freq, gui=gui, record=record,
                         obstacles=True, user_debug_gui=False, vision_attributes=False, **vec_kwargs)
        if act == ActionType.VEL:
            self.SPEED_LIMIT = 0.03 * self.MAX_SPEED_KMH * (1000 / 3600)       # BaseRLAviary.py:95

    #### kernel configuration ####################################################
    def _act_type(self):
        return _ACT[self.ACT_TYPE][0]

    def _act_width(self):
        return _ACT[self.ACT_TYPE][1]

    def _act_buffer_size(self):
        return self.ACTION_BUFFER_SIZE

    ################################################################################

    def _actionSpace(self):
        """Box(-1, 1, (NUM_DRONES, A)) (BaseRLAviary.py:132-156)."""
        size = _ACT[self.ACT_TYPE][1]
        act_lower_bound = np.array([-1 * np.ones(size) for i in range(self.NUM_DRONES)])
        act_upper_bound = np.array([+1 * np.ones(size) for i in range(self.NUM_DRONES)])
        return spaces.Box(low=act_lower_bound, high=act_upper_bound, dtype=np.float32)

    def _observationSpace(self):
        """Box of shape (NUM_DRONES, 12 + ACTION_BUFFER_SIZE*A) (BaseRLAviary.py:243-277)."""
        lo, hi = -np.inf, np.inf
        size = _ACT[self.ACT_TYPE][1]
        kin_lo = np.array([[lo, lo, 0, lo, lo, lo, lo, lo, lo, lo, lo, lo] for i in range(self.NUM_DRONES)])
        kin_hi = np.array([[hi] * 12 for i in range(self.NUM_DRONES)])
        buf = self.ACTION_BUFFER_SIZE * size
        obs_lower_bound = np.hstack([kin_lo, -np.ones((self.NUM_DRONES, buf))])
        obs_upper_bound = np.hstack([kin_hi, +np.ones((self.NUM_DRONES, buf))])
        return spaces.Box(low=obs_lower_bound, high=obs_upper_bound, dtype=np.float32)

    def _computeObs(self):
        """Current observation (BaseRLAviary.py:284-322): [D, 12+B*A] ndarray, or the [E, D, .] tensor."""
        obs = self._obs_buf[self._cur]
        return self._shape_obs(obs) if self.VECTORIZED else self._obs_to_host_single(obs)
