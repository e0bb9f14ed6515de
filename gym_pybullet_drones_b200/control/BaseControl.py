"""Controller base class (reference: gym_pybullet_drones/control/BaseControl.py:8-216)."""
import numpy as np

from ..params import DRONE_PROPERTIES, PIDCoefficients
from ..utils.enums import DroneModel


class BaseControl(object):
    """Same surface as the reference: `computeControl`, `computeControlFromState`, `reset`,
    `setPIDCoefficients`, `_getURDFParameter`; one instance drives `num_drones` drones at once."""

    def __init__(self, drone_model: DroneModel, g: float = 9.8):
        self.DRONE_MODEL = drone_model
        self.GRAVITY = g * self._getURDFParameter('m')          # BaseControl.py:35
        self.KF = self._getURDFParameter('kf')
        self.KM = self._getURDFParameter('km')
        self.reset()

    def reset(self):
        self.control_counter = 0                                # BaseControl.py:51

    def computeControlFromState(self, control_timestep, state, target_pos,
                                target_rpy=None, target_vel=None, target_rpy_rates=None):
        """Interface using the 20-float state vector(s) (BaseControl.py:55-93):
        pos = state[0:3], quat = state[3:7], vel = state[10:13], ang_vel = state[13:16]."""
        return self.computeControl(control_timestep=control_timestep,
                                   cur_pos=state[..., 0:3], cur_quat=state[..., 3:7],
                                   cur_vel=state[..., 10:13], cur_ang_vel=state[..., 13:16],
                                   target_pos=target_pos, target_rpy=target_rpy, target_vel=target_vel,
                                   target_rpy_rates=target_rpy_rates)

    def computeControl(self, control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos,
                       target_rpy=None, target_vel=None, target_rpy_rates=None):
        raise NotImplementedError

    def setPIDCoefficients(self, p_coeff_pos=None, i_coeff_pos=None, d_coeff_pos=None,
                           p_coeff_att=None, i_coeff_att=None, d_coeff_att=None):
        """BaseControl.setPIDCoefficients (BaseControl.py:138-177)."""
        ATTR_LIST = ['P_COEFF_FOR', 'I_COEFF_FOR', 'D_COEFF_FOR', 'P_COEFF_TOR', 'I_COEFF_TOR', 'D_COEFF_TOR']
        if not all(hasattr(self, attr) for attr in ATTR_LIST):
            raise AttributeError("[ERROR] in BaseControl.setPIDCoefficients(), not all PID coefficients exist as attributes in the instantiated control class.")
        for attr, val in zip(ATTR_LIST, (p_coeff_pos, i_coeff_pos, d_coeff_pos, p_coeff_att, i_coeff_att, d_coeff_att)):
            if val is not None:
                setattr(self, attr, np.asarray(val, dtype=np.float64))
        self._coefficients_changed()

    def _coefficients_changed(self):
        pass

    def _getURDFParameter(self, parameter_name: str):
        """Constants table lookup with the reference's parameter names (BaseControl.py:181-216)."""
        u = DRONE_PROPERTIES[self.DRONE_MODEL]
        if parameter_name in u:
            return u[parameter_name]
        raise KeyError(parameter_name)


_ = PIDCoefficients
