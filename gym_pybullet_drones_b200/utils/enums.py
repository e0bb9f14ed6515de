"""API constants of the reference, value-compatible (gym_pybullet_drones/utils/enums.py:3-47).

Member names and values are the contract (`ActionType('one_d_rpm')`, `DroneModel.CF2X.value + ".urdf"`, pickled
configs, argparse choices ...), so they are reproduced exactly; built with the functional Enum API from one table.
"""
from enum import Enum

_TABLE = {
    # drone models: Crazyflie 2.x in X and + configuration, and the racer
    "DroneModel": (("CF2X", "cf2x"), ("CF2P", "cf2p"), ("RACE", "racer")),
    # every member runs the explicit DYN model on the GPU; PYB_* members switch on the DYN+ aerodynamic terms
    "Physics": (("PYB", "pyb"), ("DYN", "dyn"), ("PYB_GND", "pyb_gnd"), ("PYB_DRAG", "pyb_drag"), ("PYB_DW", "pyb_dw"),
                ("PYB_GND_DRAG_DW", "pyb_gnd_drag_dw")),
    "ImageType": (("RGB", 0), ("DEP", 1), ("SEG", 2), ("BW", 3)),
    "ActionType": (("RPM", "rpm"), ("PID", "pid"), ("VEL", "vel"), ("ONE_D_RPM", "one_d_rpm"), ("ONE_D_PID", "one_d_pid")),
    "ObservationType": (("KIN", "kin"), ("RGB", "rgb")),
}

DroneModel = Enum("DroneModel", _TABLE["DroneModel"], module=__name__)
Physics = Enum("Physics", _TABLE["Physics"], module=__name__)
ImageType = Enum("ImageType", _TABLE["ImageType"], module=__name__)
ActionType = Enum("ActionType", _TABLE["ActionType"], module=__name__)
ObservationType = Enum("ObservationType", _TABLE["ObservationType"], module=__name__)
