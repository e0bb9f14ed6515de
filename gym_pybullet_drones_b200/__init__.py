"""B200-native vectorised quadrotor simulator: the Physics.DYN hot path of
utiasDSL/gym-pybullet-drones (BaseAviary.step + DSLPIDControl.computeControl) as
hand-written sm_100a CUDA kernels behind the reference's env/controller API.

Sub-packages mirror the reference layout: `envs`, `control`, `utils`.
The CUDA library (libquadsim.so, C ABI in include/quadsim.h) is mandatory: there is no CPU fallback.
"""
__version__ = "0.1.0"

try:  # register gymnasium ids like the reference (gym_pybullet_drones/__init__.py:1-21) when gymnasium exists
    from gymnasium.envs.registration import register as _register
    for _id, _ep in (("ctrl-aviary-v0", "CtrlAviary"), ("hover-aviary-v0", "HoverAviary"), ("multihover-aviary-v0", "MultiHoverAviary")):
        try:
            _register(id="b200-" + _id, entry_point="gym_pybullet_drones_b200.envs:" + _ep)
        except Exception:
            pass
except Exception:
    pass
