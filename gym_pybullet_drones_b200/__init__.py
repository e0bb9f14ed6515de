"""B200-native vectorised quadrotor simulator: the Physics.DYN hot path of
utiasDSL/gym-pybullet-drones (BaseAviary.step + DSLPIDControl.computeControl) as
hand-written sm_100a CUDA kernels behind the reference's env/controller API.

Sub-packages mirror the reference layout: `envs`, `control`, `utils`.
The CUDA library (libquadsim.so, C ABI in include/quadsim.h) is mandatory: there is no CPU fallback.
"""
__version__ = "0.1.0"

try:  # register gymnasium ids like the reference (gym_pybullet_drones/__init__.py:1-21) when gymnasium exists
    from gymnasium.envs.registration import register as _register
    # the reference's own ids (so `gym.make("hover-aviary-v0")` resolves to the GPU env once this package is imported
    # instead of the reference), and the same ids under a "b200-" prefix for side-by-side use with the reference
    for _id, _ep in (("ctrl-aviary-v0", "CtrlAviary"), ("velocity-aviary-v0", "VelocityAviary"),
                     ("hover-aviary-v0", "HoverAviary"), ("multihover-aviary-v0", "MultiHoverAviary")):
        for _prefix in ("", "b200-"):
            try:
                _register(id=_prefix + _id, entry_point="gym_pybullet_drones_b200.envs:" + _ep)
            except Exception:
                pass
except Exception:
    pass
