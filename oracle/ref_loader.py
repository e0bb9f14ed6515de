"""Tier-1 oracle loader: import the UNMODIFIED reference with test-only stand-ins.

ORACLE / TEST INFRASTRUCTURE.  Works only where /root/reference exists (the build
container); the GPU box has no reference, so nothing under `-m gpu`, smoke() or
bench.py calls this -- they use the committed golden fixtures and dyn_oracle.py.
"""
import contextlib
import io
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("QS_REFERENCE_ROOT", "/root/reference")
_STANDINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standins")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_pybullet_drones"))


def load_reference():
    """Returns a namespace with the reference's classes (DYN-capable subset)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (_STANDINS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    warnings.filterwarnings("ignore", category=UserWarning)
    import types
    with contextlib.redirect_stdout(io.StringIO()):
        import pybullet  # the stand-in  # noqa: F401
        from gym_pybullet_drones.utils.enums import DroneModel, Physics, ActionType, ObservationType
        from gym_pybullet_drones.envs.BaseAviary import BaseAviary
        from gym_pybullet_drones.envs.CtrlAviary import CtrlAviary
        from gym_pybullet_drones.envs.HoverAviary import HoverAviary
        from gym_pybullet_drones.envs.MultiHoverAviary import MultiHoverAviary
        from gym_pybullet_drones.envs.VelocityAviary import VelocityAviary
        from gym_pybullet_drones.control.DSLPIDControl import DSLPIDControl
    assert "standins" in pybullet.__file__, "real pybullet shadowed the stand-in?"
    ns = types.SimpleNamespace(**{k: v for k, v in locals().items() if not k.startswith("_")})
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints its URDF constants on every construction."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
