"""TEST-ONLY import stub: gym_pybullet_drones/envs/__init__.py:1 pulls BetaAviary,
which imports these names at module import (never called on the DYN path)."""
