"""gymnasium compatibility: use the real package when it is installed (so SB3 / `gym.make`
see genuine `gymnasium.Env` / `spaces.Box` objects), else a minimal internal stand-in with
the same surface (this image has no gymnasium)."""
import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as _gym
    from gymnasium import spaces  # noqa: F401
    Env = _gym.Env
    HAVE_GYMNASIUM = True
    try:
        from gymnasium.vector import AutoresetMode as _AM
        AUTORESET_NEXT_STEP, AUTORESET_SAME_STEP, AUTORESET_DISABLED = _AM.NEXT_STEP, _AM.SAME_STEP, _AM.DISABLED
    except Exception:
        AUTORESET_NEXT_STEP, AUTORESET_SAME_STEP, AUTORESET_DISABLED = "NextStep", "SameStep", "Disabled"
except Exception:
    HAVE_GYMNASIUM = False
    AUTORESET_NEXT_STEP, AUTORESET_SAME_STEP, AUTORESET_DISABLED = "NextStep", "SameStep", "Disabled"

    class Env:
        """Subset of gymnasium.Env used by the aviaries."""
        metadata = {}
        render_mode = None
        spec = None

        def reset(self, seed=None, options=None):
            if seed is not None:
                self.np_random = np.random.default_rng(seed)

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self

    class _Box:
        """Subset of gymnasium.spaces.Box: low/high/shape/dtype, sample(), contains()."""

        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.asarray(low).shape
            self.shape = tuple(int(s) for s in shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return "Box(%s, %s)" % (self.shape, self.dtype)

    class spaces:  # noqa: N801 - mirrors the module name
        Box = _Box


def batch_box(space, n):
    """Batched Box for VectorEnv.observation_space / action_space."""
    return spaces.Box(low=np.broadcast_to(space.low, (n,) + space.shape).copy(),
                      high=np.broadcast_to(space.high, (n,) + space.shape).copy(), dtype=space.dtype)
