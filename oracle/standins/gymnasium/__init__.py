"""TEST-ONLY minimal stand-in for `gymnasium` (absent from this image) so the
unmodified reference envs import.  Only what BaseAviary/BaseRLAviary touch:
`Env.reset(seed, options)`, `spaces.Box`, `envs.registration.register`."""
import numpy as _np
from . import spaces  # noqa: F401
from . import envs  # noqa: F401


class Env:
    metadata = {}

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.np_random = _np.random.default_rng(seed)

    def close(self):
        pass
