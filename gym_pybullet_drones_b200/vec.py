"""Stable-Baselines3 `VecEnv` protocol over a vectorised aviary.

`examples/learn.py` builds its training env with `make_vec_env(HoverAviary, env_kwargs=..., n_envs=1)`, i.e.
`DummyVecEnv([Monitor(env)])` (reference examples/learn.py:54-65).  `SB3VecAviary` gives the same protocol -- numpy
observations `[E, ...]`, `step_async/step_wait -> (obs, rewards, dones, infos)`, same-step auto-reset with
`infos[i]["terminal_observation"]` / `"TimeLimit.truncated"` and Monitor-style `infos[i]["episode"]` -- for E aviaries
that step in ONE kernel launch instead of E Python environments.  It subclasses SB3's `VecEnv` when SB3 is installed and is
duck-type compatible otherwise (this image has no SB3).
"""
import time

import numpy as np

try:  # pragma: no cover - depends on the environment
    from stable_baselines3.common.vec_env import VecEnv as _Base
    HAVE_SB3 = True
except Exception:
    _Base = object
    HAVE_SB3 = False


class SB3VecAviary(_Base):
    """`SB3VecAviary(HoverAviary, num_envs=4096, act=ActionType.ONE_D_RPM)`; extra kwargs go to the env constructor."""

    def __init__(self, env_cls, num_envs, **env_kwargs):
        env_kwargs = dict(env_kwargs)
        env_kwargs.update(num_envs=num_envs, autoreset="same_step")
        env_kwargs.setdefault("host_copy", True)
        self.env = env_cls(**env_kwargs)
        self.num_envs = int(num_envs)
        self.observation_space = self.env.single_observation_space
        self.action_space = self.env.single_action_space
        self.render_mode = None
        if HAVE_SB3:
            super().__init__(self.num_envs, self.observation_space, self.action_space)
        self._actions = None
        self._ep_ret = np.zeros(self.num_envs, np.float64)
        self._ep_len = np.zeros(self.num_envs, np.int64)
        self._t0 = time.time()
        self.reset_infos = [{} for _ in range(self.num_envs)]

    # ---- VecEnv protocol --------------------------------------------------------------------------------------------
    def reset(self):
        obs, _ = self.env.reset()
        self._ep_ret[:] = 0
        self._ep_len[:] = 0
        return obs.cpu().numpy()

    def step_async(self, actions):
        self._actions = np.asarray(actions, dtype=np.float32).reshape((self.num_envs,) + self.action_space.shape)

    def step_wait(self):
        obs, rew, term, trunc, info = self.env.step(self._actions)
        dones = term | trunc
        self._ep_ret += rew
        self._ep_len += 1
        infos = [{} for _ in range(self.num_envs)]
        if dones.any():
            idx = info.get("final_obs_env")
            if idx is None:
                raise RuntimeError("SB3VecAviary needs an env stepped through qs_step_host (autoreset='same_step', "
                                   "drones_per_env <= 128): this env did not report the finished aviaries")
            now = round(time.time() - self._t0, 6)
            for k, i in enumerate(idx):
                infos[i] = {"terminal_observation": info["final_obs"][k],
                            "TimeLimit.truncated": bool(trunc[i] and not term[i]),
                            "episode": {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i]), "t": now}}
            self._ep_ret[idx] = 0
            self._ep_len[idx] = 0
        return obs, rew.astype(np.float32), dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    def seed(self, seed=None):
        return [seed] * self.num_envs          # the simulator is deterministic (BaseAviary.reset ignores the seed too)

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    def get_attr(self, attr_name, indices=None):
        return [getattr(self.env, attr_name) for _ in self._indices(indices)]

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.env, attr_name, value)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        return [getattr(self.env, method_name)(*method_args, **method_kwargs) for _ in self._indices(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        # episode statistics are produced here, so report "wrapped by Monitor" for SB3's evaluate_policy
        name = getattr(wrapper_class, "__name__", "")
        return [name == "Monitor" for _ in self._indices(indices)]

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode=None):
        return self.env.render()
