"""On-device policy for `BaseRLAviary.rollout` (reference caller: examples/learn.py:67-95, SB3's
`PPO('MlpPolicy', env)` whose `collect_rollouts` alternates policy.forward and env.step).

`MlpPolicy` holds the weights of an SB3-MlpPolicy-shaped actor (flatten -> 64 tanh -> 64 tanh -> linear mean, state-independent
`log_std`) and optionally the critic (same trunk shape, one output) as row-major `[in][out]` float32 CUDA tensors, plus
the layout the rollout kernel reads: every matrix split once into its TF32 halves (`hi` = weights rounded to TF32, `lo` =
the rounded remainder; 3xTF32 tensor-core products then reproduce fp32), rows / last-layer columns zero-padded to 8.  `rollout(policy=...)` then evaluates it inside the kernel every tick, from the observation
window in shared memory: no policy launch, no action tensor round trip.  `forward_torch` is the same network in plain PyTorch
fp32 (what a learner would run for the gradient step, and what the tests compare the kernel with)."""
import ctypes as C

import torch

from . import _native as N


class MlpPolicy:
    HIDDEN = 64

    def __init__(self, actor, log_std, critic=None, device=None):
        """actor / critic: three (weight [in, out], bias [out]) pairs each; log_std: [out_dim]."""
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())

        def prep(net, last):
            if net is None:
                return None
            if len(net) != 3:
                raise ValueError("the on-device policy is a 3-layer MLP (in -> 64 -> 64 -> out)")
            out = []
            for k, (w, b) in enumerate(net):
                w = torch.as_tensor(w, dtype=torch.float32).to(dev).contiguous()
                b = torch.as_tensor(b, dtype=torch.float32).to(dev).contiguous()
                if w.dim() != 2 or b.shape != (w.shape[1],):
                    raise ValueError("layer %d: weight must be [in, out] and bias [out]" % k)
                out.append((w, b))
            if out[0][0].shape[1] != self.HIDDEN or out[1][0].shape != (self.HIDDEN, self.HIDDEN) or out[2][0].shape[0] != self.HIDDEN:
                raise ValueError("hidden width must be 64 (SB3 MlpPolicy default)")
            if last is not None and out[2][0].shape[1] != last:
                raise ValueError("last layer must have %d outputs" % last)
            return out

        self.actor = prep(actor, None)
        self.in_dim, self.out_dim = self.actor[0][0].shape[0], self.actor[2][0].shape[1]
        self.critic = prep(critic, 1)
        if self.critic is not None and self.critic[0][0].shape[0] != self.in_dim:
            raise ValueError("critic input width differs from the actor's")
        self.log_std = torch.as_tensor(log_std, dtype=torch.float32).to(dev).contiguous().reshape(self.out_dim)
        self.device = dev
        if self.out_dim > 32:
            raise ValueError("the on-device policy supports up to 32 action outputs per aviary (D * A)")
        self.nt3 = 1 if self.out_dim <= 8 else (2 if self.out_dim <= 16 else 4)
        self._split = {"actor": self._prepare(self.actor, 8 * self.nt3), "critic": None if self.critic is None else self._prepare(self.critic, 8)}

    @staticmethod
    def _tf32(x):
        """cvt.rna.tf32.f32: round to nearest (ties away) onto the 10-bit TF32 mantissa; the result is an fp32 with 13 zero bits."""
        i = x.contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)

    def _prepare(self, net, last_cols):
        """[(hi, lo, bias)] x 3 with zero-padded rows (multiple of 8) and last-layer columns."""
        out = []
        for k, (w, b) in enumerate(net):
            rows = (w.shape[0] + 7) // 8 * 8
            cols = last_cols if k == 2 else w.shape[1]
            wp = torch.zeros((rows, cols), dtype=torch.float32, device=w.device)
            wp[:w.shape[0], :w.shape[1]] = w
            bp = torch.zeros((cols,), dtype=torch.float32, device=w.device)
            bp[:b.shape[0]] = b
            hi = self._tf32(wp)
            lo = self._tf32(wp - hi)
            out.append((hi.contiguous(), lo.contiguous(), bp.contiguous()))
        return out

    @classmethod
    def from_linear(cls, actor_layers, log_std, critic_layers=None, device=None):
        """From torch.nn.Linear modules (weight [out, in]): e.g. SB3's mlp_extractor.policy_net[0], [2] and action_net."""
        t = lambda ls: None if ls is None else [(l.weight.detach().t().contiguous(), l.bias.detach()) for l in ls]   # noqa: E731
        return cls(t(actor_layers), log_std, t(critic_layers), device)

    @classmethod
    def random(cls, in_dim, out_dim, seed=0, critic=True, log_std=-0.5, device=None):
        g = torch.Generator().manual_seed(seed)

        def net(o):
            dims = [(in_dim, cls.HIDDEN), (cls.HIDDEN, cls.HIDDEN), (cls.HIDDEN, o)]
            return [((torch.rand(i, j, generator=g) * 2 - 1) / i ** 0.5, (torch.rand(j, generator=g) * 2 - 1) * 0.1) for i, j in dims]
        return cls(net(out_dim), torch.full((out_dim,), float(log_std)), net(1) if critic else None, device)

    # ---- the same network in PyTorch fp32 ---------------------------------------------------------------------------------
    @staticmethod
    def _mlp(net, x):
        h = torch.tanh(x @ net[0][0] + net[0][1])
        h = torch.tanh(h @ net[1][0] + net[1][1])
        return h @ net[2][0] + net[2][1]

    def forward_torch(self, obs, noise=None):
        """obs [E, D, obs_dim] (or [E, in_dim]) -> (raw action [E, out_dim], log-prob [E], value [E] or None)."""
        x = obs.reshape(obs.shape[0], -1).to(torch.float32)
        mean = self._mlp(self.actor, x)
        eps = torch.zeros_like(mean) if noise is None else noise.reshape(mean.shape)
        raw = mean + torch.exp(self.log_std) * eps
        logp = (-0.5 * eps * eps - self.log_std - 0.91893853320467274).sum(dim=1)
        val = None if self.critic is None else self._mlp(self.critic, x)[:, 0]
        return raw, logp, val

    # ---- C struct -------------------------------------------------------------------------------------------------------
    def c_struct(self, noise=None, logprob=None, values=None):
        q = N.QsPolicy()
        a = self._split["actor"]
        (q.w1, q.w1_lo, q.b1), (q.w2, q.w2_lo, q.b2), (q.w3, q.w3_lo, q.b3) = [(h.data_ptr(), l.data_ptr(), b.data_ptr()) for h, l, b in a]
        q.log_std = self.log_std.data_ptr()
        if self.critic is not None:
            c = self._split["critic"]
            (q.vw1, q.vw1_lo, q.vb1), (q.vw2, q.vw2_lo, q.vb2), (q.vw3, q.vw3_lo, q.vb3) = [(h.data_ptr(), l.data_ptr(), b.data_ptr()) for h, l, b in c]
        q.nt3 = self.nt3
        q.noise = None if noise is None else noise.data_ptr()
        q.logprob = None if logprob is None else logprob.data_ptr()
        q.values = None if values is None else values.data_ptr()
        q.in_dim, q.out_dim = self.in_dim, self.out_dim
        return q


_ = C
