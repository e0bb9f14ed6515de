"""A/B timing of the step kernels on one GPU (tools, not product): us per step for the bench workload and variants.

    python tools/ab_fast.py                 # sweeps QS_FAST / QS_FAST_WARPS in subprocesses
    python tools/ab_fast.py --one           # one configuration (the environment decides)

Pipelined = K back-to-back launches / K (programmatic dependent launch overlaps neighbours); isolated = events around single
launches separated by a stream sync and an L2-sized scrub (median)."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one():
    import torch
    from gym_pybullet_drones_b200.envs import MultiHoverAviary
    from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
    dev = torch.device("cuda:0")
    D = 2
    g = torch.Generator(device=dev).manual_seed(0)
    out = {}
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for act, A in (("RPM", 4), ("ONE_D_RPM", 1)):
        for n in ((65536,) if os.environ.get("QS_AB_SMALL") else (65536, 1048576)):
            R = 8 if n == 65536 else 2
            envs = [MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType[act], num_envs=n // D, autoreset="same_step") for _ in range(R)]
            acts = [torch.rand((n // D, D, A), device=dev, generator=g) * 2 - 1 for _ in envs]
            for e in envs:
                e.reset()
            for k in range(40):
                envs[k % R].step(acts[k % R])
            torch.cuda.synchronize()
            best = 1e9
            K = 400 if n == 65536 else 60
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for k in range(K):
                    envs[k % R].step(acts[k % R])
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K * 1e3)
            iso = []
            for k in range(30):
                scrub.fill_(k & 255)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                envs[k % R].step(acts[k % R])
                e1.record()
                torch.cuda.synchronize()
                iso.append(e0.elapsed_time(e1) * 1e3)
            iso.sort()
            out["%s_%d" % (act, n)] = {"pipelined_us": round(best, 2), "isolated_us_median": round(iso[len(iso) // 2], 2), "isolated_us_min": round(iso[0], 2)}
            del envs, acts
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("QS_")}, **out}))


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
    else:
        for env in ({"QS_STAGGER_NS": "0"}, {"QS_STAGGER_NS": "1000", "QS_STAGGER_MODE": "0"}, {"QS_STAGGER_NS": "1500", "QS_STAGGER_MODE": "0"},
                    {"QS_STAGGER_NS": "1500", "QS_STAGGER_MODE": "1"}, {"QS_STAGGER_NS": "2500", "QS_STAGGER_MODE": "1"}, {"QS_STAGGER_NS": "1500", "QS_STAGGER_MODE": "2"}):
            e = dict(os.environ)
            e.update(env)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=e, check=False)
