"""Times the UNMODIFIED reference's Physics.DYN CPU path (tier-1 oracle, stand-in pybullet) in the build container
and writes profiles/reference_cpu_r01.json.  The reference is pure Python and single threaded; the optional fan-out runs
one env per core.  PyBullet's own solver (Physics.PYB, the baseline BASELINE.json names) cannot be timed: the `pybullet`
package is not installed in this image and there is no network -- reported as such, never invented.

    python oracle/time_reference.py
"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def _run(args):
    kind, nd, cf, act, steps = args
    from oracle.ref_loader import load_reference, quiet
    R = load_reference()
    A = {"rpm": R.ActionType.RPM, "one_d_rpm": R.ActionType.ONE_D_RPM, "pid": R.ActionType.PID}[act]
    aw = {"rpm": 4, "one_d_rpm": 1, "pid": 3}[act]
    with quiet():
        if kind == "hover":
            env = R.HoverAviary(physics=R.Physics.DYN, pyb_freq=240, ctrl_freq=cf, act=A)
        else:
            env = R.MultiHoverAviary(num_drones=nd, physics=R.Physics.DYN, pyb_freq=240, ctrl_freq=cf, act=A)
        env.reset()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (steps, nd, aw)).astype(np.float32)
    if act == "pid":
        acts = (np.array([0, 0, 1], np.float32) + 0.3 * acts).astype(np.float32)
    t0 = time.perf_counter()
    for t in range(steps):
        obs, r, te, tr, _ = env.step(acts[t])
        if te or tr:
            env.reset()
    return time.perf_counter() - t0


def main():
    cases = [("hover", 1, 240, "rpm", 2000), ("hover", 1, 30, "one_d_rpm", 500), ("multihover", 2, 30, "rpm", 300),
             ("multihover", 16, 30, "rpm", 60), ("multihover", 128, 30, "rpm", 10), ("hover", 1, 30, "pid", 300)]
    out = {"host": platform.processor() or platform.machine(), "cpu_count": os.cpu_count(), "numpy": np.__version__,
           "what": "unmodified reference (gym_pybullet_drones @ /root/reference) Physics.DYN via oracle/standins, one core",
           "pybullet_PYB_baseline": "PyBullet unavailable: PYB baseline not measured", "cases": []}
    try:
        import pybullet  # noqa: F401
        if "standins" not in pybullet.__file__:
            out["pybullet_PYB_baseline"] = "real pybullet importable: time Physics.PYB separately"
    except Exception:
        pass
    for kind, nd, cf, act, steps in cases:
        dt = _run((kind, nd, cf, act, steps))
        out["cases"].append({"env": kind, "drones": nd, "pyb_freq": 240, "ctrl_freq": cf, "substeps": 240 // cf, "act": act, "steps": steps,
                             "seconds": dt, "ms_per_step": 1e3 * dt / steps, "drone_steps_per_s": nd * steps / dt})
        print(out["cases"][-1])
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        dts = pool.map(_run, [("multihover", 2, 30, "rpm", 300)] * cores)      # timed inside each worker (imports excluded)
        wall = max(dts)
    out["fan_out"] = {"processes": cores, "env": "multihover", "drones": 2, "steps_each": 300, "slowest_worker_s": wall,
                      "drone_steps_per_s": cores * 2 * 300 / wall}
    print(out["fan_out"])
    path = os.path.join(os.path.dirname(HERE), "profiles", "reference_cpu_r01.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
