"""qs_adjacency timing (BaseAviary._getAdjacencyMatrix, BaseAviary.py:658-675), every kernel version in one process (QS_ADJ_V is
read per call):

    python tools/adjacency_ab.py [versions, default 1,2,3,4,5]

    1  first version: scalar float32, FSETP + SEL packing
    2  packed float32 (FADD2 / FMUL2 / FFMA2), sign-bit decisions packed by PRMT
    3  every second column pair with scalar instructions (the packed ones issue to the FMA-heavy pipe only)
    4, 5  versions 2, 3 with registers capped for 3 CTAs per SM

One 16 384-drone aviary (268 MB matrix per query): the BASELINE config-4 lattice and uniformly random positions.  Every version's
matrix is compared bit for bit with version 1's.  Appends one JSON line per version to gpurun_out/adjacency_ab.jsonl.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gym_pybullet_drones_b200.envs import CtrlAviary
from gym_pybullet_drones_b200.utils.enums import Physics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = 6572.5
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, n, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


versions = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,2,3,4,5").split(",")]
Dn = 16384
i = np.arange(Dn)
lattice = np.stack([0.15 * (i % 128), 0.15 * (i // 128), 0.1 + 0.05 * (i % 16)], axis=1)
rnd = np.random.default_rng(5).uniform(-6, 6, (Dn, 3))
# two output matrices used alternately: 2 x 268 MB > 126 MB L2, every query writes to HBM
bufs = [torch.empty((1, Dn, Dn), dtype=torch.uint8, device="cuda") for _ in range(2)]
envs = {}
for name, xyz in (("config4_lattice", lattice), ("uniform_random", rnd)):
    envs[name] = CtrlAviary(num_drones=Dn, initial_xyzs=xyz, physics=Physics.DYN, neighbourhood_radius=1.0, num_envs=1)
    envs[name].reset()
ref = {}
os.environ["QS_ADJ_V"] = "1"
for name, env in envs.items():
    ref[name] = env.adjacency().clone()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for v in versions:
    os.environ["QS_ADJ_V"] = str(v)
    out = {"version": v, "drones": Dn, "bytes_per_query": Dn * Dn, "peak_gbs": PEAK}
    for name, env in envs.items():
        k = [0]

        def q():
            k[0] ^= 1
            env.adjacency(bufs[k[0]])

        ms = timed(q, 40)
        out[name] = {"ms_per_query": ms, "written_gbs": Dn * Dn / (ms * 1e-3) / 1e9, "frac_of_copy_peak": Dn * Dn / (ms * 1e-3) / 1e9 / PEAK,
                     "ones": int(bufs[k[0]].sum(dtype=torch.int64).item()), "equals_version_1": bool(torch.equal(bufs[k[0]], ref[name]))}
    with open(os.path.join(ROOT, "gpurun_out", "adjacency_ab.jsonl"), "a") as f:
        f.write(json.dumps(out) + "\n")
    print(json.dumps(out))
