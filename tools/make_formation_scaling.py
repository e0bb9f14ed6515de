"""profiles/r02_formation_scaling.{json,md} from gpurun_out/formation_multi_gpu_<world>_<drones>.json (tools/formation_scaling.sh)."""
import json, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sizes, worlds = (16384, 65536, 262144), (1, 2, 4, 8)
data = {}
for n in sizes:
    for w in worlds:
        p = os.path.join(root, "gpurun_out", "formation_multi_gpu_%d_%d.json" % (w, n))
        if os.path.isfile(p):
            data["world%d_drones%d" % (w, n)] = json.load(open(p))
json.dump(data, open(os.path.join(root, "profiles", "r02_formation_scaling.json"), "w"), indent=1)
L = ["# Round 2: one formation sharded by drones over the GPUs of a node (strong scaling)", "",
     "`tools/formation_scaling.sh <world>` under `gpurun --gpus <world>` (one process per GPU, torchrun).  Stacks geometry (stacks of 4,",
     "1.6 m pitch, 1.5 m between layers), Physics.PYB_GND_DRAG_DW, 240 Hz / 48 Hz = 5 substeps per control tick, hover RPM, 5 ticks",
     "from reset; CUDA events, max over ranks.  `tick` = the whole control tick of the formation: per substep the dynamics kernel,",
     "the position exchange (p2p = push + flag kernels over NVLink peer memory; nccl = all-gather + boxes kernel) and the boxed",
     "downwash kernel over this rank's rows.  Every sharded run was first checked bit-identical to the unsharded formation on rank 0",
     "(6 ticks, both exchange modes): column `identical`.", "",
     "| drones | GPUs | tick us (p2p) | speed-up vs 1 GPU | tick us (nccl) | stage us (p2p) | rows kernel us | publish kernel us | all-gather us | identical |",
     "|---|---|---|---|---|---|---|---|---|---|"]
for n in sizes:
    base = data.get("world1_drones%d" % n, {}).get("tick_us_local")
    for w in worlds:
        d = data.get("world%d_drones%d" % (w, n))
        if not d:
            continue
        t = d["tick_us_local"] if w == 1 else d["tick_us_p2p"]
        ident = d.get("p2p_bit_identical_to_unsharded") and d.get("nccl_bit_identical_to_unsharded")
        L.append("| %d | %d | %.1f%s | %.2fx | %.1f | %.1f | %.1f | %.1f | %.1f | %s |" % (
            n, w, t, " (local)" if w == 1 else "", base / t, d["tick_us_nccl"], d["stage_stacks_us_p2p"], d["part_us_downwash_rows_kernel"],
            d["part_us_publish_kernel"], d["part_us_nccl_all_gather"], "yes" if ident else "NO"))
L += ["", "Reading: the pair kernel (rows) shards with the GPUs -- 93 -> 55 -> 35 -> 27 us at 65 536 drones, 601 -> 314 -> 172 -> 95 us at 262 144 -- while",
      "the exchange does not: the publish kernel writes this rank's positions and chunk boxes into EVERY peer's buffer and raises a flag",
      "there (6.6 us alone, 14.5 us with 4 peers, 21-27 us with 8), and the dynamics kernel (8.5 us per substep) and five launches per",
      "substep are fixed.  So 16 384 drones (32 us stage) do not scale at all, 65 536 drones reach 2.06x on 4 GPUs and 2.17x on 8 (limited by",
      "publish + dynamics + launch latency: ~45 of the 56 us per substep at 8 GPUs), 262 144 drones 3.06x on 4 and 4.91x on 8 (the O(N/32)",
      "box tests per row group are the part of the rows kernel that does not shrink).  The p2p exchange beats NCCL's all-gather + boxes",
      "kernel by 7-12 % of the tick at every size.  rows/publish/all-gather columns are the stand-alone kernels on the config-4 lattice",
      "(reset positions) as in round 1; `stage` is exchange + downwash on the stacks geometry.",
      "", "Why not the config-4 lattice for whole ticks: with 0.15 m between drones of initially equal height the reference's downwash term",
      "alpha = c1 (r / (4 dz))^2 (BaseAviary.py:803) is singular once rounding makes dz a tiny positive number; the drones are thrown apart",
      "and nothing culls any more (measured: 92 us -> 1.1 ms per stage after two substeps, gpurun_out/formation_tick_kernels.csv).  That is",
      "the model, reproduced faithfully; it is not a meaningful timing workload."]
open(os.path.join(root, "profiles", "r02_formation_scaling.md"), "w").write("\n".join(L) + "\n")
print("\n".join(L[8:26]))
