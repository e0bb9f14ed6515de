"""Regenerates profiles/README.md from the latest bench / config / ncu artefacts (run in the build container)."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bench = sys.argv[1] if len(sys.argv) > 1 else sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "bench*.json")), key=os.path.getmtime)[-1]
b = json.load(open(bench))
ref = None
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "bench*_ref.json")), key=os.path.getmtime):
    ref = json.load(open(f))
cfg = {}
if os.path.isfile(os.path.join(ROOT, "gpurun_out", "configs_r01.json")):
    cfg = json.load(open(os.path.join(ROOT, "gpurun_out", "configs_r01.json")))
    json.dump(cfg, open(os.path.join(ROOT, "profiles", "r01_configs.json"), "w"), indent=1)
json.dump(b, open(os.path.join(ROOT, "profiles", "r01_bench_n1.json"), "w"), indent=1)
if ref:
    json.dump(ref, open(os.path.join(ROOT, "profiles", "r01_bench_reference_arm.json"), "w"), indent=1)
refcpu = json.load(open(os.path.join(ROOT, "profiles", "reference_cpu_r01.json")))
L = []
L.append("# profiles/ — round 1 measurements (B200, sm_100a)\n")
L.append("Everything here is copied from `gpurun_out/` (scratch) by `tools/make_profiles_readme.py` / `tools/summarize_profile.py`.\n")
L.append("## Headline bench line (`python bench.py`, N=1; file `r01_bench_n1.json`)\n")
r = b["roofline"]
L.append("| quantity | value |\n|---|---|")
L.append("| workload | %s |" % b["config"]["workload"])
L.append("| `value` (device-resident) | **%.3g drone-steps/s** (%.2f us per 65 536-drone step, %d steps) |" % (b["value"], b["ms_per_step"] * 1e3, b["steps"]))
L.append("| `e2e` (NumPy API, H2D %d B + D2H %d B per step) | **%.3g drone-steps/s** |" % (b["e2e"]["h2d_bytes_per_step"], b["e2e"]["d2h_bytes_per_step"], b["e2e"]["value"]))
L.append("| roofline (HBM, %s) | achieved %.0f GB/s of %.0f GB/s measured = **%.3f** (646 B x 65 536 / %.2f us) |" % (r.get("timing", "event pairs"), r["achieved"], r["peak"], r["frac"], r["kernel_ms"] * 1e3))
L.append("| clocks during the timed region | %s |" % json.dumps(b.get("clocks")))
if "cpu_baseline" in b:
    L.append("| `cpu_baseline` (oracle port, 1 core of the GPU box) | %.3g drone-steps/s — %s |" % (b["cpu_baseline"]["value"], b["cpu_baseline"]["sample"]))
if ref:
    L.append("| `--impl reference` (oracle port, %d processes) | %.3g drone-steps/s |" % (ref["cpu_baseline"]["cores"], ref["value"]))
L.append("| real reference, build container | %.0f drone-steps/s on one core (MultiHover D=2, S=8), %.0f on %d cores (`reference_cpu_r01.json`) |" %
         (refcpu["cases"][2]["drone_steps_per_s"], refcpu["fan_out"]["drone_steps_per_s"], refcpu["fan_out"]["processes"]))
ex = b.get("extras", {})
L.append("\n### extras of the same run\n")
L.append("| item | ms per step | drone-steps/s | algorithmic-bytes fraction of HBM peak |\n|---|---|---|---|")
if "cuda_graph_replay" in ex and "ms_per_step" in ex["cuda_graph_replay"]:
    g = ex["cuda_graph_replay"]
    L.append("| same launches replayed from a CUDA graph | %.4f | %.3g | %.3f |" % (g["ms_per_step"], g["value"], 646 * 65536 / (g["ms_per_step"] * 1e-3) / 1e9 / r["peak"]))
if "fused_rollout_T32" in ex and "ms_per_step" in ex["fused_rollout_T32"]:
    g = ex["fused_rollout_T32"]
    L.append("| `qs_rollout`, 32 ticks per launch | %.4f | %.3g | %.3f (algorithmic; the fusion removes the history re-read and the state round trip) |" % (g["ms_per_step"], g["value"], g["hbm_frac_algorithmic"]))
for n, v in ex.get("drones_per_launch_sweep", {}).items():
    if "ms_per_step" in v:
        L.append("| step kernel at %s drones per launch | %.4f | %.3g | %.3f |" % (n, v["ms_per_step"], v["value"], v["hbm_frac"]))
if "host_us_per_step_call" in ex:
    L.append("\nHost cost of one `env.step(tensor)` call (64-drone env, GPU idle): %.1f us." % ex["host_us_per_step_call"])
if cfg:
    L.append("\n## Other BASELINE.json configs (`tools/bench_configs.py`, file `r01_configs.json`)\n")
    L.append("| config | drones | S | ms per step | drone-steps/s | alg. bytes | HBM frac | note |\n|---|---|---|---|---|---|---|---|")
    for k, v in cfg.items():
        if "drone_steps_per_s" in v:
            L.append("| %s | %s | %s | %.4f | %.3g | %s | %s | %s |" % (k, v.get("drones"), v.get("S"), v["ms_per_step"], v["drone_steps_per_s"], v.get("alg_bytes", "-"),
                                                                   ("%.3f" % v["hbm_frac"]) if "hbm_frac" in v else ("%.3g pairs/s" % v["pairs_per_s"] if "pairs_per_s" in v else "-"), v.get("note", "")))
        else:
            L.append("| %s | n=%s | - | %.4f | %.3g calls/s | %s | %.3f | |" % (k, v["n"], v["ms_per_call"], v["calls_per_s"], v["alg_bytes"], v["hbm_frac"]))
L.append("""
## Optimisation log of the step kernel (65 536 drones, S=8, A=4, B=15; `bench.py` `ms_per_step`, each measured on a B200 via gpurun)

| commit milestone | us / step | what changed (evidence) |
|---|---|---|
| first parity-green kernel | 102.6 | one LDG->STG round trip per observation row (`r01_a_first_correct_kernel_ncu.md`: stall_long_sb on the row copy) |
| L2 bulk prefetch + 8-deep unrolled float4 row copies | 38.8 | `cp.async.bulk.prefetch.L2` of the CTA's history span |
| division-free FP64 tick, rpy in f32 | 31.6 | 2/|q|^2 -> 2(2-|q|^2), 1/M hoisted; 12.3k -> 7.6k instructions per warp |
| TMA bulk copy of the span into shared memory behind the physics | 24.2 | `UBLKCP.S.G` + mbarrier; Python fast path (host 13.8 -> 8.5 us per call) |
| flat shifted-span writer (rows patched in shared memory) | 21.6 | 3.8k -> 2.2k instructions per warp |
| state loads ahead of the bulk copy, LDS (not generic LD) copy-out | 18.7 | `r01_f_step_kernel_ncu.md` |
| 5-term series below 42 rad/s, fused Euler constants, diagonal-J gyro term | 17.7 | 835 -> ~640 FP64 instructions per tick (A/B on one box: 18.1 -> 17.7) |
| 64-drone, then one-warp (32-drone) CTAs | 16.2 | A/B on one box: 32/64/128-drone CTAs = 16.2/16.1/17.3 us; 1 M drones 161/162/169 us |
| programmatic dependent launch (griddepcontrol) | 14.3 | A/B on one box: 16.24 -> 14.20 us |
| terminal-observation rows by warp ballot instead of a serial flag scan | 13.1 | `LDS.U8 -> LOP3 -> BRA` chain was 25 % of the samples; 1 M drones 159 -> 143 us |
| TMA bulk store (`cp.async.bulk` S2G, `UBLKCP.G.S`) of the observation span | 12.4 | A/B on one box: 13.08 -> 12.42 us; rollout 7.57 -> 7.21 us per tick |

Tried and rejected (measured): issuing the TMA copy before the state loads (ncu 16.1 vs 14.8 us); writing the history
columns out before the physics to overlap stores with FP64 work (22.1 vs 18.7 us: the physics then waits for the bulk
copy); trimming the step kernel's fixed shared memory so 18 instead of 15 CTAs fit per SM (12.44 -> 12.64 us at 65 536
drones, 138.6 -> 136.0 us at 1 M: not adopted).  Box-to-box spread of the same binary is up to ~10 % at 1 M drones (158-176 us), so only same-box A/B numbers are
compared.
""")
fpath = os.path.join(ROOT, "profiles", "r01_formation.json")
if os.path.isfile(fpath):
    F = json.load(open(fpath))
    L.append("""
## Formations: downwash with exact chunk culling, neighbourhood query, one formation over several GPUs (`r01_formation.json`)

Downwash force kernel on the config-4 geometry (128 x 128 grid, 0.15 m pitch, z = 0.1 + 0.05 (i mod 16)), reset positions
(`r01_configs.json: config4_formation_16384_gnd_drag_dw.downwash_ms`); every variant returns the SAME bits as its own
all-pairs evaluation (`tests/test_gpu_formation.py::test_downwash_culling_is_exact`):

| kernel | all pairs | culled, row-major order | culled, Morton order |
|---|---|---|---|""")
    dm = cfg.get("config4_formation_16384_gnd_drag_dw", {}).get("downwash_ms", {})
    if dm:
        L.append("| `qs_downwash` (1024-source tiles in shared memory) | %.0f us | %.0f us | %.0f us |" % (dm["tiled_kernel_all_pairs"] * 1e3, dm["tiled_kernel_culled_row_major"] * 1e3, dm["tiled_kernel_culled_morton_order"] * 1e3))
        L.append("| `qs_downwash_boxed` (box table, 32-row CTAs; incl. the 4 us boxes kernel) | %.0f us | %.0f us | %.0f us |" % (dm["all_pairs"] * 1e3, dm["culled_row_major"] * 1e3, dm["culled_morton_order"] * 1e3))
    L.append("""
First version of the kernel (IEEE division + expf, 256-source tiles, no culling): 502 us = 5.3e11 pairs/s.  SFU
reciprocal/exp2 + per-pair underflow early-out: 207 us all pairs (1.3e12 pairs/s); with chunk culling 31-33 us.
65 536-drone formation (256 x 256 grid, Morton order): %s us per evaluation.

One formation sharded over GPUs (`tools/formation_multi_gpu.py`, one process per GPU; per physics substep: exchange of
positions + boxes, then the force of the local rows; CUDA events, max over ranks; `stage` = exchange + force,
measured on the reset geometry):

| drones | GPUs | bit-identical to the unsharded run (10 ticks x 5 substeps, gnd + drag + downwash) | stage, NCCL all-gather + boxes kernel | stage, push + flags over NVLink peer memory | parts: NCCL all-gather / publish kernel / force kernel |
|---|---|---|---|---|---|""" % ("%.0f" % (cfg.get("formation_65536_downwash_culled_morton", {}).get("ms_per_step", float("nan")) * 1e3)))
    for key in sorted(k for k in F if k.startswith("world")):
        d = F[key]
        bit = "yes (NCCL and p2p)" if d.get("nccl_bit_identical_to_unsharded") and d.get("p2p_bit_identical_to_unsharded") else ("not checked (timing only)" if "p2p_bit_identical_to_unsharded" not in d else "NO")
        L.append("| %d | %d | %s | %.1f us | %.1f us%s | %.1f / %.1f / %.1f us |" % (
            d["drones"], d["world"], bit, d["stage_us_nccl"], d["stage_us_p2p"],
            (" (unsharded single launch path: %.1f us)" % d["stage_us_local"]) if "stage_us_local" in d else "",
            d["part_us_nccl_all_gather"], d["part_us_publish_kernel"], d["part_us_downwash_rows_kernel"]))
    a = F.get("adjacency_final")
    if a:
        L.append("\n`qs_adjacency` (BaseAviary._getAdjacencyMatrix), 16 384 drones, radius 1 m: %.0f us for the 268 MB matrix = %.0f GB/s written (%.2f of the %.0f GB/s copy peak; instruction-bound: ~13 instructions per 32 pairs)." % (a["adjacency_16384_ms"] * 1e3, a["GBps_written"], a["GBps_written"] / r["peak"], r["peak"]))

L.append("\n## ncu captures (`ncu --set full --clock-control none --import-source on`, one GPU, `bench.py --steps 20`)\n")
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_*_ncu.md"))):
    L.append("* `%s` — %s" % (os.path.basename(f), open(f).read().split("\n")[2]))
L.append("* `r01_step_kernel_stall_breakdown.md` — where a warp's time goes in the step kernel at the bench size (stall reasons, opcode classes, program regions) and why the launch is latency-bound there.")
L.append("* `r01_a_launches.csv` — launch list of the first correct kernel (setup part only).")
try:
    import collections, csv
    rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r01_z_launches_final.csv"))) if len(r) > 14 and r[0].isdigit()]
    cnt, tot = collections.Counter(), collections.Counter()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[-60:]
        cnt[name] += 1
        tot[name] += float(r[-1])
    allt = sum(tot.values())
    L.append("* `r01_z_launches_final.csv` — `ncu --metrics gpu__time_duration.sum --clock-control none` over a whole `bench.py --steps 40` run (%d launches; cold-cache, serialised: shares, not absolutes):\n" % len(rows))
    L.append("  | kernel | launches | share of GPU time |\n  |---|---|---|")
    for k, v in tot.most_common(6):
        L.append("  | `%s` | %d | %.1f %% |" % (k, cnt[k], 100 * v / allt))
except Exception as ex:
    L.append("* launch list: %r" % (ex,))
L.append("\n`traffic` in the bench JSON: ncu's `dram__bytes_read.sum` is 24.3 MB per launch (state 4.2 MB + actions 1 MB + old observation span 18.9 MB = the algorithmic reads); `dram__bytes_write.sum` reads ~0 inside the kernel because the 23 MB of stores are still in the 126 MB write-back L2 when the kernel ends.")
open(os.path.join(ROOT, "profiles", "README.md"), "w").write("\n".join(L) + "\n")
print("\n".join(L))
