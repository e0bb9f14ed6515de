"""Turns a gpurun_out/*.ncu-rep (ncu --set full) into the tracked text summary under profiles/.
    python tools/summarize_profile.py gpurun_out/prof_r01_f.ncu-rep profiles/r01_step_kernel_ncu.md "note"
"""
import collections
import csv
import io
import subprocess
import sys

rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines = ["# ncu summary: %s" % rep, "", note, ""]
for r in rows[2:]:
    lines.append("## %s" % r[hdr.index("Kernel Name")])
    lines.append("")
    lines.append("| metric | value | unit |")
    lines.append("|---|---|---|")
    for k in KEYS:
        if k in hdr:
            lines.append("| %s | %s | %s |" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
sh = srows[1]
ix = {h: i for i, h in enumerate(sh)}
data = [r for r in srows[2:] if len(r) >= len(sh) and r[ix["# Samples"]].isdigit()]
stall_cols = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
ops = collections.Counter()
nwarps = None
for r in data:
    for c in stall_cols:
        tot[c] += int(r[ix[c]] or 0)
    toks = r[ix["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
ts = sum(int(r[ix["# Samples"]]) for r in data)
lines += ["## warp-state samples (source page, all launches in the report)", "", "total samples: %d" % ts, "",
          "| stall reason | samples |", "|---|---|"]
lines += ["| %s | %d |" % (k, v) for k, v in tot.most_common(10)]
lines += ["", "## executed warp-instructions by opcode (top 16)", "", "| opcode | warp-instructions |", "|---|---|"]
lines += ["| %s | %d |" % (k, v) for k, v in ops.most_common(16)]
hot = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:12]
lines += ["", "## hottest SASS lines", "", "| samples | SASS | main stall |", "|---|---|---|"]
for r in hot:
    st = max(stall_cols, key=lambda c: int(r[ix[c]] or 0))
    lines.append("| %s | `%s` | %s |" % (r[ix["# Samples"]], r[ix["Source"]].strip()[:90], st))
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
