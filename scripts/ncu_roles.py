"""Aggregate the source page of an ncu report of the frame kernel by warp role: stall samples per SASS region.
usage: python scripts/ncu_roles.py report.ncu-rep   (regions are found from marker instructions, not hard-coded)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2]
for k in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
          "launch__cluster_dim_x", "gpc__cycles_elapsed.max", "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct"):
    if k in hdr:
        print(f"{k:70s} {vals[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
data = rows[2:]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
# top instructions
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    n = int(r[ix["# Samples"]])
    st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{data.index(r):5d} {n:6d} {100 * n / tot:5.1f}% exec={r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:64]:64s} {st}")
