"""Summarise an .ncu-rep (one kernel) into a small text file for profiles/: duration, tensor-pipe
and DRAM/L2 metrics, registers, and the top stall reasons from the source page."""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "launch__shared_mem_per_block_dynamic"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    lines.append(f"kernel: {d.get('Kernel Name', '?')}")
    for k in KEYS:
        if k in d:
            lines.append(f"  {k} = {d[k]} {u.get(k, '')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr, data = rows[1], rows[2:]
    ins = hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    data = [r for r in data if len(r) > max(stall_cols + [ins])]  # (multi-kernel reports repeat the header block)

    def _int(x):
        try:
            return int(x or 0)
        except ValueError:
            return 0
    for r in data:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + _int(r[i])
    tot = sum(_int(r[ins]) for r in data)
    lines.append(f"  warp-state samples: {tot}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        lines.append(f"    {k}: {v} ({100.0 * v / max(tot, 1):.1f} %)")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
