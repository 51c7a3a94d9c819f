"""Summarise an ncu --set full report into the handful of numbers DESIGN.md / bench.py quote.
    python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep > profiles/r1_x_summary.txt"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|launch__registers_per_thread|launch__grid_size|"
    r"launch__block_size|launch__cluster_size|sm__cycles_elapsed.max|smsp__inst_executed.sum|"
    r"sm__warps_active.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct_of_peak_sustained_active|"
    r"sm__throughput.avg.pct_of_peak_sustained_elapsed|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|"
    r"lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|"
    r"sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_(xu|alu|fma|lsu|tc|tma|tmem).avg.pct_of_peak_sustained_active|"
    r"l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed|"
    r"l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed|"
    r"l1tex__m_xbar2l1tex_read_bytes.sum|lts__t_bytes.sum)$")

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    name = row[hdr.index("Kernel Name")]
    print(f"== {name}")
    for h, u, v in zip(hdr, units, row):
        if KEEP.match(h):
            print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: 0 for s in stalls}
    n = 0
    for r in rows[2:]:
        n += int(r[ix["# Samples"]] or 0)
        for s in stalls:
            tot[s] += int(r[ix[s]] or 0)
    print(f"  warp-state samples {n}: " + ", ".join(f"{k[6:]} {100 * v / max(n, 1):.0f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:6]))
    print("  hottest instructions (samples, SASS):")
    for r in sorted(rows[2:], key=lambda r: -int(r[ix["# Samples"]] or 0))[:8]:
        print(f"    {r[ix['# Samples']]:>7}  {r[ix['Source']].strip()[:90]}")
