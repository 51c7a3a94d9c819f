"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
    python profiles/summarize_launches.py profiles/r1_launches_full_step_final.csv > profiles/r1_launches_full_step_final_summary.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[ik])
    name = re.sub(r"\(.*$", "", name).replace("sgn::", "")
    v = float(r[iv].replace(",", ""))
    ms = v / 1e6 if r[iu] == "ns" else (v / 1e3 if r[iu] in ("us", "usecond") else v)
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none, ONE timed step of bench.py (cudaProfilerStart/Stop range)")
print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's by_kernel, not absolutes")
print(f"# launches {sum(cnt.values())}  total {total:.2f} ms")
print("kernel,launches,total_ms,share_pct")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k},{cnt[k]},{tot[k]:.3f},{100 * tot[k] / total:.1f}")
