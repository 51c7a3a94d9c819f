"""Per-source-line sample counts of an ncu report: python scratch/ncu_lines.py rep [file-substr]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ".cu"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur_file = None; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or not r or not r[0]: continue
    try: n = int(r[hdr["# Samples"]] or 0)
    except ValueError: continue
    if n and cur_file and sub in cur_file:
        stalls = {k[6:]: int(r[i] or 0) for k, i in hdr.items() if k.startswith("stall_") and "Not Issued" not in k}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        out.append((n, cur_file.split("/")[-1], r[0], r[1].strip()[:110], top))
tot = sum(o[0] for o in out)
print("total samples", tot)
for o in sorted(out, key=lambda o: -o[0])[:28]:
    print(f"{o[0]:6d} {o[1]}:{o[2]:>4} {o[3]}   {o[4]}")
