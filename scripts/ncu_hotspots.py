"""Summarise `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` into per-file / per-line stall-sample shares.

    ncu -i gpurun_out/prof.ncu-rep --page source --print-source cuda,sass --csv > /tmp/src.csv
    python scripts/ncu_hotspots.py /tmp/src.csv [top_n] > profiles/<name>_source_hotspots.md
"""
import csv, sys, collections, os
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
files = collections.OrderedDict(); cur = None; hdr = None
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": cur = row[1]; files[cur] = []; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if cur is None or hdr is None: continue
    if row[0] != "":                       # a CUDA source line with rolled-up metrics
        files[cur].append(row)
i_samp = 4                                   # "Warp Stall Sampling (All Samples)"
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
def num(v):
    try: return float(v)
    except ValueError: return 0.0
tot = sum(num(r[i_samp]) for rows in files.values() for r in rows)
print(f"total stall samples {tot:.0f}\n")
print("| file | share |\n|---|---|")
for f, rows in sorted(files.items(), key=lambda kv: -sum(num(r[i_samp]) for r in kv[1])):
    print(f"| {os.path.basename(f)} | {100 * sum(num(r[i_samp]) for r in rows) / tot:.1f} % |")
print("\n| stall reason | share |\n|---|---|")
st = collections.Counter()
for rows in files.values():
    for r in rows:
        for i, h in stall_cols: st[h] += num(r[i])
ts = sum(st.values())
for h, v in st.most_common(10): print(f"| {h} | {100 * v / ts:.1f} % |")
print("\n| location | share | top stalls | source |\n|---|---|---|---|")
allrows = [(num(r[i_samp]), os.path.basename(f), r) for f, rows in files.items() for r in rows]
for s, f, r in sorted(allrows, key=lambda t: -t[0])[:top]:
    sr = sorted(((num(r[i]), h) for i, h in stall_cols), reverse=True)[:2]
    ss = ", ".join(f"{h[6:]} {100 * v / max(s, 1):.0f}%" for v, h in sr if v > 0)
    print(f"| {f}:{r[0]} | {100 * s / tot:.1f} % | {ss} | `{r[1].strip()[:110]}` |")
