"""Aggregates `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by CUDA source line:
warp-instructions executed and stall samples per (file, line), with the dominant stall reasons.
  python profiles/tools/ncu_by_line.py src.csv [top_n] > by_line.csv"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rows = list(csv.reader(open(path, errors="replace")))
cur_file = None
hdr = None
agg = defaultdict(lambda: defaultdict(float))
text = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    d = {}
    for i, h in enumerate(hdr):
        d.setdefault(h, r[i])  # first "Source" = CUDA line text, second = SASS
    key = (cur_file, int(r[0]))
    text[key] = r[1].strip()
    a = agg[key]
    try:
        a["inst"] += float(d.get("Instructions Executed") or 0)
        a["samples"] += float(d.get("# Samples") or 0)
    except ValueError:
        continue
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                a[h] += float(d[h] or 0)
            except ValueError:
                pass
tot_i = sum(a["inst"] for a in agg.values()) or 1
tot_s = sum(a["samples"] for a in agg.values()) or 1
w = csv.writer(sys.stdout)
w.writerow(["file", "line", "warp_instructions", "pct_instructions", "stall_samples", "pct_samples", "top_stalls", "source"])
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((v, k) for k, v in a.items() if k.startswith("stall_")), reverse=True)[:3]
    w.writerow([key[0], key[1], int(a["inst"]), round(100 * a["inst"] / tot_i, 2), int(a["samples"]),
                round(100 * a["samples"] / tot_s, 2), " ".join(f"{k[6:]}:{int(v)}" for v, k in st if v), text[key][:110]])
print(f"# total warp instructions {int(tot_i)}, stall samples {int(tot_s)}", file=sys.stderr)
