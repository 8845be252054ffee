"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export by the device function of searcher.cuh
(or by file for everything else): warp-instructions and stall samples per distance evaluation.
  python profiles/tools/ncu_by_function.py src.csv raw.csv n_queries evals_per_query [searcher.cuh path]"""
import csv
import re
import subprocess
import sys
import os

src_csv, raw_csv, nq, epq = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
here = os.path.dirname(os.path.abspath(__file__))
cuh = sys.argv[5] if len(sys.argv) > 5 else os.path.join(here, "..", "..", "kektordb_b200", "csrc", "searcher.cuh")
lines = open(cuh).read().split("\n")
funcs = []
for i, l in enumerate(lines, 1):
    m = re.search(r"__device__.*?(\w+)\(", l)
    if m:
        funcs.append((i, m.group(1)))


def fn(line):
    name = "?"
    for s, n in funcs:
        if s <= line:
            name = n
    return name


out = subprocess.run([sys.executable, os.path.join(here, "ncu_by_line.py"), src_csv, "100000"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))[1:]
agg, samp, tot, tots = {}, {}, 0, 0
for r in rows:
    if len(r) < 8:
        continue
    key = fn(int(r[1])) if r[0] == "searcher.cuh" else r[0]
    agg[key] = agg.get(key, 0) + int(r[2])
    samp[key] = samp.get(key, 0) + int(r[4])
    tot += int(r[2])
    tots += int(r[4])
raw = list(csv.reader(open(raw_csv)))
d = dict(zip(raw[0], raw[2]))
print(f"kernel {d.get('Kernel Name')}  duration {d.get('gpu__time_duration.sum')} ms  grid {d.get('launch__grid_size')}  regs {d.get('launch__registers_per_thread')}")
print(f"warps active {float(d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0)):.1f} % of 64/SM, issue slots busy {float(d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0)):.1f} %, "
      f"dram read {d.get('dram__bytes_read.sum')} {raw[1][raw[0].index('dram__bytes_read.sum')]}")
ev = nq * epq
print(f"warp-instructions {tot}  = {tot / ev:.1f} per distance evaluation ({nq} queries x {epq} evaluations)")
print(f"{'function / file':30s} {'instr %':>8s} {'per eval':>9s} {'stall samples %':>16s}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1])[:24]:
    print(f"{key:30s} {100 * v / tot:8.2f} {v / ev:9.1f} {100 * samp[key] / max(tots, 1):16.2f}")
