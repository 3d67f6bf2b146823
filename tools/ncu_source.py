"""Per-source-line instruction and stall-sample totals from an `ncu --page source --csv --print-source cuda,sass` export.

    python tools/ncu_source.py file.csv "k_step3d_t" [top]
"""
import csv
import sys
from collections import defaultdict

path, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fn = None
fpath = None
hdr = None
agg = defaultdict(lambda: [0, 0, ""])   # (file,line) -> [inst, samples, text]
tot = [0, 0]
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1]; continue
    if row[0] == "Function Name":
        fn = row[1]; continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None or fn is None or pat not in fn:
        continue
    try:
        ln = int(row[0])
    except ValueError:
        continue
    d = dict(zip(hdr, row))
    try:
        inst = int(d.get("Instructions Executed", "0") or 0)
        smp = int(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    key = (fpath.split("/")[-1], ln, fn[:30])
    agg[key][0] += inst; agg[key][1] += smp; agg[key][2] = row[1][:110]
    tot[0] += inst; tot[1] += smp
print("total inst", tot[0], "samples", tot[1])
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{key[0]}:{key[1]:4d} inst={100*v[0]/max(tot[0],1):5.1f}% smp={100*v[1]/max(tot[1],1):5.1f}% | {v[2].strip()}")
