"""Aggregate `ncu --page source --print-source cuda,sass --csv` output by CUDA source line.
usage: ncu -i rep --page source --print-source cuda,sass --csv > x.csv; python ncu_by_line.py x.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = {}
cur, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].isdigit():
        continue
    ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
    try:
        inst, smp = int(r[ie] or 0), int(r[ns] or 0)
    except ValueError:
        continue
    stalls = {h: int(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit()}
    agg[(cur, int(r[0]))] = (inst, smp, r[1].strip()[:80], stalls)
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"total warp-instructions {ti:,}  samples {ts:,}")
byfile = collections.Counter()
for (f, l), v in agg.items():
    byfile[f] += v[1]
print("samples by file:", {k: f"{100*v/ts:.1f}%" for k, v in byfile.most_common()})
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = sorted(v[3].items(), key=lambda kv: -kv[1])[:2]
    st = " ".join(f"{k[6:]}={100*n/max(v[1],1):.0f}%" for k, n in st if n)
    print(f"{100*v[1]/ts:5.1f}% smp {100*v[0]/ti:5.1f}% inst  {f}:{l:<4d} {v[2]:80s} {st}")
