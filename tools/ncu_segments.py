#!/usr/bin/env python
"""Warp-stall samples of an .ncu-rep by code segment: the SASS is cut at barriers, mbarrier waits and tensor-memory
instructions, and every segment is printed with its share of the samples, its instruction mix and its main stall reasons.
usage: tools/ncu_segments.py prof.ncu-rep [min_pct]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
def f(r, n):
    try: return float(r[ix[n]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
CUT = ("BAR.", "SYNCS", "STTM", "LDTM", "UTMALDG", "EXIT", "WARPSYNC")
segs, cur = [], []
for r in data:
    op = r[ix["Source"]].strip()
    cur.append(r)
    if any(c in op for c in CUT):
        segs.append(cur); cur = []
if cur: segs.append(cur)
print(f"total samples {int(tot)}  segments {len(segs)}")
for s in segs:
    n = sum(f(r, "# Samples") for r in s)
    if 100 * n / tot < min_pct: continue
    mix = collections.Counter()
    for r in s:
        op = r[ix["Source"]].strip().split()
        op = [o for o in op if not o.startswith("@")][0].split(".")[0]
        mix[op] += 1
    st = sorted(((sum(f(r, k) for r in s), k[6:]) for k in stalls), reverse=True)[:4]
    ex = f(s[0], "Instructions Executed")
    print(f"{s[0][ix['Address']][-5:]}..{s[-1][ix['Address']][-5:]} {100*n/tot:5.1f}%  n={len(s):4d} exec={int(ex):9d}  end={s[-1][ix['Source']].strip()[:40]:40s} "
          + " ".join(f"{k}:{100*v/tot:.1f}" for v, k in st) + "  | " + " ".join(f"{o}{c}" for o, c in mix.most_common(6)))
