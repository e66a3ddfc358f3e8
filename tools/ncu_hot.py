"""Summarise an `ncu --page source --csv` dump: top instructions by stall samples, stall-reason totals, and samples
grouped by opcode.  Usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
S = col["Warp Stall Sampling (All Samples)"]
stalls = [h for h in hdr if h.startswith("stall_")]
tot = sum(int(r[S] or 0) for r in body)
print(f"instructions {len(body)}, samples {tot}")
agg = collections.Counter()
for r in body:
    for s in stalls:
        v = r[col[s]]
        if v: agg[s] += int(float(v))
print("stall totals:", ", ".join(f"{k[6:]} {v} ({100*v/max(tot,1):.0f}%)" for k, v in agg.most_common(10)))
byop = collections.Counter()
for r in body:
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"): op = r[col["Source"]].split()[1]
    byop[op.split(".")[0]] += int(r[S] or 0)
print("by opcode:", ", ".join(f"{k} {100*v/max(tot,1):.1f}%" for k, v in byop.most_common(14)))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:N]
for i in sorted(idx):
    r = body[i]
    top = sorted(((int(float(r[col[s]] or 0)), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {100*int(r[S] or 0)/max(tot,1):5.1f}%  {r[col['Source']].strip()[:70]:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
