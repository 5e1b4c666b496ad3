"""Top stall locations of one kernel launch from an ncu report (SASS page):
   ncu -i rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > x.csv
   python scripts/ncu_top_stalls.py x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("instructions", len(data), "total samples", tot)
agg = {}
for r in data:
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h and r[ix[h]].isdigit():
            agg[h] = agg.get(h, 0) + int(r[ix[h]])
print("stall mix:", ", ".join("%s %.1f%%" % (k, 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
    st = {h: int(r[ix[h]]) for h in hdr if h.startswith("stall_") and "Not Issued" not in h and r[ix[h]].isdigit()}
    s = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%6s %5.1f%%  %-74s %s" % (r[ix["# Samples"]], 100 * int(r[ix["# Samples"]]) / tot, r[ix["Source"]][:74], s))
