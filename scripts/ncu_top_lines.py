"""Top CUDA source lines by warp-stall samples for one launch of an ncu report:
   ncu -i rep --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 > x.csv
   python scripts/ncu_top_lines.py x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cur, out = None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 7 and r[0].isdigit() and r[6].isdigit():
        out.append((int(r[6]), cur, int(r[0]), r[1][:95]))
tot = sum(o[0] for o in out)
print("total samples", tot)
for o in sorted(out, key=lambda o: -o[0])[:n]:
    print("%6d %5.1f%% %s:%d  %s" % (o[0], 100 * o[0] / max(tot, 1), o[1], o[2], o[3]))
