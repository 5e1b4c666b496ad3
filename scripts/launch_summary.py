"""ncu launch list (scripts/gpu_launch_list.sh -> gpurun_out/launches.csv) -> per-kernel shares of
ONE eager step:  python scripts/launch_summary.py gpurun_out/launches.csv > profiles/r02/launches_step_summary.txt
The capture window holds about 3 steps; a step is delimited by SA1's FPS launch (the widest
fps_cluster_kernel instantiation)."""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
ix = {h: i for i, h in enumerate(rows[0])}
launches = []
for r in rows[1:]:
    if r[ix["ID"]].isdigit() and r[ix["Metric Name"]] == "gpu__time_duration.sum":
        v = float(r[ix["Metric Value"]].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ix["Metric Unit"]], 1.0)
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
        name = re.sub(r"\(anonymous namespace\)::|<?unnamed>::", "", name)
        launches.append((name, v))
marks = [i for i, (n, _) in enumerate(launches)
         if n.startswith("b2r::fps_cluster_kernel") or n.startswith("b2r::fps_bucket_kernel")]
# SA1's FPS is the longest FPS launch of every step
big = max(launches[i][1] for i in marks)
starts = [i for i in marks if launches[i][1] > 0.7 * big]
step = launches[starts[0]:starts[1]] if len(starts) > 1 else launches
tot = sum(v for _, v in step)
b2r = sum(v for n, v in step if n.startswith("b2r::"))
print("one eager step (bench.py --no-graph --steps 2 --warmup 3 under ncu --metrics gpu__time_duration.sum --clock-control none)")
print("per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
print("launches %d   sum %.3f ms   libb2r share %.1f%%" % (len(step), tot / 1e3, 100 * b2r / tot))
agg = {}
for n, v in step:
    c = agg.setdefault(n[:66], [0, 0.0])
    c[0] += 1
    c[1] += v
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%4d %9.1f us %5.1f%%  %s" % (c, v, 100 * v / tot, n))
