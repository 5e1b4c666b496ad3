"""ncu --csv metrics log of the fused SA-layer launches of one step (scripts/gpu_ncu_all.sh) ->
a small text table on stdout + profiles/roofline_traffic.json (mean DRAM bytes per launch of the
b2r_sa_layer_fwd / b2r_sa_layer_bwd entry points; the thin first-layer kernels of
csrc/mlp_thin.cu are launches of those entry points too):
    python scripts/ncu_layers_csv.py gpurun_out/ncu_all_layers.csv > profiles/r02/ncu_sa_layers_all_blocks.txt
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    if not r[ix["ID"]].isdigit():
        continue
    d = per.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)
    d[r[ix["Metric Name"]]] = v * scale
print("# ncu --metrics over the fused SA-layer launches of one warm step (scripts/profile_sa.py all:")
print("# sa1..sa4 + vote aggregation, B=8), block by block: 3 forward then 3 backward launches each")
print("%-34s %9s %9s %9s %7s %8s" % ("kernel", "time us", "rd MB", "wr MB", "dram%", "tensor%"))
traffic = {"sa_layer_fwd": [], "sa_layer_bwd": []}
for k in sorted(per):
    d = per[k]
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "")
    name = re.sub(r"b2r::|thin::|\(anonymous namespace\)::|<?unnamed>::", "", name)
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    print("%-34s %9.1f %9.1f %9.1f %7.1f %8.2f" % (
        name[:34], d.get("gpu__time_duration.sum", 0.0), rd / 1e6, wr / 1e6,
        d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", float("nan")),
        d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", float("nan"))))
    traffic["sa_layer_fwd" if "fwd" in name else "sa_layer_bwd"].append(rd + wr)
out = {}
for op, v in traffic.items():
    if v:
        out[op] = {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v),
                   "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the %d launches of "
                             "this entry point in one step (scripts/gpu_ncu_all.sh, "
                             "profiles/r02/ncu_sa_layers_all_blocks.txt)" % len(v)}
json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
