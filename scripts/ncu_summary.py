"""Summarise ncu reports (.ncu-rep paths) into a small text table:
    python scripts/ncu_summary.py gpurun_out/prof_sa1.ncu-rep [more.ncu-rep ...] > profiles/r01/ncu_sa_layers.txt
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [("gpu__time_duration.sum", "us", 1e-3 if False else None),
        ("dram__bytes_read.sum", "MB", None), ("dram__bytes_write.sum", "MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%", None),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "%", None),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "%", None),
        ("launch__registers_per_thread", "", None), ("launch__shared_mem_per_block_dynamic", "KB", None),
        ("lts__t_sector_hit_rate.pct", "%", None)]


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v)
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


traffic = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# %s" % os.path.basename(rep))
    print("%-28s %9s %9s %9s %7s %7s %7s %5s %8s %6s" % ("kernel", "time us", "rd MB", "wr MB", "dram%", "tensor%", "warps%", "regs", "smem KB", "L2hit%"))
    for r in data:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("unnamed>::", "")
        t = to_us(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        g = lambda k: float(r[ix[k]]) if k in ix and r[ix[k]] not in ("", "n/a") else float("nan")
        sm = to_bytes(r[ix["launch__shared_mem_per_block_dynamic"]], units[ix["launch__shared_mem_per_block_dynamic"]].split("/")[0]) / 1e3
        print("%-28s %9.1f %9.1f %9.1f %7.1f %7.2f %7.1f %5d %8.1f %6.1f" % (
            name[:28], t, rd / 1e6, wr / 1e6, g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            g("sm__warps_active.avg.pct_of_peak_sustained_active"), int(g("launch__registers_per_thread")), sm,
            g("lts__t_sector_hit_rate.pct")))
        key = re.sub(r"<.*", "", name)
        traffic.setdefault(key, []).append(rd + wr)
# profiles/roofline_traffic.json is written by scripts/ncu_layers_csv.py (all launches of one step)
