"""Time-ordered view of bench.py --trace's *_timeline.txt (kernels >= min_us)."""
import re, sys
path = sys.argv[1]; min_us = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
for ln in open(path):
    if ln.startswith('#'): continue
    m = re.match(r'\s*([\d.]+)\s+([\d.]+)\s+s(\d+)\s+(.*)', ln)
    if not m: continue
    t, d, s, n = float(m.group(1)), float(m.group(2)), int(m.group(3)), m.group(4)
    n = re.sub(r'void |b2r::|\(anonymous namespace\)::|at::native::', '', n)[:60]
    if d >= min_us: print("%8.1f -> %8.1f  %7.1f  s%-2d %s" % (t, t + d, d, s, n))
