#!/usr/bin/env python
"""Per-source-line shares of one kernel, summed over all profiled launches of an ncu report (--import-source on, -lineinfo).

    python tools/ncu_line_agg.py gpurun_out/prof.ncu-rep [top] [sort: inst|samp]
"""
import collections
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "inst"
out = subprocess.run([sys.executable, __file__.replace("ncu_line_agg", "ncu_source_summary"), rep, "100000", "inst"], capture_output=True, text=True).stdout
agg = collections.OrderedDict()
for l in out.splitlines():
    if " inst " not in l or " thr " not in l:
        print(l.strip())
        continue
    parts = l.split()
    inst, samp, thr, loc, src = float(parts[0][:-1]), float(parts[2][:-1]), float(parts[5]), parts[6], " ".join(parts[7:])
    a = agg.setdefault(loc, [0, 0, 0, src])
    a[0] += inst
    a[1] += samp
    a[2] += thr * inst
idx = 0 if key == "inst" else 1
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:top]:
    print("%5.1f%% inst %5.1f%% samp  thr %4.1f  %s  %s" % (a[0], a[1], a[2] / a[0] if a[0] else 0, k, a[3][:110]))
