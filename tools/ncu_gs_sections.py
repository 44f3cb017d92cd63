#!/usr/bin/env python
"""Shares of k_gs_pairs by phase (line ranges found from the comment markers of ppc_gs.cuh): python tools/ncu_gs_sections.py rep [kernel]"""
import collections, os, subprocess, sys
rep = sys.argv[1]
here = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(here, "..", "particlephysicsc_b200", "csrc", "ppc_gs.cuh")).read().splitlines()
marks = [(1, "helpers / kernel head")]
for i, l in enumerate(src, 1):
    t = l.strip()
    if t.startswith("// ---- ") or t.startswith("// (a)") or t.startswith("// (b)") or t.startswith("// (c)") or t.startswith("// exclusive scan of the level") \
            or t.startswith("__global__") or t.startswith("__device__ __forceinline__ void gs_pairs_subtile"):
        marks.append((i, t[:70]))
kernel = sys.argv[2] if len(sys.argv) > 2 else "k_gs_pairs"   # the first launch of this kernel in the report
out = subprocess.run([sys.executable, os.path.join(here, "ncu_source_summary.py"), rep, "100000", "inst", kernel], capture_output=True, text=True).stdout
agg = collections.OrderedDict((n, [0, 0, 0]) for _, n in marks)
agg["ppc_math.cuh"] = [0, 0, 0]
agg["other files"] = [0, 0, 0]
for l in out.splitlines():
    if " inst " not in l or " thr " not in l:
        print(l.strip())
        continue
    p = l.split()
    inst, samp, thr, loc = float(p[0][:-1]), float(p[2][:-1]), float(p[5]), p[6]
    f, ln = loc.rsplit(":", 1)
    ln = int(ln)
    if f == "ppc_gs.cuh":
        k = marks[0][1]
        for a, n in marks:
            if ln >= a:
                k = n
    elif f == "ppc_math.cuh":
        k = "ppc_math.cuh"
    else:
        k = "other files"
    agg[k][0] += inst
    agg[k][1] += samp
    agg[k][2] += thr * inst
for k, a in agg.items():
    print("%-72s %5.1f%% inst %5.1f%% samp  lanes %4.1f" % (k, a[0], a[1], a[2] / a[0] if a[0] else 0))
