#!/usr/bin/env python
"""Section-level shares of ppc_gs.cuh from an ncu report (source page): python tools/ncu_sections.py rep
Sections are line ranges of the CURRENT ppc_gs.cuh, found from `// ---- ` / `// (a)` style markers given below."""
import collections, re, subprocess, sys, os
rep = sys.argv[1]
src = open(os.path.join(os.path.dirname(__file__), "..", "particlephysicsc_b200", "csrc", "ppc_gs.cuh")).read().splitlines()
marks = []   # (line number, name): a section runs from its marker to the next one
for i, l in enumerate(src, 1):
    m = re.match(r"\s*// SECTION: (.*)", l)
    if m:
        marks.append((i, m.group(1).strip()))
marks.append((len(src) + 1, "end"))
def section(ln):
    name = "before first marker"
    for a, n in marks:
        if ln >= a:
            name = n
        else:
            break
    return name
out = subprocess.run([sys.executable, __file__.replace("ncu_sections", "ncu_source_summary"), rep, "100000", "inst"], capture_output=True, text=True).stdout
agg = collections.OrderedDict()
for l in out.splitlines():
    if ' inst ' not in l or ' thr ' not in l:
        print(l.strip())
        continue
    p = l.split()
    inst, samp, thr, loc = float(p[0][:-1]), float(p[2][:-1]), float(p[5]), p[6]
    f, ln = loc.rsplit(':', 1)
    k = section(int(ln)) if f == 'ppc_gs.cuh' else f
    a = agg.setdefault(k, [0, 0, 0])
    a[0] += inst
    a[1] += samp
    a[2] += thr * inst
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%-50s %5.1f%% inst %5.1f%% samp  lanes %4.1f' % (k, a[0], a[1], a[2] / a[0] if a[0] else 0))
