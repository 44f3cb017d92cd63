#!/usr/bin/env python
"""Section-level shares of ppc_gs.cuh from an ncu report (source page): python tools/ncu_sections.py rep"""
import collections, subprocess, sys
rep = sys.argv[1]
sec = [(0, 215, 'helpers (bins, column sums, run table)'), (216, 271, 'A setup + TMA issue'), (272, 302, 'A cell table / lut / wait'), (303, 320, 'A histogram'),
       (321, 372, 'A bin scan + scatter'), (373, 432, 'A candidate loop'), (433, 505, 'A ordered walk + pushes'), (506, 550, 'A cell walk (buckets)'),
       (551, 619, 'A bucket scan + chunk'), (620, 645, 'A records'), (646, 675, 'A kernel head'), (676, 721, 'B head'), (722, 732, 'B load velocities'),
       (733, 764, 'B fire'), (765, 800, 'B write back')]
out = subprocess.run([sys.executable, __file__.replace("ncu_sections", "ncu_source_summary"), rep, "100000", "inst"], capture_output=True, text=True).stdout
agg = collections.OrderedDict((n, [0, 0, 0]) for _, _, n in sec)
agg['ppc_math (geometry, push, impulse)'] = [0, 0, 0]
agg['other'] = [0, 0, 0]
for l in out.splitlines():
    if ' inst ' not in l or ' thr ' not in l:
        print(l.strip())
        continue
    p = l.split()
    inst, samp, thr, loc = float(p[0][:-1]), float(p[2][:-1]), float(p[5]), p[6]
    f, ln = loc.split(':')
    ln = int(ln)
    k = 'other'
    if f == 'ppc_gs.cuh':
        for a, b, n in sec:
            if a <= ln <= b:
                k = n
                break
    elif f == 'ppc_math.cuh':
        k = 'ppc_math (geometry, push, impulse)'
    agg[k][0] += inst
    agg[k][1] += samp
    agg[k][2] += thr * inst
for k, a in agg.items():
    print('%-40s %5.1f%% inst %5.1f%% samp  lanes %4.1f' % (k, a[0], a[1], a[2] / a[0] if a[0] else 0))
