#!/usr/bin/env python
"""Region breakdown of the fused kernel profile: usage ncu_regions.py src.csv dis.txt func n_samples"""
import csv, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_by_line import sass_lines
srcfile = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'contouring_uncertainty_b200', 'csrc', 'fused.cu')
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; idx = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 10]
sl = sass_lines(sys.argv[2], sys.argv[3]); n = float(sys.argv[4])
print(len(data), len(sl))
src = open(srcfile).read().split('\n')
def find(s): return [i + 1 for i, l in enumerate(src) if s in l][0]
marks = [('philox/boxmuller', find('Philox4x32-10 + Box-Muller')), ('bmp helpers/run_fill', find('bitmap helpers')), ('flood_sweep', find('One Gauss-Seidel sweep')), ('setup/zero/round', find('one sample, one warp')), ('chord/param', find('3a. chord-length')), ('tridiag PCR+coef', find('3b. not-a-knot')), ('spline eval+draw', find('3c/4. evaluate')), ('fallback', find('fallback of contour_spline')), ('closing line', find('5. closing line')), ('fill 6a', find('6. background flood')), ('fill 6b loop', find('6b. vertical propagation')), ('outputs', find('7. mask = not reached')), ('kernel rng/matvec/io', find('-- kernels'))]
marks.sort(key=lambda x: x[1])
def region(line):
    name = 'hdr'
    for nm, l in marks:
        if line >= l: name = nm
    return name
agg = {}; ti = ts = 0
for i, r in enumerate(data):
    key = sl[i][0]; ins = float(r[idx['Instructions Executed']] or 0); smp = float(r[idx['# Samples']] or 0)
    ti += ins; ts += smp
    name = 'other(inlined hdr)'
    if key and key[0] == 'fused.cu': name = region(key[1])
    a = agg.setdefault(name, [0, 0]); a[0] += ins; a[1] += smp
print('total/sample', ti / n)
for nm, (ins, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0]): print(f"{nm:28s} ins {ins/ti*100:5.1f}%  ({ins/n:7.0f}/sample)  stall-samples {smp/ts*100:5.1f}%")
