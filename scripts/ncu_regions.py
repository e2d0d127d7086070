#!/usr/bin/env python
"""Instruction / stall-sample share per code region of k_stream (ncu --page source --csv --print-source cuda,sass).
Regions are line ranges of asrd_kernels.cuh given as name:lo-hi arguments; SASS rows attributed to inlined
headers (shuffles, atomics) inherit the region of the preceding asrd_kernels.cuh row in address order."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
regions = []
for a in sys.argv[2:]:
    name, rng = a.split(':'); lo, hi = rng.split('-'); regions.append((name, int(lo), int(hi)))
def f(x):
    try: return float(x)
    except ValueError: return 0.0
hdr = None; cur_file = None; key = None; sass = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "":
        try: key = (cur_file, int(r[0]))
        except ValueError: pass
        continue
    try: ad = int(r[2], 16)
    except ValueError: continue
    sass.append((ad, key, f(r[hdr.index("Instructions Executed")]), f(r[hdr.index("# Samples")])))
sass.sort()
def region(line):
    for name, lo, hi in regions:
        if lo <= line <= hi: return name
    return 'other@%d' % (line // 50 * 50)
agg = collections.defaultdict(lambda: [0.0, 0.0]); last = '?'; tot = [0.0, 0.0]
for ad, k, n, smp in sass:
    if k[0] == 'asrd_kernels.cuh': last = region(k[1])
    agg[last][0] += n; agg[last][1] += smp; tot[0] += n; tot[1] += smp
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:18s} inst {100*v[0]/tot[0]:5.1f}%  samples {100*v[1]/tot[1]:5.1f}%")
