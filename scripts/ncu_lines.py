#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line:
share of warp-stall samples per line with the dominant stall reasons."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
out = collections.OrderedDict()
cur_file = None; hdr = None; first_kernel_done = False
agg = collections.defaultdict(lambda: collections.defaultdict(float)); src_text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    try: line = int(r[0])
    except ValueError: continue
    # source-level rows have an empty Address column
    ai = hdr.index("Address")
    if r[ai] not in ("", "-"): continue
    key = (cur_file, line)
    src_text[key] = r[1]
    for name in ("# Samples", "stall_long_sb", "stall_barrier", "stall_short_sb", "stall_lg", "stall_wait",
                 "stall_mio", "stall_branch_resolving", "stall_not_selected", "stall_math", "stall_membar",
                 "Instructions Executed", "L2 Theoretical Sectors Global"):
        if name in hdr:
            try: agg[key][name] += float(r[hdr.index(name)])
            except ValueError: pass
tot = sum(v["# Samples"] for v in agg.values()) or 1
print(f"total samples {tot:.0f}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:topn]:
    reasons = sorted(((n, x) for n, x in v.items() if n.startswith("stall_")), key=lambda t: -t[1])[:2]
    rs = " ".join(f"{n[6:]}={100*x/max(v['# Samples'],1):.0f}%" for n, x in reasons)
    print(f"{100*v['# Samples']/tot:5.1f}%  {key[0]}:{key[1]:<4} {rs:28s} | {src_text[key].strip()[:90]}")
