#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1]))); topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sort = sys.argv[3] if len(sys.argv) > 3 else "Instructions Executed"
hdr = None; cur_file = None; key = None
agg = collections.defaultdict(lambda: collections.defaultdict(float)); text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "":
        try: key = (cur_file, int(r[0])); text[key] = r[1]
        except ValueError: pass
        continue
    for i, name in enumerate(hdr):
        if name in ("# Samples", "Instructions Executed", "Thread Instructions Executed") or name.startswith("stall_"):
            try: agg[key][name] += float(r[i])
            except (ValueError, IndexError): pass
toti = sum(v["Instructions Executed"] for v in agg.values()); tots = sum(v["# Samples"] for v in agg.values())
print("total inst", toti, "samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][sort])[:topn]:
    reasons = sorted(((n, x) for n, x in v.items() if n.startswith("stall_")), key=lambda t: -t[1])[:2]
    rs = " ".join(f"{n[6:]}={100*x/max(v['# Samples'],1):.0f}%" for n, x in reasons)
    print(f"{100*v['Instructions Executed']/toti:5.1f}% inst {100*v['# Samples']/tots:5.1f}% smp thr/inst {v['Thread Instructions Executed']/max(v['Instructions Executed'],1):5.1f} {rs:26s} {k[0][:18]}:{k[1]:<5}| {text[k].strip()[:90]}")
