#!/bin/bash
# Turn the ncu artefacts of one tag in gpurun_out/ into committed text summaries under profiles/.
# usage: scripts/summarize_profiles.sh <tag> [kernel ...]
TAG=$1; shift
OUT=profiles/${TAG}_summary.txt
mkdir -p profiles
{
echo "# ncu summary, tag ${TAG} ($(date -u +%Y-%m-%dT%H:%MZ)); config 2 batch (256 streams), see scripts/profile_ncu.sh"
echo
echo "## launch list (gpu__time_duration.sum per launch, ns; cold-cache serialised replay: compare SHARES)"
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/${TAG}_launches.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki].split('(')[0]].append(float(r[vi].replace(',','')))
tot = sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:28s} launches {len(v):4d}  avg_us {sum(v)/len(v)/1e3:9.1f}  share {sum(v)/tot:.3f}")
PY
for K in "$@"; do
  F=gpurun_out/${TAG}_$K.ncu-rep
  [ -f $F ] || continue
  echo
  echo "## $K  (ncu --set full, one launch mid-utterance: 256 streams x 32 frames, sub-batch overlap off)"
  ncu -i $F --page details 2>/dev/null | grep -E "Duration|Elapsed Cycles|Executed Instructions |Registers Per|Theoretical Occ|Achieved Occ|Issue Slots Busy|L2 Hit|L1/TEX Hit|DRAM Throughput|Warp Cycles Per Issued|Grid Size|Block Size|Memory Throughput|Eligible Warps|Active Warps Per"
  ncu -i $F --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for w in ['dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread']:
    if w in hdr:
        i=hdr.index(w); print('   ', w, rows[1][i], [r[i] for r in rows[2:]])
"
  ncu -i $F --page source --csv --print-source cuda,sass 2>/dev/null > /tmp/ncu_src_$$.csv
  echo "   top stall lines (share of warp-stall samples, dominant reasons):"
  python scripts/ncu_inst_lines.py /tmp/ncu_src_$$.csv 16 "# Samples" | cut -c1-190 | sed 's/^/    /'
  rm -f /tmp/ncu_src_$$.csv
done
} > $OUT
# DRAM traffic of the dominant kernel, tied to the sources it was captured from (bench.py refuses a stale one)
if [ -f gpurun_out/${TAG}_k_stream.ncu-rep ]; then
ncu -i gpurun_out/${TAG}_k_stream.ncu-rep --page raw --csv 2>/dev/null | TAG=$TAG python -c "
import csv, sys, json, os, subprocess
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
def val(name):
    i = hdr.index(name); unit = rows[1][i]; v = float(rows[2][i].replace(',', ''))
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[unit]
tag = os.environ['TAG']
rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
sha_file = 'gpurun_out/%s_src_sha.txt' % tag
out = {'kernel': 'k_stream<true>', 'ncu_tag': tag,
       'launch': '256 streams x 32 frames (5th launch of a config-2 step, ASRD_SUBBATCH=0)',
       'source': 'profiles/%s_summary.txt (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum)' % tag,
       'kernel_src_sha': open(sha_file).read().strip() if os.path.exists(sha_file) else None,
       'git_commit_when_summarised': subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], stdout=subprocess.PIPE).stdout.decode().strip(),
       'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes_per_launch': rd + wr}
json.dump(out, open('profiles/stream_traffic.json', 'w'), indent=1)
print('profiles/stream_traffic.json:', out)
"
fi
cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches.csv
echo wrote $OUT
