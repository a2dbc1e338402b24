#!/bin/bash
mkdir -p gpurun_out
for cfg in 8,16,8,8 16,32,16,16; do
  tag=$(echo $cfg | tr , _)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_head_launches_$tag.csv python scripts/decode_probe.py $cfg 1024 1 > gpurun_out/r2_head_launches_$tag.log 2>&1
  python - <<PY
import csv
rows=[l for l in open("gpurun_out/r2_head_launches_$tag.csv") if not l.startswith("==")]
out=[]
for r in csv.DictReader(rows):
    if r.get("Metric Name")!="gpu__time_duration.sum": continue
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
    v=v/1e3 if u=="ns" else v*1e3 if u=="ms" else v
    out.append((r["Kernel Name"][:70], v))
# last decode = everything after the last-but-one k_scan_counts
idx=[i for i,(k,_) in enumerate(out) if "k_scan_counts" in k]
lo=idx[-2]+1 if len(idx)>1 else 0
tot=0
for k,v in out[lo:]:
    if v>20: print("%9.1f us  %s"%(v,k))
    tot+=v
print("last decode: %.1f us in %d launches"%(tot,len(out)-lo))
PY
done
