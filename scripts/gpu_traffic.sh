#!/bin/bash
# DRAM traffic of one whole train step (all kernels) for roofline.traffic.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/traffic_train.csv python scripts/train_once.py > gpurun_out/traffic_train.log 2>&1
tail -2 gpurun_out/traffic_train.log
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/traffic_train.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
tot={}
for r in rows[1:]:
    m=r[ix['Metric Name']]; v=float(r[ix['Metric Value']].replace(',','')); u=r[ix['Metric Unit']]
    scale={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9,'us':1,'ms':1e3,'ns':1e-3}.get(u,1)
    tot[m]=tot.get(m,0)+v*scale
print(tot)
P
