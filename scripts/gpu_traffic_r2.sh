#!/bin/bash
# DRAM traffic of one whole train step and one whole decode (all kernels) for roofline.traffic -> gpurun_out/r2_traffic.json
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/r2_traffic_train.csv python scripts/train_once.py --train-blocks 0 > gpurun_out/r2_traffic_train.log 2>&1
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/r2_traffic_decode.csv python scripts/decode_once.py > gpurun_out/r2_traffic_decode.log 2>&1
python - <<'P'
import csv, json
out = {}
for kind in ("train", "decode"):
    rows = [r for r in csv.reader(open('gpurun_out/r2_traffic_%s.csv' % kind)) if len(r) > 10]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    tot = {}; kernels = set(); nvf = {}
    for r in rows[1:]:
        m = r[ix['Metric Name']]; v = float(r[ix['Metric Value']].replace(',', '')); u = r[ix['Metric Unit']]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1, 'ms': 1e3, 'ns': 1e-3}.get(u, 1)
        tot[m] = tot.get(m, 0) + v * scale
        kernels.add(r[ix['ID']])
        if 'nvf' in r[ix['Kernel Name']] or 'fast::' in r[ix['Kernel Name']] or 'k_' in r[ix['Kernel Name']]:
            nvf[m] = nvf.get(m, 0) + v * scale
    out[kind] = dict(launches=len(kernels), read=tot.get('dram__bytes_read.sum'), write=tot.get('dram__bytes_write.sum'),
                     dram_bytes=tot.get('dram__bytes_read.sum', 0) + tot.get('dram__bytes_write.sum', 0),
                     serial_us=tot.get('gpu__time_duration.sum'))
print(json.dumps(out))
json.dump(out, open('gpurun_out/r2_traffic.json', 'w'), indent=1)
P
