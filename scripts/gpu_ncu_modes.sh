#!/bin/bash
# ncu: kernel time and DRAM bytes of the hop kernel in plain / halo-loopback / peer-loopback mode (1 GPU)
mkdir -p gpurun_out
B="bench.py --steps 6 --warmup 3 --skip-cpu --skip-cg --skip-e2e --lattice 12x48x48x48"
for m in plain loopback loopback2; do
  flag=""; [ $m != plain ] && flag="--$m"
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:hop_kernel\|ew_kernel -s 20 -c 12 --csv --log-file gpurun_out/ncu_mode_$m.csv python $B $flag > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_mode_$m.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=(r[4][:60], r[-3] if False else r[12]); 
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section Name, Metric Name, Metric Unit, Metric Value
from collections import defaultdict
d=defaultdict(lambda: defaultdict(list))
for r in rows:
    d[(r[4][:50], r[8])][r[12]].append(float(r[14].replace(',','')))
for k,v in d.items():
    print('$m', k, {m:(round(sum(x)/len(x)/ (1e6 if 'bytes' in m else 1e3),2), len(x)) for m,x in v.items()})
PY
done
