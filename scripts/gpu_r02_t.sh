#!/bin/bash
# round 2, call t (1 GPU): where the Z split's time goes - ncu launch list of the loop-back hop
mkdir -p gpurun_out
CMD="python scripts/profile_zsplit.py 24x48x48x24"
$CMD > gpurun_out/r02t_zsplit_plain.log 2>&1; echo "rc=$?"; cat gpurun_out/r02t_zsplit_plain.log | tail -3
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread --clock-control none -s 10 -c 16 --csv --log-file gpurun_out/r02t_zsplit_launches.csv $CMD > gpurun_out/r02t_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02t_zsplit_launches.csv')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn = h.index('Kernel Name'); mn = h.index('Metric Name'); mv = h.index('Metric Value'); idc = h.index('ID')
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > mv: d.setdefault((int(r[idc]), r[kn][:70]), {})[r[mn]] = float(r[mv].replace(',', ''))
for k, v in d.items():
    print(k[0], k[1], '|', round(v['gpu__time_duration.sum'] / 1e3, 1), 'us |', round((v['dram__bytes_read.sum'] + v['dram__bytes_write.sum']) / 1e6, 1), 'MB | grid', int(v['launch__grid_size']), '| regs', int(v['launch__registers_per_thread']))
PY
