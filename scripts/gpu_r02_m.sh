#!/bin/bash
# round 2, call m (1 GPU): prefetch-ahead and K6a residency experiments; warp-state ncu sections of K6a and K1
mkdir -p gpurun_out
echo "== pytest subset"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile or nd_two_flavour or nd_doublet" > gpurun_out/r02m_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02m_pytest.log
echo "== A/B"; SECONDS=0
timeout 900 python scripts/k6a_ab.py 64x32x32x32 48x24x24x24 > gpurun_out/r02m_k6a_ab.jsonl 2> gpurun_out/r02m_k6a_ab.err; echo "rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/r02m_k6a_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r02m_k6a_ab.jsonl'):
    d = json.loads(l); print(d['lattice_TxLXxLYxLZ'])
    for k in ('hop_f64', 'hop_f32', 'Qtm_pm_ndpsi', 'Qtm_pm_ndpsi_32'):
        for c, r in d[k].items():
            print(f"  {k:16s} {c:28s} burst {r['burst_us']:8.2f} us ({r['frac_burst']:.3f})  sustained {r['sustained_us']:8.2f} us ({r['frac_sustained']:.3f})")
PY
CMD3="python scripts/profile_k6a.py 48x24x24x24"
echo "== ncu"; $CMD3 > gpurun_out/r02m_profile_plain.log 2>&1 && ncu --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section InstructionStats --clock-control none -k regex:hop2_kernel\|hop_kernel -s 4 -c 5 -f -o /tmp/r02m_k6a $CMD3 > gpurun_out/r02m_ncu.log 2>&1
echo "ncu rc=$?"; ncu -i /tmp/r02m_k6a.ncu-rep --page details --csv > gpurun_out/r02m_k6a_ncu_details.csv 2>/dev/null; ncu -i /tmp/r02m_k6a.ncu-rep --page raw --csv > gpurun_out/r02m_k6a_ncu_raw.csv 2>/dev/null; ls -la gpurun_out/r02m_*
