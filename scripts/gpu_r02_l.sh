#!/bin/bash
# round 2, call l (1 GPU): the CTA tile traversal - parity tests, A/B timings (scripts/tile_ab.py), one ncu pass with the L1/L2 counters
mkdir -p gpurun_out
echo "== pytest (tile + hopping)"; SECONDS=0
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile or hopping_and_epilogues or variants or full_size or nd_two_flavour" > gpurun_out/r02l_pytest.log 2>&1; echo "rc=$? wall=${SECONDS}s"; tail -4 gpurun_out/r02l_pytest.log
echo "== A/B"; SECONDS=0
timeout 900 python scripts/tile_ab.py 48x24x24x24 12x48x48x48 64x32x32x32 > gpurun_out/r02l_tile_ab.jsonl 2> gpurun_out/r02l_tile_ab.err; echo "rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/r02l_tile_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r02l_tile_ab.jsonl'):
    d = json.loads(l); print(d['lattice_TxLXxLYxLZ'])
    for k in ('single', 'peer_loopback'):
        for tv in ('linear', 'tile'):
            r = d[k][tv]
            print(f"  {k:14s} {tv:6s}", {a: (r[a]['burst_us'], r[a]['sustained_us'], r[a]['frac_burst'], r[a]['frac_sustained']) for a in r if a.startswith('hop')}, r['cg'], r.get('bit_identical_to_linear'))
    if 'two_flavour' in d:
        for tv in ('linear', 'tile'):
            print('  two_flavour', tv, d['two_flavour'][tv])
        print('  two_flavour bit identical', d['two_flavour']['bit_identical_to_linear'])
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
CMD3="python scripts/profile_tile.py 48x24x24x24"
echo "== ncu"; $CMD3 > gpurun_out/r02l_profile_plain.log 2>&1 && ncu --metrics $M --clock-control none -k regex:hop2_kernel\|hop_kernel -c 120 --csv --log-file gpurun_out/r02l_tile_ncu_metrics.csv $CMD3 > gpurun_out/r02l_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02l_profile_plain.log
