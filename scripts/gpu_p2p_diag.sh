#!/bin/bash
# where does the peer-mode overhead come from?  (timing diagnostics; --p2p-diag bits give WRONG results across ranks and
# need TMB_P2P_DIAG=1: 1 boundary slices read the local field, 2 no end-of-hop handshake - see tmb_set_p2p_diag)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --steps 400 --warmup 20 --skip-cpu --skip-cg --skip-e2e"
pr() { python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['roofline']['avg_launch_us'],1), 'us/hop', round(d['value']), 'GFLOP/s peer', d.get('peer_mode'))"; }
python $B --lattice 12x48x48x48 2>/dev/null | pr "N=1 plain            "
python $B --lattice 12x48x48x48 --loopback 2>/dev/null | pr "N=1 loopback halo    "
python $B --lattice 12x48x48x48 --loopback2 2>/dev/null | pr "N=1 loopback peer    "
for dg in 0 1 2 3; do
  TMB_P2P_DIAG=1 $TR --nproc-per-node $N --master-port 2952$dg $B --gpus $N --p2p-diag $dg 2>/dev/null | pr "N=$N peer p2p-diag=$dg "
done
TMB_P2P=0 $TR --nproc-per-node $N --master-port 29529 $B --gpus $N 2>/dev/null | pr "N=$N NCCL halos       "
for cc in 16 32 128; do
  TMB_P2P_COPY_CTAS=$cc $TR --nproc-per-node $N --master-port 2953$((cc/16)) $B --gpus $N 2>/dev/null | pr "N=$N peer copy_ctas=$cc "
done
