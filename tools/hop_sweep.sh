#!/bin/bash
# multi-GPU hop comparison: NCCL broadcast vs NVSwitch multicast (owrx_iq_multicast_store) vs no hop (diagnostic); usage: hop_sweep.sh N [modes]
N=${1:-2}
MODES=${2:-"nccl multicast none"}
for mode in $MODES; do
OWRX_HOP=$mode OWRX_HOP_CTAS=${OWRX_HOP_CTAS:-16} timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$((RANDOM%10)) bench.py --gpus $N --steps 100 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$mode',{k:d[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step','stages_ms')}, d['e2e']['value'], d['config']['parallelism'])"
done
