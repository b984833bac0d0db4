for nch in 4 8 16; do
NCCL_MAX_NCHANNELS=$nch OWRX_HOP=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$((nch%10)) bench.py --gpus 2 --steps 100 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('nch',$nch,{k:d[k] for k in ('value','ms_per_step','stages_ms')})"
done
