#!/bin/bash
# N-GPU A/B of the gradient wire format: $1 = number of GPUs, rest = MMTG_DDP_GRAD_DTYPE values
N=$1; shift
for c in "$@"; do
  MMTG_DDP_GRAD_DTYPE=$c timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 15 --warmup 5 --no-decode --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('wire=$c', d['n_gpus'], round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline']['avg_launch_us'], {k: round(v['ms_per_step'], 2) for k, v in d['breakdown'].items()})
" | tee -a gpurun_out/r2_ddp_wire_ab.log
done
