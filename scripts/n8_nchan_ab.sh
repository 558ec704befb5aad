#!/bin/bash
# 8-GPU A/B: number of NCCL channels (= resident NCCL CTAs competing with the GEMMs for SMs)
for c in "$@"; do
  NCCL_MAX_NCHANNELS=$c timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 15 --warmup 5 --no-decode --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('nchannels=$c', d['n_gpus'], round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline']['avg_launch_us'], {k: round(v['ms_per_step'], 2) for k, v in d['breakdown'].items()})
" | tee -a gpurun_out/r2_n8_nchan_ab.log
done
