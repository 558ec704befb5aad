#!/bin/bash
# BASELINE.json configs[4]: MMTG train step with extended lyrics length and negative-sample ratio /
# curriculum-stage sweep on N GPUs (default 8). One JSON line per run into gpurun_out/config4_n$N.jsonl.
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/config4_sweep.sh 8'
N=${1:-8}
OUT=gpurun_out/config4_n$N.jsonl
mkdir -p gpurun_out
: > $OUT
run() {
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-decode "$@" 2>>gpurun_out/config4_n$N.err | tail -1 >> $OUT
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-decode "$@" 2>>gpurun_out/config4_n$N.err | tail -1 >> $OUT
  fi
}
for msl in 20 40 60 98; do run --max-sent-length $msl; done          # L = 236 / 436 / 636 / 1016, stage 3
for st in 1 2; do run --stage $st; done                              # rating filters: ragged per-rank batches
for nf in 0 0.5 0.75; do run --neg-frac $nf; done                    # negative fraction under stage 3
python - <<PY
import json
rows=[json.loads(l) for l in open("$OUT") if l.startswith("{")]
print("| L | stage | neg frac | global batch | samples/s | ms/step |")
print("|---|---|---|---|---|---|")
for r in rows:
    c=r["config"]
    print(f"| {c['seq_len']} | {c['stage']} | {c['neg_frac']} | {c['global_batch']} | {r['value']:.0f} | {r['ms_per_step']:.2f} |")
PY
