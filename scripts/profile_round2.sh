#!/bin/bash
# Round-2 profile set: ncu launch list of the train step + --set full captures of the kernels changed this
# round. Outputs under gpurun_out/r2_*; summarise with scripts/ncu_summary.py / launch_summary.py into profiles/.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode"
MMTG_GRAPH=0 timeout 300 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_train.csv $B > gpurun_out/r2_p_train.log 2>&1; echo "train list rc=$?"
for spec in "attn_fwd_tc_kernel:r2_attn_fwd_tc:30" "attn_bwd_tc_kernel:r2_attn_bwd_tc:30" "ln_bwd_rows_kernel:r2_ln_bwd_rows:40" \
            "ln_param_grad_kernel:r2_ln_param_grad:40" "gemm_bf16_tcgen05_kernel<.int.256, .unsigned int.3,:r2_gemm_cfc_fwd:40"; do
  pat="${spec%%:*}"; rest="${spec#*:}"; name="${rest%%:*}"; skip="${rest##*:}"
  MMTG_GRAPH=0 timeout 240 $NCU --set full --import-source on -k "regex:${pat}" -s $skip -c 1 -f -o gpurun_out/${name} $B > gpurun_out/p_${name}.log 2>&1
  echo "$name rc=$?"
done
MMTG_GRAPH=0 timeout 240 $NCU --set full --import-source on -k "regex:attn_bwd_tiled_tc_kernel" -s 8 -c 2 -f -o gpurun_out/r2_attn_bwd_tiled $B --batch 16 --max-sent-length 98 > gpurun_out/p_r2_attn_bwd_tiled.log 2>&1; echo "tiled rc=$?"
ls -la gpurun_out/r2_*.ncu-rep
