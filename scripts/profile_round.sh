#!/bin/bash
# Round profile set: ncu launch lists (train step, one generation) + --set full captures of the
# dominant kernels. Outputs under gpurun_out/; summarise with scripts/ncu_summary.py / launch_summary.py.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
MMTG_GRAPH=0 timeout 300 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_train.csv $B > gpurun_out/p_train.log 2>&1; echo "train list rc=$?"
timeout 300 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_decode.csv python scripts/decode_bench.py 64 > gpurun_out/p_decode.log 2>&1; echo "decode list rc=$?"
for spec in "gemm_bf16_tcgen05_kernel<.int.256, .unsigned int.3,:gemm_cfc_fwd:40" "gemm_bf16_tcgen05_kernel<.int.256, .unsigned int.36,:gemm_du_dgrad:40" \
            "gemm_bf16_tcgen05_kernel<.int.256, .unsigned int.64,:gemm_wgrad:40" "gemm_bf16_tcgen05_kernel<.int.256, .unsigned int.520,:gemm_cproj_res_drop:40" \
            "attn_fwd_tc_kernel:attn_fwd_tc:30" "attn_bwd_tc_kernel:attn_bwd_tc:30"; do
  pat="${spec%%:*}"; rest="${spec#*:}"; name="${rest%%:*}"; skip="${rest##*:}"
  MMTG_GRAPH=0 timeout 240 $NCU --set full --import-source on -k "regex:${pat}" -s $skip -c 1 -f -o gpurun_out/${name} $B > gpurun_out/p_${name}.log 2>&1
  echo "$name rc=$?"
done
timeout 300 $NCU --set full --import-source on -k regex:decode_mega_kernel -s 150 -c 1 -f -o gpurun_out/decode_mega python scripts/decode_bench.py 64 > gpurun_out/p_decode_mega.log 2>&1; echo "decode_mega rc=$?"
ls -la gpurun_out/*.ncu-rep
