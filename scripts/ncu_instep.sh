#!/bin/bash
# ncu --set full captures of in-step GEMM launches (eager bench), one launch per epilogue variant
for spec in "256, .unsigned int.3,:cfc_fwd" "256, .unsigned int.36,:du_dgrad" "256, .unsigned int.64,:wgrad"; do
  pat="${spec%%:*}"; name="${spec##*:}"
  MMTG_GRAPH=0 timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm_bf16_tcgen05_kernel<.int.${pat}" -s 40 -c 1 -o gpurun_out/gemm_instep_${name} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_instep_${name}.log 2>&1
  echo "$name rc=$?"; grep -c "PROF== Profiling" gpurun_out/ncu_instep_${name}.log
done
