#!/bin/bash
# Bisecting aid: each variant in its own process with a hard timeout (a hang must not eat the box).
python -c "import torch; torch.zeros(1).cuda(); print('warm')" 2>&1 | tail -1
for d in 1 2 0; do
  MMTG_ATTN_DBG=$d timeout 40 python scripts/attn_probe2.py 1 236 > gpurun_out/attn_dbg_$d.log 2>&1
  echo "dbg=$d rc=$?"; grep -E "dbg|mmtg:|rror" gpurun_out/attn_dbg_$d.log | sort | uniq -c | head -4
done
