#!/usr/bin/env bash
# L2-prefetch sweep of the batch-1 decode step (tools/diag.py bench_ctx = us/token at 7 context depths, graph + PDL).
# usage: bash tools/pf_sweep.sh "MODE:MB:EVICT ..."   (MB=0 = prefetch off)
mkdir -p gpurun_out
for cfg in $1; do
  IFS=: read -r mode mb ev <<< "$cfg"
  echo "=== B2L_PF_MODE=$mode B2L_PF_MB=$mb B2L_PF_EVICT=$ev"
  B2L_PF_MODE=$mode B2L_PF_MB=$mb B2L_PF_EVICT=$ev timeout 200 python tools/diag.py bench_ctx 2>&1 | grep "decode step"
done
