#!/usr/bin/env bash
# usage: bash tools/env_sweep.sh "A=1,B=2 A=0" [diag section]   -- runs tools/diag.py <section> under each comma-separated env set
sec=${2:-bench_ctx}
for cfg in $1; do
  echo "=== $cfg"
  env $(echo "$cfg" | tr ',' ' ') timeout 300 python tools/diag.py $sec 2>&1 | grep -E "decode step|us/token|tok/s"
done
