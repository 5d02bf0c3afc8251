#!/usr/bin/env bash
# One full measurement round on a 1-GPU B200 box: `gpurun -- 'bash tools/run_gpu_round.sh'`.
# Parity tests, smoke, the bench, the ncu launch list and the --set full captures that
# `python tools/summarize_profiles.py r02` (back in the build container) turns into profiles/r02_*.txt.
# Every step has its own timeout; nothing printed under ncu is a bench number.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-1800 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/prof_step.py > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:q4_gemv_kernel -c 5 -f -o gpurun_out/prof_gemv python tools/prof_step.py > gpurun_out/ncu_gemv.log 2>&1; tail -1 gpurun_out/ncu_gemv.log
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_decode -c 2 -f -o gpurun_out/prof_attn python tools/prof_step.py > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log
B2L_GEMM_M=4096 timeout 400 ncu --set full --clock-control none --import-source on -k regex:q4_gemm_kernel -s 2 -c 1 -f -o gpurun_out/prof_gemm python tools/diag.py bench_gemm > gpurun_out/ncu_gemm.log 2>&1; grep -E "^gemm|==PROF==.*[Dd]isconnected" gpurun_out/ncu_gemm.log | tail -3
