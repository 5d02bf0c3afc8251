#!/usr/bin/env bash
# One full measurement round on a B200 box (what `gpurun -- 'bash tools/run_gpu_round.sh'` ran for round 1):
# parity tests, smoke, both bench arms, the ncu launch list and the --set full captures summarised under profiles/
# by `python tools/summarize_profiles.py r01` back in the build container.  Every step has its own timeout.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cut -c1-200 gpurun_out/bench_final_ref.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/prof_step.py > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:q4_gemv_kernel -c 5 -f -o gpurun_out/prof_gemv python tools/prof_step.py > gpurun_out/ncu_gemv.log 2>&1; tail -1 gpurun_out/ncu_gemv.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_decode -c 2 -f -o gpurun_out/prof_attn python tools/prof_step.py > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log
timeout 400 python tools/diag.py bench_13b_b8 2>&1 | grep "13B" | tee gpurun_out/bench_13b_b8.log
timeout 600 python tools/diag.py bench_sizes 2>&1 | grep "gptq.int4" | tee gpurun_out/bench_sizes.log
timeout 300 python tools/diag.py bench_step_int8 2>&1 | grep "decode step"
