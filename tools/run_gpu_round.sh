mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json | cut -c1-330
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/prof_step.py > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
timeout 400 python tools/diag.py bench_13b_b8 2>&1 | grep "13B" | tee gpurun_out/bench_13b_b8.log
timeout 600 python tools/diag.py bench_sizes 2>&1 | grep "gptq.int4" | tee gpurun_out/bench_sizes.log
