mkdir -p gpurun_out
timeout 600 python tools/diag.py trace bench_layers bench_step > gpurun_out/diag2.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/prof_step.py > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:q4_linear_tc -c 5 -f -o gpurun_out/prof_q4 python tools/prof_step.py > gpurun_out/ncu_q4.log 2>&1
timeout 600 python bench.py --steps 300 --warmup 8 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err; tail -40 gpurun_out/diag2.log
