mkdir -p gpurun_out
timeout 900 python tools/diag.py gemv model bench_gemv bench_step > gpurun_out/diag4.log 2>&1
cat gpurun_out/diag4.log | head -150
