timeout 600 python -m pytest tests/test_gpu_quant.py -m gpu -q -k gemv 2>&1 | tail -2
timeout 300 python tools/diag.py bench_step 2>&1 | grep "decode step" | head -1
timeout 300 python tools/diag.py cta_times 2>&1 | grep "^---"
