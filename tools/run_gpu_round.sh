timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python tools/diag.py bench_step 2>&1 | grep "decode step" | head -1
