timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/diag.py timeline 2>&1 | grep -v "^\[" | head -16
timeout 300 python tools/diag.py bench_step 2>&1 | grep "decode step" | head -2
