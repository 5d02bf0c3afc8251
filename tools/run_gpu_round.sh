B2L_BATCH_PDL=1 timeout 100 python tools/diag.py batch_debug 2>&1 | grep "pdl=" 
B2L_BATCH_PDL=1 timeout 200 python -m pytest tests -m gpu -q -k "batch" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python tools/diag.py bench_13b_b8 2>&1 | grep "13B"
B2L_BATCH_PDL=1 timeout 300 python tools/diag.py bench_13b_b8 2>&1 | grep "decode"
