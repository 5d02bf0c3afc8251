timeout 120 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "fused_attention" 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 python tools/diag.py bench_ctx 2>&1 | grep "decode step"
B2L_ATTN_CLUSTER=1 timeout 200 python tools/diag.py bench_ctx 2>&1 | grep "decode step" | sed -n '2p;4p;5p'
