timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
B2L_TL_POS=1500 timeout 300 python tools/diag.py timeline 2>&1 | grep -E "attn|whole" | head -3
