timeout 300 python tools/diag.py timeline 2>&1 | grep -E "^  L4|^  L5 c_attn" | head -8
