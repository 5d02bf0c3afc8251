timeout 300 python tools/diag.py timeline 2>&1 | grep -E "CTA0|whole" | head -3
