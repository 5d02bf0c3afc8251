for cfg in "2 6" "3 4" "3 3"; do set -- $cfg; echo "== CTAS_PER_SM=$1 STAGES=$2"; B2L_GEMV_CTAS_PER_SM=$1 B2L_GEMV_STAGES=$2 timeout 300 python tools/diag.py bench_step 2>&1 | grep "decode step" | head -1; done
B2L_GEMV_CTAS_PER_SM=3 B2L_GEMV_STAGES=4 timeout 300 python tools/diag.py bench_gemv 2>&1 | grep "pdl=1:"
