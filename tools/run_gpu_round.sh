timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py 2>&1 | tail -1
