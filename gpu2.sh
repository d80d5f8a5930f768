timeout 1200 python -m pytest tests/test_optim_gpu.py -q -m gpu --tb=short -s 2>&1 | grep -vE "^E\s+\[|^E\s+[0-9\.\-e, ]+\]|^E\s+\+|array\(" | tail -8 | cut -c1-600
