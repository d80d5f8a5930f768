mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -vE "^E\s+\[|^E\s+[0-9\.\-e, ]+\]|^E\s+\+|array\(" | tail -12 | cut -c1-600
