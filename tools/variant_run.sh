python -m pytest tests/test_vardct_gpu.py -m gpu -q -x -k "host_entry or int16 or test_epf" 2>&1 | tail -2
python tools/variant_time.py
