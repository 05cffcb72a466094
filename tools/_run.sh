timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "packed or bootstrap or likelihood" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python tools/calib_time.py --min-gain 2>&1 | tail -12
