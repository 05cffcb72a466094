timeout -s KILL 900 python -m pytest tests/test_polish.py tests/test_consensus.py tests/test_gpu_pipeline.py tests/test_fit.py -m gpu -q 2>&1 | tail -2
timeout -s KILL 600 python tools/phase_scale.py --chunks 2000 2>&1 | tail -1
