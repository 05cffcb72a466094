timeout 600 python -m pytest tests/test_polish.py tests/test_consensus.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2
JTK_TIMING=1 timeout 900 python tools/phase_scale.py --chunks 2000 2>&1 | grep -E "polish:|phases_s" | tail -3
