mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/r2u_tests.log 2>&1; tail -3 gpurun_out/r2u_tests.log
