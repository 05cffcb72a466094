mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "small or wide_band or ragged or packed or bit_identical or very_long" > gpurun_out/v11_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/v11_memcheck.log
tail -15 gpurun_out/v11_memcheck.log
