mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_clustering.py -k "mcmc" -x -q 2>&1 | tail -2
for n in 250 1872 2072; do
  echo "== speculative $n"; JTK_MCMC_KERNEL=speculative timeout -s KILL 120 python tools/mcmc_bench.py --chains $n --host 1 2>&1 | tail -2 | head -1
done
JTK_MCMC_DEBUG=1 JTK_CLUSTER_THREADS=4 timeout -s KILL 600 python tools/phase_scale.py --chunks 250 2>&1 | grep "jtk\]" | tail -1
JTK_MCMC_DEBUG=1 timeout -s KILL 600 python tools/phase_scale.py --chunks 2000 2>&1 | tail -2
