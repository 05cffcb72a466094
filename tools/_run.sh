# the round's evidence run on one B200 (gpurun --timeout 3000 -- 'bash tools/_run.sh'): GPU tests, MCMC kernel table, ncu captures of
# the speculative kernel, bench + reference arm with the driver's arguments, ncu launch list
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
( for k in speculative subwarp; do for n in 148 250 592 1184 1872 2072; do
  echo "== $k $n"; JTK_MCMC_KERNEL=$k timeout -s KILL 120 python tools/mcmc_bench.py --chains $n --host 1 2>&1 | tail -2
done; done ) > gpurun_out/final_mcmc_bench.txt 2>&1
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:mcmc_speculative -c 1 -o gpurun_out/final_spec_2072 -f python tools/mcmc_prof.py --chains 2072 --restarts 1 > gpurun_out/final_ncu.log 2>&1
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:mcmc_speculative -c 1 -o gpurun_out/final_spec_250 -f python tools/mcmc_prof.py --chains 250 --restarts 1 >> gpurun_out/final_ncu.log 2>&1
timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout -s KILL 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --phase-chunks 80 > gpurun_out/final_launch_bench.log 2>&1
