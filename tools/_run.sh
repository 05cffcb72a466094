mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/r2aa_tests.log 2>&1; tail -3 gpurun_out/r2aa_tests.log
( for k in speculative subwarp; do for n in 148 250 592 1184 1872 2072; do
  echo "== $k $n"; JTK_MCMC_KERNEL=$k timeout -s KILL 120 python tools/mcmc_bench.py --chains $n --host 1 2>&1 | tail -2
done; done ) > gpurun_out/r2aa_mcmc_bench.txt 2>&1; grep -c identical gpurun_out/r2aa_mcmc_bench.txt
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:mcmc_speculative -c 1 -o gpurun_out/r2aa_spec_2072 -f python tools/mcmc_prof.py --chains 2072 --restarts 1 > gpurun_out/r2aa_ncu.log 2>&1
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:mcmc_speculative -c 1 -o gpurun_out/r2aa_spec_250 -f python tools/mcmc_prof.py --chains 250 --restarts 1 >> gpurun_out/r2aa_ncu.log 2>&1
timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; tail -c 300 gpurun_out/r2aa_bench.err
timeout -s KILL 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2aa_bench_ref.json 2>> gpurun_out/r2aa_bench.err
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2aa_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --phase-chunks 80 > gpurun_out/r2aa_launch_bench.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aa_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','chunks_per_s','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
print(d['extra']['chunks_phased']['phases_s'])
r=json.loads(open('gpurun_out/r2aa_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['ms_per_step'], r['cpu_baseline'])
PY
