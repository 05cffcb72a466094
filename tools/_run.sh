mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/r2y_tests.log 2>&1; tail -3 gpurun_out/r2y_tests.log
timeout -s KILL 900 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -c 300 gpurun_out/r2y_bench.err
timeout -s KILL 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2y_bench_ref.json 2>> gpurun_out/r2y_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','chunks_per_s','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
print(d['extra']['chunks_phased']['phases_s'])
r=json.loads(open('gpurun_out/r2y_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline'])
PY
