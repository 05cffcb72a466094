mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; tail -5 gpurun_out/r2o_tests.log
timeout -s KILL 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -c 600 gpurun_out/r2o_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','chunks_per_s','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'])
print(d['extra']['chunks_phased'])
PY
