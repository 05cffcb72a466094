mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2>> gpurun_out/r2_final_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_v11_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --phase-chunks 80 > gpurun_out/r2_launch_bench.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','chunks_per_s','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
print(json.dumps(d['extra'].get('band_sweep',{}).get('radius')), d['extra'].get('rows9'))
r=json.loads(open('gpurun_out/r2_final_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline'])
PY
