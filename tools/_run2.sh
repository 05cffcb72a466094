mkdir -p gpurun_out
timeout -s KILL 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2z_bench_2gpu.json 2> gpurun_out/r2z_bench_2gpu.err
nproc
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','chunks_per_s')}, 'e2e', d['e2e']['value'], d['e2e']['serial']['value'], d['extra']['chunks_phased']['phases_s'], d['extra']['chunks_phased']['host_threads'])
PY
