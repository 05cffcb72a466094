mkdir -p gpurun_out
N=$1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/r2_final_bench_${N}gpu.json 2> gpurun_out/r2_final_bench_${N}gpu.err
tail -c 400 gpurun_out/r2_final_bench_${N}gpu.err; nproc; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_final_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','chunks_per_s')}, 'e2e', d['e2e']['value'], d['extra']['chunks_phased']['phases_s'], d['extra']['chunks_phased'].get('host_threads'))
PY
