mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/v11_alltests.log
for v in rows fused; do
JTK_MODTABLE=$v timeout 900 ncu --set full --import-source on --clock-control none -f -o gpurun_out/v11_${v}_full python tools/prof.py --rows 14 --reps 1 > gpurun_out/v11_${v}_ncu.log 2>&1
ncu -i gpurun_out/v11_${v}_full.ncu-rep --page raw --csv > gpurun_out/v11_${v}_raw.csv 2>/dev/null
done
timeout 1500 python bench.py > gpurun_out/v11_bench.json 2> gpurun_out/v11_bench.err
cat gpurun_out/v11_alltests.log; head -c 1500 gpurun_out/v11_bench.json
