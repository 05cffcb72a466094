mkdir -p gpurun_out
for lib in jtk_b200/libjtkgpu.so jtk_b200/libjtkgpu_b2.so; do
  JTK_LIB_PATH=$PWD/$lib timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ab_$(basename $lib .so).csv python tools/prof.py --reps 3 > /dev/null 2>&1
  python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2ab_$(basename $lib .so).csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value')
t=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)>mv: t[r[kn][:40]].append(float(r[mv].replace(',',''))/1e6)
print('$lib', {k:round(sorted(v)[len(v)//2],3) for k,v in t.items()})
PY
done
