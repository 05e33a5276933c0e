mkdir -p gpurun_out
( timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"wt_|scatter|gather|land_kernel|halo" -s 12 -c 30 --csv --log-file gpurun_out/r02_c5_launches.csv python bench.py --config C5 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_c5_ncu.log 2>&1 )
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_c5_launches.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
d=collections.defaultdict(dict)
for r in rows[1:]:
    d[(r[iid], r[ik][:60])][r[im]]=float(r[iv].replace(',',''))
agg=collections.defaultdict(lambda:[0,0,0,0])
for (i,k),m in d.items():
    a=agg[k]; a[0]+=1; a[1]+=m.get('gpu__time_duration.sum',0); a[2]+=m.get('dram__bytes_read.sum',0); a[3]+=m.get('dram__bytes_write.sum',0)
for k,a in agg.items(): print("%-62s n=%2d  avg %.3f ms  dram r %.2f GB w %.2f GB per launch" % (k,a[0],a[1]/a[0]/1e6,a[2]/a[0]/1e9,a[3]/a[0]/1e9))
PY
