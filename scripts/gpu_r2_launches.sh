# launch list (gpu__time_duration per kernel) of a few steady-state steps of the default bench workload
cd $GRAFT_REPO_ROOT
TAG=${1:-r2l}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s ${SKIP:-130} -c ${COUNT:-52} --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 8 --warmup 6 --e2e-steps 0 --no-cpu-baseline --no-spot-check --no-ofdm-leg ${BENCH_ARGS} > gpurun_out/${TAG}_ncu_launch.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r)>5 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0]; v=float(r[-1].replace(',',''))
    unit=r[-2]
    if unit=='ns': v/=1e3
    elif unit=='ms': v*=1e3
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in agg.items(): print(f"{k:50s} n={a[0]:3d} total={a[1]:9.1f} us  avg={a[1]/a[0]:8.1f} us  {100*a[1]/tot:5.1f}%")
PY
