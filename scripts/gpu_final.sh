# final bench lines of the round (no profiler): default (with cpu_baseline), full chain at 256 and 1024 streams, reference arm
cd $GRAFT_REPO_ROOT
TAG=${1:-rX}
(timeout 900 python bench.py 2>&1 | tail -1) > gpurun_out/${TAG}_bench_default.json 2>&1
(timeout 600 python bench.py --workload full --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full.json 2>&1
(timeout 600 python bench.py --workload full --streams 1024 --steps 24 --no-cpu-baseline --e2e-steps 8 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full_1024.json 2>&1
(timeout 600 python bench.py --impl reference --steps 1 --warmup 3 2>&1 | tail -1) > gpurun_out/${TAG}_bench_reference.json 2>&1
python - <<PY
import json
for f in ('bench_default','bench_full','bench_full_1024','bench_reference'):
    try:
        d=json.loads(open('gpurun_out/${TAG}_'+f+'.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'vit',round(d.get('viterbi_mbit_s',0) or 0,1),'e2e',round(d['e2e']['value'],1),'rt',round(d.get('realtime_streams',0)))
    except Exception as e:
        print(f,'FAILED',e, open('gpurun_out/${TAG}_'+f+'.json').read()[-400:])
PY
