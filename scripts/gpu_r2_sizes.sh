# the full chain at other per-GPU stream counts (defaults as shipped)
cd $GRAFT_REPO_ROOT
for S in 256 512 2048; do
  (timeout 600 python bench.py --streams $S --no-cpu-baseline --e2e-steps 0 --no-spot-check --no-ofdm-leg --steps 40 2>/dev/null | tail -1) > gpurun_out/r2_size_$S.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_size_$S.json"))
print("$S streams: ms/step", round(d["ms_per_step"],4), "GS/s", round(d["value"]/1e3,2), {k: round(v/d["steps"],4) for k,v in d["kernel_ms"].items()})
PY
done
