for pb in 10 50 100; do
  python bench.py --patch-batch $pb --no-layer --no-cpu-baseline --steps 20 2>/dev/null > gpurun_out/pb_$pb.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/pb_$pb.json").read().strip().splitlines()[-1])
print("PB", $pb, round(d["ms_per_step"],3), round(d["value"]/1e6,1), "e2e", round(d["e2e"]["ms_per_step"],3))
PY
done
