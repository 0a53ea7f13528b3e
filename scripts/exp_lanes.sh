for cfg in "32 6" "32 12" "64 6" "64 12" "128 8" "128 16"; do
set -- $cfg
FXG_GROUPS=$2 timeout 200 python bench.py --only config2 --lanes $1 --steps 10 --warmup 3 --cpu-seconds 0.3 > gpurun_out/exp.json 2>/dev/null
python - <<PY
import json
b=json.load(open("gpurun_out/exp.json"))
print("lanes $1 groups $2: ms/batch", round(b["ms_per_batch"],3), "frac", round(b["roofline"]["frac"],3), "jobs/batch", round(b["queue"]["jobs_per_batch"],1), "e2e", round(b["e2e"]["value"]), "alloc", b["queue"]["alloc_calls_in_region"])
PY
done
