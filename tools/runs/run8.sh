mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/t8.log; tail -8 gpurun_out/t8.log
(timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --pp 0 2>&1 | tail -1) > gpurun_out/bench8.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench8.log').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"]); print("e2e",json.dumps(d["e2e"])[:900]); print("e2e_engine",d.get("e2e_engine",{}).get("value")); print(d.get("e2e_boundary_error"))
PY
