mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_ggml_graph.py tests/test_gpu_ring.py -x -q 2>&1 | tail -15) > gpurun_out/t9.log; tail -15 gpurun_out/t9.log
(timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --pp 0 2>&1 | tail -1) > gpurun_out/bench9.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench9.log').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"]); print("e2e",json.dumps(d["e2e"])[:1200]); print("e2e_engine",d.get("e2e_engine",{}).get("value")); print(d.get("e2e_boundary_error"))
PY
