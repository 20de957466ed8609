mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6) > gpurun_out/t30.log; tail -6 gpurun_out/t30.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) | tee gpurun_out/smoke30.log
(timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench30.err | tail -1) > gpurun_out/bench30.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench30.json').read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"],"whole",d["roofline"]["whole_step"]["frac"])
print("prefill",d.get("prefill",{}).get("ms"), d.get("prefill",{}).get("roofline",{}).get("frac"))
print("comparator",d.get("gpu_comparator",{}).get("value"), d.get("gpu_comparator",{}).get("b200_e2e_over_ggml_cuda"))
print("cpu",d.get("cpu_baseline",{}).get("value"))
PY
tail -2 gpurun_out/bench30.err
