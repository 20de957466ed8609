mkdir -p gpurun_out
(timeout 1700 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8) > gpurun_out/t22.log; tail -8 gpurun_out/t22.log
(timeout 1500 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench22.err | tail -1) > gpurun_out/bench22.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench22.json').read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"],"whole",d["roofline"]["whole_step"]["frac"])
print("prefill",d.get("prefill",{}).get("ms"), d.get("prefill",{}).get("roofline",{}).get("frac"))
print("comparator",d.get("gpu_comparator"))
print("cpu",d.get("cpu_baseline",{}).get("value"))
PY
tail -3 gpurun_out/bench22.err
