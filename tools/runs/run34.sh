mkdir -p gpurun_out
(timeout 100 python bench.py --model llama3-8b --steps 64 --warmup 8 --no-cpu-baseline --no-gpu-comparator 2>gpurun_out/bench34.err | tail -1) > gpurun_out/bench34_8b.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench34_8b.json').read().strip().splitlines()[-1])
print("8B value",d["value"],"e2e",d["e2e"]["value"],"whole",d["roofline"]["whole_step"]["frac"],"prefill",d.get("prefill",{}).get("ms"),d.get("prefill",{}).get("roofline",{}).get("frac"))
PY
