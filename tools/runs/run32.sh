mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
(timeout 360 $TR bench.py --gpus 2 --steps 20 --warmup 5 --pp 1024 --n-ctx 2048 2>gpurun_out/bench32.err | tail -1) > gpurun_out/bench32_pp2.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench32_pp2.json').read().strip().splitlines()[-1])
print("N=2 llama70b value",d["value"],"latency",d.get("latency_b1",{}).get("value"),"exposed",d.get("pipeline",{}).get("exposed_frac"),"prefill",d.get("prefill",{}).get("value"),d.get("prefill",{}).get("ms"))
PY
grep -i "error\|Traceback" gpurun_out/bench32.err | head -3
