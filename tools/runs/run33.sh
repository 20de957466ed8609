mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523"
(timeout 300 $TR bench.py --gpus 8 --model qwen2.5-72b --steps 32 --warmup 4 --pp 2048 --n-ctx 4096 2>gpurun_out/bench33.err | tail -1) > gpurun_out/bench33_qwen_pp8.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench33_qwen_pp8.json').read().strip().splitlines()[-1])
print("N=8 qwen72b value",d["value"],"latency",d.get("latency_b1",{}).get("value"),"exposed",d.get("pipeline",{}).get("exposed_frac"),"prefill",d.get("prefill",{}).get("value"),d.get("prefill",{}).get("ms"))
PY
grep -i "error\|Traceback" gpurun_out/bench33.err | head -3
