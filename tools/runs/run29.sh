mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
(timeout 420 $TR bench.py --gpus 2 --model qwen2.5-72b --steps 32 --warmup 4 --pp 2048 --n-ctx 4096 2>gpurun_out/bench29q.err | tail -1) > gpurun_out/bench29_qwen_pp2.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench29_qwen_pp2.json').read().strip().splitlines()[-1])
print("N=2 qwen72b value",d["value"],"latency",d.get("latency_b1",{}).get("value"),"prefill",d.get("prefill"))
PY
grep -v "^\s*$" gpurun_out/bench29q.err | grep -i "error\|failed\|Traceback" | head -5
