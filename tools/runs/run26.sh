mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_ring.py -x -q 2>&1 | tail -4) > gpurun_out/t26.log; tail -4 gpurun_out/t26.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
(timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --pp 2048 --n-ctx 4096 2>gpurun_out/bench26.err | tail -1) > gpurun_out/bench26_pp2.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench26_pp2.json').read().strip().splitlines()[-1])
print("N=2 llama70b value",d["value"],"latency",d.get("latency_b1",{}).get("value"),"prefill",d.get("prefill"))
PY
tail -3 gpurun_out/bench26.err
(timeout 900 $TR bench.py --gpus 2 --model qwen2.5-72b --steps 32 --warmup 4 --pp 2048 --n-ctx 4096 2>gpurun_out/bench26q.err | tail -1) > gpurun_out/bench26_qwen_pp2.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench26_qwen_pp2.json').read().strip().splitlines()[-1])
print("N=2 qwen72b value",d["value"],"latency",d.get("latency_b1",{}).get("value"),"prefill",d.get("prefill"))
PY
tail -3 gpurun_out/bench26q.err
