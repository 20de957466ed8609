mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "small_block or kquant_vs" 2>&1 | tail -3) | tee gpurun_out/t28.log
(timeout 200 python tools/b32_probe.py 2>&1 | tail -3) | tee gpurun_out/b32_28.log
(timeout 600 python bench.py --model qwen2.5-72b --steps 32 --warmup 4 --no-cpu-baseline --no-boundary --no-gpu-comparator 2>gpurun_out/bench28q.err | tail -1) > gpurun_out/bench28_qwen_n1.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench28_qwen_n1.json').read().strip().splitlines()[-1])
print("qwen N=1 value",d["value"],"whole",d["roofline"]["whole_step"]["frac"],"prefill",d.get("prefill",{}).get("ms"),d.get("prefill",{}).get("roofline",{}).get("frac"))
PY
tail -2 gpurun_out/bench28q.err
