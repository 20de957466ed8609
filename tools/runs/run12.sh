mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_ggml_graph.py -x -q 2>&1 | tail -6) > gpurun_out/t12.log; tail -6 gpurun_out/t12.log
(timeout 600 python bench.py --model llama3-8b --steps 64 --warmup 8 --no-cpu-baseline --pp 0 2>&1 | tail -1) > gpurun_out/bench12_8b.log; cut -c1-400 gpurun_out/bench12_8b.log
(timeout 900 python bench.py --steps 20 --warmup 5 --no-boundary --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench12.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench12.log').read().strip().splitlines()[-1])
print("value",d["value"],"prefill",d.get("prefill",{}).get("ms"), d.get("prefill",{}).get("roofline",{}).get("frac"))
PY
