mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_ggml_graph.py tests/test_gpu_mmq.py -x -q 2>&1 | tail -6) > gpurun_out/t13.log; tail -6 gpurun_out/t13.log
B="python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-boundary --pp 0"
for cfg in "X=1" "PB200_NO_CLUSTER=1" "PB200_NO_DIST=1"; do echo "== $cfg"; env $cfg timeout 200 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"value\"], d[\"ms_per_step\"], d[\"roofline\"][\"frac\"], d[\"roofline\"][\"whole_step\"][\"frac\"])" 2>&1 | tail -1; done 2>&1 | tee gpurun_out/ab13.log
(timeout 200 python tools/token_trace.py 8 128 2>&1 | tail -60) > gpurun_out/trace13.log; head -14 gpurun_out/trace13.log; tail -6 gpurun_out/trace13.log
(timeout 900 python bench.py --steps 20 --warmup 5 --no-boundary --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench13.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench13.log').read().strip().splitlines()[-1])
print("value",d["value"],"prefill",d.get("prefill",{}).get("ms"), d.get("prefill",{}).get("roofline",{}).get("frac"))
PY
