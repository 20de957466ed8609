mkdir -p gpurun_out
(timeout 240 python tools/pp_stage_probe.py 1 2 2048 2>&1 | tail -12) | tee gpurun_out/pp27.log
(timeout 240 python tools/pp_stage_probe.py 0 2 2048 2>&1 | tail -5) | tee -a gpurun_out/pp27.log
(timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -4) | tee gpurun_out/t27.log
(timeout 200 python tools/b32_probe.py 2>&1 | tail -3) | tee gpurun_out/b32_27.log
B="python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-boundary --pp 0 --no-gpu-comparator"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"])'
for cfg in "X=1"; do echo "== $cfg (Q4_K dot: 4 chains)"; env $cfg timeout 200 $B 2>&1 | tail -1 | python -c "$P" 2>&1 | tail -1; done 2>&1 | tee gpurun_out/ab27.log
