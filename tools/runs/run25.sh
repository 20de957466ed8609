mkdir -p gpurun_out
B="python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-boundary --pp 0 --no-gpu-comparator"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"])'
for cfg in "PB200_NO_TAIL=1" "X=1" "PB200_NO_TAIL=1" "X=1"; do echo "== $cfg"; env $cfg timeout 200 $B 2>&1 | tail -1 | python -c "$P" 2>&1 | tail -1; done 2>&1 | tee gpurun_out/ab25.log
(timeout 200 python tools/token_trace.py 8 128 2>&1 | tail -60) > gpurun_out/trace25.log; head -10 gpurun_out/trace25.log; tail -6 gpurun_out/trace25.log
