mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_ggml_graph.py -x -q 2>&1 | tail -4) > gpurun_out/t16.log; tail -4 gpurun_out/t16.log
B="python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-boundary --pp 0 --no-gpu-comparator"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"])'
for cfg in "PB200_LIB=$PWD/tools/ab/lib_4a129e6.so" "PB200_GEMV_L2PF=0 PB200_GEMV_NEXT_KB=0" "PB200_GEMV_L2PF=4 PB200_GEMV_NEXT_KB=0" "PB200_GEMV_L2PF=0 PB200_GEMV_NEXT_KB=24" "X=1" "PB200_GEMV_L2PF=8 PB200_GEMV_NEXT_KB=28" "PB200_GEMV_L2PF=2 PB200_GEMV_NEXT_KB=12" "PB200_GEMV_L2PF=12 PB200_GEMV_NEXT_KB=28"; do echo "== $cfg"; env $cfg timeout 200 $B 2>&1 | tail -1 | python -c "$P" 2>&1 | tail -1; done 2>&1 | tee gpurun_out/ab16.log
(timeout 200 python tools/token_trace.py 8 128 2>&1 | tail -60) > gpurun_out/trace16.log; head -14 gpurun_out/trace16.log; tail -6 gpurun_out/trace16.log
