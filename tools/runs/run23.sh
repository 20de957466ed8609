mkdir -p gpurun_out
LD_PRELOAD=$PWD/prima.cpp_b200/libggml-b200.so timeout 600 oracle/_ref/v3/test-backend-ops perf -b B200_0 -o MUL_MAT > gpurun_out/tbo_perf23.log 2>&1; echo rc=$?; tail -25 gpurun_out/tbo_perf23.log
(timeout 600 python -m pytest tests/test_gguf.py tests/test_gpu_ggml_graph.py tests/test_gpu_ring.py -q -m gpu 2>&1 | tail -6)
B="python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-boundary --pp 0 --no-gpu-comparator"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"])'
for cfg in "PB200_LIB=$PWD/tools/ab/lib_4a129e6.so" "X=1"; do echo "== $cfg"; env $cfg timeout 200 $B 2>&1 | tail -1 | python -c "$P" 2>&1 | tail -1; done 2>&1 | tee gpurun_out/ab23.log
