mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_ggml_backend.py -x -q -k "FLASH or MUL_MAT" 2>&1 | tail -8) > gpurun_out/t11.log; tail -8 gpurun_out/t11.log
(timeout 200 python tools/token_trace.py 8 128 2>&1 | tail -60) > gpurun_out/trace11.log; head -14 gpurun_out/trace11.log; tail -6 gpurun_out/trace11.log
(timeout 600 python bench.py --model llama3-8b --steps 64 --warmup 8 --no-cpu-baseline --pp 0 2>&1 | tail -1) > gpurun_out/bench11_8b.log; cut -c1-1500 gpurun_out/bench11_8b.log
(timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1) > gpurun_out/bench11.log; cut -c1-6000 gpurun_out/bench11.log
