mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_ggml_graph.py -x -q 2>&1 | tail -30) > gpurun_out/t7.log; tail -30 gpurun_out/t7.log
(timeout 300 python tests/ggml_graph_parity.py llama 8 2>&1 | tail -5) > gpurun_out/parity7.log; cat gpurun_out/parity7.log
(timeout 900 python -m pytest tests/test_gpu_ggml_backend.py -x -q 2>&1 | tail -5) > gpurun_out/t7b.log; tail -5 gpurun_out/t7b.log
(timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --pp 0 2>&1 | tail -2) > gpurun_out/bench7.log; cat gpurun_out/bench7.log
