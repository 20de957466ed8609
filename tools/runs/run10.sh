mkdir -p gpurun_out
nvidia-smi -L | head -3
(timeout 600 python -m pytest tests/test_gpu_ring.py -x -q 2>&1 | tail -12) > gpurun_out/t10.log; tail -12 gpurun_out/t10.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -4) > gpurun_out/bench10.log; tail -4 gpurun_out/bench10.log | cut -c1-3000
