mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_mmq.py -x -q 2>&1 | tail -8) > gpurun_out/t19.log; tail -8 gpurun_out/t19.log
(timeout 300 python tools/mmq_probe.py 2>&1 | tail -9) > gpurun_out/mmq19.log; cat gpurun_out/mmq19.log
echo "== CG=1"; (PB200_MMQ_CG=1 timeout 300 python tools/mmq_probe.py 2>&1 | tail -8) > gpurun_out/mmq19_cg1.log; cat gpurun_out/mmq19_cg1.log
for T in 512 2048; do echo "== prefill T=$T"; (timeout 300 python tools/prefill_probe.py $T 4 2>&1 | tail -1); done | tee gpurun_out/pf19.log
(timeout 900 python -m pytest tests/test_gpu_engine.py -x -q 2>&1 | tail -5) > gpurun_out/t19b.log; tail -5 gpurun_out/t19b.log
