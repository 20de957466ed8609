mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_mmq.py -x -q 2>&1 | tail -5) > gpurun_out/t20.log; tail -5 gpurun_out/t20.log
echo "== default (CG2, A4)"; (timeout 300 python tools/mmq_probe.py 2>&1 | tail -8) > gpurun_out/mmq20.log; cat gpurun_out/mmq20.log
echo "== CG=2 A=2"; (PB200_MMQ_A_NST=2 timeout 300 python tools/mmq_probe.py 2>&1 | tail -8) > gpurun_out/mmq20_a2.log; cat gpurun_out/mmq20_a2.log
echo "== CG=1"; (PB200_MMQ_CG=1 timeout 300 python tools/mmq_probe.py 2>&1 | tail -8) > gpurun_out/mmq20_cg1.log; cat gpurun_out/mmq20_cg1.log
for T in 512 2048; do echo "== prefill T=$T"; (timeout 300 python tools/prefill_probe.py $T 4 2>&1 | tail -1); done | tee gpurun_out/pf20.log
(timeout 900 python -m pytest tests/test_gpu_engine.py -x -q 2>&1 | tail -5) > gpurun_out/t20b.log; tail -5 gpurun_out/t20b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mmq_tc -s 12 -c 1 -o gpurun_out/mmq20_full python tools/prefill_probe.py 512 2 > gpurun_out/ncu20.log 2>&1; ls -la gpurun_out/mmq20_full.ncu-rep
