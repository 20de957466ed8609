mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_mmq.py tests/test_gpu_engine.py -x -q 2>&1 | tail -8) > gpurun_out/t18.log; tail -8 gpurun_out/t18.log
(timeout 300 python tools/mmq_probe.py 2>&1 | tail -14) > gpurun_out/mmq18.log; cat gpurun_out/mmq18.log
for T in 512 2048; do echo "== prefill T=$T"; (timeout 300 python tools/prefill_probe.py $T 4 2>&1 | tail -1); done | tee gpurun_out/pf18.log
echo "== min units 4"; PB200_MMQ_MIN_UNITS=4 timeout 300 python tools/prefill_probe.py 512 4 2>&1 | tail -1 | tee -a gpurun_out/pf18.log
echo "== min units 16"; PB200_MMQ_MIN_UNITS=16 timeout 300 python tools/prefill_probe.py 512 4 2>&1 | tail -1 | tee -a gpurun_out/pf18.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/pf18_launches.csv python tools/prefill_probe.py 512 2 > /dev/null 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/pf18_launches.csv")) if len(r)>10]
hdr=rows[0]; seq=[]
for r in rows[1:]:
    d=dict(zip(hdr,r)); seq.append((d["Kernel Name"][:40], d.get("Grid Size",""), float(d["Metric Value"])/1e3))
i=[k for k,s in enumerate(seq) if "k_iota_pos" in s[0]]
i=i[-1] if i else 0
tot=0
for s in seq[i:i+24]: print(s); tot+=s[2]
print("sum", tot)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mmq_tc -s 12 -c 2 -o gpurun_out/mmq18_full python tools/prefill_probe.py 512 2 > gpurun_out/ncu18.log 2>&1; ls -la gpurun_out/*.ncu-rep
