mkdir -p gpurun_out
B="python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-boundary --pp 0 --no-gpu-comparator"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"])'
for cfg in "PB200_LIB=$PWD/tools/ab/lib_4a129e6.so" "PB200_LIB=$PWD/tools/ab/lib_4deba25.so" "PB200_LIB=$PWD/tools/ab/lib_4deba25.so PB200_NO_CLUSTER=1" "PB200_GEMV_L2PF=0 PB200_GEMV_NEXT_KB=0" "PB200_GEMV_L2PF=0 PB200_GEMV_NEXT_KB=0 PB200_NO_CLUSTER=1" "PB200_LIB=$PWD/tools/ab/lib_4a129e6.so"; do echo "== $cfg"; env $cfg timeout 200 $B 2>&1 | tail -1 | python -c "$P" 2>&1 | tail -1; done 2>&1 | tee gpurun_out/ab15.log
for pf in 0 4; do echo "== ncu L2PF=$pf"; PB200_GEMV_L2PF=$pf timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum --clock-control none -k regex:k_gemv_kquant -s 18 -c 18 --csv --log-file gpurun_out/ncu15_pf$pf.csv python tools/gemv_probe.py 2 > /dev/null 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu15_pf$pf.csv")) if len(r)>10]
hdr=rows[0]; 
import collections
d=collections.OrderedDict()
for r in rows[1:]:
    rec=dict(zip(hdr,r)); d.setdefault(rec["ID"],{})[rec["Metric Name"]]=rec["Metric Value"]
for k,v in d.items(): print(k, v)
PY
done 2>&1 | tee gpurun_out/ncu15.log
(timeout 600 python bench.py --impl ggml-cuda --steps 32 --warmup 4 2>&1 | tail -3) > gpurun_out/cudaref15.log; tail -c 1800 gpurun_out/cudaref15.log
