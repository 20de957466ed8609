"""GPU microbenchmark: fixed overhead and marginal bandwidth of the k-quant GEMV launch (time(N) = a + b N)."""
import sys, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import pkgload
pkg = pkgload.load(); lib = pkg.Lib.get()
K = 8192
t = 12  # Q4_K
rowb = lib.c.pb200_row_bytes(t, K)
pool_bytes = 1 << 30
pool = torch.randint(0, 255, (pool_bytes,), dtype=torch.uint8, device="cuda")
# sane fp16 scales are irrelevant for timing; avoid NaN/inf paths anyway by zeroing d/dmin bytes -> outputs 0
x = torch.randn(K, device="cuda")
ws = torch.zeros(lib.c.pb200_act_workspace_bytes(K) + 64, dtype=torch.uint8, device="cuda")
lib.check(lib.c.pb200_quantize_act(t, C.c_void_p(x.data_ptr()), K, C.c_void_p(ws.data_ptr()), None), "q")
y = torch.zeros(65536, device="cuda")
torch.cuda.synchronize()
def run(N, interleave, reps=40):
    nbytes = N * rowb
    ncopies = max(1, min(pool_bytes // nbytes, 64))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def go(n):
        for i in range(n):
            off = (i % ncopies) * nbytes
            lib.c.pb200_mul_mat_vec_q(t, C.c_void_p(pool.data_ptr() + off), N, K, C.c_void_p(ws.data_ptr()), C.c_void_p(y.data_ptr()), None, None, None)
            if interleave:
                lib.c.pb200_quantize_act(t, C.c_void_p(x.data_ptr()), K, C.c_void_p(ws.data_ptr()), None)
    go(5); torch.cuda.synchronize()
    ev0.record(); go(reps); ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps * 1e3
for inter in (False, True):
    print("interleaved with a small kernel" if inter else "GEMV back to back")
    res = []
    for N in (1024, 2048, 4096, 8192, 16384, 32768, 57344):
        us = run(N, inter)
        res.append((N, us))
        print(f"  N={N:6d} {N*rowb/1e6:8.1f} MB  {us:8.2f} us  {N*rowb/us/1e6:7.2f} TB/s")
    (n0, t0), (n1, t1) = res[3], res[-1]
    b = (t1 - t0) / (n1 - n0)
    print(f"  fit: fixed {t0 - b*n0:.2f} us, marginal {rowb/b/1e6:.2f} TB/s")
