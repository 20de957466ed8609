"""GPU: per-CTA timeline of one k_gemv_kquant launch (%globaltimer stamps)."""
import sys, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import pkgload
pkg = pkgload.load(); lib = pkg.Lib.get()
lib.c.pb200_debug_set_trace.argtypes = [C.c_void_p]
K = 8192; t = 12
rowb = lib.c.pb200_row_bytes(t, K)
pool = torch.randint(0, 255, (1 << 30,), dtype=torch.uint8, device="cuda")
x = torch.randn(K, device="cuda")
ws = torch.zeros(lib.c.pb200_act_workspace_bytes(K) + 64, dtype=torch.uint8, device="cuda")
lib.check(lib.c.pb200_quantize_act(t, C.c_void_p(x.data_ptr()), K, C.c_void_p(ws.data_ptr()), None), "q")
y = torch.zeros(65536, device="cuda")
tr = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
lib.check(lib.c.pb200_debug_set_trace(C.c_void_p(tr.data_ptr())), "trace")
names = ["start", "after init+pdl_wait", "act loads issued", "ring fill issued", "act regs ready", "first tile landed", "done"]
for N in (1024, 8192, 57344):
    for rep in range(3):
        off = (rep * N * rowb) % (1 << 29)
        tr.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        lib.c.pb200_mul_mat_vec_q(t, C.c_void_p(pool.data_ptr() + off), N, K, C.c_void_p(ws.data_ptr()), C.c_void_p(y.data_ptr()), None, None, None)
        e1.record()
        torch.cuda.synchronize()
    a = tr.cpu().numpy().reshape(148, 8).astype(np.float64)
    a = a[a[:, 0] > 0]
    t0 = a[:, 0].min()
    print(f"N={N} event time {e0.elapsed_time(e1)*1e3:.1f} us; CTAs {len(a)}; stamps relative to the first CTA start (us): min / median / max")
    for k, nm in enumerate(names):
        col = (a[:, k] - t0) / 1e3
        print(f"   {nm:22s} {col.min():7.2f} {np.median(col):7.2f} {col.max():7.2f}")
