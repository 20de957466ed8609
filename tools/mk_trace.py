"""GPU: per-phase timeline of the persistent token kernel (layer 1) on the 70B bench model."""
import sys, ctypes as C, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ["PB200_PERSISTENT"] = "1"
import numpy as np, torch
import pkgload
import bench
pkg = pkgload.load(); lib = pkg.Lib.get()
lib.c.pb200_debug_set_trace.argtypes = [C.c_void_p]
cfg = bench.MODELS["llama3-70b"]
hp = dict(cfg["hp"], n_ctx=512)
hp["n_layer"] = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng = pkg.Model(pkg.HParams(**hp), 0, (0, hp["n_layer"]), True, True)
eng.synth(0, 1); eng.finalize()
tr = torch.zeros(148 * 34, dtype=torch.int64, device="cuda")
for i in range(140):
    eng.decode_async(i % 1000, i)
eng.synchronize()
lib.check(lib.c.pb200_debug_set_trace(C.c_void_p(tr.data_ptr())), "trace")
eng.decode_async(5, 140); eng.synchronize()
raw = tr.cpu().numpy().astype(np.float64)
a = raw[:148 * 16].reshape(148, 16)
b = raw[148 * 16:148 * 32].reshape(148, 16)
c = raw[148 * 32:148 * 33]
t0 = a[:, 0].min()
names = ["qkv", "wo", "gate|up", "down", "qkv(next layer)"]
print("persistent kernel, layer 1; per phase: [after barrier(s)+desc] [act regs ready] [tiles consumed]; us relative to the first stamp: min / median / max over CTAs")
for p in range(5):
    for k, nm in enumerate(("start", "act ready", "done")):
        col = (a[:, p * 3 + k] - t0) / 1e3
        print(f"  {names[p]:16s} {nm:10s} {col.min():8.2f} {np.median(col):8.2f} {col.max():8.2f}")

print("inside the qkv prologue (rmsnorm): stamps rel. to phase start; then the down prologue (staging)")
for off, base_col, nm in ((0, 0, "qkv"), (8, 9, "down")):
    st = a[:, base_col]
    for k, what in enumerate(("loads issued", "deferred refills issued", "smem stored", "after bar / compute", "act regs loaded")):
        col = (b[:, off + k] - st) / 1e3
        if (b[:, off + k] > 0).all():
            print(f"  {nm:5s} {what:26s} {col.min():7.2f} {np.median(col):7.2f} {col.max():7.2f}")
col = (c - a[:, 12]) / 1e3
print("  last rmsnorm prologue seen (next-layer qkv): sum-of-squares barrier passed at", np.median(col))

print("attention block of layer 1 (rel. to qkv phase start): phase-4 done / barrier A passed / attention done / barrier B passed")
for slot, nm in ((5, "qkv tiles done"), (6, "barrier A passed"), (7, "attention done (or skipped)"), (13, "barrier B passed")):
    col = (b[:, slot] - a[:, 0].min()) / 1e3
    attn = col[:64]; rest = col[64:]
    print(f"  {nm:28s} attention CTAs {attn.min():7.2f} {np.median(attn):7.2f} {attn.max():7.2f} | other CTAs {rest.min():7.2f} {np.median(rest):7.2f} {rest.max():7.2f}")
