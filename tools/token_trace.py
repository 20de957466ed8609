"""GPU: per-launch timeline of the GEMV kernels of ONE decoded token (PDL-chained direct launches, instrumented kernel).

    python tools/token_trace.py [n_layers=8] [prompt=128]

For every k_gemv_kquant launch of the token: when its CTAs became resident relative to the previous GEMV's end (negative =
overlap under programmatic dependent launch), when griddepcontrol.wait returned, when the activation was in registers, when
the first tile had landed, when the launch was done — min / median / max over CTAs, microseconds (%globaltimer)."""
import sys, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import pkgload
pkg = pkgload.load(); lib = pkg.Lib.get()
lib.c.pb200_debug_set_trace.argtypes = [C.c_void_p, C.c_int]
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prompt = int(sys.argv[2]) if len(sys.argv) > 2 else 128
hp = dict(n_layer=L, n_embd=8192, n_head=64, n_head_kv=8, head_dim=128, n_ff=28672, n_vocab=128256, n_ctx=512, rope_mode=0,
          n_ctx_orig=8192, rope_freq_base=500000.0, rope_freq_scale=1.0, rms_eps=1e-5)
eng = pkg.Model(pkg.HParams(**hp), 0)
eng.synth(0, 1234)
eng.finalize()
for i in range(prompt):
    eng.decode_async((i * 7919 + 13) % hp["n_vocab"], i)
eng.synchronize()
nl = 4 * L + 1
ROW = 4096
tr = torch.zeros(nl * ROW, dtype=torch.int64, device="cuda")
eng.set_use_graph(False)
for rep in range(3):
    tr.zero_()
    lib.check(lib.c.pb200_debug_set_trace(C.c_void_p(tr.data_ptr()), nl), "trace")
    eng.decode_async(4242, prompt + rep)
    eng.synchronize()
lib.c.pb200_debug_set_trace(None, 0)
a = tr.cpu().numpy().reshape(nl, ROW // 8, 8).astype(np.float64)
names = ["qkv", "wo", "gate|up", "down"]
t_tok = a[0][a[0][:, 0] > 0][:, 0].min()
print(f"{L} layers of Llama-3-70B Q4_K_M, n_kv = {prompt + 3}; times in us; per launch min/median/max over CTAs")
print("launch        ctas  start(rel prev done)      fill issued   wait returned    act in regs   first tile      done      | launch span  clk GHz")
prev_done = None
rows = []
for i in range(nl):
    r = a[i][a[i][:, 0] > 0]
    nm = "head" if i == nl - 1 else f"L{i // 4}.{names[i % 4]}"
    s0 = r[:, 0].min()
    def mm(col, base):
        c = (r[:, col] - base) / 1e3
        return f"{c.min():6.1f}/{np.median(c):6.1f}/{c.max():6.1f}"
    done = r[:, 5].max()
    clk = np.median((r[:, 7] - r[:, 6]) / np.maximum(r[:, 5] - r[:, 0], 1))
    rel = "      -       " if prev_done is None else mm(0, prev_done)
    print(f"{nm:12s} {len(r):4d}  {rel:22s} {mm(1, s0)}  {mm(2, s0)}  {mm(3, s0)}  {mm(4, s0)}  {mm(5, s0)} | {(done - s0) / 1e3:7.1f}   {clk:5.2f}")
    rows.append((nm, s0, done))
    prev_done = done
tot = (rows[-1][2] - rows[0][1]) / 1e3
print(f"token: first GEMV start -> head done {tot:.1f} us  ({tot / 1e3:.3f} ms for {L} layers + head)")
per = {}
for (nm, s0, done), nxt in zip(rows[:-1], rows[1:]):
    k = nm.split(".")[-1]
    per.setdefault(k, []).append((nxt[1] - s0) / 1e3)   # start-to-start: what the launch costs in the chain
for k, v in per.items():
    print(f"  {k:8s} start-to-next-start: median {np.median(v):6.1f} us")
