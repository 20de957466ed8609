"""GPU debugging aid: per-token error growth of the engine vs the port oracle, on a tiny model."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle_lib as O
import pkgload
from tiny_model import TinyModel
pkg = pkgload.load()
port = O.Port()
for arch, graph in (("llama", True), ("llama", False), ("qwen2", True)):
    tm = TinyModel(n_layer=3, n_embd=1024, n_head=8, n_head_kv=2, n_ff=2816, n_vocab=384, n_ctx=96, arch=arch, seed=11)
    toks = [(i * 7919 + 13) % 384 for i in range(24)]
    want, hw = tm.port_decode(port, toks)
    eng = tm.load_engine(pkg)
    eng.set_use_graph(graph)
    got = np.zeros_like(want)
    print(arch, "graph" if graph else "direct")
    for i, t in enumerate(toks):
        eng.decode(int(t), i, got[i])
        h = eng.hidden()
        print(f"  tok {i:2d} logits err {np.max(np.abs(got[i]-want[i])):.3e} hidden err {np.max(np.abs(h-hw[i])):.3e} |hidden| {np.max(np.abs(hw[i])):.2f}")
    eng.close()

# standalone attention + rope sweeps
import torch, ctypes as C
from gpu_util import dev_f32, ptr, sync
lib = pkg.Lib.get()
H, HK, D, n_ctx = 8, 2, 128, 64
rng = np.random.default_rng(0)
q = rng.standard_normal(H * D).astype(np.float32)
Kc = (rng.standard_normal((n_ctx, HK * D)) * 0.5).astype(np.float16)
Vc = rng.standard_normal((n_ctx, HK * D)).astype(np.float16)
qd, kd, vd = dev_f32(q), torch.from_numpy(Kc).cuda(), torch.from_numpy(Vc).cuda()
for n_kv in (1, 2, 5, 8, 9, 16, 17, 40):
    out = torch.zeros(H * D, device="cuda")
    pos = torch.tensor([n_kv - 1], dtype=torch.int32, device="cuda")
    lib.check(lib.c.pb200_attn_decode(ptr(qd), ptr(kd), ptr(vd), ptr(out), H, HK, D, ptr(pos), n_ctx, 1.0 / np.sqrt(D), None), "attn")
    sync()
    want = port.attention_decode(q, Kc.view(np.uint16), Vc.view(np.uint16), H, HK, D, n_kv, 1.0 / np.sqrt(D))
    print(f"attn n_kv={n_kv:3d} max err {np.max(np.abs(out.cpu().numpy()-want)):.3e}")
x = rng.standard_normal((1, 4, 128)).astype(np.float32)
xd = dev_f32(x)
for p in (0, 1, 2, 5, 9, 17, 33, 100, 1000):
    y = torch.zeros(4 * 128, device="cuda")
    pd = torch.tensor([p], dtype=torch.int32, device="cuda")
    for mode in (0, 2):
        lib.check(lib.c.pb200_rope(ptr(xd), ptr(y), 1, 4, 128, 128, mode, ptr(pd), 500000.0, 1.0, 0.0, 1.0, 32.0, 1.0, 8192, None, None), "rope")
        sync()
        want = port.rope(x[0], 4, 128, mode, p)
        print(f"rope pos={p:5d} mode={mode} max err {np.max(np.abs(y.cpu().numpy().reshape(4,128)-want)):.3e}")
