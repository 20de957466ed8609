"""GPU debugging aid: 1-layer model, every intermediate buffer of the engine vs the port oracle."""
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
def err(a, b): return float(np.max(np.abs(a - b)))
for arch in ("llama", "qwen2"):
    tm = TinyModel(n_layer=1, n_embd=1024, n_head=8, n_head_kv=2, n_ff=2816, n_vocab=384, n_ctx=96, arch=arch, seed=11)
    hp = tm.hp; E, H, HK, D, F = hp["n_embd"], hp["n_head"], hp["n_head_kv"], 128, hp["n_ff"]
    QD, EK = H * D, HK * D
    T = tm.tensors
    def mm(name, N, K, x):
        t, a = T[name]; return port.mul_mat(t, a, N, K, x)[0]
    eng = tm.load_engine(pkg)
    Kc = np.zeros((hp["n_ctx"], EK), np.uint16); Vc = np.zeros((hp["n_ctx"], EK), np.uint16)
    print(arch)
    for i in range(6):
        tok = (i * 7919 + 13) % 384
        logits = np.zeros(384, np.float32)
        eng.decode(tok, i, logits)
        t, a = T["token_embd.weight"]
        x = port.dequantize(t, a.reshape(384, -1)[tok], E)[0]
        xn = port.rms_norm(x, hp["rms_eps"]) * T["blk.0.attn_norm.weight"][1]
        q = mm("blk.0.attn_q.weight", QD, E, xn); k = mm("blk.0.attn_k.weight", EK, E, xn); v = mm("blk.0.attn_v.weight", EK, E, xn)
        if arch == "qwen2":
            q = q + T["blk.0.attn_q.bias"][1]; k = k + T["blk.0.attn_k.bias"][1]; v = v + T["blk.0.attn_v.bias"][1]
        qr = port.rope(q, H, D, hp["rope_mode"], i, freq_base=hp["rope_freq_base"], n_ctx_orig=hp["n_ctx_orig"])
        kr = port.rope(k, HK, D, hp["rope_mode"], i, freq_base=hp["rope_freq_base"], n_ctx_orig=hp["n_ctx_orig"])
        Kc[i] = O.f32_to_f16_bits(kr); Vc[i] = O.f32_to_f16_bits(v)
        att = port.attention_decode(qr, Kc, Vc, H, HK, D, i + 1, 1.0 / np.sqrt(D))
        x1 = mm("blk.0.attn_output.weight", E, QD, att) + x
        xn2 = port.rms_norm(x1, hp["rms_eps"]) * T["blk.0.ffn_norm.weight"][1]
        g = mm("blk.0.ffn_gate.weight", F, E, xn2); u = mm("blk.0.ffn_up.weight", F, E, xn2)
        act = port.silu_mul(g, u)
        x2 = mm("blk.0.ffn_down.weight", E, F, act) + x1
        # engine: x=x_a, x1=x_b, x2=xn  (see enqueue_step rotation); final copy xn -> x_b overwrites x1
        print(f"  tok {i} id {tok:3d}: x {err(eng.debug_read('x_a', E), x):.2e} k {err(eng.debug_read('k', EK), k):.2e} v {err(eng.debug_read('v', EK), v):.2e} "
              f"q(rope) {err(eng.debug_read('q', QD), qr):.2e} att {err(eng.debug_read('att', QD), att):.2e} g {err(eng.debug_read('g', F), g):.2e} "
              f"u {err(eng.debug_read('u', F), u):.2e} x2 {err(eng.debug_read('xn', E), x2):.2e}  |x1| {np.abs(x1).max():.1f}")
    eng.close()
