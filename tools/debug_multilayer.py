"""GPU debugging aid: localise the first layer/token where the engine leaves the oracle, then dump that layer's intermediates."""
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
arch = sys.argv[1] if len(sys.argv) > 1 else "llama"
NT = 6
tm = TinyModel(n_layer=3, n_embd=1024, n_head=8, n_head_kv=2, n_ff=2816, n_vocab=384, n_ctx=96, arch=arch, seed=11)
toks = [(i * 7919 + 13) % 384 for i in range(NT)]
hp = tm.hp; E, H, HK, D, F = hp["n_embd"], hp["n_head"], hp["n_head_kv"], 128, hp["n_ff"]
QD, EK = H * D, HK * D
T = tm.tensors
port_hidden = {}
for L in (1, 2, 3):
    tm.hp["n_layer"] = L
    _, hid = tm.port_decode(port, toks)
    port_hidden[L] = hid
    eng = tm.load_engine(pkg, layers=(0, L), with_embd=True, with_head=False)
    errs = []
    for i, t in enumerate(toks):
        eng.decode(int(t), i, None)
        errs.append(err(eng.hidden(), hid[i]))
    print(arch, "layers [0,%d)" % L, " ".join(f"{e:.1e}" for e in errs))
    eng.close()
tm.hp["n_layer"] = 3
def mm(name, N, K, x):
    t, a = T[name]; return port.mul_mat(t, a, N, K, x)[0]
for il in (1, 2):
    eng = tm.load_engine(pkg, layers=(il, il + 1), with_embd=False, with_head=False)
    Kc = np.zeros((hp["n_ctx"], EK), np.uint16); Vc = np.zeros((hp["n_ctx"], EK), np.uint16)
    p = f"blk.{il}."
    print(arch, "layer", il, "types", {k.split('.')[-2]: O.TYPE_NAME[T[k][0]] for k in T if k.startswith(p) and T[k][0] != 0})
    for i in range(NT):
        x = port_hidden[il][i]
        eng.set_hidden(x)
        eng.decode(0, i, None)
        xn = port.rms_norm(x, hp["rms_eps"]) * T[p + "attn_norm.weight"][1]
        q = mm(p + "attn_q.weight", QD, E, xn); k = mm(p + "attn_k.weight", EK, E, xn); v = mm(p + "attn_v.weight", EK, E, xn)
        if arch == "qwen2":
            q = q + T[p + "attn_q.bias"][1]; k = k + T[p + "attn_k.bias"][1]; v = v + T[p + "attn_v.bias"][1]
        qr = port.rope(q, H, D, hp["rope_mode"], i, freq_base=hp["rope_freq_base"], n_ctx_orig=hp["n_ctx_orig"])
        kr = port.rope(k, HK, D, hp["rope_mode"], i, freq_base=hp["rope_freq_base"], n_ctx_orig=hp["n_ctx_orig"])
        Kc[i] = O.f32_to_f16_bits(kr); Vc[i] = O.f32_to_f16_bits(v)
        att = port.attention_decode(qr, Kc, Vc, H, HK, D, i + 1, 1.0 / np.sqrt(D))
        x1 = mm(p + "attn_output.weight", E, QD, att) + x
        xn2 = port.rms_norm(x1, hp["rms_eps"]) * T[p + "ffn_norm.weight"][1]
        g = mm(p + "ffn_gate.weight", F, E, xn2); u = mm(p + "ffn_up.weight", F, E, xn2)
        act = port.silu_mul(g, u)
        x2 = mm(p + "ffn_down.weight", E, F, act) + x1
        # stage without embedding: x = x_in, x1 = x_a, x2 = x_b
        print(f"  tok {i}: k {err(eng.debug_read('k', EK), k):.2e} v {err(eng.debug_read('v', EK), v):.2e} q(rope) {err(eng.debug_read('q', QD), qr):.2e} "
              f"att {err(eng.debug_read('att', QD), att):.2e} x1 {err(eng.debug_read('x_a', E), x1):.2e} g {err(eng.debug_read('g', F), g):.2e} "
              f"u {err(eng.debug_read('u', F), u):.2e} x2 {err(eng.debug_read('x_b', E), x2):.2e} vs port-hidden {err(x2, port_hidden[il + 1][i]):.1e}")
    eng.close()
