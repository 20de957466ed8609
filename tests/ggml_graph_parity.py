"""Driver (run in a process of its own by tests/test_gpu_ggml_graph.py): a whole build_llama / build_qwen2 decode graph through
ggml_backend_graph_compute on the registered "B200_0" backend vs the reference CPU backend, same weights, same tokens, HOST tensors
in and out (ggml_backend_tensor_set / _get) — what llama_decode does.  Prints one JSON line.

    python tests/ggml_graph_parity.py <llama|qwen2> <n_tokens> [n_layer n_embd n_head n_head_kv n_ff n_vocab n_ctx]
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "host"))
from tiny_model import TinyModel   # noqa: E402  (weights as raw GGUF blocks; loads no ggml)
import host_graph as HG            # noqa: E402

arch = sys.argv[1]
n_tok = int(sys.argv[2])
dims = [int(x) for x in sys.argv[3:10]] if len(sys.argv) >= 10 else [3, 1024, 8, 2, 2816 if arch == "llama" else 3072, 384, 96]
n_layer, n_embd, n_head, n_head_kv, n_ff, n_vocab, n_ctx = dims
tm = TinyModel(n_layer=n_layer, n_embd=n_embd, n_head=n_head, n_head_kv=n_head_kv, n_ff=n_ff, n_vocab=n_vocab, n_ctx=n_ctx, arch=arch,
               ftype="q4_K_M" if arch == "llama" else "q5_K_M", freq_factors=(arch == "llama"), seed=31, branch_scale=0.1)
types = {name: t for name, (t, a) in tm.tensors.items() if name.endswith(".weight") and a.dtype == np.uint8}
has_bias = arch == "qwen2"
has_ff = "rope_freqs.weight" in tm.tensors
g, plug = HG.load(with_plugin=True)
libpb = C.CDLL(str(ROOT / "prima.cpp_b200" / "libprima_b200.so"))
libpb.pb200_kernel_launches.restype = C.c_uint64
models = {}
for be in ("CPU", "B200_0"):
    m = HG.HostModel(tm.hp, types, be, n_threads=8, has_bias=has_bias, has_freq_factors=has_ff)
    for name, (t, a) in tm.tensors.items():
        m.set_tensor(name, a)
    models[be] = m
toks = [(i * 7919 + 13) % n_vocab for i in range(n_tok)]
out = {be: np.zeros((n_tok, n_vocab), np.float32) for be in models}
hid = {be: np.zeros((n_tok, n_embd), np.float32) for be in models}
launches = []
fused0 = plug.ggml_backend_b200_fused_steps()
for i, t in enumerate(toks):
    for be, m in models.items():
        n0 = libpb.pb200_kernel_launches()
        m.decode([t], i, out[be][i])
        if be == "B200_0":
            launches.append(int(libpb.pb200_kernel_launches() - n0))
        hid[be][i] = m.hidden()[0]
EK = n_head_kv * 128
kv_err = 0.0
for il in range(n_layer):
    for nm in ("cache_k", "cache_v"):
        a = models["CPU"].get_tensor(f"blk.{il}.{nm}", EK * n_ctx * 2).view(np.float16).astype(np.float32)
        b = models["B200_0"].get_tensor(f"blk.{il}.{nm}", EK * n_ctx * 2).view(np.float16).astype(np.float32)
        kv_err = max(kv_err, float(np.max(np.abs(a - b))))
e = np.max(np.abs(out["CPU"] - out["B200_0"]), axis=1)
res = {"arch": arch, "n_tokens": n_tok, "max_abs_per_token": [float(x) for x in e], "max_abs": float(e.max()), "first_token_err": float(e[0]),
       "nmse": float(np.sum((out["CPU"] - out["B200_0"]) ** 2) / np.sum(out["CPU"] ** 2)),
       "hidden_max_abs": float(np.max(np.abs(hid["CPU"] - hid["B200_0"]))), "kv_max_abs": kv_err,
       "argmax_agree": float(np.mean(out["CPU"].argmax(1) == out["B200_0"].argmax(1))),
       "launches_per_token": launches, "n_layer": n_layer, "graph_nodes": models["B200_0"].graph_nodes,
       "fused_steps": int(plug.ggml_backend_b200_fused_steps() - fused0), "graph_builds": int(models["B200_0"].graph_builds),
       "graph_replays": int(plug.ggml_backend_b200_graph_replays())}
print(json.dumps(res))
