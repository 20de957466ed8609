"""GPU: one pb200_prefill batch of the 70B bench model (for ncu launch lists / quick timing).  python tools/prefill_probe.py [T] [layers]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import pkgload
pkg = pkgload.load()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
hp = dict(n_layer=L, n_embd=8192, n_head=64, n_head_kv=8, head_dim=128, n_ff=28672, n_vocab=128256, n_ctx=max(512, T), rope_mode=0,
          n_ctx_orig=8192, rope_freq_base=500000.0, rope_freq_scale=1.0, rms_eps=1e-5)
eng = pkg.Model(pkg.HParams(**hp))
eng.synth(0, 1234); eng.finalize()
toks = np.array([(i * 7919 + 13) % hp["n_vocab"] for i in range(T)], dtype=np.int32)
eng.prefill(toks, 0)
t0 = time.perf_counter(); eng.prefill(toks, 0); dt = time.perf_counter() - t0
print(f"prefill T={T} layers={L}: {dt * 1e3:.2f} ms  ({dt * 1e3 / L:.3f} ms/layer)")
