"""GPU (one device): what ONE pipeline rank does during bench.py --gpus N --pp P, without NCCL — the shard [l0, l1) of the 70B bench model,
micro-batches of 512 tokens through pb200_prefill_stage from a device buffer of hidden states.  python tools/pp_stage_probe.py [rank] [world] [pp]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import pkgload
pkg = pkgload.load()
rank = int(sys.argv[1]) if len(sys.argv) > 1 else 1
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pp = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
L = 80
hp = dict(n_layer=L, n_embd=8192, n_head=64, n_head_kv=8, head_dim=128, n_ff=28672, n_vocab=128256, n_ctx=4096, rope_mode=0,
          n_ctx_orig=8192, rope_freq_base=500000.0, rope_freq_scale=1.0, rms_eps=1e-5)
b = [round(r * L / world) for r in range(world + 1)]
eng = pkg.Model(pkg.HParams(**hp), 0, (b[rank], b[rank + 1]), with_embd=(rank == 0), with_head=(rank == world - 1))
eng.synth(0, 1234 + rank)
eng.set_n_seq(world)
eng.finalize()
E = hp["n_embd"]
hbuf = (torch.randn((512, E), device="cuda") * 0.1).contiguous()
toks = np.array([(i * 7919 + 13) % hp["n_vocab"] for i in range(pp)], dtype=np.int32)
ext = torch.cuda.ExternalStream(eng.stream)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(ext):
        for j in range((pp + 511) // 512):
            n = min(512, pp - j * 512)
            eng.prefill_stage(toks[j * 512:j * 512 + n] if rank == 0 else None, hbuf.data_ptr() if rank > 0 else None, n, j * 512)
    torch.cuda.synchronize()
    print(f"rank {rank}/{world} pp {pp}: pass {rep} {1e3 * (time.perf_counter() - t0):.2f} ms, aborted {eng.lib.c.pb200_aborted()}", flush=True)
