"""GPU probe: first-light correctness + timing of pb200_mul_mat_q (tcgen05 prefill GEMM).  python tools/mmq_probe.py"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O
import pkgload
from gpu_util import dev_f32, dev_u8, ptr, sync

pkg = pkgload.load()
lib = pkg.Lib.get()
port = O.Port()


def run(t, W, N, K, X):
    T = X.shape[0]
    Wd, xd = dev_u8(W), dev_f32(X)
    y = torch.full((T, N), float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.zeros(lib.c.pb200_mul_mat_q_workspace_bytes(K, T) + 64, dtype=torch.uint8, device="cuda")
    rc = lib.c.pb200_mul_mat_q(t, ptr(Wd), N, K, ptr(xd), K, T, ptr(y), None, None, ptr(ws), None)
    sync()
    return rc, y.cpu().numpy()


# timing at 70B shapes
for (t, N, K, T) in [(O.Q4_K, 8192, 8192, 512), (O.Q4_K, 28672, 8192, 512), (O.Q6_K, 8192, 28672, 512), (O.Q4_K, 8192, 8192, 128), (O.Q4_K, 28672, 8192, 2048), (O.Q4_K, 8192, 8192, 2048), (O.Q6_K, 8192, 28672, 2048), (O.Q5_K, 8192, 8192, 2048)]:
    rb = lib.c.pb200_row_bytes(t, K)
    Wd = torch.randint(0, 255, (N * rb + 64,), dtype=torch.uint8, device="cuda")
    xd = torch.randn((T, K), device="cuda")
    y = torch.zeros((T, N), device="cuda")
    ws = torch.zeros(lib.c.pb200_mul_mat_q_workspace_bytes(K, T) + 64, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        lib.c.pb200_mul_mat_q(t, ptr(Wd), N, K, ptr(xd), K, T, ptr(y), None, None, ptr(ws), None)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        lib.c.pb200_mul_mat_q(t, ptr(Wd), N, K, ptr(xd), K, T, ptr(y), None, None, ptr(ws), None)
    e1.record(); sync()
    ms = e0.elapsed_time(e1) / reps
    print(f"perf {O.TYPE_NAME[t]} N {N} K {K} T {T}: {ms:.3f} ms  {2.0 * N * K * T / ms / 1e9:.1f} TFLOP/s aborted {lib.c.pb200_aborted()}", flush=True)
