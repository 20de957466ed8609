"""Helpers for the -m gpu tests: torch supplies device memory only; every compute call goes through the C ABI."""
import ctypes as C

import numpy as np
import torch


def dev_u8(a: np.ndarray, pad: int = 64) -> torch.Tensor:
    a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    t = torch.zeros(a.size + pad, dtype=torch.uint8, device="cuda")
    t[: a.size] = torch.from_numpy(a)
    return t


def dev_f32(a) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def act_ws(lib, k: int) -> torch.Tensor:
    return torch.zeros(lib.c.pb200_act_workspace_bytes(k) + 64, dtype=torch.uint8, device="cuda")


def act_ws_fields(ws: torch.Tensor, k: int, mode: str):
    """Decodes the SoA activation workspace (csrc/common.cuh ActQ) back into the reference's block bytes."""
    kp = (k + 255) // 256 * 256
    raw = ws.cpu().numpy()
    qs = raw[:kp].view(np.int8)
    d = raw[kp: kp + kp // 32 * 4].view(np.float32)
    s = raw[kp + kp // 32 * 4: kp + kp // 32 * 8].view(np.float32)
    bs = raw[kp + kp // 32 * 8: kp + kp // 32 * 8 + kp // 16 * 2].view(np.int16)
    if mode == "q8_K":
        nb = k // 256
        out = np.zeros((nb, 292), dtype=np.uint8)
        out[:, 0:4] = d[:nb].view(np.uint8).reshape(nb, 4)
        out[:, 4:260] = qs[: nb * 256].view(np.uint8).reshape(nb, 256)
        out[:, 260:292] = bs[: nb * 16].view(np.uint8).reshape(nb, 32)
        return out.reshape(-1)
    nb = k // 32
    if mode == "q8_0":
        out = np.zeros((nb, 34), dtype=np.uint8)
        out[:, 0:2] = d[:nb].astype(np.float16).view(np.uint8).reshape(nb, 2)
        out[:, 2:] = qs[: nb * 32].view(np.uint8).reshape(nb, 32)
        return out.reshape(-1)
    out = np.zeros((nb, 36), dtype=np.uint8)
    out[:, 0:2] = d[:nb].astype(np.float16).view(np.uint8).reshape(nb, 2)
    out[:, 2:4] = s[:nb].astype(np.float16).view(np.uint8).reshape(nb, 2)
    out[:, 4:] = qs[: nb * 32].view(np.uint8).reshape(nb, 32)
    return out.reshape(-1)


def sync():
    torch.cuda.synchronize()
