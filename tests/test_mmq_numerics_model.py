"""CPU: the error bound that tests/test_gpu_mmq.py states for the tensor-core mat-mul follows from its arithmetic, it is not
fitted to measurements.  This restates, in numpy, exactly what csrc/mmq.cu feeds the tensor pipe for Q4_K weights —
    weight     = fp16( q * fp16(d * sc)  +  fp16(-(dmin * m)) )        one fused multiply-add, q exact
    activation = fp16( d8 * q8 )                                        q8_K quantization identical to the CPU backend's
— multiplies the two in float64 (the fp32 accumulation of the real kernel only adds summation-order noise) and checks the
result against the oracle's integer-dot value with the same bound and NMSE bar as the GPU test."""
import numpy as np

import oracle_lib as O


def q4k_fields(W, N, K):
    blk = W.reshape(N * K // 256, 144)
    d = blk[:, 0:2].copy().view(np.float16).astype(np.float32).reshape(-1)
    dmin = blk[:, 2:4].copy().view(np.float16).astype(np.float32).reshape(-1)
    sc12 = blk[:, 4:16].astype(np.int32)
    qs = blk[:, 16:144]
    sc = np.zeros((blk.shape[0], 8), np.int32)
    mn = np.zeros((blk.shape[0], 8), np.int32)
    for j in range(8):                                   # get_scale_min_k4, ggml-quants.c:1950-1958
        if j < 4:
            sc[:, j] = sc12[:, j] & 63
            mn[:, j] = sc12[:, j + 4] & 63
        else:
            sc[:, j] = (sc12[:, j + 4] & 0xF) | ((sc12[:, j - 4] >> 6) << 4)
            mn[:, j] = (sc12[:, j + 4] >> 4) | ((sc12[:, j] >> 6) << 4)
    q = np.zeros((blk.shape[0], 256), np.int32)
    for g in range(4):                                   # 64 elements per 32-byte group: low nibbles then high nibbles
        q[:, 64 * g: 64 * g + 32] = qs[:, 32 * g: 32 * g + 32] & 0xF
        q[:, 64 * g + 32: 64 * g + 64] = qs[:, 32 * g: 32 * g + 32] >> 4
    return d, dmin, sc, mn, q


def test_fp16_operand_model_stays_inside_the_stated_bound(port):
    t, N, K, T = O.Q4_K, 64, 1024, 24
    rng = np.random.default_rng(3)
    W = O.synth_blocks(t, N, K, seed=21)
    X = rng.standard_normal((T, K)).astype(np.float32)
    d, dmin, sc, mn, q = q4k_fields(W, N, K)
    scale_h = (d[:, None] * sc.astype(np.float32)).astype(np.float16)                    # fp32 product, rounded to fp16
    off_h = (-(dmin[:, None] * mn.astype(np.float32))).astype(np.float16)
    w64 = q.reshape(-1, 8, 32).astype(np.float64) * scale_h.astype(np.float64)[:, :, None] + off_h.astype(np.float64)[:, :, None]
    w_h = w64.astype(np.float16).astype(np.float64).reshape(N, K)                        # the fused multiply-add's single rounding
    # the model's exact-q claim: the unrounded expression equals the reference dequantization up to the two scale roundings
    Wf = port.dequantize(t, W, N * K).reshape(N, K)
    assert np.max(np.abs(w_h - Wf)) <= 2.0 ** -9 * 2 * np.max(np.abs(Wf))
    a_h = np.zeros((T, K), np.float64)
    for i in range(T):
        blocks = port.quantize_act(t, X[i]).reshape(K // 256, 292)                       # block_q8_K: f32 d | 256 x i8 | 16 x i16
        d8 = blocks[:, 0:4].copy().view(np.float32).reshape(-1)
        q8 = blocks[:, 4:260].view(np.int8).astype(np.float32)
        a_h[i] = (d8[:, None] * q8).astype(np.float16).astype(np.float64).reshape(-1)    # fp32 product, rounded to fp16
    model = a_h @ w_h.T
    want = port.mul_mat(t, W, N, K, X).astype(np.float64)
    err = np.abs(model - want)
    sub = np.repeat(np.abs(Wf).reshape(N, K // 32, 32).max(axis=2), 32, axis=1)
    bound = 2.0 ** -9 * (np.abs(X).astype(np.float64) @ (np.abs(Wf) + sub).T) + 1e-6
    assert (err <= bound).all(), float((err / bound).max())
    nmse = float(np.sum((model - want) ** 2) / np.sum(want ** 2))
    assert nmse <= 4e-6, nmse
