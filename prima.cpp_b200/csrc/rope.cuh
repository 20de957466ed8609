// prima.cpp_b200/csrc/rope.cuh — RoPE device helpers shared by ops.cu and the persistent token kernel (ggml.c:14087-14266).
#pragma once
#include "launch.h"

namespace pb {

__device__ __forceinline__ float rope_yarn_ramp(float low, float high, int i0) {
    const float y = (i0 / 2 - low) / fmaxf(0.001f, high - low);
    return 1.0f - fminf(1.0f, fmaxf(0.0f, y));
}
__device__ __forceinline__ void rope_cos_sin(const RopeParams & rp, int32_t pos, int pair, const float * freq_factors, float & c, float & s) {
    float theta = (float) pos;
    for (int j = 0; j < pair; j++) theta = __fmul_rn(theta, rp.theta_scale);
    const float ff = freq_factors ? freq_factors[pair] : 1.0f;
    const float theta_extrap = __fdiv_rn(theta, ff);
    const float theta_interp = __fmul_rn(rp.freq_scale, theta_extrap);
    float th = theta_interp, mscale = rp.attn_factor;
    if (rp.ext_factor != 0.0f) {
        const float ramp_mix = rope_yarn_ramp(rp.corr_dims[0], rp.corr_dims[1], 2 * pair) * rp.ext_factor;
        th = theta_interp * (1 - ramp_mix) + theta_extrap * ramp_mix;
        mscale *= 1.0f + 0.1f * logf(1.0f / rp.freq_scale);
    }
    c = __fmul_rn(cosf(th), mscale);
    s = __fmul_rn(sinf(th), mscale);
}
__device__ __forceinline__ void rope_rotate(float x0, float x1, float c, float s, float & y0, float & y1) {
    y0 = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
    y1 = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
}

}  // namespace pb
