// prima.cpp_b200/csrc/mmq.cu — batched (prefill) k-quant mat-mul on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces: ggml_cuda_op_mul_mat_q -> mul_mat_q<type,mmq_x,8,chk> (ggml-cuda/mmq.cu:3-98, mmq.cuh:2583-2650: int8 mma.sync
// tiles with __syncthreads ping-pong) and quantize_mmq_q8_1_cuda (quantize.cu:143-169), SURVEY §8 row a-4.
// Numerics follow the CPU backend the oracle restates: every activation row is quantized to q8_K exactly as
// quantize_row_q8_K_ref does (ggml-quants.c:3785-3822) and the weights are expanded with the dequantize_row_q{4,5,6}_K
// formulas (ggml-quants.c:2040-2065, 2390-2420, 2690-2725); both are then rounded to fp16 and multiplied on the tensor
// pipe with fp32 accumulation.  The result differs from the integer-dot CPU value only by those two fp16 roundings
// (NMSE ~1e-7; tests/test_gpu_mmq.py states the bound).
//
//   dst[t][n] = sum_k W[n][k] * X[t][k]        W: N x K k-quant rows, X: T x K f32, dst: T x N f32 (ggml layout)
//
// One CTA owns a 128-row x BN-column tile of dst (BN <= 256 tokens) and walks K in 64-element steps:
//   warp 0      producer: cp.async.bulk of the raw quantized blocks (one per row per 256-K super-block) into a 2-deep ring,
//               and of the pre-tiled fp16 activation chunk (BN x 128 B, already in the UMMA swizzle-128B image) per step
//   warp 1      owns TMEM; one lane issues 4 x tcgen05.mma (M=128, N=BN, K=16, kind::f16) per step and commits to mbarriers
//   warps 2..9  expand 128 x 64 weights per step to fp16 straight into the swizzled A stage (generic-proxy stores +
//               fence.proxy.async); afterwards warps 2..5 read the accumulator out of TMEM (tcgen05.ld 32x32b) and store dst
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "launch.h"

namespace pb {

constexpr int MMQ_BM = 128;
constexpr int MMQ_BK = 64;
constexpr int MMQ_NSTAGE = 3;
constexpr int MMQ_DQ_WARPS = 8;
constexpr int MMQ_THREADS = (2 + MMQ_DQ_WARPS) * 32;
constexpr int MMQ_A_BYTES = MMQ_BM * 128;   // one A stage: 128 rows x 64 fp16
constexpr int MMQ_CTL_BYTES = 256;

struct MmqParams {
    const uint8_t * W;
    const uint8_t * B;     // activations, fp16, tiled [T/BN][K/64][BN x 128 B swizzled]
    float * dst;           // [T][N]
    const float * bias;    // [N] or null
    int64_t row_bytes, total_bytes;
    int type, N, K, T, BN, bpb, slot;   // slot: bytes reserved per row in a raw stage (16-B aligned window around one block)
    uint32_t tmem_cols, idesc;
};

struct MmqCtl {
    uint64_t raw_full[2], raw_empty[2];
    uint64_t a_ready[MMQ_NSTAGE], b_full[MMQ_NSTAGE], stage_free[MMQ_NSTAGE];
    uint64_t acc_ready;
    uint32_t tmem_base;
    int abort;
};
static_assert(sizeof(MmqCtl) <= MMQ_CTL_BYTES, "ctl");

__device__ int g_mmq_abort;

__device__ __forceinline__ bool mmq_try(uint64_t * bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a broken pipeline must end the launch (and report through pb200_mmq_aborted), never hang the device
__device__ __forceinline__ bool mmq_wait(MmqCtl * ctl, uint64_t * bar, uint32_t parity) {
    const long long t0 = clock64();
    int spins = 0;
    while (!mmq_try(bar, parity)) {
        if ((++spins & 255) == 0) {
            if (*(volatile int *) &ctl->abort) return false;
            if (clock64() - t0 > (1ll << 27)) {
                *(volatile int *) &ctl->abort = 1;
                atomicExch(&g_mmq_abort, 1);
                return false;
            }
        }
    }
    return true;
}
__device__ __forceinline__ void bulk_g2s_plain(void * smem_dst, const void * gsrc, uint32_t bytes, uint64_t * bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    // K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
    uint64_t d = (uint64_t) ((saddr >> 4) & 0x3FFF);
    d |= (uint64_t) 1 << 16;               // leading byte offset: unused for swizzled K-major, canonical value 1
    d |= (uint64_t) (1024 >> 4) << 32;     // stride byte offset
    d |= (uint64_t) 1 << 46;               // descriptor version (Blackwell)
    d |= (uint64_t) 2 << 61;               // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t * bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
// 32-bit load from shared memory at an address that is only 2-byte aligned
__device__ __forceinline__ uint32_t lds32_u2(const uint8_t * p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t * w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t) 3);
    if ((a & 2) == 0) return w[0];
    return __funnelshift_r(w[0], w[1], 16);
}

// ---- weight expansion: thread (row, h) produces K elements [64c + 32h, 64c + 32h + 32) of its row as 16 half2 ----
__device__ __forceinline__ void expand_q4K(const uint8_t * blk, int c, int h, uint32_t (&out)[16]) {
    const float d = __half2float(*reinterpret_cast<const __half *>(blk));
    const float dmin = __half2float(*reinterpret_cast<const __half *>(blk + 2));
    const uint8_t * sc = blk + 4;
    const int j = 2 * c + h;
    int s, m;
    if (j < 4) { s = sc[j] & 63; m = sc[j + 4] & 63; }
    else { s = (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4); m = (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4); }
    const float d1 = __fmul_rn(d, (float) s), m1 = __fmul_rn(dmin, (float) m);
    const uint4 * q = reinterpret_cast<const uint4 *>(blk + 16 + 32 * c);
    const uint4 qa = q[0], qb = q[1];
    const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t v = (w[i] >> (4 * h)) & 0x0F0F0F0Fu;
        const float f0 = __fsub_rn(__fmul_rn(d1, (float) (v & 0xFF)), m1);
        const float f1 = __fsub_rn(__fmul_rn(d1, (float) ((v >> 8) & 0xFF)), m1);
        const float f2 = __fsub_rn(__fmul_rn(d1, (float) ((v >> 16) & 0xFF)), m1);
        const float f3 = __fsub_rn(__fmul_rn(d1, (float) (v >> 24)), m1);
        out[2 * i] = pack_h2(f0, f1);
        out[2 * i + 1] = pack_h2(f2, f3);
    }
}
__device__ __forceinline__ void expand_q5K(const uint8_t * blk, int c, int h, uint32_t (&out)[16]) {
    const float d = __half2float(*reinterpret_cast<const __half *>(blk));
    const float dmin = __half2float(*reinterpret_cast<const __half *>(blk + 2));
    const uint8_t * sc = blk + 4;
    const int j = 2 * c + h;
    int s, m;
    if (j < 4) { s = sc[j] & 63; m = sc[j + 4] & 63; }
    else { s = (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4); m = (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4); }
    const float d1 = __fmul_rn(d, (float) s), m1 = __fmul_rn(dmin, (float) m);
    const uint4 * qh4 = reinterpret_cast<const uint4 *>(blk + 16);
    const uint4 * q = reinterpret_cast<const uint4 *>(blk + 48 + 32 * c);
    const uint4 qa = q[0], qb = q[1], ha = qh4[0], hb = qh4[1];
    const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
    const uint32_t hh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t v = ((w[i] >> (4 * h)) & 0x0F0F0F0Fu) | (((hh[i] >> j) & 0x01010101u) << 4);
        const float f0 = __fsub_rn(__fmul_rn(d1, (float) (v & 0xFF)), m1);
        const float f1 = __fsub_rn(__fmul_rn(d1, (float) ((v >> 8) & 0xFF)), m1);
        const float f2 = __fsub_rn(__fmul_rn(d1, (float) ((v >> 16) & 0xFF)), m1);
        const float f3 = __fsub_rn(__fmul_rn(d1, (float) (v >> 24)), m1);
        out[2 * i] = pack_h2(f0, f1);
        out[2 * i + 1] = pack_h2(f2, f3);
    }
}
__device__ __forceinline__ void expand_q6K(const uint8_t * blk, int c, int h, uint32_t (&out)[16]) {
    // blk is 2-byte aligned only (210-byte blocks)
    const float d = __half2float(*reinterpret_cast<const __half *>(blk + 208));
    const int n = c >> 1, p = c & 1;
    const uint8_t * ql = blk + 64 * n + 32 * h;
    const uint8_t * qh = blk + 128 + 32 * n;
    const int8_t * sc = reinterpret_cast<const int8_t *>(blk + 192 + 8 * n + 2 * h + 4 * p);
    const float d0 = __fmul_rn(d, (float) sc[0]), d1 = __fmul_rn(d, (float) sc[1]);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t lw = lds32_u2(ql + 4 * i), hw = lds32_u2(qh + 4 * i);
        const uint32_t v = ((lw >> (4 * p)) & 0x0F0F0F0Fu) | (((hw >> (4 * p + 2 * h)) & 0x03030303u) << 4);
        const float dd = i < 4 ? d0 : d1;
        const float f0 = __fmul_rn(dd, (float) ((int) (v & 0xFF) - 32));
        const float f1 = __fmul_rn(dd, (float) ((int) ((v >> 8) & 0xFF) - 32));
        const float f2 = __fmul_rn(dd, (float) ((int) ((v >> 16) & 0xFF) - 32));
        const float f3 = __fmul_rn(dd, (float) ((int) (v >> 24) - 32));
        out[2 * i] = pack_h2(f0, f1);
        out[2 * i + 1] = pack_h2(f2, f3);
    }
}

__global__ void __launch_bounds__(MMQ_THREADS, 1) k_mmq_tc(const __grid_constant__ MmqParams P) {
    extern __shared__ uint8_t smem_raw[];
    // the swizzle-128B atoms (A and B stages) need 1024-byte alignment in the shared window
    uint8_t * smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int BN = P.BN;
    const int b_bytes = BN * 128;
    uint8_t * a_st = smem;                                       // [NSTAGE][16 KB]
    uint8_t * b_st = a_st + MMQ_NSTAGE * MMQ_A_BYTES;            // [NSTAGE][BN*128]
    uint8_t * raw = b_st + MMQ_NSTAGE * b_bytes;                 // [2][128 * slot]
    MmqCtl * ctl = reinterpret_cast<MmqCtl *>(raw + 2 * MMQ_BM * P.slot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * MMQ_BM;
    const int ttile = blockIdx.y;
    const int nsb = P.K / 256;
    const int nchunk = nsb * 4;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(&ctl->raw_full[i], 1); mbar_init(&ctl->raw_empty[i], MMQ_DQ_WARPS); }
        for (int i = 0; i < MMQ_NSTAGE; i++) { mbar_init(&ctl->a_ready[i], MMQ_DQ_WARPS); mbar_init(&ctl->b_full[i], 1); mbar_init(&ctl->stage_free[i], 1); }
        mbar_init(&ctl->acc_ready, 1);
        ctl->abort = 0;
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(P.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *(volatile uint32_t *) &ctl->tmem_base;

    if (warp == 0) {
        // ================= producer =================
        const uint8_t * Bt = P.B + (size_t) ttile * (P.K / MMQ_BK) * b_bytes;
        const int64_t lim = (P.total_bytes + 15) & ~(int64_t) 15;
        for (int sb = 0; sb < nsb; sb++) {
            const int rs = sb & 1, rr = sb >> 1;
            bool ok = true;
            if (rr > 0) ok = mmq_wait(ctl, &ctl->raw_empty[rs], (rr - 1) & 1);
            if (!ok) break;
            // one 16-B aligned window per row around its block of super-block sb
            uint32_t my_bytes = 0;
            int64_t a0[4];
            uint32_t nb[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int r = lane + 32 * i;
                const int gr = min(row0 + r, P.N - 1);
                const int64_t g0 = (int64_t) gr * P.row_bytes + (int64_t) sb * P.bpb;
                a0[i] = g0 & ~(int64_t) 15;
                int64_t a1 = (g0 + P.bpb + 15) & ~(int64_t) 15;
                if (a1 > lim) a1 = lim;
                nb[i] = (uint32_t) (a1 - a0[i]);
                my_bytes += nb[i];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_bytes += __shfl_xor_sync(0xffffffffu, my_bytes, o);
            if (lane == 0) mbar_arrive_expect_tx(&ctl->raw_full[rs], my_bytes);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int r = lane + 32 * i;
                bulk_g2s_plain(raw + (size_t) (rs * MMQ_BM + r) * P.slot, P.W + a0[i], nb[i], &ctl->raw_full[rs]);
            }
            // the 4 activation chunks of this super-block
            if (lane == 0) {
                for (int c = 0; c < 4; c++) {
                    const int u = sb * 4 + c, s = u % MMQ_NSTAGE, round = u / MMQ_NSTAGE;
                    if (round > 0 && !mmq_wait(ctl, &ctl->stage_free[s], (round - 1) & 1)) { ok = false; break; }
                    mbar_arrive_expect_tx(&ctl->b_full[s], (uint32_t) b_bytes);
                    bulk_g2s_plain(b_st + (size_t) s * b_bytes, Bt + (size_t) u * b_bytes, (uint32_t) b_bytes, &ctl->b_full[s]);
                }
            }
            ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
            if (!ok) break;
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            for (int u = 0; u < nchunk; u++) {
                const int s = u % MMQ_NSTAGE, round = u / MMQ_NSTAGE;
                if (!mmq_wait(ctl, &ctl->a_ready[s], round & 1)) break;
                if (!mmq_wait(ctl, &ctl->b_full[s], round & 1)) break;
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(smem_u32(a_st + (size_t) s * MMQ_A_BYTES));
                const uint64_t db = umma_desc_sw128(smem_u32(b_st + (size_t) s * b_bytes));
#pragma unroll
                for (int k = 0; k < MMQ_BK / 16; k++) umma_f16(tmem, da + 2 * k, db + 2 * k, P.idesc, (u | k) != 0);
                umma_commit(&ctl->stage_free[s]);
            }
            umma_commit(&ctl->acc_ready);
        }
        __syncwarp();
    } else {
        // ================= weight expansion =================
        const int dt = threadIdx.x - 64;        // 0..255
        const int r = dt >> 1, h = dt & 1;
        const int gr = min(row0 + r, P.N - 1);
        bool ok = true;
        for (int sb = 0; sb < nsb && ok; sb++) {
            const int rs = sb & 1, rr = sb >> 1;
            if (!mmq_wait(ctl, &ctl->raw_full[rs], rr & 1)) { ok = false; break; }
            const int64_t g0 = (int64_t) gr * P.row_bytes + (int64_t) sb * P.bpb;
            const uint8_t * blk = raw + (size_t) (rs * MMQ_BM + r) * P.slot + (g0 & 15);
            for (int c = 0; c < 4; c++) {
                const int u = sb * 4 + c, s = u % MMQ_NSTAGE, round = u / MMQ_NSTAGE;
                uint32_t v[16];
                if (P.type == T_Q4_K) expand_q4K(blk, c, h, v);
                else if (P.type == T_Q5_K) expand_q5K(blk, c, h, v);
                else expand_q6K(blk, c, h, v);
                if (round > 0 && !mmq_wait(ctl, &ctl->stage_free[s], (round - 1) & 1)) { ok = false; break; }
                uint8_t * arow = a_st + (size_t) s * MMQ_A_BYTES + r * 128;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int chunk = (4 * h + j) ^ (r & 7);
                    *reinterpret_cast<uint4 *>(arow + chunk * 16) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->a_ready[s]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->raw_empty[rs]);
        }
        // ================= epilogue: warps 2..5 cover the four 32-lane quadrants of TMEM =================
        if (warp < 6) {
            const bool acc_ok = mmq_wait(ctl, &ctl->acc_ready, 0);
            tc_fence_after();
            const int quad = warp & 3;
            const int n = row0 + quad * 32 + lane;
            const float bias = (P.bias && n < P.N) ? P.bias[n] : 0.f;
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem + ((uint32_t) (quad * 32) << 16) + (uint32_t) c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (acc_ok && n < P.N) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int t = ttile * BN + c0 + i;
                        if (t < P.T) P.dst[(size_t) t * P.N + n] = __fadd_rn(__uint_as_float(v[i]), bias);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(P.tmem_cols) : "memory");
    }
}

// ---- activation rows -> q8_K (exactly as the CPU backend quantizes them) -> fp16, written in the tiled UMMA image ----
__global__ void __launch_bounds__(256) k_mmq_prep(const float * __restrict__ x, int64_t ldx, int T, int K, int BN, uint8_t * __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = K / 256;
    const int t = blockIdx.x;                       // 0 .. Tpad-1
    const int b_bytes = BN * 128;
    for (int b = warp; b < nblk; b += 8) {
        float v[8];
        if (t < T) {
            const float4 * p = reinterpret_cast<const float4 *>(x + (size_t) t * ldx + (size_t) b * 256 + lane * 8);
            const float4 a = p[0], c = p[1];
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = 0.f;
        }
        float amax = 0.f, vmax = 0.f;
        int idx = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float ax = fabsf(v[i]);
            if (ax > amax) { amax = ax; vmax = v[i]; idx = lane * 8 + i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float oa = __shfl_xor_sync(0xffffffffu, amax, o);
            const float ov = __shfl_xor_sync(0xffffffffu, vmax, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (oa > amax || (oa == amax && oi < idx)) { amax = oa; vmax = ov; idx = oi; }
        }
        uint32_t h[4] = {0u, 0u, 0u, 0u};
        if (amax != 0.f) {
            const float iscale = __fdiv_rn(-127.f, vmax);
            const float d = __fdiv_rn(1.f, iscale);
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                int q = nearest_int_magic(__fmul_rn(iscale, v[i]));
                q = q < 127 ? q : 127;
                f[i] = __fmul_rn(d, (float) q);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) h[i] = pack_h2(f[2 * i], f[2 * i + 1]);
        }
        const int k = b * 256 + lane * 8;
        const int kc = k >> 6, j = (k & 63) >> 3;
        const int tt = t / BN, tl = t % BN;
        uint8_t * dstp = out + ((size_t) tt * (K / 64) + kc) * b_bytes + (size_t) tl * 128 + (size_t) ((j ^ (tl & 7)) << 4);
        *reinterpret_cast<uint4 *>(dstp) = make_uint4(h[0], h[1], h[2], h[3]);
    }
}

static int mmq_pick_bn(int T) {
    if (T >= 256) return 256;
    int bn = (T + 15) / 16 * 16;
    return bn < 16 ? 16 : bn;
}
size_t mmq_workspace_bytes(int64_t K, int64_t T) {
    const int BN = mmq_pick_bn((int) T);
    const int64_t tpad = (T + BN - 1) / BN * BN;
    return (size_t) (tpad * K * 2);
}
bool mmq_supported(int type, int64_t K) { return is_kquant(type) && K % 256 == 0 && K >= 256; }

int mmq_aborted() {
    int v = 0;
    cudaMemcpyFromSymbol(&v, g_mmq_abort, sizeof(int));
    return v;
}

cudaError_t launch_mmq(int type, const void * W, int64_t N, int64_t K, const float * x, int64_t ldx, int64_t T, float * dst, const float * bias,
                       void * ws, cudaStream_t st) {
    if (!mmq_supported(type, K) || N <= 0 || T <= 0) return cudaErrorInvalidValue;
    const int BN = mmq_pick_bn((int) T);
    const int tpad = (int) ((T + BN - 1) / BN * BN);
    k_mmq_prep<<<tpad, 256, 0, st>>>(x, ldx, (int) T, (int) K, BN, (uint8_t *) ws);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;

    MmqParams P{};
    P.W = (const uint8_t *) W;
    P.B = (const uint8_t *) ws;
    P.dst = dst;
    P.bias = bias;
    P.row_bytes = row_bytes(type, K);
    P.total_bytes = P.row_bytes * N;
    P.type = type;
    P.N = (int) N;
    P.K = (int) K;
    P.T = (int) T;
    P.BN = BN;
    P.bpb = type == T_Q4_K ? BYTES_Q4_K : (type == T_Q5_K ? BYTES_Q5_K : BYTES_Q6_K);
    P.slot = type == T_Q6_K ? 240 : P.bpb;
    uint32_t cols = 32;
    while ((int) cols < BN) cols <<= 1;
    P.tmem_cols = cols;
    // kind::f16 instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D=f32, A=B=f16, both K-major, N>>3, M>>4
    P.idesc = (1u << 4) | ((uint32_t) (BN >> 3) << 17) | ((uint32_t) (MMQ_BM >> 4) << 24);
    const size_t smem = 1024 + (size_t) MMQ_NSTAGE * (MMQ_A_BYTES + BN * 128) + 2 * (size_t) MMQ_BM * P.slot + MMQ_CTL_BYTES;
    static size_t configured = 0;
    if (smem > configured) {
        e = cudaFuncSetAttribute(k_mmq_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    dim3 grid((unsigned) ((N + MMQ_BM - 1) / MMQ_BM), (unsigned) (tpad / BN));
    k_mmq_tc<<<grid, MMQ_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace pb
