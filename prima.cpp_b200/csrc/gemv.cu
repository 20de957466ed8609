// prima.cpp_b200/csrc/gemv.cu — kernels + launchers for the decode GEMV (see gemv.cuh for the design).
#include "gemv.cuh"
#include "launch.h"
#include "quantize.cuh"
#include "rope.cuh"

#include <cstdlib>
#include <vector>

namespace pb {

// optional per-CTA timeline (debugging / profiles): 8 x %globaltimer stamps per CTA, enabled with pb200_debug_set_trace
__device__ unsigned long long * g_gemv_trace = nullptr;
__device__ __forceinline__ void trace(int k) {
    if (g_gemv_trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_gemv_trace[blockIdx.x * 8 + k] = t;
    }
}
int gemv_hang_info(unsigned long long * out32) { return (int) cudaMemcpyFromSymbol(out32, g_hang_info, 256); }
int gemv_set_trace(unsigned long long * dev_buf) { return (int) cudaMemcpyToSymbol(g_gemv_trace, &dev_buf, sizeof(dev_buf)); }

struct __align__(16) GemvSmemCtl {
    uint64_t full[GEMV_NSTAGE];
    int cnt[GEMV_NSTAGE];                 // consumer warps done with the stage; the last one refills it
    int pad_[GEMV_NSTAGE];
    uint64_t pbar[GEMV_NSTAGE][4];        // wpr > 1: "partials of this stage's row are in shared memory" per warp group
    float part[GEMV_NSTAGE][GEMV_TEAM_W]; // cross-warp partial sums, one slot per stage in flight
    double red[GEMV_NW];                  // rms_norm partial sums of squares
    int dbg[24];                          // [0..2] copies of cnt, [3..18] iteration each warp is in, [19..21] last refill iteration issued per stage
    volatile int issued[4];               // highest iteration whose tile has been REQUESTED for the stage (-1: none)
};
constexpr int GEMV_CTL_BYTES = 768;

__device__ __forceinline__ void consumer_bar() {   // the 8 consumer warps only (the producer warp never joins)
    asm volatile("bar.sync 9, %0;" ::"n"(GEMV_NW * 32) : "memory");
}
__device__ __forceinline__ float silu_f(float x) { return __fdiv_rn(x, 1.0f + expf(-x)); }   // ggml.c:2560

// Fused prologue executed by the 8 consumer warps: quantize the activation into shared memory (see PRO_* in gemv.cuh).
// Warp w owns super-blocks w, w+8, w+16, ... ; lane l owns elements 8l..8l+7 of each.  Blocks are processed four at a
// time with every global load issued up front, and for PRO_RMSNORM the same registers feed the sum of squares, so the
// vector is read exactly once (K <= 8192 in one batch; longer vectors loop over batches for the sum, then again to quantize).
__device__ __forceinline__ void load8(const float * p, float (&v)[8]) {   // .cg: L2 only (data written by other CTAs of a persistent grid)
    const float4 a0 = __ldcg(reinterpret_cast<const float4 *>(p)), a1 = __ldcg(reinterpret_cast<const float4 *>(p + 4));
    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
}
// Phase A issues the global loads of the first batch (nothing else), phase B does the arithmetic.  The ring fill is
// issued BETWEEN the two: requested first, the few KB of activation are not queued behind ~28 MB of weight prefetch
// (measured: that queueing cost every GEMV launch 4-5 us, profiles/r1_launches.md).
struct ProRegs { float x[2][8], w[2][8]; };
__device__ __forceinline__ void prologue_load(const GemvParams & P, ProRegs & R, int warp, int lane, int b0) {
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int b = b0 + warp + j * GEMV_NW;
        if (b < P.nblk) {
            load8(P.in0 + b * 256 + lane * 8, R.x[j]);
            if (P.prologue != PRO_QUANT) load8(P.in1 + b * 256 + lane * 8, R.w[j]);
        }
    }
}
__device__ __forceinline__ void prologue_compute(const GemvParams & P, GemvSmemCtl * ctl, const ActQ & sa, ProRegs & R, int warp, int lane) {
    const int tid = warp * 32 + lane;
    constexpr int B = 2;   // blocks in flight per warp (16 warps x 2 = one batch for K = 8192)
    const bool single_batch = P.nblk <= GEMV_NW * B;
    float scale = 1.f;
    if (P.prologue == PRO_RMSNORM) {
        double sum = 0.0;
        if (single_batch) {
#pragma unroll
            for (int j = 0; j < B; j++) {
                const int b = warp + j * GEMV_NW;
                if (b < P.nblk) {
#pragma unroll
                    for (int i = 0; i < 8; i++) sum += (double) __fmul_rn(R.x[j][i], R.x[j][i]);
                }
            }
        } else {
            for (int i = tid; i < P.K; i += GEMV_NW * 32) { const float v = __ldcg(P.in0 + i); sum += (double) __fmul_rn(v, v); }
        }
        sum = warp_sum_d(sum);
        if (lane == 0) ctl->red[warp] = sum;
        consumer_bar();
        if (g_gemv_trace && threadIdx.x == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); g_gemv_trace[148 * 32 + blockIdx.x] = tt; }
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < GEMV_NW; i++) t += ctl->red[i];     // every thread: same order, same result
        const float mean = (float) (t / (double) P.K);
        scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, P.eps)));
    }
    for (int b0 = 0; b0 < P.nblk; b0 += GEMV_NW * B) {
        if (b0 > 0) prologue_load(P, R, warp, lane, b0);
#pragma unroll
        for (int j = 0; j < B; j++) {
            const int b = b0 + warp + j * GEMV_NW;
            if (b < P.nblk) {
                if (P.prologue == PRO_RMSNORM) {
#pragma unroll
                    for (int i = 0; i < 8; i++) R.x[j][i] = __fmul_rn(__fmul_rn(R.x[j][i], scale), R.w[j][i]);
                } else if (P.prologue == PRO_SILU_MUL) {
#pragma unroll
                    for (int i = 0; i < 8; i++) R.x[j][i] = __fmul_rn(silu_f(R.x[j][i]), R.w[j][i]);
                }
                quantize_warp_q8K(R.x[j], lane, b, sa);
            }
        }
    }
    consumer_bar();
}

__device__ __forceinline__ void tile_info(const GemvParams & P, int t, int & m, int & r0, int & nrows) {
    m = 0;
#pragma unroll
    for (int i = 1; i < GEMV_MAX_MAT; i++)
        if (i < P.nmat && t >= P.mat[i].tile0) m = i;
    const GemvMat & M = P.mat[m];
    r0 = (t - M.tile0) * M.rows_per_tile;
    nrows = min(M.rows_per_tile, M.N - r0);
}

// one bulk copy (TMA 1-D) of tile t into ring stage s; arms the stage's mbarrier with the byte count first
__device__ __forceinline__ void issue_tile(const GemvParams & P, GemvSmemCtl * ctl, uint8_t * stages, int s, int t, uint64_t pol) {
    int m, r0, nrows;
    tile_info(P, t, m, r0, nrows);
    const GemvMat & M = P.mat[m];
    const int64_t g0 = (int64_t) r0 * M.row_bytes;
    const int64_t g1 = g0 + (int64_t) nrows * M.row_bytes;
    const int64_t a0 = g0 & ~(int64_t) 15;
    int64_t a1 = (g1 + 15) & ~(int64_t) 15;
    const int64_t lim = (M.total_bytes + 15) & ~(int64_t) 15;   // allocations are 16-B granular
    if (a1 > lim) a1 = lim;
    const uint32_t bytes = (uint32_t) (a1 - a0);
    mbar_arrive_expect_tx(&ctl->full[s], bytes);
    bulk_g2s(stages + (size_t) s * GEMV_STAGE_BYTES, M.W + a0, bytes, &ctl->full[s], pol);
}
// called by lane 0 of a consumer warp when the warp no longer needs stage s (iteration it)
__device__ __forceinline__ void release_stage(const GemvParams & P, GemvSmemCtl * ctl, uint8_t * stages, int s, int it, uint64_t pol) {
    __threadfence_block();
    const int old = atomicAdd(&ctl->cnt[s], 1);
    ctl->dbg[s] = old + 1;
    if (old == GEMV_TEAM_W - 1) {
        ctl->cnt[s] = 0;
        ctl->dbg[19 + s] = it + GEMV_NSTAGE;
        const int t = blockIdx.x + (it + GEMV_NSTAGE) * gridDim.x;
        if (t < P.ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the stage before the async-proxy refill
            issue_tile(P, ctl, stages, s, t, pol);
            __threadfence_block();
            ctl->issued[s] = it + GEMV_NSTAGE;
        } else if (P.next_W) {
            // nothing left to stream for this stage: keep the memory pipe busy with the next launch's first tiles
            const int n_my = (P.ntiles - (int) blockIdx.x + (int) gridDim.x - 1) / (int) gridDim.x;
            const int j = it + GEMV_NSTAGE - n_my;                     // 0 .. NSTAGE-1
            const int64_t off = ((int64_t) blockIdx.x + (int64_t) j * gridDim.x) * P.next_tile_bytes;
            if (j >= 0 && j < GEMV_NSTAGE && off < P.next_total_bytes) {
                int64_t a0 = off & ~(int64_t) 15;
                int64_t a1 = (off + P.next_tile_bytes + 15) & ~(int64_t) 15;
                const int64_t lim = P.next_total_bytes & ~(int64_t) 15;
                if (a1 > lim) a1 = lim;
                if (a1 > a0) bulk_prefetch_l2(P.next_W + a0, (uint32_t) (a1 - a0));
            }
        }
    }
}

__global__ void __launch_bounds__(GEMV_THREADS, 1) k_gemv_kquant(const __grid_constant__ GemvParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    GemvSmemCtl * ctl = reinterpret_cast<GemvSmemCtl *>(smem);
    uint8_t * stages = smem + GEMV_CTL_BYTES;
    uint8_t * act_smem = stages + (size_t) GEMV_NSTAGE * GEMV_STAGE_BYTES;
    static_assert(sizeof(GemvSmemCtl) <= GEMV_CTL_BYTES, "ctl block");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    trace(0);
    const uint64_t pol = policy_evict_first();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < GEMV_NSTAGE; s++) {
            mbar_init(&ctl->full[s], 1);
            ctl->cnt[s] = 0;
            ctl->issued[s] = -1;
#pragma unroll
            for (int g = 0; g < 4; g++) mbar_init(&ctl->pbar[s][g], P.wpr > 1 ? P.wpr : 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // ===== consumers (all 16 warps) =====
    const int team = warp / GEMV_TEAM_W, tw = warp % GEMV_TEAM_W;
    const int wpr = P.wpr;
    const int ngroups = GEMV_TEAM_W / wpr;
    const int group = tw / wpr, wsub = tw % wpr;
    const int blk = wsub * 32 + lane;
    const bool valid = blk < P.nblk;

    pdl_trigger();   // let the next kernel become resident as SMs drain; its own pdl_wait() orders the data
    if (P.fill_before_wait && threadIdx.x == 0) {
        // this launch follows a small kernel and is already resident while it runs: stream weights now
#pragma unroll
        for (int it = 0; it < GEMV_NSTAGE; it++) {
            const int t = blockIdx.x + it * gridDim.x;
            if (t < P.ntiles) { issue_tile(P, ctl, stages, it, t, pol); ctl->issued[it] = it; }
        }
    }
    pdl_wait();      // the activation is produced by the previous kernel in the stream
    trace(1);
    ActRegs r;
    ProRegs pr;
    ActQ sa;   // the CTA's activation in shared memory: qs[K] | bsums[K/16] i16 | d[K/256] f32
    sa.qs = reinterpret_cast<int8_t *>(act_smem);
    sa.bsums = reinterpret_cast<int16_t *>(act_smem + P.nblk * ACT_SMEM_QS_STRIDE);
    sa.d = reinterpret_cast<float *>(act_smem + P.nblk * (ACT_SMEM_QS_STRIDE + 2 * ACT_SMEM_BS_STRIDE));
    sa.s = nullptr;
    sa.qs_stride = ACT_SMEM_QS_STRIDE;
    sa.bs_stride = ACT_SMEM_BS_STRIDE;
    // 1) request the (small) activation first: ONE coalesced copy per CTA (every warp fetching its own registers from
    //    global memory moved 16x the bytes through L2 and cost ~4 us per launch, profiles/r1_gemv_timeline.md) ...
    int4 cq[4], cb;
    float cd = 0.f;
    const int nq = P.K / 16, nb16 = P.K / 128;   // int4 counts of qs and bsums
    if (P.prologue == PRO_NONE) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = threadIdx.x + j * GEMV_THREADS;
            if (i < nq) cq[j] = reinterpret_cast<const int4 *>(P.act.qs)[i];
        }
        if ((int) threadIdx.x < nb16) cb = reinterpret_cast<const int4 *>(P.act.bsums)[threadIdx.x];
        if ((int) threadIdx.x < P.nblk) cd = P.act.d[threadIdx.x];
    } else {
        prologue_load(P, pr, warp, lane, 0);
    }
    __syncthreads();   // every warp has ISSUED its loads (not waited for them)
    trace(2);
    // 2) ... then start the weight stream: fill the whole ring
    if (threadIdx.x == 0 && !P.fill_before_wait) {
#pragma unroll
        for (int it = 0; it < GEMV_NSTAGE; it++) {
            const int t = blockIdx.x + it * gridDim.x;
            if (t < P.ntiles) { issue_tile(P, ctl, stages, it, t, pol); ctl->issued[it] = it; }
        }
    }
    trace(3);
    // 3) stage / quantize the activation in shared memory while the first tiles are in flight
    if (P.prologue == PRO_NONE) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = threadIdx.x + j * GEMV_THREADS;
            if (i < nq) *reinterpret_cast<int4 *>(sa.qs + (i >> 4) * ACT_SMEM_QS_STRIDE + (i & 15) * 16) = cq[j];
        }
        if ((int) threadIdx.x < nb16) *reinterpret_cast<int4 *>(reinterpret_cast<char *>(sa.bsums) + (threadIdx.x >> 1) * (2 * ACT_SMEM_BS_STRIDE) + (threadIdx.x & 1) * 16) = cb;
        if ((int) threadIdx.x < P.nblk) sa.d[threadIdx.x] = cd;
        consumer_bar();
    } else {
        prologue_compute(P, ctl, sa, pr, warp, lane);
    }
    load_act_regs(r, sa, blk, valid);
    finish_act_regs(r);
    trace(4);

    for (int it = team, t = blockIdx.x + team * gridDim.x; t < P.ntiles; t += GEMV_NTEAM * gridDim.x, it += GEMV_NTEAM) {
        const int s = it % GEMV_NSTAGE;
        const uint32_t ph = (it / GEMV_NSTAGE) & 1;
        int m, r0, nrows;
        tile_info(P, t, m, r0, nrows);
        const GemvMat & M = P.mat[m];
        const int type = M.type;
        const int bpb = type == T_Q4_K ? BYTES_Q4_K : (type == T_Q5_K ? BYTES_Q5_K : BYTES_Q6_K);
        const uint32_t mis = (uint32_t) (((int64_t) r0 * M.row_bytes) & 15);
        const uint8_t * tile = stages + (size_t) s * GEMV_STAGE_BYTES + mis;
        if (lane == 0) ctl->dbg[3 + warp] = it;
        // With an odd ring depth the previous use of this stage belongs to the OTHER team: a parity wait alone could be
        // satisfied by the phase before it (ABA).  First make sure this iteration's tile has been requested at all.
        {
            const long long w0 = clock64();
            while (ctl->issued[s] < it) { if (clock64() - w0 > (1ll << 29)) break; }
        }
        mbar_wait(&ctl->full[s], ph, it, ctl->dbg);
        if (it == 0) trace(5);
        if (wpr == 1) {
            for (int slot = group; slot < nrows; slot += ngroups) {
                const int row = r0 + slot;
                // epilogue operands are requested before the dot so that their L2 latency is off the critical path
                float extra = 0.f;
                if (lane == 0) {
                    if (M.bias) extra = M.bias[row];
                    if (M.resid) extra += M.resid[row];
                }
                const uint8_t * bp = tile + (size_t) slot * M.row_bytes + (size_t) blk * bpb;
                float v = 0.f;
                if (valid) {
                    if (type == T_Q4_K) v = dot_q4K(bp, r);
                    else if (type == T_Q6_K) v = dot_q6K(bp, r);
                    else v = dot_q5K(bp, r);
                }
                if (slot + ngroups >= nrows) {
                    // last row of this stage for this warp: hand the buffer back to the producer before reducing
                    __syncwarp();
                    if (lane == 0) release_stage(P, ctl, stages, s, it, pol);
                }
                v = warp_sum(v);
                if (lane == 0) M.y[row] = v + extra;
            }
            if (group >= nrows) {   // this warp had no row in the tile (ragged last tile): still release the stage
                __syncwarp();
                if (lane == 0) release_stage(P, ctl, stages, s, it, pol);
            }
        } else {
            // rows split over wpr warps; rows_per_tile == ngroups, i.e. at most one row per warp group and stage.
            // Non-leader warps publish their partial and move on; only the group's leader warp waits for them
            // (mbarrier instead of bar.sync), and it releases the stage last so part[s] cannot be overwritten early.
            const int slot = group;
            const bool has_row = slot < nrows;
            const int row = r0 + slot;
            const bool lead = wsub == 0;
            float extra = 0.f;
            if (has_row && lead && lane == 0) {
                if (M.bias) extra = M.bias[row];
                if (M.resid) extra += M.resid[row];
            }
            float v = 0.f;
            if (has_row && valid) {
                const uint8_t * bp = tile + (size_t) slot * M.row_bytes + (size_t) blk * bpb;
                if (type == T_Q4_K) v = dot_q4K(bp, r);
                else if (type == T_Q6_K) v = dot_q6K(bp, r);
                else v = dot_q5K(bp, r);
            }
            if (!lead) {
                __syncwarp();
                if (lane == 0) release_stage(P, ctl, stages, s, it, pol);
                v = warp_sum(v);
                if (lane == 0) {
                    ctl->part[s][tw] = v;
                    mbar_arrive(&ctl->pbar[s][group]);   // release semantics: the partial is visible to the waiter
                }
            } else {
                v = warp_sum(v);
                uint64_t tok = 0;
                if (lane == 0) tok = mbar_arrive_token(&ctl->pbar[s][group]);
                tok = __shfl_sync(0xffffffffu, tok, 0);
                mbar_wait_token(&ctl->pbar[s][group], tok, it);
                if (lane == 0) {
                    float acc = v;
                    for (int i = 1; i < wpr; i++) acc += ctl->part[s][group * wpr + i];
                    if (has_row) M.y[row] = acc + extra;
                    release_stage(P, ctl, stages, s, it, pol);
                }
            }
        }
    }
    trace(6);
}


// =================================================================================================================
// Persistent token kernel: ALL GEMV phases of a decode step (4 per layer + lm_head) in ONE cooperative launch.
// The weight ring is never drained: a stage freed in phase g is refilled with the CTA's next tile in global order, which may
// belong to phase g+1 (or the next layer), so while the consumers sit in a grid barrier / fused prologue / the attention
// phase, up to 4 stages (192 KB per SM, 28 MB chip-wide = 3.8 us of HBM time) of the next phase are already landing.
// This removes the per-launch ramp measured in profiles/r1_gemv_timeline.txt (~9 us x 4 launches per layer).
// Phases per layer: qkv[rmsnorm fused] | barrier | rope+kv-store+attention (CTA h < n_head) | barrier | wo[quant fused] |
// barrier | gate,up[rmsnorm fused] | barrier | silu*up -> q8_K (CTA b < F/256) | barrier | down | barrier.
// =================================================================================================================
struct MkPhase {
    GemvMat mat[GEMV_MAX_MAT];
    int nmat, ntiles, K, nblk, wpr, prologue;
    const float * in0;
    const float * in1;
    float eps;
    ActQ act;   // PRO_NONE source (ffn_down: written by the silu phase)
};
struct MkLayer {
    MkPhase ph[4];                                              // qkv, wo, gate|up, down
    const float * q; const float * k; const float * v;          // attention inputs (pre-RoPE)
    __half * kc; __half * vc; float * att;
    const float * g; const float * u; ActQ actF; int F;          // silu*up -> q8_K
};
struct MkParams {
    const MkLayer * layers;
    int n_layers;
    MkPhase head;
    int with_head;
    const int32_t * pos_dev;
    RopeParams rp;
    const float * freq_factors;
    float kq_scale;
    int n_head, n_head_kv, n_ctx;
    unsigned int * barrier;   // zeroed before every launch
    int * error_flag;
    int xb_allowed;           // stages prefetched across a phase boundary before the next prologue's loads are out
};
constexpr int MK_MAX_PHASES = 4 * 160 + 1;

struct __align__(16) MkSmem {
    GemvSmemCtl ctl;
    uint16_t cnt[MK_MAX_PHASES + 7];   // tiles of this CTA per phase
    MkPhase desc[4];                   // descriptors of phases g .. g+2 (slot = phase % 4): refills never read global memory
    int xb_count;                      // refills issued across the coming phase boundary
    int deferred[GEMV_NSTAGE];         // global iterations whose refill waits until the next prologue's loads are out
};
constexpr int MK_XB_ALLOWED = GEMV_NSTAGE;   // measured: deferring buys nothing (the prologue was slow for another reason), keep the ring primed       // stages prefetched across a phase boundary before the prologue (one per team)
constexpr int MK_HDR_BYTES = 3328;
static_assert(sizeof(MkSmem) <= MK_HDR_BYTES, "MkSmem header");

__device__ __forceinline__ const MkPhase * mk_phase(const MkParams & P, int g) {
    return g < 4 * P.n_layers ? &P.layers[g >> 2].ph[g & 3] : &P.head;
}
template <class D>
__device__ __forceinline__ void tile_info_t(const D & P, int t, int & m, int & r0, int & nrows) {
    m = 0;
#pragma unroll
    for (int i = 1; i < GEMV_MAX_MAT; i++)
        if (i < P.nmat && t >= P.mat[i].tile0) m = i;
    const GemvMat & M = P.mat[m];
    r0 = (t - M.tile0) * M.rows_per_tile;
    nrows = min(M.rows_per_tile, M.N - r0);
}
__device__ __forceinline__ void mk_issue(const MkPhase * ph, GemvSmemCtl * ctl, uint8_t * stages, int s, int t, uint64_t pol) {
    int m, r0, nrows;
    tile_info_t(*ph, t, m, r0, nrows);
    const GemvMat & M = ph->mat[m];
    const int64_t g0 = (int64_t) r0 * M.row_bytes;
    const int64_t g1 = g0 + (int64_t) nrows * M.row_bytes;
    const int64_t a0 = g0 & ~(int64_t) 15;
    int64_t a1 = (g1 + 15) & ~(int64_t) 15;
    const int64_t lim = (M.total_bytes + 15) & ~(int64_t) 15;
    if (a1 > lim) a1 = lim;
    const uint32_t bytes = (uint32_t) (a1 - a0);
    mbar_arrive_expect_tx(&ctl->full[s], bytes);
    bulk_g2s(stages + (size_t) s * GEMV_STAGE_BYTES, M.W + a0, bytes, &ctl->full[s], pol);
}
// issue global iteration G of this CTA (searching forward from phase g whose first iteration is `base`)
__device__ __forceinline__ void mk_issue_iter(const MkParams & P, MkSmem * sm, uint8_t * stages, int n_phases, int g, int base, int G, uint64_t pol) {
    const int g0 = g;
    while (g < n_phases && G >= base + (int) sm->cnt[g]) { base += sm->cnt[g]; g++; }
    if (g >= n_phases) return;
    // phases g0 .. g0+2 are cached in shared memory; further look-ahead (only with very few tiles per phase) reads HBM
    const MkPhase * ph = g <= g0 + 2 ? &sm->desc[g & 3] : mk_phase(P, g);
    mk_issue(ph, &sm->ctl, stages, G % GEMV_NSTAGE, (int) blockIdx.x + (G - base) * (int) gridDim.x, pol);
    __threadfence_block();
    sm->ctl.issued[G % GEMV_NSTAGE] = G;
}
__device__ __forceinline__ void mk_load_desc(const MkParams & P, MkSmem * sm, int g, int n_phases) {
    if (g >= n_phases) return;
    const int * src = reinterpret_cast<const int *>(mk_phase(P, g));
    int * dst = reinterpret_cast<int *>(&sm->desc[g & 3]);
    for (int i = threadIdx.x; i < (int) (sizeof(MkPhase) / 4); i += GEMV_THREADS) dst[i] = src[i];
}
// A refill that belongs to a LATER phase is only issued right away for the first MK_XB_ALLOWED stages: everything an SM has
// in flight delays its own small dependent loads (measured: 192 KB of bulk copies outstanding per SM add ~4 us to the next
// prologue's 32-KB activation read, profiles/r1_persistent_timeline.txt), so the rest waits until those loads are out.
__device__ __forceinline__ void mk_release(const MkParams & P, MkSmem * sm, uint8_t * stages, int n_phases, int g, int base, int G, uint64_t pol) {
    __threadfence_block();
    const int s = G % GEMV_NSTAGE;
    if (atomicAdd(&sm->ctl.cnt[s], 1) == GEMV_TEAM_W - 1) {
        sm->ctl.cnt[s] = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int G2 = G + GEMV_NSTAGE;
        if (G2 >= base + (int) sm->cnt[g]) {                 // crosses into a later phase
            // the qkv -> wo boundary contains the attention phase, whose dependent K/V loads must not queue behind this SM's
            // own bulk refills (measured: attention 6 us alone, 15 us with 192 KB of refills in flight): defer all of them
            const int allowed = (g < 4 * P.n_layers && (g & 3) == 0) ? 0 : P.xb_allowed;
            const int k = atomicAdd(&sm->xb_count, 1);
            if (k >= allowed) { sm->deferred[k - allowed] = G2; return; }
        }
        mk_issue_iter(P, sm, stages, n_phases, g, base, G2, pol);
    }
}
// called by thread 0 of the next phase once its prologue loads have been issued
__device__ __forceinline__ void mk_issue_deferred(const MkParams & P, MkSmem * sm, uint8_t * stages, int n_phases, int g, int base, uint64_t pol) {
    const int allowed = (g > 0 && g <= 4 * P.n_layers && ((g - 1) & 3) == 0) ? 0 : P.xb_allowed;   // boundary we just crossed
    const int n = sm->xb_count - allowed;
    for (int i = 0; i < n && i < GEMV_NSTAGE; i++) mk_issue_iter(P, sm, stages, n_phases, g, base, sm->deferred[i], pol);
    sm->xb_count = 0;
}
__device__ __forceinline__ void mk_grid_barrier(const MkParams & P, unsigned & bar_idx) {
    __syncthreads();
    ++bar_idx;
    if (threadIdx.x == 0) {
        // release-add (no return value, no separate membar) then poll with acquire loads
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.barrier) : "memory");
        const unsigned target = bar_idx * gridDim.x;
        const long long t0 = clock64();
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.barrier) : "memory");
            if (clock64() - t0 > (1ll << 33)) { *P.error_flag = 1; __trap(); }   // ~4 s: never hang the GPU on a logic error
        } while (v < target);
    }
    __syncthreads();
}
// timeline of layer 1 (phases 4..8): stamp k of phase g -> slot (g-4)*3+k of this CTA's 16-entry trace row
__device__ __forceinline__ void mk_trace(int g, int k) {
    if (g_gemv_trace && threadIdx.x == 0 && g >= 4 && g < 9) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_gemv_trace[blockIdx.x * 16 + (g - 4) * 3 + k] = t;
    }
}
// finer stamps inside the prologue of phases 4 and 7 (qkv, down of layer 1): rows 148.. of the trace buffer
__device__ __forceinline__ void mk_trace2(int g, int k) {
    if (g_gemv_trace && threadIdx.x == 0 && (g == 4 || g == 7)) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_gemv_trace[148 * 16 + blockIdx.x * 16 + (g == 4 ? 0 : 8) + k] = t;
    }
}
__device__ __forceinline__ void mk_trace3(int slot) {
    if (g_gemv_trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_gemv_trace[148 * 16 + blockIdx.x * 16 + slot] = t;
    }
}
__device__ __forceinline__ void bar256() { asm volatile("bar.sync 10, 256;" ::: "memory"); }

// RoPE + KV store + attention for q head h by warps 0..7 (256 threads); same arithmetic as k_attn_fused (ops.cu)
__device__ void mk_attention(const MkParams & P, const MkLayer & L, int h, float * sm, int warp, int lane) {
    constexpr int D = 128;
    const int tid = warp * 32 + lane;
    const int pos = *P.pos_dev, n_kv = pos + 1;
    const int gqa = P.n_head / P.n_head_kv, hk = h / gqa;
    const int64_t EK = (int64_t) P.n_head_kv * D;
    float * S = sm;                                   // [n_kv padded to 32]
    float * red = sm + ((P.n_ctx + 31) & ~31);       // [8][128]
    float * q_s = red + 8 * 128;                      // [128]
    __half * k_s = reinterpret_cast<__half *>(q_s + D);
    __half * v_s = k_s + D;
    float * s_red = reinterpret_cast<float *>(v_s + D);   // [8] floats, then [8] doubles, then 2 floats
    double * s_redd = reinterpret_cast<double *>(s_red + 8);
    float * s_bc = reinterpret_cast<float *>(s_redd + 8);
    const RopeParams & rp = P.rp;
    {
        const int half_dims = rp.n_dims / 2;
        const bool neox = rp.mode & 2;
        if (tid < 128) {
            const int pair = tid & 63;
            const bool is_q = tid < 64;
            const float * src = is_q ? L.q + (int64_t) h * D : L.k + (int64_t) hk * D;
            if (pair < half_dims) {
                float c, s;
                rope_cos_sin(rp, pos, pair, P.freq_factors, c, s);
                const int i0 = neox ? pair : 2 * pair, i1 = neox ? pair + half_dims : 2 * pair + 1;
                float y0, y1;
                rope_rotate(__ldcg(src + i0), __ldcg(src + i1), c, s, y0, y1);
                if (is_q) { q_s[i0] = __half2float(__float2half_rn(y0)); q_s[i1] = __half2float(__float2half_rn(y1)); }
                else { k_s[i0] = __float2half_rn(y0); k_s[i1] = __float2half_rn(y1); }
            }
            for (int i = rp.n_dims + pair; i < D; i += 64) {
                if (is_q) q_s[i] = __half2float(__float2half_rn(__ldcg(src + i)));
                else k_s[i] = __float2half_rn(__ldcg(src + i));
            }
        } else {
            v_s[tid - 128] = __float2half_rn(__ldcg(L.v + (int64_t) hk * D + tid - 128));
        }
    }
    bar256();
    if (h % gqa == 0 && tid < 32) {
        *reinterpret_cast<uint2 *>(L.kc + (int64_t) pos * EK + (int64_t) hk * D + 4 * lane) = *reinterpret_cast<const uint2 *>(k_s + 4 * lane);
        *reinterpret_cast<uint2 *>(L.vc + (int64_t) pos * EK + (int64_t) hk * D + 4 * lane) = *reinterpret_cast<const uint2 *>(v_s + 4 * lane);
    }
    const float q0 = q_s[4 * lane], q1 = q_s[4 * lane + 1], q2 = q_s[4 * lane + 2], q3 = q_s[4 * lane + 3];
    for (int p0 = warp; p0 < n_kv; p0 += 32) {     // 4 positions per warp in flight: all K rows requested before any is used
        uint2 kraw[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int p = p0 + 8 * j;
            if (p < n_kv) {
                const __half * krow = p == pos ? k_s : L.kc + (int64_t) p * EK + (int64_t) hk * D;
                kraw[j] = *reinterpret_cast<const uint2 *>(krow + 4 * lane);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int p = p0 + 8 * j;
            if (p < n_kv) {
                const float2 k01 = __half22float2(*reinterpret_cast<const __half2 *>(&kraw[j].x));
                const float2 k23 = __half22float2(*reinterpret_cast<const __half2 *>(&kraw[j].y));
                float s = k01.x * q0;
                s = fmaf(k01.y, q1, s);
                s = fmaf(k23.x, q2, s);
                s = fmaf(k23.y, q3, s);
                s = warp_sum(s);
                if (lane == 0) S[p] = __fmul_rn(s, P.kq_scale);
            }
        }
    }
    bar256();
    float m = -INFINITY;
    for (int p = tid; p < n_kv; p += 256) m = fmaxf(m, S[p]);
    m = warp_max(m);
    if (lane == 0) s_red[warp] = m;
    bar256();
    if (tid == 0) {
        float t = s_red[0];
        for (int i = 1; i < 8; i++) t = fmaxf(t, s_red[i]);
        s_bc[0] = t;
    }
    bar256();
    const float mx = s_bc[0];
    double dsum = 0.0;
    for (int p = tid; p < n_kv; p += 256) {
        const float e = expf(__fsub_rn(S[p], mx));
        S[p] = e;
        dsum += (double) e;
    }
    dsum = warp_sum_d(dsum);
    if (lane == 0) s_redd[warp] = dsum;
    bar256();
    if (tid == 0) {
        double t = 0;
        for (int i = 0; i < 8; i++) t += s_redd[i];
        s_bc[1] = (float) (1.0 / t);
    }
    bar256();
    const float inv = s_bc[1];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int p0 = warp; p0 < n_kv; p0 += 32) {
        uint2 vraw[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int p = p0 + 8 * j;
            if (p < n_kv) {
                const __half * vrow = p == pos ? v_s : L.vc + (int64_t) p * EK + (int64_t) hk * D;
                vraw[j] = *reinterpret_cast<const uint2 *>(vrow + 4 * lane);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int p = p0 + 8 * j;
            if (p < n_kv) {
                const float w = __half2float(__float2half_rn(__fmul_rn(S[p], inv)));
                const float2 v01 = __half22float2(*reinterpret_cast<const __half2 *>(&vraw[j].x));
                const float2 v23 = __half22float2(*reinterpret_cast<const __half2 *>(&vraw[j].y));
                a0 = fmaf(v01.x, w, a0); a1 = fmaf(v01.y, w, a1); a2 = fmaf(v23.x, w, a2); a3 = fmaf(v23.y, w, a3);
            }
        }
    }
    *reinterpret_cast<float4 *>(red + warp * 128 + 4 * lane) = make_float4(a0, a1, a2, a3);
    bar256();
    if (tid < 128) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) t += red[i * 128 + tid];
        L.att[(int64_t) h * D + tid] = t;
    }
}

__global__ void __launch_bounds__(GEMV_THREADS, 1) k_token_persistent(const __grid_constant__ MkParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    MkSmem * sm = reinterpret_cast<MkSmem *>(smem);
    GemvSmemCtl * ctl = &sm->ctl;
    uint8_t * stages = smem + MK_HDR_BYTES;
    uint8_t * act_smem = stages + (size_t) GEMV_NSTAGE * GEMV_STAGE_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_phases = 4 * P.n_layers + (P.with_head ? 1 : 0);
    const uint64_t pol = policy_evict_first();

    // the wpr > 1 phases of a model all share one wpr (checked on the host): pbar counts are fixed at init
    int wpr_split = 1;
    for (int g = threadIdx.x; g < n_phases; g += GEMV_THREADS) {
        const MkPhase * ph = mk_phase(P, g);
        const int nt = ph->ntiles;
        sm->cnt[g] = (uint16_t) (nt > (int) blockIdx.x ? (nt - (int) blockIdx.x + (int) gridDim.x - 1) / (int) gridDim.x : 0);
    }
    for (int g = 0; g < min(n_phases, 5); g++) wpr_split = max(wpr_split, mk_phase(P, g)->wpr);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < GEMV_NSTAGE; s++) {
            mbar_init(&ctl->full[s], 1);
            ctl->cnt[s] = 0;
            ctl->issued[s] = -1;
#pragma unroll
            for (int gI = 0; gI < 4; gI++) mbar_init(&ctl->pbar[s][gI], wpr_split > 1 ? wpr_split : 1);
        }
        mbar_fence_init();
    }
    if (threadIdx.x == 0) sm->xb_count = 0;
    mk_load_desc(P, sm, 0, n_phases);
    mk_load_desc(P, sm, 1, n_phases);
    mk_load_desc(P, sm, 2, n_phases);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int G = 0; G < GEMV_NSTAGE; G++) mk_issue_iter(P, sm, stages, n_phases, 0, 0, G, pol);   // prime the ring
    }

    const int team = warp / GEMV_TEAM_W, tw = warp % GEMV_TEAM_W;
    unsigned bar_idx = 0;
    int base = 0;                 // global iteration index of this CTA's first tile of phase g
    ActQ sa;
    ActRegs r;
    ProRegs pr;

    for (int g = 0; g < n_phases; g++) {
        const int li = g >> 2, pi = g < 4 * P.n_layers ? (g & 3) : 4;
        // ---------------- dependencies of this phase ----------------
        if (g == 5) mk_trace3(5);
        if (g > 0) mk_grid_barrier(P, bar_idx);                       // previous GEMV phase complete everywhere
        if (g == 5) mk_trace3(6);
        if (pi == 1) {                                                // wo needs the attention output
            const MkLayer & L = P.layers[li];
            if ((int) blockIdx.x < P.n_head && warp < 8) mk_attention(P, L, blockIdx.x, reinterpret_cast<float *>(act_smem), warp, lane);
            if (g == 5) mk_trace3(7);
            mk_grid_barrier(P, bar_idx);
            if (g == 5) mk_trace3(13);
        } else if (pi == 3) {                                         // ffn_down needs silu(g)*u quantized
            const MkLayer & L = P.layers[li];
            const int b = blockIdx.x;
            if (b < L.F / 256 && warp == 0) {
                float v[8], uu[8];
                load8(L.g + b * 256 + lane * 8, v);
                load8(L.u + b * 256 + lane * 8, uu);
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = __fmul_rn(silu_f(v[i]), uu[i]);
                quantize_warp_q8K(v, lane, b, L.actF);
            }
            mk_grid_barrier(P, bar_idx);
        }
        // ---------------- descriptors: slot g is resident since phase g-2; fetch g+2 into the slot phase g-2 used ----------------
        mk_load_desc(P, sm, g + 2, n_phases);   // nobody reads slot (g+2)&3 == (g-2)&3 any more (barrier above)
        __syncthreads();
        mk_trace(g, 0);
        const MkPhase & D = sm->desc[g & 3];
        const int wpr = D.wpr;
        const int ngroups = GEMV_TEAM_W / wpr;
        const int group = tw / wpr, wsub = tw % wpr;
        const int blk = wsub * 32 + lane;
        const bool valid = blk < D.nblk;
        // ---------------- activation: stage / quantize into shared memory, then into registers ----------------
        sa.qs = reinterpret_cast<int8_t *>(act_smem);
        sa.bsums = reinterpret_cast<int16_t *>(act_smem + D.nblk * ACT_SMEM_QS_STRIDE);
        sa.d = reinterpret_cast<float *>(act_smem + D.nblk * (ACT_SMEM_QS_STRIDE + 2 * ACT_SMEM_BS_STRIDE));
        sa.s = nullptr;
        sa.qs_stride = ACT_SMEM_QS_STRIDE;
        sa.bs_stride = ACT_SMEM_BS_STRIDE;
        if (D.prologue == PRO_NONE) {
            const int nq = D.K / 16, nb16 = D.K / 128;
            int4 cq[4], cb;
            float cd = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = threadIdx.x + j * GEMV_THREADS;
                if (i < nq) cq[j] = __ldcg(reinterpret_cast<const int4 *>(D.act.qs) + i);
            }
            if ((int) threadIdx.x < nb16) cb = __ldcg(reinterpret_cast<const int4 *>(D.act.bsums) + threadIdx.x);
            if ((int) threadIdx.x < D.nblk) cd = __ldcg(D.act.d + threadIdx.x);
            mk_trace2(g, 0);
            if (threadIdx.x == 0) mk_issue_deferred(P, sm, stages, n_phases, g, base, pol);
            mk_trace2(g, 1);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = threadIdx.x + j * GEMV_THREADS;
                if (i < nq) *reinterpret_cast<int4 *>(sa.qs + (i >> 4) * ACT_SMEM_QS_STRIDE + (i & 15) * 16) = cq[j];
            }
            if ((int) threadIdx.x < nb16) *reinterpret_cast<int4 *>(reinterpret_cast<char *>(sa.bsums) + (threadIdx.x >> 1) * (2 * ACT_SMEM_BS_STRIDE) + (threadIdx.x & 1) * 16) = cb;
            if ((int) threadIdx.x < D.nblk) sa.d[threadIdx.x] = cd;
            mk_trace2(g, 2);
            consumer_bar();
            mk_trace2(g, 3);
        } else {
            GemvParams Q;   // only the prologue fields are read
            Q.prologue = D.prologue; Q.in0 = D.in0; Q.in1 = D.in1; Q.eps = D.eps; Q.K = D.K; Q.nblk = D.nblk;
            mk_trace2(g, 0);
            prologue_load(Q, pr, warp, lane, 0);
            if (threadIdx.x == 0) mk_issue_deferred(P, sm, stages, n_phases, g, base, pol);
            mk_trace2(g, 1);
            prologue_compute(Q, ctl, sa, pr, warp, lane);
            mk_trace2(g, 3);
        }
        load_act_regs(r, sa, blk, valid);
        mk_trace2(g, 4);
        finish_act_regs(r);
        mk_trace(g, 1);
        // ---------------- consume this CTA's tiles of the phase ----------------
        const int n_g = sm->cnt[g];
        for (int i = ((team - base) & 1); i < n_g; i += GEMV_NTEAM) {
            const int G = base + i;
            const int s = G % GEMV_NSTAGE;
            const uint32_t ph = (G / GEMV_NSTAGE) & 1;
            const int t = (int) blockIdx.x + i * (int) gridDim.x;
            int m, r0, nrows;
            tile_info_t(D, t, m, r0, nrows);
            const GemvMat & M = D.mat[m];
            const int type = M.type;
            const int bpb = type == T_Q4_K ? BYTES_Q4_K : (type == T_Q5_K ? BYTES_Q5_K : BYTES_Q6_K);
            const uint32_t mis = (uint32_t) (((int64_t) r0 * M.row_bytes) & 15);
            const uint8_t * tile = stages + (size_t) s * GEMV_STAGE_BYTES + mis;
            {
                const long long w0 = clock64();
                while (ctl->issued[s] < G) { if (clock64() - w0 > (1ll << 29)) break; }
            }
            mbar_wait(&ctl->full[s], ph, G);
            if (wpr == 1) {
                for (int slot = group; slot < nrows; slot += ngroups) {
                    const int row = r0 + slot;
                    float extra = 0.f;
                    if (lane == 0) {
                        if (M.bias) extra = M.bias[row];
                        if (M.resid) extra += __ldcg(M.resid + row);
                    }
                    const uint8_t * bp = tile + (size_t) slot * M.row_bytes + (size_t) blk * bpb;
                    float v = 0.f;
                    if (valid) {
                        if (type == T_Q4_K) v = dot_q4K(bp, r);
                        else if (type == T_Q6_K) v = dot_q6K(bp, r);
                        else v = dot_q5K(bp, r);
                    }
                    if (slot + ngroups >= nrows) {
                        __syncwarp();
                        if (lane == 0) mk_release(P, sm, stages, n_phases, g, base, G, pol);
                    }
                    v = warp_sum(v);
                    if (lane == 0) M.y[row] = v + extra;
                }
                if (group >= nrows) {
                    __syncwarp();
                    if (lane == 0) mk_release(P, sm, stages, n_phases, g, base, G, pol);
                }
            } else {
                const int slot = group;
                const bool has_row = slot < nrows;
                const int row = r0 + slot;
                const bool lead = wsub == 0;
                float extra = 0.f;
                if (has_row && lead && lane == 0) {
                    if (M.bias) extra = M.bias[row];
                    if (M.resid) extra += __ldcg(M.resid + row);
                }
                float v = 0.f;
                if (has_row && valid) {
                    const uint8_t * bp = tile + (size_t) slot * M.row_bytes + (size_t) blk * bpb;
                    if (type == T_Q4_K) v = dot_q4K(bp, r);
                    else if (type == T_Q6_K) v = dot_q6K(bp, r);
                    else v = dot_q5K(bp, r);
                }
                if (!lead) {
                    __syncwarp();
                    if (lane == 0) mk_release(P, sm, stages, n_phases, g, base, G, pol);
                    v = warp_sum(v);
                    if (lane == 0) {
                        ctl->part[s][tw] = v;
                        mbar_arrive(&ctl->pbar[s][group]);
                    }
                } else {
                    v = warp_sum(v);
                    uint64_t tok = 0;
                    if (lane == 0) tok = mbar_arrive_token(&ctl->pbar[s][group]);
                    tok = __shfl_sync(0xffffffffu, tok, 0);
                    mbar_wait_token(&ctl->pbar[s][group], tok);
                    if (lane == 0) {
                        float acc = v;
                        for (int j = 1; j < wpr; j++) acc += ctl->part[s][group * wpr + j];
                        if (has_row) M.y[row] = acc + extra;
                        mk_release(P, sm, stages, n_phases, g, base, G, pol);
                    }
                }
            }
        }
        base += n_g;
        mk_trace(g, 2);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Generic fallback: one warp per row, direct global loads, every supported type (incl. the 32-element block types
// Q8_0 / Q5_1 that Qwen2.5-72B's ffn_down falls back to, src/llama.cpp:19516-19551), any K.
// Follows ggml_vec_dot_q8_0_q8_0 (ggml-quants.c:5518) and ggml_vec_dot_q5_1_q8_1 (:5144).
__device__ __forceinline__ uint32_t ld_u16x2(const uint8_t * p) {   // 2-B aligned 32-bit read
    const uint16_t * q = reinterpret_cast<const uint16_t *>(p);
    return (uint32_t) q[0] | ((uint32_t) q[1] << 16);
}

struct GemvGenericParams {
    const uint8_t * W;
    float * y;
    const float * bias;
    const float * resid;
    int64_t row_bytes;
    int type, N, K;
    ActQ act;
};

__global__ void __launch_bounds__(256) k_gemv_generic(const __grid_constant__ GemvGenericParams P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    pdl_trigger();   // dependents may launch now; they still wait for this grid's completion in their own pdl_wait()
    pdl_wait();
    if (row >= P.N) return;
    const uint8_t * wrow = P.W + (int64_t) row * P.row_bytes;
    float acc = 0.f;
    if (P.type == T_Q8_0) {
        const int nb = P.K / 32;
        for (int b = lane; b < nb; b += 32) {
            const uint8_t * bp = wrow + (int64_t) b * BYTES_Q8_0;
            const float d = __half2float(__ushort_as_half(*reinterpret_cast<const uint16_t *>(bp)));
            const int4 * a = reinterpret_cast<const int4 *>(P.act.qs + (int64_t) b * 32);
            const int4 a0 = a[0], a1 = a[1];
            const int av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            int sumi = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) sumi = dp4a_ss((int) ld_u16x2(bp + 2 + 4 * i), av[i], sumi);
            acc += (float) sumi * (d * P.act.d[b]);
        }
    } else if (P.type == T_Q5_1) {
        const int nb = P.K / 32;
        for (int b = lane; b < nb; b += 32) {
            const uint8_t * bp = wrow + (int64_t) b * BYTES_Q5_1;   // 24 B: 8-B aligned rows, 4-B aligned fields
            const uint32_t dmw = *reinterpret_cast<const uint32_t *>(bp);
            const float d = __half2float(__ushort_as_half((unsigned short) (dmw & 0xffff)));
            const float mm = __half2float(__ushort_as_half((unsigned short) (dmw >> 16)));
            const uint32_t qh = *reinterpret_cast<const uint32_t *>(bp + 4);
            const int4 * a = reinterpret_cast<const int4 *>(P.act.qs + (int64_t) b * 32);
            const int4 a0 = a[0], a1 = a[1];
            const int av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            int sumi = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(bp + 8 + 4 * i);   // qs bytes 4i..4i+3
                // element j = 4i+k (low nibble) gets bit j of qh; element j+16 (high nibble) gets bit j+16
                uint32_t hb_lo = 0, hb_hi = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    hb_lo |= ((qh >> (4 * i + k)) & 1u) << (8 * k + 4);
                    hb_hi |= ((qh >> (4 * i + k + 16)) & 1u) << (8 * k + 4);
                }
                sumi = dp4a_us((w & 0x0f0f0f0fu) | hb_lo, av[i], sumi);
                sumi = dp4a_us(((w >> 4) & 0x0f0f0f0fu) | hb_hi, av[4 + i], sumi);
            }
            acc += (d * P.act.d[b]) * (float) sumi + mm * P.act.s[b];
        }
    } else {
        // k-quants without shared-memory staging (used when K > 65 536 or for tiny problems)
        const int nb = P.K / 256;
        for (int b = lane; b < nb; b += 32) {
            ActRegs r;
            load_act_regs(r, P.act, b, true);
            finish_act_regs(r);
            // stage the block through registers -> local array is avoided by reading global memory directly with the
            // same dot routines: they only need byte-addressable memory with the block's natural alignment.
            const int bpb = P.type == T_Q4_K ? BYTES_Q4_K : (P.type == T_Q5_K ? BYTES_Q5_K : BYTES_Q6_K);
            const uint8_t * bp = wrow + (int64_t) b * bpb;
            // dot_q6K uses shared-memory addressing for its alignment probe; use the global-memory variant below
            if (P.type == T_Q4_K && ((uintptr_t) bp & 15) == 0) acc += dot_q4K(bp, r);
            else if (P.type == T_Q5_K && ((uintptr_t) bp & 15) == 0) acc += dot_q5K(bp, r);
            else {
                // byte-wise scalar path (rare): dequantize on the fly against int8 activations
                const int8_t * a8 = P.act.qs + (int64_t) b * 256;
                if (P.type == T_Q6_K) {
                    const float d = __half2float(__ushort_as_half(*reinterpret_cast<const uint16_t *>(bp + 208)));
                    const int8_t * sc = reinterpret_cast<const int8_t *>(bp + 192);
                    int sumi = 0;
                    for (int n = 0; n < 2; n++)
                        for (int l = 0; l < 32; l++) {
                            const uint8_t qa = bp[64 * n + l], qb = bp[64 * n + 32 + l], h = bp[128 + 32 * n + l];
                            const int is = l / 16;
                            const int q1 = (int) ((qa & 0xF) | (((h >> 0) & 3) << 4)) - 32;
                            const int q2 = (int) ((qb & 0xF) | (((h >> 2) & 3) << 4)) - 32;
                            const int q3 = (int) ((qa >> 4) | (((h >> 4) & 3) << 4)) - 32;
                            const int q4 = (int) ((qb >> 4) | (((h >> 6) & 3) << 4)) - 32;
                            sumi += sc[8 * n + is + 0] * q1 * a8[128 * n + l] + sc[8 * n + is + 2] * q2 * a8[128 * n + 32 + l] +
                                    sc[8 * n + is + 4] * q3 * a8[128 * n + 64 + l] + sc[8 * n + is + 6] * q4 * a8[128 * n + 96 + l];
                        }
                    acc += (d * r.d) * (float) sumi;
                } else {
                    const bool q5 = P.type == T_Q5_K;
                    const __half2 dm = *reinterpret_cast<const __half2 *>(bp);
                    const uint8_t * scb = bp + 4;
                    const uint8_t * qhb = bp + 16;
                    const uint8_t * qs = bp + (q5 ? 48 : 16);
                    int sumi = 0, summ = 0;
                    for (int j = 0; j < 8; j++) {
                        int sc, mn;
                        if (j < 4) { sc = scb[j] & 63; mn = scb[j + 4] & 63; }
                        else { sc = (scb[j + 4] & 0xF) | ((scb[j - 4] >> 6) << 4); mn = (scb[j + 4] >> 4) | ((scb[j] >> 6) << 4); }
                        int dsum = 0, asum = 0;
                        for (int l = 0; l < 32; l++) {
                            const uint8_t byte = qs[32 * (j / 2) + l];
                            int q = (j & 1) ? (byte >> 4) : (byte & 0xF);
                            if (q5 && ((qhb[l] >> j) & 1)) q += 16;
                            const int a = a8[32 * j + l];
                            dsum += q * a;
                            asum += a;
                        }
                        sumi += sc * dsum;
                        summ += mn * asum;
                    }
                    acc += (__low2float(dm) * r.d) * (float) sumi - (__high2float(dm) * r.d) * (float) summ;
                }
            }
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        if (P.bias) acc += P.bias[row];
        if (P.resid) acc += P.resid[row];
        P.y[row] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming GEMV for the 32-element block types (Q8_0 34 B, Q5_1 24 B per block): Qwen2.5-72B's ffn_down (K = 29 568, not a
// multiple of 256) falls back to them (src/llama.cpp:19516-19551) and they are a third of that model's bytes.
// The activation (q8_0 / q8_1) is staged once per CTA in shared memory; every warp streams whole rows through its own 4-deep
// cp.async ring of 32-block chunks (8-byte pieces, all lanes issue, ~100 KB in flight per SM at 3 CTAs/SM); lane l owns block
// l of every chunk, i.e. the same blocks and the same per-lane order as k_gemv_generic => bit-identical results.
constexpr int B32_NST = 4;
struct GemvB32Params {
    const uint8_t * W;
    float * y;
    const float * bias;
    const float * resid;
    int64_t row_bytes;
    int type, N, K, nb, bpb;
    ActQ act;
};
__global__ void __launch_bounds__(256) k_gemv_blk32(const __grid_constant__ GemvB32Params P) {
    extern __shared__ __align__(16) uint8_t b32_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kp = (P.K + 15) & ~15;
    int8_t * a_qs = reinterpret_cast<int8_t *>(b32_smem);                        // [K]
    float * a_d = reinterpret_cast<float *>(b32_smem + kp);                      // [nb]
    float * a_s = a_d + P.nb;                                                    // [nb]  (Q5_1 only)
    const int cb = 32 * P.bpb;                                                   // chunk bytes: 1088 / 768
    uint8_t * ring = b32_smem + ((kp + 8 * P.nb + 15) & ~15) + (size_t) warp * B32_NST * 1088;
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < P.K / 16; i += 256) reinterpret_cast<int4 *>(a_qs)[i] = reinterpret_cast<const int4 *>(P.act.qs)[i];
    for (int i = threadIdx.x; i < P.nb; i += 256) { a_d[i] = P.act.d[i]; if (P.type == T_Q5_1) a_s[i] = P.act.s[i]; }
    __syncthreads();
    const int nchunk = (P.nb + 31) / 32;
    const int pieces = cb / 8;
    for (int row = blockIdx.x * 8 + warp; row < P.N; row += gridDim.x * 8) {
        const uint8_t * wrow = P.W + (int64_t) row * P.row_bytes;
        auto issue = [&](int c) {
            if (c < nchunk) {
                const int64_t off = (int64_t) c * cb;
                uint8_t * dst = ring + (size_t) (c % B32_NST) * 1088;
                for (int pc = lane; pc < pieces; pc += 32)
                    if (off + pc * 8 + 8 <= P.row_bytes)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + pc * 8)), "l"(wrow + off + pc * 8) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int c = 0; c < B32_NST - 1; c++) issue(c);
        float acc = 0.f;
        for (int c = 0; c < nchunk; c++) {
            issue(c + B32_NST - 1);
            asm volatile("cp.async.wait_group %0;" ::"n"(B32_NST - 1) : "memory");
            __syncwarp();
            const int b = c * 32 + lane;
            if (b < P.nb) {
                const uint8_t * bp = ring + (size_t) (c % B32_NST) * 1088 + lane * P.bpb;
                const int4 * a = reinterpret_cast<const int4 *>(a_qs + (int64_t) b * 32);
                const int4 a0 = a[0], a1 = a[1];
                const int av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                int sumi = 0;
                if (P.type == T_Q8_0) {
                    // 34-byte blocks are only 2-byte aligned: read the 9 aligned words around the block; qs (offset 2) is either
                    // word-aligned already (block at 4k+2) or straddles two words (block at 4k): one PRMT per word, selector per lane
                    const uint32_t * wp = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(bp) & ~(uintptr_t) 3);
                    const bool odd = (reinterpret_cast<uintptr_t>(bp) & 2) != 0;
                    const uint32_t sel = odd ? 0x7654u : 0x5432u;
                    uint32_t w[9];
#pragma unroll
                    for (int i = 0; i < 9; i++) w[i] = wp[i];
                    const float d = __half2float(__ushort_as_half((unsigned short) ((w[0] >> (odd ? 16 : 0)) & 0xffff)));
#pragma unroll
                    for (int i = 0; i < 8; i++) sumi = dp4a_ss((int) __byte_perm(w[i], w[i + 1], sel), av[i], sumi);
                    acc += (float) sumi * (d * a_d[b]);
                } else {
                    const uint32_t dmw = *reinterpret_cast<const uint32_t *>(bp);
                    const float d = __half2float(__ushort_as_half((unsigned short) (dmw & 0xffff)));
                    const float mm = __half2float(__ushort_as_half((unsigned short) (dmw >> 16)));
                    const uint32_t qh = *reinterpret_cast<const uint32_t *>(bp + 4);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t w = *reinterpret_cast<const uint32_t *>(bp + 8 + 4 * i);
                        // bit k of a nibble of qh -> bit 4 of byte k: x * 0x00204081 puts bit k at 8k (no carries), then << 4
                        const uint32_t hb_lo = ((((qh >> (4 * i)) & 0xFu) * 0x00204081u) & 0x01010101u) << 4;
                        const uint32_t hb_hi = ((((qh >> (4 * i + 16)) & 0xFu) * 0x00204081u) & 0x01010101u) << 4;
                        sumi = dp4a_us((w & 0x0f0f0f0fu) | hb_lo, av[i], sumi);
                        sumi = dp4a_us(((w >> 4) & 0x0f0f0f0fu) | hb_hi, av[4 + i], sumi);
                    }
                    acc += (d * a_d[b]) * (float) sumi + mm * a_s[b];
                }
            }
            __syncwarp();   // the slot is refilled by the next issue()
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        acc = warp_sum(acc);
        if (lane == 0) {
            if (P.bias) acc += P.bias[row];
            if (P.resid) acc += P.resid[row];
            P.y[row] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
static int pick_rows_per_tile(int64_t row_bytes, int ngroups, int N);
bool gemv_fused_prologue_ok(int K);
static int g_sm_count = 0;
static bool g_attr_set = false;

int gemv_smem_bytes() { return GEMV_CTL_BYTES + GEMV_NSTAGE * GEMV_STAGE_BYTES + GEMV_ACT_SMEM; }

int sm_count() {
    if (!g_sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_sm_count;
}

static int pick_rows_per_tile(int64_t row_bytes, int ngroups, int N) {
    int fit = (int) ((GEMV_STAGE_BYTES - 16) / row_bytes);
    if (fit < 1) return 0;
    int tr = fit >= ngroups ? (fit / ngroups) * ngroups : fit;
    if (tr > 2 * ngroups && ngroups >= 8) tr = ngroups;      // 8 rows per stage is plenty; more stages in flight instead
    if (tr > 4 * ngroups) tr = 4 * ngroups;
    if (ngroups < GEMV_TEAM_W && tr > ngroups) tr = ngroups;  // split rows (wpr > 1): one row per warp group and stage
    if (tr > N) tr = N;
    return tr;
}

// Fused launch of up to 3 k-quant matrices sharing one q8_K activation.  Returns cudaError_t as int.
int launch_gemv_kquant(const GemvDesc * d, int nmat, int K, const ActQ & act, cudaStream_t stream, bool pdl) {
    GemvFused none{};
    return launch_gemv_kquant_fused(d, nmat, K, act, none, stream, pdl);
}

// bytes of one ring tile of a [N,K] matrix of `type` in the fast kernel (0 if the fast kernel does not apply)
uint32_t gemv_tile_bytes(int type, int K, int N) {
    if (!is_kquant(type) || !gemv_fused_prologue_ok(K)) return 0;
    int wpr = 1;
    while (wpr * 32 < K / 256) wpr *= 2;
    const int64_t rb = row_bytes(type, K);
    const int tr = pick_rows_per_tile(rb, GEMV_TEAM_W / wpr, N);
    return (uint32_t) (tr * rb);
}

bool gemv_fused_prologue_ok(int K) { return K % 256 == 0 && K / 256 <= GEMV_ACT_MAX_NBLK; }

int launch_gemv_kquant_fused(const GemvDesc * d, int nmat, int K, const ActQ & act, const GemvFused & pro, cudaStream_t stream, bool pdl) {
    if (nmat < 1 || nmat > GEMV_MAX_MAT || K % 256 != 0) return (int) cudaErrorInvalidValue;
    if (pro.kind != PRO_NONE && !gemv_fused_prologue_ok(K)) return (int) cudaErrorInvalidValue;
    const int nblk = K / 256;
    bool fast = nblk <= GEMV_MAX_NBLK && gemv_fused_prologue_ok(K);   // the activation is staged in shared memory
    GemvParams P{};
    if (fast) {
        int wpr = 1;
        while (wpr * 32 < nblk) wpr *= 2;
        P.wpr = wpr;
        P.nblk = nblk;
        P.K = K;
        P.nmat = nmat;
        P.act = act;
        P.prologue = pro.kind;
        P.in0 = pro.in0;
        P.in1 = pro.in1;
        P.eps = pro.eps;
        P.next_W = (const uint8_t *) pro.next_W;
        P.next_total_bytes = pro.next_total_bytes;
        P.next_tile_bytes = pro.next_tile_bytes;
        P.fill_before_wait = pro.fill_before_wait ? 1 : 0;
        if (pro.next_W && (((uintptr_t) pro.next_W & 15) || pro.next_tile_bytes == 0)) P.next_W = nullptr;
        int tiles = 0;
        for (int i = 0; i < nmat; i++) {
            GemvMat & M = P.mat[i];
            if (!is_kquant(d[i].type)) return (int) cudaErrorInvalidValue;
            M.W = (const uint8_t *) d[i].W;
            M.y = d[i].y;
            M.bias = d[i].bias;
            M.resid = d[i].resid;
            M.type = d[i].type;
            M.N = d[i].N;
            M.row_bytes = row_bytes(d[i].type, K);
            M.total_bytes = M.row_bytes * d[i].N;
            M.rows_per_tile = pick_rows_per_tile(M.row_bytes, GEMV_TEAM_W / wpr, d[i].N);
            if (M.rows_per_tile == 0 || ((uintptr_t) M.W & 15)) { fast = false; break; }
            M.tile0 = tiles;
            tiles += (d[i].N + M.rows_per_tile - 1) / M.rows_per_tile;
        }
        P.ntiles = tiles;
    }
    if (fast) {
        if (!g_attr_set) {
            cudaError_t e = cudaFuncSetAttribute(k_gemv_kquant, cudaFuncAttributeMaxDynamicSharedMemorySize, gemv_smem_bytes());
            if (e != cudaSuccess) return (int) e;
            g_attr_set = true;
        }
        int grid = sm_count();
        if (grid > P.ntiles) grid = P.ntiles;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(GEMV_THREADS);
        cfg.dynamicSmemBytes = gemv_smem_bytes();
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return (int) cudaLaunchKernelEx(&cfg, k_gemv_kquant, P);
    }
    if (pro.kind != PRO_NONE) return (int) cudaErrorInvalidValue;   // callers must check gemv_fused_prologue_ok / alignment
    for (int i = 0; i < nmat; i++) {
        int e = launch_gemv_generic(d[i], K, act, stream, pdl);
        if (e) return e;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// host side of the persistent token kernel
struct MkHandle {
    MkParams P{};
    MkLayer * d_layers = nullptr;
    unsigned int * d_barrier = nullptr;   // [0] barrier word, [1] error flag
    int grid = 0;
};

static bool mk_fill_phase(MkPhase & ph, const MkGemvDesc & g) {
    if (g.nmat < 1 || g.nmat > GEMV_MAX_MAT || !gemv_fused_prologue_ok(g.K)) return false;
    const int nblk = g.K / 256;
    int wpr = 1;
    while (wpr * 32 < nblk) wpr *= 2;
    ph.nmat = g.nmat; ph.K = g.K; ph.nblk = nblk; ph.wpr = wpr;
    ph.prologue = g.pro.kind; ph.in0 = g.pro.in0; ph.in1 = g.pro.in1; ph.eps = g.pro.eps; ph.act = g.act;
    int tiles = 0;
    for (int i = 0; i < g.nmat; i++) {
        GemvMat & M = ph.mat[i];
        if (!is_kquant(g.d[i].type) || ((uintptr_t) g.d[i].W & 15)) return false;
        M.W = (const uint8_t *) g.d[i].W; M.y = g.d[i].y; M.bias = g.d[i].bias; M.resid = g.d[i].resid;
        M.type = g.d[i].type; M.N = g.d[i].N;
        M.row_bytes = row_bytes(g.d[i].type, g.K);
        M.total_bytes = M.row_bytes * g.d[i].N;
        M.rows_per_tile = pick_rows_per_tile(M.row_bytes, GEMV_TEAM_W / wpr, g.d[i].N);
        if (M.rows_per_tile == 0) return false;
        M.tile0 = tiles;
        tiles += (g.d[i].N + M.rows_per_tile - 1) / M.rows_per_tile;
    }
    ph.ntiles = tiles;
    return true;
}

static int mk_smem_bytes() { return MK_HDR_BYTES + GEMV_NSTAGE * GEMV_STAGE_BYTES + GEMV_ACT_SMEM; }

MkHandle * mk_build(const MkTokenDesc & t, const RopeParams & rp) {
    if (t.n_layers < 1 || 4 * t.n_layers + 1 > MK_MAX_PHASES || t.n_head > sm_count()) return nullptr;
    // attention scratch lives in the activation region of shared memory
    const size_t attn_bytes = ((size_t) ((t.n_ctx + 31) & ~31) + 8 * 128 + 128) * 4 + 2 * 128 * 2 + 8 * 4 + 8 * 8 + 16;
    if (attn_bytes > (size_t) GEMV_ACT_SMEM) return nullptr;
    std::vector<MkLayer> L((size_t) t.n_layers);
    int wsplit = 1;
    for (int il = 0; il < t.n_layers; il++) {
        const MkLayerDesc & s = t.layers[il];
        for (int p = 0; p < 4; p++) {
            if (!mk_fill_phase(L[il].ph[p], s.ph[p])) return nullptr;
            const int w = L[il].ph[p].wpr;
            if (w > 1) { if (wsplit > 1 && w != wsplit) return nullptr; wsplit = w; }
            if (L[il].ph[p].ntiles / 1 > 65535 * sm_count()) return nullptr;
        }
        if (s.F % 256 != 0 || s.F / 256 > sm_count()) return nullptr;
        L[il].q = s.q; L[il].k = s.k; L[il].v = s.v; L[il].kc = s.kc; L[il].vc = s.vc; L[il].att = s.att;
        L[il].g = s.g; L[il].u = s.u; L[il].actF = s.actF; L[il].F = s.F;
    }
    MkHandle * h = new MkHandle();
    if (t.with_head) {
        if (!mk_fill_phase(h->P.head, t.head)) { delete h; return nullptr; }
        const int w = h->P.head.wpr;
        if (w > 1 && wsplit > 1 && w != wsplit) { delete h; return nullptr; }
    }
    if (cudaMalloc(&h->d_layers, sizeof(MkLayer) * L.size()) != cudaSuccess || cudaMalloc(&h->d_barrier, 256) != cudaSuccess) {
        cudaGetLastError();
        mk_free(h);
        return nullptr;
    }
    cudaMemcpy(h->d_layers, L.data(), sizeof(MkLayer) * L.size(), cudaMemcpyHostToDevice);
    cudaMemset(h->d_barrier, 0, 256);
    h->P.layers = h->d_layers;
    h->P.n_layers = t.n_layers;
    h->P.with_head = t.with_head ? 1 : 0;
    h->P.pos_dev = t.pos_dev;
    h->P.rp = rp;
    h->P.freq_factors = t.freq_factors;
    h->P.kq_scale = t.kq_scale;
    h->P.n_head = t.n_head; h->P.n_head_kv = t.n_head_kv; h->P.n_ctx = t.n_ctx;
    h->P.barrier = h->d_barrier;
    h->P.error_flag = (int *) (h->d_barrier + 1);
    h->P.xb_allowed = getenv("PB200_XB") ? atoi(getenv("PB200_XB")) : MK_XB_ALLOWED;
    if (h->P.xb_allowed < 0) h->P.xb_allowed = 0;
    if (h->P.xb_allowed > GEMV_NSTAGE) h->P.xb_allowed = GEMV_NSTAGE;
    h->grid = sm_count();
    if (cudaFuncSetAttribute(k_token_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, mk_smem_bytes()) != cudaSuccess) {
        cudaGetLastError();
        mk_free(h);
        return nullptr;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_token_persistent, GEMV_THREADS, mk_smem_bytes()) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        mk_free(h);
        return nullptr;
    }
    return h;
}

int mk_launch(MkHandle * h, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(h->d_barrier, 0, 4, stream);
    if (e != cudaSuccess) return (int) e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(h->grid);
    cfg.blockDim = dim3(GEMV_THREADS);
    cfg.dynamicSmemBytes = mk_smem_bytes();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the grid barriers cannot deadlock on scheduling
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int) cudaLaunchKernelEx(&cfg, k_token_persistent, h->P);
}

int mk_error(MkHandle * h) {
    int f = 0;
    cudaMemcpy(&f, h->d_barrier + 1, 4, cudaMemcpyDeviceToHost);
    return f;
}

void mk_free(MkHandle * h) {
    if (!h) return;
    if (h->d_layers) cudaFree(h->d_layers);
    if (h->d_barrier) cudaFree(h->d_barrier);
    delete h;
}

static int launch_gemv_blk32(const GemvDesc & d, int K, const ActQ & act, cudaStream_t stream, bool pdl) {
    GemvB32Params P{};
    P.W = (const uint8_t *) d.W;
    P.y = d.y;
    P.bias = d.bias;
    P.resid = d.resid;
    P.type = d.type;
    P.N = d.N;
    P.K = K;
    P.nb = K / 32;
    P.bpb = d.type == T_Q8_0 ? BYTES_Q8_0 : BYTES_Q5_1;
    P.row_bytes = row_bytes(d.type, K);
    P.act = act;
    const int kp = (K + 15) & ~15;
    const size_t smem = (size_t) ((kp + 8 * P.nb + 15) & ~15) + (size_t) 8 * B32_NST * 1088;
    static size_t configured = 0;
    if (smem > configured && smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_gemv_blk32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return (int) e;
        configured = smem;
    }
    const int per_sm = (int) std::max<size_t>(1, std::min<size_t>(4, (224 * 1024) / (smem + 1024)));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(std::min((d.N + 7) / 8, sm_count() * per_sm));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int) cudaLaunchKernelEx(&cfg, k_gemv_blk32, P);
}

int launch_gemv_generic(const GemvDesc & d, int K, const ActQ & act, cudaStream_t stream, bool pdl) {
    // 32-element block types with 8-byte aligned rows and an activation that fits in shared memory: the streaming kernel
    static const bool no_b32 = getenv("PB200_NO_BLK32") != nullptr;
    if (!no_b32 && (d.type == T_Q8_0 || d.type == T_Q5_1) && K % 32 == 0 && row_bytes(d.type, K) % 8 == 0 && K % 16 == 0 && K <= 131072 &&
        ((uintptr_t) d.W & 7) == 0)
        return launch_gemv_blk32(d, K, act, stream, pdl);
    GemvGenericParams P{};
    P.W = (const uint8_t *) d.W;
    P.y = d.y;
    P.bias = d.bias;
    P.resid = d.resid;
    P.type = d.type;
    P.N = d.N;
    P.K = K;
    P.row_bytes = row_bytes(d.type, K);
    P.act = act;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((d.N + 7) / 8);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int) cudaLaunchKernelEx(&cfg, k_gemv_generic, P);
}

}  // namespace pb
