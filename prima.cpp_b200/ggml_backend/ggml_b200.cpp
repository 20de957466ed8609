// prima.cpp_b200/ggml_backend/ggml_b200.cpp — ggml backend "B200": the reference's plugin vtables (ggml-backend-impl.h:15-220)
// implemented on top of the C ABI of libprima_b200.so.  Written against the interface, not against ggml-cuda.cu: buffers are
// plain cudaMalloc regions, there is one stream per backend instance, graph_compute maps each node to one or two pb200_* calls.
//
// supports_op is exact: only what the Llama / Qwen2 decode graph needs (build_llama / build_qwen2,
// src/llama.cpp:11000-11216, 12736-12916, FA off); everything else returns false so the scheduler keeps it on the CPU.
#include "ggml-backend-impl.h"
#include "ggml-backend.h"
#include "ggml.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ggml_b200.h"
#include "../../include/prima_b200.h"

#define B200_MAX_DEVICES 16

static std::atomic<unsigned long long> g_nodes{0};

#define CUDA_OK(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            fprintf(stderr, "ggml-b200: %s failed: %s (%s:%d)\n", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            GGML_ABORT("CUDA error");   /* same convention as ggml_cuda_error, ggml-cuda.cu:62-72 */    \
        }                                                                                               \
    } while (0)

// ---------------------------------------------------------------------------------------------------- contexts
struct b200_device_ctx {
    int device;
    std::string name, description;
};
struct b200_backend_ctx {
    int device;
    cudaStream_t stream = nullptr;
    void * act_ws = nullptr;       // quantized-activation workspace (grown on demand)
    size_t act_ws_bytes = 0;
    void * mmq_ws = nullptr;       // fp16 activation tiles for the tensor-core path (grown on demand)
    size_t mmq_ws_bytes = 0;
    std::string name;
};
struct b200_buffer_ctx {
    int device;
    void * base;
};

static const int64_t MMQ_MIN_COLS = 8;
static bool mmq_ok(enum ggml_type t, int64_t k) {   // mirrors mmq_supported() of the library
    if (t == GGML_TYPE_Q4_K || t == GGML_TYPE_Q5_K || t == GGML_TYPE_Q6_K) return k % 256 == 0;
    return (t == GGML_TYPE_Q8_0 || t == GGML_TYPE_Q5_1) && k % 64 == 0 && k >= 256;
}
static bool type_is_quant(enum ggml_type t) {
    return t == GGML_TYPE_Q4_K || t == GGML_TYPE_Q5_K || t == GGML_TYPE_Q6_K || t == GGML_TYPE_Q8_0 || t == GGML_TYPE_Q5_1;
}

// ---------------------------------------------------------------------------------------------------- buffer
static const char * b200_buffer_get_name(ggml_backend_buffer_t) { return "B200"; }
static void b200_buffer_free(ggml_backend_buffer_t buffer) {
    b200_buffer_ctx * ctx = (b200_buffer_ctx *) buffer->context;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->base);
    delete ctx;
}
static void * b200_buffer_get_base(ggml_backend_buffer_t buffer) { return ((b200_buffer_ctx *) buffer->context)->base; }
static void b200_buffer_init_tensor(ggml_backend_buffer_t buffer, ggml_tensor * tensor) {
    // like ggml-cuda.cu:444-462: zero the padding behind quantized rows so that over-reads see defined bytes
    if (tensor->view_src == nullptr && ggml_is_quantized(tensor->type)) {
        b200_buffer_ctx * ctx = (b200_buffer_ctx *) buffer->context;
        const size_t sz = ggml_nbytes(tensor);
        const size_t padded = ggml_backend_buft_get_alloc_size(buffer->buft, tensor);
        if (padded > sz) {
            cudaSetDevice(ctx->device);
            CUDA_OK(cudaMemset((char *) tensor->data + sz, 0, padded - sz));
        }
    }
}
static void b200_buffer_memset_tensor(ggml_backend_buffer_t buffer, ggml_tensor * tensor, uint8_t value, size_t offset, size_t size) {
    cudaSetDevice(((b200_buffer_ctx *) buffer->context)->device);
    CUDA_OK(cudaMemset((char *) tensor->data + offset, value, size));
}
static void b200_buffer_set_tensor(ggml_backend_buffer_t buffer, ggml_tensor * tensor, const void * data, size_t offset, size_t size) {
    cudaSetDevice(((b200_buffer_ctx *) buffer->context)->device);
    CUDA_OK(cudaMemcpy((char *) tensor->data + offset, data, size, cudaMemcpyHostToDevice));   // synchronous w.r.t. the caller
}
static void b200_buffer_get_tensor(ggml_backend_buffer_t buffer, const ggml_tensor * tensor, void * data, size_t offset, size_t size) {
    cudaSetDevice(((b200_buffer_ctx *) buffer->context)->device);
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(data, (const char *) tensor->data + offset, size, cudaMemcpyDeviceToHost));
}
static bool b200_buffer_cpy_tensor(ggml_backend_buffer_t buffer, const ggml_tensor * src, ggml_tensor * dst) {
    if (src->buffer && src->buffer->iface.get_name == b200_buffer_get_name && ggml_is_contiguous(src) && ggml_is_contiguous(dst)) {
        cudaSetDevice(((b200_buffer_ctx *) buffer->context)->device);
        CUDA_OK(cudaDeviceSynchronize());
        CUDA_OK(cudaMemcpy(dst->data, src->data, ggml_nbytes(src), cudaMemcpyDeviceToDevice));
        return true;
    }
    return false;
}
static void b200_buffer_clear(ggml_backend_buffer_t buffer, uint8_t value) {
    b200_buffer_ctx * ctx = (b200_buffer_ctx *) buffer->context;
    cudaSetDevice(ctx->device);
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemset(ctx->base, value, buffer->size));
}
static const ggml_backend_buffer_i b200_buffer_iface = {
    /* .get_name      = */ b200_buffer_get_name,
    /* .free_buffer   = */ b200_buffer_free,
    /* .get_base      = */ b200_buffer_get_base,
    /* .init_tensor   = */ b200_buffer_init_tensor,
    /* .memset_tensor = */ b200_buffer_memset_tensor,
    /* .set_tensor    = */ b200_buffer_set_tensor,
    /* .get_tensor    = */ b200_buffer_get_tensor,
    /* .cpy_tensor    = */ b200_buffer_cpy_tensor,
    /* .clear         = */ b200_buffer_clear,
    /* .reset         = */ nullptr,
};

// ---------------------------------------------------------------------------------------------------- buffer type
static const char * b200_buft_get_name(ggml_backend_buffer_type_t buft) { return ((b200_device_ctx *) buft->device->context)->name.c_str(); }
static ggml_backend_buffer_t b200_buft_alloc(ggml_backend_buffer_type_t buft, size_t size) {
    b200_device_ctx * dctx = (b200_device_ctx *) buft->device->context;
    cudaSetDevice(dctx->device);
    void * p = nullptr;
    size = size > 0 ? size : 1;
    if (cudaMalloc(&p, size + 256) != cudaSuccess) {   // allocation failure -> NULL (ggml-cuda.cu:556-561), no abort
        cudaGetLastError();
        return nullptr;
    }
    return ggml_backend_buffer_init(buft, b200_buffer_iface, new b200_buffer_ctx{dctx->device, p}, size);
}
static size_t b200_buft_alignment(ggml_backend_buffer_type_t) { return 128; }
static size_t b200_buft_alloc_size(ggml_backend_buffer_type_t, const ggml_tensor * tensor) {
    size_t size = ggml_nbytes(tensor);
    if (ggml_is_quantized(tensor->type)) size = (size + 15) / 16 * 16 + 16;   // the GEMV's bulk copies move whole 16-B units
    return size;
}
static const ggml_backend_buffer_type_i b200_buft_iface = {
    /* .get_name       = */ b200_buft_get_name,
    /* .alloc_buffer   = */ b200_buft_alloc,
    /* .get_alignment  = */ b200_buft_alignment,
    /* .get_max_size   = */ nullptr,
    /* .get_alloc_size = */ b200_buft_alloc_size,
    /* .is_host        = */ nullptr,
};

// ---------------------------------------------------------------------------------------------------- op support
static bool is_noop(enum ggml_op op) {
    return op == GGML_OP_NONE || op == GGML_OP_RESHAPE || op == GGML_OP_VIEW || op == GGML_OP_PERMUTE || op == GGML_OP_TRANSPOSE;
}
// b is broadcast over a by plain repetition of its contiguous data (bias / norm weight / same shape)
static bool bcast_ok(const ggml_tensor * a, const ggml_tensor * b) {
    if (!ggml_is_contiguous(a) || !ggml_is_contiguous(b)) return false;
    bool tail = false;
    for (int d = 0; d < GGML_MAX_DIMS; d++) {
        if (b->ne[d] == a->ne[d] && !tail) continue;
        if (b->ne[d] == 1) { tail = true; continue; }
        return false;
    }
    return true;
}
static bool b200_supports_op(ggml_backend_dev_t, const ggml_tensor * op) {
    const ggml_tensor * a = op->src[0], * b = op->src[1];
    if (is_noop(op->op)) return true;
    switch (op->op) {
        case GGML_OP_RMS_NORM:
            return a->type == GGML_TYPE_F32 && op->type == GGML_TYPE_F32 && ggml_is_contiguous(a) && ggml_is_contiguous(op);
        case GGML_OP_ADD:
        case GGML_OP_MUL:
            return a->type == GGML_TYPE_F32 && b->type == GGML_TYPE_F32 && op->type == GGML_TYPE_F32 && ggml_is_contiguous(op) && bcast_ok(a, b);
        case GGML_OP_UNARY:
            return ggml_get_unary_op(op) == GGML_UNARY_OP_SILU && a->type == GGML_TYPE_F32 && ggml_is_contiguous(a) && ggml_is_contiguous(op);
        case GGML_OP_MUL_MAT: {
            if (b->type != GGML_TYPE_F32 || op->type != GGML_TYPE_F32) return false;
            if (a->ne[2] == 0 || a->ne[3] == 0 || b->ne[2] % a->ne[2] || b->ne[3] % a->ne[3]) return false;
            if (a->type == GGML_TYPE_F16) return a->nb[0] == sizeof(ggml_fp16_t) && b->nb[0] == sizeof(float) && op->nb[0] == sizeof(float);
            if (!type_is_quant(a->type)) return false;
            // quantized weights: rows of src1 must be dense.  The decode GEMV handles one activation column per launch; k-quant
            // weights with K % 256 == 0 take the tensor-core path (pb200_mul_mat_q) for any number of columns.
            if (!(ggml_is_contiguous(a) && b->nb[0] == sizeof(float) && ggml_is_contiguous(op) && a->ne[0] % ggml_blck_size(a->type) == 0)) return false;
            if (mmq_ok(a->type, a->ne[0]) && b->nb[1] % 16 == 0) return true;
            return b->ne[1] * b->ne[2] * b->ne[3] <= 64;
        }
        case GGML_OP_ROPE: {
            const int mode = ((const int32_t *) op->op_params)[2];
            return a->type == GGML_TYPE_F32 && op->type == GGML_TYPE_F32 && ggml_is_contiguous(a) && ggml_is_contiguous(op) && (mode == 0 || mode == 2) &&
                   a->ne[3] == 1 && (op->src[2] == nullptr || op->src[2]->type == GGML_TYPE_F32);
        }
        case GGML_OP_SOFT_MAX: {
            float max_bias;
            memcpy(&max_bias, (const float *) op->op_params + 1, sizeof(float));
            return a->type == GGML_TYPE_F32 && ggml_is_contiguous(a) && ggml_is_contiguous(op) && max_bias == 0.0f &&
                   (b == nullptr || (b->type == GGML_TYPE_F32 && ggml_is_contiguous(b) && b->ne[0] == a->ne[0] && b->ne[1] >= a->ne[1]));
        }
        case GGML_OP_CPY:
        case GGML_OP_DUP:
        case GGML_OP_CONT:
            return a->type == GGML_TYPE_F32 && (op->type == GGML_TYPE_F32 || op->type == GGML_TYPE_F16) && ggml_nelements(a) == ggml_nelements(op);
        case GGML_OP_GET_ROWS:
            return (a->type == GGML_TYPE_F32 || a->type == GGML_TYPE_F16 || type_is_quant(a->type)) && b->type == GGML_TYPE_I32 && op->type == GGML_TYPE_F32 &&
                   ggml_is_contiguous(a) && ggml_is_contiguous(b) && ggml_is_contiguous(op) && a->ne[2] == 1 && a->ne[3] == 1 && b->ne[2] == 1 && b->ne[3] == 1;
        default:
            return false;
    }
}

// ---------------------------------------------------------------------------------------------------- compute
#define PB_OK(expr)                                                                          \
    do {                                                                                     \
        int rc_ = (expr);                                                                    \
        if (rc_ != 0) {                                                                      \
            fprintf(stderr, "ggml-b200: %s -> %s (%d)\n", #expr, pb200_error_string(rc_), rc_); \
            return false;                                                                    \
        }                                                                                    \
    } while (0)

static bool b200_compute_node(b200_backend_ctx * ctx, ggml_tensor * dst) {
    const ggml_tensor * a = dst->src[0], * b = dst->src[1];
    void * st = ctx->stream;
    switch (dst->op) {
        case GGML_OP_RMS_NORM: {
            float eps;
            memcpy(&eps, dst->op_params, sizeof(float));
            PB_OK(pb200_rms_norm((const float *) a->data, (float *) dst->data, a->ne[0], ggml_nrows(a), eps, st));
            return true;
        }
        case GGML_OP_ADD:
        case GGML_OP_MUL:
            PB_OK(pb200_binary(dst->op == GGML_OP_ADD ? 0 : 1, (const float *) a->data, (const float *) b->data, (float *) dst->data, ggml_nelements(dst),
                               ggml_nelements(b), st));
            return true;
        case GGML_OP_UNARY:
            PB_OK(pb200_silu((const float *) a->data, (float *) dst->data, ggml_nelements(dst), st));
            return true;
        case GGML_OP_MUL_MAT: {
            if (a->type == GGML_TYPE_F16) {
                const int64_t ne[4] = {dst->ne[0], dst->ne[1], dst->ne[2], dst->ne[3]};
                const int64_t as[4] = {(int64_t) a->nb[0], (int64_t) a->nb[1], (int64_t) a->nb[2], (int64_t) a->nb[3]};
                const int64_t bs[4] = {(int64_t) b->nb[0], (int64_t) b->nb[1], (int64_t) b->nb[2], (int64_t) b->nb[3]};
                const int64_t ds[4] = {(int64_t) dst->nb[0], (int64_t) dst->nb[1], (int64_t) dst->nb[2], (int64_t) dst->nb[3]};
                PB_OK(pb200_mul_mat_f16(a->data, (const float *) b->data, (float *) dst->data, a->ne[0], ne, b->ne[2] / a->ne[2], b->ne[3] / a->ne[3], as, bs,
                                        ds, st));
                return true;
            }
            const int64_t K = a->ne[0], N = a->ne[1];
            const size_t need = pb200_act_workspace_bytes(K);
            if (need > ctx->act_ws_bytes) {
                if (ctx->act_ws) { CUDA_OK(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->act_ws); }
                CUDA_OK(cudaMalloc(&ctx->act_ws, need + 256));
                ctx->act_ws_bytes = need;
            }
            const int64_t r2 = b->ne[2] / a->ne[2], r3 = b->ne[3] / a->ne[3];
            if (b->ne[1] >= MMQ_MIN_COLS && mmq_ok(a->type, K) && b->nb[1] % 16 == 0 && (uintptr_t) b->data % 16 == 0 && b->nb[2] % 16 == 0 && b->nb[3] % 16 == 0 &&
                (uintptr_t) a->data % 16 == 0 && a->nb[2] % 16 == 0 && a->nb[3] % 16 == 0) {   // the prep kernel reads rows as float4
                // batched / prefill: the reference switches to mul_mat_q above 8 columns as well (ggml-cuda/mmq.cu:137-139)
                const size_t need_q = pb200_mul_mat_q_workspace_bytes(K, b->ne[1]);
                if (need_q > ctx->mmq_ws_bytes) {
                    if (ctx->mmq_ws) { CUDA_OK(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->mmq_ws); }
                    CUDA_OK(cudaMalloc(&ctx->mmq_ws, need_q + 256));
                    ctx->mmq_ws_bytes = need_q;
                }
                for (int64_t i3 = 0; i3 < b->ne[3]; i3++)
                    for (int64_t i2 = 0; i2 < b->ne[2]; i2++) {
                        const char * w = (const char *) a->data + (i2 / r2) * a->nb[2] + (i3 / r3) * a->nb[3];
                        const float * x = (const float *) ((const char *) b->data + i2 * b->nb[2] + i3 * b->nb[3]);
                        float * y = (float *) ((char *) dst->data + i2 * dst->nb[2] + i3 * dst->nb[3]);
                        PB_OK(pb200_mul_mat_q((int) a->type, w, N, K, x, (int64_t) (b->nb[1] / sizeof(float)), b->ne[1], y, nullptr, nullptr, ctx->mmq_ws, st));
                    }
                return true;
            }
            for (int64_t i3 = 0; i3 < b->ne[3]; i3++)
                for (int64_t i2 = 0; i2 < b->ne[2]; i2++)
                    for (int64_t i1 = 0; i1 < b->ne[1]; i1++) {
                        const char * w = (const char *) a->data + (i2 / r2) * a->nb[2] + (i3 / r3) * a->nb[3];
                        const float * x = (const float *) ((const char *) b->data + i1 * b->nb[1] + i2 * b->nb[2] + i3 * b->nb[3]);
                        float * y = (float *) ((char *) dst->data + i1 * dst->nb[1] + i2 * dst->nb[2] + i3 * dst->nb[3]);
                        PB_OK(pb200_mul_mat_vec((int) a->type, w, N, K, x, y, ctx->act_ws, st));
                    }
            return true;
        }
        case GGML_OP_ROPE: {
            const int32_t * p = (const int32_t *) dst->op_params;
            float fb, fs, ef, af, bf, bsl;
            memcpy(&fb, p + 5, 4); memcpy(&fs, p + 6, 4); memcpy(&ef, p + 7, 4); memcpy(&af, p + 8, 4); memcpy(&bf, p + 9, 4); memcpy(&bsl, p + 10, 4);
            PB_OK(pb200_rope((const float *) a->data, (float *) dst->data, a->ne[2], (int) a->ne[1], (int) a->ne[0], p[1], p[2], (const int32_t *) b->data, fb, fs,
                             ef, af, bf, bsl, p[4], dst->src[2] ? (const float *) dst->src[2]->data : nullptr, st));
            return true;
        }
        case GGML_OP_SOFT_MAX: {
            float scale;
            memcpy(&scale, dst->op_params, sizeof(float));
            PB_OK(pb200_soft_max((const float *) a->data, b ? (const float *) b->data : nullptr, (float *) dst->data, a->ne[0], ggml_nrows(a), a->ne[1], scale, st));
            return true;
        }
        case GGML_OP_CPY:
        case GGML_OP_DUP:
        case GGML_OP_CONT: {
            // enumerate in the SOURCE's logical order when shapes agree, else both sides must be contiguous-compatible
            const ggml_tensor * d = dst->op == GGML_OP_CPY ? dst->src[1] : dst;
            void * out = dst->op == GGML_OP_CPY ? dst->src[1]->data : dst->data;
            int64_t ne[4], sb[4], db[4];
            if (ggml_are_same_shape(a, d)) {
                for (int i = 0; i < 4; i++) { ne[i] = a->ne[i]; sb[i] = a->nb[i]; db[i] = d->nb[i]; }
            } else if (ggml_is_contiguous(d)) {          // e.g. cont_2d of a permuted tensor, or cpy into a flat cache view
                const int64_t es = (int64_t) ggml_type_size(d->type);
                for (int i = 0; i < 4; i++) { ne[i] = a->ne[i]; sb[i] = a->nb[i]; }
                db[0] = es; db[1] = es * ne[0]; db[2] = db[1] * ne[1]; db[3] = db[2] * ne[2];
            } else if (ggml_is_contiguous(a)) {
                for (int i = 0; i < 4; i++) { ne[i] = d->ne[i]; db[i] = d->nb[i]; }
                sb[0] = 4; sb[1] = 4 * ne[0]; sb[2] = sb[1] * ne[1]; sb[3] = sb[2] * ne[2];
            } else {
                return false;
            }
            PB_OK(pb200_copy_strided(a->data, out, d->type == GGML_TYPE_F16, ne, sb, db, st));
            return true;
        }
        case GGML_OP_GET_ROWS:
            PB_OK(pb200_get_rows((int) a->type, a->data, a->ne[0], (const int32_t *) b->data, ggml_nelements(b), (float *) dst->data, st));
            return true;
        default:
            return false;
    }
}

// ---------------------------------------------------------------------------------------------------- backend
static ggml_guid_t b200_guid() {
    static ggml_guid guid = {0xb2, 0x00, 0x5e, 0x10, 0x0a, 0x47, 0x4d, 0x41, 0x9c, 0x21, 0x70, 0x72, 0x69, 0x6d, 0x61, 0x01};
    return &guid;
}
static const char * b200_backend_get_name(ggml_backend_t backend) { return ((b200_backend_ctx *) backend->context)->name.c_str(); }
static void b200_backend_free(ggml_backend_t backend) {
    b200_backend_ctx * ctx = (b200_backend_ctx *) backend->context;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->act_ws) cudaFree(ctx->act_ws);
    if (ctx->mmq_ws) cudaFree(ctx->mmq_ws);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    delete backend;
}
static ggml_backend_buffer_type_t b200_backend_get_default_buft(ggml_backend_t backend) {
    return ggml_backend_b200_buffer_type(((b200_backend_ctx *) backend->context)->device);
}
static void b200_backend_set_tensor_async(ggml_backend_t backend, ggml_tensor * tensor, const void * data, size_t offset, size_t size) {
    b200_backend_ctx * ctx = (b200_backend_ctx *) backend->context;
    cudaSetDevice(ctx->device);
    CUDA_OK(cudaMemcpyAsync((char *) tensor->data + offset, data, size, cudaMemcpyHostToDevice, ctx->stream));
}
static void b200_backend_get_tensor_async(ggml_backend_t backend, const ggml_tensor * tensor, void * data, size_t offset, size_t size) {
    b200_backend_ctx * ctx = (b200_backend_ctx *) backend->context;
    cudaSetDevice(ctx->device);
    CUDA_OK(cudaMemcpyAsync(data, (const char *) tensor->data + offset, size, cudaMemcpyDeviceToHost, ctx->stream));
}
static void b200_backend_synchronize(ggml_backend_t backend) {
    b200_backend_ctx * ctx = (b200_backend_ctx *) backend->context;
    cudaSetDevice(ctx->device);
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
}
static enum ggml_status b200_backend_graph_compute(ggml_backend_t backend, ggml_cgraph * cgraph) {
    b200_backend_ctx * ctx = (b200_backend_ctx *) backend->context;
    cudaSetDevice(ctx->device);
    const int n = ggml_graph_n_nodes(cgraph);
    for (int i = 0; i < n; i++) {
        ggml_tensor * node = ggml_graph_node(cgraph, i);
        if (ggml_is_empty(node) || is_noop(node->op)) continue;
        if (!b200_compute_node(ctx, node)) {
            fprintf(stderr, "ggml-b200: op %s not supported inside graph_compute (supports_op must be consulted)\n", ggml_op_name(node->op));
            GGML_ABORT("unsupported op");   // ggml-cuda.cu:2671-2675
        }
        g_nodes++;
    }
    return GGML_STATUS_SUCCESS;   // asynchronous: work is enqueued on the backend stream
}
static const ggml_backend_i b200_backend_iface = {
    /* .get_name                = */ b200_backend_get_name,
    /* .free                    = */ b200_backend_free,
    /* .get_default_buffer_type = */ b200_backend_get_default_buft,
    /* .set_tensor_async        = */ b200_backend_set_tensor_async,
    /* .get_tensor_async        = */ b200_backend_get_tensor_async,
    /* .cpy_tensor_async        = */ nullptr,
    /* .synchronize             = */ b200_backend_synchronize,
    /* .graph_plan_create       = */ nullptr,
    /* .graph_plan_free         = */ nullptr,
    /* .graph_plan_update       = */ nullptr,
    /* .graph_plan_compute      = */ nullptr,
    /* .graph_compute           = */ b200_backend_graph_compute,
    /* .supports_op             = */ nullptr,
    /* .supports_buft           = */ nullptr,
    /* .offload_op              = */ nullptr,
    /* .event_record            = */ nullptr,
    /* .event_wait              = */ nullptr,
};

// ---------------------------------------------------------------------------------------------------- device + reg
static const char * b200_dev_get_name(ggml_backend_dev_t dev) { return ((b200_device_ctx *) dev->context)->name.c_str(); }
static const char * b200_dev_get_description(ggml_backend_dev_t dev) { return ((b200_device_ctx *) dev->context)->description.c_str(); }
static void b200_dev_get_memory(ggml_backend_dev_t dev, size_t * free, size_t * total) {
    cudaSetDevice(((b200_device_ctx *) dev->context)->device);
    CUDA_OK(cudaMemGetInfo(free, total));
}
static enum ggml_backend_dev_type b200_dev_get_type(ggml_backend_dev_t) { return GGML_BACKEND_DEVICE_TYPE_GPU_FULL; }
static void b200_dev_get_props(ggml_backend_dev_t dev, ggml_backend_dev_props * props) {
    props->name = b200_dev_get_name(dev);
    props->description = b200_dev_get_description(dev);
    props->type = b200_dev_get_type(dev);
    b200_dev_get_memory(dev, &props->memory_free, &props->memory_total);
    props->caps = { /* async */ true, /* host_buffer */ false, /* buffer_from_host_ptr */ false, /* events */ false };
}
static ggml_backend_t b200_dev_init_backend(ggml_backend_dev_t dev, const char *) { return ggml_backend_b200_init(((b200_device_ctx *) dev->context)->device); }
static ggml_backend_buffer_type_t b200_dev_get_buft(ggml_backend_dev_t dev) { return ggml_backend_b200_buffer_type(((b200_device_ctx *) dev->context)->device); }
static bool b200_dev_supports_buft(ggml_backend_dev_t dev, ggml_backend_buffer_type_t buft) {
    return buft->iface.get_name == b200_buft_get_name && buft->device == dev;
}
static bool b200_dev_offload_op(ggml_backend_dev_t, const ggml_tensor *) { return false; }   // never pull CPU-resident weights over PCIe (App. B)

static const ggml_backend_device_i b200_device_iface = {
    /* .get_name             = */ b200_dev_get_name,
    /* .get_description      = */ b200_dev_get_description,
    /* .get_memory           = */ b200_dev_get_memory,
    /* .get_type             = */ b200_dev_get_type,
    /* .get_props            = */ b200_dev_get_props,
    /* .init_backend         = */ b200_dev_init_backend,
    /* .get_buffer_type      = */ b200_dev_get_buft,
    /* .get_host_buffer_type = */ nullptr,
    /* .buffer_from_host_ptr = */ nullptr,
    /* .supports_op          = */ b200_supports_op,
    /* .supports_buft        = */ b200_dev_supports_buft,
    /* .offload_op           = */ b200_dev_offload_op,
    /* .event_new            = */ nullptr,
    /* .event_free           = */ nullptr,
    /* .event_synchronize    = */ nullptr,
};

struct b200_reg_ctx {
    std::vector<ggml_backend_device> devices;
    std::vector<ggml_backend_buffer_type> bufts;
};
static const char * b200_reg_get_name(ggml_backend_reg_t) { return "B200"; }
static size_t b200_reg_device_count(ggml_backend_reg_t reg) { return ((b200_reg_ctx *) reg->context)->devices.size(); }
static ggml_backend_dev_t b200_reg_get_device(ggml_backend_reg_t reg, size_t i) {
    b200_reg_ctx * ctx = (b200_reg_ctx *) reg->context;
    GGML_ASSERT(i < ctx->devices.size());
    return &ctx->devices[i];
}
static void * b200_reg_get_proc_address(ggml_backend_reg_t, const char *) { return nullptr; }   // no split buffers / host registration on this path
static const ggml_backend_reg_i b200_reg_iface = {
    /* .get_name         = */ b200_reg_get_name,
    /* .get_device_count = */ b200_reg_device_count,
    /* .get_device       = */ b200_reg_get_device,
    /* .get_proc_address = */ b200_reg_get_proc_address,
};

extern "C" {

ggml_backend_reg_t ggml_backend_b200_reg(void) {
    static ggml_backend_reg reg;
    static std::once_flag once;
    std::call_once(once, [] {
        b200_reg_ctx * ctx = new b200_reg_ctx();
        int n = pb200_device_count();
        if (n > B200_MAX_DEVICES) n = B200_MAX_DEVICES;
        ctx->devices.resize(n);
        ctx->bufts.resize(n);
        reg.iface = b200_reg_iface;
        reg.context = ctx;
        for (int i = 0; i < n; i++) {
            cudaDeviceProp prop;
            std::string desc = "CUDA device";
            if (cudaGetDeviceProperties(&prop, i) == cudaSuccess) desc = prop.name;
            ctx->devices[i].iface = b200_device_iface;
            ctx->devices[i].reg = &reg;
            ctx->devices[i].context = new b200_device_ctx{i, "B200_" + std::to_string(i), desc};
            ctx->bufts[i].iface = b200_buft_iface;
            ctx->bufts[i].device = &ctx->devices[i];
            ctx->bufts[i].context = nullptr;
        }
    });
    return &reg;
}

ggml_backend_buffer_type_t ggml_backend_b200_buffer_type(int device) {
    b200_reg_ctx * ctx = (b200_reg_ctx *) ggml_backend_b200_reg()->context;
    if (device < 0 || device >= (int) ctx->bufts.size()) return nullptr;
    return &ctx->bufts[device];
}

ggml_backend_t ggml_backend_b200_init(int device) {
    b200_reg_ctx * rctx = (b200_reg_ctx *) ggml_backend_b200_reg()->context;
    if (device < 0 || device >= (int) rctx->devices.size()) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    b200_backend_ctx * ctx = new b200_backend_ctx();
    ctx->device = device;
    ctx->name = "B200_" + std::to_string(device);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return nullptr; }
    return new ggml_backend{b200_guid(), b200_backend_iface, &rctx->devices[device], ctx};
}

int ggml_backend_is_b200(ggml_backend_t backend) { return backend != nullptr && ggml_guid_matches(backend->guid, b200_guid()); }
unsigned long long ggml_backend_b200_nodes_computed(void) { return g_nodes.load(); }

}  // extern "C"

// loading the plugin registers it with the host's registry (ggml_backend_register, ggml-backend-impl.h:220)
__attribute__((constructor)) static void ggml_b200_autoregister() {
    if (getenv("GGML_B200_NO_AUTOREG")) return;
    // only when the host process really carries a ggml registry (LD_PRELOAD also reaches unrelated helper processes)
    typedef void (*register_fn)(ggml_backend_reg_t);
    register_fn reg_fn = (register_fn) dlsym(RTLD_DEFAULT, "ggml_backend_register");
    if (!reg_fn) return;
    ggml_backend_reg_t reg = ggml_backend_b200_reg();
    if (b200_reg_device_count(reg) > 0) reg_fn(reg);
}
