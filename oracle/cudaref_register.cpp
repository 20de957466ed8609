// oracle/cudaref_register.cpp — test / measurement infrastructure (never linked or loaded by the product).
// Loading oracle/_ref/cuda/libggml-cuda-ref.so (the reference's unmodified ggml-cuda backend, Makefile.cudaref) registers its
// "CUDA" backend with the host process's ggml registry: ggml_backend_register (ggml/src/ggml-backend-impl.h:220) fed with
// ggml_backend_cuda_reg() (ggml/include/ggml-cuda.h:43) — what ggml-backend.cpp:549-552 does when built with -DGGML_USE_CUDA.
#include "ggml-backend-impl.h"
#include "ggml-cuda.h"

__attribute__((constructor)) static void cudaref_autoregister() {
    ggml_backend_register(ggml_backend_cuda_reg());
}
