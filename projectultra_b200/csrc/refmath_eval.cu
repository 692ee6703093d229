// projectultra_b200/csrc/refmath_eval.cu — evaluates the libm restatements of ref_math.cuh on the host or on the
// device over arrays, so tests can pin them against the host libm (CPU) and check device == host (GPU).
#include "pu_internal.h"
#include "ref_math.cuh"

namespace pu {

PU_RM float refmath_apply(int op, float a, float b) {
    switch (op) {
        case 0: return refmath::atan2f_ref(a, b);
        case 1: return refmath::sinf_ref(a);
        case 2: return refmath::cosf_ref(a);
        case 3: return refmath::hypotf_ref(a, b);
        default: return refmath::atanf_ref(a);
    }
}

__global__ void refmath_kernel(int op, const float* a, const float* b, float* out, size_t n) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = refmath_apply(op, a[i], b ? b[i] : 0.0f);
}

}  // namespace pu

extern "C" pu_status pu_refmath_eval(pu_ctx* ctx, int op, const float* a, const float* b, float* out, size_t n) {
    PU_REQUIRE(a && out, "pu_refmath_eval: NULL argument");
    if (!ctx) {   // host evaluation
        for (size_t i = 0; i < n; ++i) out[i] = pu::refmath_apply(op, a[i], b ? b[i] : 0.0f);
        return PU_OK;
    }
    PU_CUDA_TRY(cudaSetDevice(ctx->device));
    float *da = nullptr, *db = nullptr, *dout = nullptr;
    PU_CUDA_TRY(cudaMalloc(&da, n * sizeof(float)));
    PU_CUDA_TRY(cudaMalloc(&db, n * sizeof(float)));
    PU_CUDA_TRY(cudaMalloc(&dout, n * sizeof(float)));
    PU_CUDA_TRY(cudaMemcpy(da, a, n * sizeof(float), cudaMemcpyHostToDevice));
    if (b) PU_CUDA_TRY(cudaMemcpy(db, b, n * sizeof(float), cudaMemcpyHostToDevice));
    PU_CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));   // pageable H2D: wait for the DMA before the non-blocking stream's kernel
    pu::refmath_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(op, da, b ? db : nullptr, dout, n);
    ctx->launches.fetch_add(1);
    PU_CUDA_TRY(cudaGetLastError());
    PU_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    PU_CUDA_TRY(cudaMemcpy(out, dout, n * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return PU_OK;
}
