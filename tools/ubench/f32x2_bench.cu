// Micro-benchmark: issue rate / exactness of the sm_100 packed fp32 instructions (FFMA2 / FADD2) against scalar
// FFMA / FADD.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float a[8], b = seed, c = seed * 0.5f;
    for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
    u64 A[8];
    for (int i = 0; i < 8; ++i) A[i] = pk(a[i], a[i] + 1.0f);
    const u64 B = pk(b, b), C = pk(c, c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = __fmaf_rn(a[i], b, c);                 // scalar FFMA: 1 op/lane
                if (MODE == 1) A[i] = fma2(A[i], B, C);                      // FFMA2: 2 ops/lane
                if (MODE == 2) a[i] = __fadd_rn(a[i], c);                    // scalar FADD
                if (MODE == 3) A[i] = add2(A[i], C);                         // FADD2
                if (MODE == 4) { a[i] = __fmaf_rn(a[i], b, c); A[i] = fma2(A[i], B, C); }   // mix
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) { float x, y; upk(A[i], x, y); s += a[i] + x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// exactness: is mul.rn.f32x2 followed by add.rn.f32x2 contracted into an FFMA2 by ptxas?
__global__ void exact(const float* x, const float* y, const float* z, uint32_t* res, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float unf = __fadd_rn(__fmul_rn(x[i], y[i]), z[i]);
    const float fus = __fmaf_rn(x[i], y[i], z[i]);
    float p0, p1, q0, q1;
    upk(add2(mul2(pk(x[i], x[i]), pk(y[i], y[i])), pk(z[i], z[i])), p0, p1);                   // mul2 + add2
    upk(add2(fma2(pk(x[i], x[i]), pk(y[i], y[i]), pk(-0.0f, -0.0f)), pk(z[i], z[i])), q0, q1);  // fma2(-0) + add2
    uint32_t r = 0;
    if (__float_as_uint(p0) == __float_as_uint(unf)) r |= 1;
    if (__float_as_uint(p0) == __float_as_uint(fus)) r |= 2;
    if (__float_as_uint(q0) == __float_as_uint(unf)) r |= 4;
    if (__float_as_uint(q0) == __float_as_uint(fus)) r |= 8;
    if (__float_as_uint(unf) != __float_as_uint(fus)) r |= 16;
    res[i] = r;
}

template <int MODE> double run(const char* name, int opsPerInst) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 10, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = double(148) * 8 * 8 * iters * 64 * (MODE == 4 ? 2 : 1);   // warp instructions
    printf("%-14s %8.3f ms  %7.1f G warp-inst/s  %7.2f Tlane-op/s\n", name, ms, inst / ms * 1e-6, inst * 32 * opsPerInst / ms * 1e-9);
    cudaFree(out);
    return ms;
}

int main() {
    run<0>("FFMA", 1); run<1>("FFMA2", 2); run<2>("FADD", 1); run<3>("FADD2", 2); run<4>("FFMA+FFMA2", 1);
    const int n = 1 << 20;
    float *x, *y, *z; uint32_t* r;
    cudaMallocManaged(&x, n * 4); cudaMallocManaged(&y, n * 4); cudaMallocManaged(&z, n * 4); cudaMallocManaged(&r, n * 4);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xFFFF) / 65536.0f + 0.37f; };
    for (int i = 0; i < n; ++i) { x[i] = rnd(); y[i] = rnd(); z[i] = -x[i] * y[i] * (1.0f + 1e-3f * rnd()); }
    exact<<<n / 256, 256>>>(x, y, z, r, n);
    cudaDeviceSynchronize();
    long c[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) { if (!(r[i] & 16)) continue; c[4]++; for (int b = 0; b < 4; ++b) c[b] += (r[i] >> b) & 1; }
    printf("cases where fused != unfused: %ld; mul2+add2 == unfused %ld, == fused %ld; fma2(-0)+add2 == unfused %ld, == fused %ld\n",
           c[4], c[0], c[1], c[2], c[3]);
    return 0;
}
