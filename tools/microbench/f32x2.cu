#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 c; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ u64 mulz2(u64 a, u64 b, u64 nz) { u64 c; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(c) : "l"(a), "l"(b), "l"(nz)); return c; }
template <int MODE> __global__ void __launch_bounds__(512) k(float *o, int iters, float seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 0) {
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = seed * (tid + i);
        const float inc = seed * 0.001f;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fadd_rn(a[i], inc);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += a[i];
        o[tid] = s;
    } else if (MODE == 1) {
        u64 a[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { float2 v = make_float2(seed * (tid + 2 * i), seed * (tid + 2 * i + 1)); a[i] = *reinterpret_cast<u64 *>(&v); }
        float2 iv = make_float2(seed * 0.001f, seed * 0.001f);
        const u64 inc = *reinterpret_cast<u64 *>(&iv);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = add2(a[i], inc);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { float2 v = *reinterpret_cast<float2 *>(&a[i]); s += v.x + v.y; }
        o[tid] = s;
    } else if (MODE == 2) {   // scalar mul then add, unfused
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = seed * (tid + i);
        const float inc = seed * 0.001f, w = 1.0f + seed * 1e-6f;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fadd_rn(__fmul_rn(a[i], w), inc);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += a[i];
        o[tid] = s;
    } else {   // packed mul (fma with -0) then packed add
        u64 a[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { float2 v = make_float2(seed * (tid + 2 * i), seed * (tid + 2 * i + 1)); a[i] = *reinterpret_cast<u64 *>(&v); }
        float2 iv = make_float2(seed * 0.001f, seed * 0.001f), wv = make_float2(1.0f + seed * 1e-6f, 1.0f + seed * 1e-6f), nzv = make_float2(-0.0f, -0.0f);
        const u64 inc = *reinterpret_cast<u64 *>(&iv), w = *reinterpret_cast<u64 *>(&wv), nz = *reinterpret_cast<u64 *>(&nzv);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = add2(mulz2(a[i], w, nz), inc);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { float2 v = *reinterpret_cast<float2 *>(&a[i]); s += v.x + v.y; }
        o[tid] = s;
    }
}
template <int MODE> float run(float *o, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148 * 2, 512>>>(o, iters, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<148 * 2, 512>>>(o, iters, 1.0f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    float *o; cudaMalloc(&o, 148 * 2 * 512 * 4);
    const int iters = 20000;
    const double flops = 148.0 * 2 * 512 * 16.0 * iters;
    float t0 = run<0>(o, iters), t1 = run<1>(o, iters), t2 = run<2>(o, iters), t3 = run<3>(o, iters);
    printf("scalar FADD   %.3f ms  %.2f Tadd/s\n", t0, flops / t0 * 1e-9);
    printf("packed FADD2  %.3f ms  %.2f Tadd/s\n", t1, flops / t1 * 1e-9);
    printf("scalar FMUL+FADD  %.3f ms  %.2f T(mul+add)/s\n", t2, flops / t2 * 1e-9);
    printf("packed FFMA2(-0)+FADD2 %.3f ms  %.2f T(mul+add)/s\n", t3, flops / t3 * 1e-9);
    float h[4]; cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost); printf("%g\n", h[1]);
    return 0;
}
