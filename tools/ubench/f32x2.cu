// Microbenchmark: issue cost of packed FP32x2 (FADD2/FMUL2/FFMA2, sm_100) against scalar FADD/FMUL/FFMA.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE> __global__ void k(float *out, float a, float b) {
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    unsigned xi[8];
#pragma unroll
    for (int i = 0; i < 8; i++) xi[i] = threadIdx.x * 7 + i;
    const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.9999f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { v[i].x = __fadd_rn(v[i].x, A.x); v[i].y = __fadd_rn(v[i].y, A.y); }
            if (MODE == 1) { v[i] = __fadd2_rn(v[i], A); }
            if (MODE == 2) { v[i].x = __fmul_rn(v[i].x, B.x); v[i].y = __fmul_rn(v[i].y, B.y); }
            if (MODE == 3) { v[i] = __fmul2_rn(v[i], B); }
            if (MODE == 4) { v[i].x = __fmaf_rn(v[i].x, B.x, A.x); v[i].y = __fmaf_rn(v[i].y, B.y, A.y); }
            if (MODE == 5) { v[i] = __ffma2_rn(v[i], B, A); }
            if (MODE == 6) { v[i].x = __fadd_rn(__fmul_rn(v[i].x, B.x), A.x); v[i].y = __fadd_rn(__fmul_rn(v[i].y, B.y), A.y); }
            if (MODE == 7) { v[i] = __fadd2_rn(__fmul2_rn(v[i], B), A); }
            // issue-slot test: the same FADD work beside 8 independent ALU instructions (LOP3 / IADD3) per 16 float adds
            if (MODE == 8) { v[i].x = __fadd_rn(v[i].x, A.x); v[i].y = __fadd_rn(v[i].y, A.y); xi[i] = (xi[i] ^ it) + (xi[i] >> 3); }
            if (MODE == 9) { v[i] = __fadd2_rn(v[i], A); xi[i] = (xi[i] ^ it) + (xi[i] >> 3); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i].x + v[i].y + (MODE >= 8 ? (float)xi[i] : 0.0f);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, float *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 512>>>(d, 1e-3f, 1.0000001f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 512>>>(d, 1e-3f, 1.0000001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = 148.0 * 4 * 512 * ITERS * 16;   // float results per launch (mode 6/7: x2 operations)
    printf("%-28s %8.3f ms  %7.2f T float-results/s\n", name, ms, lane_ops / ms / 1e9);
}
int main() {
    float *d; cudaMalloc(&d, 148 * 4 * 512 * 4);
    run<0>("FADD scalar", d); run<1>("FADD2 packed", d); run<2>("FMUL scalar", d); run<3>("FMUL2 packed", d);
    run<4>("FFMA scalar", d); run<5>("FFMA2 packed", d); run<6>("FMUL+FADD scalar", d); run<7>("FMUL2+FADD2 packed", d);
    run<8>("FADD scalar + ALU mix", d); run<9>("FADD2 packed + ALU mix", d);
    return 0;
}
