// k7_blend.cuh -- frame / patch blending: JXLCodestreamDecoder.blendAdd / blendMult / blendBlend /
// blendMulAdd (J/JXLCodestreamDecoder.java:285-413), one rectangle of one channel per launch.  Elementwise and
// memory-bound; float expressions keep the Java's operand order, uncontracted.
//   a = the buffer the Java passes as `frame` (indexed at frameOffset), b = the one it passes as `ref` (at refOffset),
//   fa / ra = frameAlpha at frameOffset / refAlpha at refOffset (floats), out = canvas at patchStart.
#pragma once
#include "common.cuh"

struct BlendArgs {
    int mode, is_int, is_alpha, has_extra, clamp, premult;
    int h, w;
    const void *a, *b;
    const float *fa, *ra;
    void *out;
    // row pitches in elements (0 = compact rectangles of width w): the batched call blends inside whole device-resident planes
    long long pa, pb, pfa, pra, pout;
};

__device__ __forceinline__ float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }   // MathHelper.clampAsc

__global__ void k7_blend(BlendArgs A) {
    const long long n = (long long)A.h * A.w;
    const bool pitched = A.pout != 0;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        long long i = e, ia = e, ib = e, ifa = e, ira = e;           // element index in out / a / b / fa / ra
        if (pitched) {
            const long long r = e / A.w, c = e - r * A.w;
            i = r * A.pout + c; ia = r * A.pa + c; ib = r * A.pb + c; ifa = r * A.pfa + c; ira = r * A.pra + c;
        }
        int mode = A.mode;
        if ((mode == 2 || mode == 3) && !A.has_extra) mode = 1;       // no alpha anywhere: both resolve to blendAdd
        if (mode == 1) {
            if (A.is_int) {
                ((int *)A.out)[i] = (int)((unsigned)((const int *)A.b)[ib] + (unsigned)((const int *)A.a)[ia]);
            } else {
                ((float *)A.out)[i] = __fadd_rn(((const float *)A.b)[ib], ((const float *)A.a)[ia]);
            }
            continue;
        }
        const float fs = ((const float *)A.a)[ia], rs = ((const float *)A.b)[ib];
        float r;
        if (mode == 4) {                                               // blendMult
            const float ns = A.clamp ? clamp01(fs) : fs;
            r = __fmul_rn(ns, rs);
        } else if (mode == 2) {                                        // blendBlend
            const float old_alpha = A.is_alpha ? rs : A.ra[ira];
            float new_alpha = A.is_alpha ? fs : A.fa[ifa];
            if (A.clamp) new_alpha = clamp01(new_alpha);
            if (A.is_alpha) {
                r = __fadd_rn(old_alpha, __fmul_rn(new_alpha, __fsub_rn(1.0f, old_alpha)));
            } else if (A.premult) {
                r = __fadd_rn(fs, __fmul_rn(rs, __fsub_rn(1.0f, new_alpha)));
            } else {
                const float num = __fadd_rn(__fmul_rn(fs, new_alpha), __fmul_rn(__fmul_rn(rs, old_alpha), __fsub_rn(1.0f, new_alpha)));
                const float den = __fadd_rn(old_alpha, __fmul_rn(new_alpha, __fsub_rn(1.0f, old_alpha)));
                r = __fdiv_rn(num, den);
            }
        } else {                                                       // blendMulAdd (the alpha channel itself is a plain copy, done by the caller)
            float new_alpha = A.fa[ifa];
            if (A.clamp) new_alpha = clamp01(new_alpha);
            r = __fadd_rn(rs, __fmul_rn(new_alpha, fs));
        }
        ((float *)A.out)[i] = r;
    }
}
