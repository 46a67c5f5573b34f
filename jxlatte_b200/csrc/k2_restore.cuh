// k2_restore.cuh -- K2: Gaborish, edge-preserving filter (0-3 passes) and XYB -> linear / YCbCr -> RGB.
//
// Replaces (J/ = java/com/traneptora/jxlatte/ in the reference):
//   Frame.performGabConvolution          J/frame/Frame.java:505-542   (3x3, clamp on the padded plane)
//   Frame.performEdgePreservingFilter    :544-636, epfDistance1 :638-655, epfDistance2 :657-669, epfWeight :671-679
//   OpsinInverseMatrix.invertXYB         J/color/OpsinInverseMatrix.java:105-142
//   JXLCodestreamDecoder.performColorTransforms (YCbCr branch)  J/JXLCodestreamDecoder.java:270-282
//
// This file holds the staged version (one kernel per stage, planes round-trip through L2/HBM): simple, used as the
// reference point for the fused kernel in k2_fused.cuh and for frames whose shape the fused kernel does not take.
#pragma once
#include "common.cuh"

// MathHelper.mirrorCoordinate (J/util/MathHelper.java:323-329) on a slab: rows outside [0, rows) exist in memory when
// a neighbour rank supplied them (has_top / has_bottom); otherwise they mirror at the true frame edge.
__device__ __forceinline__ int mirror_row(int r, int rows, int has_top, int has_bottom) {
    if (r < 0 && !has_top) r = -r - 1;
    if (r >= rows && !has_bottom) r = 2 * rows - 1 - r;
    return r;
}
__device__ __forceinline__ int mirror_col(int x, int W) {
    if (x < 0) x = -x - 1;
    if (x >= W) x = 2 * W - 1 - x;
    return x;
}

// inverseSigma map, Frame.java:552-572.  One float per 8x8 block; block rows [-1, rows/8] when halos exist.
__global__ void k2_sigma(const int32_t *__restrict__ hf_mul, const int32_t *__restrict__ sharp, int wb, int br0, int br1,
                         float gscale, const float *__restrict__ lut8, float *__restrict__ inv_sigma, int *__restrict__ err) {
    const int n = (br1 - br0) * wb;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int o = br0 * wb + i;          // may be negative: hf_mul points at the slab's first own block row
        const int s = sharp[o];
        if (s < 0 || s > 7) { *err = 1; inv_sigma[o] = __int_as_float(0x7fc00000); continue; }
        const float sigma = __fdiv_rn(__fmul_rn(gscale, lut8[s]), (float)hf_mul[o]);
        inv_sigma[o] = __fdiv_rn(1.0f, sigma);
    }
}

// Modular-encoded frames have one sigma for the whole frame: invModularSigma = 1f / epfSigmaForModular (Frame.java:573-575, 604-607)
__global__ void k2_sigma_fill(float v, long long first, long long n, float *__restrict__ inv_sigma) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) inv_sigma[first + i] = v;
}

// Rows [r0, r1) of the slab (may extend into halo rows).
__global__ void k2_gab(K2Params P, const float *const in0, const float *const in1, const float *const in2,
                       float *out0, float *out1, float *out2, long long in_pitch, long long out_pitch, int r0, int r1) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = r0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.W || r >= r1) return;
    const int rn = mirror_row(r - 1, P.rows, P.has_top, P.has_bottom), rs = mirror_row(r + 1, P.rows, P.has_top, P.has_bottom);
    const int xw = x == 0 ? 0 : x - 1, xe = x + 1 == P.W ? P.W - 1 : x + 1;
    const float *in[3] = {in0, in1, in2};
    float *out[3] = {out0, out1, out2};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *R = in[c] + (long long)r * in_pitch, *N = in[c] + (long long)rn * in_pitch, *S = in[c] + (long long)rs * in_pitch;
        // Frame.java:535-537, operand order kept, no FMA contraction -> bit-identical to the Java
        const float adj = __fadd_rn(__fadd_rn(__fadd_rn(R[xw], R[xe]), N[x]), S[x]);
        const float diag = __fadd_rn(__fadd_rn(__fadd_rn(N[xw], N[xe]), S[xw]), S[xe]);
        out[c][(long long)r * out_pitch + x] = __fadd_rn(__fadd_rn(__fmul_rn(P.gab_base[c], R[x]), __fmul_rn(P.gab_adj[c], adj)),
                                                         __fmul_rn(P.gab_diag[c], diag));
    }
}

// One EPF pass over rows [r0, r1).  PASS 0: 13-point double cross with plus-shaped SADs; 1: 5-point cross with plus
// SADs; 2: 5-point cross with point differences.
template <int PASS> __global__ void k2_epf(K2Params P, const float *in0, const float *in1, const float *in2,
                                           float *out0, float *out1, float *out2, long long pitch,
                                           const float *__restrict__ inv_sigma, int r0, int r1) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = r0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.W || r >= r1) return;
    const float *in[3] = {in0, in1, in2};
    float *out[3] = {out0, out1, out2};
    const long long o = (long long)r * pitch + x;
    const float s = inv_sigma[(r >> 3) * P.wb + (x >> 3)];
    if (s != s || s > (1.0f / 0.3f)) {
#pragma unroll
        for (int c = 0; c < 3; c++) out[c][o] = in[c][o];
        return;
    }
    constexpr int NC = PASS == 0 ? 13 : 5;
    const int cy[13] = {0, 0, 0, -1, 1, -1, 1, 1, -1, 0, 0, 2, -2};
    const int cx[13] = {0, -1, 1, 0, 0, 1, 1, -1, -1, -2, 2, 0, 0};
    const int py[5] = {0, 0, 0, -1, 1};
    const int px[5] = {0, -1, 1, 0, 0};
    const int my = r & 7, mx = x & 7;
    const bool border = my == 0 || my == 7 || mx == 0 || mx == 7;
    float sumW = 0.0f, sum[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < NC; i++) {
        float dist = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (PASS == 2) {
                const int dr = mirror_row(r + cy[i], P.rows, P.has_top, P.has_bottom), dc = mirror_col(x + cx[i], P.W);
                dist = __fadd_rn(dist, __fmul_rn(fabsf(__fsub_rn(in[c][o], in[c][(long long)dr * pitch + dc])), P.ch_scale[c]));
            } else {
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    const int pr = mirror_row(r + py[j], P.rows, P.has_top, P.has_bottom), pc = mirror_col(x + px[j], P.W);
                    const int dr = mirror_row(r + cy[i] + py[j], P.rows, P.has_top, P.has_bottom), dc = mirror_col(x + cx[i] + px[j], P.W);
                    dist = __fadd_rn(dist, __fmul_rn(fabsf(__fsub_rn(in[c][(long long)pr * pitch + pc], in[c][(long long)dr * pitch + dc])), P.ch_scale[c]));
                }
            }
        }
        if (border) dist = __fmul_rn(dist, P.border_mul);
        float w = __fsub_rn(1.0f, __fmul_rn(__fmul_rn(dist, P.sigma_scale[PASS]), s));   // epfWeight :677
        w = w < 0.0f ? 0.0f : w;
        sumW = __fadd_rn(sumW, w);
        const int nr = mirror_row(r + cy[i], P.rows, P.has_top, P.has_bottom), nc = mirror_col(x + cx[i], P.W);
#pragma unroll
        for (int c = 0; c < 3; c++) sum[c] = __fadd_rn(sum[c], __fmul_rn(in[c][(long long)nr * pitch + nc], w));
    }
#pragma unroll
    for (int c = 0; c < 3; c++) out[c][o] = __fdiv_rn(sum[c], sumW);
}

__device__ __forceinline__ void color_px(const K2Params &P, float &a, float &b, float &c) {
    if (P.color_mode & 1) {
        // OpsinInverseMatrix.invertXYB :128-138 with the Java's operation order and no FMA contraction: the matrix rows
        // cancel to ~1e-3 of their terms on saturated colours, so a fused multiply-add here is visible at 16 bits
        const float gl = __fadd_rn(__fadd_rn(b, a), P.cob[0]), gm = __fadd_rn(__fsub_rn(b, a), P.cob[1]), gs = __fadd_rn(c, P.cob[2]);
        const float ml = __fadd_rn(__fmul_rn(__fmul_rn(gl, gl), gl), P.ob[0]);
        const float mm = __fadd_rn(__fmul_rn(__fmul_rn(gm, gm), gm), P.ob[1]);
        const float ms = __fadd_rn(__fmul_rn(__fmul_rn(gs, gs), gs), P.ob[2]);
        a = __fadd_rn(__fadd_rn(__fmul_rn(P.m[0], ml), __fmul_rn(P.m[1], mm)), __fmul_rn(P.m[2], ms));
        b = __fadd_rn(__fadd_rn(__fmul_rn(P.m[3], ml), __fmul_rn(P.m[4], mm)), __fmul_rn(P.m[5], ms));
        c = __fadd_rn(__fadd_rn(__fmul_rn(P.m[6], ml), __fmul_rn(P.m[7], mm)), __fmul_rn(P.m[8], ms));
    }
    if (P.color_mode & 2) {
        const float cb = a, yh = __fadd_rn(b, 0.50196078431372549019f), cr = c;
        a = __fadd_rn(yh, __fmul_rn(1.402f, cr));
        b = __fsub_rn(__fsub_rn(yh, __fmul_rn(0.34413628620102214650f, cb)), __fmul_rn(0.71413628620102214650f, cr));
        c = __fadd_rn(yh, __fmul_rn(1.772f, cb));
    }
}

__global__ void k2_color(K2Params P, const float *in0, const float *in1, const float *in2, long long in_pitch,
                         float *out0, float *out1, float *out2, long long out_pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.W || r >= P.rows) return;
    float a = in0[(long long)r * in_pitch + x], b = in1[(long long)r * in_pitch + x], c = in2[(long long)r * in_pitch + x];
    color_px(P, a, b, c);
    out0[(long long)r * out_pitch + x] = a;
    out1[(long long)r * out_pitch + x] = b;
    out2[(long long)r * out_pitch + x] = c;
}
