// common.cuh -- shared definitions of libjxlb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/jxlb200.h"

// TransformType table (J/frame/vardct/TransformType.java:10-36), device + host copies
struct TTInfo {
    uint8_t param;    // parameterIndex
    uint8_t method;   // transformMethod
    uint8_t bh, bw;   // dctSelectHeight / Width (8x8 cells)
    uint8_t flip;     // TransformType.flip() :129-131
    uint8_t pad[3];
};
static const TTInfo h_tt[27] = {
    {0, 0, 1, 1, 1}, {1, 3, 1, 1, 0}, {2, 1, 1, 1, 0}, {3, 2, 1, 1, 0}, {4, 0, 2, 2, 1}, {5, 0, 4, 4, 1},
    {6, 0, 2, 1, 1}, {6, 0, 1, 2, 0}, {7, 0, 4, 1, 1}, {7, 0, 1, 4, 0}, {8, 0, 4, 2, 1}, {8, 0, 2, 4, 0},
    {9, 5, 1, 1, 0}, {9, 4, 1, 1, 0}, {10, 6, 1, 1, 0}, {10, 6, 1, 1, 0}, {10, 6, 1, 1, 0}, {10, 6, 1, 1, 0},
    {11, 0, 8, 8, 1}, {12, 0, 8, 4, 1}, {12, 0, 4, 8, 0}, {13, 0, 16, 16, 1}, {14, 0, 16, 8, 1}, {14, 0, 8, 16, 0},
    {15, 0, 32, 32, 1}, {16, 0, 32, 16, 1}, {16, 0, 16, 32, 0},
};

// Work lists built on the device from dct_select / block_origin (k0_lists.cu)
#define N_SMALL 10
#define N_MED 8
#define N_BIGC 4      // line-length classes 32, 64, 128, 256
static const int h_small_types[N_SMALL] = {0, 1, 2, 3, 12, 13, 14, 15, 16, 17};
static const int h_med_types[N_MED] = {4, 5, 6, 7, 8, 9, 10, 11};
#define SMALL_BATCH 32          // 8x8 varblocks per CTA iteration
#define MED_COEFFS 2048         // coefficients per channel per CTA iteration of the medium kernel

struct Sched {
    int cnt[27];                // varblocks per type
    int start[27];              // first slot of the type's segment in items[]
    int cursor[27];             // scatter cursors
    int small_cum[N_SMALL + 1]; // cumulative batches over h_small_types
    int med_cum[N_MED + 1];
    int big_cum[2][N_BIGC][4];  // [pass][class]: cumulative strip-items over that class's (<= 3) types
    int ticket[2][N_BIGC];      // next work item of each k1_big launch (items differ in cost: handed out dynamically)
};

// Per-frame arguments of the stage-1 kernels
struct K1Params {
    const int32_t *q[3];
    const float *lf[3];
    float *out[3];
    long long out_pitch;
    const uint8_t *dct_select;
    const int32_t *hf_mul;
    const int32_t *xfy, *bfy;
    const int32_t *cfl_gate;    // per 64x64 tile: raster index of the varblock origin covering the tile's corner cell
    const float *wexp;          // QM weights re-laid-out per TransformType in storage orientation
    const float *cos_big;       // MathHelper.cosineLut levels 6..8 (lengths 64, 128, 256), [n-1][k]
    int W, H, wb, hb, tw;
    float sf[3];                // scaleFactor[c]  (HFCoefficients.java:270-275)
    float qb[3];                // quantBias
    float qbn;                  // quantBiasNumerator
    float base_x, base_b, color_factor;
    int vec;                    // coefficient and output planes are 16-byte aligned with pitches that are multiples of 4: 128-bit row accesses
};

struct DevTables {
    int wexp_off[27];           // float offset of (type, channel 0) in wexp; channel c at + c * H * W
};

// Stage-2 arguments
struct K2Params {
    const float *in[3];
    long long in_pitch;
    float *out[3];
    long long out_pitch;
    const int32_t *hf_mul, *sharpness;  // point at the slab's first own block row
    int W, rows;                // slab size
    int y0, frame_h;            // slab origin and whole-frame height
    int has_top, has_bottom;
    int wb;
    int gab, iters, color_mode;
    float gab_base[3], gab_adj[3], gab_diag[3];
    float gscale;               // 65536 / globalScale
    float sharp_lut[8];
    float ch_scale[3];
    float sigma_scale[3];       // per pass 0,1,2 (already times stepMultiplier)
    float border_mul;
    float m[9], ob[3], cob[3];  // scaled opsin matrix, opsin bias, -cbrt(bias)
};

#define CUDA_TRY(ctx, expr)                                                         \
    do {                                                                            \
        cudaError_t e__ = (expr);                                                   \
        if (e__ != cudaSuccess) return (ctx)->fail(JXLB200_E_CUDA, #expr, e__);     \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
