// k1_idct.cuh -- K1: HF dequantisation + chroma-from-luma + LLF-from-LF + inverse transform, fused per varblock.
//
// Replaces (J/ = java/com/traneptora/jxlatte/ in the reference):
//   HFCoefficients.dequantizeHFCoefficients  J/frame/vardct/HFCoefficients.java:267-319
//   HFCoefficients.chromaFromLuma            :146-192
//   HFCoefficients.finalizeLLF               :194-229  (+ MathHelper.forwardDCT2D, J/util/MathHelper.java:124-136)
//   PassGroup.invertVarDCT                   J/frame/group/PassGroup.java:170-331 (all 27 TransformTypes)
//
// Three kernel families, each persistent over the K0 work lists:
//   k1_small   8x8-class (DCT8, Hornuss, DCT2, DCT4, DCT4x8, DCT8x4, AFV0-3): one thread per varblock-channel, 64
//              coefficients in registers, staged through shared memory so global traffic stays row-coalesced.
//   k1_medium  16/32-class (8 shapes): tile in shared memory, one thread per line, IDCT in registers.
//   k1_big     any side >= 64: two passes (columns, then rows) over 32-line strips, lane = line, each warp owning 32
//              outputs of 32 lines; the intermediate lives in the output plane (L2-resident between the passes).
// Every float operation is an explicit round-to-nearest mul/add/sub/div in the Java's order (Java never contracts to
// FMA), so the XYB planes this stage writes are bit-identical to the reference's (see transforms.cuh for why).
#pragma once
#include "common.cuh"
#include "k0_lists.cuh"
#include "transforms.cuh"

__constant__ float c_afv[256];        // PassGroup.AFV_BASIS (J/frame/group/PassGroup.java:19-58)
__constant__ float c_cos[1302];       // MathHelper.cosineLut levels 1..5 (lengths 2..32), [n][k], J/util/MathHelper.java:17-30
__constant__ int c_cos_off[6];        // offset of level l
__constant__ float c_llf_scale[32];   // LLFScale.SCALE_F (J/frame/vardct/LLFScale.java:7-23)
__constant__ DevTables c_tab;

struct CosLut {   // packed MathHelper.cosineLut, constant-bank operand once the index is a compile-time constant
    __device__ __forceinline__ float operator()(int i) const { return c_cos[i]; }
};

__device__ __forceinline__ int ilog2_pow2(int v) { return 31 - __clz(v); }

// HFCoefficients.dequantizeHFCoefficients :310-314, one coefficient (not in the LLF corner)
__device__ __forceinline__ float dequant_one(int coeff, float qb, float qbn, float sfc, float w) {
    float quant;
    if (coeff > -2 && coeff < 2) quant = coeff == 0 ? 0.0f : (coeff > 0 ? qb : -qb);
    else quant = __fsub_rn((float)coeff, __fdiv_rn(qbn, (float)coeff));
    return __fmul_rn(__fmul_rn(quant, sfc), w);
}

// Per-varblock constants shared by the three families
struct VB {
    int by, bx, type, H, W;     // origin in cells, TransformType, pixel size
    int dsH, dsW;               // LLF corner size
    float sfc[3];               // scaleFactor[c] / hfMultiplier  (:299)
    int origin;                 // raster index of the origin cell (blockList order key)
    const float *w[3];          // weights in storage orientation
};

__device__ __forceinline__ VB make_vb(const K1Params &P, int item, int type) {
    VB v;
    v.by = item >> 16;
    v.bx = item & 0xffff;
    v.type = type;
    const TTInfo tt = c_tt[type];
    v.dsH = tt.bh; v.dsW = tt.bw;
    v.H = tt.bh * 8; v.W = tt.bw * 8;
    v.origin = v.by * P.wb + v.bx;
    const float hm = (float)P.hf_mul[v.origin];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        v.sfc[c] = __fdiv_rn(P.sf[c], hm);
        v.w[c] = P.wexp + c_tab.wexp_off[type] + c * v.H * v.W;
    }
    return v;
}

// dequantise + CfL the coefficient at local (ly, lx) of varblock v for the three channels (X, Y, B order).
// Caller guarantees (ly, lx) is outside the LLF corner.
__device__ __forceinline__ void dequant3(const K1Params &P, const VB &v, int ly, int lx, float &X, float &Y, float &B) {
    const int py = v.by * 8 + ly, px = v.bx * 8 + lx;
    const size_t gi = (size_t)py * P.W + px;
    const int wi = ly * v.W + lx;
    const int q0 = __ldg(P.q[0] + gi), q1 = __ldg(P.q[1] + gi), q2 = __ldg(P.q[2] + gi);
    X = dequant_one(q0, P.qb[0], P.qbn, v.sfc[0], __ldg(v.w[0] + wi));
    Y = dequant_one(q1, P.qb[1], P.qbn, v.sfc[1], __ldg(v.w[1] + wi));
    B = dequant_one(q2, P.qb[2], P.qbn, v.sfc[2], __ldg(v.w[2] + wi));
    // chromaFromLuma :172-188.  A tile's factor is only "cached" once the varblock covering the tile's top-left
    // pixel has been visited (raster order of origins); before that the Java reads 0.0f from its fresh arrays.
    const int tile = (py >> 6) * P.tw + (px >> 6);
    float kX = 0.0f, kB = 0.0f;
    if (__ldg(P.cfl_gate + tile) <= v.origin) {
        kX = __fadd_rn(P.base_x, __fdiv_rn((float)__ldg(P.xfy + tile), P.color_factor));
        kB = __fadd_rn(P.base_b, __fdiv_rn((float)__ldg(P.bfy + tile), P.color_factor));
    }
    X = __fadd_rn(X, __fmul_rn(kX, Y));
    B = __fadd_rn(B, __fmul_rn(kB, Y));
}

// One LLF coefficient (ky, kx) of HFCoefficients.finalizeLLF :218-226 for a dsH x dsW corner with dsH, dsW <= 4:
// forwardDCT2D = rows, then columns, each output * 1/N, same accumulation order; then * llfScale.
__device__ __forceinline__ float llf_small(const float *lf, int wb, int dsH, int dsW, int ky, int kx) {
    const int lw = ilog2_pow2(dsW), lh = ilog2_pow2(dsH);
    const float invW = 1.0f / (float)dsW, invH = 1.0f / (float)dsH;
    float col[4];
    for (int y = 0; y < dsH; y++) {
        const float *src = lf + (size_t)y * wb;
        float d2;
        if (kx == 0) {
            d2 = src[0];
            for (int n = 1; n < dsW; n++) d2 = __fadd_rn(d2, src[n]);
        } else {
            const float *lut = c_cos + c_cos_off[lw] + (kx - 1) * dsW;
            d2 = __fmul_rn(src[0], lut[0]);
            for (int n = 1; n < dsW; n++) d2 = __fadd_rn(d2, __fmul_rn(src[n], lut[n]));
        }
        col[y] = __fmul_rn(d2, invW);
    }
    float d2;
    if (ky == 0) {
        d2 = col[0];
        for (int n = 1; n < dsH; n++) d2 = __fadd_rn(d2, col[n]);
    } else {
        const float *lut = c_cos + c_cos_off[lh] + (ky - 1) * dsH;
        d2 = __fmul_rn(col[0], lut[0]);
        for (int n = 1; n < dsH; n++) d2 = __fadd_rn(d2, __fmul_rn(col[n], lut[n]));
    }
    d2 = __fmul_rn(d2, invH);
    // TransformType ctor :158-165: llfScale[y][x] = SCALE_F[y << (5 - yll)] * SCALE_F[x << (5 - xll)]
    const float sc = __fmul_rn(c_llf_scale[ky << (5 - lh)], c_llf_scale[kx << (5 - lw)]);
    return __fmul_rn(d2, sc);
}

// ------------------------------------------------------------------------------------------------------------
// k1_small: 8x8-class
// ------------------------------------------------------------------------------------------------------------
#define SMALL_PITCH 100   // 96 varblock-channels + 4: bank = (4 * elem + slot) mod 32, conflict-free for both phases

struct SmemOut {
    float *base;
    __device__ __forceinline__ void operator()(int y, int x, float val) const { base[(y * 8 + x) * SMALL_PITCH] = val; }
};
struct AfvBasis {
    __device__ __forceinline__ float operator()(int j, int i) const { return c_afv[j * 16 + i]; }
};

__global__ void __launch_bounds__(128) k1_small(K1Params P, const Sched *__restrict__ S, const int *__restrict__ items) {
    __shared__ float tile[64 * SMALL_PITCH];
    const int total = S->small_cum[N_SMALL];
    const int tid = threadIdx.x;
    for (int b = blockIdx.x; b < total; b += gridDim.x) {
        int ti = 0;
        while (b >= S->small_cum[ti + 1]) ti++;
        const int type = c_small_types[ti];
        const int first = (b - S->small_cum[ti]) * SMALL_BATCH;
        const int nvb = min(SMALL_BATCH, S->cnt[type] - first);
        const int *it = items + S->start[type] + first;

        // load + dequant + CfL: lane = (varblock, x), loop over y -> 32-byte row segments of 4 varblocks per request
        const int x = tid & 7;
        for (int pass = 0; pass < 2; pass++) {
            const int vbi = (tid >> 3) + 16 * pass;
            if (vbi < nvb) {
                const VB v = make_vb(P, it[vbi], type);
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    float X, Y, B;
                    if (y == 0 && x == 0) {   // LLF of a 1x1 corner is the LF sample itself (scale 1)
                        X = __ldg(P.lf[0] + v.origin); Y = __ldg(P.lf[1] + v.origin); B = __ldg(P.lf[2] + v.origin);
                    } else {
                        dequant3(P, v, y, x, X, Y, B);
                    }
                    float *d = tile + (y * 8 + x) * SMALL_PITCH + vbi * 3;
                    d[0] = X; d[1] = Y; d[2] = B;
                }
            }
        }
        __syncthreads();
        // transform: one thread per varblock-channel
        if (tid < nvb * 3) {
            float v[64];
#pragma unroll
            for (int e = 0; e < 64; e++) v[e] = tile[e * SMALL_PITCH + tid];
            SmemOut out{tile + tid};
            switch (type) {
            case 0: inv_dct8x8(v, CosLut(), out); break;
            case 1: inv_hornuss(v, out); break;
            case 2: inv_dct2(v, out); break;
            case 3: inv_dct4(v, CosLut(), out); break;
            case 12: inv_dct4x8<false>(v, CosLut(), out); break;
            case 13: inv_dct4x8<true>(v, CosLut(), out); break;
            default: inv_afv(v, (type == 16 || type == 17) ? 1 : 0, (type == 15 || type == 17) ? 1 : 0, AfvBasis(), CosLut(), out); break;
            }
        }
        __syncthreads();
        for (int pass = 0; pass < 2; pass++) {
            const int vbi = (tid >> 3) + 16 * pass;
            if (vbi < nvb) {
                const int item = it[vbi];
                const int by = item >> 16, bx = item & 0xffff;
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    const float *s = tile + (y * 8 + x) * SMALL_PITCH + vbi * 3;
                    const size_t o = (size_t)(by * 8 + y) * P.out_pitch + bx * 8 + x;
                    P.out[0][o] = s[0]; P.out[1][o] = s[1]; P.out[2][o] = s[2];
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// k1_medium: shapes with both sides in {8, 16, 32} except 8x8
// ------------------------------------------------------------------------------------------------------------
template <int N> __device__ __forceinline__ void line_idct(float *p, int stride) {
    float v[N];
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = p[i * stride];
    RefIDCT<N>::run(v, CosLut());
#pragma unroll
    for (int i = 0; i < N; i++) p[i * stride] = v[i];
}

template <int H, int W> __device__ void medium_batch(const K1Params &P, int type, const int *it, int nvb, float *tile, VB *s_vb) {
    constexpr int PITCH = W + 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid < nvb) s_vb[tid] = make_vb(P, it[tid], type);
    __syncthreads();
    // load + dequant + CfL (x fastest -> coalesced rows of W ints)
    for (int i = tid; i < nvb * H * W; i += nt) {
        const int vbi = i / (H * W), r = i % (H * W);
        const int y = r / W, x = r % W;
        const VB &v = s_vb[vbi];
        float X = 0.0f, Y = 0.0f, B = 0.0f;
        if (y < H / 8 && x < W / 8) {
            const int o = v.origin;
            X = llf_small(P.lf[0] + o, P.wb, H / 8, W / 8, y, x);
            Y = llf_small(P.lf[1] + o, P.wb, H / 8, W / 8, y, x);
            B = llf_small(P.lf[2] + o, P.wb, H / 8, W / 8, y, x);
        } else {
            dequant3(P, v, y, x, X, Y, B);
        }
        float *d = tile + ((vbi * 3) * H + y) * PITCH + x;
        d[0] = X; d[H * PITCH] = Y; d[2 * H * PITCH] = B;
    }
    __syncthreads();
    // columns (length H), then rows (length W): MathHelper.inverseDCT2D :110-120
    for (int l = tid; l < nvb * 3 * W; l += nt) line_idct<H>(tile + (l / W) * H * PITCH + (l % W), PITCH);
    __syncthreads();
    for (int l = tid; l < nvb * 3 * H; l += nt) line_idct<W>(tile + l * PITCH, 1);
    __syncthreads();
    for (int i = tid; i < nvb * H * W; i += nt) {
        const int vbi = i / (H * W), r = i % (H * W);
        const int y = r / W, x = r % W;
        const VB &v = s_vb[vbi];
        const size_t o = (size_t)(v.by * 8 + y) * P.out_pitch + v.bx * 8 + x;
        const float *s = tile + ((vbi * 3) * H + y) * PITCH + x;
        P.out[0][o] = s[0]; P.out[1][o] = s[H * PITCH]; P.out[2][o] = s[2 * H * PITCH];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) k1_medium(K1Params P, const Sched *__restrict__ S, const int *__restrict__ items) {
    __shared__ float tile[3 * MED_COEFFS * 9 / 8];
    __shared__ VB s_vb[MED_COEFFS / 128];
    const int total = S->med_cum[N_MED];
    for (int b = blockIdx.x; b < total; b += gridDim.x) {
        int ti = 0;
        while (b >= S->med_cum[ti + 1]) ti++;
        const int type = c_med_types[ti];
        const TTInfo tt = c_tt[type];
        const int per = MED_COEFFS / (tt.bh * tt.bw * 64);
        const int first = (b - S->med_cum[ti]) * per;
        const int nvb = min(per, S->cnt[type] - first);
        const int *it = items + S->start[type] + first;
        switch (type) {
        case 4: medium_batch<16, 16>(P, type, it, nvb, tile, s_vb); break;
        case 5: medium_batch<32, 32>(P, type, it, nvb, tile, s_vb); break;
        case 6: medium_batch<16, 8>(P, type, it, nvb, tile, s_vb); break;
        case 7: medium_batch<8, 16>(P, type, it, nvb, tile, s_vb); break;
        case 8: medium_batch<32, 8>(P, type, it, nvb, tile, s_vb); break;
        case 9: medium_batch<8, 32>(P, type, it, nvb, tile, s_vb); break;
        case 10: medium_batch<32, 16>(P, type, it, nvb, tile, s_vb); break;
        default: medium_batch<16, 32>(P, type, it, nvb, tile, s_vb); break;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k1_big: 32-line strips of varblocks with a side >= 64.  PASS 0 = columns (reads coefficients, writes the plane),
// PASS 1 = rows (reads the plane, writes the plane).  N = line length.
// MathHelper.inverseDCTHorizontal's recurrence, lane = line: a warp owns outputs k0..k0+15 and their mirrors
// N-1-k of 32 lines (32 accumulators per lane), walks n = 1..N-1 reading in[n] from shared memory and the 16 table
// entries lut[n-1][k0..k0+15] as four warp-uniform 128-bit loads, and skips n when in[n] is zero on all 32 lines
// (adding +-0 products changes nothing; the coefficient rows of pass 0 are mostly zero).
// Shared memory: A float[3][N][33] (inputs), T float[3][N][33] (pass 1 output staging), llf scratch float[3][32][33].
// ------------------------------------------------------------------------------------------------------------
#define BIG_PITCH 33
#define BIG_THREADS 256
template <int N, int PASS> struct BigSmem {
    static constexpr int kA = 3 * N * BIG_PITCH;
    static constexpr int kFloats = kA + (PASS == 1 ? kA : 0) + (PASS == 0 ? 3 * 32 * BIG_PITCH : 0);
    static constexpr int kBytes = kFloats * 4;
};
__host__ __device__ constexpr int cos_big_off(int n) { return n == 64 ? 0 : n == 128 ? 63 * 64 : 63 * 64 + 127 * 128; }
#define COS_BIG_FLOATS (63 * 64 + 127 * 128 + 255 * 256)

template <int N, int PASS> __global__ void __launch_bounds__(BIG_THREADS) k1_big(K1Params P, const Sched *__restrict__ S, const int *__restrict__ items, int cls) {
    extern __shared__ float smem[];
    float *A = smem;
    float *T = smem + BigSmem<N, PASS>::kA;     // PASS 1 only
    float *scr = smem + BigSmem<N, PASS>::kA;   // PASS 0 only
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int total = S->big_cum[PASS][cls][3];
    for (int wi = blockIdx.x; wi < total; wi += gridDim.x) {
        int j = 0;
        while (wi >= S->big_cum[PASS][cls][j + 1]) j++;
        const int type = c_big_types[PASS][cls][j];
        const TTInfo tt = c_tt[type];
        const int nstrips = (PASS == 0 ? tt.bw : tt.bh) / 4;
        const int local = wi - S->big_cum[PASS][cls][j];
        const int strip = local % nstrips;
        const VB v = make_vb(P, items[S->start[type] + local / nstrips], type);
        const int Y0 = v.by * 8, X0 = v.bx * 8;

        if (PASS == 0) {
            // rows i = 0..N-1 of 32 columns: coalesced 128-byte reads of the three coefficient planes
            for (int i = warp; i < N; i += nwarps) {
                const int lx = strip * 32 + lane;
                float X = 0.0f, Y = 0.0f, B = 0.0f;
                if (!(i < v.dsH && lx < v.dsW)) dequant3(P, v, i, lx, X, Y, B);
                A[(0 * N + i) * BIG_PITCH + lane] = X;
                A[(1 * N + i) * BIG_PITCH + lane] = Y;
                A[(2 * N + i) * BIG_PITCH + lane] = B;
            }
            if (strip == 0) {
                // finalizeLLF for a corner up to 32x32: row pass into scratch, column pass into A (overwrites)
                const int lw = ilog2_pow2(v.dsW), lh = ilog2_pow2(v.dsH);
                const float invW = 1.0f / (float)v.dsW, invH = 1.0f / (float)v.dsH;
                for (int i = tid; i < 3 * v.dsH * v.dsW; i += nt) {
                    const int c = i / (v.dsH * v.dsW), r = i % (v.dsH * v.dsW);
                    const int y = r / v.dsW, kx = r % v.dsW;
                    const float *src = P.lf[c] + (size_t)(v.by + y) * P.wb + v.bx;
                    float d2;
                    if (kx == 0) {
                        d2 = src[0];
                        for (int n = 1; n < v.dsW; n++) d2 = __fadd_rn(d2, src[n]);
                    } else {
                        const float *lut = c_cos + c_cos_off[lw] + (kx - 1) * v.dsW;
                        d2 = __fmul_rn(src[0], lut[0]);
                        for (int n = 1; n < v.dsW; n++) d2 = __fadd_rn(d2, __fmul_rn(src[n], lut[n]));
                    }
                    scr[(c * 32 + y) * BIG_PITCH + kx] = __fmul_rn(d2, invW);
                }
                __syncthreads();
                for (int i = tid; i < 3 * v.dsH * v.dsW; i += nt) {
                    const int c = i / (v.dsH * v.dsW), r = i % (v.dsH * v.dsW);
                    const int ky = r / v.dsW, kx = r % v.dsW;
                    const float *col = scr + (c * 32) * BIG_PITCH + kx;
                    float d2;
                    if (ky == 0) {
                        d2 = col[0];
                        for (int n = 1; n < v.dsH; n++) d2 = __fadd_rn(d2, col[n * BIG_PITCH]);
                    } else {
                        const float *lut = c_cos + c_cos_off[lh] + (ky - 1) * v.dsH;
                        d2 = __fmul_rn(col[0], lut[0]);
                        for (int n = 1; n < v.dsH; n++) d2 = __fadd_rn(d2, __fmul_rn(col[n * BIG_PITCH], lut[n]));
                    }
                    d2 = __fmul_rn(d2, invH);
                    const float sc = __fmul_rn(c_llf_scale[ky << (5 - lh)], c_llf_scale[kx << (5 - lw)]);
                    A[(c * N + ky) * BIG_PITCH + kx] = __fmul_rn(d2, sc);
                }
            }
        } else {
            // 32 rows x N columns of the intermediate plane, transposed into A[c][x][row]
            for (int i = tid; i < 3 * 32 * N; i += nt) {
                const int c = i / (32 * N), r = (i / N) % 32, x = i % N;
                A[(c * N + x) * BIG_PITCH + r] = P.out[c][(size_t)(Y0 + strip * 32 + r) * P.out_pitch + X0 + x];
            }
        }
        __syncthreads();

        if (N == 32) {
            for (int c = warp; c < 3; c += nwarps) {
                float vv[32];
#pragma unroll
                for (int m = 0; m < 32; m++) vv[m] = A[(c * N + m) * BIG_PITCH + lane];
                RefIDCT<32>::run(vv, CosLut());
                if (PASS == 0) {
#pragma unroll
                    for (int m = 0; m < 32; m++)
                        P.out[c][(size_t)(Y0 + m) * P.out_pitch + X0 + strip * 32 + lane] = vv[m];
                } else {
#pragma unroll
                    for (int m = 0; m < 32; m++) A[(c * N + m) * BIG_PITCH + lane] = vv[m];
                }
            }
        } else {
            const float *__restrict__ lut = P.cos_big + cos_big_off(N);
            for (int w = warp; w < 3 * (N / 32); w += nwarps) {
                const int c = w / (N / 32), k0 = (w % (N / 32)) * 16;
                const float *src = A + (c * N) * BIG_PITCH + lane;
                float lo[16], hi[16];
                const float in0 = src[0];
#pragma unroll
                for (int q = 0; q < 16; q++) { lo[q] = in0; hi[q] = in0; }
#pragma unroll 2
                for (int n = 1; n < N; n++) {
                    const float s2 = src[n * BIG_PITCH];
                    if (__ballot_sync(0xffffffffu, s2 != 0.0f) == 0u) continue;
                    const float4 *lv = reinterpret_cast<const float4 *>(lut + (n - 1) * N + k0);
                    float l[16];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float4 t4 = __ldg(lv + q);
                        l[4 * q] = t4.x; l[4 * q + 1] = t4.y; l[4 * q + 2] = t4.z; l[4 * q + 3] = t4.w;
                    }
                    if (n & 1) {
#pragma unroll
                        for (int q = 0; q < 16; q++) {
                            const float p = __fmul_rn(s2, l[q]);
                            lo[q] = __fadd_rn(lo[q], p);
                            hi[q] = __fsub_rn(hi[q], p);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 16; q++) {
                            const float p = __fmul_rn(s2, l[q]);
                            lo[q] = __fadd_rn(lo[q], p);
                            hi[q] = __fadd_rn(hi[q], p);
                        }
                    }
                }
                if (PASS == 0) {
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        P.out[c][(size_t)(Y0 + k0 + q) * P.out_pitch + X0 + strip * 32 + lane] = lo[q];
                        P.out[c][(size_t)(Y0 + N - 1 - k0 - q) * P.out_pitch + X0 + strip * 32 + lane] = hi[q];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        T[(c * N + k0 + q) * BIG_PITCH + lane] = lo[q];
                        T[(c * N + N - 1 - k0 - q) * BIG_PITCH + lane] = hi[q];
                    }
                }
            }
        }
        if (PASS == 1) {
            __syncthreads();
            const float *O = N == 32 ? A : T;
            for (int i = tid; i < 3 * 32 * N; i += nt) {
                const int c = i / (32 * N), r = (i / N) % 32, x = i % N;
                P.out[c][(size_t)(Y0 + strip * 32 + r) * P.out_pitch + X0 + x] = O[(c * N + x) * BIG_PITCH + r];
            }
        }
        __syncthreads();
    }
}
