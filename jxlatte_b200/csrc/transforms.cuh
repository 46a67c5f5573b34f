// transforms.cuh -- the arithmetic of the 27 varblock inverse transforms, written for registers.
//
// Everything here is __host__ __device__ so tests/test_transforms_host.py can run the exact same code on the CPU
// box (no GPU needed) against the oracle before it ever reaches a B200.
//
// What the reference computes (J/ = /root/reference/java/com/traneptora/jxlatte/):
//   1-D:  out[k] = in[0] + sum_{n>=1} in[n] * lut[n-1][k],  lut = (float)(sqrt2 cos(pi n (k + 1/2) / N))   J/util/MathHelper.java:17-30, 68-78
//
// The CUDA path evaluates exactly that sum: same products, same order of float additions per output, no FMA
// contraction -> the XYB planes are bit-identical to the Java's.  A fast O(N log N) factorisation (Lee, even in
// double) was measured and rejected: it differs from the Java by the Java's own accumulated rounding (X 4e-8, Y 7e-7
// on DCT64+ blocks), and 3e-8 on X already moves a saturated dark pixel by more than one 16-bit sRGB step because
// the opsin matrix rows cancel to 1e-3 of their terms there.  RefIDCT<N> is the register version for N <= 32; lines of
// 64/128/256 use the same recurrence with the table in global memory (k1_idct.cuh, k1_big).
#pragma once
#include <math.h>

#ifdef __CUDA_ARCH__
#define JXLB_HD __host__ __device__ __forceinline__
#define JXLB_MUL(a, b) __fmul_rn((a), (b))
#define JXLB_ADD(a, b) __fadd_rn((a), (b))
#define JXLB_SUB(a, b) __fsub_rn((a), (b))
#elif defined(__CUDACC__)
#define JXLB_HD __host__ __device__ __forceinline__
#define JXLB_MUL(a, b) ((a) * (b))
#define JXLB_ADD(a, b) ((a) + (b))
#define JXLB_SUB(a, b) ((a) - (b))
#else
#define JXLB_HD inline
#define JXLB_MUL(a, b) ((a) * (b))   /* host test build uses -ffp-contract=off */
#define JXLB_ADD(a, b) ((a) + (b))
#define JXLB_SUB(a, b) ((a) - (b))
#endif

// offset of level log2(N) in the packed cosine LUT (levels 1..5 = lengths 2..32): sum_{j<l} (2^j - 1) 2^j
JXLB_HD constexpr int cos_lut_off(int n) { return n == 2 ? 0 : n == 4 ? 2 : n == 8 ? 14 : n == 16 ? 70 : 310; }

// MathHelper.inverseDCTHorizontal (:68-78): same products, same order of additions per output.  The float table is
// exactly antisymmetric/symmetric, lut[n-1][N-1-k] == (-1)^n lut[n-1][k] (checked when the tables are built), so each
// product is formed once and added to out[k] and added to / subtracted from out[N-1-k]: bit-identical, 25% fewer ops.
template <int N> struct RefIDCT {
    template <class Lut> static JXLB_HD void run(float *v, Lut lut) {
        float o[N];
#pragma unroll
        for (int k = 0; k < N; k++) o[k] = v[0];
#pragma unroll
        for (int n = 1; n < N; n++) {
            const float s2 = v[n];
#pragma unroll
            for (int k = 0; k < N / 2; k++) {
                const float p = JXLB_MUL(s2, lut(cos_lut_off(N) + (n - 1) * N + k));
                o[k] = JXLB_ADD(o[k], p);
                o[N - 1 - k] = (n & 1) ? JXLB_SUB(o[N - 1 - k], p) : JXLB_ADD(o[N - 1 - k], p);
            }
        }
#pragma unroll
        for (int k = 0; k < N; k++) v[k] = o[k];
    }
};

// ---- 8x8-class varblocks: v[64] (row-major dequantised coefficients, LLF already in v[0]) -> out(y, x, value) ----
// All follow J/frame/group/PassGroup.java:88-168, 227-325 statement by statement, with explicit (uncontracted) float ops.

// METHOD_DCT 8x8  (PassGroup.java:230-233 -> MathHelper.inverseDCT2D :96-122, columns then rows)
template <class Lut, class Out> JXLB_HD void inv_dct8x8(float *v, Lut lut, Out out) {
#pragma unroll
    for (int x = 0; x < 8; x++) {
        float c[8];
#pragma unroll
        for (int y = 0; y < 8; y++) c[y] = v[y * 8 + x];
        RefIDCT<8>::run(c, lut);
#pragma unroll
        for (int y = 0; y < 8; y++) v[y * 8 + x] = c[y];
    }
#pragma unroll
    for (int y = 0; y < 8; y++) {
        RefIDCT<8>::run(v + y * 8, lut);
#pragma unroll
        for (int x = 0; x < 8; x++) out(y, x, v[y * 8 + x]);
    }
}

// 2x2 Hadamard of PassGroup.auxDCT2 (:154-165), operand order kept
JXLB_HD void aux2x2(float c00, float c01, float c10, float c11, float &r00, float &r01, float &r10, float &r11) {
    r00 = JXLB_ADD(JXLB_ADD(JXLB_ADD(c00, c01), c10), c11);
    r01 = JXLB_SUB(JXLB_SUB(JXLB_ADD(c00, c01), c10), c11);
    r10 = JXLB_SUB(JXLB_ADD(JXLB_SUB(c00, c01), c10), c11);
    r11 = JXLB_ADD(JXLB_SUB(JXLB_SUB(c00, c01), c10), c11);
}

// METHOD_DCT2  (PassGroup.java:273-277): auxDCT2 with s = 2, 4, 8
template <class Out> JXLB_HD void inv_dct2(float *v, Out out) {
    aux2x2(v[0], v[1], v[8], v[9], v[0], v[1], v[8], v[9]);
    float t[16];
#pragma unroll
    for (int iy = 0; iy < 2; iy++)
#pragma unroll
        for (int ix = 0; ix < 2; ix++)
            aux2x2(v[iy * 8 + ix], v[iy * 8 + ix + 2], v[(iy + 2) * 8 + ix], v[(iy + 2) * 8 + ix + 2],
                   t[(iy * 2) * 4 + ix * 2], t[(iy * 2) * 4 + ix * 2 + 1], t[(iy * 2 + 1) * 4 + ix * 2], t[(iy * 2 + 1) * 4 + ix * 2 + 1]);
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
        for (int x = 0; x < 4; x++) v[y * 8 + x] = t[y * 4 + x];
#pragma unroll
    for (int iy = 0; iy < 4; iy++)
#pragma unroll
        for (int ix = 0; ix < 4; ix++) {
            float r00, r01, r10, r11;
            aux2x2(v[iy * 8 + ix], v[iy * 8 + ix + 4], v[(iy + 4) * 8 + ix], v[(iy + 4) * 8 + ix + 4], r00, r01, r10, r11);
            out(iy * 2, ix * 2, r00);
            out(iy * 2, ix * 2 + 1, r01);
            out(iy * 2 + 1, ix * 2, r10);
            out(iy * 2 + 1, ix * 2 + 1, r11);
        }
}

// METHOD_HORNUSS  (PassGroup.java:278-305)
template <class Out> JXLB_HD void inv_hornuss(const float *v, Out out) {
    float lf[4];
    aux2x2(v[0], v[1], v[8], v[9], lf[0], lf[1], lf[2], lf[3]);
#pragma unroll
    for (int y = 0; y < 2; y++)
#pragma unroll
        for (int x = 0; x < 2; x++) {
            const float blockLF = lf[y * 2 + x];
            float residual = 0.0f;
#pragma unroll
            for (int iy = 0; iy < 4; iy++)
#pragma unroll
                for (int ix = (iy == 0 ? 1 : 0); ix < 4; ix++) residual = JXLB_ADD(residual, v[(y + iy * 2) * 8 + x + ix * 2]);
            const float centre = JXLB_SUB(blockLF, JXLB_MUL(residual, 0.0625f));
#pragma unroll
            for (int iy = 0; iy < 4; iy++)
#pragma unroll
                for (int ix = 0; ix < 4; ix++) {
                    if (ix == 1 && iy == 1) out(4 * y + 1, 4 * x + 1, centre);
                    else if (ix == 0 && iy == 0) out(4 * y, 4 * x, JXLB_ADD(v[(y + 2) * 8 + x + 2], centre));
                    else out(y * 4 + iy, x * 4 + ix, JXLB_ADD(v[(y + iy * 2) * 8 + x + ix * 2], centre));
                }
        }
}

// METHOD_DCT4  (PassGroup.java:306-325): four 4x4 IDCTs, transposed = true
template <class Lut, class Out> JXLB_HD void inv_dct4(const float *v, Lut lut, Out out) {
    float lf[4];
    aux2x2(v[0], v[1], v[8], v[9], lf[0], lf[1], lf[2], lf[3]);
#pragma unroll
    for (int y = 0; y < 2; y++)
#pragma unroll
        for (int x = 0; x < 2; x++) {
            float b[16];
#pragma unroll
            for (int iy = 0; iy < 4; iy++)
#pragma unroll
                for (int ix = 0; ix < 4; ix++) b[iy * 4 + ix] = v[(y + iy * 2) * 8 + x + ix * 2];
            b[0] = lf[y * 2 + x];
#pragma unroll
            for (int iy = 0; iy < 4; iy++) RefIDCT<4>::run(b + iy * 4, lut);   // over ix -> k
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float c[4] = {b[k], b[4 + k], b[8 + k], b[12 + k]};        // over iy -> m
                RefIDCT<4>::run(c, lut);
#pragma unroll
                for (int m = 0; m < 4; m++) out(4 * y + k, 4 * x + m, c[m]);
            }
        }
}

// METHOD_DCT8_4 (TransformType.DCT8_4, PassGroup.java:234-251): two 4x8 coefficient sets, transposed = true
// -> two 8-row x 4-col halves side by side.  METHOD_DCT4_8 (:252-269): transposed = false -> two 4x8 halves stacked.
template <bool kTransposed, class Lut, class Out> JXLB_HD void inv_dct4x8(const float *v, Lut lut, Out out) {
    const float coeff0 = v[0], coeff1 = v[8];
    const float lfs[2] = {JXLB_ADD(coeff0, coeff1), JXLB_SUB(coeff0, coeff1)};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        float b[32];
#pragma unroll
        for (int iy = 0; iy < 4; iy++)
#pragma unroll
            for (int ix = 0; ix < 8; ix++) b[iy * 8 + ix] = v[(h + iy * 2) * 8 + ix];
        b[0] = lfs[h];
        if (kTransposed) {
#pragma unroll
            for (int iy = 0; iy < 4; iy++) RefIDCT<8>::run(b + iy * 8, lut);  // over ix -> k (pixel row)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float c[4] = {b[k], b[8 + k], b[16 + k], b[24 + k]};      // over iy -> m (pixel column)
                RefIDCT<4>::run(c, lut);
#pragma unroll
                for (int m = 0; m < 4; m++) out(k, 4 * h + m, c[m]);
            }
        } else {
#pragma unroll
            for (int ix = 0; ix < 8; ix++) {
                float c[4] = {b[ix], b[8 + ix], b[16 + ix], b[24 + ix]};  // columns first: over iy -> m
                RefIDCT<4>::run(c, lut);
                b[ix] = c[0]; b[8 + ix] = c[1]; b[16 + ix] = c[2]; b[24 + ix] = c[3];
            }
#pragma unroll
            for (int m = 0; m < 4; m++) {
                RefIDCT<8>::run(b + m * 8, lut);                                // rows: over ix -> k
#pragma unroll
                for (int k = 0; k < 8; k++) out(4 * h + m, k, b[m * 8 + k]);
            }
        }
    }
}

// METHOD_AFV  (PassGroup.invertAFV :88-147).  basis(j, i) = AFV_BASIS[j][i] (:19-58).
template <class Basis, class Lut, class Out> JXLB_HD void inv_afv(const float *v, int flipY, int flipX, Basis basis, Lut lut, Out out) {
    float a[16];
#pragma unroll
    for (int iy = 0; iy < 4; iy++)
#pragma unroll
        for (int ix = 0; ix < 4; ix++) a[iy * 4 + ix] = v[(iy * 2) * 8 + ix * 2];
    a[0] = JXLB_MUL(JXLB_ADD(JXLB_ADD(v[0], v[8]), v[1]), 4.0f);
#pragma unroll
    for (int iy = 0; iy < 4; iy++)
#pragma unroll
        for (int ix = 0; ix < 4; ix++) {
            float sample = 0.0f;
#pragma unroll
            for (int j = 0; j < 16; j++) sample = JXLB_ADD(sample, JXLB_MUL(a[j], basis(j, iy * 4 + ix)));
            // scratch[1][iy][ix] lands at buffer[flipY*4 + (flipY ? 3-iy : iy)][flipX*4 + (flipX ? 3-ix : ix)]
            out(flipY * 4 + (flipY ? 3 - iy : iy), flipX * 4 + (flipX ? 3 - ix : ix), sample);
        }
    // "SPEC: watch signs here"
#pragma unroll
    for (int iy = 0; iy < 4; iy++)
#pragma unroll
        for (int ix = 0; ix < 4; ix++) a[iy * 4 + ix] = v[(iy * 2) * 8 + ix * 2 + 1];
    a[0] = JXLB_SUB(JXLB_ADD(v[0], v[8]), v[1]);
    {   // inverseDCT2D(4x4, transposed = false): columns (over iy) then rows (over ix)
#pragma unroll
        for (int ix = 0; ix < 4; ix++) {
            float c[4] = {a[ix], a[4 + ix], a[8 + ix], a[12 + ix]};
            RefIDCT<4>::run(c, lut);
            a[ix] = c[0]; a[4 + ix] = c[1]; a[8 + ix] = c[2]; a[12 + ix] = c[3];
        }
#pragma unroll
        for (int iy = 0; iy < 4; iy++) RefIDCT<4>::run(a + iy * 4, lut);
    }
#pragma unroll
    for (int iy = 0; iy < 4; iy++)
#pragma unroll
        for (int ix = 0; ix < 4; ix++)  // "transposed intentionally"
            out(flipY * 4 + iy, (flipX ? 0 : 4) + ix, a[ix * 4 + iy]);
    float b[32];
#pragma unroll
    for (int iy = 0; iy < 4; iy++)
#pragma unroll
        for (int ix = 0; ix < 8; ix++) b[iy * 8 + ix] = v[(1 + iy * 2) * 8 + ix];
    b[0] = JXLB_SUB(v[0], v[8]);
#pragma unroll
    for (int ix = 0; ix < 8; ix++) {
        float c[4] = {b[ix], b[8 + ix], b[16 + ix], b[24 + ix]};
        RefIDCT<4>::run(c, lut);
        b[ix] = c[0]; b[8 + ix] = c[1]; b[16 + ix] = c[2]; b[24 + ix] = c[3];
    }
#pragma unroll
    for (int iy = 0; iy < 4; iy++) {
        RefIDCT<8>::run(b + iy * 8, lut);
#pragma unroll
        for (int ix = 0; ix < 8; ix++) out((flipY ? 0 : 4) + iy, ix, b[iy * 8 + ix]);
    }
}
