// transforms.cuh -- the arithmetic of the 27 varblock inverse transforms, written for registers.
//
// Everything here is __host__ __device__ so tests/test_transforms_host.py can run the exact same code on the CPU
// box (no GPU needed) against the oracle before it ever reaches a B200.
//
// What the reference computes (J/ = /root/reference/java/com/traneptora/jxlatte/):
//   1-D:  out[k] = in[0] + sum_{n>=1} in[n] * sqrt2 * cos(pi n (k + 1/2) / N)      J/util/MathHelper.java:68-78
// jxlatte evaluates that as an O(N^2) sum; here it is Lee's recursive factorisation (even / pre-added-odd halves,
// one secant multiply per butterfly), O(N log N), fully unrolled in registers for N <= 32.  Bigger N are split in
// shared memory by lee_gather() / lee_combine() around the 32-point register kernel.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define JXLB_HD __host__ __device__ __forceinline__
#else
#define JXLB_HD inline
#endif

#define JXLB_SQRT2 1.41421356237309504880f

template <int N> JXLB_HD float lee_sec(int k);
#include "idct_tables.cuh"

// R_N: reference-convention inverse DCT of v[0..N), in place.
template <int N> struct LeeIDCT {
    static JXLB_HD void run(float *v) {
        float e[N / 2], o[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; i++) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
#pragma unroll
        for (int i = N / 2 - 1; i >= 1; i--) o[i] += o[i - 1];
        o[0] *= JXLB_SQRT2;
        LeeIDCT<N / 2>::run(e);
        LeeIDCT<N / 2>::run(o);
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            const float s = lee_sec<N>(k);
            v[k] = fmaf(s, o[k], e[k]);
            v[N - 1 - k] = fmaf(-s, o[k], e[k]);
        }
    }
};
template <> struct LeeIDCT<2> {
    static JXLB_HD void run(float *v) { const float a = v[0], b = v[1]; v[0] = a + b; v[1] = a - b; }
};
template <> struct LeeIDCT<1> {
    static JXLB_HD void run(float *) {}
};

// ---- splitting a long line (N = 32 * 2^L) around the 32-point register kernel ------------------------------
// s_0 = X;  s_{l+1}[m] = b_l ? (m ? s_l[2m+1] + s_l[2m-1] : sqrt2 * s_l[1]) : s_l[2m];   b_l = bit l of path.
// lee_gather<L>(in, path, m) = s_L[m]: element m of the 32-point sub-sequence selected by `path`.
template <int L> struct LeeGather {
    template <class In> static JXLB_HD float get(In in, int path, int m) {
        const int b = (path >> (L - 1)) & 1;
        if (!b) return LeeGather<L - 1>::get(in, path, 2 * m);
        if (m == 0) return JXLB_SQRT2 * LeeGather<L - 1>::get(in, path, 1);
        return LeeGather<L - 1>::get(in, path, 2 * m + 1) + LeeGather<L - 1>::get(in, path, 2 * m - 1);
    }
};
template <> struct LeeGather<0> {
    template <class In> static JXLB_HD float get(In in, int, int m) { return in(m); }
};

// lee_combine<L>: val[p] = R_32(sub-sequence p)[k0] for the R = 2^L paths; on return val[s] is output sample
// idx[s] of the full N-point transform (s = 0..R-1).  sec(n, k) = 1 / (2 cos(pi (2k+1) / (2n))).
template <int L, class Sec> JXLB_HD void lee_combine(float *val, int *idx, int k0, Sec sec) {
    constexpr int R = 1 << L;
    idx[0] = k0;
#pragma unroll
    for (int j = 0; j < L; j++) {
        const int n2 = 64 << j;          // length after this combine
        const int nseq = R >> j;         // sequences before this combine, each holding (1 << j) samples
        const int ns = 1 << j;
        float nv[R];
        int ni[R];
#pragma unroll
        for (int pre = 0; pre < nseq / 2; pre++) {
#pragma unroll
            for (int s = 0; s < ns; s++) {
                const float g = val[pre * ns + s];
                const float h = val[(pre + nseq / 2) * ns + s];
                const int ki = idx[s];
                const float sc = sec(n2, ki);
                nv[pre * 2 * ns + 2 * s] = fmaf(sc, h, g);
                nv[pre * 2 * ns + 2 * s + 1] = fmaf(-sc, h, g);
                ni[2 * s] = ki;
                ni[2 * s + 1] = n2 - 1 - ki;
            }
        }
#pragma unroll
        for (int i = 0; i < R; i++) val[i] = nv[i];
#pragma unroll
        for (int i = 0; i < 2 * ns; i++) idx[i] = ni[i];
    }
}

// ---- 8x8-class varblocks: v[64] (row-major dequantised coefficients, LLF already in v[0]) -> out(y, x, value) ----
// All follow J/frame/group/PassGroup.java:88-168, 227-325 statement by statement; only the 4/8-point DCTs inside are Lee.

// METHOD_DCT 8x8  (PassGroup.java:230-233 -> MathHelper.inverseDCT2D :96-122, columns then rows)
template <class Out> JXLB_HD void inv_dct8x8(float *v, Out out) {
#pragma unroll
    for (int x = 0; x < 8; x++) {
        float c[8];
#pragma unroll
        for (int y = 0; y < 8; y++) c[y] = v[y * 8 + x];
        LeeIDCT<8>::run(c);
#pragma unroll
        for (int y = 0; y < 8; y++) v[y * 8 + x] = c[y];
    }
#pragma unroll
    for (int y = 0; y < 8; y++) {
        LeeIDCT<8>::run(v + y * 8);
#pragma unroll
        for (int x = 0; x < 8; x++) out(y, x, v[y * 8 + x]);
    }
}

// 2x2 Hadamard of PassGroup.auxDCT2 (:154-165), operand order kept
JXLB_HD void aux2x2(float c00, float c01, float c10, float c11, float &r00, float &r01, float &r10, float &r11) {
    r00 = c00 + c01 + c10 + c11;
    r01 = c00 + c01 - c10 - c11;
    r10 = c00 - c01 + c10 - c11;
    r11 = c00 - c01 - c10 + c11;
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
                for (int ix = (iy == 0 ? 1 : 0); ix < 4; ix++) residual += v[(y + iy * 2) * 8 + x + ix * 2];
            const float centre = blockLF - residual * 0.0625f;
#pragma unroll
            for (int iy = 0; iy < 4; iy++)
#pragma unroll
                for (int ix = 0; ix < 4; ix++) {
                    if (ix == 1 && iy == 1) out(4 * y + 1, 4 * x + 1, centre);
                    else if (ix == 0 && iy == 0) out(4 * y, 4 * x, v[(y + 2) * 8 + x + 2] + centre);
                    else out(y * 4 + iy, x * 4 + ix, v[(y + iy * 2) * 8 + x + ix * 2] + centre);
                }
        }
}

// METHOD_DCT4  (PassGroup.java:306-325): four 4x4 IDCTs, transposed = true
template <class Out> JXLB_HD void inv_dct4(const float *v, Out out) {
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
            for (int iy = 0; iy < 4; iy++) LeeIDCT<4>::run(b + iy * 4);   // over ix -> k
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float c[4] = {b[k], b[4 + k], b[8 + k], b[12 + k]};        // over iy -> m
                LeeIDCT<4>::run(c);
#pragma unroll
                for (int m = 0; m < 4; m++) out(4 * y + k, 4 * x + m, c[m]);
            }
        }
}

// METHOD_DCT8_4 (TransformType.DCT8_4, PassGroup.java:234-251): two 4x8 coefficient sets, transposed = true
// -> two 8-row x 4-col halves side by side.  METHOD_DCT4_8 (:252-269): transposed = false -> two 4x8 halves stacked.
template <bool kTransposed, class Out> JXLB_HD void inv_dct4x8(const float *v, Out out) {
    const float coeff0 = v[0], coeff1 = v[8];
    const float lfs[2] = {coeff0 + coeff1, coeff0 - coeff1};
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
            for (int iy = 0; iy < 4; iy++) LeeIDCT<8>::run(b + iy * 8);  // over ix -> k (pixel row)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float c[4] = {b[k], b[8 + k], b[16 + k], b[24 + k]};      // over iy -> m (pixel column)
                LeeIDCT<4>::run(c);
#pragma unroll
                for (int m = 0; m < 4; m++) out(k, 4 * h + m, c[m]);
            }
        } else {
#pragma unroll
            for (int ix = 0; ix < 8; ix++) {
                float c[4] = {b[ix], b[8 + ix], b[16 + ix], b[24 + ix]};  // columns first: over iy -> m
                LeeIDCT<4>::run(c);
                b[ix] = c[0]; b[8 + ix] = c[1]; b[16 + ix] = c[2]; b[24 + ix] = c[3];
            }
#pragma unroll
            for (int m = 0; m < 4; m++) {
                LeeIDCT<8>::run(b + m * 8);                                // rows: over ix -> k
#pragma unroll
                for (int k = 0; k < 8; k++) out(4 * h + m, k, b[m * 8 + k]);
            }
        }
    }
}

// METHOD_AFV  (PassGroup.invertAFV :88-147).  basis(j, i) = AFV_BASIS[j][i] (:19-58).
template <class Basis, class Out> JXLB_HD void inv_afv(const float *v, int flipY, int flipX, Basis basis, Out out) {
    float a[16];
#pragma unroll
    for (int iy = 0; iy < 4; iy++)
#pragma unroll
        for (int ix = 0; ix < 4; ix++) a[iy * 4 + ix] = v[(iy * 2) * 8 + ix * 2];
    a[0] = (v[0] + v[8] + v[1]) * 4.0f;
#pragma unroll
    for (int iy = 0; iy < 4; iy++)
#pragma unroll
        for (int ix = 0; ix < 4; ix++) {
            float sample = 0.0f;
#pragma unroll
            for (int j = 0; j < 16; j++) sample += a[j] * basis(j, iy * 4 + ix);
            // scratch[1][iy][ix] lands at buffer[flipY*4 + (flipY ? 3-iy : iy)][flipX*4 + (flipX ? 3-ix : ix)]
            out(flipY * 4 + (flipY ? 3 - iy : iy), flipX * 4 + (flipX ? 3 - ix : ix), sample);
        }
    // "SPEC: watch signs here"
#pragma unroll
    for (int iy = 0; iy < 4; iy++)
#pragma unroll
        for (int ix = 0; ix < 4; ix++) a[iy * 4 + ix] = v[(iy * 2) * 8 + ix * 2 + 1];
    a[0] = v[0] + v[8] - v[1];
    {   // inverseDCT2D(4x4, transposed = false): columns (over iy) then rows (over ix)
#pragma unroll
        for (int ix = 0; ix < 4; ix++) {
            float c[4] = {a[ix], a[4 + ix], a[8 + ix], a[12 + ix]};
            LeeIDCT<4>::run(c);
            a[ix] = c[0]; a[4 + ix] = c[1]; a[8 + ix] = c[2]; a[12 + ix] = c[3];
        }
#pragma unroll
        for (int iy = 0; iy < 4; iy++) LeeIDCT<4>::run(a + iy * 4);
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
    b[0] = v[0] - v[8];
#pragma unroll
    for (int ix = 0; ix < 8; ix++) {
        float c[4] = {b[ix], b[8 + ix], b[16 + ix], b[24 + ix]};
        LeeIDCT<4>::run(c);
        b[ix] = c[0]; b[8 + ix] = c[1]; b[16 + ix] = c[2]; b[24 + ix] = c[3];
    }
#pragma unroll
    for (int iy = 0; iy < 4; iy++) {
        LeeIDCT<8>::run(b + iy * 8);
#pragma unroll
        for (int ix = 0; ix < 8; ix++) out((flipY ? 0 : 4) + iy, ix, b[iy * 8 + ix]);
    }
}
