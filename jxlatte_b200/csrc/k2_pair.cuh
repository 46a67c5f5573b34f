// k2_pair.cuh -- K2 fused and bit-exact like k2_exact.cuh (same stages, same operation order, nothing contracted), with
// every thread working on TWO 2x2 pixel blocks at once through the packed FP32x2 instructions of sm_100 (FADD2 / FMUL2:
// two IEEE single-precision results per issue slot, each rounded exactly like the scalar instruction).
//
// STATUS: opt-in (JXLB200_OPT_STAGE2 = 3), bit-identical to k2_exact in every parity test, but SLOWER on B200: 2.60 ms
// against 1.88 ms for the 8K frame.  The pairing doubles the live state (188 registers), so a CTA is 128 threads and an SM
// holds 8 warps instead of 16; issue-active drops to 42 % (stalls: fixed-latency waits 28 %, short scoreboard 11 %) and the
// instruction count falls only 13 %, not the hoped-for 40 %: divisions, clamps and every product stay scalar, and the
// pack / unpack moves and per-half predicates add their own.  Kept as the record of the experiment and as a second,
// independently written evaluation of the same arithmetic (profiles/r1_k2_pair_ncu.md).
//
// Why it was tried: k2_exact is bound by instruction ISSUE (70 % issue-active, FP32 pipe 48 %): about a third of its instructions are
// address arithmetic, shared-memory loads and control that exist once per block.  Pairing two blocks in one thread halves
// those and halves the issue slots of every add and subtract, while the FP32 pipe does the same work as before.
//
// Replaces Frame.performGabConvolution (J/frame/Frame.java:505-542), Frame.performEdgePreservingFilter (:544-679) and
// JXLCodestreamDecoder.performColorTransforms (J/JXLCodestreamDecoder.java:256-283).
//
// Layout.  A CTA owns the same 64 x 32 tile + 8 halo as k2_exact (80 x 48 padded).  The two blocks of a thread are 36
// columns apart: pair index j in [0, 44) holds padded column j in its low half and column j + 36 in its high half, so a
// 64-bit shared-memory word is {left block's pixel, right block's pixel} and one 128-bit load brings two pixels of both.
// Columns 36..43 therefore exist twice (high half of j = 0..3, low half of j = 40..43: the right neighbours of the left
// half and the left neighbours of the right half); `fixup` refreshes the copies after every stage together with the
// mirrored out-of-frame positions.  EPF block pairs sit at even j in [4, 40): for pass 0 (margin 4) both halves are always
// inside the region, so no lane is wasted on the most expensive stage.
//
// ptxas (12.9) contracts mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2 even under -fmad=false, which would change the
// rounding.  It does not contract across the scalar / packed boundary (checked in SASS, and the parity tests would catch
// it), so every product that feeds a sum is formed by scalar FMUL and the sums, differences and chained products are packed.
#pragma once
#include "k2_exact.cuh"

#ifndef KP_THREADS
#define KP_THREADS 128
#endif
#define KP_TH 32                          /* tile height (its own: two CTAs of 102 KB share an SM) */
#define KP_PH (KP_TH + 2 * KX_HALO)
#define KP_NBY (KP_PH / 8)
#define KP_SHIFT 36                       /* columns between the two blocks of a pair */
#define KP_PJ 44                          /* pairs per padded row */
#define KP_PWX (2 * KP_PJ)                /* floats per padded row */
#define KP_PLANE (KP_PH * KP_PWX)
#define KP_BYTES (2 * 3 * KP_PLANE * 4 + (KP_PH + KX_PW) * 4 + KP_NBY * KX_NBX * 4)

typedef unsigned long long p2;            // two floats: low = left block, high = right block

__device__ __forceinline__ p2 p2_make(float lo, float hi) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float p2_lo(p2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float p2_hi(p2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ p2 p2_add(p2 a, p2 b) { p2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 p2_sub(p2 a, p2 b) { p2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 p2_mul(p2 a, p2 b) { p2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// both halves times their own factor by SCALAR multiplies: the result may feed a packed sum without being contracted into it
__device__ __forceinline__ p2 p2_smul(p2 a, p2 b) { return p2_make(__fmul_rn(p2_lo(a), p2_lo(b)), __fmul_rn(p2_hi(a), p2_hi(b))); }
__device__ __forceinline__ p2 p2_smul1(p2 a, float s) { return p2_make(__fmul_rn(p2_lo(a), s), __fmul_rn(p2_hi(a), s)); }

// primary float index of padded column c of row y, and the index of its copy (-1: the column exists once)
__device__ __forceinline__ int kp_idx(int y, int c) { return y * KP_PWX + (c < 40 ? 2 * c : 2 * (c - KP_SHIFT) + 1); }
__device__ __forceinline__ int kp_idx2(int y, int c) {
    return (c >= KP_SHIFT && c < 40) ? y * KP_PWX + 2 * (c - KP_SHIFT) + 1 : ((c >= 40 && c < KP_PJ) ? y * KP_PWX + 2 * c : -1);
}

// After a stage wrote its in-frame outputs: out-of-frame positions take the stage's value at the mirrored coordinate (see
// mirror_fill in k2_exact.cuh for why outputs, not inputs, are mirrored) and the second copy of columns 36..43 is refreshed.
__device__ __forceinline__ void kp_fixup(float *buf, const KxTile &T, int margin, bool edge) {
    const int rh = KP_TH + 2 * margin, rw = KX_TW + 2 * margin;
    if (!edge) {
        // nothing mirrors in this tile (all but the frame's border tiles): only the eight shared columns need their copy
        for (int i = threadIdx.x; i < rh * 8; i += KP_THREADS) {
            const int ly = KX_HALO - margin + (i >> 3), lx = KP_SHIFT + (i & 7);
            const int src = kp_idx(ly, lx), dst = kp_idx2(ly, lx);
#pragma unroll
            for (int c = 0; c < 3; c++) buf[c * KP_PLANE + dst] = buf[c * KP_PLANE + src];
        }
        return;
    }
    for (int i = threadIdx.x; i < rh * rw; i += KP_THREADS) {
        const int ly = KX_HALO - margin + i / rw, lx = KX_HALO - margin + i % rw;
        const int sy = T.mrow[ly], sx = T.mcol[lx];
        const bool outside = sy != ly || sx != lx;
        const int i2 = kp_idx2(ly, lx);
        if (outside || i2 >= 0) {
            const int src = kp_idx(sy, sx), dst = kp_idx(ly, lx);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float v = buf[c * KP_PLANE + src];
                if (outside) buf[c * KP_PLANE + dst] = v;
                if (i2 >= 0) buf[c * KP_PLANE + i2] = v;
            }
        }
    }
}

// window of pairs: rows [ly - R, ly + 1 + R], pair columns [j - R, j + 1 + R] of one channel; only the diamond is read.
// j is even, so pairs (q, q + 1) starting where q - R is even are one aligned 128-bit load.
template <int R> __device__ __forceinline__ void kp_load_window(const float *__restrict__ plane, int ly, int j, p2 (&W)[2 + 2 * R][2 + 2 * R]) {
    constexpr int WN = 2 + 2 * R;
#pragma unroll
    for (int r = 0; r < WN; r++) {
        const int dr = r < R ? R - r : (r > R + 1 ? r - R - 1 : 0);
        const int reach = R - dr;
        const float *row = plane + (ly - R + r) * KP_PWX + 2 * (j - R);
#pragma unroll
        for (int q = 0; q < WN; q++) W[r][q] = 0ull;
#pragma unroll
        for (int q = (R & 1); q < WN; q += 2) {
            if (q + 1 >= R - reach && q <= R + 1 + reach) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(row + 2 * q);
                W[r][q] = v.x;
                if (q + 1 < WN) W[r][q + 1] = v.y;
            }
        }
        if ((R & 1) && reach == R) W[r][0] = *reinterpret_cast<const p2 *>(row);
    }
}

// dist_channel of k2_exact.cuh on pairs: the differences and the running sums are packed, the products by the channel
// scale are scalar (|x| is a free operand modifier there too).
template <int R, int DY, int DX, bool PLUS>
__device__ __forceinline__ void kp_dist_channel(const p2 (&W)[2 + 2 * R][2 + 2 * R], float s, p2 (&dist)[2 + DY][2 + (DX > 0 ? DX : -DX)]) {
    constexpr int DXP = DX > 0 ? DX : 0, DXN = DX < 0 ? -DX : 0;
    constexpr int E = PLUS ? 1 : 0;
    constexpr int SR0 = -DY, SC0 = -DXP, SNR = 2 + DY, SNC = 2 + DXP + DXN;
    constexpr int TR0 = SR0 - E, TC0 = SC0 - E, TNR = SNR + 2 * E, TNC = SNC + 2 * E;
    p2 T[TNR][TNC];
#pragma unroll
    for (int r = 0; r < TNR; r++)
#pragma unroll
        for (int q = 0; q < TNC; q++) {
            const bool corner = PLUS && (r == 0 || r == TNR - 1) && (q == 0 || q == TNC - 1);
            const int y = R + TR0 + r, x = R + TC0 + q;
            if (corner) {
                T[r][q] = 0ull;
            } else {
                const p2 d = p2_sub(W[y][x], W[y + DY][x + DX]);
                T[r][q] = p2_make(__fmul_rn(fabsf(p2_lo(d)), s), __fmul_rn(fabsf(p2_hi(d)), s));
            }
        }
#pragma unroll
    for (int r = 0; r < SNR; r++)
#pragma unroll
        for (int q = 0; q < SNC; q++) {
            const bool used = (r >= DY && q >= DXP && q < DXP + 2) || (r < 2 && q >= DXN && q < DXN + 2);
            if (!used) continue;
            p2 d = p2_add(dist[r][q], T[r + E][q + E]);   // centre
            if (PLUS) {
                d = p2_add(d, T[r + 1][q]);               // (0, -1)
                d = p2_add(d, T[r + 1][q + 2]);           // (0, +1)
                d = p2_add(d, T[r][q + 1]);               // (-1, 0)
                d = p2_add(d, T[r + 2][q + 1]);           // (+1, 0)
            }
            dist[r][q] = d;
        }
}

// epfWeight (Frame.java:671-679) for both halves: the chained products are packed (a product feeding a product cannot be
// contracted), the 1 - x and the clamp are scalar
__device__ __forceinline__ p2 kp_w(p2 dist, p2 m, p2 ss, p2 is) {
    const p2 x = p2_mul(p2_mul(p2_mul(dist, m), ss), is);
    return p2_make(fmaxf(__fsub_rn(1.0f, p2_lo(x)), 0.0f), fmaxf(__fsub_rn(1.0f, p2_hi(x)), 0.0f));
}

// One EPF pass for the block pair at row ly (even), pair column j (even): left block at padded column j, right block at
// j + 36.  okA / okB: the block is inside the frame and inside this pass's region (otherwise nothing is stored for it).
template <int PASS>
__device__ __forceinline__ void kp_epf_pair(const K2Params &P, const KxTile &T, const float *__restrict__ in, float *__restrict__ outp,
                                            int ly, int j, bool okA, bool okB) {
    constexpr int R = EpfGeom<PASS>::R;
    constexpr bool PLUS = PASS != 2;
    constexpr int WN = 2 + 2 * R;
    const int cA = j, cB = j + KP_SHIFT;
    const float isA = T.isig[(ly >> 3) * KX_NBX + (cA >> 3)], isB = T.isig[(ly >> 3) * KX_NBX + (cB >> 3)];
    const bool actA = okA && (isA <= (1.0f / 0.3f)), actB = okB && (isB <= (1.0f / 0.3f));   // else copied through (Frame.java:608-612); also NaN
    if (!actA && !actB) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float *i0 = in + c * KP_PLANE + ly * KP_PWX + 2 * j;
            float *o = outp + c * KP_PLANE + ly * KP_PWX + 2 * j;
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const float4 v = *reinterpret_cast<const float4 *>(i0 + rr * KP_PWX);
                if (okA) { o[rr * KP_PWX] = v.x; o[rr * KP_PWX + 2] = v.z; }
                if (okB) { o[rr * KP_PWX + 1] = v.y; o[rr * KP_PWX + 3] = v.w; }
            }
        }
        return;
    }
    p2 m[4];
    {
        const int ry = ly & 7, rxA = cA & 7, rxB = cB & 7;   // even; the two blocks differ in their column inside an 8x8 block
        const bool by0 = ry == 0, by1 = ry == 6;
        const float one = 1.0f, bm = P.border_mul;
        m[0] = p2_make((by0 || rxA == 0) ? bm : one, (by0 || rxB == 0) ? bm : one);
        m[1] = p2_make((by0 || rxA == 6) ? bm : one, (by0 || rxB == 6) ? bm : one);
        m[2] = p2_make((by1 || rxA == 0) ? bm : one, (by1 || rxB == 0) ? bm : one);
        m[3] = p2_make((by1 || rxA == 6) ? bm : one, (by1 || rxB == 6) ? bm : one);
    }
    const p2 is = p2_make(isA, isB);
    const p2 ss = p2_make(P.sigma_scale[PASS], P.sigma_scale[PASS]);
    p2 d01[2][3], d10[3][2], d11[3][3], d1m[3][3], d02[2][4], d20[4][2];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (r < 2 && q < 3) d01[r][q] = 0ull;
            if (r < 3 && q < 2) d10[r][q] = 0ull;
            if (r < 3 && q < 3) { d11[r][q] = 0ull; d1m[r][q] = 0ull; }
            if (r < 2) d02[r][q] = 0ull;
            if (q < 2) d20[r][q] = 0ull;
        }
    const int nch = P.W > 0 ? 3 : 0;   // run-time trip count: keeps the channel loop rolled (see k2_exact.cuh)
#pragma unroll 1
    for (int c = 0; c < nch; c++) {
        p2 W[WN][WN];
        kp_load_window<R>(in + c * KP_PLANE, ly, j, W);
        const float s = P.ch_scale[c];
        kp_dist_channel<R, 0, 1, PLUS>(W, s, d01);
        kp_dist_channel<R, 1, 0, PLUS>(W, s, d10);
        if (PASS == 0) {
            kp_dist_channel<R, 1, 1, PLUS>(W, s, d11);
            kp_dist_channel<R, 1, -1, PLUS>(W, s, d1m);
            kp_dist_channel<R, 0, 2, PLUS>(W, s, d02);
            kp_dist_channel<R, 2, 0, PLUS>(W, s, d20);
        }
    }
    constexpr int NW = PASS == 0 ? 12 : 4;
    p2 w[4][NW];
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const int i = p >> 1, jj = p & 1;
        w[p][0] = kp_w(d01[i][jj], m[p], ss, is);
        w[p][1] = kp_w(d01[i][jj + 1], m[p], ss, is);
        w[p][2] = kp_w(d10[i][jj], m[p], ss, is);
        w[p][3] = kp_w(d10[i + 1][jj], m[p], ss, is);
        if (PASS == 0) {
            w[p][4] = kp_w(d1m[i][jj + 1], m[p], ss, is);
            w[p][5] = kp_w(d11[i + 1][jj + 1], m[p], ss, is);
            w[p][6] = kp_w(d1m[i + 1][jj], m[p], ss, is);
            w[p][7] = kp_w(d11[i][jj], m[p], ss, is);
            w[p][8] = kp_w(d02[i][jj], m[p], ss, is);
            w[p][9] = kp_w(d02[i][jj + 2], m[p], ss, is);
            w[p][10] = kp_w(d20[i + 2][jj], m[p], ss, is);
            w[p][11] = kp_w(d20[i][jj], m[p], ss, is);
        }
    }
    p2 sumw[4];
    const p2 one2 = p2_make(1.0f, 1.0f);
#pragma unroll
    for (int p = 0; p < 4; p++) {
        p2 s = one2;
#pragma unroll
        for (int k = 0; k < NW; k++) s = p2_add(s, w[p][k]);
        sumw[p] = s;
    }
    constexpr int R2 = PASS == 0 ? 2 : 1;
    constexpr int oy[12] = {0, 0, -1, 1, -1, 1, 1, -1, 0, 0, 2, -2};
    constexpr int ox[12] = {-1, 1, 0, 0, 1, 1, -1, -1, -2, 2, 0, 0};
#pragma unroll 1
    for (int c = 0; c < nch; c++) {
        p2 W[2 + 2 * R2][2 + 2 * R2];
        kp_load_window<R2>(in + c * KP_PLANE, ly, j, W);
        float ra[4], rb[4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int i = p >> 1, jj = p & 1;
            const p2 centre = W[R2 + i][R2 + jj];
            p2 s = centre;
#pragma unroll
            for (int k = 0; k < NW; k++) s = p2_add(s, p2_smul(W[R2 + i + oy[k]][R2 + jj + ox[k]], w[p][k]));
            ra[p] = actA ? __fdiv_rn(p2_lo(s), p2_lo(sumw[p])) : p2_lo(centre);
            rb[p] = actB ? __fdiv_rn(p2_hi(s), p2_hi(sumw[p])) : p2_hi(centre);
        }
        float *o = outp + c * KP_PLANE + ly * KP_PWX + 2 * j;
        if (okA && okB) {
            *reinterpret_cast<float4 *>(o) = make_float4(ra[0], rb[0], ra[1], rb[1]);
            *reinterpret_cast<float4 *>(o + KP_PWX) = make_float4(ra[2], rb[2], ra[3], rb[3]);
        } else if (okA) {
            o[0] = ra[0]; o[2] = ra[1]; o[KP_PWX] = ra[2]; o[KP_PWX + 2] = ra[3];
        } else {
            o[1] = rb[0]; o[3] = rb[1]; o[KP_PWX + 1] = rb[2]; o[KP_PWX + 3] = rb[3];
        }
    }
}

// blockIdx.z = frame of a vertically stacked batch (see k2_exact)
template <int GAB, int ITERS> __global__ void __launch_bounds__(KP_THREADS, 2) k2_pair(K2Params P, const float *__restrict__ inv_sigma,
                                                                                     long long zpx, int zblk) {
    constexpr int M0 = GAB + (ITERS == 3 ? 3 : 0) + (ITERS >= 1 ? 2 : 0) + (ITERS >= 2 ? 1 : 0);
    extern __shared__ float sm[];
    float *bufA = sm, *bufB = sm + 3 * KP_PLANE;
    int *mrow = reinterpret_cast<int *>(sm + 6 * KP_PLANE), *mcol = mrow + KP_PH;
    float *isig = reinterpret_cast<float *>(mcol + KX_PW);
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * KX_TW, ty0 = blockIdx.y * KP_TH;
    KxTile T{isig, mrow, mcol};
    const long long zo = blockIdx.z * zpx;
    inv_sigma += (long long)blockIdx.z * zblk;
    const int rlo = P.has_top ? -JXLB200_HALO_ROWS : 0, rhi = P.rows - 1 + (P.has_bottom ? JXLB200_HALO_ROWS : 0);

    for (int i = tid; i < KP_PH; i += KP_THREADS) {
        int r = mirror_row(ty0 - KX_HALO + i, P.rows, P.has_top, P.has_bottom);
        r = min(max(r, rlo), rhi);
        mrow[i] = min(max(r - (ty0 - KX_HALO), 0), KP_PH - 1);
    }
    for (int i = tid; i < KX_PW; i += KP_THREADS) {
        int x = mirror_col(tx0 - KX_HALO + i, P.W);
        x = min(max(x, 0), P.W - 1);
        mcol[i] = min(max(x - (tx0 - KX_HALO), 0), KX_PW - 1);
    }
    if (ITERS > 0) {
        for (int i = tid; i < KP_NBY * KX_NBX; i += KP_THREADS) {
            const int gy = ty0 - KX_HALO + 8 * (i / KX_NBX), gx = tx0 - KX_HALO + 8 * (i % KX_NBX);
            const bool inside = gy >= rlo && gy <= rhi && gx >= 0 && gx < P.W;
            isig[i] = inside ? __ldg(inv_sigma + (gy >> 3) * P.wb + (gx >> 3)) : __int_as_float(0x7fc00000);
        }
    }
    // raw tile -> bufA, both copies of the shared columns.  Interior tiles: a thread takes four pairs of a row = padded columns
    // [4g, 4g + 4) and [4g + 36, 4g + 40), two 128-bit global loads and two 128-bit shared stores.
    const bool interior = tx0 >= KX_HALO && tx0 + KX_TW + KX_HALO <= P.W && ty0 - KX_HALO >= rlo && ty0 + KP_TH + KX_HALO - 1 <= rhi &&
                          (P.in_pitch & 3) == 0;
    if (interior) {
        constexpr int G = KP_PJ / 4;
        for (int i = tid; i < 3 * KP_PH * G; i += KP_THREADS) {
            const int c = i / (KP_PH * G), rem = i - c * (KP_PH * G), ly = rem / G, g = rem - ly * G;
            const float *src = P.in[c] + zo + (long long)(ty0 - KX_HALO + ly) * P.in_pitch + tx0 - KX_HALO + 4 * g;
            const float4 a = __ldg(reinterpret_cast<const float4 *>(src)), b = __ldg(reinterpret_cast<const float4 *>(src + KP_SHIFT));
            float *dst = bufA + c * KP_PLANE + ly * KP_PWX + 8 * g;
            *reinterpret_cast<float4 *>(dst) = make_float4(a.x, b.x, a.y, b.y);
            *reinterpret_cast<float4 *>(dst + 4) = make_float4(a.z, b.z, a.w, b.w);
        }
    } else {
        for (int i = tid; i < KP_PH * KP_PJ; i += KP_THREADS) {
            const int ly = i / KP_PJ, j = i - ly * KP_PJ;
            int r = mirror_row(ty0 - KX_HALO + ly, P.rows, P.has_top, P.has_bottom);
            r = min(max(r, rlo), rhi);
            int xa = mirror_col(tx0 - KX_HALO + j, P.W), xb = mirror_col(tx0 - KX_HALO + j + KP_SHIFT, P.W);
            xa = min(max(xa, 0), P.W - 1);
            xb = min(max(xb, 0), P.W - 1);
            const long long oa = zo + (long long)r * P.in_pitch + xa, ob = zo + (long long)r * P.in_pitch + xb;
#pragma unroll
            for (int c = 0; c < 3; c++)
                *reinterpret_cast<float2 *>(bufA + c * KP_PLANE + ly * KP_PWX + 2 * j) = make_float2(__ldg(P.in[c] + oa), __ldg(P.in[c] + ob));
        }
    }
    // does anything in this padded tile mirror?  (CTA-uniform)
    __syncthreads();
    bool mirrors = false;
    for (int i = tid; i < KP_PH + KX_PW; i += KP_THREADS) mirrors |= i < KP_PH ? mrow[i] != i : mcol[i - KP_PH] != i - KP_PH;
    const bool edge = __syncthreads_or(mirrors) != 0;

    float *cur = bufA, *nxt = bufB;
    if (GAB) {
        // pixel pairs j in [2, 42): left column j, right column j + 36, each stored when it lies inside margin M0 - 1 and the frame
        constexpr int m = M0 - 1, rh = KP_TH + 2 * m, jw = 40;
        for (int i = tid; i < rh * jw; i += KP_THREADS) {
            const int ly = KX_HALO - m + i / jw, j = 2 + i % jw;
            const int cA = j, cB = j + KP_SHIFT;
            const bool rowok = mrow[ly] == ly;
            const bool okA = rowok && cA >= KX_HALO - m && cA < KX_HALO + KX_TW + m && mcol[cA] == cA;
            const bool okB = rowok && cB >= KX_HALO - m && cB < KX_HALO + KX_TW + m && mcol[cB] == cB;
            if (!okA && !okB) continue;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float *R = cur + c * KP_PLANE + ly * KP_PWX + 2 * j;
                const p2 *Ru = reinterpret_cast<const p2 *>(R - KP_PWX), *Rc = reinterpret_cast<const p2 *>(R), *Rd = reinterpret_cast<const p2 *>(R + KP_PWX);
                // Frame.java:535-537: operand order kept; the sums are packed, the products scalar
                const p2 adj = p2_add(p2_add(p2_add(Rc[-1], Rc[1]), Ru[0]), Rd[0]);
                const p2 diag = p2_add(p2_add(p2_add(Ru[-1], Ru[1]), Rd[-1]), Rd[1]);
                const p2 res = p2_add(p2_add(p2_smul1(Rc[0], P.gab_base[c]), p2_smul1(adj, P.gab_adj[c])), p2_smul1(diag, P.gab_diag[c]));
                float *o = nxt + c * KP_PLANE + ly * KP_PWX + 2 * j;
                if (okA) o[0] = p2_lo(res);
                if (okB) o[1] = p2_hi(res);
            }
        }
        __syncthreads();
        if (ITERS > 0) {
            kp_fixup(nxt, T, m, edge);
            __syncthreads();
        }
        float *t = cur; cur = nxt; nxt = t;
    }

    // EPF passes over block pairs at even rows and even pair columns j in [4, 40)
#define KP_RUN_PASS(PASS, MARGIN, LAST)                                                                                   \
    {                                                                                                                     \
        constexpr int mm = ((MARGIN) + 1) & ~1, rh = KP_TH + 2 * mm, nb = (rh / 2) * 18;                                   \
        _Pragma("unroll 1") for (int b = tid; b < nb; b += KP_THREADS) {                                                  \
            const int ly = KX_HALO - mm + 2 * (b / 18), j = 4 + 2 * (b % 18);                                              \
            const int cA = j, cB = j + KP_SHIFT;                                                                          \
            const bool rowok = mrow[ly] == ly;                                                                            \
            const bool okA = rowok && cA >= KX_HALO - mm && mcol[cA] == cA;                                               \
            const bool okB = rowok && cB < KX_HALO + KX_TW + mm && mcol[cB] == cB;                                        \
            if (!okA && !okB) continue;                                                                                   \
            kp_epf_pair<PASS>(P, T, cur, nxt, ly, j, okA, okB);                                                           \
        }                                                                                                                 \
        __syncthreads();                                                                                                  \
        if (!(LAST)) {                                                                                                    \
            kp_fixup(nxt, T, mm, edge);                                                                                         \
            __syncthreads();                                                                                              \
        }                                                                                                                 \
        float *t = cur; cur = nxt; nxt = t;                                                                               \
    }
    if (ITERS == 3) KP_RUN_PASS(0, 3, false)
    if (ITERS >= 2) {
        KP_RUN_PASS(1, 1, false)
        KP_RUN_PASS(2, 0, true)
    } else if (ITERS == 1) {
        KP_RUN_PASS(1, 0, true)
    }
#undef KP_RUN_PASS

    // colour transform + store: four pixels per thread, read from their primary slots, 128-bit rows out
    for (int i = tid; i < KP_TH * (KX_TW / 4); i += KP_THREADS) {
        const int ly = KX_HALO + i / (KX_TW / 4), lx = KX_HALO + 4 * (i % (KX_TW / 4));
        const int oy = ty0 + ly - KX_HALO, ox = tx0 + lx - KX_HALO;
        if (oy >= P.rows || ox >= P.W) continue;
        const int s0 = kp_idx(ly, lx);       // the four columns share a half (40 is a multiple of 4): stride 2
        float4 a = make_float4(cur[s0], cur[s0 + 2], cur[s0 + 4], cur[s0 + 6]);
        float4 b = make_float4(cur[KP_PLANE + s0], cur[KP_PLANE + s0 + 2], cur[KP_PLANE + s0 + 4], cur[KP_PLANE + s0 + 6]);
        float4 c = make_float4(cur[2 * KP_PLANE + s0], cur[2 * KP_PLANE + s0 + 2], cur[2 * KP_PLANE + s0 + 4], cur[2 * KP_PLANE + s0 + 6]);
        color_px(P, a.x, b.x, c.x); color_px(P, a.y, b.y, c.y); color_px(P, a.z, b.z, c.z); color_px(P, a.w, b.w, c.w);
        const long long o = zo + (long long)oy * P.out_pitch + ox;
        if ((P.out_pitch & 3) == 0) {
            *reinterpret_cast<float4 *>(P.out[0] + o) = a;
            *reinterpret_cast<float4 *>(P.out[1] + o) = b;
            *reinterpret_cast<float4 *>(P.out[2] + o) = c;
        } else {
            P.out[0][o] = a.x; P.out[0][o + 1] = a.y; P.out[0][o + 2] = a.z; P.out[0][o + 3] = a.w;
            P.out[1][o] = b.x; P.out[1][o + 1] = b.y; P.out[1][o + 2] = b.z; P.out[1][o + 3] = b.w;
            P.out[2][o] = c.x; P.out[2][o + 1] = c.y; P.out[2][o + 2] = c.z; P.out[2][o + 3] = c.w;
        }
    }
}

template <int GAB, int ITERS> static cudaError_t k2_pair_attr() {
    return cudaFuncSetAttribute(k2_pair<GAB, ITERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, KP_BYTES);
}
static inline cudaError_t k2_pair_init_all() {
    cudaError_t e;
    if ((e = k2_pair_attr<1, 0>()) != cudaSuccess) return e;
    if ((e = k2_pair_attr<1, 1>()) != cudaSuccess) return e;
    if ((e = k2_pair_attr<1, 2>()) != cudaSuccess) return e;
    if ((e = k2_pair_attr<1, 3>()) != cudaSuccess) return e;
    if ((e = k2_pair_attr<0, 1>()) != cudaSuccess) return e;
    if ((e = k2_pair_attr<0, 2>()) != cudaSuccess) return e;
    if ((e = k2_pair_attr<0, 3>()) != cudaSuccess) return e;
    return cudaSuccess;
}
template <int GAB, int ITERS> static void k2_pair_go(const K2Params &K, const float *inv_sigma, cudaStream_t st, int nz, long long zpx, int zblk) {
    const dim3 grid((K.W + KX_TW - 1) / KX_TW, (K.rows + KP_TH - 1) / KP_TH, nz);
    k2_pair<GAB, ITERS><<<grid, KP_THREADS, KP_BYTES, st>>>(K, inv_sigma, zpx, zblk);
}
static inline void k2_pair_dispatch(const K2Params &K, const float *inv_sigma, cudaStream_t st, int n_frames = 1) {
    const long long zpx = (long long)K.rows * K.in_pitch;
    const int zblk = (K.rows >> 3) * K.wb;
    switch ((K.gab ? 4 : 0) + K.iters) {
    case 4: k2_pair_go<1, 0>(K, inv_sigma, st, n_frames, zpx, zblk); break;
    case 5: k2_pair_go<1, 1>(K, inv_sigma, st, n_frames, zpx, zblk); break;
    case 6: k2_pair_go<1, 2>(K, inv_sigma, st, n_frames, zpx, zblk); break;
    case 7: k2_pair_go<1, 3>(K, inv_sigma, st, n_frames, zpx, zblk); break;
    case 1: k2_pair_go<0, 1>(K, inv_sigma, st, n_frames, zpx, zblk); break;
    case 2: k2_pair_go<0, 2>(K, inv_sigma, st, n_frames, zpx, zblk); break;
    default: k2_pair_go<0, 3>(K, inv_sigma, st, n_frames, zpx, zblk); break;
    }
}
