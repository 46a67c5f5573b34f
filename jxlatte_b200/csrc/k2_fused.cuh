// k2_fused.cuh -- K2 fused: Gaborish -> EPF pass 0/1/2 -> XYB->linear / YCbCr in ONE kernel, one HBM read and one HBM
// write per pixel (k2_restore.cuh is the staged, bit-exact version of the same stages).
//
// Replaces Frame.performGabConvolution (J/frame/Frame.java:505-542), Frame.performEdgePreservingFilter (:544-679) and
// JXLCodestreamDecoder.performColorTransforms (J/JXLCodestreamDecoder.java:256-283).
//
// A CTA owns a TH x TW output tile.  It loads the tile plus a halo of M0 = gab + 3 + 2 + 1 pixels (mirror-extended at
// the true frame edges, real neighbour rows where a slab has them) into shared memory and runs every stage there,
// ping-ponging between two plane sets; each stage shrinks the valid region by its reach, the last one applies the
// colour transform and stores to global memory.
//
// EPF arithmetic.  The reference evaluates, per pixel p and offset d, SAD_d(p) = sum_c s_c sum_{k in plus} |I_c(p+k) -
// I_c(p+d+k)|.  Here D_d(q) = sum_c s_c |I_c(q) - I_c(q+d)| is formed once per position and
//   SAD_d(p)  = sum_{k in plus} D_d(p+k),        SAD_{-d}(p) = SAD_d(p-d),
// so only the 6 (pass 0) / 2 (passes 1, 2) offsets with dy > 0 or (dy == 0, dx > 0) are evaluated and each thread, which
// owns a 2x2 pixel block with its window in registers, shares the D values among its pixels.  Sums are re-associated
// and contracted to FMA, hence "within tolerance" rather than bit-exact; Gaborish and the colour transform keep the
// reference's operation order (uncontracted) because they are cheap and the colour matrix is ill-conditioned.
#pragma once
#include "common.cuh"
#include "k2_restore.cuh"

#define K2_TW 64
#define K2_TH 32
#define K2_THREADS 224

template <int GAB, int ITERS> struct K2Cfg {
    static constexpr int kR0 = ITERS == 3 ? 3 : 0, kR1 = ITERS >= 1 ? 2 : 0, kR2 = ITERS >= 2 ? 1 : 0;
    static constexpr int kM0 = GAB + kR0 + kR1 + kR2;           // halo of the raw tile
    static constexpr int kPH = K2_TH + 2 * kM0, kPW = K2_TW + 2 * kM0;
    static constexpr int kPlane = kPH * kPW;
    static constexpr int kBytes = 2 * 3 * kPlane * 4 + (kPH + kPW) * 8;
};

struct K2Tile {
    int ty0, tx0;            // frame-local (slab) coordinates of the tile's first output pixel
    const int *srow, *scol;  // per local row / column: sigma-map block row * wb, block column (mirrored at frame edges)
    const int *brow, *bcol;  // per local row / column: 1 if on an 8x8 block border row / column
};

// k = sigmaScale * invSigma * (border ? borderSadMul : 1), or a negative value for "copy through" (invSigma NaN or > 1/0.3)
template <int PASS> __device__ __forceinline__ float epf_k(const K2Params &P, const float *__restrict__ inv_sigma, const K2Tile &T, int ly, int lx) {
    const float s = __ldg(inv_sigma + T.srow[ly] + T.scol[lx]);
    if (!(s <= (1.0f / 0.3f))) return -1.0f;
    const float k = P.sigma_scale[PASS] * s;
    return (T.brow[ly] | T.bcol[lx]) ? k * P.border_mul : k;
}

// One canonical offset (DY, DX) and its negative for a 2x2 block.  I[c][r][q]: window with reach R around the block
// (block pixel (i, j) is I[c][R + i][R + j]).  PLUS: plus-shaped SAD (passes 0, 1) or point difference (pass 2).
template <int R, int DY, int DX, bool PLUS>
__device__ __forceinline__ void epf_pair(const float (&I)[3][2 + 2 * R][2 + 2 * R], const float (&sc)[3], const float (&k)[4],
                                         float (&sw)[4], float (&acc)[3][4]) {
    constexpr int DXP = DX > 0 ? DX : 0, DXN = DX < 0 ? -DX : 0;
    constexpr int E = PLUS ? 1 : 0;
    // SAD positions: rows [-DY, 1], cols [-DXP, 1 + DXN];  D positions: that box dilated by E
    constexpr int SR0 = -DY, SC0 = -DXP, SNR = 2 + DY, SNC = 2 + DXP + DXN;
    constexpr int DR0 = SR0 - E, DC0 = SC0 - E, DNR = SNR + 2 * E, DNC = SNC + 2 * E;
    float D[DNR][DNC];
#pragma unroll
    for (int r = 0; r < DNR; r++)
#pragma unroll
        for (int q = 0; q < DNC; q++) {
            const bool corner = PLUS && (r == 0 || r == DNR - 1) && (q == 0 || q == DNC - 1);
            if (corner) { D[r][q] = 0.0f; continue; }
            const int y = R + DR0 + r, x = R + DC0 + q;
            float d = fabsf(I[0][y][x] - I[0][y + DY][x + DX]) * sc[0];
            d = fmaf(fabsf(I[1][y][x] - I[1][y + DY][x + DX]), sc[1], d);
            d = fmaf(fabsf(I[2][y][x] - I[2][y + DY][x + DX]), sc[2], d);
            D[r][q] = d;
        }
    float S[SNR][SNC];
#pragma unroll
    for (int r = 0; r < SNR; r++)
#pragma unroll
        for (int q = 0; q < SNC; q++) {
            // SAD positions that no pixel of the block uses (neither as p nor as p - d) fold away
            if (PLUS) S[r][q] = (D[r + 1][q + 1] + D[r + 1][q]) + (D[r + 1][q + 2] + D[r][q + 1]) + D[r + 2][q + 1];
            else S[r][q] = D[r][q];
        }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int p = i * 2 + j;
            const float wp = fmaxf(0.0f, fmaf(-S[i - SR0][j - SC0], k[p], 1.0f));             // +d: SAD_d(p)
            const float wn = fmaxf(0.0f, fmaf(-S[i - DY - SR0][j - DX - SC0], k[p], 1.0f));   // -d: SAD_d(p - d)
            sw[p] += wp + wn;
#pragma unroll
            for (int c = 0; c < 3; c++)
                acc[c][p] = fmaf(wn, I[c][R + i - DY][R + j - DX], fmaf(wp, I[c][R + i + DY][R + j + DX], acc[c][p]));
        }
}

// One EPF pass for the 2x2 block whose top-left is local (ly, lx).  out(p, c) receives the filtered (or copied) values.
template <int PASS, class Out>
__device__ __forceinline__ void epf_block(const K2Params &P, const float *__restrict__ inv_sigma, const K2Tile &T,
                                          const float *__restrict__ in, int plane, int pitch, int ly, int lx, Out out) {
    constexpr int R = PASS == 0 ? 3 : PASS == 1 ? 2 : 1;
    constexpr int WN = 2 + 2 * R;
    float I[3][WN][WN];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < WN; r++)
#pragma unroll
            for (int q = 0; q < WN; q++) {
                // the window is a diamond: entries farther than R (Manhattan, from the block) are never read and fold away
                const int dr = r < R ? R - r : (r > R + 1 ? r - R - 1 : 0), dq = q < R ? R - q : (q > R + 1 ? q - R - 1 : 0);
                I[c][r][q] = (dr + dq <= R) ? in[c * plane + (ly - R + r) * pitch + lx - R + q] : 0.0f;
            }
    float k[4];
#pragma unroll
    for (int p = 0; p < 4; p++) k[p] = epf_k<PASS>(P, inv_sigma, T, ly + (p >> 1), lx + (p & 1));
    float sw[4] = {1.0f, 1.0f, 1.0f, 1.0f};   // centre tap: SAD 0 -> weight 1
    float acc[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int p = 0; p < 4; p++) acc[c][p] = I[c][R + (p >> 1)][R + (p & 1)];
    const float sc[3] = {P.ch_scale[0], P.ch_scale[1], P.ch_scale[2]};
    if (k[0] >= 0.0f || k[1] >= 0.0f || k[2] >= 0.0f || k[3] >= 0.0f) {
        if (PASS == 0) {
            epf_pair<R, 0, 1, true>(I, sc, k, sw, acc);
            epf_pair<R, 1, 0, true>(I, sc, k, sw, acc);
            epf_pair<R, 1, 1, true>(I, sc, k, sw, acc);
            epf_pair<R, 1, -1, true>(I, sc, k, sw, acc);
            epf_pair<R, 0, 2, true>(I, sc, k, sw, acc);
            epf_pair<R, 2, 0, true>(I, sc, k, sw, acc);
        } else if (PASS == 1) {
            epf_pair<R, 0, 1, true>(I, sc, k, sw, acc);
            epf_pair<R, 1, 0, true>(I, sc, k, sw, acc);
        } else {
            epf_pair<R, 0, 1, false>(I, sc, k, sw, acc);
            epf_pair<R, 1, 0, false>(I, sc, k, sw, acc);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const float inv = __frcp_rn(sw[p]);
        const bool copy = k[p] < 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) out(p, c, copy ? I[c][R + (p >> 1)][R + (p & 1)] : acc[c][p] * inv);
    }
}

template <int GAB, int ITERS> __global__ void __launch_bounds__(K2_THREADS) k2_fused(K2Params P, const float *__restrict__ inv_sigma) {
    using Cfg = K2Cfg<GAB, ITERS>;
    constexpr int M0 = Cfg::kM0, PH = Cfg::kPH, PW = Cfg::kPW, PLANE = Cfg::kPlane;
    extern __shared__ float sm[];
    float *bufA = sm, *bufB = sm + 3 * PLANE;
    int *srow = reinterpret_cast<int *>(sm + 6 * PLANE), *brow = srow + PH, *scol = brow + PH, *bcol = scol + PW;
    const int tid = threadIdx.x;
    K2Tile T;
    T.tx0 = blockIdx.x * K2_TW;
    T.ty0 = blockIdx.y * K2_TH;
    T.srow = srow; T.scol = scol; T.brow = brow; T.bcol = bcol;

    // per local row / column: source coordinate (mirrored at true frame edges), sigma index, block-border flag
    for (int i = tid; i < PH; i += K2_THREADS) {
        const int r = mirror_row(T.ty0 - M0 + i, P.rows, P.has_top, P.has_bottom);
        // rows that exist nowhere (beyond the halo of a short last tile) are clamped; they only feed discarded outputs
        const int rc = min(max(r, P.has_top ? -JXLB200_HALO_ROWS : 0), P.rows - 1 + (P.has_bottom ? JXLB200_HALO_ROWS : 0));
        srow[i] = (rc >> 3) * P.wb;
        brow[i] = ((rc & 7) == 0 || (rc & 7) == 7) ? 1 : 0;
    }
    for (int i = tid; i < PW; i += K2_THREADS) {
        int x = mirror_col(T.tx0 - M0 + i, P.W);
        x = min(max(x, 0), P.W - 1);
        scol[i] = x >> 3;
        bcol[i] = ((x & 7) == 0 || (x & 7) == 7) ? 1 : 0;
    }
    // raw tile -> bufA
    for (int i = tid; i < PH * PW; i += K2_THREADS) {
        const int ly = i / PW, lx = i - ly * PW;
        int r = mirror_row(T.ty0 - M0 + ly, P.rows, P.has_top, P.has_bottom);
        r = min(max(r, P.has_top ? -JXLB200_HALO_ROWS : 0), P.rows - 1 + (P.has_bottom ? JXLB200_HALO_ROWS : 0));
        int x = mirror_col(T.tx0 - M0 + lx, P.W);
        x = min(max(x, 0), P.W - 1);
        const long long o = (long long)r * P.in_pitch + x;
        bufA[i] = __ldg(P.in[0] + o);
        bufA[PLANE + i] = __ldg(P.in[1] + o);
        bufA[2 * PLANE + i] = __ldg(P.in[2] + o);
    }
    __syncthreads();

    float *cur = bufA, *nxt = bufB;
    int m = M0;   // margin around the output tile that is valid in `cur`
    if (GAB) {
        m -= 1;
        const int rh = K2_TH + 2 * m, rw = K2_TW + 2 * m;
        for (int i = tid; i < rh * rw; i += K2_THREADS) {
            const int ly = M0 - m + i / rw, lx = M0 - m + i % rw;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float *R = cur + c * PLANE + ly * PW + lx;
                // Frame.java:535-537: operand order kept, uncontracted (bit-identical to the staged kernel away from edges)
                const float adj = __fadd_rn(__fadd_rn(__fadd_rn(R[-1], R[1]), R[-PW]), R[PW]);
                const float diag = __fadd_rn(__fadd_rn(__fadd_rn(R[-PW - 1], R[-PW + 1]), R[PW - 1]), R[PW + 1]);
                nxt[c * PLANE + ly * PW + lx] = __fadd_rn(__fadd_rn(__fmul_rn(P.gab_base[c], R[0]), __fmul_rn(P.gab_adj[c], adj)),
                                                          __fmul_rn(P.gab_diag[c], diag));
            }
        }
        __syncthreads();
        float *t = cur; cur = nxt; nxt = t;
    }

    // final store (colour transform applied) of one pixel of the output tile
    auto store_px = [&](int ly, int lx, float a, float b, float c) {
        const int oy = T.ty0 + ly - M0, ox = T.tx0 + lx - M0;
        if (oy < P.rows && ox < P.W) {
            color_px(P, a, b, c);
            const long long o = (long long)oy * P.out_pitch + ox;
            P.out[0][o] = a; P.out[1][o] = b; P.out[2][o] = c;
        }
    };

    if (ITERS == 0) {
        for (int i = tid; i < K2_TH * K2_TW; i += K2_THREADS) {
            const int ly = M0 + i / K2_TW, lx = M0 + i % K2_TW;
            store_px(ly, lx, cur[ly * PW + lx], cur[PLANE + ly * PW + lx], cur[2 * PLANE + ly * PW + lx]);
        }
        return;
    }

    // EPF passes: 2x2 blocks over the region with margin m - reach
#define K2_RUN_PASS(PASS, LAST)                                                                                          \
    {                                                                                                                    \
        m -= (PASS == 0 ? 3 : PASS == 1 ? 2 : 1);                                                                        \
        const int rh = K2_TH + 2 * m, rw = K2_TW + 2 * m, bw = rw / 2, nb = (rh / 2) * bw;                               \
        _Pragma("unroll 1") for (int b = tid; b < nb; b += K2_THREADS) {                                                 \
            const int ly = M0 - m + 2 * (b / bw), lx = M0 - m + 2 * (b % bw);                                            \
            if (LAST) {                                                                                                  \
                float v[4][3];                                                                                           \
                epf_block<PASS>(P, inv_sigma, T, cur, PLANE, PW, ly, lx, [&](int p, int c, float val) { v[p][c] = val; }); \
                _Pragma("unroll") for (int p = 0; p < 4; p++) store_px(ly + (p >> 1), lx + (p & 1), v[p][0], v[p][1], v[p][2]); \
            } else {                                                                                                     \
                float *o = nxt;                                                                                          \
                epf_block<PASS>(P, inv_sigma, T, cur, PLANE, PW, ly, lx,                                                 \
                                [&](int p, int c, float val) { o[c * PLANE + (ly + (p >> 1)) * PW + lx + (p & 1)] = val; }); \
            }                                                                                                            \
        }                                                                                                                \
        if (!(LAST)) {                                                                                                   \
            __syncthreads();                                                                                             \
            float *t = cur; cur = nxt; nxt = t;                                                                          \
        }                                                                                                                \
    }
    if (ITERS == 3) K2_RUN_PASS(0, false)
    if (ITERS >= 2) {
        K2_RUN_PASS(1, false)
        K2_RUN_PASS(2, true)
    } else {
        K2_RUN_PASS(1, true)
    }
#undef K2_RUN_PASS
}

struct jxlb200_ctx;
template <int GAB, int ITERS> static cudaError_t k2_fused_attr() {
    return cudaFuncSetAttribute(k2_fused<GAB, ITERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2Cfg<GAB, ITERS>::kBytes);
}
static inline cudaError_t k2_fused_init_all() {
    cudaError_t e;
    if ((e = k2_fused_attr<1, 0>()) != cudaSuccess) return e;
    if ((e = k2_fused_attr<1, 1>()) != cudaSuccess) return e;
    if ((e = k2_fused_attr<1, 2>()) != cudaSuccess) return e;
    if ((e = k2_fused_attr<1, 3>()) != cudaSuccess) return e;
    if ((e = k2_fused_attr<0, 1>()) != cudaSuccess) return e;
    if ((e = k2_fused_attr<0, 2>()) != cudaSuccess) return e;
    if ((e = k2_fused_attr<0, 3>()) != cudaSuccess) return e;
    return cudaSuccess;
}
// the fused kernel takes every frame that has at least one neighbourhood stage and is at least one reach tall / wide
static inline bool k2_fused_supported(const K2Params &K) { return (K.gab || K.iters > 0) && K.rows >= 8 && K.W >= 8 && !(K.rows & 7) && !(K.W & 7); }

template <int GAB, int ITERS> static void k2_fused_go(const K2Params &K, const float *inv_sigma, cudaStream_t st) {
    const dim3 grid((K.W + K2_TW - 1) / K2_TW, (K.rows + K2_TH - 1) / K2_TH);
    k2_fused<GAB, ITERS><<<grid, K2_THREADS, K2Cfg<GAB, ITERS>::kBytes, st>>>(K, inv_sigma);
}
static inline void k2_fused_dispatch(const K2Params &K, const float *inv_sigma, cudaStream_t st) {
    const int key = (K.gab ? 4 : 0) + K.iters;
    switch (key) {
    case 4: k2_fused_go<1, 0>(K, inv_sigma, st); break;
    case 5: k2_fused_go<1, 1>(K, inv_sigma, st); break;
    case 6: k2_fused_go<1, 2>(K, inv_sigma, st); break;
    case 7: k2_fused_go<1, 3>(K, inv_sigma, st); break;
    case 1: k2_fused_go<0, 1>(K, inv_sigma, st); break;
    case 2: k2_fused_go<0, 2>(K, inv_sigma, st); break;
    default: k2_fused_go<0, 3>(K, inv_sigma, st); break;
    }
}
