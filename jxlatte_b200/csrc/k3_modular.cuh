// k3_modular.cuh -- K3/K4/K5: Modular inverse transforms, int32 with Java semantics (wrapping add/sub, arithmetic >>,
// truncating / and %), bit-exact.
//
// Replaces (J/ = java/com/traneptora/jxlatte/ in the reference):
//   ModularStream.applyTransforms RCT      J/frame/modular/ModularStream.java:255-326  (+ permutationLut :35-38)
//   ModularStream.applyTransforms Palette  :327-378 (+ kDeltaPalette :20-33, ModularChannel.prediction J/frame/modular/ModularChannel.java:143-183)
//   ModularChannel.inverseHorizontalSqueeze / inverseVerticalSqueeze   ModularChannel.java:361-413, tendency :23-47
#pragma once
#include "common.cuh"

// Java int arithmetic wraps; do the adds in unsigned so C++ has no UB.
__device__ __forceinline__ int jadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int jsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int jmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }

__constant__ int c_perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};

// ---- K3: RCT ----
__global__ void k3_rct(int *c0, int *c1, int *c2, long long n, int type, int perm) {
    int *ch[3] = {c0, c1, c2};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int a = c0[i], b = c1[i], c = c2[i];
        switch (type) {
        case 1: c = jadd(c, a); break;
        case 2: b = jadd(b, a); break;
        case 3: c = jadd(c, a); b = jadd(b, a); break;
        case 4: b = jadd(b, jadd(a, c) >> 1); break;
        case 5: { const int ac = jadd(a, c); b = jadd(b, jadd(a, ac) >> 1); c = ac; break; }
        case 6: {
            const int tmp = jsub(a, c >> 1);
            const int f = jsub(tmp, b >> 1);
            a = jadd(f, b); const int nb = jadd(c, tmp); c = f; b = nb;
            break;
        }
        default: break;
        }
        ch[c_perm[perm][0]][i] = a;
        ch[c_perm[perm][1]][i] = b;
        ch[c_perm[perm][2]][i] = c;
    }
}

// ---- K5: Squeeze ----
__device__ __forceinline__ int tendency(int a, int b, int c) {   // ModularChannel.tendency :23-47
    if (a >= b && b >= c) {
        int x = jadd(jsub(jsub(jmul(4, a), jmul(3, c)), b), 6) / 12;
        const int d = jmul(2, jsub(a, b));
        const int e = jmul(2, jsub(b, c));
        if (jsub(x, x & 1) > d) x = jadd(d, 1);
        if (jadd(x, x & 1) > e) x = e;
        return x;
    }
    if (a <= b && b <= c) {
        int x = jsub(jsub(jsub(jmul(4, a), jmul(3, c)), b), 6) / 12;
        const int d = jmul(2, jsub(a, b));
        const int e = jmul(2, jsub(b, c));
        if (jadd(x, x & 1) < d) x = jsub(d, 1);
        if (jsub(x, x & 1) < e) x = e;
        return x;
    }
    return 0;
}

// vertical: one thread per column, walking down (coalesced across the warp)
__global__ void k5_squeeze_v(const int *__restrict__ avg, const int *__restrict__ res, int h_avg, int h_res, int w, int *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    int top = 0;
    int a = h_avg > 0 ? avg[x] : 0;
    for (int y = 0; y < h_res; y++) {
        const int residu = res[(size_t)y * w + x];
        const int nextAvg = y + 1 < h_avg ? avg[(size_t)(y + 1) * w + x] : a;
        if (y == 0) top = a;
        const int diff = jadd(residu, tendency(top, a, nextAvg));
        const int first = jadd(a, diff / 2);
        const int second = jsub(first, diff);
        out[(size_t)(2 * y) * w + x] = first;
        out[(size_t)(2 * y + 1) * w + x] = second;
        top = second;
        a = nextAvg;
    }
    if (h_avg > h_res) out[(size_t)(2 * h_res) * w + x] = avg[(size_t)h_res * w + x];
}

// horizontal: one thread per row; 32 rows x 32 columns staged through shared memory so global traffic is coalesced
__global__ void __launch_bounds__(32) k5_squeeze_h(const int *__restrict__ avg, const int *__restrict__ res, int h, int w_avg, int w_res,
                                                   int *__restrict__ out) {
    __shared__ int sa[32][34], sr[32][33], so[32][65];
    const int lane = threadIdx.x;
    const int row0 = blockIdx.x * 32;
    const int W = w_avg + w_res;
    int left = 0;
    for (int x0 = 0; x0 < w_res; x0 += 32) {
        const int n = min(32, w_res - x0);
        for (int r = 0; r < 32 && row0 + r < h; r++) {
            const size_t ro = (size_t)(row0 + r);
            if (x0 + lane < w_avg) sa[r][lane] = avg[ro * w_avg + x0 + lane];
            if (lane == 0 && x0 + 32 < w_avg) sa[r][32] = avg[ro * w_avg + x0 + 32];
            if (lane < n) sr[r][lane] = res[ro * w_res + x0 + lane];
        }
        __syncwarp();
        if (row0 + lane < h) {
            for (int i = 0; i < n; i++) {
                const int x = x0 + i;
                const int a = sa[lane][i];
                const int nextAvg = x + 1 < w_avg ? sa[lane][i + 1] : a;
                if (x == 0) left = a;
                const int diff = jadd(sr[lane][i], tendency(left, a, nextAvg));
                const int first = jadd(a, diff / 2);
                const int second = jsub(first, diff);
                so[lane][2 * i] = first;
                so[lane][2 * i + 1] = second;
                left = second;
            }
        }
        __syncwarp();
        for (int r = 0; r < 32 && row0 + r < h; r++) {
            const size_t ro = (size_t)(row0 + r) * W + 2 * x0;
            if (lane < 2 * n) out[ro + lane] = so[r][lane];
            if (lane + 32 < 2 * n) out[ro + lane + 32] = so[r][lane + 32];
        }
        __syncwarp();
    }
    if (w_avg > w_res && row0 + lane < h)
        out[(size_t)(row0 + lane) * W + 2 * w_res] = avg[(size_t)(row0 + lane) * w_avg + w_res];
}

// ---- K4: Palette ----
__constant__ short c_delta_palette[72][3] = {
    {0, 0, 0}, {4, 4, 4}, {11, 0, 0}, {0, 0, -13}, {0, -12, 0}, {-10, -10, -10},
    {-18, -18, -18}, {-27, -27, -27}, {-18, -18, 0}, {0, 0, -32}, {-32, 0, 0}, {-37, -37, -37},
    {0, -32, -32}, {24, 24, 45}, {50, 50, 50}, {-45, -24, -24}, {-24, -45, -45}, {0, -24, -24},
    {-34, -34, 0}, {-24, 0, -24}, {-45, -45, -24}, {64, 64, 64}, {-32, 0, -32}, {0, -32, 0},
    {-32, 0, 32}, {-24, -45, -24}, {45, 24, 45}, {24, -24, -45}, {-45, -24, 24}, {80, 80, 80},
    {64, 0, 0}, {0, 0, -64}, {0, -64, -64}, {-24, -24, 45}, {96, 96, 96}, {64, 64, 0},
    {45, -24, -24}, {34, -34, 0}, {112, 112, 112}, {24, -45, -45}, {45, 45, -24}, {0, -32, 32},
    {24, -24, 45}, {0, 96, 96}, {45, -24, 24}, {24, -45, -24}, {-24, -45, 24}, {0, -64, 0},
    {96, 0, 0}, {128, 128, 128}, {64, 0, 64}, {144, 144, 144}, {96, 96, 0}, {-36, -36, 36},
    {45, -24, -45}, {45, -45, -24}, {0, 0, -96}, {0, 128, 128}, {0, 96, 0}, {45, 24, -45},
    {-128, 0, 0}, {24, -45, 24}, {-45, 24, -45}, {64, 0, -64}, {64, -64, -64}, {96, 0, 96},
    {45, -45, 24}, {24, 45, -45}, {64, 64, -64}, {128, 128, 0}, {0, 0, -128}, {-24, 45, -45},
};

struct PalArgs {
    const int *idx;
    const int *palette;
    int h, w, num_c, nb_colors, nb_deltas, d_pred, bit_depth;
};

// value of one pixel before the delta prediction is added (ModularStream.java:341-366)
__device__ __forceinline__ int palette_value(const PalArgs &A, int c, int index) {
    const int bd = A.bit_depth;
    if (index >= 0 && index < A.nb_colors) return A.palette[(size_t)c * A.nb_colors + index];
    if (index >= A.nb_colors) {
        index -= A.nb_colors;
        const int maxv = (int)((1u << (bd & 31)) - 1u);
        if (index < 64)
            return jadd(jmul((index >> ((2 * c) & 31)) % 4, maxv) / 4, (int)(1u << ((bd - 3 > 0 ? bd - 3 : 0) & 31)));
        index -= 64;
        for (int k = 0; k < c; k++) index /= 5;
        return jmul(index % 5, maxv) / 4;
    }
    if (c < 3) {
        index = (-index - 1) % 143;
        int value = c_delta_palette[(index + 1) >> 1][c];
        if ((index & 1) == 0) value = -value;
        if (bd > 8) value = (int)((unsigned)value << (((bd < 24 ? bd : 24) - 8) & 31));
        return value;
    }
    return 0;
}

// phase 1: every pixel's base value, fully parallel; flags whether any pixel needs the predictor
__global__ void k4_palette_gather(PalArgs A, int c, int *__restrict__ out, int *__restrict__ any_delta) {
    const long long n = (long long)A.h * A.w;
    int local = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int index = A.idx[i];
        out[i] = palette_value(A, c, index);
        local |= index < A.nb_deltas;
    }
    if (__any_sync(0xffffffff, local) && (threadIdx.x & 31) == 0) *any_delta = 1;
}

struct Chan {
    const int *b;
    int w;
    __device__ __forceinline__ int at(int y, int x) const { return b[(size_t)y * w + x]; }
    __device__ __forceinline__ int west(int x, int y) const { return x > 0 ? at(y, x - 1) : y > 0 ? at(y - 1, x) : 0; }
    __device__ __forceinline__ int north(int x, int y) const { return y > 0 ? at(y - 1, x) : x > 0 ? at(y, x - 1) : 0; }
    __device__ __forceinline__ int northWest(int x, int y) const {
        return x > 0 ? (y > 0 ? at(y - 1, x - 1) : at(y, x - 1)) : (y > 0 ? at(y - 1, x) : 0);
    }
    __device__ __forceinline__ int northEast(int x, int y) const { return x + 1 < w && y > 0 ? at(y - 1, x + 1) : north(x, y); }
    __device__ __forceinline__ int northNorth(int x, int y) const { return y > 1 ? at(y - 2, x) : north(x, y); }
    __device__ __forceinline__ int northEastEast(int x, int y) const { return x + 2 < w && y > 0 ? at(y - 1, x + 2) : northEast(x, y); }
    __device__ __forceinline__ int westWest(int x, int y) const { return x > 1 ? at(y, x - 2) : west(x, y); }
};
__device__ __forceinline__ int jabs(int v) { return v < 0 ? (int)(0u - (unsigned)v) : v; }

// ModularChannel.prediction :143-183 (k == 6 is rejected on the host: the reference has no WP state for palette copies)
__device__ __forceinline__ int predict(const Chan &C, int y, int x, int k) {
    int n, v, nw, w;
    switch (k) {
    case 1: return C.west(x, y);
    case 2: return C.north(x, y);
    case 3: return jadd(C.west(x, y), C.north(x, y)) / 2;
    case 4:
        w = C.west(x, y); n = C.north(x, y); nw = C.northWest(x, y);
        return jabs(jsub(n, nw)) < jabs(jsub(w, nw)) ? w : n;
    case 5: {
        w = C.west(x, y); n = C.north(x, y);
        v = jsub(jadd(w, n), C.northWest(x, y));
        const int lower = n < w ? n : w, upper = lower ^ n ^ w;
        return v < lower ? lower : v > upper ? upper : v;
    }
    case 7: return C.northEast(x, y);
    case 8: return C.northWest(x, y);
    case 9: return C.westWest(x, y);
    case 10: return jadd(C.west(x, y), C.northWest(x, y)) / 2;
    case 11: return jadd(C.north(x, y), C.northWest(x, y)) / 2;
    case 12: return jadd(C.north(x, y), C.northEast(x, y)) / 2;
    case 13:
        return jadd(jadd(jadd(jadd(jadd(jsub(jmul(6, C.north(x, y)), jmul(2, C.northNorth(x, y))), jmul(7, C.west(x, y))),
                                   C.westWest(x, y)), C.northEastEast(x, y)), jmul(3, C.northEast(x, y))), 8) / 16;
    default: return 0;
    }
}

// phase 2: delta pixels in raster-dependency order.  Every predictor reads only W, WW, N, NN, NW, NE, NEE, so pixel
// (y, x) may run at step x + 3y: one CTA per channel sweeps that wavefront, thread i owning rows i, i + T, ...
__global__ void __launch_bounds__(1024) k4_palette_delta(PalArgs A, int *out, const int *__restrict__ any_delta) {
    if (*any_delta == 0) return;
    const int T = blockDim.x;
    Chan C{out, A.w};
    for (int band = 0; band < A.h; band += T) {
        const int y = band + threadIdx.x;
        const int rows = min(T, A.h - band);
        const int steps = A.w + 3 * (rows - 1);
        for (int t = 0; t < steps; t++) {
            const int x = t - 3 * (int)threadIdx.x;
            if (y < A.h && x >= 0 && x < A.w) {
                const size_t i = (size_t)y * A.w + x;
                if (A.idx[i] < A.nb_deltas) out[i] = jadd(out[i], predict(C, y, x, A.d_pred));
            }
            __syncthreads();
        }
    }
}
