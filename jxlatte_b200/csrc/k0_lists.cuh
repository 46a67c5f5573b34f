// k0_lists.cuh -- K0: turn the varblock partition (HFMetadata.dctSelect / blockList, J/frame/vardct/HFMetadata.java:38-53)
// into per-TransformType work lists on the device, so every stage-1 CTA runs one transform kind (no divergence).
// Also builds the chroma-from-luma gate that reproduces the reference's per-group factor cache
// (J/frame/vardct/HFCoefficients.java:155-185): a tile's factor only exists once the varblock covering the tile's
// top-left pixel has been visited, and varblocks are visited in raster order of their top-left (blockList order).
#pragma once
#include "common.cuh"

__constant__ TTInfo c_tt[27];
__constant__ int c_small_types[N_SMALL];
__constant__ int c_med_types[N_MED];
__constant__ int c_big_types[2][N_BIGC][3];   // [pass][line-length class] -> up to three types (-1 = none)

// pass 0 (columns, line length = pixelHeight) and pass 1 (rows, line length = pixelWidth) classes 32/64/128/256
static const int h_big_types[2][N_BIGC][3] = {
    {{20, -1, -1}, {18, 19, 23}, {21, 22, 26}, {24, 25, -1}},
    {{19, -1, -1}, {18, 20, 22}, {21, 23, 25}, {24, 26, -1}},
};

// a varblock must stay inside the frame and inside its 256x256 group (HFMetadata.placeBlock can place nothing else); the
// kernels trust this, so a caller's inconsistent maps are rejected here instead of being written out of bounds
// hb = block rows of ONE frame: a batch of equally sized frames may be stacked vertically (block row by = frame * hb + local row)
__device__ __forceinline__ bool k0_valid_origin(int t, int i, int wb, int hb) {
    const TTInfo tt = c_tt[t];
    const int by = (i / wb) % hb, bx = i % wb;
    return bx + tt.bw <= wb && by + tt.bh <= hb && (by & 31) + tt.bh <= 32 && (bx & 31) + tt.bw <= 32;
}

__global__ void k0_count(const uint8_t *__restrict__ ds, const uint8_t *__restrict__ bo, int ncells, int wb, int hb, Sched *s, int *__restrict__ err) {
    __shared__ int h[27];
    __shared__ int bad;
    if (threadIdx.x < 27) h[threadIdx.x] = 0;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += gridDim.x * blockDim.x) {
        const int t = ds[i];
        if (t > 26) {
            bad = 1;
        } else if (bo[i]) {
            if (!k0_valid_origin(t, i, wb, hb)) bad = 1;
            else atomicAdd(&h[t], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < 27 && h[threadIdx.x]) atomicAdd(&s->cnt[threadIdx.x], h[threadIdx.x]);
    if (threadIdx.x == 0 && bad) *err = 1;   // sticky across the slabs of a pipelined frame (Sched is reset per slab)
}

__global__ void k0_plan(Sched *s) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int off = 0;
    for (int t = 0; t < 27; t++) { s->start[t] = off; s->cursor[t] = 0; off += s->cnt[t]; }
    int cum = 0;
    for (int i = 0; i < N_SMALL; i++) { s->small_cum[i] = cum; cum += (s->cnt[c_small_types[i]] + SMALL_BATCH - 1) / SMALL_BATCH; }
    s->small_cum[N_SMALL] = cum;
    cum = 0;
    for (int i = 0; i < N_MED; i++) {
        const TTInfo tt = c_tt[c_med_types[i]];
        const int per = MED_COEFFS / (tt.bh * tt.bw * 64);
        s->med_cum[i] = cum;
        cum += (s->cnt[c_med_types[i]] + per - 1) / per;
    }
    s->med_cum[N_MED] = cum;
    for (int pass = 0; pass < 2; pass++)
        for (int k = 0; k < N_BIGC; k++) {
            cum = 0;
            for (int j = 0; j < 3; j++) {
                s->big_cum[pass][k][j] = cum;
                const int t = c_big_types[pass][k][j];
                if (t >= 0) {
                    const TTInfo tt = c_tt[t];
                    // pass 0 walks 32-column strips (bw * 8 / 32 of them), pass 1 walks 32-row strips; one item per channel
                    cum += s->cnt[t] * ((pass == 0 ? tt.bw : tt.bh) / 4) * 3;
                }
            }
            s->big_cum[pass][k][3] = cum;
        }
}

// items[start[t] + i] = (by << 16) | bx of the i-th varblock of type t (order within a type is arbitrary)
// Also fills the CfL gate: gate[tile] = raster index (by * wb + bx) of the origin of the varblock that covers the tile's
// top-left cell (tiles are 8x8 cells); the gate array is preset to 0x7f7f7f7f ("never visited").
__global__ void k0_scatter(const uint8_t *__restrict__ ds, const uint8_t *__restrict__ bo, int hb_all, int wb, int hb, Sched *s,
                           int *__restrict__ items, int *__restrict__ gate, int tw) {
    __shared__ int h[27];
    __shared__ int base[27];
    if (threadIdx.x < 27) h[threadIdx.x] = 0;
    __syncthreads();
    const int ncells = hb_all * wb;
    const int per = (ncells + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(ncells, lo + per);
    // two sweeps over this CTA's cell range: count, reserve one range per type, then place
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x)
        if (bo[i] && ds[i] <= 26 && k0_valid_origin(ds[i], i, wb, hb)) atomicAdd(&h[ds[i]], 1);
    __syncthreads();
    if (threadIdx.x < 27) {
        base[threadIdx.x] = h[threadIdx.x] ? atomicAdd(&s->cursor[threadIdx.x], h[threadIdx.x]) : 0;
        h[threadIdx.x] = 0;
    }
    __syncthreads();
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int t = ds[i];
        if (bo[i] && t <= 26 && k0_valid_origin(t, i, wb, hb)) {
            const int slot = s->start[t] + base[t] + atomicAdd(&h[t], 1);
            const int by = i / wb, bx = i % wb;
            items[slot] = (by << 16) | bx;
            const TTInfo tt = c_tt[t];
            for (int cy = (by + 7) & ~7; cy < by + tt.bh; cy += 8)
                for (int cx = (bx + 7) & ~7; cx < bx + tt.bw; cx += 8) gate[(cy >> 3) * tw + (cx >> 3)] = i;
        }
    }
}

// QM weights (HFGlobal.weights: [parameterIndex][c][matrixH][matrixW]) -> per TransformType, storage orientation
// [pixelH][pixelW]: wexp[type][c][y][x] = weights[param][c][flip ? x : y][flip ? y : x]
// (index swap of HFCoefficients.dequantizeHFCoefficients :312-314 done once here so stage 1 reads rows)
__global__ void k0_expand_weights(const float *__restrict__ w, const int *__restrict__ qm_off, DevTables tab, float *__restrict__ wexp) {
    const int t = blockIdx.y;
    const TTInfo tt = c_tt[t];
    const int H = tt.bh * 8, W = tt.bw * 8;
    const int mw = max(H, W);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * H * W; i += gridDim.x * blockDim.x) {
        const int c = i / (H * W), r = i % (H * W);
        const int y = r / W, x = r % W;
        const int wy = tt.flip ? x : y, wx = tt.flip ? y : x;
        wexp[tab.wexp_off[t] + i] = w[qm_off[tt.param * 3 + c] + wy * mw + wx];
    }
}
