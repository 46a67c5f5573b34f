// k2_stream.cuh -- K2 as a persistent STREAM of rows: Gaborish -> EPF pass 0 / 1 / 2 -> colour transform in one kernel, bit-identical
// to the reference (every float operation in the reference's order, uncontracted).
//
// Replaces Frame.performGabConvolution (J/frame/Frame.java:505-542), Frame.performEdgePreservingFilter (:544-679, epfDistance1
// :638-655, epfDistance2 :657-669, epfWeight :671-679) and JXLCodestreamDecoder.performColorTransforms
// (J/JXLCodestreamDecoder.java:256-283); J/ = java/com/traneptora/jxlatte/ in the reference tree.
//
// Shape of the computation (why it is not a tile kernel).  The filters are issue-bound, not HBM-bound, so the design
// minimises INSTRUCTIONS per pixel, then shared-memory wavefronts:
//   * a CTA (one per SM, persistent, 16 warps) walks column strips of the frame, 112 useful pixels wide (+8 halo columns each side =
//     128 = 32 lanes x 4 pixels), top to bottom, in "items" of CH rows + 8 halo rows each side; the rows of all its items form one
//     continuous stream of rows S = 0, 1, 2, ...  Nothing is recomputed vertically inside an item (a tile kernel recomputes its
//     7-row halo per tile: +29 % at 64x48).
//   * the stages hand rows to each other through ring buffers in shared memory.  A TICK advances every stage by 16 rows, one stage
//     after the other (PHASES separated by a CTA barrier), each stage a fixed number of rows behind its producer; inside a phase all
//     16 warps run the same code, one row (or one pair of distance-map rows) each.  [Round 2's first version ran the stages
//     CONCURRENTLY on dedicated warps; twelve loop bodies at once overflowed the 32 KB instruction cache -- profiles/r2_k2_stream_ncu.md.]
//         TMA (cp.async.bulk.tensor, 4-row boxes)  -> RAW ring          issued a tick ahead, lands behind an mbarrier
//         G   Gaborish (or copy)                   -> GAB ring          16 rows
//         D0  pass-0 distances, d in {(0,1),(1,0),(1,1),(1,-1),(0,2),(2,0)} -> six distance maps, 48 row pairs
//         W0  pass-0 weights, sums, divide         -> P0 ring           16 rows
//         D1  pass-1 distances, d in {(0,1),(1,0)} -> two distance maps 16 row pairs
//         W1  pass-1 weights, sums, divide         -> P1 ring           16 rows (the last stage when epf_iters == 1)
//         P2  pass 2 + colour transform + store    -> HBM               16 rows
//   * a lane owns a 4-pixel quad of a row; its neighbours' values come from the neighbour lanes by shuffle, so every row is read
//     with one conflict-free 128-bit shared load per channel and the 4-way bank conflicts of strided scalar loads never occur.
//   * what is shared without changing a rounding (as in k2_exact.cuh): the 15 terms of epfDistance1 are
//     T_{c,d}(q) = fl(fl|I_c(q) - I_c(q+d)| * s_c), formed once per row pair and position by the D phases (4 T rows for 2 map rows);
//     dist_{-d}(p) == dist_d(p-d) operation for operation, so six (two) maps serve the twelve (four) offsets; the centre tap has
//     weight exactly 1.  The ordered sums are literal.
//   * frame edges: MathHelper.mirrorCoordinate (J/util/MathHelper.java:323-329) applies to a stage's OUTPUT.  Rows above the
//     frame are written behind the producer when their mirror source row is produced (they lie in the item's own upper halo);
//     rows below are copied from behind after the phase's barrier; columns are mirrored across lanes before the store.
//
// The same source compiles for the host (K2S_HOST_EMU, tests/host/k2_stream_host.cpp): 32 host threads per warp, the shuffles and
// barriers emulated, so tests/test_k2_stream_host.py holds the whole stream -- ring arithmetic, lags, mirror handling, operation
// order -- to bit-equality with the oracle on a CPU-only box.
#pragma once
#include <math.h>
#include <stdint.h>
#include "common.cuh"

#ifdef K2S_HOST_EMU
#define K2S_FN static inline
#define K2S_ADD(a, b) ((a) + (b))   /* host build uses -ffp-contract=off */
#define K2S_SUB(a, b) ((a) - (b))
#define K2S_MUL(a, b) ((a) * (b))
static inline float k2s_host_subsat(float a, float b) { const float t = a - b; return t > 0.0f ? (t > 1.0f ? 1.0f : t) : 0.0f; }
#define K2S_SUBSAT(a, b) k2s_host_subsat((a), (b))
#define K2S_RCP(b) (0.0f)             /* the host divides; the device shares one refined reciprocal per pixel (k2s_div12) */
#define K2S_LDG(p) (*(p))
#else
#include <cuda.h>
#include "k2_exact.cuh"             /* kx_rcp_refined / kx_div_fast: the three divides of a pixel share one refined reciprocal */
#define K2S_FN __device__ __forceinline__
#define K2S_ADD(a, b) __fadd_rn((a), (b))
#define K2S_SUB(a, b) __fsub_rn((a), (b))
#define K2S_MUL(a, b) __fmul_rn((a), (b))
__device__ __forceinline__ float k2s_subsat(float a, float b) { float r; asm("sub.rn.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
#define K2S_SUBSAT(a, b) k2s_subsat((a), (b))
#define K2S_RCP(b) kx_rcp_refined(b)
#define K2S_LDG(p) __ldg(p)
#endif

// trip count of the channel loops that must stay rolled: read from a kernel argument, or the compiler unrolls them whatever the pragma says
#define K2S_ROLLED_TRIPS(P) ((P).W > 0 ? 3 : 0)
#define K2S_TW 112            /* useful columns of a column strip */
#define K2S_HALO 8            /* halo columns each side and halo rows each side of an item (7 used: 1 + 3 + 2 + 1) */
#define K2S_PITCH 128         /* floats per ring row: 32 lanes x 4 */
#ifndef K2S_BAND
#define K2S_BAND 16           /* rows per tick = warps per CTA: one row (or one pair of map rows) per warp and phase.  16: one CTA per SM
                                 (8K frame, stage 2: 1.32 ms); 8: two CTAs of 8 warps per SM, one computes while the other sits at a
                                 barrier (no faster: the barriers are not what limits it); 18: 18 warps at 96 registers (1.52 ms) */
#endif
#define K2S_NWARPS K2S_BAND
#define K2S_CTAS_PER_SM (K2S_BAND == 8 ? 2 : 1)
// Ring sizes in rows.  Within a tick the writer of a ring (first row of its band: BAND t + bw) runs before its readers (BAND t + br, reading
// up to `a` rows above their own); when the writer's band lands, the readers of the same tick still need everything from row
// BAND t + br - a on, so a ring holds BAND + bw - br + a rows:
//   GAB: G -2, W0 -6 reads 2 above -> BAND + 6      D0 maps: D0 -6, W0 -6 reads 2 above -> BAND + 2
//   P0:  W0 -6, D1 / W1 -8 read 1 above -> BAND + 3  D1 maps: D1 -8, W1 -8 reads 1 above -> BAND + 1
//   P1:  W1 -8, P2 -9 reads 1 above -> BAND + 2      RAW: next band's loads are issued after G; G -2 reads 1 above -> BAND + 3, in boxes of 4
#define K2S_BOX (K2S_BAND % 4 == 0 ? 4 : 2)      /* rows per TMA box: divides the band and the 8-row granularity of the items */
#define K2S_RS_RAW ((K2S_BAND + 3 + K2S_BOX - 1) / K2S_BOX * K2S_BOX)
#define K2S_RS_GAB (K2S_BAND + 6)
#define K2S_RS_P0 (K2S_BAND + 3)
#define K2S_RS_P1 (K2S_BAND + 2)
#define K2S_RS_D0 (K2S_BAND + 2)
#define K2S_RS_D1 (K2S_BAND + 1)
#define K2S_OFF_RAW 0
#define K2S_OFF_GAB (K2S_OFF_RAW + 3 * K2S_RS_RAW * K2S_PITCH)
#define K2S_OFF_P0 (K2S_OFF_GAB + 3 * K2S_RS_GAB * K2S_PITCH)
#define K2S_OFF_P1 (K2S_OFF_P0 + 3 * K2S_RS_P0 * K2S_PITCH)
#define K2S_OFF_D0 (K2S_OFF_P1 + 3 * K2S_RS_P1 * K2S_PITCH)
#define K2S_OFF_D1 (K2S_OFF_D0 + 6 * K2S_RS_D0 * K2S_PITCH)
#define K2S_FLOATS (K2S_OFF_D1 + 2 * K2S_RS_D1 * K2S_PITCH)
#define K2S_BYTES (K2S_FLOATS * 4 + 64)      /* + two mbarriers */

// mirror margins a stage's output must carry outside the frame for the stages after it
#define K2S_MARGIN_GAB 3      /* pass 0 evaluated inside the frame reads 3 rows / columns beyond it (offset 2 + plus-shaped patch 1) */
#define K2S_MARGIN_P0 2
#define K2S_MARGIN_P1 1

struct K2SArgs {
    K2Params P;
    const float *inv_sigma;     // 1/sigma per 8x8 block, points at the slab's first own block row (k2_sigma)
    long long zpx;              // pixels between frames of a vertical stack (input and output planes)
    int zblk;                   // sigma entries between frames
    int n_frames;
    int ch;                     // rows an item produces (multiple of 8)
    int ir;                     // rows an item streams = ch + 16
    int n_cols, n_chunks;       // column strips per frame, items per column strip
    int n_items;                // n_frames * n_cols * n_chunks
    int tma_row0;               // tensor-map row of frame row 0 (8 when the slab has rows above it)
};

// In tick t a stage processes stream rows [BAND t + base, BAND t + base + BAND); the TMA loads of tick t bring rows [BAND t, BAND t + BAND).
// A stage trails its producer by the rows below its own that it reads, and by one more where a row ABOVE the frame is involved: such a
// row exists only once its mirror source has been produced (row -k is written together with row k-1).  D0 evaluates map rows down to
// y = -2 (pass 0 at row 0 uses dist_(2,0) at row -2), which read GAB rows down to -3, written with GAB row 2 = y + 4.
//   G reads RAW rows S-1 .. S+1                        -> G  = -2
//   D0 reads GAB rows S-1 .. S+3 (and see above)       -> D0 = G - 4;  W0 reads GAB S-2 .. S+2 and the maps S-2 .. S -> W0 = D0
//   D1 reads P0 rows S-1 .. S+2                        -> D1 = W0 - 2; W1 = D1
//   P2 reads P1 rows S-1 .. S+1                        -> P2 = W1 - 1
// epf_iters < 3: pass 1 reads the GAB ring: D1 = W1 = G - 2.
template <int ITERS> struct K2SCfg;
template <> struct K2SCfg<3> { static constexpr int G = -2, D0 = -6, W0 = -6, D1 = -8, W1 = -8, P2 = -9, LAST = -9; };
template <> struct K2SCfg<2> { static constexpr int G = -2, D0 = 0, W0 = 0, D1 = -4, W1 = -4, P2 = -5, LAST = -5; };
template <> struct K2SCfg<1> { static constexpr int G = -2, D0 = 0, W0 = 0, D1 = -4, W1 = -4, P2 = 0, LAST = -4; };

// ------------------------------------------------------------------------------------------------------------------------
// platform layer: lane id, shuffles, CTA barrier, TMA.  Host versions live in tests/host/k2_stream_host.cpp.
// ------------------------------------------------------------------------------------------------------------------------
#ifdef K2S_HOST_EMU
int k2s_emu_tid();
int k2s_emu_cta();
int k2s_emu_grid();
float k2s_emu_shfl(float v, int src_lane);
int k2s_emu_any(int pred);
void k2s_emu_sync();
int k2s_emu_sync_or(int pred);
struct K2STmap { const float *base; int w, rows; long long pitch; };
K2S_FN int k2s_tid() { return k2s_emu_tid(); }
K2S_FN int k2s_cta() { return k2s_emu_cta(); }
K2S_FN int k2s_grid() { return k2s_emu_grid(); }
K2S_FN float k2s_up(float v) { const int l = k2s_emu_tid() & 31; return k2s_emu_shfl(v, l > 0 ? l - 1 : l); }
K2S_FN float k2s_dn(float v) { const int l = k2s_emu_tid() & 31; return k2s_emu_shfl(v, l < 31 ? l + 1 : l); }
K2S_FN float k2s_from(float v, int src) { return k2s_emu_shfl(v, src); }
K2S_FN bool k2s_any(bool pred) { return k2s_emu_any(pred) != 0; }
K2S_FN void k2s_sync() { k2s_emu_sync(); }
K2S_FN int k2s_sync_or(int pred) { return k2s_emu_sync_or(pred); }
struct K2SQuad { float x, y, z, w; };
K2S_FN K2SQuad k2s_ld4(const float *p) { return K2SQuad{p[0], p[1], p[2], p[3]}; }
K2S_FN void k2s_st4(float *p, K2SQuad q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
K2S_FN void k2s_stg4(float *p, K2SQuad q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
// a 128 x 4 box at (x, y): zero fill outside the tensor, exactly what the TMA unit writes
K2S_FN void k2s_tma_box(float *dst, const K2STmap &m, int x, int y) {
    for (int r = 0; r < K2S_BOX; r++)
        for (int i = 0; i < K2S_PITCH; i++) {
            const int yy = y + r, xx = x + i;
            dst[r * K2S_PITCH + i] = (yy >= 0 && yy < m.rows && xx >= 0 && xx < m.w) ? m.base[(long long)yy * m.pitch + xx] : 0.0f;
        }
}
#define K2S_TMAP_PARAM const K2STmap &
#else
typedef CUtensorMap K2STmap;
typedef float4 K2SQuad;
K2S_FN int k2s_tid() { return threadIdx.x; }
K2S_FN int k2s_cta() { return blockIdx.x; }
K2S_FN int k2s_grid() { return gridDim.x; }
K2S_FN float k2s_up(float v) { return __shfl_up_sync(0xffffffffu, v, 1); }
K2S_FN float k2s_dn(float v) { return __shfl_down_sync(0xffffffffu, v, 1); }
K2S_FN float k2s_from(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
K2S_FN bool k2s_any(bool pred) { return __any_sync(0xffffffffu, pred) != 0; }
K2S_FN void k2s_sync() { __syncthreads(); }
K2S_FN int k2s_sync_or(int pred) { return __syncthreads_or(pred); }
K2S_FN K2SQuad k2s_ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
K2S_FN void k2s_st4(float *p, K2SQuad q) { *reinterpret_cast<float4 *>(p) = q; }
K2S_FN void k2s_stg4(float *p, K2SQuad q) { *reinterpret_cast<float4 *>(p) = q; }
K2S_FN uint32_t k2s_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
K2S_FN void k2s_mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k2s_smem_u32(bar)), "r"(count) : "memory");
}
K2S_FN void k2s_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k2s_smem_u32(bar)), "r"(bytes) : "memory");
}
K2S_FN void k2s_mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = k2s_smem_u32(bar);
    uint32_t done = 0;
    // bounded (about two seconds): a TMA that never completes (bad descriptor) must fault, not hang the box
    const long long t0 = clock64();
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000ll) __trap();
    }
}
K2S_FN void k2s_tma_box_async(float *dst, const K2STmap *m, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(k2s_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(k2s_smem_u32(bar)) : "memory");
}
#define K2S_TMAP_PARAM const K2STmap *
#endif

K2S_FN float k2s_get(const K2SQuad &q, int j) { return j == 0 ? q.x : j == 1 ? q.y : j == 2 ? q.z : q.w; }
K2S_FN int k2s_slot(int S, int rs) { int s = S % rs; return s < 0 ? s + rs : s; }
// the neighbouring slots of one that is known: a compare and a select instead of a division by a constant
K2S_FN int k2s_next(int s, int rs) { return s + 1 == rs ? 0 : s + 1; }
K2S_FN int k2s_prev(int s, int rs) { return s == 0 ? rs - 1 : s - 1; }

// ------------------------------------------------------------------------------------------------------------------------
// where a stream row lies: item -> (frame of the stack, column strip, chunk) -> frame row / first column
// ------------------------------------------------------------------------------------------------------------------------
struct K2SRow {
    int y;        // frame (slab) row; may lie outside [0, rows)
    int x0;       // first useful column of the strip (lane 2's first pixel)
    int z;        // frame of the stack
    int loc;      // row inside the item, 0 .. ir-1 (8 .. 8+ch-1 are the rows the item produces)
    int valid;    // the CTA has such an item
};
#ifdef K2S_HOST_EMU
#define K2S_COLD static inline
#else
#define K2S_COLD __device__ __noinline__      /* runs once per item and stage: kept out of line, every row function would carry a copy */
#endif
K2S_COLD K2SRow k2s_locate(const K2SArgs &A, int S) {
    K2SRow r;
    const int k = S / A.ir;
    r.loc = S - k * A.ir;
    const int item = k2s_cta() + k * k2s_grid();
    r.valid = item < A.n_items;
    const int per_frame = A.n_cols * A.n_chunks;
    r.z = item / per_frame;
    const int rem = item - r.z * per_frame;
    const int col = rem / A.n_chunks, chunk = rem - col * A.n_chunks;   // chunk fastest: a CTA's consecutive items are not neighbours anyway
    r.x0 = col * K2S_TW;
    r.y = chunk * A.ch - K2S_HALO + r.loc;
    return r;
}
// the divisions above are paid once per item, not once per row: a warp keeps the item its last row was in
struct K2SCursor { int lo, hi, x0, z, y_lo, valid; };
K2S_FN void k2s_cursor_init(K2SCursor &c) { c.lo = 0; c.hi = 0; c.x0 = 0; c.z = 0; c.y_lo = 0; c.valid = 0; }
K2S_FN K2SRow k2s_at(const K2SArgs &A, K2SCursor &c, int S) {
    if (S < c.lo || S >= c.hi) {
        const K2SRow r = k2s_locate(A, S);
        c.lo = S - r.loc; c.hi = c.lo + A.ir; c.x0 = r.x0; c.z = r.z; c.y_lo = r.y - r.loc; c.valid = r.valid;
    }
    K2SRow r;
    r.loc = S - c.lo; r.y = c.y_lo + r.loc; r.x0 = c.x0; r.z = c.z; r.valid = c.valid && S >= 0;
    return r;
}
// 0: the stage evaluates this row; 1: nothing to do (above the frame: filled when its mirror source is produced; or no such
// row anywhere); 2: below the frame: copy of the mirror row 2*rows-1-y from behind
K2S_FN int k2s_row_kind(const K2Params &P, int y, int margin) {
    if (y < 0) return (P.has_top && y >= -JXLB200_HALO_ROWS) ? 0 : 1;
    if (y >= P.rows) {
        if (P.has_bottom) return y < P.rows + JXLB200_HALO_ROWS ? 0 : 1;
        return y < P.rows + margin ? 2 : 1;
    }
    return 0;
}

// one row of a stage's output -> its ring: columns outside the frame are mirrored across lanes first; a row whose mirror
// image lies above the frame is also written there (behind the producer, into the item's own upper halo)
K2S_FN void k2s_emit(const K2Params &P, float *ring, int rs, int margin, const K2SRow &R, int S, int lane, K2SQuad q[3]) {
    const int nl = (P.W - R.x0 + K2S_HALO) >> 2;              // lanes whose columns lie left of the frame's right edge
    if (R.x0 == 0 || nl < 32) {                               // strip touches a frame edge (uniform over the warp)
        int src = lane;
        bool rev = false;
        if (R.x0 == 0 && lane < 2) { src = 3 - lane; rev = true; }
        else if (lane >= nl) { src = 2 * nl - 1 - lane; if (src < 0) src = 0; rev = true; }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float a = k2s_from(q[c].x, src), b = k2s_from(q[c].y, src), d = k2s_from(q[c].z, src), e = k2s_from(q[c].w, src);
            if (rev) { q[c].x = e; q[c].y = d; q[c].z = b; q[c].w = a; }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) k2s_st4(ring + (c * rs + k2s_slot(S, rs)) * K2S_PITCH + 4 * lane, q[c]);
    if (!P.has_top && R.y >= 0 && R.y < margin) {
        const int Sm = S - (2 * R.y + 1);                     // frame row -1 - y
#pragma unroll
        for (int c = 0; c < 3; c++) k2s_st4(ring + (c * rs + k2s_slot(Sm, rs)) * K2S_PITCH + 4 * lane, q[c]);
    }
}
// a row below the frame: the ring row of its mirror image, verbatim (that row was mirrored across lanes when it was emitted)
K2S_FN void k2s_copy_behind(const K2Params &P, float *ring, int rs, const K2SRow &R, int S, int lane) {
    const int Sm = S - (2 * (R.y - P.rows) + 1);
#pragma unroll
    for (int c = 0; c < 3; c++)
        k2s_st4(ring + (c * rs + k2s_slot(S, rs)) * K2S_PITCH + 4 * lane, k2s_ld4(ring + (c * rs + k2s_slot(Sm, rs)) * K2S_PITCH + 4 * lane));
}

// OpsinInverseMatrix.invertXYB (J/color/OpsinInverseMatrix.java:128-138) and the YCbCr branch of performColorTransforms
// (J/JXLCodestreamDecoder.java:270-282): the reference's operation order, uncontracted (same as color_px in k2_restore.cuh)
K2S_FN void k2s_color(const K2Params &P, float &a, float &b, float &c) {
    if (P.color_mode & 1) {
        const float gl = K2S_ADD(K2S_ADD(b, a), P.cob[0]), gm = K2S_ADD(K2S_SUB(b, a), P.cob[1]), gs = K2S_ADD(c, P.cob[2]);
        const float ml = K2S_ADD(K2S_MUL(K2S_MUL(gl, gl), gl), P.ob[0]);
        const float mm = K2S_ADD(K2S_MUL(K2S_MUL(gm, gm), gm), P.ob[1]);
        const float ms = K2S_ADD(K2S_MUL(K2S_MUL(gs, gs), gs), P.ob[2]);
        a = K2S_ADD(K2S_ADD(K2S_MUL(P.m[0], ml), K2S_MUL(P.m[1], mm)), K2S_MUL(P.m[2], ms));
        b = K2S_ADD(K2S_ADD(K2S_MUL(P.m[3], ml), K2S_MUL(P.m[4], mm)), K2S_MUL(P.m[5], ms));
        c = K2S_ADD(K2S_ADD(K2S_MUL(P.m[6], ml), K2S_MUL(P.m[7], mm)), K2S_MUL(P.m[8], ms));
    }
    if (P.color_mode & 2) {
        const float cb = a, yh = K2S_ADD(b, 0.50196078431372549019f), cr = c;
        a = K2S_ADD(yh, K2S_MUL(1.402f, cr));
        b = K2S_SUB(K2S_SUB(yh, K2S_MUL(0.34413628620102214650f, cb)), K2S_MUL(0.71413628620102214650f, cr));
        c = K2S_ADD(yh, K2S_MUL(1.772f, cb));
    }
}
// the last stage's row: colour transform and 128-bit stores of the strip's useful columns
K2S_FN void k2s_final(const K2SArgs &A, const K2SRow &R, int lane, K2SQuad q[3]) {
    const K2Params &P = A.P;
    if (R.loc < K2S_HALO || R.loc >= K2S_HALO + A.ch || R.y < 0 || R.y >= P.rows) return;
    const int col = R.x0 - K2S_HALO + 4 * lane;
    if (lane < 2 || lane >= 30 || col >= P.W) return;
    k2s_color(P, q[0].x, q[1].x, q[2].x); k2s_color(P, q[0].y, q[1].y, q[2].y);
    k2s_color(P, q[0].z, q[1].z, q[2].z); k2s_color(P, q[0].w, q[1].w, q[2].w);
    const long long o = (long long)R.z * A.zpx + (long long)R.y * P.out_pitch + col;
#pragma unroll
    for (int c = 0; c < 3; c++) k2s_stg4(P.out[c] + o, q[c]);
}

// epfWeight (Frame.java:671-679): m = borderSadMul on block-border pixels, else 1 (x * 1 is exact)
// epfWeight's clamp v < 0 ? 0 : v with v = 1 - x.  SAT: ONE saturating subtract (FADD.SAT) -- when x >= 0, v <= 1 and clamping at 1
// changes nothing (a NaN becomes 0 either way).  x is a sum of magnitudes times scales k2_stream_supported checks to be non-negative,
// times 1/sigma: the row functions take the SAT form unless a lane of the warp has 1/sigma < 0 (an HF multiplier <= 0, which a crafted
// stream can carry: HFMetadata.java:49), and the literal form for such a row.
template <bool SAT> K2S_FN float k2s_wgt(float dist, float m, float ss, float is) {
    const float x = K2S_MUL(K2S_MUL(K2S_MUL(dist, m), ss), is);
    if (SAT) return K2S_SUBSAT(1.0f, x);
    const float v = K2S_SUB(1.0f, x);
    return v < 0.0f ? 0.0f : v;                               // Frame.java:677-678, literally
}
// the twelve divides of a row quad (three channels), a[c][j] / b[j] with r[j] = K2S_RCP(b[j]): ONE range test for the twelve numerators
// (kx_div_shared's fast path is exact inside it, k2_exact.cuh) instead of a test, a branch and a reconvergence point per divide, and
// all the sums are complete before the first divide starts
K2S_FN void k2s_div12(const float a[3][4], const float b[4], const float r[4], float o[3][4]) {
#ifdef K2S_HOST_EMU
    for (int c = 0; c < 3; c++)
        for (int j = 0; j < 4; j++) o[c][j] = a[c][j] / b[j];
#else
    // smallest and largest magnitude of the twelve (a NaN numerator drops out of both and takes the fast path, where it stays a NaN)
    float lo = fabsf(a[0][0]), hi = lo;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (c + j == 0) continue;
            lo = fminf(lo, fabsf(a[c][j]));
            hi = fmaxf(hi, fabsf(a[c][j]));
        }
    if (lo >= 7.8886090522101181e-31f && hi <= 1.2676506002282294e30f) {        // kx_div_fast_ok for all of them
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int j = 0; j < 4; j++) o[c][j] = kx_div_fast(a[c][j], b[j], r[j]);
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int j = 0; j < 4; j++) o[c][j] = __fdiv_rn(a[c][j], b[j]);
    }
#endif
}
// 1/sigma of the lane's block in row y, and the border multipliers of its four pixels
K2S_FN float k2s_sigma(const K2SArgs &A, const K2SRow &R, int lane, float m[4]) {
    const K2Params &P = A.P;
    int col = R.x0 - K2S_HALO + 4 * lane;
    const bool rowb = (R.y & 7) == 0 || (R.y & 7) == 7;
    const int cb = col & 7;                                   // 0 or 4
    m[0] = (rowb || cb == 0) ? P.border_mul : 1.0f;
    m[1] = rowb ? P.border_mul : 1.0f;
    m[2] = m[1];
    m[3] = (rowb || cb == 4) ? P.border_mul : 1.0f;
    col = col < 0 ? 0 : (col >= P.W ? P.W - 1 : col);         // lanes outside the frame produce values nobody keeps
    return K2S_LDG(A.inv_sigma + (R.z * A.zblk + (R.y >> 3) * P.wb + (col >> 3)));     // 32-bit index: k2_stream_supported checks the range
}

// ------------------------------------------------------------------------------------------------------------------------
// G: Gaborish (Frame.java:505-542) of stream row S into the GAB ring: rows S-1, S, S+1 of the RAW ring, west / east neighbours by
// shuffle.  Edge = clamp on the padded plane (:526-534): at the frame's first / last row the row itself stands in for the missing
// one, at the first / last column the lane's own outer pixel.  Returns the row's kind (k2s_row_kind).
// ------------------------------------------------------------------------------------------------------------------------
template <int GAB> K2S_FN int k2s_g_row(const K2SArgs &A, float *sm, K2SCursor &cur, int S, int lane) {
    const K2Params &P = A.P;
    const K2SRow R = k2s_at(A, cur, S);
    if (!R.valid) return 1;
    const int kind = k2s_row_kind(P, R.y, K2S_MARGIN_GAB);
    if (kind != 0) return kind;
    const float *raw = sm + K2S_OFF_RAW + 4 * lane;
    float *gab = sm + K2S_OFF_GAB;
    const int sr = k2s_slot(S, K2S_RS_RAW);
    K2SQuad out[3];
    if (!GAB) {
#pragma unroll
        for (int c = 0; c < 3; c++) out[c] = k2s_ld4(raw + (c * K2S_RS_RAW + sr) * K2S_PITCH);
    } else {
        const int col = R.x0 - K2S_HALO + 4 * lane;
        const bool first_col = col == 0, last_col = col + 4 == P.W;
        const bool clamp_n = R.y == 0 && !P.has_top, clamp_s = R.y == P.rows - 1 && !P.has_bottom;
        const int sn = clamp_n ? sr : (sr == 0 ? K2S_RS_RAW - 1 : sr - 1), ss = clamp_s ? sr : (sr + 1 == K2S_RS_RAW ? 0 : sr + 1);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const K2SQuad qn = k2s_ld4(raw + (c * K2S_RS_RAW + sn) * K2S_PITCH), qr = k2s_ld4(raw + (c * K2S_RS_RAW + sr) * K2S_PITCH);
            const K2SQuad qs = k2s_ld4(raw + (c * K2S_RS_RAW + ss) * K2S_PITCH);
            float n6[6] = {k2s_up(qn.w), qn.x, qn.y, qn.z, qn.w, k2s_dn(qn.x)};
            float r6[6] = {k2s_up(qr.w), qr.x, qr.y, qr.z, qr.w, k2s_dn(qr.x)};
            float s6[6] = {k2s_up(qs.w), qs.x, qs.y, qs.z, qs.w, k2s_dn(qs.x)};
            if (first_col) { n6[0] = n6[1]; r6[0] = r6[1]; s6[0] = s6[1]; }
            if (last_col) { n6[5] = n6[4]; r6[5] = r6[4]; s6[5] = s6[4]; }
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // Frame.java:535-537: operand order kept, uncontracted
                const float adj = K2S_ADD(K2S_ADD(K2S_ADD(r6[j], r6[j + 2]), n6[j + 1]), s6[j + 1]);
                const float diag = K2S_ADD(K2S_ADD(K2S_ADD(n6[j], n6[j + 2]), s6[j]), s6[j + 2]);
                o[j] = K2S_ADD(K2S_ADD(K2S_MUL(P.gab_base[c], r6[j + 1]), K2S_MUL(P.gab_adj[c], adj)), K2S_MUL(P.gab_diag[c], diag));
            }
            out[c].x = o[0]; out[c].y = o[1]; out[c].z = o[2]; out[c].w = o[3];
        }
    }
    k2s_emit(P, gab, K2S_RS_GAB, K2S_MARGIN_GAB, R, S, lane, out);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// D: rows S0 and S0+1 of the distance map of one canonical offset d = (DY, DX) (epfDistance1, Frame.java:638-655).
// T(q) = fl(fl|I(q) - I(q+d)| * s_c) for the four rows S0-1 .. S0+2 (input rows S0-1 .. S0+2+DY, one 128-bit load each; the shifted
// row's columns come from the neighbour lanes), then dist(y) = sum over c of T(y,x) + T(y,x-1) + T(y,x+1) + T(y-1,x) + T(y+1,x), in
// that order.  Every stream row is evaluated, inside the frame or not: the input ring carries the mirrored rows and columns.
// ------------------------------------------------------------------------------------------------------------------------
template <int DY, int DX> K2S_FN void k2s_d_pair(const K2Params &P, const float *in, int rs_in, float *outmap, int rs_out, int S0, int lane) {
    float dist[2][4];
    int slot[4 + DY];
    slot[0] = k2s_slot(S0 - 1, rs_in);
#pragma unroll
    for (int i = 1; i < 4 + DY; i++) slot[i] = k2s_next(slot[i - 1], rs_in);
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int j = 0; j < 4; j++) dist[k][j] = 0.0f;
    // the channel loop stays rolled (trip count read from a kernel argument, or the compiler unrolls it whatever the pragma says): the
    // D0 phase runs six instances of this body at once, and unrolled they are two instruction caches' worth of code
    const int nch = K2S_ROLLED_TRIPS(P);
#pragma unroll
    for (int i = 0; i < 4 + DY; i++) slot[i] = slot[i] * K2S_PITCH + 4 * lane;       // float offset of the lane's quad in the row
#pragma unroll 1
    for (int c = 0; c < nch; c++) {
        const float *pl = in + c * rs_in * K2S_PITCH;
        float I[4 + DY][4];
#pragma unroll
        for (int i = 0; i < 4 + DY; i++) {
            const K2SQuad q = k2s_ld4(pl + slot[i]);
            I[i][0] = q.x; I[i][1] = q.y; I[i][2] = q.z; I[i][3] = q.w;
        }
        const float s = P.ch_scale[c];
        float T[4][4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const float *b = I[r], *v = I[r + DY];
            float sh[4];
            if (DX == 0) { sh[0] = v[0]; sh[1] = v[1]; sh[2] = v[2]; sh[3] = v[3]; }
            else if (DX == 1) { sh[0] = v[1]; sh[1] = v[2]; sh[2] = v[3]; sh[3] = k2s_dn(v[0]); }
            else if (DX == 2) { sh[0] = v[2]; sh[1] = v[3]; sh[2] = k2s_dn(v[0]); sh[3] = k2s_dn(v[1]); }
            else { sh[0] = k2s_up(v[3]); sh[1] = v[0]; sh[2] = v[1]; sh[3] = v[2]; }
#pragma unroll
            for (int j = 0; j < 4; j++) T[r][j] = K2S_MUL(fabsf(K2S_SUB(b[j], sh[j])), s);
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const float *t = T[k + 1];
            const float t6[6] = {k2s_up(t[3]), t[0], t[1], t[2], t[3], k2s_dn(t[0])};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float d = K2S_ADD(dist[k][j], t6[j + 1]);                      // (0, 0); the running sum starts at 0f like the Java's
                d = K2S_ADD(d, t6[j]);                                         // (0, -1)
                d = K2S_ADD(d, t6[j + 2]);                                     // (0, +1)
                d = K2S_ADD(d, T[k][j]);                                       // (-1, 0)
                d = K2S_ADD(d, T[k + 2][j]);                                   // (+1, 0)
                dist[k][j] = d;
            }
        }
    }
    int so = k2s_slot(S0, rs_out);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        K2SQuad q; q.x = dist[k][0]; q.y = dist[k][1]; q.z = dist[k][2]; q.w = dist[k][3];
        k2s_st4(outmap + so * K2S_PITCH + 4 * lane, q);
        so = so + 1 == rs_out ? 0 : so + 1;
    }
}

// Pass 0: THREE maps of the same row pair in one task.  SET 0 = offsets (0,1), (1,1), (0,2) -> maps 0, 2, 4; SET 1 = (1,0), (1,-1), (2,0)
// -> maps 1, 3, 5.  The maps of a set read the same input rows, so each row is loaded once per channel and its east (west) neighbours
// are fetched once for all three; arithmetic and its order are those of k2s_d_pair, map by map.
template <int SET> K2S_FN void k2s_d_triple(const K2Params &P, const float *in, float *maps, int S0, int lane) {
    constexpr int DYS[3] = {SET == 0 ? 0 : 1, 1, SET == 0 ? 0 : 2};
    constexpr int DXS[3] = {SET == 0 ? 1 : 0, SET == 0 ? 1 : -1, SET == 0 ? 2 : 0};
    constexpr int NR = SET == 0 ? 5 : 6;                       // input rows S0-1 .. S0+2+max dy
    float dist[3][2][4];
    int off[NR];
    off[0] = k2s_slot(S0 - 1, K2S_RS_GAB);
#pragma unroll
    for (int i = 1; i < NR; i++) off[i] = k2s_next(off[i - 1], K2S_RS_GAB);
#pragma unroll
    for (int i = 0; i < NR; i++) off[i] = off[i] * K2S_PITCH + 4 * lane;
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int j = 0; j < 4; j++) dist[m][k][j] = 0.0f;
    const int nch = K2S_ROLLED_TRIPS(P);
#pragma unroll 1
    for (int c = 0; c < nch; c++) {
        const float *pl = in + c * K2S_RS_GAB * K2S_PITCH;
        float I[NR][4], e1[NR], e2[NR], w1[NR];                // the row's quad; columns x+4, x+5 and x-1 where a map of the set needs them
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const K2SQuad q = k2s_ld4(pl + off[i]);
            I[i][0] = q.x; I[i][1] = q.y; I[i][2] = q.z; I[i][3] = q.w;
        }
#pragma unroll
        for (int i = 0; i < NR; i++) {
            e1[i] = 0.0f; e2[i] = 0.0f; w1[i] = 0.0f;
            if (SET == 0) { e1[i] = k2s_dn(I[i][0]); if (i < 4) e2[i] = k2s_dn(I[i][1]); }     // (0,2) shifts rows 0..3 only
            else if (i >= 1 && i < 5) w1[i] = k2s_up(I[i][3]);                                    // (1,-1) shifts rows 1..4
        }
        const float s = P.ch_scale[c];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            const int DY = DYS[m], DX = DXS[m];
            float T[4][4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const float *b = I[r], *v = I[r + DY];
                float sh[4];
                if (DX == 0) { sh[0] = v[0]; sh[1] = v[1]; sh[2] = v[2]; sh[3] = v[3]; }
                else if (DX == 1) { sh[0] = v[1]; sh[1] = v[2]; sh[2] = v[3]; sh[3] = e1[r + DY]; }
                else if (DX == 2) { sh[0] = v[2]; sh[1] = v[3]; sh[2] = e1[r + DY]; sh[3] = e2[r + DY]; }
                else { sh[0] = w1[r + DY]; sh[1] = v[0]; sh[2] = v[1]; sh[3] = v[2]; }
#pragma unroll
                for (int j = 0; j < 4; j++) T[r][j] = K2S_MUL(fabsf(K2S_SUB(b[j], sh[j])), s);
            }
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const float *t = T[k + 1];
                const float t6[6] = {k2s_up(t[3]), t[0], t[1], t[2], t[3], k2s_dn(t[0])};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float d = K2S_ADD(dist[m][k][j], t6[j + 1]);                   // (0, 0); the running sum starts at 0f like the Java's
                    d = K2S_ADD(d, t6[j]);                                         // (0, -1)
                    d = K2S_ADD(d, t6[j + 2]);                                     // (0, +1)
                    d = K2S_ADD(d, T[k][j]);                                       // (-1, 0)
                    d = K2S_ADD(d, T[k + 2][j]);                                   // (+1, 0)
                    dist[m][k][j] = d;
                }
            }
        }
    }
    const int so0 = k2s_slot(S0, K2S_RS_D0), so1 = k2s_next(so0, K2S_RS_D0);
#pragma unroll
    for (int m = 0; m < 3; m++) {
        float *mp = maps + (2 * m + SET) * K2S_RS_D0 * K2S_PITCH + 4 * lane;
        K2SQuad q0, q1;
        q0.x = dist[m][0][0]; q0.y = dist[m][0][1]; q0.z = dist[m][0][2]; q0.w = dist[m][0][3];
        q1.x = dist[m][1][0]; q1.y = dist[m][1][1]; q1.z = dist[m][1][2]; q1.w = dist[m][1][3];
        k2s_st4(mp + so0 * K2S_PITCH, q0);
        k2s_st4(mp + so1 * K2S_PITCH, q1);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// W0: pass 0 (13-point double cross, Frame.java:44-55 crossList order) of stream row S: weights from the six maps, channel sums,
// divide -> P0 ring.  Map order: 0 (0,1), 1 (1,0), 2 (1,1), 3 (1,-1), 4 (0,2), 5 (2,0); dist_{-d}(p) = map_d(p - d).
// ------------------------------------------------------------------------------------------------------------------------
template <bool SAT> K2S_FN void k2s_w0_body(const K2SArgs &A, float *sm, const K2SRow &R, int S, int lane, float is, const float (&m)[4]) {
    const K2Params &P = A.P;
    const float *in = sm + K2S_OFF_GAB;
    float *ring = sm + K2S_OFF_P0;
    const float *d0 = sm + K2S_OFF_D0;
    const int s0 = k2s_slot(S, K2S_RS_D0), s1 = k2s_prev(s0, K2S_RS_D0), s2 = k2s_prev(s1, K2S_RS_D0);
    const int g2 = k2s_slot(S, K2S_RS_GAB), g1 = k2s_prev(g2, K2S_RS_GAB), g0 = k2s_prev(g1, K2S_RS_GAB), g3 = k2s_next(g2, K2S_RS_GAB),
              g4 = k2s_next(g3, K2S_RS_GAB);
#define K2S_MAP(m, s) k2s_ld4(d0 + ((m) * K2S_RS_D0 + (s)) * K2S_PITCH + 4 * lane)
    const K2SQuad q01 = K2S_MAP(0, s0), q10 = K2S_MAP(1, s0), q11 = K2S_MAP(2, s0), q1m = K2S_MAP(3, s0), q02 = K2S_MAP(4, s0), q20 = K2S_MAP(5, s0);
    const K2SQuad p10 = K2S_MAP(1, s1), p11 = K2S_MAP(2, s1), p1m = K2S_MAP(3, s1), p20 = K2S_MAP(5, s2);
#undef K2S_MAP
    const float e01 = k2s_up(q01.w), e02a = k2s_up(q02.z), e02b = k2s_up(q02.w), e11 = k2s_up(p11.w), e1m = k2s_dn(p1m.x);
    const float v01[5] = {e01, q01.x, q01.y, q01.z, q01.w};              // map (0,1) at columns x-1 .. x+3
    const float v02[6] = {e02a, e02b, q02.x, q02.y, q02.z, q02.w};       // map (0,2) at columns x-2 .. x+3
    const float u11[5] = {e11, p11.x, p11.y, p11.z, p11.w};              // map (1,1), row y-1, columns x-1 .. x+3
    const float u1m[5] = {p1m.x, p1m.y, p1m.z, p1m.w, e1m};              // map (1,-1), row y-1, columns x .. x+4
    const bool pass = !(is <= (1.0f / 0.3f));                            // copied through (Frame.java:608-612); also NaN
    const float ss = P.sigma_scale[0];
    float w[4][12], sumw[4], rsum[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        w[j][0] = k2s_wgt<SAT>(v01[j], m[j], ss, is);                         // (0,-1) = -(0,1): map at (y, x-1)
        w[j][1] = k2s_wgt<SAT>(v01[j + 1], m[j], ss, is);                     // (0, 1)
        w[j][2] = k2s_wgt<SAT>(k2s_get(p10, j), m[j], ss, is);                // (-1,0) = -(1,0): map at (y-1, x)
        w[j][3] = k2s_wgt<SAT>(k2s_get(q10, j), m[j], ss, is);                // (1, 0)
        w[j][4] = k2s_wgt<SAT>(u1m[j + 1], m[j], ss, is);                     // (-1,1) = -(1,-1): map at (y-1, x+1)
        w[j][5] = k2s_wgt<SAT>(k2s_get(q11, j), m[j], ss, is);                // (1, 1)
        w[j][6] = k2s_wgt<SAT>(k2s_get(q1m, j), m[j], ss, is);                // (1,-1)
        w[j][7] = k2s_wgt<SAT>(u11[j], m[j], ss, is);                         // (-1,-1) = -(1,1): map at (y-1, x-1)
        w[j][8] = k2s_wgt<SAT>(v02[j], m[j], ss, is);                         // (0,-2) = -(0,2): map at (y, x-2)
        w[j][9] = k2s_wgt<SAT>(v02[j + 2], m[j], ss, is);                     // (0, 2)
        w[j][10] = k2s_wgt<SAT>(k2s_get(q20, j), m[j], ss, is);               // (2, 0)
        w[j][11] = k2s_wgt<SAT>(k2s_get(p20, j), m[j], ss, is);               // (-2,0) = -(2,0): map at (y-2, x)
        float s = 1.0f;                                                  // 0 + weight(centre) = 1
#pragma unroll
        for (int k = 0; k < 12; k++) s = K2S_ADD(s, w[j][k]);
        sumw[j] = s;
        rsum[j] = K2S_RCP(s);
    }
    K2SQuad out[3];
    float sv[3][4], ctr[3][4], o[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *pl = in + c * K2S_RS_GAB * K2S_PITCH + 4 * lane;
        const K2SQuad a = k2s_ld4(pl + g0 * K2S_PITCH), b = k2s_ld4(pl + g1 * K2S_PITCH);
        const K2SQuad r = k2s_ld4(pl + g2 * K2S_PITCH);
        const K2SQuad d = k2s_ld4(pl + g3 * K2S_PITCH), e = k2s_ld4(pl + g4 * K2S_PITCH);
        const float b6[6] = {k2s_up(b.w), b.x, b.y, b.z, b.w, k2s_dn(b.x)};
        const float d6[6] = {k2s_up(d.w), d.x, d.y, d.z, d.w, k2s_dn(d.x)};
        const float r8[8] = {k2s_up(r.z), k2s_up(r.w), r.x, r.y, r.z, r.w, k2s_dn(r.x), k2s_dn(r.y)};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = r8[j + 2];                                          // 0 + I * 1
            s = K2S_ADD(s, K2S_MUL(r8[j + 1], w[j][0]));
            s = K2S_ADD(s, K2S_MUL(r8[j + 3], w[j][1]));
            s = K2S_ADD(s, K2S_MUL(b6[j + 1], w[j][2]));
            s = K2S_ADD(s, K2S_MUL(d6[j + 1], w[j][3]));
            s = K2S_ADD(s, K2S_MUL(b6[j + 2], w[j][4]));
            s = K2S_ADD(s, K2S_MUL(d6[j + 2], w[j][5]));
            s = K2S_ADD(s, K2S_MUL(d6[j], w[j][6]));
            s = K2S_ADD(s, K2S_MUL(b6[j], w[j][7]));
            s = K2S_ADD(s, K2S_MUL(r8[j], w[j][8]));
            s = K2S_ADD(s, K2S_MUL(r8[j + 4], w[j][9]));
            s = K2S_ADD(s, K2S_MUL(k2s_get(e, j), w[j][10]));
            s = K2S_ADD(s, K2S_MUL(k2s_get(a, j), w[j][11]));
            sv[c][j] = s;
            ctr[c][j] = r8[j + 2];
        }
    }
    k2s_div12(sv, sumw, rsum, o);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        out[c].x = pass ? ctr[c][0] : o[c][0]; out[c].y = pass ? ctr[c][1] : o[c][1];
        out[c].z = pass ? ctr[c][2] : o[c][2]; out[c].w = pass ? ctr[c][3] : o[c][3];
    }
    k2s_emit(P, ring, K2S_RS_P0, K2S_MARGIN_P0, R, S, lane, out);
}
// the literal-clamp forms run only for rows with a non-positive HF multiplier: out of line, away from the hot code
K2S_COLD void k2s_w0_body_literal(const K2SArgs &A, float *sm, const K2SRow &R, int S, int lane, float is, const float (&m)[4]) {
    k2s_w0_body<false>(A, sm, R, S, lane, is, m);
}
K2S_FN int k2s_w0_row(const K2SArgs &A, float *sm, K2SCursor &cur, int S, int lane) {
    const K2SRow R = k2s_at(A, cur, S);
    const int kind = R.valid ? k2s_row_kind(A.P, R.y, K2S_MARGIN_P0) : 1;
    if (kind != 0) return kind;          // 2: a row below the frame, copied from behind after the phase's barrier (k2s_fixup)
    float m[4];
    const float is = k2s_sigma(A, R, lane, m);                           // a global load: issued before anything else of the row
    if (k2s_any(is < 0.0f)) k2s_w0_body_literal(A, sm, R, S, lane, is, m);
    else k2s_w0_body<true>(A, sm, R, S, lane, is, m);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// W1: pass 1 (5-point cross) of stream row S from the two maps; input ring = P0 (epf_iters == 3) or GAB.  LAST: epf_iters == 1,
// the row goes through the colour transform to HBM instead of the P1 ring.
// ------------------------------------------------------------------------------------------------------------------------
template <int RS_IN, bool LAST, bool SAT>
K2S_FN void k2s_w1_body(const K2SArgs &A, float *sm, const float *in, const K2SRow &R, int S, int lane, float is, const float (&m)[4]) {
    const K2Params &P = A.P;
    float *ring = sm + K2S_OFF_P1;
    const float *d1 = sm + K2S_OFF_D1;
    const int s0 = k2s_slot(S, K2S_RS_D1), s1 = k2s_prev(s0, K2S_RS_D1);
    const int i1 = k2s_slot(S, RS_IN), i0 = k2s_prev(i1, RS_IN), i2 = k2s_next(i1, RS_IN);
    const K2SQuad q01 = k2s_ld4(d1 + (0 * K2S_RS_D1 + s0) * K2S_PITCH + 4 * lane);
    const K2SQuad q10 = k2s_ld4(d1 + (1 * K2S_RS_D1 + s0) * K2S_PITCH + 4 * lane);
    const K2SQuad p10 = k2s_ld4(d1 + (1 * K2S_RS_D1 + s1) * K2S_PITCH + 4 * lane);
    const float v01[5] = {k2s_up(q01.w), q01.x, q01.y, q01.z, q01.w};
    const bool pass = !(is <= (1.0f / 0.3f));
    const float ss = P.sigma_scale[1];
    float w[4][4], sumw[4], rsum[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        w[j][0] = k2s_wgt<SAT>(v01[j], m[j], ss, is);
        w[j][1] = k2s_wgt<SAT>(v01[j + 1], m[j], ss, is);
        w[j][2] = k2s_wgt<SAT>(k2s_get(p10, j), m[j], ss, is);
        w[j][3] = k2s_wgt<SAT>(k2s_get(q10, j), m[j], ss, is);
        sumw[j] = K2S_ADD(K2S_ADD(K2S_ADD(K2S_ADD(1.0f, w[j][0]), w[j][1]), w[j][2]), w[j][3]);
        rsum[j] = K2S_RCP(sumw[j]);
    }
    K2SQuad out[3];
    float sv[3][4], ctr[3][4], o[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *pl = in + c * RS_IN * K2S_PITCH + 4 * lane;
        const K2SQuad b = k2s_ld4(pl + i0 * K2S_PITCH), r = k2s_ld4(pl + i1 * K2S_PITCH);
        const K2SQuad d = k2s_ld4(pl + i2 * K2S_PITCH);
        const float r6[6] = {k2s_up(r.w), r.x, r.y, r.z, r.w, k2s_dn(r.x)};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = r6[j + 1];
            s = K2S_ADD(s, K2S_MUL(r6[j], w[j][0]));
            s = K2S_ADD(s, K2S_MUL(r6[j + 2], w[j][1]));
            s = K2S_ADD(s, K2S_MUL(k2s_get(b, j), w[j][2]));
            s = K2S_ADD(s, K2S_MUL(k2s_get(d, j), w[j][3]));
            sv[c][j] = s;
            ctr[c][j] = r6[j + 1];
        }
    }
    k2s_div12(sv, sumw, rsum, o);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        out[c].x = pass ? ctr[c][0] : o[c][0]; out[c].y = pass ? ctr[c][1] : o[c][1];
        out[c].z = pass ? ctr[c][2] : o[c][2]; out[c].w = pass ? ctr[c][3] : o[c][3];
    }
    if (LAST) k2s_final(A, R, lane, out);
    else k2s_emit(P, ring, K2S_RS_P1, K2S_MARGIN_P1, R, S, lane, out);
}
template <int RS_IN, bool LAST>
K2S_COLD void k2s_w1_body_literal(const K2SArgs &A, float *sm, const float *in, const K2SRow &R, int S, int lane, float is, const float (&m)[4]) {
    k2s_w1_body<RS_IN, LAST, false>(A, sm, in, R, S, lane, is, m);
}
template <int RS_IN, bool LAST> K2S_FN int k2s_w1_row(const K2SArgs &A, float *sm, const float *in, K2SCursor &cur, int S, int lane) {
    const K2SRow R = k2s_at(A, cur, S);
    const int kind = R.valid ? k2s_row_kind(A.P, R.y, LAST ? 0 : K2S_MARGIN_P1) : 1;
    if (kind != 0) return kind;
    float m[4];
    const float is = k2s_sigma(A, R, lane, m);
    if (k2s_any(is < 0.0f)) k2s_w1_body_literal<RS_IN, LAST>(A, sm, in, R, S, lane, is, m);
    else k2s_w1_body<RS_IN, LAST, true>(A, sm, in, R, S, lane, is, m);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// P2: pass 2 (5-point cross, point differences: epfDistance2, Frame.java:657-669) of stream row S from the P1 ring, then the
// colour transform and the store.  Everything stays in registers.
// ------------------------------------------------------------------------------------------------------------------------
template <bool SAT> K2S_FN void k2s_p2_body(const K2SArgs &A, float *sm, const K2SRow &R, int S, int lane, float is, const float (&m)[4]) {
    const K2Params &P = A.P;
    const float *in = sm + K2S_OFF_P1;
    float r6[3][6], up[3][4], dn[3][4];
    const int i1 = k2s_slot(S, K2S_RS_P1), i0 = k2s_prev(i1, K2S_RS_P1), i2 = k2s_next(i1, K2S_RS_P1);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *pl = in + c * K2S_RS_P1 * K2S_PITCH + 4 * lane;
        const K2SQuad b = k2s_ld4(pl + i0 * K2S_PITCH), r = k2s_ld4(pl + i1 * K2S_PITCH);
        const K2SQuad d = k2s_ld4(pl + i2 * K2S_PITCH);
        r6[c][0] = k2s_up(r.w); r6[c][1] = r.x; r6[c][2] = r.y; r6[c][3] = r.z; r6[c][4] = r.w; r6[c][5] = k2s_dn(r.x);
        up[c][0] = b.x; up[c][1] = b.y; up[c][2] = b.z; up[c][3] = b.w;
        dn[c][0] = d.x; dn[c][1] = d.y; dn[c][2] = d.z; dn[c][3] = d.w;
    }
    float h[5], vu[4], vd[4];                                 // (0,1) distances at columns x-1 .. x+3; (1,0) at rows y-1 and y
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float s = P.ch_scale[c];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float t = K2S_MUL(fabsf(K2S_SUB(r6[c][k], r6[c][k + 1])), s);
            h[k] = c == 0 ? t : K2S_ADD(h[k], t);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float tu = K2S_MUL(fabsf(K2S_SUB(up[c][j], r6[c][j + 1])), s), td = K2S_MUL(fabsf(K2S_SUB(r6[c][j + 1], dn[c][j])), s);
            vu[j] = c == 0 ? tu : K2S_ADD(vu[j], tu);
            vd[j] = c == 0 ? td : K2S_ADD(vd[j], td);
        }
    }
    const bool pass = !(is <= (1.0f / 0.3f));
    const float ss = P.sigma_scale[2];
    float o[3][4], sv[3][4], sumw[4], rsum[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float w0 = k2s_wgt<SAT>(h[j], m[j], ss, is), w1 = k2s_wgt<SAT>(h[j + 1], m[j], ss, is);
        const float w2 = k2s_wgt<SAT>(vu[j], m[j], ss, is), w3 = k2s_wgt<SAT>(vd[j], m[j], ss, is);
        sumw[j] = K2S_ADD(K2S_ADD(K2S_ADD(K2S_ADD(1.0f, w0), w1), w2), w3);
        rsum[j] = K2S_RCP(sumw[j]);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float s = r6[c][j + 1];
            s = K2S_ADD(s, K2S_MUL(r6[c][j], w0));
            s = K2S_ADD(s, K2S_MUL(r6[c][j + 2], w1));
            s = K2S_ADD(s, K2S_MUL(up[c][j], w2));
            s = K2S_ADD(s, K2S_MUL(dn[c][j], w3));
            sv[c][j] = s;
        }
    }
    k2s_div12(sv, sumw, rsum, o);
    if (pass) {
#pragma unroll
        for (int c = 0; c < 3; c++) { o[c][0] = r6[c][1]; o[c][1] = r6[c][2]; o[c][2] = r6[c][3]; o[c][3] = r6[c][4]; }
    }
    K2SQuad out[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { out[c].x = o[c][0]; out[c].y = o[c][1]; out[c].z = o[c][2]; out[c].w = o[c][3]; }
    k2s_final(A, R, lane, out);
}
K2S_COLD void k2s_p2_body_literal(const K2SArgs &A, float *sm, const K2SRow &R, int S, int lane, float is, const float (&m)[4]) {
    k2s_p2_body<false>(A, sm, R, S, lane, is, m);
}
K2S_FN void k2s_p2_row(const K2SArgs &A, float *sm, K2SCursor &cur, int S, int lane) {
    const K2SRow R = k2s_at(A, cur, S);
    if (!R.valid || k2s_row_kind(A.P, R.y, 0) != 0) return;
    float m[4];
    const float is = k2s_sigma(A, R, lane, m);
    if (k2s_any(is < 0.0f)) k2s_p2_body_literal(A, sm, R, S, lane, is, m);
    else k2s_p2_body<true>(A, sm, R, S, lane, is, m);
}

// ------------------------------------------------------------------------------------------------------------------------
// the tick loop: every stage advances by one band, one stage after the other
// ------------------------------------------------------------------------------------------------------------------------
K2S_FN int k2s_total_rows(const K2SArgs &A) {
    const int cta = k2s_cta(), g = k2s_grid();
    const int mine = cta < A.n_items ? (A.n_items - cta + g - 1) / g : 0;
    return mine * A.ir;
}

// rows [BAND t, BAND t + BAND) of the stream -> RAW ring: BAND / BOX boxes per plane (an item is a multiple of 8 rows, a box never
// straddles two).  One thread issues (twelve lanes issuing at once were measured no faster); the slots were last read by G a tick ago,
// which a CTA barrier separates from this call.
K2S_FN void k2s_load_band(const K2SArgs &A, float *sm, uint64_t *bars, K2S_TMAP_PARAM t0, K2S_TMAP_PARAM t1, K2S_TMAP_PARAM t2, int t, int total,
                              K2SCursor &cur) {
    int boxes = (total - K2S_BAND * t + K2S_BOX - 1) / K2S_BOX;
    if (boxes <= 0) return;
    if (boxes > K2S_BAND / K2S_BOX) boxes = K2S_BAND / K2S_BOX;
#ifndef K2S_HOST_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of tick t - 1 before async-proxy writes
    uint64_t *bar = bars + (t & 1);
    k2s_mbar_expect_tx(bar, boxes * 3 * K2S_BOX * K2S_PITCH * 4);
#endif
#pragma unroll 1
    for (int h = 0; h < boxes; h++) {
        const int S = K2S_BAND * t + K2S_BOX * h;
        const K2SRow R = k2s_at(A, cur, S);       // the issuing thread's warp is on the phase's critical path: no divisions here
        const int ty = R.z * A.P.rows + R.y + A.tma_row0, slot = k2s_slot(S, K2S_RS_RAW);
        float *dst = sm + K2S_OFF_RAW + slot * K2S_PITCH;
#ifdef K2S_HOST_EMU
        k2s_tma_box(dst, t0, R.x0 - K2S_HALO, ty);
        k2s_tma_box(dst + K2S_RS_RAW * K2S_PITCH, t1, R.x0 - K2S_HALO, ty);
        k2s_tma_box(dst + 2 * K2S_RS_RAW * K2S_PITCH, t2, R.x0 - K2S_HALO, ty);
#else
        k2s_tma_box_async(dst, t0, R.x0 - K2S_HALO, ty, bar);
        k2s_tma_box_async(dst + K2S_RS_RAW * K2S_PITCH, t1, R.x0 - K2S_HALO, ty, bar);
        k2s_tma_box_async(dst + 2 * K2S_RS_RAW * K2S_PITCH, t2, R.x0 - K2S_HALO, ty, bar);
#endif
    }
}

// End of a phase that writes a plane ring: CTA barrier, then the rows below the frame (kind 2) are copied from their mirror rows --
// which another warp may have produced in this very phase -- and a second barrier publishes them.  Only the ticks that hold a frame's
// last rows pay for the second one.
K2S_FN void k2s_fixup(const K2SArgs &A, float *ring, int rs, K2SCursor &cur, int kind, int S, int lane) {
    if (k2s_sync_or(kind == 2)) {
        if (kind == 2) k2s_copy_behind(A.P, ring, rs, k2s_at(A, cur, S), S, lane);
        k2s_sync();
    }
}

template <int DY, int DX> K2S_FN void k2s_d_task(const K2SArgs &A, const float *in, int rs_in, float *outmap, int rs_out, int S0, int total, int lane) {
    if (S0 + 1 >= 0 && S0 < total) k2s_d_pair<DY, DX>(A.P, in, rs_in, outmap, rs_out, S0, lane);
}

// -DK2S_PHASE_TIMING (diagnostic builds only): cycles warp 1 spends working in each stage and waiting at the barrier that ends it,
// summed over ticks and CTAs into k2s_phase_clk[2 * stage + {0 work, 1 wait}] (tools/phase_timing.py reads it)
#if defined(K2S_PHASE_TIMING) && !defined(K2S_HOST_EMU)
__device__ unsigned long long k2s_phase_clk[16];
#define K2S_CLK_BEGIN long long clk_a = clock64(), clk_b;
#define K2S_CLK_WORK(ph) clk_b = clock64(); if (tid == 32) atomicAdd(&k2s_phase_clk[2 * (ph)], (unsigned long long)(clk_b - clk_a)); clk_a = clk_b;
#define K2S_CLK_WAIT(ph) clk_b = clock64(); if (tid == 32) atomicAdd(&k2s_phase_clk[2 * (ph) + 1], (unsigned long long)(clk_b - clk_a)); clk_a = clk_b;
#else
#define K2S_CLK_BEGIN
#define K2S_CLK_WORK(ph)
#define K2S_CLK_WAIT(ph)
#endif

template <int GAB, int ITERS>
K2S_FN void k2s_body(const K2SArgs &A, float *sm, uint64_t *bars, K2S_TMAP_PARAM t0, K2S_TMAP_PARAM t1, K2S_TMAP_PARAM t2) {
    using Cfg = K2SCfg<ITERS>;
    const int tid = k2s_tid(), warp = tid >> 5, lane = tid & 31;
    const int total = k2s_total_rows(A);
    const int n_ticks = (total - Cfg::LAST + K2S_BAND - 1) / K2S_BAND;
#ifndef K2S_HOST_EMU
    if (tid == 0) {
        k2s_mbar_init(bars, 1);
        k2s_mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#endif
    float *gab = sm + K2S_OFF_GAB, *p0 = sm + K2S_OFF_P0, *p1 = sm + K2S_OFF_P1, *d0 = sm + K2S_OFF_D0, *d1 = sm + K2S_OFF_D1;
    const float *in1 = ITERS == 3 ? p0 : gab;                 // what pass 1 reads
    constexpr int RS1 = ITERS == 3 ? K2S_RS_P0 : K2S_RS_GAB;
    K2SCursor cur;
    k2s_cursor_init(cur);
    if (tid == 0) k2s_load_band(A, sm, bars, t0, t1, t2, 0, total, cur);
#ifdef K2S_HOST_EMU
    k2s_sync();
#endif
    // G of band `tb`: the band's rows have landed (mbarrier), one row per warp, then the barrier and the copies of rows below the frame
    // (k2s_fixup), then the loads of the next band are issued -- they land while the other stages run
#define K2S_G_PHASE(tb)                                                                                          \
    {                                                                                                            \
        K2S_WAIT_BAND(tb)                                                                                        \
        K2S_CLK_WAIT(6)                                                                                          \
        const int Sg = K2S_BAND * (tb) + Cfg::G + warp;                                                          \
        const int kindg = (Sg >= 0 && Sg < total) ? k2s_g_row<GAB>(A, sm, cur, Sg, lane) : 1;                    \
        K2S_CLK_WORK(0)                                                                                          \
        k2s_fixup(A, gab, K2S_RS_GAB, cur, kindg, Sg, lane);                                                     \
        K2S_CLK_WAIT(0)                                                                                          \
        if (tid == 0) k2s_load_band(A, sm, bars, t0, t1, t2, (tb) + 1, total, cur);                              \
    }
#ifdef K2S_HOST_EMU
#define K2S_WAIT_BAND(tb)
#else
#define K2S_WAIT_BAND(tb) if (K2S_BAND * (tb) < total) k2s_mbar_wait(bars + ((tb) & 1), ((tb) >> 1) & 1);
#endif
    // K2S_MERGE_G (epf_iters == 3): G runs one band ahead, in the same phase as D1 (they touch different rings; G's barrier is D1's too),
    // so a tick has four barriers instead of five.  Measured SLOWER (8K frame: 1.406 ms against 1.370 ms), so it is off.
#ifndef K2S_MERGE_G
#define K2S_MERGE_G 0
#endif
    constexpr bool MERGE = ITERS == 3 && K2S_MERGE_G;
    K2S_CLK_BEGIN
    if (MERGE) K2S_G_PHASE(0)
#pragma unroll 1
    for (int t = 0; t < n_ticks; t++) {
        const int S16 = K2S_BAND * t;
        if (!MERGE) K2S_G_PHASE(t)
        if (ITERS == 3) {
            // ---- D0: 6 maps x BAND / 2 row pairs; a warp forms three maps of its row pair at once (they read the same input rows) ----
            {
                const int S0 = S16 + Cfg::D0 + 2 * (warp % (K2S_BAND / 2));
                if (S0 + 1 >= 0 && S0 < total) {
                    if (warp < K2S_BAND / 2) k2s_d_triple<0>(A.P, gab, d0, S0, lane);
                    else k2s_d_triple<1>(A.P, gab, d0, S0, lane);
                }
            }
            K2S_CLK_WORK(1)
            k2s_sync();
            K2S_CLK_WAIT(1)
            // ---- W0 ----
            {
                const int S = S16 + Cfg::W0 + warp;
                const int kind = (S >= 0 && S < total) ? k2s_w0_row(A, sm, cur, S, lane) : 1;
                K2S_CLK_WORK(2)
                k2s_fixup(A, p0, K2S_RS_P0, cur, kind, S, lane);
                K2S_CLK_WAIT(2)
            }
        }
        // ---- D1: 2 maps x BAND / 2 row pairs ----
        {
            const int S0 = S16 + Cfg::D1 + 2 * (warp % (K2S_BAND / 2));
            if (warp < K2S_BAND / 2) k2s_d_task<0, 1>(A, in1, RS1, d1, K2S_RS_D1, S0, total, lane);
            else k2s_d_task<1, 0>(A, in1, RS1, d1 + K2S_RS_D1 * K2S_PITCH, K2S_RS_D1, S0, total, lane);
        }
        K2S_CLK_WORK(3)
        if (MERGE) K2S_G_PHASE(t + 1)      // GAB slots of band t + 1 were last read by W0 of this tick, two barriers ago
        else k2s_sync();
        K2S_CLK_WAIT(3)
        // ---- W1 ----
        {
            const int S = S16 + Cfg::W1 + warp;
            const int kind = (S >= 0 && S < total) ? k2s_w1_row<RS1, ITERS == 1>(A, sm, in1, cur, S, lane) : 1;
            K2S_CLK_WORK(4)
            if (ITERS > 1) k2s_fixup(A, p1, K2S_RS_P1, cur, kind, S, lane);
            K2S_CLK_WAIT(4)
        }
        // ---- P2 ----
        if (ITERS > 1) {
            const int S = S16 + Cfg::P2 + warp;
            if (S >= 0 && S < total) k2s_p2_row(A, sm, cur, S, lane);
            K2S_CLK_WORK(5)
        }
        // epf_iters > 1: no barrier here -- the next phase (D0, or G when epf_iters == 2) writes rings whose last readers ran before
        // the barriers above.  epf_iters == 1: W1 has just read the GAB ring, and G is about to write it.
        if (ITERS == 1) k2s_sync();
    }
#undef K2S_G_PHASE
#undef K2S_WAIT_BAND
}

// ------------------------------------------------------------------------------------------------------------------------
// host side shared by the library and the emulator: how a frame is cut into items
// ------------------------------------------------------------------------------------------------------------------------
// Picks the rows per item: items are dealt round-robin to min(items, n_cta) CTAs, each streams ir = ch + 16 rows per item, so the
// kernel's length is about ceil(items / ctas) * (ch + 16) rows; short items balance better, tall ones waste fewer halo rows.
static inline void k2s_plan(int W, int rows, int n_frames, int n_cta, K2SArgs &A) {
    A.n_cols = (W + K2S_TW - 1) / K2S_TW;
    A.n_frames = n_frames;
    long long best = -1;
    int best_chunks = 1;
    for (int n = 1; n <= (rows + 7) / 8 && n <= 4096; n++) {
        const int ch = (((rows + n - 1) / n) + 7) & ~7;
        if (ch < 32 && n > 1) break;
        const int chunks = (rows + ch - 1) / ch;
        const long long items = (long long)n_frames * A.n_cols * chunks;
        const long long ctas = items < n_cta ? items : n_cta;
        const long long cost = ((items + ctas - 1) / ctas) * (ch + 2 * K2S_HALO);
        if (best < 0 || cost < best) { best = cost; best_chunks = chunks; A.ch = ch; }
    }
    A.n_chunks = best_chunks;
    A.ir = A.ch + 2 * K2S_HALO;
    A.n_items = n_frames * A.n_cols * A.n_chunks;
}

#ifndef K2S_HOST_EMU
template <int GAB, int ITERS>
__global__ void __launch_bounds__(32 * K2S_NWARPS, K2S_CTAS_PER_SM)
k2_stream(const __grid_constant__ K2SArgs A, const __grid_constant__ CUtensorMap t0, const __grid_constant__ CUtensorMap t1,
          const __grid_constant__ CUtensorMap t2) {
    extern __shared__ __align__(128) float k2s_smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(k2s_smem + K2S_FLOATS);
    k2s_body<GAB, ITERS>(A, k2s_smem, bars, &t0, &t1, &t2);
}

typedef CUresult (*k2s_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline k2s_encode_fn k2s_encoder() {
    static k2s_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (k2s_encode_fn)p;
    }
    return fn;
}

template <int GAB, int ITERS> static cudaError_t k2s_attr() {
    return cudaFuncSetAttribute(k2_stream<GAB, ITERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2S_BYTES);
}
static inline cudaError_t k2_stream_init_all() {
    cudaError_t e;
    if ((e = k2s_attr<1, 1>()) != cudaSuccess) return e;
    if ((e = k2s_attr<1, 2>()) != cudaSuccess) return e;
    if ((e = k2s_attr<1, 3>()) != cudaSuccess) return e;
    if ((e = k2s_attr<0, 1>()) != cudaSuccess) return e;
    if ((e = k2s_attr<0, 2>()) != cudaSuccess) return e;
    if ((e = k2s_attr<0, 3>()) != cudaSuccess) return e;
    return cudaSuccess;
}
// the TMA path needs 16-byte aligned planes and pitches that are multiples of 4 floats; anything else stays on k2_exact
static inline bool k2_stream_supported(const K2Params &K, int n_frames) {
    if (K.iters < 1 || K.rows < 8 || K.W < 8 || (K.rows & 7) || (K.W & 7) || (K.in_pitch & 3) || (K.out_pitch & 3) || !k2s_encoder()) return false;
    for (int c = 0; c < 3; c++)
        if (((uintptr_t)K.in[c] | (uintptr_t)K.out[c]) & 15) return false;
    if ((long long)K.rows * n_frames + 16 > 0x7fffffffll) return false;
    if ((long long)((K.rows >> 3) + 2) * n_frames * K.wb > 0x7fffffffll) return false;       // k2s_sigma indexes the 1/sigma map with an int
    // k2s_wgt clamps 1 - x with one saturating subtract, which equals the reference's max(1 - x, 0) when x >= 0: every frame-level factor
    // of x must be non-negative (they are in every real stream; a frame with a negative scale goes to the tile kernel); the per-block
    // factor 1/sigma is checked row by row in the kernel
    if (!(K.border_mul >= 0.0f) || !(K.gscale > 0.0f)) return false;
    for (int i = 0; i < 3; i++)
        if (!(K.ch_scale[i] >= 0.0f) || !(K.sigma_scale[i] >= 0.0f)) return false;
    for (int i = 0; i < 8; i++)
        if (!(K.sharp_lut[i] >= 0.0f)) return false;
    return true;
}

// returns 0, or -1 when a tensor map could not be encoded (the caller falls back to k2_exact)
static inline int k2_stream_launch(const K2Params &K, const float *inv_sigma, cudaStream_t st, int n_frames, int sms) {
    K2SArgs A;
    A.P = K;
    A.inv_sigma = inv_sigma;
    A.zpx = (long long)K.rows * K.in_pitch;
    A.zblk = (K.rows >> 3) * K.wb;
    const int n_cta = sms * K2S_CTAS_PER_SM;
    k2s_plan(K.W, K.rows, n_frames, n_cta, A);
    A.tma_row0 = K.has_top ? JXLB200_HALO_ROWS : 0;
    const long long map_rows = (long long)K.rows * n_frames + (K.has_top ? JXLB200_HALO_ROWS : 0) + (K.has_bottom ? JXLB200_HALO_ROWS : 0);
    CUtensorMap tm[3];
    for (int c = 0; c < 3; c++) {
        const cuuint64_t dims[2] = {(cuuint64_t)K.W, (cuuint64_t)map_rows};
        const cuuint64_t strides[1] = {(cuuint64_t)K.in_pitch * 4};
        const cuuint32_t box[2] = {K2S_PITCH, K2S_BOX}, es[2] = {1, 1};
        void *base = (void *)(K.in[c] - (K.has_top ? (long long)JXLB200_HALO_ROWS * K.in_pitch : 0));
        if (k2s_encoder()(&tm[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
    }
    const int grid = A.n_items < n_cta ? A.n_items : n_cta;
#define K2S_GO(G, I) k2_stream<G, I><<<grid, 32 * K2S_NWARPS, K2S_BYTES, st>>>(A, tm[0], tm[1], tm[2])
    switch ((K.gab ? 4 : 0) + K.iters) {
    case 5: K2S_GO(1, 1); break;
    case 6: K2S_GO(1, 2); break;
    case 7: K2S_GO(1, 3); break;
    case 1: K2S_GO(0, 1); break;
    case 2: K2S_GO(0, 2); break;
    default: K2S_GO(0, 3); break;
    }
#undef K2S_GO
    return 0;
}
#endif
