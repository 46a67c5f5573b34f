// k2_stream.cuh -- K2 as a persistent, warp-specialised STREAM: Gaborish -> EPF pass 0 / 1 / 2 -> colour transform in one kernel,
// bit-identical to the reference (every float operation in the reference's order, uncontracted).
//
// Replaces Frame.performGabConvolution (J/frame/Frame.java:505-542), Frame.performEdgePreservingFilter (:544-679, epfDistance1
// :638-655, epfDistance2 :657-669, epfWeight :671-679) and JXLCodestreamDecoder.performColorTransforms
// (J/JXLCodestreamDecoder.java:256-283); J/ = java/com/traneptora/jxlatte/ in the reference tree.
//
// Shape of the computation (why it is not a tile kernel).  The filters are issue-bound, not HBM-bound, so the design
// minimises INSTRUCTIONS per pixel, then shared-memory wavefronts:
//   * a CTA (one per SM, persistent) walks column strips of the frame, 112 useful pixels wide (+8 halo columns each side = 128 =
//     32 lanes x 4 pixels), top to bottom, in "items" of CH rows + 8 halo rows each side; the rows of all its items form one
//     continuous stream of rows S = 0, 1, 2, ...  Nothing is recomputed vertically inside an item (a tile kernel recomputes its
//     7-row halo per tile: +29 % at 64x48).
//   * each stage of the chain is a ROLE held by dedicated warps; all roles run concurrently, each on its own 8-row band of the
//     stream, a fixed number of rows behind its producer, and hand rows to each other through ring buffers in shared memory.
//     One bar.sync per 8-row tick is the only synchronisation (plus the mbarrier the TMA loads complete on).
//         TMA (cp.async.bulk.tensor, 4-row boxes)  -> RAW ring
//         G   Gaborish (or copy)                   -> GAB ring          1 warp, rolling 3-row window in registers
//         D0  pass-0 distances, one warp per canonical offset d in {(0,1),(1,0),(1,1),(1,-1),(0,2),(2,0)} -> six distance maps
//         W0  pass-0 weights, sums, divide         -> P0 ring           4 warps
//         D1  pass-1 distances, d in {(0,1),(1,0)} -> two distance maps 2 warps
//         W1  pass-1 weights, sums, divide         -> P1 ring           2 warps (the last stage when epf_iters == 1)
//         P2  pass 2 + colour transform + store    -> HBM               3 warps
//   * a lane owns a 4-pixel quad of a row; its neighbours' values come from the neighbour lanes by shuffle, so every row is read
//     with one conflict-free 128-bit shared load per channel and the 4-way bank conflicts of strided scalar loads never occur.
//   * what is shared without changing a rounding (as in k2_exact.cuh): the 15 terms of epfDistance1 are
//     T_{c,d}(q) = fl(fl|I_c(q) - I_c(q+d)| * s_c), formed ONCE per position by the D warps and kept in registers as a rolling
//     3-row window; dist_{-d}(p) == dist_d(p-d) operation for operation, so six (two) maps serve the twelve (four) offsets; the
//     centre tap has weight exactly 1.  The ordered sums are literal.
//   * frame edges: MathHelper.mirrorCoordinate (J/util/MathHelper.java:323-329) applies to a stage's OUTPUT.  Rows above the
//     frame are written behind the producer when their mirror source row is produced (they lie in the item's own upper halo);
//     rows below are copied from behind when the producer reaches them; columns are mirrored across lanes before the store.
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
#define K2S_DIV(a, b) ((a) / (b))
#define K2S_LDG(p) (*(p))
#else
#include <cuda.h>
#define K2S_FN __device__ __forceinline__
#define K2S_ADD(a, b) __fadd_rn((a), (b))
#define K2S_SUB(a, b) __fsub_rn((a), (b))
#define K2S_MUL(a, b) __fmul_rn((a), (b))
#define K2S_DIV(a, b) __fdiv_rn((a), (b))
#define K2S_LDG(p) __ldg(p)
#endif

#define K2S_TW 112            /* useful columns of a column strip */
#define K2S_HALO 8            /* halo columns each side and halo rows each side of an item (7 used: 1 + 3 + 2 + 1) */
#define K2S_PITCH 128         /* floats per ring row: 32 lanes x 4 */
#define K2S_BAND 8            /* rows per tick */
#define K2S_RS_RAW 20
#define K2S_RS_GAB 34
#define K2S_RS_P0 30
#define K2S_RS_P1 18
#define K2S_RS_D0 18
#define K2S_RS_D1 17
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

// Stage s processes stream rows [8t + base_s, 8t + base_s + 8) in tick t.  A consumer trails its producer by one band (what
// it reads was complete before the tick began) plus the rows below its own that it reads.  The D roles trail further: a D warp
// loads row r at step r - 1 - DY (rolling window), and a row above the frame exists only once its mirror source row k - 1 has been
// produced (k <= margin), 2k + DY rows later in stream order: 8 more rows behind G, 5 more behind W0.
// The W bases are also chosen modulo 8: a row below the frame is a copy of its mirror row (k2s_copy_behind), and the warp that copies
// must not run ahead of the warp that produces the source within a tick.  Items start on multiples of 8, so a band covers frame rows
// y == base .. base + 7 (mod 8) and the frame's last row is == 7: with W0 == 7 (mod 8) and two rows per warp the pair (rows-1, rows) is
// one warp's and (rows-2, rows+1) straddles two ticks; with W1 == 2 (mod 8) and four rows per warp, or 1 (mod 8) and two rows per
// warp, rows-1 and rows are one warp's again.
template <int ITERS> struct K2SCfg;
template <> struct K2SCfg<3> {
    static constexpr int G = -9, D0 = -25, W0 = -33, D1 = -46, W1 = -54, P2 = -63, LAST = -63;
    static constexpr int NWARPS = 18;
};
template <> struct K2SCfg<2> {
    static constexpr int G = -9, D0 = 0, W0 = 0, D1 = -23, W1 = -31, P2 = -40, LAST = -40;
    static constexpr int NWARPS = 11;
};
template <> struct K2SCfg<1> {
    static constexpr int G = -9, D0 = 0, W0 = 0, D1 = -23, W1 = -31, P2 = 0, LAST = -31;
    static constexpr int NWARPS = 11;
};

// ------------------------------------------------------------------------------------------------------------------------
// platform layer: lane id, shuffles, CTA barrier, TMA.  Host versions live in tests/host/k2_stream_host.cpp.
// ------------------------------------------------------------------------------------------------------------------------
#ifdef K2S_HOST_EMU
int k2s_emu_tid();
int k2s_emu_cta();
int k2s_emu_grid();
float k2s_emu_shfl(float v, int src_lane);
void k2s_emu_sync();
struct K2STmap { const float *base; int w, rows; long long pitch; };
K2S_FN int k2s_tid() { return k2s_emu_tid(); }
K2S_FN int k2s_cta() { return k2s_emu_cta(); }
K2S_FN int k2s_grid() { return k2s_emu_grid(); }
K2S_FN float k2s_up(float v) { const int l = k2s_emu_tid() & 31; return k2s_emu_shfl(v, l > 0 ? l - 1 : l); }
K2S_FN float k2s_dn(float v) { const int l = k2s_emu_tid() & 31; return k2s_emu_shfl(v, l < 31 ? l + 1 : l); }
K2S_FN float k2s_from(float v, int src) { return k2s_emu_shfl(v, src); }
K2S_FN void k2s_sync() { k2s_emu_sync(); }
struct K2SQuad { float x, y, z, w; };
K2S_FN K2SQuad k2s_ld4(const float *p) { return K2SQuad{p[0], p[1], p[2], p[3]}; }
K2S_FN void k2s_st4(float *p, K2SQuad q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
K2S_FN void k2s_stg4(float *p, K2SQuad q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
// a 128 x 4 box at (x, y): zero fill outside the tensor, exactly what the TMA unit writes
K2S_FN void k2s_tma_box(float *dst, const K2STmap &m, int x, int y) {
    for (int r = 0; r < 4; r++)
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
// every warp of the CTA arrives at barrier 0 once per tick, each from its own role's loop
K2S_FN void k2s_sync() { asm volatile("bar.sync 0;" ::: "memory"); }
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
K2S_FN K2SRow k2s_locate(const K2SArgs &A, int S) {
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
// the divisions above are paid once per item, not once per row: a role keeps the item its last row was in
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
K2S_FN float k2s_wgt(float dist, float m, float ss, float is) {
    return fmaxf(K2S_SUB(1.0f, K2S_MUL(K2S_MUL(K2S_MUL(dist, m), ss), is)), 0.0f);
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
    return K2S_LDG(A.inv_sigma + (long long)R.z * A.zblk + (long long)(R.y >> 3) * P.wb + (col >> 3));
}

// ------------------------------------------------------------------------------------------------------------------------
// G: Gaborish (Frame.java:505-542) of stream row S into the GAB ring; one warp, rolling window: rows y-1 and y stay in registers
// with their west / east neighbours, row y+1 is the one load of the step.  Edge = clamp on the padded plane (:526-534).
// ------------------------------------------------------------------------------------------------------------------------
struct K2SGabState { float n[3][6], r[3][6]; };
template <int GAB> K2S_FN void k2s_g_row(const K2SArgs &A, float *sm, K2SGabState &st, K2SCursor &cur, int S, int lane) {
    const K2Params &P = A.P;
    const K2SRow R = k2s_at(A, cur, S);
    const float *raw = sm + K2S_OFF_RAW;
    float *gab = sm + K2S_OFF_GAB;
    K2SQuad out[3];
    if (!GAB) {
#pragma unroll
        for (int c = 0; c < 3; c++) out[c] = k2s_ld4(raw + (c * K2S_RS_RAW + k2s_slot(S, K2S_RS_RAW)) * K2S_PITCH + 4 * lane);
    } else {
        const int col = R.x0 - K2S_HALO + 4 * lane;
        const bool first_col = col == 0, last_col = col + 4 == P.W;
        const bool clamp_n = R.y == 0 && !P.has_top, clamp_s = R.y == P.rows - 1 && !P.has_bottom;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const K2SQuad q = k2s_ld4(raw + (c * K2S_RS_RAW + k2s_slot(S + 1, K2S_RS_RAW)) * K2S_PITCH + 4 * lane);
            float s6[6] = {k2s_up(q.w), q.x, q.y, q.z, q.w, k2s_dn(q.x)};
            if (first_col) s6[0] = s6[1];
            if (last_col) s6[5] = s6[4];
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float *Rr = st.r[c];
                const float nw = clamp_n ? Rr[j] : st.n[c][j], nx = clamp_n ? Rr[j + 1] : st.n[c][j + 1], ne = clamp_n ? Rr[j + 2] : st.n[c][j + 2];
                const float sw = clamp_s ? Rr[j] : s6[j], sx = clamp_s ? Rr[j + 1] : s6[j + 1], se = clamp_s ? Rr[j + 2] : s6[j + 2];
                // Frame.java:535-537: operand order kept, uncontracted
                const float adj = K2S_ADD(K2S_ADD(K2S_ADD(Rr[j], Rr[j + 2]), nx), sx);
                const float diag = K2S_ADD(K2S_ADD(K2S_ADD(nw, ne), sw), se);
                o[j] = K2S_ADD(K2S_ADD(K2S_MUL(P.gab_base[c], Rr[j + 1]), K2S_MUL(P.gab_adj[c], adj)), K2S_MUL(P.gab_diag[c], diag));
            }
            out[c].x = o[0]; out[c].y = o[1]; out[c].z = o[2]; out[c].w = o[3];
#pragma unroll
            for (int i = 0; i < 6; i++) { st.n[c][i] = st.r[c][i]; st.r[c][i] = s6[i]; }
        }
    }
    if (!R.valid) return;
    const int kind = k2s_row_kind(P, R.y, K2S_MARGIN_GAB);
    if (kind == 0) k2s_emit(P, gab, K2S_RS_GAB, K2S_MARGIN_GAB, R, S, lane, out);
    else if (kind == 2) k2s_copy_behind(P, gab, K2S_RS_GAB, R, S, lane);
}

// ------------------------------------------------------------------------------------------------------------------------
// D: the distance map of one canonical offset d = (DY, DX) for stream row S (epfDistance1, Frame.java:638-655).
// T(q) = fl(fl|I(q) - I(q+d)| * s_c).  Rolling: A = T(y-1) (own 4 columns), B = T(y) (columns x-1 .. x+4), this step forms
// T(y+1) from the one row it loads (y+1+DY) and the rows it kept; dist(y) = sum over c of T(y,x) + T(y,x-1) + T(y,x+1) + T(y-1,x)
// + T(y+1,x), in that order.
// ------------------------------------------------------------------------------------------------------------------------
// ONE body for every canonical offset (dy, dx are run-time, warp-uniform): eight warps run it at once, and eight template instances
// of it were 27 KB of hot code in front of a 32 KB instruction cache (profiles/r2_k2_stream_ncu.md).
struct K2SDistState { float a[3][4], b[3][6], i1[3][4], i2[3][4]; };
K2S_FN void k2s_d_row(const K2Params &P, const float *in_row, int plane_stride, float *out_row, int dy, int dx, K2SDistState &st, int lane) {
    float dist[4];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const K2SQuad nq = k2s_ld4(in_row + c * plane_stride + 4 * lane);      // row y + 1 + dy
        const float nv[4] = {nq.x, nq.y, nq.z, nq.w};
        float base[4], sh[4];
#pragma unroll
        for (int j = 0; j < 4; j++) base[j] = dy == 0 ? nv[j] : st.i1[c][j];
        if (dx == 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) sh[j] = nv[j];
        } else if (dx == 1) {
            const float e4 = k2s_dn(nv[0]);
            sh[0] = nv[1]; sh[1] = nv[2]; sh[2] = nv[3]; sh[3] = e4;
        } else if (dx == 2) {
            const float e4 = k2s_dn(nv[0]), e5 = k2s_dn(nv[1]);
            sh[0] = nv[2]; sh[1] = nv[3]; sh[2] = e4; sh[3] = e5;
        } else {
            const float em = k2s_up(nv[3]);
            sh[0] = em; sh[1] = nv[0]; sh[2] = nv[1]; sh[3] = nv[2];
        }
        const float s = P.ch_scale[c];
        float t[4];
#pragma unroll
        for (int j = 0; j < 4; j++) t[j] = K2S_MUL(fabsf(K2S_SUB(base[j], sh[j])), s);
        const float tl = k2s_up(t[3]), tr = k2s_dn(t[0]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float d = c == 0 ? st.b[c][j + 1] : K2S_ADD(dist[j], st.b[c][j + 1]);   // (0, 0); the running sum starts at 0f and 0 + t == t
            d = K2S_ADD(d, st.b[c][j]);                                             // (0, -1)
            d = K2S_ADD(d, st.b[c][j + 2]);                                         // (0, +1)
            d = K2S_ADD(d, st.a[c][j]);                                             // (-1, 0)
            d = K2S_ADD(d, t[j]);                                                   // (+1, 0)
            dist[j] = d;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            st.a[c][j] = st.b[c][j + 1]; st.b[c][j + 1] = t[j];
            st.i1[c][j] = dy == 2 ? st.i2[c][j] : nv[j];
            st.i2[c][j] = nv[j];
        }
        st.b[c][0] = tl; st.b[c][5] = tr;
    }
    K2SQuad q; q.x = dist[0]; q.y = dist[1]; q.z = dist[2]; q.w = dist[3];
    k2s_st4(out_row + 4 * lane, q);
}

// ------------------------------------------------------------------------------------------------------------------------
// W0: pass 0 (13-point double cross, Frame.java:44-55 crossList order) of stream row S: weights from the six maps, channel sums,
// divide -> P0 ring.  Map order: 0 (0,1), 1 (1,0), 2 (1,1), 3 (1,-1), 4 (0,2), 5 (2,0); dist_{-d}(p) = map_d(p - d).
// ------------------------------------------------------------------------------------------------------------------------
K2S_FN void k2s_w0_row(const K2SArgs &A, float *sm, K2SCursor &cur, int S, int lane) {
    const K2Params &P = A.P;
    const K2SRow R = k2s_at(A, cur, S);
    const float *in = sm + K2S_OFF_GAB;
    float *ring = sm + K2S_OFF_P0;
    const int kind = R.valid ? k2s_row_kind(P, R.y, K2S_MARGIN_P0) : 1;
    if (kind == 1) return;
    if (kind == 2) { k2s_copy_behind(P, ring, K2S_RS_P0, R, S, lane); return; }
    const float *d0 = sm + K2S_OFF_D0;
    const int s0 = k2s_slot(S, K2S_RS_D0), s1 = k2s_slot(S - 1, K2S_RS_D0), s2 = k2s_slot(S - 2, K2S_RS_D0);
#define K2S_MAP(m, s) k2s_ld4(d0 + ((m) * K2S_RS_D0 + (s)) * K2S_PITCH + 4 * lane)
    const K2SQuad q01 = K2S_MAP(0, s0), q10 = K2S_MAP(1, s0), q11 = K2S_MAP(2, s0), q1m = K2S_MAP(3, s0), q02 = K2S_MAP(4, s0), q20 = K2S_MAP(5, s0);
    const K2SQuad p10 = K2S_MAP(1, s1), p11 = K2S_MAP(2, s1), p1m = K2S_MAP(3, s1), p20 = K2S_MAP(5, s2);
#undef K2S_MAP
    const float e01 = k2s_up(q01.w), e02a = k2s_up(q02.z), e02b = k2s_up(q02.w), e11 = k2s_up(p11.w), e1m = k2s_dn(p1m.x);
    const float v01[5] = {e01, q01.x, q01.y, q01.z, q01.w};              // map (0,1) at columns x-1 .. x+3
    const float v02[6] = {e02a, e02b, q02.x, q02.y, q02.z, q02.w};       // map (0,2) at columns x-2 .. x+3
    const float u11[5] = {e11, p11.x, p11.y, p11.z, p11.w};              // map (1,1), row y-1, columns x-1 .. x+3
    const float u1m[5] = {p1m.x, p1m.y, p1m.z, p1m.w, e1m};              // map (1,-1), row y-1, columns x .. x+4
    float m[4];
    const float is = k2s_sigma(A, R, lane, m);
    const bool pass = !(is <= (1.0f / 0.3f));                            // copied through (Frame.java:608-612); also NaN
    const float ss = P.sigma_scale[0];
    float w[4][12], sumw[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        w[j][0] = k2s_wgt(v01[j], m[j], ss, is);                         // (0,-1) = -(0,1): map at (y, x-1)
        w[j][1] = k2s_wgt(v01[j + 1], m[j], ss, is);                     // (0, 1)
        w[j][2] = k2s_wgt(k2s_get(p10, j), m[j], ss, is);                // (-1,0) = -(1,0): map at (y-1, x)
        w[j][3] = k2s_wgt(k2s_get(q10, j), m[j], ss, is);                // (1, 0)
        w[j][4] = k2s_wgt(u1m[j + 1], m[j], ss, is);                     // (-1,1) = -(1,-1): map at (y-1, x+1)
        w[j][5] = k2s_wgt(k2s_get(q11, j), m[j], ss, is);                // (1, 1)
        w[j][6] = k2s_wgt(k2s_get(q1m, j), m[j], ss, is);                // (1,-1)
        w[j][7] = k2s_wgt(u11[j], m[j], ss, is);                         // (-1,-1) = -(1,1): map at (y-1, x-1)
        w[j][8] = k2s_wgt(v02[j], m[j], ss, is);                         // (0,-2) = -(0,2): map at (y, x-2)
        w[j][9] = k2s_wgt(v02[j + 2], m[j], ss, is);                     // (0, 2)
        w[j][10] = k2s_wgt(k2s_get(q20, j), m[j], ss, is);               // (2, 0)
        w[j][11] = k2s_wgt(k2s_get(p20, j), m[j], ss, is);               // (-2,0) = -(2,0): map at (y-2, x)
        float s = 1.0f;                                                  // 0 + weight(centre) = 1
#pragma unroll
        for (int k = 0; k < 12; k++) s = K2S_ADD(s, w[j][k]);
        sumw[j] = s;
    }
    K2SQuad out[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *pl = in + c * K2S_RS_GAB * K2S_PITCH + 4 * lane;
        const K2SQuad a = k2s_ld4(pl + k2s_slot(S - 2, K2S_RS_GAB) * K2S_PITCH), b = k2s_ld4(pl + k2s_slot(S - 1, K2S_RS_GAB) * K2S_PITCH);
        const K2SQuad r = k2s_ld4(pl + k2s_slot(S, K2S_RS_GAB) * K2S_PITCH);
        const K2SQuad d = k2s_ld4(pl + k2s_slot(S + 1, K2S_RS_GAB) * K2S_PITCH), e = k2s_ld4(pl + k2s_slot(S + 2, K2S_RS_GAB) * K2S_PITCH);
        const float b6[6] = {k2s_up(b.w), b.x, b.y, b.z, b.w, k2s_dn(b.x)};
        const float d6[6] = {k2s_up(d.w), d.x, d.y, d.z, d.w, k2s_dn(d.x)};
        const float r8[8] = {k2s_up(r.z), k2s_up(r.w), r.x, r.y, r.z, r.w, k2s_dn(r.x), k2s_dn(r.y)};
        float o[4];
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
            o[j] = pass ? r8[j + 2] : K2S_DIV(s, sumw[j]);
        }
        out[c].x = o[0]; out[c].y = o[1]; out[c].z = o[2]; out[c].w = o[3];
    }
    k2s_emit(P, ring, K2S_RS_P0, K2S_MARGIN_P0, R, S, lane, out);
}

// ------------------------------------------------------------------------------------------------------------------------
// W1: pass 1 (5-point cross) of stream row S from the two maps; input ring = P0 (epf_iters == 3) or GAB.  LAST: epf_iters == 1,
// the row goes through the colour transform to HBM instead of the P1 ring.
// ------------------------------------------------------------------------------------------------------------------------
template <int RS_IN, bool LAST> K2S_FN void k2s_w1_row(const K2SArgs &A, float *sm, const float *in, K2SCursor &cur, int S, int lane) {
    const K2Params &P = A.P;
    const K2SRow R = k2s_at(A, cur, S);
    float *ring = sm + K2S_OFF_P1;
    const int kind = R.valid ? k2s_row_kind(P, R.y, LAST ? 0 : K2S_MARGIN_P1) : 1;
    if (kind == 1) return;
    if (kind == 2) { if (!LAST) k2s_copy_behind(P, ring, K2S_RS_P1, R, S, lane); return; }
    const float *d1 = sm + K2S_OFF_D1;
    const K2SQuad q01 = k2s_ld4(d1 + (0 * K2S_RS_D1 + k2s_slot(S, K2S_RS_D1)) * K2S_PITCH + 4 * lane);
    const K2SQuad q10 = k2s_ld4(d1 + (1 * K2S_RS_D1 + k2s_slot(S, K2S_RS_D1)) * K2S_PITCH + 4 * lane);
    const K2SQuad p10 = k2s_ld4(d1 + (1 * K2S_RS_D1 + k2s_slot(S - 1, K2S_RS_D1)) * K2S_PITCH + 4 * lane);
    const float v01[5] = {k2s_up(q01.w), q01.x, q01.y, q01.z, q01.w};
    float m[4];
    const float is = k2s_sigma(A, R, lane, m);
    const bool pass = !(is <= (1.0f / 0.3f));
    const float ss = P.sigma_scale[1];
    float w[4][4], sumw[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        w[j][0] = k2s_wgt(v01[j], m[j], ss, is);
        w[j][1] = k2s_wgt(v01[j + 1], m[j], ss, is);
        w[j][2] = k2s_wgt(k2s_get(p10, j), m[j], ss, is);
        w[j][3] = k2s_wgt(k2s_get(q10, j), m[j], ss, is);
        sumw[j] = K2S_ADD(K2S_ADD(K2S_ADD(K2S_ADD(1.0f, w[j][0]), w[j][1]), w[j][2]), w[j][3]);
    }
    K2SQuad out[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *pl = in + c * RS_IN * K2S_PITCH + 4 * lane;
        const K2SQuad b = k2s_ld4(pl + k2s_slot(S - 1, RS_IN) * K2S_PITCH), r = k2s_ld4(pl + k2s_slot(S, RS_IN) * K2S_PITCH);
        const K2SQuad d = k2s_ld4(pl + k2s_slot(S + 1, RS_IN) * K2S_PITCH);
        const float r6[6] = {k2s_up(r.w), r.x, r.y, r.z, r.w, k2s_dn(r.x)};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = r6[j + 1];
            s = K2S_ADD(s, K2S_MUL(r6[j], w[j][0]));
            s = K2S_ADD(s, K2S_MUL(r6[j + 2], w[j][1]));
            s = K2S_ADD(s, K2S_MUL(k2s_get(b, j), w[j][2]));
            s = K2S_ADD(s, K2S_MUL(k2s_get(d, j), w[j][3]));
            o[j] = pass ? r6[j + 1] : K2S_DIV(s, sumw[j]);
        }
        out[c].x = o[0]; out[c].y = o[1]; out[c].z = o[2]; out[c].w = o[3];
    }
    if (LAST) k2s_final(A, R, lane, out);
    else k2s_emit(P, ring, K2S_RS_P1, K2S_MARGIN_P1, R, S, lane, out);
}

// ------------------------------------------------------------------------------------------------------------------------
// P2: pass 2 (5-point cross, point differences: epfDistance2, Frame.java:657-669) of stream row S from the P1 ring, then the
// colour transform and the store.  Everything stays in registers.
// ------------------------------------------------------------------------------------------------------------------------
K2S_FN void k2s_p2_row(const K2SArgs &A, float *sm, K2SCursor &cur, int S, int lane) {
    const K2Params &P = A.P;
    const K2SRow R = k2s_at(A, cur, S);
    if (!R.valid || k2s_row_kind(P, R.y, 0) != 0) return;
    const float *in = sm + K2S_OFF_P1;
    float r6[3][6], up[3][4], dn[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *pl = in + c * K2S_RS_P1 * K2S_PITCH + 4 * lane;
        const K2SQuad b = k2s_ld4(pl + k2s_slot(S - 1, K2S_RS_P1) * K2S_PITCH), r = k2s_ld4(pl + k2s_slot(S, K2S_RS_P1) * K2S_PITCH);
        const K2SQuad d = k2s_ld4(pl + k2s_slot(S + 1, K2S_RS_P1) * K2S_PITCH);
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
    float m[4];
    const float is = k2s_sigma(A, R, lane, m);
    const bool pass = !(is <= (1.0f / 0.3f));
    const float ss = P.sigma_scale[2];
    float o[3][4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float w0 = k2s_wgt(h[j], m[j], ss, is), w1 = k2s_wgt(h[j + 1], m[j], ss, is);
        const float w2 = k2s_wgt(vu[j], m[j], ss, is), w3 = k2s_wgt(vd[j], m[j], ss, is);
        const float sumw = K2S_ADD(K2S_ADD(K2S_ADD(K2S_ADD(1.0f, w0), w1), w2), w3);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float s = r6[c][j + 1];
            s = K2S_ADD(s, K2S_MUL(r6[c][j], w0));
            s = K2S_ADD(s, K2S_MUL(r6[c][j + 2], w1));
            s = K2S_ADD(s, K2S_MUL(up[c][j], w2));
            s = K2S_ADD(s, K2S_MUL(dn[c][j], w3));
            o[c][j] = pass ? r6[c][j + 1] : K2S_DIV(s, sumw);
        }
    }
    K2SQuad out[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { out[c].x = o[c][0]; out[c].y = o[c][1]; out[c].z = o[c][2]; out[c].w = o[c][3]; }
    k2s_final(A, R, lane, out);
}

// ------------------------------------------------------------------------------------------------------------------------
// the tick loops of the roles.  n_ticks is the same for every warp; each role's loop carries its own rolling state, so the
// kernel's register count is the maximum over the roles, not their sum.
// ------------------------------------------------------------------------------------------------------------------------
K2S_FN int k2s_total_rows(const K2SArgs &A) {
    const int cta = k2s_cta(), g = k2s_grid();
    const int mine = cta < A.n_items ? (A.n_items - cta + g - 1) / g : 0;
    return mine * A.ir;
}

template <int GAB, int ITERS>
K2S_FN void k2s_role_g(const K2SArgs &A, float *sm, uint64_t *bars, K2S_TMAP_PARAM t0, K2S_TMAP_PARAM t1, K2S_TMAP_PARAM t2, int n_ticks, int total, int lane) {
    using Cfg = K2SCfg<ITERS>;
    K2SGabState st;
    K2SCursor cur;
    k2s_cursor_init(cur);
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int i = 0; i < 6; i++) { st.n[c][i] = 0.0f; st.r[c][i] = 0.0f; }
    for (int t = 0; t < n_ticks; t++) {
        // rows [8t, 8t+8) of the stream -> RAW ring, two 4-row boxes per plane; the slots were last read in tick t-1
        if (8 * t < total) {
#ifdef K2S_HOST_EMU
            if (lane == 0)
                for (int h = 0; h < 2; h++) {
                    const K2SRow R = k2s_locate(A, 8 * t + 4 * h);
                    const int ty = R.z * A.P.rows + R.y + A.tma_row0, slot = k2s_slot(8 * t + 4 * h, K2S_RS_RAW);
                    k2s_tma_box(sm + K2S_OFF_RAW + (0 * K2S_RS_RAW + slot) * K2S_PITCH, t0, R.x0 - K2S_HALO, ty);
                    k2s_tma_box(sm + K2S_OFF_RAW + (1 * K2S_RS_RAW + slot) * K2S_PITCH, t1, R.x0 - K2S_HALO, ty);
                    k2s_tma_box(sm + K2S_OFF_RAW + (2 * K2S_RS_RAW + slot) * K2S_PITCH, t2, R.x0 - K2S_HALO, ty);
                }
#else
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of tick t-1 before async-proxy writes
                uint64_t *bar = bars + (t & 1);
                k2s_mbar_expect_tx(bar, 6 * 4 * K2S_PITCH * 4);
                for (int h = 0; h < 2; h++) {
                    const K2SRow R = k2s_locate(A, 8 * t + 4 * h);
                    const int ty = R.z * A.P.rows + R.y + A.tma_row0, slot = k2s_slot(8 * t + 4 * h, K2S_RS_RAW);
                    k2s_tma_box_async(sm + K2S_OFF_RAW + (0 * K2S_RS_RAW + slot) * K2S_PITCH, t0, R.x0 - K2S_HALO, ty, bar);
                    k2s_tma_box_async(sm + K2S_OFF_RAW + (1 * K2S_RS_RAW + slot) * K2S_PITCH, t1, R.x0 - K2S_HALO, ty, bar);
                    k2s_tma_box_async(sm + K2S_OFF_RAW + (2 * K2S_RS_RAW + slot) * K2S_PITCH, t2, R.x0 - K2S_HALO, ty, bar);
                }
            }
#endif
        }
#ifndef K2S_HOST_EMU
        // the rows this tick reads arrived with the loads of tick t-1
        if (t >= 1 && 8 * (t - 1) < total) k2s_mbar_wait(bars + ((t - 1) & 1), ((t - 1) >> 1) & 1);
#endif
#pragma unroll 1
        for (int r = 0; r < K2S_BAND; r++) {
            const int S = 8 * t + Cfg::G + r;
            // S == -1 primes the rolling window with stream row 0 (the row itself is not emitted)
            if (S >= -1 && S < total) k2s_g_row<GAB>(A, sm, st, cur, S, lane);
        }
        k2s_sync();
    }
}

K2S_FN void k2s_role_d(const K2SArgs &A, const float *in, int rs_in, float *outmap, int rs_out, int dy, int dx, int base, int n_ticks, int total, int lane) {
    K2SDistState st;
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int i = 0; i < 6; i++) st.b[c][i] = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) { st.a[c][i] = 0.0f; st.i1[c][i] = 0.0f; st.i2[c][i] = 0.0f; }
    }
    const int plane_stride = rs_in * K2S_PITCH;
    for (int t = 0; t < n_ticks; t++) {
        int si = k2s_slot(8 * t + base + 1 + dy, rs_in), so = k2s_slot(8 * t + base, rs_out);     // ring slots advance with the rows
#pragma unroll 1
        for (int r = 0; r < K2S_BAND; r++) {
            const int S = 8 * t + base + r;
            // steps -3 .. -1 prime the rolling rows (row r is loaded at step r - 1 - dy); what they store lands in ring slots nobody has used yet
            if (S >= -3 && S < total) k2s_d_row(A.P, in + si * K2S_PITCH, plane_stride, outmap + so * K2S_PITCH, dy, dx, st, lane);
            si = si + 1 == rs_in ? 0 : si + 1;
            so = so + 1 == rs_out ? 0 : so + 1;
        }
        k2s_sync();
    }
}

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
    float *gab = sm + K2S_OFF_GAB, *p0 = sm + K2S_OFF_P0, *d0 = sm + K2S_OFF_D0, *d1 = sm + K2S_OFF_D1;
    const float *in1 = ITERS == 3 ? p0 : gab;                 // what pass 1 reads
    K2SCursor cur;
    k2s_cursor_init(cur);
    constexpr int RS1 = ITERS == 3 ? K2S_RS_P0 : K2S_RS_GAB;
    // role of each warp.  Warp w issues on scheduler w % 4; heavy and light roles are interleaved so the four schedulers carry about
    // the same number of instructions per tick (DESIGN.md has the budget).
    if (ITERS == 3) {
        //  warp:  0  1  2  3 | 4  5  6  7 | 8  9  10 11 | 12 13 14 15 | 16 17
        //  role:  W0 W0 W0 W0| D0 D0 D0 D0| P2 P2 W1 W1 | D0 D0 P2 G  | D1 D1
        if (warp < 4) {
            for (int t = 0; t < n_ticks; t++) {
#pragma unroll 1
                for (int r = 0; r < 2; r++) {
                    const int S = 8 * t + Cfg::W0 + 2 * warp + r;
                    if (S >= 0 && S < total) k2s_w0_row(A, sm, cur, S, lane);
                }
                k2s_sync();
            }
        } else if ((warp >= 4 && warp < 8) || warp == 12 || warp == 13 || warp >= 16) {
            // distance roles: pass 0 offsets (0,1) (1,0) (1,1) (1,-1) (0,2) (2,0) on warps 4-7, 12, 13; pass 1 offsets (0,1) (1,0) on warps 16, 17
            const bool p1 = warp >= 16;
            const int m = p1 ? warp - 16 : (warp < 8 ? warp - 4 : warp - 8);          // map index
            const int dy = p1 ? m : (m == 0 || m == 4 ? 0 : (m == 5 ? 2 : 1));
            const int dx = p1 ? 1 - m : (m == 0 ? 1 : m == 2 ? 1 : m == 3 ? -1 : m == 4 ? 2 : 0);
            k2s_role_d(A, p1 ? in1 : gab, p1 ? RS1 : K2S_RS_GAB, p1 ? d1 + m * K2S_RS_D1 * K2S_PITCH : d0 + m * K2S_RS_D0 * K2S_PITCH,
                       p1 ? K2S_RS_D1 : K2S_RS_D0, dy, dx, p1 ? Cfg::D1 : Cfg::D0, n_ticks, total, lane);
        }
        else if (warp == 8 || warp == 9 || warp == 14) {
            const int first = warp == 8 ? 0 : warp == 9 ? 3 : 6, count = warp == 14 ? 2 : 3;
            for (int t = 0; t < n_ticks; t++) {
#pragma unroll 1
                for (int r = 0; r < count; r++) {
                    const int S = 8 * t + Cfg::P2 + first + r;
                    if (S >= 0 && S < total) k2s_p2_row(A, sm, cur, S, lane);
                }
                k2s_sync();
            }
        } else if (warp == 10 || warp == 11) {
            for (int t = 0; t < n_ticks; t++) {
#pragma unroll 1
                for (int r = 0; r < 4; r++) {
                    const int S = 8 * t + Cfg::W1 + 4 * (warp - 10) + r;
                    if (S >= 0 && S < total) k2s_w1_row<RS1, false>(A, sm, in1, cur, S, lane);
                }
                k2s_sync();
            }
        } else k2s_role_g<GAB, ITERS>(A, sm, bars, t0, t1, t2, n_ticks, total, lane);      // warp 15
    } else {
        //  warp:  0  1  2  3 | 4  5  6  7 | 8  9  10        (epf_iters 2: W1 x4 then P2 x4;  epf_iters 1: W1 x8, the last stage)
        //  role:  W1 W1 W1 W1| P2 P2 P2 P2| D1 D1 G
        if (warp < 8 && (ITERS == 1 || warp < 4)) {
            constexpr int per = ITERS == 1 ? 1 : 2;
            for (int t = 0; t < n_ticks; t++) {
#pragma unroll 1
                for (int r = 0; r < per; r++) {
                    const int S = 8 * t + Cfg::W1 + per * warp + r;
                    if (S >= 0 && S < total) k2s_w1_row<RS1, ITERS == 1>(A, sm, in1, cur, S, lane);
                }
                k2s_sync();
            }
        } else if (warp < 8) {
            for (int t = 0; t < n_ticks; t++) {
#pragma unroll 1
                for (int r = 0; r < 2; r++) {
                    const int S = 8 * t + Cfg::P2 + 2 * (warp - 4) + r;
                    if (S >= 0 && S < total) k2s_p2_row(A, sm, cur, S, lane);
                }
                k2s_sync();
            }
        } else if (warp == 8 || warp == 9) {
            const int m = warp - 8;
            k2s_role_d(A, in1, RS1, d1 + m * K2S_RS_D1 * K2S_PITCH, K2S_RS_D1, m, 1 - m, Cfg::D1, n_ticks, total, lane);
        } else k2s_role_g<GAB, ITERS>(A, sm, bars, t0, t1, t2, n_ticks, total, lane);
    }
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
__global__ void __launch_bounds__(32 * K2SCfg<ITERS>::NWARPS, 1)
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
    if (K.iters < 1 || K.rows < 8 || K.W < 8 || (K.in_pitch & 3) || (K.out_pitch & 3) || !k2s_encoder()) return false;
    for (int c = 0; c < 3; c++)
        if (((uintptr_t)K.in[c] | (uintptr_t)K.out[c]) & 15) return false;
    if ((long long)K.rows * n_frames + 16 > 0x7fffffffll) return false;
    return true;
}

// returns 0, or -1 when a tensor map could not be encoded (the caller falls back to k2_exact)
static inline int k2_stream_launch(const K2Params &K, const float *inv_sigma, cudaStream_t st, int n_frames, int sms) {
    K2SArgs A;
    A.P = K;
    A.inv_sigma = inv_sigma;
    A.zpx = (long long)K.rows * K.in_pitch;
    A.zblk = (K.rows >> 3) * K.wb;
    k2s_plan(K.W, K.rows, n_frames, sms, A);
    A.tma_row0 = K.has_top ? JXLB200_HALO_ROWS : 0;
    const long long map_rows = (long long)K.rows * n_frames + (K.has_top ? JXLB200_HALO_ROWS : 0) + (K.has_bottom ? JXLB200_HALO_ROWS : 0);
    CUtensorMap tm[3];
    for (int c = 0; c < 3; c++) {
        const cuuint64_t dims[2] = {(cuuint64_t)K.W, (cuuint64_t)map_rows};
        const cuuint64_t strides[1] = {(cuuint64_t)K.in_pitch * 4};
        const cuuint32_t box[2] = {K2S_PITCH, 4}, es[2] = {1, 1};
        void *base = (void *)(K.in[c] - (K.has_top ? (long long)JXLB200_HALO_ROWS * K.in_pitch : 0));
        if (k2s_encoder()(&tm[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
    }
    const int grid = A.n_items < sms ? A.n_items : sms;
#define K2S_GO(G, I) k2_stream<G, I><<<grid, 32 * K2SCfg<I>::NWARPS, K2S_BYTES, st>>>(A, tm[0], tm[1], tm[2])
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
