// Host emulation of the stage-2 stream kernel (jxlatte_b200/csrc/k2_stream.cuh): the SAME source, compiled with
// K2S_HOST_EMU, run with one host thread per CUDA thread -- 32 per warp, the shuffles and the CTA barrier emulated with
// pthread barriers, the TMA boxes with a zero-filling copy.  The roles of a CTA run truly concurrently, as on the SM, so a
// row handed over a tick too early shows up as a mismatch against the oracle.  tests/test_k2_stream_host.py drives it.
// Built with -ffp-contract=off: the reference-order sums must not be contracted into FMAs.
#define K2S_HOST_EMU 1
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../jxlatte_b200/csrc/k2_stream.cuh"

namespace {
struct Emu {
    int n_warps = 0, cta = 0, grid = 1;
    pthread_barrier_t cta_bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<float> xch;
    int any[2] = {0, 0};
    std::vector<int> any_phase;
} g;
thread_local int t_tid = 0;
}  // namespace

int k2s_emu_tid() { return t_tid; }
int k2s_emu_cta() { return g.cta; }
int k2s_emu_grid() { return g.grid; }
float k2s_emu_shfl(float v, int src_lane) {
    const int w = t_tid >> 5, l = t_tid & 31;
    g.xch[w * 32 + l] = v;
    pthread_barrier_wait(&g.warp_bar[w]);
    const float r = g.xch[w * 32 + (src_lane & 31)];
    pthread_barrier_wait(&g.warp_bar[w]);
    return r;
}
int k2s_emu_any(int pred) {
    const int w = t_tid >> 5, l = t_tid & 31;
    g.xch[w * 32 + l] = pred ? 1.0f : 0.0f;
    pthread_barrier_wait(&g.warp_bar[w]);
    int r = 0;
    for (int i = 0; i < 32; i++) r |= g.xch[w * 32 + i] != 0.0f;
    pthread_barrier_wait(&g.warp_bar[w]);
    return r;
}
void k2s_emu_sync() { pthread_barrier_wait(&g.cta_bar); }
int k2s_emu_sync_or(int pred) {
    if (pred) __atomic_store_n(&g.any[g.any_phase[t_tid] & 1], 1, __ATOMIC_RELAXED);
    pthread_barrier_wait(&g.cta_bar);
    const int r = __atomic_load_n(&g.any[g.any_phase[t_tid] & 1], __ATOMIC_RELAXED);
    pthread_barrier_wait(&g.cta_bar);
    // the flag of this parity is cleared by thread 0 only after everybody has read it; the next sync_or uses the other flag
    if (t_tid == 0) __atomic_store_n(&g.any[g.any_phase[t_tid] & 1], 0, __ATOMIC_RELAXED);
    g.any_phase[t_tid]++;
    return r;
}

namespace {
struct Job {
    const K2SArgs *A;
    float *sm;
    const K2STmap *tm;
    int gab, iters, tid;
};
void *thread_main(void *arg) {
    const Job *j = (const Job *)arg;
    t_tid = j->tid;
    const K2STmap &a = j->tm[0], &b = j->tm[1], &c = j->tm[2];
    switch (j->gab * 4 + j->iters) {
    case 5: k2s_body<1, 1>(*j->A, j->sm, nullptr, a, b, c); break;
    case 6: k2s_body<1, 2>(*j->A, j->sm, nullptr, a, b, c); break;
    case 7: k2s_body<1, 3>(*j->A, j->sm, nullptr, a, b, c); break;
    case 1: k2s_body<0, 1>(*j->A, j->sm, nullptr, a, b, c); break;
    case 2: k2s_body<0, 2>(*j->A, j->sm, nullptr, a, b, c); break;
    default: k2s_body<0, 3>(*j->A, j->sm, nullptr, a, b, c); break;
    }
    return nullptr;
}

// fill_k2 of jxlb200.cu, restated for the harness (Frame.performGabConvolution :510-517, Frame.java:545, invertXYB :108-119)
void fill(K2Params &K, const jxlb200_frame_params *p, const jxlb200_slab *slab) {
    memset(&K, 0, sizeof(K));
    K.W = p->width;
    K.rows = slab ? slab->rows : p->height;
    K.y0 = slab ? slab->y0 : 0;
    K.frame_h = slab ? slab->frame_height : p->height;
    K.has_top = slab ? slab->has_top : 0;
    K.has_bottom = slab ? slab->has_bottom : 0;
    K.wb = p->width >> 3;
    K.gab = p->gab; K.iters = p->epf_iters; K.color_mode = p->color_mode;
    for (int c = 0; c < 3; c++) {
        const float w1 = p->gab_w1[c], w2 = p->gab_w2[c];
        const float mult = 1.0f / (1.0f + 4.0f * (w1 + w2));
        K.gab_base[c] = mult; K.gab_adj[c] = w1 * mult; K.gab_diag[c] = w2 * mult;
        K.ch_scale[c] = p->epf_channel_scale[c];
    }
    K.gscale = 65536.0f / p->global_scale;
    for (int i = 0; i < 8; i++) K.sharp_lut[i] = p->epf_sharp_lut[i];
    const float step = 1.65f * 4.0f * (1.0f - (float)sqrt(0.5));
    K.sigma_scale[0] = step * p->epf_pass0_sigma_scale;
    K.sigma_scale[1] = step;
    K.sigma_scale[2] = step * p->epf_pass2_sigma_scale;
    K.border_mul = p->epf_border_sad_mul;
    const float it = 255.0f / p->intensity_target;
    for (int i = 0; i < 9; i++) K.m[i] = p->opsin_matrix[i] * it;
    for (int c = 0; c < 3; c++) { K.ob[c] = p->opsin_bias[c]; K.cob[c] = -(float)cbrt((double)p->opsin_bias[c]); }
}
}  // namespace

// in[c] / hf_mul / sharpness point at the slab's first own row / block row; when the slab has a neighbour the 8 rows / one
// block row beyond must be there (the layout jxlb200_restore_dev takes).  n_frames > 1: frames stacked vertically, no slab.
// grid = number of CTAs to emulate (they run one after the other).
extern "C" int k2s_host_run(const jxlb200_frame_params *p, const jxlb200_slab *slab, int n_frames, const float *const in[3],
                            long long in_pitch, const int32_t *hf_mul, const int32_t *sharpness, float *const out[3], int grid) {
    if (p->epf_iters < 1 || p->epf_iters > 3) return -1;
    K2SArgs A;
    memset(&A, 0, sizeof(A));
    fill(A.P, p, slab);
    K2Params &K = A.P;
    for (int c = 0; c < 3; c++) { K.in[c] = in[c]; K.out[c] = out[c]; }
    K.in_pitch = in_pitch; K.out_pitch = K.W;
    // k2_sigma (Frame.java:552-572): one float per block, one extra block row towards each neighbour slab
    const int br0 = K.has_top ? -1 : 0, br1 = (K.rows / 8) * n_frames + (K.has_bottom ? 1 : 0);
    std::vector<float> sigma((size_t)(br1 - br0 + 2) * K.wb, 0.0f);
    float *inv_sigma = sigma.data() + K.wb;
    for (int i = br0 * K.wb; i < br1 * K.wb; i++) {
        const int s = sharpness[i];
        if (s < 0 || s > 7) return -2;
        const float sg = (K.gscale * K.sharp_lut[s]) / (float)hf_mul[i];
        inv_sigma[i] = 1.0f / sg;
    }
    A.inv_sigma = inv_sigma;
    A.zpx = (long long)K.rows * K.in_pitch;
    A.zblk = (K.rows >> 3) * K.wb;
    k2s_plan(K.W, K.rows, n_frames, grid, A);
    A.tma_row0 = K.has_top ? JXLB200_HALO_ROWS : 0;
    K2STmap tm[3];
    for (int c = 0; c < 3; c++) {
        tm[c].base = in[c] - (K.has_top ? (long long)JXLB200_HALO_ROWS * in_pitch : 0);
        tm[c].w = K.W;
        tm[c].rows = K.rows * n_frames + (K.has_top ? JXLB200_HALO_ROWS : 0) + (K.has_bottom ? JXLB200_HALO_ROWS : 0);
        tm[c].pitch = in_pitch;
    }
    const int nw = K2S_NWARPS;
    g.n_warps = nw;
    g.grid = A.n_items < grid ? A.n_items : grid;
    g.xch.assign((size_t)nw * 32, 0.0f);
    g.any_phase.assign((size_t)nw * 32, 0);
    g.any[0] = g.any[1] = 0;
    g.warp_bar.resize(nw);
    std::vector<float> sm(K2S_FLOATS + 16);
    for (int cta = 0; cta < g.grid; cta++) {
        g.cta = cta;
        // garbage, not zeros: nothing may depend on what a ring held before it was written
        for (size_t i = 0; i < sm.size(); i++) sm[i] = 1.0e30f * (float)((i * 2654435761u) & 0xff);
        pthread_barrier_init(&g.cta_bar, nullptr, nw * 32);
        for (int w = 0; w < nw; w++) pthread_barrier_init(&g.warp_bar[w], nullptr, 32);
        std::vector<pthread_t> th(nw * 32);
        std::vector<Job> jobs(nw * 32);
        pthread_attr_t at;
        pthread_attr_init(&at);
        pthread_attr_setstacksize(&at, 256 * 1024);
        for (int t = 0; t < nw * 32; t++) {
            jobs[t] = Job{&A, sm.data(), tm, p->gab ? 1 : 0, p->epf_iters, t};
            if (pthread_create(&th[t], &at, thread_main, &jobs[t]) != 0) return -3;
        }
        for (int t = 0; t < nw * 32; t++) pthread_join(th[t], nullptr);
        pthread_attr_destroy(&at);
        pthread_barrier_destroy(&g.cta_bar);
        for (int w = 0; w < nw; w++) pthread_barrier_destroy(&g.warp_bar[w]);
    }
    return 0;
}

// the plan alone (how a frame is cut into items), for the CPU tests
extern "C" void k2s_host_plan(int W, int rows, int n_frames, int n_cta, int out[5]) {
    K2SArgs A;
    memset(&A, 0, sizeof(A));
    k2s_plan(W, rows, n_frames, n_cta, A);
    out[0] = A.ch; out[1] = A.ir; out[2] = A.n_cols; out[3] = A.n_chunks; out[4] = A.n_items;
}
