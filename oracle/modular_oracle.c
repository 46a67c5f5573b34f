/*
 * oracle/modular_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see jxl_oracle.h).
 *
 * Literal restatement of jxlatte's Modular inverse transforms (RCT, Palette, Squeeze).  int32 with Java
 * semantics: wrapping overflow (-fwrapv), arithmetic >>, truncating / and %.
 * "J/" = /root/reference/java/com/traneptora/jxlatte/.  PARITY UNPINNED (no JVM here).
 */
#include "jxl_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ModularChannel.tendency (J/frame/modular/ModularChannel.java:23-47) */
int32_t orc_tendency(int32_t a, int32_t b, int32_t c) {
    if (a >= b && b >= c) {
        int x = (4 * a - 3 * c - b + 6) / 12;
        int d = 2 * (a - b);
        int e = 2 * (b - c);
        if ((x - (x & 1)) > d) x = d + 1;
        if ((x + (x & 1)) > e) x = e;
        return x;
    }
    if (a <= b && b <= c) {
        int x = (4 * a - 3 * c - b - 6) / 12;
        int d = 2 * (a - b);
        int e = 2 * (b - c);
        if ((x + (x & 1)) < d) x = d - 1;
        if ((x - (x & 1)) < e) x = e;
        return x;
    }
    return 0;
}

/* ModularChannel.inverseHorizontalSqueeze :361-387 / inverseVerticalSqueeze :389-413.
 * avg: h_avg x w_avg, res: h_res x w_res, out: (h_avg x (w_avg+w_res)) or ((h_avg+h_res) x w_avg). */
int32_t orc_modular_squeeze(const int32_t *avg, const int32_t *res, int32_t h_avg, int32_t w_avg,
    int32_t h_res, int32_t w_res, int32_t horizontal, int32_t *out) {
    if (horizontal) {
        const int W = w_avg + w_res, H = h_avg;
        if ((w_avg != w_res && w_avg != 1 + w_res) || h_res != h_avg) return -1; /* "Corrupted squeeze transform" */
        for (int y = 0; y < H; y++) {
            for (int x = 0; x < w_res; x++) {
                int a = avg[(size_t)y * w_avg + x];
                int residu = res[(size_t)y * w_res + x];
                int nextAvg = x + 1 < w_avg ? avg[(size_t)y * w_avg + x + 1] : a;
                int left = x > 0 ? out[(size_t)y * W + 2 * x - 1] : a;
                int diff = residu + orc_tendency(left, a, nextAvg);
                int first = a + diff / 2;
                out[(size_t)y * W + 2 * x] = first;
                out[(size_t)y * W + 2 * x + 1] = first - diff;
            }
        }
        if (w_avg > w_res) {
            const int xs = 2 * w_res;
            for (int y = 0; y < H; y++) out[(size_t)y * W + xs] = avg[(size_t)y * w_avg + w_res];
        }
    } else {
        const int W = w_avg;
        if ((h_avg != h_res && h_avg != 1 + h_res) || w_res != w_avg) return -1;
        for (int y = 0; y < h_res; y++) {
            for (int x = 0; x < W; x++) {
                int a = avg[(size_t)y * W + x];
                int residu = res[(size_t)y * W + x];
                int nextAvg = y + 1 < h_avg ? avg[(size_t)(y + 1) * W + x] : a;
                int top = y > 0 ? out[(size_t)(2 * y - 1) * W + x] : a;
                int diff = residu + orc_tendency(top, a, nextAvg);
                int first = a + diff / 2;
                out[(size_t)(2 * y) * W + x] = first;
                out[(size_t)(2 * y + 1) * W + x] = first - diff;
            }
        }
        if (h_avg > h_res)
            memcpy(out + (size_t)(2 * h_res) * W, avg + (size_t)h_res * W, sizeof(int32_t) * W);
    }
    return 0;
}

/* Forward squeeze, derived from the inverse above (tests only: inverse(forward(x)) == x). */
void orc_modular_forward_squeeze(const int32_t *in, int32_t h, int32_t w, int32_t horizontal, int32_t *avg, int32_t *res) {
    if (horizontal) {
        const int wa = (w + 1) / 2, wr = w / 2;
        for (int y = 0; y < h; y++) {
            const int32_t *row = in + (size_t)y * w;
            for (int x = 0; x < wr; x++) {
                int A = row[2 * x], B = row[2 * x + 1];
                avg[(size_t)y * wa + x] = (A + B + (A > B ? 1 : 0)) >> 1;
            }
            if (wa > wr) avg[(size_t)y * wa + wr] = row[2 * wr];
            for (int x = 0; x < wr; x++) {
                int A = row[2 * x], B = row[2 * x + 1];
                int a = avg[(size_t)y * wa + x];
                int nextAvg = x + 1 < wa ? avg[(size_t)y * wa + x + 1] : a;
                int left = x > 0 ? row[2 * x - 1] : a;
                res[(size_t)y * wr + x] = (A - B) - orc_tendency(left, a, nextAvg);
            }
        }
    } else {
        const int ha = (h + 1) / 2, hr = h / 2;
        for (int x = 0; x < w; x++) {
            for (int y = 0; y < hr; y++) {
                int A = in[(size_t)(2 * y) * w + x], B = in[(size_t)(2 * y + 1) * w + x];
                avg[(size_t)y * w + x] = (A + B + (A > B ? 1 : 0)) >> 1;
            }
            if (ha > hr) avg[(size_t)hr * w + x] = in[(size_t)(2 * hr) * w + x];
            for (int y = 0; y < hr; y++) {
                int A = in[(size_t)(2 * y) * w + x], B = in[(size_t)(2 * y + 1) * w + x];
                int a = avg[(size_t)y * w + x];
                int nextAvg = y + 1 < ha ? avg[(size_t)(y + 1) * w + x] : a;
                int top = y > 0 ? in[(size_t)(2 * y - 1) * w + x] : a;
                res[(size_t)y * w + x] = (A - B) - orc_tendency(top, a, nextAvg);
            }
        }
    }
}

/* ModularStream.applyTransforms RCT branch (J/frame/modular/ModularStream.java:35-38, 255-326).
 * In place: afterwards ch[k] holds what channels.get(beginC + k) holds in the reference. */
static const int permutationLut[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};

void orc_modular_rct(int32_t *const ch[3], int32_t h, int32_t w, int32_t rct_type, int32_t *const out[3]) {
    int permutation = rct_type / 7;
    int type = rct_type % 7;
    int32_t *v0 = ch[0], *v1 = ch[1], *v2 = ch[2];
    const size_t n = (size_t)h * w;
    switch (type) {
    case 0: break;
    case 1: for (size_t i = 0; i < n; i++) v2[i] += v0[i]; break;
    case 2: for (size_t i = 0; i < n; i++) v1[i] += v0[i]; break;
    case 3: for (size_t i = 0; i < n; i++) { const int a = v0[i]; v2[i] += a; v1[i] += a; } break;
    case 4: for (size_t i = 0; i < n; i++) v1[i] += (v0[i] + v2[i]) >> 1; break;
    case 5:
        for (size_t i = 0; i < n; i++) {
            const int a = v0[i];
            const int ac = a + v2[i];
            v1[i] += (a + ac) >> 1;
            v2[i] = ac;
        }
        break;
    case 6:
        for (size_t i = 0; i < n; i++) {
            const int b = v1[i];
            const int c = v2[i];
            const int tmp = v0[i] - (c >> 1);
            const int f = tmp - (b >> 1);
            v0[i] = f + b;
            v1[i] = c + tmp;
            v2[i] = f;
        }
        break;
    }
    /* channels.set(start + permutationLut[permutation][j], v[j]) :324-325 */
    for (int j = 0; j < 3; j++)
        memcpy(out[permutationLut[permutation][j]], ch[j], sizeof(int32_t) * n);
}

/* ---- Palette (J/frame/modular/ModularStream.java:20-33, 327-378; ModularChannel.java:95-183) ---- */
static const int kDeltaPalette[72][3] = {
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

typedef struct { int32_t *buffer; int h, w; } chan_t;
#define B(c, y, x) ((c)->buffer[(size_t)(y) * (c)->w + (x)])
static int west(const chan_t *c, int x, int y) { return x > 0 ? B(c, y, x - 1) : y > 0 ? B(c, y - 1, x) : 0; }
static int north(const chan_t *c, int x, int y) { return y > 0 ? B(c, y - 1, x) : x > 0 ? B(c, y, x - 1) : 0; }
static int northWest(const chan_t *c, int x, int y) {
    return x > 0 ? (y > 0 ? B(c, y - 1, x - 1) : B(c, y, x - 1)) : (y > 0 ? B(c, y - 1, x) : 0);
}
static int northEast(const chan_t *c, int x, int y) { return x + 1 < c->w && y > 0 ? B(c, y - 1, x + 1) : north(c, x, y); }
static int northNorth(const chan_t *c, int x, int y) { return y > 1 ? B(c, y - 2, x) : north(c, x, y); }
static int northEastEast(const chan_t *c, int x, int y) { return x + 2 < c->w && y > 0 ? B(c, y - 1, x + 2) : northEast(c, x, y); }
static int westWest(const chan_t *c, int x, int y) { return x > 1 ? B(c, y, x - 2) : west(c, x, y); }
static int iabs(int v) { return v < 0 ? -v : v; }
/* MathHelper.clamp(int v, int a, int b) (J/util/MathHelper.java:209-213) */
static int clamp3(int v, int a, int b) {
    int lower = a < b ? a : b;
    int upper = lower ^ a ^ b;
    return v < lower ? lower : v > upper ? upper : v;
}

/* ModularChannel.prediction :143-183 (k == 6 needs the WP state the palette copies never have) */
static int prediction(const chan_t *c, int y, int x, int k) {
    int n, v, nw, w;
    switch (k) {
    case 0: return 0;
    case 1: return x > 0 ? B(c, y, x - 1) : y > 0 ? B(c, y - 1, x) : 0;
    case 2: return y > 0 ? B(c, y - 1, x) : x > 0 ? B(c, y, x - 1) : 0;
    case 3: return (west(c, x, y) + north(c, x, y)) / 2;
    case 4:
        w = west(c, x, y); n = north(c, x, y); nw = northWest(c, x, y);
        return iabs(n - nw) < iabs(w - nw) ? w : n;
    case 5:
        w = west(c, x, y); n = north(c, x, y);
        v = w + n - northWest(c, x, y);
        return clamp3(v, n, w);
    case 7: return northEast(c, x, y);
    case 8: return northWest(c, x, y);
    case 9: return westWest(c, x, y);
    case 10: return (west(c, x, y) + northWest(c, x, y)) / 2;
    case 11: return (north(c, x, y) + northWest(c, x, y)) / 2;
    case 12: return (north(c, x, y) + northEast(c, x, y)) / 2;
    case 13:
        return (6 * north(c, x, y) - 2 * northNorth(c, x, y) + 7 * west(c, x, y) + westWest(c, x, y)
                + northEastEast(c, x, y) + 3 * northEast(c, x, y) + 8) / 16;
    default: return 0;
    }
}

/* palette: num_c x nb_colors (meta channel 0, row c = output channel c). out[c]: h x w, c < num_c. */
int32_t orc_modular_palette(const int32_t *idx, const int32_t *palette, int32_t h, int32_t w,
    int32_t num_c, int32_t nb_colors, int32_t nb_deltas, int32_t d_pred, int32_t bit_depth, int32_t *const out[]) {
    if (d_pred < 0 || d_pred > 13) return -1;
    if (d_pred == 6 && nb_deltas > 0) return -3; /* reference dereferences pred[][] == null */
    for (int c = 0; c < num_c; c++) {
        chan_t chan = {out[c], h, w};
        memcpy(out[c], idx, sizeof(int32_t) * (size_t)h * w); /* new ModularChannel(firstChannel) :334-336 */
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                int index = B(&chan, y, x);
                int isDelta = index < nb_deltas;
                int value;
                if (index >= 0 && index < nb_colors) {
                    value = palette[(size_t)c * nb_colors + index];
                } else if (index >= nb_colors) {
                    index -= nb_colors;
                    if (index < 64) {
                        /* Java masks int shift counts to 5 bits */
                        value = ((index >> ((2 * c) & 31)) % 4) * ((1 << (bit_depth & 31)) - 1) / 4
                            + (1 << ((bit_depth - 3 > 0 ? bit_depth - 3 : 0) & 31));
                    } else {
                        index -= 64;
                        for (int k = 0; k < c; k++) index /= 5;
                        value = (index % 5) * ((1 << (bit_depth & 31)) - 1) / 4;
                    }
                } else if (c < 3) {
                    index = (-index - 1) % 143;
                    value = kDeltaPalette[(index + 1) >> 1][c];
                    if ((index & 1) == 0) value = -value;
                    if (bit_depth > 8) value <<= (bit_depth < 24 ? bit_depth : 24) - 8;
                } else {
                    value = 0;
                }
                B(&chan, y, x) = value;
                if (isDelta) B(&chan, y, x) += prediction(&chan, y, x, d_pred);
            }
        }
    }
    return 0;
}
