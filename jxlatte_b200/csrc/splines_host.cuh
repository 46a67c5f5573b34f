// splines_host.cuh -- host half of spline rendering: from control points and quantised DCT-32 coefficient
// tracks to the list of unit-spaced arcs the pixel kernel (k8_splines) consumes.  Sequential and tiny (one entry per pixel of
// curve length), so it runs on the host in double/float exactly as the reference does:
//   Spline.upsampleControlPoints  J/frame/features/spline/Spline.java:26-88   (centripetal Catmull-Rom, 16 steps per segment)
//   Spline.computeIntermediarySamples :89-123                                  (walk at render distance 1)
//   Spline.computeCoeffs :131-150, fourierICT :124-130, renderSpline :151-170  (per-arc colour, sigma, bounding box)
// Reference quirks kept on purpose: every spline is drawn with spline 0's coefficients (the constructor drops splineID),
// and MathHelper.max(0.01f, ...) returns the minimum.
#pragma once
#include <cmath>
#include <vector>

#include "k8_features.cuh"

namespace splines_host {

struct Track { float c[32]; };

inline int f2i(float v) {        // Java (int) cast: NaN -> 0, saturating
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return -2147483647 - 1;
    return (int)v;
}

inline float ict32(const Track &tr, float t) {
    float total = (float)std::sqrt(0.5) * tr.c[0];
    for (int i = 1; i < 32; i++) total += tr.c[i] * (float)std::cos(i * (M_PI / 32.0) * (t + 0.5));
    return total;
}

// points: (x, y) pairs.  Appends this spline's arcs.
inline void build(const int32_t *points, int n, const Track (&trk)[4], int h, int w, std::vector<SplineArcDev> &out) {
    std::vector<float> uy, ux;
    if (n == 1) {
        uy.push_back((float)points[1]);
        ux.push_back((float)points[0]);
    } else {
        std::vector<int> ey(n + 2), ex(n + 2);
        for (int i = 0; i < n; i++) { ex[i + 1] = points[2 * i]; ey[i + 1] = points[2 * i + 1]; }
        ex[0] = ex[1] * 2 - ex[2];             ey[0] = ey[1] * 2 - ey[2];
        ex[n + 1] = ex[n] * 2 - ex[n - 1];     ey[n + 1] = ey[n] * 2 - ey[n - 1];
        const int segments = n - 1;
        uy.resize(16 * segments + 1);
        ux.resize(16 * segments + 1);
        for (int s = 0; s < segments; s++) {
            float py[4], px[4], dy[3], dx[3], t[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; k++) { py[k] = (float)ey[s + k]; px[k] = (float)ex[s + k]; }
            uy[16 * s] = py[1];
            ux[16 * s] = px[1];
            for (int k = 0; k < 3; k++) {
                dy[k] = py[k + 1] - py[k];
                dx[k] = px[k + 1] - px[k];
                t[k + 1] = t[k] + (float)std::pow((double)(dy[k] * dy[k] + dx[k] * dx[k]), 0.25);
            }
            for (int step = 1; step < 16; step++) {
                const float knot = t[1] + 0.0625f * step * (t[2] - t[1]);
                float ay[3], ax[3], by[2], bx[2];
                for (int k = 0; k < 3; k++) {
                    const float f = (knot - t[k]) / (t[k + 1] - t[k]);
                    ay[k] = dy[k] * f + py[k];
                    ax[k] = dx[k] * f + px[k];
                }
                for (int k = 0; k < 2; k++) {
                    const float f = (knot - t[k]) / (t[k + 2] - t[k]);
                    by[k] = (ay[k + 1] - ay[k]) * f + ay[k];
                    bx[k] = (ax[k + 1] - ax[k]) * f + ax[k];
                }
                const float f = (knot - t[1]) / (t[2] - t[1]);
                uy[16 * s + step] = (by[1] - by[0]) * f + by[0];
                ux[16 * s + step] = (bx[1] - bx[0]) * f + bx[0];
            }
        }
        uy.back() = (float)points[2 * (n - 1) + 1];
        ux.back() = (float)points[2 * (n - 1)];
    }
    // unit-distance samples along the polyline
    struct Arc { float y, x, len; };
    std::vector<Arc> arcs;
    float cy = uy[0], cx = ux[0];
    arcs.push_back({cy, cx, 1.0f});
    size_t next = 0;
    while (next < uy.size()) {
        float py = cy, px = cx, walked = 0.0f;
        for (;;) {
            if (next >= uy.size()) { arcs.push_back({py, px, walked}); break; }
            const float dy = uy[next] - py, dx = ux[next] - px;
            const float step = (float)std::sqrt((double)(dy * dy + dx * dx));
            if (walked + step >= 1.0f) {
                const float f = (1.0f - walked) / step;
                cy = dy * f + py;
                cx = dx * f + px;
                arcs.push_back({cy, cx, 1.0f});
                break;
            }
            walked += step;
            py = uy[next];
            px = ux[next];
            next++;
        }
    }
    const float total = ((float)arcs.size() - 2.0f) * 1.0f + arcs.back().len;
    if (total <= 0.0) return;
    for (size_t i = 0; i < arcs.size(); i++) {
        const float progress = std::fmin(1.0f, (float)i * 1.0f / total);
        const float t = 31.0f * progress;
        SplineArcDev a;
        a.y = arcs[i].y;
        a.x = arcs[i].x;
        for (int c = 0; c < 3; c++) a.value[c] = ict32(trk[c], t) * arcs[i].len;
        a.sigma = ict32(trk[3], t);
        a.inv_sigma = 1.0f / a.sigma;
        float m = 0.01f;
        for (int c = 0; c < 3; c++) m = a.value[c] < m ? a.value[c] : m;
        const float reach = (float)std::sqrt((double)(-2.0f * a.sigma * a.sigma * ((float)std::log(0.1) * 3.0f - m)));
        a.x0 = std::max(0, f2i(a.x - reach + 0.5f));
        a.x1 = std::min(w - 1, f2i(a.x + reach + 0.5f));
        a.y0 = std::max(0, f2i(a.y - reach + 0.5f));
        a.y1 = std::min(h - 1, f2i(a.y + reach + 0.5f));
        out.push_back(a);
    }
}

}  // namespace splines_host
