// Host build of jxlatte_b200/csrc/transforms.cuh so the transform arithmetic can be checked against the oracle on
// a CPU-only box (tests/test_transforms_host.py).  Mirrors how the kernels drive the same templates.
// Built with -ffp-contract=off: the reference-order transforms must not be contracted into FMAs.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "../../jxlatte_b200/csrc/transforms.cuh"

static float g_cos[1302];
static void init_cos() {
    static bool done = false;
    if (done) return;
    int o = 0;
    const double root2 = sqrt(2.0);
    for (int l = 1; l <= 5; l++) {
        const int s = 1 << l;
        for (int n = 0; n < s - 1; n++)
            for (int k = 0; k < s; k++) g_cos[o++] = (float)(root2 * cos(M_PI * (n + 1) * (k + 0.5) / s));
    }
    done = true;
}
static float lut_host(int i) { return g_cos[i]; }

extern "C" int jxlb_test_idct1d(const float *in, float *out, int n) {
    init_cos();
    float v[32];
    switch (n) {
    case 1: out[0] = in[0]; return 0;
    case 2: memcpy(v, in, 8); RefIDCT<2>::run(v, lut_host); memcpy(out, v, 8); return 0;
    case 4: memcpy(v, in, 16); RefIDCT<4>::run(v, lut_host); memcpy(out, v, 16); return 0;
    case 8: memcpy(v, in, 32); RefIDCT<8>::run(v, lut_host); memcpy(out, v, 32); return 0;
    case 16: memcpy(v, in, 64); RefIDCT<16>::run(v, lut_host); memcpy(out, v, 64); return 0;
    case 32: memcpy(v, in, 128); RefIDCT<32>::run(v, lut_host); memcpy(out, v, 128); return 0;
    }
    return -1;
}

extern "C" int jxlb_test_block8(int type, const float *in64, float *out64, const float *afv_basis) {
    init_cos();
    float v[64];
    memcpy(v, in64, sizeof(v));
    auto out = [&](int y, int x, float val) { out64[y * 8 + x] = val; };
    auto basis = [&](int j, int i) { return afv_basis[j * 16 + i]; };
    switch (type) {
    case 0: inv_dct8x8(v, lut_host, out); return 0;
    case 1: inv_hornuss(v, out); return 0;
    case 2: inv_dct2(v, out); return 0;
    case 3: inv_dct4(v, lut_host, out); return 0;
    case 12: inv_dct4x8<false>(v, lut_host, out); return 0;   // TransformType.DCT4_8 -> METHOD_DCT4_8
    case 13: inv_dct4x8<true>(v, lut_host, out); return 0;    // TransformType.DCT8_4 -> METHOD_DCT8_4
    case 14: case 15: case 16: case 17:
        inv_afv(v, (type == 16 || type == 17) ? 1 : 0, (type == 15 || type == 17) ? 1 : 0, basis, lut_host, out); return 0;
    }
    return -1;
}
