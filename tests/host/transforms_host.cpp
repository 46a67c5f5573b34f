// Host build of jxlatte_b200/csrc/transforms.cuh so the transform arithmetic can be checked against the oracle on
// a CPU-only box (tests/test_transforms_host.py).  Mirrors how the kernels drive the same templates.
#include <stdint.h>
#include <string.h>
#include "../../jxlatte_b200/csrc/transforms.cuh"

static float sec_host(int n, int k) { return n == 64 ? h_sec64[k] : n == 128 ? h_sec128[k] : h_sec256[k]; }

template <int L> static void long_line(const float *in, float *out) {
    constexpr int R = 1 << L;
    float T[R][32];
    for (int p = 0; p < R; p++) {
        float v[32];
        for (int m = 0; m < 32; m++) v[m] = LeeGather<L>::get([&](int i) { return in[i]; }, p, m);
        LeeIDCT<32>::run(v);
        memcpy(T[p], v, sizeof(v));
    }
    for (int k0 = 0; k0 < 32; k0++) {
        float val[R];
        int idx[R];
        for (int p = 0; p < R; p++) val[p] = T[p][k0];
        lee_combine<L>(val, idx, k0, sec_host);
        for (int s = 0; s < R; s++) out[idx[s]] = val[s];
    }
}

extern "C" int jxlb_test_idct1d(const float *in, float *out, int n) {
    float v[32];
    switch (n) {
    case 1: out[0] = in[0]; return 0;
    case 2: memcpy(v, in, 8); LeeIDCT<2>::run(v); memcpy(out, v, 8); return 0;
    case 4: memcpy(v, in, 16); LeeIDCT<4>::run(v); memcpy(out, v, 16); return 0;
    case 8: memcpy(v, in, 32); LeeIDCT<8>::run(v); memcpy(out, v, 32); return 0;
    case 16: memcpy(v, in, 64); LeeIDCT<16>::run(v); memcpy(out, v, 64); return 0;
    case 32: memcpy(v, in, 128); LeeIDCT<32>::run(v); memcpy(out, v, 128); return 0;
    case 64: long_line<1>(in, out); return 0;
    case 128: long_line<2>(in, out); return 0;
    case 256: long_line<3>(in, out); return 0;
    }
    return -1;
}

extern "C" int jxlb_test_block8(int type, const float *in64, float *out64, const float *afv_basis) {
    float v[64];
    memcpy(v, in64, sizeof(v));
    auto out = [&](int y, int x, float val) { out64[y * 8 + x] = val; };
    auto basis = [&](int j, int i) { return afv_basis[j * 16 + i]; };
    switch (type) {
    case 0: inv_dct8x8(v, out); return 0;
    case 1: inv_hornuss(v, out); return 0;
    case 2: inv_dct2(v, out); return 0;
    case 3: inv_dct4(v, out); return 0;
    case 12: inv_dct4x8<false>(v, out); return 0;   // TransformType.DCT4_8 -> METHOD_DCT4_8
    case 13: inv_dct4x8<true>(v, out); return 0;    // TransformType.DCT8_4 -> METHOD_DCT8_4
    case 14: case 15: case 16: case 17:
        inv_afv(v, (type == 16 || type == 17) ? 1 : 0, (type == 15 || type == 17) ? 1 : 0, basis, out); return 0;
    }
    return -1;
}
