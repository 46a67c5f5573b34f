// qm_tables.cuh -- host-side build of the dequantisation (QM) weight tables, once per frame, exactly where the
// reference builds them (HFGlobal constructor).  Replaces HFGlobal.getDefaultParams / generateWeights
// (J/frame/vardct/HFGlobal.java:79-188, 304-432): same float/double mix (Math.pow and Math.sqrt in double, everything
// else float), tables stored as reciprocals except MODE_RAW.
#pragma once
#include <math.h>
#include <string.h>
#include "common.cuh"

namespace qm {

enum { kLibrary = 0, kHornuss = 1, kDct2 = 2, kDct4 = 3, kDct4x8 = 4, kAfv = 5, kDct = 6, kRaw = 7 };

struct Row3 { float v[3][17]; int n; };

static inline void put(float dst[3][17], int32_t *n, const float *r, int len) {
    *n = len;
    for (int c = 0; c < 3; c++) memcpy(dst[c], r + c * len, sizeof(float) * len);
}
static inline void put9(float dst[3][9], int32_t *n, const float *r, int len) {
    *n = len;
    for (int c = 0; c < 3; c++) memcpy(dst[c], r + c * len, sizeof(float) * len);
}

// The 17 default parameter sets (HFGlobal.java:79-188).  Rows are X, Y, B.
static void default_params(jxlb200_qm_params p[17]) {
    memset(p, 0, sizeof(jxlb200_qm_params) * 17);
    for (int i = 0; i < 17; i++) p[i].denominator = 1.0f;
    static const float dct8[18] = {3150.0f, 0.0f, -0.4f, -0.4f, -0.4f, -2.0f, 560.0f, 0.0f, -0.3f, -0.3f, -0.3f, -0.3f,
                                   512.0f, -2.0f, -1.0f, 0.0f, -1.0f, -2.0f};
    put(p[0].dct_param, &p[0].n_dct, dct8, 6); p[0].mode = kDct;
    static const float hornuss[9] = {280.0f, 3160.0f, 3160.0f, 60.0f, 864.0f, 864.0f, 18.0f, 200.0f, 200.0f};
    put9(p[1].param, &p[1].n_param, hornuss, 3); p[1].mode = kHornuss;
    static const float dct2[18] = {3840.0f, 2560.0f, 1280.0f, 640.0f, 480.0f, 300.0f, 960.0f, 640.0f, 320.0f, 180.0f, 140.0f, 120.0f,
                                   640.0f, 320.0f, 128.0f, 64.0f, 32.0f, 16.0f};
    put9(p[2].param, &p[2].n_param, dct2, 6); p[2].mode = kDct2;
    static const float d44[12] = {2200.0f, 0.0f, 0.0f, 0.0f, 392.0f, 0.0f, 0.0f, 0.0f, 112.0f, -0.25f, -0.25f, -0.5f};
    static const float ones2[6] = {1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f};
    put(p[3].dct_param, &p[3].n_dct, d44, 4); put9(p[3].param, &p[3].n_param, ones2, 2);
    put(p[3].params4x4, &p[3].n_4x4, d44, 4); p[3].mode = kDct4;
    static const float dct16[21] = {
        8996.8725711814115328f, -1.3000777393353804f, -0.49424529824571225f, -0.439093774457103443f, -0.6350101832695744f, -0.90177264050827612f, -1.6162099239887414f,
        3191.48366296844234752f, -0.67424582104194355f, -0.80745813428471001f, -0.44925837484843441f, -0.35865440981033403f, -0.31322389111877305f, -0.37615025315725483f,
        1157.50408145487200256f, -2.0531423165804414f, -1.4f, -0.50687130033378396f, -0.42708730624733904f, -1.4856834539296244f, -4.9209142884401604f};
    put(p[4].dct_param, &p[4].n_dct, dct16, 7); p[4].mode = kDct;
    static const float dct32[24] = {
        15718.40830982518931456f, -1.025f, -0.98f, -0.9012f, -0.4f, -0.48819395464f, -0.421064f, -0.27f,
        7305.7636810695983104f, -0.8041958212306401f, -0.7633036457487539f, -0.55660379990111464f, -0.49785304658857626f, -0.43699592683512467f, -0.40180866526242109f, -0.27321683125358037f,
        3803.53173721215041536f, -3.060733579805728f, -2.0413270132490346f, -2.0235650159727417f, -0.5495389509954993f, -0.4f, -0.4f, -0.3f};
    put(p[5].dct_param, &p[5].n_dct, dct32, 8); p[5].mode = kDct;
    static const float dct8x16[21] = {
        7240.7734393502f, -0.7f, -0.7f, -0.2f, -0.2f, -0.2f, -0.5f,
        1448.15468787004f, -0.5f, -0.5f, -0.5f, -0.2f, -0.2f, -0.2f,
        506.854140754517f, -1.4f, -0.2f, -0.5f, -0.5f, -1.5f, -3.6f};
    put(p[6].dct_param, &p[6].n_dct, dct8x16, 7); p[6].mode = kDct;
    static const float dct8x32[24] = {
        16283.2494710648897f, -1.7812845336559429f, -1.6309059012653515f, -1.0382179034313539f, -0.85f, -0.7f, -0.9f, -1.2360638576849587f,
        5089.15750884921511936f, -0.320049391452786891f, -0.35362849922161446f, -0.30340000000000003f, -0.61f, -0.5f, -0.5f, -0.6f,
        3397.77603275308720128f, -0.321327362693153371f, -0.34507619223117997f, -0.70340000000000003f, -0.9f, -1.0f, -1.0f, -1.1754605576265209f};
    put(p[7].dct_param, &p[7].n_dct, dct8x32, 8); p[7].mode = kDct;
    static const float dct16x32[24] = {
        13844.97076442300573f, -0.97113799999999995f, -0.658f, -0.42026f, -0.22712f, -0.2206f, -0.226f, -0.6f,
        4798.964084220744293f, -0.61125308982767057f, -0.83770786552491361f, -0.79014862079498627f, -0.2692727459704829f, -0.38272769465388551f, -0.22924222653091453f, -0.20719098826199578f,
        1807.236946760964614f, -1.2f, -1.2f, -0.7f, -0.7f, -0.7f, -0.4f, -0.5f};
    put(p[8].dct_param, &p[8].n_dct, dct16x32, 8); p[8].mode = kDct;
    static const float d48[12] = {
        2198.050556016380522f, -0.96269623020744692f, -0.76194253026666783f, -0.6551140670773547f,
        764.3655248643528689f, -0.92630200888366945f, -0.9675229603596517f, -0.27845290869168118f,
        527.107573587542228f, -1.4594385811273854f, -1.450082094097871593f, -1.5843722511996204f};
    static const float ones1[3] = {1.0f, 1.0f, 1.0f};
    put(p[9].dct_param, &p[9].n_dct, d48, 4); put9(p[9].param, &p[9].n_param, ones1, 1); p[9].mode = kDct4x8;
    static const float afv[27] = {
        3072.0f, 3072.0f, 256.0f, 256.0f, 256.0f, 414.0f, 0.0f, 0.0f, 0.0f,
        1024.0f, 1024.0f, 50.0f, 50.0f, 50.0f, 58.0f, 0.0f, 0.0f, 0.0f,
        384.0f, 384.0f, 12.0f, 12.0f, 12.0f, 22.0f, -0.25f, -0.25f, -0.25f};
    put(p[10].dct_param, &p[10].n_dct, d48, 4); put9(p[10].param, &p[10].n_param, afv, 9);
    put(p[10].params4x4, &p[10].n_4x4, d44, 4); p[10].mode = kAfv;
    static const float tail[3][7] = {
        {-1.025f, -0.78f, -0.65012f, -0.19041574084286472f, -0.20819395464f, -0.421064f, -0.32733845535848671f},
        {-0.3041958212306401f, -0.3633036457487539f, -0.35660379990111464f, -0.3443074455424403f, -0.33699592683512467f, -0.30180866526242109f, -0.27321683125358037f},
        {-1.2f, -1.2f, -0.8f, -0.7f, -0.7f, -0.4f, -0.5f}};
    static const float head[6][3] = {
        {23966.1665298448605f, 8380.19148390090414f, 4493.02378009847706f},
        {15358.89804933239925f, 5597.360516150652990f, 2919.961618960011210f},
        {47932.3330596897210f, 16760.38296780180828f, 8986.04756019695412f},
        {30717.796098664792f, 11194.72103230130598f, 5839.92323792002242f},
        {95864.6661193794420f, 33520.76593560361656f, 17972.09512039390824f},
        {61435.5921973295970f, 24209.44206460261196f, 12979.84647584004484f}};
    for (int i = 0; i < 6; i++) {
        jxlb200_qm_params &q = p[11 + i];
        for (int c = 0; c < 3; c++) {
            q.dct_param[c][0] = head[i][c];
            memcpy(&q.dct_param[c][1], tail[c], sizeof(float) * 7);
        }
        q.n_dct = 8;
        q.mode = kDct;
    }
}

// HFGlobal.quantMult :55-57, interpolate :42-53
static inline float mult(float v) { return v >= 0 ? 1.0f + v : 1.0f / (1.0f - v); }
static inline float interp(float pos, const float *bands, int nb) {
    const int last = nb - 1;
    if (last == 0) return bands[0];
    const int i = (int)pos;
    const float frac = pos - i;
    if (i + 1 > last) return bands[last];
    const float a = bands[i], b = bands[i + 1];
    return a * (float)pow((double)(b / a), (double)frac);   // Math.pow is double
}
// HFGlobal.getDCTQuantWeights :59-77
static void dct_weights(int h, int w, const float *prm, int n, float *out) {
    float bands[17];
    bands[0] = prm[0];
    for (int i = 1; i < n; i++) bands[i] = bands[i - 1] * mult(prm[i]);
    const float scale = (n - 1) / ((float)sqrt(2.0) + 1e-6f);
    for (int y = 0; y < h; y++) {
        const float dy = (float)y * scale / (h - 1);
        const float dy2 = dy * dy;
        for (int x = 0; x < w; x++) {
            const float dx = (float)x * scale / (w - 1);
            out[y * w + x] = interp((float)sqrt((double)(dx * dx + dy2)), bands, n);   // Math.sqrt is double
        }
    }
}

static const float kAfvFreqs[16] = {0, 0, 0.8517778890324296f, 5.37778436506804f, 0, 0, 4.734747904497923f, 5.449245381693219f,
                                    1.6598270267479331f, 4, 7.275749096817861f, 10.423227632456525f, 2.662932286148962f,
                                    7.630657783650829f, 8.962388608184032f, 12.97166202570235f};

// all 17 tables; returns 0 or JXLB200_E_STREAM ("Negative or infinite weight", "Illegal negative band value")
static int generate(const jxlb200_qm_params prm[17], float *weights, int32_t offsets[51]) {
    int off = 0;
    for (int idx = 0; idx < 17; idx++) {
        int mh = 0, mw = 0;   // TransformType.getByParameterIndex: first non-vertical type with this parameterIndex
        for (int t = 0; t < 27; t++)
            if (h_tt[t].param == idx && h_tt[t].bh <= h_tt[t].bw) {
                mh = h_tt[t].bh * 8; mw = h_tt[t].bw * 8;
                break;
            }
        const jxlb200_qm_params &q = prm[idx];
        for (int c = 0; c < 3; c++) {
            float *wt = weights + off;
            offsets[idx * 3 + c] = off;
            off += mh * mw;
            switch (q.mode) {
            case kDct:
                dct_weights(mh, mw, q.dct_param[c], q.n_dct, wt);
                break;
            case kDct4: {
                float w4[16];
                dct_weights(4, 4, q.dct_param[c], q.n_dct, w4);
                for (int y = 0; y < 8; y++)
                    for (int x = 0; x < 8; x++) wt[y * 8 + x] = w4[(y / 2) * 4 + x / 2];
                wt[8] /= q.param[c][0];
                wt[1] /= q.param[c][0];
                wt[9] /= q.param[c][1];
                break;
            }
            case kDct2: {
                const float *pr = q.param[c];
                for (int y = 0; y < 8; y++)
                    for (int x = 0; x < 8; x++) {
                        const int m = y > x ? y : x;   // band by the larger coordinate: {0}, {1}, {2,3}, {4..7}
                        float v;
                        if (m == 0) v = 1.0f;
                        else if (m == 1) v = (y == 1 && x == 1) ? pr[1] : pr[0];
                        else if (m < 4) v = (y >= 2 && x >= 2) ? pr[3] : pr[2];
                        else v = (y >= 4 && x >= 4) ? pr[5] : pr[4];
                        wt[y * 8 + x] = v;
                    }
                break;
            }
            case kHornuss:
                for (int i = 0; i < 64; i++) wt[i] = q.param[c][0];
                wt[9] = q.param[c][2];
                wt[1] = wt[8] = q.param[c][1];
                wt[0] = 1.0f;
                break;
            case kDct4x8: {
                float w48[32];
                dct_weights(4, 8, q.dct_param[c], q.n_dct, w48);
                for (int y = 0; y < 8; y++)
                    for (int x = 0; x < 8; x++) wt[y * 8 + x] = w48[(y / 2) * 8 + x];
                wt[8] /= q.param[c][0];
                break;
            }
            case kAfv: {   // getAFVTransformWeights :304-345
                float w48[32], w44[16], bands[4];
                dct_weights(4, 8, q.dct_param[c], q.n_dct, w48);
                dct_weights(4, 4, q.params4x4[c], q.n_4x4, w44);
                const float low = 0.8517778890324296f, high = 12.97166202570235f;
                bands[0] = q.param[c][5];
                if (bands[0] < 0) return JXLB200_E_STREAM;
                for (int i = 1; i < 4; i++) {
                    bands[i] = bands[i - 1] * mult(q.param[c][i + 5]);
                    if (bands[i] < 0) return JXLB200_E_STREAM;
                }
                memset(wt, 0, sizeof(float) * 64);
                wt[0] = 1.0f;
                wt[8] = q.param[c][0]; wt[1] = q.param[c][1];
                wt[16] = q.param[c][2]; wt[2] = q.param[c][3];
                wt[18] = q.param[c][4];
                for (int y = 0; y < 4; y++) {
                    for (int x = 0; x < 4; x++) {
                        if (x < 2 && y < 2) continue;
                        wt[(2 * x) * 8 + 2 * y] = interp((kAfvFreqs[y * 4 + x] - low) / (high - low), bands, 4);
                    }
                    for (int x = 0; x < 8; x++)
                        if (x || y) wt[(2 * y + 1) * 8 + x] = w48[y * 8 + x];
                    for (int x = 0; x < 4; x++)
                        if (x || y) wt[(2 * y) * 8 + 2 * x + 1] = w44[y * 4 + x];
                }
                break;
            }
            case kRaw:
                if (!q.raw[c]) return JXLB200_E_ARG;
                for (int i = 0; i < mh * mw; i++) wt[i] = q.raw[c][i] * q.denominator;
                break;
            default:
                return JXLB200_E_ARG;
            }
        }
        if (q.mode != kRaw)
            for (int c = 0; c < 3; c++) {
                float *wt = weights + offsets[idx * 3 + c];
                for (int i = 0; i < mh * mw; i++) {
                    if (!(wt[i] > 0.0f) || !isfinite(wt[i])) return JXLB200_E_STREAM;
                    wt[i] = 1.0f / wt[i];
                }
            }
    }
    return 0;
}

}  // namespace qm
