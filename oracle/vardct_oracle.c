/*
 * oracle/vardct_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see jxl_oracle.h).
 *
 * Literal restatement of jxlatte's VarDCT reconstruction: same temporaries, same loop order, float32,
 * no FMA contraction.  "J/" = /root/reference/java/com/traneptora/jxlatte/.  PARITY UNPINNED (no JVM here).
 *
 * The only liberty taken: Java's per-group int[3][256][256] / float[3][256][256] arrays are addressed inside
 * frame-level planes (group-local index + group origin), and independent groups / rows may run on several
 * OpenMP threads for the CPU baseline.  Neither changes any arithmetic.
 */
#define _GNU_SOURCE
#include "jxl_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * TransformType  (J/frame/vardct/TransformType.java:10-36, 129-131, 145-166)
 * ---------------------------------------------------------------------------------------------- */
enum { METHOD_DCT = 0, METHOD_DCT2 = 1, METHOD_DCT4 = 2, METHOD_HORNUSS = 3, METHOD_DCT8_4 = 4,
       METHOD_DCT4_8 = 5, METHOD_AFV = 6 };
enum { MODE_LIBRARY = 0, MODE_HORNUSS = 1, MODE_DCT2 = 2, MODE_DCT4 = 3, MODE_DCT4_8 = 4, MODE_AFV = 5,
       MODE_DCT = 6, MODE_RAW = 7 };

typedef struct { int type, parameterIndex, orderID, method, pixelH, pixelW; } tt_t;
/* name, type, parameterIndex, orderID, transformMethod, pixelHeight, pixelWidth */
static const tt_t TT[27] = {
    {0, 0, 0, 0, 8, 8},       /* DCT8 */
    {1, 1, 1, 3, 8, 8},       /* HORNUSS */
    {2, 2, 1, 1, 8, 8},       /* DCT2 */
    {3, 3, 1, 2, 8, 8},       /* DCT4 */
    {4, 4, 2, 0, 16, 16},     /* DCT16 */
    {5, 5, 3, 0, 32, 32},     /* DCT32 */
    {6, 6, 4, 0, 16, 8},      /* DCT16_8 */
    {7, 6, 4, 0, 8, 16},      /* DCT8_16 */
    {8, 7, 5, 0, 32, 8},      /* DCT32_8 */
    {9, 7, 5, 0, 8, 32},      /* DCT8_32 */
    {10, 8, 6, 0, 32, 16},    /* DCT32_16 */
    {11, 8, 6, 0, 16, 32},    /* DCT16_32 */
    {12, 9, 1, 5, 8, 8},      /* DCT4_8 */
    {13, 9, 1, 4, 8, 8},      /* DCT8_4 */
    {14, 10, 1, 6, 8, 8},     /* AFV0 */
    {15, 10, 1, 6, 8, 8},     /* AFV1 */
    {16, 10, 1, 6, 8, 8},     /* AFV2 */
    {17, 10, 1, 6, 8, 8},     /* AFV3 */
    {18, 11, 7, 0, 64, 64},   /* DCT64 */
    {19, 12, 8, 0, 64, 32},   /* DCT64_32 */
    {20, 12, 8, 0, 32, 64},   /* DCT32_64 */
    {21, 13, 9, 0, 128, 128}, /* DCT128 */
    {22, 14, 10, 0, 128, 64}, /* DCT128_64 */
    {23, 14, 10, 0, 64, 128}, /* DCT64_128 */
    {24, 15, 11, 0, 256, 256},/* DCT256 */
    {25, 16, 12, 0, 256, 128},/* DCT256_128 */
    {26, 16, 12, 0, 128, 256},/* DCT128_256 */
};

static int tt_flip(const tt_t *t) { /* TransformType.flip :129-131 */
    return t->pixelH > t->pixelW || (t->method == METHOD_DCT && t->pixelH == t->pixelW);
}
static int tt_matrix_h(const tt_t *t) { return t->pixelH < t->pixelW ? t->pixelH : t->pixelW; }
static int tt_matrix_w(const tt_t *t) { return t->pixelH > t->pixelW ? t->pixelH : t->pixelW; }
static int tt_is_vertical(const tt_t *t) { return t->pixelH > t->pixelW; }

/* TransformType.getByParameterIndex :69-72,100-102 -- first non-vertical type with that parameterIndex */
static const tt_t *tt_by_param(int pi) {
    for (int i = 0; i < 27; i++)
        if (TT[i].parameterIndex == pi && !tt_is_vertical(&TT[i])) return &TT[i];
    return NULL;
}

int32_t orc_tt_info(int32_t type, int32_t *param_index, int32_t *method, int32_t *pixel_h, int32_t *pixel_w, int32_t *flip) {
    if (type < 0 || type > 26) return -2;
    const tt_t *t = &TT[type];
    if (param_index) *param_index = t->parameterIndex;
    if (method) *method = t->method;
    if (pixel_h) *pixel_h = t->pixelH;
    if (pixel_w) *pixel_w = t->pixelW;
    if (flip) *flip = tt_flip(t);
    return 0;
}

/* MathHelper.ceilLog2 (J/util/MathHelper.java:156-162) */
static int ceil_log2(long x) {
    long v = x - 1;
    int n = 0;
    while (v > 0) { n++; v >>= 1; }
    return n;
}

/* LLFScale (J/frame/vardct/LLFScale.java:7-23) */
static const float SCALE_F[32] = {
    1.0000000000000000000f, 1.0003954307206444720f, 1.0015830492063566798f,
    1.0035668445359847378f, 1.0063534990068075448f, 1.0099524393750471170f,
    1.0143759095929498827f, 1.0196390660646908181f, 1.0257600967811994622f,
    1.0327603660498609462f, 1.0406645869479269795f, 1.0495010240726261235f,
    1.0593017296818027804f, 1.0701028169146909598f, 1.0819447744633102634f,
    1.0948728278735071820f, 1.1089373535928257701f, 1.1241943530045446156f,
    1.1407059950032801390f, 1.1585412372562662921f, 1.1777765381971696030f,
    1.1984966740821024139f, 1.2207956782314713353f, 1.2447779229495839992f,
    1.2705593687655135089f, 1.2982690107340108228f, 1.3280505578212198723f,
    1.3600643892400108061f, 1.3944898413648201160f, 1.4315278911623840964f,
    1.4714043176060183528f, 1.5143734423313919909f,
};
static float scale_f(int x, int xll) { return SCALE_F[x << (5 - xll)]; }

/* TransformType ctor :158-165 */
float orc_llf_scale(int32_t type, int32_t y, int32_t x) {
    const tt_t *t = &TT[type];
    int yll = ceil_log2(t->pixelH >> 3), xll = ceil_log2(t->pixelW >> 3);
    return scale_f(y, yll) * scale_f(x, xll);
}

/* ------------------------------------------------------------------------------------------------
 * MathHelper  (J/util/MathHelper.java:17-30, 68-145, 323-329)
 * ---------------------------------------------------------------------------------------------- */
static float *cosine_lut[9]; /* [l][n*(1<<l) + k] */
static void init_cosine_lut(void) {
    static volatile int done = 0;
    if (done) return;
#pragma omp critical(orc_lut)
    {
        if (!done) {
            const double root2 = sqrt(2.0);
            for (int l = 0; l < 9; l++) {
                int s = 1 << l;
                float *t = (float *)malloc(sizeof(float) * (size_t)(s > 1 ? (s - 1) * s : 1));
                for (int n = 0; n < s - 1; n++)
                    for (int k = 0; k < s; k++)
                        t[n * s + k] = (float)(root2 * cos(M_PI * (n + 1) * (k + 0.5) / s));
                cosine_lut[l] = t;
            }
            done = 1;
        }
    }
}

/* MathHelper.inverseDCTHorizontal :68-78 */
static void inverse_dct_horizontal(const float *src, float *dest, int xLogLength, int xLength) {
    for (int k = 0; k < xLength; k++) dest[k] = src[0];
    const float *lutX = cosine_lut[xLogLength];
    for (int n = 1; n < xLength; n++) {
        const float *lut = lutX + (size_t)(n - 1) * xLength;
        const float s2 = src[n];
        for (int k = 0; k < xLength; k++)
            dest[k] += s2 * lut[k];
    }
}

/* MathHelper.forwardDCTHorizontal :80-94 */
static void forward_dct_horizontal(const float *src, float *dest, int xLogLength, int xLength) {
    const float invLength = 1.0f / xLength;
    float d2 = src[0];
    for (int x = 1; x < xLength; ++x) d2 += src[x];
    dest[0] = d2 * invLength;
    for (int k = 1; k < xLength; ++k) {
        const float *lut = cosine_lut[xLogLength] + (size_t)(k - 1) * xLength;
        d2 = src[0] * lut[0];
        for (int n = 1; n < xLength; ++n) d2 += src[n] * lut[n];
        dest[k] = d2 * invLength;
    }
}

/* MathHelper.transposeMatrixInto :138-145 */
static void transpose_into(const float *src, int sp, float *dest, int dp, int srcHeight, int srcWidth) {
    for (int y = 0; y < srcHeight; y++)
        for (int x = 0; x < srcWidth; x++)
            dest[x * dp + y] = src[y * sp + x];
}

#define SCR 256 /* scratch pitch: Java scratch blocks are float[256][256] */

/* MathHelper.inverseDCT2D :96-122.  src/dest are (pointer to start element, pitch). */
static void inverse_dct_2d(const float *src, int sp, float *dest, int dp, int height, int width,
                           float *scratch0, float *scratch1, int transposed) {
    int logHeight = ceil_log2(height);
    int logWidth = ceil_log2(width);
    if (transposed) {
        for (int y = 0; y < height; y++)
            inverse_dct_horizontal(src + (size_t)y * sp, scratch1 + (size_t)y * SCR, logWidth, width);
        transpose_into(scratch1, SCR, scratch0, SCR, height, width);
        for (int y = 0; y < width; y++)
            inverse_dct_horizontal(scratch0 + (size_t)y * SCR, dest + (size_t)y * dp, logHeight, height);
    } else {
        transpose_into(src, sp, scratch0, SCR, height, width);
        for (int y = 0; y < width; y++)
            inverse_dct_horizontal(scratch0 + (size_t)y * SCR, scratch1 + (size_t)y * SCR, logHeight, height);
        transpose_into(scratch1, SCR, scratch0, SCR, width, height);
        for (int y = 0; y < height; y++)
            inverse_dct_horizontal(scratch0 + (size_t)y * SCR, dest + (size_t)y * dp, logWidth, width);
    }
}

/* MathHelper.forwardDCT2D :124-136 (scratch pitch 32: Java uses float[2][32][32], HFCoefficients.java:195) */
static void forward_dct_2d(const float *src, int sp, float *dest, int dp, int height, int width,
                           float *scratch0, float *scratch1, int scrp) {
    const int yLogLength = ceil_log2(height);
    const int xLogLength = ceil_log2(width);
    for (int y = 0; y < height; y++)
        forward_dct_horizontal(src + (size_t)y * sp, scratch0 + (size_t)y * scrp, xLogLength, width);
    transpose_into(scratch0, scrp, scratch1, scrp, height, width);
    for (int x = 0; x < width; x++)
        forward_dct_horizontal(scratch1 + (size_t)x * scrp, scratch0 + (size_t)x * scrp, yLogLength, height);
    transpose_into(scratch0, scrp, dest, dp, width, height);
}

/* MathHelper.mirrorCoordinate :323-329 */
int32_t orc_mirror_coordinate(int32_t coordinate, int32_t size) {
    while (coordinate < 0 || coordinate >= size) {
        int tc = ~coordinate;
        coordinate = tc >= 0 ? tc : (size << 1) + tc;
    }
    return coordinate;
}

void orc_inverse_dct_1d(const float *src, float *dest, int32_t n) {
    init_cosine_lut();
    inverse_dct_horizontal(src, dest, ceil_log2(n), n);
}
void orc_forward_dct_1d(const float *src, float *dest, int32_t n) {
    init_cosine_lut();
    forward_dct_horizontal(src, dest, ceil_log2(n), n);
}
void orc_inverse_dct_2d(const float *src, float *dest, int32_t h, int32_t w, int32_t transposed) {
    init_cosine_lut();
    float *s0 = (float *)malloc(sizeof(float) * SCR * SCR), *s1 = (float *)malloc(sizeof(float) * SCR * SCR);
    inverse_dct_2d(src, w, dest, transposed ? h : w, h, w, s0, s1, transposed);
    free(s0); free(s1);
}
void orc_forward_dct_2d(const float *src, float *dest, int32_t h, int32_t w) {
    init_cosine_lut();
    float *s0 = (float *)malloc(sizeof(float) * SCR * SCR), *s1 = (float *)malloc(sizeof(float) * SCR * SCR);
    forward_dct_2d(src, w, dest, w, h, w, s0, s1, SCR);
    free(s0); free(s1);
}

/* ------------------------------------------------------------------------------------------------
 * HFGlobal: QM weight tables  (J/frame/vardct/HFGlobal.java:19-21, 42-77, 79-188, 304-432)
 * ---------------------------------------------------------------------------------------------- */
static const float afvFreqs[16] = {0, 0, 0.8517778890324296f, 5.37778436506804f,
    0, 0, 4.734747904497923f, 5.449245381693219f, 1.6598270267479331f, 4, 7.275749096817861f,
    10.423227632456525f, 2.662932286148962f, 7.630657783650829f, 8.962388608184032f, 12.97166202570235f};

static void set_rows(float dst[3][17], int *n, const float *r0, const float *r1, const float *r2, int len) {
    *n = len;
    memcpy(dst[0], r0, sizeof(float) * len);
    memcpy(dst[1], r1, sizeof(float) * len);
    memcpy(dst[2], r2, sizeof(float) * len);
}
static void set_param_rows(float dst[3][9], int *n, const float *r0, const float *r1, const float *r2, int len) {
    *n = len;
    memcpy(dst[0], r0, sizeof(float) * len);
    memcpy(dst[1], r1, sizeof(float) * len);
    memcpy(dst[2], r2, sizeof(float) * len);
}
static void prepend(float *dst, float a, const float *seq) { dst[0] = a; memcpy(dst + 1, seq, sizeof(float) * 7); }

/* HFGlobal.getDefaultParams :79-188 */
void orc_qm_default_params(orc_qm_params p[17]) {
    memset(p, 0, sizeof(orc_qm_params) * 17);
    for (int i = 0; i < 17; i++) p[i].denominator = 1.0f;
    {
        static const float a[] = {3150.0f, 0.0f, -0.4f, -0.4f, -0.4f, -2.0f};
        static const float b[] = {560.0f, 0.0f, -0.3f, -0.3f, -0.3f, -0.3f};
        static const float c[] = {512.0f, -2.0f, -1.0f, 0.0f, -1.0f, -2.0f};
        set_rows(p[0].dct_param, &p[0].n_dct, a, b, c, 6); p[0].mode = MODE_DCT;
    }
    {
        static const float a[] = {280.0f, 3160.0f, 3160.0f};
        static const float b[] = {60.0f, 864.0f, 864.0f};
        static const float c[] = {18.0f, 200.0f, 200.0f};
        set_param_rows(p[1].param, &p[1].n_param, a, b, c, 3); p[1].mode = MODE_HORNUSS;
    }
    {
        static const float a[] = {3840.0f, 2560.0f, 1280.0f, 640.0f, 480.0f, 300.0f};
        static const float b[] = {960.0f, 640.0f, 320.0f, 180.0f, 140.0f, 120.0f};
        static const float c[] = {640.0f, 320.0f, 128.0f, 64.0f, 32.0f, 16.0f};
        set_param_rows(p[2].param, &p[2].n_param, a, b, c, 6); p[2].mode = MODE_DCT2;
    }
    static const float d44a[] = {2200.0f, 0.0f, 0.0f, 0.0f};
    static const float d44b[] = {392.0f, 0.0f, 0.0f, 0.0f};
    static const float d44c[] = {112.0f, -0.25f, -0.25f, -0.5f};
    {
        static const float one2[] = {1.0f, 1.0f};
        set_rows(p[3].dct_param, &p[3].n_dct, d44a, d44b, d44c, 4);
        set_param_rows(p[3].param, &p[3].n_param, one2, one2, one2, 2);
        set_rows(p[3].params4x4, &p[3].n_4x4, d44a, d44b, d44c, 4);
        p[3].mode = MODE_DCT4;
    }
    {
        static const float a[] = {8996.8725711814115328f, -1.3000777393353804f, -0.49424529824571225f, -0.439093774457103443f,
            -0.6350101832695744f, -0.90177264050827612f, -1.6162099239887414f};
        static const float b[] = {3191.48366296844234752f, -0.67424582104194355f, -0.80745813428471001f, -0.44925837484843441f,
            -0.35865440981033403f, -0.31322389111877305f, -0.37615025315725483f};
        static const float c[] = {1157.50408145487200256f, -2.0531423165804414f, -1.4f, -0.50687130033378396f,
            -0.42708730624733904f, -1.4856834539296244f, -4.9209142884401604f};
        set_rows(p[4].dct_param, &p[4].n_dct, a, b, c, 7); p[4].mode = MODE_DCT;
    }
    {
        static const float a[] = {15718.40830982518931456f, -1.025f, -0.98f, -0.9012f, -0.4f, -0.48819395464f, -0.421064f, -0.27f};
        static const float b[] = {7305.7636810695983104f, -0.8041958212306401f, -0.7633036457487539f, -0.55660379990111464f,
            -0.49785304658857626f, -0.43699592683512467f, -0.40180866526242109f, -0.27321683125358037f};
        static const float c[] = {3803.53173721215041536f, -3.060733579805728f, -2.0413270132490346f, -2.0235650159727417f,
            -0.5495389509954993f, -0.4f, -0.4f, -0.3f};
        set_rows(p[5].dct_param, &p[5].n_dct, a, b, c, 8); p[5].mode = MODE_DCT;
    }
    {
        static const float a[] = {7240.7734393502f, -0.7f, -0.7f, -0.2f, -0.2f, -0.2f, -0.5f};
        static const float b[] = {1448.15468787004f, -0.5f, -0.5f, -0.5f, -0.2f, -0.2f, -0.2f};
        static const float c[] = {506.854140754517f, -1.4f, -0.2f, -0.5f, -0.5f, -1.5f, -3.6f};
        set_rows(p[6].dct_param, &p[6].n_dct, a, b, c, 7); p[6].mode = MODE_DCT;
    }
    {
        static const float a[] = {16283.2494710648897f, -1.7812845336559429f, -1.6309059012653515f,
            -1.0382179034313539f, -0.85f, -0.7f, -0.9f, -1.2360638576849587f};
        static const float b[] = {5089.15750884921511936f, -0.320049391452786891f, -0.35362849922161446f,
            -0.30340000000000003f, -0.61f, -0.5f, -0.5f, -0.6f};
        static const float c[] = {3397.77603275308720128f, -0.321327362693153371f, -0.34507619223117997f,
            -0.70340000000000003f, -0.9f, -1.0f, -1.0f, -1.1754605576265209f};
        set_rows(p[7].dct_param, &p[7].n_dct, a, b, c, 8); p[7].mode = MODE_DCT;
    }
    {
        static const float a[] = {13844.97076442300573f, -0.97113799999999995f, -0.658f, -0.42026f, -0.22712f, -0.2206f, -0.226f, -0.6f};
        static const float b[] = {4798.964084220744293f, -0.61125308982767057f, -0.83770786552491361f, -0.79014862079498627f,
            -0.2692727459704829f, -0.38272769465388551f, -0.22924222653091453f, -0.20719098826199578f};
        static const float c[] = {1807.236946760964614f, -1.2f, -1.2f, -0.7f, -0.7f, -0.7f, -0.4f, -0.5f};
        set_rows(p[8].dct_param, &p[8].n_dct, a, b, c, 8); p[8].mode = MODE_DCT;
    }
    static const float d48a[] = {2198.050556016380522f, -0.96269623020744692f, -0.76194253026666783f, -0.6551140670773547f};
    static const float d48b[] = {764.3655248643528689f, -0.92630200888366945f, -0.9675229603596517f, -0.27845290869168118f};
    static const float d48c[] = {527.107573587542228f, -1.4594385811273854f, -1.450082094097871593f, -1.5843722511996204f};
    {
        static const float one1[] = {1.0f};
        set_rows(p[9].dct_param, &p[9].n_dct, d48a, d48b, d48c, 4);
        set_param_rows(p[9].param, &p[9].n_param, one1, one1, one1, 1);
        p[9].mode = MODE_DCT4_8;
    }
    {
        static const float a[] = {3072.0f, 3072.0f, 256.0f, 256.0f, 256.0f, 414.0f, 0.0f, 0.0f, 0.0f};
        static const float b[] = {1024.0f, 1024.0f, 50.0f, 50.0f, 50.0f, 58.0f, 0.0f, 0.0f, 0.0f};
        static const float c[] = {384.0f, 384.0f, 12.0f, 12.0f, 12.0f, 22.0f, -0.25f, -0.25f, -0.25f};
        set_rows(p[10].dct_param, &p[10].n_dct, d48a, d48b, d48c, 4);
        set_param_rows(p[10].param, &p[10].n_param, a, b, c, 9);
        set_rows(p[10].params4x4, &p[10].n_4x4, d44a, d44b, d44c, 4);
        p[10].mode = MODE_AFV;
    }
    static const float seqA[] = {-1.025f, -0.78f, -0.65012f, -0.19041574084286472f,
        -0.20819395464f, -0.421064f, -0.32733845535848671f};
    static const float seqB[] = {-0.3041958212306401f, -0.3633036457487539f, -0.35660379990111464f, -0.3443074455424403f,
        -0.33699592683512467f, -0.30180866526242109f, -0.27321683125358037f};
    static const float seqC[] = {-1.2f, -1.2f, -0.8f, -0.7f, -0.7f, -0.4f, -0.5f};
    static const float heads[6][3] = {
        {23966.1665298448605f, 8380.19148390090414f, 4493.02378009847706f},
        {15358.89804933239925f, 5597.360516150652990f, 2919.961618960011210f},
        {47932.3330596897210f, 16760.38296780180828f, 8986.04756019695412f},
        {30717.796098664792f, 11194.72103230130598f, 5839.92323792002242f},
        {95864.6661193794420f, 33520.76593560361656f, 17972.09512039390824f},
        {61435.5921973295970f, 24209.44206460261196f, 12979.84647584004484f},
    };
    for (int i = 0; i < 6; i++) {
        orc_qm_params *q = &p[11 + i];
        prepend(q->dct_param[0], heads[i][0], seqA);
        prepend(q->dct_param[1], heads[i][1], seqB);
        prepend(q->dct_param[2], heads[i][2], seqC);
        q->n_dct = 8;
        q->mode = MODE_DCT;
    }
}

/* HFGlobal.interpolate :42-53 */
static float qm_interpolate(float scaledPos, const float *bands, int nbands) {
    int len = nbands - 1;
    if (len == 0) return bands[0];
    int scaledIndex = (int)scaledPos;
    float fracIndex = scaledPos - scaledIndex;
    if (scaledIndex + 1 > len) return bands[len];
    float a = bands[scaledIndex];
    float b = bands[scaledIndex + 1];
    return a * (float)pow(b / a, fracIndex);
}
/* HFGlobal.quantMult :55-57 */
static float quant_mult(float v) { return v >= 0 ? 1.0f + v : 1.0f / (1.0f - v); }

/* HFGlobal.getDCTQuantWeights :59-77 -> weights[height][width] (pitch width) */
static void get_dct_quant_weights(int height, int width, const float *params, int nparams, float *weights) {
    float bands[17];
    bands[0] = params[0];
    for (int i = 1; i < nparams; i++) bands[i] = bands[i - 1] * quant_mult(params[i]);
    const float SQRT_2 = (float)sqrt(2.0);
    float scale = (nparams - 1) / (SQRT_2 + 1e-6f);
    for (int y = 0; y < height; y++) {
        float dy = (float)y * scale / (height - 1);
        float dy2 = dy * dy;
        for (int x = 0; x < width; x++) {
            float dx = (float)x * scale / (width - 1);
            float dist = (float)sqrt(dx * dx + dy2);
            weights[y * width + x] = qm_interpolate(dist, bands, nparams);
        }
    }
}

/* HFGlobal.getAFVTransformWeights :304-345 -> weight[8][8] */
static int get_afv_weights(const orc_qm_params *prm, int c, float *weight) {
    float weights4x8[4 * 8], weights4x4[4 * 4];
    get_dct_quant_weights(4, 8, prm->dct_param[c], prm->n_dct, weights4x8);
    get_dct_quant_weights(4, 4, prm->params4x4[c], prm->n_4x4, weights4x4);
    float low = 0.8517778890324296f;
    float high = 12.97166202570235f;
    float bands[4];
    bands[0] = prm->param[c][5];
    if (bands[0] < 0) return -2;
    for (int i = 1; i < 4; i++) {
        bands[i] = bands[i - 1] * quant_mult(prm->param[c][i + 5]);
        if (bands[i] < 0) return -2;
    }
    memset(weight, 0, sizeof(float) * 64);
    weight[0 * 8 + 0] = 1.0f;
    weight[1 * 8 + 0] = prm->param[c][0];
    weight[0 * 8 + 1] = prm->param[c][1];
    weight[2 * 8 + 0] = prm->param[c][2];
    weight[0 * 8 + 2] = prm->param[c][3];
    weight[2 * 8 + 2] = prm->param[c][4];
    for (int y = 0; y < 4; y++) {
        for (int x = 0; x < 4; x++) {
            if (x < 2 && y < 2) continue;
            float pos = (afvFreqs[y * 4 + x] - low) / (high - low);
            weight[(2 * x) * 8 + 2 * y] = qm_interpolate(pos, bands, 4);
        }
        for (int x = 0; x < 8; x++) {
            if (x == 0 && y == 0) continue;
            weight[(2 * y + 1) * 8 + x] = weights4x8[y * 8 + x];
        }
        for (int x = 0; x < 4; x++) {
            if (x == 0 && y == 0) continue;
            weight[(2 * y) * 8 + 2 * x + 1] = weights4x4[y * 4 + x];
        }
    }
    return 0;
}

/* HFGlobal.generateWeights :347-432, for all 17 parameter sets */
int32_t orc_qm_generate(const orc_qm_params params[17], float *weights, int32_t offsets[51]) {
    int32_t off = 0;
    for (int index = 0; index < 17; index++) {
        const tt_t *tt = tt_by_param(index);
        const int mh = tt_matrix_h(tt), mw = tt_matrix_w(tt);
        const orc_qm_params *prm = &params[index];
        for (int c = 0; c < 3; c++) {
            float *wt = weights + off;
            offsets[index * 3 + c] = off;
            off += mh * mw;
            float w[64];
            switch (prm->mode) {
            case MODE_DCT:
                get_dct_quant_weights(mh, mw, prm->dct_param[c], prm->n_dct, wt);
                break;
            case MODE_DCT4:
                get_dct_quant_weights(4, 4, prm->dct_param[c], prm->n_dct, w);
                for (int y = 0; y < 8; y++)
                    for (int x = 0; x < 8; x++)
                        wt[y * 8 + x] = w[(y / 2) * 4 + x / 2];
                wt[1 * 8 + 0] /= prm->param[c][0];
                wt[0 * 8 + 1] /= prm->param[c][0];
                wt[1 * 8 + 1] /= prm->param[c][1];
                break;
            case MODE_DCT2:
                memset(w, 0, sizeof(w));
                w[0] = 1.0f;
                w[0 * 8 + 1] = w[1 * 8 + 0] = prm->param[c][0];
                w[1 * 8 + 1] = prm->param[c][1];
                for (int y = 0; y < 2; y++)
                    for (int x = 0; x < 2; x++) {
                        w[y * 8 + x + 2] = w[(x + 2) * 8 + y] = prm->param[c][2];
                        w[(y + 2) * 8 + x + 2] = prm->param[c][3];
                    }
                for (int y = 0; y < 4; y++)
                    for (int x = 0; x < 4; x++) {
                        w[y * 8 + x + 4] = w[(x + 4) * 8 + y] = prm->param[c][4];
                        w[(y + 4) * 8 + x + 4] = prm->param[c][5];
                    }
                memcpy(wt, w, sizeof(w));
                break;
            case MODE_HORNUSS:
                for (int i = 0; i < 64; i++) w[i] = prm->param[c][0];
                w[1 * 8 + 1] = prm->param[c][2];
                w[0 * 8 + 1] = w[1 * 8 + 0] = prm->param[c][1];
                w[0] = 1.0f;
                memcpy(wt, w, sizeof(w));
                break;
            case MODE_DCT4_8: {
                float w48[32];
                get_dct_quant_weights(4, 8, prm->dct_param[c], prm->n_dct, w48);
                for (int y = 0; y < 8; y++)
                    for (int x = 0; x < 8; x++)
                        wt[y * 8 + x] = w48[(y / 2) * 8 + x];
                wt[1 * 8 + 0] /= prm->param[c][0];
                break;
            }
            case MODE_AFV:
                if (get_afv_weights(prm, c, wt)) return -2;
                break;
            case MODE_RAW:
                for (int y = 0; y < mh; y++)
                    for (int x = 0; x < mw; x++)
                        wt[y * mw + x] = prm->raw[c][y * mw + x] * prm->denominator;
                break;
            default:
                return -2;
            }
        }
        if (prm->mode != MODE_RAW) {
            for (int c = 0; c < 3; c++) {
                float *wt = weights + offsets[index * 3 + c];
                for (int i = 0; i < mh * mw; i++) {
                    if (wt[i] <= 0.0f || !isfinite(wt[i])) return -2;
                    wt[i] = 1.0f / wt[i];
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * HFMetadata.placeBlock  (J/frame/vardct/HFMetadata.java:38-53, 93-119)
 * ---------------------------------------------------------------------------------------------- */
int32_t orc_place_blocks(int32_t hb, int32_t wb, int32_t n_blocks, const int32_t *types, const int32_t *muls,
    uint8_t *dct_select, uint8_t *block_origin, int32_t *hf_mul, int32_t pitch) {
    int lastY = 0, lastX = 0;
    for (int i = 0; i < n_blocks; i++) {
        int type = types[i];
        if (type > 26 || type < 0) return -2;
        const int bh = TT[type].pixelH >> 3, bw = TT[type].pixelW >> 3;
        int placed = 0;
        for (int y = lastY, x = lastX; y < hb && !placed; y++, x = 0) {
            for (; x < wb; x++) {
                if (bw + x > wb) break;               /* "block too big to put here": continue outerY */
                int occupied = 0;
                for (int ix = 0; ix < bw; ix++) {
                    uint8_t t = dct_select[y * pitch + x + ix];
                    if (t != 255) {
                        x += (TT[t].pixelW >> 3) - 1;
                        occupied = 1;
                        break;
                    }
                }
                if (occupied) continue;
                if (y + bh > hb) return -2;           /* Java would throw ArrayIndexOutOfBounds here */
                for (int iy = 0; iy < bh; iy++)
                    for (int ix = 0; ix < bw; ix++) {
                        dct_select[(y + iy) * pitch + x + ix] = (uint8_t)type;
                        hf_mul[(y + iy) * pitch + x + ix] = muls[i];
                    }
                block_origin[y * pitch + x] = 1;
                lastY = y; lastX = x;
                placed = 1;
                break;
            }
        }
        if (!placed) return -2;
    }
    return n_blocks;
}

/* ------------------------------------------------------------------------------------------------
 * PassGroup  (J/frame/group/PassGroup.java:19-58, 83-331)
 * ---------------------------------------------------------------------------------------------- */
static const float AFV_BASIS[16][16] = {{0.25f, 0.25f, 0.25f, 0.25f, 0.25f, 0.25f, 0.25f, 0.25f, 0.25f, 0.25f, 0.25f,
    0.25f, 0.25f, 0.25f, 0.25f, 0.25f}, {0.876902929799142f, 0.2206518106944235f, -0.10140050393753763f,
    -0.1014005039375375f, 0.2206518106944236f, -0.10140050393753777f, -0.10140050393753772f, -0.10140050393753763f,
    -0.10140050393753758f, -0.10140050393753769f, -0.1014005039375375f, -0.10140050393753768f, -0.10140050393753768f,
    -0.10140050393753759f, -0.10140050393753763f, -0.10140050393753741f}, {0.0f, 0.0f, 0.40670075830260755f,
    0.44444816619734445f, 0.0f, 0.0f, 0.19574399372042936f, 0.2929100136981264f, -0.40670075830260716f,
    -0.19574399372042872f, 0.0f, 0.11379074460448091f, -0.44444816619734384f, -0.29291001369812636f,
    -0.1137907446044814f, 0.0f}, {0.0f, 0.0f, -0.21255748058288748f, 0.3085497062849767f, 0.0f, 0.4706702258572536f,
    -0.1621205195722993f, 0.0f, -0.21255748058287047f, -0.16212051957228327f, -0.47067022585725277f,
    -0.1464291867126764f, 0.3085497062849487f, 0.0f, -0.14642918671266536f, 0.4251149611657548f}, {0.0f,
    -0.7071067811865474f, 0.0f, 0.0f, 0.7071067811865476f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f},
    {-0.4105377591765233f, 0.6235485373547691f, -0.06435071657946274f, -0.06435071657946266f, 0.6235485373547694f,
    -0.06435071657946284f, -0.0643507165794628f, -0.06435071657946274f, -0.06435071657946272f, -0.06435071657946279f,
    -0.06435071657946266f, -0.06435071657946277f, -0.06435071657946277f, -0.06435071657946273f, -0.06435071657946274f,
    -0.0643507165794626f}, {0.0f, 0.0f, -0.4517556589999482f, 0.15854503551840063f, 0.0f, -0.04038515160822202f,
    0.0074182263792423875f, 0.39351034269210167f, -0.45175565899994635f, 0.007418226379244351f, 0.1107416575309343f,
    0.08298163094882051f, 0.15854503551839705f, 0.3935103426921022f, 0.0829816309488214f, -0.45175565899994796f},
    {0.0f, 0.0f, -0.304684750724869f, 0.5112616136591823f, 0.0f, 0.0f, -0.290480129728998f, -0.06578701549142804f,
    0.304684750724884f, 0.2904801297290076f, 0.0f, -0.23889773523344604f, -0.5112616136592012f, 0.06578701549142545f,
    0.23889773523345467f, 0.0f}, {0.0f, 0.0f, 0.3017929516615495f, 0.25792362796341184f, 0.0f, 0.16272340142866204f,
    0.09520022653475037f, 0.0f, 0.3017929516615503f, 0.09520022653475055f, -0.16272340142866173f, -0.35312385449816297f,
    0.25792362796341295f, 0.0f, -0.3531238544981624f, -0.6035859033230976f}, {0.0f, 0.0f, 0.40824829046386274f, 0.0f, 0.0f,
    0.0f, 0.0f, -0.4082482904638628f, -0.4082482904638635f, 0.0f, 0.0f, -0.40824829046386296f, 0.0f, 0.4082482904638634f,
    0.408248290463863f, 0.0f}, {0.0f, 0.0f, 0.1747866975480809f, 0.0812611176717539f, 0.0f, 0.0f, -0.3675398009862027f,
    -0.307882213957909f, -0.17478669754808135f, 0.3675398009862011f, 0.0f, 0.4826689115059883f, -0.08126111767175039f,
    0.30788221395790305f, -0.48266891150598584f, 0.0f}, {0.0f, 0.0f, -0.21105601049335784f, 0.18567180916109802f, 0.0f, 0.0f,
    0.49215859013738733f, -0.38525013709251915f, 0.21105601049335806f, -0.49215859013738905f, 0.0f, 0.17419412659916217f,
    -0.18567180916109904f, 0.3852501370925211f, -0.1741941265991621f, 0.0f}, {0.0f, 0.0f, -0.14266084808807264f,
    -0.3416446842253372f, 0.0f, 0.7367497537172237f, 0.24627107722075148f, -0.08574019035519306f, -0.14266084808807344f,
    0.24627107722075137f, 0.14883399227113567f, -0.04768680350229251f, -0.3416446842253373f, -0.08574019035519267f,
    -0.047686803502292804f, -0.14266084808807242f}, {0.0f, 0.0f, -0.13813540350758585f, 0.3302282550303788f, 0.0f,
    0.08755115000587084f, -0.07946706605909573f, -0.4613374887461511f, -0.13813540350758294f, -0.07946706605910261f,
    0.49724647109535086f, 0.12538059448563663f, 0.3302282550303805f, -0.4613374887461554f, 0.12538059448564315f,
    -0.13813540350758452f}, {0.0f, 0.0f, -0.17437602599651067f, 0.0702790691196284f, 0.0f, -0.2921026642334881f,
    0.3623817333531167f, 0.0f, -0.1743760259965108f, 0.36238173335311646f, 0.29210266423348785f, -0.4326608024727445f,
    0.07027906911962818f, 0.0f, -0.4326608024727457f, 0.34875205199302267f}, {0.0f, 0.0f, 0.11354987314994337f,
    -0.07417504595810355f, 0.0f, 0.19402893032594343f, -0.435190496523228f, 0.21918684838857466f, 0.11354987314994257f,
    -0.4351904965232251f, 0.5550443808910661f, -0.25468277124066463f, -0.07417504595810233f, 0.2191868483885728f,
    -0.25468277124066413f, 0.1135498731499429f},
};
const float *orc_afv_basis(void) { return &AFV_BASIS[0][0]; }

/* scratchBlock[i] : float[256][256] */
#define SB(i) (scratch + (size_t)(i) * SCR * SCR)

/* PassGroup.layBlock :83-86 */
static void lay_block(const float *block, int bp, float *buffer, int dp, int h, int w) {
    for (int y = 0; y < h; y++) memcpy(buffer + (size_t)y * dp, block + (size_t)y * bp, sizeof(float) * w);
}

/* PassGroup.invertAFV :88-147.  coeffs -> element (ppg), cp pitch; buffer -> element (ppf), dp pitch. */
static void invert_afv(const float *coeffs, int cp, float *buffer, int dp, int type, float *scratch) {
    float *s0 = SB(0), *s1 = SB(1);
    s0[0] = (coeffs[0] + coeffs[cp] + coeffs[1]) * 4.0f;
    for (int iy = 0; iy < 4; iy++)
        for (int ix = (iy == 0 ? 1 : 0); ix < 4; ix++)
            s0[iy * SCR + ix] = coeffs[(iy * 2) * cp + ix * 2];
    int flipY = (type == 16 || type == 17) ? 1 : 0; /* AFV2 || AFV3 */
    int flipX = (type == 15 || type == 17) ? 1 : 0; /* AFV1 || AFV3 */
    for (int iy = 0; iy < 4; iy++) {
        for (int ix = 0; ix < 4; ix++) {
            float sample = 0.0f;
            for (int j = 0; j < 16; j++) {
                int jy = j >> 2;
                int jx = j & 3;
                sample += s0[jy * SCR + jx] * AFV_BASIS[j][iy * 4 + ix];
            }
            s1[iy * SCR + ix] = sample;
        }
    }
    for (int iy = 0; iy < 4; iy++)
        for (int ix = 0; ix < 4; ix++)
            buffer[(flipY * 4 + iy) * dp + flipX * 4 + ix] = s1[(flipY == 1 ? 3 - iy : iy) * SCR + (flipX == 1 ? 3 - ix : ix)];
    /* SPEC: watch signs here */
    s0[0] = coeffs[0] + coeffs[cp] - coeffs[1];
    for (int iy = 0; iy < 4; iy++)
        for (int ix = (iy == 0 ? 1 : 0); ix < 4; ix++)
            s0[iy * SCR + ix] = coeffs[(iy * 2) * cp + ix * 2 + 1];
    inverse_dct_2d(s0, SCR, s1, SCR, 4, 4, SB(2), SB(3), 0);
    for (int iy = 0; iy < 4; iy++)
        for (int ix = 0; ix < 4; ix++) /* transposed intentionally */
            buffer[(flipY * 4 + iy) * dp + (flipX == 1 ? 0 : 4) + ix] = s1[ix * SCR + iy];
    s0[0] = coeffs[0] - coeffs[cp];
    for (int iy = 0; iy < 4; iy++)
        for (int ix = (iy == 0 ? 1 : 0); ix < 8; ix++)
            s0[iy * SCR + ix] = coeffs[(1 + iy * 2) * cp + ix];
    inverse_dct_2d(s0, SCR, s1, SCR, 4, 8, SB(2), SB(3), 0);
    for (int iy = 0; iy < 4; iy++)
        for (int ix = 0; ix < 8; ix++)
            buffer[((flipY == 1 ? 0 : 4) + iy) * dp + ix] = s1[iy * SCR + ix];
}

/* PassGroup.auxDCT2 :149-168 */
static void aux_dct2(const float *coeffs, int cp, float *result, int rp, int s) {
    lay_block(coeffs, cp, result, rp, 8, 8);
    int num = s / 2;
    for (int iy = 0; iy < num; iy++) {
        for (int ix = 0; ix < num; ix++) {
            float c00 = coeffs[iy * cp + ix];
            float c01 = coeffs[iy * cp + ix + num];
            float c10 = coeffs[(iy + num) * cp + ix];
            float c11 = coeffs[(iy + num) * cp + ix + num];
            float r00 = c00 + c01 + c10 + c11;
            float r01 = c00 + c01 - c10 - c11;
            float r10 = c00 - c01 + c10 - c11;
            float r11 = c00 - c01 - c10 + c11;
            result[(iy * 2) * rp + ix * 2] = r00;
            result[(iy * 2) * rp + ix * 2 + 1] = r01;
            result[(iy * 2 + 1) * rp + ix * 2] = r10;
            result[(iy * 2 + 1) * rp + ix * 2 + 1] = r11;
        }
    }
}

/* One varblock x channel of PassGroup.invertVarDCT :227-328.
 * coeffs -> dequantHFCoeff[c] at ppg (pitch cp), frame -> frameBuffer[c] at ppf (pitch fp). */
static int invert_varblock(const float *coeffs, int cp, float *frame, int fp, int type, float *scratch) {
    const tt_t *tt = &TT[type];
    float coeff0, coeff1;
    float lfs[2];
    switch (tt->method) {
    case METHOD_DCT:
        inverse_dct_2d(coeffs, cp, frame, fp, tt->pixelH, tt->pixelW, SB(0), SB(1), 0);
        break;
    case METHOD_DCT8_4:
        coeff0 = coeffs[0];
        coeff1 = coeffs[cp];
        lfs[0] = coeff0 + coeff1;
        lfs[1] = coeff0 - coeff1;
        for (int x = 0; x < 2; x++) {
            float *s0 = SB(0);
            s0[0] = lfs[x];
            for (int iy = 0; iy < 4; iy++)
                for (int ix = (iy == 0 ? 1 : 0); ix < 8; ix++)
                    s0[iy * SCR + ix] = coeffs[(x + iy * 2) * cp + ix];
            inverse_dct_2d(s0, SCR, frame + (x << 2), fp, 4, 8, SB(1), SB(2), 1);
        }
        break;
    case METHOD_DCT4_8:
        coeff0 = coeffs[0];
        coeff1 = coeffs[cp];
        lfs[0] = coeff0 + coeff1;
        lfs[1] = coeff0 - coeff1;
        for (int y = 0; y < 2; y++) {
            float *s0 = SB(0);
            s0[0] = lfs[y];
            for (int iy = 0; iy < 4; iy++)
                for (int ix = (iy == 0 ? 1 : 0); ix < 8; ix++)
                    s0[iy * SCR + ix] = coeffs[(y + iy * 2) * cp + ix];
            inverse_dct_2d(s0, SCR, frame + (size_t)(y << 2) * fp, fp, 4, 8, SB(1), SB(2), 0);
        }
        break;
    case METHOD_AFV:
        invert_afv(coeffs, cp, frame, fp, type, scratch);
        break;
    case METHOD_DCT2:
        aux_dct2(coeffs, cp, SB(0), SCR, 2);
        aux_dct2(SB(0), SCR, SB(1), SCR, 4);
        aux_dct2(SB(1), SCR, frame, fp, 8);
        break;
    case METHOD_HORNUSS: {
        float *s0 = SB(0), *s1 = SB(1);
        aux_dct2(coeffs, cp, s1, SCR, 2);
        for (int y = 0; y < 2; y++) {
            for (int x = 0; x < 2; x++) {
                float blockLF = s1[y * SCR + x];
                float residual = 0.0f;
                for (int iy = 0; iy < 4; iy++)
                    for (int ix = (iy == 0 ? 1 : 0); ix < 4; ix++)
                        residual += coeffs[(y + iy * 2) * cp + x + ix * 2];
                s0[(4 * y + 1) * SCR + 4 * x + 1] = blockLF - residual * 0.0625f;
                for (int iy = 0; iy < 4; iy++) {
                    for (int ix = 0; ix < 4; ix++) {
                        if (ix == 1 && iy == 1) continue;
                        s0[(y * 4 + iy) * SCR + x * 4 + ix] =
                            coeffs[(y + iy * 2) * cp + x + ix * 2] + s0[(4 * y + 1) * SCR + 4 * x + 1];
                    }
                }
                s0[(4 * y) * SCR + 4 * x] = coeffs[(y + 2) * cp + x + 2] + s0[(4 * y + 1) * SCR + 4 * x + 1];
            }
        }
        lay_block(s0, SCR, frame, fp, 8, 8);
        break;
    }
    case METHOD_DCT4: {
        float *s0 = SB(0), *s1 = SB(1), *s4 = SB(4);
        aux_dct2(coeffs, cp, s4, SCR, 2);
        for (int y = 0; y < 2; y++) {
            for (int x = 0; x < 2; x++) {
                s0[0] = s4[y * SCR + x];
                for (int iy = 0; iy < 4; iy++)
                    for (int ix = (iy == 0 ? 1 : 0); ix < 4; ix++)
                        s0[iy * SCR + ix] = coeffs[(y + iy * 2) * cp + x + ix * 2];
                inverse_dct_2d(s0, SCR, s1, SCR, 4, 4, SB(2), SB(3), 1);
                for (int iy = 0; iy < 4; iy++)
                    for (int ix = 0; ix < 4; ix++)
                        frame[(4 * y + iy) * fp + 4 * x + ix] = s1[iy * SCR + ix];
            }
        }
        break;
    }
    default:
        return -3;
    }
    return 0;
}

/* One varblock, one channel: exported for known-answer tests of every TransformType. */
int32_t orc_invert_varblock(const float *coeffs, int32_t cp, float *frame, int32_t fp, int32_t type) {
    init_cosine_lut();
    if (type < 0 || type > 26) return -2;
    float *scratch = (float *)malloc(sizeof(float) * 5 * SCR * SCR);
    int r = invert_varblock(coeffs, cp, frame, fp, type, scratch);
    free(scratch);
    return r;
}

/* ------------------------------------------------------------------------------------------------
 * HFCoefficients.bakeDequantizedCoeffs for one group + PassGroup.invertVarDCT
 * (J/frame/vardct/HFCoefficients.java:140-229, 267-319; J/frame/group/PassGroup.java:202-330)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int y, x; } pt_t; /* block position in frame block units */

static int process_group(const orc_frame_params *p, int gy, int gx,
    const int32_t *const qcoeff[3], const float *const lf[3],
    const uint8_t *dct_select, const uint8_t *block_origin, const int32_t *hf_mul,
    const int32_t *x_from_y, const int32_t *b_from_y,
    const float *qm_weights, const int32_t *qm_offsets,
    float *const out[3], float *const dq[3], float *scratch, pt_t *blocks, float *xFactors, float *bFactors) {
    const int W = p->width, H = p->height;
    const int wb = W >> 3, hb = H >> 3;
    const int tw = (W + 63) >> 6;
    /* blocks[] of this group in blockList order = raster order of top-left (HFMetadata.java:38-53; HFCoefficients.java:74-83) */
    int nblocks = 0;
    const int by0 = gy << 5, bx0 = gx << 5;
    const int by1 = by0 + 32 < hb ? by0 + 32 : hb, bx1 = bx0 + 32 < wb ? bx0 + 32 : wb;
    for (int y = by0; y < by1; y++)
        for (int x = bx0; x < bx1; x++)
            if (block_origin[y * wb + x]) { blocks[nblocks].y = y; blocks[nblocks].x = x; nblocks++; }

    int subsampled = 0;
    for (int c = 0; c < 3; c++) if (p->shift_x[c] != 0 || p->shift_y[c] != 0) subsampled = 1;
    /* chroma subsampling (jpegUpsamplingY/X): channel c lives in a (H >> sy) x (W >> sx) plane; a varblock takes part in
     * channel c only when its block position is a multiple of the subsampling factor, and then sits at the shifted
     * position with the SAME TransformType (HFCoefficients.java:290-303, PassGroup.java:217-226). */
    int Wc[3], wbc[3];
    for (int c = 0; c < 3; c++) { Wc[c] = W >> p->shift_x[c]; wbc[c] = wb >> p->shift_x[c]; }
#define SUB_SKIP(pos, c) ((((pos).y >> p->shift_y[c]) << p->shift_y[c]) != (pos).y || (((pos).x >> p->shift_x[c]) << p->shift_x[c]) != (pos).x)
#define SUB_ORG(pos, c) ((size_t)(((pos).y >> p->shift_y[c]) << 3) * Wc[c] + (((pos).x >> p->shift_x[c]) << 3))

    /* ---- dequantizeHFCoefficients :267-319 ---- */
    float globalScale = 65536.0f / p->global_scale;
    float scaleFactor[3] = {
        globalScale * (float)pow(0.8, p->xqm_scale - 2.0),
        globalScale,
        globalScale * (float)pow(0.8, p->bqm_scale - 2.0),
    };
    float qbclut[3][3];
    for (int c = 0; c < 3; c++) { qbclut[c][0] = -p->quant_bias[c]; qbclut[c][1] = 0.0f; qbclut[c][2] = p->quant_bias[c]; }
    for (int i = 0; i < nblocks; i++) {
        pt_t pos = blocks[i];
        int type = dct_select[pos.y * wb + pos.x];
        if (type > 26) return -2;
        const tt_t *tt = &TT[type];
        int flip = tt_flip(tt);
        const int mw = tt_matrix_w(tt);
        const int dsH = tt->pixelH >> 3, dsW = tt->pixelW >> 3;
        for (int c = 0; c < 3; c++) {
            if (SUB_SKIP(pos, c)) continue; /* subsampled block */
            const float *w3 = qm_weights + qm_offsets[tt->parameterIndex * 3 + c];
            float sfc = scaleFactor[c] / hf_mul[pos.y * wb + pos.x];
            const float *qbc = qbclut[c];
            const size_t org = SUB_ORG(pos, c);
            for (int y = 0; y < tt->pixelH; y++) {
                for (int x = 0; x < tt->pixelW; x++) {
                    if (y < dsH && x < dsW) continue;
                    size_t idx = org + (size_t)y * Wc[c] + x;
                    int coeff = qcoeff[c][idx];
                    float quant = (coeff > -2 && coeff < 2) ? qbc[coeff + 1] : coeff - p->quant_bias_numerator / coeff;
                    int wy = flip ? x : y;
                    int wx = x ^ y ^ wy;
                    dq[c][idx] = quant * sfc * w3[wy * mw + wx];
                }
            }
        }
    }

    /* ---- chromaFromLuma :146-192 (xFactors/bFactors are fresh zero arrays per call, like the Java);
     * skipped entirely when any channel is subsampled (:149-151) ---- */
    if (!subsampled) {
        const int th = (H + 63) >> 6;
        memset(xFactors, 0, sizeof(float) * (size_t)th * tw);
        memset(bFactors, 0, sizeof(float) * (size_t)th * tw);
        for (int i = 0; i < nblocks; i++) {
            pt_t pos = blocks[i];
            const tt_t *tt = &TT[dct_select[pos.y * wb + pos.x]];
            int pPosY = pos.y << 3;
            int pPosX = pos.x << 3;
            for (int iy = 0; iy < tt->pixelH; iy++) {
                int y = pPosY + iy;
                int fy = y >> 6;
                int by = (fy << 6) == y;
                float *xF = xFactors + (size_t)fy * tw;
                float *bF = bFactors + (size_t)fy * tw;
                const int32_t *hfX = x_from_y + (size_t)fy * tw;
                const int32_t *hfB = b_from_y + (size_t)fy * tw;
                for (int ix = 0; ix < tt->pixelW; ix++) {
                    int x = pPosX + ix;
                    int fx = x >> 6;
                    float kX, kB;
                    if (by && (fx << 6) == x) {
                        kX = p->base_corr_x + hfX[fx] / (float)p->color_factor;
                        kB = p->base_corr_b + hfB[fx] / (float)p->color_factor;
                        xF[fx] = kX;
                        bF[fx] = kB;
                    } else {
                        kX = xF[fx];
                        kB = bF[fx];
                    }
                    size_t idx = (size_t)y * W + x;
                    float dequantY = dq[1][idx];
                    dq[0][idx] += kX * dequantY;
                    dq[2][idx] += kB * dequantY;
                }
            }
        }
    }

    /* ---- finalizeLLF :194-229 ---- */
    {
        float *s0 = scratch, *s1 = scratch + 32 * 32;
        for (int i = 0; i < nblocks; i++) {
            pt_t pos = blocks[i];
            int type = dct_select[pos.y * wb + pos.x];
            const tt_t *tt = &TT[type];
            const int dsH = tt->pixelH >> 3, dsW = tt->pixelW >> 3;
            for (int c = 0; c < 3; c++) {
                if (SUB_SKIP(pos, c)) continue;
                const float *dqlf = lf[c] + (size_t)(pos.y >> p->shift_y[c]) * wbc[c] + (pos.x >> p->shift_x[c]);
                float *d = dq[c] + SUB_ORG(pos, c);
                forward_dct_2d(dqlf, wbc[c], d, Wc[c], dsH, dsW, s0, s1, 32);
                for (int y = 0; y < dsH; y++)
                    for (int x = 0; x < dsW; x++)
                        d[(size_t)y * Wc[c] + x] *= orc_llf_scale(type, y, x);
            }
        }
    }

    /* ---- invertVarDCT :209-330 ---- */
    for (int i = 0; i < nblocks; i++) {
        pt_t pos = blocks[i];
        int type = dct_select[pos.y * wb + pos.x];
        for (int c = 0; c < 3; c++) {
            if (SUB_SKIP(pos, c)) continue;
            size_t o = SUB_ORG(pos, c);
            int r = invert_varblock(dq[c] + o, Wc[c], out[c] + o, Wc[c], type, scratch);
            if (r) return r;
        }
    }
#undef SUB_SKIP
#undef SUB_ORG
    return 0;
}

int32_t orc_vardct_invert(const orc_frame_params *p,
    const int32_t *const qcoeff[3], const float *const lf[3],
    const uint8_t *dct_select, const uint8_t *block_origin, const int32_t *hf_mul,
    const int32_t *x_from_y, const int32_t *b_from_y,
    const float *qm_weights, const int32_t *qm_offsets,
    float *const out[3], float *const dequant_out[3], int32_t nthreads) {
    init_cosine_lut();
    const int W = p->width, H = p->height;
    if (W <= 0 || H <= 0 || (W & 7) || (H & 7)) return -1;
    const int groupRows = (H + 255) >> 8, groupCols = (W + 255) >> 8; /* Frame.java:127-129, groupDim 256 */
    float *dq[3];
    int own = dequant_out == NULL || dequant_out[0] == NULL;
    size_t plane_n[3];
    for (int c = 0; c < 3; c++) plane_n[c] = (size_t)(W >> p->shift_x[c]) * (H >> p->shift_y[c]);
    for (int c = 0; c < 3; c++)
        dq[c] = own ? (float *)calloc(plane_n[c], sizeof(float)) : dequant_out[c];
    if (!own) for (int c = 0; c < 3; c++) memset(dq[c], 0, sizeof(float) * plane_n[c]);
    int rc = 0;
    const int tw = (W + 63) >> 6, th = (H + 63) >> 6;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        float *scratch = (float *)malloc(sizeof(float) * 5 * SCR * SCR);
        pt_t *blocks = (pt_t *)malloc(sizeof(pt_t) * 1024);
        float *xF = (float *)malloc(sizeof(float) * (size_t)tw * th);
        float *bF = (float *)malloc(sizeof(float) * (size_t)tw * th);
#pragma omp for schedule(dynamic, 1)
        for (int g = 0; g < groupRows * groupCols; g++) {
            int gy = g / groupCols, gx = g % groupCols; /* Frame.getGroupLocation :883-885 */
            int r = process_group(p, gy, gx, qcoeff, lf, dct_select, block_origin, hf_mul, x_from_y, b_from_y,
                                  qm_weights, qm_offsets, out, dq, scratch, blocks, xF, bF);
            if (r) {
#pragma omp critical(orc_rc)
                rc = r;
            }
        }
        free(scratch); free(blocks); free(xF); free(bF);
    }
    if (own) for (int c = 0; c < 3; c++) free(dq[c]);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Frame.performGabConvolution  (J/frame/Frame.java:505-542)
 * ---------------------------------------------------------------------------------------------- */
void orc_gab(const orc_frame_params *p, const float *const in[3], float *const out[3], int32_t nthreads) {
    const int width = p->width, height = p->height;
    float normGabBase[3], normGabAdj[3], normGabDiag[3];
    for (int c = 0; c < 3; c++) {
        float gabW1 = p->gab_w1[c];
        float gabW2 = p->gab_w2[c];
        float mult = 1.0f / (1.0f + 4.0f * (gabW1 + gabW2));
        normGabBase[c] = mult;
        normGabAdj[c] = gabW1 * mult;
        normGabDiag[c] = gabW2 * mult;
    }
    if (nthreads < 1) nthreads = 1;
    for (int c = 0; c < 3; c++) {
        const float *buffC = in[c];
        float *newBufferF = out[c];
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int y = 0; y < height; y++) {
            int north = (y == 0 ? 0 : y - 1);
            int south = (y + 1 == height) ? height - 1 : y + 1;
            const float *buffR = buffC + (size_t)y * width;
            const float *buffN = buffC + (size_t)north * width;
            const float *buffS = buffC + (size_t)south * width;
            float *newBuffR = newBufferF + (size_t)y * width;
            for (int x = 0; x < width; x++) {
                int west = (x == 0 ? 0 : x - 1);
                int east = (x + 1 == width ? width - 1 : x + 1);
                float adj = buffR[west] + buffR[east] + buffN[x] + buffS[x];
                float diag = buffN[west] + buffN[east] + buffS[west] + buffS[east];
                newBuffR[x] = normGabBase[c] * buffR[x] + normGabAdj[c] * adj + normGabDiag[c] * diag;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Frame.performEdgePreservingFilter  (J/frame/Frame.java:44-55, 544-679)
 * ---------------------------------------------------------------------------------------------- */
static const int epfCross[5][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}}; /* Point(y, x) */
static const int epfDoubleCross[13][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}, {1, -1}, {-1, -1},
    {0, -2}, {0, 2}, {2, 0}, {-2, 0}};

/* epfDistance1 :638-655 */
static float epf_distance1(const orc_frame_params *p, float *const buffer[3], int basePosY, int basePosX, const int *dCross) {
    const int H = p->height, W = p->width;
    float dist = 0.0f;
    for (int c = 0; c < 3; c++) {
        const float *buffC = buffer[c];
        float scale = p->epf_channel_scale[c];
        for (int k = 0; k < 5; k++) {
            int pY = orc_mirror_coordinate(basePosY + epfCross[k][0], H);
            int pX = orc_mirror_coordinate(basePosX + epfCross[k][1], W);
            int dY = orc_mirror_coordinate(basePosY + dCross[0] + epfCross[k][0], H);
            int dX = orc_mirror_coordinate(basePosX + dCross[1] + epfCross[k][1], W);
            dist += fabsf(buffC[(size_t)pY * W + pX] - buffC[(size_t)dY * W + dX]) * scale;
        }
    }
    return dist;
}
/* epfDistance2 :657-669 */
static float epf_distance2(const orc_frame_params *p, float *const buffer[3], int basePosY, int basePosX, const int *cross) {
    const int H = p->height, W = p->width;
    float dist = 0.0f;
    for (int c = 0; c < 3; c++) {
        const float *buffC = buffer[c];
        int dY = orc_mirror_coordinate(basePosY + cross[0], H);
        int dX = orc_mirror_coordinate(basePosX + cross[1], W);
        dist += fabsf(buffC[(size_t)basePosY * W + basePosX] - buffC[(size_t)dY * W + dX]) * p->epf_channel_scale[c];
    }
    return dist;
}
/* epfWeight :671-679 */
static float epf_weight(const orc_frame_params *p, float sigmaScale, float distance, float inverseSigma, int refY, int refX) {
    int modY = refY & 7;
    int modX = refX & 7;
    if (modY == 0 || modY == 7 || modX == 0 || modX == 7)
        distance *= p->epf_border_sad_mul;
    float v = 1.0f - distance * sigmaScale * inverseSigma;
    return v < 0.0f ? 0.0f : v;
}

static int32_t epf_core(const orc_frame_params *p, float *const buf[3], float *inverseSigma, int32_t nthreads);

int32_t orc_epf(const orc_frame_params *p, float *const buf[3], const int32_t *hf_mul, const int32_t *sharpness, int32_t nthreads) {
    const int H = p->height, W = p->width;
    int blockHeight = (H + 7) >> 3;
    int blockWidth = (W + 7) >> 3;
    float *inverseSigma = (float *)malloc(sizeof(float) * (size_t)blockHeight * blockWidth);
    float globalScale = 65536.0f / p->global_scale;
    for (int y = 0; y < blockHeight; y++) {
        for (int x = 0; x < blockWidth; x++) {
            int hf = hf_mul[y * blockWidth + x];
            int sharp = sharpness[y * blockWidth + x];
            if (sharp < 0 || sharp > 7) { free(inverseSigma); return -2; }
            float sigma = globalScale * p->epf_sharp_lut[sharp] / hf;
            inverseSigma[y * blockWidth + x] = 1.0f / sigma;
        }
    }
    return epf_core(p, buf, inverseSigma, nthreads);
}

/* Modular-encoded frames: invModularSigma = 1f / header.restorationFilter.epfSigmaForModular for every pixel (Frame.java:573-575,
 * 604-607).  A one-colour frame (`colors == 1 ? 0 : c`, :642, 661) is this function on three copies of the channel. */
int32_t orc_epf_uniform(const orc_frame_params *p, float *const buf[3], float epf_sigma_for_modular, int32_t nthreads) {
    const int H = p->height, W = p->width;
    const size_t nb = (size_t)((H + 7) >> 3) * ((W + 7) >> 3);
    float *inverseSigma = (float *)malloc(sizeof(float) * nb);
    const float inv = 1.0f / epf_sigma_for_modular;
    for (size_t i = 0; i < nb; i++) inverseSigma[i] = inv;
    return epf_core(p, buf, inverseSigma, nthreads);
}

/* takes ownership of inverseSigma */
static int32_t epf_core(const orc_frame_params *p, float *const buf[3], float *inverseSigma, int32_t nthreads) {
    const float SQRT_H = (float)sqrt(0.5);
    float stepMultiplier = 1.65f * 4.0f * (1.0f - SQRT_H);
    const int H = p->height, W = p->width;
    int blockWidth = (W + 7) >> 3;
    float *bufA[3], *bufB[3];
    for (int c = 0; c < 3; c++) {
        bufA[c] = buf[c];
        bufB[c] = (float *)malloc(sizeof(float) * (size_t)W * H);
    }
    if (nthreads < 1) nthreads = 1;
    int swaps = 0;
    for (int i = 0; i < 3; i++) {
        if (i == 0 && p->epf_iters < 3) continue;
        if (i == 2 && p->epf_iters < 2) break;
        float *inputBuffers[3] = {bufA[0], bufA[1], bufA[2]};
        float *outputBuffers[3] = {bufB[0], bufB[1], bufB[2]};
        float sigmaScale;
        if (i == 0) sigmaScale = stepMultiplier * p->epf_pass0_sigma_scale;
        else if (i == 2) sigmaScale = stepMultiplier * p->epf_pass2_sigma_scale;
        else sigmaScale = stepMultiplier;
        const int (*crossList)[2] = i == 0 ? epfDoubleCross : epfCross;
        const int ncross = i == 0 ? 13 : 5;
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int y = 0; y < H; y++) {
            float sumChannels[3];
            for (int x = 0; x < W; x++) {
                float s = inverseSigma[(y >> 3) * blockWidth + (x >> 3)];
                if (s != s || s > (1.0f / 0.3f)) {
                    for (int c = 0; c < 3; c++)
                        outputBuffers[c][(size_t)y * W + x] = inputBuffers[c][(size_t)y * W + x];
                    continue;
                }
                float sumWeights = 0.0f;
                sumChannels[0] = sumChannels[1] = sumChannels[2] = 0.0f;
                for (int k = 0; k < ncross; k++) {
                    const int *cross = crossList[k];
                    float dist = i == 2 ? epf_distance2(p, inputBuffers, y, x, cross)
                                        : epf_distance1(p, inputBuffers, y, x, cross);
                    float weight = epf_weight(p, sigmaScale, dist, s, y, x);
                    sumWeights += weight;
                    int mY = orc_mirror_coordinate(y + cross[0], H);
                    int mX = orc_mirror_coordinate(x + cross[1], W);
                    for (int c = 0; c < 3; c++)
                        sumChannels[c] += inputBuffers[c][(size_t)mY * W + mX] * weight;
                }
                for (int c = 0; c < 3; c++)
                    outputBuffers[c][(size_t)y * W + x] = sumChannels[c] / sumWeights;
            }
        }
        for (int c = 0; c < 3; c++) { float *t = bufA[c]; bufA[c] = bufB[c]; bufB[c] = t; }
        swaps++;
    }
    /* result lives in bufA; copy back into the caller's planes if an odd number of swaps happened */
    if (swaps & 1)
        for (int c = 0; c < 3; c++) memcpy(buf[c], bufA[c], sizeof(float) * (size_t)W * H);
    for (int c = 0; c < 3; c++) free((swaps & 1) ? bufA[c] : bufB[c]);
    free(inverseSigma);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * performColorTransforms + OpsinInverseMatrix.invertXYB
 * (J/JXLCodestreamDecoder.java:256-283; J/color/OpsinInverseMatrix.java:81-84, 105-142)
 * ---------------------------------------------------------------------------------------------- */
void orc_color(const orc_frame_params *p, float *const buf[3], int32_t nthreads) {
    const size_t n = (size_t)p->width * p->height;
    if (nthreads < 1) nthreads = 1;
    if (p->color_mode & 1) {
        const float itScale = 255.0f / p->intensity_target;
        float scaledMatrix[9];
        for (int i = 0; i < 9; i++) scaledMatrix[i] = p->opsin_matrix[i] * itScale;
        const float ob0 = p->opsin_bias[0], ob1 = p->opsin_bias[1], ob2 = p->opsin_bias[2];
        const float cob0 = -(float)cbrt(p->opsin_bias[0]);
        const float cob1 = -(float)cbrt(p->opsin_bias[1]);
        const float cob2 = -(float)cbrt(p->opsin_bias[2]);
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (size_t i = 0; i < n; i++) {
            const float xybX = buf[0][i];
            const float xybY = buf[1][i];
            const float xybB = buf[2][i];
            const float gammaL = xybY + xybX + cob0;
            const float gammaM = xybY - xybX + cob1;
            const float gammaS = xybB + cob2;
            const float mixL = (gammaL * gammaL) * gammaL + ob0;
            const float mixM = (gammaM * gammaM) * gammaM + ob1;
            const float mixS = (gammaS * gammaS) * gammaS + ob2;
            buf[0][i] = scaledMatrix[0] * mixL + scaledMatrix[1] * mixM + scaledMatrix[2] * mixS;
            buf[1][i] = scaledMatrix[3] * mixL + scaledMatrix[4] * mixM + scaledMatrix[5] * mixS;
            buf[2][i] = scaledMatrix[6] * mixL + scaledMatrix[7] * mixM + scaledMatrix[8] * mixS;
        }
    }
    if (p->color_mode & 2) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (size_t i = 0; i < n; i++) {
            float cb = buf[0][i];
            float yh = buf[1][i] + 0.50196078431372549019f;
            float cr = buf[2][i];
            buf[0][i] = yh + 1.402f * cr;
            buf[1][i] = yh - 0.34413628620102214650f * cb - 0.71413628620102214650f * cr;
            buf[2][i] = yh + 1.772f * cb;
        }
    }
}

/* Frame.invertSubsampling (J/frame/Frame.java:681-723): per channel, xShift horizontal doublings then yShift vertical
 * doublings; out = 0.75f * centre + 0.25f * neighbour, neighbour clamped at the plane edge.
 * in[c]: (H >> sy) x (W >> sx); out[c]: H x W (may be the same pointer when the channel is not subsampled). */
void orc_invert_subsampling(const orc_frame_params *p, const float *const in[3], float *const out[3]) {
    const int W = p->width, H = p->height;
    for (int c = 0; c < 3; c++) {
        int w = W >> p->shift_x[c], h = H >> p->shift_y[c];
        float *cur = (float *)malloc(sizeof(float) * (size_t)w * h);
        memcpy(cur, in[c], sizeof(float) * (size_t)w * h);
        int xShift = p->shift_x[c];
        while (xShift-- > 0) {
            float *nw = (float *)malloc(sizeof(float) * (size_t)w * 2 * h);
            for (int y = 0; y < h; y++) {
                const float *oldRow = cur + (size_t)y * w;
                float *newRow = nw + (size_t)y * w * 2;
                for (int x = 0; x < w; x++) {
                    float b75 = 0.75f * oldRow[x];
                    newRow[2 * x] = b75 + 0.25f * oldRow[x == 0 ? 0 : x - 1];
                    newRow[2 * x + 1] = b75 + 0.25f * oldRow[x + 1 == w ? w - 1 : x + 1];
                }
            }
            free(cur); cur = nw; w *= 2;
        }
        int yShift = p->shift_y[c];
        while (yShift-- > 0) {
            float *nw = (float *)malloc(sizeof(float) * (size_t)w * h * 2);
            for (int y = 0; y < h; y++) {
                const float *oldRow = cur + (size_t)y * w;
                const float *oldRowPrev = cur + (size_t)(y == 0 ? 0 : y - 1) * w;
                const float *oldRowNext = cur + (size_t)(y + 1 == h ? h - 1 : y + 1) * w;
                float *firstNewRow = nw + (size_t)(2 * y) * w;
                float *secondNewRow = nw + (size_t)(2 * y + 1) * w;
                for (int x = 0; x < w; x++) {
                    float b75 = 0.75f * oldRow[x];
                    firstNewRow[x] = b75 + 0.25f * oldRowPrev[x];
                    secondNewRow[x] = b75 + 0.25f * oldRowNext[x];
                }
            }
            free(cur); cur = nw; h *= 2;
        }
        memcpy(out[c], cur, sizeof(float) * (size_t)W * H);
        free(cur);
    }
}

/* Frame.decodeFrame tail :457-461 + JXLCodestreamDecoder.decode :637 */
int32_t orc_vardct_reconstruct(const orc_frame_params *p,
    const int32_t *const qcoeff[3], const float *const lf[3],
    const uint8_t *dct_select, const uint8_t *block_origin, const int32_t *hf_mul,
    const int32_t *x_from_y, const int32_t *b_from_y, const int32_t *sharpness,
    const float *qm_weights, const int32_t *qm_offsets,
    float *const out[3], int32_t nthreads) {
    int subsampled = 0;
    for (int c = 0; c < 3; c++) if (p->shift_x[c] != 0 || p->shift_y[c] != 0) subsampled = 1;
    int rc;
    if (!subsampled) {
        rc = orc_vardct_invert(p, qcoeff, lf, dct_select, block_origin, hf_mul, x_from_y, b_from_y,
                               qm_weights, qm_offsets, out, NULL, nthreads);
        if (rc) return rc;
    } else {
        float *sub[3];
        for (int c = 0; c < 3; c++)
            sub[c] = (float *)calloc((size_t)(p->width >> p->shift_x[c]) * (p->height >> p->shift_y[c]), sizeof(float));
        rc = orc_vardct_invert(p, qcoeff, lf, dct_select, block_origin, hf_mul, x_from_y, b_from_y,
                               qm_weights, qm_offsets, sub, NULL, nthreads);
        if (!rc) {
            const float *in[3] = {sub[0], sub[1], sub[2]};
            orc_invert_subsampling(p, in, out);
        }
        for (int c = 0; c < 3; c++) free(sub[c]);
        if (rc) return rc;
    }
    if (p->gab) {
        const size_t n = (size_t)p->width * p->height;
        float *tmp[3];
        for (int c = 0; c < 3; c++) tmp[c] = (float *)malloc(sizeof(float) * n);
        const float *in[3] = {out[0], out[1], out[2]};
        orc_gab(p, in, tmp, nthreads);
        for (int c = 0; c < 3; c++) { memcpy(out[c], tmp[c], sizeof(float) * n); free(tmp[c]); }
    }
    if (p->epf_iters > 0) {
        rc = orc_epf(p, out, hf_mul, sharpness, nthreads);
        if (rc) return rc;
    }
    orc_color(p, out, nthreads);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Frame / patch blending  (J/JXLCodestreamDecoder.java:285-413): blendAdd, blendMult, blendBlend, blendMulAdd on one
 * rectangle of one channel.  a = the Java's `frame` argument at frameOffset, b = its `ref` argument at refOffset.
 * ---------------------------------------------------------------------------------------------- */
static float clamp_asc(float v, float lo, float hi) { return v < lo ? lo : v > hi ? hi : v; }

int32_t orc_blend(int32_t mode, int32_t is_int, int32_t is_alpha, int32_t has_extra, int32_t clamp, int32_t premult,
    int32_t h, int32_t w, void *canvas, int64_t cp, const void *a, int64_t ap, const void *b, int64_t bp,
    const float *fa, int64_t fap, const float *ra, int64_t rap) {
    if (mode < 1 || mode > 4) return -2;
    if ((mode == 2 || mode == 3) && !has_extra) mode = 1;
    if (is_int && mode != 1) return -1;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            if (mode == 1) {
                if (is_int)
                    ((int32_t *)canvas)[y * cp + x] = ((const int32_t *)b)[y * bp + x] + ((const int32_t *)a)[y * ap + x];
                else
                    ((float *)canvas)[y * cp + x] = ((const float *)b)[y * bp + x] + ((const float *)a)[y * ap + x];
                continue;
            }
            float *cf = (float *)canvas + y * cp + x;
            const float ff = ((const float *)a)[y * ap + x], rf = ((const float *)b)[y * bp + x];
            if (mode == 4) {
                float newSample = ff;
                if (clamp) newSample = clamp_asc(newSample, 0.0f, 1.0f);
                *cf = newSample * rf;
            } else if (mode == 2) {
                float oldSample = rf, newSample = ff;
                float oldAlpha = is_alpha ? oldSample : ra[y * rap + x];
                float newAlpha = is_alpha ? newSample : fa[y * fap + x];
                if (clamp) newAlpha = clamp_asc(newAlpha, 0.0f, 1.0f);
                if (is_alpha) *cf = oldAlpha + newAlpha * (1.0f - oldAlpha);
                else if (premult) *cf = newSample + oldSample * (1.0f - newAlpha);
                else *cf = (newSample * newAlpha + oldSample * oldAlpha * (1.0f - newAlpha)) / (oldAlpha + newAlpha * (1.0f - oldAlpha));
            } else {
                float oldSample = rf, newSample = ff, newAlpha = fa[y * fap + x];
                if (clamp) newAlpha = clamp_asc(newAlpha, 0.0f, 1.0f);
                *cf = oldSample + newAlpha * newSample;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Frame.performUpsampling  (J/frame/Frame.java:217-260): k x k upsampling of one float channel with the per-phase 5x5
 * kernels weights[ky][kx][iy][ix] (ImageHeader.getUpWeights), mirrored edges, result clamped to the [min, max] of the 25
 * samples -- with the Java's initial values (min = Float.MAX_VALUE, max = Float.MIN_VALUE, the smallest POSITIVE float).
 * ---------------------------------------------------------------------------------------------- */
void orc_upsample(const float *in, int32_t h, int32_t w, int32_t k, const float *weights, float *out) {
    for (int y = 0; y < h; y++) {
        for (int ky = 0; ky < k; ky++) {
            for (int x = 0; x < w; x++) {
                for (int kx = 0; kx < k; kx++) {
                    const float *wt = weights + (size_t)(ky * k + kx) * 25;
                    float total = 0.0f;
                    float min = 3.4028234663852886e38f;
                    float max = 1.401298464324817e-45f;
                    for (int iy = 0; iy < 5; iy++) {
                        for (int ix = 0; ix < 5; ix++) {
                            int newY = orc_mirror_coordinate(y + iy - 2, h);
                            int newX = orc_mirror_coordinate(x + ix - 2, w);
                            float sample = in[(size_t)newY * w + newX];
                            if (sample < min) min = sample;
                            if (sample > max) max = sample;
                            total += wt[iy * 5 + ix] * sample;
                        }
                    }
                    out[(size_t)(y * k + ky) * w * k + x * k + kx] = total < min ? min : total > max ? max : total;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Noise: XorShiro (J/frame/features/XorShiro.java), Frame.initializeNoise / synthesizeNoise (J/frame/Frame.java:748-835)
 * ---------------------------------------------------------------------------------------------- */
static uint64_t split_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
typedef struct { uint64_t state0[8], state1[8]; uint32_t batch[16]; int batchPos; } xorshiro_t;
static void xs_init(xorshiro_t *r, uint64_t seed0, uint64_t seed1) {
    r->state0[0] = split_mix64(seed0 + 0x9e3779b97f4a7c15ULL);
    r->state1[0] = split_mix64(seed1 + 0x9e3779b97f4a7c15ULL);
    for (int i = 1; i < 8; i++) {
        r->state0[i] = split_mix64(r->state0[i - 1]);
        r->state1[i] = split_mix64(r->state1[i - 1]);
    }
    r->batchPos = 16;
}
static void xs_fill_batch(xorshiro_t *r) {
    for (int i = 0; i < 8; i++) {
        const uint64_t a = r->state1[i];
        uint64_t b = r->state0[i];
        const uint64_t c = a + b;
        r->state0[i] = a;
        b ^= b << 23;
        r->state1[i] = b ^ a ^ (b >> 18) ^ (a >> 5);
        r->batch[2 * i] = (uint32_t)(c & 0xffffffffULL);
        r->batch[2 * i + 1] = (uint32_t)(c >> 32);
    }
    r->batchPos = 0;
}
static void xs_fill(xorshiro_t *r, uint32_t *bits, int n) {
    for (int i = 0; i < n; i++) {
        if (r->batchPos >= 16) xs_fill_batch(r);
        bits[i] = r->batch[r->batchPos++];
    }
}

/* planes[3]: X, Y, B of the (upsampled) frame, h x w, modified in place.  seed0 = (visibleFrames << 32) | invisibleFrames. */
void orc_noise(float *const planes[3], int32_t h, int32_t w, int32_t group_dim, int64_t seed0, const float *lut,
    float base_x, float base_b) {
    static const float laplacian[5][5] = {
        {0.16f, 0.16f, 0.16f, 0.16f, 0.16f}, {0.16f, 0.16f, 0.16f, 0.16f, 0.16f}, {0.16f, 0.16f, -3.84f, 0.16f, 0.16f},
        {0.16f, 0.16f, 0.16f, 0.16f, 0.16f}, {0.16f, 0.16f, 0.16f, 0.16f, 0.16f}};
    const size_t n = (size_t)h * w;
    float *local[3], *noise[3];
    for (int c = 0; c < 3; c++) { local[c] = (float *)calloc(n, sizeof(float)); noise[c] = (float *)calloc(n, sizeof(float)); }
    const int log_dim = group_dim == 128 ? 7 : group_dim == 256 ? 8 : group_dim == 512 ? 9 : 10;
    const int groupRowStride = (w + group_dim - 1) / group_dim;
    const int numGroups = groupRowStride * ((h + group_dim - 1) / group_dim);
    for (int group = 0; group < numGroups; group++) {
        int y0 = (group / groupRowStride) << log_dim;
        int x0 = (group % groupRowStride) << log_dim;
        uint64_t seed1 = (((uint64_t)(uint32_t)x0) << 32) | (uint64_t)(uint32_t)y0;
        int ySize = group_dim < h - y0 ? group_dim : h - y0;
        int xSize = group_dim < w - x0 ? group_dim : w - x0;
        xorshiro_t rng;
        xs_init(&rng, (uint64_t)seed0, seed1);
        uint32_t bits[16];
        for (int c = 0; c < 3; c++)
            for (int y = 0; y < ySize; y++)
                for (int x = 0; x < xSize; x += 16) {
                    xs_fill(&rng, bits, 16);
                    for (int i = 0; i < 16 && x + i < xSize; i++) {
                        uint32_t f = (bits[i] >> 9) | 0x3f800000u;
                        float v;
                        memcpy(&v, &f, 4);
                        local[c][(size_t)(y0 + y) * w + x0 + x + i] = v;
                    }
                }
    }
    for (int c = 0; c < 3; c++)
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                float acc = 0.0f;
                for (int iy = 0; iy < 5; iy++)
                    for (int ix = 0; ix < 5; ix++) {
                        int cy = orc_mirror_coordinate(y + iy - 2, h);
                        int cx = orc_mirror_coordinate(x + ix - 2, w);
                        acc += local[c][(size_t)cy * w + cx] * laplacian[iy][ix];
                    }
                noise[c][(size_t)y * w + x] = acc;
            }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const size_t i = (size_t)y * w + x;
            float inScaledR = planes[1][i] + planes[0][i];
            inScaledR = inScaledR < 0.0f ? 0.0f : 3.0f * inScaledR;
            float inScaledG = planes[1][i] - planes[0][i];
            inScaledG = inScaledG < 0.0f ? 0.0f : 3.0f * inScaledG;
            int intInR, intInG;
            float fracInR, fracInG;
            if (inScaledR >= 7.0f) { intInR = 6; fracInR = 1.0f; } else { intInR = (int)inScaledR; fracInR = inScaledR - intInR; }
            if (inScaledG >= 7.0f) { intInG = 6; fracInG = 1.0f; } else { intInG = (int)inScaledG; fracInG = inScaledG - intInG; }
            float sr = (lut[intInR + 1] - lut[intInR]) * fracInR + lut[intInR];
            float sg = (lut[intInG + 1] - lut[intInG]) * fracInG + lut[intInG];
            sr = clamp_asc(sr, 0.0f, 1.0f);
            sg = clamp_asc(sg, 0.0f, 1.0f);
            float nr = sr * (0.00171875f * noise[0][i] + 0.21828125f * noise[2][i]);
            float ng = sg * (0.00171875f * noise[1][i] + 0.21828125f * noise[2][i]);
            float nrg = nr + ng;
            planes[1][i] += nrg;
            planes[0][i] += base_x * nrg + nr - ng;
            planes[2][i] += base_b * nrg;
        }
    for (int c = 0; c < 3; c++) { free(local[c]); free(noise[c]); }
}

/* ------------------------------------------------------------------------------------------------
 * Splines  (J/frame/features/spline/Spline.java, J/frame/Frame.java:739-746, MathHelper.erf :40-66)
 * Quirks kept: Spline's constructor drops its splineID, so every spline is drawn with spline 0's coefficients (:22-24,
 * :141); MathHelper.max returns the MINIMUM of its arguments (J/util/MathHelper.java:190-195).
 * ---------------------------------------------------------------------------------------------- */
static float orc_erf(float z) {
    const float az = fabsf(z);
    float absErf;
    if (az > 1e-4f) {
        const float t = 1.0f / (az * 0.5f + 1.0f);
        const float u = t * (t * (t * (t * (t * (t * (t * (t * (t * 0.17087277f - 0.82215223f) + 1.48851587f) - 1.13520398f)
                          + 0.27886807f) - 0.18628806f) + 0.09678418f) + 0.37409196f) + 1.00002368f) - 1.26551223f;
        absErf = 1.0f - t * (float)exp(-z * z + u);
    } else {
        const float t = 1.0f / (az * 0.47047f + 1.0f);
        const float u = t * (t * (t * 0.7478556f - 0.0958798f) + 0.3480242f);
        absErf = 1.0f - u * (float)exp(-z * z);
    }
    if (z < 0) return -absErf;
    return absErf;
}
static int java_f2i(float v) { if (v != v) return 0; if (v >= 2147483648.0f) return 2147483647; if (v <= -2147483648.0f) return -2147483647 - 1; return (int)v; }
static float fourier_ict(const float *coeffs, float t) {
    const float SQRT_H = (float)sqrt(0.5);
    float total = SQRT_H * coeffs[0];
    for (int i = 1; i < 32; i++) total += coeffs[i] * (float)cos(i * (M_PI / 32.0) * (t + 0.5));
    return total;
}
typedef struct { float locationY, locationX, arcLength; } arc_t;

/* points: (x, y) pairs of all splines back to back, npoints[s] pairs each; coeff: [num_splines][4][32] = X, Y, B, sigma */
int32_t orc_splines(float *const planes[3], int32_t h, int32_t w, int32_t num_splines, const int32_t *npoints, const int32_t *points,
    const int32_t *coeff, int32_t quant_adjust, float base_x, float base_b) {
    const float SQRT_F = (float)sqrt(0.125);
    const int32_t *pts = points;
    for (int s = 0; s < num_splines; s++) {
        const int n = npoints[s];
        /* computeCoeffs(splineID = 0) */
        float coeffX[32], coeffY[32], coeffB[32], coeffSigma[32];
        {
            const int32_t *c0 = coeff; /* spline 0 */
            float quantAdjust = quant_adjust / 8.0f;
            float invQa = quantAdjust >= 0 ? 1.0f / (1.0f + quantAdjust) : 1.0f - quantAdjust;
            float yAdjust = 0.106066017f * invQa, xAdjust = 0.005939697f * invQa, bAdjust = 0.098994949f * invQa, sigmaAdjust = 0.47135738f * invQa;
            for (int i = 0; i < 32; i++) {
                coeffY[i] = c0[32 + i] * yAdjust;
                coeffX[i] = c0[i] * xAdjust + base_x * coeffY[i];
                coeffB[i] = c0[64 + i] * bAdjust + base_b * coeffY[i];
                coeffSigma[i] = c0[96 + i] * sigmaAdjust;
            }
        }
        /* upsampleControlPoints */
        int nup;
        float *upY, *upX;
        if (n == 1) {
            nup = 1;
            upY = (float *)malloc(sizeof(float)); upX = (float *)malloc(sizeof(float));
            upY[0] = (float)pts[1]; upX[0] = (float)pts[0];
        } else {
            const int ne = n + 2;
            int *ey = (int *)malloc(sizeof(int) * ne), *ex = (int *)malloc(sizeof(int) * ne);
            ey[0] = pts[1] * 2 - pts[3]; ex[0] = pts[0] * 2 - pts[2];
            for (int i = 0; i < n; i++) { ey[i + 1] = pts[2 * i + 1]; ex[i + 1] = pts[2 * i]; }
            ey[ne - 1] = pts[2 * (n - 1) + 1] * 2 - pts[2 * (n - 2) + 1];
            ex[ne - 1] = pts[2 * (n - 1)] * 2 - pts[2 * (n - 2)];
            nup = 16 * (ne - 3) + 1;
            upY = (float *)malloc(sizeof(float) * nup); upX = (float *)malloc(sizeof(float) * nup);
            float t[4], pY[4], pX[4], dY[3], dX[3], aY[3], aX[3], bY[2], bX[2];
            for (int i = 0; i < ne - 3; i++) {
                for (int k = 0; k < 4; k++) { pY[k] = (float)ey[i + k]; pX[k] = (float)ex[i + k]; }
                upY[i << 4] = pY[1];
                upX[i << 4] = pX[1];
                t[0] = 0.0f;
                for (int k = 0; k < 3; k++) {
                    dY[k] = pY[k + 1] - pY[k];
                    dX[k] = pX[k + 1] - pX[k];
                    t[k + 1] = t[k] + (float)pow(dY[k] * dY[k] + dX[k] * dX[k], 0.25);
                }
                for (int step = 1; step < 16; step++) {
                    float knot = t[1] + 0.0625f * step * (t[2] - t[1]);
                    for (int k = 0; k < 3; k++) {
                        float f = (knot - t[k]) / (t[k + 1] - t[k]);
                        aY[k] = dY[k] * f + pY[k];
                        aX[k] = dX[k] * f + pX[k];
                    }
                    for (int k = 0; k < 2; k++) {
                        float f = (knot - t[k]) / (t[k + 2] - t[k]);
                        bY[k] = (aY[k + 1] - aY[k]) * f + aY[k];
                        bX[k] = (aX[k + 1] - aX[k]) * f + aX[k];
                    }
                    float f = (knot - t[1]) / (t[2] - t[1]);
                    upY[i * 16 + step] = (bY[1] - bY[0]) * f + bY[0];
                    upX[i * 16 + step] = (bX[1] - bX[0]) * f + bX[0];
                }
            }
            upY[nup - 1] = (float)pts[2 * (n - 1) + 1];
            upX[nup - 1] = (float)pts[2 * (n - 1)];
            free(ey); free(ex);
        }
        /* computeIntermediarySamples(renderDistance = 1) */
        const float renderDistance = 1.0f;
        int cap = 1024, narcs = 0;
        arc_t *arcs = (arc_t *)malloc(sizeof(arc_t) * cap);
#define PUSH(y_, x_, l_) do { if (narcs == cap) { cap *= 2; arcs = (arc_t *)realloc(arcs, sizeof(arc_t) * cap); } \
        arcs[narcs].locationY = (y_); arcs[narcs].locationX = (x_); arcs[narcs].arcLength = (l_); narcs++; } while (0)
        float currentY = upY[0], currentX = upX[0];
        int nextID = 0;
        PUSH(currentY, currentX, renderDistance);
        while (nextID < nup) {
            float prevY = currentY, prevX = currentX, arcLengthFromPrevious = 0.0f;
            while (1) {
                if (nextID >= nup) { PUSH(prevY, prevX, arcLengthFromPrevious); break; }
                float nextY = upY[nextID], nextX = upX[nextID];
                float dY = nextY - prevY, dX = nextX - prevX;
                float arcLengthToNext = (float)sqrt(dY * dY + dX * dX);
                if (arcLengthFromPrevious + arcLengthToNext >= renderDistance) {
                    float f = (renderDistance - arcLengthFromPrevious) / arcLengthToNext;
                    currentY = dY * f + prevY;
                    currentX = dX * f + prevX;
                    PUSH(currentY, currentX, renderDistance);
                    break;
                }
                arcLengthFromPrevious += arcLengthToNext;
                prevY = nextY;
                prevX = nextX;
                nextID++;
            }
        }
#undef PUSH
        free(upY); free(upX);
        /* renderSpline */
        float arcLength = (narcs - 2.0f) * renderDistance + arcs[narcs - 1].arcLength;
        if (!(arcLength <= 0.0)) {
            for (int i = 0; i < narcs; i++) {
                arc_t arc = arcs[i];
                float progressAlongArc = fminf(1.0f, i * renderDistance / arcLength);
                float t = 31.0f * progressAlongArc;
                float values[3];
                values[0] = fourier_ict(coeffX, t) * arc.arcLength;
                values[1] = fourier_ict(coeffY, t) * arc.arcLength;
                values[2] = fourier_ict(coeffB, t) * arc.arcLength;
                float sigma = fourier_ict(coeffSigma, t);
                float inverseSigma = 1.0f / sigma;
                float maxColor = 0.01f; /* MathHelper.max(...) is a minimum */
                for (int c = 0; c < 3; c++) maxColor = values[c] < maxColor ? values[c] : maxColor;
                float maxDist = (float)sqrt(-2.0f * sigma * sigma * ((float)log(0.1) * 3.0f - maxColor));
                int xBegin = java_f2i(arc.locationX - maxDist + 0.5f); if (xBegin < 0) xBegin = 0;
                int xEnd = java_f2i(arc.locationX + maxDist + 0.5f); if (xEnd > w - 1) xEnd = w - 1;
                int yBegin = java_f2i(arc.locationY - maxDist + 0.5f); if (yBegin < 0) yBegin = 0;
                int yEnd = java_f2i(arc.locationY + maxDist + 0.5f); if (yEnd > h - 1) yEnd = h - 1;
                for (int c = 0; c < 3; c++)
                    for (int y = yBegin; y <= yEnd; y++)
                        for (int x = xBegin; x <= xEnd; x++) {
                            float dY = y - arc.locationY, dX = x - arc.locationX;
                            float distance = (float)sqrt(dY * dY + dX * dX);
                            float factor = orc_erf((0.5f * distance + SQRT_F) * inverseSigma);
                            factor -= orc_erf((0.5f * distance - SQRT_F) * inverseSigma);
                            float extra = 0.25f * values[c] * sigma * factor * factor;
                            planes[c][(size_t)y * w + x] += extra;
                        }
            }
        }
        free(arcs);
        pts += 2 * n;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * PNG samples: TF_SRGB.fromLinearF (J/color/TransferFunction.java:39-43), ImageBuffer.castToInt0 / clamp
 * (J/util/ImageBuffer.java:129-160), PNGWriter sample order (J/io/PNGWriter.java:191-203).  out: big-endian when bits == 16.
 * ---------------------------------------------------------------------------------------------- */
void orc_pack_samples(const void *const *planes, const int32_t *is_int, const int32_t *depth, int32_t n_channels, int32_t n_color,
    int32_t linear, int32_t h, int32_t w, int32_t bits, uint8_t *out) {
    const int maxValue = ~(~0 << bits), bytes = bits > 8 ? 2 : 1;
    const size_t n = (size_t)h * w;
    for (size_t i = 0; i < n; i++)
        for (int c = 0; c < n_channels; c++) {
            int v;
            if (is_int[c] && depth[c] == bits) {
                v = ((const int32_t *)planes[c])[i];
            } else {
                float f;
                if (is_int[c]) { float scaleFactor = 1.0f / (float)(~(~0 << depth[c])); f = ((const int32_t *)planes[c])[i] * scaleFactor; }
                else f = ((const float *)planes[c])[i];
                if (linear && c < n_color) {
                    if (f < 0.00313066844250063f) f = f * 12.92f;
                    else f = 1.055f * (float)pow(f, 0.4166666666666667) + -0.055f;
                }
                v = java_f2i(f * (float)maxValue + 0.5f);
            }
            v = v < 0 ? 0 : v > maxValue ? maxValue : v;
            uint8_t *o = out + (i * n_channels + c) * bytes;
            if (bytes == 2) { o[0] = (uint8_t)(v >> 8); o[1] = (uint8_t)(v & 255); } else o[0] = (uint8_t)v;
        }
}

/* ------------------------------------------------------------------------------------------------
 * LF coefficients: dequantisation, LF chroma-from-luma, adaptive smoothing
 * (J/frame/vardct/LFCoefficients.java:61-103 and adaptiveSmooth :113-179), per LF group (256 x 256 blocks).
 * lf_quant[i]: frame-level hb x wb planes in FRAME order X, Y, B (i.e. lfQuant[cMap[i]] stitched over LF groups);
 * extra_precision[g]: the 2 bits read per LF group; kx / kb: baseCorrelation + (factorLF - 128) / colorFactor (:81-82).
 * ---------------------------------------------------------------------------------------------- */
void orc_lf_dequant(int32_t hb, int32_t wb, const float scaled_dequant[3], float kx, float kb, int32_t cfl, int32_t smooth,
                    const int32_t *const lf_quant[3], const uint8_t *extra_precision, float *const out[3]) {
    const int gcols = (wb + 255) >> 8, grows = (hb + 255) >> 8;
    for (int gy = 0; gy < grows; gy++)
        for (int gx = 0; gx < gcols; gx++) {
            const int y0 = gy << 8, x0 = gx << 8;
            const int h = hb - y0 < 256 ? hb - y0 : 256, w = wb - x0 < 256 ? wb - x0 : 256;
            const int ep = extra_precision[gy * gcols + gx];
            float *co[3], *weighted[3], *gap = (float *)malloc(sizeof(float) * (size_t)h * w);
            for (int i = 0; i < 3; i++) {
                co[i] = (float *)malloc(sizeof(float) * (size_t)h * w);
                weighted[i] = (float *)calloc((size_t)h * w, sizeof(float));
                const float sd = scaled_dequant[i] / (1 << ep);                       /* :67 */
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++)
                        co[i][y * w + x] = lf_quant[i][(size_t)(y0 + y) * wb + x0 + x] * sd;   /* :72 */
            }
            if (cfl)                                                                    /* :77-94 */
                for (int k = 0; k < h * w; k++) {
                    co[0][k] += kx * co[1][k];
                    co[2][k] += kb * co[1][k];
                }
            if (smooth && h >= 3 && w >= 3) {                                           /* adaptiveSmooth :113-179 */
                for (int k = 0; k < h * w; k++) gap[k] = 0.5f;
                for (int i = 0; i < 3; i++) {
                    const float sd = scaled_dequant[i];
                    for (int y = 1; y < h - 1; y++)
                        for (int x = 1; x < w - 1; x++) {
                            const float *c = co[i] + y * w + x;
                            const float sample = c[0];
                            const float adjacent = c[-1] + c[1] + c[-w] + c[w];
                            const float diag = c[-w - 1] + c[-w + 1] + c[w - 1] + c[w + 1];
                            const float wv = 0.05226273532324128f * sample + 0.20345139757231578f * adjacent + 0.0334829185968739f * diag;
                            weighted[i][y * w + x] = wv;
                            const float g = fabsf(sample - wv) * sd;
                            if (g > gap[y * w + x]) gap[y * w + x] = g;
                        }
                }
                for (int k = 0; k < h * w; k++) {
                    const float v = 3.0f - 4.0f * gap[k];
                    gap[k] = v > 0.0f ? v : 0.0f;                                       /* Math.max(0f, ...) :155 */
                }
                for (int i = 0; i < 3; i++)
                    for (int y = 0; y < h; y++)
                        for (int x = 0; x < w; x++) {
                            const int k = y * w + x;
                            float v = co[i][k];
                            if (!(y == 0 || y + 1 == h || x == 0 || x + 1 == w)) v = (co[i][k] - weighted[i][k]) * gap[k] + weighted[i][k];
                            out[i][(size_t)(y0 + y) * wb + x0 + x] = v;
                        }
            } else {
                for (int i = 0; i < 3; i++)
                    for (int y = 0; y < h; y++)
                        for (int x = 0; x < w; x++) out[i][(size_t)(y0 + y) * wb + x0 + x] = co[i][y * w + x];
            }
            for (int i = 0; i < 3; i++) { free(co[i]); free(weighted[i]); }
            free(gap);
        }
}
