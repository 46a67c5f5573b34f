/*
 * oracle/jxl_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A literal, operation-for-operation C restatement of jxlatte's (Traneptora/jxlatte, pure Java)
 * post-entropy VarDCT reconstruction and Modular inverse transforms.  It exists so the CUDA path
 * can be checked against the reference's arithmetic.  It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or decoded outputs, and no JVM
 * exists in this image, so this restatement cannot be compared with the reference running here.
 * It is pinned only by (a) reading the Java side by side (file:line cited at every function) and
 * (b) independent known-answer checks in tests/ (scipy DCTs, closed forms, inverse-of-forward).
 *
 * Java semantics kept: float32 everywhere Java uses float, no FMA contraction
 * (-ffp-contract=off), same accumulation order, int32 wrapping (-fwrapv), truncating / and %.
 *
 * Citation convention: J/ = /root/reference/java/com/traneptora/jxlatte/
 */
#ifndef JXL_ORACLE_H
#define JXL_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order as include/jxlb200.h:jxlb200_frame_params (kept separate on purpose: the product
 * never includes anything from oracle/). */
typedef struct {
    int32_t width, height;            /* padded frame size, Frame.getPaddedFrameSize (J/frame/Frame.java:924-941) */
    int32_t global_scale;             /* LFGlobal.globalScale */
    int32_t xqm_scale, bqm_scale;     /* FrameHeader.xqmScale / bqmScale */
    float quant_bias[3];              /* OpsinInverseMatrix.quantBias (J/color/OpsinInverseMatrix.java:23-25) */
    float quant_bias_numerator;       /* :27 */
    int32_t color_factor;             /* LFChannelCorrelation.colorFactor */
    float base_corr_x, base_corr_b;
    int32_t shift_x[3], shift_y[3];   /* jpegUpsamplingX/Y */
    int32_t gab;
    float gab_w1[3], gab_w2[3];
    int32_t epf_iters;
    float epf_sharp_lut[8];           /* already multiplied by epfQuantMul (RestorationFilter.java:42-43) */
    float epf_channel_scale[3];
    float epf_pass0_sigma_scale, epf_pass2_sigma_scale, epf_border_sad_mul;
    int32_t color_mode;               /* 0 none, 1 invertXYB, 2 YCbCr->RGB, 3 both (XYB then YCbCr) */
    float opsin_matrix[9];
    float opsin_bias[3];
    float intensity_target;
} orc_frame_params;

/* HFGlobal DCTParams (J/frame/vardct/DCTParams.java) flattened. */
typedef struct {
    int32_t mode;                     /* TransformType.MODE_* */
    int32_t n_dct, n_param, n_4x4;
    float denominator;
    float dct_param[3][17];
    float param[3][9];
    float params4x4[3][17];
    const float *raw[3];              /* MODE_RAW: matrixH*matrixW values per channel */
} orc_qm_params;

#define ORC_QM_TOTAL_PER_CHANNEL 131584   /* sum over 17 parameter sets of matrixH*matrixW */

/* TransformType table lookups (J/frame/vardct/TransformType.java:10-36) */
int32_t orc_tt_info(int32_t type, int32_t *param_index, int32_t *method, int32_t *pixel_h, int32_t *pixel_w, int32_t *flip);

/* HFGlobal.getDefaultParams (J/frame/vardct/HFGlobal.java:79-188) */
void orc_qm_default_params(orc_qm_params out[17]);
/* HFGlobal.generateWeights for all 17 sets (:347-432). weights: 3*131584 floats, laid out
 * [param][channel][matrixH][matrixW]; offsets[p*3+c] = float offset. returns 0, or -2 on invalid weight. */
int32_t orc_qm_generate(const orc_qm_params params[17], float *weights, int32_t offsets[51]);

/* Stage 1: HFCoefficients.bakeDequantizedCoeffs + PassGroup.invertVarDCT over every group of the frame
 * (J/frame/Frame.java:367-373).  All planes are frame-level, row-major, pitch = width (or width/8, ceil(width/64)).
 * dequant_out (optional, may be NULL): receives dequantHFCoeff[3][H][W] after CfL+LLF for inspection. */
int32_t orc_vardct_invert(const orc_frame_params *p,
    const int32_t *const qcoeff[3], const float *const lf[3],
    const uint8_t *dct_select, const uint8_t *block_origin, const int32_t *hf_mul,
    const int32_t *x_from_y, const int32_t *b_from_y,
    const float *qm_weights, const int32_t *qm_offsets,
    float *const out[3], float *const dequant_out[3], int32_t nthreads);

/* PassGroup.invertVarDCT switch body for one varblock x channel (J/frame/group/PassGroup.java:227-328) */
int32_t orc_invert_varblock(const float *coeffs, int32_t cp, float *frame, int32_t fp, int32_t type);

/* Frame.performGabConvolution (J/frame/Frame.java:505-542): in[3] -> out[3] */
void orc_gab(const orc_frame_params *p, const float *const in[3], float *const out[3], int32_t nthreads);
/* Frame.performEdgePreservingFilter (:544-636). buf[3] updated in place (ping-pong handled inside).
 * returns -2 on sharpness outside [0,7]. */
int32_t orc_epf(const orc_frame_params *p, float *const buf[3], const int32_t *hf_mul, const int32_t *sharpness, int32_t nthreads);
/* LFCoefficients.java:61-103, 113-179: LF dequantisation + LF chroma-from-luma + adaptive smoothing, per LF group */
void orc_lf_dequant(int32_t hb, int32_t wb, const float scaled_dequant[3], float kx, float kb, int32_t cfl, int32_t smooth,
                    const int32_t *const lf_quant[3], const uint8_t *extra_precision, float *const out[3]);
/* Modular-encoded frames: one sigma for the frame (Frame.java:573-575, 604-607) */
int32_t orc_epf_uniform(const orc_frame_params *p, float *const buf[3], float epf_sigma_for_modular, int32_t nthreads);
/* JXLCodestreamDecoder.performColorTransforms (J/JXLCodestreamDecoder.java:256-283) */
void orc_color(const orc_frame_params *p, float *const buf[3], int32_t nthreads);

/* Whole path: invert -> gab -> epf -> color.  out[3] receives the final planes. */
void orc_invert_subsampling(const orc_frame_params *p, const float *const in[3], float *const out[3]);
int32_t orc_blend(int32_t mode, int32_t is_int, int32_t is_alpha, int32_t has_extra, int32_t clamp, int32_t premult,
    int32_t h, int32_t w, void *canvas, int64_t cp, const void *a, int64_t ap, const void *b, int64_t bp,
    const float *fa, int64_t fap, const float *ra, int64_t rap);
void orc_upsample(const float *in, int32_t h, int32_t w, int32_t k, const float *weights, float *out);
void orc_noise(float *const planes[3], int32_t h, int32_t w, int32_t group_dim, int64_t seed0, const float *lut,
    float base_x, float base_b);
int32_t orc_splines(float *const planes[3], int32_t h, int32_t w, int32_t num_splines, const int32_t *npoints, const int32_t *points,
    const int32_t *coeff, int32_t quant_adjust, float base_x, float base_b);
void orc_pack_samples(const void *const *planes, const int32_t *is_int, const int32_t *depth, int32_t n_channels, int32_t n_color,
    int32_t linear, int32_t h, int32_t w, int32_t bits, uint8_t *out);
int32_t orc_vardct_reconstruct(const orc_frame_params *p,
    const int32_t *const qcoeff[3], const float *const lf[3],
    const uint8_t *dct_select, const uint8_t *block_origin, const int32_t *hf_mul,
    const int32_t *x_from_y, const int32_t *b_from_y, const int32_t *sharpness,
    const float *qm_weights, const int32_t *qm_offsets,
    float *const out[3], int32_t nthreads);

/* 1-D / 2-D primitives exposed for known-answer tests (J/util/MathHelper.java:68-136) */
void orc_inverse_dct_1d(const float *src, float *dest, int32_t n);
void orc_forward_dct_1d(const float *src, float *dest, int32_t n);
void orc_inverse_dct_2d(const float *src, float *dest, int32_t h, int32_t w, int32_t transposed);
void orc_forward_dct_2d(const float *src, float *dest, int32_t h, int32_t w);
int32_t orc_mirror_coordinate(int32_t coordinate, int32_t size);
float orc_llf_scale(int32_t type, int32_t y, int32_t x);
const float *orc_afv_basis(void);

/* HFMetadata.placeBlock replay (J/frame/vardct/HFMetadata.java:38-53,93-119): place n_blocks of the given
 * types first-fit in an lf-group of hb x wb blocks. Writes dct_select (255 = empty), block_origin, hf_mul.
 * returns number placed, or -2 if one does not fit. */
int32_t orc_place_blocks(int32_t hb, int32_t wb, int32_t n_blocks, const int32_t *types, const int32_t *muls,
    uint8_t *dct_select, uint8_t *block_origin, int32_t *hf_mul, int32_t pitch);

/* ---- Modular (J/frame/modular/ModularStream.java:224-380, ModularChannel.java) ---- */
void orc_modular_rct(int32_t *const ch[3], int32_t h, int32_t w, int32_t rct_type, int32_t *const out[3]);
int32_t orc_modular_palette(const int32_t *idx, const int32_t *palette, int32_t h, int32_t w,
    int32_t num_c, int32_t nb_colors, int32_t nb_deltas, int32_t d_pred, int32_t bit_depth, int32_t *const out[]);
int32_t orc_modular_squeeze(const int32_t *avg, const int32_t *res, int32_t h_avg, int32_t w_avg,
    int32_t h_res, int32_t w_res, int32_t horizontal, int32_t *out);
/* forward transforms written from the inverses, for round-trip tests only */
void orc_modular_forward_squeeze(const int32_t *in, int32_t h, int32_t w, int32_t horizontal, int32_t *avg, int32_t *res);
int32_t orc_tendency(int32_t a, int32_t b, int32_t c);

#ifdef __cplusplus
}
#endif
#endif
