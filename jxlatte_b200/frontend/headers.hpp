// headers.hpp -- image-level and frame-level headers.  Field order and defaults follow jxlatte:
// J/bundle/ImageHeader.java:168-300, BitDepthHeader.java, ExtraChannelInfo.java, Extensions.java, AnimationHeader.java,
// J/color/ColorEncodingBundle.java, ToneMapping.java, OpsinInverseMatrix.java:61-79, CIEXY.java:14-20,
// J/frame/FrameHeader.java:84-200, J/bundle/PassesInfo.java, BlendingInfo.java, J/frame/features/RestorationFilter.java.
#pragma once
#include "entropy.hpp"

namespace jxlf {

struct BitDepth {
    bool is_float = false;
    int bits = 8, exp_bits = 0;
    void read(BitReader &br) {
        is_float = br.flag();
        if (is_float) {
            bits = (int)br.u32(32, 0, 16, 0, 24, 0, 1, 6);
            exp_bits = 1 + (int)br.bits(4);
        } else {
            bits = (int)br.u32(8, 0, 10, 0, 12, 0, 1, 6);
            exp_bits = 0;
        }
    }
};

struct ExtraChannel {
    int type = 0;                // 0 alpha, 1 depth, 2 spot colour, 3 selection mask, 4 black, 5 CFA, 6 thermal, 15/16 (non-)optional
    BitDepth depth;
    int dim_shift = 0;
    std::string name;
    bool alpha_associated = false;
    float spot[4] = {0, 0, 0, 0};
    int cfa_index = 1;
    void read(BitReader &br) {
        const bool default_alpha = br.flag();
        if (!default_alpha) {
            type = (int)br.enumeration();
            if (!((type >= 0 && type <= 6) || type == 15 || type == 16)) throw StreamError("illegal extra channel type");
            depth.read(br);
            dim_shift = (int)br.u32(0, 0, 3, 0, 4, 0, 1, 3);
            const int len = (int)br.u32(0, 0, 0, 4, 16, 5, 48, 10);
            name.resize(len);
            for (auto &ch : name) ch = (char)br.bits(8);
            alpha_associated = type == 0 && br.flag();
        }
        if (type == 2) for (float &v : spot) v = br.f16();
        if (type == 5) cfa_index = (int)br.u32(1, 0, 0, 2, 3, 4, 19, 8);
    }
};

inline void skip_extensions(BitReader &br) {           // Extensions.readExtensions: payload lengths are taken as bytes there
    const uint64_t key = br.u64();
    uint64_t len[64] = {0};
    for (int i = 0; i < 64; i++)
        if (key >> i & 1) len[i] = br.u64();
    for (int i = 0; i < 64; i++)
        if (key >> i & 1) br.skip_bits(len[i] * 8);
}

struct ColorEncoding {
    bool all_default = true, use_icc = false;
    int color_space = 0;         // 0 RGB, 1 grey, 2 XYB, 3 unknown
    int white_point = 1;         // 1 D65, 2 custom, 10 E, 11 DCI
    int primaries = 1;           // 1 sRGB, 2 custom, 9 BT.2100, 11 P3
    int transfer = (1 << 24) + 13;   // gamma*1e7 below 2^24, else 2^24 + enum (13 = sRGB)
    int rendering_intent = 1;
    float white_xy[2] = {0.3127f, 0.3290f};
    float prim_xy[3][2] = {{0.639998686f, 0.330010138f}, {0.300003784f, 0.600003357f}, {0.150002046f, 0.059997204f}};

    static void custom_xy(BitReader &br, float (&xy)[2]) {
        for (float &v : xy) v = (float)unpack_signed(br.u32(0, 19, 524288, 19, 1048576, 20, 2097152, 21)) * 1e-6f;
    }
    void read(BitReader &br) {
        all_default = br.flag();
        if (all_default) return;
        use_icc = br.flag();
        color_space = (int)br.enumeration();
        if (color_space > 3) throw StreamError("invalid colour space");
        if (!use_icc && color_space != 2) white_point = (int)br.enumeration();
        switch (white_point) {
        case 1: break;
        case 2: custom_xy(br, white_xy); break;
        case 10: white_xy[0] = white_xy[1] = 1.0f / 3.0f; break;
        case 11: white_xy[0] = 0.314f; white_xy[1] = 0.351f; break;
        default: throw StreamError("invalid white point");
        }
        if (!use_icc && color_space != 2 && color_space != 1) primaries = (int)br.enumeration();
        switch (primaries) {
        case 1: break;
        case 2: for (auto &p : prim_xy) custom_xy(br, p); break;
        case 9: { const float v[3][2] = {{0.708f, 0.292f}, {0.170f, 0.797f}, {0.131f, 0.046f}}; std::memcpy(prim_xy, v, sizeof v); break; }
        case 11: { const float v[3][2] = {{0.680f, 0.320f}, {0.265f, 0.690f}, {0.150f, 0.060f}}; std::memcpy(prim_xy, v, sizeof v); break; }
        default: throw StreamError("invalid primaries");
        }
        if (!use_icc) {
            if (br.flag()) {
                transfer = (int)br.bits(24);
                if (transfer > 10000000) throw StreamError("illegal gamma");
            } else {
                const int e = (int)br.enumeration();
                if (!(e == 1 || e == 2 || e == 8 || e == 13 || e == 16 || e == 17 || e == 18)) throw StreamError("illegal transfer function");
                transfer = (1 << 24) + e;
            }
            rendering_intent = (int)br.enumeration();
            if (rendering_intent > 3) throw StreamError("invalid rendering intent");
        }
    }
};

struct ImageHeader {
    int level = 5;
    int height = 0, width = 0;
    int orientation = 1;
    int intrinsic_h = 0, intrinsic_w = 0, preview_h = 0, preview_w = 0;
    bool have_animation = false, have_timecodes = false;
    uint32_t tps_num = 0, tps_den = 0, num_loops = 0;
    BitDepth depth;
    bool modular_16bit = true;
    std::vector<ExtraChannel> extra;
    bool xyb_encoded = true;
    ColorEncoding color;
    float intensity_target = 255.0f, min_nits = 0.0f, linear_below = 0.0f;
    bool relative_to_max_display = false;
    // OpsinInverseMatrix
    float opsin_inv[9] = {11.031566901960783f, -9.866943921568629f, -0.16462299647058826f, -3.254147380392157f, 4.418770392156863f,
                          -0.16462299647058826f, -3.6588512862745097f, 2.7129230470588235f, 1.9459282392156863f};
    float opsin_bias[3] = {-0.0037930732552754493f, -0.0037930732552754493f, -0.0037930732552754493f};
    float quant_bias[3] = {0.945349926692846f, 0.9299455010825141f, 0.9500648966626564f};
    float quant_bias_numerator = 0.145f;
    std::vector<float> up2, up4, up8;       // empty = defaults
    std::vector<uint8_t> encoded_icc;

    int color_channels() const { return color.color_space == 1 ? 1 : 3; }

    static void size_header(BitReader &br, int level, int &h, int &w) {
        const bool div8 = br.flag();
        h = div8 ? (int)(1 + br.bits(5)) << 3 : (int)br.u32(1, 9, 1, 13, 1, 18, 1, 30);
        const int ratio = (int)br.bits(3);
        if (ratio) w = width_from_ratio(ratio, h);
        else w = div8 ? (int)(1 + br.bits(5)) << 3 : (int)br.u32(1, 9, 1, 13, 1, 18, 1, 30);
        const int64_t max_dim = level <= 5 ? 1ll << 18 : 1ll << 28, max_area = level <= 5 ? 1ll << 30 : 1ll << 40;
        if (w > max_dim || h > max_dim || (int64_t)w * h > max_area) throw StreamError("image dimensions exceed the codestream level");
    }
    static int width_from_ratio(int ratio, int h) {
        switch (ratio) {
        case 1: return h;
        case 2: return (int)(h * 6ll / 5);
        case 3: return (int)(h * 4ll / 3);
        case 4: return (int)(h * 3ll / 2);
        case 5: return (int)(h * 16ll / 9);
        case 6: return (int)(h * 5ll / 4);
        default: return h * 2;
        }
    }

    void read(BitReader &br, int lvl) {
        level = lvl;
        if (br.bits(16) != 0x0aff) throw StreamError("not a JPEG XL codestream (FF 0A signature missing)");
        size_header(br, level, height, width);
        const bool all_default = br.flag();
        const bool extra_fields = all_default ? false : br.flag();
        if (extra_fields) {
            orientation = 1 + (int)br.bits(3);
            if (br.flag()) size_header(br, level, intrinsic_h, intrinsic_w);
            if (br.flag()) {
                const bool div8 = br.flag();
                auto dim = [&]() { return div8 ? (int)br.u32(16, 0, 32, 0, 1, 5, 33, 9) : (int)br.u32(1, 6, 65, 8, 321, 10, 1345, 12); };
                preview_h = dim();
                const int ratio = (int)br.bits(3);
                preview_w = ratio ? width_from_ratio(ratio, preview_h) : dim();
                if (preview_w > 4096 || preview_h > 4096) throw StreamError("preview too large");
            }
            if (br.flag()) {
                have_animation = true;
                tps_num = br.u32(100, 0, 1000, 0, 1, 10, 1, 30);
                tps_den = br.u32(1, 0, 1001, 0, 1, 8, 1, 10);
                num_loops = br.u32(0, 0, 0, 3, 0, 16, 0, 32);
                have_timecodes = br.flag();
            }
        }
        if (!all_default) {
            depth.read(br);
            modular_16bit = br.flag();
            const int n_extra = (int)br.u32(0, 0, 1, 0, 2, 4, 1, 12);
            extra.resize(n_extra);
            for (auto &e : extra) e.read(br);
            xyb_encoded = br.flag();
            color.read(br);
        }
        if (extra_fields && !br.flag()) {
            intensity_target = br.f16();
            if (intensity_target <= 0) throw StreamError("intensity target must be positive");
            min_nits = br.f16();
            if (min_nits < 0 || min_nits > intensity_target) throw StreamError("min nits out of range");
            relative_to_max_display = br.flag();
            linear_below = br.f16();
            if (linear_below < 0 || (relative_to_max_display && linear_below > 1)) throw StreamError("linear_below out of range");
        }
        if (!all_default) skip_extensions(br);
        const bool default_matrix = br.flag();
        if (!default_matrix && xyb_encoded && !br.flag()) {
            for (float &v : opsin_inv) v = br.f16();
            for (float &v : opsin_bias) v = br.f16();
            for (float &v : quant_bias) v = br.f16();
            quant_bias_numerator = br.f16();
        }
        const int cw_mask = default_matrix ? 0 : (int)br.bits(3);
        if (cw_mask & 1) { up2.resize(15); for (float &v : up2) v = br.f16(); }
        if (cw_mask & 2) { up4.resize(55); for (float &v : up4) v = br.f16(); }
        if (cw_mask & 4) { up8.resize(210); for (float &v : up8) v = br.f16(); }
        if (color.use_icc) {
            const uint64_t n = br.u64();
            if (n > (1u << 28)) throw StreamError("ICC profile too large");
            encoded_icc.resize((size_t)n);
            EntropyStream es(br, 41);
            for (size_t i = 0; i < encoded_icc.size(); i++) encoded_icc[i] = (uint8_t)es.read(br, icc_context(i));
            es.expect_final_state("ICC profile");
        }
        br.align();
    }

  private:
    int icc_context(size_t i) const {                 // ImageHeader.getICCContext :79-116
        if (i <= 128) return 0;
        const int b1 = encoded_icc[i - 1], b2 = encoded_icc[i - 2];
        auto alpha = [](int b) { return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z'); };
        auto digit = [](int b) { return (b >= '0' && b <= '9') || b == '.' || b == ','; };
        int p1, p2;
        if (alpha(b1)) p1 = 0;
        else if (digit(b1)) p1 = 1;
        else if (b1 <= 1) p1 = 2 + b1;
        else if (b1 < 16) p1 = 4;
        else if (b1 > 240 && b1 < 255) p1 = 5;
        else if (b1 == 255) p1 = 6;
        else p1 = 7;
        if (alpha(b2)) p2 = 0;
        else if (digit(b2)) p2 = 1;
        else if (b2 < 16) p2 = 2;
        else if (b2 > 240) p2 = 3;
        else p2 = 4;
        return 1 + p1 + 8 * p2;
    }
};

// ---------------------------------------------------------------------------------------------------------------
enum { FRAME_REGULAR = 0, FRAME_LF = 1, FRAME_REFERENCE_ONLY = 2, FRAME_SKIP_PROGRESSIVE = 3 };
enum { ENC_VARDCT = 0, ENC_MODULAR = 1 };
enum : uint64_t { FLAG_NOISE = 1, FLAG_PATCHES = 2, FLAG_SPLINES = 16, FLAG_USE_LF_FRAME = 32, FLAG_SKIP_ADAPTIVE_LF_SMOOTHING = 128 };

struct Blending {
    int mode = 0, alpha_channel = 0, source = 0;
    bool clamp = false;
    void read(BitReader &br, bool extra, bool full_frame) {
        mode = (int)br.u32(0, 0, 1, 0, 2, 0, 3, 2);
        if (extra && (mode == 2 || mode == 3)) alpha_channel = (int)br.u32(0, 0, 1, 0, 2, 0, 3, 3);
        if (extra && (mode == 2 || mode == 4 || mode == 3)) clamp = br.flag();
        if (mode != 0 || !full_frame) source = (int)br.bits(2);
    }
};

struct Restoration {
    bool gab = true;
    float gab_w1[3] = {0.115169525f, 0.115169525f, 0.115169525f};
    float gab_w2[3] = {0.061248592f, 0.061248592f, 0.061248592f};
    int epf_iters = 2;
    float sharp_lut[8] = {0, 1.0f / 7, 2.0f / 7, 3.0f / 7, 4.0f / 7, 5.0f / 7, 6.0f / 7, 1.0f};   // multiplied by quant_mul after read()
    float channel_scale[3] = {40.0f, 5.0f, 3.5f};
    float quant_mul = 0.46f, pass0_sigma_scale = 0.9f, pass2_sigma_scale = 6.5f, border_sad_mul = 2.0f / 3.0f, sigma_for_modular = 1.0f;
    void read(BitReader &br, int encoding, bool header_all_default) {
        const bool all_default = header_all_default ? true : br.flag();
        if (!all_default) {
            gab = br.flag();
            if (gab && br.flag())
                for (int i = 0; i < 3; i++) { gab_w1[i] = br.f16(); gab_w2[i] = br.f16(); }
            epf_iters = (int)br.bits(2);
            if (epf_iters > 0 && encoding == ENC_VARDCT && br.flag())
                for (float &v : sharp_lut) v = br.f16();
            if (epf_iters > 0 && br.flag()) {
                for (float &v : channel_scale) v = br.f16();
                br.bits(32);
            }
            if (epf_iters > 0 && br.flag()) {
                if (encoding == ENC_VARDCT) quant_mul = br.f16();
                pass0_sigma_scale = br.f16();
                pass2_sigma_scale = br.f16();
                border_sad_mul = br.f16();
            }
            if (epf_iters > 0 && encoding == ENC_MODULAR) sigma_for_modular = br.f16();
            skip_extensions(br);
        }
        for (float &v : sharp_lut) v *= quant_mul;
    }
};

struct FrameHeader {
    int type = FRAME_REGULAR, encoding = ENC_VARDCT;
    uint64_t flags = 0;
    bool do_ycbcr = false;
    int shift_y[3] = {0, 0, 0}, shift_x[3] = {0, 0, 0};      // jpegUpsamplingY/X after normalisation (max - own)
    int upsampling = 1;
    std::vector<int> ec_upsampling;
    int group_size_shift = 1, group_dim = 256;
    int xqm_scale = 3, bqm_scale = 2;
    int num_passes = 1, num_ds = 0;
    bool passes_coded = false;                              // PassesInfo() leaves lastPass EMPTY (PassesInfo.java:14-20): Pass.java then finds no entry
    int pass_shift[11] = {0}, downsample[4] = {1, 1, 1, 1}, last_pass[4] = {0, 0, 0, 0};
    int lf_level = 0;
    bool have_crop = false;
    int x0 = 0, y0 = 0, width = 0, height = 0;              // bounds (size after upsampling / LF-level division, padded for subsampling)
    Blending blending;
    std::vector<Blending> ec_blending;
    uint32_t duration = 0, timecode = 0;
    bool is_last = true;
    int save_as_reference = 0;
    bool save_before_ct = false;
    std::string name;
    Restoration rf;

    void read(BitReader &br, const ImageHeader &ih) {
        const bool all_default = br.flag();
        if (!all_default) {
            type = (int)br.bits(2);
            encoding = (int)br.bits(1);
            flags = br.u64();
            if (!ih.xyb_encoded) do_ycbcr = br.flag();
        }
        int raw_y[3] = {0, 0, 0}, raw_x[3] = {0, 0, 0};
        if (do_ycbcr && !(flags & FLAG_USE_LF_FRAME))
            for (int i = 0; i < 3; i++) {
                const int mode = (int)br.bits(2);
                raw_y[i] = mode == 1 || mode == 3;
                raw_x[i] = mode == 1 || mode == 2;
            }
        ec_upsampling.assign(ih.extra.size(), 1);
        if (!all_default && !(flags & FLAG_USE_LF_FRAME)) {
            upsampling = 1 << br.bits(2);
            for (auto &u : ec_upsampling) u = 1 << br.bits(2);
        }
        group_size_shift = encoding == ENC_MODULAR ? (int)br.bits(2) : 1;
        group_dim = 128 << group_size_shift;
        if (ih.xyb_encoded && encoding == ENC_VARDCT) {
            if (!all_default) { xqm_scale = (int)br.bits(3); bqm_scale = (int)br.bits(3); }
        } else {
            xqm_scale = bqm_scale = 2;
        }
        if (!all_default && type != FRAME_REFERENCE_ONLY) {
            passes_coded = true;
            num_passes = (int)br.u32(1, 0, 2, 0, 3, 0, 4, 3);
            num_ds = num_passes != 1 ? (int)br.u32(0, 0, 1, 0, 2, 0, 3, 1) : 0;
            if (num_ds >= num_passes) throw StreamError("num_ds must be below num_passes");
            for (int i = 0; i < num_passes - 1; i++) pass_shift[i] = (int)br.bits(2);
            pass_shift[num_passes - 1] = 0;
            for (int i = 0; i < num_ds; i++) downsample[i] = 1 << br.bits(2);
            for (int i = 0; i < num_ds; i++) last_pass[i] = (int)br.u32(0, 0, 1, 0, 2, 0, 0, 3);
        }
        downsample[num_ds] = 1;
        last_pass[num_ds] = num_passes - 1;
        lf_level = type == FRAME_LF ? 1 + (int)br.bits(2) : 0;
        have_crop = (!all_default && type != FRAME_LF) ? br.flag() : false;
        if (have_crop && type != FRAME_REFERENCE_ONLY) {
            x0 = unpack_signed(br.u32(0, 8, 256, 11, 2304, 14, 18688, 30));
            y0 = unpack_signed(br.u32(0, 8, 256, 11, 2304, 14, 18688, 30));
        }
        if (have_crop) {
            width = (int)br.u32(0, 8, 256, 11, 2304, 14, 18688, 30);
            height = (int)br.u32(0, 8, 256, 11, 2304, 14, 18688, 30);
        } else {
            width = ih.width;
            height = ih.height;
        }
        const bool normal = !all_default && (type == FRAME_REGULAR || type == FRAME_SKIP_PROGRESSIVE);
        const bool full_frame = y0 <= 0 && x0 <= 0 && y0 + height >= ih.height && x0 + width >= ih.width;
        height = ceil_div(height, upsampling);
        width = ceil_div(width, upsampling);
        height = ceil_div(height, 1 << (3 * lf_level));
        width = ceil_div(width, 1 << (3 * lf_level));
        ec_blending.assign(ih.extra.size(), Blending());
        if (normal) {
            blending.read(br, !ih.extra.empty(), full_frame);
            for (auto &b : ec_blending) b.read(br, true, full_frame);
            if (ih.have_animation) duration = br.u32(0, 0, 1, 0, 0, 8, 0, 32);
            if (ih.have_animation && ih.have_timecodes) timecode = br.bits(32);
            is_last = br.flag();
        } else {
            is_last = type == FRAME_REGULAR;
        }
        save_as_reference = (!all_default && type != FRAME_LF && !is_last) ? (int)br.bits(2) : 0;
        if (!all_default && (type == FRAME_REFERENCE_ONLY ||
                             (full_frame && (type == FRAME_REGULAR || type == FRAME_SKIP_PROGRESSIVE) && (duration == 0 || save_as_reference != 0) &&
                              !is_last && blending.mode == 0)))
            save_before_ct = br.flag();
        if (!all_default) {
            const int len = (int)br.u32(0, 0, 0, 4, 16, 5, 48, 10);
            name.resize(len);
            for (auto &ch : name) ch = (char)br.bits(8);
        }
        rf.read(br, encoding, all_default);
        if (!all_default) skip_extensions(br);
        const int my = std::max(raw_y[0], std::max(raw_y[1], raw_y[2])), mx = std::max(raw_x[0], std::max(raw_x[1], raw_x[2]));
        height = ceil_div(height, 1 << my) << my;
        width = ceil_div(width, 1 << mx) << mx;
        for (int i = 0; i < 3; i++) {
            shift_y[i] = my - raw_y[i];
            shift_x[i] = mx - raw_x[i];
        }
    }
};

}  // namespace jxlf
