// modular.hpp -- Modular sub-bitstreams: meta-adaptive tree, per-pixel context walk, predictors (incl. the
// self-correcting weighted predictor), channel-list bookkeeping for the transforms, and host versions of the inverse
// transforms for the SMALL internal streams (LF quant, HF metadata, quant tables, per-group streams).  The frame-level
// image channels are handed to the caller with their transform list so RCT / Palette / Squeeze run on the GPU
// (jxlb200_modular_*).
//
// Behaviour follows jxlatte: J/frame/modular/ModularStream.java:66-380, ModularChannel.java:95-413, MATree.java:23-82,
// WPParams.java, TransformInfo.java, SqueezeParam.java.
#pragma once
#include <functional>

#include "entropy.hpp"

namespace jxlf {

struct Channel {
    int h = 0, w = 0;
    int vshift = 0, hshift = 0;
    int oy = 0, ox = 0;              // origin inside the frame-level channel (group streams)
    bool force_wp = false;
    bool decoded = false;
    std::vector<int32_t> px;         // h * w once allocated
    Channel() = default;
    Channel(int h_, int w_, int vs, int hs) : h(h_), w(w_), vshift(vs), hshift(hs) {}
    void allocate() { if (px.size() != (size_t)h * w) px.assign((size_t)h * w, 0); }
    int32_t &at(int y, int x) { return px[(size_t)y * w + x]; }
    int32_t at(int y, int x) const { return px[(size_t)y * w + x]; }
    bool same_shape(const Channel &o) const { return h == o.h && w == o.w && vshift == o.vshift && hshift == o.hshift; }
};

struct SqueezeStep { bool horizontal, in_place; int begin_c, num_c; };
struct Transform {
    int tr = 0;                      // 0 RCT, 1 palette, 2 squeeze
    int begin_c = 0, rct_type = 0, num_c = 0, nb_colors = 0, nb_deltas = 0, d_pred = 0;
    std::vector<SqueezeStep> sp;     // explicit list, or the default list once the stream header has been replayed
};
enum { TR_RCT = 0, TR_PALETTE = 1, TR_SQUEEZE = 2 };

struct WPParams {
    int p1 = 16, p2 = 10, p3a = 7, p3b = 7, p3c = 7, p3d = 0, p3e = 0;
    int w[4] = {13, 12, 12, 12};
    void read(BitReader &br) {
        if (br.flag()) return;
        p1 = br.bits(5); p2 = br.bits(5); p3a = br.bits(5); p3b = br.bits(5); p3c = br.bits(5); p3d = br.bits(5); p3e = br.bits(5);
        for (int i = 0; i < 4; i++) w[i] = br.bits(4);
    }
};

// Meta-adaptive tree, nodes in the breadth-first order they are coded in
struct MATree {
    struct Node {
        int property = -1;           // < 0: leaf
        int32_t value = 0;
        int left = 0, right = 0;
        int context = 0, predictor = 0;
        int32_t offset = 0, multiplier = 1;
    };
    std::vector<Node> nodes;
    EntropyStream stream;            // the symbol stream's tables (forked per modular stream)
    bool uses_wp = false;

    void read(BitReader &br) {
        EntropyStream ts(br, 6);
        int next_ctx = 0;
        size_t pending = 1;
        while (pending-- > 0) {
            if (nodes.size() > (1u << 20)) throw StreamError("MA tree too large");
            Node n;
            const int property = (int)ts.read(br, 1) - 1;
            if (property >= 0) {
                n.property = property;
                n.value = unpack_signed(ts.read(br, 0));
                n.left = (int)(nodes.size() + pending + 1);
                n.right = n.left + 1;
                pending += 2;
                if (property == 15) uses_wp = true;
            } else {
                n.context = next_ctx++;
                n.predictor = (int)ts.read(br, 2);
                if (n.predictor > 13) throw StreamError("MA tree: invalid predictor");
                n.offset = unpack_signed(ts.read(br, 3));
                const uint32_t mul_log = ts.read(br, 4);
                if (mul_log > 30) throw StreamError("MA tree: mul_log too large");
                const uint32_t mul_bits = ts.read(br, 5);
                if (mul_bits > (1u << (31 - mul_log)) - 2) throw StreamError("MA tree: mul_bits too large");
                n.multiplier = (int32_t)((mul_bits + 1) << mul_log);
                if (n.predictor == 6) uses_wp = true;
            }
            nodes.push_back(n);
        }
        ts.expect_final_state("MA tree");
        stream = EntropyStream(br, (int)(nodes.size() + 1) / 2);
    }
};

namespace detail {
inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
inline int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
inline int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
inline int32_t wrap_shl(int32_t a, int s) { return (int32_t)((uint32_t)a << (s & 31)); }
inline int32_t iabs(int32_t v) { return v < 0 ? (int32_t)(0u - (uint32_t)v) : v; }
inline int32_t clamp2(int32_t v, int32_t a, int32_t b) {
    const int32_t lo = a < b ? a : b, hi = a < b ? b : a;
    return v < lo ? lo : v > hi ? hi : v;
}
inline int32_t clamp3(int32_t v, int32_t a, int32_t b, int32_t c) {
    int32_t lo = a < b ? a : b, hi = a < b ? b : a;
    lo = lo < c ? lo : c;
    hi = hi > c ? hi : c;
    return v < lo ? lo : v > hi ? hi : v;
}
inline int floor_log1p(uint64_t x) {        // floor(log2(x + 1))
    const int c = ceil_log1p(x);
    return ((x + 1) & x) != 0 ? c - 1 : c;
}

// neighbourhood of one channel with the edge rules of ModularChannel.java:95-121
struct Nb {
    const Channel &c;
    int32_t W(int x, int y) const { return x > 0 ? c.at(y, x - 1) : y > 0 ? c.at(y - 1, x) : 0; }
    int32_t N(int x, int y) const { return y > 0 ? c.at(y - 1, x) : x > 0 ? c.at(y, x - 1) : 0; }
    int32_t NW(int x, int y) const { return x > 0 ? (y > 0 ? c.at(y - 1, x - 1) : c.at(y, x - 1)) : (y > 0 ? c.at(y - 1, x) : 0); }
    int32_t NE(int x, int y) const { return x + 1 < c.w && y > 0 ? c.at(y - 1, x + 1) : N(x, y); }
    int32_t NN(int x, int y) const { return y > 1 ? c.at(y - 2, x) : N(x, y); }
    int32_t NEE(int x, int y) const { return x + 2 < c.w && y > 0 ? c.at(y - 1, x + 2) : NE(x, y); }
    int32_t WW(int x, int y) const { return x > 1 ? c.at(y, x - 2) : W(x, y); }
};

// predictors 0..13 except 6 (ModularChannel.prediction :143-183); Java int arithmetic (wrapping, truncating division)
inline int32_t predict(const Channel &ch, int y, int x, int k, int32_t wp_pred) {
    const Nb n{ch};
    switch (k) {
    case 0: return 0;
    case 1: return n.W(x, y);
    case 2: return n.N(x, y);
    case 3: return wrap_add(n.W(x, y), n.N(x, y)) / 2;
    case 4: {
        const int32_t w = n.W(x, y), nn = n.N(x, y), nw = n.NW(x, y);
        return iabs(wrap_sub(nn, nw)) < iabs(wrap_sub(w, nw)) ? w : nn;
    }
    case 5: {
        const int32_t w = n.W(x, y), nn = n.N(x, y);
        return clamp2(wrap_sub(wrap_add(w, nn), n.NW(x, y)), nn, w);
    }
    case 6: return wrap_add(wp_pred, 3) >> 3;
    case 7: return n.NE(x, y);
    case 8: return n.NW(x, y);
    case 9: return n.WW(x, y);
    case 10: return wrap_add(n.W(x, y), n.NW(x, y)) / 2;
    case 11: return wrap_add(n.N(x, y), n.NW(x, y)) / 2;
    case 12: return wrap_add(n.N(x, y), n.NE(x, y)) / 2;
    case 13: {
        int32_t s = wrap_mul(6, n.N(x, y));
        s = wrap_sub(s, wrap_mul(2, n.NN(x, y)));
        s = wrap_add(s, wrap_mul(7, n.W(x, y)));
        s = wrap_add(s, n.WW(x, y));
        s = wrap_add(s, n.NEE(x, y));
        s = wrap_add(s, wrap_mul(3, n.NE(x, y)));
        return wrap_add(s, 8) / 16;
    }
    default: throw StreamError("invalid predictor");
    }
}

inline int32_t tendency(int32_t a, int32_t b, int32_t c) {     // ModularChannel.tendency :23-47
    if (a >= b && b >= c) {
        int32_t x = wrap_add(wrap_sub(wrap_sub(wrap_mul(4, a), wrap_mul(3, c)), b), 6) / 12;
        const int32_t d = wrap_mul(2, wrap_sub(a, b)), e = wrap_mul(2, wrap_sub(b, c));
        if (x - (x & 1) > d) x = d + 1;
        if (x + (x & 1) > e) x = e;
        return x;
    }
    if (a <= b && b <= c) {
        int32_t x = wrap_sub(wrap_sub(wrap_sub(wrap_mul(4, a), wrap_mul(3, c)), b), 6) / 12;
        const int32_t d = wrap_mul(2, wrap_sub(a, b)), e = wrap_mul(2, wrap_sub(b, c));
        if (x + (x & 1) < d) x = d - 1;
        if (x - (x & 1) < e) x = e;
        return x;
    }
    return 0;
}
}  // namespace detail

struct FrameContext {                 // what a modular stream needs to know about its frame
    int group_dim = 256;
    int modular_h = 0, modular_w = 0; // FrameHeader bounds (Frame.getModularFrameSize)
    std::vector<int> ec_dim_shift;    // per extra channel
    const MATree *global_tree = nullptr;
    int bit_depth = 8;
};

class ModularStream {
  public:
    std::vector<Channel> channels;
    std::vector<Transform> transforms;
    int nb_meta = 0;
    int stream_index = 0;
    bool transformed = false;

    ModularStream() = default;

    // frame-level stream: channel_count channels of the frame size (colour first, then extra channels with dim shifts)
    void init_global(BitReader &br, const FrameContext &fc, int index, int channel_count, int ec_start) {
        if ((int64_t)channel_count * fc.modular_h * fc.modular_w > (1ll << 31)) throw Unsupported("more than 2^31 modular samples in one frame");
        std::vector<Channel> list;
        for (int i = 0; i < channel_count; i++) {
            const int shift = i < ec_start ? 0 : fc.ec_dim_shift[i - ec_start];
            list.emplace_back(fc.modular_h, fc.modular_w, shift, shift);     // jxlatte keeps the full size (ModularStream.java:86-89)
        }
        init(br, fc, index, std::move(list));
    }
    void init(BitReader &br, const FrameContext &fc, int index, std::vector<Channel> list) {
        fc_ = &fc;
        stream_index = index;
        channels = std::move(list);
        if (channels.empty()) { dist_multiplier_ = 1; return; }
        const bool use_global_tree = br.flag();
        wp_.read(br);
        const int nb_transforms = (int)br.u32(0, 0, 1, 0, 2, 4, 18, 8);
        transforms.resize(nb_transforms);
        for (auto &t : transforms) read_transform(br, t);
        for (auto &t : transforms) replay_transform(t);
        for (auto &t : transforms) {
            if (t.tr == TR_SQUEEZE) coverage().squeeze++;
            else if (t.tr == TR_RCT) coverage().rct++;
            else { coverage().palette++; if (t.nb_deltas > 0) coverage().delta_palette++; }
        }
        (use_global_tree ? coverage().global_trees : coverage().local_trees)++;
        if (!use_global_tree) {
            local_tree_.read(br);
            tree_ = &local_tree_;
        } else {
            if (!fc.global_tree) throw StreamError("modular stream wants a global tree but the frame has none");
            tree_ = fc.global_tree;
        }
        symbols_ = tree_->stream.fork();
        dist_multiplier_ = 0;
        for (auto &c : channels) dist_multiplier_ = std::max(dist_multiplier_, c.w);
    }

    // partial = the frame-level stream inside LFGlobal: stop at the first non-meta channel larger than a group
    void decode_channels(BitReader &br, bool partial = false) {
        int coded_index = 0;
        for (size_t i = 0; i < channels.size(); i++) {
            Channel &c = channels[i];
            if (partial && (int)i >= nb_meta && (c.h > fc_->group_dim || c.w > fc_->group_dim)) break;
            if (c.w == 0 || c.h == 0) {
                c.allocate();
            } else {
                decode_channel(br, (int)i, coded_index);
                coded_index++;
            }
        }
        if (symbols_.valid()) symbols_.expect_final_state("modular stream");
        if (!partial) apply_transforms();
    }

    // Host inverse transforms (ModularStream.applyTransforms :224-380), for the small internal streams
    void apply_transforms() {
        if (transformed) return;
        transformed = true;
        for (int i = (int)transforms.size() - 1; i >= 0; i--) {
            const Transform &t = transforms[i];
            if (t.tr == TR_SQUEEZE) {
                for (int j = (int)t.sp.size() - 1; j >= 0; j--) {
                    const SqueezeStep &s = t.sp[j];
                    const int begin = s.begin_c, end = begin + s.num_c - 1;
                    const int offset = s.in_place ? end + 1 : (int)channels.size() + begin - end - 1;
                    for (int c = begin; c <= end; c++) {
                        const int r = offset + c - begin;
                        channels[c] = s.horizontal ? unsqueeze_h(channels[c], channels[r]) : unsqueeze_v(channels[c], channels[r]);
                    }
                    channels.erase(channels.begin() + offset, channels.begin() + offset + (end - begin + 1));
                }
            } else if (t.tr == TR_RCT) {
                inverse_rct(t);
            } else {
                inverse_palette(t);
            }
        }
    }

  private:
    const FrameContext *fc_ = nullptr;
    WPParams wp_;
    MATree local_tree_;
    const MATree *tree_ = nullptr;
    EntropyStream symbols_;
    int dist_multiplier_ = 1;

    static void read_transform(BitReader &br, Transform &t) {
        t.tr = (int)br.bits(2);
        if (t.tr == 3) throw StreamError("illegal modular transform");
        if (t.tr != TR_SQUEEZE) t.begin_c = (int)br.u32(0, 3, 8, 6, 72, 10, 1096, 13);
        if (t.tr == TR_RCT) t.rct_type = (int)br.u32(6, 0, 0, 2, 2, 4, 10, 6);
        if (t.tr == TR_PALETTE) {
            t.num_c = (int)br.u32(1, 0, 3, 0, 4, 0, 1, 13);
            t.nb_colors = (int)br.u32(0, 8, 256, 10, 1280, 12, 5376, 16);
            t.nb_deltas = (int)br.u32(0, 0, 1, 8, 257, 10, 1281, 16);
            t.d_pred = (int)br.bits(4);
        }
        if (t.tr == TR_SQUEEZE) {
            const int n = (int)br.u32(0, 0, 1, 4, 9, 6, 41, 8);
            t.sp.resize(n);
            for (auto &s : t.sp) {
                s.horizontal = br.flag();
                s.in_place = br.flag();
                s.begin_c = (int)br.u32(0, 3, 8, 6, 72, 10, 1096, 13);
                s.num_c = (int)br.u32(1, 0, 2, 0, 3, 0, 4, 4);
            }
        }
    }

    // channel-list effect of one forward transform (ModularStream.java:95-173)
    void replay_transform(Transform &t) {
        auto need = [&](int idx) {
            if (idx < 0 || idx >= (int)channels.size()) throw StreamError("modular transform refers to a missing channel");
        };
        if (t.tr == TR_PALETTE) {
            need(t.begin_c);
            need(t.begin_c + t.num_c - 1);
            if (t.begin_c < nb_meta) nb_meta += 2 - t.num_c;
            else nb_meta++;
            channels.erase(channels.begin() + t.begin_c + 1, channels.begin() + t.begin_c + t.num_c);
            if (t.nb_deltas > 0 && t.d_pred == 6) channels[t.begin_c].force_wp = true;
            channels.insert(channels.begin(), Channel(t.num_c, t.nb_colors, -1, -1));
        } else if (t.tr == TR_SQUEEZE) {
            if (t.sp.empty()) {
                const int first = nb_meta, count = (int)channels.size() - first;
                need(0);
                int h = channels[0].h, w = channels[0].w;         // jxlatte reads channel 0 here (:112)
                if (count > 2 && first + 1 < (int)channels.size() && channels[first + 1].h == h && channels[first + 1].w == w) {
                    t.sp.push_back({true, false, first + 1, 2});
                    t.sp.push_back({false, false, first + 1, 2});
                }
                if (h >= w && h > 8) {
                    t.sp.push_back({false, true, first, count});
                    h = (h + 1) / 2;
                }
                while (w > 8 || h > 8) {
                    if (w > 8) { t.sp.push_back({true, true, first, count}); w = (w + 1) / 2; }
                    if (h > 8) { t.sp.push_back({false, true, first, count}); h = (h + 1) / 2; }
                }
            }
            for (const SqueezeStep &s : t.sp) {
                const int begin = s.begin_c, end = begin + s.num_c - 1;
                need(begin);
                need(end);
                const int offset = s.in_place ? end + 1 : (int)channels.size();
                if (begin < nb_meta) {
                    if (!s.in_place) throw StreamError("squeeze of meta channels must be in place");
                    if (end >= nb_meta) throw StreamError("squeeze of meta channels must end in the meta channels");
                    nb_meta += s.num_c;
                }
                for (int k = begin; k <= end; k++) {
                    Channel &c = channels[k];
                    Channel r = c;
                    if (s.horizontal) {
                        const int w = c.w;
                        c.w = (w + 1) / 2;
                        c.hshift++;
                        r = c;
                        r.w = w / 2;
                    } else {
                        const int h = c.h;
                        c.h = (h + 1) / 2;
                        c.vshift++;
                        r = c;
                        r.h = h / 2;
                    }
                    r.px.clear();
                    channels.insert(channels.begin() + offset + k - begin, r);
                }
            }
        }
    }

    // ---- per-pixel decode (ModularChannel.decode :321-357) ----
    void decode_channel(BitReader &br, int list_index, int coded_index) {
        Channel &ch = channels[list_index];
        if (ch.decoded) throw std::logic_error("modular channel decoded twice");
        ch.decoded = true;
        ch.allocate();
        const bool use_wp = ch.force_wp || tree_->uses_wp;
        if (use_wp) coverage().wp_channels++;
        const int H = ch.h, W = ch.w;
        // The weighted predictor looks at its errors on the current and the previous row only (W, WW, N, NW, NE), and a pixel's
        // own prediction is used at that pixel only: two rows per error plane (row y lives at (y & 1) * W) and one scalar, not
        // the six H x W planes of the Java.  Positions of row y are read only after they were written in row y, so the stale
        // row y - 2 underneath never shows.
        std::vector<int32_t> err[5];
        int32_t wp_pred = 0;
        if (use_wp)
            for (auto &e : err) e.assign((size_t)2 * W, 0);
        const std::vector<MATree::Node> &nodes = tree_->nodes;
        // properties 0 (channel) and 1 (stream) are constant here: skip straight through those nodes
        int root = 0;
        while (nodes[root].property == 0 || nodes[root].property == 1) {
            const int32_t v = nodes[root].property == 0 ? coded_index : stream_index;
            root = v > nodes[root].value ? nodes[root].left : nodes[root].right;
        }
        int32_t sub[4] = {0, 0, 0, 0};
        const detail::Nb nb{ch};
        for (int y = 0; y < H; y++) {
            for (int x = 0; x < W; x++) {
                int32_t max_error = 0;
                if (use_wp) max_error = wp_prepare(ch, err, wp_pred, sub, x, y);
                int at = root;
                while (nodes[at].property >= 0) {
                    const int32_t v = property(ch, nb, nodes[at].property, coded_index, max_error, y, x);
                    at = v > nodes[at].value ? nodes[at].left : nodes[at].right;
                }
                const MATree::Node &leaf = nodes[at];
                const uint32_t sym = symbols_.read(br, leaf.context, dist_multiplier_);
                const int32_t diff = detail::wrap_add(detail::wrap_mul(unpack_signed(sym), leaf.multiplier), leaf.offset);
                const int32_t value = detail::wrap_add(diff, detail::predict(ch, y, x, leaf.predictor, use_wp ? wp_pred : 0));
                ch.at(y, x) = value;
                if (use_wp) {
                    const int32_t v3 = detail::wrap_shl(value, 3);
                    const size_t at_e = (size_t)(y & 1) * W + x;
                    for (int e = 0; e < 4; e++)
                        err[e][at_e] = detail::wrap_add(detail::iabs(detail::wrap_sub(sub[e], v3)), 3) >> 3;
                    err[4][at_e] = detail::wrap_sub(wp_pred, v3);
                }
            }
        }
    }

    // ModularChannel.prePredictWP :185-236
    int32_t wp_prepare(const Channel &ch, const std::vector<int32_t> (&err)[5], int32_t &pred, int32_t (&sub)[4], int x, int y) const {
        using namespace detail;
        const int W = ch.w;
        const Nb n{ch};
        auto E = [&](int e, int yy, int xx) { return err[e][(size_t)(yy & 1) * W + xx]; };     // two-row ring, see decode_channel
        auto eW = [&](int e) { return x > 0 ? E(e, y, x - 1) : 0; };
        auto eN = [&](int e) { return y > 0 ? E(e, y - 1, x) : 0; };
        auto eWW = [&](int e) { return x > 1 ? E(e, y, x - 2) : 0; };
        auto eNW = [&](int e) { return x > 0 && y > 0 ? E(e, y - 1, x - 1) : eN(e); };
        auto eNE = [&](int e) { return x + 1 < W && y > 0 ? E(e, y - 1, x + 1) : eN(e); };
        const int32_t n3 = wrap_shl(n.N(x, y), 3), nw3 = wrap_shl(n.NW(x, y), 3), ne3 = wrap_shl(n.NE(x, y), 3);
        const int32_t w3 = wrap_shl(n.W(x, y), 3), nn3 = wrap_shl(n.NN(x, y), 3);
        const int32_t tN = eN(4), tW = eW(4), tNE = eNE(4), tNW = eNW(4);
        sub[0] = wrap_sub(wrap_add(w3, ne3), n3);
        sub[1] = wrap_sub(n3, wrap_mul(wrap_add(wrap_add(tW, tN), tNE), wp_.p1) >> 5);
        sub[2] = wrap_sub(w3, wrap_mul(wrap_add(wrap_add(tW, tN), tNW), wp_.p2) >> 5);
        int32_t acc = wrap_mul(tNW, wp_.p3a);
        acc = wrap_add(acc, wrap_mul(tN, wp_.p3b));
        acc = wrap_add(acc, wrap_mul(tNE, wp_.p3c));
        acc = wrap_add(acc, wrap_mul(wrap_sub(nn3, n3), wp_.p3d));
        acc = wrap_add(acc, wrap_mul(wrap_sub(nw3, w3), wp_.p3e));
        sub[3] = wrap_sub(n3, acc >> 5);
        int32_t weight[4];
        int32_t wsum = 0;
        for (int e = 0; e < 4; e++) {
            int32_t s32 = wrap_add(wrap_add(wrap_add(wrap_add(eN(e), eW(e)), eNW(e)), eWW(e)), eNE(e));
            int64_t s = s32;
            if (x + 1 == W) s += eW(e);
            const uint64_t es = (uint64_t)s & 0xffffffffull;
            int shift = floor_log1p(es) - 5;
            if (shift < 0) shift = 0;
            const uint32_t prod = (uint32_t)wp_.w[e] * (uint32_t)((1 << 24) / (int)((es >> shift) + 1));
            weight[e] = (int32_t)(4 + (prod >> shift));
            wsum = wrap_add(wsum, weight[e]);
        }
        const int log_weight = floor_log1p((uint64_t)(int64_t)(wsum - 1)) - 4;
        wsum = 0;
        for (int e = 0; e < 4; e++) {
            weight[e] = (int32_t)((uint32_t)weight[e] >> (log_weight & 31));
            wsum += weight[e];
        }
        int64_t s = (int64_t)((uint32_t)wsum >> 1) - 1;
        for (int e = 0; e < 4; e++) s += wrap_mul(sub[e], weight[e]);
        int32_t p = (int32_t)((s * ((1 << 24) / wsum)) >> 24);
        if (((tN ^ tW) | (tN ^ tNW)) <= 0) p = clamp3(p, w3, n3, ne3);
        pred = p;
        int32_t m = tW;
        if (iabs(tN) > iabs(m)) m = tN;
        if (iabs(tNW) > iabs(m)) m = tNW;
        if (iabs(tNE) > iabs(m)) m = tNE;
        return m;
    }

    // ModularChannel.propertyExpand :238-308
    int32_t property(const Channel &ch, const detail::Nb &n, int k, int coded_index, int32_t max_error, int y, int x) const {
        using namespace detail;
        switch (k) {
        case 0: return coded_index;
        case 1: return stream_index;
        case 2: return y;
        case 3: return x;
        case 4: return iabs(n.N(x, y));
        case 5: return iabs(n.W(x, y));
        case 6: return n.N(x, y);
        case 7: return n.W(x, y);
        case 8:
            return x > 0 ? wrap_sub(n.W(x, y), wrap_sub(wrap_add(n.W(x - 1, y), n.N(x - 1, y)), n.NW(x - 1, y))) : n.W(x, y);
        case 9: return wrap_sub(wrap_add(n.W(x, y), n.N(x, y)), n.NW(x, y));
        case 10: return wrap_sub(n.W(x, y), n.NW(x, y));
        case 11: return wrap_sub(n.NW(x, y), n.N(x, y));
        case 12: return wrap_sub(n.N(x, y), n.NE(x, y));
        case 13: return wrap_sub(n.N(x, y), n.NN(x, y));
        case 14: return wrap_sub(n.W(x, y), n.WW(x, y));
        case 15: return max_error;
        default: break;
        }
        if (k - 16 >= 4 * coded_index) return 0;
        int k2 = 16;
        for (int j = coded_index - 1; j >= 0; j--) {      // jxlatte indexes the channel list with the coded index (:280-282)
            const Channel &o = channels[j];
            if (!ch.same_shape(o)) continue;
            if (k2 + 4 <= k) { k2 += 4; continue; }
            const int32_t rC = o.at(y, x);
            if (k2++ == k) return iabs(rC);
            if (k2++ == k) return rC;
            const int32_t rW = x > 0 ? o.at(y, x - 1) : 0;
            const int32_t rN = y > 0 ? o.at(y - 1, x) : rW;
            const int32_t rNW = x > 0 && y > 0 ? o.at(y - 1, x - 1) : rW;
            const int32_t rG = wrap_sub(rC, clamp2(wrap_sub(wrap_add(rW, rN), rNW), rN, rW));
            if (k2++ == k) return iabs(rG);
            if (k2++ == k) return rG;
        }
        return 0;
    }

    // ---- host inverse transforms ----
    static Channel unsqueeze_h(const Channel &avg, const Channel &res) {       // ModularChannel.java:361-387
        using namespace detail;
        if ((avg.w != res.w && avg.w != res.w + 1) || avg.h != res.h) throw StreamError("corrupted squeeze transform");
        Channel out(avg.h, avg.w + res.w, avg.vshift, avg.hshift - 1);
        out.oy = avg.oy; out.ox = avg.ox;
        out.allocate();
        for (int y = 0; y < out.h; y++) {
            for (int x = 0; x < res.w; x++) {
                const int32_t a = avg.at(y, x), r = res.at(y, x);
                const int32_t next = x + 1 < avg.w ? avg.at(y, x + 1) : a;
                const int32_t left = x > 0 ? out.at(y, 2 * x - 1) : a;
                const int32_t diff = wrap_add(r, tendency(left, a, next));
                const int32_t first = wrap_add(a, diff / 2);
                out.at(y, 2 * x) = first;
                out.at(y, 2 * x + 1) = wrap_sub(first, diff);
            }
            if (avg.w > res.w) out.at(y, 2 * res.w) = avg.at(y, res.w);
        }
        out.decoded = true;
        return out;
    }
    static Channel unsqueeze_v(const Channel &avg, const Channel &res) {       // :389-413
        using namespace detail;
        if ((avg.h != res.h && avg.h != res.h + 1) || avg.w != res.w) throw StreamError("corrupted squeeze transform");
        Channel out(avg.h + res.h, avg.w, avg.vshift - 1, avg.hshift);
        out.oy = avg.oy; out.ox = avg.ox;
        out.allocate();
        for (int y = 0; y < res.h; y++)
            for (int x = 0; x < out.w; x++) {
                const int32_t a = avg.at(y, x), r = res.at(y, x);
                const int32_t next = y + 1 < avg.h ? avg.at(y + 1, x) : a;
                const int32_t top = y > 0 ? out.at(2 * y - 1, x) : a;
                const int32_t diff = wrap_add(r, tendency(top, a, next));
                const int32_t first = wrap_add(a, diff / 2);
                out.at(2 * y, x) = first;
                out.at(2 * y + 1, x) = wrap_sub(first, diff);
            }
        if (avg.h > res.h)
            for (int x = 0; x < out.w; x++) out.at(2 * res.h, x) = avg.at(res.h, x);
        out.decoded = true;
        return out;
    }

    void inverse_rct(const Transform &t) {                                    // ModularStream.java:255-326
        using namespace detail;
        static const int kPerm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};
        const int perm = t.rct_type / 7, type = t.rct_type % 7, b = t.begin_c;
        if (b + 2 >= (int)channels.size() || perm > 5) throw StreamError("RCT refers to missing channels");
        Channel v[3] = {std::move(channels[b]), std::move(channels[b + 1]), std::move(channels[b + 2])};
        if (!v[0].same_shape(v[1]) || !v[0].same_shape(v[2])) throw StreamError("RCT over channels of different shapes");
        const size_t n = v[0].px.size();
        int32_t *p0 = v[0].px.data(), *p1 = v[1].px.data(), *p2 = v[2].px.data();
        for (size_t i = 0; i < n; i++) {
            const int32_t a = p0[i], bb = p1[i], c = p2[i];
            switch (type) {
            case 1: p2[i] = wrap_add(c, a); break;
            case 2: p1[i] = wrap_add(bb, a); break;
            case 3: p2[i] = wrap_add(c, a); p1[i] = wrap_add(bb, a); break;
            case 4: p1[i] = wrap_add(bb, wrap_add(a, c) >> 1); break;
            case 5: { const int32_t ac = wrap_add(a, c); p1[i] = wrap_add(bb, wrap_add(a, ac) >> 1); p2[i] = ac; break; }
            case 6: {
                const int32_t tmp = wrap_sub(a, c >> 1), f = wrap_sub(tmp, bb >> 1);
                p0[i] = wrap_add(f, bb); p1[i] = wrap_add(c, tmp); p2[i] = f;
                break;
            }
            default: break;
            }
        }
        for (int j = 0; j < 3; j++) channels[b + kPerm[perm][j]] = std::move(v[j]);
    }

    void inverse_palette(const Transform &t);                                  // defined below (needs the delta table)
};

namespace detail {
inline const int16_t kDeltaPalette[72][3] = {
    {0, 0, 0}, {4, 4, 4}, {11, 0, 0}, {0, 0, -13}, {0, -12, 0}, {-10, -10, -10}, {-18, -18, -18}, {-27, -27, -27},
    {-18, -18, 0}, {0, 0, -32}, {-32, 0, 0}, {-37, -37, -37}, {0, -32, -32}, {24, 24, 45}, {50, 50, 50}, {-45, -24, -24},
    {-24, -45, -45}, {0, -24, -24}, {-34, -34, 0}, {-24, 0, -24}, {-45, -45, -24}, {64, 64, 64}, {-32, 0, -32}, {0, -32, 0},
    {-32, 0, 32}, {-24, -45, -24}, {45, 24, 45}, {24, -24, -45}, {-45, -24, 24}, {80, 80, 80}, {64, 0, 0}, {0, 0, -64},
    {0, -64, -64}, {-24, -24, 45}, {96, 96, 96}, {64, 64, 0}, {45, -24, -24}, {34, -34, 0}, {112, 112, 112}, {24, -45, -45},
    {45, 45, -24}, {0, -32, 32}, {24, -24, 45}, {0, 96, 96}, {45, -24, 24}, {24, -45, -24}, {-24, -45, 24}, {0, -64, 0},
    {96, 0, 0}, {128, 128, 128}, {64, 0, 64}, {144, 144, 144}, {96, 96, 0}, {-36, -36, 36}, {45, -24, -45}, {45, -45, -24},
    {0, 0, -96}, {0, 128, 128}, {0, 96, 0}, {45, 24, -45}, {-128, 0, 0}, {24, -45, 24}, {-45, 24, -45}, {64, 0, -64},
    {64, -64, -64}, {96, 0, 96}, {45, -45, 24}, {24, 45, -45}, {64, 64, -64}, {128, 128, 0}, {0, 0, -128}, {-24, 45, -45}};
}

inline void ModularStream::inverse_palette(const Transform &t) {               // ModularStream.java:327-378
    using namespace detail;
    const int first = t.begin_c + 1, bd = fc_->bit_depth;
    if (first >= (int)channels.size() || channels[0].h < t.num_c || channels[0].w < t.nb_colors) throw StreamError("palette refers to missing channels");
    const Channel pal = channels[0];
    const Channel idx = channels[first];
    std::vector<Channel> outs;
    for (int c = 0; c < t.num_c; c++) {
        Channel o = idx;
        for (int y = 0; y < idx.h; y++)
            for (int x = 0; x < idx.w; x++) {
                int32_t index = idx.at(y, x);
                const bool is_delta = index < t.nb_deltas;
                int32_t value;
                if (index >= 0 && index < t.nb_colors) {
                    value = pal.at(c, index);
                } else if (index >= t.nb_colors) {
                    index -= t.nb_colors;
                    const int32_t maxv = (int32_t)((1u << (bd & 31)) - 1u);
                    if (index < 64) {
                        value = wrap_add(wrap_mul((index >> ((2 * c) & 31)) % 4, maxv) / 4, (int32_t)(1u << (std::max(0, bd - 3) & 31)));
                    } else {
                        index -= 64;
                        for (int k = 0; k < c; k++) index /= 5;
                        value = wrap_mul(index % 5, maxv) / 4;
                    }
                } else if (c < 3) {
                    index = (-index - 1) % 143;
                    value = kDeltaPalette[(index + 1) >> 1][c];
                    if ((index & 1) == 0) value = -value;
                    if (bd > 8) value = wrap_shl(value, std::min(bd, 24) - 8);
                } else {
                    value = 0;
                }
                o.at(y, x) = value;
                if (is_delta) {
                    if (t.d_pred == 6) throw Unsupported("delta palette with the weighted predictor (the reference has no predictor state there)");
                    o.at(y, x) = wrap_add(value, predict(o, y, x, t.d_pred, 0));
                }
            }
        o.decoded = true;
        outs.push_back(std::move(o));
    }
    channels.erase(channels.begin() + first);
    channels.insert(channels.begin() + first, outs.begin(), outs.end());
    channels.erase(channels.begin());
}

}  // namespace jxlf
