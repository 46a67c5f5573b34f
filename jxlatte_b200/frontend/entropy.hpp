// entropy.hpp -- JPEG XL entropy-coded streams: clustered contexts, ANS (alias tables) or Brotli-style prefix codes,
// hybrid integers, optional LZ77.  Behaviour follows jxlatte's entropy/ package (J/entropy/EntropyStream.java:88-246,
// ANSSymbolDistribution.java:44-215, PrefixSymbolDistribution.java:27-203, HybridIntegerConfig.java:21-35).
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <cstring>
#include <memory>
#include <new>
#include <sys/mman.h>

#include "bits.hpp"

namespace jxlf {

// which optional coding tools a stream actually used (reported in jxlf_describe: tells what the samples do and do not exercise)
struct Coverage {
    typedef std::atomic<long> N;      // group sections are decoded on several threads
    N streams{0}, lz77_streams{0}, prefix_streams{0}, ans_streams{0}, mtf_context_maps{0}, lz77_copies{0};
    N permuted_toc{0}, coded_orders{0}, multi_pass_frames{0}, custom_block_ctx{0}, wp_channels{0}, global_trees{0}, local_trees{0};
    N squeeze{0}, palette{0}, delta_palette{0}, rct{0}, raw_quant{0}, custom_quant{0}, lf_smoothing{0};
    void reset() {
        for (N *p : {&streams, &lz77_streams, &prefix_streams, &ans_streams, &mtf_context_maps, &lz77_copies, &permuted_toc, &coded_orders,
                     &multi_pass_frames, &custom_block_ctx, &wp_channels, &global_trees, &local_trees, &squeeze, &palette, &delta_palette,
                     &rct, &raw_quant, &custom_quant, &lf_smoothing}) *p = 0;
    }
};
// Diagnostic counters only: process-wide and reset by every jxlf_decode, so two decodes running at once (ctypes releases the
// GIL) report each other's counts; no parsed state depends on them and every field is atomic.
inline Coverage &coverage() { static Coverage c; return c; }

struct HybridConfig {
    int split_exp = 0, msb = 0, lsb = 0;
    void read(BitReader &br, int log_alphabet) {
        split_exp = (int)br.bits(ceil_log1p(log_alphabet));
        msb = lsb = 0;
        if (split_exp == log_alphabet) return;
        msb = (int)br.bits(ceil_log1p(split_exp));
        if (msb > split_exp) throw StreamError("hybrid integer: msb_in_token too large");
        lsb = (int)br.bits(ceil_log1p(split_exp - msb));
        if (msb + lsb > split_exp) throw StreamError("hybrid integer: msb + lsb too large");
    }
    uint32_t decode(BitReader &br, uint32_t token) const {
        const uint32_t split = 1u << split_exp;
        if (token < split) return token;
        const int n = split_exp - lsb - msb + (int)((token - split) >> (msb + lsb));
        if (n > 32) throw StreamError("hybrid integer: too many extra bits");
        const uint32_t low = token & ((1u << lsb) - 1);
        uint32_t hi = ((token >> lsb) & ((1u << msb) - 1)) | (1u << msb);
        const uint64_t body = n == 32 ? (((uint64_t)hi << 32) | br.bits(32)) : (((uint64_t)hi << n) | br.bits(n));
        return (uint32_t)((body << lsb) | low);
    }
};

// The LZ77 window: 1 Mi symbols, zero-initialised like the Java's `new int[1 << 20]`.  Anonymous pages straight from mmap
// are zero without being touched, so a stream pays only for the part of the window it writes (a vector's zero-fill of
// the 4 MB cost 0.8 ms per LZ77 stream: a fifth of the front-end time of a small photo).
class LzWindow {
  public:
    LzWindow() = default;
    LzWindow(const LzWindow &o) { if (o.p_) { reset(); std::memcpy(p_, o.p_, kBytes); } }
    LzWindow(LzWindow &&o) noexcept : p_(o.p_) { o.p_ = nullptr; }
    LzWindow &operator=(LzWindow o) noexcept { std::swap(p_, o.p_); return *this; }
    ~LzWindow() { release(); }
    void reset() {
        release();
        void *m = mmap(nullptr, kBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m == MAP_FAILED) throw std::bad_alloc();
        p_ = static_cast<uint32_t *>(m);
    }
    uint32_t &operator[](uint32_t i) { return p_[i]; }

  private:
    static constexpr size_t kBytes = (size_t)4 << 20;
    void release() { if (p_) munmap(p_, kBytes); p_ = nullptr; }
    uint32_t *p_ = nullptr;
};

// One clustered distribution: either an ANS alias table over a 12-bit state slice or a prefix-code lookup table.
struct Distribution {
    HybridConfig cfg;
    // ANS
    std::vector<uint16_t> freq, cutoff, alias;
    std::vector<int32_t> offset;
    int log_bucket = 0;
    // prefix
    std::vector<uint32_t> lut;       // (symbol << 8) | length, indexed by the next `lut_bits` stream bits
    int lut_bits = 0;
    uint32_t only_symbol = 0;        // when the code has a single symbol (zero-length code)
};

class PrefixBuilder {
  public:
    // Canonical code from lengths given in assignment order (entries with length 0 are skipped): shorter codes first is
    // the caller's responsibility.  Table is indexed by LSB-first stream bits.
    static void build(Distribution &d, int lut_bits, const std::vector<int> &lens, const std::vector<uint32_t> &syms) {
        d.lut_bits = lut_bits;
        d.lut.assign((size_t)1 << lut_bits, 0);
        uint64_t code = 0;          // left-aligned in 32 bits
        int used = 0;
        for (size_t i = 0; i < lens.size(); i++) {
            const int len = lens[i];
            if (len <= 0) continue;
            if (len > lut_bits) throw StreamError("prefix code longer than its table");
            uint32_t c = (uint32_t)(code >> (32 - len));          // len-bit canonical code, MSB first
            uint32_t rev = 0;
            for (int b = 0; b < len; b++) rev |= ((c >> b) & 1u) << (len - 1 - b);
            for (uint32_t idx = rev; idx < d.lut.size(); idx += 1u << len) d.lut[idx] = (syms[i] << 8) | (uint32_t)len;
            code += 1ull << (32 - len);
            if (code > (1ull << 32)) throw StreamError("over-subscribed prefix code");
            used++;
        }
        if (used == 0) throw StreamError("empty prefix code");
        if (code != (1ull << 32)) throw StreamError("incomplete prefix code");
    }
};

class EntropyStream {
  public:
    EntropyStream() = default;
    EntropyStream(BitReader &br, int num_dists) { read_header(br, num_dists, false); }

    // A second reader over the same tables with fresh ANS / LZ77 state (jxlatte's copy constructor, :108-119)
    EntropyStream fork() const {
        EntropyStream e;
        e.shared_ = shared_;
        if (shared_ && shared_->lz77) e.window_.reset();
        return e;
    }
    bool valid() const { return (bool)shared_; }

    uint32_t read(BitReader &br, int ctx, int dist_multiplier = 0) {
        Shared &s = *shared_;
        if (to_copy_ > 0) return copy_one();
        if (ctx < 0 || (size_t)ctx >= s.cluster.size()) throw std::logic_error("entropy context out of range");
        const Distribution &d = s.dists[s.cluster[ctx]];
        uint32_t token = symbol(br, d);
        if (s.lz77 && token >= s.lz_min_symbol) {
            const Distribution &ld = s.dists[s.cluster.back()];
            to_copy_ = s.lz_min_length + s.lz_len_cfg.decode(br, token - s.lz_min_symbol);
            const uint32_t dtoken = symbol(br, ld);
            int64_t distance = ld.cfg.decode(br, dtoken);
            if (dist_multiplier == 0) {
                distance++;
            } else if (distance < 120) {
                const int8_t *sd = kSpecial[distance];
                distance = sd[0] + (int64_t)dist_multiplier * sd[1];
                if (distance < 1) distance = 1;
            } else {
                distance -= 119;
            }
            distance = std::min<int64_t>(distance, 1 << 20);
            distance = std::min<int64_t>(distance, decoded_);
            copy_pos_ = decoded_ - (uint32_t)distance;
            coverage().lz77_copies++;
            if (to_copy_ == 0) throw StreamError("LZ77 copy of zero length");
            return copy_one();
        }
        const uint32_t v = d.cfg.decode(br, token);
        if (s.lz77) window_[decoded_++ & 0xfffff] = v;
        return v;
    }

    // ANS streams must end in the initial state (EntropyStream.java:133-141)
    bool final_state_ok() const { return !has_state_ || state_ == 0x130000u; }
    void expect_final_state(const char *what) const {
        if (!final_state_ok()) throw StreamError(std::string("ANS final state check failed: ") + what);
    }

    // Context map (EntropyStream.java:51-106).  Returns the number of clusters.
    static int read_cluster_map(BitReader &br, std::vector<uint8_t> &map, int num_dists, int max_clusters) {
        map.assign(num_dists, 0);
        if (num_dists > 1) {
            if (br.flag()) {
                const int nbits = (int)br.bits(2);
                for (auto &m : map) m = (uint8_t)br.bits(nbits);
            } else {
                const bool mtf = br.flag();
                if (mtf) coverage().mtf_context_maps++;
                EntropyStream nested;
                nested.read_header(br, 1, num_dists <= 2);
                std::vector<uint32_t> raw(num_dists);
                for (auto &r : raw) r = nested.read(br, 0);
                nested.expect_final_state("context map");
                if (mtf) {
                    uint8_t order[256];
                    for (int i = 0; i < 256; i++) order[i] = (uint8_t)i;
                    for (auto &r : raw) {
                        if (r > 255) throw StreamError("context map: MTF index above 255");
                        const uint8_t v = order[r];
                        std::memmove(order + 1, order, r);
                        order[0] = v;
                        r = v;
                    }
                }
                for (int i = 0; i < num_dists; i++) {
                    if (raw[i] > 255) throw StreamError("context map: cluster index above 255");
                    map[i] = (uint8_t)raw[i];
                }
            }
        }
        int clusters = 0;
        for (auto m : map) clusters = std::max(clusters, m + 1);
        if (clusters > max_clusters) throw StreamError("context map: too many clusters");
        return clusters;
    }

  private:
    struct Shared {
        bool lz77 = false;
        uint32_t lz_min_symbol = 0, lz_min_length = 0;
        HybridConfig lz_len_cfg;
        std::vector<uint8_t> cluster;
        std::vector<Distribution> dists;
        bool prefix = false;
        int log_alphabet = 0;
    };
    std::shared_ptr<Shared> shared_;
    LzWindow window_;
    uint32_t to_copy_ = 0, copy_pos_ = 0, decoded_ = 0;
    uint32_t state_ = 0;
    bool has_state_ = false;

    static const int8_t kSpecial[120][2];

    uint32_t copy_one() {
        const uint32_t v = window_[copy_pos_++ & 0xfffff];
        to_copy_--;
        window_[decoded_++ & 0xfffff] = v;
        return v;
    }

    uint32_t symbol(BitReader &br, const Distribution &d) {
        if (shared_->prefix) {
            if (d.lut_bits == 0) return d.only_symbol;
            const uint32_t e = d.lut[br.peek(d.lut_bits)];
            if ((e & 0xff) == 0) throw StreamError("invalid prefix code word");
            br.drop((int)(e & 0xff));
            return e >> 8;
        }
        if (!has_state_) {
            state_ = br.bits(32);
            has_state_ = true;
        }
        const uint32_t idx = state_ & 0xfff;
        const uint32_t i = idx >> d.log_bucket, pos = idx & ((1u << d.log_bucket) - 1);
        uint32_t sym, off;
        if (pos >= d.cutoff[i]) {
            sym = d.alias[i];
            off = (uint32_t)(d.offset[i] + (int32_t)pos);
        } else {
            sym = i;
            off = pos;
        }
        state_ = (uint32_t)d.freq[sym] * (state_ >> 12) + off;
        if (state_ < (1u << 16)) state_ = (state_ << 16) | br.bits(16);
        return sym;
    }

    void read_header(BitReader &br, int num_dists, bool forbid_lz77) {
        if (num_dists <= 0) throw std::logic_error("entropy stream needs at least one context");
        auto s = std::make_shared<Shared>();
        s->lz77 = br.flag();
        coverage().streams++;
        if (s->lz77) {
            coverage().lz77_streams++;
            if (forbid_lz77) throw StreamError("nested entropy stream may not use LZ77");
            s->lz_min_symbol = br.u32(224, 0, 512, 0, 4096, 0, 8, 15);
            s->lz_min_length = br.u32(3, 0, 4, 0, 5, 2, 9, 8);
            num_dists++;
            s->lz_len_cfg.read(br, 8);
            window_.reset();
        }
        const int clusters = read_cluster_map(br, s->cluster, num_dists, num_dists);
        s->dists.resize(clusters);
        s->prefix = br.flag();
        (s->prefix ? coverage().prefix_streams : coverage().ans_streams)++;
        s->log_alphabet = s->prefix ? 15 : 5 + (int)br.bits(2);
        for (auto &d : s->dists) d.cfg.read(br, s->log_alphabet);
        if (s->prefix) {
            std::vector<int> sizes(clusters, 1);
            for (auto &n : sizes)
                if (br.flag()) {
                    const int k = (int)br.bits(4);
                    n = 1 + (1 << k) + (int)br.bits(k);
                }
            for (int i = 0; i < clusters; i++) read_prefix_code(br, s->dists[i], sizes[i]);
        } else {
            for (auto &d : s->dists) read_ans_distribution(br, d, s->log_alphabet);
        }
        shared_ = std::move(s);
    }

    // ---- ANS histogram + alias table (ANSSymbolDistribution.java:44-215) ----
    static void read_ans_distribution(BitReader &br, Distribution &d, int log_alphabet) {
        const int table = 1 << log_alphabet;
        std::vector<int> f;
        int uniq = -1;
        if (br.flag()) {
            if (br.flag()) {                               // two symbols
                const int a = (int)br.u8(), b = (int)br.u8();
                if (a == b) throw StreamError("ANS: dual-peak symbols coincide");
                f.assign(1 + std::max(a, b), 0);
                if ((int)f.size() > table) throw StreamError("ANS: alphabet too large");
                f[a] = (int)br.bits(12);
                f[b] = 4096 - f[a];
                if (f[a] == 0) uniq = b;
            } else {                                       // one symbol
                const int a = (int)br.u8();
                f.assign(1 + a, 0);
                f[a] = 4096;
                uniq = a;
            }
        } else if (br.flag()) {                            // flat
            const int n = 1 + (int)br.u8();
            if (n > table) throw StreamError("ANS: alphabet too large");
            if (n == 1) uniq = 0;
            f.assign(n, 4096 / n);
            for (int i = 0; i < 4096 % n; i++) f[i]++;
        } else {
            int len = 0;
            while (len < 3 && br.flag()) len++;
            const int shift = (int)((br.bits(len) | (1u << len)) - 1);
            if (shift > 13) throw StreamError("ANS: shift above 13");
            const int n = 3 + (int)br.u8();
            if (n > table) throw StreamError("ANS: alphabet too large");
            f.assign(n, 0);
            std::vector<int> logc(n, 0), same(n, 0);
            int omit_log = -1, omit_pos = -1;
            for (int i = 0; i < n; i++) {
                logc[i] = log_count(br);
                if (logc[i] == 13) {
                    const int rle = (int)br.u8();
                    same[i] = rle + 5;
                    i += rle + 3;
                    continue;
                }
                if (logc[i] > omit_log) { omit_log = logc[i]; omit_pos = i; }
            }
            if (omit_pos < 0 || (omit_pos + 1 < n && logc[omit_pos + 1] == 13)) throw StreamError("ANS: invalid omitted position");
            int total = 0, run = 0, prev = 0;
            for (int i = 0; i < n; i++) {
                if (same[i]) {
                    run = same[i] - 1;
                    prev = i > 0 ? f[i - 1] : 0;
                }
                if (run) {
                    f[i] = prev;
                    run--;
                } else {
                    if (i == omit_pos || logc[i] == 0) continue;
                    if (logc[i] == 1) {
                        f[i] = 1;
                    } else {
                        int bc = shift - ((12 - logc[i] + 1) >> 1);
                        bc = std::max(0, std::min(bc, logc[i] - 1));
                        f[i] = (1 << (logc[i] - 1)) + ((int)br.bits(bc) << (logc[i] - 1 - bc));
                    }
                }
                total += f[i];
            }
            f[omit_pos] = 4096 - total;
            if (f[omit_pos] < 0) throw StreamError("ANS: frequencies exceed 4096");
        }
        // alias table (Vose with stacks, in the order the format prescribes)
        d.log_bucket = 12 - log_alphabet;
        const int bucket = 1 << d.log_bucket;
        d.freq.assign(std::max((size_t)table, f.size()), 0);      // a one-symbol histogram may name a symbol beyond the table (the Java never checks it)
        for (size_t i = 0; i < f.size(); i++) d.freq[i] = (uint16_t)f[i];
        d.cutoff.assign(table, 0);
        d.alias.assign(table, 0);
        d.offset.assign(table, 0);
        if (uniq >= 0) {
            for (int i = 0; i < table; i++) {
                d.alias[i] = (uint16_t)uniq;
                d.offset[i] = i * bucket;
            }
            return;
        }
        std::vector<int> cut(table, 0), over, under;
        for (int i = 0; i < (int)f.size(); i++) {
            cut[i] = f[i];
            d.alias[i] = (uint16_t)i;
            if (cut[i] > bucket) over.push_back(i);
            else if (cut[i] < bucket) under.push_back(i);
        }
        for (int i = (int)f.size(); i < table; i++) under.push_back(i);
        while (!over.empty()) {
            if (under.empty()) throw StreamError("ANS: inconsistent histogram");
            const int u = under.back(); under.pop_back();
            const int o = over.back(); over.pop_back();
            const int by = bucket - cut[u];
            cut[o] -= by;
            d.alias[u] = (uint16_t)o;
            d.offset[u] = cut[o];
            if (cut[o] < bucket) under.push_back(o);
            else if (cut[o] > bucket) over.push_back(o);
        }
        for (int i = 0; i < table; i++) {
            if (cut[i] == bucket) {
                d.alias[i] = (uint16_t)i;
                d.offset[i] = 0;
                cut[i] = 0;
            } else {
                d.offset[i] -= cut[i];
            }
            d.cutoff[i] = (uint16_t)cut[i];
        }
    }

    // the fixed 7-bit prefix code of the log-counts (ANSSymbolDistribution.java:14-32)
    static int log_count(BitReader &br) {
        static const uint8_t kLen[14] = {5, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 6, 7, 7};
        static const uint8_t kCode[14] = {0x11, 0x0b, 0x0f, 0x03, 0x09, 0x07, 0x04, 0x02, 0x05, 0x06, 0x00, 0x21, 0x01, 0x41};
        const uint32_t w = br.peek(7);
        for (int s = 0; s < 14; s++)
            if ((w & ((1u << kLen[s]) - 1)) == kCode[s]) {
                br.drop(kLen[s]);
                return s;
            }
        throw StreamError("ANS: bad log-count code");
    }

    // ---- Brotli-style prefix code (PrefixSymbolDistribution.java:27-203) ----
    static void read_prefix_code(BitReader &br, Distribution &d, int alphabet) {
        d.lut_bits = 0;
        d.only_symbol = 0;
        if (alphabet == 1) return;
        const int hskip = (int)br.bits(2);
        if (hskip == 1) {
            const int sym_bits = ceil_log1p((uint64_t)alphabet - 1);
            const int n = 1 + (int)br.bits(2);
            uint32_t sym[4] = {0, 0, 0, 0};
            for (int i = 0; i < n; i++) sym[i] = br.bits(sym_bits);
            const bool tree = n == 4 ? br.flag() : false;
            std::vector<int> lens;
            if (n == 1) { d.only_symbol = sym[0]; return; }
            if (n == 2) { lens = {1, 1}; if (sym[0] > sym[1]) std::swap(sym[0], sym[1]); }
            else if (n == 3) { lens = {1, 2, 2}; if (sym[1] > sym[2]) std::swap(sym[1], sym[2]); }
            else if (tree) { lens = {1, 2, 3, 3}; if (sym[2] > sym[3]) std::swap(sym[2], sym[3]); }
            else { lens = {2, 2, 2, 2}; std::sort(sym, sym + 4); }
            PrefixBuilder::build(d, lens.back(), lens, std::vector<uint32_t>(sym, sym + n));
            return;
        }
        // code lengths of the code-length alphabet
        static const uint8_t kOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
        static const uint8_t kL0Sym[16] = {0, 4, 3, 2, 0, 4, 3, 1, 0, 4, 3, 2, 0, 4, 3, 5};
        static const uint8_t kL0Len[16] = {2, 2, 2, 3, 2, 2, 2, 4, 2, 2, 2, 3, 2, 2, 2, 4};
        int l1[18] = {0};
        int space = 0, nonzero = 0;
        for (int i = hskip; i < 18; i++) {
            const uint32_t w = br.peek(4);
            br.drop(kL0Len[w]);
            const int len = kL0Sym[w];
            l1[kOrder[i]] = len;
            if (len) {
                space += 32 >> len;
                nonzero++;
            }
            if (space >= 32) break;
        }
        if ((space != 32 && nonzero >= 2) || nonzero < 1) throw StreamError("prefix code: bad code-length code");
        Distribution cl;
        if (nonzero == 1) {
            for (int i = 0; i < 18; i++) if (l1[i]) cl.only_symbol = (uint32_t)i;
        } else {
            std::vector<int> lens;
            std::vector<uint32_t> syms;
            for (int len = 1; len <= 5; len++)
                for (int i = 0; i < 18; i++)
                    if (l1[i] == len) { lens.push_back(len); syms.push_back((uint32_t)i); }
            PrefixBuilder::build(cl, 5, lens, syms);
        }
        std::vector<int> l2(alphabet, 0);
        int total = 0, prev = 8, rep_nz = 0, rep_z = 0, count_nz = 0;
        for (int i = 0; i < alphabet; i++) {
            uint32_t code;
            if (cl.lut_bits == 0) {
                code = cl.only_symbol;
            } else {
                const uint32_t e = cl.lut[br.peek(5)];
                if ((e & 0xff) == 0) throw StreamError("prefix code: bad code-length word");
                br.drop((int)(e & 0xff));
                code = e >> 8;
            }
            if (code == 16) {
                int extra = 3 + (int)br.bits(2);
                if (rep_nz > 0) extra = 4 * (rep_nz - 2) - rep_nz + extra;
                if (extra < 0 || i + extra > alphabet) throw StreamError("prefix code: repeat runs past the alphabet");
                for (int j = 0; j < extra; j++) l2[i + j] = prev;
                total += (32768 >> prev) * extra;
                count_nz += extra;
                i += extra - 1;
                rep_nz += extra;
                rep_z = 0;
            } else if (code == 17) {
                int extra = 3 + (int)br.bits(3);
                if (rep_z > 0) extra = 8 * (rep_z - 2) - rep_z + extra;
                if (extra < 0 || i + extra > alphabet) throw StreamError("prefix code: zero run past the alphabet");
                i += extra - 1;
                rep_nz = 0;
                rep_z += extra;
            } else {
                l2[i] = (int)code;
                rep_nz = rep_z = 0;
                if (code) {
                    total += 32768 >> code;
                    prev = (int)code;
                    count_nz++;
                }
            }
            if (total >= 32768) break;
        }
        if (total != 32768 && count_nz != 1) throw StreamError("prefix code: lengths do not fill the code space");
        if (count_nz == 1) {
            for (int i = 0; i < alphabet; i++) if (l2[i]) d.only_symbol = (uint32_t)i;
            return;
        }
        std::vector<int> lens;
        std::vector<uint32_t> syms;
        for (int len = 1; len <= 15; len++)
            for (int i = 0; i < alphabet; i++)
                if (l2[i] == len) { lens.push_back(len); syms.push_back((uint32_t)i); }
        PrefixBuilder::build(d, lens.back(), lens, syms);    // table as wide as the longest code (was 15 bits = 128 KB each)
    }
};

inline const int8_t EntropyStream::kSpecial[120][2] = {
    {0, 1}, {1, 0}, {1, 1}, {-1, 1}, {0, 2}, {2, 0}, {1, 2}, {-1, 2}, {2, 1}, {-2, 1}, {2, 2}, {-2, 2}, {0, 3}, {3, 0}, {1, 3},
    {-1, 3}, {3, 1}, {-3, 1}, {2, 3}, {-2, 3}, {3, 2}, {-3, 2}, {0, 4}, {4, 0}, {1, 4}, {-1, 4}, {4, 1}, {-4, 1}, {3, 3}, {-3, 3},
    {2, 4}, {-2, 4}, {4, 2}, {-4, 2}, {0, 5}, {3, 4}, {-3, 4}, {4, 3}, {-4, 3}, {5, 0}, {1, 5}, {-1, 5}, {5, 1}, {-5, 1}, {2, 5},
    {-2, 5}, {5, 2}, {-5, 2}, {4, 4}, {-4, 4}, {3, 5}, {-3, 5}, {5, 3}, {-5, 3}, {0, 6}, {6, 0}, {1, 6}, {-1, 6}, {6, 1}, {-6, 1},
    {2, 6}, {-2, 6}, {6, 2}, {-6, 2}, {4, 5}, {-4, 5}, {5, 4}, {-5, 4}, {3, 6}, {-3, 6}, {6, 3}, {-6, 3}, {0, 7}, {7, 0}, {1, 7},
    {-1, 7}, {5, 5}, {-5, 5}, {7, 1}, {-7, 1}, {4, 6}, {-4, 6}, {6, 4}, {-6, 4}, {2, 7}, {-2, 7}, {7, 2}, {-7, 2}, {3, 7}, {-3, 7},
    {7, 3}, {-7, 3}, {5, 6}, {-5, 6}, {6, 5}, {-6, 5}, {8, 0}, {4, 7}, {-4, 7}, {7, 4}, {-7, 4}, {8, 1}, {8, 2}, {6, 6}, {-6, 6},
    {8, 3}, {5, 7}, {-5, 7}, {7, 5}, {-7, 5}, {8, 4}, {6, 7}, {-6, 7}, {7, 6}, {-7, 6}, {8, 5}, {7, 7}, {-7, 7}, {8, 6}, {8, 7}};

}  // namespace jxlf
