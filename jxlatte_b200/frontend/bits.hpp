// bits.hpp -- byte-buffer bit reader (LSB first) and the JPEG XL container demux for the C++ front end.
//
// The front end is the sequential host half of the decoder: it turns a .jxl file into the post-entropy
// frame state the CUDA reconstruction consumes.  It mirrors the behaviour of jxlatte's io/ package
// (J/io/Bitreader.java:26-98 for the field codings, J/io/Demuxer.java:55-150 for the box walk) but works on one
// in-memory buffer instead of a pull stream.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace jxlf {

struct StreamError : std::runtime_error {      // InvalidBitstreamException
    explicit StreamError(const std::string &m) : std::runtime_error(m) {}
};
struct Unsupported : std::runtime_error {      // UnsupportedOperationException
    explicit Unsupported(const std::string &m) : std::runtime_error(m) {}
};

class BitReader {
  public:
    BitReader() = default;
    BitReader(const uint8_t *data, size_t size) : p_(data), n_(size) {}

    // 0..32 bits, least significant first
    uint32_t bits(int count) {
        if (count == 0) return 0;
        fill();
        if (have_ < count) throw StreamError("unexpected end of stream");
        const uint32_t v = (uint32_t)(acc_ & (count == 64 ? ~0ull : ((1ull << count) - 1)));
        acc_ >>= count;
        have_ -= count;
        consumed_ += count;
        return v;
    }
    // up to 32 bits without consuming; bits past the end read as zero
    uint32_t peek(int count) {
        fill();
        return (uint32_t)(acc_ & ((1ull << count) - 1));
    }
    void drop(int count) {
        fill();
        if (have_ < count) throw StreamError("unexpected end of stream");
        acc_ >>= count;
        have_ -= count;
        consumed_ += count;
    }
    bool flag() { return bits(1) != 0; }
    uint32_t u32(uint32_t c0, int u0, uint32_t c1, int u1, uint32_t c2, int u2, uint32_t c3, int u3) {
        switch (bits(2)) {
        case 0: return c0 + bits(u0);
        case 1: return c1 + bits(u1);
        case 2: return c2 + bits(u2);
        default: return c3 + bits(u3);
        }
    }
    uint64_t u64() {
        switch (bits(2)) {
        case 0: return 0;
        case 1: return 1 + bits(4);
        case 2: return 17 + bits(8);
        default: break;
        }
        uint64_t v = bits(12);
        int shift = 12;
        while (flag()) {
            if (shift == 60) { v |= (uint64_t)bits(4) << shift; break; }
            v |= (uint64_t)bits(8) << shift;
            shift += 8;
        }
        return v;
    }
    float f16() {
        const uint32_t h = bits(16);
        const uint32_t mant = h & 0x3ff, e = (h >> 10) & 0x1f, sign = h >> 15;
        if (e == 31) throw StreamError("non-finite float16");
        if (e == 0) return (sign ? -1.0f : 1.0f) * (float)mant / 16777216.0f;
        const uint32_t w = (sign << 31) | ((e + 112) << 23) | (mant << 13);
        float f;
        std::memcpy(&f, &w, 4);
        return f;
    }
    uint32_t enumeration() {
        const uint32_t v = u32(0, 0, 1, 0, 2, 4, 18, 6);
        if (v > 63) throw StreamError("enum value above 63");
        return v;
    }
    uint32_t u8() {                      // the 0 / 1 / 2^n + u(n) code used inside histograms
        if (!flag()) return 0;
        const int n = (int)bits(3);
        return n == 0 ? 1 : bits(n) + (1u << n);
    }
    void align() {                       // zero_pad_to_byte
        const int r = (int)(consumed_ & 7);
        if (r && bits(8 - r) != 0) throw StreamError("non-zero padding bits");
    }
    void skip_bits(uint64_t count) {
        while (count > 32) { drop(32); count -= 32; }
        drop((int)count);
    }
    uint64_t position() const { return consumed_; }     // in bits
    bool at_end() {
        fill();
        return have_ == 0;
    }
    // a reader over the next `bytes` bytes (must be byte aligned); this reader skips past them
    BitReader section(size_t bytes) {
        if (consumed_ & 7) throw std::logic_error("section() on an unaligned reader");
        const size_t at = (size_t)(consumed_ >> 3);
        if (at + bytes > n_) throw StreamError("section runs past the end of the stream");
        BitReader r(p_ + at, bytes);
        seek_bytes(at + bytes);
        return r;
    }
    void seek_bytes(size_t at) {
        pos_ = at;
        acc_ = 0;
        have_ = 0;
        consumed_ = (uint64_t)at * 8;
    }
    const uint8_t *data() const { return p_; }
    size_t size() const { return n_; }

  private:
    void fill() {
        while (have_ <= 56 && pos_ < n_) {
            acc_ |= (uint64_t)p_[pos_++] << have_;
            have_ += 8;
        }
    }
    const uint8_t *p_ = nullptr;
    size_t n_ = 0, pos_ = 0;
    uint64_t acc_ = 0, consumed_ = 0;
    int have_ = 0;
};

// Raw codestream (FF 0A) or ISO-BMFF container: concatenates the jxlc / jxlp payloads.  Returns the codestream level.
inline int extract_codestream(const std::vector<uint8_t> &file, std::vector<uint8_t> &out) {
    static const uint8_t kSig[12] = {0, 0, 0, 0x0c, 'J', 'X', 'L', ' ', 0x0d, 0x0a, 0x87, 0x0a};
    int level = 5;
    if (file.size() < 12 || std::memcmp(file.data(), kSig, 12) != 0) {
        out = file;
        return level;
    }
    out.clear();
    size_t at = 12;
    auto be = [&](size_t o, int n) {
        uint64_t v = 0;
        for (int i = 0; i < n; i++) v = (v << 8) | file[o + i];
        return v;
    };
    while (at + 8 <= file.size()) {
        uint64_t size = be(at, 4);
        const uint32_t tag = (uint32_t)be(at + 4, 4);
        size_t header = 8;
        if (size == 1) {
            if (at + 16 > file.size()) throw StreamError("truncated extended box size");
            size = be(at + 8, 8);
            header = 16;
        }
        // compare against the bytes that are left, never form at + size first: a 64-bit extended size near 2^64 would wrap
        if (size != 0 && (size < header || size > (uint64_t)(file.size() - at))) throw StreamError("illegal box size");
        const size_t payload_end = size == 0 ? file.size() : at + (size_t)size;
        if (payload_end <= at) throw StreamError("illegal box size");       // the walk always moves forward
        size_t body = at + header;
        if (tag == 0x6a786c6c) {                 // jxll
            if (payload_end - body != 1) throw StreamError("jxll box must hold one byte");
            level = file[body];
            if (level != 5 && level != 10) throw StreamError("invalid codestream level");
        } else if (tag == 0x6a786c63) {          // jxlc
            out.insert(out.end(), file.begin() + body, file.begin() + payload_end);
        } else if (tag == 0x6a786c70) {          // jxlp: 4-byte sequence number first
            if (payload_end - body < 4) throw StreamError("truncated jxlp box");
            out.insert(out.end(), file.begin() + body + 4, file.begin() + payload_end);
        }
        at = payload_end;
    }
    return level;
}

inline int ceil_log1p(uint64_t x) { return x == 0 ? 0 : 64 - __builtin_clzll(x); }     // bits needed to write x
inline int ceil_log2(uint64_t x) { return x <= 1 ? 0 : ceil_log1p(x - 1); }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int32_t unpack_signed(uint32_t v) { return (v & 1) ? (int32_t)(-(int64_t)(((uint64_t)v + 1) >> 1)) : (int32_t)(v >> 1); }

}  // namespace jxlf
