// ORACLE (test infrastructure) -- CPU restatement of the loro wire codec primitives.
//
// This header is part of oracle/: it is the *checker*, never the product.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// The arithmetic restated here lives in third-party crates that are NOT vendored under
// /root/reference (Cargo.lock pins): serde_columnar 0.3.14, postcard 1.1.3, leb128 0.2.5,
// xxhash-rust 0.8.15.  The restatement follows their in-tree descriptions:
//   docs/encoding.md:861-1398 (LEB128, BoolRle, AnyRle, DeltaRle, DeltaOfDelta, postcard, columnar)
//   moon/loro_codec/serde_columnar_*.mbt, postcard_varint.mbt, leb128.mbt, xxhash32.mbt
// and is pinned by the known-answer vectors of moon/loro_codec/*_test.mbt and by the golden
// blobs under crates/loro/tests + crates/examples (see tests/test_oracle_codec.py).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>

namespace lo {

typedef unsigned __int128 u128;
typedef __int128 i128;

struct DecodeError : std::runtime_error {
    explicit DecodeError(const std::string& m) : std::runtime_error(m) {}
};

// ---------------------------------------------------------------- reader
struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    Reader(const uint8_t* b, size_t n) : p(b), end(b + n) {}
    size_t remaining() const { return (size_t)(end - p); }
    bool empty() const { return p >= end; }
    uint8_t u8() {
        if (p >= end) throw DecodeError("eof");
        return *p++;
    }
    const uint8_t* take(size_t n) {
        if (remaining() < n) throw DecodeError("eof(take)");
        const uint8_t* r = p;
        p += n;
        return r;
    }
    // postcard varint / ULEB128 share the 7-bit little-endian group layout
    // (docs/encoding.md:869-946, 1220-1239; moon/loro_codec/postcard_varint.mbt, leb128.mbt)
    u128 varint_u128() {
        u128 v = 0;
        int shift = 0;
        for (int i = 0; i < 19; i++) {
            uint8_t b = u8();
            v |= (u128)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
        throw DecodeError("varint too long");
    }
    uint64_t varint() {
        uint64_t v = 0;
        int shift = 0;
        for (int i = 0; i < 10; i++) {
            uint8_t b = u8();
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
        throw DecodeError("varint too long");
    }
    uint64_t uleb() { return varint(); }
    // postcard signed = zigzag (docs/encoding.md:1241-1260)
    int64_t zigzag() {
        uint64_t v = varint();
        return (int64_t)(v >> 1) ^ -(int64_t)(v & 1);
    }
    i128 zigzag128() {
        u128 v = varint_u128();
        return (i128)(v >> 1) ^ -(i128)(v & 1);
    }
    // SLEB128: two's complement, NOT zigzag (docs/encoding.md:948-1054; value.rs:858-888)
    int64_t sleb() {
        int64_t result = 0;
        int shift = 0;
        uint8_t b;
        int n = 0;
        do {
            b = u8();
            if (shift < 64) result |= (int64_t)(uint64_t)(b & 0x7f) << shift;
            shift += 7;
            if (++n > 10) throw DecodeError("sleb too long");
        } while (b & 0x80);
        if (shift < 64 && (b & 0x40)) result |= -((int64_t)1 << shift);
        return result;
    }
};

// ---------------------------------------------------------------- writer
struct Writer {
    std::vector<uint8_t> buf;
    void u8(uint8_t b) { buf.push_back(b); }
    void bytes(const uint8_t* b, size_t n) { buf.insert(buf.end(), b, b + n); }
    void bytes(const std::vector<uint8_t>& v) { buf.insert(buf.end(), v.begin(), v.end()); }
    void bytes(const std::string& s) { buf.insert(buf.end(), s.begin(), s.end()); }
    void varint(uint64_t v) {
        while (v >= 0x80) {
            buf.push_back((uint8_t)(v | 0x80));
            v >>= 7;
        }
        buf.push_back((uint8_t)v);
    }
    void varint_u128(u128 v) {
        while (v >= 0x80) {
            buf.push_back((uint8_t)((uint8_t)v | 0x80));
            v >>= 7;
        }
        buf.push_back((uint8_t)v);
    }
    void uleb(uint64_t v) { varint(v); }
    void zigzag(int64_t v) { varint(((uint64_t)v << 1) ^ (uint64_t)(v >> 63)); }
    void zigzag128(i128 v) { varint_u128(((u128)v << 1) ^ (u128)(v >> 127)); }
    void sleb(int64_t v) {
        bool more = true;
        while (more) {
            uint8_t b = v & 0x7f;
            v >>= 7;  // arithmetic
            if ((v == 0 && !(b & 0x40)) || (v == -1 && (b & 0x40)))
                more = false;
            else
                b |= 0x80;
            buf.push_back(b);
        }
    }
    // postcard bytes / str: varint len + payload (docs/encoding.md:1262-1284)
    void len_bytes(const std::vector<uint8_t>& v) {
        varint(v.size());
        bytes(v);
    }
};

// ---------------------------------------------------------------- xxHash32
// moon/loro_codec/xxhash32.mbt; docs/encoding-xxhash32.md; call site encoding.rs:278-295, 397-416.
static const uint32_t XXH_SEED_LORO = 0x4F524F4Cu;
inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
inline uint32_t rd32le(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
inline uint32_t xxh32(const uint8_t* d, size_t len, uint32_t seed) {
    const uint32_t P1 = 0x9E3779B1u, P2 = 0x85EBCA77u, P3 = 0xC2B2AE3Du, P4 = 0x27D4EB2Fu,
                   P5 = 0x165667B1u;
    size_t off = 0;
    uint32_t h;
    if (len >= 16) {
        uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        size_t limit = len - 16;
        while (off <= limit) {
            v1 = rotl32(v1 + rd32le(d + off) * P2, 13) * P1;
            v2 = rotl32(v2 + rd32le(d + off + 4) * P2, 13) * P1;
            v3 = rotl32(v3 + rd32le(d + off + 8) * P2, 13) * P1;
            v4 = rotl32(v4 + rd32le(d + off + 12) * P2, 13) * P1;
            off += 16;
        }
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else {
        h = seed + P5;
    }
    h += (uint32_t)len;
    while (off + 4 <= len) {
        h = rotl32(h + rd32le(d + off) * P3, 17) * P4;
        off += 4;
    }
    while (off < len) {
        h = rotl32(h + d[off] * P5, 11) * P1;
        off++;
    }
    h ^= h >> 15;
    h *= P2;
    h ^= h >> 13;
    h *= P3;
    h ^= h >> 16;
    return h;
}

// ---------------------------------------------------------------- BoolRle
// docs/encoding.md:1058-1082, 1316-1340; moon/loro_codec/serde_columnar_bool_rle.mbt.
// Alternating run lengths starting with `false` (first may be 0).
inline std::vector<bool> bool_rle_take_n(Reader& r, size_t n) {
    std::vector<bool> out;
    bool state = false;
    while (out.size() < n) {
        uint64_t len = r.varint();
        if (out.size() + len > n) throw DecodeError("boolrle: too many elements");
        for (uint64_t i = 0; i < len; i++) out.push_back(state);
        state = !state;
    }
    return out;
}
inline void bool_rle_encode(Writer& w, const std::vector<bool>& v) {
    if (v.empty()) return;
    bool state = false;
    uint64_t run = 0;
    for (bool b : v) {
        if (b == state)
            run++;
        else {
            w.varint(run);
            state = !state;
            run = 1;
        }
    }
    w.varint(run);
}

// ---------------------------------------------------------------- AnyRle<T>
// docs/encoding.md:1084-1115, 1342-1365; moon/loro_codec/serde_columnar_any_rle.mbt.
// Segment header zigzag(len): len>0 run of one value, len<0 |len| literal values.
enum class VK { U8, VARU, ZIG, ZIG128 };  // how one value is written

template <class T, class RD>
inline void any_rle_decode_all(Reader& r, std::vector<T>& out, RD rd) {
    while (!r.empty()) {
        int64_t sl = r.zigzag();
        if (sl == 0) throw DecodeError("anyrle: zero length segment");
        if (sl > 0) {
            T v = rd(r);
            for (int64_t i = 0; i < sl; i++) out.push_back(v);
        } else {
            for (int64_t i = 0; i < -sl; i++) out.push_back(rd(r));
        }
    }
}
template <class T, class RD>
inline void any_rle_take_n(Reader& r, size_t n, std::vector<T>& out, RD rd) {
    size_t got = 0;
    while (got < n) {
        int64_t sl = r.zigzag();
        if (sl == 0) throw DecodeError("anyrle: zero length segment");
        uint64_t len = sl > 0 ? (uint64_t)sl : (uint64_t)(-sl);
        if (got + len > n) throw DecodeError("anyrle: too many elements");
        if (sl > 0) {
            T v = rd(r);
            for (uint64_t i = 0; i < len; i++) out.push_back(v);
        } else {
            for (uint64_t i = 0; i < len; i++) out.push_back(rd(r));
        }
        got += len;
    }
}

// Encoder state machine (upstream source not in tree; verified against
// every ops/delete_start_ids/header/change_meta section of the in-tree golden blobs by
// tests/test_oracle_golden.py).  States Empty / LoneVal / Run / LiteralRun.
template <class T, class WR>
struct AnyRleEncoder {
    Writer& w;
    WR wr;
    enum { Empty, Lone, Run, Lit } st = Empty;
    T last{};
    uint64_t run = 0;
    std::vector<T> lit;
    AnyRleEncoder(Writer& w_, WR wr_) : w(w_), wr(wr_) {}
    void flush_run(const T& v, uint64_t n) {
        w.zigzag((int64_t)n);
        wr(w, v);
    }
    void flush_lit() {
        w.zigzag(-(int64_t)lit.size());
        for (auto& v : lit) wr(w, v);
        lit.clear();
    }
    void append(const T& x) {
        switch (st) {
            case Empty:
                st = Lone;
                last = x;
                break;
            case Lone:
                if (x == last) {
                    st = Run;
                    run = 2;
                } else {
                    lit.clear();
                    lit.push_back(last);
                    last = x;
                    st = Lit;
                }
                break;
            case Run:
                if (x == last)
                    run++;
                else {
                    flush_run(last, run);
                    st = Lone;
                    last = x;
                }
                break;
            case Lit:
                if (x == last) {
                    flush_lit();
                    st = Run;
                    run = 2;
                } else {
                    lit.push_back(last);
                    last = x;
                }
                break;
        }
    }
    void finish() {
        switch (st) {
            case Empty: break;
            case Lone:
                lit.clear();
                lit.push_back(last);
                flush_lit();
                break;
            case Run: flush_run(last, run); break;
            case Lit:
                lit.push_back(last);
                flush_lit();
                break;
        }
        st = Empty;
    }
};

// value read/write functors
struct RdU8 { uint64_t operator()(Reader& r) const { return r.u8(); } };
struct RdVar { uint64_t operator()(Reader& r) const { return r.varint(); } };
struct RdZig { int64_t operator()(Reader& r) const { return r.zigzag(); } };
struct RdZig128 { i128 operator()(Reader& r) const { return r.zigzag128(); } };
struct WrU8 { void operator()(Writer& w, uint64_t v) const { w.u8((uint8_t)v); } };
struct WrVar { void operator()(Writer& w, uint64_t v) const { w.varint(v); } };
struct WrZig { void operator()(Writer& w, int64_t v) const { w.zigzag(v); } };
struct WrZig128 { void operator()(Writer& w, i128 v) const { w.zigzag128(v); } };

// DeltaRle = AnyRle<i128> over deltas from 0 (docs/encoding.md:1119-1124, 1367-1398).
inline std::vector<int64_t> delta_rle_decode_all(Reader& r) {
    std::vector<i128> d;
    any_rle_decode_all<i128>(r, d, RdZig128());
    std::vector<int64_t> out;
    i128 acc = 0;
    for (auto x : d) {
        acc += x;
        out.push_back((int64_t)acc);
    }
    return out;
}
inline void delta_rle_encode(Writer& w, const std::vector<int64_t>& v) {
    AnyRleEncoder<i128, WrZig128> e(w, WrZig128());
    i128 prev = 0;
    for (auto x : v) {
        e.append((i128)x - prev);
        prev = x;
    }
    e.finish();
}

// ---------------------------------------------------------------- DeltaOfDelta
// docs/encoding.md:1126-1172; moon/loro_codec/serde_columnar_delta_of_delta_{decode,encode,bits}.mbt.
// Option<i64> first value, u8 last_used_bits, MSB-first prefix-coded delta-of-deltas.
struct BitCursor {
    const uint8_t* b;
    size_t nbytes;
    size_t bitpos = 0;
    BitCursor(const uint8_t* b_, size_t n) : b(b_), nbytes(n) {}
    bool bit() {
        if (bitpos >= nbytes * 8) throw DecodeError("dod: eof");
        bool v = (b[bitpos >> 3] >> (7 - (bitpos & 7))) & 1;
        bitpos++;
        return v;
    }
    uint64_t bits(int n) {
        uint64_t v = 0;
        for (int i = 0; i < n; i++) v = (v << 1) | (bit() ? 1 : 0);
        return v;
    }
};
// take n values; leaves r positioned at the next byte boundary after the consumed bits
// (DeltaOfDeltaDecoder::take_n_finalize as used by block_meta_encode.rs:121-157).
inline std::vector<int64_t> dod_take_n(Reader& r, size_t n) {
    uint8_t tag = r.u8();
    std::vector<int64_t> out;
    int64_t first = 0;
    if (tag == 1)
        first = r.zigzag();
    else if (tag != 0)
        throw DecodeError("dod: bad option tag");
    uint8_t last_used_bits = r.u8();
    if (tag == 0) {
        if (n != 0) throw DecodeError("dod: not enough elements");
        if (last_used_bits != 0) throw DecodeError("dod: invalid empty");
        return out;
    }
    if (n == 0) throw DecodeError("dod: too many elements");
    if (last_used_bits > 8) throw DecodeError("dod: invalid last_used_bits");
    BitCursor bc(r.p, r.remaining());
    int64_t prev = first, delta = 0;
    out.push_back(first);
    for (size_t i = 1; i < n; i++) {
        int64_t dod;
        if (!bc.bit())
            dod = 0;
        else if (!bc.bit())
            dod = (int64_t)bc.bits(7) - 63;
        else if (!bc.bit())
            dod = (int64_t)bc.bits(9) - 255;
        else if (!bc.bit())
            dod = (int64_t)bc.bits(12) - 2047;
        else if (!bc.bit())
            dod = (int64_t)bc.bits(21) - 1048575;
        else
            dod = (int64_t)bc.bits(64);
        delta += dod;
        prev += delta;
        out.push_back(prev);
    }
    r.take((bc.bitpos + 7) / 8);
    return out;
}
struct BitWriter {
    std::vector<uint8_t> bytes;
    uint32_t cur = 0;
    int nbits = 0;
    void bit(bool b) {
        cur = (cur << 1) | (b ? 1 : 0);
        if (++nbits == 8) {
            bytes.push_back((uint8_t)cur);
            cur = 0;
            nbits = 0;
        }
    }
    void bits(uint64_t v, int n) {
        for (int i = n - 1; i >= 0; i--) bit((v >> i) & 1);
    }
};
inline void dod_encode(Writer& w, const std::vector<int64_t>& v) {
    if (v.empty()) {
        w.u8(0);
        w.u8(0);
        return;
    }
    w.u8(1);
    w.zigzag(v[0]);
    if (v.size() == 1) {
        w.u8(0);
        return;
    }
    BitWriter bw;
    int64_t prev_delta = 0;
    for (size_t i = 1; i < v.size(); i++) {
        int64_t d = v[i] - v[i - 1];
        int64_t x = d - prev_delta;
        prev_delta = d;
        if (x == 0)
            bw.bit(false);
        else if (x >= -63 && x <= 64) {
            bw.bits(0b10, 2);
            bw.bits((uint64_t)(x + 63), 7);
        } else if (x >= -255 && x <= 256) {
            bw.bits(0b110, 3);
            bw.bits((uint64_t)(x + 255), 9);
        } else if (x >= -2047 && x <= 2048) {
            bw.bits(0b1110, 4);
            bw.bits((uint64_t)(x + 2047), 12);
        } else if (x >= -1048575 && x <= 1048576) {
            bw.bits(0b11110, 5);
            bw.bits((uint64_t)(x + 1048575), 21);
        } else {
            bw.bits(0b11111, 5);
            bw.bits((uint64_t)x, 64);
        }
    }
    if (bw.nbits == 0) {
        w.u8(8);
        w.bytes(bw.bytes);
        return;
    }
    int used = bw.nbits;
    bw.bytes.push_back((uint8_t)((bw.cur & 0xFF) << (8 - used)));
    w.u8((uint8_t)used);
    w.bytes(bw.bytes);
}

// ---------------------------------------------------------------- columnar wrapper
// docs/encoding.md:1296-1314; moon/loro_codec/serde_columnar.mbt.
// `#[columnar(ser,de)] struct S { #[columnar(class="vec")] f: Vec<Row> }` serialises as
// varint(1 field) + [varint ncols + per column (varint nbytes, payload)].
inline std::vector<std::pair<const uint8_t*, size_t>> columnar_take_wrapped(const uint8_t* b,
                                                                            size_t n,
                                                                            size_t expect_cols) {
    Reader r(b, n);
    uint64_t fields = r.varint();
    if (fields != 1) throw DecodeError("columnar: expected 1-field wrapper");
    uint64_t ncols = r.varint();
    if (ncols != expect_cols) throw DecodeError("columnar: column count mismatch");
    std::vector<std::pair<const uint8_t*, size_t>> cols;
    for (uint64_t i = 0; i < ncols; i++) {
        uint64_t len = r.varint();
        cols.push_back({r.take(len), (size_t)len});
    }
    if (!r.empty()) throw DecodeError("columnar: trailing bytes");
    return cols;
}
inline void columnar_write_wrapped(Writer& w, const std::vector<std::vector<uint8_t>>& cols) {
    w.varint(1);
    w.varint(cols.size());
    for (auto& c : cols) w.len_bytes(c);
}

}  // namespace lo
