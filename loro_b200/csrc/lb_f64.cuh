// loro_b200 -- f64 -> shortest round-trip decimal text, as serde_json prints LoroValue::Double
// (crates/loro-common/src/value.rs:692-711 -> serde_json -> ryu 1.x `format64`, "pretty" layout).
//
// The reference's digits come from Ryu: the shortest decimal that reads back as the same double, the closest one
// when several of that length exist, an exact half rounding to the even digit, interval ends accepted when the
// mantissa is even.  This is the same function computed the exact way (Steele-White / Burger-Dybvig free-format
// generation on big integers): doubles are rare in CRDT documents and each costs a few thousand integer operations
// at worst, so the 10 KB power-of-five tables of Ryu are not worth their constant memory.  One thread per value.
// Layout rules restated from ryu's pretty printer (length = digits, kk = position of the decimal point):
//   0 < kk <= 16 and no fraction digits: "1234000.0" ; 0 < kk <= 16: "12.34" ; -5 < kk <= 0: "0.001234" ;
//   otherwise scientific "1.234e33" / "1e-7" (no '+', no padding).  Non-finite values print as "null" (serde_json).
#pragma once
#include "lb_dev.cuh"

#define F64_WORDS 40   // 1280 bits: the largest intermediate is below 2^1140

struct F64Big {
    u32 w[F64_WORDS];
    int n;   // used words (no leading zero words)
    __device__ void set(u64 v) { n = 0; while (v) { w[n++] = (u32)v; v >>= 32; } }
    __device__ void mul_small(u32 m) {
        u64 c = 0;
        for (int i = 0; i < n; i++) { u64 x = (u64)w[i] * m + c; w[i] = (u32)x; c = x >> 32; }
        if (c && n < F64_WORDS) w[n++] = (u32)c;
    }
    __device__ void shl(int bits) {
        int ws = bits >> 5, bs = bits & 31;
        if (n == 0) return;
        if (bs) {
            u32 c = 0;
            for (int i = 0; i < n; i++) { u32 x = w[i]; w[i] = (x << bs) | c; c = x >> (32 - bs); }
            if (c && n < F64_WORDS) w[n++] = c;
        }
        if (ws) {
            for (int i = n - 1; i >= 0; i--) if (i + ws < F64_WORDS) w[i + ws] = w[i];
            for (int i = 0; i < ws; i++) w[i] = 0;
            n = n + ws < F64_WORDS ? n + ws : F64_WORDS;
        }
    }
    __device__ void add(const F64Big& o) {
        u64 c = 0;
        int m = n > o.n ? n : o.n;
        for (int i = 0; i < m; i++) {
            u64 x = (u64)(i < n ? w[i] : 0) + (i < o.n ? o.w[i] : 0) + c;
            w[i] = (u32)x;
            c = x >> 32;
        }
        n = m;
        if (c && n < F64_WORDS) w[n++] = (u32)c;
    }
    __device__ void sub(const F64Big& o) {   // *this >= o
        i64 c = 0;
        for (int i = 0; i < n; i++) {
            i64 x = (i64)w[i] - (i < o.n ? o.w[i] : 0) + c;
            w[i] = (u32)x;
            c = x >> 32;
        }
        while (n > 0 && w[n - 1] == 0) n--;
    }
};
__device__ inline int f64_cmp(const F64Big& a, const F64Big& b) {
    if (a.n != b.n) return a.n < b.n ? -1 : 1;
    for (int i = a.n - 1; i >= 0; i--)
        if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
}
// compare a + b with c
__device__ inline int f64_cmp_sum(const F64Big& a, const F64Big& b, const F64Big& c) {
    F64Big t = a;
    t.add(b);
    return f64_cmp(t, c);
}

// writes the text into out (at most 25 bytes), returns its length
__device__ __noinline__ int f64_format(u64 bits, char* out) {
    int len = 0;
    const u32 ieee_e = (u32)((bits >> 52) & 0x7FF);
    const u64 ieee_m = bits & 0xFFFFFFFFFFFFFull;
    if (ieee_e == 0x7FF) { out[0] = 'n'; out[1] = 'u'; out[2] = 'l'; out[3] = 'l'; return 4; }
    if (bits >> 63) out[len++] = '-';
    if (ieee_e == 0 && ieee_m == 0) { out[len++] = '0'; out[len++] = '.'; out[len++] = '0'; return len; }
    u64 f;
    int e2;
    if (ieee_e == 0) { f = ieee_m; e2 = -1074; }
    else { f = ieee_m | (1ull << 52); e2 = (int)ieee_e - 1075; }
    const bool even = (f & 1) == 0;
    const bool asym = ieee_m == 0 && ieee_e > 1;   // the lower neighbour is half as far away
    F64Big r, s, mp, mm;
    r.set(f);
    s.set(1);
    mp.set(1);
    mm.set(1);
    if (e2 >= 0) {
        r.shl(e2 + (asym ? 2 : 1));
        s.shl(asym ? 2 : 1);
        mp.shl(e2 + (asym ? 1 : 0));
        mm.shl(e2);
    } else {
        r.shl(asym ? 2 : 1);
        s.shl(-e2 + (asym ? 2 : 1));
        if (asym) mp.shl(1);
    }
    int blen = 64 - __clzll((long long)f);
    // k = ceil(log10(v)) estimated from the binary exponent: never too large, at most one too small (fixed below);
    // floor(n * log10(2)) == (n * 78913) >> 18 for |n| <= 1650
    const int nb = e2 + blen - 1;
    int k = nb == 0 ? 0 : ((nb * 78913) >> 18) + 1;
    if (k >= 0) for (int i = 0; i < k; i++) s.mul_small(10);
    else for (int i = 0; i < -k; i++) { r.mul_small(10); mp.mul_small(10); mm.mul_small(10); }
    {   // the estimate may be one too small
        int c = f64_cmp_sum(r, mp, s);
        if (even ? c >= 0 : c > 0) k++;
        else { r.mul_small(10); mp.mul_small(10); mm.mul_small(10); }
    }
    char dig[24];
    int nd = 0;
    while (nd < 20) {
        int d = 0;
        while (f64_cmp(r, s) >= 0) { r.sub(s); d++; }
        int c1 = f64_cmp(r, mm), c2 = f64_cmp_sum(r, mp, s);
        bool tc1 = even ? c1 <= 0 : c1 < 0;
        bool tc2 = even ? c2 >= 0 : c2 > 0;
        if (!tc1 && !tc2) {
            dig[nd++] = (char)('0' + d);
            r.mul_small(10); mp.mul_small(10); mm.mul_small(10);
            continue;
        }
        if (tc1 && tc2) {
            F64Big t2 = r;
            t2.shl(1);
            int c = f64_cmp(t2, s);
            if (c > 0 || (c == 0 && (d & 1))) d++;   // an exact half goes to the even digit (Ryu)
        } else if (tc2) d++;
        dig[nd++] = (char)('0' + d);
        break;
    }
    for (int i = nd - 1; i > 0 && dig[i] > '9'; i--) { dig[i] -= 10; dig[i - 1]++; }   // (defensive: a carry cannot occur)
    if (dig[0] > '9') { dig[0] = '1'; for (int i = 1; i < nd; i++) dig[i] = '0'; k++; }
    while (nd > 1 && dig[nd - 1] == '0') nd--;
    const int kk = k;   // value = 0.DIGITS x 10^kk
    if (0 < kk && kk <= 16 && nd <= kk) {
        for (int i = 0; i < nd; i++) out[len++] = dig[i];
        for (int i = nd; i < kk; i++) out[len++] = '0';
        out[len++] = '.';
        out[len++] = '0';
    } else if (0 < kk && kk <= 16) {
        for (int i = 0; i < kk; i++) out[len++] = dig[i];
        out[len++] = '.';
        for (int i = kk; i < nd; i++) out[len++] = dig[i];
    } else if (-5 < kk && kk <= 0) {
        out[len++] = '0';
        out[len++] = '.';
        for (int i = 0; i < -kk; i++) out[len++] = '0';
        for (int i = 0; i < nd; i++) out[len++] = dig[i];
    } else {
        out[len++] = dig[0];
        if (nd > 1) { out[len++] = '.'; for (int i = 1; i < nd; i++) out[len++] = dig[i]; }
        out[len++] = 'e';
        int ex = kk - 1;
        if (ex < 0) { out[len++] = '-'; ex = -ex; }
        char tmp[4];
        int tn = 0;
        do { tmp[tn++] = (char)('0' + ex % 10); ex /= 10; } while (ex);
        while (tn) out[len++] = tmp[--tn];
    }
    return len;
}
