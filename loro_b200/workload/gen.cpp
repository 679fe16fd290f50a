// loro_b200 workload generator -- seeded synthetic update streams in the reference's FastUpdates format.
//
// Produces what `export(ExportMode::all_updates())` of a fully synced replica yields for a document edited by
// several peers (mixed List/Map ops, peers fork from a common prefix, edit
// concurrently and sync pairwise).  It is a host-side *input* tool for bench.py and the tests: it never runs
// inside the measured path and shares no code with oracle/ (the checker) -- its replicas integrate remote
// ops with an origin-based Fugue list, an algorithm formulation independent of the eg-walker replay used by
// both the oracle and the CUDA engine, so a three-way agreement on final states is a real cross-check.
//
// Wire format restated from /root/reference/docs/encoding.md (change block :467-678, values :680-860,
// column codecs :1056-1398, header/checksum :44-90); local-op shapes follow
// crates/loro-internal/src/handler.rs:2744-2779 (list delete = one op per element, run-merged) and
// txn.rs:560-625 (counter / lamport assignment); Fugue sibling rule from
// container/richtext/tracker/crdt_rope.rs:138-217.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint8_t u8;
typedef unsigned __int128 u128;
typedef __int128 i128;

// ------------------------------------------------------------------ rng: xoshiro256**
struct Rng {
    u64 s[4];
    static u64 splitmix(u64& x) {
        u64 z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(u64 seed) { for (int i = 0; i < 4; i++) s[i] = splitmix(seed); }
    static u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
    u64 next() {
        u64 r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    u32 below(u32 n) { return n ? (u32)(next() % n) : 0; }
    double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

// ------------------------------------------------------------------ byte writer + codecs
struct W {
    std::vector<u8> b;
    void u8_(u8 x) { b.push_back(x); }
    void bytes(const u8* p, size_t n) { b.insert(b.end(), p, p + n); }
    void bytes(const std::vector<u8>& v) { b.insert(b.end(), v.begin(), v.end()); }
    void varint(u64 v) { while (v >= 0x80) { b.push_back((u8)(v | 0x80)); v >>= 7; } b.push_back((u8)v); }
    void varint128(u128 v) { while (v >= 0x80) { b.push_back((u8)((u8)v | 0x80)); v >>= 7; } b.push_back((u8)v); }
    void zig(i64 v) { varint(((u64)v << 1) ^ (u64)(v >> 63)); }
    void zig128(i128 v) { varint128(((u128)v << 1) ^ (u128)(v >> 127)); }
    void sleb(i64 v) {
        bool more = true;
        while (more) {
            u8 x = v & 0x7f;
            v >>= 7;
            if ((v == 0 && !(x & 0x40)) || (v == -1 && (x & 0x40))) more = false; else x |= 0x80;
            b.push_back(x);
        }
    }
    void lenbytes(const std::vector<u8>& v) { varint(v.size()); bytes(v); }
};

// AnyRle encoder (LoneVal / Run / LiteralRun machine of serde_columnar; docs/encoding.md:1084-1115)
template <class T, class WR>
struct AnyRle {
    W& w; WR wr;
    int st = 0;  // 0 empty 1 lone 2 run 3 lit
    T last{}; u64 run = 0; std::vector<T> lit;
    AnyRle(W& w_, WR wr_) : w(w_), wr(wr_) {}
    void flush_lit() { w.zig(-(i64)lit.size()); for (auto& v : lit) wr(w, v); lit.clear(); }
    void push(const T& x) {
        switch (st) {
            case 0: st = 1; last = x; break;
            case 1: if (x == last) { st = 2; run = 2; } else { lit.clear(); lit.push_back(last); last = x; st = 3; } break;
            case 2: if (x == last) run++; else { w.zig((i64)run); wr(w, last); st = 1; last = x; } break;
            case 3: if (x == last) { flush_lit(); st = 2; run = 2; } else { lit.push_back(last); last = x; } break;
        }
    }
    void finish() {
        if (st == 1) { lit.clear(); lit.push_back(last); flush_lit(); }
        else if (st == 2) { w.zig((i64)run); wr(w, last); }
        else if (st == 3) { lit.push_back(last); flush_lit(); }
        st = 0;
    }
};
struct WrU8 { void operator()(W& w, u64 v) const { w.u8_((u8)v); } };
struct WrVar { void operator()(W& w, u64 v) const { w.varint(v); } };
struct WrZ128 { void operator()(W& w, i128 v) const { w.zig128(v); } };
void delta_rle(W& w, const std::vector<i64>& v) {
    AnyRle<i128, WrZ128> e(w, WrZ128());
    i128 prev = 0;
    for (auto x : v) { e.push((i128)x - prev); prev = x; }
    e.finish();
}
void bool_rle(W& w, const std::vector<bool>& v) {
    if (v.empty()) return;
    bool state = false; u64 run = 0;
    for (bool x : v) { if (x == state) run++; else { w.varint(run); state = !state; run = 1; } }
    w.varint(run);
}
// DeltaOfDelta (docs/encoding.md:1126-1172)
void dod(W& w, const std::vector<i64>& v) {
    if (v.empty()) { w.u8_(0); w.u8_(0); return; }
    w.u8_(1); w.zig(v[0]);
    if (v.size() == 1) { w.u8_(0); return; }
    std::vector<u8> out; u32 cur = 0; int nb = 0;
    auto bit = [&](bool x) { cur = (cur << 1) | (x ? 1 : 0); if (++nb == 8) { out.push_back((u8)cur); cur = 0; nb = 0; } };
    auto bits = [&](u64 x, int n) { for (int i = n - 1; i >= 0; i--) bit((x >> i) & 1); };
    i64 pd = 0;
    for (size_t i = 1; i < v.size(); i++) {
        i64 d = v[i] - v[i - 1], x = d - pd; pd = d;
        if (x == 0) bit(false);
        else if (x >= -63 && x <= 64) { bits(2, 2); bits((u64)(x + 63), 7); }
        else if (x >= -255 && x <= 256) { bits(6, 3); bits((u64)(x + 255), 9); }
        else if (x >= -2047 && x <= 2048) { bits(14, 4); bits((u64)(x + 2047), 12); }
        else if (x >= -1048575 && x <= 1048576) { bits(30, 5); bits((u64)(x + 1048575), 21); }
        else { bits(31, 5); bits((u64)x, 64); }
    }
    if (nb == 0) { w.u8_(8); w.bytes(out); return; }
    int used = nb;
    out.push_back((u8)((cur & 0xFF) << (8 - used)));
    w.u8_((u8)used); w.bytes(out);
}
u32 rotl32(u32 x, int r) { return (x << r) | (x >> (32 - r)); }
u32 rd32(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
u32 xxh32(const u8* d, size_t len, u32 seed) {
    const u32 P1 = 0x9E3779B1u, P2 = 0x85EBCA77u, P3 = 0xC2B2AE3Du, P4 = 0x27D4EB2Fu, P5 = 0x165667B1u;
    size_t off = 0; u32 h;
    if (len >= 16) {
        u32 v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        while (off + 16 <= len) {
            v1 = rotl32(v1 + rd32(d + off) * P2, 13) * P1; v2 = rotl32(v2 + rd32(d + off + 4) * P2, 13) * P1;
            v3 = rotl32(v3 + rd32(d + off + 8) * P2, 13) * P1; v4 = rotl32(v4 + rd32(d + off + 12) * P2, 13) * P1;
            off += 16;
        }
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else h = seed + P5;
    h += (u32)len;
    while (off + 4 <= len) { h = rotl32(h + rd32(d + off) * P3, 17) * P4; off += 4; }
    while (off < len) { h = rotl32(h + d[off] * P5, 11) * P1; off++; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

// ------------------------------------------------------------------ ops / changes
struct Id { int peer; i32 ctr; bool operator==(const Id& o) const { return peer == o.peer && ctr == o.ctr; } };
const Id NO_ID{-1, -1};
enum Kind : u8 { K_LIST_INS, K_LIST_DEL, K_MAP_SET, K_MAP_DEL, K_TREE, K_TEXT_INS, K_TEXT_DEL };
struct Val { bool is_str; i64 i; char s[9]; u8 slen; };
// values of a list insert: the first value inline (runs are rare with random positions), the rest on the heap
struct Vals {
    Val first; u32 n = 0; std::vector<Val> rest;
    size_t size() const { return n; }
    void push_back(const Val& v) { if (n == 0) first = v; else rest.push_back(v); n++; }
    const Val& operator[](size_t i) const { return i == 0 ? first : rest[i - 1]; }
    void append(const Vals& o) { for (size_t i = 0; i < o.size(); i++) push_back(o[i]); }
};
struct Op {
    Kind kind; i32 ctr;
    i32 pos;                   // list ops
    Vals vals;                 // list insert
    Id ol, orr;                // origins of the first inserted atom (replication only, not on the wire)
    Id del_start; i32 del_len; // signed
    u8 key; Val mapval;        // map ops
    Id target, parent;         // tree ops: parent.peer -1 = root, -2 = DELETED_TREE_ROOT
    std::string position;      // fractional index bytes
    std::string text;          // text insert (ASCII here: unicode length == bytes)
    u32 arena_gen = 0;         // string arena buffer generation (append-only buffer, capacity doubling from 32)
    u64 arena_start = 0, arena_end = 0;
    int atoms() const {
        return kind == K_LIST_INS ? (int)vals.size() : (kind == K_LIST_DEL || kind == K_TEXT_DEL) ? (del_len < 0 ? -del_len : del_len)
               : kind == K_TEXT_INS ? (int)text.size() : 1;
    }
};
struct Change {
    int peer; i32 ctr; u32 lamport; std::vector<Id> deps; std::vector<Op> ops;
    int atoms() const { return ops.empty() ? 0 : ops.back().ctr + ops.back().atoms() - ctr; }
};

// DeleteSpan merge rules (reference: container/list/list_op.rs:189-249, 381-434)
bool del_mergable(const Op& a, const Op& b) {
    auto bid = [](const Op& o) { return o.del_len == 1 || o.del_len == -1; };
    auto start = [](const Op& o) { return o.del_len > 0 ? o.pos : o.pos + 1 + o.del_len; };
    auto next_pos = [&](const Op& o) { return o.del_len > 0 ? start(o) : start(o) - 1; };
    auto prev_pos = [](const Op& o) { return o.del_len > 0 ? o.pos : o.pos + 1; };
    auto id_end = [](const Op& o) { return Id{o.del_start.peer, o.del_start.ctr + (o.del_len < 0 ? -o.del_len : o.del_len)}; };
    auto inc1 = [](Id x) { return Id{x.peer, x.ctr + 1}; };
    if (bid(a) && bid(b)) return (a.pos == b.pos && inc1(a.del_start) == b.del_start) || (a.pos == b.pos + 1 && a.del_start == inc1(b.del_start));
    if (bid(a)) { if (a.pos == prev_pos(b)) return b.del_len > 0 ? inc1(a.del_start) == b.del_start : a.del_start == id_end(b); return false; }
    if (bid(b)) { if (next_pos(a) == b.pos) return a.del_len > 0 ? id_end(a) == b.del_start : a.del_start == inc1(b.del_start); return false; }
    if (next_pos(a) == b.pos && (a.del_len > 0) == (b.del_len > 0)) return a.del_len > 0 ? id_end(a) == b.del_start : a.del_start == id_end(b);
    return false;
}
void del_merge(Op& a, const Op& b) {
    auto bid = [](const Op& o) { return o.del_len == 1 || o.del_len == -1; };
    a.del_start.ctr = std::min(a.del_start.ctr, b.del_start.ctr);
    if (bid(a) && bid(b)) a.del_len = a.pos == b.pos ? 2 : -2;
    else if (bid(a)) a.del_len = b.del_len + (b.del_len > 0 ? 1 : -1);
    else if (bid(b)) a.del_len += a.del_len > 0 ? 1 : -1;
    else a.del_len += b.del_len;
}
bool op_mergable(const Op& a, const Op& b) {
    if (a.kind != b.kind || a.ctr + a.atoms() != b.ctr) return false;
    if (a.kind == K_LIST_INS) return a.pos + (i32)a.vals.size() == b.pos && a.arena_end == b.arena_start;
    if (a.kind == K_TEXT_INS) return a.pos + (i32)a.text.size() == b.pos && a.arena_end == b.arena_start && a.arena_gen == b.arena_gen;
    if (a.kind == K_LIST_DEL || a.kind == K_TEXT_DEL) return del_mergable(a, b);
    return false;
}
bool rle_push(std::vector<Op>& ops, const Op& op) {
    if (!ops.empty() && op_mergable(ops.back(), op)) {
        Op& a = ops.back();
        if (a.kind == K_LIST_INS) { a.vals.append(op.vals); a.arena_end = op.arena_end; }
        else if (a.kind == K_TEXT_INS) { a.text += op.text; a.arena_end = op.arena_end; }
        else del_merge(a, op);
        return true;
    }
    ops.push_back(op);
    return false;
}
size_t op_estimate(const Op& o) {
    return o.kind == K_LIST_INS ? 4 * o.vals.size() : o.kind == K_TEXT_INS ? o.text.size() : (o.kind == K_LIST_DEL || o.kind == K_TREE || o.kind == K_TEXT_DEL) ? 8 : 3;
}
size_t change_estimate(const Change& c) {
    size_t s = 4 + (std::max<size_t>(c.deps.size(), 1) - 1) * 4;
    for (auto& o : c.ops) s += op_estimate(o);
    return s;
}

// ------------------------------------------------------------------ origin-based Fugue list replica
struct Item { Id id; Id ol, orr; bool deleted; Val v; struct Blk* blk; };
struct Blk { std::vector<Item*> it; int vis = 0; int idx = 0; };
struct FList {
    std::vector<Blk*> blocks;
    std::vector<std::vector<Item*>> by_id;  // per peer: atom counter -> item (sparse: only list atoms)
    std::vector<Item*> chunks;
    size_t chunk_used = 1024;
    explicit FList(int npeers) : by_id(npeers) { blocks.push_back(new Blk()); }
    ~FList() { for (auto b : blocks) delete b; for (auto c : chunks) delete[] c; }
    Item* alloc_item() {
        if (chunk_used == 1024) { chunks.push_back(new Item[1024]); chunk_used = 0; }
        return &chunks.back()[chunk_used++];
    }
    int visible = 0;
    Item* find(Id id) {
        if (id.peer < 0) return nullptr;
        auto& v = by_id[id.peer];
        return (size_t)id.ctr < v.size() ? v[id.ctr] : nullptr;
    }
    // global position (block index, index in block)
    std::pair<int, int> locate(Item* x) {
        int bi = x->blk->idx;
        auto& v = x->blk->it;
        int k = 0;
        while (v[k] != x) k++;
        return {bi, k};
    }
    i64 order(Item* x) { auto p = locate(x); return ((i64)p.first << 20) | p.second; }
    void insert_at(int bi, int k, Item* x) {
        Blk* b = blocks[bi];
        b->it.insert(b->it.begin() + k, x);
        x->blk = b;
        if (!x->deleted) { b->vis++; visible++; }
        if (b->it.size() > 64) {
            Blk* nb = new Blk();
            nb->it.assign(b->it.begin() + 32, b->it.end());
            b->it.resize(32);
            for (auto y : nb->it) { y->blk = nb; if (!y->deleted) { nb->vis++; b->vis--; } }
            blocks.insert(blocks.begin() + bi + 1, nb);
            for (size_t q = bi + 1; q < blocks.size(); q++) blocks[q]->idx = (int)q;
        }
    }
    // k-th visible item
    Item* visible_at(int pos) {
        for (auto b : blocks) {
            if (pos < b->vis) { for (auto x : b->it) if (!x->deleted) { if (pos == 0) return x; pos--; } }
            pos -= b->vis;
        }
        return nullptr;
    }
    // successor in total order (tombstones included)
    Item* next_of(Item* x) {
        auto p = locate(x);
        if (p.second + 1 < (int)blocks[p.first]->it.size()) return blocks[p.first]->it[p.second + 1];
        for (int bi = p.first + 1; bi < (int)blocks.size(); bi++) if (!blocks[bi]->it.empty()) return blocks[bi]->it[0];
        return nullptr;
    }
    Item* first() { for (auto b : blocks) if (!b->it.empty()) return b->it[0]; return nullptr; }
    void reg(Item* x) {
        auto& v = by_id[x->id.peer];
        if (v.size() <= (size_t)x->id.ctr) v.resize(x->id.ctr + 1, nullptr);
        v[x->id.ctr] = x;
    }
    Id ol_of(Id id) { Item* x = find(id); return x ? x->ol : NO_ID; }
    // integrate one atom with known origins (local or remote); peers[] give the real ids for tie-breaks
    void integrate(Item* nw, const std::vector<u64>& peer_ids) {
        reg(nw);
        Item* L = find(nw->ol);
        Item* R = find(nw->orr);
        // start right after L
        int bi = 0, k = 0;
        if (L) { auto p = locate(L); bi = p.first; k = p.second + 1; }
        // my right parent: R if its origin_left equals mine; a missing R is the end-of-document placeholder whose
        // origin_left is "none" (reference: tracker.rs:38-63)
        const i64 INF = (i64)1 << 60;
        bool has_pr; i64 pr_key = 0;
        if (R) { has_pr = R->ol == nw->ol; if (has_pr) pr_key = order(R); }
        else { has_pr = nw->ol == NO_ID; pr_key = INF; }
        int ins_b = bi, ins_k = k;
        bool scanning = false;
        std::vector<Item*> visited;
        int cb = bi, ck = k;
        while (true) {
            while (cb < (int)blocks.size() && ck >= (int)blocks[cb]->it.size()) { cb++; ck = 0; }
            if (cb >= (int)blocks.size()) break;
            Item* o = blocks[cb]->it[ck];
            if (o == R) break;
            bool same_ol = o->ol == nw->ol;
            if (!same_ol) {
                bool in_vis = false;
                for (auto v : visited) if (v->id == o->ol) { in_vis = true; break; }
                if (!in_vis) break;
            }
            visited.push_back(o);
            if (same_ol) {
                if (o->orr == nw->orr) {
                    if (peer_ids[o->id.peer] > peer_ids[nw->id.peer]) break;
                    scanning = false;
                } else {
                    bool o_has; i64 o_key = 0;
                    Item* orr = find(o->orr);
                    if (orr) { o_has = orr->ol == nw->ol; if (o_has) o_key = order(orr); }
                    else { o_has = nw->ol == NO_ID; o_key = INF; }
                    int cmp;
                    if (o_has && has_pr) cmp = o_key < pr_key ? -1 : (o_key > pr_key ? 1 : 0);
                    else if (o_has) cmp = -1;
                    else if (has_pr) cmp = 1;
                    else cmp = 0;
                    if (cmp < 0) scanning = true;
                    else if (cmp == 0 && peer_ids[o->id.peer] > peer_ids[nw->id.peer]) break;
                    else scanning = false;
                }
            }
            ck++;
            if (!scanning) { ins_b = cb; ins_k = ck; }
        }
        insert_at(ins_b, ins_k, nw);
    }
    void mark_deleted(Id id) {
        Item* x = find(id);
        if (x && !x->deleted) { x->deleted = true; x->blk->vis--; visible--; }
    }
};

struct MapSlot { bool set = false; bool has = false; Val v; u32 lamport = 0; u64 peer = 0; };

struct Replica {
    int me; int npeers;
    FList list;
    std::vector<i32> vv;            // per peer: end counter known
    std::vector<Id> frontiers;
    u64 arena = 0;                  // values arena position (adjacency model for run-merging)
    Change txn; bool open = false;
    Replica(int me_, int np) : me(me_), npeers(np), list(np), vv(np, 0) {}
};

struct DocGen {
    int np; std::vector<u64> peer_ids;
    std::vector<Replica*> reps;
    std::vector<std::vector<Change>> log;       // per peer: committed changes, counter order
    std::vector<std::map<i32, u32>> lam_index;  // per peer: change start ctr -> lamport (for dep lamports)
    Rng rng;
    DocGen(int np_, u64 seed) : np(np_), rng(seed) {
        for (int p = 0; p < np; p++) { peer_ids.push_back((Rng::splitmix(seed) | 1) + p * 2); }
        for (int p = 0; p < np; p++) reps.push_back(new Replica(p, np));
        log.resize(np);
        lam_index.resize(np);
    }
    ~DocGen() { for (auto r : reps) delete r; }
    u32 lamport_of(Id id) {
        auto& m = lam_index[id.peer];
        auto it = m.upper_bound(id.ctr);
        --it;
        return it->second + (u32)(id.ctr - it->first);
    }
    void begin(Replica& r) {
        if (r.open) return;
        r.txn = Change();
        r.txn.peer = r.me;
        r.txn.ctr = r.vv[r.me];
        r.txn.deps = r.frontiers;
        u32 l = 0;
        for (auto& d : r.frontiers) l = std::max(l, lamport_of(d) + 1);
        r.txn.lamport = l;
        r.open = true;
    }
    i32 next_ctr(Replica& r) { return r.open ? r.txn.ctr + r.txn.atoms() : r.vv[r.me]; }
    void commit(Replica& r) {
        if (!r.open) return;
        r.open = false;
        if (r.txn.ops.empty()) return;
        Change c = std::move(r.txn);
        lam_index[r.me][c.ctr] = c.lamport;
        r.vv[r.me] = c.ctr + c.atoms();
        r.frontiers.clear();
        r.frontiers.push_back(Id{r.me, r.vv[r.me] - 1});
        log[r.me].push_back(std::move(c));
    }
    Val rand_val() {
        Val v{};
        if (rng.unit() < 0.7) { v.is_str = false; v.i = (i64)(rng.next() % 2000001) - 1000000; }
        else {
            v.is_str = true;
            v.slen = (u8)rng.below(9);
            for (int i = 0; i < v.slen; i++) v.s[i] = "abcdefghijklmnopqrstuvwxyz \"\\"[rng.below(29)];
        }
        return v;
    }
    // ---- local edits
    void list_insert(Replica& r, int pos, const Val& v) {
        begin(r);
        Op op{};
        op.kind = K_LIST_INS;
        op.ctr = next_ctr(r);
        op.pos = pos;
        op.vals.push_back(v);
        op.arena_start = r.arena;
        op.arena_end = ++r.arena;
        Item* it = r.list.alloc_item();
        it->id = Id{r.me, op.ctr};
        Item* L = pos > 0 ? r.list.visible_at(pos - 1) : nullptr;
        it->ol = L ? L->id : NO_ID;
        Item* R = L ? r.list.next_of(L) : r.list.first();
        it->orr = R ? R->id : NO_ID;
        it->deleted = false;
        it->v = v;
        op.ol = it->ol;
        op.orr = it->orr;
        r.list.integrate(it, peer_ids);
        rle_push(r.txn.ops, op);
    }
    void list_delete(Replica& r, int pos, int len) {
        begin(r);
        for (int k = 0; k < len; k++) {  // one op per element at the same position (handler.rs:2744-2779)
            Item* x = r.list.visible_at(pos);
            Op op{};
            op.kind = K_LIST_DEL;
            op.ctr = next_ctr(r);
            op.pos = pos;
            op.del_start = x->id;
            op.del_len = 1;
            r.list.mark_deleted(x->id);
            rle_push(r.txn.ops, op);
        }
    }
    void map_op(Replica& r, int key, bool del, const Val& v) {
        begin(r);
        Op op{};
        op.kind = del ? K_MAP_DEL : K_MAP_SET;
        op.ctr = next_ctr(r);
        op.key = (u8)key;
        op.mapval = v;
        rle_push(r.txn.ops, op);
    }
    // ---- replication: apply every change `dst` lacks from `src`'s knowledge, in causal (lamport) order
    void pull(Replica& dst, Replica& src) {
        commit(dst);
        commit(src);
        std::vector<const Change*> todo;
        for (int p = 0; p < np; p++) {
            if (src.vv[p] <= dst.vv[p]) continue;
            for (auto& c : log[p]) if (c.ctr >= dst.vv[p] && c.ctr < src.vv[p]) todo.push_back(&c);
        }
        if (todo.empty()) return;
        std::sort(todo.begin(), todo.end(), [&](const Change* a, const Change* b) {
            return a->lamport != b->lamport ? a->lamport < b->lamport : a->peer < b->peer; });
        for (const Change* c : todo) {
            for (const Op& op : c->ops) {
                if (op.kind == K_LIST_INS) {
                    for (size_t k = 0; k < op.vals.size(); k++) {
                        Item* it = dst.list.alloc_item();
                        it->id = Id{c->peer, op.ctr + (i32)k};
                        it->ol = k == 0 ? op.ol : Id{c->peer, op.ctr + (i32)k - 1};
                        it->orr = op.orr;
                        it->deleted = false;
                        it->v = op.vals[k];
                        dst.list.integrate(it, peer_ids);
                    }
                    dst.arena += op.vals.size();  // decode allocates in the importer's arena
                } else if (op.kind == K_LIST_DEL) {
                    int n = op.del_len < 0 ? -op.del_len : op.del_len;
                    for (int k = 0; k < n; k++) dst.list.mark_deleted(Id{op.del_start.peer, op.del_start.ctr + k});
                }
            }
            dst.vv[c->peer] = c->ctr + c->atoms();
            // frontiers: drop deps, add last id
            for (auto& d : c->deps)
                dst.frontiers.erase(std::remove(dst.frontiers.begin(), dst.frontiers.end(), d), dst.frontiers.end());
            dst.frontiers.push_back(Id{c->peer, c->ctr + c->atoms() - 1});
        }
    }
    // one user action; returns the number of atom ops it produced
    int random_op(Replica& r, int max_atoms = 4) {
        double x = rng.unit();
        int n = r.list.visible;
        if (x < 0.60 || (x < 0.75 && n == 0)) list_insert(r, (int)rng.below((u32)n + 1), rand_val());
        else if (x < 0.75) {
            int pos = (int)rng.below((u32)n);
            int len = std::min<int>(std::min<int>(1 + (int)rng.below(4), n - pos), max_atoms);
            list_delete(r, pos, len);
            return len;
        } else {
            Val v{};
            v.is_str = false;
            v.i = (i64)rng.below(100000);
            map_op(r, (int)rng.below(16), rng.unit() < 0.10, v);
        }
        return 1;
    }
};

// ------------------------------------------------------------------ encoding of one peer's changes into blocks
template <class T>
struct Reg { std::vector<T> v; size_t reg(const T& x) { for (size_t i = 0; i < v.size(); i++) if (v[i] == x) return i; v.push_back(x); return v.size() - 1; } };

void write_val(W& w, const Val& v) {
    if (v.is_str) { w.u8_(5); w.varint(v.slen); w.bytes((const u8*)v.s, v.slen); }
    else { w.u8_(3); w.sleb(v.i); }
}

std::vector<u8> encode_block(const std::vector<Change>& blk, const std::vector<u64>& peer_ids) {
    Reg<u64> peers;
    Reg<std::string> keys;
    Reg<int> cids;  // 0 = List "list", 1 = Map "map", 2 = Tree "tree"
    peers.reg(peer_ids[blk[0].peer]);
    // position register pre-filled in sorted order (reference: block_encode.rs:156-178)
    std::vector<std::string> positions;
    for (auto& c : blk)
        for (auto& op : c.ops)
            if (op.kind == K_TREE && op.parent.peer != -2) positions.push_back(op.position);
    std::sort(positions.begin(), positions.end());
    positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
    const u64 DEL_PEER = ~0ull;
    std::vector<i64> c_cidx, c_prop, d_peer, d_ctr, d_len;
    std::vector<u64> c_vt, c_len;
    W vw;
    for (auto& c : blk)
        for (auto& op : c.ops) {
            int cid = op.kind == K_TREE ? 2 : (op.kind == K_TEXT_INS || op.kind == K_TEXT_DEL) ? 3 : (op.kind == K_LIST_INS || op.kind == K_LIST_DEL) ? 0 : 1;
            size_t ci = cids.reg(cid);
            i64 prop = op.kind == K_TREE ? 0 : op.pos;
            u64 vt;
            if (cid == 1) {   // map ops: the key
                char kb[8];
                int n = snprintf(kb, sizeof kb, "k%d", (int)op.key);
                prop = (i64)keys.reg(std::string(kb, n));
            }
            switch (op.kind) {
                case K_LIST_INS:
                    vt = 11; vw.u8_(7); vw.varint(op.vals.size());
                    for (size_t q = 0; q < op.vals.size(); q++) write_val(vw, op.vals[q]);
                    break;
                case K_TEXT_INS:
                    vt = 5; vw.varint(op.text.size()); vw.bytes((const u8*)op.text.data(), op.text.size());
                    break;
                case K_LIST_DEL: case K_TEXT_DEL:
                    vt = 9;
                    d_peer.push_back((i64)peers.reg(peer_ids[op.del_start.peer]));
                    d_ctr.push_back(op.del_start.ctr);
                    d_len.push_back(op.del_len);
                    break;
                case K_MAP_SET: vt = 11; write_val(vw, op.mapval); break;
                case K_TREE: {   // RawTreeMove (reference: encoding/value.rs:969-989, block_encode.rs:324-362)
                    vt = 16;
                    vw.varint(peers.reg(peer_ids[op.target.peer]));
                    vw.varint((u32)op.target.ctr);
                    if (op.parent.peer == -2) {
                        size_t pp = peers.reg(DEL_PEER);
                        vw.varint(0); vw.u8_(0); vw.varint(pp); vw.varint(0x7FFFFFFFu);
                    } else {
                        size_t pp = op.parent.peer >= 0 ? peers.reg(peer_ids[op.parent.peer]) : 0;
                        size_t pi = std::lower_bound(positions.begin(), positions.end(), op.position) - positions.begin();
                        vw.varint(pi);
                        vw.u8_(op.parent.peer >= 0 ? 0 : 1);
                        if (op.parent.peer >= 0) { vw.varint(pp); vw.varint((u32)op.parent.ctr); }
                    }
                    break;
                }
                default: vt = 8;
            }
            c_cidx.push_back((i64)ci);
            c_prop.push_back(prop);
            c_vt.push_back(vt);
            c_len.push_back((u64)op.atoms());
        }
    // container arena: roots register their names as keys
    W cw;
    cw.varint(cids.v.size());
    for (int cid : cids.v) {
        cw.varint(4); cw.u8_(1); cw.u8_(cid == 0 ? 1 : cid == 1 ? 0 : cid == 2 ? 3 : 2); cw.varint(0);
        cw.zig((i64)keys.reg(cid == 0 ? "list" : cid == 1 ? "map" : cid == 2 ? "tree" : "text"));
    }
    W kw;
    for (auto& k : keys.v) { kw.varint(k.size()); kw.bytes((const u8*)k.data(), k.size()); }
    W ow;
    {
        W a, b, c, d;
        delta_rle(a, c_cidx); delta_rle(b, c_prop);
        AnyRle<u64, WrU8> e1(c, WrU8()); for (auto x : c_vt) e1.push(x); e1.finish();
        AnyRle<u64, WrVar> e2(d, WrVar()); for (auto x : c_len) e2.push(x); e2.finish();
        ow.varint(1); ow.varint(4); ow.lenbytes(a.b); ow.lenbytes(b.b); ow.lenbytes(c.b); ow.lenbytes(d.b);
    }
    W pw;   // PositionArena (reference: encoding/arena.rs:159-224): common prefix lengths + rests
    if (!positions.empty()) {
        W c0, c1;
        AnyRle<u64, WrVar> e(c0, WrVar());
        c1.varint(positions.size());
        const std::string* last = nullptr;
        for (auto& p : positions) {
            size_t common = 0;
            if (last) while (common < last->size() && common < p.size() && (*last)[common] == p[common]) common++;
            e.push(common);
            c1.varint(p.size() - common);
            c1.bytes((const u8*)p.data() + common, p.size() - common);
            last = &p;
        }
        e.finish();
        pw.varint(1); pw.varint(2); pw.lenbytes(c0.b); pw.lenbytes(c1.b);
    }
    W dw;
    if (!d_peer.empty()) {
        W a, b, c;
        delta_rle(a, d_peer); delta_rle(b, d_ctr); delta_rle(c, d_len);
        dw.varint(1); dw.varint(3); dw.lenbytes(a.b); dw.lenbytes(b.b); dw.lenbytes(c.b);
    }
    // header + meta (reference: block_meta_encode.rs:13-88)
    W lens, dsw, dlw, dpw, dcw, lw, tw, mw;
    {
        std::vector<bool> dep_self; std::vector<i64> dep_ctrs, lams, tss;
        AnyRle<u64, WrVar> dl(dlw, WrVar()), dp(dpw, WrVar()), ml(mw, WrVar());
        for (size_t i = 0; i < blk.size(); i++) {
            const Change& c = blk[i];
            if (i + 1 < blk.size()) { lens.varint((u64)c.atoms()); lams.push_back(c.lamport); }
            tss.push_back(0);
            ml.push(0);
            bool ds = false; u64 others = 0;
            for (auto& d : c.deps) {
                if (d.peer == c.peer) ds = true;
                else { dp.push(peers.reg(peer_ids[d.peer])); dep_ctrs.push_back(d.ctr); others++; }
            }
            dep_self.push_back(ds);
            dl.push(others);
        }
        dl.finish(); dp.finish(); ml.finish();
        bool_rle(dsw, dep_self); dod(dcw, dep_ctrs); dod(lw, lams); dod(tw, tss);
    }
    W hw;
    hw.varint(peers.v.size());
    for (auto p : peers.v) for (int k = 0; k < 8; k++) hw.u8_((u8)(p >> (8 * k)));
    hw.bytes(lens.b); hw.bytes(dsw.b); hw.bytes(dlw.b); hw.bytes(dpw.b); hw.bytes(dcw.b); hw.bytes(lw.b);
    W meta;
    meta.bytes(tw.b); meta.bytes(mw.b);
    W out;
    const Change& f = blk.front(); const Change& l = blk.back();
    out.varint((u32)f.ctr);
    out.varint((u32)(l.ctr + l.atoms() - f.ctr));
    out.varint(f.lamport);
    out.varint(l.lamport + (u32)l.atoms() - f.lamport);
    out.varint(blk.size());
    std::vector<u8> empty;
    out.lenbytes(hw.b); out.lenbytes(meta.b); out.lenbytes(cw.b); out.lenbytes(kw.b); out.lenbytes(pw.b);
    out.lenbytes(ow.b); out.lenbytes(dw.b); out.lenbytes(vw.b);
    return out.b;
}

// change merge + block packing (reference: change_store.rs:711-764,1244-1291 ; change.rs:268-283)
std::vector<u8> export_all(DocGen& g) {  // consumes g.log
    W body;
    for (int p = 0; p < g.np; p++) {
        if (g.log[p].empty()) continue;
        // blocks are emitted in (peer id, counter) order: collect then sort peers by real id below
    }
    std::vector<int> order(g.np);
    for (int p = 0; p < g.np; p++) order[p] = p;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return g.peer_ids[a] < g.peer_ids[b]; });
    for (int p : order) {
        std::vector<Change> blk;
        size_t est = 0;
        auto flush = [&]() {
            if (blk.empty()) return;
            std::vector<u8> bytes = encode_block(blk, g.peer_ids);
            body.varint(bytes.size());
            body.bytes(bytes);
            blk.clear();
            est = 0;
        };
        for (Change& c0 : g.log[p]) {
            Change c = std::move(c0);
            size_t new_size = change_estimate(c);
            if (blk.empty()) { est = new_size; blk.push_back(std::move(c)); continue; }
            bool is_full = new_size + est > 4096;
            Change& last = blk.back();
            bool can_merge = c.deps.size() == 1 && c.deps[0].peer == c.peer && c.ctr == last.ctr + last.atoms();
            if (can_merge && (!is_full || (c.ops.size() == 1 && op_mergable(last.ops.back(), c.ops[0])))) {
                for (auto& op : c.ops) { size_t s = op_estimate(op); if (!rle_push(last.ops, op)) est += s; }
            } else if (is_full) {
                flush();
                est = new_size;
                blk.push_back(std::move(c));
            } else {
                est += new_size;
                blk.push_back(std::move(c));
            }
        }
        flush();
    }
    std::vector<u8> out(22 + body.b.size(), 0);
    memcpy(out.data(), "loro", 4);
    out[20] = 0; out[21] = 4;
    memcpy(out.data() + 22, body.b.data(), body.b.size());
    u32 h = xxh32(out.data() + 20, out.size() - 20, 0x4F524F4Cu);
    out[16] = (u8)h; out[17] = (u8)(h >> 8); out[18] = (u8)(h >> 16); out[19] = (u8)(h >> 24);
    return out;
}

void json_escape(std::string& o, const char* s, size_t n) {
    static const char* hex = "0123456789abcdef";
    o.push_back('"');
    for (size_t i = 0; i < n; i++) {
        unsigned char c = (unsigned char)s[i];
        if (c == '"') o += "\\\""; else if (c == '\\') o += "\\\\";
        else if (c < 0x20) { o += "\\u00"; o.push_back(hex[c >> 4]); o.push_back(hex[c & 15]); }
        else o.push_back((char)c);
    }
    o.push_back('"');
}
void json_val(std::string& o, const Val& v) { if (v.is_str) json_escape(o, v.s, v.slen); else o += std::to_string(v.i); }

// expected deep value (keys sorted) from the generator's own converged replica + LWW over the op log
std::string expected_json(DocGen& g) {
    std::string o = "{\"list\":[";
    bool first = true;
    for (auto b : g.reps[0]->list.blocks)
        for (auto x : b->it) if (!x->deleted) { if (!first) o.push_back(','); first = false; json_val(o, x->v); }
    o += "],\"map\":{";
    MapSlot slots[16];
    for (int p = 0; p < g.np; p++)
        for (auto& c : g.log[p])
            for (auto& op : c.ops) {
                if (op.kind != K_MAP_SET && op.kind != K_MAP_DEL) continue;
                u32 lam = c.lamport + (u32)(op.ctr - c.ctr);
                MapSlot& s = slots[op.key];
                if (!s.set || lam > s.lamport || (lam == s.lamport && g.peer_ids[p] > s.peer)) {
                    s.set = true; s.has = op.kind == K_MAP_SET; s.v = op.mapval; s.lamport = lam; s.peer = g.peer_ids[p];
                }
            }
    std::vector<std::pair<std::string, int>> ks;
    for (int k = 0; k < 16; k++) if (slots[k].set && slots[k].has) ks.push_back({"k" + std::to_string(k), k});
    std::sort(ks.begin(), ks.end());
    first = true;
    for (auto& kv : ks) {
        if (!first) o.push_back(',');
        first = false;
        json_escape(o, kv.first.data(), kv.first.size());
        o.push_back(':');
        json_val(o, slots[kv.second].v);
    }
    o += "}}";
    return o;
}

struct DocOut { std::vector<u8> blob; std::string json; u64 atoms; };

// SURVEY.md 8d, config C3
DocOut gen_c3(u64 seed, int n_ops, int n_peers, int prefix_ops, int sync_every, int txn_ops, bool want_json) {
    DocGen g(n_peers, seed);
    int done = 0;
    Replica& r0 = *g.reps[0];
    int since_commit = 0;
    while (done < prefix_ops && done < n_ops) {
        done += g.random_op(r0, std::min(prefix_ops, n_ops) - done);
        if (++since_commit >= txn_ops) { g.commit(r0); since_commit = 0; }
    }
    g.commit(r0);
    for (int p = 1; p < n_peers; p++) g.pull(*g.reps[p], r0);
    int since_sync = 0;
    while (done < n_ops) {
        int p = (int)g.rng.below((u32)n_peers);
        Replica& r = *g.reps[p];
        // a burst of edits by one peer, like a user typing
        int burst = 1 + (int)g.rng.below(20);
        for (int k = 0; k < burst && done < n_ops; k++) {
            int a = g.random_op(r, n_ops - done);
            done += a;
            since_sync += a;
            if (g.rng.below((u32)txn_ops) == 0) g.commit(r);
        }
        if (since_sync >= sync_every && n_peers > 1) {
            since_sync = 0;
            int a = (int)g.rng.below((u32)n_peers), b = (int)g.rng.below((u32)n_peers - 1);
            if (b >= a) b++;
            g.pull(*g.reps[a], *g.reps[b]);
            g.pull(*g.reps[b], *g.reps[a]);
        }
    }
    for (int round = 0; round < 2; round++)
        for (int a = 0; a < n_peers; a++)
            for (int b = 0; b < n_peers; b++) if (a != b) g.pull(*g.reps[a], *g.reps[b]);
    DocOut out;
    if (want_json) out.json = expected_json(g);
    out.blob = export_all(g);
    out.atoms = 0;
    for (int p = 0; p < n_peers; p++) out.atoms += (u64)g.reps[0]->vv[p];
    return out;
}

// ------------------------------------------------------------------ config C5: movable trees
// fractional index, jitter 0 (reference: crates/fractional_index/src/lib.rs:52-127); strings carry the terminator
const u8 FI_TERM = 128;
std::string fi_after(const std::string& b) {
    for (size_t i = 0; i < b.size(); i++) {
        u8 c = (u8)b[i];
        if (c < FI_TERM) return b.substr(0, i);
        if (c < 255) { std::string a = b.substr(0, i + 1); a[i] = (char)(c + 1); return a; }
    }
    return b;
}
std::string fi_append_after(const std::string* last) {   // FractionalIndex::new(Some(last), None) / default
    if (!last) return std::string(1, (char)FI_TERM);
    std::string a = fi_after(*last);
    a.push_back((char)FI_TERM);
    return a;
}
struct TNode { int parent = -3; std::string pos; u32 lamport = 0; int peer = 0; };   // parent: node, -1 root, -3 unexist
struct TreeRep {
    std::vector<TNode> nodes;                 // indexed by the create op's counter (all creates are peer 0's)
    std::vector<std::vector<int>> kids;       // per node, sibling order; kids.back() = the root list
    const std::vector<u64>* peer_ids = nullptr;
    explicit TreeRep(size_t n) : nodes(n), kids(n + 1) {}
    std::vector<int>& list_of(int parent) { return parent < 0 ? kids.back() : kids[(size_t)parent]; }
    bool before(int a, int b) const {
        const TNode &x = nodes[(size_t)a], &y = nodes[(size_t)b];
        if (x.pos != y.pos) return x.pos < y.pos;
        if (x.lamport != y.lamport) return x.lamport < y.lamport;
        return (*peer_ids)[(size_t)x.peer] < (*peer_ids)[(size_t)y.peer];
    }
    bool is_ancestor(int anc, int node) const {   // anc is node or one of its ancestors
        if (nodes[(size_t)anc].parent == -3) return false;
        while (node >= 0) {
            if (node == anc) return true;
            node = nodes[(size_t)node].parent;
        }
        return false;
    }
    void place(int t, int parent, const std::string& pos, u32 lamport, int peer) {
        TNode& n = nodes[(size_t)t];
        if (n.parent != -3) { auto& old = list_of(n.parent); old.erase(std::find(old.begin(), old.end(), t)); }
        n.parent = parent; n.pos = pos; n.lamport = lamport; n.peer = peer;
        auto& l = list_of(parent);
        size_t i = 0;
        while (i < l.size() && before(l[i], t)) i++;
        l.insert(l.begin() + (long)i, t);
    }
};
void tree_json(std::string& o, const TreeRep& tr, int parent, const std::vector<u64>& peer_ids) {
    static const char* HEX = "0123456789ABCDEF";
    const auto& l = parent < 0 ? tr.kids.back() : tr.kids[(size_t)parent];
    o.push_back('[');
    for (size_t i = 0; i < l.size(); i++) {
        if (i) o.push_back(',');
        const TNode& n = tr.nodes[(size_t)l[i]];
        o += "{\"children\":";
        tree_json(o, tr, l[i], peer_ids);
        o += ",\"fractional_index\":\"";
        for (unsigned char c : n.pos) { o.push_back(HEX[c >> 4]); o.push_back(HEX[c & 15]); }
        o += "\",\"id\":\"" + std::to_string(l[i]) + "@" + std::to_string(peer_ids[0]) + "\",\"index\":" + std::to_string(i) + ",\"meta\":{},\"parent\":";
        if (parent < 0) o += "null"; else o += "\"" + std::to_string(parent) + "@" + std::to_string(peer_ids[0]) + "\"";
        o.push_back('}');
    }
    o.push_back(']');
}

// peer 0 builds a tree of n_nodes (fan-out <= max_fanout), every peer starts from it and
// issues n_moves moves concurrently (random target, random new parent; cycles between peers' moves are left in --
// the merge has to resolve them).  One FastUpdates blob = export(all_updates) of a fully synced replica.
DocOut gen_c5(u64 seed, int n_nodes, int n_peers, int n_moves, int max_fanout, int txn_ops, bool want_json) {
    DocGen g(n_peers, seed);
    std::vector<TreeRep> trs((size_t)n_peers, TreeRep((size_t)n_nodes));
    for (auto& t : trs) t.peer_ids = &g.peer_ids;
    auto tree_op = [&](int p, int target, int parent) {
        Replica& r = *g.reps[p];
        TreeRep& tr = trs[(size_t)p];
        g.begin(r);
        Op op{};
        op.kind = K_TREE;
        op.ctr = g.next_ctr(r);
        op.pos = 0;
        op.target = Id{0, target};
        op.parent = parent < 0 ? Id{-1, 0} : Id{0, parent};
        // handler/tree.rs:548-566 `mov`: append after the last child (the target itself not counted)
        const std::string* last = nullptr;
        for (int k : tr.list_of(parent)) if (k != target) last = &tr.nodes[(size_t)k].pos;
        op.position = fi_append_after(last);
        u32 lam = r.txn.lamport + (u32)(op.ctr - r.txn.ctr);
        tr.place(target, parent, op.position, lam, p);
        r.txn.ops.push_back(op);
    };
    // base tree by peer 0: node k is created by the op with counter k
    int since = 0;
    for (int k = 0; k < n_nodes; k++) {
        int parent = -1;
        if (k > 0 && g.rng.unit() < 0.9) {
            for (int tries = 0; tries < 8; tries++) {
                int c = (int)g.rng.below((u32)k);
                if ((int)trs[0].kids[(size_t)c].size() < max_fanout) { parent = c; break; }
            }
        }
        if (parent < 0 && (int)trs[0].kids.back().size() >= max_fanout && k > 0) {
            // the root list is full too: first node with room
            for (int c = 0; c < k; c++) if ((int)trs[0].kids[(size_t)c].size() < max_fanout) { parent = c; break; }
        }
        tree_op(0, k, parent);
        if (++since >= txn_ops) { g.commit(*g.reps[0]); since = 0; }
    }
    g.commit(*g.reps[0]);
    for (int p = 1; p < n_peers; p++) {
        trs[(size_t)p] = trs[0];
        trs[(size_t)p].peer_ids = &g.peer_ids;
        Replica& r = *g.reps[p];
        r.vv = g.reps[0]->vv;
        r.frontiers = g.reps[0]->frontiers;
    }
    // concurrent moves
    for (int p = 0; p < n_peers; p++) {
        since = 0;
        for (int m = 0; m < n_moves; m++) {
            int target = 0, parent = -1;
            for (int tries = 0; tries < 16; tries++) {
                target = (int)g.rng.below((u32)n_nodes);
                parent = g.rng.unit() < 0.05 ? -1 : (int)g.rng.below((u32)n_nodes);
                if (parent < 0 || !trs[(size_t)p].is_ancestor(target, parent)) break;   // a local cycle is an API error
                parent = -1;
            }
            tree_op(p, target, parent);
            if (++since >= txn_ops) { g.commit(*g.reps[p]); since = 0; }
        }
        g.commit(*g.reps[p]);
    }
    DocOut out;
    out.atoms = 0;
    for (int p = 0; p < n_peers; p++) out.atoms += (u64)g.reps[p]->vv[p];
    if (want_json) {
        // the merge, independently: every op in (lamport, peer id) order, a move whose new parent sits below the
        // target is skipped (reference: diff_calc/tree.rs:445-508)
        struct M { u32 lam; u64 pid; int peer; const Op* op; };
        std::vector<M> all;
        for (int p = 0; p < n_peers; p++)
            for (auto& c : g.log[p])
                for (auto& op : c.ops) all.push_back(M{c.lamport + (u32)(op.ctr - c.ctr), g.peer_ids[p], p, &op});
        std::sort(all.begin(), all.end(), [](const M& a, const M& b) { return a.lam != b.lam ? a.lam < b.lam : a.pid < b.pid; });
        TreeRep fin((size_t)n_nodes);
        fin.peer_ids = &g.peer_ids;
        for (auto& m : all) {
            int parent = m.op->parent.peer < 0 ? -1 : m.op->parent.ctr;
            if (parent >= 0 && fin.is_ancestor(m.op->target.ctr, parent)) continue;
            fin.place(m.op->target.ctr, parent, m.op->position, m.lam, m.peer);
        }
        out.json = "{\"tree\":";
        tree_json(out.json, fin, -1, g.peer_ids);
        out.json += "}";
    }
    out.blob = export_all(g);
    return out;
}

// ------------------------------------------------------------------ config C4: one rich-text document, many peers
// A replica's view of the text: the visible characters as runs of consecutive ids, in blocks (positions are looked
// up by scanning block totals).  Peers never sync in this config, so a replica holds the base plus its own edits and
// deleted characters simply leave the view.
struct TRun { int peer; i32 ctr; i32 len; };
struct TextView {
    std::vector<std::vector<TRun>> blocks;
    std::vector<i32> tot;
    i32 total = 0;
    void init(int peer, i32 len) { blocks.assign(1, {TRun{peer, 0, len}}); tot.assign(1, len); total = len; }
    void rebalance(size_t b) {
        if (blocks[b].size() <= 128) return;
        size_t half = blocks[b].size() / 2;
        std::vector<TRun> hi(blocks[b].begin() + (long)half, blocks[b].end());
        blocks[b].resize(half);
        i32 t = 0;
        for (auto& r : hi) t += r.len;
        blocks.insert(blocks.begin() + (long)b + 1, std::move(hi));
        tot.insert(tot.begin() + (long)b + 1, t);
        tot[b] -= t;
    }
    // (block, run, offset) of visible position pos (the run that contains character pos; end = past the last run)
    void locate(i32 pos, size_t* b, size_t* r, i32* off) const {
        size_t bi = 0;
        while (bi + 1 < blocks.size() && pos >= tot[bi]) { pos -= tot[bi]; bi++; }
        size_t ri = 0;
        while (ri < blocks[bi].size() && pos >= blocks[bi][ri].len) { pos -= blocks[bi][ri].len; ri++; }
        *b = bi; *r = ri; *off = pos;
    }
    void insert(i32 pos, TRun nw) {
        size_t b, r; i32 off;
        locate(pos, &b, &r, &off);
        auto& v = blocks[b];
        if (r < v.size() && off > 0) {   // split the run under the cursor
            TRun tail{v[r].peer, v[r].ctr + off, v[r].len - off};
            v[r].len = off;
            v.insert(v.begin() + (long)r + 1, tail);
            r++;
        }
        v.insert(v.begin() + (long)r, nw);
        tot[b] += nw.len;
        total += nw.len;
        rebalance(b);
    }
    // remove [pos, pos + len): the id spans that disappear, left to right
    void erase(i32 pos, i32 len, std::vector<TRun>* gone) {
        while (len > 0) {
            size_t b, r; i32 off;
            locate(pos, &b, &r, &off);
            auto& v = blocks[b];
            TRun& x = v[r];
            i32 take = std::min(len, x.len - off);
            gone->push_back(TRun{x.peer, x.ctr + off, take});
            if (off == 0 && take == x.len) v.erase(v.begin() + (long)r);
            else if (off == 0) { x.ctr += take; x.len -= take; }
            else if (off + take == x.len) x.len = off;
            else {
                TRun tail{x.peer, x.ctr + off + take, x.len - off - take};
                x.len = off;
                v.insert(v.begin() + (long)r + 1, tail);
            }
            tot[b] -= take;
            total -= take;
            len -= take;
            if (v.empty() && blocks.size() > 1) { blocks.erase(blocks.begin() + (long)b); tot.erase(tot.begin() + (long)b); }
        }
    }
};

// peer 0 inserts `base_chars` ASCII characters (one op), then `n_peers` other peers each make
// `edits` edits on their own copy of that base (70 % insert of 1-8 characters, 30 % delete of 1-8), never syncing:
// one document whose merge has n_peers fully concurrent branches.  Deletes are emitted the way the text handler does
// (handler.rs:1897-1957): one op per run of consecutive ids, rightmost first.
DocOut gen_c4(u64 seed, int base_chars, int n_peers, int edits, int txn_ops) {
    const int np = n_peers + 1;
    DocGen g(np, seed);
    {   // the base, one change of one op
        Replica& r0 = *g.reps[0];
        g.begin(r0);
        Op op{};
        op.kind = K_TEXT_INS;
        op.ctr = 0;
        op.pos = 0;
        op.text.resize((size_t)base_chars);
        for (int i = 0; i < base_chars; i++) op.text[(size_t)i] = "abcdefghijklmnopqrstuvwxyz     .,\n"[g.rng.below(35)];
        op.arena_start = 0;
        op.arena_end = (u64)base_chars;
        r0.txn.ops.push_back(op);
        g.commit(r0);
    }
    std::vector<std::thread> ts;
    std::vector<std::vector<Change>> logs((size_t)np);
    std::atomic<int> next(1);
    const std::vector<Id> base_frontier = g.reps[0]->frontiers;
    const u32 base_lamport_end = (u32)base_chars;
    auto work = [&]() {
        while (true) {
            int p = next.fetch_add(1);
            if (p >= np) break;
            Rng rng(seed * 1000003ull + (u64)p);
            TextView view;
            view.init(0, base_chars);
            std::vector<Change>& log = logs[(size_t)p];
            Change txn;
            bool open = false;
            i32 ctr = 0;
            u32 lamport = base_lamport_end;   // every change of this peer follows the base (and its own predecessor)
            u64 arena = 0, cap = 0;
            u32 gen = 0;
            int since = 0;
            auto begin = [&]() {
                if (open) return;
                txn = Change();
                txn.peer = p;
                txn.ctr = ctr;
                txn.lamport = lamport;
                if (ctr == 0) txn.deps = base_frontier; else txn.deps = {Id{p, ctr - 1}};
                open = true;
            };
            auto commit = [&]() {
                if (!open) return;
                open = false;
                if (txn.ops.empty()) return;
                lamport += (u32)txn.atoms();
                log.push_back(std::move(txn));
            };
            for (int e = 0; e < edits; e++) {
                begin();
                if (rng.unit() < 0.7 || view.total < 8) {
                    int len = 1 + (int)rng.below(8);
                    Op op{};
                    op.kind = K_TEXT_INS;
                    op.ctr = ctr;
                    op.pos = (i32)rng.below((u32)view.total + 1);
                    op.text.resize((size_t)len);
                    for (int i = 0; i < len; i++) op.text[(size_t)i] = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"[rng.below(26)];
                    // local string arena: contiguous allocations, the buffer doubles from 32 (same model as the checker)
                    op.arena_start = arena;
                    if (arena + (u64)len > cap) { u64 nc = std::max<u64>(cap * 2, 32); while (nc < arena + (u64)len) nc *= 2; cap = nc; gen++; }
                    arena += (u64)len;
                    op.arena_end = arena;
                    op.arena_gen = gen;
                    view.insert(op.pos, TRun{p, ctr, len});
                    ctr += len;
                    rle_push(txn.ops, op);
                } else {
                    int len = 1 + (int)rng.below(8);
                    i32 pos = (i32)rng.below((u32)(view.total - len + 1));
                    std::vector<TRun> gone;
                    view.erase(pos, len, &gone);
                    i32 end = pos + len;
                    for (size_t k = gone.size(); k-- > 0;) {   // rightmost run first
                        Op op{};
                        op.kind = K_TEXT_DEL;
                        op.ctr = ctr;
                        op.pos = end - gone[k].len;
                        op.del_start = Id{gone[k].peer, gone[k].ctr};
                        op.del_len = gone[k].len;
                        end -= gone[k].len;
                        ctr += gone[k].len;
                        rle_push(txn.ops, op);
                    }
                }
                if (++since >= txn_ops) { commit(); since = 0; }
            }
            commit();
        }
    };
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    for (unsigned t = 1; t < hw && t < (unsigned)np; t++) ts.emplace_back(work);
    work();
    for (auto& t : ts) t.join();
    DocOut out;
    out.atoms = (u64)base_chars;
    for (int p = 1; p < np; p++) {
        for (auto& c : logs[(size_t)p]) out.atoms += (u64)c.atoms();
        g.log[(size_t)p] = std::move(logs[(size_t)p]);
    }
    out.blob = export_all(g);
    return out;
}

}  // namespace

extern "C" {

struct lw_batch {
    std::vector<u8> bytes;           // blobs at 16-byte aligned starts
    std::vector<u64> offs;
    std::vector<u32> lens;
    std::vector<std::string> json;
    u64 atoms = 0;
};

// Generate docs [first_doc, first_doc + n_docs) of config C3 (seed = doc index + seed_base).
lw_batch* lw_generate_c3(u64 seed_base, u64 first_doc, u64 n_docs, int n_ops, int n_peers, int prefix_ops,
                         int sync_every, int txn_ops, int want_json, int threads) {
    std::vector<DocOut> docs(n_docs);
    std::atomic<u64> next(0);
    auto work = [&]() {
        while (true) {
            u64 i = next.fetch_add(1);
            if (i >= n_docs) break;
            docs[i] = gen_c3(seed_base + first_doc + i, n_ops, n_peers, prefix_ops, sync_every, txn_ops, want_json != 0);
        }
    };
    if (threads < 1) threads = 1;
    std::vector<std::thread> ts;
    for (int t = 1; t < threads; t++) ts.emplace_back(work);
    work();
    for (auto& t : ts) t.join();
    lw_batch* b = new lw_batch();
    u64 total = 0;
    for (auto& d : docs) { b->offs.push_back(total); b->lens.push_back((u32)d.blob.size()); total += (d.blob.size() + 15) & ~(u64)15; b->atoms += d.atoms; }
    b->bytes.assign(total + 64, 0);
    for (size_t i = 0; i < docs.size(); i++) {
        memcpy(b->bytes.data() + b->offs[i], docs[i].blob.data(), docs[i].blob.size());
        if (want_json) b->json.push_back(std::move(docs[i].json));
    }
    return b;
}
// Generate docs [first_doc, first_doc + n_docs) of config C5 (seed = doc index + seed_base).
lw_batch* lw_generate_c5(u64 seed_base, u64 first_doc, u64 n_docs, int n_nodes, int n_peers, int n_moves, int max_fanout,
                         int txn_ops, int want_json, int threads) {
    std::vector<DocOut> docs(n_docs);
    std::atomic<u64> next(0);
    auto work = [&]() {
        while (true) {
            u64 i = next.fetch_add(1);
            if (i >= n_docs) break;
            docs[i] = gen_c5(seed_base + first_doc + i, n_nodes, n_peers, n_moves, max_fanout, txn_ops, want_json != 0);
        }
    };
    if (threads < 1) threads = 1;
    std::vector<std::thread> ts;
    for (int t = 1; t < threads; t++) ts.emplace_back(work);
    work();
    for (auto& t : ts) t.join();
    lw_batch* b = new lw_batch();
    u64 total = 0;
    for (auto& d : docs) { b->offs.push_back(total); b->lens.push_back((u32)d.blob.size()); total += (d.blob.size() + 15) & ~(u64)15; b->atoms += d.atoms; }
    b->bytes.assign(total + 64, 0);
    for (size_t i = 0; i < docs.size(); i++) {
        memcpy(b->bytes.data() + b->offs[i], docs[i].blob.data(), docs[i].blob.size());
        if (want_json) b->json.push_back(std::move(docs[i].json));
    }
    return b;
}
// Config C4: ONE document (seed picks it).
lw_batch* lw_generate_c4(u64 seed, int base_chars, int n_peers, int edits, int txn_ops) {
    DocOut d = gen_c4(seed, base_chars, n_peers, edits, txn_ops);
    lw_batch* b = new lw_batch();
    b->offs.push_back(0);
    b->lens.push_back((u32)d.blob.size());
    b->atoms = d.atoms;
    b->bytes.assign(((d.blob.size() + 15) & ~(size_t)15) + 64, 0);
    memcpy(b->bytes.data(), d.blob.data(), d.blob.size());
    return b;
}
const u8* lw_bytes(lw_batch* b, u64* total) { *total = b->bytes.size(); return b->bytes.data(); }
const u64* lw_offsets(lw_batch* b) { return b->offs.data(); }
const u32* lw_lens(lw_batch* b) { return b->lens.data(); }
u64 lw_atoms(lw_batch* b) { return b->atoms; }
const char* lw_json(lw_batch* b, u64 i, u64* len) { *len = b->json[i].size(); return b->json[i].data(); }
void lw_free(lw_batch* b) { delete b; }
}
