// ORACLE (test infrastructure) -- a minimal CPU restatement of LoroDoc's import/merge/export path.
// Never linked into the product; only tests/, smoke() and bench.py's CPU-baseline legs use it.
//
// Restates (file:line under /root/reference/crates/loro-internal/src unless noted):
//   oplog import            encoding/outdated_encode_reordered.rs:40-83, oplog.rs:130-196,
//                           oplog/pending_changes.rs:31-140, oplog/loro_dag.rs:935-954
//   change store / packing  oplog/change_store.rs:494-576 (export_blocks_from), :711-764 (insert_change),
//                           :913-1000 (split), :1244-1291 (push_change); change.rs:128-139,203-283
//   op run-merge rules      container/list/list_op.rs:189-278,381-434,516-658; op.rs:143-172
//   eg-walker (Fugue)       container/richtext/tracker.rs:84-232,330-526 ; tracker/crdt_rope.rs:43-315,
//                           484-640 ; fugue_span.rs:192-386 ; diff_calc.rs:140-236,585-620
//   map LWW                 diff_calc.rs:423-551 ; delta/map_delta.rs:19-46
//   tree                    diff_calc/tree.rs:21-148,445-508 ; state/tree_state.rs:592-595,1229-1275,1424-1452
//   deep value / JSON       state.rs:894-924,1039 ; loro-common/src/value.rs:692-711
// Simplifications (do not change results): state is recomputed by replaying the whole oplog from the
// empty version (the reference replays from the LCA; final state is a function of the oplog), spans in
// the tracker are never re-merged (FugueSpan::can_merge is lossless), pending changes are kept in a flat
// list and retried to a fixpoint.
#pragma once
#include <algorithm>
#include <cassert>
#include <charconv>
#include <cmath>
#include <functional>
#include <list>
#include "block.hpp"

namespace lo {

static const size_t MAX_BLOCK_SIZE = 4096;  // change_store.rs:37

// ------------------------------------------------------------------ size estimates + merge rules
inline size_t value_estimate(const Value& v);
inline size_t op_estimate(const Op& op, uint8_t ctype) {
    switch (op.kind) {  // list_op.rs:109-123, op/content.rs:70-77
        case OP_LIST_INSERT: return 4 * (size_t)op.atom_len();
        case OP_TEXT_INSERT: return op.text.size();
        case OP_DELETE: return 8;
        case OP_LIST_MOVE: return 8;
        case OP_LIST_SET: return 7;
        case OP_STYLE_START: return 10;
        case OP_STYLE_END: return 1;
        case OP_MAP_SET: case OP_MAP_DEL: return 3;
        case OP_TREE_CREATE: case OP_TREE_MOVE: case OP_TREE_DELETE: return 8;
        default: return ctype == CT_COUNTER ? 4 : 6;
    }
}

// DeleteSpan::is_mergable / DeleteSpanWithId::is_mergable (list_op.rs:189-249, 381-396)
inline bool del_mergable(const Op& a, const Op& b) {
    bool ab = a.del_bidirectional(), bb = b.del_bidirectional();
    int64_t apos = a.prop, bpos = b.prop;
    if (ab && bb)
        return (apos == bpos && a.del_start.inc(1) == b.del_start) ||
               (apos == bpos + 1 && a.del_start == b.del_start.inc(1));
    if (ab && !bb) {
        if (apos == b.del_prev_pos())
            return b.del_len > 0 ? a.del_start.inc(1) == b.del_start : a.del_start == b.del_id_end();
        return false;
    }
    if (!ab && bb) {
        if (a.del_next_pos() == bpos)
            return a.del_len > 0 ? a.del_id_end() == b.del_start : a.del_start == b.del_start.inc(1);
        return false;
    }
    if (a.del_next_pos() == bpos && a.del_direction() == b.del_direction())
        return a.del_len > 0 ? a.del_id_end() == b.del_start : a.del_start == b.del_id_end();
    return false;
}
inline void del_merge(Op& a, const Op& b) {  // list_op.rs:244-249, 398-434
    a.del_start.counter = std::min(a.del_start.counter, b.del_start.counter);
    bool ab = a.del_bidirectional(), bb = b.del_bidirectional();
    if (ab && bb) {
        if (a.prop == b.prop) a.del_len = 2;
        else a.del_len = -2;
    } else if (ab && !bb) {
        a.del_len = b.del_len + b.del_direction();
    } else if (!ab && bb) {
        a.del_len += a.del_direction();
    } else {
        a.del_len += b.del_len;
    }
}
inline bool op_mergable(const Op& a, const Op& b) {  // op.rs:143-148 + list_op.rs:516-552
    if (a.ctr_end() != b.counter || a.cidx != b.cidx || a.kind != b.kind) return false;
    switch (a.kind) {
        case OP_LIST_INSERT:
            return a.prop + a.atom_len() == b.prop && a.arena_end == b.arena_start;
        case OP_TEXT_INSERT:
            return a.prop + (int64_t)a.unicode_len == b.prop && a.arena_gen == b.arena_gen &&
                   a.arena_end == b.arena_start && a.arena_ustart + a.unicode_len == b.arena_ustart;
        case OP_DELETE: return del_mergable(a, b);
        default: return false;
    }
}
inline void op_merge(Op& a, const Op& b) {
    switch (a.kind) {
        case OP_LIST_INSERT:
            a.values.insert(a.values.end(), b.values.begin(), b.values.end());
            a.arena_end = b.arena_end;
            break;
        case OP_TEXT_INSERT:
            a.text += b.text;
            a.unicode_len += b.unicode_len;
            a.arena_end = b.arena_end;
            break;
        case OP_DELETE: del_merge(a, b); break;
        default: break;
    }
}
// RleVec::push: returns true when merged into the last element (crates/rle/src/rle_vec.rs)
inline bool rle_push(std::vector<Op>& ops, const Op& op) {
    if (!ops.empty() && op_mergable(ops.back(), op)) {
        op_merge(ops.back(), op);
        return true;
    }
    ops.push_back(op);
    return false;
}
inline bool rle_push(std::vector<Op>& ops, Op&& op) {
    if (!ops.empty() && op_mergable(ops.back(), op)) {
        op_merge(ops.back(), op);
        return true;
    }
    ops.push_back(std::move(op));
    return false;
}
inline size_t utf8_byte_index(const std::string& s, size_t chars) {
    size_t i = 0, c = 0;
    while (i < s.size() && c < chars) {
        i++;
        while (i < s.size() && ((unsigned char)s[i] & 0xC0) == 0x80) i++;
        c++;
    }
    return i;
}
// Op::slice (op.rs:161-172, list_op.rs:603-658, 251-278, 436-444)
inline Op op_slice(const Op& op, int from, int to) {
    Op r = op;
    r.counter = op.counter + from;
    switch (op.kind) {
        case OP_LIST_INSERT:
            r.values.assign(op.values.begin() + from, op.values.begin() + to);
            r.prop = op.prop + from;
            r.arena_start = op.arena_start + from;
            r.arena_end = op.arena_start + to;
            break;
        case OP_TEXT_INSERT: {
            size_t fb = utf8_byte_index(op.text, from), tb = utf8_byte_index(op.text, to);
            r.text = op.text.substr(fb, tb - fb);
            r.unicode_len = (uint32_t)(to - from);
            r.prop = op.prop + from;
            r.arena_start = op.arena_start + fb;
            r.arena_end = op.arena_start + tb;
            r.arena_ustart = op.arena_ustart + from;
            break;
        }
        case OP_DELETE: {
            int n = op.atom_len();
            if (op.del_len > 0) {
                r.del_start = op.del_start.inc(from);
                r.del_len = to - from;
                r.prop = op.prop;
            } else {
                r.del_start = op.del_start.inc(n - to);
                r.prop = op.prop - from;
                r.del_len = from - to;
            }
            break;
        }
        default: break;
    }
    return r;
}

// ------------------------------------------------------------------ tracker (eg-walker for one List/Text)
struct TSpan {
    int peer;  // index into the replay's peer table; -1 = UNKNOWN_PEER_ID sentinel (tracker.rs:38-63)
    Counter ctr;
    int len;
    bool future = false;
    int del = 0;
    bool has_ol = false, has_or = false;
    int ol_peer = 0, or_peer = 0;
    Counter ol_ctr = 0, or_ctr = 0;
    struct TBlock* blk = nullptr;
    bool active() const { return !future && del == 0; }
};
struct TSuper;
struct TBlock {
    std::vector<TSpan*> spans;
    int64_t vis = 0;
    std::list<TBlock>::iterator self;
    TSuper* sup = nullptr;
};
// second level of the rope (the reference's is a B-tree, crdt_rope.rs): a run of consecutive blocks with its totals, so
// that position queries and order keys skip 128 blocks at a time instead of walking the whole list (a single
// document with millions of spans -- BASELINE config C4 -- took hours otherwise)
struct TSuper {
    int64_t vis = 0;       // visible atoms of its blocks
    int64_t nspans = 0;    // spans of its blocks
    int nblocks = 0;
    std::list<TBlock>::iterator first;
    std::list<TSuper>::iterator self;
};
struct DelRec {  // Cursor::Delete(id_span) (tracker/id_to_cursor.rs); op atoms [ctr, ctr+len)
    int len;
    int tpeer;
    Counter tctr;  // leftmost target counter
    bool reversed;  // atom k deletes tctr + (len-1-k) when reversed
};

struct Tracker {
    static const int BLK = 96;
    static const int SUP = 256;     // blocks per superblock before it is cut in two
    std::list<TBlock> blocks;
    std::list<TSuper> supers;
    std::vector<std::map<Counter, TSpan*>> idmap;  // per peer: span start counter -> span
    std::vector<std::map<Counter, DelRec>> delmap;  // per peer: delete-op start counter -> record
    std::vector<Counter> cur_vv;                    // current_vv (tracker.rs:26)
    TSpan* unknown = nullptr;
    std::vector<TSpan*> pool;
    bool inconsistent_delete = false;

    explicit Tracker(int npeers) : idmap(npeers), delmap(npeers), cur_vv(npeers, 0) {
        blocks.emplace_back();
        blocks.back().self = std::prev(blocks.end());
        supers.emplace_back();
        supers.back().self = std::prev(supers.end());
        supers.back().first = blocks.begin();
        supers.back().nblocks = 1;
        blocks.back().sup = &supers.back();
        unknown = new_span(-1, 0, (int)(UINT32_MAX / 4));
        unknown->blk = &blocks.back();
        blocks.back().spans.push_back(unknown);
        blocks.back().vis = unknown->len;
        supers.back().vis = unknown->len;
        supers.back().nspans = 1;
    }
    void add_vis(TBlock* b, int64_t d) { b->vis += d; b->sup->vis += d; }
    // a superblock that grew past SUP blocks hands its second half to a new one
    void split_super(TSuper* su) {
        if (su->nblocks <= SUP) return;
        auto ns = supers.emplace(std::next(su->self));
        ns->self = ns;
        int half = su->nblocks / 2;
        auto b = su->first;
        int64_t vis = 0, nsp = 0;
        for (int k = 0; k < half; k++, ++b) { vis += b->vis; nsp += (int64_t)b->spans.size(); }
        ns->first = b;
        ns->nblocks = su->nblocks - half;
        ns->vis = su->vis - vis;
        ns->nspans = su->nspans - nsp;
        su->nblocks = half;
        su->vis = vis;
        su->nspans = nsp;
        for (int k = 0; k < ns->nblocks; k++, ++b) b->sup = &*ns;
    }
    ~Tracker() {
        for (auto s : pool) delete s;
    }
    Tracker(const Tracker&) = delete;
    TSpan* new_span(int peer, Counter c, int len) {
        TSpan* s = new TSpan();
        s->peer = peer;
        s->ctr = c;
        s->len = len;
        pool.push_back(s);
        return s;
    }
    int64_t vlen(const TSpan* s) const { return s->active() ? s->len : 0; }

    typedef std::list<TBlock>::iterator BIt;
    struct Cur {
        BIt b;
        int i;    // index in block
        int off;  // offset in span
    };
    BIt block_of(TSpan* s) { return s->blk->self; }
    int index_in_block(TSpan* s) {
        auto& v = s->blk->spans;
        for (size_t i = 0; i < v.size(); i++)
            if (v[i] == s) return (int)i;
        assert(false);
        return -1;
    }
    // global order key for cmp_pos (crdt_rope.rs:433-446)
    int64_t order_key(TSpan* s) {
        int64_t k = 0;
        TSuper* su = s->blk->sup;
        for (auto it = supers.begin(); &*it != su; ++it) k += it->nspans;
        for (auto it = su->first; &*it != s->blk; ++it) k += (int64_t)it->spans.size();
        return k + index_in_block(s);
    }
    void insert_span_at(BIt b, int i, TSpan* s) {
        b->spans.insert(b->spans.begin() + i, s);
        s->blk = &*b;
        add_vis(&*b, vlen(s));
        b->sup->nspans++;
        if ((int)b->spans.size() > BLK) {
            auto nb = blocks.emplace(std::next(b));
            nb->self = nb;
            nb->sup = b->sup;
            b->sup->nblocks++;
            int half = (int)b->spans.size() / 2;
            nb->spans.assign(b->spans.begin() + half, b->spans.end());
            b->spans.resize(half);
            nb->vis = 0;
            for (auto x : nb->spans) {
                x->blk = &*nb;
                nb->vis += vlen(x);
            }
            b->vis -= nb->vis;          // (both blocks stay in the same superblock: its totals are unchanged)
            split_super(b->sup);
        }
    }
    // split span s at offset k (0<k<len); returns the right part (FugueSpan::_slice, fugue_span.rs:257-279)
    TSpan* split(TSpan* s, int k) {
        assert(k > 0 && k < s->len);
        TSpan* r = new_span(s->peer, s->ctr + k, s->len - k);
        r->future = s->future;
        r->del = s->del;
        r->has_ol = true;
        r->ol_peer = s->peer;
        r->ol_ctr = s->ctr + k - 1;
        r->has_or = s->has_or;
        r->or_peer = s->or_peer;
        r->or_ctr = s->or_ctr;
        s->len = k;
        BIt b = block_of(s);
        int i = index_in_block(s);
        add_vis(&*b, -vlen(r));  // insert_span_at adds it back
        insert_span_at(b, i + 1, r);
        if (s->peer >= 0) idmap[s->peer][r->ctr] = r;
        return r;
    }
    TSpan* find_span(int peer, Counter c) {
        if (peer < 0) return unknown;
        auto& m = idmap[peer];
        auto it = m.upper_bound(c);
        if (it == m.begin()) return nullptr;
        --it;
        TSpan* s = it->second;
        return (c < s->ctr + s->len) ? s : nullptr;
    }
    // isolate ids [a,b) of `peer` into whole spans and call f on each
    void for_id_range(int peer, Counter a, Counter b, const std::function<void(TSpan*)>& f) {
        Counter c = a;
        while (c < b) {
            TSpan* s = find_span(peer, c);
            if (!s) {  // ids of other containers / not inserts: skip to next known span
                auto& m = idmap[peer];
                auto it = m.upper_bound(c);
                if (it == m.end() || it->first >= b) return;
                c = it->first;
                continue;
            }
            if (s->ctr < c) s = split(s, c - s->ctr);
            if (s->ctr + s->len > b) split(s, b - s->ctr);
            f(s);
            c = s->ctr + s->len;
        }
    }
    void set_status(TSpan* s, int set_future, int del_diff) {
        int64_t before = vlen(s);
        if (set_future >= 0) s->future = set_future != 0;
        s->del += del_diff;
        add_vis(s->blk, vlen(s) - before);
    }
    // retreat (dir=-1) / forward (dir=+1) the ops of `peer` with counters in [a,b)
    // (tracker.rs:334-441 _checkout, :448-526 forward)
    void toggle_range(int peer, Counter a, Counter b, int dir) {
        for_id_range(peer, a, b, [&](TSpan* s) { set_status(s, dir < 0 ? 1 : 0, 0); });
        auto& dm = delmap[peer];
        auto it = dm.upper_bound(a);
        if (it != dm.begin()) --it;
        for (; it != dm.end() && it->first < b; ++it) {
            Counter s0 = it->first, s1 = s0 + it->second.len;
            Counter lo = std::max(a, s0), hi = std::min(b, s1);
            if (lo >= hi) continue;
            const DelRec& d = it->second;
            Counter t0, t1;  // target counters
            if (!d.reversed) {
                t0 = d.tctr + (lo - s0);
                t1 = d.tctr + (hi - s0);
            } else {
                t0 = d.tctr + (d.len - (hi - s0));
                t1 = d.tctr + (d.len - (lo - s0));
            }
            for_id_range(d.tpeer, t0, t1, [&](TSpan* s) { set_status(s, -1, dir); });
        }
    }
    void checkout(const std::vector<Counter>& vv) {
        for (size_t p = 0; p < vv.size(); p++) {
            if (cur_vv[p] > vv[p]) toggle_range((int)p, vv[p], cur_vv[p], -1);
            else if (cur_vv[p] < vv[p]) toggle_range((int)p, cur_vv[p], vv[p], +1);
            cur_vv[p] = vv[p];
        }
    }
    // ActiveLenQueryPreferLeft (crdt_rope.rs:542-598): cursor right after the pos-th visible atom
    Cur query_left(int64_t pos) {
        if (pos == 0) return Cur{blocks.begin(), 0, 0};
        int64_t left = pos;
        auto su = supers.begin();
        while (su != supers.end() && left > su->vis) { left -= su->vis; ++su; }   // the first block with left <= vis is in here
        if (su == supers.end()) throw std::runtime_error("tracker: insert pos out of range");
        for (auto b = su->first; b != blocks.end(); ++b) {
            if (left <= b->vis) {
                for (int i = 0; i < (int)b->spans.size(); i++) {
                    int64_t v = vlen(b->spans[i]);
                    if (v > 0 && left <= v) return Cur{b, i, (int)left};
                    left -= v;
                }
                assert(false);
            }
            left -= b->vis;
        }
        throw std::runtime_error("tracker: insert pos out of range");
    }
    // ActiveLenQueryPreferRight (crdt_rope.rs:600-652): first visible atom with index == pos
    Cur query_right(int64_t pos) {
        int64_t left = pos;
        auto su = supers.begin();
        while (su != supers.end() && left >= su->vis) { left -= su->vis; ++su; }  // the first block with left < vis is in here
        if (su == supers.end()) throw std::runtime_error("tracker: delete pos out of range");
        for (auto b = su->first; b != blocks.end(); ++b) {
            if (left < b->vis) {
                for (int i = 0; i < (int)b->spans.size(); i++) {
                    int64_t v = vlen(b->spans[i]);
                    if (left < v) return Cur{b, i, (int)left};
                    left -= v;
                }
                assert(false);
            }
            left -= b->vis;
        }
        throw std::runtime_error("tracker: delete pos out of range");
    }
    bool next_span(BIt& b, int& i) {  // advance (b,i) to the next span; false at end
        i++;
        while (b != blocks.end() && i >= (int)b->spans.size()) {
            ++b;
            i = 0;
        }
        return b != blocks.end();
    }
    bool prev_span(BIt& b, int& i) {
        i--;
        while (i < 0) {
            if (b == blocks.begin()) return false;
            --b;
            i = (int)b->spans.size() - 1;
        }
        return true;
    }

    // CrdtRope::insert (crdt_rope.rs:43-227)
    void insert(int peer, Counter ctr, int len, int64_t pos, uint64_t peer_id_real,
                const std::vector<uint64_t>& peer_ids) {
        Cur start = query_left(pos);
        TSpan* ns = new_span(peer, ctr, len);
        // origin_left
        if (start.off == 0) {
            BIt b = start.b;
            int i = start.i;
            if (prev_span(b, i)) {
                TSpan* l = b->spans[i];
                ns->has_ol = true;
                ns->ol_peer = l->peer;
                ns->ol_ctr = l->ctr + l->len - 1;
            }
        } else {
            TSpan* l = start.b->spans[start.i];
            ns->has_ol = true;
            ns->ol_peer = l->peer;
            ns->ol_ctr = l->ctr + start.off - 1;
        }
        // origin_right + in_between
        std::vector<TSpan*> in_between;
        TSpan* parent_right = nullptr;  // leaf of the right parent, if any
        {
            BIt b = start.b;
            int i = start.i;
            bool first = true;
            bool ok = b != blocks.end() && i < (int)b->spans.size();
            if (!ok) ok = next_span(b, i);
            while (ok) {
                TSpan* e = b->spans[i];
                int off = first ? start.off : 0;
                bool is_first = first;
                first = false;
                if (is_first && off >= e->len) {
                    ok = next_span(b, i);
                    continue;
                }
                if (!e->future) {
                    ns->has_or = true;
                    ns->or_peer = e->peer;
                    ns->or_ctr = e->ctr + off;
                    if (is_first && off > 0)
                        parent_right = e;
                    else {
                        bool same = (e->has_ol == ns->has_ol) &&
                                    (!e->has_ol || (e->ol_peer == ns->ol_peer && e->ol_ctr == ns->ol_ctr));
                        if (same) parent_right = e;
                    }
                    break;
                }
                in_between.push_back(e);
                ok = next_span(b, i);
            }
        }
        // insertion point
        TSpan* after = nullptr;  // insert right after this span (whole) when set
        if (!in_between.empty()) {
            bool scanning = false;
            std::vector<TSpan*> visited;
            auto in_visited = [&](int p, Counter c) {
                for (auto v : visited)
                    if (v->peer == p && c >= v->ctr && c < v->ctr + v->len) return true;
                return false;
            };
            auto peer_real = [&](int p) -> uint64_t { return p < 0 ? UINT64_MAX : peer_ids[p]; };
            for (TSpan* o : in_between) {
                bool same_ol = (o->has_ol == ns->has_ol) &&
                               (!o->has_ol || (o->ol_peer == ns->ol_peer && o->ol_ctr == ns->ol_ctr));
                if (!same_ol && (!o->has_ol || !in_visited(o->ol_peer, o->ol_ctr))) break;
                visited.push_back(o);
                if (same_ol) {
                    bool same_or = (o->has_or == ns->has_or) &&
                                   (!o->has_or || (o->or_peer == ns->or_peer && o->or_ctr == ns->or_ctr));
                    if (same_or) {
                        if (peer_real(o->peer) > peer_id_real) break;
                        scanning = false;
                    } else {
                        TSpan* other_pr = nullptr;
                        if (o->has_or) {
                            TSpan* e = find_span(o->or_peer, o->or_ctr);
                            assert(e);
                            // origin_left of the atom o->or (span start or interior)
                            bool e_has_ol;
                            int e_olp = 0;
                            Counter e_olc = 0;
                            if (e->ctr == o->or_ctr) {
                                e_has_ol = e->has_ol;
                                e_olp = e->ol_peer;
                                e_olc = e->ol_ctr;
                            } else {
                                e_has_ol = true;
                                e_olp = e->peer;
                                e_olc = o->or_ctr - 1;
                            }
                            bool eq = (e_has_ol == ns->has_ol) &&
                                      (!e_has_ol || (e_olp == ns->ol_peer && e_olc == ns->ol_ctr));
                            if (eq) other_pr = e;
                        }
                        int cmp;  // cmp_pos(other_parent_right, parent_right)
                        if (other_pr && parent_right) {
                            int64_t a = order_key(other_pr), b2 = order_key(parent_right);
                            cmp = a < b2 ? -1 : (a > b2 ? 1 : 0);
                        } else if (other_pr && !parent_right)
                            cmp = -1;
                        else if (!other_pr && parent_right)
                            cmp = 1;
                        else
                            cmp = 0;
                        if (cmp < 0)
                            scanning = true;
                        else if (cmp == 0 && peer_real(o->peer) > peer_id_real)
                            break;
                        else
                            scanning = false;
                    }
                }
                if (!scanning) after = o;
            }
        }
        if (after) {
            BIt b = block_of(after);
            insert_span_at(b, index_in_block(after) + 1, ns);
        } else if (start.b == blocks.end() || start.i >= (int)start.b->spans.size()) {
            insert_span_at(std::prev(blocks.end()), (int)blocks.back().spans.size(), ns);
        } else {
            TSpan* e = start.b->spans[start.i];
            if (start.off == 0)
                insert_span_at(start.b, start.i, ns);
            else if (start.off >= e->len)
                insert_span_at(block_of(e), index_in_block(e) + 1, ns);
            else {
                split(e, start.off);
                insert_span_at(block_of(e), index_in_block(e) + 1, ns);
            }
        }
        idmap[peer][ctr] = ns;
        cur_vv[peer] = std::max(cur_vv[peer], ctr + len);
    }

    // CrdtRope::delete + Tracker::_delete (crdt_rope.rs:236-315, tracker.rs:173-232)
    void del(int op_peer, Counter op_ctr, int target_peer, Counter target_ctr, int64_t pos, int len,
             bool reversed) {
        if (reversed && len > 1) {
            Counter cur = op_ctr;
            for (int i = len - 1; i >= 0; i--) {
                del_forward(op_peer, cur, target_peer, target_ctr + i, pos + i, 1, true);
                cur += 1;
            }
        } else
            del_forward(op_peer, op_ctr, target_peer, target_ctr, pos, len, reversed);
        cur_vv[op_peer] = std::max(cur_vv[op_peer], op_ctr + len);
    }
    void del_forward(int op_peer, Counter op_ctr, int target_peer, Counter target_ctr, int64_t pos,
                     int len, bool reversed) {
        int remaining = len;
        Counter cur_id = op_ctr;
        Counter expect = target_ctr;
        while (remaining > 0) {
            Cur c = query_right(pos);
            TSpan* s = c.b->spans[c.i];
            if (c.off > 0) s = split(s, c.off);
            if (s->len > remaining) split(s, remaining);
            if (s->peer != target_peer || s->ctr != expect) inconsistent_delete = true;
            set_status(s, -1, +1);
            delmap[op_peer][cur_id] = DelRec{s->len, s->peer, s->ctr, reversed};
            cur_id += s->len;
            expect += s->len;
            remaining -= s->len;
            // the deleted atoms vanish from the visible sequence, so `pos` stays
        }
    }
};

// ------------------------------------------------------------------ materialised state
struct SeqItem {
    ID id;          // atom id
    int op_index;   // index into Replay::ins_ops (content lookup)
    int off;        // atom offset inside that op
};
struct MapEntry {
    bool has = false;
    Value v;
    Lamport lamport = 0;
    PeerID peer = 0;
    bool set = false;
};
struct TreeNodeState {
    ID id;
    bool parent_null = true;  // root
    ID parent;
    bool deleted = false;         // dead: the node or one of its ancestors sits under DELETED_TREE_ROOT
    bool direct_deleted = false;  // the node's own parent is DELETED_TREE_ROOT
    std::string position;
    Lamport lamport;  // of the last effective move
    PeerID peer;
};

// ------------------------------------------------------------------ fractional index (local tree ops only)
// crates/fractional_index/src/lib.rs:52-127, jitter 0.  Strings hold the full bytes incl. the terminator 0x80.
namespace fi {
static const uint8_t TERM = 128;
inline std::string new_before(const std::string& b) {
    for (size_t i = 0; i < b.size(); i++) {
        uint8_t c = (uint8_t)b[i];
        if (c > TERM) return b.substr(0, i);
        if (c > 0) { std::string a = b.substr(0, i + 1); a[i] = (char)(c - 1); return a; }
    }
    throw std::runtime_error("fractional index: new_before");
}
inline std::string new_after(const std::string& b) {
    for (size_t i = 0; i < b.size(); i++) {
        uint8_t c = (uint8_t)b[i];
        if (c < TERM) return b.substr(0, i);
        if (c < 255) { std::string a = b.substr(0, i + 1); a[i] = (char)(c + 1); return a; }
    }
    throw std::runtime_error("fractional index: new_after");
}
inline bool new_between(const std::string& l, const std::string& r, std::string* out) {
    size_t shorter = std::min(l.size(), r.size()) - 1;
    for (size_t i = 0; i < shorter; i++) {
        int a = (uint8_t)l[i], b = (uint8_t)r[i];
        if (a < b - 1) { std::string x = l.substr(0, i + 1); x[i] = (char)(a + (b - a) / 2); *out = x; return true; }
        if (a == b - 1) { *out = l.substr(0, i + 1) + new_after(l.substr(i + 1)); return true; }
        if (a > b) return false;
    }
    if (l.size() < r.size()) {
        std::string prefix = r.substr(0, shorter + 1);
        if ((uint8_t)prefix.back() < TERM) return false;
        *out = prefix + new_before(r.substr(shorter + 1));
        return true;
    }
    if (l.size() == r.size()) return false;
    std::string prefix = l.substr(0, shorter + 1);
    if ((uint8_t)prefix.back() >= TERM) return false;
    *out = prefix + new_after(l.substr(shorter + 1));
    return true;
}
// FractionalIndex::new (lib.rs:128-136): nullptr = no bound; the result is terminated
inline bool make(const std::string* lower, const std::string* upper, std::string* out) {
    std::string body;
    if (lower && upper) { if (!new_between(*lower, *upper, &body)) return false; }
    else if (lower) body = new_after(*lower);
    else if (upper) body = new_before(*upper);
    else { *out = std::string(1, (char)TERM); return true; }
    body.push_back((char)TERM);
    *out = body;
    return true;
}
inline void gen_evenly(const std::string* lower, const std::string* upper, size_t n, std::vector<std::string>& out) {
    if (n == 0) return;  // lib.rs:155-195
    size_t mid = n / 2;
    std::string m;
    if (!make(lower, upper, &m)) throw std::runtime_error("fractional index: generate_n_evenly");
    if (n == 1) { out.push_back(m); return; }
    gen_evenly(lower, &m, mid, out);
    out.push_back(m);
    if (n - mid - 1 == 0) return;
    gen_evenly(&m, upper, n - mid - 1, out);
}
}  // namespace fi
struct ContainerState {
    uint8_t type = 0;
    // list/text
    std::vector<Value> list_values;
    std::vector<ID> ids;
    std::string text;
    // map: sorted by key
    std::map<std::string, MapEntry> map;
    // tree
    std::vector<TreeNodeState> tree;  // alive + deleted nodes
    bool unsupported = false;         // saw ops the oracle does not merge (styles, moves, counter)
};

// ------------------------------------------------------------------ change store model
struct StoreBlock {
    PeerID peer;
    Counter c0, c1;
    Lamport l0, l1;
    size_t est;
    std::vector<Change> changes;
};

struct Doc : ArenaCtx {
    PeerID peer;
    std::vector<ContainerID> containers;
    std::map<ContainerID, int> cid_index;
    // arena adjacency model (arena.rs:237-263; arena/str_arena.rs; append-only-bytes 0.1.12)
    uint64_t arena_values = 0, arena_str_bytes = 0, arena_str_unicode = 0, str_cap = 0;
    uint32_t str_gen = 0;
    std::map<ID, StoreBlock> store;  // ChangeStore::mem_parsed_kv keyed by block start id
    std::map<PeerID, Counter> vv;
    std::vector<ID> frontiers;
    std::vector<Change> pending;
    // local transaction
    Change txn;
    bool txn_open = false;
    // state cache
    std::vector<ContainerState> state;
    bool state_valid = false;
    bool inconsistent_delete = false;
    uint64_t replay_ops = 0;

    explicit Doc(PeerID p) : peer(p) {}

    // ---------------- ArenaCtx
    int register_container(const ContainerID& c) override {
        auto it = cid_index.find(c);
        if (it != cid_index.end()) return it->second;
        int i = (int)containers.size();
        containers.push_back(c);
        cid_index[c] = i;
        return i;
    }
    const ContainerID& container_id(int cidx) const override { return containers[(size_t)cidx]; }
    void alloc_values(Op& op) override {
        op.arena_start = arena_values;
        arena_values += op.values.size();
        op.arena_end = arena_values;
    }
    void alloc_str(Op& op) override {
        // StrArena::alloc pushes the string in chunks of >128 bytes (str_arena.rs:46-66); the
        // AppendOnlyBytes buffer doubles from 32 (restated from append-only-bytes 0.1.12, not in tree;
        // UNPINNED: only affects whether two adjacent text inserts re-merge on export).
        op.arena_start = arena_str_bytes;
        op.arena_ustart = arena_str_unicode;
        uint64_t target = arena_str_bytes + op.text.size();
        if (target > str_cap) {
            uint64_t nc = std::max<uint64_t>(str_cap * 2, 32);
            while (nc < target) nc *= 2;
            str_cap = nc;
            str_gen++;
        }
        arena_str_bytes = target;
        arena_str_unicode += op.unicode_len;
        op.arena_end = arena_str_bytes;
        op.arena_gen = str_gen;
    }

    // ---------------- estimates
    size_t change_estimate(const Change& c) const {  // change.rs:128-139
        size_t ops = 0;
        for (auto& op : c.ops) ops += op_estimate(op, containers[(size_t)op.cidx].type);
        size_t deps = (std::max<size_t>(c.deps.size(), 1) - 1) * 4;
        return 2 + 1 + 1 + ops + deps;
    }
    static bool can_merge_right(const Change& a, const Change& b, int64_t merge_interval) {  // change.rs:268-283
        return b.id.peer == a.id.peer && b.id.counter == a.id.counter + a.atom_len() &&
               b.deps.size() == 1 && b.deps[0].peer == a.id.peer &&
               b.timestamp - a.timestamp <= merge_interval && a.has_msg == b.has_msg && a.msg == b.msg;
    }
    static Change change_slice(const Change& c, int from, int to) {  // change.rs:203-258
        Change r;
        Counter fc = c.id.counter + from, tc = c.id.counter + to;
        for (auto& op : c.ops) {
            if (op.counter >= tc) break;
            if (op.ctr_end() <= fc) continue;
            int so = std::min(std::max(fc - op.counter, 0), op.atom_len());
            int eo = std::min(std::max(tc - op.counter, 0), op.atom_len());
            r.ops.push_back(op_slice(op, so, eo));
        }
        if (from > 0) r.deps = {c.id.inc(from - 1)};
        else r.deps = c.deps;
        r.id = c.id.inc(from);
        r.lamport = c.lamport + (Lamport)from;
        r.timestamp = c.timestamp;
        r.has_msg = c.has_msg;
        r.msg = c.msg;
        return r;
    }

    // ---------------- ChangeStore::insert_change (change_store.rs:711-764) on an arbitrary store
    void store_insert(std::map<ID, StoreBlock>& st, Change change, bool split_when_exceeds,
                      int64_t merge_interval) {
        size_t est = change_estimate(change);
        if (est > MAX_BLOCK_SIZE && split_when_exceeds) {
            split_change_then_insert(st, change);
            return;
        }
        ID id = change.id;
        auto it = st.lower_bound(id);
        if (it != st.begin()) {
            --it;
            StoreBlock& b = it->second;
            if (b.peer == id.peer) {
                if (b.c1 != id.counter) throw std::runtime_error("counter should be continuous");
                if (push_change(b, change, est, merge_interval)) return;
            }
        }
        StoreBlock nb;
        nb.peer = id.peer;
        nb.c0 = id.counter;
        nb.c1 = change.ctr_end();
        nb.l0 = change.lamport;
        nb.l1 = change.lamport_end();
        nb.est = est;
        nb.changes.push_back(std::move(change));
        st[id] = std::move(nb);
    }
    // ChangesBlock::push_change (change_store.rs:1244-1291)
    bool push_change(StoreBlock& b, Change& change, size_t new_size, int64_t merge_interval) {
        if (b.c1 != change.id.counter) return false;
        int atom_len = change.atom_len();
        bool is_full = new_size + b.est > MAX_BLOCK_SIZE;
        Change& last = b.changes.back();
        if (can_merge_right(last, change, merge_interval) &&
            (!is_full || (change.ops.size() == 1 && op_mergable(last.ops.back(), change.ops[0])))) {
            for (auto& op : change.ops) {
                size_t size = op_estimate(op, containers[(size_t)op.cidx].type);
                if (!rle_push(last.ops, op)) b.est += size;
            }
        } else {
            if (is_full) return false;
            b.est += new_size;
            b.changes.push_back(change);
        }
        b.c1 = change.id.counter + atom_len;
        b.l1 = change.lamport + (Lamport)atom_len;
        return true;
    }
    // change_store.rs:913-1000
    void split_change_then_insert(std::map<ID, StoreBlock>& st, const Change& change) {
        Change nc;
        nc.deps = change.deps;
        nc.id = change.id;
        nc.lamport = change.lamport;
        nc.timestamp = change.timestamp;
        nc.has_msg = change.has_msg;
        nc.msg = change.msg;
        size_t est = change_estimate(nc);
        auto flush = [&]() {
            if (nc.atom_len() == 0) return;
            Counter ctr_end = nc.ctr_end();
            Lamport next_l = nc.lamport_end();
            Change ans;
            ans.deps = {ID{nc.id.peer, ctr_end - 1}};
            ans.id = ID{nc.id.peer, ctr_end};
            ans.lamport = next_l;
            ans.timestamp = nc.timestamp;
            ans.has_msg = nc.has_msg;
            ans.msg = nc.msg;
            store_insert(st, nc, false, 0);
            nc = ans;
            est = change_estimate(nc);
        };
        for (const Op& op0 : change.ops) {
            Op op = op0;
            uint8_t ct = containers[(size_t)op.cidx].type;
            if (op_estimate(op, ct) >= MAX_BLOCK_SIZE - est) flush();
            bool consumed = false;
            while (true) {
                size_t room = MAX_BLOCK_SIZE - est;
                if (op_estimate(op, ct) <= room) break;
                if (op.kind != OP_LIST_INSERT && op.kind != OP_TEXT_INSERT) break;
                size_t end = op.kind == OP_TEXT_INSERT ? std::min<size_t>(room, op.atom_len())
                                                       : std::min<size_t>(room / 4, op.atom_len());
                if (end == 0) break;
                rle_push(nc.ops, op_slice(op, 0, (int)end));
                flush();
                if ((int)end < op.atom_len())
                    op = op_slice(op, (int)end, op.atom_len());
                else {
                    consumed = true;
                    break;
                }
            }
            if (consumed) continue;
            est += op_estimate(op, ct);
            if (est > MAX_BLOCK_SIZE && !nc.ops.empty()) {
                flush();
                rle_push(nc.ops, op);
            } else
                rle_push(nc.ops, op);
        }
        if (!nc.ops.empty()) store_insert(st, nc, false, 0);
    }

    // ---------------- dag helpers
    const Change* find_change(ID id) const {
        auto it = store.upper_bound(id);
        if (it == store.begin()) return nullptr;
        --it;
        const StoreBlock& b = it->second;
        if (b.peer != id.peer || id.counter >= b.c1) return nullptr;
        for (auto& c : b.changes)
            if (id.counter >= c.id.counter && id.counter < c.ctr_end()) return &c;
        return nullptr;
    }
    bool get_lamport(ID id, Lamport* out) const {  // loro_dag.rs:935-944
        const Change* c = find_change(id);
        if (!c) return false;
        *out = c->lamport + (Lamport)(id.counter - c->id.counter);
        return true;
    }
    bool lamport_from_deps(const std::vector<ID>& deps, Lamport* out) const {  // loro_dag.rs:946-954
        Lamport l = 0;
        for (auto& d : deps) {
            Lamport x;
            if (!get_lamport(d, &x)) return false;
            l = std::max(l, x + 1);
        }
        *out = l;
        return true;
    }
    Counter vv_get(PeerID p) const {
        auto it = vv.find(p);
        return it == vv.end() ? 0 : it->second;
    }
    // OpLog::insert_new_change (oplog.rs:130-145) minus dag-node bookkeeping
    void insert_new_change(Change change, bool from_local) {
        // Frontiers::update_frontiers_on_new_change (version/frontiers.rs:233-246)
        for (auto& d : change.deps)
            frontiers.erase(std::remove(frontiers.begin(), frontiers.end(), d), frontiers.end());
        frontiers.push_back(change.id_last());
        vv[change.id.peer] = change.ctr_end();
        store_insert(store, std::move(change), true, from_local ? 1000LL * 1000 : 0);
        if (!from_local) state_valid = false;  // local ops keep `state` in sync themselves
    }

    // ---------------- import (encoding.rs:232-270, fast_snapshot.rs:270-288)
    struct ImportStatus {
        std::map<PeerID, std::pair<Counter, Counter>> success, pending;
    };
    static void range_extend(std::map<PeerID, std::pair<Counter, Counter>>& r, PeerID p, Counter a, Counter b) {
        auto it = r.find(p);
        if (it == r.end()) r[p] = {a, b};
        else {
            it->second.first = std::min(it->second.first, a);
            it->second.second = std::max(it->second.second, b);
        }
    }
    // returns BlobErr-like code: 0 ok, 1..4 header errors, 10 decode error, 11 unsupported mode(snapshot)
    int import(const uint8_t* bytes, size_t n, ImportStatus* status, std::string* err = nullptr) {
        commit();
        uint16_t mode;
        const uint8_t* body;
        size_t body_len;
        BlobErr e = parse_blob(bytes, n, &mode, &body, &body_len);
        if (e != BLOB_OK) return (int)e;
        if (mode != MODE_FAST_UPDATES) return 11;
        std::vector<Change> changes;
        try {
            auto blocks = split_updates_body(body, body_len);
            for (auto& blk : blocks) {
                std::vector<Change> raw = decode_block(blk.first, blk.second, *this);
                if (raw.empty()) continue;
                Counter start = vv_get(raw[0].id.peer);  // change_store.rs:244-267
                for (auto& c0 : raw) {
                    Change c = std::move(c0);  // re-push ops through the RleVec merge (block_encode.rs:651)
                    std::vector<Op> src;
                    src.swap(c.ops);
                    c.ops.reserve(src.size());
                    for (auto& op : src) rle_push(c.ops, std::move(op));
                    if (c.id.counter >= start)
                        changes.push_back(std::move(c));
                    else if (c.ctr_end() > start)
                        changes.push_back(change_slice(c, start - c.id.counter, c.atom_len()));
                }
            }
        } catch (DecodeError& ex) {
            if (err) *err = ex.what();
            return 10;
        }
        std::stable_sort(changes.begin(), changes.end(),
                         [](const Change& a, const Change& b) { return a.lamport < b.lamport; });
        ImportStatus st;
        std::vector<Change> pend;
        for (auto& change : changes) {  // import_changes_to_oplog
            if (change.ctr_end() <= vv_get(change.id.peer)) continue;
            Lamport l;
            if (!lamport_from_deps(change.deps, &l)) {
                pend.push_back(std::move(change));
                continue;
            }
            change.lamport = l;
            apply_remote(std::move(change), &st);
        }
        for (auto& c : pend) range_extend(st.pending, c.id.peer, c.id.counter, c.ctr_end());
        for (auto& c : pend) pending.push_back(std::move(c));
        try_apply_pending(&st);
        if (status) *status = st;
        return 0;
    }
    void apply_remote(Change change, ImportStatus* st) {
        Counter end = vv_get(change.id.peer);  // trim_the_known_part_of_change (oplog.rs:181-196)
        if (change.id.counter < end) {
            if (change.ctr_end() <= end) return;
            change = change_slice(change, end - change.id.counter, change.atom_len());
        } else if (change.id.counter > end) {
            // gap on own peer: cannot happen once deps are satisfied (self dep covers it)
        }
        if (st) range_extend(st->success, change.id.peer, change.id.counter, change.ctr_end());
        insert_new_change(std::move(change), false);
    }
    void try_apply_pending(ImportStatus* st) {  // pending_changes.rs:63-140 (fixpoint form)
        bool progress = true;
        while (progress && !pending.empty()) {
            progress = false;
            std::stable_sort(pending.begin(), pending.end(),
                             [](const Change& a, const Change& b) { return a.lamport < b.lamport; });
            std::vector<Change> rest;
            for (auto& c : pending) {
                if (c.ctr_end() <= vv_get(c.id.peer)) {
                    progress = true;
                    continue;
                }
                Lamport l;
                bool self_gap = c.id.counter > vv_get(c.id.peer);
                if (!self_gap && lamport_from_deps(c.deps, &l)) {
                    c.lamport = l;
                    apply_remote(std::move(c), st);
                    progress = true;
                } else
                    rest.push_back(std::move(c));
            }
            pending.swap(rest);
        }
    }

    // ---------------- export (encoding.rs:350-416, change_store.rs:494-576)
    std::vector<uint8_t> export_updates(const std::map<PeerID, Counter>& from) {
        commit();
        std::map<ID, StoreBlock> ns;
        for (auto& kv : vv) {
            PeerID p = kv.first;
            auto f = from.find(p);
            Counter start = f == from.end() ? 0 : f->second;
            Counter end = kv.second;
            if (start >= end) continue;
            for (auto it = store.lower_bound(ID{p, 0}); it != store.end() && it->first.peer == p; ++it) {
                for (auto& c : it->second.changes) {
                    if (c.ctr_end() <= start) continue;
                    int s = std::min(std::max(start - c.id.counter, 0), c.atom_len());
                    int e = std::min(std::max(end - c.id.counter, 0), c.atom_len());
                    if (s == e) continue;
                    Change ch = (s == 0 && e == c.atom_len()) ? c : change_slice(c, s, e);
                    store_insert(ns, ch, false, 0);
                }
            }
        }
        Writer body;
        for (auto& kv : ns) {
            std::vector<uint8_t> b = encode_block(kv.second.changes, *this);
            body.uleb(b.size());
            body.bytes(b);
        }
        return wrap_blob(MODE_FAST_UPDATES, body.buf);
    }

    // ---------------- local ops (handler.rs / txn.rs; workload generation only)
    Counter next_counter() { return txn_open ? txn.ctr_end() : vv_get(peer); }
    void txn_begin() {
        if (txn_open) return;
        txn = Change();
        txn.id = ID{peer, vv_get(peer)};
        txn.deps = frontiers;
        Lamport l = 0;
        lamport_from_deps(frontiers, &l);
        txn.lamport = l;
        txn.timestamp = 0;  // record_timestamp defaults to false (configure.rs:23-25)
        txn_open = true;
    }
    void commit() {
        if (!txn_open) return;
        txn_open = false;
        if (txn.ops.empty()) return;
        Change c = std::move(txn);
        insert_new_change(std::move(c), true);
    }
    int get_container(const std::string& name, uint8_t type) {
        ContainerID c;
        c.root = true;
        c.name = name;
        c.type = type;
        return register_container(c);
    }
    void push_local(Op op) {
        txn_begin();
        op.counter = txn.ctr_end();
        if (txn.ops.empty()) op.counter = txn.id.counter;
        rle_push(txn.ops, op);
    }
    void ensure_state() {
        if (!state_valid) replay();
    }
    // local edits keep `state` in sync so that generation stays O(n) per op
    bool text_insert(int cidx, size_t pos, const std::string& s) {
        ensure_state();
        ContainerState& st = cstate(cidx);
        size_t n = st.ids.size();
        if (pos > n || s.empty()) return false;
        Op op;
        op.cidx = cidx;
        op.kind = OP_TEXT_INSERT;
        op.prop = (int32_t)pos;
        op.text = s;
        op.unicode_len = (uint32_t)utf8_chars(s);
        alloc_str(op);
        Counter c0 = next_counter();
        push_local(op);
        size_t bpos = utf8_byte_index(st.text, pos);
        st.text.insert(bpos, s);
        std::vector<ID> ids;
        for (uint32_t i = 0; i < op.unicode_len; i++) ids.push_back(ID{peer, c0 + (Counter)i});
        st.ids.insert(st.ids.begin() + pos, ids.begin(), ids.end());
        return true;
    }
    bool list_insert(int cidx, size_t pos, const std::vector<Value>& vals) {
        ensure_state();
        ContainerState& st = cstate(cidx);
        if (pos > st.ids.size() || vals.empty()) return false;
        Op op;
        op.cidx = cidx;
        op.kind = OP_LIST_INSERT;
        op.prop = (int32_t)pos;
        op.values = vals;
        Counter c0 = next_counter();
        for (size_t i = 0; i < op.values.size(); i++)
            if (op.values[i].k == Value::Container) {
                op.values[i].cid.root = false;
                op.values[i].cid.peer = peer;
                op.values[i].cid.counter = c0 + (Counter)i;
                register_container(op.values[i].cid);
            }
        alloc_values(op);
        push_local(op);
        std::vector<ID> ids;
        for (size_t i = 0; i < vals.size(); i++) ids.push_back(ID{peer, c0 + (Counter)i});
        st.ids.insert(st.ids.begin() + pos, ids.begin(), ids.end());
        st.list_values.insert(st.list_values.begin() + pos, op.values.begin(), op.values.end());
        return true;
    }
    // delete [pos,pos+len): text emits one op per contiguous id run, rightmost first
    // (handler.rs:1897-1957); list emits one op per element (handler.rs:2744-2779)
    bool seq_delete(int cidx, size_t pos, size_t len) {
        ensure_state();
        ContainerState& st = cstate(cidx);
        if (len == 0 || pos + len > st.ids.size()) return false;
        bool is_text = containers[(size_t)cidx].type == CT_TEXT;
        if (is_text) {
            size_t end = pos + len;
            while (end > pos) {
                size_t s = end - 1;
                while (s > pos && st.ids[s - 1].peer == st.ids[s].peer &&
                       st.ids[s - 1].counter + 1 == st.ids[s].counter)
                    s--;
                Op op;
                op.cidx = cidx;
                op.kind = OP_DELETE;
                op.prop = (int32_t)s;
                op.del_start = st.ids[s];
                op.del_len = (int64_t)(end - s);
                push_local(op);
                end = s;
            }
            size_t b0 = utf8_byte_index(st.text, pos), b1 = utf8_byte_index(st.text, pos + len);
            st.text.erase(b0, b1 - b0);
        } else {
            for (size_t i = 0; i < len; i++) {
                Op op;
                op.cidx = cidx;
                op.kind = OP_DELETE;
                op.prop = (int32_t)pos;
                op.del_start = st.ids[pos + i];
                op.del_len = 1;
                push_local(op);
            }
            st.list_values.erase(st.list_values.begin() + pos, st.list_values.begin() + pos + len);
        }
        st.ids.erase(st.ids.begin() + pos, st.ids.begin() + pos + len);
        return true;
    }
    bool map_set(int cidx, const std::string& key, const Value* v) {
        ensure_state();
        Op op;
        op.cidx = cidx;
        op.kind = v ? OP_MAP_SET : OP_MAP_DEL;
        op.key = key;
        Counter c0 = next_counter();
        if (v) {
            op.mapval = *v;
            if (op.mapval.k == Value::Container) {
                op.mapval.cid.root = false;
                op.mapval.cid.peer = peer;
                op.mapval.cid.counter = c0;
                register_container(op.mapval.cid);
            }
        }
        txn_begin();
        Lamport l = txn.lamport + (Lamport)(c0 - txn.id.counter);
        push_local(op);
        MapEntry& e = cstate(cidx).map[key];
        e.set = true;
        e.has = v != nullptr;
        if (v) e.v = op.mapval;
        e.lamport = l;
        e.peer = peer;
        return true;
    }
    // ---------------- local tree ops (handler/tree.rs:292-355,518-723; state/tree_state.rs:175-236,690-723)
    TreeNodeState* tree_find(ContainerState& st, ID id) {
        for (auto& n : st.tree)
            if (n.id == id) return &n;
        return nullptr;
    }
    // children of an alive parent in sibling order (position, lamport, peer)
    std::vector<TreeNodeState*> tree_children(ContainerState& st, bool root, ID parent) {
        std::vector<TreeNodeState*> kids;
        for (auto& n : st.tree)
            if (!n.deleted && n.parent_null == root && (root || n.parent == parent)) kids.push_back(&n);
        std::sort(kids.begin(), kids.end(), [](const TreeNodeState* a, const TreeNodeState* b) {
            if (a->position != b->position) return a->position < b->position;
            if (a->lamport != b->lamport) return a->lamport < b->lamport;
            return a->peer < b->peer;
        });
        return kids;
    }
    bool tree_is_ancestor(ContainerState& st, ID anc, bool node_root, ID node) {  // tree_state.rs:727-747
        if (!tree_find(st, anc)) return false;
        while (!node_root) {
            if (node == anc) return true;
            TreeNodeState* n = tree_find(st, node);
            if (!n || n->direct_deleted) return false;
            node_root = n->parent_null;
            node = n->parent;
        }
        return false;
    }
    // NodeChildren::generate_fi_at (tree_state.rs:175-236): positions for `target` at index `pos` among `kids`
    // (the target itself already taken out); more than one entry = siblings with equal positions get new ones
    std::vector<std::pair<ID, std::string>> tree_positions_at(const std::vector<TreeNodeState*>& kids, size_t pos, ID target) {
        std::vector<std::pair<ID, std::string>> out;
        if (kids.empty()) { out.push_back({target, std::string(1, (char)fi::TERM)}); return out; }
        const std::string* left = pos > 0 ? &kids[pos - 1]->position : nullptr;
        const std::string* right = pos < kids.size() ? &kids[pos]->position : nullptr;
        std::vector<ID> reset;
        const std::string* next_right = nullptr;
        if (left && right && *left == *right) {
            reset.push_back(kids[pos]->id);
            for (size_t i = pos + 1; i < kids.size(); i++) {
                if (kids[i]->position == *left) reset.push_back(kids[i]->id);
                else { next_right = &kids[i]->position; break; }
            }
        }
        if (reset.empty()) {
            std::string p;
            if (!fi::make(left, right, &p)) throw std::runtime_error("fractional index: no room");
            out.push_back({target, p});
            return out;
        }
        std::vector<std::string> ps;
        fi::gen_evenly(left, next_right, reset.size() + 1, ps);
        out.push_back({target, ps[0]});
        for (size_t i = 0; i < reset.size(); i++) out.push_back({reset[i], ps[i + 1]});
        return out;
    }
    void tree_push_op(int cidx, OpKind kind, ID target, bool parent_root, ID parent, const std::string& position) {
        Op op;
        op.cidx = cidx;
        op.kind = kind;
        op.target = target;
        op.parent_null = parent_root;
        op.parent = parent;
        op.position = position;
        txn_begin();
        Counter c0 = next_counter();
        Lamport l = txn.lamport + (Lamport)(c0 - txn.id.counter);
        push_local(op);
        ContainerState& st = cstate(cidx);
        TreeNodeState* n = tree_find(st, target);
        if (!n) { st.tree.push_back(TreeNodeState()); n = &st.tree.back(); n->id = target; }
        n->lamport = l;
        n->peer = peer;
        if (kind == OP_TREE_DELETE) {
            n->direct_deleted = true;
        } else {
            n->direct_deleted = false;
            n->parent_null = parent_root;
            n->parent = parent;
            n->position = position;
        }
        // liveness of everything below a deleted (or revived) node
        for (auto& x : st.tree) {
            bool dead = false;
            const TreeNodeState* cur = &x;
            for (int guard = 0; guard < 1000000; guard++) {
                if (cur->direct_deleted) { dead = true; break; }
                if (cur->parent_null) break;
                cur = tree_find(st, cur->parent);
                if (!cur) { dead = true; break; }
            }
            x.deleted = dead;
        }
    }
    // index < 0 = append.  Returns false when the request is invalid (dead / missing parent, index out of range).
    bool tree_create(int cidx, bool parent_root, ID parent, int index, ID* out) {
        ensure_state();
        ContainerState& st = cstate(cidx);
        if (!parent_root) { TreeNodeState* p = tree_find(st, parent); if (!p || p->deleted) return false; }
        auto kids = tree_children(st, parent_root, parent);
        size_t pos = index < 0 ? kids.size() : (size_t)index;
        if (pos > kids.size()) return false;
        ID target{peer, next_counter()};
        auto ps = tree_positions_at(kids, pos, target);
        for (size_t i = 0; i < ps.size(); i++)
            tree_push_op(cidx, i == 0 ? OP_TREE_CREATE : OP_TREE_MOVE, ps[i].first, parent_root, parent, ps[i].second);
        if (out) *out = target;
        return true;
    }
    bool tree_move(int cidx, ID target, bool parent_root, ID parent, int index) {
        ensure_state();
        ContainerState& st = cstate(cidx);
        TreeNodeState* t = tree_find(st, target);
        if (!t || t->deleted) return false;
        if (!parent_root) { TreeNodeState* p = tree_find(st, parent); if (!p || p->deleted) return false; }
        if (tree_is_ancestor(st, target, parent_root, parent)) return false;  // CyclicMoveError
        auto kids = tree_children(st, parent_root, parent);
        bool already = t->parent_null == parent_root && (parent_root || t->parent == parent);
        size_t len = kids.size();
        if (already) {
            size_t cur = 0;
            while (kids[cur]->id != target) cur++;
            if (index >= 0 && cur == (size_t)index) return true;  // nothing to do
            kids.erase(kids.begin() + cur);
            len--;
        }
        size_t pos = index < 0 ? len : (size_t)index;
        if (pos > len) return false;
        auto ps = tree_positions_at(kids, pos, target);
        for (auto& pr : ps) tree_push_op(cidx, OP_TREE_MOVE, pr.first, parent_root, parent, pr.second);
        return true;
    }
    bool tree_delete(int cidx, ID target) {
        ensure_state();
        ContainerState& st = cstate(cidx);
        TreeNodeState* t = tree_find(st, target);
        if (!t || t->deleted) return false;
        tree_push_op(cidx, OP_TREE_DELETE, target, false, ID{DELETED_TREE_ROOT_PEER, DELETED_TREE_ROOT_CTR}, "");
        return true;
    }
    int tree_meta(ID target) {  // TreeID::associated_meta_container (loro-common/src/lib.rs)
        ContainerID c;
        c.root = false;
        c.peer = target.peer;
        c.counter = target.counter;
        c.type = CT_MAP;
        return register_container(c);
    }
    ContainerState& cstate(int cidx) {
        if (state.size() < containers.size()) state.resize(containers.size());
        ContainerState& s = state[(size_t)cidx];
        s.type = containers[(size_t)cidx].type;
        return s;
    }

    // ---------------- replay: oplog -> state
    void replay();
    std::string to_json();
    void json_container(std::string& out, int cidx, int depth);
    void json_value(std::string& out, const Value& v, int depth);
};

// ---- JSON helpers (serde_json compact output; loro-common/src/value.rs:692-711)
inline void json_escape(std::string& out, const std::string& s) {
    static const char* hex = "0123456789abcdef";
    out.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) {
                    out += "\\u00";
                    out.push_back(hex[c >> 4]);
                    out.push_back(hex[c & 15]);
                } else
                    out.push_back((char)c);
        }
    }
    out.push_back('"');
}
inline void json_f64(std::string& out, double d) {  // serde_json: ryu shortest, null for non-finite
    if (!std::isfinite(d)) {
        out += "null";
        return;
    }
    char buf[64];
    auto r = std::to_chars(buf, buf + 64, d, std::chars_format::scientific);
    std::string s(buf, r.ptr);  // d.ddddde[+-]xx
    bool neg = s[0] == '-';
    if (neg) s = s.substr(1);
    size_t epos = s.find('e');
    std::string mant = s.substr(0, epos);
    int exp = std::atoi(s.c_str() + epos + 1);
    std::string digits;
    for (char c : mant)
        if (c != '.') digits.push_back(c);
    int nd = (int)digits.size();
    int kk = exp + 1;  // decimal point position
    std::string o;
    if (neg) o.push_back('-');
    if (d == 0) {
        o += "0.0";
    } else if (0 < kk && kk <= 16 && nd <= kk) {  // integer-valued
        o += digits;
        o.append((size_t)(kk - nd), '0');
        o += ".0";
    } else if (0 < kk && kk <= 16) {
        o += digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
    } else if (-5 < kk && kk <= 0) {
        o += "0.";
        o.append((size_t)(-kk), '0');
        o += digits;
    } else {
        o.push_back(digits[0]);
        if (nd > 1) o += "." + digits.substr(1);
        o += "e" + std::to_string(kk - 1);
    }
    out += o;
}
inline std::string id_string(ID id) { return std::to_string(id.counter) + "@" + std::to_string(id.peer); }

inline void Doc::json_value(std::string& out, const Value& v, int depth) {
    switch (v.k) {
        case Value::Null: out += "null"; break;
        case Value::True: out += "true"; break;
        case Value::False: out += "false"; break;
        case Value::I64: out += std::to_string(v.i); break;
        case Value::F64: json_f64(out, v.f); break;
        case Value::Str: json_escape(out, v.s); break;
        case Value::Binary: {
            out.push_back('[');
            for (size_t i = 0; i < v.s.size(); i++) {
                if (i) out.push_back(',');
                out += std::to_string((unsigned)(unsigned char)v.s[i]);
            }
            out.push_back(']');
            break;
        }
        case Value::List: {
            out.push_back('[');
            for (size_t i = 0; i < v.list.size(); i++) {
                if (i) out.push_back(',');
                json_value(out, v.list[i], depth);
            }
            out.push_back(']');
            break;
        }
        case Value::Map: {
            std::vector<const std::pair<std::string, Value>*> es;
            for (auto& kv : v.map) es.push_back(&kv);
            std::stable_sort(es.begin(), es.end(), [](auto a, auto b) { return a->first < b->first; });
            out.push_back('{');
            bool first = true;
            for (size_t i = 0; i < es.size(); i++) {
                if (i + 1 < es.size() && es[i + 1]->first == es[i]->first) continue;  // last wins
                if (!first) out.push_back(',');
                first = false;
                json_escape(out, es[i]->first);
                out.push_back(':');
                json_value(out, es[i]->second, depth);
            }
            out.push_back('}');
            break;
        }
        case Value::Container: {
            auto it = cid_index.find(v.cid);
            if (it == cid_index.end() || depth > 64) {
                // an id never seen as an op target: empty container of its type
                switch (v.cid.type) {
                    case CT_TEXT: out += "\"\""; break;
                    case CT_MAP: out += "{}"; break;
                    case CT_COUNTER: out += "0.0"; break;
                    default: out += "[]";
                }
            } else
                json_container(out, it->second, depth + 1);
            break;
        }
    }
}
inline void Doc::json_container(std::string& out, int cidx, int depth) {
    ContainerState& s = cstate(cidx);
    switch (containers[(size_t)cidx].type) {
        case CT_TEXT: json_escape(out, s.text); break;
        case CT_LIST: case CT_MOVABLE: {
            out.push_back('[');
            for (size_t i = 0; i < s.list_values.size(); i++) {
                if (i) out.push_back(',');
                json_value(out, s.list_values[i], depth);
            }
            out.push_back(']');
            break;
        }
        case CT_MAP: {
            out.push_back('{');
            bool first = true;
            for (auto& kv : s.map) {
                if (!kv.second.has) continue;
                if (!first) out.push_back(',');
                first = false;
                json_escape(out, kv.first);
                out.push_back(':');
                json_value(out, kv.second.v, depth);
            }
            out.push_back('}');
            break;
        }
        case CT_TREE: {
            // state/tree_state.rs:1424-1452: nested nodes ordered by (fractional index, lamport, peer)
            std::function<void(bool, ID)> emit = [&](bool root, ID parent) {
                std::vector<const TreeNodeState*> kids;
                for (auto& n : s.tree)
                    if (!n.deleted && n.parent_null == root && (root || n.parent == parent)) kids.push_back(&n);
                std::sort(kids.begin(), kids.end(), [](const TreeNodeState* a, const TreeNodeState* b) {
                    if (a->position != b->position) return a->position < b->position;
                    if (a->lamport != b->lamport) return a->lamport < b->lamport;
                    return a->peer < b->peer;
                });
                out.push_back('[');
                for (size_t i = 0; i < kids.size(); i++) {
                    if (i) out.push_back(',');
                    const TreeNodeState* n = kids[i];
                    out += "{\"children\":";
                    emit(false, n->id);
                    out += ",\"fractional_index\":\"";
                    static const char* HEX = "0123456789ABCDEF";
                    for (unsigned char c : n->position) {
                        out.push_back(HEX[c >> 4]);
                        out.push_back(HEX[c & 15]);
                    }
                    out += "\",\"id\":\"" + id_string(n->id) + "\",\"index\":" + std::to_string(i) + ",\"meta\":";
                    ContainerID mc;
                    mc.root = false;
                    mc.peer = n->id.peer;
                    mc.counter = n->id.counter;
                    mc.type = CT_MAP;
                    json_value(out, Value::container(mc), depth);
                    out += ",\"parent\":";
                    if (root) out += "null";
                    else out += "\"" + id_string(parent) + "\"";
                    out.push_back('}');
                }
                out.push_back(']');
            };
            emit(true, ID{});
            break;
        }
        default: out += "null";
    }
}
// get_deep_value (state.rs:894-924): object keyed by root container name; keys emitted sorted.
inline std::string Doc::to_json() {
    commit();
    ensure_state();
    std::map<std::string, int> roots;
    for (size_t i = 0; i < containers.size(); i++)
        if (containers[i].root) roots[containers[i].name] = (int)i;
    std::string out = "{";
    bool first = true;
    for (auto& kv : roots) {
        if (!first) out.push_back(',');
        first = false;
        json_escape(out, kv.first);
        out.push_back(':');
        json_container(out, kv.second, 0);
    }
    out.push_back('}');
    return out;
}

// ------------------------------------------------------------------ replay
inline void Doc::replay() {
    state.clear();
    state.resize(containers.size());
    for (size_t i = 0; i < containers.size(); i++) state[i].type = containers[i].type;
    inconsistent_delete = false;
    // peer table
    std::vector<PeerID> peer_ids;
    std::map<PeerID, int> pidx;
    for (auto& kv : vv) {
        pidx[kv.first] = (int)peer_ids.size();
        peer_ids.push_back(kv.first);
    }
    int P = (int)peer_ids.size();
    // per-peer change lists (counter order)
    std::vector<std::vector<const Change*>> per_peer((size_t)P);
    for (auto& kv : store)
        for (auto& c : kv.second.changes) per_peer[(size_t)pidx[c.id.peer]].push_back(&c);
    // version vector per change (deps closure), computed in topological order
    std::vector<std::vector<std::vector<Counter>>> cvv((size_t)P);
    for (int p = 0; p < P; p++) cvv[(size_t)p].resize(per_peer[(size_t)p].size());
    auto find_idx = [&](int p, Counter c) -> int {  // index of the change of peer p containing counter c
        auto& v = per_peer[(size_t)p];
        int lo = 0, hi = (int)v.size() - 1, ans = -1;
        while (lo <= hi) {
            int mid = (lo + hi) / 2;
            if (v[(size_t)mid]->id.counter <= c) { ans = mid; lo = mid + 1; } else hi = mid - 1;
        }
        return ans;
    };
    std::vector<size_t> next((size_t)P, 0);
    std::vector<Counter> applied((size_t)P, 0);
    auto ready = [&](const Change* c) {
        for (auto& d : c->deps) {
            auto it = pidx.find(d.peer);
            if (it == pidx.end() || applied[(size_t)it->second] <= d.counter) return false;
        }
        return true;
    };
    // trackers / per-container accumulators
    std::vector<std::unique_ptr<Tracker>> trackers(containers.size());
    struct InsRef { const Op* op; };
    std::vector<std::map<std::pair<int, Counter>, const Op*>> ins_index(containers.size());  // (peer idx, op ctr) -> op
    struct TreeMove { Lamport lamport; PeerID peer; const Op* op; ID op_id; };
    std::vector<std::vector<TreeMove>> tree_moves(containers.size());

    size_t total = 0;
    for (auto& v : per_peer) total += v.size();
    int cur_peer = -1;
    for (size_t done = 0; done < total; done++) {
        int pick = -1;
        if (cur_peer >= 0 && next[(size_t)cur_peer] < per_peer[(size_t)cur_peer].size() &&
            ready(per_peer[(size_t)cur_peer][next[(size_t)cur_peer]]))
            pick = cur_peer;
        else {
            Lamport best = 0;
            for (int p = 0; p < P; p++) {
                if (next[(size_t)p] >= per_peer[(size_t)p].size()) continue;
                const Change* c = per_peer[(size_t)p][next[(size_t)p]];
                if (!ready(c)) continue;
                if (pick < 0 || c->lamport < best) {
                    pick = p;
                    best = c->lamport;
                }
            }
        }
        if (pick < 0) throw std::runtime_error("replay: no ready change (cyclic deps?)");
        cur_peer = pick;
        size_t ci = next[(size_t)pick]++;
        const Change* c = per_peer[(size_t)pick][ci];
        // vv of the change = closure of deps
        std::vector<Counter> v((size_t)P, 0);
        for (auto& d : c->deps) {
            int dp = pidx[d.peer];
            int di = find_idx(dp, d.counter);
            const std::vector<Counter>& dv = cvv[(size_t)dp][(size_t)di];
            for (int q = 0; q < P; q++) v[(size_t)q] = std::max(v[(size_t)q], dv[(size_t)q]);
            v[(size_t)dp] = std::max(v[(size_t)dp], d.counter + 1);
        }
        cvv[(size_t)pick][ci] = v;
        std::vector<char> visited(containers.size(), 0);
        for (const Op& op : c->ops) {
            replay_ops++;
            Lamport op_lamport = c->lamport + (Lamport)(op.counter - c->id.counter);
            uint8_t ct = containers[(size_t)op.cidx].type;
            ContainerState& st = state[(size_t)op.cidx];
            if (op.kind == OP_LIST_INSERT || op.kind == OP_TEXT_INSERT || op.kind == OP_DELETE) {
                if (!trackers[(size_t)op.cidx]) trackers[(size_t)op.cidx].reset(new Tracker(P));
                Tracker& t = *trackers[(size_t)op.cidx];
                if (!visited[(size_t)op.cidx]) {  // diff_calc.rs:215-226: checkout once per change
                    std::vector<Counter> ov = v;
                    ov[(size_t)pick] = std::max(ov[(size_t)pick], op.counter);
                    t.checkout(ov);
                    visited[(size_t)op.cidx] = 1;
                }
                if (op.kind == OP_DELETE) {
                    auto tp = pidx.find(op.del_start.peer);
                    int tpi = tp == pidx.end() ? -2 : tp->second;
                    t.del(pick, op.counter, tpi, op.del_start.counter, op.del_start_pos(), op.atom_len(),
                          op.del_len < 0);
                } else {
                    t.insert(pick, op.counter, op.atom_len(), op.prop, peer_ids[(size_t)pick], peer_ids);
                    ins_index[(size_t)op.cidx][{pick, op.counter}] = &op;
                }
            } else if (op.kind == OP_MAP_SET || op.kind == OP_MAP_DEL) {
                MapEntry& e = st.map[op.key];  // diff_calc.rs:450-473 / map_delta.rs:19-46
                bool win = !e.set || op_lamport > e.lamport || (op_lamport == e.lamport && c->id.peer > e.peer);
                if (win) {
                    e.set = true;
                    e.has = op.kind == OP_MAP_SET;
                    e.v = op.mapval;
                    e.lamport = op_lamport;
                    e.peer = c->id.peer;
                }
            } else if (op.kind == OP_TREE_CREATE || op.kind == OP_TREE_MOVE || op.kind == OP_TREE_DELETE) {
                tree_moves[(size_t)op.cidx].push_back(TreeMove{op_lamport, c->id.peer, &op, ID{c->id.peer, op.counter}});
            } else {
                st.unsupported = true;
            }
            (void)ct;
        }
        applied[(size_t)pick] = c->ctr_end();
    }
    // final version: everything applied
    std::vector<Counter> allv((size_t)P);
    for (int p = 0; p < P; p++) allv[(size_t)p] = applied[(size_t)p];
    for (size_t ci = 0; ci < containers.size(); ci++) {
        ContainerState& st = state[ci];
        if (trackers[ci]) {
            Tracker& t = *trackers[ci];
            t.checkout(allv);
            if (t.inconsistent_delete) inconsistent_delete = true;
            auto& idx = ins_index[ci];
            for (auto& blk : t.blocks)
                for (TSpan* s : blk.spans) {
                    if (s->peer < 0 || !s->active()) continue;
                    // content lookup: op containing (peer, ctr)
                    Counter c = s->ctr;
                    while (c < s->ctr + s->len) {
                        auto it = idx.upper_bound({s->peer, c});
                        --it;
                        const Op* op = it->second;
                        int off = c - op->counter;
                        int take = std::min(op->atom_len() - off, s->ctr + s->len - c);
                        if (op->kind == OP_TEXT_INSERT) {
                            size_t b0 = utf8_byte_index(op->text, (size_t)off);
                            size_t b1 = utf8_byte_index(op->text, (size_t)(off + take));
                            st.text.append(op->text, b0, b1 - b0);
                        } else {
                            for (int k = 0; k < take; k++) st.list_values.push_back(op->values[(size_t)(off + k)]);
                        }
                        for (int k = 0; k < take; k++) st.ids.push_back(ID{peer_ids[(size_t)s->peer], c + k});
                        c += take;
                    }
                }
        }
        if (!tree_moves[ci].empty()) {
            // diff_calc/tree.rs:445-508: apply moves in (lamport, peer) order; a move takes effect iff the
            // new parent is not a descendant of the target (cycle check) and the parent exists/not deleted
            auto& mv = tree_moves[ci];
            std::sort(mv.begin(), mv.end(), [](const TreeMove& a, const TreeMove& b) {
                return a.lamport != b.lamport ? a.lamport < b.lamport : a.peer < b.peer;
            });
            std::map<ID, TreeNodeState> nodes;
            auto is_ancestor = [&](ID anc, ID node) {  // anc is an ancestor of (or equal to) node
                ID cur = node;
                int guard = 0;
                while (true) {
                    if (cur == anc) return true;
                    auto it = nodes.find(cur);
                    if (it == nodes.end() || it->second.parent_null || it->second.deleted) return false;
                    cur = it->second.parent;
                    if (++guard > 10000000) return false;
                }
            };
            for (auto& m : mv) {
                const Op& op = *m.op;
                TreeNodeState n;
                n.id = op.target;
                n.lamport = m.lamport;
                n.peer = m.peer;
                if (op.kind == OP_TREE_DELETE) {
                    auto it = nodes.find(op.target);
                    if (it == nodes.end()) {
                        n.deleted = n.direct_deleted = true;
                        nodes[op.target] = n;
                    } else {
                        it->second.deleted = it->second.direct_deleted = true;
                        it->second.lamport = m.lamport;
                        it->second.peer = m.peer;
                    }
                    continue;
                }
                n.parent_null = op.parent_null;
                n.parent = op.parent;
                n.position = op.position;
                if (!op.parent_null && is_ancestor(op.target, op.parent)) continue;  // would form a cycle
                nodes[op.target] = n;
            }
            // a node is alive iff no ancestor (incl. itself) is deleted and every parent exists
            for (auto& kv : nodes) {
                TreeNodeState n = kv.second;
                ID cur = n.id;
                bool dead = false;
                int guard = 0;
                while (true) {
                    auto it = nodes.find(cur);
                    if (it == nodes.end() || it->second.deleted) { dead = true; break; }
                    if (it->second.parent_null) break;
                    cur = it->second.parent;
                    if (++guard > 10000000) { dead = true; break; }
                }
                n.deleted = dead;
                st.tree.push_back(n);
            }
        }
    }
    state_valid = true;
}

}  // namespace lo
