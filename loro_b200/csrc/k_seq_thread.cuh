// loro_b200 -- phase 5 (large batches): eg-walker (Fugue) integration, ONE THREAD per document.
//
// Same algorithm and data model as k_seq.cuh (reference map there), different mapping to the machine: the
// warp-per-document kernel spends ~900 warp instructions per op on 32-wide scans of one document; when the
// batch holds thousands of documents it is far better to give every lane its own document and let the
// whole batch advance concurrently (10^5 documents = 10^5 threads in flight, every memory latency hidden
// by the other documents).  Nodes are narrower here (16 entries = one 64-byte line per field) so that a
// sequential scan of a node stays inside one or two sectors.
#pragma once
#include "lb_defs.h"
#include "k_seq.cuh"

#define LB_TF 16   // slots per leaf / children per node in the thread-per-document layout

struct TSeq {
    const SeqPools& p;   // kernel parameters stay in the constant bank (__grid_constant__)
    const SeqTables* t;
    __device__ TSeq(const SeqPools& p_, const SeqTables* t_) : p(p_), t(t_) {}
    const DocInfo* di;
    u32 err;
    u32 cidx;
    u64 leaf0, node0, cvv0;
    u32 leaf_cap, node_cap, n_leaves, n_nodes, root, height, first_leaf, unk_leaf;

    __device__ __forceinline__ u64 atom_index(u32 peer, i32 ctr) const {
        return di->atom0 + t->dpeer[di->peer0 + peer].atom_base + (u32)ctr;
    }
    // leaf slot = one uint4 {x: ctr, y: len, z: peer | st << 16, w: unused}; a leaf is 16 slots = 256 B
    __device__ __forceinline__ u64 ls(u32 leaf, int slot) const { return (leaf0 + leaf) * LB_TF + slot; }
    __device__ __forceinline__ u64 ns(u32 nd, int i) const { return (node0 + nd) * LB_TF + i; }
    static __device__ __forceinline__ u32 s_peer(const uint4& v) { return v.z & 0xFFFFu; }
    static __device__ __forceinline__ u32 s_st(const uint4& v) { return v.z >> 16; }
    static __device__ __forceinline__ i32 s_ctr(const uint4& v) { return (i32)v.x; }
    static __device__ __forceinline__ i32 s_len(const uint4& v) { return (i32)v.y; }
    static __device__ __forceinline__ i32 s_vis(const uint4& v) { return (v.z >> 16) == 0 ? (i32)v.y : 0; }
    static __device__ __forceinline__ uint4 mk_slot(u32 peer, i32 ctr, i32 len, u32 st) {
        uint4 v; v.x = (u32)ctr; v.y = (u32)len; v.z = (peer & 0xFFFFu) | (st << 16); v.w = 0; return v;
    }
    // whole-leaf image in thread-local storage: all 16 loads are independent -> one memory round trip
    __device__ __forceinline__ u32 leaf_load(u32 leaf, uint4* L) const {
        const uint4* g = p.tleaf + (leaf0 + leaf) * LB_TF;
#pragma unroll
        for (int i = 0; i < LB_TF; i++) L[i] = g[i];
        return p.leaf_n[leaf0 + leaf];
    }
    __device__ __forceinline__ void leaf_store(u32 leaf, const uint4* L, u32 n) {
        uint4* g = p.tleaf + (leaf0 + leaf) * LB_TF;
#pragma unroll
        for (int i = 0; i < LB_TF; i++) if (i < (int)n) g[i] = L[i];
        p.leaf_n[leaf0 + leaf] = n;
    }
    // node image: child[16] + vis[16] via vector loads
    __device__ __forceinline__ u32 node_load_vis(u32 nd, i32* v) const {
        const int4* g = (const int4*)(p.node_vis + (node0 + nd) * LB_TF);
#pragma unroll
        for (int i = 0; i < LB_TF / 4; i++) { int4 x = g[i]; v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w; }
        return p.node_n[node0 + nd];
    }
    __device__ __forceinline__ void node_load_child(u32 nd, u32* c) const {
        const uint4* g = (const uint4*)(p.node_child + (node0 + nd) * LB_TF);
#pragma unroll
        for (int i = 0; i < LB_TF / 4; i++) { uint4 x = g[i]; c[4 * i] = x.x; c[4 * i + 1] = x.y; c[4 * i + 2] = x.z; c[4 * i + 3] = x.w; }
    }

    __device__ int nd_find(u32 nd, u32 child) const {
        u32 c[LB_TF];
        node_load_child(nd, c);
        u32 n = p.node_n[node0 + nd];
        int idx = -1;
#pragma unroll
        for (int i = 0; i < LB_TF; i++) if (i < (int)n && c[i] == child && idx < 0) idx = i;
        return idx;
    }
    __device__ void add_vis(u32 leaf, i32 delta) {
        if (delta == 0) return;
        u32 child = leaf;
        u32 nd = p.leaf_parent[leaf0 + leaf];
        while (nd != NODE_NONE) {
            int idx = nd_find(nd, child);
            if (idx < 0) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
            p.node_vis[ns(nd, idx)] += delta;
            child = nd;
            nd = p.node_parent[node0 + nd];
        }
    }
    __device__ i32 node_total(u32 nd) const {
        i32 v[LB_TF];
        u32 n = node_load_vis(nd, v);
        i32 s = 0;
#pragma unroll
        for (int i = 0; i < LB_TF; i++) if (i < (int)n) s += v[i];
        return s;
    }
    __device__ void node_insert_no_split(u32 nd, int after, u32 child, i32 vis, bool kids_are_leaves) {
        u32 n = p.node_n[node0 + nd];
        for (int i = (int)n - 1; i > after; i--) {
            p.node_child[ns(nd, i + 1)] = p.node_child[ns(nd, i)];
            p.node_vis[ns(nd, i + 1)] = p.node_vis[ns(nd, i)];
        }
        p.node_child[ns(nd, after + 1)] = child;
        p.node_vis[ns(nd, after + 1)] = vis;
        p.node_n[node0 + nd] = n + 1;
        if (kids_are_leaves) p.leaf_parent[leaf0 + child] = nd; else p.node_parent[node0 + child] = nd;
    }
    __device__ void node_insert(u32 nd, int after, u32 child, i32 vis, bool kids_are_leaves) {
        while (true) {
            u32 n = p.node_n[node0 + nd];
            if (n < LB_TF) { node_insert_no_split(nd, after, child, vis, kids_are_leaves); return; }
            if (n_nodes >= node_cap) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
            u32 nn = n_nodes++;
            const int H = LB_TF / 2;
            for (int i = H; i < LB_TF; i++) {
                u32 c = p.node_child[ns(nd, i)];
                p.node_child[ns(nn, i - H)] = c;
                p.node_vis[ns(nn, i - H)] = p.node_vis[ns(nd, i)];
                if (kids_are_leaves) p.leaf_parent[leaf0 + c] = nn; else p.node_parent[node0 + c] = nn;
            }
            p.node_n[node0 + nd] = H;
            p.node_n[node0 + nn] = H;
            if (after >= H) node_insert_no_split(nn, after - H, child, vis, kids_are_leaves);
            else node_insert_no_split(nd, after, child, vis, kids_are_leaves);
            i32 tot_old = node_total(nd), tot_new = node_total(nn);
            u32 parent = p.node_parent[node0 + nd];
            if (parent == NODE_NONE) {
                if (n_nodes >= node_cap) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
                u32 nr = n_nodes++;
                p.node_child[ns(nr, 0)] = nd;
                p.node_vis[ns(nr, 0)] = tot_old;
                p.node_child[ns(nr, 1)] = nn;
                p.node_vis[ns(nr, 1)] = tot_new;
                p.node_n[node0 + nr] = 2;
                p.node_parent[node0 + nr] = NODE_NONE;
                p.node_parent[node0 + nd] = nr;
                p.node_parent[node0 + nn] = nr;
                root = nr;
                height++;
                return;
            }
            int idx = nd_find(parent, nd);
            if (idx < 0) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
            p.node_vis[ns(parent, idx)] = tot_old;
            child = nn;
            vis = tot_new;
            after = idx;
            nd = parent;
            kids_are_leaves = false;
        }
    }
    static __device__ __forceinline__ int slot_in(const uint4* L, u32 n, u32 peer, i32 c) {
        int r = -1;
#pragma unroll
        for (int s = 0; s < LB_TF; s++)
            if (s < (int)n && r < 0 && s_peer(L[s]) == (peer & 0xFFFFu) && c >= s_ctr(L[s]) && c < s_ctr(L[s]) + s_len(L[s])) r = s;
        return r;
    }
    __device__ int slot_of(u32 leaf, u32 peer, i32 c) const {
        uint4 L[LB_TF];
        u32 n = leaf_load(leaf, L);
        return slot_in(L, n, peer, c);
    }
    __device__ void leaf_split(u32 leaf) {
        if (n_leaves >= leaf_cap) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
        u32 nl = n_leaves++;
        const int H = LB_TF / 2;
        uint4 L[LB_TF];
        leaf_load(leaf, L);
        i32 moved = 0;
        uint4* g = p.tleaf + (leaf0 + nl) * LB_TF;
#pragma unroll
        for (int s = H; s < LB_TF; s++) {
            g[s - H] = L[s];
            moved += s_vis(L[s]);
        }
        for (int s = H; s < LB_TF; s++) {
            u32 pe = s_peer(L[s]);
            if (pe == PEER_UNKNOWN) unk_leaf = nl;
            else {
                u64 a0 = atom_index(pe, s_ctr(L[s]));
                i32 ln = s_len(L[s]);
                for (i32 i = 0; i < ln; i++) p.atom_leaf[a0 + i] = nl;
            }
        }
        p.leaf_n[leaf0 + leaf] = H;
        p.leaf_n[leaf0 + nl] = H;
        p.leaf_next[leaf0 + nl] = p.leaf_next[leaf0 + leaf];
        p.leaf_next[leaf0 + leaf] = nl;
        u32 parent = p.leaf_parent[leaf0 + leaf];
        int idx = nd_find(parent, leaf);
        if (idx < 0) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
        p.node_vis[ns(parent, idx)] -= moved;
        node_insert(parent, idx, nl, moved, true);
    }
    // insert slot value `nv` at index `at` of `leaf` (splitting first if full); returns the leaf that holds it
    __device__ u32 leaf_insert(u32 leaf, int at, uint4 nv) {
        if (p.leaf_n[leaf0 + leaf] >= LB_TF) {
            leaf_split(leaf);
            if (err) return leaf;
            if (at > LB_TF / 2) { leaf = p.leaf_next[leaf0 + leaf]; at -= LB_TF / 2; }
        }
        uint4 L[LB_TF];
        u32 n = leaf_load(leaf, L);
#pragma unroll
        for (int s = LB_TF - 1; s > 0; s--) if (s > at && s <= (int)n) L[s] = L[s - 1];
#pragma unroll
        for (int s = 0; s < LB_TF; s++) if (s == at) L[s] = nv;
        leaf_store(leaf, L, n + 1);
        return leaf;
    }
    // split the span containing atom (peer, c) right before that atom (no-op at a span start)
    __device__ void split_before(u32 peer, i32 c) {
        u32 leaf = p.atom_leaf[atom_index(peer, c)];
        uint4 L[LB_TF];
        u32 n = leaf_load(leaf, L);
        int slot = slot_in(L, n, peer, c);
        if (slot < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
        uint4 sv = L[0];
#pragma unroll
        for (int s = 0; s < LB_TF; s++) if (s == slot) sv = L[s];
        i32 ctr = s_ctr(sv);
        if (ctr == c) return;
        if (n >= LB_TF) {
            leaf_split(leaf);
            if (err) return;
            leaf = p.atom_leaf[atom_index(peer, c)];
            n = leaf_load(leaf, L);
            slot = slot_in(L, n, peer, c);
            if (slot < 0) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
        }
        i32 len = s_len(sv);
        u32 st = s_st(sv);
        i32 k = c - ctr;
        // shorten the left part in place, open the next slot for the right part
#pragma unroll
        for (int s = LB_TF - 1; s > 0; s--) if (s > slot + 1 && s <= (int)n) L[s] = L[s - 1];
#pragma unroll
        for (int s = 0; s < LB_TF; s++) {
            if (s == slot) L[s] = mk_slot(peer, ctr, k, st);
            if (s == slot + 1) L[s] = mk_slot(peer, c, len - k, st);
        }
        leaf_store(leaf, L, n + 1);
        u64 a_old = atom_index(peer, ctr), a_new = atom_index(peer, c);
        p.a_ol_peer[a_new] = (u16)peer;
        p.a_ol_ctr[a_new] = c - 1;
        p.a_or_peer[a_new] = p.a_or_peer[a_old];
        p.a_or_ctr[a_new] = p.a_or_ctr[a_old];
    }
    __device__ void range_set(u32 peer, i32 lo, i32 hi, int set_future, int del_diff) {
        i32 c = lo;
        while (c < hi && !err) {
            u64 ai = atom_index(peer, c);
            u32 leaf = p.atom_leaf[ai];
            if (leaf == LEAF_NONE) { c++; continue; }
            uint4 L[LB_TF];
            u32 n = leaf_load(leaf, L);
            int slot = slot_in(L, n, peer, c);
            if (slot < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
            uint4 sv = L[0];
#pragma unroll
            for (int s = 0; s < LB_TF; s++) if (s == slot) sv = L[s];
            if (s_ctr(sv) != c || c + s_len(sv) > hi) {
                // boundaries do not line up with the span: cut, then look again
                if (s_ctr(sv) != c) split_before(peer, c);
                if (!err && s_ctr(sv) + s_len(sv) > hi) split_before(peer, hi);
                if (err) return;
                leaf = p.atom_leaf[ai];
                n = leaf_load(leaf, L);
                slot = slot_in(L, n, peer, c);
                if (slot < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
#pragma unroll
                for (int s = 0; s < LB_TF; s++) if (s == slot) sv = L[s];
            }
            i32 s_len_ = s_len(sv);
            u32 st = s_st(sv);
            u32 nst = st;
            if (set_future == 1) nst |= ST_FUTURE;
            if (set_future == 0) nst &= ~(u32)ST_FUTURE;
            nst = (nst & ST_FUTURE) | (((nst & 0x7FFF) + del_diff) & 0x7FFF);
            p.tleaf[ls(leaf, slot)] = mk_slot(peer, c, s_len_, nst);
            i32 before = st == 0 ? s_len_ : 0, after = nst == 0 ? s_len_ : 0;
            add_vis(leaf, after - before);
            c += s_len_;
        }
    }
    __device__ void toggle_ops(u32 peer, i32 a, i32 b, int dir) {
        i32 c = a;
        while (c < b && !err) {
            u32 row = t->atom_row[atom_index(peer, c)];
            i32 r_ctr = t->op_counter[row];
            i32 r_end = r_ctr + (i32)t->op_len[row];
            i32 hi = r_end < b ? r_end : b;
            u8 kind = t->op_kind[row];
            if (t->op_cidx[row] == cidx) {
                if (kind == OPK_SEQ_INS) range_set(peer, c, hi, dir < 0 ? 1 : 0, 0);
                else if (kind == OPK_SEQ_DEL) {
                    u32 dl = t->op_del[row];
                    i32 dlen = t->del_len[dl];
                    i32 n = dlen < 0 ? -dlen : dlen;
                    const BlockInfo& bi = t->blocks[t->ch_block[t->op_change[row]]];
                    u32 tp = t->peer_map[bi.peer0 + t->del_peer_idx[dl]];
                    i32 tc = t->del_counter[dl];
                    i32 t0, t1;
                    if (dlen > 0) { t0 = tc + (c - r_ctr); t1 = tc + (hi - r_ctr); }
                    else { t0 = tc + (n - (hi - r_ctr)); t1 = tc + (n - (c - r_ctr)); }
                    range_set(tp, t0, t1, -1, dir);
                }
            }
            c = hi;
        }
    }
    __device__ void checkout(const i32* vv, u32 own_peer, i32 own_ctr) {
        u32 P = di->P;
        for (u32 q = 0; q < P && !err; q++) {
            i32 tgt = vv ? vv[q] : 0;
            if (q == own_peer && own_ctr > tgt) tgt = own_ctr;
            i32 cur = p.cvv[cvv0 + q];
            if (cur > tgt) toggle_ops(q, tgt, cur, -1);
            else if (cur < tgt) toggle_ops(q, cur, tgt, +1);
            if (cur != tgt) p.cvv[cvv0 + q] = tgt;
        }
    }
    __device__ u64 order_key(u32 leaf, int slot) const {
        u64 key = (u64)slot;
        int shift = 6;
        u32 child = leaf;
        u32 nd = p.leaf_parent[leaf0 + leaf];
        while (nd != NODE_NONE) {
            key |= (u64)nd_find(nd, child) << shift;
            shift += 6;
            child = nd;
            nd = p.node_parent[node0 + nd];
        }
        return key;
    }
    __device__ u64 order_key_of_atom(u32 peer, i32 c) const {
        if (peer == PEER_UNKNOWN) {
            u32 n = p.leaf_n[leaf0 + unk_leaf];
            for (u32 s = 0; s < n; s++)
                if (s_peer(p.tleaf[ls(unk_leaf, s)]) == PEER_UNKNOWN) return order_key(unk_leaf, (int)s);
            return ~0ull;
        }
        u32 leaf = p.atom_leaf[atom_index(peer, c)];
        return order_key(leaf, slot_of(leaf, peer, c));
    }
    __device__ void atom_origin_left(u32 peer, i32 c, u16* op, i32* oc) const {
        if (peer == PEER_UNKNOWN) { *op = PEER_NONE; *oc = -1; return; }
        u32 leaf = p.atom_leaf[atom_index(peer, c)];
        int slot = slot_of(leaf, peer, c);
        if (s_ctr(p.tleaf[ls(leaf, slot)]) == c) { u64 a = atom_index(peer, c); *op = p.a_ol_peer[a]; *oc = p.a_ol_ctr[a]; }
        else { *op = (u16)peer; *oc = c - 1; }
    }

    // CrdtRope::insert (crdt_rope.rs:43-227)
    __device__ void insert(u32 peer, i32 ctr, i32 len, i32 pos) {
        u32 leaf = first_leaf;
        int slot = 0;
        i32 off = 0;
        u16 ol_peer = PEER_NONE;
        i32 ol_ctr = -1;
        u32 cur_peer = PEER_NONE;
        i32 cur_ctr = 0, cur_len = 0;
        uint4 L[LB_TF];
        u32 ln_ = 0;
        if (pos > 0) {
            i32 rem = pos;
            u32 nd = root;
            for (u32 lvl = height; lvl >= 1; lvl--) {
                i32 v[LB_TF];
                u32 n = node_load_vis(nd, v);
                int idx = -1;
#pragma unroll
                for (int i = 0; i < LB_TF; i++) {
                    if (i < (int)n && idx < 0) {
                        if (rem <= v[i]) idx = i; else rem -= v[i];
                    }
                }
                if (idx < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
                nd = p.node_child[ns(nd, idx)];
            }
            leaf = nd;
            ln_ = leaf_load(leaf, L);
            int found = -1;
#pragma unroll
            for (int s = 0; s < LB_TF; s++) {
                if (s < (int)ln_ && found < 0) {
                    i32 v = s_vis(L[s]);
                    if (v > 0 && rem <= v) found = s; else rem -= v;
                }
            }
            if (found < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
            slot = found;
            off = rem;
            uint4 sv = L[0];
#pragma unroll
            for (int s = 0; s < LB_TF; s++) if (s == slot) sv = L[s];
            cur_peer = s_peer(sv);
            cur_ctr = s_ctr(sv);
            cur_len = s_len(sv);
            if (cur_peer == PEER_UNKNOWN) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
            ol_peer = (u16)cur_peer;
            ol_ctr = cur_ctr + off - 1;
        } else {
            ln_ = leaf_load(leaf, L);
        }
        // origin_right + in-between (future) spans
        u16 or_peer = PEER_NONE;
        i32 or_ctr = -1;
        bool pr_valid = false;
        u32 pr_leaf = 0;
        int pr_slot = 0;
        u32 n_between = 0;
        int scan_from = (pos > 0 && off >= cur_len) ? slot + 1 : slot;
        if (pos > 0 && off < cur_len) {
            or_peer = (u16)cur_peer;
            or_ctr = cur_ctr + off;
        } else {
            u32 l2 = leaf;
            int from = scan_from;
            bool found = false;
            u32 n2 = ln_;
            while (true) {
#pragma unroll
                for (int s = 0; s < LB_TF; s++) {
                    if (s >= from && s < (int)n2 && !found) {
                        if (!(s_st(L[s]) & ST_FUTURE)) {
                            or_peer = (u16)s_peer(L[s]);
                            or_ctr = s_ctr(L[s]);
                            pr_leaf = l2;
                            pr_slot = s;
                            pr_valid = true;
                            found = true;
                        } else n_between++;
                    }
                }
                if (found) break;
                l2 = p.leaf_next[leaf0 + l2];
                if (l2 == LEAF_NONE) break;
                from = 0;
                n2 = leaf_load(l2, L);
            }
            if (l2 != leaf) ln_ = leaf_load(leaf, L);   // restore the image of the cursor leaf
        }
        bool after_valid = false;
        u32 after_peer = 0;
        i32 after_ctr = 0;
        if (n_between) {
            u64 pr_key = 0;
            if (pr_valid) {
                u16 e_olp;
                i32 e_olc;
                if (or_peer == PEER_UNKNOWN) { e_olp = PEER_NONE; e_olc = -1; }
                else { u64 a = atom_index(or_peer, or_ctr); e_olp = p.a_ol_peer[a]; e_olc = p.a_ol_ctr[a]; }
                pr_valid = e_olp == ol_peer && (ol_peer == PEER_NONE || e_olc == ol_ctr);
                if (pr_valid) pr_key = order_key(pr_leaf, pr_slot);
            }
            bool scanning = false;
            u64 my_peer_id = t->dpeer[di->peer0 + peer].id;
            u64 first_key = 0;
            bool have_first = false;
            u32 l2 = leaf;
            int from = scan_from;
            u32 seen = 0;
            bool stop = false;
            while (l2 != LEAF_NONE && seen < n_between && !stop) {
                u32 n = p.leaf_n[leaf0 + l2];
                for (int s = from; s < (int)n && seen < n_between && !stop; s++) {
                    uint4 ov = p.tleaf[ls(l2, s)];
                    u32 o_peer = s_peer(ov);
                    i32 o_ctr = s_ctr(ov);
                    seen++;
                    u64 o_key = order_key(l2, s);
                    if (!have_first) { first_key = o_key; have_first = true; }
                    u64 oa = atom_index(o_peer, o_ctr);
                    u16 o_olp = p.a_ol_peer[oa];
                    i32 o_olc = p.a_ol_ctr[oa];
                    bool same_ol = o_olp == ol_peer && (ol_peer == PEER_NONE || o_olc == ol_ctr);
                    if (!same_ol) {
                        bool in_visited = false;
                        if (o_olp != PEER_NONE && o_olp != PEER_UNKNOWN && p.atom_leaf[atom_index(o_olp, o_olc)] != LEAF_NONE) {
                            u64 lk = order_key_of_atom(o_olp, o_olc);
                            in_visited = lk >= first_key && lk < o_key;
                        }
                        if (!in_visited) { stop = true; break; }
                    }
                    if (same_ol) {
                        u16 o_orp = p.a_or_peer[oa];
                        i32 o_orc = p.a_or_ctr[oa];
                        bool same_or = o_orp == or_peer && (or_peer == PEER_NONE || o_orc == or_ctr);
                        u64 o_peer_id = t->dpeer[di->peer0 + o_peer].id;
                        if (same_or) {
                            if (o_peer_id > my_peer_id) { stop = true; break; }
                            scanning = false;
                        } else {
                            bool o_pr = false;
                            u64 o_pr_key = 0;
                            if (o_orp != PEER_NONE) {
                                u16 e_olp;
                                i32 e_olc;
                                atom_origin_left(o_orp, o_orc, &e_olp, &e_olc);
                                if (e_olp == ol_peer && (ol_peer == PEER_NONE || e_olc == ol_ctr)) {
                                    o_pr = true;
                                    o_pr_key = order_key_of_atom(o_orp, o_orc);
                                }
                            }
                            int cmp;
                            if (o_pr && pr_valid) cmp = o_pr_key < pr_key ? -1 : (o_pr_key > pr_key ? 1 : 0);
                            else if (o_pr) cmp = -1;
                            else if (pr_valid) cmp = 1;
                            else cmp = 0;
                            if (cmp < 0) scanning = true;
                            else if (cmp == 0 && o_peer_id > my_peer_id) { stop = true; break; }
                            else scanning = false;
                        }
                    }
                    if (!scanning) { after_valid = true; after_peer = o_peer; after_ctr = o_ctr; }
                }
                l2 = p.leaf_next[leaf0 + l2];
                from = 0;
            }
        }
        u32 tgt_leaf;
        int at;
        if (after_valid) {
            tgt_leaf = p.atom_leaf[atom_index(after_peer, after_ctr)];
            at = slot_of(tgt_leaf, after_peer, after_ctr) + 1;
        } else if (pos == 0) {
            tgt_leaf = first_leaf;
            at = 0;
        } else if (off < cur_len) {
            split_before(cur_peer, cur_ctr + off);
            if (err) return;
            tgt_leaf = p.atom_leaf[atom_index(cur_peer, cur_ctr + off)];
            at = slot_of(tgt_leaf, cur_peer, cur_ctr + off);
        } else {
            tgt_leaf = leaf;
            at = slot + 1;
        }
        if (at < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
        tgt_leaf = leaf_insert(tgt_leaf, at, mk_slot(peer, ctr, len, 0));
        if (err) return;
        u64 a0 = atom_index(peer, ctr);
        p.a_ol_peer[a0] = ol_peer;
        p.a_ol_ctr[a0] = ol_ctr;
        p.a_or_peer[a0] = or_peer;
        p.a_or_ctr[a0] = or_ctr;
        for (i32 i = 0; i < len; i++) p.atom_leaf[a0 + i] = tgt_leaf;
        add_vis(tgt_leaf, len);
    }

    __device__ void load_container(u32 c) {
        DocContainer& dc = t->dcont[di->cid0 + c];
        cidx = c;
        leaf0 = dc.leaf0;
        node0 = dc.node0;
        leaf_cap = dc.leaf_cap;
        node_cap = dc.node_cap;
        n_leaves = dc.n_leaves;
        n_nodes = dc.n_nodes;
        root = dc.root;
        height = dc.height;
        first_leaf = dc.first_leaf;
        cvv0 = dc.cvv0;
        unk_leaf = dc.unk_sid;
        if (n_leaves == 0) {
            if (leaf_cap < 1 || node_cap < 1) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
            n_leaves = 1;
            n_nodes = 1;
            root = 0;
            height = 1;
            first_leaf = 0;
            unk_leaf = 0;
            p.tleaf[ls(0, 0)] = mk_slot(PEER_UNKNOWN, 0, UNKNOWN_LEN, 0);
            p.leaf_n[leaf0] = 1;
            p.leaf_parent[leaf0] = 0;
            p.leaf_next[leaf0] = LEAF_NONE;
            p.node_child[ns(0, 0)] = 0;
            p.node_vis[ns(0, 0)] = UNKNOWN_LEN;
            p.node_n[node0] = 1;
            p.node_parent[node0] = NODE_NONE;
            for (u32 q = 0; q < di->P; q++) p.cvv[cvv0 + q] = 0;
        }
    }
    __device__ void store_container() {
        if (cidx == 0xFFFFFFFFu) return;
        DocContainer& dc = t->dcont[di->cid0 + cidx];
        dc.n_leaves = n_leaves;
        dc.n_nodes = n_nodes;
        dc.root = root;
        dc.height = height;
        dc.first_leaf = first_leaf;
        dc.unk_sid = unk_leaf;
    }
    __device__ void emit_output() {
        DocContainer& dc = t->dcont[di->cid0 + cidx];
        u32 n_out = 0, total = 0;
        for (u32 l2 = first_leaf; l2 != LEAF_NONE; l2 = p.leaf_next[leaf0 + l2]) {
            u32 n = p.leaf_n[leaf0 + l2];
            for (u32 s = 0; s < n; s++) {
                uint4 sv = p.tleaf[ls(l2, s)];
                u32 pe = s_peer(sv);
                if (s_st(sv) != 0 || pe == PEER_UNKNOWN) continue;
                i32 ct = s_ctr(sv), ln = s_len(sv);
                if (n_out < dc.out_cap) {
                    u32 row = t->atom_row[atom_index(pe, ct)];
                    p.out_row[dc.out0 + n_out] = row;
                    p.out_off[dc.out0 + n_out] = (u32)(ct - t->op_counter[row]);
                    p.out_len[dc.out0 + n_out] = (u32)ln;
                }
                n_out++;
                total += (u32)ln;
            }
        }
        if (n_out > dc.out_cap) err = LB_ERR(DOC_ERR_CAPACITY);
        dc.n_out = n_out < dc.out_cap ? n_out : dc.out_cap;
        dc.seq_len = total;
    }
};

// one thread per document
__global__ void k_seq_integrate_thread(DocInfo* __restrict__ docs, u32 n_docs, const __grid_constant__ SeqPools pools,
                                       const __grid_constant__ SeqTables tables) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    DocInfo& di = docs[d];
    if (di.code != DOC_OK || di.n_applied == 0) return;
    bool any = false;
    for (u32 c = 0; c < di.C; c++)
        if (tables.dcont[di.cid0 + c].leaf_cap) any = true;
    if (!any) return;
    TSeq s(pools, &tables);
    s.di = &di;
    s.err = 0;
    s.cidx = 0xFFFFFFFFu;
    for (u64 i = 0; i < di.atom_total; i++) pools.atom_leaf[di.atom0 + i] = LEAF_NONE;
    for (u32 c = 0; c < di.C; c++) pools.cont_epoch[di.cid0 + c] = 0xFFFFFFFFu;
    u32 P = di.P;
    u32 prev_peer = 0xFFFFFFFFu;
    u32 cur_epoch = 0xFFFFFFFFu;   // epoch of the active container (kept in a register)
    // Flat state machine: every loop iteration handles exactly one op row (or one change transition), so
    // the lanes of a warp -- which walk different documents -- reconverge at every row instead of waiting
    // for the longest change of the warp.
    u32 k = 0, kcur = 0;           // next change to open / index of the open change
    u32 ch = 0, peer = 0, nr = 0, r = 0;
    u64 r0 = 0;
    bool chain = false;
    const i32* vv = nullptr;
    u8 n_kind = 0; u32 n_c = 0; i32 n_ctr = 0, n_len = 0, n_prop = 0; u32 n_del = 0;
    while (true) {
        if (r >= nr) {
            if (k >= di.n_applied || s.err) break;
            if (k > 0) prev_peer = peer;
            kcur = k;
            ch = tables.ch_walk[di.ch0 + k];
            peer = tables.ch_peer[ch];
            r0 = tables.ch_op0[ch];
            nr = tables.ch_nops[ch];
            chain = tables.ch_dep_self[ch] && tables.ch_ndeps[ch] == 0 && prev_peer == peer && k > 0;
            vv = nullptr;
            r = 0;
            k++;
            if (nr) {
                n_kind = tables.op_kind[r0]; n_c = tables.op_cidx[r0]; n_ctr = tables.op_counter[r0];
                n_len = (i32)tables.op_len[r0]; n_prop = tables.op_prop[r0]; n_del = tables.op_del[r0];
            }
            continue;
        }
        u8 kind = n_kind; u32 c = n_c; i32 ctr = n_ctr, len = n_len, prop = n_prop; u32 dl = n_del;
        r++;
        if (r < nr) {   // request the next row before working on this one
            u64 nx = r0 + r;
            n_kind = tables.op_kind[nx]; n_c = tables.op_cidx[nx]; n_ctr = tables.op_counter[nx];
            n_len = (i32)tables.op_len[nx]; n_prop = tables.op_prop[nx]; n_del = tables.op_del[nx];
        }
        if (kind != OPK_SEQ_INS && kind != OPK_SEQ_DEL) continue;
        if (c != s.cidx) {
            if (s.cidx != 0xFFFFFFFFu) pools.cont_epoch[di.cid0 + s.cidx] = cur_epoch;
            s.store_container();
            s.load_container(c);
            if (s.err) break;
            cur_epoch = pools.cont_epoch[di.cid0 + c];
        }
        if (cur_epoch != kcur) {
            if (!(chain && cur_epoch == kcur - 1)) {
                if (!vv) {
                    const DocPeer& dp = tables.dpeer[di.peer0 + peer];
                    i32 cc = tables.ch_counter[ch];
                    u32 lo = 0, hi = dp.ch_count;
                    while (hi - lo > 1) {
                        u32 mid = (lo + hi) >> 1;
                        if (tables.ch_counter[tables.ch_order[di.ch0 + dp.ch_first + mid]] <= cc) lo = mid; else hi = mid;
                    }
                    vv = tables.ch_vv + di.vv0 + (u64)(dp.ch_first + lo) * P;
                }
                s.checkout(vv, peer, ctr);
            }
            cur_epoch = kcur;
        }
        if (kind == OPK_SEQ_INS) s.insert(peer, ctr, len, prop);
        else {
            const BlockInfo& bi = tables.blocks[tables.ch_block[ch]];
            u32 tp = tables.peer_map[bi.peer0 + tables.del_peer_idx[dl]];
            i32 tc = tables.del_counter[dl];
            s.range_set(tp, tc, tc + len, -1, +1);
        }
        // current_vv follows the tracker's own ops; in causal order this entry only grows
        pools.cvv[s.cvv0 + peer] = ctr + len;
    }
    for (u32 c = 0; c < di.C && !s.err; c++) {
        const DocContainer& dc = tables.dcont[di.cid0 + c];
        if (!dc.leaf_cap || (dc.n_leaves == 0 && c != s.cidx)) continue;
        if (c != s.cidx) {
            s.store_container();
            s.load_container(c);
        }
        for (u32 q = 0; q < P && !s.err; q++) {
            i32 tgt = tables.dpeer[di.peer0 + q].end_counter;
            i32 cur = pools.cvv[s.cvv0 + q];
            if (cur > tgt) s.toggle_ops(q, tgt, cur, -1);
            else if (cur < tgt) s.toggle_ops(q, cur, tgt, +1);
            pools.cvv[s.cvv0 + q] = tgt;
        }
        if (!s.err) s.emit_output();
    }
    s.store_container();
    if (s.err) di.code = s.err;
}
