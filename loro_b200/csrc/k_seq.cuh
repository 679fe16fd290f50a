// loro_b200 -- phase 5: eg-walker (Fugue) integration of List/Text containers, one warp per document.
//
// Replaces (reference, relative to crates/loro-internal/src):
//   container/richtext/tracker.rs:84-232 (insert/delete), :330-526 (checkout: retreat / forward)
//   container/richtext/tracker/crdt_rope.rs:43-227 (Fugue origin reconstruction + sibling scan),
//     :236-315 (delete), :325-361 (status updates), :542-652 (active-length queries)
//   container/richtext/fugue_span.rs:192-386 (span status: future / delete_times ; slicing)
//   container/richtext/tracker/id_to_cursor.rs (id -> span lookup)  ->  dense atom -> leaf array
//   diff_calc.rs:175-236,585-620 (per-change checkout then apply)
//
// Data structure: per container a B+tree.
//   * leaves (HBM, SoA): up to 32 spans each as (peer, counter, len, state); one warp reads a whole leaf with
//     four coalesced accesses and resolves "k-th visible atom" with a shuffle scan + ballot;
//   * internal nodes: up to 32 (child, visible-length) pairs; the first NS nodes of the container being
//     integrated live in SHARED memory (one private region per warp), the rest spill to HBM, so the descent
//     and the visible-length updates of an op normally touch no global memory at all;
//   * per document, dense atom-indexed arrays (HBM): atom -> leaf, atom -> op row, and the Fugue origins of
//     every span start.  Spans never merge or disappear, they only split, and a split never moves atoms, so
//     only a leaf split rewrites atom -> leaf entries.
// Deletes are applied by target id (for well-formed histories the reference's by-position deletion hits the
// same atoms).  Every function is warp-synchronous: all 32 lanes call it with identical arguments; a
// __syncwarp() separates reads by all lanes from a following write by one lane to the same location.
#pragma once
#include "lb_defs.h"

#define NODE_NONE 0xFFFFFFFFu
#define LEAF_NONE 0xFFFFFFFFu
#define ST_FUTURE 0x8000u
#ifndef LB_SEQ_NS
#define LB_SEQ_NS 8           // internal nodes cached in shared memory per warp (more smem = less L1)
#endif
#define LB_SEQ_WARPS 4        // warps (documents) per CTA

struct SeqPools {
    // leaves
    u32* leaf_ps;    // warp layout: peer | state << 16 per slot
    i32* leaf_ctr; i32* leaf_len; u32* leaf_n; u32* leaf_parent; u32* leaf_next;
    uint4* tleaf;   // thread-per-document layout: one uint4 per slot (k_seq_thread.cuh)
    // internal nodes (global home; nodes < NS of the active container are cached in shared memory)
    u32* node_child; i32* node_vis; u32* node_n; u32* node_parent;
    // per-document atom-indexed arrays
    u32* atom_leaf;                                  // LEAF_NONE = not an inserted list/text atom (yet)
    u16* a_ol_peer; i32* a_ol_ctr; u16* a_or_peer; i32* a_or_ctr;   // origins, valid at span starts
    i32* cvv;
    u32* cont_epoch;   // per container: last walk index that checked out / applied an op
    u32* out_row; u32* out_off; u32* out_len;
};

struct SeqTables {
    const DocPeer* dpeer; DocContainer* dcont;
    const u32* ch_walk; const u64* ch_op0; const u32* ch_nops; const u16* ch_peer; const i32* ch_vv;
    const u32* ch_order; const i32* ch_counter; const u32* ch_ndeps; const u8* ch_dep_self;
    const u8* op_kind; const u32* op_cidx; const i32* op_prop; const u32* op_len; const i32* op_counter;
    const u32* op_del; const u32* op_change;
    const u32* del_peer_idx; const i32* del_counter; const i32* del_len;
    const u32* peer_map; const BlockInfo* blocks; const u32* ch_block;
    const u32* atom_row;
};

struct SeqSmem {   // one per warp
    u32 child[LB_SEQ_NS][32];
    i32 vis[LB_SEQ_NS][32];
    u32 n[LB_SEQ_NS];
    u32 parent[LB_SEQ_NS];
    u32 abase[32];   // atom_base of the document's first 32 peers
};

struct LeafImg { u32 n; u16 peer; i32 ctr; i32 len; u16 st; };   // lane i holds slot i of a leaf

struct Seq {
    const SeqPools& p;   // kernel parameters stay in the constant bank (__grid_constant__)
    const SeqTables* t;
    __device__ Seq(const SeqPools& p_, const SeqTables* t_) : p(p_), t(t_) {}
    const DocInfo* di;
    SeqSmem* sm;
    int lane;
    u32 err;
    // current container
    u32 cidx;
    u64 leaf0, node0, cvv0;
    u32 leaf_cap, node_cap, n_leaves, n_nodes, root, height, first_leaf, unk_leaf;

    u64 atom0;
    __device__ __forceinline__ u64 atom_index(u32 peer, i32 ctr) const {
        u32 base = peer < 32 ? sm->abase[peer] : t->dpeer[di->peer0 + peer].atom_base;
        return atom0 + base + (u32)ctr;
    }
    // ---- node accessors (shared-memory cache for node ids < NS)
    __device__ __forceinline__ u32 nd_n(u32 nd) const { return nd < LB_SEQ_NS ? sm->n[nd] : p.node_n[node0 + nd]; }
    __device__ __forceinline__ void nd_set_n(u32 nd, u32 v) { if (nd < LB_SEQ_NS) sm->n[nd] = v; else p.node_n[node0 + nd] = v; }
    __device__ __forceinline__ u32 nd_parent(u32 nd) const { return nd < LB_SEQ_NS ? sm->parent[nd] : p.node_parent[node0 + nd]; }
    __device__ __forceinline__ void nd_set_parent(u32 nd, u32 v) { if (nd < LB_SEQ_NS) sm->parent[nd] = v; else p.node_parent[node0 + nd] = v; }
    __device__ __forceinline__ u32 nd_child(u32 nd, int i) const { return nd < LB_SEQ_NS ? sm->child[nd][i] : p.node_child[(node0 + nd) * 32 + i]; }
    __device__ __forceinline__ i32 nd_vis(u32 nd, int i) const { return nd < LB_SEQ_NS ? sm->vis[nd][i] : p.node_vis[(node0 + nd) * 32 + i]; }
    __device__ __forceinline__ void nd_set(u32 nd, int i, u32 c, i32 v) {
        if (nd < LB_SEQ_NS) { sm->child[nd][i] = c; sm->vis[nd][i] = v; }
        else { p.node_child[(node0 + nd) * 32 + i] = c; p.node_vis[(node0 + nd) * 32 + i] = v; }
    }
    __device__ __forceinline__ void nd_add_vis(u32 nd, int i, i32 d) {
        if (nd < LB_SEQ_NS) sm->vis[nd][i] += d; else p.node_vis[(node0 + nd) * 32 + i] += d;
    }
    __device__ __forceinline__ int nd_find(u32 nd, u32 child) {   // index of `child` in node (warp-wide)
        u32 n = nd_n(nd);
        u32 c = lane < (int)n ? nd_child(nd, lane) : NODE_NONE;
        unsigned m = __ballot_sync(LB_FULL, c == child);
        return __ffs(m) - 1;
    }
    // Parent links carry the position inside the parent: link = (parent << 5) | index, NODE_NONE for the root.
    // ---- add `delta` visible atoms on the path leaf -> root
    __device__ void add_vis(u32 leaf, i32 delta) {
        if (delta == 0) return;
        u32 link = p.leaf_parent[leaf0 + leaf];
        while (link != NODE_NONE) {
            u32 nd = link >> 5;
            if (lane == 0) nd_add_vis(nd, (int)(link & 31), delta);
            link = nd_parent(nd);
        }
        __syncwarp();
    }
    // ---- insert (child, vis) into node `nd` right after index `after`; room must exist
    __device__ void node_insert_no_split(u32 nd, int after, u32 child, i32 vis, bool kids_are_leaves) {
        u32 n = nd_n(nd);
        u32 c = lane < (int)n ? nd_child(nd, lane) : 0;
        i32 v = lane < (int)n ? nd_vis(nd, lane) : 0;
        u32 c_up = __shfl_up_sync(LB_FULL, c, 1);
        i32 v_up = __shfl_up_sync(LB_FULL, v, 1);
        int at = after + 1;
        if (lane == at) { c = child; v = vis; }
        else if (lane > at) { c = c_up; v = v_up; }
        if (lane <= (int)n) nd_set(nd, lane, c, v);
        if (lane >= at && lane <= (int)n) {   // the new child and every child that moved one place up
            u32 link = (nd << 5) | (u32)lane;
            if (kids_are_leaves) p.leaf_parent[leaf0 + c] = link; else nd_set_parent(c, link);
        }
        if (lane == 0) nd_set_n(nd, n + 1);
        __syncwarp();
    }
    __device__ i32 node_total(u32 nd) {
        u32 n = nd_n(nd);
        return warp_sum(lane < (int)n ? nd_vis(nd, lane) : 0);
    }
    // ---- insert with splits propagating upward
    __device__ void node_insert(u32 nd, int after, u32 child, i32 vis, bool kids_are_leaves) {
        while (true) {
            u32 n = nd_n(nd);
            if (n < 32) { node_insert_no_split(nd, after, child, vis, kids_are_leaves); return; }
            if (n_nodes >= node_cap) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
            __syncwarp();
            u32 nn = n_nodes++;
            u32 c = nd_child(nd, lane);
            i32 v = nd_vis(nd, lane);
            __syncwarp();
            if (lane >= 16) {
                nd_set(nn, lane - 16, c, v);
                u32 link = (nn << 5) | (u32)(lane - 16);
                if (kids_are_leaves) p.leaf_parent[leaf0 + c] = link; else nd_set_parent(c, link);
            }
            if (lane == 0) { nd_set_n(nd, 16); nd_set_n(nn, 16); }
            __syncwarp();
            if (after >= 16) node_insert_no_split(nn, after - 16, child, vis, kids_are_leaves);
            else node_insert_no_split(nd, after, child, vis, kids_are_leaves);
            i32 tot_old = node_total(nd), tot_new = node_total(nn);
            u32 plink = nd_parent(nd);
            __syncwarp();   // every lane has read the parent link before it is rewritten
            if (plink == NODE_NONE) {
                if (n_nodes >= node_cap) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
                u32 nr = n_nodes++;
                if (lane == 0) {
                    nd_set(nr, 0, nd, tot_old);
                    nd_set(nr, 1, nn, tot_new);
                    nd_set_n(nr, 2);
                    nd_set_parent(nr, NODE_NONE);
                    nd_set_parent(nd, (nr << 5) | 0u);
                    nd_set_parent(nn, (nr << 5) | 1u);
                }
                __syncwarp();
                root = nr;
                height++;
                return;
            }
            u32 parent = plink >> 5;
            int idx = (int)(plink & 31);
            if (lane == 0) nd_set(parent, idx, nd, tot_old);
            __syncwarp();
            child = nn;
            vis = tot_new;
            after = idx;
            nd = parent;
            kids_are_leaves = false;
        }
    }
    // ---- leaf helpers
    __device__ __forceinline__ LeafImg leaf_load(u32 leaf) {
        LeafImg L;
        L.n = p.leaf_n[leaf0 + leaf];
        u64 b = (leaf0 + leaf) * 32 + lane;
        bool in = lane < (int)L.n;
        u32 ps = in ? p.leaf_ps[b] : ((u32)PEER_NONE | (1u << 16));
        L.peer = (u16)ps;
        L.st = (u16)(ps >> 16);
        L.ctr = in ? p.leaf_ctr[b] : 0;
        L.len = in ? p.leaf_len[b] : 0;
        return L;
    }
    __device__ __forceinline__ void leaf_store(u32 leaf, const LeafImg& L, u32 new_n) {
        u64 b = (leaf0 + leaf) * 32 + lane;
        if (lane < (int)new_n) {
            p.leaf_ps[b] = (u32)L.peer | ((u32)L.st << 16);
            p.leaf_ctr[b] = L.ctr;
            p.leaf_len[b] = L.len;
        }
        if (lane == 0) p.leaf_n[leaf0 + leaf] = new_n;
        __syncwarp();
    }
    // slot of the span containing atom (peer, c) inside a loaded leaf image
    __device__ __forceinline__ int slot_of(const LeafImg& L, u32 peer, i32 c) {
        unsigned m = __ballot_sync(LB_FULL, lane < (int)L.n && L.peer == (u16)peer && c >= L.ctr && c < L.ctr + L.len);
        return __ffs(m) - 1;
    }
    // ---- split a full leaf: upper 16 slots move to a new leaf (their atom -> leaf entries follow)
    __device__ void leaf_split(u32 leaf) {
        if (n_leaves >= leaf_cap) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
        __syncwarp();
        u32 nl = n_leaves++;
        LeafImg L = leaf_load(leaf);
        i32 vis = L.st == 0 ? L.len : 0;
        if (lane >= 16) {
            u64 dst = (leaf0 + nl) * 32 + lane - 16;
            p.leaf_ps[dst] = (u32)L.peer | ((u32)L.st << 16);
            p.leaf_ctr[dst] = L.ctr;
            p.leaf_len[dst] = L.len;
            if (L.peer != PEER_UNKNOWN) p.atom_leaf[atom_index(L.peer, L.ctr)] = nl;
        }
        // long spans: their remaining atoms cooperatively
        unsigned longm = __ballot_sync(LB_FULL, lane >= 16 && L.len > 1 && L.peer != PEER_UNKNOWN);
        while (longm) {
            int s = __ffs(longm) - 1;
            longm &= longm - 1;
            u32 sp = __shfl_sync(LB_FULL, (u32)L.peer, s);
            i32 sc = __shfl_sync(LB_FULL, L.ctr, s);
            i32 sl = __shfl_sync(LB_FULL, L.len, s);
            u64 a0 = atom_index(sp, sc);
            for (i32 i = 1 + lane; i < sl; i += 32) p.atom_leaf[a0 + i] = nl;
        }
        unsigned unk = __ballot_sync(LB_FULL, lane >= 16 && L.peer == PEER_UNKNOWN);
        if (unk) unk_leaf = nl;
        i32 moved = warp_sum(lane >= 16 ? vis : 0);
        if (lane == 0) {
            p.leaf_n[leaf0 + leaf] = 16;
            p.leaf_n[leaf0 + nl] = 16;
            p.leaf_next[leaf0 + nl] = p.leaf_next[leaf0 + leaf];
            p.leaf_next[leaf0 + leaf] = nl;
        }
        __syncwarp();
        u32 plink = p.leaf_parent[leaf0 + leaf];
        u32 parent = plink >> 5;
        int idx = (int)(plink & 31);
        if (lane == 0) nd_add_vis(parent, idx, -moved);
        __syncwarp();
        node_insert(parent, idx, nl, moved, true);
    }
    // ---- open one empty slot at index `at` (0..n) of `leaf`, splitting the leaf first when it is full.
    // On return (leaf, at) name the opened slot, T is the shifted register image (slot `at` to be filled by
    // the caller) and the leaf must be written back with leaf_store(leaf, T, T.n + 1).
    __device__ void leaf_open(u32& leaf, int& at, LeafImg& T) {
        if (T.n >= 32) {
            leaf_split(leaf);
            if (err) return;
            if (at > 16) { leaf = p.leaf_next[leaf0 + leaf]; at -= 16; }
            T = leaf_load(leaf);
        }
        u16 pe = __shfl_up_sync(LB_FULL, T.peer, 1);
        i32 ct = __shfl_up_sync(LB_FULL, T.ctr, 1);
        i32 ln = __shfl_up_sync(LB_FULL, T.len, 1);
        u16 st = __shfl_up_sync(LB_FULL, T.st, 1);
        if (lane > at) { T.peer = pe; T.ctr = ct; T.len = ln; T.st = st; }
    }
    // ---- split the span containing atom (peer, c) right before that atom (FugueSpan::_slice,
    // fugue_span.rs:257-279); visible totals unchanged.  No-op when (peer, c) already starts a span.
    __device__ void split_before(u32 peer, i32 c) {
        u32 leaf = p.atom_leaf[atom_index(peer, c)];
        LeafImg L = leaf_load(leaf);
        int slot = slot_of(L, peer, c);
        if (slot < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
        i32 ctr = __shfl_sync(LB_FULL, L.ctr, slot);
        if (ctr == c) return;
        if (L.n >= 32) {   // make room first; the span (still whole) may move to the new leaf
            leaf_split(leaf);
            if (err) return;
            leaf = p.atom_leaf[atom_index(peer, c)];
            L = leaf_load(leaf);
            slot = slot_of(L, peer, c);
            if (slot < 0 || L.n >= 32) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
        }
        i32 len = __shfl_sync(LB_FULL, L.len, slot);
        u16 st = __shfl_sync(LB_FULL, L.st, slot);
        i32 k = c - ctr;
        if (lane == slot) L.len = k;               // left part keeps its slot
        u32 lf = leaf;
        int at = slot + 1;
        u16 pe = __shfl_up_sync(LB_FULL, L.peer, 1);
        i32 ct = __shfl_up_sync(LB_FULL, L.ctr, 1);
        i32 ln = __shfl_up_sync(LB_FULL, L.len, 1);
        u16 s2 = __shfl_up_sync(LB_FULL, L.st, 1);
        if (lane > at) { L.peer = pe; L.ctr = ct; L.len = ln; L.st = s2; }
        if (lane == at) { L.peer = (u16)peer; L.ctr = c; L.len = len - k; L.st = st; }
        leaf_store(lf, L, L.n + 1);
        // origins of the new span start; the right part's atoms may need a new home leaf
        u64 a_old = atom_index(peer, ctr), a_new = atom_index(peer, c);
        u16 orp = p.a_or_peer[a_old];
        i32 orc = p.a_or_ctr[a_old];
        if (lane == 0) {
            p.a_ol_peer[a_new] = (u16)peer;
            p.a_ol_ctr[a_new] = c - 1;
            p.a_or_peer[a_new] = orp;
            p.a_or_ctr[a_new] = orc;
        }
        if (lf != leaf)
            for (i32 i = lane; i < len - k; i += 32) p.atom_leaf[a_new + i] = lf;
        __syncwarp();
    }
    // ---- apply a status change to the inserted atoms [lo,hi) of `peer`
    __device__ void range_set(u32 peer, i32 lo, i32 hi, int set_future, int del_diff) {
        i32 c = lo;
        while (c < hi && !err) {
            u64 ai = atom_index(peer, c);
            u32 leaf = p.atom_leaf[ai];
            if (leaf == LEAF_NONE) { c++; continue; }
            LeafImg L = leaf_load(leaf);
            int slot = slot_of(L, peer, c);
            if (slot < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
            i32 s_ctr = __shfl_sync(LB_FULL, L.ctr, slot);
            i32 s_len = __shfl_sync(LB_FULL, L.len, slot);
            if (s_ctr != c || c + s_len > hi) {
                // boundaries do not line up with the span: cut it, then look again (rare)
                if (s_ctr != c) split_before(peer, c);
                if (!err && s_ctr + s_len > hi) split_before(peer, hi);
                if (err) return;
                leaf = p.atom_leaf[ai];
                L = leaf_load(leaf);
                slot = slot_of(L, peer, c);
                if (slot < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
                s_len = __shfl_sync(LB_FULL, L.len, slot);
            }
            // status change of the whole span + visible-length propagation
            u16 st = __shfl_sync(LB_FULL, L.st, slot);
            u16 nst = st;
            if (set_future == 1) nst |= ST_FUTURE;
            if (set_future == 0) nst &= (u16)~ST_FUTURE;
            nst = (u16)((nst & ST_FUTURE) | (((nst & 0x7FFF) + del_diff) & 0x7FFF));
            if (lane == 0) p.leaf_ps[(leaf0 + leaf) * 32 + slot] = (peer & 0xFFFFu) | ((u32)nst << 16);
            __syncwarp();
            i32 before = st == 0 ? s_len : 0, after = nst == 0 ? s_len : 0;
            add_vis(leaf, after - before);
            c += s_len;
        }
    }
    // ---- retreat (dir=-1) / forward (dir=+1) the ops of `peer` with counters [a,b) that touch this container
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
            __syncwarp();
            if (cur != tgt && lane == 0) p.cvv[cvv0 + q] = tgt;
        }
        __syncwarp();
    }
    // ---- position key of slot (leaf, slot) for cmp_pos (crdt_rope.rs:433-446)
    __device__ u64 order_key(u32 leaf, int slot) {
        u64 key = (u64)slot;
        int shift = 6;
        u32 link = p.leaf_parent[leaf0 + leaf];
        while (link != NODE_NONE) {
            key |= (u64)(link & 31) << shift;
            shift += 6;
            link = nd_parent(link >> 5);
        }
        return key;
    }
    __device__ u64 order_key_of_atom(u32 peer, i32 c) {
        if (peer == PEER_UNKNOWN) {
            LeafImg L = leaf_load(unk_leaf);
            unsigned m = __ballot_sync(LB_FULL, lane < (int)L.n && L.peer == PEER_UNKNOWN);
            return order_key(unk_leaf, __ffs(m) - 1);
        }
        u32 leaf = p.atom_leaf[atom_index(peer, c)];
        LeafImg L = leaf_load(leaf);
        return order_key(leaf, slot_of(L, peer, c));
    }
    // origin_left of atom (peer, c): stored for span starts, implied inside a span
    __device__ void atom_origin_left(u32 peer, i32 c, u16* op, i32* oc) {
        if (peer == PEER_UNKNOWN) { *op = PEER_NONE; *oc = -1; return; }
        u32 leaf = p.atom_leaf[atom_index(peer, c)];
        LeafImg L = leaf_load(leaf);
        int slot = slot_of(L, peer, c);
        i32 s_ctr = __shfl_sync(LB_FULL, L.ctr, slot);
        if (s_ctr == c) { u64 a = atom_index(peer, c); *op = p.a_ol_peer[a]; *oc = p.a_ol_ctr[a]; }
        else { *op = (u16)peer; *oc = c - 1; }
    }

    // ---- CrdtRope::insert (crdt_rope.rs:43-227)
    __device__ void insert(u32 peer, i32 ctr, i32 len, i32 pos) {
        // 1. cursor: right after the pos-th visible atom (prefer-left)
        u32 leaf = first_leaf;
        int slot = 0;
        i32 off = 0, rem = pos;
        if (pos > 0) {
            u32 nd = root;
            for (u32 lvl = height; lvl >= 1; lvl--) {
                u32 n = nd_n(nd);
                i32 v = lane < (int)n ? nd_vis(nd, lane) : 0;
                i32 incl = warp_incl_scan(v, lane);
                unsigned m = __ballot_sync(LB_FULL, lane < (int)n && incl >= rem);
                int idx = __ffs(m) - 1;
                if (idx < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
                rem -= __shfl_sync(LB_FULL, incl - v, idx);
                nd = nd_child(nd, idx);
            }
            leaf = nd;
        }
        LeafImg L = leaf_load(leaf);
        u16 ol_peer = PEER_NONE;
        i32 ol_ctr = -1;
        u32 cur_peer = PEER_NONE;
        i32 cur_ctr = 0, cur_len = 0;
        if (pos > 0) {
            i32 v = L.st == 0 ? L.len : 0;
            i32 incl = warp_incl_scan(v, lane);
            unsigned m = __ballot_sync(LB_FULL, lane < (int)L.n && incl >= rem);
            slot = __ffs(m) - 1;
            if (slot < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
            off = rem - __shfl_sync(LB_FULL, incl - v, slot);
            cur_peer = __shfl_sync(LB_FULL, (u32)L.peer, slot);
            cur_ctr = __shfl_sync(LB_FULL, L.ctr, slot);
            cur_len = __shfl_sync(LB_FULL, L.len, slot);
            if (cur_peer == PEER_UNKNOWN) { err = LB_ERR(DOC_ERR_CORRUPT); return; }  // beyond the content
            ol_peer = (u16)cur_peer;
            ol_ctr = cur_ctr + off - 1;
        }
        // 2. origin_right: first non-future span at/after the cursor; skipped spans are "in between"
        u16 or_peer = PEER_NONE;
        i32 or_ctr = -1;
        bool pr_valid = false;        // is there a right parent?
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
            LeafImg S = L;
            while (true) {
                bool cand = lane >= from && lane < (int)S.n;
                unsigned nonfut = __ballot_sync(LB_FULL, cand && !(S.st & ST_FUTURE));
                unsigned fut = __ballot_sync(LB_FULL, cand && (S.st & ST_FUTURE));
                if (nonfut) {
                    int f = __ffs(nonfut) - 1;
                    n_between += __popc(fut & ((1u << f) - 1));
                    or_peer = (u16)__shfl_sync(LB_FULL, (u32)S.peer, f);
                    or_ctr = __shfl_sync(LB_FULL, S.ctr, f);
                    pr_leaf = l2;
                    pr_slot = f;
                    pr_valid = true;   // provisional: confirmed below only when needed
                    break;
                }
                n_between += __popc(fut);
                l2 = p.leaf_next[leaf0 + l2];
                if (l2 == LEAF_NONE) break;
                from = 0;
                S = leaf_load(l2);
            }
        }
        // 3. Fugue sibling scan among the concurrent (future) spans (rare path; uniform serial code)
        bool after_valid = false;
        u32 after_peer = 0;
        i32 after_ctr = 0;   // insert right after the span starting at this atom
        if (n_between) {
            // right parent of the new span: origin_right counts only if its origin_left equals ours
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
                    u64 si = (leaf0 + l2) * 32 + s;
                    u32 o_peer = p.leaf_ps[si] & 0xFFFFu;
                    i32 o_ctr = p.leaf_ctr[si];
                    seen++;
                    u64 o_key = order_key(l2, s);
                    if (!have_first) { first_key = o_key; have_first = true; }
                    u64 oa = atom_index(o_peer, o_ctr);
                    u16 o_olp = p.a_ol_peer[oa];
                    i32 o_olc = p.a_ol_ctr[oa];
                    bool same_ol = o_olp == ol_peer && (ol_peer == PEER_NONE || o_olc == ol_ctr);
                    if (!same_ol) {
                        // "visited" is a prefix of the in-between spans: membership is a position test
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
        // 4. physical insertion
        u32 tgt_leaf;
        int at;
        LeafImg T;
        if (after_valid) {
            tgt_leaf = p.atom_leaf[atom_index(after_peer, after_ctr)];
            T = leaf_load(tgt_leaf);
            at = slot_of(T, after_peer, after_ctr) + 1;
        } else if (pos == 0) {
            tgt_leaf = first_leaf;
            T = L;
            at = 0;
        } else if (off < cur_len) {
            split_before(cur_peer, cur_ctr + off);
            if (err) return;
            tgt_leaf = p.atom_leaf[atom_index(cur_peer, cur_ctr + off)];   // right part: insert before it
            T = leaf_load(tgt_leaf);
            at = slot_of(T, cur_peer, cur_ctr + off);
        } else {
            tgt_leaf = leaf;
            T = L;
            at = slot + 1;
        }
        if (at < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
        leaf_open(tgt_leaf, at, T);
        if (err) return;
        if (lane == at) { T.peer = (u16)peer; T.ctr = ctr; T.len = len; T.st = 0; }
        leaf_store(tgt_leaf, T, T.n + 1);
        u64 a0 = atom_index(peer, ctr);
        if (lane == 0) {
            p.a_ol_peer[a0] = ol_peer;
            p.a_ol_ctr[a0] = ol_ctr;
            p.a_or_peer[a0] = or_peer;
            p.a_or_ctr[a0] = or_ctr;
        }
        for (i32 i = lane; i < len; i += 32) p.atom_leaf[a0 + i] = tgt_leaf;
        __syncwarp();
        add_vis(tgt_leaf, len);
    }

    // ---- delete by target id (crdt_rope.rs:236-315 ; tracker.rs:173-232)
    __device__ void del(u32 tpeer, i32 tctr, i32 n) { range_set(tpeer, tctr, tctr + n, -1, +1); }

    // ---- container switching: internal nodes < NS live in shared memory while a container is active
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
        if (n_leaves == 0) { init_container(); return; }
        u32 cached = n_nodes < LB_SEQ_NS ? n_nodes : LB_SEQ_NS;
        for (u32 nd = 0; nd < cached; nd++) {
            sm->child[nd][lane] = p.node_child[(node0 + nd) * 32 + lane];
            sm->vis[nd][lane] = p.node_vis[(node0 + nd) * 32 + lane];
            if (lane == 0) { sm->n[nd] = p.node_n[node0 + nd]; sm->parent[nd] = p.node_parent[node0 + nd]; }
        }
        __syncwarp();
    }
    __device__ void store_container() {
        if (cidx == 0xFFFFFFFFu) return;
        __syncwarp();
        u32 cached = n_nodes < LB_SEQ_NS ? n_nodes : LB_SEQ_NS;
        for (u32 nd = 0; nd < cached; nd++) {
            p.node_child[(node0 + nd) * 32 + lane] = sm->child[nd][lane];
            p.node_vis[(node0 + nd) * 32 + lane] = sm->vis[nd][lane];
            if (lane == 0) { p.node_n[node0 + nd] = sm->n[nd]; p.node_parent[node0 + nd] = sm->parent[nd]; }
        }
        if (lane == 0) {
            DocContainer& dc = t->dcont[di->cid0 + cidx];
            dc.n_leaves = n_leaves;
            dc.n_nodes = n_nodes;
            dc.root = root;
            dc.height = height;
            dc.first_leaf = first_leaf;
            dc.unk_sid = unk_leaf;
        }
        __syncwarp();
    }
    // Tracker::new_with_unknown (tracker.rs:38-63): one placeholder span of length u32::MAX/4
    __device__ void init_container() {
        if (leaf_cap < 1 || node_cap < 1) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
        n_leaves = 1;
        n_nodes = 1;
        root = 0;
        height = 1;
        first_leaf = 0;
        unk_leaf = 0;
        if (lane == 0) {
            p.leaf_ps[leaf0 * 32] = PEER_UNKNOWN;
            p.leaf_ctr[leaf0 * 32] = 0;
            p.leaf_len[leaf0 * 32] = UNKNOWN_LEN;
            p.leaf_n[leaf0] = 1;
            p.leaf_parent[leaf0] = 0;   // (node 0 << 5) | index 0
            p.leaf_next[leaf0] = LEAF_NONE;
            nd_set(0, 0, 0, UNKNOWN_LEN);
            nd_set_n(0, 1);
            nd_set_parent(0, NODE_NONE);
        }
        for (u32 q = lane; q < di->P; q += 32) p.cvv[cvv0 + q] = 0;
        __syncwarp();
    }
    // ---- emit the final visible runs of the current container (after checkout to the final version)
    __device__ void emit_output() {
        DocContainer& dc = t->dcont[di->cid0 + cidx];
        u32 n_out = 0;
        u32 total = 0;
        u32 l2 = first_leaf;
        while (l2 != LEAF_NONE) {
            LeafImg L = leaf_load(l2);
            bool live = lane < (int)L.n && L.st == 0 && L.peer != PEER_UNKNOWN;
            unsigned m = __ballot_sync(LB_FULL, live);
            if (live) {
                u32 o = n_out + __popc(m & ((1u << lane) - 1));
                if (o < dc.out_cap) {
                    u32 row = t->atom_row[atom_index(L.peer, L.ctr)];
                    p.out_row[dc.out0 + o] = row;
                    p.out_off[dc.out0 + o] = (u32)(L.ctr - t->op_counter[row]);
                    p.out_len[dc.out0 + o] = (u32)L.len;
                }
            }
            total += (u32)warp_sum(live ? L.len : 0);
            n_out += __popc(m);
            l2 = p.leaf_next[leaf0 + l2];
        }
        if (n_out > dc.out_cap) err = LB_ERR(DOC_ERR_CAPACITY);
        __syncwarp();
        if (lane == 0) {
            dc.n_out = n_out < dc.out_cap ? n_out : dc.out_cap;
            dc.seq_len = total;
        }
        __syncwarp();
    }
};

// one warp per document, LB_SEQ_WARPS documents per CTA
#ifndef LB_SEQ_MINB
#define LB_SEQ_MINB 8         // 8 CTAs x 4 warps = 32 resident documents per SM (64 registers/thread)
#endif
__global__ void __launch_bounds__(32 * LB_SEQ_WARPS, LB_SEQ_MINB)
k_seq_integrate(DocInfo* __restrict__ docs, u32 n_docs, const __grid_constant__ SeqPools pools,
                const __grid_constant__ SeqTables tables) {
    __shared__ SeqSmem smem[LB_SEQ_WARPS];
    u32 warp_in_cta = threadIdx.x >> 5;
    u32 warp_global = blockIdx.x * LB_SEQ_WARPS + warp_in_cta;
    int lane = threadIdx.x & 31;
    if (warp_global >= n_docs) return;
    DocInfo& di = docs[warp_global];
    if (di.code != DOC_OK || di.n_applied == 0) return;
    bool any = false;
    for (u32 c = 0; c < di.C; c++)
        if (tables.dcont[di.cid0 + c].leaf_cap) any = true;
    if (!any) return;
    Seq s(pools, &tables);
    s.di = &di;
    s.sm = &smem[warp_in_cta];
    s.lane = lane;
    s.err = 0;
    s.cidx = 0xFFFFFFFFu;
    s.atom0 = di.atom0;
    if (lane < (int)di.P) s.sm->abase[lane] = tables.dpeer[di.peer0 + lane].atom_base;
    __syncwarp();
    for (u64 i = lane; i < di.atom_total; i += 32) pools.atom_leaf[di.atom0 + i] = LEAF_NONE;
    for (u32 c = lane; c < di.C; c += 32) pools.cont_epoch[di.cid0 + c] = 0xFFFFFFFFu;
    __syncwarp();
    u32 P = di.P;
    u32 prev_peer = 0xFFFFFFFFu;
    u32 cur_epoch = 0xFFFFFFFFu;   // epoch of the active container (register; spilled on container switch)
    for (u32 k = 0; k < di.n_applied && !s.err; k++) {
        u32 ch = tables.ch_walk[di.ch0 + k];
        u32 peer = tables.ch_peer[ch];
        u64 r0 = tables.ch_op0[ch];
        u32 nr = tables.ch_nops[ch];
        // fast path (no checkout): the change only depends on its predecessor, which was the previous change
        // of the walk, and the container's tracker sat at that version when it was last touched
        bool chain = tables.ch_dep_self[ch] && tables.ch_ndeps[ch] == 0 && prev_peer == peer && k > 0;
        const i32* vv = nullptr;
        for (u32 r = 0; r < nr && !s.err; r++) {
            u64 row = r0 + r;
            u8 kind = tables.op_kind[row];
            if (kind != OPK_SEQ_INS && kind != OPK_SEQ_DEL) continue;
            u32 c = tables.op_cidx[row];
            i32 ctr = tables.op_counter[row];
            i32 len = (i32)tables.op_len[row];
            if (c != s.cidx) {
                if (s.cidx != 0xFFFFFFFFu && lane == 0) pools.cont_epoch[di.cid0 + s.cidx] = cur_epoch;
                s.store_container();
                s.load_container(c);
                if (s.err) break;
                cur_epoch = pools.cont_epoch[di.cid0 + c];
            }
            if (cur_epoch != k) {
                if (!(chain && cur_epoch == k - 1)) {
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
                cur_epoch = k;
            }
            if (kind == OPK_SEQ_INS) s.insert(peer, ctr, len, tables.op_prop[row]);
            else {
                u32 dl = tables.op_del[row];
                const BlockInfo& bi = tables.blocks[tables.ch_block[ch]];
                u32 tp = tables.peer_map[bi.peer0 + tables.del_peer_idx[dl]];
                s.del(tp, tables.del_counter[dl], len);
            }
            // current_vv of the tracker follows its own ops (tracker.rs:131-139, 228-231); in causal order this
            // entry only grows, and insert()/del() end with a warp barrier
            if (lane == 0) pools.cvv[s.cvv0 + peer] = ctr + len;
            __syncwarp();
        }
        prev_peer = peer;
    }
    // final version = everything applied
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
            __syncwarp();
            if (lane == 0) pools.cvv[s.cvv0 + q] = tgt;
            __syncwarp();
        }
        if (!s.err) s.emit_output();
    }
    s.store_container();
    if (lane == 0 && s.err) di.code = s.err;
}
