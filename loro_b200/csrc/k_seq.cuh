// loro_b200 -- phase 5: eg-walker (Fugue) integration of List/Text containers, one warp per document.
//
// Replaces (reference, relative to crates/loro-internal/src):
//   container/richtext/tracker.rs:84-232 (insert/delete), :330-526 (checkout: retreat / forward)
//   container/richtext/tracker/crdt_rope.rs:43-227 (Fugue origin reconstruction + sibling scan),
//     :236-315 (delete), :325-361 (status updates), :542-652 (active-length queries)
//   container/richtext/fugue_span.rs:192-386 (span status: future / delete_times ; slicing)
//   container/richtext/tracker/id_to_cursor.rs (id -> span lookup)  ->  dense atom_sid / atom_row arrays
//   diff_calc.rs:175-236,585-620 (per-change checkout then apply)
//
// Data structure (all in HBM, SoA): per container a B+tree whose leaves hold up to 32 spans
// (sid, len, state) and whose internal nodes hold up to 32 (child, visible-length) pairs, so that one
// warp reads a whole node with one coalesced access and resolves "k-th visible atom" with a shuffle
// scan + ballot per level.  Spans never merge or disappear, they only split; ids resolve to spans through
// a dense per-document atom -> span array.  Deletes are applied by target id (for well-formed histories
// the reference's by-position deletion hits the same atoms).
//
// Every function below is warp-synchronous: all 32 lanes call it with identical arguments.
#pragma once
#include "lb_defs.h"

#define SID_NONE 0xFFFFFFFFu
#define NODE_NONE 0xFFFFFFFFu
#define ST_FUTURE 0x80000000u

struct SeqPools {
    // leaves
    u32* leaf_sid; i32* leaf_len; u32* leaf_st; u32* leaf_n; u32* leaf_parent; u32* leaf_next;
    // internal nodes
    u32* node_child; i32* node_vis; u32* node_n; u32* node_parent;
    // spans (doc-level pools)
    u16* sp_peer; i32* sp_ctr; i32* sp_len; u32* sp_leaf;
    u16* sp_ol_peer; i32* sp_ol_ctr; u16* sp_or_peer; i32* sp_or_ctr;
    u32* atom_sid;
    i32* cvv;
    u32* cont_epoch;   // per container: last walk index that did a checkout
    u32* out_row; u32* out_off; u32* out_len;
};

struct SeqTables {
    const DocPeer* dpeer; DocContainer* dcont;
    const u32* ch_walk; const u64* ch_op0; const u32* ch_nops; const u16* ch_peer; const i32* ch_vv;
    const u32* ch_order; const i32* ch_counter;
    const u8* op_kind; const u32* op_cidx; const i32* op_prop; const u32* op_len; const i32* op_counter;
    const u32* op_del; const u32* op_change;
    const u32* del_peer_idx; const i32* del_counter; const i32* del_len;
    const u32* peer_map; const BlockInfo* blocks; const u32* ch_block;
    const u32* atom_row;
};

// per-warp working state (uniform across lanes)
struct Seq {
    SeqPools p;
    const SeqTables* t;
    const DocInfo* di;
    int lane;
    // doc level
    u32 n_spans;
    u32 err;
    // current container
    u32 cidx;
    u64 leaf0, node0;
    u32 leaf_cap, node_cap, n_leaves, n_nodes, root, height, first_leaf;
    u64 cvv0;

    __device__ __forceinline__ u64 atom_index(u32 peer, i32 ctr) const {
        return di->atom0 + t->dpeer[di->peer0 + peer].atom_base + (u32)ctr;
    }
    __device__ __forceinline__ u32 alloc_span() {
        if (n_spans >= di->span_cap) { err = LB_ERR(DOC_ERR_CAPACITY); return SID_NONE; }
        return n_spans++;
    }
    // ---- add `delta` visible atoms on the path leaf -> root
    __device__ void add_vis(u32 leaf, i32 delta) {
        if (delta == 0) return;
        u32 child = leaf;
        u32 node = p.leaf_parent[leaf0 + leaf];
        while (node != NODE_NONE) {
            u32 n = p.node_n[node0 + node];
            u32 c = lane < (int)n ? p.node_child[(node0 + node) * 32 + lane] : NODE_NONE;
            unsigned m = __ballot_sync(LB_FULL, c == child);
            int idx = __ffs(m) - 1;
            if (idx < 0) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
            if (lane == idx) p.node_vis[(node0 + node) * 32 + lane] += delta;
            child = node;
            node = p.node_parent[node0 + node];
        }
        __syncwarp();
    }
    // ---- insert (child, vis) into `node` right after position `after` (-1 = front); splits upward.
    // `is_leaf_level`: the children of `node` are leaves.  Returns the node that finally holds `child`.
    __device__ u32 node_insert(u32 node, int after, u32 child, i32 vis, bool children_are_leaves) {
        u32 result = NODE_NONE;
        while (true) {
            u32 n = p.node_n[node0 + node];
            if (n == 32) {
                // split: upper 16 children move to a new node
                if (n_nodes >= node_cap) { err = LB_ERR(DOC_ERR_CAPACITY); return result; }
                __syncwarp();
                u32 nn = n_nodes++;
                u32 c = p.node_child[(node0 + node) * 32 + lane];
                i32 v = p.node_vis[(node0 + node) * 32 + lane];
                if (lane >= 16) {
                    p.node_child[(node0 + nn) * 32 + lane - 16] = c;
                    p.node_vis[(node0 + nn) * 32 + lane - 16] = v;
                    if (children_are_leaves) p.leaf_parent[leaf0 + c] = nn;
                    else p.node_parent[node0 + c] = nn;
                }
                i32 moved = warp_sum(lane >= 16 ? v : 0);
                if (lane == 0) {
                    p.node_n[node0 + node] = 16;
                    p.node_n[node0 + nn] = 16;
                }
                __syncwarp();
                // place the pending child in the proper half
                u32 target;
                int t_after;
                if (after >= 16) { target = nn; t_after = after - 16; }
                else { target = node; t_after = after; }
                insert_no_split(target, t_after, child, vis, children_are_leaves);
                if (result == NODE_NONE) result = target;
                // now insert `nn` into the parent of `node`
                u32 parent = p.node_parent[node0 + node];
                if (parent == NODE_NONE) {
                    if (n_nodes >= node_cap) { err = LB_ERR(DOC_ERR_CAPACITY); return result; }
                    u32 nr = n_nodes++;
                    i32 tot_old = warp_sum(lane < 32 ? (lane < (int)p.node_n[node0 + node] ? p.node_vis[(node0 + node) * 32 + lane] : 0) : 0);
                    i32 tot_new = warp_sum(lane < (int)p.node_n[node0 + nn] ? p.node_vis[(node0 + nn) * 32 + lane] : 0);
                    if (lane == 0) {
                        p.node_child[(node0 + nr) * 32 + 0] = node;
                        p.node_vis[(node0 + nr) * 32 + 0] = tot_old;
                        p.node_child[(node0 + nr) * 32 + 1] = nn;
                        p.node_vis[(node0 + nr) * 32 + 1] = tot_new;
                        p.node_n[node0 + nr] = 2;
                        p.node_parent[node0 + nr] = NODE_NONE;
                        p.node_parent[node0 + node] = nr;
                        p.node_parent[node0 + nn] = nr;
                    }
                    __syncwarp();
                    root = nr;
                    height++;
                    return result;
                }
                // fix the parent's entry of `node` (its vis shrank by what moved, adjusted for the new child)
                u32 pn = p.node_n[node0 + parent];
                u32 pc = lane < (int)pn ? p.node_child[(node0 + parent) * 32 + lane] : NODE_NONE;
                unsigned m = __ballot_sync(LB_FULL, pc == node);
                int idx = __ffs(m) - 1;
                i32 tot_new = warp_sum(lane < (int)p.node_n[node0 + nn] ? p.node_vis[(node0 + nn) * 32 + lane] : 0);
                i32 tot_old = warp_sum(lane < (int)p.node_n[node0 + node] ? p.node_vis[(node0 + node) * 32 + lane] : 0);
                if (lane == idx) p.node_vis[(node0 + parent) * 32 + lane] = tot_old;
                __syncwarp();
                (void)moved;
                // continue one level up: pending child = nn with vis tot_new, after idx.
                // NOTE: the caller adds the vis of the *original* pending child along the path afterwards
                // via add_vis, so the parent totals written here exclude it consistently: tot_old/tot_new
                // are recomputed from the children (which already include the pending child).  To keep the
                // invariant "parent entry == sum of children" we therefore must not add it again: callers
                // use vis=0 for structural inserts and account visibility separately.
                child = nn;
                vis = tot_new;
                after = idx;
                node = parent;
                children_are_leaves = false;
                continue;
            }
            insert_no_split(node, after, child, vis, children_are_leaves);
            if (result == NODE_NONE) result = node;
            return result;
        }
    }
    __device__ void insert_no_split(u32 node, int after, u32 child, i32 vis, bool children_are_leaves) {
        u32 n = p.node_n[node0 + node];
        u32 c = lane < (int)n ? p.node_child[(node0 + node) * 32 + lane] : 0;
        i32 v = lane < (int)n ? p.node_vis[(node0 + node) * 32 + lane] : 0;
        u32 c_up = __shfl_up_sync(LB_FULL, c, 1);
        i32 v_up = __shfl_up_sync(LB_FULL, v, 1);
        int at = after + 1;
        if (lane == at) { c = child; v = vis; }
        else if (lane > at) { c = c_up; v = v_up; }
        if (lane <= (int)n) {
            p.node_child[(node0 + node) * 32 + lane] = c;
            p.node_vis[(node0 + node) * 32 + lane] = v;
        }
        if (lane == 0) {
            p.node_n[node0 + node] = n + 1;
            if (children_are_leaves) p.leaf_parent[leaf0 + child] = node;
            else p.node_parent[node0 + child] = node;
        }
        __syncwarp();
    }
    // ---- make room in `leaf` (split when full).  Afterwards every span keeps a valid sp_leaf.
    __device__ void leaf_make_room(u32 leaf) {
        if (p.leaf_n[leaf0 + leaf] < 32) return;
        if (n_leaves >= leaf_cap) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
        __syncwarp();  // callers' reads (sp_leaf, slots) are complete before spans start moving
        u32 nl = n_leaves++;
        u64 src = (leaf0 + leaf) * 32 + lane;
        u32 sid = p.leaf_sid[src];
        i32 len = p.leaf_len[src];
        u32 st = p.leaf_st[src];
        i32 vis = st == 0 ? len : 0;
        if (lane >= 16) {
            u64 dst = (leaf0 + nl) * 32 + lane - 16;
            p.leaf_sid[dst] = sid;
            p.leaf_len[dst] = len;
            p.leaf_st[dst] = st;
            p.sp_leaf[di->span0 + sid] = nl;
        }
        i32 moved = warp_sum(lane >= 16 ? vis : 0);
        if (lane == 0) {
            p.leaf_n[leaf0 + leaf] = 16;
            p.leaf_n[leaf0 + nl] = 16;
            p.leaf_next[leaf0 + nl] = p.leaf_next[leaf0 + leaf];
            p.leaf_next[leaf0 + leaf] = nl;
        }
        __syncwarp();
        // parent bookkeeping: old leaf's entry loses `moved`, the new leaf enters right after it
        u32 parent = p.leaf_parent[leaf0 + leaf];
        u32 pn = p.node_n[node0 + parent];
        u32 pc = lane < (int)pn ? p.node_child[(node0 + parent) * 32 + lane] : NODE_NONE;
        unsigned m = __ballot_sync(LB_FULL, pc == leaf);
        int idx = __ffs(m) - 1;
        if (lane == idx) p.node_vis[(node0 + parent) * 32 + lane] -= moved;
        __syncwarp();
        node_insert(parent, idx, nl, moved, true);
    }
    // ---- insert a slot at index `at` of `leaf` (room must exist); no path update
    __device__ void leaf_insert_slot(u32 leaf, int at, u32 sid, i32 len, u32 st) {
        u32 n = p.leaf_n[leaf0 + leaf];
        u64 base = (leaf0 + leaf) * 32;
        u32 s = lane < (int)n ? p.leaf_sid[base + lane] : 0;
        i32 l = lane < (int)n ? p.leaf_len[base + lane] : 0;
        u32 x = lane < (int)n ? p.leaf_st[base + lane] : 0;
        u32 s_up = __shfl_up_sync(LB_FULL, s, 1);
        i32 l_up = __shfl_up_sync(LB_FULL, l, 1);
        u32 x_up = __shfl_up_sync(LB_FULL, x, 1);
        if (lane == at) { s = sid; l = len; x = st; }
        else if (lane > at) { s = s_up; l = l_up; x = x_up; }
        if (lane <= (int)n) {
            p.leaf_sid[base + lane] = s;
            p.leaf_len[base + lane] = l;
            p.leaf_st[base + lane] = x;
        }
        if (lane == 0) {
            p.leaf_n[leaf0 + leaf] = n + 1;
            p.sp_leaf[di->span0 + sid] = leaf;
        }
        __syncwarp();
    }
    // ---- locate the slot of span `sid` inside its leaf
    __device__ __forceinline__ int slot_of(u32 leaf, u32 sid) {
        u32 n = p.leaf_n[leaf0 + leaf];
        u32 s = lane < (int)n ? p.leaf_sid[(leaf0 + leaf) * 32 + lane] : SID_NONE;
        unsigned m = __ballot_sync(LB_FULL, s == sid);
        return __ffs(m) - 1;
    }
    // ---- split span `sid` at offset k (0<k<len): the right part becomes a new span placed right after it
    // (FugueSpan::_slice, fugue_span.rs:257-279).  Visible totals are unchanged.  Returns the new sid.
    __device__ u32 span_split(u32 sid, i32 k) {
        u64 g = di->span0 + sid;
        u32 leaf = p.sp_leaf[g];
        leaf_make_room(leaf);
        if (err) return SID_NONE;
        leaf = p.sp_leaf[g];
        int slot = slot_of(leaf, sid);
        u32 nsid = alloc_span();
        if (nsid == SID_NONE || slot < 0) { err = err ? err : LB_ERR(DOC_ERR_CAPACITY); return SID_NONE; }
        u64 ng = di->span0 + nsid;
        u16 peer = p.sp_peer[g];
        i32 ctr = p.sp_ctr[g];
        i32 len = p.sp_len[g];
        u32 st = p.leaf_st[(leaf0 + leaf) * 32 + slot];
        u16 orp = p.sp_or_peer[g];
        i32 orc = p.sp_or_ctr[g];
        __syncwarp();  // all lanes hold the old span fields before lane 0 rewrites them
        if (lane == 0) {
            p.sp_peer[ng] = peer;
            p.sp_ctr[ng] = ctr + k;
            p.sp_len[ng] = len - k;
            p.sp_ol_peer[ng] = peer;
            p.sp_ol_ctr[ng] = ctr + k - 1;
            p.sp_or_peer[ng] = orp;
            p.sp_or_ctr[ng] = orc;
            p.sp_len[g] = k;
            p.leaf_len[(leaf0 + leaf) * 32 + slot] = k;
        }
        __syncwarp();
        leaf_insert_slot(leaf, slot + 1, nsid, len - k, st);
        if (peer != PEER_UNKNOWN) {
            u64 a0 = atom_index(peer, ctr + k);
            for (i32 i = lane; i < len - k; i += 32) p.atom_sid[a0 + i] = nsid;
        }
        __syncwarp();
        return nsid;
    }
    // ---- change the status of span `sid` (whole span) and propagate the visible-length delta
    __device__ void span_set(u32 sid, int set_future, int del_diff) {
        u64 g = di->span0 + sid;
        u32 leaf = p.sp_leaf[g];
        int slot = slot_of(leaf, sid);
        if (slot < 0) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
        u64 si = (leaf0 + leaf) * 32 + slot;
        u32 st = p.leaf_st[si];
        i32 len = p.leaf_len[si];
        u32 nst = st;
        if (set_future == 1) nst |= ST_FUTURE;
        if (set_future == 0) nst &= ~ST_FUTURE;
        nst = (nst & ST_FUTURE) | (u32)(((i32)(nst & 0xFFFF) + del_diff) & 0xFFFF);
        __syncwarp();  // every lane has read the old state before lane 0 overwrites it
        if (lane == 0) p.leaf_st[si] = nst;
        __syncwarp();
        i32 before = st == 0 ? len : 0, after = nst == 0 ? len : 0;
        add_vis(leaf, after - before);
    }
    // ---- apply f to the insert-spans covering ids [lo,hi) of `peer`
    __device__ void range_set(u32 peer, i32 lo, i32 hi, int set_future, int del_diff) {
        i32 c = lo;
        while (c < hi && !err) {
            u32 sid = p.atom_sid[atom_index(peer, c)];
            if (sid == SID_NONE) { c++; continue; }
            u64 g = di->span0 + sid;
            i32 s_ctr = p.sp_ctr[g], s_len = p.sp_len[g];
            if (c < s_ctr || c >= s_ctr + s_len) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
            if (s_ctr < c) {
                sid = span_split(sid, c - s_ctr);
                if (err) return;
                g = di->span0 + sid;
                s_ctr = c;
                s_len = p.sp_len[g];
            }
            if (s_ctr + s_len > hi) {
                span_split(sid, hi - s_ctr);
                if (err) return;
                s_len = hi - s_ctr;
            }
            span_set(sid, set_future, del_diff);
            c = s_ctr + s_len;
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
            __syncwarp();  // all lanes have read cvv[q]
            if (cur != tgt && lane == 0) p.cvv[cvv0 + q] = tgt;
        }
        __syncwarp();
    }
    // ---- position key of a span for cmp_pos (crdt_rope.rs:433-446): child indices along the root path
    __device__ u64 order_key(u32 sid) {
        u64 g = di->span0 + sid;
        u32 leaf = p.sp_leaf[g];
        u64 key = (u64)slot_of(leaf, sid);
        int shift = 6;
        u32 child = leaf;
        u32 node = p.leaf_parent[leaf0 + leaf];
        while (node != NODE_NONE) {
            u32 n = p.node_n[node0 + node];
            u32 c = lane < (int)n ? p.node_child[(node0 + node) * 32 + lane] : NODE_NONE;
            unsigned m = __ballot_sync(LB_FULL, c == child);
            key |= (u64)(__ffs(m) - 1) << shift;
            shift += 6;
            child = node;
            node = p.node_parent[node0 + node];
        }
        return key;
    }
    // origin_left of the atom (peer, ctr) living in span `sid`
    __device__ __forceinline__ void atom_origin_left(u32 sid, i32 ctr, u16* op, i32* oc) {
        u64 g = di->span0 + sid;
        if (p.sp_ctr[g] == ctr) { *op = p.sp_ol_peer[g]; *oc = p.sp_ol_ctr[g]; }
        else { *op = p.sp_peer[g]; *oc = ctr - 1; }
    }

    // ---- CrdtRope::insert (crdt_rope.rs:43-227)
    __device__ void insert(u32 peer, i32 ctr, i32 len, i32 pos) {
        // 1. cursor: right after the pos-th visible atom, preferring the left-most position
        u32 leaf = first_leaf;
        int slot = 0;
        i32 off = 0;
        if (pos > 0) {
            u32 node = root;
            i32 rem = pos;
            for (u32 lvl = height; lvl >= 1; lvl--) {
                u32 n = p.node_n[node0 + node];
                i32 v = lane < (int)n ? p.node_vis[(node0 + node) * 32 + lane] : 0;
                i32 incl = warp_incl_scan(v, lane);
                unsigned m = __ballot_sync(LB_FULL, lane < (int)n && incl >= rem);
                int idx = __ffs(m) - 1;
                if (idx < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
                i32 before = __shfl_sync(LB_FULL, incl - v, idx);
                u32 child = p.node_child[(node0 + node) * 32 + idx];
                rem -= before;
                node = child;
            }
            leaf = node;
            u32 n = p.leaf_n[leaf0 + leaf];
            u64 base = (leaf0 + leaf) * 32;
            i32 l = lane < (int)n ? p.leaf_len[base + lane] : 0;
            u32 st = lane < (int)n ? p.leaf_st[base + lane] : 1;
            i32 v = st == 0 ? l : 0;
            i32 incl = warp_incl_scan(v, lane);
            unsigned m = __ballot_sync(LB_FULL, lane < (int)n && incl >= rem);
            slot = __ffs(m) - 1;
            if (slot < 0) { err = LB_ERR(DOC_ERR_CORRUPT); return; }
            off = rem - __shfl_sync(LB_FULL, incl - v, slot);
        }
        // 2. origin_left
        u16 ol_peer = PEER_NONE;
        i32 ol_ctr = -1;
        u32 cur_sid = p.leaf_sid[(leaf0 + leaf) * 32 + slot];
        i32 cur_len = p.leaf_len[(leaf0 + leaf) * 32 + slot];
        if (pos > 0) {
            if (cur_sid == unk_sid) { err = LB_ERR(DOC_ERR_CORRUPT); return; }  // position beyond the content
            u64 g = di->span0 + cur_sid;
            ol_peer = p.sp_peer[g];
            ol_ctr = p.sp_ctr[g] + off - 1;
        }
        // 3. origin_right: first non-future span at/after the cursor; spans skipped are "in between"
        u16 or_peer = PEER_NONE;
        i32 or_ctr = -1;
        u32 parent_right = SID_NONE;
        u32 n_between = 0;
        u32 scan_leaf = leaf;
        int scan_from = slot;
        if (pos > 0 && off >= cur_len) scan_from = slot + 1;   // cursor sits at the end of cur span
        if (pos > 0 && off < cur_len) {
            // inside an active span: it is its own right neighbour
            u64 g = di->span0 + cur_sid;
            or_peer = p.sp_peer[g];
            or_ctr = p.sp_ctr[g] + off;
            parent_right = cur_sid;
        } else {
            u32 l2 = scan_leaf;
            int from = scan_from;
            while (l2 != NODE_NONE) {
                u32 n = p.leaf_n[leaf0 + l2];
                u32 st = lane < (int)n ? p.leaf_st[(leaf0 + l2) * 32 + lane] : 0;
                bool cand = lane >= from && lane < (int)n;
                unsigned nonfut = __ballot_sync(LB_FULL, cand && !(st & ST_FUTURE));
                unsigned fut = __ballot_sync(LB_FULL, cand && (st & ST_FUTURE));
                if (nonfut) {
                    int f = __ffs(nonfut) - 1;
                    n_between += __popc(fut & ((1u << f) - 1));
                    u32 sid = p.leaf_sid[(leaf0 + l2) * 32 + f];
                    u64 g = di->span0 + sid;
                    or_peer = p.sp_peer[g];
                    or_ctr = p.sp_ctr[g];
                    if (p.sp_ol_peer[g] == ol_peer && (ol_peer == PEER_NONE || p.sp_ol_ctr[g] == ol_ctr))
                        parent_right = sid;
                    break;
                }
                n_between += __popc(fut);
                l2 = p.leaf_next[leaf0 + l2];
                from = 0;
            }
        }
        // 4. Fugue sibling scan among the concurrent (future) spans (rare path; uniform serial code)
        u32 after_sid = SID_NONE;
        if (n_between) {
            bool scanning = false;
            u64 my_peer_id = t->dpeer[di->peer0 + peer].id;
            u64 pr_key = parent_right != SID_NONE ? order_key(parent_right) : 0;
            // "visited" (crdt_rope.rs:141-160) is always a prefix of the in-between spans, so membership of
            // an origin_left is a position test: first in-between span <= span(origin_left) < current span
            u64 first_key = 0;
            bool have_first = false;
            u32 l2 = scan_leaf;
            int from = scan_from;
            u32 seen = 0;
            bool stop = false;
            while (l2 != NODE_NONE && seen < n_between && !stop) {
                u32 n = p.leaf_n[leaf0 + l2];
                for (int s = from; s < (int)n && seen < n_between && !stop; s++) {
                    u32 osid = p.leaf_sid[(leaf0 + l2) * 32 + s];
                    u64 og = di->span0 + osid;
                    seen++;
                    u64 o_key = order_key(osid);
                    if (!have_first) { first_key = o_key; have_first = true; }
                    u16 o_olp = p.sp_ol_peer[og];
                    i32 o_olc = p.sp_ol_ctr[og];
                    bool same_ol = o_olp == ol_peer && (ol_peer == PEER_NONE || o_olc == ol_ctr);
                    if (!same_ol) {
                        bool in_visited = false;
                        if (o_olp != PEER_NONE && o_olp != PEER_UNKNOWN) {
                            u32 lsid = p.atom_sid[atom_index(o_olp, o_olc)];
                            if (lsid != SID_NONE) {
                                u64 lk = order_key(lsid);
                                in_visited = lk >= first_key && lk < o_key;
                            }
                        }
                        if (!in_visited) { stop = true; break; }
                    }
                    if (same_ol) {
                        u16 o_orp = p.sp_or_peer[og];
                        i32 o_orc = p.sp_or_ctr[og];
                        bool same_or = o_orp == or_peer && (or_peer == PEER_NONE || o_orc == or_ctr);
                        u64 o_peer_id = t->dpeer[di->peer0 + p.sp_peer[og]].id;
                        if (same_or) {
                            if (o_peer_id > my_peer_id) { stop = true; break; }
                            scanning = false;
                        } else {
                            u32 other_pr = SID_NONE;
                            if (o_orp != PEER_NONE) {
                                u32 esid = (o_orp == PEER_UNKNOWN) ? unk_sid : p.atom_sid[atom_index(o_orp, o_orc)];
                                u16 e_olp;
                                i32 e_olc;
                                atom_origin_left(esid, o_orc, &e_olp, &e_olc);
                                if (e_olp == ol_peer && (ol_peer == PEER_NONE || e_olc == ol_ctr)) other_pr = esid;
                            }
                            int cmp;
                            if (other_pr != SID_NONE && parent_right != SID_NONE) {
                                u64 a = order_key(other_pr);
                                cmp = a < pr_key ? -1 : (a > pr_key ? 1 : 0);
                            } else if (other_pr != SID_NONE) cmp = -1;
                            else if (parent_right != SID_NONE) cmp = 1;
                            else cmp = 0;
                            if (cmp < 0) scanning = true;
                            else if (cmp == 0 && o_peer_id > my_peer_id) { stop = true; break; }
                            else scanning = false;
                        }
                    }
                    if (!scanning) after_sid = osid;
                }
                l2 = p.leaf_next[leaf0 + l2];
                from = 0;
            }
        }
        // 5. physical insertion
        u32 nsid = alloc_span();
        if (nsid == SID_NONE) return;
        u64 ng = di->span0 + nsid;
        if (lane == 0) {
            p.sp_peer[ng] = (u16)peer;
            p.sp_ctr[ng] = ctr;
            p.sp_len[ng] = len;
            p.sp_ol_peer[ng] = ol_peer;
            p.sp_ol_ctr[ng] = ol_ctr;
            p.sp_or_peer[ng] = or_peer;
            p.sp_or_ctr[ng] = or_ctr;
        }
        u32 tgt_leaf;
        int at;
        if (after_sid != SID_NONE) {
            tgt_leaf = p.sp_leaf[di->span0 + after_sid];
            leaf_make_room(tgt_leaf);
            if (err) return;
            tgt_leaf = p.sp_leaf[di->span0 + after_sid];
            at = slot_of(tgt_leaf, after_sid) + 1;
        } else if (pos == 0) {
            leaf_make_room(first_leaf);
            if (err) return;
            tgt_leaf = first_leaf;
            at = 0;
        } else {
            if (off < cur_len) {
                span_split(cur_sid, off);
                if (err) return;
            }
            tgt_leaf = p.sp_leaf[di->span0 + cur_sid];
            leaf_make_room(tgt_leaf);
            if (err) return;
            tgt_leaf = p.sp_leaf[di->span0 + cur_sid];
            at = slot_of(tgt_leaf, cur_sid) + 1;
        }
        leaf_insert_slot(tgt_leaf, at, nsid, len, 0);
        add_vis(tgt_leaf, len);
        u64 a0 = atom_index(peer, ctr);
        for (i32 i = lane; i < len; i += 32) p.atom_sid[a0 + i] = nsid;
        if (lane == 0) {
            i32 cur = p.cvv[cvv0 + peer];
            if (ctr + len > cur) p.cvv[cvv0 + peer] = ctr + len;
        }
        __syncwarp();
    }
    u32 unk_sid;   // sid of the placeholder span of the current container

    // ---- delete by target id (crdt_rope.rs:236-315 ; tracker.rs:173-232)
    __device__ void del(u32 op_peer, i32 op_ctr, i32 n, u32 tpeer, i32 tctr) {
        range_set(tpeer, tctr, tctr + n, -1, +1);
        if (lane == 0) {
            i32 cur = p.cvv[cvv0 + op_peer];
            if (op_ctr + n > cur) p.cvv[cvv0 + op_peer] = op_ctr + n;
        }
        __syncwarp();
    }

    // ---- container switching
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
        unk_sid = dc.unk_sid;
        if (n_leaves == 0) init_container();
    }
    __device__ void store_container() {
        if (cidx == 0xFFFFFFFFu) return;
        __syncwarp();
        if (lane == 0) {
            DocContainer& dc = t->dcont[di->cid0 + cidx];
            dc.n_leaves = n_leaves;
            dc.n_nodes = n_nodes;
            dc.root = root;
            dc.height = height;
            dc.first_leaf = first_leaf;
            dc.unk_sid = unk_sid;
        }
        __syncwarp();
    }
    // Tracker::new_with_unknown (tracker.rs:38-63): one placeholder span of length u32::MAX/4
    __device__ void init_container() {
        if (leaf_cap < 1 || node_cap < 1) { err = LB_ERR(DOC_ERR_CAPACITY); return; }
        u32 sid = alloc_span();
        if (sid == SID_NONE) return;
        u64 g = di->span0 + sid;
        n_leaves = 1;
        n_nodes = 1;
        root = 0;
        height = 1;
        first_leaf = 0;
        unk_sid = sid;
        if (lane == 0) {
            p.sp_peer[g] = PEER_UNKNOWN;
            p.sp_ctr[g] = 0;
            p.sp_len[g] = UNKNOWN_LEN;
            p.sp_leaf[g] = 0;
            p.sp_ol_peer[g] = PEER_NONE;
            p.sp_ol_ctr[g] = -1;
            p.sp_or_peer[g] = PEER_NONE;
            p.sp_or_ctr[g] = -1;
            p.leaf_sid[leaf0 * 32] = sid;
            p.leaf_len[leaf0 * 32] = UNKNOWN_LEN;
            p.leaf_st[leaf0 * 32] = 0;
            p.leaf_n[leaf0] = 1;
            p.leaf_parent[leaf0] = 0;
            p.leaf_next[leaf0] = NODE_NONE;
            p.node_child[node0 * 32] = 0;
            p.node_vis[node0 * 32] = UNKNOWN_LEN;
            p.node_n[node0] = 1;
            p.node_parent[node0] = NODE_NONE;
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
        while (l2 != NODE_NONE) {
            u32 n = p.leaf_n[leaf0 + l2];
            u64 base = (leaf0 + l2) * 32;
            u32 sid = lane < (int)n ? p.leaf_sid[base + lane] : SID_NONE;
            u32 st = lane < (int)n ? p.leaf_st[base + lane] : 1;
            bool live = lane < (int)n && st == 0 && sid != unk_sid;
            unsigned m = __ballot_sync(LB_FULL, live);
            if (live) {
                u64 g = di->span0 + sid;
                u32 o = n_out + __popc(m & ((1u << lane) - 1));
                if (o < dc.out_cap) {
                    u32 row = t->atom_row[atom_index(p.sp_peer[g], p.sp_ctr[g])];
                    p.out_row[dc.out0 + o] = row;
                    p.out_off[dc.out0 + o] = (u32)(p.sp_ctr[g] - t->op_counter[row]);
                    p.out_len[dc.out0 + o] = (u32)p.sp_len[g];
                }
            }
            i32 l = live ? p.sp_len[di->span0 + sid] : 0;
#ifdef LB_SIMT_EMU
            if (getenv("LB_EMU_TRACE") && lane < (int)n) fprintf(stderr, "   emit c=%u leaf=%u lane=%d sid=%u st=%x live=%d l=%d unk=%u\n", cidx, l2, lane, sid, st, (int)live, l, unk_sid);
#endif
            total += (u32)warp_sum(l);
            n_out += __popc(m);
            l2 = p.leaf_next[leaf0 + l2];
        }
        if (n_out > dc.out_cap) err = LB_ERR(DOC_ERR_CAPACITY);
        __syncwarp();
        if (lane == 0) {
            dc.n_out = n_out < dc.out_cap ? n_out : dc.out_cap;
            dc.seq_len = total;
#ifdef LB_SIMT_EMU
            if (getenv("LB_EMU_TRACE")) fprintf(stderr, "   emit store c=%u n_out=%u total=%u\n", cidx, n_out, total);
#endif
        }
        __syncwarp();
    }
};

// one warp per document
__global__ void k_seq_integrate(DocInfo* __restrict__ docs, u32 n_docs, SeqPools pools, SeqTables tables) {
    u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp_global >= n_docs) return;
    DocInfo& di = docs[warp_global];
    if (di.code != DOC_OK || di.n_applied == 0) return;
    // does this document have any sequence container with work?
    bool any = false;
    for (u32 c = 0; c < di.C; c++) {
        const DocContainer& dc = tables.dcont[di.cid0 + c];
        if (dc.leaf_cap) any = true;
    }
    if (!any) return;
    Seq s;
    s.p = pools;
    s.t = &tables;
    s.di = &di;
    s.lane = lane;
    s.n_spans = 0;
    s.err = 0;
    s.cidx = 0xFFFFFFFFu;
    s.unk_sid = 0;
    // atom_sid of this doc starts as NONE
    for (u64 i = lane; i < di.atom_total; i += 32) pools.atom_sid[di.atom0 + i] = SID_NONE;
    for (u32 c = 0; c < di.C; c++)
        if (lane == 0) pools.cont_epoch[di.cid0 + c] = 0xFFFFFFFFu;
    __syncwarp();
    u32 P = di.P;
    for (u32 k = 0; k < di.n_applied && !s.err; k++) {
        u32 ch = tables.ch_walk[di.ch0 + k];
        u32 peer = tables.ch_peer[ch];
        // row of this change inside the doc's ch_vv: position in the per-peer ordered list
        const DocPeer& dp = tables.dpeer[di.peer0 + peer];
        i32 cc = tables.ch_counter[ch];
        u32 lo = 0, hi = dp.ch_count;
        while (hi - lo > 1) {
            u32 mid = (lo + hi) >> 1;
            if (tables.ch_counter[tables.ch_order[di.ch0 + dp.ch_first + mid]] <= cc) lo = mid; else hi = mid;
        }
        const i32* vv = tables.ch_vv + di.vv0 + (u64)(dp.ch_first + lo) * P;
        u64 r0 = tables.ch_op0[ch];
        u32 nr = tables.ch_nops[ch];
        for (u32 r = 0; r < nr && !s.err; r++) {
            u64 row = r0 + r;
            u8 kind = tables.op_kind[row];
            if (kind != OPK_SEQ_INS && kind != OPK_SEQ_DEL) continue;
            u32 c = tables.op_cidx[row];
            if (c != s.cidx) {
                s.store_container();
                s.load_container(c);
                if (s.err) break;
            }
            i32 ctr = tables.op_counter[row];
            if (pools.cont_epoch[di.cid0 + c] != k) {
                s.checkout(vv, peer, ctr);
                if (lane == 0) pools.cont_epoch[di.cid0 + c] = k;
                __syncwarp();
            }
            i32 len = (i32)tables.op_len[row];
            if (kind == OPK_SEQ_INS) s.insert(peer, ctr, len, tables.op_prop[row]);
            else {
                u32 dl = tables.op_del[row];
                const BlockInfo& bi = tables.blocks[tables.ch_block[ch]];
                u32 tp = tables.peer_map[bi.peer0 + tables.del_peer_idx[dl]];
                s.del(peer, ctr, len, tp, tables.del_counter[dl]);
            }
        }
    }
    // final version = everything applied
    for (u32 c = 0; c < di.C && !s.err; c++) {
        const DocContainer& dc = tables.dcont[di.cid0 + c];
        if (!dc.leaf_cap || (dc.n_leaves == 0 && c != s.cidx)) continue;
        if (c != s.cidx) {
            s.store_container();
            s.load_container(c);
        }
        // final vv: every peer at its end counter
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
#ifdef LB_SIMT_EMU
    if (lane == 0 && getenv("LB_EMU_TRACE")) {
        fprintf(stderr, "seq doc %u: err=%u spans=%u/%u applied=%u\n", warp_global, s.err, s.n_spans, di.span_cap, di.n_applied);
        for (u32 c = 0; c < di.C; c++) {
            const DocContainer& dc = tables.dcont[di.cid0 + c];
            fprintf(stderr, "  cont %u type=%u leaf_cap=%u n_leaves=%u nodes=%u/%u out=%u/%u seq_len=%u ins=%u del=%u\n", c, dc.type,
                    dc.leaf_cap, dc.n_leaves, dc.n_nodes, dc.node_cap, dc.n_out, dc.out_cap, dc.seq_len, dc.n_ins_rows, dc.n_del_rows);
            for (u32 o = 0; o < dc.n_out && o < 8; o++)
                fprintf(stderr, "    run row=%u off=%u len=%u\n", pools.out_row[dc.out0 + o], pools.out_off[dc.out0 + o], pools.out_len[dc.out0 + o]);
        }
    }
#endif
    if (lane == 0) {
        di.n_spans = s.n_spans;
        if (s.err) di.code = s.err;
    }
}
