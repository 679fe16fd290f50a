// loro_b200 -- phase 5b: movable tree (one warp per document).
//
// Replaces (reference, relative to crates/loro-internal/src):
//   diff_calc/tree.rs:21-148 (TreeDiffCalculator: ops of a tree in (lamport, peer) order), :417-452
//     (MoveLamportAndID ordering), :471-508 (TreeCacheForDiff::apply / is_ancestor_of: a move whose new parent is
//     a descendant of the target is recorded but not effected)
//   state/tree_state.rs:690-747 (TreeState::mov with the cycle check, used for imports into a fresh document),
//     :592-595 + :61-67 (children ordered by (fractional index bytes, lamport, peer) = NodePosition),
//     :749-760 (is_node_deleted: a node is alive iff its parent chain reaches the root)
//   loro-common/src/lib.rs:631 (DELETED_TREE_ROOT: a delete is a move under it)
//
// Shape: a tree's ops are totally ordered by (lamport, peer) and every op depends on the tree the earlier ones
// left, so the apply is sequential per tree; the parallelism is across documents (config C5: 10^4 of them).
// Everything around the sequential core is lane-parallel: the document's RawTreeMove records are sorted by a
// 64-bit key with a warp bitonic network over global memory (the records are L2-resident), 32 records are fetched
// per round trip, and the sibling lists come out of a second sort of the nodes by (parent slot, position prefix)
// with the full NodePosition comparison as tie-break.  Node tables are dense arrays over the document's atoms
// (a TreeID is the id of its create op), so "node -> parent" is one load.
// Nodes of different tree containers share the atom-indexed tables: in a well-formed document their ids are
// disjoint; a hostile blob that moves a node of one tree inside another gets a memory-safe, cycle-free result.
#pragma once
#include "lb_defs.h"

struct TreeTables {
    const DocPeer* dpeer;
    const BlockInfo* blocks;
    const u32* op_cidx; const u32* op_lamport;
    const uint4* tr_rec;      // per tree op (k_op_classify): target atom, parent atom | TREE_ROOT | TREE_DELETED, position, row
    const u64* tr_key;        // (lamport << 32 | peer rank << 16), ~0 for ops that are not applied
    u64* ts_key; u32* ts_val; // sort space, one entry per tree op
    const u64* pos_off; const u32* pos_len; const u8* pos_pool;
    // per document: S = atom_total + C slots starting at DocInfo::tree0.  Slots [0, atom_total) are nodes, slot
    // atom_total + c is the root of tree container c.
    u32* tn_parent;           // [node] TREE_UNEXIST | TREE_ROOT | TREE_DELETED | parent node
    u32* tn_move;             // [node] tree op of the last effective move (position, lamport, peer of the node)
    u32* tn_base;             // [slot] first child in tn_child
    u32* tn_cnt;              // [slot] number of children
    u32* tn_sib;              // [node] index among its siblings
    u64* ns_key;              // [node] sort key: parent slot << 32 | first four position bytes
    u32* tn_child;            // [node] after the sort: nodes grouped by parent slot, in sibling order
};

// ---- warp bitonic sort of (key, val) pairs in global memory, any n (partners beyond n act as +inf: with every
// comparator pointing the same way they never have to move).  `tie(a, b)` orders two vals whose keys are equal.
template <class Tie>
__device__ inline void warp_sort_pairs(u64* key, u32* val, u32 n, int lane, Tie tie) {
    if (n < 2) return;
    for (u32 k = 2; (k >> 1) < n; k <<= 1) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            bool flip = j == (k >> 1);
            for (u32 i = (u32)lane; i < n; i += 32) {
                u32 l = flip ? (i ^ (k - 1)) : (i ^ j);
                if (l > i && l < n) {
                    u64 ki = key[i], kl = key[l];
                    bool sw = ki > kl;
                    u32 vi = val[i], vl = val[l];
                    if (ki == kl && ki != ~0ull) sw = tie(vl, vi);   // +inf entries carry no payload to compare
                    if (sw) { key[i] = kl; key[l] = ki; val[i] = vl; val[l] = vi; }
                }
            }
            __syncwarp();
        }
    }
}
struct NoTie { __device__ bool operator()(u32, u32) const { return false; } };

// lexicographic comparison of two fractional indexes (FractionalIndex derives Ord on its bytes)
__device__ inline int pos_cmp(const TreeTables& t, u32 pa, u32 pb) {
    if (pa == pb) return 0;
    const u8* a = t.pos_pool + t.pos_off[pa];
    const u8* b = t.pos_pool + t.pos_off[pb];
    u32 la = t.pos_len[pa], lb = t.pos_len[pb];
    u32 n = la < lb ? la : lb;
    for (u32 i = 0; i < n; i++)
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return la < lb ? -1 : (la > lb ? 1 : 0);
}

__global__ void k_tree_build(DocInfo* __restrict__ docs, u32 n_docs, TreeTables t) {
    u32 d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (d >= n_docs) return;
    const DocInfo& di = docs[d];
    if (di.code != DOC_OK || !di.has_tree) return;
    const u32 A = (u32)di.atom_total, C = di.C;
    const u64 base = di.tree0;
    u32* parent = t.tn_parent + base;
    u32* move = t.tn_move + base;
    for (u32 i = lane; i < A + C; i += 32) { parent[i] = TREE_UNEXIST; t.tn_cnt[base + i] = 0; t.tn_base[base + i] = 0; }
    // ---- the document's tree ops (contiguous: blocks of a document are) in (lamport, peer) order
    const u64 tr_lo = t.blocks[di.b0].tr0, tr_hi = t.blocks[di.b1].tr0;
    const u32 n_tr = (u32)(tr_hi - tr_lo);
    u64* skey = t.ts_key + tr_lo;
    u32* sval = t.ts_val + tr_lo;
    for (u32 i = lane; i < n_tr; i += 32) { skey[i] = t.tr_key[tr_lo + i]; sval[i] = i; }
    __syncwarp();
    warp_sort_pairs(skey, sval, n_tr, lane, NoTie());
    // ---- sequential apply, 32 records per round trip
    bool stop = false;
    for (u32 j0 = 0; j0 < n_tr && !stop; j0 += 32) {
        u32 j = j0 + (u32)lane;
        uint4 rec;
        rec.x = rec.y = rec.z = 0; rec.w = 0xFFFFFFFFu;
        u32 ti_l = 0;
        if (j < n_tr && skey[j] != ~0ull) { ti_l = sval[j]; rec = t.tr_rec[tr_lo + ti_l]; }
        u32 cnt = n_tr - j0 < 32 ? n_tr - j0 : 32;
        for (u32 s = 0; s < cnt; s++) {
            u32 row = __shfl_sync(LB_FULL, rec.w, (int)s);
            if (row == 0xFFFFFFFFu) { stop = true; break; }   // not applied: these sort last
            u32 target = __shfl_sync(LB_FULL, rec.x, (int)s);
            u32 np = __shfl_sync(LB_FULL, rec.y, (int)s);
            u32 ti = __shfl_sync(LB_FULL, ti_l, (int)s);
            bool effected = true;
            if (np < TREE_UNEXIST && parent[target] != TREE_UNEXIST) {
                // is the target an ancestor of (or equal to) the new parent?  (tree.rs:477-508, tree_state.rs:727-747)
                u32 cur = np;
                for (u32 guard = 0; guard <= A; guard++) {
                    if (cur == target) { effected = false; break; }
                    u32 pp = parent[cur];
                    if (pp >= TREE_UNEXIST) break;
                    cur = pp;
                }
            }
            __syncwarp();
            if (effected && lane == 0) { parent[target] = np; move[target] = ti; }
            __syncwarp();
        }
    }
    __syncwarp();
    // ---- sibling lists: sort the nodes by (parent slot, position, lamport, peer)
    u64* nkey = t.ns_key + base;
    u32* child = t.tn_child + base;
    for (u32 a = lane; a < A; a += 32) {
        u32 p = parent[a];
        u64 key = ~0ull;
        if (p != TREE_UNEXIST && p != TREE_DELETED) {
            uint4 rec = t.tr_rec[tr_lo + move[a]];
            u32 slot = p == TREE_ROOT ? A + t.op_cidx[rec.w] : p;
            const u8* pb = t.pos_pool + t.pos_off[rec.z];
            u32 pl = t.pos_len[rec.z];
            u32 pre = 0;
            for (u32 k = 0; k < 4; k++) pre = (pre << 8) | (k < pl ? pb[k] : 0u);
            key = ((u64)slot << 32) | pre;
        }
        nkey[a] = key;
        child[a] = a;
    }
    __syncwarp();
    warp_sort_pairs(nkey, child, A, lane, [&](u32 a, u32 b) -> bool {   // a before b ?
        uint4 ra = t.tr_rec[tr_lo + move[a]], rb = t.tr_rec[tr_lo + move[b]];
        int c = pos_cmp(t, ra.z, rb.z);
        if (c) return c < 0;
        return t.tr_key[tr_lo + move[a]] < t.tr_key[tr_lo + move[b]];   // (lamport, peer): NodePosition.idlp
    });
    // ---- slot -> (first child, count), node -> sibling index
    for (u32 j = lane; j < A; j += 32) {
        u64 k = nkey[j];
        if (k == ~0ull) continue;
        u32 slot = (u32)(k >> 32);
        if (j == 0 || (u32)(nkey[j - 1] >> 32) != slot) t.tn_base[base + slot] = j;
    }
    __syncwarp();
    for (u32 j = lane; j < A; j += 32) {
        u64 k = nkey[j];
        if (k == ~0ull) continue;
        u32 slot = (u32)(k >> 32);
        t.tn_sib[base + child[j]] = j - t.tn_base[base + slot];
        if (j + 1 == A || (u32)(nkey[j + 1] >> 32) != slot) t.tn_cnt[base + slot] = j + 1 - t.tn_base[base + slot];
    }
}
