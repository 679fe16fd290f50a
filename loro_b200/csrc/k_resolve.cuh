// loro_b200 -- phase 3: per-document resolution and causal scan.
//
// Replaces (reference, relative to crates/loro-internal/src):
//   arena.rs register_container / keys / peers interning (doc-level tables)
//   encoding/outdated_encode_reordered.rs:40-83 import_changes_to_oplog (dedupe, lamport from deps, pending)
//   oplog/loro_dag.rs:935-954 get_lamport / get_change_lamport_from_deps
//   oplog/pending_changes.rs:31-140 (changes whose deps are missing stay pending)
//   oplog.rs:402-470 + dag/iter.rs:180-339 (a causal iteration order; the engine always replays from the
//   empty version, and prefers to stay on one peer's chain so that the tracker rarely has to retreat)
//   diff_calc.rs:175-236 (the version vector handed to each calculator before a change)
// Round-1 shape: one thread per document (documents are independent; a batch has 10^3..10^5 of them).
#pragma once
#include "lb_defs.h"

struct ResolveTables {
    // block-level inputs
    const u64* peer_id;
    const u64* key_off; const u32* key_len;
    const u8* cid_root; const u8* cid_type; const u32* cid_peer_idx; const i32* cid_koc;
    const u32* ch_block; const i32* ch_counter; const u32* ch_len; const u32* ch_lamport_wire;
    const u64* ch_dep0; const u32* ch_ndeps; const u8* ch_dep_self;
    const u32* dep_peer_idx; const i32* dep_counter;
    // doc-level outputs (index spaces: peers <-> block peer entries, containers <-> block cid entries,
    // keys <-> block key entries, changes <-> batch-wide change index)
    DocPeer* dpeer; u32* peer_map;
    DocContainer* dcont; u32* cid_map;
    u64* dkey_off; u32* dkey_len; u32* key_map;
    u32* blk_order;      // per doc: its blocks sorted by (peer, counter_start)
    u32* ch_order;       // per doc: changes grouped by peer, counter order (batch-wide change ids)
    u16* ch_peer;        // doc peer idx of each change
    u8* ch_applied;
    u32* ch_lamport;     // recomputed lamport
    u32* ch_walk;        // per doc: applied changes in replay order
    i32* ch_vv;          // per doc: n_changes * P
    u32* ch_pos;         // per change: its position in the doc's ch_order (= row of ch_vv)
    u32* ch_trim;        // per change: leading atoms the document already had when the change arrived
                         // (OpLog::trim_the_known_part_of_change, oplog.rs:181-196: the rest is applied as a slice)
};

__device__ inline bool bytes_eq(const u8* a, const u8* b, u32 n) {
    for (u32 i = 0; i < n; i++)
        if (a[i] != b[i]) return false;
    return true;
}

// thread per doc: intern peers / containers / keys, order the changes per peer.
__global__ void k_doc_tables(const u8* __restrict__ bytes, DocInfo* __restrict__ docs, u32 n_docs,
                             const BlockInfo* __restrict__ blocks, ResolveTables t) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    DocInfo di = docs[d];
    di.P = di.C = di.K = 0;
    di.n_changes = 0;
    di.n_ops = 0;
    if (di.code != DOC_OK) { docs[d] = di; return; }
    for (u32 b = di.b0; b < di.b1; b++)
        if (blocks[b].err) { di.code = blocks[b].err; break; }
    if (di.code != DOC_OK || di.b0 == di.b1) { docs[d] = di; return; }
    const BlockInfo& first = blocks[di.b0];
    const BlockInfo& end = blocks[di.b1];  // sentinel / next doc's first block holds the scan totals
    di.peer0 = first.peer0;
    di.cid0 = first.cid0;
    di.key0 = first.key0;
    di.ch0 = first.ch0;
    di.op0 = first.op0;
    di.n_changes = (u32)(end.ch0 - first.ch0);
    di.n_ops = end.op0 - first.op0;
    di.n_deps = (u32)(end.dep0 - first.dep0);
    di.P_cap = (u32)(end.peer0 - first.peer0);
    di.C_cap = (u32)(end.cid0 - first.cid0);
    di.K_cap = (u32)(end.key0 - first.key0);
    // ---- peers
    u32 P = 0;
    for (u32 b = di.b0; b < di.b1; b++) {
        const BlockInfo& bi = blocks[b];
        for (u32 j = 0; j < bi.n_peers; j++) {
            u64 id = t.peer_id[bi.peer0 + j];
            u32 f = 0;
            while (f < P && t.dpeer[di.peer0 + f].id != id) f++;
            if (f == P) {
                DocPeer np;
                np.id = id;
                np.rank = 0;
                np.first_counter = 0;
                np.end_counter = 0;
                np.max_counter = 0;
                np.atom_base = 0;
                np.ch_first = 0;
                np.ch_count = 0;
                np.pend_lo = 0;
                np.pend_hi = 0;
                t.dpeer[di.peer0 + P] = np;
                P++;
            }
            t.peer_map[bi.peer0 + j] = f;
        }
    }
    if (P > 0xFFF0) { di.code = LB_ERR(DOC_ERR_CAPACITY); docs[d] = di; return; }
    for (u32 i = 0; i < P; i++) {
        u32 r = 0;
        u64 id = t.dpeer[di.peer0 + i].id;
        for (u32 j = 0; j < P; j++) r += t.dpeer[di.peer0 + j].id < id;
        t.dpeer[di.peer0 + i].rank = r;
    }
    // ---- keys (dedupe by bytes)
    u32 K = 0;
    for (u32 b = di.b0; b < di.b1; b++) {
        const BlockInfo& bi = blocks[b];
        for (u32 j = 0; j < bi.n_keys; j++) {
            u64 off = t.key_off[bi.key0 + j];
            u32 len = t.key_len[bi.key0 + j];
            u32 f = 0;
            while (f < K && !(t.dkey_len[di.key0 + f] == len && bytes_eq(bytes + t.dkey_off[di.key0 + f], bytes + off, len))) f++;
            if (f == K) {
                t.dkey_off[di.key0 + K] = off;
                t.dkey_len[di.key0 + K] = len;
                K++;
            }
            t.key_map[bi.key0 + j] = f;
        }
    }
    // ---- containers
    u32 C = 0;
    for (u32 b = di.b0; b < di.b1; b++) {
        const BlockInfo& bi = blocks[b];
        for (u32 j = 0; j < bi.n_cids; j++) {
            u8 is_root = t.cid_root[bi.cid0 + j], type = t.cid_type[bi.cid0 + j];
            i32 koc = t.cid_koc[bi.cid0 + j];
            u64 peer = 0, noff = 0;
            u32 nlen = 0;
            if (is_root) {
                noff = t.key_off[bi.key0 + (u32)koc];
                nlen = t.key_len[bi.key0 + (u32)koc];
            } else
                peer = t.peer_id[bi.peer0 + t.cid_peer_idx[bi.cid0 + j]];
            u32 f = 0;
            for (; f < C; f++) {
                const DocContainer& dc = t.dcont[di.cid0 + f];
                if (dc.is_root != is_root || dc.type != type) continue;
                if (is_root) {
                    if (dc.name_len == nlen && bytes_eq(bytes + dc.name_off, bytes + noff, nlen)) break;
                } else if (dc.peer == peer && dc.counter == koc)
                    break;
            }
            if (f == C) {
                DocContainer dc;
                memset(&dc, 0, sizeof(dc));
                dc.is_root = is_root;
                dc.type = type;
                dc.name_off = noff;
                dc.name_len = nlen;
                dc.peer = peer;
                dc.counter = is_root ? 0 : koc;
                dc.key_or_peer = is_root ? t.key_map[bi.key0 + (u32)koc] : t.peer_map[bi.peer0 + t.cid_peer_idx[bi.cid0 + j]];
                t.dcont[di.cid0 + C] = dc;
                C++;
            }
            t.cid_map[bi.cid0 + j] = f;
        }
    }
    di.P = P;
    di.C = C;
    di.K = K;
    // ---- order blocks by (peer, counter_start) : insertion sort on the doc's slice of blk_order
    u32 nb = di.b1 - di.b0;
    for (u32 i = 0; i < nb; i++) {
        u32 b = di.b0 + i;
        u32 bp = t.peer_map[blocks[b].peer0];
        u32 bc = blocks[b].counter_start;
        u32 j = i;
        while (j > 0) {
            u32 o = t.blk_order[di.b0 + j - 1];
            u32 op = t.peer_map[blocks[o].peer0];
            if (op < bp || (op == bp && blocks[o].counter_start <= bc)) break;
            t.blk_order[di.b0 + j] = o;
            j--;
        }
        t.blk_order[di.b0 + j] = b;
    }
    // ---- per-peer change lists
    u32 k = 0;
    u32 cur_peer = 0xFFFFFFFFu;
    for (u32 i = 0; i < nb; i++) {
        const BlockInfo& bi = blocks[t.blk_order[di.b0 + i]];
        u32 p = t.peer_map[bi.peer0];
        if (p != cur_peer) {
            t.dpeer[di.peer0 + p].ch_first = k;
            cur_peer = p;
        }
        for (u32 c = 0; c < bi.n_changes; c++) {
            u32 ch = (u32)(bi.ch0 + c);
            t.ch_order[di.ch0 + k] = ch;
            t.ch_peer[ch] = (u16)p;
            k++;
        }
        t.dpeer[di.peer0 + p].ch_count += bi.n_changes;
    }
    // blocks of one blob never overlap, blocks of several blobs (import_batch) may: keep every peer's list ordered
    // by counter (stable, so a re-delivered change follows the first copy)
    for (u32 p = 0; p < P; p++) {
        const DocPeer& dp = t.dpeer[di.peer0 + p];
        u32* lst = t.ch_order + di.ch0 + dp.ch_first;
        for (u32 i = 1; i < dp.ch_count; i++) {
            u32 c = lst[i];
            i32 cc = t.ch_counter[c];
            u32 j = i;
            while (j > 0 && t.ch_counter[lst[j - 1]] > cc) { lst[j] = lst[j - 1]; j--; }
            lst[j] = c;
        }
    }
    docs[d] = di;
}

// lamport of atom (peer p, counter c) if applied; returns false when unknown
__device__ inline bool lamport_of(const DocInfo& di, const ResolveTables& t, u32 p, i32 c, u32* out,
                                  u32* ch_out) {
    const DocPeer& dp = t.dpeer[di.peer0 + p];
    if (c < 0 || c >= dp.end_counter) return false;
    // binary search in the peer's ordered change list
    u32 lo = 0, hi = dp.ch_count;
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (t.ch_counter[t.ch_order[di.ch0 + dp.ch_first + mid]] <= c) lo = mid;
        else hi = mid;
    }
    // duplicates (the same change delivered twice inside an import_batch) sort next to each other: the first one
    // is the one that was applied
    while (lo > 0 && t.ch_counter[t.ch_order[di.ch0 + dp.ch_first + lo - 1]] == t.ch_counter[t.ch_order[di.ch0 + dp.ch_first + lo]]) lo--;
    // a re-delivered or sliced copy (export from a version vector cuts a change: A[3..10) next to A[0..10)) was
    // dropped, not applied: the applied change that covers c sorts before it
    while (lo > 0 && !t.ch_applied[t.ch_order[di.ch0 + dp.ch_first + lo]]) lo--;
    u32 ch = t.ch_order[di.ch0 + dp.ch_first + lo];
    if (!t.ch_applied[ch] || c < t.ch_counter[ch] || c >= t.ch_counter[ch] + (i32)t.ch_len[ch]) return false;
    *out = t.ch_lamport[ch] + (u32)(c - t.ch_counter[ch]);
    *ch_out = ch;
    return true;
}

// thread per doc: pending detection, lamport recomputation, replay order, per-change version vectors.
__global__ void k_doc_causal(DocInfo* __restrict__ docs, u32 n_docs, const BlockInfo* __restrict__ blocks,
                             ResolveTables t, u32* __restrict__ peer_cursor) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    DocInfo di = docs[d];
    di.n_applied = 0;
    di.n_pending = 0;
    di.atom_ops = 0;
    di.atom_total = 0;
    if (di.code != DOC_OK || di.n_changes == 0) { docs[d] = di; return; }
    u32 P = di.P;
    u32* cursor = peer_cursor + di.peer0;  // per peer: index of the next unapplied change
    for (u32 p = 0; p < P; p++) cursor[p] = 0;
    u32 cur = 0xFFFFFFFFu;
    u32 walk_n = 0;
    while (true) {
        // ---- pick a ready change: stay on the current peer if possible, else min (lamport, peer)
        u32 pick = 0xFFFFFFFFu, pick_lam = 0, pick_ch = 0;
        for (u32 step = 0; step < P + 1; step++) {
            u32 p;
            if (step == 0) { if (cur == 0xFFFFFFFFu) continue; p = cur; }
            else { p = step - 1; if (p == cur) continue; }
            DocPeer& dp = t.dpeer[di.peer0 + p];
            if (cursor[p] >= dp.ch_count) continue;
            u32 ch = t.ch_order[di.ch0 + dp.ch_first + cursor[p]];
            i32 ctr = t.ch_counter[ch];
            if (ctr != dp.end_counter) {
                if (ctr < dp.end_counter) {  // overlapping duplicate: drop if fully covered
                    if (ctr + (i32)t.ch_len[ch] <= dp.end_counter) { cursor[p]++; step--; continue; }
                    // partial overlap (an update exported from an older version vector): the known part is trimmed and
                    // the rest applies as Change::slice -- it depends on its own predecessor only (change.rs:248-252)
                    u32 l, c2;
                    if (!lamport_of(di, t, p, dp.end_counter - 1, &l, &c2)) continue;
                    u32 lam_t = l + 1;
                    if (step == 0) { pick = p; pick_lam = lam_t; pick_ch = ch; break; }
                    if (pick == 0xFFFFFFFFu || lam_t < pick_lam ||
                        (lam_t == pick_lam && t.dpeer[di.peer0 + p].rank < t.dpeer[di.peer0 + pick].rank)) {
                        pick = p; pick_lam = lam_t; pick_ch = ch;
                    }
                }
                continue;  // gap: predecessor missing -> pending
            }
            bool ok = true;
            u32 lam = 0;
            if (t.ch_dep_self[ch]) {
                u32 l, c2;
                if (!lamport_of(di, t, p, ctr - 1, &l, &c2)) ok = false; else lam = l + 1;
            }
            u64 d0 = t.ch_dep0[ch];
            u32 bpeer0 = (u32)0;
            const BlockInfo& bi = blocks[t.ch_block[ch]];
            (void)bpeer0;
            for (u32 k = 0; ok && k < t.ch_ndeps[ch]; k++) {
                u32 dp_idx = t.peer_map[bi.peer0 + t.dep_peer_idx[d0 + k]];
                u32 l, c2;
                if (!lamport_of(di, t, dp_idx, t.dep_counter[d0 + k], &l, &c2)) ok = false;
                else if (l + 1 > lam) lam = l + 1;
            }
            if (!ok) continue;
            if (step == 0) { pick = p; pick_lam = lam; pick_ch = ch; break; }
            if (pick == 0xFFFFFFFFu || lam < pick_lam ||
                (lam == pick_lam && t.dpeer[di.peer0 + p].rank < t.dpeer[di.peer0 + pick].rank)) {
                pick = p; pick_lam = lam; pick_ch = ch;
            }
        }
        if (pick == 0xFFFFFFFFu || di.code != DOC_OK) break;
        // ---- apply
        u32 ch = pick_ch;
        DocPeer& dp = t.dpeer[di.peer0 + pick];
        i32 ctr = t.ch_counter[ch];
        const u32 trim = ctr < dp.end_counter ? (u32)(dp.end_counter - ctr) : 0u;
        u32 local = dp.ch_first + cursor[pick];  // row of this change in the doc's ch_vv
        i32* v = t.ch_vv + di.vv0 + (u64)local * P;
        for (u32 q = 0; q < P; q++) v[q] = 0;
        const BlockInfo& bi = blocks[t.ch_block[ch]];
        u32 ndeps = trim ? 1u : t.ch_ndeps[ch] + (t.ch_dep_self[ch] ? 1 : 0);
        for (u32 k = 0; k < ndeps; k++) {
            u32 dpi;
            i32 dc;
            if (trim) { dpi = pick; dc = dp.end_counter - 1; }
            else if (k == t.ch_ndeps[ch]) { dpi = pick; dc = ctr - 1; }
            else { dpi = t.peer_map[bi.peer0 + t.dep_peer_idx[t.ch_dep0[ch] + k]]; dc = t.dep_counter[t.ch_dep0[ch] + k]; }
            u32 l = 0, dch = ch;
            lamport_of(di, t, dpi, dc, &l, &dch);
            const i32* dv = t.ch_vv + di.vv0 + (u64)t.ch_pos[dch] * P;   // the dependency is applied: its row is known
            for (u32 q = 0; q < P; q++) if (dv[q] > v[q]) v[q] = dv[q];
            if (dc + 1 > v[dpi]) v[dpi] = dc + 1;
        }
        t.ch_lamport[ch] = pick_lam - trim;   // lamport of the change's (trimmed) first atom: counters and lamports run in step
        t.ch_trim[ch] = trim;
        t.ch_applied[ch] = 1;
        t.ch_pos[ch] = local;
        t.ch_walk[di.ch0 + walk_n++] = ch;
        dp.end_counter = ctr + (i32)t.ch_len[ch];
        di.atom_ops += t.ch_len[ch] - trim;
        cursor[pick]++;
        cur = pick;
    }
    di.n_applied = walk_n;
    // ---- pending bookkeeping + atom bases
    u32 base = 0;
    for (u32 p = 0; p < P; p++) {
        DocPeer& dp = t.dpeer[di.peer0 + p];
        dp.atom_base = base;
        base += (u32)dp.end_counter;
        dp.pend_lo = dp.pend_hi = 0;
        for (u32 k = cursor[p]; k < dp.ch_count; k++) {
            u32 ch = t.ch_order[di.ch0 + dp.ch_first + k];
            i32 c0 = t.ch_counter[ch], c1 = c0 + (i32)t.ch_len[ch];
            if (c1 <= dp.end_counter) continue;
            if (dp.pend_lo == dp.pend_hi) { dp.pend_lo = c0 < dp.end_counter ? dp.end_counter : c0; dp.pend_hi = c1; }
            else { if (c0 < dp.pend_lo) dp.pend_lo = c0; if (c1 > dp.pend_hi) dp.pend_hi = c1; }
            di.n_pending++;
        }
    }
    di.atom_total = base;
    docs[d] = di;
}

// thread per doc: frontiers = the heads of the causal graph (reference: version/frontiers.rs:233-246
// update_frontiers_on_new_change, oplog/loro_dag.rs:251-269).  The last id of peer p is a head unless it lies in the
// causal past of another peer's change; version vectors grow along a peer's chain, so looking at every peer's LAST
// applied change is enough.
__global__ void k_doc_frontiers(const DocInfo* __restrict__ docs, u32 n_docs, ResolveTables t) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const DocInfo& di = docs[d];
    if (di.code != DOC_OK) return;
    u32 P = di.P;
    for (u32 p = 0; p < P; p++) t.dpeer[di.peer0 + p].is_head = t.dpeer[di.peer0 + p].end_counter > 0 ? 1u : 0u;
    for (u32 q = 0; q < P; q++) {
        const DocPeer& dq = t.dpeer[di.peer0 + q];
        // last applied change of q
        u32 last = 0xFFFFFFFFu;
        for (u32 k = dq.ch_count; k-- > 0;) {
            u32 ch = t.ch_order[di.ch0 + dq.ch_first + k];
            if (t.ch_applied[ch]) { last = ch; break; }
        }
        if (last == 0xFFFFFFFFu) continue;
        const i32* v = t.ch_vv + di.vv0 + (u64)t.ch_pos[last] * P;
        for (u32 p = 0; p < P; p++)
            if (p != q && v[p] >= t.dpeer[di.peer0 + p].end_counter && t.dpeer[di.peer0 + p].end_counter > 0) t.dpeer[di.peer0 + p].is_head = 0;
    }
}
