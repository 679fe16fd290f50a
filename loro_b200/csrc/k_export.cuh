// loro_b200 -- phase 7: re-export of every document as a FastUpdates blob (ExportMode::all_updates).
//
// Replaces (reference, relative to crates/loro-internal/src):
//   encoding.rs:350-416 (export, header + xxHash32), encoding/fast_snapshot.rs:257-260 (body framing),
//   oplog/change_store.rs:494-576,711-764,1244-1291 (export_blocks_from -> a fresh store: block packing with
//     MAX_BLOCK_SIZE, ChangesBlock::push_change), change.rs:128-139,268-283 (estimate_storage_size,
//     can_merge_right), op.rs:143-148 + container/list/list_op.rs:189-278,381-434,516-658 (RleVec merge rules),
//   oplog/change_store/block_encode.rs:137-278 (encode_block), block_meta_encode.rs:13-88 (encode_changes),
//   encoding/arena.rs:103-147 (ContainerArena), serde_columnar 0.3.14 column encoders (BoolRle, AnyRle,
//   DeltaRle, DeltaOfDelta).
//
// What the imported document's change store looks like is re-derived from the decoded tables, per peer:
//   A  ops of each decoded change go through the RleVec merge (block_encode.rs:651),
//   B  changes enter the store in counter order (ChangeStore::insert_change, merge_interval 0 for imports),
//   C  export re-inserts the stored changes into a fresh store (export_blocks_from) -- same rules, new sizes.
// A thread per document runs A-C and leaves compact lists of merged ops / merged changes / output blocks;
// a thread per output block then encodes it (two passes: sizes, bytes).
// Whether two neighbouring inserts merge depends on where their payloads landed in the importing document's
// arenas (arena.rs:237-263): values are adjacent when nothing else was allocated in between (decode order),
// strings additionally need the append-only buffer not to have been reallocated (capacity doubles from 32;
// restated from append-only-bytes 0.1.12, same model as the oracle, unpinned by reference tests).
// Not yet covered (lb_doc_export_updates answers LB_ERR_UNSUPPORTED for the document): changes whose size
// estimate exceeds one block (split_change_then_insert), values containing nested maps (block-local key indices
// inside the payload), documents with pending changes; Tree/MovableList/styles never reach this phase.
#pragma once
#include "lb_defs.h"

enum { XK_NONE = 0, XK_LIST = 1, XK_TEXT = 2, XK_DEL = 3, XK_MAPSET = 4, XK_MAPDEL = 5 };
#define LB_MAX_BLOCK_SIZE 4096   // change_store.rs:37

struct XDoc {          // per document
    u32 n_st, n_mo, n_mc, n_mb;   // store positions, merged ops, merged changes, output blocks
    u32 flags;                    // bit0: export unsupported for this document
    u32 pad;
    u64 ob0;                      // first output block (batch-wide)
    u64 scratch0;                 // first scratch word of this doc's blocks
    u64 exp_off;                  // offset of the blob in the export buffer
    u32 exp_len;
    u32 n_blocks_pad;
};
struct XBlock {        // one output block
    u32 doc;
    u32 mc0, mc1;      // merged-change range (absolute indices into mc_*)
    u32 len;           // encoded bytes (without the ULEB length prefix)
    u32 sec_len[8];
    u32 col_len[8];    // ops columns 0-3, delete columns 4-6
    u64 off;           // offset of the block bytes inside the export buffer
    u64 scratch;       // scratch words of this block
};

struct ExportTables {
    const u8* bytes; const BlockInfo* blocks; const DocPeer* dpeer; const DocContainer* dcont;
    const u64* dkey_off; const u32* dkey_len; const u32* key_map; const u32* peer_map;
    const u32* ch_order; const u8* ch_applied; const u32* ch_block; const i32* ch_counter; const u32* ch_len;
    const u32* ch_lamport; const i64* ch_ts; const u64* ch_op0; const u32* ch_nops;
    const u64* ch_dep0; const u32* ch_ndeps; const u8* ch_dep_self; const u32* dep_peer_idx; const i32* dep_counter;
    const u64* ch_msg_off; const u32* ch_msg_len;
    const u8* op_kind; const u32* op_cidx; const i32* op_prop; const u32* op_len; const i32* op_counter;
    const u64* op_val_off; const u32* op_val_len; const u32* op_del; const u32* op_aux;
    const i32* del_counter; const i32* del_len;
    // per row
    u32* r_astart;     // arena position (values: atoms, strings: bytes) relative to the document
    u32* r_bytes;      // text rows: payload bytes
    u32* st_row;       // store position -> row (per doc region [op0, op0 + n_st))
    // merged ops (per doc region [op0, op0 + n_mo))
    u8* mo_xk; u32* mo_cidx; i32* mo_ctr; u32* mo_atoms; i32* mo_prop; u32* mo_f0; u32* mo_f1; i32* mo_f2;
    u32* mo_st0; u32* mo_nst;
    // merged changes (per doc region [ch0, ch0 + n_mc)) and output blocks (first merged change, same region)
    u32* mc_src; u32* mc_o0; u32* mc_no; u32* mc_atoms;
    u32* mb_first;
    XDoc* xdoc;
};

// ---------------------------------------------------------------------------------------------- byte sink
struct XSink {
    u8* dst;   // nullptr = counting
    u64 n;
    __device__ __forceinline__ void put(u8 c) { if (dst) dst[n] = c; n++; }
    __device__ __forceinline__ void varint(u64 v) { while (v >= 0x80) { put((u8)(v | 0x80)); v >>= 7; } put((u8)v); }
    __device__ __forceinline__ void zigzag(i64 v) { varint(((u64)v << 1) ^ (u64)(v >> 63)); }
    __device__ __forceinline__ void copy(const u8* s, u64 len) {
        if (dst) for (u64 i = 0; i < len; i++) dst[n + i] = s[i];
        n += len;
    }
};
__device__ __forceinline__ u32 varint_len(u64 v) { u32 k = 1; while (v >= 0x80) { v >>= 7; k++; } return k; }

// ---------------------------------------------------------------------------------------------- merge rules
struct XOp {   // one (possibly merged) op: the fields the merge rules and the encoder need
    u8 xk; u32 cidx; i32 ctr; u32 atoms; i32 prop; u32 f0, f1; i32 f2; u32 st0, nst;
    // LIST/TEXT: f0 = arena start, f1 = arena end (TEXT: bytes; f1 - f0 = payload bytes)
    // DEL: f0 = target peer (doc-level), f1 = lowest target counter, f2 = signed length
};
__device__ __forceinline__ XOp xop_load(const ExportTables& t, u64 i) {
    XOp o;
    o.xk = t.mo_xk[i]; o.cidx = t.mo_cidx[i]; o.ctr = t.mo_ctr[i]; o.atoms = t.mo_atoms[i]; o.prop = t.mo_prop[i];
    o.f0 = t.mo_f0[i]; o.f1 = t.mo_f1[i]; o.f2 = t.mo_f2[i]; o.st0 = t.mo_st0[i]; o.nst = t.mo_nst[i];
    return o;
}
__device__ __forceinline__ void xop_store(const ExportTables& t, u64 i, const XOp& o) {
    t.mo_xk[i] = o.xk; t.mo_cidx[i] = o.cidx; t.mo_ctr[i] = o.ctr; t.mo_atoms[i] = o.atoms; t.mo_prop[i] = o.prop;
    t.mo_f0[i] = o.f0; t.mo_f1[i] = o.f1; t.mo_f2[i] = o.f2; t.mo_st0[i] = o.st0; t.mo_nst[i] = o.nst;
}
__device__ __forceinline__ u32 xop_estimate(const XOp& o) {   // list_op.rs:109-123, op/content.rs:70-77
    switch (o.xk) {
        case XK_LIST: return 4 * o.atoms;
        case XK_TEXT: return o.f1 - o.f0;
        case XK_DEL: return 8;
        default: return 3;
    }
}
// string arena generation: number of capacity doublings (from 32) needed to hold `end` bytes
__device__ __forceinline__ u32 str_gen(u32 end) {
    u32 k = 0;
    u64 cap = 32;
    while (cap < end) { cap <<= 1; k++; }
    return k;
}
// DeleteSpan helpers (list_op.rs:381-444); prop = pos, f2 = signed len, (f0, f1) = id_start
__device__ __forceinline__ bool d_bidi(const XOp& o) { return o.f2 == 1 || o.f2 == -1; }
__device__ __forceinline__ i64 d_start_pos(const XOp& o) { return o.f2 > 0 ? o.prop : (i64)o.prop + 1 + o.f2; }
__device__ __forceinline__ i64 d_next_pos(const XOp& o) { return o.f2 > 0 ? d_start_pos(o) : d_start_pos(o) - 1; }
__device__ __forceinline__ i64 d_prev_pos(const XOp& o) { return o.f2 > 0 ? o.prop : (i64)o.prop + 1; }
__device__ __forceinline__ i64 d_id_end(const XOp& o) { return (i64)(i32)o.f1 + (o.f2 < 0 ? -o.f2 : o.f2); }
__device__ inline bool xop_mergable(const XOp& a, const XOp& b) {   // op.rs:143-148 + list_op.rs:516-552
    if (a.ctr + (i32)a.atoms != b.ctr || a.cidx != b.cidx || a.xk != b.xk) return false;
    switch (a.xk) {
        case XK_LIST: return (i64)a.prop + a.atoms == b.prop && a.f1 == b.f0;
        case XK_TEXT: return (i64)a.prop + a.atoms == b.prop && a.f1 == b.f0 && str_gen(a.f1) == str_gen(b.f1);
        case XK_DEL: {
            if (a.f0 != b.f0) return false;    // ids of different peers never line up
            bool ab = d_bidi(a), bb = d_bidi(b);
            i64 as = (i32)a.f1, bs = (i32)b.f1;
            if (ab && bb) return (a.prop == b.prop && as + 1 == bs) || ((i64)a.prop == (i64)b.prop + 1 && as == bs + 1);
            if (ab && !bb) {
                if (a.prop == d_prev_pos(b)) return b.f2 > 0 ? as + 1 == bs : as == d_id_end(b);
                return false;
            }
            if (!ab && bb) {
                if (d_next_pos(a) == b.prop) return a.f2 > 0 ? d_id_end(a) == bs : as == bs + 1;
                return false;
            }
            if (d_next_pos(a) == b.prop && (a.f2 > 0) == (b.f2 > 0)) return a.f2 > 0 ? d_id_end(a) == bs : as == d_id_end(b);
            return false;
        }
        default: return false;
    }
}
__device__ inline void xop_merge(XOp& a, const XOp& b) {
    switch (a.xk) {
        case XK_LIST: case XK_TEXT: a.f1 = b.f1; break;
        case XK_DEL: {   // list_op.rs:244-249, 398-434
            bool ab = d_bidi(a), bb = d_bidi(b);
            i32 as = (i32)a.f1, bs = (i32)b.f1;
            i32 nl;
            if (ab && bb) nl = a.prop == b.prop ? 2 : -2;
            else if (ab && !bb) nl = b.f2 + (b.f2 > 0 ? 1 : -1);
            else if (!ab && bb) nl = a.f2 + (a.f2 > 0 ? 1 : -1);
            else nl = a.f2 + b.f2;
            a.f1 = (u32)(as < bs ? as : bs);
            a.f2 = nl;
            break;
        }
        default: break;
    }
    a.atoms += b.atoms;
    a.nst += b.nst;
}

// ---------------------------------------------------------------------------------------------- X1: arenas
// thread per document: arena positions of every row in decode order (the importing document allocates while it
// decodes: block_encode.rs:619-657)
__global__ void k_exp_arena(const DocInfo* __restrict__ docs, u32 n_docs, ExportTables t) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const DocInfo& di = docs[d];
    XDoc x;
    x.n_st = x.n_mo = x.n_mc = x.n_mb = 0;
    x.flags = 0; x.pad = 0; x.ob0 = 0; x.scratch0 = 0; x.exp_off = 0; x.exp_len = 0; x.n_blocks_pad = 0;
    if (di.code == DOC_OK) {
        u32 vals = 0, strs = 0;
        for (u64 r = di.op0; r < di.op0 + di.n_ops; r++) {
            u8 k = t.op_kind[r];
            if (k != OPK_SEQ_INS) continue;   // kind of a skipped (pending) row is OPK_SKIP: see k_exp_pack
            const DocContainer& dc = t.dcont[di.cid0 + t.op_cidx[r]];
            if (dc.type == CT_TEXT) {
                Cur c(t.bytes + t.op_val_off[r], t.op_val_len[r]);
                u32 n = (u32)c.varint();
                t.r_astart[r] = strs;
                t.r_bytes[r] = n;
                strs += n;
            } else {
                t.r_astart[r] = vals;
                vals += t.op_len[r];
            }
        }
    }
    t.xdoc[d] = x;
}

// ---------------------------------------------------------------------------------------------- X2: packing
struct XChange {   // a (possibly merged) change during packing
    u32 src;       // first source change: id, lamport, deps, timestamp, message
    u32 o0, no;    // ops [o0, o0+no) in the mo arrays (absolute)
    u32 atoms;
};
struct XStore {    // the store of one peer while changes are inserted in counter order
    bool have_block;
    u32 blk_est;
    u32 last_src;  // first source change of the block's last (merged) change
    u32 w_op;      // next free merged-op slot (absolute)
    u32 w_ch;      // next free merged-change slot (absolute)
};
__device__ __forceinline__ bool xmsg_same(const ExportTables& t, u32 a, u32 b) {
    u32 la = t.ch_msg_len[a], lb = t.ch_msg_len[b];
    if (la != lb) return false;
    const u8* pa = t.bytes + t.ch_msg_off[a];
    const u8* pb = t.bytes + t.ch_msg_off[b];
    for (u32 i = 0; i < la; i++)
        if (pa[i] != pb[i]) return false;
    return true;
}
// ChangeStore::insert_change + ChangesBlock::push_change (change_store.rs:711-764, 1244-1291) for the next change
// of the peer.  Ops of X sit at [X.o0, X.o0 + X.no) at or after s.w_op; they are compacted down to s.w_op.
// `mark_blocks`: record block starts (final pass only).  Returns false when the change would have to be split.
__device__ inline bool xstore_insert(const ExportTables& t, XStore& s, XChange X, bool mark_blocks, bool split_when_exceeds) {
    u32 ndeps = t.ch_ndeps[X.src] + (t.ch_dep_self[X.src] ? 1u : 0u);
    u32 est = 4 + (ndeps > 1 ? (ndeps - 1) * 4 : 0);
    for (u32 i = 0; i < X.no; i++) est += xop_estimate(xop_load(t, X.o0 + i));
    bool ok = true;
    if (est > LB_MAX_BLOCK_SIZE && split_when_exceeds) ok = false;   // split_change_then_insert: not restated yet
    if (s.have_block) {
        bool is_full = est + s.blk_est > LB_MAX_BLOCK_SIZE;
        bool can = t.ch_dep_self[X.src] && t.ch_ndeps[X.src] == 0 && t.ch_ts[X.src] <= t.ch_ts[s.last_src] &&
                   xmsg_same(t, s.last_src, X.src);
        bool single = false;
        if (can && is_full && X.no == 1) single = xop_mergable(xop_load(t, s.w_op - 1), xop_load(t, X.o0));
        if (can && (!is_full || single)) {
            XOp back = xop_load(t, s.w_op - 1);
            for (u32 i = 0; i < X.no; i++) {
                XOp o = xop_load(t, X.o0 + i);
                if (xop_mergable(back, o)) xop_merge(back, o);
                else {
                    xop_store(t, s.w_op - 1, back);
                    s.blk_est += xop_estimate(o);
                    back = o;
                    s.w_op++;
                }
            }
            xop_store(t, s.w_op - 1, back);
            u32 lc = s.w_ch - 1;
            t.mc_no[lc] = s.w_op - t.mc_o0[lc];
            t.mc_atoms[lc] += X.atoms;
            return ok;
        }
        if (!is_full) {
            s.blk_est += est;
            goto append;
        }
    }
    // a new block starts with this change
    s.have_block = true;
    s.blk_est = est;
    if (mark_blocks) t.mb_first[s.w_ch] = 1;
append:
    for (u32 i = 0; i < X.no; i++) {
        if (s.w_op + i != X.o0 + i) xop_store(t, s.w_op + i, xop_load(t, X.o0 + i));
    }
    t.mc_src[s.w_ch] = X.src;
    t.mc_o0[s.w_ch] = s.w_op;
    t.mc_no[s.w_ch] = X.no;
    t.mc_atoms[s.w_ch] = X.atoms;
    s.w_op += X.no;
    s.w_ch++;
    s.last_src = X.src;
    return ok;
}

// thread per document
__global__ void k_exp_pack(const DocInfo* __restrict__ docs, u32 n_docs, ExportTables t) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const DocInfo& di = docs[d];
    if (di.code != DOC_OK) return;
    XDoc x = t.xdoc[d];
    u32 op_base = (u32)di.op0, ch_base = (u32)di.ch0;
    u32 w_st = op_base, w_op = op_base, w_ch = ch_base;   // absolute write positions
    bool unsupported = false;
    for (u32 rank = 0; rank < di.P; rank++) {   // blocks are keyed by (peer id, counter): ascending peer id
        u32 p = 0;
        while (p < di.P && t.dpeer[di.peer0 + p].rank != rank) p++;
        if (p == di.P) break;
        const DocPeer& dp = t.dpeer[di.peer0 + p];
        u32 a_op0 = w_op, a_ch0 = w_ch;
        // ---- stage A: decoded changes, ops re-pushed through the RleVec merge
        u32 a_op = a_op0, a_ch = a_ch0;
        for (u32 k = 0; k < dp.ch_count; k++) {
            u32 ch = t.ch_order[di.ch0 + dp.ch_first + k];
            if (!t.ch_applied[ch]) continue;
            u64 r0 = t.ch_op0[ch];
            u32 nr = t.ch_nops[ch];
            u32 first = a_op;
            for (u32 r = 0; r < nr; r++) {
                u64 row = r0 + r;
                XOp o;
                u8 kind = t.op_kind[row];
                o.cidx = t.op_cidx[row];
                o.ctr = t.op_counter[row];
                o.atoms = t.op_len[row];
                o.prop = t.op_prop[row];
                o.f0 = o.f1 = 0; o.f2 = 0;
                o.st0 = w_st; o.nst = 1;
                t.st_row[w_st++] = (u32)row;
                switch (kind) {
                    case OPK_SEQ_INS:
                        if (t.dcont[di.cid0 + o.cidx].type == CT_TEXT) { o.xk = XK_TEXT; o.f0 = t.r_astart[row]; o.f1 = o.f0 + t.r_bytes[row]; }
                        else { o.xk = XK_LIST; o.f0 = t.r_astart[row]; o.f1 = o.f0 + o.atoms; }
                        break;
                    case OPK_SEQ_DEL: {
                        u32 dl = t.op_del[row];
                        o.xk = XK_DEL; o.f0 = t.op_aux[row]; o.f1 = (u32)t.del_counter[dl]; o.f2 = t.del_len[dl];
                        break;
                    }
                    case OPK_MAP_SET: case OPK_MAP_DEL:
                        o.xk = kind == OPK_MAP_SET ? XK_MAPSET : XK_MAPDEL;
                        o.prop = (i32)t.key_map[t.blocks[t.ch_block[ch]].key0 + (u32)o.prop];
                        break;
                    default: o.xk = XK_NONE; unsupported = true;
                }
                if (a_op > first) {
                    XOp back = xop_load(t, a_op - 1);
                    if (xop_mergable(back, o)) { xop_merge(back, o); xop_store(t, a_op - 1, back); continue; }
                }
                xop_store(t, a_op++, o);
            }
            t.mc_src[a_ch] = ch;
            t.mc_o0[a_ch] = first;
            t.mc_no[a_ch] = a_op - first;
            t.mc_atoms[a_ch] = t.ch_len[ch];
            a_ch++;
        }
        // ---- stage B: import (insert_change, merge_interval 0, split_when_exceeds)
        XStore s;
        s.have_block = false; s.blk_est = 0; s.last_src = 0; s.w_op = a_op0; s.w_ch = a_ch0;
        for (u32 k = a_ch0; k < a_ch; k++) {
            XChange X;
            X.src = t.mc_src[k]; X.o0 = t.mc_o0[k]; X.no = t.mc_no[k]; X.atoms = t.mc_atoms[k];
            if (!xstore_insert(t, s, X, false, true)) unsupported = true;
        }
        u32 b_op = s.w_op, b_ch = s.w_ch;
        (void)b_op;
        // ---- stage C: export (export_blocks_from -> fresh store, no splitting)
        s.have_block = false; s.blk_est = 0; s.last_src = 0; s.w_op = a_op0; s.w_ch = a_ch0;
        for (u32 k = a_ch0; k < b_ch; k++) t.mb_first[k] = 0;
        for (u32 k = a_ch0; k < b_ch; k++) {
            XChange X;
            X.src = t.mc_src[k]; X.o0 = t.mc_o0[k]; X.no = t.mc_no[k]; X.atoms = t.mc_atoms[k];
            xstore_insert(t, s, X, true, false);
        }
        for (u32 k = a_ch0; k < s.w_ch; k++) x.n_mb += t.mb_first[k];
        w_op = s.w_op;
        w_ch = s.w_ch;
    }
    x.n_st = w_st - op_base;
    x.n_mo = w_op - op_base;
    x.n_mc = w_ch - ch_base;
    if (unsupported || (di.has_unsupported & 0x7FFFFFFFu) || di.n_pending) x.flags |= 1;
    if (x.flags & 1) x.n_mb = 0;
    t.xdoc[d] = x;
}

// thread per document: list the output blocks (after the scans of n_mb and scratch sizes)
__global__ void k_exp_list(const DocInfo* __restrict__ docs, u32 n_docs, ExportTables t, XBlock* __restrict__ xb) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const DocInfo& di = docs[d];
    const XDoc& x = t.xdoc[d];
    if (di.code != DOC_OK || x.n_mb == 0) return;
    u32 ch_base = (u32)di.ch0;
    u64 ob = x.ob0;
    u32 per = 2 * (di.P + di.K + di.C);
    u32 open = 0xFFFFFFFFu;
    u32 idx = 0;
    for (u32 k = ch_base; k < ch_base + x.n_mc; k++) {
        if (t.mb_first[k]) {
            if (open != 0xFFFFFFFFu) { xb[ob + idx].mc1 = k; idx++; }
            XBlock b;
            b.doc = d; b.mc0 = k; b.mc1 = k; b.len = 0; b.off = 0;
            for (int i = 0; i < 8; i++) { b.sec_len[i] = 0; b.col_len[i] = 0; }
            b.scratch = x.scratch0 + (u64)idx * per;
            xb[ob + idx] = b;
            open = k;
        }
    }
    if (open != 0xFFFFFFFFu) xb[ob + idx].mc1 = ch_base + x.n_mc;
}

// ---------------------------------------------------------------------------------------------- column encoders
// All work on an index range with a value functor, so that a literal segment's length is known before its values
// are written (serde_columnar AnyRle state machine: maximal runs of >= 2 equal values become runs, the values
// between them literal segments; a lone value is a literal of one).
template <class F, class W>
__device__ inline void enc_anyrle(XSink& s, u32 n, F val, W wr) {
    u32 i = 0;
    while (i < n) {
        i64 v = val(i);
        u32 j = i;
        while (j + 1 < n && val(j + 1) == v) j++;
        if (j > i) {
            s.zigzag((i64)(j - i + 1));
            wr(s, v);
            i = j + 1;
            continue;
        }
        u32 k = i;
        i64 cur = v;
        while (k < n) {
            if (k + 1 < n) {
                i64 nx = val(k + 1);
                if (nx == cur) break;
                cur = nx;
            }
            k++;
        }
        s.zigzag(-(i64)(k - i));
        for (u32 q = i; q < k; q++) wr(s, val(q));
        i = k;
    }
}
struct WrVarint { __device__ void operator()(XSink& s, i64 v) const { s.varint((u64)v); } };
struct WrByte { __device__ void operator()(XSink& s, i64 v) const { s.put((u8)v); } };
struct WrZigzag { __device__ void operator()(XSink& s, i64 v) const { s.zigzag(v); } };
template <class F>
__device__ inline void enc_deltarle(XSink& s, u32 n, F val) {   // AnyRle over the deltas from 0
    enc_anyrle(s, n, [&](u32 i) -> i64 { i64 prev = i ? val(i - 1) : 0; return val(i) - prev; }, WrZigzag());
}
template <class F>
__device__ inline void enc_boolrle(XSink& s, u32 n, F val) {
    if (n == 0) return;
    bool state = false;
    u64 run = 0;
    for (u32 i = 0; i < n; i++) {
        bool b = val(i);
        if (b == state) run++;
        else { s.varint(run); state = !state; run = 1; }
    }
    s.varint(run);
}
struct XBits {   // MSB-first bit packer on top of a sink
    XSink& s;
    u32 cur;
    int nbits;
    __device__ XBits(XSink& s_) : s(s_), cur(0), nbits(0) {}
    __device__ void bit(bool b) {
        cur = (cur << 1) | (b ? 1u : 0u);
        if (++nbits == 8) { s.put((u8)cur); cur = 0; nbits = 0; }
    }
    __device__ void bits(u64 v, int n) { for (int i = n - 1; i >= 0; i--) bit((v >> i) & 1); }
};
// DeltaOfDelta (docs/encoding.md:1126-1172): Option<i64> first, u8 bits used in the last byte, prefix codes
template <class F>
__device__ inline void enc_dod(XSink& s, u32 n, F val) {
    if (n == 0) { s.put(0); s.put(0); return; }
    s.put(1);
    s.zigzag(val(0));
    if (n == 1) { s.put(0); return; }
    // the "bits used" byte precedes the packed bits: count them first
    u64 total_bits = 0;
    {
        i64 prev_delta = 0;
        for (u32 i = 1; i < n; i++) {
            i64 dl = val(i) - val(i - 1);
            i64 x = dl - prev_delta;
            prev_delta = dl;
            if (x == 0) total_bits += 1;
            else if (x >= -63 && x <= 64) total_bits += 9;
            else if (x >= -255 && x <= 256) total_bits += 12;
            else if (x >= -2047 && x <= 2048) total_bits += 16;
            else if (x >= -1048575 && x <= 1048576) total_bits += 26;
            else total_bits += 69;
        }
    }
    int used = (int)(total_bits & 7);
    s.put((u8)(used == 0 ? 8 : used));
    XBits bw(s);
    i64 prev_delta = 0;
    for (u32 i = 1; i < n; i++) {
        i64 dl = val(i) - val(i - 1);
        i64 x = dl - prev_delta;
        prev_delta = dl;
        if (x == 0) bw.bit(false);
        else if (x >= -63 && x <= 64) { bw.bits(2, 2); bw.bits((u64)(x + 63), 7); }
        else if (x >= -255 && x <= 256) { bw.bits(6, 3); bw.bits((u64)(x + 255), 9); }
        else if (x >= -2047 && x <= 2048) { bw.bits(14, 4); bw.bits((u64)(x + 2047), 12); }
        else if (x >= -1048575 && x <= 1048576) { bw.bits(30, 5); bw.bits((u64)(x + 1048575), 21); }
        else { bw.bits(31, 5); bw.bits((u64)x, 64); }
    }
    if (bw.nbits) s.put((u8)((bw.cur & 0xFF) << (8 - bw.nbits)));
}

// ---------------------------------------------------------------------------------------------- X4/X6: encode
// First-use registers of one block (encoding/value_register.rs): order lists + inverse maps in scratch.
struct XReg {
    u32* ord; u32* inv; u32 n;
    __device__ u32 reg(u32 v) {
        if (inv[v] != 0xFFFFFFFFu) return inv[v];
        inv[v] = n;
        ord[n] = v;
        return n++;
    }
};
// cross-peer deps of the block's changes as one flat sequence (cursor: accesses are almost monotonic)
struct XDeps {
    const ExportTables& t; u32 mc0, N; u32 j; u32 base;
    __device__ XDeps(const ExportTables& t_, u32 mc0_, u32 N_) : t(t_), mc0(mc0_), N(N_), j(0), base(0) {}
    __device__ u64 at(u32 i) {   // index of flat dep i in the dep tables
        if (i < base) { j = 0; base = 0; }
        while (j < N) {
            u32 src = t.mc_src[mc0 + j];
            u32 nd = t.ch_ndeps[src];
            if (i < base + nd) return t.ch_dep0[src] + (i - base);
            base += nd;
            j++;
        }
        return 0;
    }
};
__device__ __forceinline__ u8 xk_value_type(u8 xk) {
    switch (xk) {
        case XK_LIST: case XK_MAPSET: return VK_LORO_VALUE;
        case XK_TEXT: return VK_STR;
        case XK_DEL: return VK_DELETE_SEQ;
        default: return VK_DELETE_ONCE;
    }
}

// thread per output block.  pass 0: registers + section sizes ; pass 1: bytes.
__global__ void k_exp_encode(const DocInfo* __restrict__ docs, u64 n_blocks, ExportTables t, XBlock* __restrict__ xb,
                             u32* __restrict__ scratch, u8* __restrict__ out, int pass) {
    u64 bi_ = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (bi_ >= n_blocks) return;
    XBlock B = xb[bi_];
    const DocInfo& di = docs[B.doc];
    const u32 P = di.P, K = di.K, C = di.C;
    u32* sc = scratch + B.scratch;
    XReg peers, keys, cids;
    peers.ord = sc; peers.inv = sc + P;
    keys.ord = sc + 2 * P; keys.inv = keys.ord + K;
    cids.ord = keys.ord + 2 * K; cids.inv = cids.ord + C;
    const u32 mc0 = B.mc0, N = B.mc1 - B.mc0;
    const u32 o0 = t.mc_o0[mc0];
    const u32 o1 = t.mc_o0[B.mc1 - 1] + t.mc_no[B.mc1 - 1];
    const u32 n_ops = o1 - o0;
    const u32 first_src = t.mc_src[mc0], last_src = t.mc_src[B.mc1 - 1];
    u32 n_dep = 0;
    for (u32 j = 0; j < N; j++) n_dep += t.ch_ndeps[t.mc_src[mc0 + j]];
    u32 n_del = 0;
    if (pass == 0) {
        for (u32 i = 0; i < P; i++) peers.inv[i] = 0xFFFFFFFFu;
        for (u32 i = 0; i < K; i++) keys.inv[i] = 0xFFFFFFFFu;
        for (u32 i = 0; i < C; i++) cids.inv[i] = 0xFFFFFFFFu;
        peers.n = keys.n = cids.n = 0;
    }
    // peer of the block: the author of its changes
    u32 block_peer;
    {
        const BlockInfo& sb = t.blocks[t.ch_block[first_src]];
        block_peer = t.peer_map[sb.peer0];
    }
    if (pass == 0) {
        peers.reg(block_peer);
        // ops in order: containers, map keys, delete targets (block_encode.rs:180-236); values are made block-local
        for (u32 i = o0; i < o1; i++) {
            u8 xk = t.mo_xk[i];
            t.mo_cidx[i] = cids.reg(t.mo_cidx[i]);
            if (xk == XK_MAPSET || xk == XK_MAPDEL) t.mo_prop[i] = (i32)keys.reg((u32)t.mo_prop[i]);
            else if (xk == XK_DEL) { t.mo_f0[i] = peers.reg(t.mo_f0[i]); n_del++; }
        }
        // ContainerArena::from_containers (arena.rs:103-147): roots register their name, normals their peer
        for (u32 i = 0; i < cids.n; i++) {
            const DocContainer& dc = t.dcont[di.cid0 + cids.ord[i]];
            if (dc.is_root) keys.reg(dc.key_or_peer); else peers.reg(dc.key_or_peer);
        }
        // encode_changes (block_meta_encode.rs:13-88): dependency peers
        for (u32 j = 0; j < N; j++) {
            u32 src = t.mc_src[mc0 + j];
            const BlockInfo& sb = t.blocks[t.ch_block[src]];
            for (u32 k = 0; k < t.ch_ndeps[src]; k++) peers.reg(t.peer_map[sb.peer0 + t.dep_peer_idx[t.ch_dep0[src] + k]]);
        }
        B.col_len[7] = peers.n | (keys.n << 16);   // register sizes for pass 1 (cids.n lives in sec_len scratch below)
    } else {
        peers.n = B.col_len[7] & 0xFFFFu;
        keys.n = B.col_len[7] >> 16;
        for (u32 i = o0; i < o1; i++) n_del += t.mo_xk[i] == XK_DEL;
    }
    // number of containers: recount from the inverse map in pass 1
    if (pass == 1) { u32 n = 0; for (u32 i = 0; i < C; i++) n += cids.inv[i] != 0xFFFFFFFFu; cids.n = n; }

    // ------------------------------------------------------------------ section writers (count or write)
    auto dep_local = [&](XDeps& dc, u32 i) -> i64 {
        u64 di_ = dc.at(i);
        // source block of the owning change: dc.j is positioned on it after at()
        const BlockInfo& sb = t.blocks[t.ch_block[t.mc_src[mc0 + dc.j]]];
        return (i64)peers.inv[t.peer_map[sb.peer0 + t.dep_peer_idx[di_]]];
    };
    auto w_header = [&](XSink& s) {
        s.varint(peers.n);
        for (u32 i = 0; i < peers.n; i++) {
            u64 id = t.dpeer[di.peer0 + peers.ord[i]].id;
            for (int k = 0; k < 8; k++) s.put((u8)(id >> (8 * k)));
        }
        for (u32 j = 0; j + 1 < N; j++) s.varint(t.mc_atoms[mc0 + j]);
        enc_boolrle(s, N, [&](u32 j) { return t.ch_dep_self[t.mc_src[mc0 + j]] != 0; });
        enc_anyrle(s, N, [&](u32 j) -> i64 { return (i64)t.ch_ndeps[t.mc_src[mc0 + j]]; }, WrVarint());
        { XDeps dc(t, mc0, N); enc_anyrle(s, n_dep, [&](u32 i) -> i64 { return dep_local(dc, i); }, WrVarint()); }
        { XDeps dc(t, mc0, N); enc_dod(s, n_dep, [&](u32 i) -> i64 { return (i64)t.dep_counter[dc.at(i)]; }); }
        enc_dod(s, N - 1, [&](u32 j) -> i64 { return (i64)t.ch_lamport[t.mc_src[mc0 + j]]; });
    };
    auto w_meta = [&](XSink& s) {
        enc_dod(s, N, [&](u32 j) -> i64 { return t.ch_ts[t.mc_src[mc0 + j]]; });
        enc_anyrle(s, N, [&](u32 j) -> i64 { return (i64)t.ch_msg_len[t.mc_src[mc0 + j]]; }, WrVarint());
        for (u32 j = 0; j < N; j++) {
            u32 src = t.mc_src[mc0 + j];
            s.copy(t.bytes + t.ch_msg_off[src], t.ch_msg_len[src]);
        }
    };
    auto w_cids = [&](XSink& s) {
        s.varint(cids.n);
        for (u32 i = 0; i < cids.n; i++) {
            const DocContainer& dc = t.dcont[di.cid0 + cids.ord[i]];
            s.varint(4);
            s.put(dc.is_root ? 1 : 0);
            s.put(dc.type);
            if (dc.is_root) { s.varint(0); s.zigzag((i64)keys.inv[dc.key_or_peer]); }
            else { s.varint(peers.inv[dc.key_or_peer]); s.zigzag((i64)dc.counter); }
        }
    };
    auto w_keys = [&](XSink& s) {
        for (u32 i = 0; i < keys.n; i++) {
            u32 k = keys.ord[i];
            s.varint(t.dkey_len[di.key0 + k]);
            s.copy(t.bytes + t.dkey_off[di.key0 + k], t.dkey_len[di.key0 + k]);
        }
    };
    auto w_opcol = [&](XSink& s, int col) {
        switch (col) {
            case 0: enc_deltarle(s, n_ops, [&](u32 i) -> i64 { return (i64)t.mo_cidx[o0 + i]; }); break;
            case 1: enc_deltarle(s, n_ops, [&](u32 i) -> i64 { return (i64)t.mo_prop[o0 + i]; }); break;
            case 2: enc_anyrle(s, n_ops, [&](u32 i) -> i64 { return (i64)xk_value_type(t.mo_xk[o0 + i]); }, WrByte()); break;
            default: enc_anyrle(s, n_ops, [&](u32 i) -> i64 { return (i64)t.mo_atoms[o0 + i]; }, WrVarint());
        }
    };
    // delete rows as a flat sequence: cursor over the ops
    u32 dcur_i = 0, dcur_op = o0;   // dcur_op = op index of delete number dcur_i (when valid)
    bool dcur_valid = false;
    auto del_op = [&](u32 i) -> u32 {
        if (!dcur_valid || i < dcur_i) { dcur_i = 0; dcur_op = o0; while (t.mo_xk[dcur_op] != XK_DEL) dcur_op++; dcur_valid = true; }
        while (dcur_i < i) { dcur_op++; while (t.mo_xk[dcur_op] != XK_DEL) dcur_op++; dcur_i++; }
        return dcur_op;
    };
    auto w_delcol = [&](XSink& s, int col) {
        dcur_valid = false;
        switch (col) {
            case 0: enc_deltarle(s, n_del, [&](u32 i) -> i64 { return (i64)t.mo_f0[del_op(i)]; }); break;
            case 1: enc_deltarle(s, n_del, [&](u32 i) -> i64 { return (i64)(i32)t.mo_f1[del_op(i)]; }); break;
            default: enc_deltarle(s, n_del, [&](u32 i) -> i64 { return (i64)t.mo_f2[del_op(i)]; });
        }
    };
    auto w_values = [&](XSink& s) {
        for (u32 i = o0; i < o1; i++) {
            u8 xk = t.mo_xk[i];
            u32 st0 = t.mo_st0[i], nst = t.mo_nst[i];
            if (xk == XK_LIST) {
                s.put(7);
                s.varint(t.mo_atoms[i]);
                for (u32 q = 0; q < nst; q++) {
                    u32 row = t.st_row[st0 + q];
                    Cur c(t.bytes + t.op_val_off[row], t.op_val_len[row]);
                    (void)c.get();
                    (void)c.varint();
                    s.copy(c.p, c.left());
                }
            } else if (xk == XK_TEXT) {
                s.varint(t.mo_f1[i] - t.mo_f0[i]);
                for (u32 q = 0; q < nst; q++) {
                    u32 row = t.st_row[st0 + q];
                    Cur c(t.bytes + t.op_val_off[row], t.op_val_len[row]);
                    (void)c.varint();
                    s.copy(c.p, c.left());
                }
            } else if (xk == XK_MAPSET) {
                u32 row = t.st_row[st0];
                s.copy(t.bytes + t.op_val_off[row], t.op_val_len[row]);
            }
        }
    };

    if (pass == 0) {
        XSink s;
        s.dst = nullptr;
        s.n = 0; w_header(s); B.sec_len[0] = (u32)s.n;
        s.n = 0; w_meta(s); B.sec_len[1] = (u32)s.n;
        s.n = 0; w_cids(s); B.sec_len[2] = (u32)s.n;
        s.n = 0; w_keys(s); B.sec_len[3] = (u32)s.n;
        B.sec_len[4] = 0;
        u32 tot = 2;   // varint(1) varint(4)
        for (int c = 0; c < 4; c++) { s.n = 0; w_opcol(s, c); B.col_len[c] = (u32)s.n; tot += varint_len(s.n) + (u32)s.n; }
        B.sec_len[5] = tot;
        if (n_del) {
            tot = 2;
            for (int c = 0; c < 3; c++) { s.n = 0; w_delcol(s, c); B.col_len[4 + c] = (u32)s.n; tot += varint_len(s.n) + (u32)s.n; }
            B.sec_len[6] = tot;
        } else B.sec_len[6] = 0;
        s.n = 0; w_values(s); B.sec_len[7] = (u32)s.n;
        u32 counter_len = 0;
        for (u32 j = 0; j < N; j++) counter_len += t.mc_atoms[mc0 + j];
        u32 lam0 = t.ch_lamport[first_src];
        u32 lam_len = t.ch_lamport[last_src] + t.mc_atoms[B.mc1 - 1] - lam0;
        u32 len = varint_len((u32)t.ch_counter[first_src]) + varint_len(counter_len) + varint_len(lam0) + varint_len(lam_len) + varint_len(N);
        for (int i = 0; i < 8; i++) len += varint_len(B.sec_len[i]) + B.sec_len[i];
        B.len = len;
        xb[bi_] = B;
        return;
    }
    // ---- pass 1: ULEB length prefix + block bytes at the document's slot
    u64 base = t.xdoc[B.doc].exp_off + B.off;
    XSink s;
    s.dst = out + base - varint_len(B.len);
    s.n = 0;
    s.varint(B.len);
    u32 counter_len = 0;
    for (u32 j = 0; j < N; j++) counter_len += t.mc_atoms[mc0 + j];
    u32 lam0 = t.ch_lamport[first_src];
    u32 lam_len = t.ch_lamport[last_src] + t.mc_atoms[B.mc1 - 1] - lam0;
    s.varint((u32)t.ch_counter[first_src]);
    s.varint(counter_len);
    s.varint(lam0);
    s.varint(lam_len);
    s.varint(N);
    s.varint(B.sec_len[0]); w_header(s);
    s.varint(B.sec_len[1]); w_meta(s);
    s.varint(B.sec_len[2]); w_cids(s);
    s.varint(B.sec_len[3]); w_keys(s);
    s.varint(0);
    s.varint(B.sec_len[5]);
    s.varint(1); s.varint(4);
    for (int c = 0; c < 4; c++) { s.varint(B.col_len[c]); w_opcol(s, c); }
    s.varint(B.sec_len[6]);
    if (n_del) {
        s.varint(1); s.varint(3);
        for (int c = 0; c < 3; c++) { s.varint(B.col_len[4 + c]); w_delcol(s, c); }
    }
    s.varint(B.sec_len[7]); w_values(s);
}

// thread per document: block offsets inside the blob, blob length (after encode pass 0)
__global__ void k_exp_layout(const DocInfo* __restrict__ docs, u32 n_docs, ExportTables t, XBlock* __restrict__ xb,
                             u32* __restrict__ padded_len) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    XDoc& x = t.xdoc[d];
    u32 len = 0;
    if (docs[d].code == DOC_OK && !(x.flags & 1)) {
        len = 22;
        for (u32 i = 0; i < x.n_mb; i++) {
            XBlock& b = xb[x.ob0 + i];
            len += varint_len(b.len);
            b.off = len;
            len += b.len;
        }
    }
    x.exp_len = len;
    padded_len[d] = (len + 15u) & ~15u;
}

// thread per document: header, mode, checksum (encoding.rs:397-416)
__global__ void k_exp_finish(const DocInfo* __restrict__ docs, u32 n_docs, ExportTables t, u8* __restrict__ out) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const XDoc& x = t.xdoc[d];
    if (x.exp_len == 0) return;
    u8* b = out + x.exp_off;
    b[0] = 'l'; b[1] = 'o'; b[2] = 'r'; b[3] = 'o';
    for (int i = 4; i < 20; i++) b[i] = 0;
    b[20] = 0; b[21] = 4;   // FastUpdates, big endian
    u32 h = xxh32_dev(b + 20, x.exp_len - 20, XX_SEED_LORO);
    b[16] = (u8)h; b[17] = (u8)(h >> 8); b[18] = (u8)(h >> 16); b[19] = (u8)(h >> 24);
}
