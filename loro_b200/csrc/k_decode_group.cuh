// loro_b200 -- phase 2, one warp per GROUP of four change blocks, one lane per column stream.
//
// Replaces the same reference code as k_decode.cuh (block_encode.rs:527-659, serde_columnar 0.3.14 AnyRle / DeltaRle,
// encoding/value.rs:603-700).  Shape of the work:
//   * the four blocks of a group are staged in shared memory by bulk asynchronous copies (TMA 1-D: cp.async.bulk +
//     mbarrier): HBM sees one aligned, coalesced read of every block (the thread-per-block decoder fetched 8-byte
//     windows out of 32-byte sectors through cursors 4 KB apart: 8x the algorithmic traffic, profiles/r1_ncu_decode.md)
//     and every byte the decoders look at afterwards costs a shared-memory access instead of an L1/L2/HBM round trip;
//   * a block is eight independent byte streams -- the four ops columns, the three delete-start columns and the
//     header group (peers, change lengths, deps, lamports, timestamps, keys, cids, positions) -- so eight lanes decode
//     one block at the same time and the warp's 32 lanes are busy on four blocks.  Run-length columns are sequential by
//     nature (a value's position depends on every run before it) and on real update streams the runs are short, so a
//     lane per stream beats a warp per run (k_decode_warp.cuh: same traffic, 50x the instructions);
//   * the values walk (a chain: a value starts where the previous one ends) and the counter / change assignment run
//     as a second pair of streams per block once the kind and length columns are there (kinds kept in shared memory).
// Anything the fast path does not cover -- a block larger than the staging slot, more rows than the kind buffer,
// malformed input -- is handed to the single-lane decoder (decode_block_small / decode_block_rows of k_decode.cuh),
// which rewrites the same rows and produces the precise error code, so the fast path never has to explain a failure.
#pragma once
#include "k_decode.cuh"

#define DG_G 4             // blocks per warp
#define DG_WARPS 4         // warps per CTA
#define DG_BYTES 4608      // staging slot per block (block bytes + 16-byte alignment slack)
#define DG_ROWS 1536       // op rows per block whose kinds fit the shared-memory buffer

struct DgWarp {
    alignas(16) u8 bytes[DG_G][DG_BYTES];
    u8 vt[DG_G][DG_ROWS];
    alignas(8) u64 bar;
    u32 pad_[2];
};

// one column stream of a staged block: AnyRle segments of raw bytes (mode 0), varints (1) or zig-zag deltas that are
// accumulated (2, DeltaRle); values outside [lo, hi] flag the block (and are replaced by `lo`, like the thread-per-block
// decoder does, so that later passes stay in range); kinds are written as bytes (to the table and to shared memory),
// everything else as 32-bit words.  Returns false when the column is malformed or does not hold exactly n values.
__device__ inline bool dg_stream(const u8* col, u32 len, int mode, u32 n, i64 lo, i64 hi, u32* out32, u8* out8, u8* out8s,
                                 bool* corrupt) {
    RleCur c(col, len, mode);
    i64 acc = 0;
    u32 r = 0;
    for (; r < n; r++) {
        i64 v;
        if (!c.next(&v)) break;
        if (mode == 2) { acc += v; v = acc; }
        if (v < lo || v > hi) { *corrupt = true; v = lo; }
        if (out8) { out8[r] = (u8)v; out8s[r] = (u8)v; }
        else out32[r] = (u32)v;
    }
    if (c.c.err || r != n) return false;
    i64 extra;
    return !c.next(&extra);
}

__global__ void __launch_bounds__(32 * DG_WARPS)
k_block_decode_group(const u8* __restrict__ bytes, BlockInfo* __restrict__ blocks, u64 n_blocks, Tables t) {
#ifdef LB_SIMT_EMU
    LB_DYN_SMEM(DgWarp, smem);
#else
    extern __shared__ __align__(16) u8 dg_smem_raw[];
    DgWarp* smem = (DgWarp*)dg_smem_raw;
#endif
    const int lane = threadIdx.x & 31;
    const u32 w = threadIdx.x >> 5;
    const int sb = lane >> 3, role = lane & 7;
    const u64 g0 = ((u64)blockIdx.x * DG_WARPS + w) * DG_G;
    if (g0 >= n_blocks) return;
    DgWarp& S = smem[w];
    const u64 i = g0 + (u64)sb;
    const bool valid = i < n_blocks;
    const BlockInfo& bi = blocks[valid ? i : g0];
    const bool live = valid && !bi.err;
    const u32 shift = (u32)(bi.off & 15);
    const u32 span = (shift + bi.len + 15u) & ~15u;
    const bool staged = live && span <= DG_BYTES && bi.n_ops <= DG_ROWS && (((size_t)bytes) & 15) == 0;
    // ---- stage the blocks of the group: one bulk asynchronous copy each, completion on the warp's mbarrier
#ifdef LB_SIMT_EMU
    if (staged)
        for (u32 k = (u32)role; k < span; k += 8) S.bytes[sb][k] = bytes[bi.off - shift + k];
    __syncwarp();
#else
    {
        const u32 bar = (u32)__cvta_generic_to_shared(&S.bar);
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(DG_G));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        if (role == 0) {
            if (staged) {
                const u32 dst = (u32)__cvta_generic_to_shared(S.bytes[sb]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(span) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(bytes + (bi.off - shift)), "r"(span), "r"(bar) : "memory");
            } else
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        }
        __syncwarp();
        u32 done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        }
    }
#endif
    const u8* b = S.bytes[sb] + shift;
    u8* vts = S.vt[sb];
    u32 err_small = 0, n_maps = 0;
    bool bad = false;           // this lane's stream failed: the block goes to the single-lane decoder
    bool corrupt = false;
    // ---- phase 1: eight streams per block
    if (staged) {
        if (role == 7) err_small = decode_block_small(b, bi, i, t);
        else if (role < 4) {
            const u8* col[4];
            u32 cl[4];
            if (!columnar_open(b + bi.sec_off[5], bi.sec_len[5], 4, col, cl)) bad = true;
            else {
                const u64 r0 = bi.op0;
                const u32 R = bi.n_ops;
                bool ok;
                if (role == 0) ok = dg_stream(col[0], cl[0], 2, R, 0, (i64)bi.n_cids - 1, t.op_cid + r0, nullptr, nullptr, &corrupt);
                else if (role == 1) ok = dg_stream(col[1], cl[1], 2, R, -(i64)0x80000000ll, (i64)0x7FFFFFFF, (u32*)t.op_prop + r0, nullptr, nullptr, &corrupt);
                else if (role == 2) ok = dg_stream(col[2], cl[2], 0, R, 0, 255, nullptr, t.op_vtype + r0, vts, &corrupt);
                else ok = dg_stream(col[3], cl[3], 1, R, 1, (i64)0x7FFFFFFF, t.op_len + r0, nullptr, nullptr, &corrupt);
                if (!ok) bad = true;
            }
        } else if (bi.n_dels) {
            const u8* col[3];
            u32 cl[3];
            if (!bi.sec_len[6] || !columnar_open(b + bi.sec_off[6], bi.sec_len[6], 3, col, cl)) bad = true;
            else {
                const u64 d0 = bi.del0;
                bool ok;
                if (role == 4) ok = dg_stream(col[0], cl[0], 2, bi.n_dels, 0, (i64)bi.n_peers - 1, t.del_peer_idx + d0, nullptr, nullptr, &corrupt);
                else if (role == 5) ok = dg_stream(col[1], cl[1], 2, bi.n_dels, -(i64)0x80000000ll, (i64)0x7FFFFFFF, (u32*)t.del_counter + d0, nullptr, nullptr, &corrupt);
                else {
                    // a delete span of length 0 is corrupt: [lo, hi] cannot express "anything but 0", checked below
                    ok = dg_stream(col[2], cl[2], 2, bi.n_dels, -(i64)0x80000000ll, (i64)0x7FFFFFFF, (u32*)t.del_len + d0, nullptr, nullptr, &corrupt);
                    for (u32 q = 0; ok && q < bi.n_dels; q++) if (t.del_len[d0 + q] == 0) corrupt = true;
                }
                if (!ok) bad = true;
            }
        } else if (bi.sec_len[6]) bad = true;
    }
    __syncwarp();
    unsigned badm = __ballot_sync(LB_FULL, bad || corrupt);
    err_small = (u32)__shfl_sync(LB_FULL, (int)err_small, (sb << 3) | 7);
    bool blk_bad = ((badm >> (sb << 3)) & 0xFFu) != 0 || err_small != 0;
    // ---- phase 2: the values chain (role 0) and counters / changes (role 1) of every block that is still clean
    bool p2_bad = false;
    if (staged && !blk_bad && role == 0) {
        const u64 r0 = bi.op0;
        const u32 R = bi.n_ops;
        Cur v(b + bi.sec_off[7], bi.sec_len[7]);
        u32 ndel = 0, ntree = 0;
        for (u32 r = 0; r < R; r++) {
            const u64 row = r0 + r;
            const u8 vt = vts[r];
            const u8* v0 = v.p;
            u32 aux_idx = 0xFFFFFFFFu;
            if (vt == VK_RAW_TREE_MOVE) {
                // read_raw_tree_move (value.rs:969-989): subject peer idx / counter, position idx, parent (null = root)
                u64 sp = v.varint(), sc = v.varint(), pi = v.varint();
                u8 pn = v.get();
                u64 pp = 0, pcn = 0;
                if (!pn) { pp = v.varint(); pcn = v.varint(); }
                if (ntree >= bi.n_tree || sp >= bi.n_peers || (!pn && pp >= bi.n_peers) || sc > 0x7FFFFFFFull || pcn > 0x7FFFFFFFull) { p2_bad = true; break; }
                u8 pk = pn ? TRP_ROOT : TRP_NODE;
                if (!pn && t.peer_id[bi.peer0 + (u32)pp] == DELETED_ROOT_PEER && (i32)pcn == DELETED_ROOT_CTR) pk = TRP_DELETED;
                if (pk != TRP_DELETED && pi >= bi.n_pos) { p2_bad = true; break; }
                u64 ti = bi.tr0 + ntree++;
                t.tr_target_peer[ti] = (u32)sp;
                t.tr_target_ctr[ti] = (i32)sc;
                t.tr_parent_kind[ti] = pk;
                t.tr_parent_peer[ti] = (u32)pp;
                t.tr_parent_ctr[ti] = (i32)pcn;
                t.tr_pos[ti] = pk == TRP_DELETED ? 0xFFFFFFFFu : (u32)(bi.pos0 + pi);
                aux_idx = (u32)ti;
            } else
                skip_value(v, vt, &n_maps);
            t.op_val_off[row] = bi.off + (u64)(v0 - b);
            t.op_val_len[row] = (u32)(v.p - v0);
            if (vt == VK_DELETE_SEQ) aux_idx = (u32)(bi.del0 + ndel++);
            t.op_del[row] = aux_idx;
        }
        if (v.err || !v.empty() || ndel != bi.n_dels || ntree != bi.n_tree) p2_bad = true;
    } else if (staged && !blk_bad && role == 1) {
        const u64 r0 = bi.op0;
        const u32 R = bi.n_ops, N = bi.n_changes;
        i32 counter = (i32)bi.counter_start;
        u32 change = 0, ch_first_row = 0;
        i32 next_boundary = (i32)bi.counter_start + (i32)t.ch_len[bi.ch0];
        t.ch_op0[bi.ch0] = r0;
        for (u32 r = 0; r < R; r++) {
            if (change >= N) { p2_bad = true; break; }
            t.op_counter[r0 + r] = counter;
            t.op_change[r0 + r] = (u32)(bi.ch0 + change);
            counter += (i32)t.op_len[r0 + r];
            if (counter > next_boundary) { p2_bad = true; break; }      // a row never straddles a change boundary
            if (counter == next_boundary) {
                t.ch_nops[bi.ch0 + change] = r + 1 - ch_first_row;
                change++;
                ch_first_row = r + 1;
                if (change < N) { t.ch_op0[bi.ch0 + change] = r0 + r + 1; next_boundary += (i32)t.ch_len[bi.ch0 + change]; }
            }
        }
        if (change != N || counter != (i32)(bi.counter_start + bi.counter_len)) p2_bad = true;
    }
    __syncwarp();
    badm = __ballot_sync(LB_FULL, p2_bad);
    blk_bad = blk_bad || ((badm >> (sb << 3)) & 0xFFu) != 0;
    // ---- the block leader finishes: clean blocks are done, the others are decoded again by the single-lane decoder
    if (role == 0 && live) {
        u32 err = 0;
        if (!staged) {
            const u8* g = bytes + bi.off;
            err = decode_block_small(g, bi, i, t);
            n_maps = 0;
            err = decode_block_rows_cols(g, bi, t, err, &n_maps);
            atomicAdd(&t.dw_stats[2], 1ull);
        } else if (blk_bad) {
            n_maps = 0;
            err = decode_block_rows(b, bi, t, err_small, &n_maps);
            if (!err) err = err_small;
            atomicAdd(&t.dw_stats[1], 1ull);
        } else
            atomicAdd(&t.dw_stats[0], 1ull);
        blocks[i].n_value_maps = n_maps;
        if (err) decode_block_fail(bi, i, t, blocks, err);
    }
}
