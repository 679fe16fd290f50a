// loro_b200 -- table layouts shared by host orchestration and kernels (all tables live in HBM).
//
// Vocabulary follows the reference: blob (one exported update), block (a ~4 KB change block),
// change, op row (one run-merged op = one row of the block's `ops` columns), atom (one counter unit).
#pragma once
#include "lb_dev.cuh"

// One decoded change block (reference: block_encode.rs:95-119 EncodedBlock).
struct BlockInfo {
    u32 doc;
    u32 err;            // DOC_* code raised while parsing this block
    u32 blob_rank;      // index of the block's blob among the blobs of its document (import order)
    u32 pad_rank;
    u64 off;            // byte offset of the block inside the batch byte buffer
    u32 len;
    u32 counter_start, counter_len, lamport_start, lamport_len, n_changes;
    u32 sec_off[8];     // header, change_meta, cids, keys, positions, ops, delete_start_ids, values
    u32 sec_len[8];     //   (offsets relative to `off`)
    u32 n_peers, n_keys, n_cids, n_ops, n_dels, n_deps;
    u32 values_bytes;
    u32 n_value_maps;   // LoroValue::Map levels inside the values section (their keys are block-local indices)
    u32 n_pos;          // fractional indexes in the positions arena (encoding/arena.rs:159-233)
    u32 pos_bytes;      // their total size once the common prefixes are expanded
    u32 n_tree;         // RawTreeMove rows (encoding/value.rs:969-989)
    // exclusive-scan bases into the batch-wide tables (filled by the host after the count pass)
    u64 peer0, key0, cid0, ch0, dep0, op0, del0, pos0, posb0, tr0;
};

// One document of the batch (one blob).
struct DocInfo {
    u32 code;           // DOC_* result
    u32 b0, b1;         // block range [b0,b1)
    u32 P, C, K;        // distinct peers / containers / keys (after resolve)
    u32 P_cap, C_cap, K_cap;  // upper bounds used for allocation (sums over blocks)
    u64 peer0, cid0, key0;    // bases into doc-level peer / container / key tables (capacity-sized)
    u64 ch0;            // first change (batch-wide change index) ; changes of a doc are contiguous
    u32 n_changes;
    u32 n_applied;      // changes applied (not pending)
    u64 op0;            // first op row
    u64 n_ops;
    u64 vv0;            // base into ch_vv (n_changes * P entries)
    u64 atom0;          // base into atom_row / atom_sid (sum over peers of imported counter range)
    u64 atom_total;
    u64 mapslot0;       // base into map LWW slots (C * K)
    u64 span0;          // base into the doc's span pool
    u32 span_cap;
    u32 n_spans;
    u64 atom_ops;       // merged atoms
    u32 n_pending;      // pending changes
    u32 n_deps;         // cross-peer deps of all changes (capacity estimate)
    u32 n_blobs;        // blobs imported into this document (import_batch)
    u32 has_unsupported;
    u32 has_tree;       // any applied movable-tree op (k_tree.cuh)
    u32 n_prior;        // the first n_prior blobs are the document's EARLIER state (lb_docset_import): the import
                        // status reports what the remaining blobs added to it
    u64 tree0;          // base into the per-document tree node tables (atom_total + C slots)
    u64 json_off;
    u32 json_len;
    u32 pad;
};

// doc-level peer entry
struct DocPeer {
    u64 id;
    u32 rank;          // rank of id among the doc's peers (ascending)
    i32 succ_lo;       // ImportStatus.success = [succ_lo, end_counter) when has_succ (first counter the import added)
    i32 end_counter;   // vv after import (exclusive)
    u32 n_app;         // applied copies of this peer's changes (the first n_app entries of its ch_aorder list)
    u32 atom_base;     // offset of this peer's atoms inside the doc's atom arrays
    u32 ch_first;      // index into doc_change_order of this peer's first change
    u32 ch_count;
    i32 pend_lo, pend_hi;  // pending counter range (lo<hi when any)
    u32 is_head;           // the peer's last imported id is a frontier of the document (version/frontiers.rs:233-246)
    u32 has_succ;          // some change of this peer was applied by the (new) blobs of this import
};

// doc-level container entry (reference: ContainerID, loro-common/src/lib.rs:114-180)
struct DocContainer {
    u8 is_root, type, pad0, pad1;
    u32 name_len;
    u64 name_off;      // root: name bytes inside the batch buffer
    u64 peer;          // normal
    i32 counter;
    u32 n_ins_rows, n_del_rows, n_ins_atoms, n_map_rows;
    // sequence tracker pools (k_seq)
    u64 leaf0;  u32 leaf_cap;  u32 n_leaves;
    u64 node0;  u32 node_cap;  u32 n_nodes;
    u32 root, height, first_leaf;
    u64 out0;   u32 out_cap;   u32 n_out;   // final visible runs (row, off, len)
    u32 seq_len;        // visible atoms
    u32 unk_sid;        // span id of the tracker's placeholder span
    u32 key_or_peer;    // root: doc-level key index of the name ; normal: doc-level peer index of the creator
    u32 unsupported;
    u64 cvv0;           // tracker current_vv (P entries) base
};
