/* loro_b200 -- C ABI of the B200-native batched CRDT merge engine.
 *
 * Drop-in boundary for ONE hot path of loro-dev/loro: batched `LoroDoc::import` / `import_batch` of FastUpdates blobs
 * into fresh documents (lb_import_batch*) or into documents that already hold history (lb_docset_*) -> (decode, causal
 * scan, eg-walker merge of List / Text / Map / Tree) -> deep JSON state / import status / version vector / frontiers
 * -> re-export of every document (`export(ExportMode::all_updates)`, `export(updates(from))` on demand).
 * The reference has no C FFI for this path; every entry point cites the Rust interface it
 * replaces (paths relative to /root/reference):
 *
 *   lb_import_batch          crates/loro/src/lib.rs:639  LoroDoc::import(&self, &[u8]) -> Result<ImportStatus>
 *                            crates/loro/src/lib.rs:425  LoroDoc::import_batch (blobs sharing a doc_id -> one document)
 *                            crates/loro-internal/src/loro.rs:562-643 (header/checksum/mode checks first)
 *   lb_doc_status            crates/loro-internal/src/encoding.rs:226-230 ImportStatus{success,pending}
 *                            crates/loro-common/src/error.rs:8-105 (LoroError variants -> lb_doc_code)
 *   lb_doc_json              crates/loro/src/lib.rs:866  LoroDoc::get_deep_value() (serde_json text, keys sorted)
 *   lb_doc_vv                crates/loro/src/lib.rs:816  LoroDoc::oplog_vv()
 *   lb_doc_frontiers         crates/loro/src/lib.rs:881  LoroDoc::oplog_frontiers()
 *   lb_doc_export_updates    crates/loro/src/lib.rs:1235 LoroDoc::export(ExportMode::all_updates() / updates(from))
 *                            crates/loro-internal/src/encoding.rs:79-83, 350-416, oplog/change_store.rs:494-576
 *   lb_docset_import         crates/loro/src/lib.rs:639, :425 on a document that already holds history
 *                            (crates/loro-internal/src/loro.rs:562-643, 1183-1290, oplog.rs:130-196)
 *   lb_batch_counters        crates/loro-internal/src/loro.rs:1458 len_ops / len_changes (summed over the batch)
 *
 * Conventions (mirroring the reference): input buffers are borrowed for the duration of the call only;
 * outputs are owned by the batch handle until lb_batch_free; a bad blob never aborts the batch -- it yields a
 * per-document error code; checksum / mode are verified before anything else; missing dependencies are not
 * errors (they are reported as `pending`).  Thread-safe for distinct handles.
 *
 * All compute runs in hand-written CUDA kernels (sm_100a).  There is no CPU fallback: without a CUDA device
 * every entry point returns LB_ERR_NO_DEVICE.
 */
#ifndef LORO_B200_H
#define LORO_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum lb_status {
    LB_OK = 0,
    LB_ERR_INVALID_ARG = 1,
    LB_ERR_NO_DEVICE = 2,
    LB_ERR_CUDA = 3,
    LB_ERR_OOM = 4,
    LB_ERR_INTERNAL = 5,
    LB_ERR_UNSUPPORTED = 6   /* valid request the engine does not cover yet (see lb_last_error) */
} lb_status;

/* per-document result code; the LoroError variant it corresponds to is given on the right */
typedef enum lb_doc_code {
    LB_DOC_OK = 0,
    LB_DOC_ERR_DECODE = 1,          /* LoroError::DecodeError (short blob, bad magic, malformed block)  */
    LB_DOC_ERR_CHECKSUM = 2,        /* LoroError::DecodeChecksumMismatchError                            */
    LB_DOC_ERR_MODE = 3,            /* IncompatibleFutureEncodingError / ImportUnsupportedEncodingMode   */
                                    /* (a mode other than FastSnapshot 3 / FastUpdates 4)                */
    LB_DOC_ERR_CORRUPT = 4,         /* LoroError::DecodeDataCorruptionError                               */
    LB_DOC_ERR_UNSUPPORTED = 5,     /* well-formed, but outside this path: an intact FastSnapshot blob    */
                                    /* (mode 3), or ops the engine does not merge yet        */
                                    /* (rich-text styles, movable list, counter)                          */
    LB_DOC_ERR_CAPACITY = 6         /* internal capacity bound exceeded (engine bug or adversarial input) */
} lb_doc_code;

typedef struct lb_blob {
    const uint8_t* ptr; /* host pointer, borrowed */
    size_t len;
    uint64_t doc_id;    /* blobs with the same doc_id are imported into ONE document (LoroDoc::import_batch,
                           crates/loro/src/lib.rs:425); documents are numbered by first appearance */
} lb_blob;

typedef struct lb_options {
    int device;          /* CUDA device ordinal */
    uint32_t flags;      /* LB_FLAG_* */
    uint32_t reserved[6];
} lb_options;
#define LB_FLAG_NO_JSON 1u      /* skip deep-value JSON materialisation */
#define LB_FLAG_KEEP_DEVICE 2u  /* keep intermediate device tables for lb_debug_* (tests) */
#define LB_FLAG_EXPORT 4u       /* also re-export every document (phase 7) for lb_doc_export_updates */
#define LB_FLAG_COMPACT 8u      /* lb_docset_import: afterwards keep each touched document as its own export (see below) */

typedef struct lb_id_span {
    uint64_t peer;
    int32_t start; /* inclusive counter */
    int32_t end;   /* exclusive counter */
} lb_id_span;

typedef struct lb_import_status {
    lb_doc_code code;
    size_t n_success;
    const lb_id_span* success; /* ImportStatus.success (VersionRange) */
    size_t n_pending;
    const lb_id_span* pending; /* ImportStatus.pending */
} lb_import_status;

typedef struct lb_counters {
    uint64_t docs, docs_ok;
    uint64_t blob_bytes;
    uint64_t blocks, changes, op_rows; /* decoded */
    uint64_t atom_ops;                 /* merged ops (sum of change.atom_len of applied changes) */
    uint64_t pending_changes;
    uint64_t json_bytes;
    uint64_t state_hash;               /* xor over docs of (xxh32(json) << 32 | len): order independent */
} lb_counters;

typedef struct lb_timings { /* device time per phase in milliseconds (CUDA events on the batch stream) */
    float h2d, frame, decode, resolve, classify, integrate, materialise, d2h, total_device, reexport;
    /* algorithmic bytes of the decode phase: blob bytes read + SoA bytes written */
    uint64_t decode_bytes_read, decode_bytes_written;
    uint32_t kernel_launches;
    uint64_t export_bytes;             /* bytes written by the re-export phase */
    float tree;                        /* movable-tree phase (sort, apply, sibling lists) */
    uint32_t reserved0;
    uint64_t tree_ops;                 /* RawTreeMove rows decoded */
    /* change blocks by decode path: lane-parallel rows / staged in shared memory with the rows on one lane /
     * larger than the staging buffer (k_decode_warp.cuh) */
    uint64_t decode_fast_blocks, decode_lane_blocks, decode_unstaged_blocks;
    /* host milliseconds spent inside the device allocator while the batch was built (they are part of the wall time of
     * a step but of no device phase) and the bytes it handed out */
    float alloc_host_ms;
    uint32_t reserved1;
    uint64_t device_bytes;
    /* host wall time of the whole import call, and of its tail: from the moment the last kernel was enqueued (results
     * download, status tables) -- what a step costs beyond `total_device` */
    float host_call_ms, host_tail_ms;
} lb_timings;

typedef struct lb_batch lb_batch;

/* Import a batch of FastUpdates blobs from HOST memory: one fresh document per distinct doc_id (documents are
 * numbered in order of first appearance; lb_doc_count tells how many there are). */
lb_status lb_import_batch(const lb_blob* blobs, size_t n_blobs, const lb_options* opt, lb_batch** out);

/* Same, with the blobs already resident in device memory: `d_bytes` is one device buffer holding all blobs,
 * blob i occupying [offsets[i], offsets[i] + lens[i]) (HOST arrays of n_docs entries; every offset a multiple
 * of 16).  Nothing is copied host->device except these two small arrays. */
lb_status lb_import_batch_device(const uint8_t* d_bytes, const uint64_t* offsets, const uint32_t* lens,
                                 size_t n_docs, const lb_options* opt, lb_batch** out);

size_t lb_doc_count(const lb_batch* b);
lb_status lb_doc_status(const lb_batch* b, size_t doc, lb_import_status* out);
lb_status lb_doc_json(const lb_batch* b, size_t doc, const char** utf8, size_t* len);
lb_status lb_doc_vv(const lb_batch* b, size_t doc, const lb_id_span** spans, size_t* n); /* start=0,end=vv[peer] */
/* LoroDoc::oplog_frontiers() (crates/loro/src/lib.rs:881; version/frontiers.rs:233-246): the heads of the causal graph,
 * one span [counter, counter + 1) per head id. */
lb_status lb_doc_frontiers(const lb_batch* b, size_t doc, const lb_id_span** spans, size_t* n);
/* LoroDoc::export(ExportMode::updates(from)) of document `doc` (crates/loro/src/lib.rs:1235, encoding.rs:79-83, 350-416,
 * oplog/change_store.rs:494-528): the FastUpdates blob a fresh reference document would export after importing the same
 * input.  `from` = NULL / n_from = 0 is ExportMode::all_updates() (computed for every document at import time);
 * otherwise `from` is a version vector (one span per peer, `end` = the first counter the receiver lacks; peers not
 * listed start at 0) and the stored changes are cut there (Change::slice) -- computed on demand, the returned buffer
 * stays valid until the next from-export of the same document or lb_batch_free.  Needs LB_FLAG_EXPORT at import time.
 * LB_ERR_UNSUPPORTED: the document uses something the export phase does not cover (lb_last_error). */
lb_status lb_doc_export_updates(const lb_batch* b, size_t doc, const lb_id_span* from, size_t n_from,
                                const uint8_t** bytes, size_t* len);
lb_status lb_batch_counters(const lb_batch* b, lb_counters* out);
lb_status lb_batch_timings(const lb_batch* b, lb_timings* out);
const char* lb_last_error(void); /* thread-local, human readable */
/* Device blocks that lived until lb_batch_free are kept (per device, by size class, at most LB_DEV_CACHE_GB gigabytes,
 * default 85 % of the device's memory) for the next batch of similar shape; this gives them back to the driver. */
lb_status lb_device_trim(int device);

/* One process per GPU: pin the calling thread -- and the staging / download threads the engine creates from it -- to
 * the CPUs of the NUMA node `device` is attached to (sysfs).  Call before building the input buffers so that they,
 * the pinned staging ring and the gather threads all sit next to the GPU.  LB_ERR_UNSUPPORTED: topology unknown. */
lb_status lb_numa_bind(int device);
void lb_batch_free(lb_batch* b);

/* ---- persistent documents: imports against an EXISTING document state ------------------------------------------------
 * LoroDoc::import / import_batch on a document that already holds history (crates/loro/src/lib.rs:639, :425;
 * crates/loro-internal/src/loro.rs:562-643, 1183-1290; oplog.rs:130-196: changes the document knows are skipped or
 * trimmed, pending changes wait in the oplog until a later import brings their dependencies).
 * A docset keeps, per doc_id, the document's change store in wire form IN DEVICE MEMORY: the update blobs it has
 * imported, in order (what the reference's ChangeStore keeps are encoded blocks too, change_store.rs:60-110).
 * lb_docset_import lays the stored blobs of every touched document in front of the new ones (device-to-device) and
 * replays the document; the batch it returns answers exactly like the reference's import on the existing document:
 *   lb_doc_status    ImportStatus of THIS import: success = what the new blobs added (a change the document already
 *                    held is not reported, a stored pending change released by this import is), pending = what the
 *                    new blobs parked;
 *   lb_doc_json / lb_doc_vv / lb_doc_frontiers / lb_doc_export_updates    the document after the import.
 * Several blobs with one doc_id in one call = import_batch on that document (sorted by mode, then number of changes
 * descending, among the new blobs).  A document whose import fails (checksum, decode error, ...) keeps its earlier
 * state, like the reference (loro.rs:584: checked before any state change).
 * LB_FLAG_COMPACT in `opt->flags`: after this import every touched document without pending changes is replaced by a
 * fresh document that imported its own export -- `fresh.import(doc.export(all_updates))`, the way a host drops
 * redundant history; the stored form shrinks to one blob.  State, vv and frontiers are unaffected; later exports equal
 * those of a reference document that was re-created the same way (the op segmentation of an export depends on which
 * blob brought which piece of a change, so they may differ by a few bytes from an uncompacted document's).
 * The engine merges by replaying the
 * document's whole history, not from the common ancestor of the two versions (dag.rs:488-667): same results, the
 * cost of an import grows with the history.
 * One docset serves one device; calls on the same docset are serialised.  LB_FLAG_EXPORT is implied. */
typedef struct lb_docset lb_docset;
lb_status lb_docset_new(const lb_options* opt, lb_docset** out);
lb_status lb_docset_import(lb_docset* set, const lb_blob* blobs, size_t n_blobs, const lb_options* opt, lb_batch** out);
size_t lb_docset_doc_count(const lb_docset* set);
uint64_t lb_docset_stored_bytes(const lb_docset* set);   /* device bytes of the stored documents */
void lb_docset_free(lb_docset* set);

/* test hooks (need LB_FLAG_KEEP_DEVICE): copy one decoded SoA table to the host.
 * name in {"op_cid","op_prop","op_vtype","op_len","op_counter","ch_counter","ch_len","ch_lamport",
 *          "ch_ts","dep_peer","dep_counter","blk_doc","blk_nchanges"}; returns element count. */
lb_status lb_debug_table(const lb_batch* b, const char* name, void* dst, size_t dst_bytes, size_t* n_elems,
                         size_t* elem_size);

#ifdef __cplusplus
}
#endif
#endif
