// loro_b200 -- host orchestration + C ABI (include/loro_b200.h).
//
// The host side only sizes tables, launches kernels and copies results; every byte of decode / merge /
// materialisation work happens in the kernels of k_*.cuh.  There is no CPU fallback: with no CUDA device
// the entry points fail with LB_ERR_NO_DEVICE.
#include "../../include/loro_b200.h"

#include <algorithm>
#include <chrono>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <sched.h>

#include "lb_defs.h"
#include "k_frame.cuh"
#include "k_decode.cuh"
#include "k_decode_warp.cuh"
#include "k_decode_group.cuh"
#include "k_resolve.cuh"
#include "k_classify.cuh"
#include "k_seq.cuh"
#include "k_tree.cuh"
#include "k_state.cuh"
#include "k_export.cuh"
#include "host_stage.hpp"

static thread_local std::string g_last_error;
const char* lb_last_error(void) { return g_last_error.c_str(); }

#define CK(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            g_last_error = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            throw lb_status(LB_ERR_CUDA);                                                    \
        }                                                                                    \
    } while (0)

// thread per doc helpers -------------------------------------------------------------------------
__global__ void k_doc_sizes(const DocInfo* __restrict__ docs, u32 n_docs, u32* __restrict__ vvsize,
                            u32* __restrict__ atoms, u32* __restrict__ mapslots, int which) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const DocInfo& di = docs[d];
    bool ok = di.code == DOC_OK;
    if (which == 0) vvsize[d] = ok ? di.n_changes * di.P : 0;
    else if (which == 2) {   // tree node slots (k_tree.cuh) ; atoms[n_docs] collects the largest tree document
        vvsize[d] = ok && di.has_tree ? (u32)di.atom_total + di.C : 0;
        if (ok && di.has_tree) atomicMax(&atoms[n_docs], (u32)di.atom_total);
    }
    else {
        atoms[d] = ok ? (u32)di.atom_total : 0;
        mapslots[d] = ok ? di.C * di.K : 0;
    }
}
__global__ void k_pack_peers(const DocInfo* __restrict__ docs, u32 n_docs, const DocPeer* __restrict__ dpeer,
                             const u64* __restrict__ base, DocPeer* __restrict__ out) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const DocInfo& di = docs[d];
    for (u32 p = 0; p < di.P; p++) out[base[d] + p] = dpeer[di.peer0 + p];
}
// multi-field scan: CTA f scans field f (offsets in bytes inside the strided records)
struct ScanJob { const u8* in; u8* out; size_t in_stride, out_stride; u64 n; };
struct ScanJobs { ScanJob j[8]; };
__global__ void k_excl_scan_multi(ScanJobs jobs) {
    const ScanJob& jb = jobs.j[blockIdx.x];
    const u8* in = jb.in; u8* out = jb.out; size_t in_stride = jb.in_stride, out_stride = jb.out_stride; u64 n = jb.n;
    __shared__ u64 warp_tot[32];
    __shared__ u64 carry_s;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (u64 base = 0; base < n; base += blockDim.x) {
        u64 i = base + threadIdx.x;
        u64 v = i < n ? (u64) * (const u32*)(in + i * in_stride) : 0;
        u64 s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 t = __shfl_up_sync(LB_FULL, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) warp_tot[w] = s;
        __syncthreads();
        if (w == 0) {
            u64 t = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
            u64 ts = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 u = __shfl_up_sync(LB_FULL, ts, d);
                if (lane >= d) ts += u;
            }
            warp_tot[lane] = ts - t;
        }
        __syncthreads();
        u64 carry = carry_s;
        if (i < n) *(u64*)(out + i * out_stride) = carry + warp_tot[w] + s - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_tot[w] + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) *(u64*)(out + n * out_stride) = carry_s;
}

// warp per segment: copy blobs between device buffers (lb_docset: the stored state of a document enters the next batch,
// the re-exported blobs leave the batch).  Sources and destinations are 16-byte aligned.
struct CopySeg { const u8* src; u8* dst; u64 len; };
__global__ void k_copy_segments(const CopySeg* __restrict__ segs, u32 n) {
    u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= n) return;
    CopySeg sg = segs[w];
    u64 n16 = sg.len >> 4;
    const uint4* s4 = (const uint4*)sg.src;
    uint4* d4 = (uint4*)sg.dst;
    for (u64 i = lane; i < n16; i += 32) d4[i] = s4[i];
    for (u64 i = (n16 << 4) + lane; i < sg.len; i += 32) sg.dst[i] = sg.src[i];
}

#ifndef LB_DECODE_DEFAULT
#define LB_DECODE_DEFAULT 1   // 0 rows, 1 cols, 2 warp, 3 group (k_decode*.cuh)
#endif

namespace {

// Device blocks that lived until their batch was freed are kept per device, by size class, and handed to the next batch
// that asks for the same class: a service imports batch after batch of similar shape, and the stream-ordered pool took
// up to hundreds of milliseconds (host-blocking) to find or map room for the multi-GB row tables of a step -- 68 ms per
// step on config C5.  Blocks a batch gives back early (Dev::release) still go to the pool, so their memory stays
// available to every later size.  lb_device_trim() empties the cache.
struct BlockCache {
    std::mutex mu;
    std::unordered_map<size_t, std::vector<void*>> free_;
    size_t bytes = 0;
};
BlockCache& block_cache(int device) {
    static BlockCache caches[64];
    return caches[(unsigned)device % 64];
}
size_t cache_cap_bytes() {   // LB_DEV_CACHE_GB, default 85 % of the device's memory
    static const size_t cap = [] {
        const char* e = getenv("LB_DEV_CACHE_GB");
        if (e) return (size_t)(atof(e) * 1e9);
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) tot = (size_t)64e9;
        return (size_t)((double)tot * 0.85);
    }();
    return cap;
}
inline size_t size_class(size_t sz) {   // 8 classes per power of two (at most 12.5 % above the request), 256-byte granules
    sz = (sz + 255) & ~(size_t)255;
#ifndef LB_SIMT_EMU
    if (sz > 4096) {
        int e = 63 - __builtin_clzll((unsigned long long)(sz - 1));
        size_t step = (size_t)1 << (e - 3);
        sz = (sz + step - 1) & ~(step - 1);
    }
#endif
    return sz;
}
void cache_flush(int device, cudaStream_t st) {
    BlockCache& bc = block_cache(device);
    std::lock_guard<std::mutex> g(bc.mu);
    for (auto& kv : bc.free_) for (void* p : kv.second) cudaFreeAsync(p, st);
    bc.free_.clear();
    bc.bytes = 0;
}

// Batch streams are recycled per device: memory a batch frees early goes back to the stream-ordered pool, and the pool
// serves a later request fastest when it comes from the stream the memory was freed on.
struct StreamCache {
    std::mutex mu;
    std::vector<cudaStream_t> idle;
};
StreamCache& stream_cache(int device) {
    static StreamCache caches[64];
    return caches[(unsigned)device % 64];
}
cudaStream_t stream_take(int device) {
    StreamCache& sc = stream_cache(device);
    {
        std::lock_guard<std::mutex> g(sc.mu);
        if (!sc.idle.empty()) { cudaStream_t s = sc.idle.back(); sc.idle.pop_back(); return s; }
    }
    cudaStream_t s = nullptr;
    if (cudaStreamCreate(&s) != cudaSuccess) return nullptr;
    return s;
}
void stream_give(int device, cudaStream_t s) {
    StreamCache& sc = stream_cache(device);
    std::lock_guard<std::mutex> g(sc.mu);
    if (sc.idle.size() < 4 && !getenv("LB_NO_STREAM_REUSE")) sc.idle.push_back(s);
    else cudaStreamDestroy(s);
}

struct Dev {  // owns every device allocation of a batch
    cudaStream_t stream = nullptr;
    int device = 0;
    std::vector<std::pair<void*, size_t>> ptrs;   // whole blocks in use
    // Blocks the batch is done with before it ends (the tracker pools after phase 5) stay with the batch: later tables
    // are carved out of them (same stream, so the order of use is the order of enqueueing) and at the end they go to
    // the block cache whole.  Handing them to the stream-ordered pool and asking it for the export tables made the
    // multi-GB requests block the host at random (0.3 ms on one step, 120 ms -- once 1.2 s -- on the next).
    struct Carve { size_t off, need; bool live; };
    struct Region { char* base; size_t size, used; std::vector<Carve> stack; };
    std::vector<Region> released;
    size_t bytes = 0;
    double alloc_ms = 0;   // host time inside the allocator (it blocks when the pool has to map memory)
    template <class T>
    T* alloc(size_t n, bool zero = false) {
        void* p = nullptr;
        const size_t want = (n ? n : 1) * sizeof(T);
        const size_t sz = size_class(want);
        auto t0 = std::chrono::steady_clock::now();
        bool carved = false;
#ifndef LB_SIMT_EMU
        {
            const size_t need = (want + 255) & ~(size_t)255;
            for (auto& r : released)
                if (r.size - r.used >= need) { p = r.base + r.used; r.stack.push_back(Carve{r.used, need, true}); r.used += need; carved = true; break; }
        }
        if (!p) {
            BlockCache& bc = block_cache(device);
            std::lock_guard<std::mutex> g(bc.mu);
            auto it = bc.free_.find(sz);
            if (it != bc.free_.end() && !it->second.empty()) { p = it->second.back(); it->second.pop_back(); bc.bytes -= sz; }
        }
#endif
        if (!p) {
            cudaError_t e = cudaMallocAsync(&p, sz, stream);
            if (e != cudaSuccess) {   // out of memory with blocks parked in the cache: give them back and try once more
                cudaGetLastError();
                cache_flush(device, stream);
                cudaStreamSynchronize(stream);
                e = cudaMallocAsync(&p, sz, stream);
            }
            if (e != cudaSuccess) {
                g_last_error = std::string("cudaMallocAsync(") + std::to_string(sz) + "): " + cudaGetErrorString(e);
                throw lb_status(LB_ERR_OOM);
            }
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        alloc_ms += ms;
        if (ms > 1.0 && getenv("LB_PHASE_TRACE")) fprintf(stderr, "[trace] slow alloc: %zu bytes took %.3f ms\n", sz, ms);
        if (!carved) { ptrs.push_back({p, sz}); bytes += sz; }
        if (zero) CK(cudaMemsetAsync(p, 0, want, stream));
        return (T*)p;
    }
    // the batch has enqueued the last consumer of a table: its block becomes room for later tables (see `released`)
    template <class T>
    void release(T*& p) {
        if (!p) return;
        for (size_t i = 0; i < ptrs.size(); i++)
            if (ptrs[i].first == (void*)p) {
#ifndef LB_SIMT_EMU
                released.push_back(Region{(char*)p, ptrs[i].second, 0, {}});
#else
                cudaFreeAsync((void*)p, stream);
#endif
                ptrs[i] = ptrs.back();
                ptrs.pop_back();
                break;
            }
        // a table carved out of a released block: its room is reusable once everything carved after it is gone too
        for (auto& r : released) {
            if ((char*)p < r.base || (char*)p >= r.base + r.size) continue;
            const size_t off = (size_t)((char*)p - r.base);
            for (auto& c : r.stack) if (c.off == off) c.live = false;
            while (!r.stack.empty() && !r.stack.back().live) { r.used = r.stack.back().off; r.stack.pop_back(); }
            break;
        }
        p = nullptr;
    }
    // the batch is gone and its stream has been synchronised: the blocks are free for any stream
    void free_all() {
#ifndef LB_SIMT_EMU
        for (auto& r : released) ptrs.push_back({(void*)r.base, r.size});
        released.clear();
        BlockCache& bc = block_cache(device);
        std::lock_guard<std::mutex> g(bc.mu);
        for (auto& pr : ptrs) {
            if (bc.bytes + pr.second <= cache_cap_bytes()) { bc.free_[pr.second].push_back(pr.first); bc.bytes += pr.second; }
            else cudaFreeAsync(pr.first, stream);
        }
#else
        for (auto& pr : ptrs) cudaFreeAsync(pr.first, stream);
#endif
        ptrs.clear();
    }
};

}  // namespace

struct lb_batch {
    Dev dev;
    size_t n_docs = 0;
    uint32_t flags = 0;
    bool owns_bytes = false;
    // device state
    const u8* d_bytes = nullptr;
    u64* d_offs = nullptr;
    u32* d_lens = nullptr;
    size_t n_blobs = 0;                       // blobs in the byte buffer (>= n_docs: import_batch groups)
    std::vector<u32> blob_doc, doc_blob0;     // blob -> document ; document -> first blob (n_docs + 1)
    std::vector<u32> doc_nprior;              // lb_docset_import: leading blobs of each document that restate its earlier state
    DocInfo* d_docs = nullptr;
    BlockInfo* d_blocks = nullptr;
    DocPeer* d_dpeer = nullptr;
    u8* d_json = nullptr;
    u8* d_export = nullptr;      // phase 7 output: one FastUpdates blob per document
    XDoc* d_xdoc = nullptr;
    u64 export_total = 0;
    std::vector<XDoc> xdocs;
    ExportTables xt{};            // phase-7 tables kept for export(updates(from)) on demand
    bool have_xt = false;
    std::unordered_map<size_t, std::vector<uint8_t>> from_exports;   // last on-demand export per document
    uint8_t* exported = nullptr;  // malloc'ed host copy (lbstage::download)
    bool export_fetched = false;
    u64 n_blocks = 0, n_changes = 0, n_rows = 0, n_peers_tot = 0, json_total = 0, n_deps = 0;
    Tables tb{};
    // host results
    std::vector<DocInfo> docs;
    std::vector<DocPeer> dpeer;          // packed: document d owns [peer_base[d], peer_base[d] + P)
    std::vector<u64> peer_base;
    char* json = nullptr;   // from lbstage::host_cache (never zero-filled): filled by lbstage::download
    bool json_fetched = false;
    // host-buffer entry point: the JSON goes home on a second stream while the export phase still computes
    bool eager_json = false, json_ok = true;
    int device = 0;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t json_ev = nullptr;
    std::thread json_thread;
    std::vector<lb_id_span> spans[4];          // success, pending, vv, frontiers of all documents, flat
    std::vector<size_t> span_off[4];           // document d owns [span_off[k][d], span_off[k][d + 1])
    std::vector<uint64_t> doc_ids;
    lb_counters counters{};
    lb_timings timings{};
    std::chrono::steady_clock::time_point t_call = std::chrono::steady_clock::now(), t_tail = t_call;
    cudaEvent_t ev[16];
    int n_ev = 0;
    bool ev_created = false;
};

// Persistent documents (lb_docset_*): what a document keeps between imports is its change store in wire form -- the
// FastUpdates blob it would export (ExportMode::all_updates), resident in device memory -- exactly what the reference's
// ChangeStore keeps (encoded blocks in a kv store, change_store.rs:60-110).  A document that still has pending changes
// keeps the blobs it was built from instead (pending changes are not part of an export).
struct DocsetBuf {     // one device buffer per import generation, shared by the documents stored in it
    u8* d = nullptr;
    size_t bytes = 0;
    ~DocsetBuf() { if (d) cudaFree(d); }
};
struct DocsetBlob { std::shared_ptr<DocsetBuf> buf; u64 off; u32 len; };
struct DocsetDoc { std::vector<DocsetBlob> blobs; };
struct lb_docset {
    int device = 0;
    std::mutex mu;
    std::unordered_map<u64, DocsetDoc> docs;
    u64 stored_bytes = 0;
};

namespace {

const int TPB = 128;
inline unsigned nblk(u64 n, int tpb = TPB) { return (unsigned)((n + tpb - 1) / tpb); }

// LB_PHASE_TRACE=1: host wall clock between named points (each one synchronises the stream: diagnosis only)
void trace_point(lb_batch* b, const char* name);
void mark(lb_batch* b) {
    if (b->n_ev < 16) CK(cudaEventRecord(b->ev[b->n_ev++], b->dev.stream));
}

template <class T>
T d2h_one(lb_batch* b, const T* src) {
    T v;
    CK(cudaMemcpyAsync(&v, src, sizeof(T), cudaMemcpyDeviceToHost, b->dev.stream));
    CK(cudaStreamSynchronize(b->dev.stream));
    return v;
}

void trace_point(lb_batch* b, const char* name) {
    static const bool on = getenv("LB_PHASE_TRACE") != nullptr;
    if (!on) return;
    static std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    cudaStreamSynchronize(b->dev.stream);
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[trace] %-24s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}

void run_scans(lb_batch* b, std::vector<ScanJob> jobs) {
    for (size_t i = 0; i < jobs.size(); i += 8) {
        ScanJobs sj;
        memset(&sj, 0, sizeof(sj));
        unsigned n = 0;
        for (; n < 8 && i + n < jobs.size(); n++) sj.j[n] = jobs[i + n];
        LB_LAUNCH(k_excl_scan_multi, n, 1024, 0, b->dev.stream, sj);
        b->timings.kernel_launches++;
    }
}
#define FIELD_JOB(base, type, in_field, out_field, count)                                              \
    ScanJob{(const u8*)(base) + offsetof(type, in_field), (u8*)(base) + offsetof(type, out_field), \
            sizeof(type), sizeof(type), (u64)(count)}

void pipeline(lb_batch* b) {
    Dev& dv = b->dev;
    cudaStream_t st = dv.stream;
    u32 D = (u32)b->n_docs;
    lb_timings& tm = b->timings;
    if (D == 0) {
        for (int i = 0; i < 8; i++) mark(b);
        b->docs.resize(1);
        return;
    }
    // ------------------------------------------------------------ phase 1: frame
    u32 Q = (u32)b->n_blobs;
    b->d_docs = dv.alloc<DocInfo>(D + 1, true);
    u32* d_blob_code = dv.alloc<u32>(Q + 1, true);
    u32* d_blob_nblocks = dv.alloc<u32>(Q + 1, true);
    u64* d_blob_block0 = dv.alloc<u64>(Q + 2, true);
    u32* d_blob_doc = dv.alloc<u32>(Q + 1);
    u32* d_doc_blob0 = dv.alloc<u32>(D + 2);
    CK(cudaMemcpyAsync(d_blob_doc, b->blob_doc.data(), sizeof(u32) * Q, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_doc_blob0, b->doc_blob0.data(), sizeof(u32) * (D + 1), cudaMemcpyHostToDevice, st));
    LB_LAUNCH(k_frame_count, nblk((u64)Q * 32, 128), 128, 0, st, b->d_bytes, b->d_offs, b->d_lens, Q, d_blob_code, d_blob_nblocks);
    run_scans(b, {ScanJob{(const u8*)d_blob_nblocks, (u8*)d_blob_block0, 4, 8, Q}});
    u32* d_doc_nprior = nullptr;
    if (!b->doc_nprior.empty()) {
        d_doc_nprior = dv.alloc<u32>(D + 1);
        CK(cudaMemcpyAsync(d_doc_nprior, b->doc_nprior.data(), sizeof(u32) * D, cudaMemcpyHostToDevice, st));
    }
    LB_LAUNCH(k_frame_docs, nblk(D), TPB, 0, st, D, d_doc_blob0, d_blob_code, d_blob_block0, d_doc_nprior, b->d_docs);
    tm.kernel_launches += 2;
    u64 B = d2h_one(b, d_blob_block0 + Q);
    b->n_blocks = B;
    b->d_blocks = dv.alloc<BlockInfo>(B + 1, true);
    LB_LAUNCH(k_frame_fill, nblk(Q), TPB, 0, st, b->d_bytes, b->d_offs, b->d_lens, Q, d_blob_doc, d_blob_code, d_blob_block0, d_doc_blob0, b->d_blocks);
    tm.kernel_launches += 1;
    mark(b);  // [1] frame done
    // ------------------------------------------------------------ phase 2: decode
    BlockInfo* blk = b->d_blocks;
    if (B) {
        LB_LAUNCH(k_block_count, nblk(B, 64), 64, 0, st, b->d_bytes, blk, B);
        tm.kernel_launches += 1;
    }
    run_scans(b, {FIELD_JOB(blk, BlockInfo, n_peers, peer0, B), FIELD_JOB(blk, BlockInfo, n_keys, key0, B),
                  FIELD_JOB(blk, BlockInfo, n_cids, cid0, B), FIELD_JOB(blk, BlockInfo, n_changes, ch0, B),
                  FIELD_JOB(blk, BlockInfo, n_deps, dep0, B), FIELD_JOB(blk, BlockInfo, n_ops, op0, B),
                  FIELD_JOB(blk, BlockInfo, n_dels, del0, B), FIELD_JOB(blk, BlockInfo, n_pos, pos0, B),
                  FIELD_JOB(blk, BlockInfo, pos_bytes, posb0, B), FIELD_JOB(blk, BlockInfo, n_tree, tr0, B)});
    BlockInfo tot = d2h_one(b, blk + B);
    u64 NP = tot.peer0, NK = tot.key0, NC = tot.cid0, NCH = tot.ch0, ND = tot.dep0, NR = tot.op0, NDEL = tot.del0;
    u64 NPOS = tot.pos0, NPOSB = tot.posb0, NTR = tot.tr0;
    if (NTR >= 0xFFFFFFFFull || NPOS >= 0xFFFFFFFFull) {
        g_last_error = "batch too large: tree ops / positions must fit 32 bits";
        throw lb_status(LB_ERR_INVALID_ARG);
    }
    if (NR >= 0xFFFFFFFFull || NCH >= 0xFFFFFFFFull) {
        g_last_error = "batch too large: op rows / changes must fit 32 bits";
        throw lb_status(LB_ERR_INVALID_ARG);
    }
    b->n_changes = NCH;
    b->n_rows = NR;
    b->n_peers_tot = NP;
    b->n_deps = ND;
    Tables& t = b->tb;
    t.peer_id = dv.alloc<u64>(NP);
    t.key_off = dv.alloc<u64>(NK); t.key_len = dv.alloc<u32>(NK);
    t.cid_root = dv.alloc<u8>(NC); t.cid_type = dv.alloc<u8>(NC); t.cid_peer_idx = dv.alloc<u32>(NC); t.cid_koc = dv.alloc<i32>(NC);
    t.ch_block = dv.alloc<u32>(NCH); t.ch_counter = dv.alloc<i32>(NCH); t.ch_len = dv.alloc<u32>(NCH);
    t.ch_lamport = dv.alloc<u32>(NCH); t.ch_ts = dv.alloc<i64>(NCH); t.ch_dep0 = dv.alloc<u64>(NCH);
    t.ch_msg_off = dv.alloc<u64>(NCH); t.ch_msg_len = dv.alloc<u32>(NCH, true);
    t.ch_ndeps = dv.alloc<u32>(NCH); t.ch_dep_self = dv.alloc<u8>(NCH); t.ch_op0 = dv.alloc<u64>(NCH);
    t.ch_nops = dv.alloc<u32>(NCH, true);
    t.dep_peer_idx = dv.alloc<u32>(ND); t.dep_counter = dv.alloc<i32>(ND);
    t.op_cid = dv.alloc<u32>(NR); t.op_prop = dv.alloc<i32>(NR); t.op_vtype = dv.alloc<u8>(NR); t.op_len = dv.alloc<u32>(NR);
    t.op_counter = dv.alloc<i32>(NR); t.op_change = dv.alloc<u32>(NR); t.op_val_off = dv.alloc<u64>(NR);
    t.op_val_len = dv.alloc<u32>(NR); t.op_del = dv.alloc<u32>(NR);
    t.del_peer_idx = dv.alloc<u32>(NDEL); t.del_counter = dv.alloc<i32>(NDEL); t.del_len = dv.alloc<i32>(NDEL);
    t.pos_off = dv.alloc<u64>(NPOS); t.pos_len = dv.alloc<u32>(NPOS); t.pos_pool = dv.alloc<u8>(NPOSB + 8);
    t.tr_target_peer = dv.alloc<u32>(NTR); t.tr_target_ctr = dv.alloc<i32>(NTR); t.tr_parent_kind = dv.alloc<u8>(NTR);
    t.tr_parent_peer = dv.alloc<u32>(NTR); t.tr_parent_ctr = dv.alloc<i32>(NTR); t.tr_pos = dv.alloc<u32>(NTR);
    t.dw_stats = dv.alloc<unsigned long long>(4, true);
    if (B) {
        // decoder variants (A/B switch LB_DECODE = rows | cols | warp | group): thread per block with all cursors at once,
        // thread per block one column at a time, warp per block on TMA-staged shared memory (lane-parallel run
        // expansion), warp per four TMA-staged blocks with a lane per column stream
        static const char* mode_env = getenv("LB_DECODE");
        static const int mode = !mode_env ? LB_DECODE_DEFAULT
                                : (!strcmp(mode_env, "rows") ? 0 : (!strcmp(mode_env, "warp") ? 2 : (!strcmp(mode_env, "group") ? 3 : 1)));
        if (mode == 0) LB_LAUNCH(k_block_decode, nblk(B, 64), 64, 0, st, b->d_bytes, blk, B, t);
        else if (mode == 1) LB_LAUNCH(k_block_decode_cols, nblk(B, 64), 64, 0, st, b->d_bytes, blk, B, t);
        else if (mode == 3) {
            const size_t smem = sizeof(DgWarp) * DG_WARPS;
#ifndef LB_SIMT_EMU
            static bool attr_set_g = false;
            if (!attr_set_g) { CK(cudaFuncSetAttribute(k_block_decode_group, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set_g = true; }
#endif
            LB_LAUNCH(k_block_decode_group, nblk(B, DG_G * DG_WARPS), 32 * DG_WARPS, smem, st, b->d_bytes, blk, B, t);
        } else {
            const size_t smem = sizeof(DwWarp) * DW_WARPS;
#ifndef LB_SIMT_EMU
            static bool attr_set = false;
            if (!attr_set) { CK(cudaFuncSetAttribute(k_block_decode_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; }
#endif
            LB_LAUNCH(k_block_decode_warp, nblk(B, DW_WARPS), 32 * DW_WARPS, smem, st, b->d_bytes, blk, B, t);
        }
        tm.kernel_launches += 1;
    }
    mark(b);  // [2] decode done
    // SURVEY 8d algorithmic bytes of decode: blob bytes read + SoA written
    tm.decode_bytes_written = NR * 13 + NCH * (4 + 4 + 8 + 4) + ND * 12;
    // ------------------------------------------------------------ phase 3: resolve
    ResolveTables rt;
    memset(&rt, 0, sizeof(rt));
    rt.peer_id = t.peer_id; rt.key_off = t.key_off; rt.key_len = t.key_len;
    rt.cid_root = t.cid_root; rt.cid_type = t.cid_type; rt.cid_peer_idx = t.cid_peer_idx; rt.cid_koc = t.cid_koc;
    rt.ch_block = t.ch_block; rt.ch_counter = t.ch_counter; rt.ch_len = t.ch_len; rt.ch_lamport_wire = t.ch_lamport;
    rt.ch_dep0 = t.ch_dep0; rt.ch_ndeps = t.ch_ndeps; rt.ch_dep_self = t.ch_dep_self;
    rt.dep_peer_idx = t.dep_peer_idx; rt.dep_counter = t.dep_counter;
    b->d_dpeer = dv.alloc<DocPeer>(NP, true);
    rt.dpeer = b->d_dpeer; rt.peer_map = dv.alloc<u32>(NP);
    DocContainer* dcont = dv.alloc<DocContainer>(NC + 1, true);
    rt.dcont = dcont; rt.cid_map = dv.alloc<u32>(NC);
    rt.dkey_off = dv.alloc<u64>(NK); rt.dkey_len = dv.alloc<u32>(NK); rt.key_map = dv.alloc<u32>(NK);
    rt.blk_order = dv.alloc<u32>(B);
    rt.ch_order = dv.alloc<u32>(NCH);
    rt.ch_aorder = dv.alloc<u32>(NCH);
    rt.ch_peer = dv.alloc<u16>(NCH);
    rt.ch_applied = dv.alloc<u8>(NCH, true);
    rt.ch_lamport = dv.alloc<u32>(NCH, true);
    rt.ch_walk = dv.alloc<u32>(NCH);
    rt.ch_pos = dv.alloc<u32>(NCH, true);
    rt.ch_trim = dv.alloc<u32>(NCH, true);
    // status of multi-blob documents (import_batch groups, lb_docset_import): per-copy epochs, per-blob pending hulls
    i32* d_pend_scratch = nullptr;
    if (Q > D) {
        rt.ch_epoch = dv.alloc<u32>(NCH);
        rt.ch_maxend = dv.alloc<i32>(NCH);
        rt.head_lamport = dv.alloc<u32>(NP);
        d_pend_scratch = dv.alloc<i32>(2 * (u64)Q + 2);
    }
    LB_LAUNCH(k_doc_tables, nblk(D, 64), 64, 0, st, b->d_bytes, b->d_docs, D, blk, rt);
    u32* d_tmp_a = dv.alloc<u32>(D + 1, true);
    u32* d_tmp_b = dv.alloc<u32>(D + 1, true);
    u32* d_tmp_c = dv.alloc<u32>(D + 1, true);
    LB_LAUNCH(k_doc_sizes, nblk(D), TPB, 0, st, b->d_docs, D, d_tmp_a, d_tmp_b, d_tmp_c, 0);
    tm.kernel_launches += 2;
    run_scans(b, {ScanJob{(const u8*)d_tmp_a, (u8*)b->d_docs + offsetof(DocInfo, vv0), 4, sizeof(DocInfo), D}});
    u64 VV = d2h_one(b, &b->d_docs[D].vv0);
    rt.ch_vv = dv.alloc<i32>(VV);
    u32* d_cursor = dv.alloc<u32>(NP);
    LB_LAUNCH(k_doc_causal, nblk(D, 64), 64, 0, st, b->d_docs, D, blk, rt, d_cursor, d_doc_blob0, d_pend_scratch);
    LB_LAUNCH(k_doc_frontiers, nblk(D, 64), 64, 0, st, b->d_docs, D, rt);
    LB_LAUNCH(k_doc_sizes, nblk(D), TPB, 0, st, b->d_docs, D, d_tmp_a, d_tmp_b, d_tmp_c, 1);
    tm.kernel_launches += 3;
    run_scans(b, {ScanJob{(const u8*)d_tmp_b, (u8*)b->d_docs + offsetof(DocInfo, atom0), 4, sizeof(DocInfo), D},
                  ScanJob{(const u8*)d_tmp_c, (u8*)b->d_docs + offsetof(DocInfo, mapslot0), 4, sizeof(DocInfo), D}});
    DocInfo dtot = d2h_one(b, &b->d_docs[D]);
    u64 NATOM = dtot.atom0, NSLOT = dtot.mapslot0;
    mark(b);  // [3] resolve done
    // ------------------------------------------------------------ phase 4: classify + map LWW
    ClassifyTables ct;
    memset(&ct, 0, sizeof(ct));
    ct.blocks = blk; ct.ch_block = t.ch_block; ct.ch_applied = rt.ch_applied; ct.ch_lamport = rt.ch_lamport;
    ct.ch_counter = t.ch_counter; ct.ch_peer = rt.ch_peer; ct.ch_trim = rt.ch_trim;
    ct.bytes = b->d_bytes; ct.op_val_off = t.op_val_off; ct.op_val_len = t.op_val_len;
    ct.op_cid = t.op_cid; ct.op_prop = t.op_prop; ct.op_vtype = t.op_vtype; ct.op_len = t.op_len;
    ct.op_counter = t.op_counter; ct.op_change = t.op_change;
    ct.op_del = t.op_del; ct.del_peer_idx = t.del_peer_idx; ct.del_counter = t.del_counter; ct.del_len = t.del_len;
    ct.peer_map = rt.peer_map;
    ct.tr_target_peer = t.tr_target_peer; ct.tr_target_ctr = t.tr_target_ctr; ct.tr_parent_kind = t.tr_parent_kind;
    ct.tr_parent_peer = t.tr_parent_peer; ct.tr_parent_ctr = t.tr_parent_ctr; ct.tr_pos = t.tr_pos;
    ct.tr_rec = dv.alloc<uint4>(NTR); ct.tr_key = dv.alloc<u64>(NTR); ct.tr_ids = dv.alloc<uint4>(NTR);
    ct.cid_map = rt.cid_map; ct.key_map = rt.key_map; ct.dcont = dcont; ct.dpeer = b->d_dpeer;
    ct.op_kind = dv.alloc<u8>(NR); ct.op_cidx = dv.alloc<u32>(NR); ct.op_lamport = dv.alloc<u32>(NR);
    ct.atom_row = dv.alloc<u32>(NATOM);
    ct.op_rec = dv.alloc<uint4>(NR); ct.op_aux = dv.alloc<u32>(NR);
    ct.map_best = dv.alloc<unsigned long long>(NSLOT, true);
    ct.map_row = dv.alloc<u32>(NSLOT);
    if (NR) {
        LB_LAUNCH(k_op_classify, nblk(NR, 256), 256, 0, st, b->d_docs, NR, ct);
        LB_LAUNCH(k_map_winner, nblk(NR, 256), 256, 0, st, b->d_docs, NR, ct);
        tm.kernel_launches += 2;
    }
    // capacities -> pools
    u32* cap_leaf = dv.alloc<u32>(NC + 1, true);
    u32* cap_node = dv.alloc<u32>(NC + 1, true);
    u32* cap_out = dv.alloc<u32>(NC + 1, true);
    u32* cap_cvv = dv.alloc<u32>(NC + 1, true);
    u32* span_cap = dv.alloc<u32>(D + 1, true);
    const u32 leaf_w = 32;   // slots per leaf = lanes per warp (k_seq.cuh)
    LB_LAUNCH(k_container_caps, nblk(D), TPB, 0, st, b->d_docs, D, dcont, cap_leaf, cap_node, cap_out, cap_cvv, span_cap, leaf_w);
    tm.kernel_launches += 1;
    run_scans(b, {ScanJob{(const u8*)cap_leaf, (u8*)dcont + offsetof(DocContainer, leaf0), 4, sizeof(DocContainer), NC},
                  ScanJob{(const u8*)cap_node, (u8*)dcont + offsetof(DocContainer, node0), 4, sizeof(DocContainer), NC},
                  ScanJob{(const u8*)cap_out, (u8*)dcont + offsetof(DocContainer, out0), 4, sizeof(DocContainer), NC},
                  ScanJob{(const u8*)cap_cvv, (u8*)dcont + offsetof(DocContainer, cvv0), 4, sizeof(DocContainer), NC},
                  ScanJob{(const u8*)span_cap, (u8*)b->d_docs + offsetof(DocInfo, span0), 4, sizeof(DocInfo), D}});
    DocContainer ctot = d2h_one(b, dcont + NC);
    DocInfo dtot2 = d2h_one(b, &b->d_docs[D]);
    u64 NLEAF = ctot.leaf0, NNODE = ctot.node0, NOUT = ctot.out0, NCVV = ctot.cvv0;
    (void)dtot2;
    mark(b);  // [4] classify done
    // ------------------------------------------------------------ phase 5: sequence integration
    SeqPools sp;
    memset(&sp, 0, sizeof(sp));
    sp.leaf = dv.alloc<uint4>(NLEAF * leaf_w);
    sp.node = dv.alloc<uint2>(NNODE * leaf_w);
    sp.node_parent = dv.alloc<u32>(NNODE);
    sp.atom_leaf = dv.alloc<u32>(NATOM);
    CK(cudaMemsetAsync(sp.atom_leaf, 0xFF, sizeof(u32) * NATOM, st));   // LEAF_NONE everywhere
    sp.a_org = dv.alloc<uint4>(NATOM);
    sp.cvv = dv.alloc<i32>(NCVV, true);
    sp.cont_epoch = dv.alloc<u32>(NC + 1);
    sp.out_row = dv.alloc<u32>(NOUT); sp.out_off = dv.alloc<u32>(NOUT); sp.out_len = dv.alloc<u32>(NOUT);
    SeqTables sq;
    memset(&sq, 0, sizeof(sq));
    sq.dpeer = b->d_dpeer; sq.dcont = dcont;
    sq.ch_walk = rt.ch_walk; sq.ch_op0 = t.ch_op0; sq.ch_nops = t.ch_nops; sq.ch_peer = rt.ch_peer; sq.ch_vv = rt.ch_vv;
    sq.ch_order = rt.ch_aorder; sq.ch_counter = t.ch_counter; sq.ch_ndeps = t.ch_ndeps; sq.ch_dep_self = t.ch_dep_self;
    sq.ch_pos = rt.ch_pos;
    sq.op_rec = ct.op_rec; sq.op_aux = ct.op_aux; sq.op_change = t.op_change; sq.op_counter = t.op_counter;
    sq.atom_row = ct.atom_row;
    LB_LAUNCH(k_seq_integrate, nblk(D, LB_SEQ_WARPS), 32 * LB_SEQ_WARPS, 0, st, b->d_docs, D, sp, sq);
    tm.kernel_launches += 1;
    if (!(b->flags & LB_FLAG_KEEP_DEVICE)) {   // the tracker pools are the largest tables of the batch: free them early
        dv.release(sp.leaf); dv.release(sp.node); dv.release(sp.node_parent); dv.release(sp.atom_leaf); dv.release(sp.a_org);
        dv.release(sp.cvv); dv.release(sp.cont_epoch); dv.release(ct.atom_row); dv.release(ct.op_rec);
    }
    mark(b);  // [5] list/text integration done
    // ------------------------------------------------------------ phase 5b: movable trees
    TreeTables tt;
    memset(&tt, 0, sizeof(tt));
    if (NTR) {
        CK(cudaMemsetAsync(d_tmp_b + D, 0, sizeof(u32), st));
        LB_LAUNCH(k_doc_sizes, nblk(D), TPB, 0, st, b->d_docs, D, d_tmp_a, d_tmp_b, d_tmp_c, 2);
        run_scans(b, {ScanJob{(const u8*)d_tmp_a, (u8*)b->d_docs + offsetof(DocInfo, tree0), 4, sizeof(DocInfo), D}});
        u64 NTS = d2h_one(b, &b->d_docs[D].tree0);
        const u32 max_atoms = d2h_one(b, d_tmp_b + D);
        tt.dpeer = b->d_dpeer; tt.blocks = blk; tt.op_cidx = ct.op_cidx; tt.op_lamport = ct.op_lamport;
        tt.tr_rec = ct.tr_rec; tt.tr_key = ct.tr_key;
        tt.ts_key = dv.alloc<u64>(NTR); tt.ts_val = dv.alloc<u32>(NTR); tt.ts_rec = dv.alloc<uint4>(NTR);
        tt.pos_off = t.pos_off; tt.pos_len = t.pos_len; tt.pos_pool = t.pos_pool;
        tt.tn_parent = dv.alloc<u32>(NTS); tt.tn_move = dv.alloc<u32>(NTS); tt.tn_base = dv.alloc<u32>(NTS);
        tt.tn_cnt = dv.alloc<u32>(NTS); tt.tn_sib = dv.alloc<u32>(NTS); tt.ns_key = dv.alloc<u64>(NTS);
        tt.tn_child = dv.alloc<u32>(NTS);
        tt.tn_root = dv.alloc<u32>(NTS); tt.tn_aopen = dv.alloc<u32>(NTS); tt.tn_aclose = dv.alloc<u32>(NTS);
        tt.dcont = dcont;
        // 16-bit parent links of one document in shared memory, sized for the largest tree document of the batch:
        // the number of resident documents (one sequential chain each) is what the apply kernel's speed depends on
        u32 s_nodes = max_atoms < TREE_S_NODES_MAX ? max_atoms : (u32)TREE_S_NODES_MAX;
        s_nodes = (s_nodes + 63u) & ~63u;
        const size_t tree_smem = (size_t)TREE_WARPS * s_nodes * sizeof(u16);
#ifndef LB_SIMT_EMU
        CK(cudaFuncSetAttribute(k_tree_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tree_smem));
        CK(cudaFuncSetAttribute(k_tree_apply, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        if (getenv("LB_PHASE_TRACE")) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tree_apply, 32 * TREE_WARPS, tree_smem);
            fprintf(stderr, "[trace] k_tree_apply: %zu bytes of shared memory per document, %d documents resident per SM\n", tree_smem, nb);
        }
#endif
        LB_LAUNCH(k_tree_sort, nblk((u64)D * 32, 128), 128, 0, st, b->d_docs, D, tt);
        LB_LAUNCH(k_tree_apply, nblk((u64)D * 32, 32 * TREE_WARPS), 32 * TREE_WARPS, tree_smem, st, b->d_docs, D, tt, s_nodes);
        LB_LAUNCH(k_tree_layout, nblk((u64)D * 32, 128), 128, 0, st, b->d_docs, D, tt);
        tm.kernel_launches += 4;
    }
    mark(b);  // [5b] trees done
    tm.tree_ops = NTR;
    // ------------------------------------------------------------ phase 6: JSON
    StateTables stt;
    memset(&stt, 0, sizeof(stt));
    stt.bytes = b->d_bytes; stt.dpeer = b->d_dpeer; stt.dcont = dcont;
    stt.dkey_off = rt.dkey_off; stt.dkey_len = rt.dkey_len; stt.map_row = ct.map_row; stt.map_best = ct.map_best;
    stt.op_kind = ct.op_kind; stt.op_vtype = t.op_vtype; stt.op_len = t.op_len; stt.op_counter = t.op_counter;
    stt.op_change = t.op_change; stt.op_val_off = t.op_val_off; stt.op_val_len = t.op_val_len; stt.ch_peer = rt.ch_peer;
    stt.ch_block = t.ch_block; stt.bkey_off = t.key_off; stt.bkey_len = t.key_len;
    stt.out_row = sp.out_row; stt.out_off = sp.out_off; stt.out_len = sp.out_len;
    stt.blocks = blk; stt.tn_parent = tt.tn_parent; stt.tn_move = tt.tn_move; stt.tn_base = tt.tn_base; stt.tn_cnt = tt.tn_cnt;
    stt.tn_sib = tt.tn_sib; stt.tn_child = tt.tn_child; stt.tr_rec = ct.tr_rec;
    stt.tn_root = tt.tn_root; stt.tn_aopen = tt.tn_aopen; stt.tn_aclose = tt.tn_aclose; stt.ns_key = tt.ns_key;
    stt.pos_off = t.pos_off; stt.pos_len = t.pos_len; stt.pos_pool = t.pos_pool;
    unsigned long long* d_acc = dv.alloc<unsigned long long>(4, true);
    if (!(b->flags & LB_FLAG_NO_JSON)) {
        LB_LAUNCH(k_json, nblk((u64)D * 32, 128), 128, 0, st, b->d_docs, D, stt, (u8*)nullptr, 0);
        LB_LAUNCH(k_json_padlen, nblk(D), TPB, 0, st, b->d_docs, D, d_tmp_a);
        tm.kernel_launches += 2;
        run_scans(b, {ScanJob{(const u8*)d_tmp_a, (u8*)b->d_docs + offsetof(DocInfo, json_off), 4, sizeof(DocInfo), D}});
        u64 JT = d2h_one(b, &b->d_docs[D].json_off);
        b->json_total = JT;
        b->d_json = dv.alloc<u8>(JT + 16, true);
        LB_LAUNCH(k_json, nblk((u64)D * 32, 128), 128, 0, st, b->d_docs, D, stt, b->d_json, 1);
        tm.kernel_launches += 1;
    }
    LB_LAUNCH(k_doc_hash, nblk((u64)D * 32, 128), 128, 0, st, b->d_docs, D, (const u8*)b->d_json, d_acc);
    tm.kernel_launches += 1;
    mark(b);  // [6] materialise done
    if (b->eager_json && b->json_total && !(b->flags & LB_FLAG_NO_JSON)) {
        CK(cudaStreamCreate(&b->stream2));
        CK(cudaEventCreate(&b->json_ev));
        CK(cudaEventRecord(b->json_ev, st));
        b->json = (char*)lbstage::host_cache().take(b->json_total + 1);
        if (!b->json) { g_last_error = "out of host memory"; throw lb_status(LB_ERR_OOM); }
        b->json_thread = std::thread([b] {
            bool ok = cudaSetDevice(b->device) == cudaSuccess && cudaStreamWaitEvent(b->stream2, b->json_ev, 0) == cudaSuccess &&
                      lbstage::download(b->d_json, (u8*)b->json, b->json_total, b->stream2);
            b->json[b->json_total] = 0;
            b->json_ok = ok;
        });
    }
    // ------------------------------------------------------------ phase 7: re-export (all_updates per document)
    if (b->flags & LB_FLAG_EXPORT) {
        trace_point(b, "before export");
        ExportTables xt;
        memset(&xt, 0, sizeof(xt));
        xt.bytes = b->d_bytes; xt.blocks = blk; xt.dpeer = b->d_dpeer; xt.dcont = dcont;
        xt.dkey_off = rt.dkey_off; xt.dkey_len = rt.dkey_len; xt.key_map = rt.key_map; xt.peer_map = rt.peer_map;
        xt.ch_order = rt.ch_aorder; xt.ch_applied = rt.ch_applied; xt.ch_block = t.ch_block; xt.ch_counter = t.ch_counter;
        xt.ch_len = t.ch_len; xt.ch_lamport = rt.ch_lamport; xt.ch_ts = t.ch_ts; xt.ch_op0 = t.ch_op0; xt.ch_nops = t.ch_nops;
        xt.ch_dep0 = t.ch_dep0; xt.ch_ndeps = t.ch_ndeps; xt.ch_dep_self = t.ch_dep_self;
        xt.dep_peer_idx = t.dep_peer_idx; xt.dep_counter = t.dep_counter;
        xt.ch_msg_off = t.ch_msg_off; xt.ch_msg_len = t.ch_msg_len;
        xt.op_kind = ct.op_kind; xt.op_vtype = t.op_vtype; xt.op_cidx = ct.op_cidx; xt.op_prop = t.op_prop; xt.op_len = t.op_len;
        xt.op_counter = t.op_counter; xt.op_val_off = t.op_val_off; xt.op_val_len = t.op_val_len; xt.op_del = t.op_del;
        xt.op_aux = ct.op_aux; xt.del_counter = t.del_counter; xt.del_len = t.del_len;
        xt.tr_ids = ct.tr_ids; xt.tr_pos = t.tr_pos; xt.pos_off = t.pos_off; xt.pos_len = t.pos_len; xt.pos_pool = t.pos_pool;
        if (NTR) {
            xt.pos_rank = dv.alloc<u32>(NPOS); xt.pos_rep = dv.alloc<u32>(NPOS);
            xt.ps_key = dv.alloc<u64>(NPOS); xt.ps_val = dv.alloc<u32>(NPOS);
        }
        xt.x_rec = dv.alloc<uint4>(NR); xt.r_bytes = dv.alloc<u32>(NR); xt.r_flag = dv.alloc<u8>(NR);
        xt.ch_nseg = dv.alloc<u32>(NCH + 1, true); xt.ch_novf = dv.alloc<u32>(NCH + 1, true);
        xt.ch_seg0 = dv.alloc<u64>(NCH + 2, true);
        xt.n_changes = NCH;
        xt.n_rows = NR;
        xt.ch_syn = dv.alloc<u32>(NCH + 1, true); xt.ch_syn0 = dv.alloc<u64>(NCH + 2, true);
        xt.xdoc = dv.alloc<XDoc>(D + 1, true);
        b->d_xdoc = xt.xdoc;
        xt.ch_aval = dv.alloc<u32>(NCH + 1, true); xt.ch_astr = dv.alloc<u32>(NCH + 1, true);
        xt.ch_aval0 = dv.alloc<u64>(NCH + 2, true); xt.ch_astr0 = dv.alloc<u64>(NCH + 2, true);
        // segment / final-change records: one slot per change + one per extra segment of a split change; the
        // extras are counted by pass 0, so the arrays are sized with a bound first and checked after the scan
        u64 SEGCAP = NCH + NCH / 4 + 1024;
        if (getenv("LB_EXPORT_TIGHT_SEGCAP")) SEGCAP = NCH;   // testing hook: force the growth path
        xt.sg_src = dv.alloc<u32>(SEGCAP); xt.sg_r0 = dv.alloc<u32>(SEGCAP); xt.sg_from = dv.alloc<u32>(SEGCAP);
        xt.sg_atoms = dv.alloc<u32>(SEGCAP); xt.sg_est = dv.alloc<u32>(SEGCAP); xt.sg_nmops = dv.alloc<u32>(SEGCAP);
        xt.sg_ndel = dv.alloc<u32>(SEGCAP); xt.sg_nrows = dv.alloc<u32>(SEGCAP); xt.sg_last_head = dv.alloc<u32>(SEGCAP);
        xt.sg_skip = dv.alloc<u32>(SEGCAP, true); xt.ch_trim = rt.ch_trim;
        xt.fc_src = dv.alloc<u32>(SEGCAP); xt.fc_pos = dv.alloc<u32>(SEGCAP); xt.fc_r0 = dv.alloc<u32>(SEGCAP);
        xt.fc_from = dv.alloc<u32>(SEGCAP); xt.fc_atoms = dv.alloc<u32>(SEGCAP); xt.fc_nrows = dv.alloc<u32>(SEGCAP);
        xt.fc_ndel = dv.alloc<u32>(SEGCAP); xt.fc_block = dv.alloc<u8>(SEGCAP); xt.fc_skip = dv.alloc<u32>(SEGCAP, true);
        xt.only_doc = 0xFFFFFFFFu; xt.from_ctr = nullptr;
        trace_point(b, "export allocs");
        LB_LAUNCH(k_exp_init, nblk(D), TPB, 0, st, b->d_docs, D, xt);
        if (NTR) { LB_LAUNCH(k_exp_posrank, nblk((u64)D * 32, 128), 128, 0, st, b->d_docs, D, xt); tm.kernel_launches += 1; }
        if (NCH) LB_LAUNCH(k_exp_arena, nblk(NCH, 64), 64, 0, st, NCH, xt, b->d_docs);
        run_scans(b, {ScanJob{(const u8*)xt.ch_aval, (u8*)xt.ch_aval0, 4, 8, NCH}, ScanJob{(const u8*)xt.ch_astr, (u8*)xt.ch_astr0, 4, 8, NCH}});
        trace_point(b, "posrank+arena");
        if (NCH) LB_LAUNCH(k_exp_changes, nblk(NCH, 64), 64, 0, st, b->d_docs, NCH, xt, 0);
        tm.kernel_launches += 3;
        trace_point(b, "changes pass 0");
        run_scans(b, {ScanJob{(const u8*)xt.ch_novf, (u8*)xt.ch_seg0, 4, 8, NCH}, ScanJob{(const u8*)xt.ch_syn, (u8*)xt.ch_syn0, 4, 8, NCH}});
        u64 NOVF = d2h_one(b, xt.ch_seg0 + NCH);
        u64 NSYN = d2h_one(b, xt.ch_syn0 + NCH);
        xt.has_syn = NSYN ? 1 : 0;
        xt.s_rec = dv.alloc<uint4>(NSYN); xt.s_len = dv.alloc<u32>(NSYN); xt.s_bytes = dv.alloc<u32>(NSYN);
        xt.s_flag = dv.alloc<u8>(NSYN); xt.s_voff = dv.alloc<u64>(NSYN); xt.s_vlen = dv.alloc<u32>(NSYN); xt.s_aux = dv.alloc<u32>(NSYN);
        if (NCH + NOVF > SEGCAP) {   // unusually many split changes: grow the tables, keep what pass 0 wrote
            u64 cap = NCH + NOVF;
            u32** sgs[10] = {&xt.sg_src, &xt.sg_r0, &xt.sg_from, &xt.sg_atoms, &xt.sg_est, &xt.sg_nmops, &xt.sg_ndel, &xt.sg_nrows, &xt.sg_last_head, &xt.sg_skip};
            for (auto pp : sgs) {
                u32* nw = dv.alloc<u32>(cap);
                CK(cudaMemcpyAsync(nw, *pp, sizeof(u32) * NCH, cudaMemcpyDeviceToDevice, st));
                dv.release(*pp);
                *pp = nw;
            }
            u32** fcs[8] = {&xt.fc_src, &xt.fc_pos, &xt.fc_r0, &xt.fc_from, &xt.fc_atoms, &xt.fc_nrows, &xt.fc_ndel, &xt.fc_skip};
            for (auto pp : fcs) { dv.release(*pp); *pp = dv.alloc<u32>(cap, true); }
            dv.release(xt.fc_block);
            xt.fc_block = dv.alloc<u8>(cap);
        }
        if (NOVF) { LB_LAUNCH(k_exp_changes, nblk(NCH, 64), 64, 0, st, b->d_docs, NCH, xt, 1); tm.kernel_launches += 1; }
        LB_LAUNCH(k_exp_store, nblk(D, 64), 64, 0, st, b->d_docs, D, xt);
        LB_LAUNCH(k_exp_sizes, nblk(D), TPB, 0, st, b->d_docs, D, xt, d_tmp_a, d_tmp_b);
        tm.kernel_launches += 2;
        run_scans(b, {ScanJob{(const u8*)d_tmp_a, (u8*)xt.xdoc + offsetof(XDoc, ob0), 4, sizeof(XDoc), D},
                      ScanJob{(const u8*)d_tmp_b, (u8*)xt.xdoc + offsetof(XDoc, scratch0), 4, sizeof(XDoc), D}});
        XDoc xtot = d2h_one(b, xt.xdoc + D);
        u64 NOB = xtot.ob0, NSCR = xtot.scratch0;
        XBlock* xb = dv.alloc<XBlock>(NOB + 1);
        u32* xscratch = dv.alloc<u32>(NSCR + 1);
        trace_point(b, "store+sizes");
        LB_LAUNCH(k_exp_list, nblk(D), TPB, 0, st, b->d_docs, D, xt, xb);
        if (NOB) { if (NOB < LB_XENC_CAP_BLOCKS) LB_LAUNCH(k_exp_encode<1>, nblk(NOB, 64), 64, 0, st, b->d_docs, NOB, xt, xb, xscratch, (u8*)nullptr, 0); else LB_LAUNCH(k_exp_encode<0>, nblk(NOB, 64), 64, 0, st, b->d_docs, NOB, xt, xb, xscratch, (u8*)nullptr, 0); }
        LB_LAUNCH(k_exp_layout, nblk(D), TPB, 0, st, b->d_docs, D, xt, xb, d_tmp_a);
        tm.kernel_launches += 3;
        trace_point(b, "encode pass 0");
        run_scans(b, {ScanJob{(const u8*)d_tmp_a, (u8*)xt.xdoc + offsetof(XDoc, exp_off), 4, sizeof(XDoc), D}});
        u64 XT = d2h_one(b, &xt.xdoc[D].exp_off);
        b->export_total = XT;
        trace_point(b, "layout scan + size d2h");
        b->d_export = dv.alloc<u8>(XT + 16, true);
        trace_point(b, "export buffer alloc+zero");
        if (NOB) { if (NOB < LB_XENC_CAP_BLOCKS) LB_LAUNCH(k_exp_encode<1>, nblk(NOB, 64), 64, 0, st, b->d_docs, NOB, xt, xb, xscratch, b->d_export, 1); else LB_LAUNCH(k_exp_encode<0>, nblk(NOB, 64), 64, 0, st, b->d_docs, NOB, xt, xb, xscratch, b->d_export, 1); }
        trace_point(b, "encode pass 1 kernel");
        LB_LAUNCH(k_exp_finish, nblk((u64)D * 32, 128), 128, 0, st, b->d_docs, D, xt, b->d_export);
        tm.kernel_launches += 2;
        trace_point(b, "encode pass 1");
        tm.export_bytes = XT;
        b->xt = xt;
        b->have_xt = true;
    }
    mark(b);  // [7] export done
    // ------------------------------------------------------------ results to host
    b->t_tail = std::chrono::steady_clock::now();
    b->docs.resize(D + 1);
    CK(cudaMemcpyAsync(b->docs.data(), b->d_docs, sizeof(DocInfo) * (D + 1), cudaMemcpyDeviceToHost, st));
    unsigned long long acc[4];
    CK(cudaMemcpyAsync(acc, d_acc, sizeof(acc), cudaMemcpyDeviceToHost, st));
    unsigned long long dws[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(dws, t.dw_stats, sizeof(dws), cudaMemcpyDeviceToHost, st));
    if (b->d_xdoc) {
        b->xdocs.resize(D);
        CK(cudaMemcpyAsync(b->xdocs.data(), b->d_xdoc, sizeof(XDoc) * D, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    // the peer table is sized by the blocks' peer registers (a few entries per BLOCK: hundreds of MB for 10^6 blocks) but a
    // document uses only its first P entries: those are packed on the device and only they travel
    {
        b->peer_base.assign(D + 1, 0);
        for (u32 d = 0; d < D; d++) b->peer_base[d + 1] = b->peer_base[d] + b->docs[d].P;
        const u64 total = b->peer_base[D];
        b->dpeer.resize(total);
        if (total) {
            u64* d_pbase = dv.alloc<u64>(D + 1);
            DocPeer* d_packed = dv.alloc<DocPeer>(total);
            CK(cudaMemcpyAsync(d_pbase, b->peer_base.data(), sizeof(u64) * (D + 1), cudaMemcpyHostToDevice, st));
            LB_LAUNCH(k_pack_peers, nblk(D), TPB, 0, st, b->d_docs, D, b->d_dpeer, d_pbase, d_packed);
            tm.kernel_launches += 1;
            CK(cudaMemcpyAsync(b->dpeer.data(), d_packed, sizeof(DocPeer) * total, cudaMemcpyDeviceToHost, st));
        }
    }
    mark(b);  // [8] d2h queued
    CK(cudaStreamSynchronize(st));
    lb_counters& c = b->counters;
    c.docs = D;
    c.docs_ok = acc[3];
    c.blocks = B;
    c.changes = NCH;
    c.op_rows = NR;
    c.atom_ops = acc[1];
    c.pending_changes = acc[2];
    c.state_hash = acc[0];
    tm.decode_fast_blocks = dws[0]; tm.decode_lane_blocks = dws[1]; tm.decode_unstaged_blocks = dws[2];
    c.json_bytes = 0;
    for (u32 d = 0; d < D; d++) c.json_bytes += b->docs[d].json_len;
    (void)NDEL;
}

// ImportStatus / vv / frontiers spans of every document, flat: [off[d], off[d + 1]) of one array per kind (a vector per
// document and kind cost ~0.1 s of host time per 10^5-document batch in allocations alone)
void build_status(lb_batch* b) {
    size_t D = b->n_docs;
    for (int k = 0; k < 4; k++) { b->span_off[k].assign(D + 1, 0); b->spans[k].clear(); }
    size_t total_peers = 0;
    for (size_t d = 0; d < D; d++) total_peers += b->docs[d].P;
    for (int k = 0; k < 4; k++) b->spans[k].reserve(k == 1 ? 16 : total_peers);
    for (size_t d = 0; d < D; d++) {
        DocInfo& di = b->docs[d];
        if (di.code == DOC_OK && di.has_unsupported) di.code = DOC_ERR_UNSUPPORTED;
        if (di.code == DOC_OK || di.code == DOC_ERR_UNSUPPORTED) {
            for (u32 p = 0; p < di.P; p++) {
                const DocPeer& dp = b->dpeer[b->peer_base[d] + p];
                if (dp.has_succ) b->spans[0].push_back(lb_id_span{dp.id, dp.succ_lo, dp.end_counter});
                if (dp.pend_hi > dp.pend_lo) b->spans[1].push_back(lb_id_span{dp.id, dp.pend_lo, dp.pend_hi});
                if (dp.end_counter > 0) b->spans[2].push_back(lb_id_span{dp.id, 0, dp.end_counter});
                if (dp.is_head && dp.end_counter > 0) b->spans[3].push_back(lb_id_span{dp.id, dp.end_counter - 1, dp.end_counter});
            }
        }
        for (int k = 0; k < 4; k++) b->span_off[k][d + 1] = b->spans[k].size();
    }
}

void timings_from_events(lb_batch* b) {
    auto el = [&](int a, int c) {
        float ms = 0;
        if (a < b->n_ev && c < b->n_ev) cudaEventElapsedTime(&ms, b->ev[a], b->ev[c]);
        return ms;
    };
    lb_timings& t = b->timings;
    t.h2d = el(0, 1);
    t.frame = el(1, 2);
    t.decode = el(2, 3);
    t.resolve = el(3, 4);
    t.classify = el(4, 5);
    t.integrate = el(5, 6);
    t.tree = el(6, 7);
    t.materialise = el(7, 8);
    t.reexport = el(8, 9);
    t.d2h = el(9, 10);
    t.total_device = el(1, 9);
}

lb_status run_batch(lb_batch* b) {
    try {
        pipeline(b);
        build_status(b);
        timings_from_events(b);
        auto now = std::chrono::steady_clock::now();
        b->timings.host_call_ms = std::chrono::duration<float, std::milli>(now - b->t_call).count();
        b->timings.host_tail_ms = std::chrono::duration<float, std::milli>(now - b->t_tail).count();
    } catch (lb_status s) {
        return s;
    }
    return LB_OK;
}

// host-side peek at a blob (only what import_batch's ordering needs: mode and the number of changes)
u32 blob_mode(const uint8_t* p, size_t n) { return n >= 22 ? ((u32)p[20] << 8) | p[21] : 0xFFFFu; }
u64 blob_change_count(const uint8_t* p, size_t n) {
    if (n < 22) return 0;
    size_t i = 22;
    auto varint = [&](u64* v) {
        *v = 0;
        for (int s = 0; s < 70 && i < n; s += 7) {
            uint8_t c = p[i++];
            *v |= (u64)(c & 0x7f) << (s < 64 ? s : 63);
            if (!(c & 0x80)) return true;
        }
        return false;
    };
    u64 total = 0;
    while (i < n) {
        u64 len, x;
        if (!varint(&len) || len > n - i) break;
        size_t end = i + (size_t)len;
        bool ok = true;
        for (int k = 0; k < 5 && ok; k++) ok = varint(&x) && i <= end;
        if (ok) total += x;   // the fifth varint of the envelope is n_changes
        i = end;
    }
    return total;
}

lb_status check_device(const lb_options* opt) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        g_last_error = "no CUDA device: loro_b200 has no CPU fallback";
        return LB_ERR_NO_DEVICE;
    }
    int dev = opt ? opt->device : 0;
    if (dev < 0 || dev >= n) {
        g_last_error = "bad device ordinal";
        return LB_ERR_INVALID_ARG;
    }
    if (cudaSetDevice(dev) != cudaSuccess) {
        g_last_error = "cudaSetDevice failed";
        return LB_ERR_CUDA;
    }
#ifndef LB_SIMT_EMU
    {   // keep freed table memory cached in the stream-ordered pool between batches
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
#endif
    return LB_OK;
}

void init_batch(lb_batch* b) {
    b->dev.device = b->device;
    b->dev.stream = stream_take(b->device);
    if (!b->dev.stream) { g_last_error = "cudaStreamCreate failed"; throw lb_status(LB_ERR_CUDA); }
    for (int i = 0; i < 16; i++) CK(cudaEventCreate(&b->ev[i]));
    b->ev_created = true;
}

// export(ExportMode::updates(from)) of one document (encoding.rs:79-83 ; change_store.rs:494-528 export_blocks_from):
// the import store is rebuilt for that document only, its changes are cut at `from` (Change::slice) on their way into
// the fresh export store, and the result is encoded like the import-time export.  The phase-7 tables of the batch are
// reused; only the per-call pieces (cut positions, block list, scratch, output) are allocated.
lb_status export_from(lb_batch* b, size_t doc, const lb_id_span* from, size_t n_from, std::vector<uint8_t>& out) {
    try {
        Dev& dv = b->dev;
        cudaStream_t st = dv.stream;
        const u32 D = (u32)b->n_docs;
        const u64 NCH = b->n_changes;
        const DocInfo& di = b->docs[doc];
        ExportTables xt = b->xt;
        // only the document's own peer slots are read (every kernel is restricted to `only_doc`)
        std::vector<i32> h_from(di.P + 1, 0);
        for (size_t k = 0; k < n_from; k++)
            for (u32 p = 0; p < di.P; p++)
                if (b->dpeer[b->peer_base[doc] + p].id == from[k].peer) h_from[p] = from[k].end;
        i32* d_from = dv.alloc<i32>(b->n_peers_tot + 1);
        CK(cudaMemcpyAsync(d_from + di.peer0, h_from.data(), sizeof(i32) * di.P, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // h_from is pageable host memory
        xt.from_ctr = d_from;
        xt.only_doc = (u32)doc;
        XDoc* xdoc = dv.alloc<XDoc>(D + 1, true);
        xt.xdoc = xdoc;
        u32* tmp_a = dv.alloc<u32>(D + 1, true);
        u32* tmp_b = dv.alloc<u32>(D + 1, true);
        LB_LAUNCH(k_exp_init, nblk(D), TPB, 0, st, b->d_docs, D, xt);
        if (NCH) {
            LB_LAUNCH(k_exp_changes, nblk(NCH, 64), 64, 0, st, b->d_docs, NCH, xt, 0);
            LB_LAUNCH(k_exp_changes, nblk(NCH, 64), 64, 0, st, b->d_docs, NCH, xt, 1);
        }
        LB_LAUNCH(k_exp_store, nblk(D, 64), 64, 0, st, b->d_docs, D, xt);
        LB_LAUNCH(k_exp_sizes, nblk(D), TPB, 0, st, b->d_docs, D, xt, tmp_a, tmp_b);
        run_scans(b, {ScanJob{(const u8*)tmp_a, (u8*)xdoc + offsetof(XDoc, ob0), 4, sizeof(XDoc), D},
                      ScanJob{(const u8*)tmp_b, (u8*)xdoc + offsetof(XDoc, scratch0), 4, sizeof(XDoc), D}});
        XDoc xtot = d2h_one(b, xdoc + D);
        u64 NOB = xtot.ob0, NSCR = xtot.scratch0;
        XBlock* xb = dv.alloc<XBlock>(NOB + 1);
        u32* xscratch = dv.alloc<u32>(NSCR + 1);
        LB_LAUNCH(k_exp_list, nblk(D), TPB, 0, st, b->d_docs, D, xt, xb);
        if (NOB) { if (NOB < LB_XENC_CAP_BLOCKS) LB_LAUNCH(k_exp_encode<1>, nblk(NOB, 64), 64, 0, st, b->d_docs, NOB, xt, xb, xscratch, (u8*)nullptr, 0); else LB_LAUNCH(k_exp_encode<0>, nblk(NOB, 64), 64, 0, st, b->d_docs, NOB, xt, xb, xscratch, (u8*)nullptr, 0); }
        LB_LAUNCH(k_exp_layout, nblk(D), TPB, 0, st, b->d_docs, D, xt, xb, tmp_a);
        run_scans(b, {ScanJob{(const u8*)tmp_a, (u8*)xdoc + offsetof(XDoc, exp_off), 4, sizeof(XDoc), D}});
        u64 XT = d2h_one(b, &xdoc[D].exp_off);
        u8* d_out = dv.alloc<u8>(XT + 16, true);
        if (NOB) { if (NOB < LB_XENC_CAP_BLOCKS) LB_LAUNCH(k_exp_encode<1>, nblk(NOB, 64), 64, 0, st, b->d_docs, NOB, xt, xb, xscratch, d_out, 1); else LB_LAUNCH(k_exp_encode<0>, nblk(NOB, 64), 64, 0, st, b->d_docs, NOB, xt, xb, xscratch, d_out, 1); }
        LB_LAUNCH(k_exp_finish, nblk((u64)D * 32, 128), 128, 0, st, b->d_docs, D, xt, d_out);
        XDoc x = d2h_one(b, xdoc + doc);
        lb_status rc = LB_OK;
        if ((x.flags & 1) || x.exp_len == 0) { g_last_error = "document uses features the export phase does not cover"; rc = LB_ERR_UNSUPPORTED; }
        else {
            out.resize(x.exp_len);
            CK(cudaMemcpyAsync(out.data(), d_out + x.exp_off, x.exp_len, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        dv.release(d_from); dv.release(xdoc); dv.release(tmp_a); dv.release(tmp_b); dv.release(xb); dv.release(xscratch); dv.release(d_out);
        return rc;
    } catch (lb_status s) {
        return s;
    }
}

}  // namespace

// After an import into a docset: every document of the batch whose import succeeded gets its new stored form -- the
// blobs it was built from (earlier ones first), copied out of the batch's byte buffer, or, with LB_FLAG_COMPACT, the
// blob it re-exports (ExportMode::all_updates) when nothing is pending and the export phase covers it.  A document
// whose import failed (checksum, decode, ...) keeps its earlier state: the reference rejects such an import before any
// state change.
void docset_store(lb_docset* set, lb_batch* b, const std::vector<u64>& offs, const std::vector<u32>& lens) {
    const size_t nd = b->n_docs;
    struct Pick { size_t doc; bool exported; };
    std::vector<Pick> picks;
    u64 total = 0;
    for (size_t d = 0; d < nd; d++) {
        const DocInfo& di = b->docs[d];
        if (di.code != DOC_OK && di.code != DOC_ERR_UNSUPPORTED) continue;
        // LB_FLAG_COMPACT: the document is replaced by a fresh one that imported its own export (what a host does to drop
        // redundant history: `fresh.import(doc.export(all_updates))`); otherwise it keeps every blob it ever imported, in
        // order, so that a later export is byte-identical to the reference's after the same sequence of imports (the
        // op segmentation of an export depends on which blob brought which piece of a change)
        bool exported = (b->flags & LB_FLAG_COMPACT) && di.code == DOC_OK && di.n_pending == 0 && d < b->xdocs.size() &&
                        !(b->xdocs[d].flags & 1) && b->xdocs[d].exp_len > 0;
        picks.push_back(Pick{d, exported});
        if (exported) total += ((u64)b->xdocs[d].exp_len + 15) & ~(u64)15;
        else for (u32 q = b->doc_blob0[d]; q < b->doc_blob0[d + 1]; q++) total += ((u64)lens[q] + 15) & ~(u64)15;
    }
    if (picks.empty()) return;
    auto buf = std::make_shared<DocsetBuf>();
    buf->bytes = total + 64;
    if (cudaMalloc((void**)&buf->d, buf->bytes) != cudaSuccess) {
        cudaGetLastError();
        g_last_error = "docset: out of device memory for the stored documents";
        throw lb_status(LB_ERR_OOM);
    }
    std::vector<CopySeg> segs;
    std::vector<std::pair<u64, DocsetDoc>> fresh;
    u64 w = 0;
    for (const Pick& pk : picks) {
        DocsetDoc nd_;
        auto push = [&](const u8* src, u32 len) {
            segs.push_back(CopySeg{src, buf->d + w, len});
            nd_.blobs.push_back(DocsetBlob{buf, w, len});
            w += ((u64)len + 15) & ~(u64)15;
        };
        if (pk.exported) push(b->d_export + b->xdocs[pk.doc].exp_off, b->xdocs[pk.doc].exp_len);
        else for (u32 q = b->doc_blob0[pk.doc]; q < b->doc_blob0[pk.doc + 1]; q++) push(b->d_bytes + offs[q], lens[q]);
        fresh.push_back({b->doc_ids[pk.doc], std::move(nd_)});
    }
    CopySeg* d_segs = b->dev.alloc<CopySeg>(segs.size());
    CK(cudaMemcpyAsync(d_segs, segs.data(), sizeof(CopySeg) * segs.size(), cudaMemcpyHostToDevice, b->dev.stream));
    LB_LAUNCH(k_copy_segments, nblk((u64)segs.size() * 32, 128), 128, 0, b->dev.stream, d_segs, (u32)segs.size());
    CK(cudaStreamSynchronize(b->dev.stream));
    for (auto& kv : fresh) {
        DocsetDoc& slot = set->docs[kv.first];
        for (const DocsetBlob& ob : slot.blobs) set->stored_bytes -= ob.len;
        slot = std::move(kv.second);
        for (const DocsetBlob& nb : slot.blobs) set->stored_bytes += nb.len;
    }
}

extern "C" {

// Host-buffer import, shared by lb_import_batch (fresh documents) and lb_docset_import (documents with an earlier state:
// their stored blobs come first, already in device memory, and count as `n_prior` for the import status).
static lb_status import_host(const lb_blob* blobs, size_t n_blobs, const lb_options* opt, lb_docset* set, lb_batch** out) {
    if (!out || (!blobs && n_blobs)) { g_last_error = "null argument"; return LB_ERR_INVALID_ARG; }
    *out = nullptr;
    lb_options o2;
    memset(&o2, 0, sizeof(o2));
    if (opt) o2 = *opt;
    if (set) { o2.device = set->device; o2.flags |= LB_FLAG_EXPORT; }   // the re-export is what a stored document keeps
    lb_status s = check_device(&o2);
    if (s != LB_OK) return s;
    if (n_blobs >= 0x7FFFFFFFull) { g_last_error = "too many blobs"; return LB_ERR_INVALID_ARG; }
    lb_batch* b = new lb_batch();
    b->n_docs = n_blobs;
    b->flags = o2.flags;
    b->device = o2.device;
    b->eager_json = true;   // host buffers in, host results expected
    try {
        init_batch(b);
        // blobs with the same doc_id form one document (LoroDoc::import_batch); documents are numbered in order of
        // first appearance and their blobs laid out consecutively, in the order given
        std::vector<u32> order(n_blobs);
        std::vector<u32> host_count;
        {
            std::unordered_map<u64, u32> doc_of;
            std::vector<u32> doc_idx(n_blobs);
            std::vector<u32>& count = host_count;
            for (size_t i = 0; i < n_blobs; i++) {
                auto it = doc_of.find(blobs[i].doc_id);
                if (it == doc_of.end()) {
                    it = doc_of.emplace(blobs[i].doc_id, (u32)count.size()).first;
                    count.push_back(0);
                    b->doc_ids.push_back(blobs[i].doc_id);
                }
                doc_idx[i] = it->second;
                count[it->second]++;
            }
            size_t nd = count.size();
            std::vector<u32> first(nd + 1, 0);
            for (size_t d = 0; d < nd; d++) first[d + 1] = first[d] + count[d];
            std::vector<u32> cursor(first.begin(), first.end() - 1);
            for (size_t i = 0; i < n_blobs; i++) order[cursor[doc_idx[i]]++] = (u32)i;
            // import_batch imports its blobs sorted by (mode, number of changes descending), stably
            // (loro.rs:1194-1202): the order decides where payloads land in the document's arenas, which the
            // re-export merge rules look at
            for (size_t d = 0; d < nd; d++) {
                u32 q0 = first[d], q1 = first[d + 1];
                if (q1 - q0 < 2) continue;
                std::vector<std::pair<std::pair<u32, i64>, u32>> keyed;
                for (u32 q = q0; q < q1; q++) {
                    const lb_blob& bl = blobs[order[q]];
                    keyed.push_back({{blob_mode(bl.ptr, bl.len), -(i64)blob_change_count(bl.ptr, bl.len)}, order[q]});
                }
                std::stable_sort(keyed.begin(), keyed.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
                for (u32 q = q0; q < q1; q++) order[q] = keyed[q - q0].second;
            }
            b->n_docs = nd;
        }
        const size_t nd = b->n_docs;
        // the blob list of the batch: per document, its stored blobs (device) then the new ones (host)
        std::vector<DocsetDoc*> prior(nd, nullptr);
        size_t n_prior_total = 0;
        if (set) {
            b->doc_nprior.assign(nd, 0);
            for (size_t d = 0; d < nd; d++) {
                auto it = set->docs.find(b->doc_ids[d]);
                if (it == set->docs.end()) continue;
                prior[d] = &it->second;
                b->doc_nprior[d] = (u32)it->second.blobs.size();
                n_prior_total += it->second.blobs.size();
            }
        }
        const size_t Q = n_blobs + n_prior_total;
        if (Q >= 0x7FFFFFFFull) { g_last_error = "too many blobs"; throw lb_status(LB_ERR_INVALID_ARG); }
        b->n_blobs = Q;
        b->doc_blob0.assign(nd + 1, 0);
        b->blob_doc.resize(Q);
        std::vector<u64> offs(Q + 1);
        std::vector<u32> lens(Q + 1, 0);
        std::vector<lbstage::BlobView> views;      // host blobs, with their offsets in the batch buffer
        std::vector<u64> view_offs;
        std::vector<CopySeg> segs;                 // stored blobs: device-to-device (dst filled in below)
        std::vector<u64> seg_offs;
        views.reserve(n_blobs);
        view_offs.reserve(n_blobs + 1);
        // the host blobs fill [0, H) of the batch buffer in one contiguous upload, the stored blobs follow
        u64 total = 0, ptotal = 0;
        for (size_t i = 0; i < n_blobs; i++) ptotal += (blobs[i].len + 15) & ~(u64)15;
        size_t q = 0, hq = 0;
        for (size_t d = 0; d < nd; d++) {
            b->doc_blob0[d] = (u32)q;
            if (prior[d])
                for (const DocsetBlob& sb : prior[d]->blobs) {
                    offs[q] = ptotal;
                    lens[q] = sb.len;
                    b->blob_doc[q] = (u32)d;
                    segs.push_back(CopySeg{sb.buf->d + sb.off, nullptr, sb.len});
                    seg_offs.push_back(ptotal);
                    ptotal += ((u64)sb.len + 15) & ~(u64)15;
                    b->counters.blob_bytes += sb.len;
                    q++;
                }
            for (u32 k = 0; k < host_count[d]; k++, hq++) {
                const lb_blob& bl = blobs[order[hq]];
                if (bl.len > 0xFFFFFFF0ull || (!bl.ptr && bl.len)) {
                    g_last_error = "blob too large or null";
                    throw lb_status(LB_ERR_INVALID_ARG);
                }
                offs[q] = total;
                lens[q] = (u32)bl.len;
                b->blob_doc[q] = (u32)d;
                views.push_back(lbstage::BlobView{bl.ptr, bl.len});
                view_offs.push_back(total);
                total += (bl.len + 15) & ~(u64)15;
                b->counters.blob_bytes += bl.len;
                q++;
            }
        }
        b->doc_blob0[nd] = (u32)q;
        offs[Q] = ptotal;
        view_offs.push_back(total);
        total = ptotal;
        // stage through the pinned ring: host gather of slot k overlaps the DMA of slot k-1 (host_stage.hpp)
        CK(cudaEventRecord(b->ev[b->n_ev++], b->dev.stream));  // [0]
        u8* d_bytes = b->dev.alloc<u8>(total + 64);
        b->d_offs = b->dev.alloc<u64>(Q + 1);
        b->d_lens = b->dev.alloc<u32>(Q + 1);
        if (!segs.empty()) {
            for (size_t k = 0; k < segs.size(); k++) segs[k].dst = d_bytes + seg_offs[k];
            CopySeg* d_segs = b->dev.alloc<CopySeg>(segs.size());
            CK(cudaMemcpyAsync(d_segs, segs.data(), sizeof(CopySeg) * segs.size(), cudaMemcpyHostToDevice, b->dev.stream));
            LB_LAUNCH(k_copy_segments, nblk((u64)segs.size() * 32, 128), 128, 0, b->dev.stream, d_segs, (u32)segs.size());
            CK(cudaStreamSynchronize(b->dev.stream));   // `segs` is pageable host memory
        }
        if (!views.empty() && !lbstage::upload_blobs(views.data(), view_offs.data(), views.size(), d_bytes, b->dev.stream)) {
            g_last_error = "h2d staging failed";
            throw lb_status(LB_ERR_CUDA);
        }
        CK(cudaMemcpyAsync(b->d_offs, offs.data(), sizeof(u64) * (Q + 1), cudaMemcpyHostToDevice, b->dev.stream));
        CK(cudaMemcpyAsync(b->d_lens, lens.data(), sizeof(u32) * (Q + 1), cudaMemcpyHostToDevice, b->dev.stream));
        b->d_bytes = d_bytes;
        b->timings.decode_bytes_read = b->counters.blob_bytes;
        mark(b);  // [1] h2d done (index 0 = start)
        // event indices: 0 start,1 h2d,2 frame,3 decode,4 resolve,5 classify,6 integrate,7 materialise,8 d2h
        s = run_batch(b);
        CK(cudaStreamSynchronize(b->dev.stream));
        if (s == LB_OK && set) docset_store(set, b, offs, lens);
    } catch (lb_status e) {
        s = e;
    }
    if (s != LB_OK) { lb_batch_free(b); return s; }
    *out = b;
    return LB_OK;
}

lb_status lb_import_batch(const lb_blob* blobs, size_t n_blobs, const lb_options* opt, lb_batch** out) {
    return import_host(blobs, n_blobs, opt, nullptr, out);
}

lb_status lb_docset_new(const lb_options* opt, lb_docset** out) {
    if (!out) { g_last_error = "null argument"; return LB_ERR_INVALID_ARG; }
    *out = nullptr;
    lb_status s = check_device(opt);
    if (s != LB_OK) return s;
    lb_docset* set = new lb_docset();
    set->device = opt ? opt->device : 0;
    *out = set;
    return LB_OK;
}

void lb_docset_free(lb_docset* set) { delete set; }

size_t lb_docset_doc_count(const lb_docset* set) { return set ? set->docs.size() : 0; }

uint64_t lb_docset_stored_bytes(const lb_docset* set) { return set ? set->stored_bytes : 0; }

lb_status lb_docset_import(lb_docset* set, const lb_blob* blobs, size_t n_blobs, const lb_options* opt, lb_batch** out) {
    if (!set) { g_last_error = "null argument"; return LB_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> g(set->mu);
    return import_host(blobs, n_blobs, opt, set, out);
}

lb_status lb_import_batch_device(const uint8_t* d_bytes, const uint64_t* offsets, const uint32_t* blob_lens,
                                 size_t n_docs, const lb_options* opt, lb_batch** out) {
    if (!out || ((!offsets || !blob_lens || !d_bytes) && n_docs)) { g_last_error = "null argument"; return LB_ERR_INVALID_ARG; }
    *out = nullptr;
    lb_status s = check_device(opt);
    if (s != LB_OK) return s;
    lb_batch* b = new lb_batch();
    b->n_docs = n_docs;
    b->flags = opt ? opt->flags : 0;
    b->device = opt ? opt->device : 0;
    try {
        init_batch(b);
        std::vector<u64> offs(n_docs + 1);
        std::vector<u32> lens(n_docs + 1, 0);
        for (size_t i = 0; i < n_docs; i++) {
            if (offsets[i] & 15) { g_last_error = "blob offsets must be multiples of 16"; throw lb_status(LB_ERR_INVALID_ARG); }
            offs[i] = offsets[i];
            lens[i] = blob_lens[i];
            b->doc_ids.push_back(i);
            b->blob_doc.push_back((u32)i);
            b->doc_blob0.push_back((u32)i);
            b->counters.blob_bytes += lens[i];
        }
        offs[n_docs] = n_docs ? offsets[n_docs - 1] + lens[n_docs - 1] : 0;
        b->doc_blob0.push_back((u32)n_docs);
        b->n_blobs = n_docs;
        CK(cudaEventRecord(b->ev[b->n_ev++], b->dev.stream));  // [0]
        b->d_offs = b->dev.alloc<u64>(n_docs + 1);
        b->d_lens = b->dev.alloc<u32>(n_docs + 1);
        CK(cudaMemcpyAsync(b->d_offs, offs.data(), sizeof(u64) * (n_docs + 1), cudaMemcpyHostToDevice, b->dev.stream));
        CK(cudaMemcpyAsync(b->d_lens, lens.data(), sizeof(u32) * (n_docs + 1), cudaMemcpyHostToDevice, b->dev.stream));
        b->d_bytes = d_bytes;
        b->timings.decode_bytes_read = b->counters.blob_bytes;
        mark(b);  // [1]
        s = run_batch(b);
    } catch (lb_status e) {
        s = e;
    }
    if (s != LB_OK) { lb_batch_free(b); return s; }
    *out = b;
    return LB_OK;
}

size_t lb_doc_count(const lb_batch* b) { return b ? b->n_docs : 0; }

lb_status lb_doc_status(const lb_batch* b, size_t doc, lb_import_status* out) {
    if (!b || !out || doc >= b->n_docs) { g_last_error = "bad argument"; return LB_ERR_INVALID_ARG; }
    out->code = (lb_doc_code)b->docs[doc].code;
    out->n_success = b->span_off[0][doc + 1] - b->span_off[0][doc];
    out->success = b->spans[0].data() + b->span_off[0][doc];
    out->n_pending = b->span_off[1][doc + 1] - b->span_off[1][doc];
    out->pending = b->spans[1].data() + b->span_off[1][doc];
    return LB_OK;
}

lb_status lb_doc_vv(const lb_batch* b, size_t doc, const lb_id_span** spans, size_t* n) {
    if (!b || !spans || !n || doc >= b->n_docs) { g_last_error = "bad argument"; return LB_ERR_INVALID_ARG; }
    *spans = b->spans[2].data() + b->span_off[2][doc];
    *n = b->span_off[2][doc + 1] - b->span_off[2][doc];
    return LB_OK;
}

lb_status lb_doc_frontiers(const lb_batch* b, size_t doc, const lb_id_span** spans, size_t* n) {
    if (!b || !spans || !n || doc >= b->n_docs) { g_last_error = "bad argument"; return LB_ERR_INVALID_ARG; }
    *spans = b->spans[3].data() + b->span_off[3][doc];
    *n = b->span_off[3][doc + 1] - b->span_off[3][doc];
    return LB_OK;
}

lb_status lb_doc_json(const lb_batch* cb, size_t doc, const char** utf8, size_t* len) {
    lb_batch* b = const_cast<lb_batch*>(cb);
    if (!b || !utf8 || !len || doc >= b->n_docs) { g_last_error = "bad argument"; return LB_ERR_INVALID_ARG; }
    if (b->flags & LB_FLAG_NO_JSON) { g_last_error = "batch was imported with LB_FLAG_NO_JSON"; return LB_ERR_INVALID_ARG; }
    if (b->json_thread.joinable()) {
        b->json_thread.join();
        if (!b->json_ok) { g_last_error = "json d2h failed"; return LB_ERR_CUDA; }
        b->json_fetched = true;
    }
    if (!b->json_fetched) {
        b->json = (char*)lbstage::host_cache().take(b->json_total + 1);
        if (!b->json) { g_last_error = "out of host memory"; return LB_ERR_OOM; }
        if (b->json_total && !lbstage::download(b->d_json, (u8*)b->json, b->json_total, b->dev.stream)) {
            g_last_error = "json d2h failed";
            return LB_ERR_CUDA;
        }
        b->json[b->json_total] = 0;
        b->json_fetched = true;
    }
    const DocInfo& di = b->docs[doc];
    if (di.code != DOC_OK) { *utf8 = ""; *len = 0; return LB_OK; }
    *utf8 = b->json + di.json_off;
    *len = di.json_len;
    return LB_OK;
}

lb_status lb_doc_export_updates(const lb_batch* cb, size_t doc, const lb_id_span* from, size_t n_from,
                                const uint8_t** bytes, size_t* len) {
    lb_batch* b = const_cast<lb_batch*>(cb);
    if (!b || !bytes || !len || doc >= b->n_docs) { g_last_error = "bad argument"; return LB_ERR_INVALID_ARG; }
    if (!(b->flags & LB_FLAG_EXPORT)) { g_last_error = "batch was imported without LB_FLAG_EXPORT"; return LB_ERR_INVALID_ARG; }
    if (b->docs[doc].code != DOC_OK) { g_last_error = "document failed to import"; return LB_ERR_INVALID_ARG; }
    if (from && n_from) {   // export(ExportMode::updates(from)): computed on demand for this document
        if (!b->have_xt) { g_last_error = "batch holds no export tables"; return LB_ERR_INVALID_ARG; }
        std::vector<uint8_t>& buf = b->from_exports[doc];
        lb_status rc = export_from(b, doc, from, n_from, buf);
        if (rc != LB_OK) return rc;
        *bytes = buf.data();
        *len = buf.size();
        return LB_OK;
    }
    const XDoc& x = b->xdocs[doc];
    if ((x.flags & 1) || x.exp_len == 0) { g_last_error = "document uses features the export phase does not cover"; return LB_ERR_UNSUPPORTED; }
    if (!b->export_fetched) {
        b->exported = (uint8_t*)lbstage::host_cache().take(b->export_total + 1);
        if (!b->exported) { g_last_error = "out of host memory"; return LB_ERR_OOM; }
        if (b->export_total && !lbstage::download(b->d_export, b->exported, b->export_total, b->dev.stream)) {
            g_last_error = "export d2h failed";
            return LB_ERR_CUDA;
        }
        b->export_fetched = true;
    }
    *bytes = b->exported + x.exp_off;
    *len = x.exp_len;
    return LB_OK;
}

lb_status lb_batch_counters(const lb_batch* b, lb_counters* out) {
    if (!b || !out) return LB_ERR_INVALID_ARG;
    *out = b->counters;
    return LB_OK;
}
lb_status lb_batch_timings(const lb_batch* b, lb_timings* out) {
    if (!b || !out) return LB_ERR_INVALID_ARG;
    *out = b->timings;
    out->alloc_host_ms = (float)b->dev.alloc_ms;
    out->device_bytes = b->dev.bytes;
    return LB_OK;
}

lb_status lb_debug_table(const lb_batch* b, const char* name, void* dst, size_t dst_bytes, size_t* n_elems,
                         size_t* elem_size) {
    if (!b || !name || !n_elems || !elem_size) return LB_ERR_INVALID_ARG;
    if (!(b->flags & LB_FLAG_KEEP_DEVICE)) { g_last_error = "needs LB_FLAG_KEEP_DEVICE"; return LB_ERR_INVALID_ARG; }
    std::string nm(name);
    const void* src = nullptr;
    size_t n = 0, es = 0;
    const Tables& t = b->tb;
#define TAB(str, ptr, cnt) if (nm == str) { src = ptr; n = cnt; es = sizeof(*ptr); }
    TAB("op_cid", t.op_cid, b->n_rows) TAB("op_prop", t.op_prop, b->n_rows) TAB("op_vtype", t.op_vtype, b->n_rows)
    TAB("op_len", t.op_len, b->n_rows) TAB("op_counter", t.op_counter, b->n_rows)
    TAB("ch_counter", t.ch_counter, b->n_changes) TAB("ch_len", t.ch_len, b->n_changes)
    TAB("ch_lamport", t.ch_lamport, b->n_changes) TAB("ch_ts", t.ch_ts, b->n_changes)
    TAB("dep_peer", t.dep_peer_idx, b->n_deps) TAB("dep_counter", t.dep_counter, b->n_deps)
#undef TAB
    if (nm == "blk_doc" || nm == "blk_nchanges") {   // fields of the block descriptors
        *n_elems = b->n_blocks;
        *elem_size = 4;
        if (dst) {
            if (dst_bytes < b->n_blocks * 4) { g_last_error = "buffer too small"; return LB_ERR_INVALID_ARG; }
            std::vector<BlockInfo> hb(b->n_blocks);
            if (b->n_blocks && cudaMemcpy(hb.data(), b->d_blocks, sizeof(BlockInfo) * b->n_blocks, cudaMemcpyDeviceToHost) != cudaSuccess) return LB_ERR_CUDA;
            for (size_t i = 0; i < hb.size(); i++) ((u32*)dst)[i] = nm == "blk_doc" ? hb[i].doc : hb[i].n_changes;
        }
        return LB_OK;
    }
    if (!src) { g_last_error = "unknown table"; return LB_ERR_INVALID_ARG; }
    *n_elems = n;
    *elem_size = es;
    if (dst) {
        if (dst_bytes < n * es) { g_last_error = "buffer too small"; return LB_ERR_INVALID_ARG; }
        if (n && cudaMemcpy(dst, src, n * es, cudaMemcpyDeviceToHost) != cudaSuccess) return LB_ERR_CUDA;
    }
    return LB_OK;
}

void lb_batch_free(lb_batch* b) {
    if (!b) return;
    if (b->json_thread.joinable()) b->json_thread.join();
    if (b->json_ev) cudaEventDestroy(b->json_ev);
    if (b->stream2) cudaStreamDestroy(b->stream2);
    if (b->ev_created) cudaStreamSynchronize(b->dev.stream);   // the cached blocks must be idle
    b->dev.free_all();
    if (b->ev_created) {   // the stream exists whenever the events do (init_batch)
        cudaStreamSynchronize(b->dev.stream);
        for (int i = 0; i < 16; i++) cudaEventDestroy(b->ev[i]);
        stream_give(b->device, b->dev.stream);
    }
    lbstage::host_cache().give(b->json);
    lbstage::host_cache().give(b->exported);
    delete b;
}

// Give the device blocks kept for the next batch (BlockCache) back to the driver.
lb_status lb_device_trim(int device) {
    lb_options o;
    memset(&o, 0, sizeof(o));
    o.device = device;
    lb_status st = check_device(&o);
    if (st != LB_OK) return st;
    cache_flush(device, nullptr);
    cudaStreamSynchronize(nullptr);
    return LB_OK;
}

// Pin the CALLING thread (and every thread it creates afterwards: the staging workers, the JSON download thread) to
// the CPUs of the NUMA node the device hangs off, so that the pinned staging ring and the gather threads of a rank stay
// next to its GPU.  One process per GPU calls this once, before it builds its input buffers.  LB_ERR_UNSUPPORTED when
// the topology cannot be read (no sysfs entry, single node): nothing is changed.
lb_status lb_numa_bind(int device) {
#ifdef LB_SIMT_EMU
    (void)device;
    g_last_error = "no topology in the emulated build";
    return LB_ERR_UNSUPPORTED;
#else
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { g_last_error = "cudaDeviceGetPCIBusId failed"; return LB_ERR_CUDA; }
    for (char* c = bus; *c; c++) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    int node = -1;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) { g_last_error = "no NUMA node recorded for " + std::string(bus); return LB_ERR_UNSUPPORTED; }
    path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
    f = fopen(path.c_str(), "r");
    if (!f) { g_last_error = "cannot read " + path; return LB_ERR_UNSUPPORTED; }
    char list[1024] = {0};
    if (!fgets(list, sizeof(list), f)) list[0] = 0;
    fclose(f);
    cpu_set_t want, have;
    CPU_ZERO(&want);
    for (char* p = list; *p;) {   // "0-31,64-95"
        char* e;
        long a = strtol(p, &e, 10), b_ = a;
        if (e == p) break;
        if (*e == '-') { p = e + 1; b_ = strtol(p, &e, 10); }
        for (long c = a; c <= b_ && c < CPU_SETSIZE; c++) CPU_SET((int)c, &want);
        p = *e == ',' ? e + 1 : e;
        if (*e != ',' ) break;
    }
    if (sched_getaffinity(0, sizeof(have), &have) == 0) {   // stay inside what the container allows
        cpu_set_t both;
        CPU_AND(&both, &want, &have);
        if (CPU_COUNT(&both) == 0) { g_last_error = "the device's NUMA node has no CPU this process may use"; return LB_ERR_UNSUPPORTED; }
        want = both;
    }
    if (sched_setaffinity(0, sizeof(want), &want) != 0) { g_last_error = "sched_setaffinity failed"; return LB_ERR_UNSUPPORTED; }
    return LB_OK;
#endif
}

#ifdef LB_SIMT_EMU
// test hook of the emulated build only: the device-side f64 formatter on the host (tests/test_f64_format.py)
int lb_emu_format_f64(double d, char* out) {
    u64 bits;
    memcpy(&bits, &d, 8);
    return f64_format(bits, out);
}
#endif
}  // extern "C"
