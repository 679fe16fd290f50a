// loro_b200 -- phase 1: blob header check, xxHash32 checksum, FastUpdates framing.
//
// Replaces (reference, relative to crates/loro-internal/src):
//   encoding.rs:299-330 parse_header_and_body, :278-295 check_checksum
//   encoding/fast_snapshot.rs:270-288 decode_updates (the ULEB128 length-prefixed block walk)
// One thread per blob: xxHash32 is a sequential recurrence per 16-byte stripe (4 independent
// accumulators), so the parallelism is across blobs; loads are 32-bit and the stripe loop is unrolled
// so that several independent loads are in flight per thread.
#pragma once
#include "lb_defs.h"

__device__ __forceinline__ u32 ld32(const u8* p) {  // p is 4-byte aligned
    return *(const u32*)p;
}

// the digest once the stripes before `off` are folded into v1..v4 (off == 0: nothing folded yet)
__device__ inline u32 xxh32_tail(const u8* d, size_t len, u32 seed, size_t off, u32 v1, u32 v2, u32 v3, u32 v4) {
    u32 h;
    if (len >= 16) {
        size_t limit = len - 16;
        // 64-byte steps: 16 independent loads issued before the dependent multiply chains
        while (off + 64 <= len) {
            u32 w[16];
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = ld32(d + off + 4 * i);
#pragma unroll
            for (int s = 0; s < 4; s++) {
                v1 = rotl32(v1 + w[4 * s + 0] * XXP2, 13) * XXP1;
                v2 = rotl32(v2 + w[4 * s + 1] * XXP2, 13) * XXP1;
                v3 = rotl32(v3 + w[4 * s + 2] * XXP2, 13) * XXP1;
                v4 = rotl32(v4 + w[4 * s + 3] * XXP2, 13) * XXP1;
            }
            off += 64;
        }
        while (off <= limit) {
            v1 = rotl32(v1 + ld32(d + off) * XXP2, 13) * XXP1;
            v2 = rotl32(v2 + ld32(d + off + 4) * XXP2, 13) * XXP1;
            v3 = rotl32(v3 + ld32(d + off + 8) * XXP2, 13) * XXP1;
            v4 = rotl32(v4 + ld32(d + off + 12) * XXP2, 13) * XXP1;
            off += 16;
        }
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else {
        h = seed + XXP5;
    }
    h += (u32)len;
    while (off + 4 <= len) {
        h = rotl32(h + ld32(d + off) * XXP3, 17) * XXP4;
        off += 4;
    }
    while (off < len) {
        h = rotl32(h + d[off] * XXP5, 11) * XXP1;
        off++;
    }
    h ^= h >> 15;
    h *= XXP2;
    h ^= h >> 13;
    h *= XXP3;
    h ^= h >> 16;
    return h;
}
__device__ u32 xxh32_dev(const u8* d, size_t len, u32 seed) {
    return xxh32_tail(d, len, seed, 0, seed + XXP1 + XXP2, seed + XXP2, seed, seed - XXP1);
}
// The same digest computed by a warp (every lane returns it).  The recurrence stays sequential -- four accumulators,
// one step per 16-byte stripe -- but the bytes arrive as coalesced 128-byte loads: lane L holds words L, L+32, L+64,
// L+96 of a 512-byte piece, i.e. component L&3 of stripes (L>>2) + 8k, and the lanes with (L&3) == c all run
// accumulator c, fetching stripe s from lane ((s&7)<<2)|c by shuffle.  A thread walking a 600 KB document alone
// waits for memory at every stripe; here the loads of the next piece are in flight while the chain runs.
__device__ inline u32 xxh32_warp(const u8* d, size_t len, u32 seed, int lane) {
    const int c = lane & 3;
    u32 v = c == 0 ? seed + XXP1 + XXP2 : c == 1 ? seed + XXP2 : c == 2 ? seed : seed - XXP1;
    size_t off = 0;
    if (len >= 1024) {
        u32 w0 = ld32(d + 4 * lane), w1 = ld32(d + 4 * (lane + 32)), w2 = ld32(d + 4 * (lane + 64)), w3 = ld32(d + 4 * (lane + 96));
        while (off + 512 <= len) {
            u32 n0 = 0, n1 = 0, n2 = 0, n3 = 0;
            if (off + 1024 <= len) {   // the next piece, before the chain of this one
                const u8* q = d + off + 512;
                n0 = ld32(q + 4 * lane); n1 = ld32(q + 4 * (lane + 32)); n2 = ld32(q + 4 * (lane + 64)); n3 = ld32(q + 4 * (lane + 96));
            }
#pragma unroll
            for (int s = 0; s < 8; s++) v = rotl32(v + __shfl_sync(LB_FULL, w0, (s << 2) | c) * XXP2, 13) * XXP1;
#pragma unroll
            for (int s = 0; s < 8; s++) v = rotl32(v + __shfl_sync(LB_FULL, w1, (s << 2) | c) * XXP2, 13) * XXP1;
#pragma unroll
            for (int s = 0; s < 8; s++) v = rotl32(v + __shfl_sync(LB_FULL, w2, (s << 2) | c) * XXP2, 13) * XXP1;
#pragma unroll
            for (int s = 0; s < 8; s++) v = rotl32(v + __shfl_sync(LB_FULL, w3, (s << 2) | c) * XXP2, 13) * XXP1;
            w0 = n0; w1 = n1; w2 = n2; w3 = n3;
            off += 512;
        }
    }
    u32 v1 = __shfl_sync(LB_FULL, v, 0), v2 = __shfl_sync(LB_FULL, v, 1), v3 = __shfl_sync(LB_FULL, v, 2), v4 = __shfl_sync(LB_FULL, v, 3);
    u32 h = 0;
    if (lane == 0) h = xxh32_tail(d, len, seed, off, v1, v2, v3, v4);
    return __shfl_sync(LB_FULL, h, 0);
}

// warp per blob: validate header + checksum (warp-cooperative xxHash32), count blocks (lane 0).
__global__ void k_frame_count(const u8* __restrict__ bytes, const u64* __restrict__ offs,
                              const u32* __restrict__ lens, u32 n_blobs, u32* __restrict__ blob_code,
                              u32* __restrict__ blob_nblocks) {
    u32 q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (q >= n_blobs) return;
    const u8* b = bytes + offs[q];   // blob starts are 16-byte aligned
    size_t n = lens[q];
    u32 code = DOC_OK;
    u32 nb = 0;
    if (n < 22) code = LB_ERR(DOC_ERR_DECODE);
    else if (b[0] != 'l' || b[1] != 'o' || b[2] != 'r' || b[3] != 'o') code = LB_ERR(DOC_ERR_DECODE);
    else {
        u32 mode = ((u32)b[20] << 8) | b[21];
        // the reference verifies the checksum of every mode it knows before looking further (encoding.rs:299-330), so a
        // damaged snapshot is a checksum error; an intact FastSnapshot (mode 3) is a scope limit of this engine, not an
        // unknown encoding
        if (mode != 3 && mode != 4) code = LB_ERR(DOC_ERR_MODE);
        else {
            u32 expect = (u32)b[16] | ((u32)b[17] << 8) | ((u32)b[18] << 16) | ((u32)b[19] << 24);
            if (xxh32_warp(b + 20, n - 20, XX_SEED_LORO, lane) != expect) code = LB_ERR(DOC_ERR_CHECKSUM);
            else if (mode == 3) code = LB_ERR(DOC_ERR_UNSUPPORTED);
            else if (lane == 0) {
                Cur c(b + 22, n - 22);
                while (!c.empty()) {
                    u64 len = c.varint();
                    c.skip(len);
                    if (c.err) break;
                    nb++;
                }
                if (c.err) { code = LB_ERR(DOC_ERR_DECODE); nb = 0; }
            }
        }
    }
    if (lane) return;
    blob_code[q] = code;
    blob_nblocks[q] = code == DOC_OK ? nb : 0;
}

// thread per document: a document is a run of consecutive blobs (LoroDoc::import_batch: loro.rs:1183-1290); the
// first blob that fails decides the document's code (header, checksum and mode are checked before any state
// change: loro.rs:584)
__global__ void k_frame_docs(u32 n_docs, const u32* __restrict__ doc_blob0, const u32* __restrict__ blob_code,
                             const u64* __restrict__ blob_block0, const u32* __restrict__ doc_nprior,
                             DocInfo* __restrict__ docs) {
    u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    u32 q0 = doc_blob0[d], q1 = doc_blob0[d + 1];
    u32 code = DOC_OK;
    for (u32 q = q0; q < q1 && code == DOC_OK; q++) code = blob_code[q];
    docs[d].code = code;
    docs[d].b0 = (u32)blob_block0[q0];
    docs[d].b1 = (u32)blob_block0[q1];
    docs[d].n_blobs = q1 - q0;
    docs[d].n_prior = doc_nprior ? doc_nprior[d] : 0;   // blobs that restate the document's earlier state (lb_docset_import)
}

// thread per blob: emit block descriptors at the scanned positions.
__global__ void k_frame_fill(const u8* __restrict__ bytes, const u64* __restrict__ offs,
                             const u32* __restrict__ lens, u32 n_blobs, const u32* __restrict__ blob_doc,
                             const u32* __restrict__ blob_code, const u64* __restrict__ blob_block0,
                             const u32* __restrict__ doc_blob0, BlockInfo* __restrict__ blocks) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_blobs) return;
    if (blob_code[q] != DOC_OK) return;
    const u8* b = bytes + offs[q];
    size_t n = lens[q];
    Cur c(b + 22, n - 22);
    u64 i = blob_block0[q];
    while (!c.empty()) {
        u64 len = c.varint();
        BlockInfo& bi = blocks[i++];
        bi.doc = blob_doc[q];
        bi.err = 0;
        bi.blob_rank = q - doc_blob0[blob_doc[q]];
        bi.off = offs[q] + (u64)(c.p - b);
        bi.len = (u32)len;
        c.skip(len);
    }
}
