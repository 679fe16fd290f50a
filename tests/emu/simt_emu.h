// TEST INFRASTRUCTURE ONLY -- a tiny SIMT emulator so that the product's .cu sources can be compiled
// as plain C++ and *functionally* exercised in the GPU-less build container (pytest -m "not gpu").
//
// It is NOT a product code path: loro_b200/ never loads the emulated library, and the real library
// (libloro_b200.so, built by nvcc for sm_100a) has no CPU fallback.  The emulator runs every CUDA
// thread of a CTA as a fiber on one OS thread; warp collectives and __syncthreads are rendezvous
// points.  It checks logic (indexing, scans, state machines), not memory-model races or performance.
#pragma once
#ifndef LB_SIMT_EMU
#error "simt_emu.h is only for the LB_SIMT_EMU test build"
#endif
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_ { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) int4 { int x, y, z, w; };

namespace simt {

struct WarpSync {
    uint64_t gen = 0;
    unsigned arrived = 0;
    unsigned alive = 0;
    uint64_t slot[2][32];
    unsigned part[2] = {0, 0};  // participants of the rendezvous stored in slot[g]
    unsigned site[32];          // call-site tag of each lane's pending collective (divergence check)
    void* bt[32][12];
    int bt_n[32];
};
struct CtaSync {
    uint64_t gen = 0;
    unsigned arrived = 0;
    unsigned alive = 0;
};
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = false;
    unsigned tid = 0;
    unsigned lane = 0, warp = 0;
};
struct Ctx {
    uint3_ tid, bid;
    dim3 bdim, gdim;
    Fiber* fiber = nullptr;
    WarpSync* warp = nullptr;
    CtaSync* cta = nullptr;
    char* dyn_smem = nullptr;
    uint64_t progress = 0;  // bumped whenever a rendezvous completes or a fiber exits
};
extern thread_local Ctx* cur;
extern thread_local void* sched_sp;
extern "C" void simt_switch(void** save_sp, void* load_sp);
void yield();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

extern thread_local unsigned cur_site;
void divergence_abort(unsigned a, unsigned b, unsigned lane_a, unsigned lane_b);
extern bool capture_bt;
int capture_backtrace(void** out, int n);
inline uint64_t warp_exchange(unsigned mask, uint64_t v, int src_lane_valid_dummy = 0) {
    (void)src_lane_valid_dummy;
    WarpSync& w = *cur->warp;
    unsigned lane = cur->fiber->lane;
    unsigned expect = mask & w.alive;
    int g = (int)(w.gen & 1);
    w.slot[g][lane] = v;
    w.site[lane] = cur_site;
    if (capture_bt) w.bt_n[lane] = capture_backtrace(w.bt[lane], 12);
    w.arrived |= 1u << lane;
    if ((w.arrived & expect) == expect) {
        for (unsigned i = 0; i < 32; i++)
            if (((expect >> i) & 1) && w.site[i] != cur_site) divergence_abort(cur_site, w.site[i], lane, i);
        w.part[g] = expect;
        w.arrived = 0;
        w.gen++;
        cur->progress++;
    } else {
        uint64_t my = w.gen;
        while (w.gen == my) yield();
    }
    return (uint64_t)g;
}
}  // namespace simt

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define threadIdx (simt::cur->tid)
#define blockIdx (simt::cur->bid)
#define blockDim (simt::cur->bdim)
#define gridDim (simt::cur->gdim)
#define warpSize 32

// ---- warp collectives
inline void simt_syncwarp(unsigned mask = 0xffffffffu) { simt::warp_exchange(mask, 0); }
inline unsigned simt_ballot_sync(unsigned mask, int pred) {
    int g = (int)simt::warp_exchange(mask, pred ? 1 : 0);
    simt::WarpSync& w = *simt::cur->warp;
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if ((w.part[g] >> i) & 1)
            if (w.slot[g][i]) r |= 1u << i;
    return r;
}
inline int simt_any_sync(unsigned mask, int pred) { return simt_ballot_sync(mask, pred) != 0; }
inline int simt_all_sync(unsigned mask, int pred) {
    int g = (int)simt::warp_exchange(mask, pred ? 1 : 0);
    simt::WarpSync& w = *simt::cur->warp;
    for (int i = 0; i < 32; i++)
        if (((w.part[g] >> i) & 1) && !w.slot[g][i]) return 0;
    return 1;
}
inline unsigned __activemask() { return simt::cur->warp->alive; }
template <class T>
inline T simt_shfl_read(int g, int src) {
    uint64_t raw = simt::cur->warp->slot[g][src & 31];
    T out;
    std::memcpy(&out, &raw, sizeof(T));
    return out;
}
template <class T>
inline uint64_t simt_pack(T v) {
    static_assert(sizeof(T) <= 8, "shfl of >8 bytes");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    return raw;
}
template <class T>
inline T simt_shfl_sync(unsigned mask, T v, int src, int width = 32) {
    int g = (int)simt::warp_exchange(mask, simt_pack(v));
    int lane = (int)simt::cur->fiber->lane;
    int base = lane & ~(width - 1);
    return simt_shfl_read<T>(g, base + (src & (width - 1)));
}
template <class T>
inline T simt_shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int g = (int)simt::warp_exchange(mask, simt_pack(v));
    int lane = (int)simt::cur->fiber->lane;
    int base = lane & ~(width - 1);
    int src = lane - (int)delta;
    return src < base ? v : simt_shfl_read<T>(g, src);
}
template <class T>
inline T simt_shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int g = (int)simt::warp_exchange(mask, simt_pack(v));
    int lane = (int)simt::cur->fiber->lane;
    int base = lane & ~(width - 1);
    int src = lane + (int)delta;
    return src >= base + width ? v : simt_shfl_read<T>(g, src);
}
template <class T>
inline T simt_shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    int g = (int)simt::warp_exchange(mask, simt_pack(v));
    int lane = (int)simt::cur->fiber->lane;
    (void)width;
    return simt_shfl_read<T>(g, lane ^ lanemask);
}
inline void __syncthreads() {
    simt::CtaSync& c = *simt::cur->cta;
    c.arrived++;
    if (c.arrived >= c.alive) {
        c.arrived = 0;
        c.gen++;
        simt::cur->progress++;
    } else {
        uint64_t my = c.gen;
        while (c.gen == my) simt::yield();
    }
}
inline void __threadfence() {}
inline void __threadfence_block() {}

#define SIMT_SITE_ (simt::cur_site = (unsigned)(__LINE__ + 100000u * (unsigned)(sizeof(__FILE__) % 1000)))
#define __syncwarp(...) (SIMT_SITE_, simt_syncwarp(__VA_ARGS__))
#define __ballot_sync(...) (SIMT_SITE_, simt_ballot_sync(__VA_ARGS__))
#define __any_sync(...) (SIMT_SITE_, simt_any_sync(__VA_ARGS__))
#define __all_sync(...) (SIMT_SITE_, simt_all_sync(__VA_ARGS__))
#define __shfl_sync(...) (SIMT_SITE_, simt_shfl_sync(__VA_ARGS__))
#define __shfl_up_sync(...) (SIMT_SITE_, simt_shfl_up_sync(__VA_ARGS__))
#define __shfl_down_sync(...) (SIMT_SITE_, simt_shfl_down_sync(__VA_ARGS__))
#define __shfl_xor_sync(...) (SIMT_SITE_, simt_shfl_xor_sync(__VA_ARGS__))

// ---- bit intrinsics
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
    shift &= 31;
    return shift ? (hi << shift) | (lo >> (32 - shift)) : hi;
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 7;
        r |= (unsigned)((v >> (8 * sel)) & 0xff) << (8 * i);
    }
    return r;
}
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }

// ---- atomics (CTAs may run on several OS threads)
template <class T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicSub(T* p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicXor(T* p, T v) { return __atomic_fetch_xor(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicCAS(T* p, T cmp, T v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return cmp;
}
template <class T> inline T atomicMax(T* p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class T> inline T atomicMin(T* p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ---- a fake CUDA runtime (host side of the emulated build)
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct simt_event_* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
// LB_EMU_GUARD=1: every "device" allocation ends right before an inaccessible page (and starts right after one),
// so an out-of-bounds index faults at once instead of silently touching a neighbour (used by the fuzz test).
cudaError_t simt_guard_malloc(void** p, size_t n);
cudaError_t simt_guard_free(void* p);
bool simt_guard_enabled();
inline cudaError_t cudaMalloc(void** p, size_t n) {
    if (simt_guard_enabled()) return simt_guard_malloc(p, n);
    *p = std::malloc(n ? n : 1);
    return *p ? 0 : 2;
}
inline cudaError_t cudaFree(void* p) {
    if (simt_guard_enabled()) return simt_guard_free(p);
    std::free(p);
    return 0;
}
inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) std::memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { static int dummy; *s = &dummy; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaMemGetInfo(size_t* fr, size_t* tot) { *fr = *tot = (size_t)8 << 30; return 0; }
inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emu error" : "no error"; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }

// kernel launch: LB_LAUNCH(kernel, grid, block, smem_bytes, stream, args...)
#define LB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ((getenv("LB_EMU_KTRACE") ? fprintf(stderr, "simt_emu: launch %s\n", #kernel) : 0), \
     simt::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); }))
#define LB_DYN_SMEM(type, name) type* name = (type*)simt::cur->dyn_smem
