// TEST INFRASTRUCTURE: self-test of the SIMT emulator.
#include "simt_emu.h"
#include <numeric>

__global__ void k_scan(const int* in, int* out, int n) {
    __shared__ int warp_tot[32];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int v = i < n ? in[i] : 0;
    int s = v;
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += t;
    }
    if (lane == 31) warp_tot[w] = s;
    __syncthreads();
    if (w == 0) {
        int t = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
        int ts = t;
        for (int d = 1; d < 32; d <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, ts, d);
            if (lane >= d) ts += u;
        }
        warp_tot[lane] = ts - t;
    }
    __syncthreads();
    if (i < n) out[i] = s - v + warp_tot[w];
    unsigned b = __ballot_sync(0xffffffffu, v & 1);
    if (i < n && lane == 0) atomicAdd(&out[n], __popc(b));
}

int main() {
    int n = 1000, block = 256, grid = (n + block - 1) / block;
    std::vector<int> in(n), out(n + 1, 0);
    for (int i = 0; i < n; i++) in[i] = (i * 7) % 13;
    LB_LAUNCH(k_scan, grid, block, 0, 0, in.data(), out.data(), n);
    int odd = 0;
    for (int b = 0; b < grid; b++) {
        int acc = 0;
        for (int i = b * block; i < std::min(n, (b + 1) * block); i++) {
            if (out[i] != acc) { printf("FAIL at %d: %d vs %d\n", i, out[i], acc); return 1; }
            acc += in[i];
        }
    }
    for (int i = 0; i < n; i++) odd += in[i] & 1;
    if (out[n] != odd) { printf("FAIL ballot %d vs %d\n", out[n], odd); return 1; }
    printf("emu selftest ok\n");
    return 0;
}
