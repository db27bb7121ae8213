// NVLink P2P flag ping-pong between GPU0 and GPU1 (one process, peer access enabled): measures the
// store -> remote poll latency that bounds the per-round exchange of the sharded sumcheck, as a
// function of the idle gap between messages, with and without a background heartbeat store stream.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void pingpong(volatile unsigned long long* mine, volatile unsigned long long* peer, volatile unsigned long long* peer_scratch,
                         int iters, int first, int gap_ns, int heartbeat, unsigned long long* wait_ns) {
    if (threadIdx.x == 32) {   // heartbeat warp lane: keep traffic flowing to the peer
        if (!heartbeat) return;
        unsigned long long i = 0;
        while (*mine < (unsigned long long)iters) { peer_scratch[8] = ++i; unsigned long long t = gt(); while (gt() - t < (unsigned long long)heartbeat) {} }
        return;
    }
    if (threadIdx.x != 0) return;
    unsigned long long total = 0;
    for (int i = 1; i <= iters; i++) {
        if (first) {
            unsigned long long t = gt(); while (gt() - t < (unsigned long long)gap_ns) {}
            unsigned long long t0 = gt();
            *peer = i;
            while (*mine != (unsigned long long)i) {}
            total += gt() - t0;
        } else {
            while (*mine != (unsigned long long)i) {}
            *peer = i;
        }
    }
    if (first) *wait_ns = total / iters;
}
int main() {
    int n = 0; cudaGetDeviceCount(&n); if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    unsigned long long *f0, *f1, *w0, *w1;
    cudaSetDevice(0); cudaDeviceEnablePeerAccess(1, 0); cudaMalloc(&f0, 4096); cudaMalloc(&w0, 8);
    cudaSetDevice(1); cudaDeviceEnablePeerAccess(0, 0); cudaMalloc(&f1, 4096); cudaMalloc(&w1, 8);
    for (int hb = 0; hb <= 2000; hb = hb ? hb * 4 : 500)
        for (int gap = 0; gap <= 40000; gap = gap ? gap * 3 : 1000) {
            const int iters = 300;
            cudaSetDevice(0); cudaMemset(f0, 0, 4096); cudaSetDevice(1); cudaMemset(f1, 0, 4096); cudaDeviceSynchronize(); cudaSetDevice(0); cudaDeviceSynchronize();
            pingpong<<<1, 64>>>(f0, f1, f1, iters, 1, gap, hb, w0);
            cudaSetDevice(1); pingpong<<<1, 64>>>(f1, f0, f0, iters, 0, gap, hb, w1);
            cudaDeviceSynchronize(); cudaSetDevice(0); cudaDeviceSynchronize();
            unsigned long long w; cudaMemcpy(&w, w0, 8, cudaMemcpyDeviceToHost);
            printf("heartbeat %5d ns  idle gap %6d ns : round trip %6llu ns\n", hb, gap, w);
        }
    return 0;
}
