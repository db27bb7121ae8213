// Integer-pipe microbenchmark for sm_100a: issue rate of the instructions the field kernels are made of.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each test runs ITER iterations of a body with CH independent dependency chains per thread (one block per SM) and
// reports warp-instructions per cycle per SM sub-partition (SMSP) at 1, 2, 4, 8 warps per SMSP.
// CAVEAT (read before quoting a number): ptxas rewrites these bodies — the IMAD.WIDE loop below compiles to 7 IMAD.WIDE +
// 7 carry adds + 2 moves per iteration, not 8 bare IMAD.WIDE — so the printed rates are LOWER BOUNDS of the per-instruction
// issue rate (measured on B200: IMAD.WIDE >= 0.175, IMAD / IADD3 / LOP3 ~ 0.5 per cycle per SMSP).  The figures DESIGN.md
// relies on come from ncu's pipe counters on the real kernels instead (sm__pipe_fmaheavy_cycles_active: IMAD.WIDE occupies
// the heavy FMA pipe for ~4 cycles per warp, the other IMAD forms for 2).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
#define CH 8

template <int MODE>
__global__ void k(uint64_t* out, uint32_t seed, long long* cycles) {
    uint64_t acc[CH];
    uint32_t a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed + i * 7 + threadIdx.x; b[i] = seed * 3 + i + threadIdx.x; acc[i] = i; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) acc[i] = (uint64_t)a[i] * (uint32_t)acc[i] + acc[i];                    // IMAD.WIDE.U32 (dependent chain)
            if (MODE == 1) { uint32_t lo = (uint32_t)acc[i]; lo = lo + a[i] + b[i]; acc[i] = (acc[i] & 0xFFFFFFFF00000000ULL) | lo; }   // IADD3
            if (MODE == 2) { uint32_t lo = (uint32_t)acc[i]; lo = lo * a[i] + b[i]; acc[i] = (acc[i] & 0xFFFFFFFF00000000ULL) | lo; }   // IMAD
            if (MODE == 3) { acc[i] = (uint64_t)a[i] * (uint32_t)acc[i] + acc[i]; b[i] = (b[i] ^ a[i]) & (uint32_t)it; }    // IMAD.WIDE + LOP3
            if (MODE == 4) { acc[i] = (uint64_t)a[i] * (uint32_t)acc[i] + acc[i]; b[i] = (b[i] ^ a[i]) & (uint32_t)it; a[i] = (a[i] | b[i]) ^ seed; }   // + 2 LOP3
            if (MODE == 5) { uint32_t lo = (uint32_t)acc[i]; lo = (lo ^ a[i]) & b[i]; acc[i] = (acc[i] & 0xFFFFFFFF00000000ULL) | lo; }   // LOP3
        }
    }
    const long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += acc[i] + a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int MODE>
void run(const char* name, int instr) {
    uint64_t* out; long long* cyc; cudaMalloc(&out, 1 << 23); cudaMalloc(&cyc, 8);
    printf("%-40s", name);
    for (int threads = 128; threads <= 1024; threads *= 2) {
        k<MODE><<<148, threads>>>(out, 12345, cyc);
        cudaDeviceSynchronize();
        k<MODE><<<148, threads>>>(out, 12345, cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  %dw: %.3f", threads / 128, (double)ITER * CH * instr * (threads / 32) / 4.0 / (double)c);
    }
    printf("   [warp-instr/cycle/SMSP]\n");
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("IMAD.WIDE.U32", 1);
    run<2>("IMAD", 1);
    run<1>("IADD3", 1);
    run<5>("LOP3", 1);
    run<3>("IMAD.WIDE + LOP3", 2);
    run<4>("IMAD.WIDE + 2 LOP3", 3);
    return 0;
}
