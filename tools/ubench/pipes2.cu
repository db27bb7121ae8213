// Integer-pipe microbenchmark for sm_100a, second take: every body is inline PTX (asm volatile) so the loop is exactly the
// listed instructions (check: cuobjdump -sass pipes2 | grep -A40 "Function : _Z1kILi<MODE>").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu && ./pipes2
// Reports warp-instructions per cycle per SM sub-partition (SMSP) at 1, 2, 4, 8 warps per SMSP; `instr` = instructions
// credited per chain per iteration.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
#define CH 8

template <int MODE>
__global__ void k(uint64_t* out, uint32_t seed, long long* cycles) {
    uint64_t acc[CH];
    uint32_t a[CH], b[CH], c[CH], d[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
        a[i] = seed + i * 7 + threadIdx.x; b[i] = seed * 3 + i + threadIdx.x; acc[i] = i + seed;
        c[i] = seed ^ (i * 11); d[i] = seed + i;
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b[i]));   // IMAD.WIDE.U32, 64-bit accumulate chain
            if (MODE == 1) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[i]) : "r"(a[i]), "r"(b[i]));        // IMAD.WIDE.U32 no addend, independent
            if (MODE == 2) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(a[i]), "r"(b[i]));        // IMAD
            if (MODE == 3) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(c[i]) : "r"(a[i]), "r"(b[i]));            // IMAD.HI.U32
            if (MODE == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(a[i]));                          // IADD3
            if (MODE == 5) {   // the kernels' multiply-accumulate: 2x (mad.lo.cc + madc.hi.cc) -> IMAD.WIDE.U32 / IMAD.WIDE.U32.X ?
                uint32_t lo = (uint32_t)acc[i], hi = (uint32_t)(acc[i] >> 32);
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;\n\taddc.u32 %4, %4, 0;"
                             : "+r"(lo), "+r"(hi), "+r"(c[i]) : "r"(a[i]), "r"(b[i]));
                acc[i] = ((uint64_t)hi << 32) | lo;
            }
            if (MODE == 6) {   // 1 IMAD.WIDE : 1 IADD3
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(a[i]));
            }
            if (MODE == 7) {   // 1 IMAD.WIDE : 2 IADD3
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(a[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(d[i]) : "r"(b[i]));
            }
            if (MODE == 8) {   // 1 IMAD.WIDE : 3 IADD3
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(a[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(d[i]) : "r"(b[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(b[i]));
            }
            if (MODE == 9) {   // 64-bit add with carry: IADD3 + IADD3.X
                uint32_t lo = (uint32_t)acc[i], hi = (uint32_t)(acc[i] >> 32);
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(a[i]), "r"(b[i]));
                acc[i] = ((uint64_t)hi << 32) | lo;
            }
            if (MODE == 10) asm volatile("mad.lo.u64 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"((uint64_t)a[i] | ((uint64_t)b[i] << 32)), "l"((uint64_t)b[i] | ((uint64_t)a[i] << 32)));   // 64-bit IMAD (lo)
            if (MODE == 11) {  // double-precision FMA
                double x = __longlong_as_double((long long)acc[i]);
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(1.0000001), "d"(0.5));
                acc[i] = (uint64_t)__double_as_longlong(x);
            }
            if (MODE == 12) {  // FFMA
                float x = __uint_as_float(c[i]);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(1.0001f), "f"(0.5f));
                c[i] = __float_as_uint(x);
            }
        }
    }
    const long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += acc[i] + a[i] + b[i] + c[i] + d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int MODE>
void run(const char* name, int instr) {
    uint64_t* out; long long* cyc; cudaMalloc(&out, 1 << 23); cudaMalloc(&cyc, 8);
    printf("%-44s", name);
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
    run<0>("IMAD.WIDE.U32 (acc chain)", 1);
    run<1>("IMAD.WIDE.U32 (mul.wide, independent)", 1);
    run<2>("IMAD (mad.lo.u32)", 1);
    run<3>("IMAD.HI.U32", 1);
    run<4>("IADD3", 1);
    run<5>("mad.lo.cc+madc.hi.cc+addc (3 PTX)", 3);
    run<6>("IMAD.WIDE + 1 IADD3", 2);
    run<7>("IMAD.WIDE + 2 IADD3", 3);
    run<8>("IMAD.WIDE + 3 IADD3", 4);
    run<9>("IADD3 + IADD3.X", 2);
    run<10>("mad.lo.u64", 1);
    run<11>("DFMA", 1);
    run<12>("FFMA", 1);
    return 0;
}
