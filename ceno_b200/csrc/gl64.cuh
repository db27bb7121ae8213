// gl64.cuh — branch-free Goldilocks (p = 2^64 - 2^32 + 1) and its quadratic extension
// F_p[X]/(X^2 - 7) in registers, for sm_100a.
//
// Restates the field of p3-goldilocks =0.4.3 / ff_ext::GoldilocksExt2 (reference Cargo.toml:34,
// W = 7).
//
// Design (the kernels are integer-issue bound, so instruction count is the budget):
//  * products are accumulated UNREDUCED in a 160-bit accumulator (acc160: two u64 + a carry word);
//    one reduction per dot product instead of one per multiply;
//  * the reduction uses 2^64 = EPS, 2^96 = -1, 2^128 = -2^32 (mod p) and is written as PTX
//    carry chains on 32-bit limbs (11 instructions, no compares/selects);
//  * "weak" results are any u64 congruent to the value (fine as a multiplicand); "canonical" results
//    are < p (required by gl_sub/gl_add operands and for everything that leaves the device);
//  * round evaluation points use subtraction only: with nd = lo - hi,  f(1) = hi, f(2) = hi - nd,
//    f(3) = f(2) - nd  (gl_sub is 5 instructions, a canonical gl_add 9).
//
// The same header compiles on the host (g++) with portable carry emulation so tests/test_field_host.py
// can check the algebra of every sequence against big-int arithmetic without a GPU; the device path
// is the PTX one.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define GL_DEV __device__ __forceinline__
#define GL_HD __host__ __device__ inline
#else
#define GL_DEV inline
#define GL_HD inline
#define __align__(x) __attribute__((aligned(x)))
#endif

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL

struct __align__(16) ext_t {
    uint64_t c0, c1;
};

// ---------------------------------------------------------------- 64x64 -> 128 product
GL_DEV void gl_mulwide(uint64_t a, uint64_t b, uint64_t& lo, uint64_t& hi) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);   // nvcc fuses the pair into 4 IMAD.WIDE.U32 + 2 IADD3
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    lo = (uint64_t)x;
    hi = (uint64_t)(x >> 64);
#endif
}

// ---------------------------------------------------------------- lazy accumulators (32-bit limbs)
// acc_t holds a sum of 64x64 products WITHOUT carry propagation between the product columns:
//   value = E + M 2^32,  E = e0 + e1 2^32 + e2 2^64 + e3 2^96 + e4 2^128,  M = m0 + m1 2^32 + m2 2^64.
// a*b adds a0 b0 into (e1:e0), a1 b1 into (e3:e2) [one carry chain, tail in e4] and a0 b1, a1 b0 into
// (m1:m0) [carry in m2].  Every window is an aligned register pair, so ptxas fuses each mad.lo.cc /
// madc.hi pair into one IMAD.WIDE.U32: a multiply-accumulate is 4 IMAD.WIDE + 3 carry adds.
struct acc_t {
    uint32_t e0, e1, e2, e3, e4, m0, m1, m2;
};
GL_DEV void acc_zero(acc_t& A) { A.e0 = A.e1 = A.e2 = A.e3 = A.e4 = A.m0 = A.m1 = A.m2 = 0; }
GL_DEV void acc_set64(acc_t& A, uint64_t x) { acc_zero(A); A.e0 = (uint32_t)x; A.e1 = (uint32_t)(x >> 32); }
// A += a*b   (any u64 operands; up to 2^31 accumulations)
GL_DEV void acc_mac(acc_t& A, uint64_t a, uint64_t b) {
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %8, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %5, %8, %11, %5;\n\t"
        "madc.hi.cc.u32 %6, %8, %11, %6;\n\t"
        "addc.u32 %7, %7, 0;\n\t"
        "mad.lo.cc.u32 %5, %9, %10, %5;\n\t"
        "madc.hi.cc.u32 %6, %9, %10, %6;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+r"(A.e0), "+r"(A.e1), "+r"(A.e2), "+r"(A.e3), "+r"(A.e4), "+r"(A.m0), "+r"(A.m1), "+r"(A.m2)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
#else
    unsigned __int128 E = ((unsigned __int128)A.e3 << 96) | ((unsigned __int128)A.e2 << 64) | ((unsigned __int128)A.e1 << 32) | A.e0;
    unsigned __int128 add = (unsigned __int128)((uint64_t)a0 * b0) + ((unsigned __int128)((uint64_t)a1 * b1) << 64);
    unsigned __int128 Es = E + add;
    A.e4 += (Es < E) ? 1u : 0u;
    A.e0 = (uint32_t)Es; A.e1 = (uint32_t)(Es >> 32); A.e2 = (uint32_t)(Es >> 64); A.e3 = (uint32_t)(Es >> 96);
    unsigned __int128 M = ((unsigned __int128)A.m2 << 64) | ((unsigned __int128)A.m1 << 32) | A.m0;
    M += (uint64_t)a0 * b1;
    M += (uint64_t)a1 * b0;
    A.m0 = (uint32_t)M; A.m1 = (uint32_t)(M >> 32); A.m2 = (uint32_t)(M >> 64);
#endif
}
// A = a*b into a FRESH accumulator: no zero-fill, the only possible carry is the second cross product's
// (4 IMAD.WIDE + 1 carry add instead of 8 register clears + 4 IMAD.WIDE + 3 carry adds)
GL_DEV void acc_mul(acc_t& A, uint64_t a, uint64_t b) {
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
#if defined(__CUDA_ARCH__)
    asm("{\n\t"
        ".reg .b64 t;\n\t"
        "mul.wide.u32 t, %8, %10;\n\t"
        "mov.b64 {%0, %1}, t;\n\t"
        "mul.wide.u32 t, %9, %11;\n\t"
        "mov.b64 {%2, %3}, t;\n\t"
        "mul.wide.u32 t, %8, %11;\n\t"
        "mov.b64 {%5, %6}, t;\n\t"
        "mad.lo.cc.u32 %5, %9, %10, %5;\n\t"
        "madc.hi.cc.u32 %6, %9, %10, %6;\n\t"
        "addc.u32 %7, 0, 0;\n\t"
        "mov.u32 %4, 0;\n\t"
        "}"
        : "=&r"(A.e0), "=&r"(A.e1), "=&r"(A.e2), "=&r"(A.e3), "=&r"(A.e4), "=&r"(A.m0), "=&r"(A.m1), "=&r"(A.m2)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
#else
    acc_zero(A);
    acc_mac(A, a, b);
#endif
}
// A = x + a*b into a FRESH accumulator (the fold's x + r d): x rides in the low product's addend, the carry
// runs through the high product and cannot leave it (a1 b1 + 1 < 2^64)
GL_DEV void acc_fma_first(acc_t& A, uint64_t x, uint64_t a, uint64_t b) {
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
#if defined(__CUDA_ARCH__)
    const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32);
    asm("{\n\t"
        ".reg .b64 t;\n\t"
        "mad.lo.cc.u32 %0, %8, %10, %12;\n\t"
        "madc.hi.cc.u32 %1, %8, %10, %13;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, 0;\n\t"
        "madc.hi.u32 %3, %9, %11, 0;\n\t"
        "mul.wide.u32 t, %8, %11;\n\t"
        "mov.b64 {%5, %6}, t;\n\t"
        "mad.lo.cc.u32 %5, %9, %10, %5;\n\t"
        "madc.hi.cc.u32 %6, %9, %10, %6;\n\t"
        "addc.u32 %7, 0, 0;\n\t"
        "mov.u32 %4, 0;\n\t"
        "}"
        : "=&r"(A.e0), "=&r"(A.e1), "=&r"(A.e2), "=&r"(A.e3), "=&r"(A.e4), "=&r"(A.m0), "=&r"(A.m1), "=&r"(A.m2)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(x0), "r"(x1));
#else
    acc_set64(A, x);
    acc_mac(A, a, b);
#endif
}
// x = l0 + l1 2^32 + l2 2^64 + l3 2^96 + l4 2^128 (l4 < 2^31) -> some u64 congruent to x ("weak"):
//   x = (l1:l0) + (l2 << 32) - (l2 + l3 + (l4 << 32))   (2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32)
//   T = (l1:l0) + (l2 << 32) (carry c), U = T - Z (borrow b), result = U + (c - b) EPS  — never wraps twice.
GL_DEV uint64_t gl_reduce_limbs(uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3, uint32_t l4) {
#if defined(__CUDA_ARCH__)
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 t1, c, z0, z1, u0, u1, d;\n\t"
        ".reg .b64 uu;\n\t"
        "add.cc.u32 t1, %2, %3;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "add.cc.u32 z0, %3, %4;\n\t"
        "addc.u32 z1, %5, 0;\n\t"
        "sub.cc.u32 u0, %1, z0;\n\t"
        "subc.cc.u32 u1, t1, z1;\n\t"
        "subc.u32 d, c, 0;\n\t"
        // result = U + d EPS = (U - d) + (d << 32), d in {-1,0,1}: both steps on the FMA pipe
        "mov.b64 uu, {u0, u1};\n\t"
        "mad.wide.s32 uu, d, -1, uu;\n\t"
        "mov.b64 {u0, u1}, uu;\n\t"
        "mad.lo.s32 u1, d, 1, u1;\n\t"
        "mov.b64 %0, {u0, u1};\n\t"
        "}"
        : "=l"(r) : "r"(l0), "r"(l1), "r"(l2), "r"(l3), "r"(l4));
    return r;
#else
    uint64_t t = (uint64_t)l1 + l2;
    const uint32_t t1 = (uint32_t)t, c = (uint32_t)(t >> 32);
    t = (uint64_t)l2 + l3;
    const uint32_t z0 = (uint32_t)t, z1 = l4 + (uint32_t)(t >> 32);
    int64_t u = (int64_t)l0 - z0;
    const uint32_t u0 = (uint32_t)u;
    int64_t bor = u < 0 ? 1 : 0;
    u = (int64_t)t1 - z1 - bor;
    const uint32_t u1 = (uint32_t)u;
    bor = u < 0 ? 1 : 0;
    const int32_t d = (int32_t)c - (int32_t)bor;
    const uint32_t nd = (uint32_t)(-d), dh = (uint32_t)(d >> 31);
    t = (uint64_t)u0 + nd;
    const uint32_t r0 = (uint32_t)t;
    const uint32_t r1 = u1 + dh + (uint32_t)(t >> 32);
    return ((uint64_t)r1 << 32) | r0;
#endif
}
GL_DEV uint64_t acc_reduce_weak(uint64_t s0, uint64_t s1, uint32_t s2) {
    return gl_reduce_limbs((uint32_t)s0, (uint32_t)(s0 >> 32), (uint32_t)s1, (uint32_t)(s1 >> 32), s2);
}
// fold M into E (one carry chain), then reduce
GL_DEV uint64_t acc_weak(const acc_t& A) {
    uint32_t l1, l2, l3, l4;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %4, %8;\n\t"
        "addc.cc.u32 %1, %5, %9;\n\t"
        "addc.cc.u32 %2, %6, %10;\n\t"
        "addc.u32 %3, %7, 0;"
        : "=r"(l1), "=r"(l2), "=r"(l3), "=r"(l4) : "r"(A.e1), "r"(A.e2), "r"(A.e3), "r"(A.e4), "r"(A.m0), "r"(A.m1), "r"(A.m2));
#else
    uint64_t t = (uint64_t)A.e1 + A.m0; l1 = (uint32_t)t;
    t = (uint64_t)A.e2 + A.m1 + (t >> 32); l2 = (uint32_t)t;
    t = (uint64_t)A.e3 + A.m2 + (t >> 32); l3 = (uint32_t)t;
    l4 = A.e4 + (uint32_t)(t >> 32);
#endif
    return gl_reduce_limbs(A.e0, l1, l2, l3, l4);
}
// compact accumulator (5 limbs) for long-lived sums: adding an acc_t costs 9 carry adds
struct cacc_t {
    uint32_t l0, l1, l2, l3, l4;
};
GL_DEV void cacc_zero(cacc_t& C) { C.l0 = C.l1 = C.l2 = C.l3 = C.l4 = 0; }
GL_DEV void cacc_add(cacc_t& C, const acc_t& A) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, %9;\n\t"
        "add.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(C.l0), "+r"(C.l1), "+r"(C.l2), "+r"(C.l3), "+r"(C.l4)
        : "r"(A.e0), "r"(A.e1), "r"(A.e2), "r"(A.e3), "r"(A.e4), "r"(A.m0), "r"(A.m1), "r"(A.m2));
#else
    uint64_t t = (uint64_t)C.l0 + A.e0; C.l0 = (uint32_t)t;
    t = (uint64_t)C.l1 + A.e1 + (t >> 32); C.l1 = (uint32_t)t;
    t = (uint64_t)C.l2 + A.e2 + (t >> 32); C.l2 = (uint32_t)t;
    t = (uint64_t)C.l3 + A.e3 + (t >> 32); C.l3 = (uint32_t)t;
    C.l4 += A.e4 + (uint32_t)(t >> 32);
    t = (uint64_t)C.l1 + A.m0; C.l1 = (uint32_t)t;
    t = (uint64_t)C.l2 + A.m1 + (t >> 32); C.l2 = (uint32_t)t;
    t = (uint64_t)C.l3 + A.m2 + (t >> 32); C.l3 = (uint32_t)t;
    C.l4 += (uint32_t)(t >> 32);
#endif
}
GL_DEV uint64_t cacc_weak(const cacc_t& C) { return gl_reduce_limbs(C.l0, C.l1, C.l2, C.l3, C.l4); }

// 96-bit lazy SUM of u64 values (linear layers without products: Poseidon2's MDS / internal sums): 3 carry adds per
// term, one reduction per linear combination.  Up to 2^32 terms.
struct wsum_t {
    uint32_t l0, l1, l2;
};
GL_DEV void wsum_set(wsum_t& W, uint64_t x) { W.l0 = (uint32_t)x; W.l1 = (uint32_t)(x >> 32); W.l2 = 0; }
GL_DEV void wsum_add(wsum_t& W, uint64_t x) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %0, %3;\n\t"
        "addc.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(W.l0), "+r"(W.l1), "+r"(W.l2) : "r"((uint32_t)x), "r"((uint32_t)(x >> 32)));
#else
    uint64_t t = (uint64_t)W.l0 + (uint32_t)x; W.l0 = (uint32_t)t;
    t = (uint64_t)W.l1 + (uint32_t)(x >> 32) + (t >> 32); W.l1 = (uint32_t)t;
    W.l2 += (uint32_t)(t >> 32);
#endif
}
GL_DEV void wsum_addw(wsum_t& W, const wsum_t& V) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %0, %3;\n\t"
        "addc.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 %2, %2, %5;"
        : "+r"(W.l0), "+r"(W.l1), "+r"(W.l2) : "r"(V.l0), "r"(V.l1), "r"(V.l2));
#else
    uint64_t t = (uint64_t)W.l0 + V.l0; W.l0 = (uint32_t)t;
    t = (uint64_t)W.l1 + V.l1 + (t >> 32); W.l1 = (uint32_t)t;
    W.l2 += V.l2 + (uint32_t)(t >> 32);
#endif
}
GL_DEV uint64_t wsum_weak(const wsum_t& W) { return gl_reduce_limbs(W.l0, W.l1, W.l2, 0, 0); }

// any u64 -> canonical:  x >= p  <=>  hi == 0xFFFFFFFF and lo != 0, and then x - p = lo - 1
// Device form: the same borrow chain as gl_sub(x, p) — x - p, and p is added back (as "- EPS") when that borrowed: five
// carry-chain instructions, no compares or selects (the compare/select form was 8 SASS instructions and a quarter of the
// round-1 kernel's issue slots).
GL_HD uint64_t gl_canon(uint64_t x) {
#if defined(__CUDA_ARCH__)
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 x0, x1, m;\n\t"
        "mov.b64 {x0, x1}, %1;\n\t"
        "sub.cc.u32 x0, x0, 1;\n\t"
        "subc.cc.u32 x1, x1, 0xFFFFFFFF;\n\t"
        "subc.u32 m, 0, 0;\n\t"          // m = -borrow: x < p
        "sub.cc.u32 x0, x0, m;\n\t"      // borrow: subtract EPS (== add p)
        "subc.u32 x1, x1, 0;\n\t"
        "mov.b64 %0, {x0, x1};\n\t"
        "}"
        : "=l"(r) : "l"(x));
    return r;
#else
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return (hi == 0xFFFFFFFFu && lo != 0) ? (uint64_t)(lo - 1) : x;
#endif
}
GL_DEV uint64_t acc_canon(const acc_t& A) { return gl_canon(acc_weak(A)); }
GL_DEV uint64_t cacc_canon(const cacc_t& C) { return gl_canon(cacc_weak(C)); }

// ---------------------------------------------------------------- base field
// a, b canonical -> canonical (5 instructions, ALU pipe)
GL_DEV uint64_t gl_sub(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, m;\n\t"
        "mov.b64 {a0, a1}, %1;\n\t"
        "mov.b64 {b0, b1}, %2;\n\t"
        "sub.cc.u32 a0, a0, b0;\n\t"
        "subc.cc.u32 a1, a1, b1;\n\t"
        "subc.u32 m, 0, 0;\n\t"          // m = -borrow
        "sub.cc.u32 a0, a0, m;\n\t"      // borrow: subtract EPS (== add p)
        "subc.u32 a1, a1, 0;\n\t"
        "mov.b64 %0, {a0, a1};\n\t"
        "}"
        : "=l"(r) : "l"(a), "l"(b));
    return r;
#else
    uint64_t d = a - b;
    return (a < b) ? d - GL_EPS : d;
#endif
}
// a, b canonical -> canonical
GL_DEV uint64_t gl_add(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, m;\n\t"
        ".reg .pred q;\n\t"
        "mov.b64 {a0, a1}, %1;\n\t"
        "mov.b64 {b0, b1}, %2;\n\t"
        // nb = p - b  (b = 0 gives p, which the subtraction below handles)
        "sub.cc.u32 b0, 1, b0;\n\t"
        "subc.u32 b1, 0xFFFFFFFF, b1;\n\t"
        // a - nb
        "sub.cc.u32 a0, a0, b0;\n\t"
        "subc.cc.u32 a1, a1, b1;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 a0, a0, m;\n\t"
        "subc.u32 a1, a1, 0;\n\t"
        "mov.b64 %0, {a0, a1};\n\t"
        "}"
        : "=l"(r) : "l"(a), "l"(b));
    return r;
#else
    uint64_t nb = GL_P - b;          // in [1, p]
    uint64_t d = a - nb;
    return (a < nb) ? d - GL_EPS : d;
#endif
}
GL_DEV uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }

// any u64 operands -> weak / canonical product
GL_DEV uint64_t gl_mul_weak(uint64_t a, uint64_t b) {
    uint64_t lo, hi;
    gl_mulwide(a, b, lo, hi);
    return acc_reduce_weak(lo, hi, 0);
}
GL_DEV uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_canon(gl_mul_weak(a, b)); }
// 7 a for any u64 a -> weak   (7a < 2^67)
GL_DEV uint64_t gl_mul7_weak(uint64_t a) {
#if defined(__CUDA_ARCH__)
    // 7a = (u0:t0) + u1 2^64 with u1 <= 6, and 2^64 = EPS: two wrap-free corrections (5 instructions)
    uint64_t r;
    asm("{\n\t"
        ".reg .u32 x0, x1, t0, t1, u0, u1, v0, v1, c;\n\t"
        ".reg .b64 t, u, v;\n\t"
        "mov.b64 {x0, x1}, %1;\n\t"
        "mul.wide.u32 t, x0, 7;\n\t"
        "mov.b64 {t0, t1}, t;\n\t"
        "mov.b64 u, {t1, 0};\n\t"
        "mad.wide.u32 u, x1, 7, u;\n\t"
        "mov.b64 {u0, u1}, u;\n\t"
        "mad.lo.cc.u32 v0, u1, 0xFFFFFFFF, t0;\n\t"
        "madc.hi.cc.u32 v1, u1, 0xFFFFFFFF, u0;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "mov.b64 v, {v0, v1};\n\t"
        "mad.wide.u32 %0, c, 0xFFFFFFFF, v;\n\t"
        "}"
        : "=l"(r) : "l"(a));
    return r;
#else
    unsigned __int128 x = (unsigned __int128)a * 7;
    unsigned __int128 v = (unsigned __int128)(uint64_t)x + (unsigned __int128)(uint64_t)(x >> 64) * GL_EPS;
    uint64_t lo = (uint64_t)v;
    if ((uint64_t)(v >> 64)) lo += GL_EPS;   // cannot wrap again
    return lo;
#endif
}

// ---------------------------------------------------------------- extension
GL_DEV ext_t ext_make(uint64_t a, uint64_t b) { ext_t r; r.c0 = a; r.c1 = b; return r; }
GL_DEV ext_t ext_zero() { return ext_make(0, 0); }
GL_DEV ext_t ext_one() { return ext_make(1, 0); }
GL_DEV ext_t ext_add(ext_t a, ext_t b) { return ext_make(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
GL_DEV ext_t ext_sub(ext_t a, ext_t b) { return ext_make(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
GL_DEV ext_t ext_canon(ext_t a) { return ext_make(gl_canon(a.c0), gl_canon(a.c1)); }

// lazy extension accumulator: value = A0 + A1 X, both unreduced
struct eacc {
    acc_t A0, A1;
};
struct ecacc {       // compact, for the per-thread round sums
    cacc_t A0, A1;
};
GL_DEV void ecacc_zero(ecacc& E) { cacc_zero(E.A0); cacc_zero(E.A1); }
GL_DEV void eacc_zero(eacc& E) { acc_zero(E.A0); acc_zero(E.A1); }
// E += a * b with b's 7*c1 supplied:  (a0 + a1 X)(b0 + b1 X) = a0 b0 + a1 (7 b1) + (a0 b1 + a1 b0) X
GL_DEV void eacc_mac(eacc& E, ext_t a, ext_t b, uint64_t b1_7) {
    acc_mac(E.A0, a.c0, b.c0);
    acc_mac(E.A0, a.c1, b1_7);
    acc_mac(E.A1, a.c0, b.c1);
    acc_mac(E.A1, a.c1, b.c0);
}
// E = a * b (fresh accumulators)
GL_DEV void eacc_mul(eacc& E, ext_t a, ext_t b, uint64_t b1_7) {
    acc_mul(E.A0, a.c0, b.c0);
    acc_mac(E.A0, a.c1, b1_7);
    acc_mul(E.A1, a.c0, b.c1);
    acc_mac(E.A1, a.c1, b.c0);
}
GL_DEV void ecacc_add(ecacc& C, const eacc& E) { cacc_add(C.A0, E.A0); cacc_add(C.A1, E.A1); }
GL_DEV ext_t ecacc_canon(const ecacc& C) { return ext_make(cacc_canon(C.A0), cacc_canon(C.A1)); }
GL_DEV ext_t eacc_weak(const eacc& E) { return ext_make(acc_weak(E.A0), acc_weak(E.A1)); }
GL_DEV ext_t eacc_canon(const eacc& E) { return ext_make(acc_canon(E.A0), acc_canon(E.A1)); }

// fixed right operand with 7*c1 precomputed (fold challenge r, alphas)
struct extmul_t {
    uint64_t c0, c1, c1_7;
};
GL_DEV extmul_t extmul_prep(ext_t b) {
    extmul_t m; m.c0 = gl_canon(b.c0); m.c1 = gl_canon(b.c1); m.c1_7 = gl_canon(gl_mul7_weak(m.c1)); return m;
}
GL_DEV void eacc_mac_prep(eacc& E, ext_t a, const extmul_t& b) {
    acc_mac(E.A0, a.c0, b.c0);
    acc_mac(E.A0, a.c1, b.c1_7);
    acc_mac(E.A1, a.c0, b.c1);
    acc_mac(E.A1, a.c1, b.c0);
}
// a * b for any u64 limbs; weak / canonical result
GL_DEV ext_t ext_mul_weak(ext_t a, ext_t b) {
    eacc E;
    eacc_mul(E, a, b, gl_mul7_weak(b.c1));
    return eacc_weak(E);
}
GL_DEV ext_t ext_mul(ext_t a, ext_t b) { return ext_canon(ext_mul_weak(a, b)); }
GL_DEV void eacc_mul_prep(eacc& E, ext_t a, const extmul_t& b) {
    acc_mul(E.A0, a.c0, b.c0);
    acc_mac(E.A0, a.c1, b.c1_7);
    acc_mul(E.A1, a.c0, b.c1);
    acc_mac(E.A1, a.c1, b.c0);
}
GL_DEV ext_t ext_mul_prep(ext_t a, const extmul_t& b) {
    eacc E;
    eacc_mul_prep(E, a, b);
    return eacc_canon(E);
}
GL_DEV ext_t ext_mul_prep_weak(ext_t a, const extmul_t& b) {
    eacc E;
    eacc_mul_prep(E, a, b);
    return eacc_weak(E);
}
GL_DEV ext_t ext_mul_base(ext_t a, uint64_t b) { return ext_make(gl_mul(a.c0, b), gl_mul(a.c1, b)); }
// x + d * r  (the fold), one reduction per limb; x canonical or not, result canonical
GL_DEV ext_t ext_fma_prep(ext_t x, ext_t d, const extmul_t& r) {
    eacc E;                       // x rides in the accumulator: no separate modular add
    acc_fma_first(E.A0, x.c0, d.c0, r.c0);
    acc_mac(E.A0, d.c1, r.c1_7);
    acc_fma_first(E.A1, x.c1, d.c0, r.c1);
    acc_mac(E.A1, d.c1, r.c0);
    return eacc_canon(E);
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------- memory helpers
// 16-byte and 32-byte vector accesses (LDG.E.128 / LDG.E.ENL2.256 on sm_100a).
GL_DEV ext_t ld_ext(const ext_t* p) {
    ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    return ext_make(v.x, v.y);
}
GL_DEV void st_ext(ext_t* p, ext_t v) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(v.c0, v.c1); }
GL_DEV void ld_ext2(const ext_t* p, ext_t& a, ext_t& b) {   // p 32-byte aligned
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a.c0), "=l"(a.c1), "=l"(b.c0), "=l"(b.c1) : "l"(p));
}
GL_DEV void st_ext2(ext_t* p, ext_t a, ext_t b) {           // p 32-byte aligned
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a.c0), "l"(a.c1), "l"(b.c0), "l"(b.c1) : "memory");
}

// ---------------------------------------------------------------- warp reduction
// Sum of the 32 lanes' values mod p with the hardware integer warp reduction (REDUX.SUM): the value is cut into four
// 16-bit limbs, each limb is summed over the warp in ONE instruction (32 * 65535 < 2^21, no overflow), and the four sums are
// recombined and reduced once.  Four independent REDUX instead of a five-step dependent chain of shuffles + modular adds;
// the result (canonical) is valid in every lane.  Any u64 input is accepted.
GL_DEV uint64_t shfl_down_u64(uint64_t v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
GL_DEV uint64_t warp_sum_gl(uint64_t v) {
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    const uint32_t s0 = __reduce_add_sync(0xffffffffu, lo & 0xFFFFu);
    const uint32_t s1 = __reduce_add_sync(0xffffffffu, lo >> 16);
    const uint32_t s2 = __reduce_add_sync(0xffffffffu, hi & 0xFFFFu);
    const uint32_t s3 = __reduce_add_sync(0xffffffffu, hi >> 16);
    const uint64_t a = (uint64_t)s0 + ((uint64_t)s1 << 16) + ((uint64_t)s2 << 32);   // < 2^54
    const uint64_t b = (uint64_t)(s3 & 0xFFFFu) << 48;
    const uint64_t t = a + b;
    const uint32_t top = (s3 >> 16) + (t < a ? 1u : 0u);                             // value = t + top 2^64, top < 64
    return gl_canon(gl_reduce_limbs((uint32_t)t, (uint32_t)(t >> 32), top, 0, 0));
}
GL_DEV ext_t warp_reduce_ext(ext_t v) { return ext_make(warp_sum_gl(v.c0), warp_sum_gl(v.c1)); }
#endif

// ---------------------------------------------------------------- stand-in transcript (device + host)
// The documented stand-in sponge (NOT Poseidon2): see include/ceno_b200.h cg_standin_*.
#if defined(__CUDACC__)
#define GL_HDC constexpr __host__ __device__ inline
#else
#define GL_HDC constexpr inline
#endif
GL_HDC uint64_t cg_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
GL_HD void cg_tr_absorb(uint64_t& h, uint64_t x) { h = cg_splitmix64(h ^ cg_splitmix64(x)); }
GL_HD uint64_t cg_tr_squeeze(uint64_t& h) {
    h = cg_splitmix64(h + 0xD1B54A32D192ED03ULL);
    return h >= GL_P ? h - GL_P : h;
}
GL_HD void cg_tr_append_message(uint64_t& h, const uint8_t* msg, uint64_t len) {
    cg_tr_absorb(h, 0x6D73670000000000ULL ^ len);
    for (uint64_t i = 0; i < len; i += 8) {
        uint64_t w = 0;
        for (uint64_t j = 0; j < 8 && i + j < len; j++) w |= (uint64_t)msg[i + j] << (8 * j);
        cg_tr_absorb(h, w);
    }
}
// One sumcheck round of the stand-in transcript — absorb the D evaluations, absorb the label "Internal round", squeeze the
// challenge — arranged for a single GPU lane: absorb(h, x) = mix(h ^ mix(x)), and the inner mix(x) of the 2D message words
// (independent of h, so they overlap) and of the three label words (compile-time constants) leave a serial chain of
// 2D + 3 + 2 mixes instead of 4D + 6 + 2.  Bit-identical to cg_tr_absorb / cg_tr_append_message / cg_tr_squeeze.
GL_HDC uint64_t cg_tr_pack8(const char* s, int n) {
    uint64_t w = 0;
    for (int j = 0; j < 8 && j < n; j++) w |= (uint64_t)(uint8_t)s[j] << (8 * j);
    return w;
}
template <int D>
GL_HD ext_t cg_tr_round(uint64_t& h, const ext_t (&res)[D]) {
    constexpr uint64_t L0 = cg_splitmix64(0x6D73670000000000ULL ^ 14ULL);
    constexpr uint64_t L1 = cg_splitmix64(cg_tr_pack8("Internal", 8));
    constexpr uint64_t L2 = cg_splitmix64(cg_tr_pack8(" round", 6));
    uint64_t sx[2 * D];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int x = 0; x < D; x++) { sx[2 * x] = cg_splitmix64(res[x].c0); sx[2 * x + 1] = cg_splitmix64(res[x].c1); }
    uint64_t g = h;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 2 * D; i++) g = cg_splitmix64(g ^ sx[i]);
    g = cg_splitmix64(g ^ L0);
    g = cg_splitmix64(g ^ L1);
    g = cg_splitmix64(g ^ L2);
    ext_t r;
    r.c0 = cg_tr_squeeze(g);
    r.c1 = cg_tr_squeeze(g);
    h = g;
    return r;
}
