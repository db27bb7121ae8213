// gl64.cuh — branch-free Goldilocks (p = 2^64 - 2^32 + 1) and its quadratic extension
// F_p[X]/(X^2 - 7) in registers, for sm_100a.
//
// Restates the field of p3-goldilocks =0.4.3 / ff_ext::GoldilocksExt2 (reference Cargo.toml:34,
// W = 7).  Representation: canonical u64 at every function boundary that leaves
// the device; inside product chains values may be any u64 (the wide product + reduction is
// correct for non-canonical operands).
//
// Identities used:  2^64 = 2^32 - 1 =: EPS,  2^96 = -1,  2^128 = -2^32   (mod p).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL

#define GL_DEV __device__ __forceinline__

struct __align__(16) ext_t {
    uint64_t c0, c1;
};

// ---------------------------------------------------------------- base field
GL_DEV uint64_t gl_canon(uint64_t x) { return x >= GL_P ? x - GL_P : x; }

// a, b canonical -> canonical
GL_DEV uint64_t gl_add(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    // overflow past 2^64, or landing in [p, 2^64): both are fixed by adding EPS (== subtracting p mod 2^64)
    return ((s < a) | (s >= GL_P)) ? s + GL_EPS : s;
}
GL_DEV uint64_t gl_sub(uint64_t a, uint64_t b) {
    uint64_t d = a - b;
    return (a < b) ? d - GL_EPS : d;
}
GL_DEV uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }

// x = lo + hi 2^64 + top 2^128  (top < 2^31)  ->  canonical residue
GL_DEV uint64_t gl_reduce160(uint64_t lo, uint64_t hi, uint64_t top) {
    uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t sub = hi_hi + (top << 32);      // hi_hi 2^96 = -hi_hi ; top 2^128 = -top 2^32
    uint64_t t0 = lo - sub;
    if (lo < sub) t0 -= GL_EPS;              // wrapped by 2^64 = EPS
    uint64_t t1 = hi_lo * GL_EPS;            // hi_lo 2^64 = hi_lo EPS  (< 2^64)
    uint64_t r = t0 + t1;
    if (r < t1) r += GL_EPS;
    return gl_canon(r);
}
GL_DEV uint64_t gl_reduce128(uint64_t lo, uint64_t hi) { return gl_reduce160(lo, hi, 0); }

// any u64 operands -> canonical
GL_DEV uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128(a * b, __umul64hi(a, b)); }

// a0*b0 + a1*b1 with ONE reduction (129-bit intermediate)
GL_DEV uint64_t gl_dot2(uint64_t a0, uint64_t b0, uint64_t a1, uint64_t b1) {
    uint64_t l0 = a0 * b0, h0 = __umul64hi(a0, b0);
    uint64_t l1 = a1 * b1, h1 = __umul64hi(a1, b1);
    uint64_t lo = l0 + l1;
    uint64_t c = lo < l0;
    uint64_t hi = h0 + h1;
    uint64_t top = hi < h0;
    hi += c;
    top += (hi < c);
    return gl_reduce160(lo, hi, top);
}

// 7 * a  (a canonical) -> canonical.  7a < 2^67: 7a = lo + t 2^64, t < 8 -> lo + t EPS.
GL_DEV uint64_t gl_mul7(uint64_t a) {
    uint64_t lo = a * 7ULL, t = __umul64hi(a, 7ULL);
    uint64_t add = t * GL_EPS;               // < 2^35
    uint64_t r = lo + add;
    if (r < add) r += GL_EPS;
    return gl_canon(r);
}

// ---------------------------------------------------------------- extension
GL_DEV ext_t ext_make(uint64_t a, uint64_t b) { ext_t r; r.c0 = a; r.c1 = b; return r; }
GL_DEV ext_t ext_zero() { return ext_make(0, 0); }
GL_DEV ext_t ext_one() { return ext_make(1, 0); }
GL_DEV ext_t ext_add(ext_t a, ext_t b) { return ext_make(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
GL_DEV ext_t ext_sub(ext_t a, ext_t b) { return ext_make(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
GL_DEV ext_t ext_canon(ext_t a) { return ext_make(gl_canon(a.c0), gl_canon(a.c1)); }

// (a0 + a1 X)(b0 + b1 X) = a0 b0 + 7 a1 b1 + (a0 b1 + a1 b0) X ; 4 wide products, 2 reductions
GL_DEV ext_t ext_mul(ext_t a, ext_t b) {
    uint64_t b1_7 = gl_mul7(gl_canon(b.c1));
    return ext_make(gl_dot2(a.c0, b.c0, a.c1, b1_7), gl_dot2(a.c0, b.c1, a.c1, b.c0));
}
// multiplier with 7*c1 precomputed, for a fixed right operand (the fold challenge r)
struct extmul_t {
    uint64_t c0, c1, c1_7;
};
GL_DEV extmul_t extmul_prep(ext_t b) {
    extmul_t m; m.c0 = gl_canon(b.c0); m.c1 = gl_canon(b.c1); m.c1_7 = gl_mul7(m.c1); return m;
}
GL_DEV ext_t ext_mul_prep(ext_t a, const extmul_t& b) {
    return ext_make(gl_dot2(a.c0, b.c0, a.c1, b.c1_7), gl_dot2(a.c0, b.c1, a.c1, b.c0));
}
GL_DEV ext_t ext_mul_base(ext_t a, uint64_t b) { return ext_make(gl_mul(a.c0, b), gl_mul(a.c1, b)); }

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

// ---------------------------------------------------------------- warp / block reduction
GL_DEV uint64_t shfl_down_u64(uint64_t v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
GL_DEV ext_t warp_reduce_ext(ext_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ext_t o = ext_make(shfl_down_u64(v.c0, d), shfl_down_u64(v.c1, d));
        v = ext_add(v, o);
    }
    return v;
}

// ---------------------------------------------------------------- stand-in transcript (device + host)
// The documented stand-in sponge (NOT Poseidon2): see include/ceno_b200.h cg_standin_*.
__host__ __device__ inline uint64_t cg_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__host__ __device__ inline void cg_tr_absorb(uint64_t& h, uint64_t x) { h = cg_splitmix64(h ^ cg_splitmix64(x)); }
__host__ __device__ inline uint64_t cg_tr_squeeze(uint64_t& h) {
    h = cg_splitmix64(h + 0xD1B54A32D192ED03ULL);
    return h >= GL_P ? h - GL_P : h;
}
__host__ __device__ inline void cg_tr_append_message(uint64_t& h, const uint8_t* msg, uint64_t len) {
    cg_tr_absorb(h, 0x6D73670000000000ULL ^ len);
    for (uint64_t i = 0; i < len; i += 8) {
        uint64_t w = 0;
        for (uint64_t j = 0; j < 8 && i + j < len; j++) w |= (uint64_t)msg[i + j] << (8 * j);
        cg_tr_absorb(h, w);
    }
}
