// ntt_kernels.cuh — batched radix-2 NTT over Goldilocks for the Reed-Solomon step of the Basefold commitment
// (PCS::batch_commit encodes every witness column before the Merkle hash; EXTERNAL mpcs + p3-dft,
// call site ceno_zkvm/src/scheme/cpu/mod.rs:559-584; reference GPU path basefold.batch_commit_*,
// ceno_zkvm/src/scheme/gpu/mod.rs:1062-1509).
//
// Transform: X[k] = sum_j x[j] w^(jk),  w = two_adic_generator(log_n) = g^(2^(32-log_n)),  g = 7^((p-1)/2^32).
//
// B200 layout: the kernel is a multi-pass ("four-step", generalised) decimation in frequency.  log_n is split into
// digits of T <= 12 bits, top digit first.  One pass = one digit: a CTA stages a tile of 2^T x 2^C elements
// (2^T points of the digit, 2^C >= 8 neighbouring columns of the lower index bits so every global access is a
// 64..32768-byte contiguous run) in 32 KB of shared memory, runs the T butterfly stages there, and multiplies by the
// inter-digit twiddle w_{2^(lo+T)}^(k * j_rest) on the way out.  HBM traffic is 16 B per element per pass: 2 passes up
// to 2^21, 3 up to 2^27.  Small-transform outputs stay in bit-reversed position, which makes the overall output order
// exactly the bit reversal of k — the order a Basefold-style folding prover wants (adjacent pairs fold) — so no
// permutation pass is needed on the fast path; natural order costs one extra permutation kernel.
// The inverse runs the same passes backwards (decimation in time, inverse twiddles on the way in, 1/n on the way out).
// Twiddles: w_{2^27}^E = A[E & 8191] * B[E >> 13] (two L2-resident tables, 192 KB), digit twiddles from a 4096-entry table.
#pragma once
#include "gl64.cuh"

#define CG_NTT_MAX_LOG 27
#define CG_NTT_TILE_LOG 12
#define CG_NTT_A_BITS 13

struct NttPassArgs {
    const uint64_t* in;
    uint64_t* out;               // may alias `in`
    uint64_t in_col_stride, out_col_stride;   // in elements (of `estride` u64 each)
    uint64_t in_len;             // valid input elements per column; the rest reads as zero (RS zero padding)
    uint32_t estride;            // 1: base arrays; 2: one limb of an ext array
    uint32_t log_n, lo, T, C;    // this pass transforms index bits [lo, lo+T) with 2^C neighbours per tile
    int inverse;
    int scale;                   // inverse, last pass: multiply by n_inv
    uint64_t n_inv;
    const uint64_t* A;           // w_{2^27}^i, i < 2^13
    const uint64_t* B;           // w_{2^27}^(i << 13), i < 2^14
    const uint64_t* W12;         // w_{2^12}^i, i < 2^12
};

GL_DEV uint64_t ntt_tw(const NttPassArgs& a, uint32_t E) {   // w_{2^27}^E, E < 2^27
    return gl_mul(__ldg(a.A + (E & ((1u << CG_NTT_A_BITS) - 1))), __ldg(a.B + (E >> CG_NTT_A_BITS)));
}

// Digit geometry (T, C) is a template parameter: every index computation is a shift by a constant and every loop has a
// compile-time trip count (the run-time version spent most of its instructions on index arithmetic: 436 per element per pass).
// FORWARD = decimation in frequency (natural in, bit-reversed out), else decimation in time (bit-reversed in, natural out).
// When C == 0 the last two DIF stages (first two DIT stages) run in registers on 4 consecutive elements per thread
// (two 128-bit shared-memory accesses, conflict-free) instead of two more shared-memory round trips.
template <int T, int C, bool FORWARD>
__global__ void __launch_bounds__(256) ntt_pass_kernel(const __grid_constant__ NttPassArgs a) {
    constexpr uint32_t E_t = 1u << (T + C), cmask = (1u << C) - 1, NP = E_t >> 1;
    constexpr int REG2 = (C == 0 && T >= 2) ? 2 : 0;       // stages handled in registers
    __shared__ __align__(16) uint64_t tile[E_t];
    __shared__ uint64_t tws[(1u << T) / 2];
    const uint32_t lo = a.lo;
    const uint32_t tiles_per_col_log = a.log_n - T - C;
    const uint64_t tile_id = blockIdx.x;
    const uint64_t col = tile_id >> tiles_per_col_log;
    const uint64_t tau = tile_id & ((1ULL << tiles_per_col_log) - 1);
    const uint32_t ig_bits = lo - C;
    const uint64_t ig = tau & ((1ULL << ig_bits) - 1), outer = tau >> ig_bits;
    const uint64_t base = (outer << (lo + T)) | (ig << C);
    const uint32_t jr0 = (uint32_t)(ig << C);             // j_rest of column c = jr0 + c  (< 2^lo)
    const uint32_t sub_shift = CG_NTT_MAX_LOG - (lo + T);  // w_{2^(lo+T)}^e = w_{2^27}^(e << sub_shift)
    constexpr uint32_t emask = (1u << CG_NTT_MAX_LOG) - 1;
    // digit twiddles w_{2^T}^(+-i), i < 2^(T-1)
    for (uint32_t i = threadIdx.x; i < (1u << T) / 2; i += 256) {
        uint32_t e = i << (CG_NTT_TILE_LOG - T);
        if (!FORWARD) e = (4096u - e) & 4095u;
        tws[i] = __ldg(a.W12 + e);
    }
    const uint64_t* src = a.in + col * a.in_col_stride * a.estride;
#pragma unroll 4
    for (uint32_t e0 = 0; e0 < E_t; e0 += 256) {
        const uint32_t e = e0 + threadIdx.x;
        if (E_t >= 256 || e < E_t) {
            const uint32_t t = e >> C, c = e & cmask;
            const uint64_t idx = base + ((uint64_t)t << lo) + c;
            uint64_t v = idx < a.in_len ? gl_canon(src[idx * a.estride]) : 0ULL;
            if (!FORWARD && lo) {
                const uint32_t k = __brev(t) >> (32 - T);
                const uint32_t ex = (uint32_t)(((uint64_t)k * (jr0 + c)) << sub_shift) & emask;
                v = gl_mul(v, ntt_tw(a, ((1u << CG_NTT_MAX_LOG) - ex) & emask));
            }
            tile[e] = v;
        }
    }
    __syncthreads();
    if (FORWARD) {
#pragma unroll
        for (int s = T - 1; s >= REG2; s--) {
#pragma unroll 2
            for (uint32_t q0 = 0; q0 < NP; q0 += 256) {
                const uint32_t q = q0 + threadIdx.x;
                if (NP >= 256 || q < NP) {
                    const uint32_t c = q & cmask, pt = q >> C;
                    const uint32_t low = pt & ((1u << s) - 1), high = pt >> s;
                    const uint32_t i0 = ((((high << 1) << s) | low) << C) | c, i1 = i0 + (1u << (s + C));
                    const uint64_t x = tile[i0], y = tile[i1];
                    tile[i0] = gl_add(x, y);
                    tile[i1] = gl_mul(gl_sub(x, y), tws[low << (T - 1 - s)]);
                }
            }
            __syncthreads();
        }
        if (REG2) {   // stages 1 and 0 on elements 4q .. 4q+3:  stage 1 pairs (0,2),(1,3) with w^(low << (T-2)), stage 0 pairs (0,1),(2,3)
            const uint64_t w1 = tws[1u << (T - 2)];   // w_4^1 for T >= 2 (low = 1 at stage 1)
#pragma unroll 2
            for (uint32_t q0 = 0; q0 < E_t / 4; q0 += 256) {
                const uint32_t q = q0 + threadIdx.x;
                if (E_t / 4 >= 256 || q < E_t / 4) {
                    ulonglong2* p = reinterpret_cast<ulonglong2*>(tile + 4 * q);
                    const ulonglong2 u = p[0], v = p[1];
                    const uint64_t a0 = gl_add(u.x, v.x), a2 = gl_sub(u.x, v.x);
                    const uint64_t a1 = gl_add(u.y, v.y), a3 = gl_mul(gl_sub(u.y, v.y), w1);
                    p[0] = make_ulonglong2(gl_add(a0, a1), gl_sub(a0, a1));
                    p[1] = make_ulonglong2(gl_add(a2, a3), gl_sub(a2, a3));
                }
            }
            __syncthreads();
        }
    } else {
        if (REG2) {   // DIT stages 0 and 1 on elements 4q .. 4q+3 with the inverse twiddles
            const uint64_t w1 = tws[1u << (T - 2)];
#pragma unroll 2
            for (uint32_t q0 = 0; q0 < E_t / 4; q0 += 256) {
                const uint32_t q = q0 + threadIdx.x;
                if (E_t / 4 >= 256 || q < E_t / 4) {
                    ulonglong2* p = reinterpret_cast<ulonglong2*>(tile + 4 * q);
                    const ulonglong2 u = p[0], v = p[1];
                    const uint64_t a0 = gl_add(u.x, u.y), a1 = gl_sub(u.x, u.y);
                    const uint64_t a2 = gl_add(v.x, v.y), a3 = gl_mul(gl_sub(v.x, v.y), w1);
                    p[0] = make_ulonglong2(gl_add(a0, a2), gl_add(a1, a3));
                    p[1] = make_ulonglong2(gl_sub(a0, a2), gl_sub(a1, a3));
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int s = REG2; s < T; s++) {
#pragma unroll 2
            for (uint32_t q0 = 0; q0 < NP; q0 += 256) {
                const uint32_t q = q0 + threadIdx.x;
                if (NP >= 256 || q < NP) {
                    const uint32_t c = q & cmask, pt = q >> C;
                    const uint32_t low = pt & ((1u << s) - 1), high = pt >> s;
                    const uint32_t i0 = ((((high << 1) << s) | low) << C) | c, i1 = i0 + (1u << (s + C));
                    const uint64_t x = tile[i0], y = gl_mul(tile[i1], tws[low << (T - 1 - s)]);
                    tile[i0] = gl_add(x, y);
                    tile[i1] = gl_sub(x, y);
                }
            }
            __syncthreads();
        }
    }
    uint64_t* dst = a.out + col * a.out_col_stride * a.estride;
#pragma unroll 4
    for (uint32_t e0 = 0; e0 < E_t; e0 += 256) {
        const uint32_t e = e0 + threadIdx.x;
        if (E_t >= 256 || e < E_t) {
            const uint32_t t = e >> C, c = e & cmask;
            const uint64_t idx = base + ((uint64_t)t << lo) + c;
            uint64_t v = tile[e];
            if (FORWARD && lo) {
                const uint32_t k = __brev(t) >> (32 - T);
                const uint32_t ex = (uint32_t)(((uint64_t)k * (jr0 + c)) << sub_shift) & emask;
                v = gl_mul(v, ntt_tw(a, ex));
            }
            if (a.scale) v = gl_mul(v, a.n_inv);
            dst[idx * a.estride] = v;
        }
    }
}

// tables: A[i] = W^i, B[i] = W^(i << 13), W12[i] = W^(i << 15)  with W = w_{2^27} = g^(2^5)
__global__ void ntt_tables_kernel(uint64_t* A, uint64_t* B, uint64_t* W12, uint64_t omega27) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    auto pw = [&](uint64_t e) {
        uint64_t r = 1, b = omega27;
        while (e) { if (e & 1) r = gl_mul(r, b); b = gl_mul(b, b); e >>= 1; }
        return r;
    };
    if (i < (1u << CG_NTT_A_BITS)) A[i] = pw(i);
    if (i < (1u << (CG_NTT_MAX_LOG - CG_NTT_A_BITS))) B[i] = pw((uint64_t)i << CG_NTT_A_BITS);
    if (i < 4096u) W12[i] = pw((uint64_t)i << (CG_NTT_MAX_LOG - 12));
}

// in-place bit-reversal permutation of every column (natural <-> bit-reversed order)
__global__ void __launch_bounds__(256) ntt_bitrev_kernel(uint64_t* data, uint32_t log_n, uint64_t n_cols, uint64_t col_stride, uint32_t estride) {
    const uint64_t n = 1ULL << log_n, total = n * n_cols, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const uint64_t col = g >> log_n, i = g & (n - 1);
        const uint64_t j = log_n ? (__brevll(i) >> (64 - log_n)) : 0;
        if (i < j) {
            uint64_t* p = data + col * col_stride * estride;
            const uint64_t x = p[i * estride], y = p[j * estride];
            p[i * estride] = y;
            p[j * estride] = x;
        }
    }
}
