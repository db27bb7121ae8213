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

__global__ void __launch_bounds__(256) ntt_pass_kernel(const __grid_constant__ NttPassArgs a) {
    __shared__ uint64_t tile[1 << CG_NTT_TILE_LOG];
    __shared__ uint64_t tws[1 << (CG_NTT_TILE_LOG - 1)];
    const uint32_t T = a.T, C = a.C, lo = a.lo;
    const uint32_t E_t = 1u << (T + C), cmask = (1u << C) - 1;
    const uint32_t tiles_per_col_log = a.log_n - T - C;
    const uint64_t tile_id = blockIdx.x;
    const uint64_t col = tile_id >> tiles_per_col_log;
    const uint64_t tau = tile_id & ((1ULL << tiles_per_col_log) - 1);
    const uint32_t ig_bits = lo - C;
    const uint64_t ig = tau & ((1ULL << ig_bits) - 1), outer = tau >> ig_bits;
    const uint64_t base = (outer << (lo + T)) | (ig << C);
    const uint32_t jr0 = (uint32_t)(ig << C);             // j_rest of column c = jr0 + c  (< 2^lo)
    const uint32_t sub_shift = CG_NTT_MAX_LOG - (lo + T);  // w_{2^(lo+T)}^e = w_{2^27}^(e << sub_shift)
    const uint32_t emask = (1u << CG_NTT_MAX_LOG) - 1;
    // digit twiddles w_{2^T}^(+-i), i < 2^(T-1)
    for (uint32_t i = threadIdx.x; i < (1u << T) / 2; i += blockDim.x) {
        uint32_t e = i << (CG_NTT_TILE_LOG - T);
        if (a.inverse) e = (4096u - e) & 4095u;
        tws[i] = __ldg(a.W12 + e);
    }
    const uint64_t* src = a.in + col * a.in_col_stride * a.estride;
    for (uint32_t e = threadIdx.x; e < E_t; e += blockDim.x) {
        const uint32_t t = e >> C, c = e & cmask;
        const uint64_t idx = base + ((uint64_t)t << lo) + c;
        uint64_t v = idx < a.in_len ? gl_canon(src[idx * a.estride]) : 0ULL;
        if (a.inverse && lo) {
            const uint32_t k = __brev(t) >> (32 - T);
            const uint32_t ex = (uint32_t)(((uint64_t)k * (jr0 + c)) << sub_shift) & emask;
            v = gl_mul(v, ntt_tw(a, ((1u << CG_NTT_MAX_LOG) - ex) & emask));
        }
        tile[e] = v;
    }
    __syncthreads();
    const uint32_t npairs = E_t >> 1;
    if (!a.inverse) {
        for (int s = (int)T - 1; s >= 0; s--) {
            for (uint32_t q = threadIdx.x; q < npairs; q += blockDim.x) {
                const uint32_t c = q & cmask, pt = q >> C;
                const uint32_t low = pt & ((1u << s) - 1), high = pt >> s;
                const uint32_t i0 = ((((high << 1) << s) | low) << C) | c, i1 = i0 + (1u << (s + C));
                const uint64_t x = tile[i0], y = tile[i1];
                tile[i0] = gl_add(x, y);
                tile[i1] = gl_mul(gl_sub(x, y), tws[low << (T - 1 - s)]);
            }
            __syncthreads();
        }
    } else {
        for (uint32_t s = 0; s < T; s++) {
            for (uint32_t q = threadIdx.x; q < npairs; q += blockDim.x) {
                const uint32_t c = q & cmask, pt = q >> C;
                const uint32_t low = pt & ((1u << s) - 1), high = pt >> s;
                const uint32_t i0 = ((((high << 1) << s) | low) << C) | c, i1 = i0 + (1u << (s + C));
                const uint64_t x = tile[i0], y = gl_mul(tile[i1], tws[low << (T - 1 - s)]);
                tile[i0] = gl_add(x, y);
                tile[i1] = gl_sub(x, y);
            }
            __syncthreads();
        }
    }
    uint64_t* dst = a.out + col * a.out_col_stride * a.estride;
    for (uint32_t e = threadIdx.x; e < E_t; e += blockDim.x) {
        const uint32_t t = e >> C, c = e & cmask;
        const uint64_t idx = base + ((uint64_t)t << lo) + c;
        uint64_t v = tile[e];
        if (!a.inverse && lo) {
            const uint32_t k = __brev(t) >> (32 - T);
            const uint32_t ex = (uint32_t)(((uint64_t)k * (jr0 + c)) << sub_shift) & emask;
            v = gl_mul(v, ntt_tw(a, ex));
        }
        if (a.scale) v = gl_mul(v, a.n_inv);
        dst[idx * a.estride] = v;
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
