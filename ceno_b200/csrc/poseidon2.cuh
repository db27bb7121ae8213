// poseidon2.cuh — Poseidon2 over Goldilocks (width 8, x^7, 8 external + 22 internal rounds), the
// padding-free sponge leaf hash and the 2-to-1 compression of the Merkle commitment (SURVEY.md §8 a9:
// TraceCommitter::commit_traces -> PCS::batch_commit, ceno_zkvm/src/scheme/cpu/mod.rs:559-584).
//
// PARITY UNPINNED: the round constants, the internal diagonal and Basefold's leaf arrangement are defined in
// un-vendored crates (p3-goldilocks 0.4.3, gkr-backend `poseidon` / `mpcs`).  They are
// parameters here (cg_poseidon2_set_params); the structure follows the published Poseidon2 construction as
// Plonky3 instantiates it for Goldilocks:
//   external layer: M4 on each 4-chunk, then every lane += sum of the lanes at the same chunk offset;
//   internal layer: s = sum(state); state[i] = state[i] * diag[i] + s;
//   permutation   : external, 4 x (rc, sbox all, external), 22 x (rc0, sbox lane 0, internal), 4 x (...).
// Leaf = PaddingFreeSponge<8, rate 4, out 4> over one matrix row (inputs overwrite the rate lanes);
// node = TruncatedPermutation<2, 4, 8>.
//
// One thread per permutation; all arithmetic through the lazy accumulators of gl64.cuh (one reduction per
// linear combination).  The leaf kernel reads a COLUMN-major matrix (what the reference keeps on the device
// after matrix_transpose, ceno_zkvm/src/scheme/gpu/mod.rs:933-989), so consecutive threads read consecutive
// addresses of every column.
#pragma once
#include "gl64.cuh"

struct P2Params {
    uint64_t ext_rc[8][8];
    uint64_t int_rc[22];
    uint64_t diag[8];
    uint32_t mds_variant;   // 0: circ(2,3,1,1)   1: Horizen-Labs M4
    uint32_t pad;
};

// Values between the steps are "weak" (any u64 congruent to the field element): multiplicands and summands need no
// canonical form; the permutation canonicalises its 8 outputs once at the end.
GL_DEV uint64_t p2_sbox(uint64_t x) {
    const uint64_t x2 = gl_mul_weak(x, x), x3 = gl_mul_weak(x2, x), x4 = gl_mul_weak(x2, x2);
    return gl_mul_weak(x4, x3);
}
// one 4-chunk of the external layer as unreduced 96-bit sums
//   variant 0, circ(2,3,1,1):  o_i = S + x_i + 2 x_{i+1},  S = x0+x1+x2+x3           (additions only)
//   variant 1, Horizen-Labs M4 [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]]: small-constant multiples by repeated addition
GL_DEV void p2_m4_wide(const uint64_t* x, uint32_t variant, wsum_t* o) {
    if (variant == 0) {
        wsum_t S;
        wsum_set(S, x[0]); wsum_add(S, x[1]); wsum_add(S, x[2]); wsum_add(S, x[3]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            o[i] = S;
            wsum_add(o[i], x[i]);
            wsum_add(o[i], x[(i + 1) & 3]);
            wsum_add(o[i], x[(i + 1) & 3]);
        }
    } else {
        const uint32_t M1[4][4] = {{5, 7, 1, 3}, {4, 6, 1, 1}, {1, 3, 5, 7}, {1, 1, 4, 6}};
        wsum_t m[4][3];   // x_j, 2 x_j, 4 x_j
#pragma unroll
        for (int j = 0; j < 4; j++) {
            wsum_set(m[j][0], x[j]);
            m[j][1] = m[j][0]; wsum_addw(m[j][1], m[j][0]);
            m[j][2] = m[j][1]; wsum_addw(m[j][2], m[j][1]);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            wsum_set(o[i], 0);
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int b = 0; b < 3; b++)
                    if ((M1[i][j] >> b) & 1) wsum_addw(o[i], m[j][b]);
        }
    }
}
// external layer: M4 on each 4-chunk, then lane i of chunk c += o[0][i] + o[1][i]; `rc` (optional) = the NEXT round's
// constants, added inside the same sums (one reduction per lane for linear layer + constant)
GL_DEV void p2_external(uint64_t* s, uint32_t variant, const uint64_t* rc) {
    wsum_t o0[4], o1[4];
    p2_m4_wide(s, variant, o0);
    p2_m4_wide(s + 4, variant, o1);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        wsum_t a = o0[i], b = o1[i];
        wsum_addw(a, o0[i]); wsum_addw(a, o1[i]);     // 2 o0 + o1
        wsum_addw(b, o1[i]); wsum_addw(b, o0[i]);     // o0 + 2 o1
        if (rc) { wsum_add(a, rc[i]); wsum_add(b, rc[4 + i]); }
        s[i] = wsum_weak(a);
        s[4 + i] = wsum_weak(b);
    }
}
// any u64 in, canonical out
GL_DEV void p2_permute(const P2Params& p, uint64_t* s) {
    p2_external(s, p.mds_variant, p.ext_rc[0]);
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = p2_sbox(s[i]);
        p2_external(s, p.mds_variant, r < 3 ? p.ext_rc[r + 1] : nullptr);
    }
    for (int r = 0; r < 22; r++) {
        wsum_t t;
        wsum_set(t, s[0]);
        wsum_add(t, p.int_rc[r]);
        s[0] = p2_sbox(wsum_weak(t));
        wsum_t S;
        wsum_set(S, s[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) wsum_add(S, s[i]);
        const uint64_t sum = wsum_weak(S);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            acc_t A;
            acc_fma_first(A, sum, s[i], p.diag[i]);
            s[i] = acc_weak(A);
        }
    }
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            wsum_t t;
            wsum_set(t, s[i]);
            wsum_add(t, p.ext_rc[r][i]);
            s[i] = p2_sbox(wsum_weak(t));
        }
        p2_external(s, p.mds_variant, nullptr);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = gl_canon(s[i]);
}

#if defined(__CUDACC__)
GL_DEV void p2_load_params(P2Params& sp, const P2Params* gp) {
    uint64_t* d = reinterpret_cast<uint64_t*>(&sp);
    const uint64_t* g = reinterpret_cast<const uint64_t*>(gp);
    for (int i = threadIdx.x; i < (int)(sizeof(P2Params) / 8); i += blockDim.x) d[i] = g[i];
    __syncthreads();
}
__global__ void __launch_bounds__(128) p2_permute_kernel(const P2Params* __restrict__ gp, uint64_t* __restrict__ states, uint64_t n) {
    __shared__ P2Params sp;
    p2_load_params(sp, gp);
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s[8];
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = gl_canon(states[8 * i + k]);
    p2_permute(sp, s);
#pragma unroll
    for (int k = 0; k < 8; k++) states[8 * i + k] = s[k];
}
// leaf digests: one thread per row; matrix element (row i, column c) at  col_major ? m[c*height + i] : m[i*width + c]
__global__ void __launch_bounds__(128) p2_leaf_kernel(const P2Params* __restrict__ gp, const uint64_t* __restrict__ m, uint64_t width,
                                                      uint64_t height, int col_major, uint64_t* __restrict__ digests) {
    __shared__ P2Params sp;
    p2_load_params(sp, gp);
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= height) return;
    uint64_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint64_t c = 0; c < width; c += 4) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (c + k < width) s[k] = gl_canon(col_major ? m[(c + k) * height + i] : m[i * width + c + k]);
        p2_permute(sp, s);
    }
    *reinterpret_cast<ulonglong4*>(digests + 4 * i) = make_ulonglong4(s[0], s[1], s[2], s[3]);
}
__global__ void __launch_bounds__(128) p2_compress_kernel(const P2Params* __restrict__ gp, const uint64_t* __restrict__ in, uint64_t n_out,
                                                          uint64_t* __restrict__ out) {
    __shared__ P2Params sp;
    p2_load_params(sp, gp);
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const ulonglong4 l = *reinterpret_cast<const ulonglong4*>(in + 8 * i), r = *reinterpret_cast<const ulonglong4*>(in + 8 * i + 4);
    uint64_t s[8] = {l.x, l.y, l.z, l.w, r.x, r.y, r.z, r.w};
    p2_permute(sp, s);
    *reinterpret_cast<ulonglong4*>(out + 4 * i) = make_ulonglong4(s[0], s[1], s[2], s[3]);
}
#endif
