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

GL_DEV uint64_t p2_sbox(uint64_t x) {
    const uint64_t x2 = gl_mul_weak(x, x), x3 = gl_mul_weak(x2, x), x4 = gl_mul_weak(x2, x2);
    return gl_mul(x4, x3);
}
GL_DEV void p2_m4(uint64_t* x, uint32_t variant) {
    const uint32_t M0[4][4] = {{2, 3, 1, 1}, {1, 2, 3, 1}, {1, 1, 2, 3}, {3, 1, 1, 2}};
    const uint32_t M1[4][4] = {{5, 7, 1, 3}, {4, 6, 1, 1}, {1, 3, 5, 7}, {1, 1, 4, 6}};
    uint64_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        acc_t A;
        acc_zero(A);
#pragma unroll
        for (int j = 0; j < 4; j++) acc_mac(A, variant ? M1[i][j] : M0[i][j], x[j]);
        o[i] = acc_canon(A);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = o[i];
}
GL_DEV void p2_external(uint64_t* s, uint32_t variant) {
    p2_m4(s, variant);
    p2_m4(s + 4, variant);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint64_t t = gl_add(s[i], s[4 + i]);
        s[i] = gl_add(s[i], t);
        s[4 + i] = gl_add(s[4 + i], t);
    }
}
// state canonical in, canonical out
GL_DEV void p2_permute(const P2Params& p, uint64_t* s) {
    p2_external(s, p.mds_variant);
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = p2_sbox(gl_add(s[i], p.ext_rc[r][i]));
        p2_external(s, p.mds_variant);
    }
    for (int r = 0; r < 22; r++) {
        s[0] = p2_sbox(gl_add(s[0], p.int_rc[r]));
        acc_t S;
        acc_zero(S);
#pragma unroll
        for (int i = 0; i < 8; i++) acc_mac(S, 1ULL, s[i]);
        const uint64_t sum = acc_canon(S);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            acc_t A;
            acc_set64(A, sum);
            acc_mac(A, s[i], p.diag[i]);
            s[i] = acc_canon(A);
        }
    }
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = p2_sbox(gl_add(s[i], p.ext_rc[r][i]));
        p2_external(s, p.mds_variant);
    }
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
