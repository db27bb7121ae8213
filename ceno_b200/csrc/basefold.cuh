// basefold.cuh — Basefold PCS on the device: commit (RS-encode + Merkle) and batch_open (f-2, a9).
//
// Call sites replaced: TraceCommitter::commit_traces -> PCS::batch_commit (ceno_zkvm/src/scheme/cpu/mod.rs:559-584; GPU
// basefold.batch_commit_*, ceno_zkvm/src/scheme/gpu/mod.rs:1062-1509) and OpeningProver::open -> PCS::batch_open
// (ceno_zkvm/src/scheme/cpu/mod.rs:1415-1457; GPU ceno_zkvm/src/scheme/gpu/mod.rs:3324-3413).  The protocol is the one the
// in-tree verifier restatement pins (ceno_recursion_v2/src/pcs/mod.rs:1111-1317 replay_basefold, :7494-7727 query checks,
// :7765-7781 fold rule, :444-592 final claim, :138-145 basecode_log == 0):
//   batch coefficients 1, a, a^2, ... over every committed column; running codeword = RLC of the codewords, smaller codewords
//   join when the folded codeword reaches their height; per round a degree-2 sumcheck message [p(1), p(2)] of
//   sum_p eq(point_p, .) g_p(.), g_p = RLC of the opening's columns, the challenge, then the Merkle commitment of the
//   current codeword as (even, odd) pairs; fold  lo = (a+b)/2, hi = (a-b) g_h^{-bitrev(i)}/2, lo + r (hi - lo); the final
//   message is one element per opening; queries open one row per input commitment and one sibling per round.
// Kernels here: column RLC (codewords and evaluation vectors), the codeword fold, the fold twiddle table and the query gather;
// RS-encode, Poseidon2 leaf/compress kernels and the sumcheck come from the rest of the library.
#pragma once

// out[i] (+)= sum_j coeff[j] * cols[j * stride + i]: base-field columns, ext coefficients, one lazy reduction per limb
__global__ void __launch_bounds__(CG_THREADS) bf_rlc_cols_kernel(const uint64_t* __restrict__ cols, uint64_t stride, uint32_t width,
                                                                 const ext_t* __restrict__ coeff, uint64_t n, ext_t* __restrict__ out, int accumulate) {
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        acc_t A0, A1;
        acc_zero(A0); acc_zero(A1);
        if (accumulate) { const ext_t o = ld_ext(out + i); acc_set64(A0, o.c0); acc_set64(A1, o.c1); }
        for (uint32_t j = 0; j < width; j++) {
            const uint64_t v = cols[(uint64_t)j * stride + i];
            const ext_t cj = coeff[j];
            acc_mac(A0, cj.c0, v);
            acc_mac(A1, cj.c1, v);
        }
        st_ext(out + i, ext_make(acc_canon(A0), acc_canon(A1)));
    }
}
// tw[i] = g_inv^{bitrev(i, bits)} / 2 for i < 2^bits (bits = log2 height - 1 of the tallest codeword); the table of a lower
// codeword is a prefix of it (bitrev_{b+1}(i) = 2 bitrev_b(i) for i < 2^b and g_{h-1} = g_h^2)
struct BfTwArgs {
    uint64_t pw[32];   // g_inv^(2^b)
    uint64_t inv2;
    uint32_t bits;
    uint64_t* tw;
};
__global__ void __launch_bounds__(CG_THREADS) bf_twiddle_kernel(const __grid_constant__ BfTwArgs a) {
    const uint64_t n = 1ULL << a.bits, step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        uint64_t v = a.inv2;
        for (uint32_t b = 0; b < a.bits; b++)
            if ((i >> b) & 1) v = gl_mul(v, a.pw[a.bits - 1 - b]);   // bit b of i is bit (bits-1-b) of bitrev(i)
        a.tw[i] = v;
    }
}
// fold_codeword_pair (ceno_recursion_v2/src/pcs/mod.rs:7771-7781) over a whole codeword
__global__ void __launch_bounds__(CG_THREADS) bf_fold_kernel(const ext_t* __restrict__ in, uint64_t n_out, ext_t r, const uint64_t* __restrict__ tw,
                                                             ext_t* __restrict__ out) {
    const extmul_t rm = extmul_prep(r);
    const uint64_t inv2 = 0x7FFFFFFF80000001ULL;   // (p + 1) / 2
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += step) {
        ext_t a, b;
        ld_ext2(in + 2 * i, a, b);
        const ext_t lo = ext_mul_base(ext_add(a, b), inv2);
        const ext_t hi = ext_mul_base(ext_sub(a, b), tw[i]);
        st_ext(out + i, ext_fma_prep(lo, ext_sub(hi, lo), rm));
    }
}
// query gather: one descriptor per contiguous or strided run of u64 words
struct BfCopy {
    const uint64_t* src;
    uint64_t dst;      // word offset into the proof buffer
    uint64_t stride;   // words between consecutive source words
    uint64_t n;
};
__global__ void bf_gather_kernel(const BfCopy* __restrict__ list, uint64_t n_entries, uint64_t* __restrict__ proof) {
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entries; e += step) {
        const BfCopy c = list[e];
        for (uint64_t w = 0; w < c.n; w++) proof[c.dst + w] = c.src[w * c.stride];
    }
}
