// Host build of ceno_b200/csrc/gl64.cuh (portable carry emulation) so the ALGEBRA of the device
// sequences (weak reduction, lazy accumulation, subtraction-only evaluation points) can be checked
// against big-int arithmetic without a GPU.  The PTX transcription itself is covered by the -m gpu tests.
#include "../../ceno_b200/csrc/gl64.cuh"
#include "../../ceno_b200/csrc/poseidon2.cuh"
extern "C" {
void h_p2_permute(const P2Params* p, uint64_t* st) { uint64_t s[8]; for (int i = 0; i < 8; i++) s[i] = gl_canon(st[i]); p2_permute(*p, s); for (int i = 0; i < 8; i++) st[i] = s[i]; }
uint64_t h_reduce_weak(uint64_t s0, uint64_t s1, uint32_t s2) { return acc_reduce_weak(s0, s1, s2); }
// compact accumulation of n products a_i*b_i through the aligned-limb accumulator
uint64_t h_cacc_dot(const uint64_t* a, const uint64_t* b, uint32_t n, uint32_t per) {
    cacc_t C; cacc_zero(C);
    for (uint32_t i = 0; i < n; i += per) {
        acc_t A; acc_zero(A);
        for (uint32_t j = i; j < n && j < i + per; j++) acc_mac(A, a[j], b[j]);
        cacc_add(C, A);
    }
    return cacc_canon(C);
}
uint64_t h_canon(uint64_t x) { return gl_canon(x); }
uint64_t h_sub(uint64_t a, uint64_t b) { return gl_sub(a, b); }
uint64_t h_add(uint64_t a, uint64_t b) { return gl_add(a, b); }
uint64_t h_mul(uint64_t a, uint64_t b) { return gl_mul(a, b); }
uint64_t h_mul7_weak(uint64_t a) { return gl_mul7_weak(a); }
void h_ext_mul(const uint64_t* a, const uint64_t* b, uint64_t* o) { ext_t r = ext_mul(ext_make(a[0], a[1]), ext_make(b[0], b[1])); o[0] = r.c0; o[1] = r.c1; }
void h_ext_fma(const uint64_t* x, const uint64_t* d, const uint64_t* r, uint64_t* o) {
    ext_t v = ext_fma_prep(ext_make(x[0], x[1]), ext_make(d[0], d[1]), extmul_prep(ext_make(r[0], r[1])));
    o[0] = v.c0; o[1] = v.c1;
}
// sum_i a_i * b_i accumulated lazily, reduced once
void h_eacc_dot(const uint64_t* a, const uint64_t* b, uint32_t n, uint64_t* o) {
    eacc E; eacc_zero(E);
    for (uint32_t i = 0; i < n; i++) {
        ext_t bb = ext_make(b[2 * i], b[2 * i + 1]);
        eacc_mac(E, ext_make(a[2 * i], a[2 * i + 1]), bb, gl_mul7_weak(bb.c1));
    }
    ext_t v = eacc_canon(E); o[0] = v.c0; o[1] = v.c1;
}
}
