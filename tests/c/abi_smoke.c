/* A plain-C consumer of include/ceno_b200.h — what a Rust `extern "C"` binding sees.
 * Links only libceno_b200.so (+ the CPU oracle as the checker).  Builds eq(w,.), proves the degree-3
 * sumcheck eq*A*B with the transcript behind a C callback, and compares every output with the oracle.
 * Exit code 0 = bit-exact. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ceno_b200.h"

/* oracle (test infrastructure) */
typedef struct { const uint64_t* data; uint32_t num_vars; uint32_t is_ext; } or_mle;
typedef struct { uint64_t h; } or_transcript;
void or_fill_ext(uint64_t seed, uint64_t n, uint64_t* out);
void or_build_eq_x_r_vec(const uint64_t* r, uint32_t k, uint64_t* out);
void or_tr_init(or_transcript* t, const uint8_t* label, uint64_t len);
void or_tr_append_message(or_transcript* t, const uint8_t* msg, uint64_t len);
int or_sumcheck_prove_standin(const or_mle* mles, uint32_t n_mles, const uint64_t* term_coeff, const uint32_t* term_off,
                              const uint32_t* term_idx, uint32_t n_terms, uint32_t num_vars, uint32_t degree, or_transcript* tr,
                              uint64_t* round_evals, uint64_t* final_evals, uint64_t* challenges);

#define CK(x) do { int rc__ = (x); if (rc__) { fprintf(stderr, "%s -> %d: %s\n", #x, rc__, ctx ? cg_last_error(ctx) : "?"); return 2; } } while (0)

/* the "Rust transcript": absorb the round message, sample the challenge (stand-in sponge) */
static void my_challenge(void* user, uint32_t round, const uint64_t* evals, uint32_t degree, uint64_t out_r[2]) {
    (void)round;
    cg_standin_append_ext((uint64_t*)user, evals, degree);
    cg_standin_sample((uint64_t*)user, "Internal round", out_r);
}

int main(void) {
    const uint32_t k = 14, deg = 3;
    const uint64_t n = 1ULL << k;
    cg_ctx* ctx = NULL;
    int rc = cg_init(0, &ctx);
    if (rc == CG_ERR_NO_DEVICE) { printf("no sm_100 device: cg_init refuses (no CPU fallback)\n"); return 77; }
    CK(rc);
    uint64_t* w = malloc(16 * k); uint64_t* a = malloc(16 * n); uint64_t* b = malloc(16 * n); uint64_t* eq = malloc(16 * n);
    or_fill_ext(0xE9, k, w); or_fill_ext(1, n, a); or_fill_ext(2, n, b); or_build_eq_x_r_vec(w, k, eq);
    void *d_a, *d_b, *d_eq;
    CK(cg_alloc(ctx, 16 * n, &d_a)); CK(cg_alloc(ctx, 16 * n, &d_b)); CK(cg_alloc(ctx, 16 * n, &d_eq));
    CK(cg_h2d(ctx, d_a, a, 16 * n, NULL)); CK(cg_h2d(ctx, d_b, b, 16 * n, NULL));
    CK(cg_build_eq(ctx, w, k, (uint64_t*)d_eq, 0, n, NULL));            /* kernel (iii-eq) */
    cg_mle_desc mles[3] = {{d_eq, n, k, 1}, {d_a, n, k, 1}, {d_b, n, k, 1}};
    const uint64_t coeff[2] = {1, 0};
    const uint32_t off[2] = {0, 3}, idx[3] = {0, 1, 2};
    uint64_t state, nv = k, dg = deg;
    cg_standin_init(&state, (const uint8_t*)"c-abi", 5);
    cg_standin_append_message(&state, (const uint8_t*)&nv, 8);
    cg_standin_append_message(&state, (const uint8_t*)&dg, 8);
    uint64_t* rounds = calloc(2 * deg * k, 8); uint64_t fin[6], *chal = calloc(2 * k, 8);
    CK(cg_sumcheck_prove(ctx, mles, 3, coeff, off, idx, 1, k, deg, CG_SC_DEFAULT, my_challenge, &state, rounds, fin, chal, NULL));
    /* oracle */
    or_mle om[3] = {{eq, k, 1}, {a, k, 1}, {b, k, 1}};
    or_transcript tr; or_tr_init(&tr, (const uint8_t*)"c-abi", 5);
    uint64_t* r2 = calloc(2 * deg * k, 8); uint64_t f2[6], *c2 = calloc(2 * k, 8);
    if (or_sumcheck_prove_standin(om, 3, coeff, off, idx, 1, k, deg, &tr, r2, f2, c2)) return 3;
    int ok = !memcmp(rounds, r2, 16 * deg * k) && !memcmp(fin, f2, 48) && !memcmp(chal, c2, 16 * k) && state == tr.h;
    /* the device eq table itself */
    uint64_t* eq_back = malloc(16 * n);
    CK(cg_d2h(ctx, eq_back, d_eq, 16 * n, NULL)); CK(cg_stream_sync(ctx, NULL));
    ok = ok && !memcmp(eq_back, eq, 16 * n);
    printf("c-abi smoke: %s (launches=%llu)\n", ok ? "bit-exact" : "MISMATCH", (unsigned long long)cg_launch_count(ctx));
    cg_free(ctx, d_a); cg_free(ctx, d_b); cg_free(ctx, d_eq);
    cg_destroy(ctx);
    return ok ? 0 : 1;
}
