/*
 * ceno_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's GKR-sumcheck hot path
 * (scroll-tech/ceno @ ac16425).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (ceno_b200/, include/) never links, imports or calls it.
 *
 * PARITY STATUS: "parity unpinned" for byte-level sumcheck messages.
 *   The arithmetic of IOPProverState::prove / fix_variables / build_eq_x_r_vec
 *   lives in the un-vendored dependency scroll-tech/gkr-backend tag
 *   v1.0.0-alpha.35 (commit 5c9c8a61; crates sumcheck, multilinear_extensions,
 *   transcript) over p3-goldilocks =0.4.3 (reference Cargo.toml:30-40,
 *   Cargo.lock:5767-5770).  That source is not under /root/reference and no
 *   Rust toolchain exists here, so this file restates the PUBLISHED algorithm
 *   and is pinned against every known-answer relation the reference tree
 *   holds for the path (tests/test_oracle_kat.py):
 *     - gkr_iop/src/utils.rs:332-441   closed forms vs MultilinearExtension::evaluate
 *     - gkr_iop/src/selector.rs:396-435 quark selector literal vector
 *     - ceno_zkvm/src/scheme/utils.rs:934-1195 tower-witness literal vectors
 *     - verifier relations p(0)+p(1)=claim, claim'=interp(p)(r), final check
 *       (ceno_recursion_v2/src/main/mod.rs:3488-3531; zerocheck_layer.rs:233-385)
 *   For uniform-size instances the round messages are mathematically
 *   determined by those facts (field arithmetic is exact), so any correct
 *   implementation is bit-identical given the same challenges.  The challenge
 *   source here is a documented STAND-IN (splitmix64), not the reference's
 *   Poseidon2 BasicTranscript (constants are upstream-only).
 *
 * Field: Goldilocks p = 2^64 - 2^32 + 1; extension F_p[X]/(X^2 - 7)
 * (p3-goldilocks 0.4.3 BinomiallyExtendable<2>, W = 7 — upstream fact, see
 * SURVEY.md §A8).  Elements are canonical u64; ext = {c0, c1}.
 *
 * Index convention: MLE index b = sum b_i 2^i, round j binds
 * variable j, i.e. round 0 folds adjacent pairs (2b, 2b+1)
 * (gkr_iop/src/utils.rs:209-232 and tests :356-374).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL
typedef unsigned __int128 u128;
typedef uint64_t gl;
typedef struct { gl c0, c1; } ext;

#define OR_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ field */
static inline gl gl_add(gl a, gl b) { u128 s = (u128)a + b; return (gl)(s >= GL_P ? s - GL_P : s); }
static inline gl gl_sub(gl a, gl b) { return a >= b ? a - b : a + (GL_P - b); }
static inline gl gl_neg(gl a) { return a ? GL_P - a : 0; }
/* slow-but-obvious product, used to pin the fast one in tests */
static inline gl gl_mul_slow(gl a, gl b) { return (gl)(((u128)a * b) % GL_P); }
/* standard Goldilocks reduce128 (the algorithm p3-goldilocks uses) */
static inline gl gl_mul(gl a, gl b) {
    u128 x = (u128)a * b;
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;
    uint64_t t1 = hi_lo * GL_EPS;
    uint64_t r = t0 + t1;
    if (r < t1) r += GL_EPS;
    if (r >= GL_P) r -= GL_P;
    return r;
}
static gl gl_pow(gl a, uint64_t e) { gl r = 1; while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; } return r; }
static inline gl gl_inv(gl a) { return gl_pow(a, GL_P - 2); }

static inline ext E(gl a, gl b) { ext r = {a, b}; return r; }
static inline ext ext_from(gl a) { return E(a, 0); }
static inline ext ext_add(ext a, ext b) { return E(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
static inline ext ext_sub(ext a, ext b) { return E(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
static inline ext ext_neg(ext a) { return E(gl_neg(a.c0), gl_neg(a.c1)); }
static inline ext ext_mul(ext a, ext b) {
    gl v0 = gl_mul(a.c0, b.c0), v1 = gl_mul(a.c1, b.c1);
    gl c0 = gl_add(v0, gl_mul(7, v1));
    gl c1 = gl_add(gl_mul(a.c0, b.c1), gl_mul(a.c1, b.c0));
    return E(c0, c1);
}
static inline ext ext_mul_base(ext a, gl b) { return E(gl_mul(a.c0, b), gl_mul(a.c1, b)); }
static ext ext_inv(ext a) {
    /* 1/(a0 + a1 X) = (a0 - a1 X) / (a0^2 - 7 a1^2) */
    gl n = gl_sub(gl_mul(a.c0, a.c0), gl_mul(7, gl_mul(a.c1, a.c1)));
    gl ni = gl_inv(n);
    return E(gl_mul(a.c0, ni), gl_mul(gl_neg(a.c1), ni));
}
static const ext EXT_ONE = {1, 0};
static const ext EXT_ZERO = {0, 0};

OR_API uint64_t or_gl_add(uint64_t a, uint64_t b) { return gl_add(a, b); }
OR_API uint64_t or_gl_sub(uint64_t a, uint64_t b) { return gl_sub(a, b); }
OR_API uint64_t or_gl_mul(uint64_t a, uint64_t b) { return gl_mul(a, b); }
OR_API uint64_t or_gl_mul_slow(uint64_t a, uint64_t b) { return gl_mul_slow(a, b); }
OR_API uint64_t or_gl_inv(uint64_t a) { return gl_inv(a); }
OR_API void or_ext_mul(const uint64_t* a, const uint64_t* b, uint64_t* o) { ext r = ext_mul(E(a[0], a[1]), E(b[0], b[1])); o[0] = r.c0; o[1] = r.c1; }
OR_API void or_ext_inv(const uint64_t* a, uint64_t* o) { ext r = ext_inv(E(a[0], a[1])); o[0] = r.c0; o[1] = r.c1; }

/* -------------------------------------------------- deterministic inputs */
/* splitmix64 counter-mode generator, SURVEY §8d: element i, limb l <-
 * splitmix64(seed + 2 i + l) mod p (bias <= 2^-32, documented). */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline gl to_canon(uint64_t x) { return x >= GL_P ? x - GL_P : x; }
OR_API void or_fill_ext(uint64_t seed, uint64_t n, uint64_t* out) {
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < n; i++) {
        out[2 * i] = to_canon(splitmix64(seed + 2 * i));
        out[2 * i + 1] = to_canon(splitmix64(seed + 2 * i + 1));
    }
}
OR_API void or_fill_base(uint64_t seed, uint64_t n, uint64_t* out) {
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < n; i++) out[i] = to_canon(splitmix64(seed + 2 * i));
}

/* ----------------------------------------------- stand-in transcript
 * NOT the reference's Poseidon2 BasicTranscript (constants upstream-only,
 * SURVEY §A8/§C-2).  Same call ORDER as the reference so the
 * real transcript can replace it behind the challenge callback. */
typedef struct { uint64_t h; } or_transcript;
static inline void tr_absorb(or_transcript* t, uint64_t x) { t->h = splitmix64(t->h ^ splitmix64(x)); }
static inline uint64_t tr_squeeze(or_transcript* t) { t->h = splitmix64(t->h + 0xD1B54A32D192ED03ULL); return to_canon(t->h); }
OR_API void or_tr_init(or_transcript* t, const uint8_t* label, uint64_t len) {
    t->h = 0x43454E4F42323030ULL; /* "CENOB200" */
    tr_absorb(t, len);
    for (uint64_t i = 0; i < len; i += 8) {
        uint64_t w = 0;
        for (uint64_t j = 0; j < 8 && i + j < len; j++) w |= (uint64_t)label[i + j] << (8 * j);
        tr_absorb(t, w);
    }
}
OR_API void or_tr_append_message(or_transcript* t, const uint8_t* msg, uint64_t len) {
    tr_absorb(t, 0x6D73670000000000ULL ^ len);
    for (uint64_t i = 0; i < len; i += 8) {
        uint64_t w = 0;
        for (uint64_t j = 0; j < 8 && i + j < len; j++) w |= (uint64_t)msg[i + j] << (8 * j);
        tr_absorb(t, w);
    }
}
OR_API void or_tr_append_ext(or_transcript* t, const uint64_t* e, uint64_t n) {
    for (uint64_t i = 0; i < 2 * n; i++) tr_absorb(t, e[i]);
}
OR_API void or_tr_challenge(or_transcript* t, uint64_t* out) { out[0] = tr_squeeze(t); out[1] = tr_squeeze(t); }
/* read_challenge after a label: transcript.sample_and_append_challenge(label) */
OR_API void or_tr_sample(or_transcript* t, const char* label, uint64_t* out) {
    or_tr_append_message(t, (const uint8_t*)label, strlen(label));
    or_tr_challenge(t, out);
}

/* --------------------------------------------------------- eq / MLE ops */
/* build_eq_x_r_vec: eq[b] = prod_i (b_i r_i + (1-b_i)(1-r_i)), b_i = i-th LSB
 * (gkr_iop/src/selector.rs:419-427, gkr_iop/src/utils.rs:91-97). */
OR_API void or_build_eq_x_r_vec(const uint64_t* r, uint32_t k, uint64_t* out) {
    ext* o = (ext*)out;
    o[0] = EXT_ONE;
    for (uint32_t i = 0; i < k; i++) {
        ext ri = E(r[2 * i], r[2 * i + 1]);
        uint64_t n = 1ULL << i;
        for (uint64_t b = 0; b < n; b++) {
            ext hi = ext_mul(o[b], ri);
            o[b + n] = hi;
            o[b] = ext_sub(o[b], hi);
        }
    }
}
/* eq_eval(x, y) = prod_i (x_i y_i + (1-x_i)(1-y_i))  (gkr_iop/src/utils.rs:168-176) */
static ext eq_eval(const ext* a, const ext* b, uint32_t n) {
    ext acc = EXT_ONE;
    for (uint32_t i = 0; i < n; i++) {
        ext xy = ext_mul(a[i], b[i]);
        ext t = ext_add(ext_add(xy, xy), ext_sub(EXT_ONE, ext_add(a[i], b[i]))); /* 2xy + 1 - x - y */
        acc = ext_mul(acc, t);
    }
    return acc;
}
OR_API void or_eq_eval(const uint64_t* a, const uint64_t* b, uint32_t n, uint64_t* out) {
    ext r = eq_eval((const ext*)a, (const ext*)b, n); out[0] = r.c0; out[1] = r.c1;
}
/* eq_eval_less_or_equal_than (gkr_iop/src/utils.rs:166-208) — literal restatement */
static ext eq_eval_le(uint64_t max_idx, const ext* a, uint32_t alen, const ext* b, uint32_t blen) {
    ext* rp = (ext*)malloc(sizeof(ext) * (blen + 1));
    ext* rp2 = (ext*)malloc(sizeof(ext) * (blen + 1));
    rp[0] = EXT_ONE;
    for (uint32_t i = 0; i < blen; i++) {
        ext t = ext_add(ext_mul(a[i], b[i]), ext_mul(ext_sub(EXT_ONE, a[i]), ext_sub(EXT_ONE, b[i])));
        rp[i + 1] = ext_mul(rp[i], t);
    }
    rp2[blen] = EXT_ONE;
    for (int32_t i = (int32_t)blen - 1; i >= 0; i--) {
        ext bit = ext_from((max_idx >> i) & 1);
        ext t = ext_add(ext_mul(ext_mul(a[i], b[i]), bit),
                        ext_mul(ext_mul(ext_sub(EXT_ONE, a[i]), ext_sub(EXT_ONE, b[i])), ext_sub(EXT_ONE, bit)));
        rp2[i] = ext_mul(rp2[i + 1], t);
    }
    ext ans = rp[blen];
    for (uint32_t i = 0; i < blen; i++) {
        if ((max_idx >> i) & 1) continue;
        ans = ext_sub(ans, ext_mul(ext_mul(ext_mul(rp[i], rp2[i + 1]), a[i]), b[i]));
    }
    for (uint32_t i = blen; i < alen; i++) ans = ext_mul(ans, ext_sub(EXT_ONE, a[i]));
    free(rp); free(rp2);
    return ans;
}
OR_API void or_eq_eval_less_or_equal_than(uint64_t max_idx, const uint64_t* a, uint32_t alen,
                                          const uint64_t* b, uint32_t blen, uint64_t* out) {
    ext r = eq_eval_le(max_idx, (const ext*)a, alen, (const ext*)b, blen); out[0] = r.c0; out[1] = r.c1;
}

/* MultilinearExtension::evaluate: point[i] pairs with index bit i. */
OR_API void or_mle_evaluate(const uint64_t* evals, uint32_t is_ext, uint32_t num_vars,
                            const uint64_t* point, uint64_t* out) {
    uint64_t n = 1ULL << num_vars;
    ext* w = (ext*)malloc(sizeof(ext) * n);
    for (uint64_t i = 0; i < n; i++) w[i] = is_ext ? E(evals[2 * i], evals[2 * i + 1]) : ext_from(evals[i]);
    for (uint32_t j = 0; j < num_vars; j++) {
        ext r = E(point[2 * j], point[2 * j + 1]);
        n >>= 1;
        for (uint64_t b = 0; b < n; b++) w[b] = ext_add(w[2 * b], ext_mul(r, ext_sub(w[2 * b + 1], w[2 * b])));
    }
    out[0] = w[0].c0; out[1] = w[0].c1;
    free(w);
}

/* fix_variables (one variable, LSB): f'[b] = f[2b] + r (f[2b+1] - f[2b]).
 * Output is always ext (a base MLE becomes ext after the first fold). */
OR_API void or_fix_variable(const uint64_t* evals, uint32_t is_ext, uint64_t len, const uint64_t* r, uint64_t* out) {
    ext rr = E(r[0], r[1]);
    ext* o = (ext*)out;
    uint64_t half = len / 2;
#pragma omp parallel for schedule(static)
    for (uint64_t b = 0; b < half; b++) {
        if (is_ext) {
            ext lo = E(evals[4 * b], evals[4 * b + 1]), hi = E(evals[4 * b + 2], evals[4 * b + 3]);
            o[b] = ext_add(lo, ext_mul(rr, ext_sub(hi, lo)));
        } else {
            gl lo = evals[2 * b], hi = evals[2 * b + 1];
            o[b] = ext_add(ext_from(lo), ext_mul_base(rr, gl_sub(hi, lo)));
        }
    }
}

/* closed forms from gkr_iop/src/utils.rs:210-308 (KAT targets) */
static ext eval_wellform_address_vec(uint64_t offset, uint64_t scaled, const ext* r, uint32_t n, int descending) {
    ext sum = EXT_ZERO, state = EXT_ONE;
    for (uint32_t i = 0; i < n; i++) { sum = ext_add(sum, ext_mul(r[i], state)); state = ext_mul(state, ext_from(2)); }
    ext tmp = ext_mul(ext_from(to_canon(scaled)), sum);
    if (descending) tmp = ext_neg(tmp);
    return ext_add(ext_from(to_canon(offset)), tmp);
}
OR_API void or_eval_wellform_address_vec(uint64_t offset, uint64_t scaled, const uint64_t* r, uint32_t n, int desc, uint64_t* out) {
    ext v = eval_wellform_address_vec(offset, scaled, (const ext*)r, n, desc); out[0] = v.c0; out[1] = v.c1;
}
OR_API void or_eval_stacked_wellform_address_vec(const uint64_t* r_, uint32_t n, uint64_t* out) {
    const ext* r = (const ext*)r_;
    ext res = EXT_ZERO;
    if (n >= 2) for (uint32_t i = 1; i < n; i++)
        res = ext_add(ext_mul(res, ext_sub(EXT_ONE, r[i])), ext_mul(eval_wellform_address_vec(0, 1, r, i, 0), r[i]));
    out[0] = res.c0; out[1] = res.c1;
}
OR_API void or_eval_stacked_constant_vec(const uint64_t* r_, uint32_t n, uint64_t* out) {
    const ext* r = (const ext*)r_;
    ext res = EXT_ZERO;
    if (n >= 2) for (uint32_t i = 1; i < n; i++)
        res = ext_add(ext_mul(res, ext_sub(EXT_ONE, r[i])), ext_mul(ext_from(i), r[i]));
    out[0] = res.c0; out[1] = res.c1;
}

/* ------------------------------------------------------------ selectors
 * SelectorType::compute (gkr_iop/src/selector.rs:131-245).
 * kind: 0 Whole, 1 Prefix(offset,num_instances), 2 OrderedSparse(indices,num_vars_inner),
 *       3 QuarkBinaryTreeLessThan(num_instances). */
OR_API int or_selector_compute(int kind, const uint64_t* point, uint32_t num_vars, uint64_t offset,
                               uint64_t num_instances, const uint64_t* indices, uint32_t n_indices,
                               uint32_t inner_vars, uint64_t* out) {
    uint64_t n = 1ULL << num_vars;
    ext* sel = (ext*)out;
    or_build_eq_x_r_vec(point, num_vars, out);
    if (kind == 0) return 0;
    if (kind == 1) {
        uint64_t start = offset, end = offset + num_instances;
        if (end > n) return -1;
        for (uint64_t i = 0; i < start; i++) sel[i] = EXT_ZERO;
        for (uint64_t i = end; i < n; i++) sel[i] = EXT_ZERO;
        return 0;
    }
    if (kind == 2) {
        uint64_t chunk = 1ULL << inner_vars;
        for (uint64_t c = 0; c < n / chunk; c++) {
            ext* ch = sel + c * chunk;
            if (c >= num_instances) { for (uint64_t i = 0; i < chunk; i++) ch[i] = EXT_ZERO; continue; }
            uint32_t it = 0;
            for (uint64_t i = 0; i < chunk; i++) {
                if (it < n_indices && indices[it] == i) it++; else ch[i] = EXT_ZERO;
            }
        }
        return 0;
    }
    if (kind == 3) {
        if (offset != 0) return -1;
        uint64_t ninst = num_instances, start = 0, chunk_len = n / 2;
        uint32_t i = 0;
        while (chunk_len > 0) {
            uint64_t cur = 0;
            if (i < num_vars) { cur = ninst / 2; ninst = (ninst + 1) / 2; }
            uint64_t zs = cur < chunk_len ? cur : chunk_len;
            for (uint64_t x = zs; x < chunk_len; x++) sel[start + x] = EXT_ZERO;
            start += chunk_len; chunk_len /= 2; i++;
        }
        sel[n - 1] = EXT_ZERO;
        return 0;
    }
    return -2;
}

/* -------------------------------------------------------------- sumcheck
 * Restatement of IOPProverState::prove for uniform-size instances
 * (SURVEY §8a1, §A1-A3, A7): P(x) = sum_t c_t prod_{i in S_t} f_i(x).
 * Round j message = [p_j(1) .. p_j(d)]; p_j(0) is not sent
 * (ceno_recursion_v2/src/main/mod.rs:3513-3526).  Parallel decomposition =
 * contiguous hypercube chunks per thread, per-thread partial sums merged on
 * the main thread (mirrors num_threads = optimal_sumcheck_threads(k),
 * ceno_zkvm/src/scheme/cpu/mod.rs:89,411). */
typedef struct { const uint64_t* data; uint32_t num_vars; uint32_t is_ext; } or_mle;
typedef void (*or_challenge_fn)(void* user, uint32_t round, const uint64_t* evals, uint32_t degree, uint64_t* out_r);

#define OR_MAX_DEG 16

OR_API int or_sumcheck_prove(const or_mle* mles, uint32_t n_mles, const uint64_t* term_coeff,
                             const uint32_t* term_off, const uint32_t* term_idx, uint32_t n_terms,
                             uint32_t num_vars, uint32_t degree, or_challenge_fn cb, void* user,
                             uint64_t* round_evals, uint64_t* final_evals, uint64_t* challenges) {
    if (degree > OR_MAX_DEG || degree == 0) return -1;
    for (uint32_t i = 0; i < n_mles; i++) if (mles[i].num_vars > num_vars) return -2;
    /* Mixed sizes ("frontload", the cross-chip batched main sumcheck, ceno_zkvm/src/scheme/cpu/mod.rs:1332-1360):
     * an MLE f with k' < k variables stands for  F(x) = f(x_0..x_{k'-1}) * prod_{j >= k'} x_j , i.e. the dense
     * table that holds f in its LAST 2^k' entries and zero elsewhere.  Pinned in-tree by the verifier's final
     * claim, restated in ceno_recursion_v2/src/main/mod.rs:3414-3448 (every factor's evaluation is multiplied by
     * every tail challenge) and by ceno_zkvm/src/scheme/verifier.rs:233-237.  The final evaluation reported for
     * such an MLE is the raw f(r_0..r_{k'-1}) (the callers multiply the tail themselves, cpu/mod.rs:1346-1358).
     * This restatement simply runs the uniform prover on the dense embedding. */
    uint64_t n = 1ULL << num_vars;
    ext** f = (ext**)malloc(sizeof(ext*) * n_mles);
    ext** g = (ext**)malloc(sizeof(ext*) * n_mles);
    ext** raw = (ext**)calloc(n_mles, sizeof(ext*));
    for (uint32_t i = 0; i < n_mles; i++) {
        f[i] = (ext*)malloc(sizeof(ext) * n);
        g[i] = (ext*)malloc(sizeof(ext) * (n / 2 ? n / 2 : 1));
        const uint64_t* d = mles[i].data;
        const uint64_t ni = 1ULL << mles[i].num_vars, off = n - ni;
        if (off) memset(f[i], 0, sizeof(ext) * off);
        if (mles[i].is_ext) memcpy(f[i] + off, d, sizeof(ext) * ni);
        else {
#pragma omp parallel for schedule(static)
            for (uint64_t b = 0; b < ni; b++) f[i][off + b] = ext_from(d[b]);
        }
        if (off) { raw[i] = (ext*)malloc(sizeof(ext) * ni); memcpy(raw[i], f[i] + off, sizeof(ext) * ni); }
    }
    const ext* coeff = (const ext*)term_coeff;
    int nthr = 1;
#ifdef _OPENMP
    nthr = omp_get_max_threads();
#endif
    ext* part = (ext*)malloc(sizeof(ext) * OR_MAX_DEG * nthr);
    for (uint32_t j = 0; j < num_vars; j++) {
        uint64_t half = n >> 1;
        for (int t = 0; t < nthr * OR_MAX_DEG; t++) part[t] = EXT_ZERO;
#pragma omp parallel
        {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            ext acc[OR_MAX_DEG];
            for (uint32_t t = 0; t < degree; t++) acc[t] = EXT_ZERO;
#pragma omp for schedule(static)
            for (uint64_t b = 0; b < half; b++) {
                for (uint32_t t = 0; t < n_terms; t++) {
                    ext prod[OR_MAX_DEG];
                    for (uint32_t x = 0; x < degree; x++) prod[x] = coeff[t];
                    for (uint32_t q = term_off[t]; q < term_off[t + 1]; q++) {
                        const ext* fi = f[term_idx[q]];
                        ext lo = fi[2 * b], hi = fi[2 * b + 1];
                        ext dl = ext_sub(hi, lo), v = hi;
                        for (uint32_t x = 0; x < degree; x++) { prod[x] = ext_mul(prod[x], v); v = ext_add(v, dl); }
                    }
                    for (uint32_t x = 0; x < degree; x++) acc[x] = ext_add(acc[x], prod[x]);
                }
            }
            for (uint32_t t = 0; t < degree; t++) part[tid * OR_MAX_DEG + t] = acc[t];
        }
        ext* msg = (ext*)(round_evals + (uint64_t)j * degree * 2);
        for (uint32_t t = 0; t < degree; t++) {
            ext s = EXT_ZERO;
            for (int q = 0; q < nthr; q++) s = ext_add(s, part[q * OR_MAX_DEG + t]);
            msg[t] = s;
        }
        uint64_t r_[2];
        cb(user, j, (const uint64_t*)msg, degree, r_);
        challenges[2 * j] = r_[0]; challenges[2 * j + 1] = r_[1];
        ext r = E(r_[0], r_[1]);
        for (uint32_t i = 0; i < n_mles; i++) {
            ext* src = f[i]; ext* dst = g[i];
#pragma omp parallel for schedule(static)
            for (uint64_t b = 0; b < half; b++) dst[b] = ext_add(src[2 * b], ext_mul(r, ext_sub(src[2 * b + 1], src[2 * b])));
            f[i] = dst; g[i] = src;
            if (raw[i] && j < mles[i].num_vars) {   /* the raw small MLE follows its own first k' challenges */
                const uint64_t h2 = 1ULL << (mles[i].num_vars - j - 1);
                for (uint64_t b = 0; b < h2; b++) raw[i][b] = ext_add(raw[i][2 * b], ext_mul(r, ext_sub(raw[i][2 * b + 1], raw[i][2 * b])));
            }
        }
        n = half;
    }
    for (uint32_t i = 0; i < n_mles; i++) {
        const ext v = raw[i] ? raw[i][0] : f[i][0];
        final_evals[2 * i] = v.c0; final_evals[2 * i + 1] = v.c1;
    }
    for (uint32_t i = 0; i < n_mles; i++) { free(f[i]); free(g[i]); free(raw[i]); }
    free(f); free(g); free(raw); free(part);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * CPU-baseline variant: the reference's parallel decomposition.
 * `num_threads = optimal_sumcheck_threads(k)` contiguous hypercube chunks, one per thread; each thread
 * evaluates and folds its own chunk IN PLACE (sequential within the chunk, so the LSB fold is safe) and
 * owns a shrinking prefix of it; partial round sums are merged on the main thread; once every chunk is
 * one element the T survivors are compacted and a single thread finishes (the tail).  Ext MLEs given
 * with consume != 0 are folded inside the caller's buffers, like Either::Right(&mut mle) in the
 * reference (ceno_zkvm/src/scheme/cpu/mod.rs:417-418, gkr_iop/src/gkr/layer/cpu/mod.rs:186-200) — no
 * allocation or copy on the timed path.  The inner loop is specialised for one product of three ext
 * MLEs (the tower layer / T3 shape), everything else takes the generic term loop.
 * Bit-identical to or_sumcheck_prove (tests/test_oracle_kat.py). */
static inline gl gl_add_fast(gl a, gl b) { gl s = a + b; return (s < a || s >= GL_P) ? s - GL_P : s; }
static inline ext ext_add_f(ext a, ext b) { return E(gl_add_fast(a.c0, b.c0), gl_add_fast(a.c1, b.c1)); }
static inline ext ext_sub_f(ext a, ext b) { return E(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
static inline gl gl_mul7(gl a) { u128 x = (u128)a * 7; uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64); uint64_t t = hi * GL_EPS; uint64_t r = lo + t; if (r < t) r += GL_EPS; return r >= GL_P ? r - GL_P : r; }
static inline ext ext_mul_f(ext a, ext b) {
    /* a0 b0 + 7 a1 b1 and a0 b1 + a1 b0 with one reduction each (129-bit sums) */
    u128 p0 = (u128)a.c0 * b.c0, p1 = (u128)a.c1 * gl_mul7(b.c1);
    u128 s = p0 + p1; uint64_t top = s < p0;
    u128 q0 = (u128)a.c0 * b.c1, q1 = (u128)a.c1 * b.c0;
    u128 t = q0 + q1; uint64_t top2 = t < q0;
    ext r;
    {   uint64_t lo = (uint64_t)s, hi = (uint64_t)(s >> 64), hh = hi >> 32, hl = hi & GL_EPS;
        uint64_t sub = hh + (top << 32); uint64_t t0 = lo - sub; if (lo < sub) t0 -= GL_EPS;
        uint64_t t1 = hl * GL_EPS; uint64_t x = t0 + t1; if (x < t1) x += GL_EPS; r.c0 = x >= GL_P ? x - GL_P : x; }
    {   uint64_t lo = (uint64_t)t, hi = (uint64_t)(t >> 64), hh = hi >> 32, hl = hi & GL_EPS;
        uint64_t sub = hh + (top2 << 32); uint64_t t0 = lo - sub; if (lo < sub) t0 -= GL_EPS;
        uint64_t t1 = hl * GL_EPS; uint64_t x = t0 + t1; if (x < t1) x += GL_EPS; r.c1 = x >= GL_P ? x - GL_P : x; }
    return r;
}
OR_API int or_sumcheck_prove_chunked(const or_mle* mles, uint32_t n_mles, const uint64_t* term_coeff,
                                     const uint32_t* term_off, const uint32_t* term_idx, uint32_t n_terms,
                                     uint32_t num_vars, uint32_t degree, or_challenge_fn cb, void* user,
                                     uint64_t* round_evals, uint64_t* final_evals, uint64_t* challenges, int consume) {
    if (degree > OR_MAX_DEG || degree == 0) return -1;
    for (uint32_t i = 0; i < n_mles; i++) if (mles[i].num_vars != num_vars) return -2;
    const uint64_t n = 1ULL << num_vars;
    int maxthr = 1;
#ifdef _OPENMP
    maxthr = omp_get_max_threads();
#endif
    uint32_t lt = 0;
    while ((2u << lt) <= (uint32_t)maxthr && lt + 1 < num_vars) lt++;   /* T = 2^lt chunks, each >= 2 elements */
    if (num_vars == 0) lt = 0;
    const uint32_t T = 1u << lt;
    const uint64_t chunk = n >> lt;
    ext** f = (ext**)malloc(sizeof(ext*) * (n_mles ? n_mles : 1));
    int* owned = (int*)calloc(n_mles ? n_mles : 1, sizeof(int));
    for (uint32_t i = 0; i < n_mles; i++) {
        if (mles[i].is_ext && consume) { f[i] = (ext*)mles[i].data; continue; }
        f[i] = (ext*)malloc(sizeof(ext) * n); owned[i] = 1;
        const uint64_t* d = mles[i].data;
#pragma omp parallel for schedule(static)
        for (uint64_t b = 0; b < n; b++) f[i][b] = mles[i].is_ext ? E(d[2 * b], d[2 * b + 1]) : ext_from(d[b]);
    }
    const ext* coeff = (const ext*)term_coeff;
    const int t3 = (n_terms == 1 && degree == 3 && term_off[1] - term_off[0] == 3);
    ext* part = (ext*)malloc(sizeof(ext) * OR_MAX_DEG * T);
    uint64_t live = chunk;          /* live elements per chunk */
    uint32_t j = 0;
    for (; j < num_vars && live >= 2; j++) {
        const uint64_t half = live >> 1;
#pragma omp parallel for schedule(static) num_threads(T)
        for (uint32_t t = 0; t < T; t++) {
            ext acc[OR_MAX_DEG];
            for (uint32_t x = 0; x < degree; x++) acc[x] = EXT_ZERO;
            const uint64_t base = (uint64_t)t * chunk;
            if (t3) {
                const ext* A = f[term_idx[term_off[0]]] + base; const ext* B = f[term_idx[term_off[0] + 1]] + base;
                const ext* Cc = f[term_idx[term_off[0] + 2]] + base;
                const ext c = coeff[0];
                const int c_one = (c.c0 == 1 && c.c1 == 0);
                ext h1 = EXT_ZERO, h2 = EXT_ZERO, h3 = EXT_ZERO;
                for (uint64_t b = 0; b < half; b++) {
                    ext a1 = A[2 * b + 1], ad = ext_sub_f(a1, A[2 * b]);
                    ext b1 = B[2 * b + 1], bd = ext_sub_f(b1, B[2 * b]);
                    ext c1 = Cc[2 * b + 1], cd = ext_sub_f(c1, Cc[2 * b]);
                    ext a2 = ext_add_f(a1, ad), b2 = ext_add_f(b1, bd), c2 = ext_add_f(c1, cd);
                    ext a3 = ext_add_f(a2, ad), b3 = ext_add_f(b2, bd), c3 = ext_add_f(c2, cd);
                    h1 = ext_add_f(h1, ext_mul_f(ext_mul_f(a1, b1), c1));
                    h2 = ext_add_f(h2, ext_mul_f(ext_mul_f(a2, b2), c2));
                    h3 = ext_add_f(h3, ext_mul_f(ext_mul_f(a3, b3), c3));
                }
                if (!c_one) { h1 = ext_mul_f(h1, c); h2 = ext_mul_f(h2, c); h3 = ext_mul_f(h3, c); }
                acc[0] = h1; acc[1] = h2; acc[2] = h3;
            } else {
                for (uint64_t b = 0; b < half; b++)
                    for (uint32_t q = 0; q < n_terms; q++) {
                        ext prod[OR_MAX_DEG];
                        for (uint32_t x = 0; x < degree; x++) prod[x] = coeff[q];
                        for (uint32_t z = term_off[q]; z < term_off[q + 1]; z++) {
                            const ext* fi = f[term_idx[z]] + base;
                            ext hi = fi[2 * b + 1], dl = ext_sub_f(hi, fi[2 * b]), v = hi;
                            for (uint32_t x = 0; x < degree; x++) { prod[x] = ext_mul_f(prod[x], v); v = ext_add_f(v, dl); }
                        }
                        for (uint32_t x = 0; x < degree; x++) acc[x] = ext_add_f(acc[x], prod[x]);
                    }
            }
            for (uint32_t x = 0; x < degree; x++) part[t * OR_MAX_DEG + x] = acc[x];
        }
        ext* msg = (ext*)(round_evals + (uint64_t)j * degree * 2);
        for (uint32_t x = 0; x < degree; x++) {
            ext sacc = EXT_ZERO;
            for (uint32_t t = 0; t < T; t++) sacc = ext_add_f(sacc, part[t * OR_MAX_DEG + x]);
            msg[x] = sacc;
        }
        uint64_t r_[2];
        cb(user, j, (const uint64_t*)msg, degree, r_);
        challenges[2 * j] = r_[0]; challenges[2 * j + 1] = r_[1];
        const ext r = E(r_[0], r_[1]);
#pragma omp parallel for schedule(static) num_threads(T) collapse(2)
        for (uint32_t i = 0; i < n_mles; i++)
            for (uint32_t t = 0; t < T; t++) {
                ext* c = f[i] + (uint64_t)t * chunk;
                for (uint64_t b = 0; b < half; b++) c[b] = ext_add_f(c[2 * b], ext_mul_f(ext_sub_f(c[2 * b + 1], c[2 * b]), r));
            }
        live = half;
    }
    /* tail: one survivor per chunk -> compact, finish single-threaded */
    uint64_t m = T;
    ext** g = (ext**)malloc(sizeof(ext*) * (n_mles ? n_mles : 1));
    for (uint32_t i = 0; i < n_mles; i++) {
        g[i] = (ext*)malloc(sizeof(ext) * m);
        for (uint64_t t = 0; t < m; t++) g[i][t] = f[i][t * chunk];
    }
    for (; j < num_vars; j++) {
        const uint64_t half = m >> 1;
        ext acc[OR_MAX_DEG];
        for (uint32_t x = 0; x < degree; x++) acc[x] = EXT_ZERO;
        for (uint64_t b = 0; b < half; b++)
            for (uint32_t q = 0; q < n_terms; q++) {
                ext prod[OR_MAX_DEG];
                for (uint32_t x = 0; x < degree; x++) prod[x] = coeff[q];
                for (uint32_t z = term_off[q]; z < term_off[q + 1]; z++) {
                    const ext* fi = g[term_idx[z]];
                    ext hi = fi[2 * b + 1], dl = ext_sub_f(hi, fi[2 * b]), v = hi;
                    for (uint32_t x = 0; x < degree; x++) { prod[x] = ext_mul_f(prod[x], v); v = ext_add_f(v, dl); }
                }
                for (uint32_t x = 0; x < degree; x++) acc[x] = ext_add_f(acc[x], prod[x]);
            }
        ext* msg = (ext*)(round_evals + (uint64_t)j * degree * 2);
        for (uint32_t x = 0; x < degree; x++) msg[x] = acc[x];
        uint64_t r_[2];
        cb(user, j, (const uint64_t*)msg, degree, r_);
        challenges[2 * j] = r_[0]; challenges[2 * j + 1] = r_[1];
        const ext r = E(r_[0], r_[1]);
        for (uint32_t i = 0; i < n_mles; i++)
            for (uint64_t b = 0; b < half; b++) g[i][b] = ext_add_f(g[i][2 * b], ext_mul_f(ext_sub_f(g[i][2 * b + 1], g[i][2 * b]), r));
        m = half;
    }
    for (uint32_t i = 0; i < n_mles; i++) { final_evals[2 * i] = g[i][0].c0; final_evals[2 * i + 1] = g[i][0].c1; }
    for (uint32_t i = 0; i < n_mles; i++) { free(g[i]); if (owned[i]) free(f[i]); }
    free(f); free(g); free(owned); free(part);
    return 0;
}
OR_API int or_sumcheck_prove_chunked_standin(const or_mle* mles, uint32_t n_mles, const uint64_t* term_coeff,
                                             const uint32_t* term_off, const uint32_t* term_idx, uint32_t n_terms,
                                             uint32_t num_vars, uint32_t degree, or_transcript* tr,
                                             uint64_t* round_evals, uint64_t* final_evals, uint64_t* challenges, int consume);

/* Stand-in transcript driven sumcheck (SURVEY §A2 order):
 *   append_message(num_vars LE u64), append_message(degree LE u64);
 *   per round: absorb d ext evals, label "Internal round", sample challenge. */
static void standin_cb(void* user, uint32_t round, const uint64_t* evals, uint32_t degree, uint64_t* out_r) {
    (void)round;
    or_transcript* t = (or_transcript*)user;
    or_tr_append_ext(t, evals, degree);
    or_tr_sample(t, "Internal round", out_r);
}
OR_API int or_sumcheck_prove_standin(const or_mle* mles, uint32_t n_mles, const uint64_t* term_coeff,
                                     const uint32_t* term_off, const uint32_t* term_idx, uint32_t n_terms,
                                     uint32_t num_vars, uint32_t degree, or_transcript* tr,
                                     uint64_t* round_evals, uint64_t* final_evals, uint64_t* challenges) {
    uint64_t nv = num_vars, dg = degree;
    or_tr_append_message(tr, (const uint8_t*)&nv, 8);
    or_tr_append_message(tr, (const uint8_t*)&dg, 8);
    return or_sumcheck_prove(mles, n_mles, term_coeff, term_off, term_idx, n_terms, num_vars, degree,
                             standin_cb, tr, round_evals, final_evals, challenges);
}

OR_API int or_sumcheck_prove_chunked_standin(const or_mle* mles, uint32_t n_mles, const uint64_t* term_coeff,
                                             const uint32_t* term_off, const uint32_t* term_idx, uint32_t n_terms,
                                             uint32_t num_vars, uint32_t degree, or_transcript* tr,
                                             uint64_t* round_evals, uint64_t* final_evals, uint64_t* challenges, int consume) {
    uint64_t nv = num_vars, dg = degree;
    or_tr_append_message(tr, (const uint8_t*)&nv, 8);
    or_tr_append_message(tr, (const uint8_t*)&dg, 8);
    return or_sumcheck_prove_chunked(mles, n_mles, term_coeff, term_off, term_idx, n_terms, num_vars, degree,
                                     standin_cb, tr, round_evals, final_evals, challenges, consume);
}

/* verifier-side helpers (ceno_recursion_v2/src/main/mod.rs:3513-3526):
 * interpolate p through (0,e0),(1,ev[0]),...,(d,ev[d-1]) and evaluate at r. */
OR_API void or_extrapolate_uni_poly(const uint64_t* eval0, const uint64_t* evals, uint32_t degree,
                                    const uint64_t* r_, uint64_t* out) {
    ext ys[OR_MAX_DEG + 1];
    ys[0] = E(eval0[0], eval0[1]);
    for (uint32_t i = 0; i < degree; i++) ys[i + 1] = E(evals[2 * i], evals[2 * i + 1]);
    ext r = E(r_[0], r_[1]);
    ext acc = EXT_ZERO;
    for (uint32_t i = 0; i <= degree; i++) {
        ext num = EXT_ONE; gl den = 1;
        for (uint32_t j = 0; j <= degree; j++) {
            if (j == i) continue;
            num = ext_mul(num, ext_sub(r, ext_from(j)));
            den = gl_mul(den, gl_sub(i, j));
        }
        acc = ext_add(acc, ext_mul(ys[i], ext_mul_base(num, gl_inv(den))));
    }
    out[0] = acc.c0; out[1] = acc.c1;
}

/* ----------------------------------------------------- tower witness build
 * interleaving_mles_to_mles (ceno_zkvm/src/scheme/utils.rs:402-462).
 * mles: n_mles arrays, each `mle_len` elements (ext or base), num_limbs outputs
 * of length out_len each (caller allocates num_limbs*out_len ext). */
OR_API uint64_t or_interleave_out_len(uint32_t n_mles, uint64_t num_instances, uint32_t num_limbs) {
    uint64_t np2 = 1; while (np2 < num_instances) np2 <<= 1;
    if (np2 < 2) np2 = 2; /* next_pow2_instance_padding: minimum 2 */
    uint32_t l2i = 0; while ((1ULL << l2i) < np2) l2i++;
    uint32_t l2m = 0; while ((1ULL << l2m) < n_mles) l2m++;
    uint32_t l2l = 0; while ((1ULL << l2l) < num_limbs) l2l++;
    uint32_t e = l2m + (l2i > l2l ? l2i - l2l : 0);
    return 1ULL << e;
}
OR_API void or_interleaving_mles_to_mles(const uint64_t* const* mles, const uint32_t* is_ext, uint32_t n_mles,
                                         uint64_t mle_len, uint64_t num_instances, uint32_t num_limbs,
                                         const uint64_t* default_, uint64_t* out) {
    uint64_t out_len = or_interleave_out_len(n_mles, num_instances, num_limbs);
    uint64_t per_fanin_len = mle_len / num_limbs; if (per_fanin_len < 1) per_fanin_len = 1;
    uint32_t l2m = 0; while ((1ULL << l2m) < n_mles) l2m++;
    uint64_t per_instance = 1ULL << l2m;
    ext def = E(default_[0], default_[1]);
    for (uint32_t fi = 0; fi < num_limbs; fi++) {
        ext* ev = (ext*)out + (uint64_t)fi * out_len;
        for (uint64_t x = 0; x < out_len; x++) ev[x] = def;
        uint64_t start = per_fanin_len * fi;
        if (start < num_instances) {
            uint64_t valid = per_fanin_len < num_instances - start ? per_fanin_len : num_instances - start;
            for (uint32_t i = 0; i < n_mles; i++) {
                /* Ext arm takes valid_instances_len, Base arm per_fanin_len (utils.rs:433-456);
                 * both are clipped by the slice `.get(..)` and by the chunk count. */
                uint64_t cnt = is_ext[i] ? valid : per_fanin_len;
                if (start + cnt > mle_len) cnt = 0; /* `.get(range)` out of range -> None -> &[] */
                uint64_t maxc = out_len / per_instance;
                if (cnt > maxc) cnt = maxc;
                for (uint64_t s = 0; s < cnt; s++) {
                    const uint64_t* d = mles[i];
                    ev[s * per_instance + i] = is_ext[i] ? E(d[2 * (start + s)], d[2 * (start + s) + 1]) : ext_from(d[start + s]);
                }
            }
        }
    }
}

/* infer_tower_product_witness (ceno_zkvm/src/scheme/utils.rs:588-659), fanin 2.
 * in: last layer f1,f2 each 2^(num_vars-1) ext.  out: layers[0..num_vars) where
 * layer l (l=0 output) holds 2 arrays of 2^l ext, packed consecutively:
 * offset(l) = 2*(2^l - 1), array 0 then array 1. */
OR_API void or_infer_tower_product_witness(uint32_t num_vars, const uint64_t* f1, const uint64_t* f2, uint64_t* out) {
    ext* o = (ext*)out;
    uint64_t len = 1ULL << (num_vars - 1);
    uint64_t off = 2 * (len - 1);
    memcpy(o + off, f1, sizeof(ext) * len);
    memcpy(o + off + len, f2, sizeof(ext) * len);
    for (int32_t l = (int32_t)num_vars - 2; l >= 0; l--) {
        uint64_t ilen = 1ULL << (l + 1), olen = 1ULL << l;
        ext* in1 = o + 2 * (ilen - 1); ext* in2 = in1 + ilen;
        ext* out0 = o + 2 * (olen - 1);
        for (uint32_t idx = 0; idx < 2; idx++) {
            uint64_t start = idx * olen;
#pragma omp parallel for schedule(static)
            for (uint64_t x = 0; x < olen; x++) out0[idx * olen + x] = ext_mul(in1[start + x], in2[start + x]);
        }
    }
}
/* infer_tower_logup_witness (ceno_zkvm/src/scheme/utils.rs:488-581).
 * q1,q2 (and optional p1,p2; NULL -> numerators all one) each 2^nv ext.
 * out: layers l=0..nv, layer l holds [p1,p2,q1,q2] each 2^l: offset(l)=4*(2^l-1). */
OR_API void or_infer_tower_logup_witness(uint32_t nv, const uint64_t* p1, const uint64_t* p2,
                                         const uint64_t* q1, const uint64_t* q2, uint64_t* out) {
    ext* o = (ext*)out;
    uint64_t len = 1ULL << nv;
    ext* L = o + 4 * (len - 1);
    for (uint64_t x = 0; x < len; x++) {
        L[x] = p1 ? ((const ext*)p1)[x] : EXT_ONE;
        L[len + x] = p2 ? ((const ext*)p2)[x] : EXT_ONE;
        L[2 * len + x] = ((const ext*)q1)[x];
        L[3 * len + x] = ((const ext*)q2)[x];
    }
    int have_p = p1 != NULL;
    for (int32_t l = (int32_t)nv - 1; l >= 0; l--) {
        uint64_t ilen = 1ULL << (l + 1), olen = 1ULL << l;
        ext* I = o + 4 * (ilen - 1); ext* O = o + 4 * (olen - 1);
        ext *ip1 = I, *ip2 = I + ilen, *iq1 = I + 2 * ilen, *iq2 = I + 3 * ilen;
        for (uint32_t idx = 0; idx < 2; idx++) {
            uint64_t start = idx * olen;
#pragma omp parallel for schedule(static)
            for (uint64_t x = 0; x < olen; x++) {
                ext a1 = iq1[start + x], a2 = iq2[start + x], p, q;
                if (have_p || l != (int32_t)nv - 1) p = ext_add(ext_mul(a1, ip2[start + x]), ext_mul(a2, ip1[start + x]));
                else p = ext_add(a1, a2);
                q = ext_mul(a1, a2);
                O[idx * olen + x] = p;           /* next p_{idx} */
                O[(2 + idx) * olen + x] = q;     /* next q_{idx} */
            }
        }
    }
}

/* ----------------------------------------------------------- tower prover
 * CpuTowerProver::create_proof (ceno_zkvm/src/scheme/cpu/mod.rs:346-554),
 * num_fanin = 2, stand-in transcript.  Witness layout = the packed layouts
 * produced by or_infer_tower_{product,logup}_witness.  prod spec i has
 * prod_nv[i] layers (witness.len()), logup spec i has logup_nv[i]+1.
 * Outputs: proofs flattened as rounds: for round=1..max_round:
 *   sumcheck messages (round * 3 ext), then per prod spec present: 2 ext evals,
 *   per logup spec present: 4 ext evals.  point: final rt (max_round+... ext).
 * Returns number of u64 written to `proof`. */
typedef struct { or_transcript* t; } tower_cb_ctx;
OR_API int64_t or_tower_create_proof(uint32_t n_prod, const uint32_t* prod_layers, const uint64_t* const* prod_wit,
                                     uint32_t n_logup, const uint32_t* logup_layers, const uint64_t* const* logup_wit,
                                     or_transcript* tr, uint64_t* proof, uint64_t* point_out, uint32_t* point_len) {
    uint32_t max_round_index = 0;
    for (uint32_t i = 0; i < n_prod; i++) if (prod_layers[i] - 1 > max_round_index) max_round_index = prod_layers[i] - 1;
    for (uint32_t i = 0; i < n_logup; i++) if (logup_layers[i] - 1 > max_round_index) max_round_index = logup_layers[i] - 1;
    uint32_t n_alpha = n_prod + 2 * n_logup;
    ext* alpha = (ext*)malloc(sizeof(ext) * (n_alpha ? n_alpha : 1));
    /* get_challenge_pows: label "combine subset evals", ONE alpha, powers */
    uint64_t a_[2];
    or_tr_sample(tr, "combine subset evals", a_);
    { ext a = E(a_[0], a_[1]), p = EXT_ONE; for (uint32_t i = 0; i < n_alpha; i++) { alpha[i] = p; p = ext_mul(p, a); } }
    /* initial_rt = sample_and_append_vec("product_sum", 1) */
    ext* rt = (ext*)malloc(sizeof(ext) * (max_round_index + 2));
    uint32_t rt_len = 1;
    { uint64_t c[2]; or_tr_sample(tr, "product_sum", c); rt[0] = E(c[0], c[1]); }
    uint64_t w = 0;
    for (uint32_t round = 1; round <= max_round_index; round++) {
        uint32_t nv = rt_len; /* == round */
        uint64_t n = 1ULL << nv;
        /* collect MLE list: eq, then per present prod spec (a,b), per logup (p1,p2,q1,q2) */
        uint32_t cap = 1 + 2 * n_prod + 4 * n_logup;
        or_mle* mles = (or_mle*)malloc(sizeof(or_mle) * cap);
        uint64_t* eq = (uint64_t*)malloc(sizeof(ext) * n);
        or_build_eq_x_r_vec((const uint64_t*)rt, nv, eq);
        uint32_t m = 0;
        mles[m].data = eq; mles[m].num_vars = nv; mles[m].is_ext = 1; m++;
        uint32_t max_terms = n_prod + 3 * n_logup;
        ext* coeff = (ext*)malloc(sizeof(ext) * (max_terms ? max_terms : 1));
        uint32_t* toff = (uint32_t*)malloc(sizeof(uint32_t) * (max_terms + 1));
        uint32_t* tidx = (uint32_t*)malloc(sizeof(uint32_t) * 3 * (max_terms ? max_terms : 1));
        uint32_t nt = 0, q = 0;
        int* prod_present = (int*)calloc(n_prod ? n_prod : 1, sizeof(int));
        int* lk_present = (int*)calloc(n_logup ? n_logup : 1, sizeof(int));
        uint32_t* prod_first = (uint32_t*)malloc(sizeof(uint32_t) * (n_prod ? n_prod : 1));
        uint32_t* lk_first = (uint32_t*)malloc(sizeof(uint32_t) * (n_logup ? n_logup : 1));
        for (uint32_t i = 0; i < n_prod; i++) {
            if (round >= prod_layers[i]) continue; /* spec has no layer `round` */
            prod_present[i] = 1; prod_first[i] = m;
            const uint64_t* base = prod_wit[i] + 2 * (2 * (n - 1));
            mles[m].data = base; mles[m].num_vars = nv; mles[m].is_ext = 1; m++;
            mles[m].data = base + 2 * n; mles[m].num_vars = nv; mles[m].is_ext = 1; m++;
            coeff[nt] = alpha[i]; toff[nt] = q; tidx[q++] = 0; tidx[q++] = m - 2; tidx[q++] = m - 1; nt++;
        }
        for (uint32_t i = 0; i < n_logup; i++) {
            if (round >= logup_layers[i]) continue;
            lk_present[i] = 1; lk_first[i] = m;
            const uint64_t* base = logup_wit[i] + 2 * (4 * (n - 1));
            for (uint32_t z = 0; z < 4; z++) { mles[m].data = base + 2 * n * z; mles[m].num_vars = nv; mles[m].is_ext = 1; m++; }
            uint32_t p1 = m - 4, p2 = m - 3, q1 = m - 2, q2 = m - 1;
            ext an = alpha[n_prod + 2 * i], ad = alpha[n_prod + 2 * i + 1];
            coeff[nt] = an; toff[nt] = q; tidx[q++] = 0; tidx[q++] = p1; tidx[q++] = q2; nt++;
            coeff[nt] = an; toff[nt] = q; tidx[q++] = 0; tidx[q++] = p2; tidx[q++] = q1; nt++;
            coeff[nt] = ad; toff[nt] = q; tidx[q++] = 0; tidx[q++] = q1; tidx[q++] = q2; nt++;
        }
        toff[nt] = q;
        uint64_t* fin = (uint64_t*)malloc(sizeof(ext) * m);
        uint64_t* chal = (uint64_t*)malloc(sizeof(ext) * nv);
        int rc = or_sumcheck_prove_standin(mles, m, (const uint64_t*)coeff, toff, tidx, nt, nv, 3, tr,
                                           proof + w, fin, chal);
        if (rc) return rc;
        w += (uint64_t)nv * 3 * 2;
        for (uint32_t i = 0; i < n_prod; i++) if (prod_present[i]) {
            or_tr_append_ext(tr, fin + 2 * prod_first[i], 2);
            memcpy(proof + w, fin + 2 * prod_first[i], sizeof(ext) * 2); w += 4;
        }
        for (uint32_t i = 0; i < n_logup; i++) if (lk_present[i]) {
            or_tr_append_ext(tr, fin + 2 * lk_first[i], 4);
            memcpy(proof + w, fin + 2 * lk_first[i], sizeof(ext) * 4); w += 8;
        }
        /* rt' = challenges || r_merge */
        uint64_t rm[2]; or_tr_sample(tr, "merge", rm);
        for (uint32_t i = 0; i < nv; i++) rt[i] = E(chal[2 * i], chal[2 * i + 1]);
        rt[nv] = E(rm[0], rm[1]); rt_len = nv + 1;
        or_tr_sample(tr, "combine subset evals", a_);
        { ext a = E(a_[0], a_[1]), p = EXT_ONE; for (uint32_t i = 0; i < n_alpha; i++) { alpha[i] = p; p = ext_mul(p, a); } }
        free(mles); free(eq); free(coeff); free(toff); free(tidx); free(fin); free(chal);
        free(prod_present); free(lk_present); free(prod_first); free(lk_first);
    }
    for (uint32_t i = 0; i < rt_len; i++) { point_out[2 * i] = rt[i].c0; point_out[2 * i + 1] = rt[i].c1; }
    *point_len = rt_len;
    free(alpha); free(rt);
    return (int64_t)w;
}

OR_API void or_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
OR_API int or_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Poseidon2 over Goldilocks, width 8, and the Merkle commitment built on it.
 * PARITY UNPINNED: round constants, the internal diagonal and the Basefold leaf arrangement live in
 * the un-vendored crates (p3-goldilocks 0.4.3 / gkr-backend `poseidon`, `mpcs`), so the
 * constants are PARAMETERS here.  Structure restated from the published Poseidon2 construction as
 * Plonky3 instantiates it for Goldilocks: S-box x^7, 8 external + 22 internal rounds;
 *   external layer: M4 on each 4-chunk, then every lane += the sum of the lanes at the same offset;
 *   internal layer: s = sum(state); state[i] = state[i] * diag[i] + s;
 *   order: external layer, 4 x (rc, sbox all, external), 22 x (rc0, sbox lane 0, internal), 4 x (...).
 * Hashing (p3-symmetric): PaddingFreeSponge<8, rate 4, out 4> for a leaf row (inputs OVERWRITE the
 * rate lanes, permute per chunk), TruncatedPermutation<2, 4, 8> for an inner node. */
#define P2_W 8
#define P2_RF 8
#define P2_RP 22
typedef struct {
    uint64_t ext_rc[P2_RF][P2_W];
    uint64_t int_rc[P2_RP];
    uint64_t diag[P2_W];
    uint32_t mds_variant;   /* 0: circ(2,3,1,1)   1: Horizen-Labs M4 [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]] */
    uint32_t pad;
} or_p2_params;
static inline gl p2_sbox(gl x) { gl x2 = gl_mul(x, x), x3 = gl_mul(x2, x), x4 = gl_mul(x2, x2); return gl_mul(x4, x3); }
static void p2_m4(gl* x, uint32_t variant) {
    static const uint64_t M0[4][4] = {{2, 3, 1, 1}, {1, 2, 3, 1}, {1, 1, 2, 3}, {3, 1, 1, 2}};
    static const uint64_t M1[4][4] = {{5, 7, 1, 3}, {4, 6, 1, 1}, {1, 3, 5, 7}, {1, 1, 4, 6}};
    gl o[4];
    for (int i = 0; i < 4; i++) {
        gl acc = 0;
        for (int j = 0; j < 4; j++) acc = gl_add(acc, gl_mul((variant ? M1 : M0)[i][j], x[j]));
        o[i] = acc;
    }
    for (int i = 0; i < 4; i++) x[i] = o[i];
}
static void p2_external(gl* s, uint32_t variant) {
    p2_m4(s, variant); p2_m4(s + 4, variant);
    for (int i = 0; i < 4; i++) { gl t = gl_add(s[i], s[4 + i]); s[i] = gl_add(s[i], t); s[4 + i] = gl_add(s[4 + i], t); }
}
OR_API void or_poseidon2_permute(const or_p2_params* p, uint64_t* state) {
    gl s[P2_W];
    for (int i = 0; i < P2_W; i++) s[i] = to_canon(state[i]);
    p2_external(s, p->mds_variant);
    for (int r = 0; r < P2_RF / 2; r++) {
        for (int i = 0; i < P2_W; i++) s[i] = p2_sbox(gl_add(s[i], to_canon(p->ext_rc[r][i])));
        p2_external(s, p->mds_variant);
    }
    for (int r = 0; r < P2_RP; r++) {
        s[0] = p2_sbox(gl_add(s[0], to_canon(p->int_rc[r])));
        gl sum = 0;
        for (int i = 0; i < P2_W; i++) sum = gl_add(sum, s[i]);
        for (int i = 0; i < P2_W; i++) s[i] = gl_add(gl_mul(s[i], to_canon(p->diag[i])), sum);
    }
    for (int r = P2_RF / 2; r < P2_RF; r++) {
        for (int i = 0; i < P2_W; i++) s[i] = p2_sbox(gl_add(s[i], to_canon(p->ext_rc[r][i])));
        p2_external(s, p->mds_variant);
    }
    for (int i = 0; i < P2_W; i++) state[i] = s[i];
}
static void p2_hash_row(const or_p2_params* p, const uint64_t* row, uint64_t width, uint64_t* out4) {
    uint64_t st[P2_W] = {0};
    for (uint64_t c = 0; c < width; c += 4) {
        for (uint64_t i = 0; i < 4 && c + i < width; i++) st[i] = to_canon(row[c + i]);
        or_poseidon2_permute(p, st);
    }
    for (int i = 0; i < 4; i++) out4[i] = st[i];
}
static void p2_compress(const or_p2_params* p, const uint64_t* l, const uint64_t* r, uint64_t* out4) {
    uint64_t st[P2_W];
    for (int i = 0; i < 4; i++) { st[i] = l[i]; st[4 + i] = r[i]; }
    or_poseidon2_permute(p, st);
    for (int i = 0; i < 4; i++) out4[i] = st[i];
}
/* matrix: height rows (power of two) x width base elements, row-major.  tree: (2*height - 1) digests of 4 u64,
 * leaves first (level 0 = leaf digests), root last.  Returns the root in root4. */
OR_API int or_merkle_commit(const or_p2_params* p, const uint64_t* matrix, uint64_t width, uint64_t height,
                            uint64_t* tree, uint64_t* root4) {
    if (height == 0 || (height & (height - 1))) return -1;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < height; i++) p2_hash_row(p, matrix + i * width, width, tree + 4 * i);
    uint64_t off = 0, n = height;
    while (n > 1) {
        uint64_t* dst = tree + 4 * (off + n);
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < n / 2; i++) p2_compress(p, tree + 4 * (off + 2 * i), tree + 4 * (off + 2 * i + 1), dst + 4 * i);
        off += n; n /= 2;
    }
    for (int i = 0; i < 4; i++) root4[i] = tree[4 * off + i];
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Rotation pre-passes: BooleanHypercube cyclic group (gkr_iop/src/gkr/booleanhypercube.rs:10-193),
 * rotation_next_base_mle / rotation_selector (gkr_iop/src/utils.rs:19-76).  The group table is the sequence
 * X^i mod (X^5 + X^2 + 1) resp. (X^6 + X + 1), generated here by the shift register instead of a literal table. */
static void bh_table(uint32_t nv, uint64_t* tab) {   /* 2^nv entries: 1, X, X^2, ..., back to 1 */
    const uint64_t modulus = nv == 5 ? 0x25 : 0x43;
    uint64_t cur = 1;
    for (uint64_t i = 0; i < (1ULL << nv); i++) {
        tab[i] = cur;
        cur <<= 1;
        if (cur >> nv) cur ^= modulus;
    }
}
OR_API int or_bh_table(uint32_t nv, uint64_t* tab) { if (nv != 5 && nv != 6) return -1; bh_table(nv, tab); return 0; }
/* literal restatement of the reference loop (including its overwrite order) */
OR_API int or_rotation_next_base_mle(const uint64_t* evals, uint64_t len, uint32_t log2, uint64_t* out) {
    if (log2 != 5 && log2 != 6) return -1;
    const uint64_t size = 1ULL << log2;
    uint64_t idx[64];
    bh_table(log2, idx);
    memset(out, 0, sizeof(uint64_t) * len);
    for (uint64_t c = 0; c + size <= len; c += size) {
        const uint64_t* o = evals + c; uint64_t* r = out + c;
        const uint64_t first = idx[0], last = idx[size - 1];
        if (first == last) r[last] = o[first];
        r[0] = o[0];
        for (int64_t i = (int64_t)size - 2; i >= 0; i--) r[idx[i]] = o[idx[i + 1]];
    }
    return 0;
}
OR_API int or_rotation_selector(const uint64_t* eq, uint64_t total_len, uint32_t subgroup_size, uint32_t log2, uint64_t* out) {
    if ((log2 != 5 && log2 != 6) || subgroup_size > (1u << log2)) return -1;
    const uint64_t size = 1ULL << log2;
    uint64_t idx[64];
    bh_table(log2, idx);
    memset(out, 0, sizeof(ext) * total_len);
    for (uint64_t c = 0; c + size <= total_len; c += size)
        for (int64_t i = (int64_t)subgroup_size - 1; i >= 0; i--) {
            ((ext*)out)[c + idx[i]] = ((const ext*)eq)[c + idx[i]];
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Radix-2 NTT over Goldilocks and Reed-Solomon encoding of witness columns (the RS-encode
 * step of PCS::batch_commit, EXTERNAL mpcs::Basefold over p3-dft; call site ceno_zkvm/src/scheme/cpu/mod.rs:559-584).
 * PARITY UNPINNED for the arrangement (rate, coset, leaf order are upstream-only); what IS fixed is the
 * transform itself: X[k] = sum_j x[j] w^(jk) with w = two_adic_generator(log_n) = g^(2^(32 - log_n)),
 * g = 7^((p-1)/2^32) = 1753635133440165772 (p3-goldilocks TWO_ADIC_GENERATOR; checked in tests against 7^((p-1)/2^32)).
 * Textbook decimation-in-time: bit-reversal permutation, then log_n butterfly stages — deliberately a different
 * formulation from the device's multi-pass decimation-in-frequency kernels. */
#define GL_TWO_ADIC_GEN 1753635133440165772ULL
static uint64_t bitrev64(uint64_t x, uint32_t bits) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1ULL) << (bits - 1 - i);
    return r;
}
OR_API uint64_t or_two_adic_generator(uint32_t bits) { return bits > 32 ? 0 : gl_pow(GL_TWO_ADIC_GEN, 1ULL << (32 - bits)); }
static void ntt_one(gl* a, uint32_t log_n, int inverse) {
    const uint64_t n = 1ULL << log_n;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t j = bitrev64(i, log_n);
        if (i < j) { gl t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (uint32_t s = 1; s <= log_n; s++) {
        const uint64_t m = 1ULL << s, h = m >> 1;
        gl wm = or_two_adic_generator(s);
        if (inverse) wm = gl_inv(wm);
        for (uint64_t k = 0; k < n; k += m) {
            gl w = 1;
            for (uint64_t j = 0; j < h; j++) {
                const gl t = gl_mul(w, a[k + j + h]), u = a[k + j];
                a[k + j] = gl_add(u, t);
                a[k + j + h] = gl_sub(u, t);
                w = gl_mul(w, wm);
            }
        }
    }
    if (inverse) {
        const gl ninv = gl_inv(to_canon(n));
        for (uint64_t i = 0; i < n; i++) a[i] = gl_mul(a[i], ninv);
    }
}
/* data: n_cols columns of 2^log_n base elements, column c at data + c * 2^log_n (column-major), natural order in
 * and out; bitrev != 0: forward writes bit-reversed order / inverse reads bit-reversed order. */
OR_API int or_ntt(uint64_t* data, uint32_t log_n, uint64_t n_cols, int inverse, int bitrev) {
    if (log_n > 32) return -1;
    const uint64_t n = 1ULL << log_n;
#pragma omp parallel for schedule(dynamic)
    for (uint64_t c = 0; c < n_cols; c++) {
        gl* a = data + c * n;
        for (uint64_t i = 0; i < n; i++) a[i] = to_canon(a[i]);
        if (inverse && bitrev)
            for (uint64_t i = 0; i < n; i++) { const uint64_t j = bitrev64(i, log_n); if (i < j) { gl t = a[i]; a[i] = a[j]; a[j] = t; } }
        ntt_one(a, log_n, inverse);
        if (!inverse && bitrev)
            for (uint64_t i = 0; i < n; i++) { const uint64_t j = bitrev64(i, log_n); if (i < j) { gl t = a[i]; a[i] = a[j]; a[j] = t; } }
    }
    return 0;
}
/* Reed-Solomon encode: every column (2^log_n message symbols = the MLE's evaluation vector taken as coefficients)
 * is zero-padded to 2^(log_n + rate_log) and transformed; code column c at out + c * 2^(log_n + rate_log). */
OR_API int or_rs_encode(const uint64_t* msg, uint64_t width, uint32_t log_n, uint32_t rate_log, uint64_t* out, int bitrev) {
    if (log_n + rate_log > 32) return -1;
    const uint64_t n = 1ULL << log_n, m = 1ULL << (log_n + rate_log);
    for (uint64_t c = 0; c < width; c++) {
        memcpy(out + c * m, msg + c * n, n * sizeof(uint64_t));
        memset(out + c * m + n, 0, (m - n) * sizeof(uint64_t));
    }
    return or_ntt(out, log_n + rate_log, width, 0, bitrev);
}
