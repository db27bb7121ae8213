// ecc_host.cuh — host side of the EC-sum Quark prover, included by cabi.cu.
//
// cg_ecc_quark_terms: the zerocheck expression of CpuEccProver::create_ecc_proof (reference
// ceno_zkvm/src/scheme/cpu/mod.rs:153-262) in monomial form — what `expr_builder.to_virtual_polys(&[exprs_add + exprs_bypass +
// export_expr])` hands to IOPProverState::prove — over the MLE order
//     [sel_add, sel_bypass, sel_export, s(7), x0(7), y0(7), x1(7), y1(7), x3(7), y3(7)]
// with the septic-extension products expanded by z^7 = 2z + 5 (ceno_zkvm/src/scheme/septic_curve.rs:681-705):
//     add    (sel_add)   : s (x0 - x1) - (y0 - y1),   s^2 - x0 - x1 - x3,   s (x0 - x3) - (y0 + y3)      alpha[0..21)
//     bypass (sel_bypass): x3 - x0,  y3 - y0                                                            alpha[21..35)
//     export (sel_export): x3 - final_sum.x,  y3 - final_sum.y                                          alpha[35..49)
// Pure host bookkeeping (no device work): terms are merged per monomial and emitted in lexicographic order of their sorted
// factor lists, so the table is deterministic.
#pragma once

namespace ecc_host {
typedef unsigned __int128 u128;
struct E2 { uint64_t c0, c1; };
static inline uint64_t fadd(uint64_t a, uint64_t b) { u128 s = (u128)a + b; return (uint64_t)(s >= GL_P ? s - GL_P : s); }
static inline uint64_t fmul(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) % GL_P); }
static inline E2 scale(E2 a, uint64_t c) { return E2{fmul(a.c0, c), fmul(a.c1, c)}; }
static inline E2 neg(E2 a) { return E2{a.c0 ? GL_P - a.c0 : 0, a.c1 ? GL_P - a.c1 : 0}; }
}   // namespace ecc_host

CG_EXPORT int cg_ecc_quark_terms(const uint64_t* alpha_pows_ext, const uint64_t* final_x, const uint64_t* final_y, uint64_t* coeff_out,
                                 uint32_t* off_out, uint32_t* idx_out, uint32_t cap_terms, uint32_t cap_idx, uint32_t* n_terms, uint32_t* n_idx) {
    using namespace ecc_host;
    if (!alpha_pows_ext || !final_x || !final_y || !n_terms || !n_idx) return CG_ERR_INVALID;
    constexpr uint32_t D = 7, SEL_ADD = 0, SEL_BYP = 1, SEL_EXP = 2;
    enum { Sg = 0, X0, Y0, X1, Y1, X3, Y3 };
    auto V = [](uint32_t group, uint32_t limb) { return 3 + D * group + limb; };
    std::map<std::vector<uint32_t>, E2> acc;
    auto add_term = [&](E2 c, std::vector<uint32_t> f) {
        std::sort(f.begin(), f.end());
        E2& t = acc[f];
        t.c0 = fadd(t.c0, c.c0);
        t.c1 = fadd(t.c1, c.c1);
    };
    auto alpha = [&](uint32_t i) { return E2{alpha_pows_ext[2 * i] % GL_P, alpha_pows_ext[2 * i + 1] % GL_P}; };
    // a (septic) * b (septic), limb k of the product: sum over (i, j) with z^(i+j) reduced by z^7 = 2z + 5
    auto septic_product = [&](E2 a, uint32_t sel, uint32_t ga, uint32_t gb, uint32_t k, bool negate) {
        for (uint32_t i = 0; i < D; i++)
            for (uint32_t j = 0; j < D; j++) {
                const uint32_t d = i + j;
                uint64_t c = 0;
                if (d < D) { if (d == k) c = 1; }
                else { if (d - D == k) c = 5; else if (d - D + 1 == k) c = 2; }
                if (!c) continue;
                E2 t = scale(a, c);
                add_term(negate ? neg(t) : t, {sel, V(ga, i), V(gb, j)});
            }
    };
    uint32_t ai = 0;
    for (uint32_t k = 0; k < D; k++) {   // s (x0 - x1) - (y0 - y1)
        const E2 a = alpha(ai + k);
        septic_product(a, SEL_ADD, Sg, X0, k, false);
        septic_product(a, SEL_ADD, Sg, X1, k, true);
        add_term(neg(a), {SEL_ADD, V(Y0, k)});
        add_term(a, {SEL_ADD, V(Y1, k)});
    }
    ai += D;
    for (uint32_t k = 0; k < D; k++) {   // s^2 - x0 - x1 - x3
        const E2 a = alpha(ai + k);
        septic_product(a, SEL_ADD, Sg, Sg, k, false);
        for (uint32_t g : {(uint32_t)X0, (uint32_t)X1, (uint32_t)X3}) add_term(neg(a), {SEL_ADD, V(g, k)});
    }
    ai += D;
    for (uint32_t k = 0; k < D; k++) {   // s (x0 - x3) - (y0 + y3)
        const E2 a = alpha(ai + k);
        septic_product(a, SEL_ADD, Sg, X0, k, false);
        septic_product(a, SEL_ADD, Sg, X3, k, true);
        add_term(neg(a), {SEL_ADD, V(Y0, k)});
        add_term(neg(a), {SEL_ADD, V(Y3, k)});
    }
    ai += D;
    for (int pass = 0; pass < 2; pass++) {   // bypass: x3 - x0, y3 - y0
        const uint32_t g3 = pass ? Y3 : X3, g0 = pass ? Y0 : X0;
        for (uint32_t k = 0; k < D; k++) {
            const E2 a = alpha(ai + k);
            add_term(a, {SEL_BYP, V(g3, k)});
            add_term(neg(a), {SEL_BYP, V(g0, k)});
        }
        ai += D;
    }
    for (int pass = 0; pass < 2; pass++) {   // export: x3 - final_sum.x, y3 - final_sum.y
        const uint32_t g3 = pass ? Y3 : X3;
        const uint64_t* fin = pass ? final_y : final_x;
        for (uint32_t k = 0; k < D; k++) {
            const E2 a = alpha(ai + k);
            add_term(a, {SEL_EXP, V(g3, k)});
            add_term(neg(scale(a, fin[k] % GL_P)), {SEL_EXP});
        }
        ai += D;
    }
    uint32_t nt = 0, ni = 0;
    for (const auto& kv : acc) {
        if (kv.second.c0 == 0 && kv.second.c1 == 0) continue;
        if (coeff_out && off_out && idx_out) {
            if (nt >= cap_terms || ni + kv.first.size() > cap_idx) return CG_ERR_INVALID;
            coeff_out[2 * nt] = kv.second.c0;
            coeff_out[2 * nt + 1] = kv.second.c1;
            off_out[nt] = ni;
            for (uint32_t f : kv.first) idx_out[ni++] = f;
        } else {
            ni += (uint32_t)kv.first.size();
        }
        nt++;
    }
    if (off_out && nt <= cap_terms) off_out[nt] = ni;
    *n_terms = nt;
    *n_idx = ni;
    return CG_OK;
}
