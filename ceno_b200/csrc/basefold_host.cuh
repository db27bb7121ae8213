// basefold_host.cuh — host side of the Basefold entry points (included by cabi.cu; kernels in basefold.cuh).
#pragma once

struct cg_pcs_commitment {
    cg_ctx* ctx = nullptr;
    uint64_t width = 0;
    uint32_t nv = 0, rate_log = 0;
    const uint64_t* d_msg = nullptr;   // the caller's column-major evaluation matrix (must outlive the opening)
    uint64_t* d_code = nullptr;        // width x 2^(nv + rate_log), column-major, rows in bit-reversed order
    uint64_t* d_tree = nullptr;        // 2 * 2^(nv + rate_log) - 1 digests, leaf level first
    uint64_t root[4] = {0, 0, 0, 0};
};

CG_EXPORT int cg_basefold_commit(cg_ctx* c, const uint64_t* d_msg, uint64_t width, uint32_t num_vars, const cg_basefold_params* prm,
                                 cg_stream s, cg_pcs_commitment** out) {
    if (!c || !d_msg || !prm || !out || width == 0) return set_err(c, CG_ERR_INVALID, "cg_basefold_commit: bad argument");
    if (num_vars == 0) return set_err(c, CG_ERR_UNSUPPORTED, "cg_basefold_commit: num_vars must be >= 1 (an opening needs at least one fold round)");
    if (num_vars + prm->rate_log > CG_NTT_MAX_LOG) return set_err(c, CG_ERR_UNSUPPORTED, "cg_basefold_commit: num_vars + rate_log > 27");
    if (!c->d_p2) return set_err(c, CG_ERR_STATE, "cg_poseidon2_set_params has not been called");
    cg_pcs_commitment* cm = new cg_pcs_commitment();
    cm->ctx = c; cm->width = width; cm->nv = num_vars; cm->rate_log = prm->rate_log; cm->d_msg = d_msg;
    const uint64_t h = 1ULL << (num_vars + prm->rate_log);
    int rc = cg_alloc(c, sizeof(uint64_t) * width * h, (void**)&cm->d_code);
    if (rc == CG_OK) rc = cg_alloc(c, 32 * (2 * h - 1), (void**)&cm->d_tree);
    if (rc == CG_OK) rc = cg_rs_encode(c, d_msg, width, num_vars, prm->rate_log, cm->d_code, CG_NTT_BITREV, s);
    if (rc == CG_OK) rc = cg_merkle_commit(c, cm->d_code, width, h, 1, cm->d_tree, cm->root, s);
    if (rc != CG_OK) {
        if (cm->d_code) cg_free(c, cm->d_code);
        if (cm->d_tree) cg_free(c, cm->d_tree);
        delete cm;
        return rc;
    }
    *out = cm;
    return CG_OK;
}
CG_EXPORT int cg_basefold_commitment_root(const cg_pcs_commitment* cm, uint64_t h_root[4]) {
    if (!cm || !h_root) return CG_ERR_INVALID;
    memcpy(h_root, cm->root, 32);
    return CG_OK;
}
CG_EXPORT int cg_basefold_commitment_codeword(const cg_pcs_commitment* cm, const uint64_t** d_code, const uint64_t** d_tree) {
    if (!cm) return CG_ERR_INVALID;
    if (d_code) *d_code = cm->d_code;
    if (d_tree) *d_tree = cm->d_tree;
    return CG_OK;
}
CG_EXPORT int cg_basefold_commitment_free(cg_pcs_commitment* cm) {
    if (!cm) return CG_ERR_INVALID;
    cudaSetDevice(cm->ctx->device);
    cudaDeviceSynchronize();
    cg_free(cm->ctx, cm->d_code);
    cg_free(cm->ctx, cm->d_tree);
    delete cm;
    return CG_OK;
}

// ---- stand-in transcript events of the PCS (same sponge as cg_standin_*; NOT the Poseidon2 duplex challenger)
static void spcs_label(void* u, const char* l) { cg_tr_append_message(*(uint64_t*)u, (const uint8_t*)l, strlen(l)); }
static void spcs_sample(void* u, uint64_t o[2]) { o[0] = cg_tr_squeeze(*(uint64_t*)u); o[1] = cg_tr_squeeze(*(uint64_t*)u); }
static void spcs_obs_ext(void* u, const uint64_t* e, uint64_t n) { for (uint64_t i = 0; i < 2 * n; i++) cg_tr_absorb(*(uint64_t*)u, e[i]); }
static void spcs_obs_base(void* u, const uint64_t* e, uint64_t n) { for (uint64_t i = 0; i < n; i++) cg_tr_absorb(*(uint64_t*)u, e[i]); }
static uint64_t spcs_bits(void* u, uint32_t bits) { uint64_t o[2]; spcs_sample(u, o); return bits >= 64 ? o[0] : (o[0] & ((1ULL << bits) - 1)); }
static uint64_t spcs_grind(void* u, uint32_t bits) {
    if (bits == 0) return 0;
    for (uint64_t w = 0;; w++) {
        uint64_t probe = *(uint64_t*)u;
        const uint64_t ww[2] = {w, 0};
        spcs_obs_base(&probe, ww, 2);
        if (spcs_bits(&probe, bits) == 0) { spcs_obs_base(u, ww, 2); (void)spcs_bits(u, bits); return w; }
    }
}
CG_EXPORT void cg_standin_pcs_vt(uint64_t* st, cg_pcs_transcript_vt* o) {
    o->user = st;
    o->observe_label = spcs_label;
    o->sample_ext = spcs_sample;
    o->observe_exts = spcs_obs_ext;
    o->observe_base = spcs_obs_base;
    o->sample_bits = spcs_bits;
    o->grind = spcs_grind;
}

static int bf_shape(const cg_basefold_opening* ops, uint32_t n_ops, const cg_basefold_params* prm, uint32_t* max_nv, uint64_t* words) {
    uint32_t mx = 0;
    for (uint32_t i = 0; i < n_ops; i++) {
        if (!ops[i].commit || ops[i].commit->rate_log != prm->rate_log || ops[i].commit->nv == 0) return CG_ERR_INVALID;
        mx = std::max(mx, ops[i].commit->nv);
    }
    const uint32_t lh = mx + prm->rate_log;
    uint64_t w = (uint64_t)mx * 4 + (uint64_t)mx * 4 + 2ULL * n_ops + 1;
    uint64_t per_q = 1;
    for (uint32_t i = 0; i < n_ops; i++) per_q += ops[i].commit->width + 4ULL * (ops[i].commit->nv + prm->rate_log);
    for (uint32_t r = 0; r < mx; r++) per_q += 2 + 4ULL * (lh - r - 1);
    w += per_q * prm->n_queries;
    *max_nv = mx;
    *words = w;
    return CG_OK;
}
CG_EXPORT uint64_t cg_basefold_proof_len(const cg_basefold_opening* ops, uint32_t n_ops, const cg_basefold_params* prm) {
    uint32_t mx = 0;
    uint64_t w = 0;
    if (!ops || !prm || n_ops == 0 || bf_shape(ops, n_ops, prm, &mx, &w) != CG_OK) return 0;
    return w;
}

CG_EXPORT int cg_basefold_batch_open(cg_ctx* c, const cg_basefold_opening* ops, uint32_t n_ops, const cg_basefold_params* prm,
                                     const cg_pcs_transcript_vt* tr, uint64_t* h_proof, uint64_t proof_cap_words, cg_stream s) {
    if (!c || !ops || !n_ops || !prm || !tr || !h_proof) return set_err(c, CG_ERR_INVALID, "cg_basefold_batch_open: null argument");
    uint32_t max_nv = 0;
    uint64_t words = 0;
    if (bf_shape(ops, n_ops, prm, &max_nv, &words) != CG_OK)
        return set_err(c, CG_ERR_INVALID, "cg_basefold_batch_open: every opening needs a commitment with num_vars >= 1 and the same rate_log");
    if (proof_cap_words < words) return set_err(c, CG_ERR_INVALID, "cg_basefold_batch_open: proof buffer too small (cg_basefold_proof_len)");
    for (uint32_t i = 0; i < n_ops; i++)
        if (!ops[i].h_point_ext || !ops[i].h_evals_ext) return set_err(c, CG_ERR_INVALID, "cg_basefold_batch_open: opening without point / evaluations");
    cudaStream_t st = S(c, s);
    CU(c, cudaSetDevice(c->device));
    const uint32_t rate = prm->rate_log, lh_max = max_nv + rate, num_rounds = max_nv;

    // ---- batch coefficients 1, a, a^2, ... over every committed column, in opening order
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_ops; i++) total += ops[i].commit->width;
    tr->observe_label(tr->user, "batch coeffs");
    uint64_t al[2];
    tr->sample_ext(tr->user, al);
    const ext_t alpha{al[0] % GL_P, al[1] % GL_P};
    std::vector<ext_t> coeffs(total);
    {
        ext_t acc{1, 0};
        for (uint64_t i = 0; i < total; i++) { coeffs[i] = acc; acc = hx_mul(acc, alpha); }
    }
    std::vector<void*> owned;       // pooled device buffers of this call
    std::vector<cg_sumcheck*> scs(n_ops, nullptr);
    auto cleanup = [&]() {
        for (auto* sc : scs) if (sc) cg_sumcheck_destroy(sc);
        cudaStreamSynchronize(st);
        for (void* p : owned) cg_free(c, p);
    };
    auto alloc = [&](size_t bytes, void** p) -> int {
        int rc = cg_alloc(c, bytes, p);
        if (rc == CG_OK) owned.push_back(*p);
        return rc;
    };
#define BF_CHK(expr) do { int rc__ = (expr); if (rc__ != CG_OK) { cleanup(); return rc__; } } while (0)
    ext_t* d_coeff = nullptr;
    BF_CHK(alloc(sizeof(ext_t) * total, (void**)&d_coeff));
    if (cudaMemcpyAsync(d_coeff, coeffs.data(), sizeof(ext_t) * total, cudaMemcpyHostToDevice, st) != cudaSuccess) { cleanup(); return set_err(c, CG_ERR_CUDA, "coefficient upload failed"); }

    // ---- per opening: g = RLC of its columns (ext), eq(point, .), the claimed sum S = sum_j coeff_j eval_j, its sumcheck
    std::vector<ext_t*> d_g(n_ops, nullptr), d_eq(n_ops, nullptr);
    std::vector<ext_t> S(n_ops, ext_t{0, 0});
    std::vector<uint64_t> cofs(n_ops, 0);
    {
        uint64_t off = 0;
        const uint64_t one[2] = {1, 0};
        const uint32_t t_off[2] = {0, 2}, t_idx[2] = {0, 1};
        for (uint32_t i = 0; i < n_ops; i++) {
            const cg_pcs_commitment* cm = ops[i].commit;
            const uint64_t n = 1ULL << cm->nv;
            cofs[i] = off;
            BF_CHK(alloc(sizeof(ext_t) * n, (void**)&d_g[i]));
            BF_CHK(alloc(sizeof(ext_t) * n, (void**)&d_eq[i]));
            bf_rlc_cols_kernel<<<grid_for(c, n, 8), CG_THREADS, 0, st>>>(cm->d_msg, n, (uint32_t)cm->width, d_coeff + off, n, d_g[i], 0);
            LAUNCHED(c);
            BF_CHK(cg_build_eq(c, ops[i].h_point_ext, cm->nv, (uint64_t*)d_eq[i], 0, n, s));
            for (uint64_t j = 0; j < cm->width; j++)
                S[i] = hx_add(S[i], hx_mul(coeffs[off + j], ext_t{ops[i].h_evals_ext[2 * j] % GL_P, ops[i].h_evals_ext[2 * j + 1] % GL_P}));
            cg_mle_desc md[2] = {{d_eq[i], n, cm->nv, CG_MLE_EXT}, {d_g[i], n, cm->nv, CG_MLE_EXT}};
            BF_CHK(cg_sumcheck_create(c, md, 2, one, t_off, t_idx, 1, cm->nv, 2, CG_SC_DEFAULT, s, &scs[i]));
            off += cm->width;
        }
    }
    // ---- running codeword: RLC of the tallest codewords; fold twiddles
    std::vector<ext_t*> cw(num_rounds + 1, nullptr);
    std::vector<uint64_t*> trees(num_rounds, nullptr);
    auto add_codewords = [&](uint32_t nv, ext_t* dst, bool first) {
        const uint64_t h = 1ULL << (nv + rate);
        for (uint32_t i = 0; i < n_ops; i++) {
            const cg_pcs_commitment* cm = ops[i].commit;
            if (cm->nv != nv) continue;
            bf_rlc_cols_kernel<<<grid_for(c, h, 8), CG_THREADS, 0, st>>>(cm->d_code, h, (uint32_t)cm->width, d_coeff + cofs[i], h, dst, first ? 0 : 1);
            LAUNCHED(c);
            first = false;
        }
    };
    BF_CHK(alloc(sizeof(ext_t) << lh_max, (void**)&cw[0]));
    add_codewords(max_nv, cw[0], true);
    uint64_t* d_tw = nullptr;
    BF_CHK(alloc(sizeof(uint64_t) << (lh_max - 1), (void**)&d_tw));
    {
        BfTwArgs ta;
        memset(&ta, 0, sizeof(ta));
        const uint64_t g = hx_powmod(hx_powmod(7, (GL_P - 1) >> 32), 1ULL << (32 - lh_max));   // two_adic_generator(lh_max)
        uint64_t pw = hx_powmod(g, GL_P - 2);
        for (uint32_t b = 0; b < 32; b++) { ta.pw[b] = pw; pw = hx_mulmod(pw, pw); }
        ta.inv2 = (GL_P + 1) / 2;
        ta.bits = lh_max - 1;
        ta.tw = d_tw;
        bf_twiddle_kernel<<<grid_for(c, 1ULL << ta.bits, 8), CG_THREADS, 0, st>>>(ta);
        LAUNCHED(c);
    }

    // ---- rounds
    uint64_t* p_sum = h_proof;
    uint64_t* p_com = h_proof + 4ULL * num_rounds;
    uint64_t* p_fin = p_com + 4ULL * num_rounds;
    uint64_t* p_pow = p_fin + 2ULL * n_ops;
    uint64_t* p_q = p_pow + 1;
    for (uint32_t r = 0; r < num_rounds; r++) {
        ext_t e1{0, 0}, e2{0, 0};
        for (uint32_t i = 0; i < n_ops; i++) {
            const uint32_t join = max_nv - ops[i].commit->nv;
            if (r < join) {   // not joined yet: constant in X, 2^(join - r - 1) copies of its claimed sum
                const uint64_t k = hx_powmod(2, join - r - 1);
                const ext_t v{hx_mulmod(S[i].c0, k), hx_mulmod(S[i].c1, k)};
                e1 = hx_add(e1, v);
                e2 = hx_add(e2, v);
            } else {
                uint64_t m[4];
                BF_CHK(cg_sumcheck_round_eval(scs[i], m));
                e1 = hx_add(e1, ext_t{m[0], m[1]});
                e2 = hx_add(e2, ext_t{m[2], m[3]});
            }
        }
        uint64_t* m = p_sum + 4ULL * r;
        m[0] = e1.c0; m[1] = e1.c1; m[2] = e2.c0; m[3] = e2.c1;
        tr->observe_exts(tr->user, m, 2);
        tr->observe_label(tr->user, "commit round");
        uint64_t ch[2];
        tr->sample_ext(tr->user, ch);
        const ext_t chal{ch[0] % GL_P, ch[1] % GL_P};
        // commit the current codeword as (even, odd) pairs: a row-major matrix of 4 base elements per leaf
        const uint64_t h = 1ULL << (lh_max - r);
        BF_CHK(alloc(32 * (h - 1), (void**)&trees[r]));
        BF_CHK(cg_merkle_commit(c, (const uint64_t*)cw[r], 4, h / 2, 0, trees[r], p_com + 4ULL * r, s));
        tr->observe_base(tr->user, p_com + 4ULL * r, 4);
        // fold; codewords of the next height join
        BF_CHK(alloc(sizeof(ext_t) * (h / 2), (void**)&cw[r + 1]));
        bf_fold_kernel<<<grid_for(c, h / 2, 8), CG_THREADS, 0, st>>>(cw[r], h / 2, chal, d_tw, cw[r + 1]);
        LAUNCHED(c);
        if (r + 1 < num_rounds) add_codewords(max_nv - r - 1, cw[r + 1], false);
        for (uint32_t i = 0; i < n_ops; i++)
            if (r >= max_nv - ops[i].commit->nv) BF_CHK(cg_sumcheck_bind(scs[i], ch));
    }
    if (cudaGetLastError() != cudaSuccess) { cleanup(); return set_err(c, CG_ERR_CUDA, "basefold kernel launch failed"); }
    // ---- final message: one element per opening (basecode_log = 0): g_p at the fold challenges
    for (uint32_t i = 0; i < n_ops; i++) {
        uint64_t fin[4];
        BF_CHK(cg_sumcheck_final_evals(scs[i], fin));
        p_fin[2 * i] = fin[2];
        p_fin[2 * i + 1] = fin[3];
    }
    tr->observe_exts(tr->user, p_fin, n_ops);
    *p_pow = prm->pow_bits ? tr->grind(tr->user, prm->pow_bits) : 0;
    tr->observe_label(tr->user, "query indices");
    // ---- queries: one gather launch over a descriptor list
    std::vector<BfCopy> list;
    uint64_t w = 0;
    std::vector<uint64_t> qidx(prm->n_queries);
    for (uint32_t q = 0; q < prm->n_queries; q++) {
        const uint64_t idx0 = tr->sample_bits(tr->user, lh_max);
        qidx[q] = idx0;
        w += 1;   // the index itself is written by the host below
        for (uint32_t i = 0; i < n_ops; i++) {
            const cg_pcs_commitment* cm = ops[i].commit;
            const uint32_t lh = cm->nv + rate;
            const uint64_t hh = 1ULL << lh, red = idx0 >> (max_nv - cm->nv);
            list.push_back(BfCopy{cm->d_code + red, w, hh, cm->width});
            w += cm->width;
            uint64_t lvl_off = 0, lvl_n = hh, x = red;
            for (uint32_t l = 0; l < lh; l++) {
                list.push_back(BfCopy{cm->d_tree + 4 * (lvl_off + (x ^ 1)), w, 1, 4});
                w += 4;
                lvl_off += lvl_n; lvl_n >>= 1; x >>= 1;
            }
        }
        uint64_t idx = idx0;
        for (uint32_t r = 0; r < num_rounds; r++) {
            list.push_back(BfCopy{(const uint64_t*)(cw[r] + (idx ^ 1)), w, 1, 2});
            w += 2;
            const uint32_t depth = lh_max - r - 1;
            uint64_t lvl_off = 0, lvl_n = 1ULL << depth, x = idx >> 1;
            for (uint32_t l = 0; l < depth; l++) {
                list.push_back(BfCopy{trees[r] + 4 * (lvl_off + (x ^ 1)), w, 1, 4});
                w += 4;
                lvl_off += lvl_n; lvl_n >>= 1; x >>= 1;
            }
            idx >>= 1;
        }
    }
    if (w) {
        BfCopy* d_list = nullptr;
        uint64_t* d_out = nullptr;
        BF_CHK(alloc(sizeof(BfCopy) * list.size(), (void**)&d_list));
        BF_CHK(alloc(sizeof(uint64_t) * w, (void**)&d_out));
        bool ok = cudaMemcpyAsync(d_list, list.data(), sizeof(BfCopy) * list.size(), cudaMemcpyHostToDevice, st) == cudaSuccess;
        if (ok) {
            bf_gather_kernel<<<grid_for(c, list.size(), 8), CG_THREADS, 0, st>>>(d_list, list.size(), d_out);
            LAUNCHED(c);
            ok = cudaMemcpyAsync(p_q, d_out, sizeof(uint64_t) * w, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
        }
        if (!ok) { cleanup(); return set_err(c, CG_ERR_CUDA, "basefold query gather failed"); }
        const uint64_t per_q = w / prm->n_queries;
        for (uint32_t q = 0; q < prm->n_queries; q++) p_q[per_q * q] = qidx[q];
    }
#undef BF_CHK
    cleanup();
    return CG_OK;
}
