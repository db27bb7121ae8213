// ntt_host.cuh — host side of cg_ntt / cg_rs_encode (included by cabi.cu; kernels in ntt_kernels.cuh).
#pragma once

static uint64_t host_gl_mul(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) % GL_P); }
static uint64_t host_gl_pow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    while (e) { if (e & 1) r = host_gl_mul(r, a); a = host_gl_mul(a, a); e >>= 1; }
    return r;
}
#define CG_GL_TWO_ADIC_GEN 1753635133440165772ULL   /* 7^((p-1)/2^32): p3-goldilocks two_adic_generator(32) */

static int ntt_tables(cg_ctx* c, cudaStream_t st) {
    std::lock_guard<std::mutex> g(c->mu);
    if (c->d_ntt_tab) return CG_OK;
    const size_t nA = 1u << CG_NTT_A_BITS, nB = 1u << (CG_NTT_MAX_LOG - CG_NTT_A_BITS), nW = 4096;
    uint64_t* tab = nullptr;
    cudaError_t e = cudaMalloc((void**)&tab, (nA + nB + nW) * sizeof(uint64_t));
    if (e != cudaSuccess) { c->err = std::string("cudaMalloc(ntt tables): ") + cudaGetErrorString(e); return CG_ERR_OOM; }
    const uint64_t omega27 = host_gl_pow(CG_GL_TWO_ADIC_GEN, 1ULL << (32 - CG_NTT_MAX_LOG));
    ntt_tables_kernel<<<(unsigned)(nB / 256), 256, 0, st>>>(tab, tab + nA, tab + nA + nB, omega27);
    c->launches++;
    e = cudaStreamSynchronize(st);   // other streams may use the tables right after this returns
    if (e != cudaSuccess) { cudaFree(tab); c->err = std::string("ntt_tables_kernel: ") + cudaGetErrorString(e); return CG_ERR_CUDA; }
    c->d_ntt_tab = tab;
    return CG_OK;
}

// the instantiated digit geometries: last digit (C = 0, T = 1..12) and leading digits (T = 1..9, C = 12 - T)
template <int T, int C>
static void ntt_launch_tc(const NttPassArgs& a, unsigned tiles, cudaStream_t st) {
    if (a.inverse) ntt_pass_kernel<T, C, false><<<tiles, 256, 0, st>>>(a);
    else ntt_pass_kernel<T, C, true><<<tiles, 256, 0, st>>>(a);
}
static bool ntt_launch(const NttPassArgs& a, unsigned tiles, cudaStream_t st) {
#define CG_NTT_CASE(T_, C_) if (a.T == T_ && a.C == C_) { ntt_launch_tc<T_, C_>(a, tiles, st); return true; }
    CG_NTT_CASE(1, 0) CG_NTT_CASE(2, 0) CG_NTT_CASE(3, 0) CG_NTT_CASE(4, 0) CG_NTT_CASE(5, 0) CG_NTT_CASE(6, 0)
    CG_NTT_CASE(7, 0) CG_NTT_CASE(8, 0) CG_NTT_CASE(9, 0) CG_NTT_CASE(10, 0) CG_NTT_CASE(11, 0) CG_NTT_CASE(12, 0)
    CG_NTT_CASE(1, 11) CG_NTT_CASE(2, 10) CG_NTT_CASE(3, 9) CG_NTT_CASE(4, 8) CG_NTT_CASE(5, 7) CG_NTT_CASE(6, 6)
    CG_NTT_CASE(7, 5) CG_NTT_CASE(8, 4) CG_NTT_CASE(9, 3)
#undef CG_NTT_CASE
    return false;
}

// digits of log_n, top digit first: the last (lowest) digit takes up to 12 bits, the rest is split evenly (<= 8 each)
static int ntt_plan(uint32_t log_n, uint32_t digits[4]) {
    if (log_n <= CG_NTT_TILE_LOG) { digits[0] = log_n; return 1; }
    const uint32_t rest = log_n - CG_NTT_TILE_LOG;
    if (rest <= 9) { digits[0] = rest; digits[1] = CG_NTT_TILE_LOG; return 2; }
    digits[0] = (rest + 1) / 2; digits[1] = rest / 2; digits[2] = CG_NTT_TILE_LOG;
    return 3;
}

// one limb array: passes over `n_cols` columns.  in != out only for the first pass (RS zero padding reads a shorter input).
static int ntt_run(cg_ctx* c, const uint64_t* in, uint64_t in_col_stride, uint64_t in_len, uint64_t* out, uint64_t out_col_stride,
                   uint32_t estride, uint32_t log_n, uint64_t n_cols, bool inverse, cudaStream_t st) {
    uint32_t digits[4];
    const int np = ntt_plan(log_n, digits);
    uint32_t los[4];
    {
        uint32_t hi = log_n;
        for (int p = 0; p < np; p++) { hi -= digits[p]; los[p] = hi; }
    }
    const uint64_t n = 1ULL << log_n;
    NttPassArgs a;
    memset(&a, 0, sizeof(a));
    a.estride = estride;
    a.log_n = log_n;
    a.inverse = inverse ? 1 : 0;
    a.n_inv = host_gl_pow(n % GL_P, GL_P - 2);
    a.A = c->d_ntt_tab;
    a.B = a.A + (1u << CG_NTT_A_BITS);
    a.W12 = a.B + (1u << (CG_NTT_MAX_LOG - CG_NTT_A_BITS));
    for (int step = 0; step < np; step++) {
        const int p = inverse ? np - 1 - step : step;
        a.in = step == 0 ? in : out;
        a.in_col_stride = step == 0 ? in_col_stride : out_col_stride;
        a.in_len = step == 0 ? in_len : n;
        a.out = out;
        a.out_col_stride = out_col_stride;
        a.T = digits[p];
        a.lo = los[p];
        a.C = std::min<uint32_t>(CG_NTT_TILE_LOG - a.T, a.lo);
        a.scale = (inverse && step == np - 1) ? 1 : 0;
        const uint64_t tiles = n_cols << (log_n - a.T - a.C);
        if (tiles > 0x7FFFFFFFULL) return set_err(c, CG_ERR_INVALID, "cg_ntt: too many tiles for one launch");
        if (!ntt_launch(a, (unsigned)tiles, st)) return set_err(c, CG_ERR_STATE, "cg_ntt: no kernel for this digit geometry");
        LAUNCHED(c);
    }
    CU(c, cudaGetLastError());
    return CG_OK;
}

CG_EXPORT int cg_ntt(cg_ctx* c, uint64_t* d_data, uint32_t log_n, uint64_t n_cols, uint64_t col_stride, uint32_t flags, cg_stream s) {
    if (!c || !d_data) return CG_ERR_INVALID;
    if (log_n > CG_NTT_MAX_LOG) return set_err(c, CG_ERR_UNSUPPORTED, "cg_ntt: log_n > 27 (MAX_NUM_VARIABLES 24 + rate_log 3)");
    if (col_stride < (1ULL << log_n)) return set_err(c, CG_ERR_INVALID, "cg_ntt: col_stride < 2^log_n");
    if (n_cols == 0 || log_n == 0) return CG_OK;
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = S(c, s);
    CHK(ntt_tables(c, st));
    const bool inverse = flags & CG_NTT_INVERSE, bitrev = flags & CG_NTT_BITREV, ext = flags & CG_NTT_EXT;
    const uint32_t estride = ext ? 2 : 1;
    const unsigned pg = grid_for(c, n_cols << log_n, 8);
    // the kernels map natural -> bit-reversed (forward) and bit-reversed -> natural (inverse)
    for (uint32_t limb = 0; limb < estride; limb++) {
        uint64_t* p = d_data + limb;
        if (inverse && !bitrev) { ntt_bitrev_kernel<<<pg, 256, 0, st>>>(p, log_n, n_cols, col_stride, estride); LAUNCHED(c); }
        CHK(ntt_run(c, p, col_stride, 1ULL << log_n, p, col_stride, estride, log_n, n_cols, inverse, st));
        if (!inverse && !bitrev) { ntt_bitrev_kernel<<<pg, 256, 0, st>>>(p, log_n, n_cols, col_stride, estride); LAUNCHED(c); }
    }
    CU(c, cudaGetLastError());
    return CG_OK;
}

CG_EXPORT int cg_rs_encode(cg_ctx* c, const uint64_t* d_msg, uint64_t width, uint32_t log_n, uint32_t rate_log, uint64_t* d_code,
                           uint32_t flags, cg_stream s) {
    if (!c || !d_msg || !d_code) return CG_ERR_INVALID;
    if (flags & (CG_NTT_INVERSE | CG_NTT_EXT)) return set_err(c, CG_ERR_INVALID, "cg_rs_encode: only CG_NTT_BITREV is accepted");
    const uint32_t L = log_n + rate_log;
    if (L > CG_NTT_MAX_LOG) return set_err(c, CG_ERR_UNSUPPORTED, "cg_rs_encode: log_n + rate_log > 27");
    if (width == 0) return CG_OK;
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = S(c, s);
    if (L == 0) { CU(c, cudaMemcpyAsync(d_code, d_msg, width * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st)); return CG_OK; }
    CHK(ntt_tables(c, st));
    // the first pass reads the message (2^log_n symbols per column, the rest of the 2^L block is zero) and writes the code buffer
    CHK(ntt_run(c, d_msg, 1ULL << log_n, 1ULL << log_n, d_code, 1ULL << L, 1, L, width, false, st));
    if (!(flags & CG_NTT_BITREV)) {
        ntt_bitrev_kernel<<<grid_for(c, width << L, 8), 256, 0, st>>>(d_code, L, width, 1ULL << L, 1);
        LAUNCHED(c);
    }
    CU(c, cudaGetLastError());
    return CG_OK;
}
