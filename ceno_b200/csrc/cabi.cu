// cabi.cu — host side of the C ABI declared in include/ceno_b200.h.
//
// Owns: context (device, stream, pooled allocator, pinned staging), the sumcheck round loop
// (the device-side counterpart of IOPProverState, external sumcheck crate),
// eq/selector builders, tower build + prover.  All arithmetic runs in the kernels of
// sumcheck_kernels.cuh; there is no CPU fallback — every entry point fails with an error code
// when the device is missing.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ceno_b200.h"
#include "sumcheck_kernels.cuh"
#include "poseidon2.cuh"
#include "ntt_kernels.cuh"
#include "basefold.cuh"

#define CG_EXPORT extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------------------------------ ctx
struct cg_ctx {
    int device = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t max_smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    std::mutex mu;
    std::string err;
    std::multimap<size_t, void*> free_blocks;
    struct PendingFree { void* p; size_t sz; cudaEvent_t ev; };
    std::vector<PendingFree> pending_free;               // cg_free_async: blocks whose owning stream may still use them
    std::unordered_map<void*, size_t> live;
    size_t used = 0, reserved = 0;
    std::atomic<uint64_t> launches{0};
    std::atomic<int> live_sc{0};                          // sumchecks alive right now (lanes): persistent kernels shrink when > 1
    std::vector<std::pair<size_t, void*>> pinned_cache;   // reusable pinned staging buffers
    P2Params* d_p2 = nullptr;                             // Poseidon2 constants (caller-supplied, cg_poseidon2_set_params)
    std::vector<float> profile_ms;                        // per-round device time of the last CG_SC_PROFILE run
    uint64_t* d_ntt_tab = nullptr;                        // NTT twiddle tables A | B | W12 (lazy, cg_ntt)
    uint32_t tail_max_c = 1;                              // largest cluster the tail kernel can be launched with (16, 8, ... 1)
    bool host_wait_ok = true;                             // false: kernel launches block the host (profiler / sanitizer) — no kernel may wait for the host
    unsigned long long wait_timeout_cycles = 8000000000ULL;   // device-side limit of every wait on the host or a peer (CG_WAIT_TIMEOUT_MS)
};

static int set_err(cg_ctx* c, int code, const std::string& msg) {
    if (c) { std::lock_guard<std::mutex> g(c->mu); c->err = msg; }
    return code;
}
#define CU(c, call)                                                                                  \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return set_err((c), e__ == cudaErrorMemoryAllocation ? CG_ERR_OOM : CG_ERR_CUDA,         \
                           std::string(#call) + ": " + cudaGetErrorString(e__));                     \
    } while (0)
#define CHK(expr)                     \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != CG_OK) return rc__; \
    } while (0)

static inline cudaStream_t S(cg_ctx* c, cg_stream s) { return s ? (cudaStream_t)s : c->own_stream; }
static inline unsigned grid_for(cg_ctx* c, uint64_t items, unsigned per_sm = 4) {
    uint64_t b = (items + CG_THREADS - 1) / CG_THREADS;
    uint64_t cap = (uint64_t)c->sm_count * per_sm;
    if (cap > CG_MAX_BLOCKS) cap = CG_MAX_BLOCKS;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}
#define LAUNCHED(c) ((c)->launches++)

CG_EXPORT const char* cg_version(void) { return "ceno_b200 0.1 (sm_100a)"; }

// Per-context kernel preparation.  (1) Dynamic shared-memory opt-ins are function attributes, i.e. process-global state:
// they are raised ONCE here to the device maximum instead of per launch (two lanes setting different sizes for the same
// kernel would race).  (2) cudaFuncGetAttributes forces the load of the persistent kernels now: with lazy module loading
// the first launch of a function can otherwise happen while another persistent kernel is resident and waiting for the host.
template <typename K>
static cudaError_t optin_smem(K kernel, size_t max_optin) {
    cudaFuncAttributes at;
    cudaError_t e = cudaFuncGetAttributes(&at, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(max_optin - at.sharedSizeBytes));
}
template <typename K>
static cudaError_t preload(K kernel) {
    cudaFuncAttributes at;
    return cudaFuncGetAttributes(&at, kernel);
}
static cudaError_t prepare_kernels(size_t max_optin) {
    cudaError_t e;
#define CG_PREP(call) if ((e = (call)) != cudaSuccess) return e
    CG_PREP(optin_smem(tower_ctail_kernel<true>, max_optin));
    CG_PREP(optin_smem(tower_ctail_kernel<false>, max_optin));
    CG_PREP(cudaFuncSetAttribute(tower_ctail_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));    // clusters of 16 CTAs
    CG_PREP(cudaFuncSetAttribute(tower_ctail_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CG_PREP(optin_smem(eq_small_kernel, max_optin));
    CG_PREP(optin_smem(veq_tma_kernel<false, false, 2>, max_optin));
    CG_PREP(optin_smem(veq_tma_kernel<true, true, 2>, max_optin));
    CG_PREP(optin_smem(veq_tma_kernel<true, false, 2>, max_optin));
    CG_PREP(optin_smem(veq_tma_kernel<false, false, 3>, max_optin));
    CG_PREP(optin_smem(veq_tma_kernel<true, true, 3>, max_optin));
    CG_PREP(optin_smem(veq_tma_kernel<true, false, 3>, max_optin));
    CG_PREP(optin_smem(veq_persist_kernel, max_optin));
    CG_PREP(optin_smem(veq_tma_kernel<true, true, 2, true>, max_optin));
    CG_PREP(optin_smem(veq_tma_kernel<true, false, 2, true>, max_optin));
    CG_PREP(preload(tower_mid_kernel<true>));
    CG_PREP(preload(tower_mid_kernel<false>));
    CG_PREP(preload(fold_kernel));
    CG_PREP(preload(veq_materialise_kernel));
    CG_PREP(preload(veq_tables_kernel));
#undef CG_PREP
    return cudaSuccess;
}
// ask for eager module loading when the CUDA runtime has not been initialised yet in this process (no effect otherwise)
__attribute__((constructor)) static void cg_module_loading_eager() { setenv("CUDA_MODULE_LOADING", "EAGER", 0); }

// Can a running kernel be answered by the host?  Under ncu / compute-sanitizer every launch blocks the calling thread until
// the kernel has finished, so a persistent kernel that waits for the host transcript (tail / mid mailbox protocol) would
// only ever see its own timeout.  cg_init finds out by experiment: a one-thread kernel waits (<= 50 ms) for a flag in mapped
// pinned memory that the host sets right after the launch call returns.  When the flag never arrives in time the context
// runs host-transcript sumchecks with one launch per round (CG_SC_NO_TAIL | CG_SC_NO_MID semantics); the device-challenger
// path keeps its persistent kernels (they never wait for the host).  CG_NO_HOST_WAIT=1 forces that mode, =0 skips the probe.
__global__ void host_wait_probe_kernel(const volatile unsigned* flag, unsigned* result, unsigned long long timeout_cycles) {
    const long long t0 = clock64();
    while (*flag == 0u) {
        if ((unsigned long long)(clock64() - t0) > timeout_cycles) { *result = 2u; return; }
    }
    *result = 1u;
}
static bool probe_host_wait(cudaStream_t st) {
    if (const char* e = getenv("CG_NO_HOST_WAIT")) return atoi(e) == 0;
    if (getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NV_SANITIZER_INJECTION_PORT_BASE")) return false;   // ncu / compute-sanitizer announce themselves
    unsigned* h = nullptr;
    if (cudaHostAlloc((void**)&h, 64, cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return true; }
    h[0] = 0; h[8] = 0;
    __sync_synchronize();
    host_wait_probe_kernel<<<1, 1, 0, st>>>(h, h + 8, 100000000ULL);   // ~50 ms
    ((volatile unsigned*)h)[0] = 1u;
    __sync_synchronize();
    const bool ok = cudaStreamSynchronize(st) == cudaSuccess && ((volatile unsigned*)h)[8] == 1u;
    cudaGetLastError();
    cudaFreeHost(h);
    return ok;
}

CG_EXPORT int cg_init(int device_id, cg_ctx** out) {
    if (!out) return CG_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device_id < 0 || device_id >= n) return CG_ERR_NO_DEVICE;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device_id) != cudaSuccess) return CG_ERR_NO_DEVICE;
    if (p.major != 10) return CG_ERR_NO_DEVICE;   // built for sm_100a only — no other code path exists
    cg_ctx* c = new cg_ctx();
    c->device = device_id;
    c->sm_count = p.multiProcessorCount;
    c->cc_major = p.major;
    c->cc_minor = p.minor;
    c->max_smem_optin = p.sharedMemPerBlockOptin;
    if (cudaSetDevice(device_id) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return CG_ERR_CUDA;
    }
    if (prepare_kernels(c->max_smem_optin) != cudaSuccess) {
        cudaGetLastError();
        cudaStreamDestroy(c->own_stream);
        delete c;
        return CG_ERR_CUDA;
    }
    c->host_wait_ok = probe_host_wait(c->own_stream);
    // largest thread-block cluster the tail kernel can run with at its full shared-memory footprint (16 is non-portable)
    for (uint32_t cs = CG_CT_MAX_C; cs >= 2; cs >>= 1) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(cs);
        cfg.blockDim = dim3(CG_CT_THREADS);
        cfg.dynamicSmemBytes = c->max_smem_optin - 8192;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, tower_ctail_kernel<false>, &cfg) == cudaSuccess && n_clusters >= 1) { c->tail_max_c = cs; break; }
        cudaGetLastError();
    }
    if (const char* e = getenv("CG_WAIT_TIMEOUT_MS")) {   // the device-side wait limit is configurable (default ~4 s)
        const double ms = atof(e);
        if (ms > 0) c->wait_timeout_cycles = (unsigned long long)(ms * 2.0e6);
    }
    // internal temporaries are stream-ordered (cudaMallocAsync); keep freed memory cached in the pool
    cudaMemPool_t mp;
    if (cudaDeviceGetDefaultMemPool(&mp, device_id) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = c;
    return CG_OK;
}
// stream-ordered temporaries: safe under the reference's one-stream-per-thread concurrency
static int tmp_alloc(cg_ctx* c, size_t bytes, void** p, cudaStream_t st) {
    if (bytes == 0) bytes = 256;
    cudaError_t e = cudaMallocAsync(p, bytes, st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(c, e == cudaErrorMemoryAllocation ? CG_ERR_OOM : CG_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    }
    return CG_OK;
}
static void tmp_free(void* p, cudaStream_t st) { if (p) cudaFreeAsync(p, st); }
static void* pinned_get(cg_ctx* c, size_t bytes) {
    {
        std::lock_guard<std::mutex> g(c->mu);
        for (size_t i = 0; i < c->pinned_cache.size(); i++)
            if (c->pinned_cache[i].first >= bytes) {
                void* p = c->pinned_cache[i].second;
                c->pinned_cache.erase(c->pinned_cache.begin() + i);
                return p;
            }
    }
    void* p = nullptr;
    if (bytes < 4096) bytes = 4096;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
static void pinned_put(cg_ctx* c, void* p, size_t bytes) {
    std::lock_guard<std::mutex> g(c->mu);
    c->pinned_cache.emplace_back(bytes < 4096 ? 4096 : bytes, p);
}
static void reap_pending_free(cg_ctx* c, bool wait);
CG_EXPORT int cg_pool_trim(cg_ctx* c) {
    if (!c) return CG_ERR_INVALID;
    std::lock_guard<std::mutex> g(c->mu);
    cudaSetDevice(c->device);
    reap_pending_free(c, true);
    for (auto& kv : c->free_blocks) { cudaFree(kv.second); c->reserved -= kv.first; }
    c->free_blocks.clear();
    return CG_OK;
}
CG_EXPORT int cg_destroy(cg_ctx* c) {
    if (!c) return CG_ERR_INVALID;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cg_pool_trim(c);
    for (auto& kv : c->live) cudaFree(kv.first);
    for (auto& pc : c->pinned_cache) cudaFreeHost(pc.second);
    if (c->d_p2) cudaFree(c->d_p2);
    if (c->d_ntt_tab) cudaFree(c->d_ntt_tab);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return CG_OK;
}
CG_EXPORT const char* cg_last_error(cg_ctx* c) {   // a per-thread copy: other lanes may set a new message concurrently
    static thread_local std::string tl;
    if (!c) return "null context";
    {
        std::lock_guard<std::mutex> g(c->mu);
        tl = c->err;
    }
    return tl.c_str();
}
CG_EXPORT uint64_t cg_launch_count(cg_ctx* c) { return c ? c->launches.load() : 0ULL; }
CG_EXPORT int cg_device_info(cg_ctx* c, int* sm, int* maj, int* min, size_t* fr, size_t* tot) {
    if (!c) return CG_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    size_t f = 0, t = 0;
    CU(c, cudaMemGetInfo(&f, &t));
    if (sm) *sm = c->sm_count;
    if (maj) *maj = c->cc_major;
    if (min) *min = c->cc_minor;
    if (fr) *fr = f;
    if (tot) *tot = t;
    return CG_OK;
}

// pooled allocator: exact-size free lists rounded to 2 MiB (sumcheck workspaces recur per layer)
static size_t round_size(size_t b) {
    const size_t g = b <= (1u << 16) ? 256 : (b <= (1u << 21) ? (1u << 16) : (1u << 21));
    return (b + g - 1) / g * g;
}
// blocks released by cg_free_async become reusable once their stream has passed the release point (caller holds c->mu)
static void reap_pending_free(cg_ctx* c, bool wait) {
    for (size_t i = 0; i < c->pending_free.size();) {
        cg_ctx::PendingFree& pf = c->pending_free[i];
        cudaError_t q = wait ? cudaEventSynchronize(pf.ev) : cudaEventQuery(pf.ev);
        if (q == cudaErrorNotReady) { i++; continue; }
        cudaGetLastError();
        cudaEventDestroy(pf.ev);
        c->free_blocks.emplace(pf.sz, pf.p);
        c->pending_free[i] = c->pending_free.back();
        c->pending_free.pop_back();
    }
}
CG_EXPORT int cg_alloc(cg_ctx* c, size_t bytes, void** dptr) {
    if (!c || !dptr) return CG_ERR_INVALID;
    if (bytes == 0) bytes = 256;
    const size_t sz = round_size(bytes);
    {
        std::lock_guard<std::mutex> g(c->mu);
        if (!c->pending_free.empty()) reap_pending_free(c, false);
        auto it = c->free_blocks.find(sz);
        if (it != c->free_blocks.end()) {
            *dptr = it->second;
            c->free_blocks.erase(it);
            c->live[*dptr] = sz;
            c->used += sz;
            return CG_OK;
        }
    }
    CU(c, cudaSetDevice(c->device));
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sz);
    if (e == cudaErrorMemoryAllocation) {   // retry once after releasing the cache
        cudaGetLastError();
        cg_pool_trim(c);
        e = cudaMalloc(&p, sz);
    }
    if (e != cudaSuccess) return set_err(c, e == cudaErrorMemoryAllocation ? CG_ERR_OOM : CG_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    std::lock_guard<std::mutex> g(c->mu);
    c->live[p] = sz;
    c->used += sz;
    c->reserved += sz;
    *dptr = p;
    return CG_OK;
}
CG_EXPORT int cg_free(cg_ctx* c, void* p) {
    if (!c) return CG_ERR_INVALID;
    if (!p) return CG_OK;
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->live.find(p);
    if (it == c->live.end()) { c->err = "cg_free: pointer not owned by this context"; return CG_ERR_INVALID; }
    c->free_blocks.emplace(it->second, p);
    c->used -= it->second;
    c->live.erase(it);
    return CG_OK;
}
// stream-ordered release: the block returns to the pool only after everything enqueued on `s` so far has finished, so a
// lane may release a buffer its own (non-blocking) stream is still using without synchronising first
CG_EXPORT int cg_free_async(cg_ctx* c, void* p, cg_stream s) {
    if (!c) return CG_ERR_INVALID;
    if (!p) return CG_OK;
    CU(c, cudaSetDevice(c->device));
    cudaEvent_t ev;
    CU(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (cudaEventRecord(ev, S(c, s)) != cudaSuccess) { cudaEventDestroy(ev); return set_err(c, CG_ERR_CUDA, "cg_free_async: cudaEventRecord failed"); }
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->live.find(p);
    if (it == c->live.end()) { cudaEventDestroy(ev); c->err = "cg_free_async: pointer not owned by this context"; return CG_ERR_INVALID; }
    c->pending_free.push_back(cg_ctx::PendingFree{p, it->second, ev});
    c->used -= it->second;
    c->live.erase(it);
    return CG_OK;
}
CG_EXPORT int cg_pool_stats(cg_ctx* c, size_t* used, size_t* reserved) {
    if (!c) return CG_ERR_INVALID;
    std::lock_guard<std::mutex> g(c->mu);
    if (used) *used = c->used;
    if (reserved) *reserved = c->reserved;
    return CG_OK;
}
CG_EXPORT int cg_h2d(cg_ctx* c, void* dst, const void* src, size_t bytes, cg_stream s) {
    if (!c) return CG_ERR_INVALID;
    CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, S(c, s)));
    return CG_OK;
}
CG_EXPORT int cg_d2h(cg_ctx* c, void* dst, const void* src, size_t bytes, cg_stream s) {
    if (!c) return CG_ERR_INVALID;
    CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, S(c, s)));
    return CG_OK;
}
CG_EXPORT int cg_d2d(cg_ctx* c, void* dst, const void* src, size_t bytes, cg_stream s) {
    if (!c) return CG_ERR_INVALID;
    CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, S(c, s)));
    return CG_OK;
}
CG_EXPORT int cg_stream_sync(cg_ctx* c, cg_stream s) {
    if (!c) return CG_ERR_INVALID;
    CU(c, cudaStreamSynchronize(S(c, s)));
    return CG_OK;
}
CG_EXPORT int cg_host_alloc_pinned(cg_ctx* c, size_t bytes, void** h) {
    if (!c || !h) return CG_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaHostAlloc(h, bytes, cudaHostAllocDefault));
    return CG_OK;
}
CG_EXPORT int cg_host_free_pinned(cg_ctx* c, void* h) {
    if (!c) return CG_ERR_INVALID;
    CU(c, cudaFreeHost(h));
    return CG_OK;
}

// ------------------------------------------------------------------------ stand-in transcript
CG_EXPORT void cg_standin_init(uint64_t* st, const uint8_t* label, uint64_t len) {
    uint64_t h = 0x43454E4F42323030ULL;
    cg_tr_absorb(h, len);
    for (uint64_t i = 0; i < len; i += 8) {
        uint64_t w = 0;
        for (uint64_t j = 0; j < 8 && i + j < len; j++) w |= (uint64_t)label[i + j] << (8 * j);
        cg_tr_absorb(h, w);
    }
    *st = h;
}
CG_EXPORT void cg_standin_append_message(uint64_t* st, const uint8_t* msg, uint64_t len) { cg_tr_append_message(*st, msg, len); }
CG_EXPORT void cg_standin_append_ext(uint64_t* st, const uint64_t* e, uint64_t n) {
    for (uint64_t i = 0; i < 2 * n; i++) cg_tr_absorb(*st, e[i]);
}
CG_EXPORT void cg_standin_sample(uint64_t* st, const char* label, uint64_t out[2]) {
    cg_tr_append_message(*st, (const uint8_t*)label, strlen(label));
    out[0] = cg_tr_squeeze(*st);
    out[1] = cg_tr_squeeze(*st);
}
CG_EXPORT void cg_standin_challenge_cb(void* user, uint32_t, const uint64_t* evals, uint32_t degree, uint64_t out[2]) {
    uint64_t* st = (uint64_t*)user;
    cg_standin_append_ext(st, evals, degree);
    cg_standin_sample(st, "Internal round", out);
}
static void standin_vt_sample(void* u, const char* l, uint64_t o[2]) { cg_standin_sample((uint64_t*)u, l, o); }
static void standin_vt_append(void* u, const uint64_t* e, uint64_t n) { cg_standin_append_ext((uint64_t*)u, e, n); }
static void standin_vt_begin(void* u, uint64_t nv, uint64_t dg) {
    cg_standin_append_message((uint64_t*)u, (const uint8_t*)&nv, 8);
    cg_standin_append_message((uint64_t*)u, (const uint8_t*)&dg, 8);
}
CG_EXPORT void cg_standin_vt(uint64_t* st, cg_transcript_vt* o) {
    o->user = st;
    o->sample = standin_vt_sample;
    o->append_exts = standin_vt_append;
    o->sumcheck_begin = standin_vt_begin;
    o->round_challenge = cg_standin_challenge_cb;
}

// ------------------------------------------------------------------------------------- eq build
static int upload_small(cg_ctx* c, const void* h, size_t bytes, void** d, cudaStream_t st) {
    CHK(tmp_alloc(c, bytes, d, st));
    CU(c, cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, st));   // pageable source: staged synchronously
    return CG_OK;
}
static int eq_small(cg_ctx* c, const ext_t* d_point, uint32_t k, ext_t* d_out, cudaStream_t st) {
    const size_t smem = sizeof(ext_t) << k;
    // (dynamic shared-memory opt-in raised once per context in prepare_kernels)
    unsigned threads = k >= 10 ? 1024 : (k >= 5 ? (1u << k) : 32);
    eq_small_kernel<<<1, threads, smem, st>>>(d_point, k, d_out);
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    return CG_OK;
}
// d_point: k ext on device.  Recursive two-level build.
static int build_eq_dev(cg_ctx* c, const ext_t* d_point, uint32_t k, ext_t* d_out, uint64_t start, uint64_t end, cudaStream_t st) {
    const uint64_t n = 1ULL << k;
    if (k <= CG_EQ_SMALL_K) {
        CHK(eq_small(c, d_point, k, d_out, st));
        if (start > 0 || end < n) {
            prefix_mask_kernel<<<grid_for(c, n), CG_THREADS, 0, st>>>(d_out, n, start, end);
            LAUNCHED(c);
            CU(c, cudaGetLastError());
        }
        return CG_OK;
    }
    const uint32_t lo_k = CG_EQ_SMALL_K, hi_k = k - lo_k;
    void *L = nullptr, *H = nullptr;
    CHK(tmp_alloc(c, sizeof(ext_t) << lo_k, &L, st));
    CHK(tmp_alloc(c, sizeof(ext_t) << hi_k, &H, st));
    int rc = eq_small(c, d_point, lo_k, (ext_t*)L, st);
    if (rc == CG_OK) rc = build_eq_dev(c, d_point + lo_k, hi_k, (ext_t*)H, 0, 1ULL << hi_k, st);
    if (rc == CG_OK) {
        eq_outer_kernel<<<grid_for(c, n / CG_EQ_PER_THREAD, 8), CG_THREADS, 0, st>>>((const ext_t*)L, (const ext_t*)H, lo_k, n, start, end, d_out);
        LAUNCHED(c);
        if (cudaGetLastError() != cudaSuccess) rc = set_err(c, CG_ERR_CUDA, "eq_outer_kernel launch failed");
    }
    tmp_free(L, st);   // stream-ordered: released after the kernels above
    tmp_free(H, st);
    return rc;
}
CG_EXPORT int cg_build_eq(cg_ctx* c, const uint64_t* h_point, uint32_t k, uint64_t* d_out, uint64_t offset,
                          uint64_t num_instances, cg_stream s) {
    if (!c || !d_out || (k && !h_point) || k > 40) return set_err(c, CG_ERR_INVALID, "cg_build_eq: bad argument");
    const uint64_t n = 1ULL << k;
    if (offset > n || num_instances > n - offset) return set_err(c, CG_ERR_INVALID, "cg_build_eq: offset + num_instances > 2^k");
    if ((uintptr_t)d_out & (k >= 1 ? 31 : 15)) return set_err(c, CG_ERR_INVALID, "cg_build_eq: d_out_ext must be 32-byte aligned (256-bit stores)");
    cudaStream_t st = S(c, s);
    CU(c, cudaSetDevice(c->device));
    void* d_point = nullptr;
    uint64_t dummy[2] = {0, 0};
    CHK(upload_small(c, k ? (const void*)h_point : (const void*)dummy, k ? sizeof(ext_t) * k : 16, &d_point, st));
    int rc = build_eq_dev(c, (const ext_t*)d_point, k, (ext_t*)d_out, offset, offset + num_instances, st);
    tmp_free(d_point, st);
    return rc;
}
CG_EXPORT int cg_selector_compute(cg_ctx* c, int kind, const uint64_t* h_point, uint32_t nv, uint64_t offset,
                                  uint64_t num_instances, const uint64_t* h_indices, uint32_t n_indices,
                                  uint32_t inner_vars, uint64_t* d_out, cg_stream s) {
    if (!c) return CG_ERR_INVALID;
    const uint64_t n = 1ULL << nv;
    cudaStream_t st = S(c, s);
    switch (kind) {
        case CG_SEL_WHOLE: return cg_build_eq(c, h_point, nv, d_out, 0, n, s);
        case CG_SEL_PREFIX:
            if (offset + num_instances > n) return set_err(c, CG_ERR_INVALID, "selector Prefix: end > 2^num_vars (selector.rs:144-150)");
            return cg_build_eq(c, h_point, nv, d_out, offset, num_instances, s);
        case CG_SEL_ORDERED_SPARSE: {
            if (inner_vars > nv || inner_vars > 20) return set_err(c, CG_ERR_INVALID, "selector OrderedSparse: bad inner_vars");
            CHK(cg_build_eq(c, h_point, nv, d_out, 0, n, s));
            std::vector<uint8_t> keep(1ULL << inner_vars, 0);
            // the reference walks `indices` in order and keeps i only when it equals the NEXT
            // expected index (selector.rs:176-186): out-of-order entries stall the iterator.
            uint32_t it = 0;
            for (uint64_t i = 0; i < keep.size(); i++)
                if (it < n_indices && h_indices[it] == i) { keep[i] = 1; it++; }
            void* d_keep = nullptr;
            CHK(upload_small(c, keep.data(), keep.size(), &d_keep, st));
            sparse_mask_kernel<<<grid_for(c, n), CG_THREADS, 0, st>>>((ext_t*)d_out, n, inner_vars, num_instances, (const uint8_t*)d_keep);
            LAUNCHED(c);
            tmp_free(d_keep, st);
            CU(c, cudaGetLastError());
            return CG_OK;
        }
        case CG_SEL_QUARK_LT: {
            if (offset != 0) return set_err(c, CG_ERR_INVALID, "selector QuarkBinaryTreeLessThan: offset must be 0 (selector.rs:192)");
            if (nv == 0 || nv > 63) return set_err(c, CG_ERR_INVALID, "selector QuarkBinaryTreeLessThan: num_vars out of range");
            CHK(cg_build_eq(c, h_point, nv, d_out, 0, n, s));
            QuarkArgs q;
            memset(&q, 0, sizeof(q));
            q.num_vars = nv;
            uint64_t ni = num_instances;
            for (uint32_t i = 0; i < nv; i++) { q.seq[i] = ni / 2; ni = (ni + 1) / 2; }
            quark_mask_kernel<<<grid_for(c, n), CG_THREADS, 0, st>>>((ext_t*)d_out, n, q);
            LAUNCHED(c);
            CU(c, cudaGetLastError());
            return CG_OK;
        }
        default: return set_err(c, CG_ERR_INVALID, "cg_selector_compute: unknown kind");
    }
}

// EC-sum Quark pre-passes (f-3): the three selector MLEs of CpuEccProver::create_ecc_proof and the even/odd split
CG_EXPORT int cg_ecc_quark_selectors(cg_ctx* c, const uint64_t* h_out_rt, uint32_t n_vars, uint64_t num_instances, uint64_t* d_sel_add,
                                     uint64_t* d_sel_bypass, uint64_t* d_sel_export, cg_stream s) {
    if (!c || !d_sel_add || !d_sel_bypass || !d_sel_export) return CG_ERR_INVALID;
    if (n_vars == 0 || n_vars > 40) return set_err(c, CG_ERR_INVALID, "cg_ecc_quark_selectors: num_vars out of range");
    const uint64_t n = 1ULL << n_vars;
    if (num_instances > n) return set_err(c, CG_ERR_INVALID, "cg_ecc_quark_selectors: num_instances > 2^num_vars");
    if (((uintptr_t)d_sel_add | (uintptr_t)d_sel_bypass | (uintptr_t)d_sel_export) & 31) return set_err(c, CG_ERR_INVALID, "cg_ecc_quark_selectors: outputs must be 32-byte aligned");
    cudaStream_t st = S(c, s);
    CHK(cg_selector_compute(c, CG_SEL_QUARK_LT, h_out_rt, n_vars, 0, num_instances, nullptr, 0, 0, d_sel_add, s));
    CHK(cg_build_eq(c, h_out_rt, n_vars, d_sel_bypass, 0, n, s));
    ecc_selectors_kernel<<<grid_for(c, n), CG_THREADS, 0, st>>>((const ext_t*)d_sel_add, (ext_t*)d_sel_bypass, (ext_t*)d_sel_export, n);
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    return CG_OK;
}
CG_EXPORT int cg_split_even_odd(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, uint64_t* const* d_even, uint64_t* const* d_odd, cg_stream s) {
    if (!c || (n_mles && (!mles || !d_even || !d_odd))) return CG_ERR_INVALID;
    if (n_mles == 0) return CG_OK;
    const uint32_t nv = mles[0].num_vars;
    if (nv == 0) return set_err(c, CG_ERR_INVALID, "cg_split_even_odd: MLE has no variables");
    std::vector<const void*> ptrs(3 * (size_t)n_mles);
    for (uint32_t i = 0; i < n_mles; i++) {
        if (mles[i].is_ext != CG_MLE_BASE) return set_err(c, CG_ERR_INVALID, "cg_split_even_odd: base-field MLEs only (get_base_field_vec, cpu/mod.rs:141)");
        if (mles[i].num_vars != nv || mles[i].len != (1ULL << nv)) return set_err(c, CG_ERR_INVALID, "cg_split_even_odd: all MLEs must be full and of one size");
        if ((uintptr_t)mles[i].dptr & 15) return set_err(c, CG_ERR_INVALID, "cg_split_even_odd: input must be 16-byte aligned");
        ptrs[i] = mles[i].dptr;
        ptrs[n_mles + i] = d_even[i];
        ptrs[2 * (size_t)n_mles + i] = d_odd[i];
    }
    cudaStream_t st = S(c, s);
    CU(c, cudaSetDevice(c->device));
    void* d_ptrs = nullptr;
    CHK(upload_small(c, ptrs.data(), sizeof(void*) * ptrs.size(), &d_ptrs, st));
    SplitArgs a;
    a.in = (const uint64_t* const*)d_ptrs;
    a.even = (uint64_t* const*)d_ptrs + n_mles;
    a.odd = (uint64_t* const*)d_ptrs + 2 * (size_t)n_mles;
    a.n_out = 1ULL << (nv - 1);
    split_even_odd_kernel<<<dim3(grid_for(c, a.n_out), n_mles), 256, 0, st>>>(a);
    LAUNCHED(c);
    tmp_free(d_ptrs, st);
    CU(c, cudaGetLastError());
    return CG_OK;
}


// ------------------------------------------------------------------------------------- comm
// Multi-GPU mailbox (the reference has no multi-GPU path — docs/src/optimizations.md:3-5
// lists distributed sumcheck as TODO).  One cudaMalloc'd CommBuf per rank, exported through CUDA IPC;
// the host framework (torch.distributed, MPI, ...) only moves the 64-byte handles once at start-up.
struct cg_comm {
    cg_ctx* ctx = nullptr;
    int rank = 0, nranks = 1;
    CommBuf* mine = nullptr;
    CommBuf* peers[CG_MAX_RANKS] = {nullptr};
    uint64_t seq = 1;        // next exchange sequence number (identical on every rank)
    uint64_t gather_calls = 0;
    uint64_t bar_seq = 0;    // rank barriers so far (identical on every rank)
    // peer-mapped arena (sharded tower layers are written by the partner ranks' layer kernels over NVLink): every rank allocates
    // the same size and carves it with the same sequence of requests, so an offset means the same buffer on every rank
    char* arena[CG_MAX_RANKS] = {nullptr};
    size_t arena_bytes = 0, arena_off = 0;
    int* d_error = nullptr;
    unsigned long long* d_dbg = nullptr;
};
CG_EXPORT int cg_comm_create(cg_ctx* c, int rank, int nranks, cg_comm** out, uint8_t handle_out[64]) {
    if (!c || !out || !handle_out || nranks < 1 || nranks > CG_MAX_RANKS || rank < 0 || rank >= nranks || (nranks & (nranks - 1)))
        return set_err(c, CG_ERR_INVALID, "cg_comm_create: nranks must be a power of two <= 8 and 0 <= rank < nranks");
    CU(c, cudaSetDevice(c->device));
    cg_comm* cm = new cg_comm();
    cm->ctx = c;
    cm->rank = rank;
    cm->nranks = nranks;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    if (cudaMalloc(&p, sizeof(CommBuf)) != cudaSuccess || cudaMemset(p, 0, sizeof(CommBuf)) != cudaSuccess ||
        cudaMalloc((void**)&cm->d_error, 256) != cudaSuccess || cudaMemset(cm->d_error, 0, 256) != cudaSuccess) {
        delete cm;
        return set_err(c, CG_ERR_CUDA, "cg_comm_create: cudaMalloc failed");
    }
    cm->mine = (CommBuf*)p;
    cm->peers[rank] = cm->mine;
    if (getenv("CG_COMM_DEBUG")) { cudaMalloc((void**)&cm->d_dbg, 1024 * 4 * 8); cudaMemset(cm->d_dbg, 0, 1024 * 4 * 8); }
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (nranks > 1 && cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
        delete cm;
        return set_err(c, CG_ERR_CUDA, "cg_comm_create: cudaIpcGetMemHandle failed");
    }
    memcpy(handle_out, &h, 64);
    CU(c, cudaDeviceSynchronize());
    *out = cm;
    return CG_OK;
}
CG_EXPORT int cg_comm_connect(cg_comm* cm, const uint8_t* all_handles) {
    if (!cm || !all_handles) return CG_ERR_INVALID;
    cg_ctx* c = cm->ctx;
    CU(c, cudaSetDevice(c->device));
    for (int p = 0; p < cm->nranks; p++) {
        if (p == cm->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + 64 * (size_t)p, 64);
        void* ptr = nullptr;
        CU(c, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        cm->peers[p] = (CommBuf*)ptr;
    }
    return CG_OK;
}
CG_EXPORT int cg_comm_arena_create(cg_comm* cm, size_t bytes, uint8_t handle_out[64]) {
    if (!cm || !handle_out || bytes == 0) return CG_ERR_INVALID;
    cg_ctx* c = cm->ctx;
    if (cm->arena[cm->rank]) return set_err(c, CG_ERR_STATE, "cg_comm_arena_create: the arena exists already");
    CU(c, cudaSetDevice(c->device));
    void* p = nullptr;
    CU(c, cudaMalloc(&p, bytes));
    cm->arena[cm->rank] = (char*)p;
    cm->arena_bytes = bytes;
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (cm->nranks > 1) CU(c, cudaIpcGetMemHandle(&h, p));
    memcpy(handle_out, &h, 64);
    return CG_OK;
}
CG_EXPORT int cg_comm_arena_connect(cg_comm* cm, const uint8_t* all_handles) {
    if (!cm || !all_handles || !cm->arena[cm->rank]) return CG_ERR_INVALID;
    cg_ctx* c = cm->ctx;
    CU(c, cudaSetDevice(c->device));
    for (int p = 0; p < cm->nranks; p++) {
        if (p == cm->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + 64 * (size_t)p, 64);
        void* ptr = nullptr;
        CU(c, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        cm->arena[p] = (char*)ptr;
    }
    return CG_OK;
}
// bump allocation: the same call sequence on every rank yields the same offsets
static int arena_alloc(cg_comm* cm, size_t bytes, size_t* off) {
    const size_t a = (cm->arena_off + 255) & ~(size_t)255;
    if (!cm->arena[cm->rank] || a + bytes > cm->arena_bytes) return set_err(cm->ctx, CG_ERR_OOM, "sharded tower: the peer arena is too small (cg_comm_arena_create)");
    *off = a;
    cm->arena_off = a + bytes;
    return CG_OK;
}
static int comm_barrier(cg_comm* cm, cudaStream_t st);
CG_EXPORT int cg_comm_destroy(cg_comm* cm) {
    if (!cm) return CG_ERR_INVALID;
    cudaSetDevice(cm->ctx->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < cm->nranks; p++)
        if (cm->arena[p]) { if (p == cm->rank) cudaFree(cm->arena[p]); else cudaIpcCloseMemHandle(cm->arena[p]); }
    if (cm->d_dbg) {   // CG_COMM_DEBUG: dump the last exchanges' wait times
        std::vector<unsigned long long> h(1024 * 4);
        cudaMemcpy(h.data(), cm->d_dbg, h.size() * 8, cudaMemcpyDeviceToHost);
        const uint64_t last = cm->seq - 1;
        for (uint64_t q = last > 40 ? last - 40 : 1; q <= last; q++)
            fprintf(stderr, "[comm rank %d] seq %llu wait_self %llu ns wait_peer %llu ns polls %llu start %llu\n", cm->rank, (unsigned long long)q,
                    h[(q % 1024) * 4], h[(q % 1024) * 4 + 1], h[(q % 1024) * 4 + 2], h[(q % 1024) * 4 + 3]);
        cudaFree(cm->d_dbg);
    }
    for (int p = 0; p < cm->nranks; p++)
        if (p != cm->rank && cm->peers[p]) cudaIpcCloseMemHandle(cm->peers[p]);
    cudaFree(cm->mine);
    cudaFree(cm->d_error);
    delete cm;
    return CG_OK;
}
static void comm_dev(cg_comm* cm, CommDev& d, uint64_t n_exchanges) {
    memset(&d, 0, sizeof(d));
    if (!cm || cm->nranks <= 1) return;
    d.rank = cm->rank;
    d.nranks = cm->nranks;
    d.seq = cm->seq;
    cm->seq += n_exchanges;
    for (int p = 0; p < cm->nranks; p++) d.peers[p] = cm->peers[p];
    d.d_error = cm->d_error;
    d.timeout_cycles = cm->ctx->wait_timeout_cycles;
    d.dbg = cm->d_dbg;
}
static int comm_barrier(cg_comm* cm, cudaStream_t st) {
    if (!cm || cm->nranks <= 1) return CG_OK;
    CommDev d;
    comm_dev(cm, d, 0);
    comm_barrier_kernel<<<1, 32, 0, st>>>(d, ++cm->bar_seq);
    cm->ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) return set_err(cm->ctx, CG_ERR_CUDA, "comm_barrier_kernel launch failed");
    return CG_OK;
}
static int comm_check(cg_comm* cm, cudaStream_t st) {
    if (!cm || cm->nranks <= 1) return CG_OK;
    int e = 0;
    if (cudaMemcpyAsync(&e, cm->d_error, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
        return set_err(cm->ctx, CG_ERR_CUDA, "comm error flag read failed");
    if (e) return set_err(cm->ctx, CG_ERR_CUDA, "multi-GPU exchange timed out waiting for a peer rank");
    return CG_OK;
}

// host-side field helpers (transcript-side scalar work: a handful of operations per sumcheck, next to the callbacks)
static inline uint64_t hx_mulmod(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) % GL_P); }
static inline uint64_t hx_addmod(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a + b) % GL_P); }
static inline uint64_t hx_submod(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)(a % GL_P) + GL_P - (b % GL_P)) % GL_P); }
static inline ext_t hx_mul(ext_t a, ext_t b) {
    return ext_t{hx_addmod(hx_mulmod(a.c0, b.c0), hx_mulmod(7, hx_mulmod(a.c1, b.c1))), hx_addmod(hx_mulmod(a.c0, b.c1), hx_mulmod(a.c1, b.c0))};
}
static inline ext_t hx_add(ext_t a, ext_t b) { return ext_t{hx_addmod(a.c0, b.c0), hx_addmod(a.c1, b.c1)}; }
static inline uint64_t hx_powmod(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    for (; e; e >>= 1, a = hx_mulmod(a, a)) if (e & 1) r = hx_mulmod(r, a);
    return r;
}
// (a0 + a1 X)^-1 = (a0 - a1 X) / (a0^2 - 7 a1^2)   (X^2 = 7; the norm is non-zero for a != 0 because 7 is a non-residue)
static inline ext_t hx_inv(ext_t a) {
    const uint64_t norm = hx_submod(hx_mulmod(a.c0, a.c0), hx_mulmod(7, hx_mulmod(a.c1, a.c1)));
    const uint64_t ni = hx_powmod(norm, GL_P - 2);
    return ext_t{hx_mulmod(a.c0 % GL_P, ni), hx_mulmod(hx_submod(0, a.c1), ni)};
}

// --------------------------------------------------------------------------------- sumcheck
struct MleState {
    const void* orig = nullptr;
    uint32_t orig_is_ext = 0;
    void* padded = nullptr;   // pooled zero-padded copy when len < 2^num_vars
};
struct TowerLayout {           // MLE indices of the specialised (tower-shaped) kernel
    bool on = false;
    uint32_t eq = 0;
    std::vector<uint32_t> prod;   // 2 per spec
    std::vector<ext_t> prod_alpha;
    std::vector<uint32_t> lk;     // 4 per spec
    std::vector<ext_t> lk_an, lk_ad;
    bool alpha_one = false;
};
#define CG_SC_INT_DEFER_VEQ (1u << 30)   // internal: the caller decides about split-eq rounds after create (tower layers)
struct VeqState {               // a virtual eq MLE (CG_MLE_EQ) handled by the split-eq kernels
    bool split = false;         // split rounds active: the eq state is not materialised yet
    bool have = false;          // a virtual eq was declared (point recorded)
    uint32_t idx = 0;           // its MLE index
    uint32_t J = 0;             // rounds 0 .. J-1 are split rounds
    std::vector<uint64_t> h_point, h_up;
    ext_t* d_w = nullptr;       // the point on the device
    ext_t* d_prefix = nullptr;  // P_folds = prod_{i < folds} eq(w_i, r_i)   (the rank factor `scale` is kept separate)
    ext_t* d_inv1mw = nullptr;  // 1 / (1 - w_j), host-computed
    ext_t* d_qstate = nullptr;  // coefficients of the previous round's q(X): the running claim of the claim-derived rounds
    bool derive_ok = false;     // every 1 - w_j is invertible (else the rounds accumulate all three sums)
    bool have_claim = false;    // the caller knows the claimed sum (tower layers): round 0 is claim-derived as well
    ext_t claim0{0, 0};
    ext_t scale{1, 0};          // sharded prove: eq(w_top, rank) — the constant factor of eq on this rank's slice
    ulonglong4 *d_L = nullptr, *d_H = nullptr;
    uint64_t h_off[CG_VEQ_MAX_ROUNDS + 1] = {0};
    // general tower layouts (tveq_round_kernel): S alpha-folded copies of every H entry; launches that read virtual leaves
    // (rounds 0 / 1) use the lanes-over-high mapping with their own low / high tables
    bool general = false;
    uint32_t S = 0;
    ulonglong4* d_UA = nullptr;
    bool m1 = false;
    uint32_t m1_lo[2] = {0, 0};
    ulonglong4 *d_LR[2] = {nullptr, nullptr}, *d_HS[2] = {nullptr, nullptr};
};
struct cg_sumcheck {
    cg_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr;
    uint32_t n_mles = 0, num_vars = 0, degree = 0, flags = 0, n_terms = 0;
    uint32_t round = 0;        // binds so far
    uint32_t folds = 0;        // folds applied so far
    bool evaluated = false;    // round_eval done for the current round
    bool pending = false;
    ext_t pending_r;
    const ext_t* pending_r_ptr = nullptr;   // device challenger: challenge lives on the device
    std::vector<MleState> mles;
    std::vector<const VirtLeaf*> virt;   // per MLE: its ORIGINAL input is a virtual tower leaf array (null / empty: a plain array)
    uint32_t virt_l2m = 0;               // log2 of the padded record count of the virtual leaves (index layout of their rows)
    struct VirtHost { uint32_t n_records = 0, l2m = 0; ext_t def{0, 0}; };
    std::vector<VirtHost> virt_h;        // host copy of what the split-eq launches need to know about each virtual slot
    TowerLayout tl;
    VeqState veq;
    // device state
    void* ws = nullptr;        // workspace: per MLE [n/2 | n/4] ext
    ext_t* d_final = nullptr;  // n_mles ext
    ext_t* d_coeff = nullptr;
    uint32_t *d_off = nullptr, *d_idx = nullptr;
    MleSlot* d_slots = nullptr;   // [(num_vars+1) * n_mles]  state after f folds
    FoldSlot* d_fold = nullptr;   // [num_vars * n_mles]      fold f -> f+1
    RoundOut out{};
    ext_t* d_msgs = nullptr;      // num_vars * degree (device challenger) or degree
    uint64_t* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
    // device challenger
    uint64_t* d_tr_state = nullptr;
    ext_t* d_chal = nullptr;
    std::vector<void*> owned;
    int* d_error = nullptr;
    unsigned *d_mid_ticket = nullptr, *d_mid_flag = nullptr;
    unsigned *d_per_ticket = nullptr, *d_per_flag = nullptr;   // persistent split-eq rounds
    bool mid_used = false;
    std::vector<uint64_t> h_coeff;   // term tables stay on the host until a generic kernel needs them
    std::vector<uint32_t> h_off, h_idx;
    bool tables_ready = false;
    // grouped plan (terms grouped by their round-0 ext factors), host side and device side
    bool plan_on = false;
    std::vector<uint32_t> p_g_term_off, p_g_ext_off, p_g_ext_idx, p_t_off, p_t_idx;
    std::vector<uint32_t> pf_g_term_off, pf_g_ext_off, pf_g_ext_idx;   // fine plan: groups split into 8-term chunks
    uint32_t *d_fg_term_off = nullptr, *d_fg_ext_off = nullptr, *d_fg_ext_idx = nullptr;
    std::vector<uint64_t> p_t_coeff;
    uint32_t *d_g_term_off = nullptr, *d_g_ext_off = nullptr, *d_g_ext_idx = nullptr, *d_pt_off = nullptr, *d_pt_idx = nullptr;
    uint64_t* d_pt_coeff = nullptr;
    uint32_t extra_rounds = 0;     // sharded prove: replicated rounds the tail kernel runs after the all-gather
    bool extra_done = false;
    bool profile_append = false;   // sharded prove: the replicated tail appends to the local rounds' profile
    cg_comm* comm = nullptr;       // multi-GPU: partial sums are combined in-kernel over NVLink
    std::vector<cudaEvent_t> ev;   // CG_SC_PROFILE: 2 events per round on the launching stream
};

// ext elements of workspace per MLE: [n/2 | n/4], kept even so every buffer stays 32-byte aligned
static uint64_t ws_per(uint64_t n) { uint64_t p = n / 2 + n / 4; p = (p + 1) & ~1ULL; return p < 2 ? 2 : p; }
// state of MLE i after f folds
static const void* mle_buf(const cg_sumcheck* sc, uint32_t i, uint32_t f) {
    if (f == 0) return sc->mles[i].padded ? sc->mles[i].padded : sc->mles[i].orig;
    const uint64_t n = 1ULL << sc->num_vars;
    const uint64_t per = ws_per(n);
    ext_t* base = (ext_t*)sc->ws + (uint64_t)i * per;
    return (f & 1) ? (const void*)base : (const void*)(base + n / 2);
}
static uint32_t mle_is_ext(const cg_sumcheck* sc, uint32_t i, uint32_t f) { return f == 0 ? sc->mles[i].orig_is_ext : 1u; }

static int sc_alloc(cg_sumcheck* sc, size_t bytes, void** p) {
    CHK(tmp_alloc(sc->ctx, bytes, p, sc->stream));
    sc->owned.push_back(*p);
    return CG_OK;
}

CG_EXPORT int cg_sumcheck_destroy(cg_sumcheck* sc) {
    if (!sc) return CG_ERR_INVALID;
    for (void* p : sc->owned) tmp_free(p, sc->stream);   // stream-ordered: no synchronisation needed
    if (sc->h_pinned) pinned_put(sc->ctx, sc->h_pinned, sc->h_pinned_bytes);
    sc->ctx->live_sc--;
    delete sc;
    return CG_OK;
}

// virtual eq -> ordinary table (shapes / flags the split-eq kernels do not cover)
static int veq_build_full(cg_sumcheck* sc, uint32_t i, const uint64_t* h_point) {
    void* buf = nullptr;
    CHK(sc_alloc(sc, sizeof(ext_t) << sc->num_vars, &buf));
    CHK(cg_build_eq(sc->ctx, h_point, sc->num_vars, (uint64_t*)buf, 0, 1ULL << sc->num_vars, (cg_stream)sc->stream));
    if (i == sc->veq.idx && sc->veq.have && !(sc->veq.scale.c0 == 1 && sc->veq.scale.c1 == 0)) {
        const uint64_t n = 1ULL << sc->num_vars;
        scale_ext_kernel<<<grid_for(sc->ctx, n, 8), CG_THREADS, 0, sc->stream>>>((ext_t*)buf, n, sc->veq.scale);
        LAUNCHED(sc->ctx);
        CU(sc->ctx, cudaGetLastError());
    }
    sc->mles[i].orig = buf;
    sc->mles[i].orig_is_ext = 1;
    return CG_OK;
}
static uint32_t log2_u64(uint64_t x);
static uint64_t tail_cap_n0(const cg_ctx* c, size_t n_slots, bool sharded);
// split rounds run until the cluster tail can take the (materialised) state over: it enters at round J with a pending
// fold, i.e. with 2^(k - J) elements per MLE (cluster-wide; sharded: times the rank count).  0: too small for split rounds.
static uint32_t veq_split_rounds(const cg_ctx* c, uint32_t k, size_t n_slots, bool sharded, uint32_t extra_rounds) {
    if (k < 2 + CG_VEQ_LO_BITS) return 0;
    const uint32_t cap_log = log2_u64(std::max<uint64_t>(2, tail_cap_n0(c, n_slots, sharded)));
    const uint32_t loc_log = cap_log > extra_rounds ? cap_log - extra_rounds : 1;
    uint32_t J = k > loc_log ? k - loc_log : 1;
    if (J > k - 1 - CG_VEQ_LO_BITS) J = k - 1 - CG_VEQ_LO_BITS;   // a round needs at least one row of 256 pairs
    if (J > CG_VEQ_MAX_ROUNDS) J = CG_VEQ_MAX_ROUNDS;
    return J;
}
// split mode: upload the point, build every round's L/H tables in one launch
static int veq_setup_split(cg_sumcheck* sc) {
    cg_ctx* c = sc->ctx;
    VeqState& v = sc->veq;
    const uint32_t k = sc->num_vars;
    const bool sharded = sc->comm && sc->comm->nranks > 1;
    const TowerLayout& tl = sc->tl;
    bool any_virt = false;
    for (const VirtLeaf* q : sc->virt) any_virt |= q != nullptr;
    v.general = !(tl.alpha_one && tl.prod.size() == 2 && tl.lk.empty() && !any_virt);
    v.J = veq_split_rounds(c, k, v.general ? sc->n_mles : 3, sharded, sc->extra_rounds);
    if (v.J == 0 || (any_virt && (v.J < 2 || sc->virt_l2m < 3))) return set_err(c, CG_ERR_STATE, "split-eq rounds: shape not eligible");
    v.h_off[0] = 0;
    for (uint32_t j = 0; j < v.J; j++) v.h_off[j + 1] = v.h_off[j] + (1ULL << (k - j - 1 - CG_VEQ_LO_BITS));
    // one upload: [w (k ext) | 1 / (1 - w_j) (k ext) | prefix = 1 | q-state (3 ext)]
    void* p = nullptr;
    CHK(sc_alloc(sc, sizeof(ext_t) * (2 * (size_t)k + 4) + 256, &p));
    v.d_w = (ext_t*)p;
    v.d_inv1mw = v.d_w + k;
    v.d_prefix = v.d_w + 2 * (size_t)k;
    v.d_qstate = v.d_prefix + 1;
    v.h_up.assign(2 * (2 * (size_t)k + 4), 0);
    memcpy(v.h_up.data(), v.h_point.data(), sizeof(ext_t) * k);
    v.derive_ok = !(sc->flags & CG_SC_NO_DERIVE);
    for (uint32_t j = 0; j < k; j++) {
        const ext_t om{hx_submod(1, v.h_point[2 * j]), hx_submod(0, v.h_point[2 * j + 1])};
        if (om.c0 == 0 && om.c1 == 0) { if (j >= 1 && j < v.J) v.derive_ok = false; if (j == 0) v.have_claim = false; continue; }
        const ext_t iv = hx_inv(om);
        v.h_up[2 * (k + j)] = iv.c0;
        v.h_up[2 * (k + j) + 1] = iv.c1;
    }
    v.h_up[2 * (2 * (size_t)k)] = 1;   // prefix = 1
    if (v.have_claim) { v.h_up[2 * (2 * (size_t)k + 1)] = v.claim0.c0; v.h_up[2 * (2 * (size_t)k + 1) + 1] = v.claim0.c1; }   // q-state: q_{-1}(0) = the claimed sum
    CU(c, cudaMemcpyAsync(v.d_w, v.h_up.data(), sizeof(uint64_t) * v.h_up.size(), cudaMemcpyHostToDevice, sc->stream));   // v lives as long as sc
    CHK(sc_alloc(sc, sizeof(ulonglong4) * ((size_t)v.J << CG_VEQ_LO_BITS), &p));
    v.d_L = (ulonglong4*)p;
    CHK(sc_alloc(sc, sizeof(ulonglong4) * v.h_off[v.J], &p));
    v.d_H = (ulonglong4*)p;
    VeqTabArgs a;
    memset(&a, 0, sizeof(a));
    a.w = v.d_w; a.k = k; a.J = v.J; a.L = v.d_L; a.H = v.d_H;
    for (uint32_t j = 0; j <= v.J; j++) a.h_off[j] = v.h_off[j];
    if (v.general) {   // weight slots: one per product spec (alpha), two per logup spec (numerator / denominator alpha)
        v.S = (uint32_t)(tl.prod_alpha.size() + 2 * tl.lk_an.size());
        if (v.S == 0 || v.S > CG_TVEQ_MAX_SLOTS) return set_err(c, CG_ERR_STATE, "split-eq rounds: spec count out of range");
        for (size_t i = 0; i < tl.prod_alpha.size(); i++) a.alpha[i] = tl.prod_alpha[i];
        for (size_t l = 0; l < tl.lk_an.size(); l++) {
            a.alpha[tl.prod_alpha.size() + 2 * l] = tl.lk_an[l];
            a.alpha[tl.prod_alpha.size() + 2 * l + 1] = tl.lk_ad[l];
        }
        CHK(sc_alloc(sc, sizeof(ulonglong4) * v.h_off[v.J] * v.S, &p));
        v.d_UA = (ulonglong4*)p;
        a.UA = v.d_UA;
        a.S = v.S;
    }
    veq_tables_kernel<<<grid_for(c, ((uint64_t)v.J << CG_VEQ_LO_BITS) + v.h_off[v.J], 8), CG_THREADS, 0, sc->stream>>>(a);
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    if (any_virt) {   // rounds 0 and 1 read the leaves: lanes over the record rows, loop over the record index
        v.m1 = true;
        for (uint32_t j = 0; j < 2; j++) {
            v.m1_lo[j] = sc->virt_l2m - 1 - j;   // item bits below the row index
            if (v.m1_lo[j] > k - j - 1) v.m1_lo[j] = k - j - 1;
            VeqRangeArgs ra;
            memset(&ra, 0, sizeof(ra));
            ra.w = v.d_w;
            ra.v0 = j + 1; ra.nb = v.m1_lo[j]; ra.S = v.S;
            memcpy(ra.alpha, a.alpha, sizeof(ra.alpha));
            CHK(sc_alloc(sc, sizeof(ulonglong4) * v.S << ra.nb, &p));
            v.d_LR[j] = ra.out = (ulonglong4*)p;
            veq_range_table_kernel<<<grid_for(c, 1ULL << ra.nb, 8), CG_THREADS, 0, sc->stream>>>(ra);
            LAUNCHED(c);
            ra.v0 = j + 1 + v.m1_lo[j]; ra.nb = k - ra.v0; ra.S = 0;
            CHK(sc_alloc(sc, sizeof(ulonglong4) << ra.nb, &p));
            v.d_HS[j] = ra.out = (ulonglong4*)p;
            veq_range_table_kernel<<<grid_for(c, 1ULL << ra.nb, 8), CG_THREADS, 0, sc->stream>>>(ra);
            LAUNCHED(c);
        }
        CU(c, cudaGetLastError());
    }
    v.split = true;
    return CG_OK;
}
static const void* mle_buf(const cg_sumcheck* sc, uint32_t i, uint32_t f);
// leave split mode: write the eq state of fold level sc->folds where a materialised table would be now
static int veq_leave_split(cg_sumcheck* sc) {
    cg_ctx* c = sc->ctx;
    VeqState& v = sc->veq;
    if (!v.split) return CG_OK;
    const uint32_t f = sc->folds;
    const uint64_t n = 1ULL << (sc->num_vars - f);
    if (f == 0) {   // nothing bound yet: the plain table
        CHK(veq_build_full(sc, v.idx, v.h_point.data()));
    } else {
        if (f > v.J) return set_err(c, CG_ERR_STATE, "split-eq: no tables for this fold level");
        const uint32_t j = f - 1;   // tables of round j cover variables [j+1, k) = [f, k)
        veq_materialise_kernel<<<grid_for(c, n / 2, 8), CG_THREADS, 0, sc->stream>>>(v.d_L + ((size_t)j << CG_VEQ_LO_BITS), v.d_H + v.h_off[j],
                                                                                  v.d_prefix, v.scale, n, (ext_t*)mle_buf(sc, v.idx, f));
        LAUNCHED(c);
        CU(c, cudaGetLastError());
    }
    v.split = false;
    return CG_OK;
}
static int veq_prepare(cg_sumcheck* sc) {   // before evaluating round sc->round
    if (sc->veq.split && sc->round >= sc->veq.J) return veq_leave_split(sc);
    return CG_OK;
}

// A caller that runs many sumchecks back to back on one stream (the tower prover: one per layer, sizes doubling) lends ONE
// workspace instead of a fresh pooled block per sumcheck: the stream-ordered pool fragments under a doubling size sequence and
// re-maps gigabytes per allocation (measured on the keccak tower: 107 ms of allocation against 129 ms of kernels).
struct ScWorkspace {
    void* ptr = nullptr;
    size_t bytes = 0;
};
static thread_local const ScWorkspace* tl_lent_ws = nullptr;
static int sc_create_common(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, uint32_t num_vars, uint32_t degree,
                            uint32_t flags, cudaStream_t st, cg_sumcheck** out) {
    if (!c || !out || (n_mles && !mles)) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_create: null argument");
    if (degree == 0 || degree > CG_MAX_DEGREE) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_create: degree must be in 1..8");
    if (num_vars > 34) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_create: num_vars too large");
    for (uint32_t i = 0; i < n_mles; i++) {
        if (mles[i].num_vars != num_vars)
            return set_err(c, CG_ERR_UNSUPPORTED,
                           "mixed num_vars (the frontloaded batched main sumcheck) is served by cg_sumcheck_prove, not by the step API");
        if (mles[i].is_ext == CG_MLE_EQ) {   // virtual eq: dptr is the HOST point
            if (!mles[i].dptr && num_vars) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_create: virtual eq MLE without a point");
            cudaPointerAttributes pa;          // every other kind carries a device pointer in this field: catch the mix-up
            if (mles[i].dptr && cudaPointerGetAttributes(&pa, mles[i].dptr) == cudaSuccess && pa.type == cudaMemoryTypeDevice)
                return set_err(c, CG_ERR_INVALID, "cg_sumcheck_create: a CG_MLE_EQ descriptor carries its point in HOST memory (dptr is a device pointer)");
            cudaGetLastError();
            continue;
        }
        if (mles[i].is_ext > CG_MLE_EQ) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_create: unknown MLE kind");
        if (!mles[i].dptr || mles[i].len > (1ULL << num_vars) || ((uintptr_t)mles[i].dptr & 15))
            return set_err(c, CG_ERR_INVALID, "cg_sumcheck_create: MLE pointer null/unaligned or len > 2^num_vars");
    }
    CU(c, cudaSetDevice(c->device));
    cg_sumcheck* sc = new cg_sumcheck();
    sc->ctx = c;
    c->live_sc++;
    sc->stream = st;
    sc->n_mles = n_mles;
    sc->num_vars = num_vars;
    sc->degree = degree;
    sc->flags = flags;
    sc->mles.resize(n_mles);
    const uint64_t n = 1ULL << num_vars;
    int rc = CG_OK;
    for (uint32_t i = 0; i < n_mles && rc == CG_OK; i++) {
        if (mles[i].is_ext == CG_MLE_EQ) {
            sc->mles[i].orig = nullptr;
            sc->mles[i].orig_is_ext = 1;
            const uint64_t* pt = (const uint64_t*)mles[i].dptr;
            if (!sc->veq.have) {   // the first one may run split-eq rounds (decided once the term shape is known)
                sc->veq.have = true;
                sc->veq.idx = i;
                sc->veq.h_point.resize(2 * (size_t)num_vars);
                for (size_t q = 0; q < sc->veq.h_point.size(); q++) sc->veq.h_point[q] = pt[q] >= GL_P ? pt[q] - GL_P : pt[q];
            } else {
                rc = veq_build_full(sc, i, pt);
            }
            continue;
        }
        sc->mles[i].orig = mles[i].dptr;
        sc->mles[i].orig_is_ext = mles[i].is_ext ? 1 : 0;
        const size_t es = mles[i].is_ext ? 16 : 8;
        const bool misaligned32 = ((uintptr_t)mles[i].dptr & 31) != 0;   // 256-bit loads need 32 B
        if (mles[i].len < n || misaligned32) {   // occupied prefix / odd alignment: pooled zero-padded copy
            rc = sc_alloc(sc, es * n, &sc->mles[i].padded);
            if (rc == CG_OK && cudaMemsetAsync(sc->mles[i].padded, 0, es * n, st) != cudaSuccess) rc = set_err(c, CG_ERR_CUDA, "memset failed");
            if (rc == CG_OK && cudaMemcpyAsync(sc->mles[i].padded, mles[i].dptr, es * mles[i].len, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                rc = set_err(c, CG_ERR_CUDA, "pad copy failed");
        }
    }
    if (rc == CG_OK && num_vars >= 1) {
        const size_t need = (size_t)n_mles * ws_per(n) * sizeof(ext_t);
        if (tl_lent_ws && tl_lent_ws->ptr && tl_lent_ws->bytes >= need) sc->ws = tl_lent_ws->ptr;   // lent by the caller (not owned)
        else rc = sc_alloc(sc, need, &sc->ws);
    }
    if (rc == CG_OK) {
        // one slab for the small per-sumcheck state: [ticket | error | final | msgs | chal | partials]
        auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
        const size_t o_ticket = 0, o_err = 256, o_final = 512;
        const size_t o_msgs = o_final + up(sizeof(ext_t) * (n_mles + 1));
        const size_t o_chal = o_msgs + up(sizeof(ext_t) * (size_t)(num_vars + 4) * degree);
        const size_t o_part = o_chal + up(sizeof(ext_t) * (num_vars + 4));
        const size_t total = o_part + sizeof(ext_t) * CG_MAX_BLOCKS * CG_MAX_DEGREE;
        void* slab = nullptr;
        rc = sc_alloc(sc, total, &slab);
        if (rc == CG_OK && cudaMemsetAsync(slab, 0, 512, st) != cudaSuccess) rc = set_err(c, CG_ERR_CUDA, "memset failed");
        char* b = (char*)slab;
        sc->out.ticket = (unsigned*)(b + o_ticket);
        sc->d_error = (int*)(b + o_err);
        sc->d_mid_ticket = (unsigned*)(b + 128);
        sc->d_mid_flag = (unsigned*)(b + 192);
        sc->d_per_ticket = (unsigned*)(b + 320);
        sc->d_per_flag = (unsigned*)(b + 384);
        sc->d_final = (ext_t*)(b + o_final);
        sc->d_msgs = (ext_t*)(b + o_msgs);
        sc->d_chal = (ext_t*)(b + o_chal);
        sc->out.partials = (ext_t*)(b + o_part);
    }
    if (rc == CG_OK) {
        sc->h_pinned_bytes = sizeof(ext_t) * (CG_MAX_DEGREE + 4) + sizeof(TailMailbox) + 64;
        sc->h_pinned = (uint64_t*)pinned_get(c, sc->h_pinned_bytes);
        if (!sc->h_pinned) rc = set_err(c, CG_ERR_CUDA, "cudaHostAlloc failed");
    }
    if (rc != CG_OK) { cg_sumcheck_destroy(sc); return rc; }
    *out = sc;
    return CG_OK;
}

// upload per-round slot tables for the generic kernels
static int sc_ensure_tables(cg_sumcheck* sc) {
    if (sc->tables_ready) return CG_OK;
    cg_ctx* c = sc->ctx;
    const uint32_t m = sc->n_mles, nv = sc->num_vars;
    std::vector<MleSlot> slots((size_t)(nv + 1) * m);
    std::vector<FoldSlot> folds((size_t)(nv ? nv : 1) * m);
    for (uint32_t f = 0; f <= nv; f++)
        for (uint32_t i = 0; i < m; i++) {
            if (f == nv && nv > 0) { slots[(size_t)f * m + i] = MleSlot{sc->d_final + i, 1u, 0u}; continue; }
            slots[(size_t)f * m + i] = MleSlot{mle_buf(sc, i, f), mle_is_ext(sc, i, f), f == 0 ? 1u : 0u};
        }
    for (uint32_t f = 0; f < nv; f++)
        for (uint32_t i = 0; i < m; i++) {
            ext_t* dst = (f + 1 == nv) ? sc->d_final + i : (ext_t*)mle_buf(sc, i, f + 1);
            folds[(size_t)f * m + i] = FoldSlot{mle_buf(sc, i, f), dst, mle_is_ext(sc, i, f), f == 0 ? 1u : 0u};
        }
    void* p = nullptr;
    CHK(sc_alloc(sc, sizeof(MleSlot) * slots.size() + 16, &p));
    sc->d_slots = (MleSlot*)p;
    CU(c, cudaMemcpyAsync(p, slots.data(), sizeof(MleSlot) * slots.size(), cudaMemcpyHostToDevice, sc->stream));
    CHK(sc_alloc(sc, sizeof(FoldSlot) * folds.size() + 16, &p));
    sc->d_fold = (FoldSlot*)p;
    CU(c, cudaMemcpyAsync(p, folds.data(), sizeof(FoldSlot) * folds.size(), cudaMemcpyHostToDevice, sc->stream));
    // term tables (generic kernels only)
    CHK(sc_alloc(sc, sizeof(ext_t) * (sc->n_terms + 1), &p)); sc->d_coeff = (ext_t*)p;
    CHK(sc_alloc(sc, sizeof(uint32_t) * (sc->n_terms + 2), &p)); sc->d_off = (uint32_t*)p;
    CHK(sc_alloc(sc, sizeof(uint32_t) * (sc->h_idx.size() + 1), &p)); sc->d_idx = (uint32_t*)p;
    if (sc->n_terms) CU(c, cudaMemcpyAsync(sc->d_coeff, sc->h_coeff.data(), sizeof(ext_t) * sc->n_terms, cudaMemcpyHostToDevice, sc->stream));
    CU(c, cudaMemcpyAsync(sc->d_off, sc->h_off.data(), sizeof(uint32_t) * (sc->n_terms + 1), cudaMemcpyHostToDevice, sc->stream));
    if (!sc->h_idx.empty()) CU(c, cudaMemcpyAsync(sc->d_idx, sc->h_idx.data(), sizeof(uint32_t) * sc->h_idx.size(), cudaMemcpyHostToDevice, sc->stream));
    if (sc->plan_on) {
        auto up32 = [&](const std::vector<uint32_t>& v, uint32_t** d) -> int {
            void* q = nullptr;
            CHK(sc_alloc(sc, sizeof(uint32_t) * (v.size() + 1), &q));
            *d = (uint32_t*)q;
            if (!v.empty()) CU(c, cudaMemcpyAsync(q, v.data(), sizeof(uint32_t) * v.size(), cudaMemcpyHostToDevice, sc->stream));
            return CG_OK;
        };
        CHK(up32(sc->p_g_term_off, &sc->d_g_term_off));
        CHK(up32(sc->p_g_ext_off, &sc->d_g_ext_off));
        CHK(up32(sc->p_g_ext_idx, &sc->d_g_ext_idx));
        CHK(up32(sc->pf_g_term_off, &sc->d_fg_term_off));
        CHK(up32(sc->pf_g_ext_off, &sc->d_fg_ext_off));
        CHK(up32(sc->pf_g_ext_idx, &sc->d_fg_ext_idx));
        CHK(up32(sc->p_t_off, &sc->d_pt_off));
        CHK(up32(sc->p_t_idx, &sc->d_pt_idx));
        CHK(sc_alloc(sc, sizeof(uint64_t) * (sc->p_t_coeff.size() + 1), &p));
        sc->d_pt_coeff = (uint64_t*)p;
        if (!sc->p_t_coeff.empty()) CU(c, cudaMemcpyAsync(p, sc->p_t_coeff.data(), sizeof(uint64_t) * sc->p_t_coeff.size(), cudaMemcpyHostToDevice, sc->stream));
    }
    CU(c, cudaStreamSynchronize(sc->stream));   // the local slot vectors go out of scope
    sc->tables_ready = true;
    return CG_OK;
}

static int sc_create_terms(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, const uint64_t* coeff,
                           const uint32_t* off, const uint32_t* idx, uint32_t n_terms, uint32_t num_vars,
                           uint32_t degree, uint32_t flags, cg_stream s, const ext_t* veq_scale, cg_sumcheck** out,
                           cg_comm* comm = nullptr, uint32_t extra_rounds = 0);
CG_EXPORT int cg_sumcheck_create(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, const uint64_t* coeff,
                                 const uint32_t* off, const uint32_t* idx, uint32_t n_terms, uint32_t num_vars,
                                 uint32_t degree, uint32_t flags, cg_stream s, cg_sumcheck** out) {
    return sc_create_terms(c, mles, n_mles, coeff, off, idx, n_terms, num_vars, degree, flags, s, nullptr, out);
}
static int sc_create_terms(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, const uint64_t* coeff,
                           const uint32_t* off, const uint32_t* idx, uint32_t n_terms, uint32_t num_vars,
                           uint32_t degree, uint32_t flags, cg_stream s, const ext_t* veq_scale, cg_sumcheck** out,
                           cg_comm* comm, uint32_t extra_rounds) {
    if (!c) return CG_ERR_INVALID;
    if (n_terms && (!coeff || !off || !idx)) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_create: null term table");
    for (uint32_t t = 0; t < n_terms; t++) {
        if (off[t + 1] < off[t] || off[t + 1] - off[t] > degree) return set_err(c, CG_ERR_INVALID, "term has more factors than `degree`");
        for (uint32_t q = off[t]; q < off[t + 1]; q++)
            if (idx[q] >= n_mles) return set_err(c, CG_ERR_INVALID, "term references an MLE index out of range");
    }
    cudaStream_t st = S(c, s);
    cg_sumcheck* sc = nullptr;
    CHK(sc_create_common(c, mles, n_mles, num_vars, degree, flags, st, &sc));
    sc->n_terms = n_terms;
    sc->comm = comm;                   // sharded prove: known before the split-eq rounds are planned
    sc->extra_rounds = extra_rounds;
    int rc = CG_OK;
    const uint32_t n_idx = n_terms ? off[n_terms] : 0;
    sc->h_coeff.resize(2 * (size_t)n_terms);   // canonicalised on the host (inputs may be any u64)
    for (size_t i = 0; i < sc->h_coeff.size(); i++) sc->h_coeff[i] = coeff[i] >= GL_P ? coeff[i] - GL_P : coeff[i];
    sc->h_off.assign(n_terms + 1, 0);
    if (n_terms) memcpy(sc->h_off.data(), off, sizeof(uint32_t) * (n_terms + 1));
    sc->h_idx.assign(n_idx, 0);
    if (n_idx) memcpy(sc->h_idx.data(), idx, sizeof(uint32_t) * n_idx);
    // grouped plan: terms that share the same set of round-0 ext factors (selectors / eq) are summed first
    if (rc == CG_OK && n_terms && degree <= 4) {
        std::map<std::vector<uint32_t>, std::vector<uint32_t>> groups;   // ext-factor multiset -> term ids
        for (uint32_t tm = 0; tm < n_terms; tm++) {
            std::vector<uint32_t> key;
            for (uint32_t q = off[tm]; q < off[tm + 1]; q++) if (mles[idx[q]].is_ext) key.push_back(idx[q]);
            std::sort(key.begin(), key.end());
            groups[key].push_back(tm);
        }
        // two plans over the same term order: whole groups (large rounds: one selector multiply per group) and
        // groups split into chunks of 8 terms (small rounds: the chunks are spread over blockIdx.y)
        sc->p_g_term_off.push_back(0);
        sc->p_g_ext_off.push_back(0);
        sc->p_t_off.push_back(0);
        sc->pf_g_term_off.push_back(0);
        sc->pf_g_ext_off.push_back(0);
        for (auto& kv : groups) {
            const uint32_t first_term = (uint32_t)sc->p_t_off.size() - 1;
            for (uint32_t tm : kv.second) {
                const uint64_t c0 = sc->h_coeff[2 * tm], c1 = sc->h_coeff[2 * tm + 1];
                sc->p_t_coeff.push_back(c0);
                sc->p_t_coeff.push_back(c1);
                sc->p_t_coeff.push_back((uint64_t)(((unsigned __int128)c1 * 7) % GL_P));
                for (uint32_t q = off[tm]; q < off[tm + 1]; q++) if (!mles[idx[q]].is_ext) sc->p_t_idx.push_back(idx[q]);
                sc->p_t_off.push_back((uint32_t)sc->p_t_idx.size());
            }
            const uint32_t last_term = (uint32_t)sc->p_t_off.size() - 1;
            for (uint32_t e : kv.first) sc->p_g_ext_idx.push_back(e);
            sc->p_g_ext_off.push_back((uint32_t)sc->p_g_ext_idx.size());
            sc->p_g_term_off.push_back(last_term);
            for (uint32_t b = first_term; b < last_term; b += 8) {
                for (uint32_t e : kv.first) sc->pf_g_ext_idx.push_back(e);
                sc->pf_g_ext_off.push_back((uint32_t)sc->pf_g_ext_idx.size());
                sc->pf_g_term_off.push_back(std::min(b + 8, last_term));
            }
        }
        sc->plan_on = true;
    }
    // shape detection: one degree-3 product of three distinct ext MLEs -> tower kernel (T3).  Only when the list holds nothing
    // else (n_mles == 3): the specialised kernels fold the term's MLEs only, and an unreferenced MLE (a zerocheck layer passes
    // every witin/fixed/structural column) must still be folded for get_mle_flatten_final_evaluations.
    if (rc == CG_OK && !(flags & CG_SC_FORCE_GENERIC) && n_terms == 1 && n_mles == 3 && degree == 3 && off[1] - off[0] == 3 && num_vars >= 1) {
        const uint32_t a = idx[off[0]], b = idx[off[0] + 1], d = idx[off[0] + 2];
        if (a != b && b != d && a != d && mles[a].is_ext && mles[b].is_ext && mles[d].is_ext) {
            sc->tl.on = true;
            sc->tl.eq = a;
            sc->tl.prod = {b, d};
            if (sc->veq.have && sc->veq.idx == b) { sc->tl.eq = b; sc->tl.prod = {a, d}; }   // the factors commute:
            if (sc->veq.have && sc->veq.idx == d) { sc->tl.eq = d; sc->tl.prod = {a, b}; }   // the virtual eq takes the eq slot
            ext_t al = ext_t{coeff[0] >= GL_P ? coeff[0] - GL_P : coeff[0], coeff[1] >= GL_P ? coeff[1] - GL_P : coeff[1]};
            sc->tl.prod_alpha = {al};
            sc->tl.alpha_one = (al.c0 == 1 && al.c1 == 0);
        }
    }
    if (rc == CG_OK && sc->veq.have && veq_scale) sc->veq.scale = *veq_scale;
    if (rc == CG_OK && sc->veq.have && !(flags & CG_SC_INT_DEFER_VEQ)) {   // (deferred: the tower prover sets its layout first)
        const bool split_ok = sc->tl.on && sc->tl.eq == sc->veq.idx && sc->tl.prod.size() == 2 && sc->tl.lk.empty() &&
                              !(flags & (CG_SC_NO_FUSE | CG_SC_FORCE_GENERIC)) && num_vars >= 20 && num_vars <= 32;
        rc = split_ok ? veq_setup_split(sc) : veq_build_full(sc, sc->veq.idx, sc->veq.h_point.data());
    }
    if (rc != CG_OK) { cg_sumcheck_destroy(sc); return rc; }
    *out = sc;
    return CG_OK;
}

CG_EXPORT uint32_t cg_sumcheck_round(const cg_sumcheck* sc) { return sc ? sc->round : 0; }

static int launch_fold(cg_sumcheck* sc, uint32_t f) {   // fold state f -> f+1
    cg_ctx* c = sc->ctx;
    CHK(sc_ensure_tables(sc));
    const uint64_t n_out = 1ULL << (sc->num_vars - f - 1);
    dim3 grid(grid_for(c, n_out), sc->n_mles);
    if (sc->n_mles == 0) return CG_OK;
    fold_kernel<<<grid, CG_THREADS, 0, sc->stream>>>(sc->d_fold + (size_t)f * sc->n_mles, n_out, sc->pending_r, sc->pending_r_ptr);
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    return CG_OK;
}
template <int D>
static void launch_generic_d(cg_sumcheck* sc, const GenericArgs& a, unsigned grid) {
    generic_round_kernel<D><<<grid, CG_THREADS, 0, sc->stream>>>(a);
}
static int launch_generic_eval(cg_sumcheck* sc, uint32_t f, const RoundOut& ro) {   // evaluate state f
    cg_ctx* c = sc->ctx;
    CHK(sc_ensure_tables(sc));
    GenericArgs a;
    a.mles = sc->d_slots + (size_t)f * sc->n_mles;
    a.coeff = sc->d_coeff;
    a.off = sc->d_off;
    a.idx = sc->d_idx;
    a.n_terms = sc->n_terms;
    a.n_pairs = 1ULL << (sc->num_vars - f - 1);
    a.out = ro;
    const unsigned grid = grid_for(c, a.n_pairs);
    if (sc->plan_on && !(sc->flags & CG_SC_NO_PLAN)) {
        GroupedArgs ga;
        ga.mles = a.mles;
        ga.g_term_off = sc->d_g_term_off;
        ga.g_ext_off = sc->d_g_ext_off;
        ga.g_ext_idx = sc->d_g_ext_idx;
        ga.t_coeff = sc->d_pt_coeff;
        ga.t_off = sc->d_pt_off;
        ga.t_idx = sc->d_pt_idx;
        ga.n_groups = (uint32_t)sc->p_g_term_off.size() - 1;
        ga.n_pairs = a.n_pairs;
        ga.out = ro;
        unsigned gx = grid_for(c, ga.n_pairs, 2), gy = 1;
        if (gx < (unsigned)c->sm_count) {   // too few items to fill the chip: spread 8-term chunks over blockIdx.y
            ga.g_term_off = sc->d_fg_term_off;
            ga.g_ext_off = sc->d_fg_ext_off;
            ga.g_ext_idx = sc->d_fg_ext_idx;
            ga.n_groups = (uint32_t)sc->pf_g_term_off.size() - 1;
            gy = std::min<unsigned>(ga.n_groups, std::max(1u, 2u * (unsigned)c->sm_count / gx));
            if ((uint64_t)gx * gy > CG_MAX_BLOCKS) gy = CG_MAX_BLOCKS / gx;
        }
        const dim3 ggrid(gx, gy);
#define CG_GROUPED(DD)                                                                                              \
    do {                                                                                                            \
        if (f == 0) grouped_round_kernel<DD, true><<<ggrid, CG_THREADS, 0, sc->stream>>>(ga);                      \
        else grouped_round_kernel<DD, false><<<ggrid, CG_THREADS, 0, sc->stream>>>(ga);                            \
    } while (0)
        switch (sc->degree) {
            case 1: CG_GROUPED(1); break;
            case 2: CG_GROUPED(2); break;
            case 3: CG_GROUPED(3); break;
            default: CG_GROUPED(4); break;
        }
#undef CG_GROUPED
        LAUNCHED(c);
        CU(c, cudaGetLastError());
        return CG_OK;
    }
    switch (sc->degree) {
        case 1: launch_generic_d<1>(sc, a, grid); break;
        case 2: launch_generic_d<2>(sc, a, grid); break;
        case 3: launch_generic_d<3>(sc, a, grid); break;
        case 4: launch_generic_d<4>(sc, a, grid); break;
        case 5: launch_generic_d<5>(sc, a, grid); break;
        case 6: launch_generic_d<6>(sc, a, grid); break;
        case 7: launch_generic_d<7>(sc, a, grid); break;
        default: launch_generic_d<8>(sc, a, grid); break;
    }
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    return CG_OK;
}
#ifndef CG_VEQ_MINB_DEFAULT
#define CG_VEQ_MINB_DEFAULT 2
#endif
// split-eq round (virtual eq): evaluate state f, folding f-1 -> f first when `fold`
static int launch_veq(cg_sumcheck* sc, uint32_t f, bool fold, const RoundOut& ro) {
    cg_ctx* c = sc->ctx;
    const VeqState& v = sc->veq;
    VeqArgs a;
    memset(&a, 0, sizeof(a));
    const uint32_t src = fold ? f - 1 : f;
    for (int z = 0; z < 2; z++) {
        a.in[z] = (const ext_t*)mle_buf(sc, sc->tl.prod[z], src);
        a.out[z] = fold ? (ext_t*)mle_buf(sc, sc->tl.prod[z], f) : nullptr;
    }
    a.L = v.d_L + ((size_t)f << CG_VEQ_LO_BITS);
    a.H = v.d_H + v.h_off[f];
    a.n_rows = 1ULL << (sc->num_vars - f - 1 - CG_VEQ_LO_BITS);
    a.r = sc->pending_r;
    a.r_ptr = sc->pending_r_ptr;
    a.fin.w = v.d_w;
    a.fin.inv1mw = v.d_inv1mw;
    a.fin.prefix = v.d_prefix;
    a.fin.qstate = v.d_qstate;
    a.fin.scale = v.scale;
    a.fin.sharded = (sc->comm && sc->comm->nranks > 1) ? 1 : 0;
    a.fin.round = f;
    a.fin.fold = fold ? 1 : 0;
    a.fin.r = sc->pending_r;
    a.fin.r_ptr = sc->pending_r_ptr;
    a.out_ = ro;
    uint64_t blocks = (uint64_t)c->sm_count * 2;
    if (blocks > a.n_rows) blocks = a.n_rows;
    const unsigned grid = (unsigned)blocks;
    const bool canon = (src == 0);
#define CG_VEQ_LAUNCH(MB)                                                                        \
    do {                                                                                         \
        if (!fold) veq_round_kernel<false, false, MB><<<grid, 256, 0, sc->stream>>>(a);          \
        else if (canon) veq_round_kernel<true, true, MB><<<grid, 256, 0, sc->stream>>>(a);       \
        else veq_round_kernel<true, false, MB><<<grid, 256, 0, sc->stream>>>(a);                 \
    } while (0)
    static const int use_tma = []() { const char* e = getenv("CG_VEQ_TMA"); return e ? atoi(e) : 1; }();
    static const int minb = []() { const char* e = getenv("CG_VEQ_MINB"); return e ? atoi(e) : CG_VEQ_MINB_DEFAULT; }();
    // claim-derived round (two products per pair): every round after the first, default configuration
    const bool derive = fold && v.derive_ok && use_tma && minb != 3;
    a.fin.derive = derive ? 1 : 0;
    if (derive) {
        uint64_t tb = (uint64_t)c->sm_count * 2;
        if (tb > a.n_rows) tb = a.n_rows;
        if (canon) veq_tma_kernel<true, true, 2, true><<<(unsigned)tb, 256, VeqTmaCfg<true, 2>::SMEM, sc->stream>>>(a);
        else veq_tma_kernel<true, false, 2, true><<<(unsigned)tb, 256, VeqTmaCfg<true, 2>::SMEM, sc->stream>>>(a);
    } else if (use_tma) {   // rows staged through shared memory by cp.async.bulk (default)
        // CG_VEQ_MINB = 2 | 3 resident blocks per SM (A/B switch; see VeqTmaCfg)
        uint64_t tb = (uint64_t)c->sm_count * (minb == 3 ? 3 : 2);
        if (tb > a.n_rows) tb = a.n_rows;
#define CG_VEQ_TMA_LAUNCH(MB)                                                                                                   \
    do {                                                                                                                        \
        if (!fold) veq_tma_kernel<false, false, MB><<<(unsigned)tb, 256, VeqTmaCfg<false, MB>::SMEM, sc->stream>>>(a);          \
        else if (canon) veq_tma_kernel<true, true, MB><<<(unsigned)tb, 256, VeqTmaCfg<true, MB>::SMEM, sc->stream>>>(a);        \
        else veq_tma_kernel<true, false, MB><<<(unsigned)tb, 256, VeqTmaCfg<true, MB>::SMEM, sc->stream>>>(a);                  \
    } while (0)
        if (minb == 3) CG_VEQ_TMA_LAUNCH(3); else CG_VEQ_TMA_LAUNCH(2);
#undef CG_VEQ_TMA_LAUNCH
    } else {         // CG_VEQ_TMA=0: the same round with plain 256-bit loads (A/B comparison)
        CG_VEQ_LAUNCH(2);
    }
#undef CG_VEQ_LAUNCH
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    return CG_OK;
}
// the layer's arrays for a launch that evaluates state f (fold==false) or folds f-1 -> f and evaluates (fold==true)
static void fill_tower_args(cg_sumcheck* sc, uint32_t f, bool fold, bool with_eq, TowerArgs& a, bool& any_virt) {
    const TowerLayout& tl = sc->tl;
    memset(&a, 0, sizeof(a));
    const uint32_t src = fold ? f - 1 : f;
    auto in = [&](uint32_t i) { return (const ext_t*)mle_buf(sc, i, src); };
    auto outp = [&](uint32_t i) { return fold ? (ext_t*)mle_buf(sc, i, f) : (ext_t*)nullptr; };
    if (with_eq) {
        a.eq_in = in(tl.eq);
        a.eq_out = outp(tl.eq);
    }
    a.n_prod = (int)tl.prod_alpha.size();
    a.n_logup = (int)tl.lk_an.size();
    for (int p = 0; p < a.n_prod; p++) {
        for (int z = 0; z < 2; z++) { a.prod_in[p][z] = in(tl.prod[2 * p + z]); a.prod_out[p][z] = outp(tl.prod[2 * p + z]); }
        a.alpha_prod[p] = tl.prod_alpha[p];
    }
    for (int l = 0; l < a.n_logup; l++) {
        for (int z = 0; z < 4; z++) { a.lk_in[l][z] = in(tl.lk[4 * l + z]); a.lk_out[l][z] = outp(tl.lk[4 * l + z]); }
        a.alpha_num[l] = tl.lk_an[l];
        a.alpha_den[l] = tl.lk_ad[l];
    }
    a.alpha_one = tl.alpha_one ? 1 : 0;
    a.n_pairs = 1ULL << (sc->num_vars - f - 1);
    a.r = sc->pending_r;
    a.r_ptr = sc->pending_r_ptr;
    any_virt = false;                // virtual tower leaves: only a launch that reads the original inputs sees them
    if (src == 0 && !sc->virt.empty()) {
        for (int p = 0; p < a.n_prod; p++)
            for (int z = 0; z < 2; z++) { a.virt[1 + 2 * p + z] = sc->virt[tl.prod[2 * p + z]]; any_virt |= a.virt[1 + 2 * p + z] != nullptr; }
        for (int l = 0; l < a.n_logup; l++)
            for (int z = 0; z < 4; z++) { a.virt[1 + 2 * a.n_prod + 4 * l + z] = sc->virt[tl.lk[4 * l + z]]; any_virt |= a.virt[1 + 2 * a.n_prod + 4 * l + z] != nullptr; }
    }
}
// split-eq round of a general tower layout (tveq_round_kernel)
static int launch_tveq(cg_sumcheck* sc, uint32_t f, bool fold, const RoundOut& ro) {
    cg_ctx* c = sc->ctx;
    const VeqState& v = sc->veq;
    TVeqArgs a;
    memset(&a, 0, sizeof(a));
    bool any_virt = false;
    fill_tower_args(sc, f, fold, false, a.t, any_virt);
    const uint32_t src = fold ? f - 1 : f;
    a.S = v.S;
    if (any_virt) {
        if (!v.m1 || f > 1) return set_err(c, CG_ERR_STATE, "split-eq rounds: no tables for a virtual-leaf launch");
        a.U = v.d_LR[f];
        a.F = v.d_HS[f];
        a.lo_bits = v.m1_lo[f];
        const TowerLayout& tl = sc->tl;
        const bool info = sc->virt_h.size() == sc->virt.size();
        auto is_one = [](const ext_t& e) { return e.c0 == 1 && e.c1 == 0; };
        for (size_t l = 0; info && l < tl.lk_an.size(); l++) {   // numerators described as "no records, default one"
            const uint32_t i0 = tl.lk[4 * l], i1 = tl.lk[4 * l + 1];
            if (sc->virt[i0] && sc->virt[i1] && !sc->virt_h[i0].n_records && !sc->virt_h[i1].n_records && is_one(sc->virt_h[i0].def) && is_one(sc->virt_h[i1].def))
                a.t.pone_mask |= 1u << l;
        }
        // record padding: items past every slot's records are constant pairs; their sum is added per chunk
        const uint64_t n_lo = 1ULL << a.lo_bits, chunk = std::min<uint64_t>(n_lo, 256);
        a.pad_lo = (uint32_t)n_lo;
        bool all_virt = info;
        uint64_t pad = 0;
        for (size_t i = 1; all_virt && i < sc->virt.size(); i++) {
            if (!sc->virt[i] || sc->virt_h[i].l2m != sc->virt_l2m) all_virt = false;   // (one index layout for every slot)
            else pad = std::max<uint64_t>(pad, ((uint64_t)sc->virt_h[i].n_records + (2ULL << f) - 1) >> (f + 1));
        }
        if (all_virt && pad < n_lo && n_lo / chunk <= CG_TVEQ_PAD_CHUNKS && (n_lo << f) << 1 == (1ULL << sc->virt_l2m)) {
            // K = value of the layer polynomial on an all-default pair (per unit of eq weight), e[lo] = eq over the low variables
            ext_t K{0, 0};
            auto D = [&](uint32_t i) { return sc->virt_h[i].def; };
            for (size_t p = 0; p < tl.prod_alpha.size(); p++) K = hx_add(K, hx_mul(tl.prod_alpha[p], hx_mul(D(tl.prod[2 * p]), D(tl.prod[2 * p + 1]))));
            for (size_t l = 0; l < tl.lk_an.size(); l++) {
                const uint32_t p1 = tl.lk[4 * l], p2 = tl.lk[4 * l + 1], q1 = tl.lk[4 * l + 2], q2 = tl.lk[4 * l + 3];
                K = hx_add(K, hx_mul(tl.lk_an[l], hx_add(hx_mul(D(p1), D(q2)), hx_mul(D(p2), D(q1)))));
                K = hx_add(K, hx_mul(tl.lk_ad[l], hx_mul(D(q1), D(q2))));
            }
            std::vector<ext_t> e(n_lo);
            e[0] = ext_t{1, 0};
            for (uint32_t b = 0; b < a.lo_bits; b++) {
                const ext_t w{v.h_point[2 * (f + 1 + b)], v.h_point[2 * (f + 1 + b) + 1]}, om{hx_submod(1, w.c0), hx_submod(0, w.c1)};
                for (uint64_t x = 0; x < (1ULL << b); x++) {
                    e[x | (1ULL << b)] = hx_mul(e[x], w);
                    e[x] = hx_mul(e[x], om);
                }
            }
            for (uint64_t ch = 0; ch < n_lo / chunk; ch++) {
                ext_t sum{0, 0};
                for (uint64_t lo = std::max(ch * chunk, pad); lo < (ch + 1) * chunk; lo++) sum = hx_add(sum, e[lo]);
                a.pad_sum[ch] = hx_mul(K, sum);
            }
            a.pad_lo = (uint32_t)pad;
        }
    } else {
        a.U = v.d_UA + v.h_off[f] * v.S;
        a.F = v.d_L + ((size_t)f << CG_VEQ_LO_BITS);
    }
    a.fin.w = v.d_w;
    a.fin.inv1mw = v.d_inv1mw;
    a.fin.prefix = v.d_prefix;
    a.fin.qstate = v.d_qstate;
    a.fin.scale = v.scale;
    a.fin.sharded = (sc->comm && sc->comm->nranks > 1) ? 1 : 0;
    a.fin.round = f;
    a.fin.fold = fold ? 1 : 0;
    a.fin.r = sc->pending_r;
    a.fin.r_ptr = sc->pending_r_ptr;
    const bool derive = v.derive_ok && (fold || (f == 0 && v.have_claim));
    a.fin.derive = derive ? 1 : 0;
    a.out_ = ro;
    const bool canon = (src == 0);
    uint64_t blocks = (uint64_t)c->sm_count * 2;   // one resident wave (2 blocks per SM)
    if (!any_virt) blocks = std::min<uint64_t>(blocks, a.t.n_pairs >> CG_VEQ_LO_BITS);
    else {
        const uint64_t n_lo = 1ULL << a.lo_bits, units = (a.t.n_pairs >> a.lo_bits) * (n_lo / std::min<uint64_t>(n_lo, 256));
        blocks = std::min<uint64_t>(blocks, (4 * units + 255) / 256);   // four lanes per unit
    }
    const unsigned grid = (unsigned)(blocks ? blocks : 1);
#define CG_TVEQ(F, CN, DR, VR) tveq_round_kernel<F, CN, DR, VR><<<grid, 256, 0, sc->stream>>>(a)
    if (any_virt) {
        if (!fold) { if (derive) CG_TVEQ(false, true, true, true); else CG_TVEQ(false, true, false, true); }
        else if (derive) CG_TVEQ(true, true, true, true);
        else CG_TVEQ(true, true, false, true);
    } else if (!fold) {
        if (derive) CG_TVEQ(false, true, true, false);   // round 0 with a known claim (f == 0: caller-provided buffers)
        else if (canon) CG_TVEQ(false, true, false, false);
        else CG_TVEQ(false, false, false, false);
    } else if (derive) {
        if (canon) CG_TVEQ(true, true, true, false); else CG_TVEQ(true, false, true, false);
    } else {
        if (canon) CG_TVEQ(true, true, false, false); else CG_TVEQ(true, false, false, false);
    }
#undef CG_TVEQ
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    return CG_OK;
}
// tower-shaped kernel: evaluate state f (fold==false) or fold f-1 -> f and evaluate (fold==true)
static int launch_tower(cg_sumcheck* sc, uint32_t f, bool fold, const RoundOut& ro) {
    if (sc->veq.split) return sc->veq.general ? launch_tveq(sc, f, fold, ro) : launch_veq(sc, f, fold, ro);
    cg_ctx* c = sc->ctx;
    TowerArgs a;
    bool any_virt = false;
    fill_tower_args(sc, f, fold, true, a, any_virt);
    a.out = ro;
    const uint32_t src = fold ? f - 1 : f;
    const bool canon = (src == 0);   // reading caller-provided buffers
    const bool simple = a.n_prod == 1 && a.n_logup == 0 && a.alpha_one && !any_virt;
    // launch shape: threads x resident blocks per SM (one persistent wave, grid-stride loop)
    static const int cfg = []() { const char* e = getenv("CG_TOWER_CFG"); return e ? atoi(e) : 0; }();
    const int threads = simple ? (cfg == 1 || cfg == 2 ? 128 : (cfg == 3 ? 192 : 256)) : 256;
    const int minb = simple ? (cfg == 1 ? 5 : (cfg == 2 ? 6 : (cfg == 3 ? 3 : 2))) : 2;
    uint64_t blocks = (a.n_pairs + threads - 1) / threads;
    const uint64_t cap = (uint64_t)c->sm_count * minb * (simple ? 1 : 2);
    if (blocks > cap) blocks = cap;
    if (blocks > CG_MAX_BLOCKS) blocks = CG_MAX_BLOCKS;
    const unsigned grid = (unsigned)(blocks ? blocks : 1);
#define CG_TOWER_LAUNCH2(F, CN, SI, TH, MB) tower_round_kernel<F, CN, SI, TH, MB><<<grid, TH, 0, sc->stream>>>(a)
#define CG_TOWER_LAUNCH(F, CN)                                                   \
    do {                                                                         \
        if (!simple) CG_TOWER_LAUNCH2(F, CN, false, 256, 2);                     \
        else if (cfg == 1) CG_TOWER_LAUNCH2(F, CN, true, 128, 5);                \
        else if (cfg == 2) CG_TOWER_LAUNCH2(F, CN, true, 128, 6);                \
        else if (cfg == 3) CG_TOWER_LAUNCH2(F, CN, true, 192, 3);                \
        else CG_TOWER_LAUNCH2(F, CN, true, 256, 2);                              \
    } while (0)
    if (any_virt) {
        if (fold) tower_round_kernel<true, true, false, 256, 2, true><<<grid, 256, 0, sc->stream>>>(a);
        else tower_round_kernel<false, true, false, 256, 2, true><<<grid, 256, 0, sc->stream>>>(a);
    } else if (fold) { if (canon) CG_TOWER_LAUNCH(true, true); else CG_TOWER_LAUNCH(true, false); }
    else { if (canon) CG_TOWER_LAUNCH(false, true); else CG_TOWER_LAUNCH(false, false); }
#undef CG_TOWER_LAUNCH
#undef CG_TOWER_LAUNCH2
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    return CG_OK;
}

// enqueue the kernels of the current round's evaluation (applying a pending fold first)
static int sc_enqueue_round(cg_sumcheck* sc, const RoundOut& ro) {
    if (sc->round >= sc->num_vars) return set_err(sc->ctx, CG_ERR_STATE, "round_eval: all variables are already bound");
    CHK(veq_prepare(sc));
    if (sc->pending) {
        const uint32_t f = sc->folds;   // fold f -> f+1, then evaluate state f+1
        if (sc->tl.on && !(sc->flags & CG_SC_NO_FUSE)) {
            CHK(launch_tower(sc, f + 1, true, ro));
        } else {
            CHK(launch_fold(sc, f));
            if (sc->tl.on) CHK(launch_tower(sc, f + 1, false, ro));
            else CHK(launch_generic_eval(sc, f + 1, ro));
        }
        sc->folds++;
        sc->pending = false;
    } else {
        if (sc->tl.on) CHK(launch_tower(sc, sc->folds, false, ro));
        else CHK(launch_generic_eval(sc, sc->folds, ro));
    }
    sc->evaluated = true;
    return CG_OK;
}

static RoundOut make_ro(cg_sumcheck* sc) {
    RoundOut ro = sc->out;
    comm_dev(sc->comm, ro.comm, 1);   // one exchange per round-evaluation launch
    return ro;
}
CG_EXPORT int cg_sumcheck_attach_comm(cg_sumcheck* sc, cg_comm* cm) {
    if (!sc) return CG_ERR_INVALID;
    if (sc->round != 0 || sc->evaluated) return set_err(sc->ctx, CG_ERR_STATE, "attach_comm: must be called before the first round");
    if (sc->veq.have && cm) return set_err(sc->ctx, CG_ERR_UNSUPPORTED, "attach_comm: virtual eq MLEs are not supported with a comm yet");
    sc->comm = cm;
    return CG_OK;
}
CG_EXPORT int cg_sumcheck_round_eval(cg_sumcheck* sc, uint64_t* h_out) {
    if (!sc || !h_out) return CG_ERR_INVALID;
    cg_ctx* c = sc->ctx;
    CU(c, cudaSetDevice(c->device));
    RoundOut ro = make_ro(sc);
    ro.d_out = sc->d_msgs;
    ro.d_tr_state = nullptr;
    ro.d_r_out = nullptr;
    CHK(sc_enqueue_round(sc, ro));
    CU(c, cudaMemcpyAsync(sc->h_pinned, sc->d_msgs, sizeof(ext_t) * sc->degree, cudaMemcpyDeviceToHost, sc->stream));
    CU(c, cudaStreamSynchronize(sc->stream));
    memcpy(h_out, sc->h_pinned, sizeof(ext_t) * sc->degree);
    return CG_OK;
}

static int sc_apply_pending_fold_only(cg_sumcheck* sc) {
    if (!sc->pending) return CG_OK;
    CHK(veq_leave_split(sc));   // a fold without an evaluation cannot advance the split-eq prefix
    CHK(launch_fold(sc, sc->folds));
    sc->folds++;
    sc->pending = false;
    return CG_OK;
}

static int sc_bind_common(cg_sumcheck* sc) {
    sc->round++;
    sc->evaluated = false;
    if (sc->round == sc->num_vars) CHK(sc_apply_pending_fold_only(sc));   // last variable: produce the final evaluations
    return CG_OK;
}
CG_EXPORT int cg_sumcheck_bind(cg_sumcheck* sc, const uint64_t r[2]) {
    if (!sc || !r) return CG_ERR_INVALID;
    if (sc->round >= sc->num_vars) return set_err(sc->ctx, CG_ERR_STATE, "bind: all variables are already bound");
    CU(sc->ctx, cudaSetDevice(sc->ctx->device));
    CHK(sc_apply_pending_fold_only(sc));   // bind without an intervening round_eval: plain fold
    sc->pending = true;
    sc->pending_r = ext_t{r[0] >= GL_P ? r[0] - GL_P : r[0], r[1] >= GL_P ? r[1] - GL_P : r[1]};
    sc->pending_r_ptr = nullptr;
    return sc_bind_common(sc);
}

CG_EXPORT int cg_sumcheck_final_evals(cg_sumcheck* sc, uint64_t* h_out) {
    if (!sc || !h_out) return CG_ERR_INVALID;
    cg_ctx* c = sc->ctx;
    if (sc->round != sc->num_vars) return set_err(c, CG_ERR_STATE, "final_evals: not all variables are bound yet");
    CU(c, cudaSetDevice(c->device));
    if (sc->num_vars == 0) {   // zero-variable sumcheck is a no-op: the single value of each MLE
        for (uint32_t i = 0; i < sc->n_mles; i++) {
            uint64_t v[2] = {0, 0};
            CU(c, cudaMemcpyAsync(v, sc->mles[i].padded ? sc->mles[i].padded : sc->mles[i].orig, sc->mles[i].orig_is_ext ? 16 : 8,
                                  cudaMemcpyDeviceToHost, sc->stream));
            CU(c, cudaStreamSynchronize(sc->stream));
            h_out[2 * i] = v[0] >= GL_P ? v[0] - GL_P : v[0];
            h_out[2 * i + 1] = v[1] >= GL_P ? v[1] - GL_P : v[1];
        }
        return CG_OK;
    }
    CU(c, cudaMemcpyAsync(h_out, sc->d_final, sizeof(ext_t) * sc->n_mles, cudaMemcpyDeviceToHost, sc->stream));
    CU(c, cudaStreamSynchronize(sc->stream));
    return CG_OK;
}

CG_EXPORT int cg_sumcheck_peek(cg_sumcheck* sc, uint32_t i, const void** dptr, uint64_t* len, uint32_t* is_ext) {
    if (!sc || i >= sc->n_mles) return CG_ERR_INVALID;
    CHK(sc_apply_pending_fold_only(sc));
    CHK(veq_leave_split(sc));
    CU(sc->ctx, cudaStreamSynchronize(sc->stream));
    const uint32_t f = sc->folds;
    if (dptr) *dptr = (f == sc->num_vars && f > 0) ? (const void*)(sc->d_final + i) : mle_buf(sc, i, f);
    if (len) *len = 1ULL << (sc->num_vars - f);
    if (is_ext) *is_ext = mle_is_ext(sc, i, f);
    return CG_OK;
}


// ---- persistent tail (all remaining rounds in one CTA, arrays in shared memory)
static TailMailbox* sc_mailbox(cg_sumcheck* sc) {
    return (TailMailbox*)((char*)sc->h_pinned + ((sizeof(ext_t) * (CG_MAX_DEGREE + 4) + 63) & ~(size_t)63));
}
// host side of the mailbox protocol (see TailMailbox)
static void mailbox_reset(TailMailbox* mb) {
    for (int i = 0; i < 16; i++) mb->msg[i] = 0;
    uint64_t z[16] = {0};
    mb->seq_msg = ~0ULL ^ cg_mb_mix(z, 16);   // validates for no (round, word count): a stale buffer can never be accepted
    mb->r[0] = mb->r[1] = 0;
    mb->seq_r = 0;
    mb->abort = 0;
    mb->fin_flag = 0;   // a valid flag is (rounds + 1) ^ checksum: zero words + zero flag never validate for seq >= 1
    for (int i = 0; i < 2 * CG_COMM_GATHER_SLOTS; i++) mb->fin[i] = 0;
    __sync_synchronize();
}
// true when the message of sequence `seq` (n words) is complete; the words are copied to `out`
static bool mailbox_take(TailMailbox* mb, uint32_t n, uint64_t seq, uint64_t* out) {
    const uint64_t flag = mb->seq_msg;
    uint64_t w[16];
    const volatile uint64_t* m = mb->msg;
    for (uint32_t i = 0; i < n; i++) w[i] = m[i];
    if ((flag ^ cg_mb_mix(w, n)) != seq) return false;
    for (uint32_t i = 0; i < n; i++) out[i] = w[i];
    return true;
}
// final evaluations posted by the tail kernel (n ext in the caller's MLE order)
static bool mailbox_take_final(TailMailbox* mb, uint32_t n, uint64_t seq, uint64_t* out) {
    const uint64_t flag = mb->fin_flag;
    uint64_t w[2 * CG_COMM_GATHER_SLOTS], h = 0;
    const volatile uint64_t* f = mb->fin;
    for (uint32_t i = 0; i < 2 * n; i++) { w[i] = f[i]; h ^= cg_mb_word(w[i], i); }
    if ((flag ^ h) != seq) return false;
    if (out) for (uint32_t i = 0; i < 2 * n; i++) out[i] = w[i];
    return true;
}
static void mailbox_reply(TailMailbox* mb, const uint64_t r[2], uint64_t seq) {
    ((volatile uint64_t*)mb->r)[0] = r[0];
    ((volatile uint64_t*)mb->r)[1] = r[1];
    __sync_synchronize();
    mb->seq_r = seq;
}
// Shape of a cluster-tail launch for a layer with n_slots MLEs entering with n0 elements each (cluster-wide):
// n_loc0 = elements per MLE per CTA, C = cluster size, nt = size at which the slices are gathered into CTA 0.
struct TailPlan {
    bool ok = false;
    uint32_t C = 1, n_loc0 = 0, nt = 0;
    size_t smem = 0;
};
static uint32_t pow2_floor_u32(uint64_t x) { uint32_t r = 1; while (2ULL * r <= x) r *= 2; return r; }
// per-CTA capacity (elements per MLE) and cluster-wide capacity for n_slots MLEs
static uint32_t tail_nloc_cap(const cg_ctx* c, size_t n_slots) {
    const uint64_t cap = (c->max_smem_optin - 8192) / (n_slots * sizeof(ext_t));
    return cap < 2 ? 0 : std::min<uint32_t>(CG_CT_MAX_NLOC, pow2_floor_u32(cap));
}
// Concurrent lanes (ChipScheduler, up to 8 chip proofs at once) share the SMs: a persistent tail that waits for its host
// transcript on 16 SMs per lane, or a cooperative mid kernel that needs the whole chip to itself, starves the other lanes'
// streaming and hashing kernels (measured: 6 chips on 8 lanes took LONGER than sequentially).  So the cluster shrinks with
// the number of sumchecks alive on the context, and the mid kernel is left out when there is more than one.
static uint32_t lane_cluster_cap(const cg_ctx* c) {
    const int live = c->live_sc.load();
    // measured on the 6-chip shard: clusters of 4-8 CTAs (each needing a whole SM of one GPC at the same instant) wait behind
    // the other lanes' streaming waves — 4 lanes took 120 ms against 61 ms sequentially; pairs are scheduled readily
    return live <= 1 ? CG_CT_MAX_C : 2u;
}
static TailPlan tail_plan(const cg_ctx* c, size_t n_slots, uint64_t n0, bool sharded) {
    TailPlan p;
    const uint32_t cap = tail_nloc_cap(c, n_slots);
    if (n0 < 1 || cap < 2 || n_slots > CG_COMM_GATHER_SLOTS) return p;
    if (sharded && n_slots * n0 > CG_GATHER_EXT) return p;
    const char* fe = getenv("CG_TAIL_CLUSTER");   // testing: cap the cluster size (read per call so one process can sweep it)
    const int force_c = fe ? atoi(fe) : 0;
    const uint32_t max_c = std::min(lane_cluster_cap(c), force_c > 0 ? std::min<uint32_t>((uint32_t)force_c, c->tail_max_c) : c->tail_max_c);
    if (n0 <= std::min<uint32_t>(cap, CG_TAIL_START_N) || max_c == 1) {   // one CTA
        if (n0 > cap) return p;
        p.C = 1; p.n_loc0 = (uint32_t)n0; p.nt = (uint32_t)n0;
    } else {
        uint64_t C = n0 / cap;
        if (C < 2) C = 2;
        if (C > max_c) return p;
        p.C = (uint32_t)C;
        p.n_loc0 = (uint32_t)(n0 / C);
        p.nt = std::min<uint32_t>(p.n_loc0 / 2, CG_CT_GATHER_N);   // stay distributed while a CTA still has work for its warps
        if (p.nt < p.C || p.n_loc0 < 4) return p;
    }
    p.smem = n_slots * (size_t)p.n_loc0 * sizeof(ext_t);
    p.ok = true;
    return p;
}
// largest entry size (elements per MLE, cluster-wide, after the entry fold) the tail accepts for n_slots MLEs
static uint64_t tail_cap_n0(const cg_ctx* c, size_t n_slots, bool sharded) {
    const char* fe = getenv("CG_TAIL_CLUSTER");
    const uint32_t max_c = std::min(lane_cluster_cap(c), fe && atoi(fe) > 0 ? std::min<uint32_t>((uint32_t)atoi(fe), c->tail_max_c) : c->tail_max_c);
    uint64_t n = (uint64_t)tail_nloc_cap(c, n_slots) * max_c;
    if (max_c == 1) n = std::min<uint64_t>(n, CG_TAIL_START_N);
    if (sharded) while (n > 1 && n_slots * n > CG_GATHER_EXT) n >>= 1;
    return n;
}
static uint64_t tail_entry_n0(const cg_sumcheck* sc) {
    const uint64_t cur = 1ULL << (sc->num_vars - sc->folds);
    return (sc->pending ? cur / 2 : cur) << sc->extra_rounds;       // sharded: the gathered (global) array
}
static bool tail_eligible(const cg_sumcheck* sc) {
    if (sc->veq.split) return false;
    if (!sc->tl.on || (sc->flags & (CG_SC_NO_FUSE | CG_SC_NO_TAIL)) || sc->round >= sc->num_vars) return false;
    const uint64_t cur = 1ULL << (sc->num_vars - sc->folds);
    const uint64_t n_loc = sc->pending ? cur / 2 : cur;
    if (n_loc < 2) return false;
    if (sc->folds == 0)                                // virtual tower leaves can only be read by the streaming kernels
        for (const VirtLeaf* v : sc->virt) if (v) return false;
    const bool sharded = sc->comm && sc->comm->nranks > 1;
    if (sharded && !sc->extra_rounds) return false;   // step API with a comm: per-round exchange only
    const size_t n_slots = 1 + sc->tl.prod.size() + sc->tl.lk.size();
    return tail_plan(sc->ctx, n_slots, tail_entry_n0(sc), sharded).ok;
}
// launches the tail for rounds sc->round .. num_vars-1; d_tr_state == nullptr -> host mailbox
static int launch_tail(cg_sumcheck* sc, uint64_t* d_tr_state, ext_t* d_msgs, ext_t* d_chal) {
    cg_ctx* c = sc->ctx;
    const TowerLayout& tl = sc->tl;
    CTailArgs ca;
    memset(&ca, 0, sizeof(ca));
    TailArgs& a = ca.t;
    const uint32_t f = sc->folds;
    auto in = [&](uint32_t i) { return (const ext_t*)mle_buf(sc, i, f); };
    a.t.eq_in = in(tl.eq);
    a.t.n_prod = (int)tl.prod_alpha.size();
    a.t.n_logup = (int)tl.lk_an.size();
    int slot = 0;
    a.final_idx[slot++] = (uint16_t)tl.eq;
    for (int p = 0; p < a.t.n_prod; p++) {
        for (int z = 0; z < 2; z++) { a.t.prod_in[p][z] = in(tl.prod[2 * p + z]); a.final_idx[slot++] = (uint16_t)tl.prod[2 * p + z]; }
        a.t.alpha_prod[p] = tl.prod_alpha[p];
    }
    for (int l = 0; l < a.t.n_logup; l++) {
        for (int z = 0; z < 4; z++) { a.t.lk_in[l][z] = in(tl.lk[4 * l + z]); a.final_idx[slot++] = (uint16_t)tl.lk[4 * l + z]; }
        a.t.alpha_num[l] = tl.lk_an[l];
        a.t.alpha_den[l] = tl.lk_ad[l];
    }
    a.t.alpha_one = tl.alpha_one ? 1 : 0;
    a.t.r = sc->pending_r;
    a.t.r_ptr = sc->pending_r_ptr;
    a.entry_fold = sc->pending ? 1 : 0;
    a.canon = (f == 0) ? 1 : 0;
    a.n0 = (uint32_t)tail_entry_n0(sc);   // sharded: size after the entry all-gather
    a.first_round = sc->round;
    a.num_rounds = sc->num_vars + sc->extra_rounds;
    a.d_msgs = d_msgs;
    a.d_chal = d_chal;
    a.d_final = sc->d_final;
    a.d_tr_state = d_tr_state;
    a.mail = d_tr_state ? nullptr : sc_mailbox(sc);
    a.d_error = sc->d_error;
    a.timeout_cycles = c->wait_timeout_cycles;   // default ~4 s: a dead host must not hang the GPU
    const bool sharded = sc->comm && sc->comm->nranks > 1 && sc->extra_rounds;
    if (sharded) {   // sharded prove: all-gather on entry, then everything replicated
        comm_dev(sc->comm, a.comm, 1);
        a.gather_par = (int)(sc->comm->gather_calls++ & 1);
        a.gather_seq = a.comm.seq;
        for (int p = 0; p < sc->comm->nranks; p++) {
            ca.gbuf[p] = &sc->comm->peers[p]->gather.big[a.gather_par][0];
            ca.gflag[p] = &sc->comm->peers[p]->gather.big_seq[a.gather_par][0];
        }
        sc->extra_done = true;
    }
    const TailPlan tp = tail_plan(c, (size_t)slot, a.n0, sharded);
    if (!tp.ok) return set_err(c, CG_ERR_STATE, "launch_tail: layer does not fit the tail kernel");
    ca.n_loc0 = tp.n_loc0;
    ca.nt = tp.nt;
    static long long* d_dbg = nullptr;
    if (getenv("CG_TAIL_DEBUG") && d_tr_state) {   // device challenger only: the dump below synchronises the stream, and a kernel
                                                   // that waits for the host's answer would never get it
        if (!d_dbg) cudaMalloc((void**)&d_dbg, 64 * 8 * sizeof(long long));
        cudaMemsetAsync(d_dbg, 0, 64 * 8 * sizeof(long long), sc->stream);
        ca.dbg = d_dbg;
    }
    const bool simple = a.t.n_prod == 1 && a.t.n_logup == 0 && a.t.alpha_one;
    // the dynamic shared-memory opt-in of both instantiations is raised once per context (prepare_kernels): a per-launch
    // cudaFuncSetAttribute is process-global state and races between lanes
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(tp.C);
    cfg.blockDim = dim3(CG_CT_THREADS);
    cfg.dynamicSmemBytes = tp.smem;
    cfg.stream = sc->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = tp.C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (simple) CU(c, cudaLaunchKernelEx(&cfg, tower_ctail_kernel<true>, ca));
    else CU(c, cudaLaunchKernelEx(&cfg, tower_ctail_kernel<false>, ca));
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    if (ca.dbg) {   // CG_TAIL_DEBUG: print the phase breakdown of every round (cycles of CTA 0)
        cudaStreamSynchronize(sc->stream);
        std::vector<long long> hdbg(64 * 8);
        cudaMemcpy(hdbg.data(), d_dbg, hdbg.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        const uint32_t nr = a.num_rounds - a.first_round;
        fprintf(stderr, "[ctail] C=%u n0=%u n_loc0=%u nt=%u rounds=%u\n", tp.C, a.n0, tp.n_loc0, tp.nt, nr);
        for (uint32_t q = 0; q < nr && q < 64; q++) {
            const long long* d = &hdbg[q * 8];
            fprintf(stderr, "[ctail] round %2u: eval %6lld  exch %6lld  chal %6lld (reduce %lld transcript %lld)  fold %6lld  | gap %6lld cycles\n", q, d[1] - d[0], d[2] - d[1], d[3] - d[2],
                    d[5] - d[2], d[6] - d[5], d[4] - d[3], q + 1 < nr ? hdbg[(q + 1) * 8] - d[4] : 0LL);
        }
    }
    return CG_OK;
}
static void sc_mark_done(cg_sumcheck* sc) {
    sc->round = sc->num_vars;
    sc->folds = sc->num_vars;
    sc->pending = false;
    sc->evaluated = false;
}

// ---- cooperative mid kernel (rounds between the streaming rounds and the tail)
#define CG_MID_MAX_LOG_PAIRS 17
// first round the tail kernel can take over, given the state at round sc->round
static uint32_t log2_u64(uint64_t x) { uint32_t l = 0; while ((2ULL << l) <= x) l++; return l; }
static uint32_t tail_entry_round(const cg_sumcheck* sc) {
    const uint32_t left = sc->num_vars - sc->folds;                  // log2(elements) of the current state
    const size_t n_slots = 1 + sc->tl.prod.size() + sc->tl.lk.size();
    const bool sharded = sc->comm && sc->comm->nranks > 1;
    const uint32_t cap_log = log2_u64(std::max<uint64_t>(1, tail_cap_n0(sc->ctx, n_slots, sharded)));   // global entry size after the entry fold
    const uint32_t glob = cap_log + 1;                               // ... so the state before it may hold twice that
    const uint32_t lim = glob > sc->extra_rounds ? glob - sc->extra_rounds : 0;
    const uint32_t skip = left > lim ? left - lim : 0;
    const uint32_t jt = sc->round + skip;
    return jt < sc->num_vars ? jt : sc->num_vars;
}
static bool mid_eligible(const cg_sumcheck* sc) {
    if (sc->veq.split) return false;
    if (!sc->tl.on || sc->mid_used || (sc->flags & (CG_SC_NO_FUSE | CG_SC_NO_TAIL | CG_SC_NO_MID))) return false;
    if (sc->ctx->live_sc.load() > 1) return false;   // a cooperative launch needs every SM: not while other lanes are proving
    if (!sc->pending || sc->folds < 1 || sc->round >= sc->num_vars) return false;
    const uint32_t left = sc->num_vars - sc->folds;
    if (left < 3 || left - 2 > CG_MID_MAX_LOG_PAIRS) return false;    // fused round: 2^left elements -> 2^(left-2) pairs
    const size_t n_slots = 1 + sc->tl.prod.size() + sc->tl.lk.size();
    if (n_slots > CG_COMM_GATHER_SLOTS) return false;
    return tail_entry_round(sc) > sc->round;
}
static int launch_mid(cg_sumcheck* sc, uint64_t* d_tr_state, uint32_t* jt_out) {
    cg_ctx* c = sc->ctx;
    const TowerLayout& tl = sc->tl;
    MidArgs a;
    memset(&a, 0, sizeof(a));
    a.t.n_prod = (int)tl.prod_alpha.size();
    a.t.n_logup = (int)tl.lk_an.size();
    int slot = 0;
    auto put = [&](uint32_t mle) { a.buf_odd[slot] = (const ext_t*)mle_buf(sc, mle, 1); a.buf_even[slot] = (const ext_t*)mle_buf(sc, mle, 2); slot++; };
    put(tl.eq);
    for (int p = 0; p < a.t.n_prod; p++) { put(tl.prod[2 * p]); put(tl.prod[2 * p + 1]); a.t.alpha_prod[p] = tl.prod_alpha[p]; }
    for (int l = 0; l < a.t.n_logup; l++) {
        for (int z = 0; z < 4; z++) put(tl.lk[4 * l + z]);
        a.t.alpha_num[l] = tl.lk_an[l];
        a.t.alpha_den[l] = tl.lk_ad[l];
    }
    a.t.alpha_one = tl.alpha_one ? 1 : 0;
    a.t.r = sc->pending_r;
    a.t.r_ptr = sc->pending_r_ptr;
    a.f0 = sc->folds;
    a.log_n_in = sc->num_vars - sc->folds;
    a.first_round = sc->round;
    a.end_round = tail_entry_round(sc);
    a.d_msgs = sc->d_msgs;
    a.d_chal = sc->d_chal;
    a.d_tr_state = d_tr_state;
    a.mail = d_tr_state ? nullptr : sc_mailbox(sc);
    a.d_error = sc->d_error;
    a.timeout_cycles = c->wait_timeout_cycles;
    comm_dev(sc->comm, a.comm, a.end_round - a.first_round);
    a.partials = sc->out.partials;
    a.ticket = sc->d_mid_ticket;
    a.round_flag = sc->d_mid_flag;
    const bool simple = a.t.n_prod == 1 && a.t.n_logup == 0 && a.t.alpha_one;
    const void* fn = simple ? (const void*)tower_mid_kernel<true> : (const void*)tower_mid_kernel<false>;
    int occ = 0;
    CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, CG_THREADS, 0));
    if (occ < 1) return set_err(c, CG_ERR_CUDA, "mid kernel does not fit on an SM");
    const uint64_t first_pairs = 1ULL << (a.log_n_in - 2);
    uint64_t blocks = (first_pairs + CG_THREADS - 1) / CG_THREADS;
    const uint64_t cap = (uint64_t)c->sm_count * (occ > 2 ? 2 : occ);
    if (blocks > cap) blocks = cap;
    void* args[] = {&a};
    CU(c, cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(CG_THREADS), args, 0, sc->stream));
    LAUNCHED(c);
    // bookkeeping: the device now owns rounds [first_round, end_round); the last challenge sits in d_chal
    const uint32_t steps = a.end_round - a.first_round;
    sc->folds += steps;
    sc->round = a.end_round;
    sc->pending = true;
    sc->pending_r_ptr = sc->d_chal + (a.end_round - 1);
    sc->evaluated = false;
    sc->mid_used = true;
    *jt_out = a.end_round;
    return CG_OK;
}

// ---- persistent split-eq rounds (veq_persist_kernel): every claim-derived streaming round in one cooperative launch
static bool persist_eligible(const cg_sumcheck* sc) {
    static const int on = []() { const char* e = getenv("CG_VEQ_PERSIST"); return e ? atoi(e) : 1; }();
    static const int use_tma = []() { const char* e = getenv("CG_VEQ_TMA"); return e ? atoi(e) : 1; }();
    static const int minb = []() { const char* e = getenv("CG_VEQ_MINB"); return e ? atoi(e) : CG_VEQ_MINB_DEFAULT; }();
    const VeqState& v = sc->veq;
    if (!on || !use_tma || minb == 3 || !v.split || v.general || !v.derive_ok) return false;
    if (sc->flags & (CG_SC_NO_FUSE | CG_SC_NO_MID | CG_SC_NO_PERSIST | CG_SC_FORCE_GENERIC)) return false;
    if (!sc->pending || sc->round < 1 || sc->round + 1 >= v.J) return false;   // at least two rounds, after round 0
    if (sc->ctx->live_sc.load() > 1) return false;                              // cooperative: needs the whole chip
    return true;
}
static int launch_veq_persist(cg_sumcheck* sc, uint64_t* d_tr_state, uint32_t* upto_out) {
    cg_ctx* c = sc->ctx;
    const VeqState& v = sc->veq;
    VeqPersistArgs a;
    memset(&a, 0, sizeof(a));
    const uint32_t j0 = sc->round, j1 = v.J, steps = j1 - j0;
    for (uint32_t i = 0; i <= steps; i++) {
        a.bufA[i] = (const ext_t*)mle_buf(sc, sc->tl.prod[0], j0 - 1 + i);
        a.bufB[i] = (const ext_t*)mle_buf(sc, sc->tl.prod[1], j0 - 1 + i);
    }
    a.L = v.d_L;
    a.H = v.d_H;
    for (uint32_t j = 0; j <= v.J; j++) a.h_off[j] = v.h_off[j];
    a.k = sc->num_vars;
    a.j0 = j0;
    a.j1 = j1;
    a.canon_first = (j0 == 1) ? 1 : 0;
    a.r = sc->pending_r;
    a.r_ptr = sc->pending_r_ptr;
    a.fin.w = v.d_w;
    a.fin.inv1mw = v.d_inv1mw;
    a.fin.prefix = v.d_prefix;
    a.fin.qstate = v.d_qstate;
    a.fin.scale = v.scale;
    a.fin.sharded = (sc->comm && sc->comm->nranks > 1) ? 1 : 0;
    a.d_msgs = sc->d_msgs;
    a.d_chal = sc->d_chal;
    a.d_tr_state = d_tr_state;
    a.mail = d_tr_state ? nullptr : sc_mailbox(sc);
    a.d_error = sc->d_error;
    a.timeout_cycles = c->wait_timeout_cycles;
    comm_dev(sc->comm, a.comm, steps);
    a.partials = sc->out.partials;
    a.ticket = sc->d_per_ticket;
    a.round_flag = sc->d_per_flag;
    const uint64_t first_rows = 1ULL << (sc->num_vars - j0 - 1 - CG_VEQ_LO_BITS);
    uint64_t blocks = (uint64_t)c->sm_count * 2;
    if (blocks > first_rows) blocks = first_rows;
    if (blocks > CG_MAX_BLOCKS) blocks = CG_MAX_BLOCKS;
    void* args[] = {&a};
    CU(c, cudaLaunchCooperativeKernel((const void*)veq_persist_kernel, dim3((unsigned)blocks), dim3(256), args, VeqTmaCfg<true, 2>::SMEM, sc->stream));
    LAUNCHED(c);
    sc->folds += steps;
    sc->round = j1;
    sc->pending = true;
    sc->pending_r_ptr = sc->d_chal + (j1 - 1);
    sc->evaluated = false;
    *upto_out = j1;
    return CG_OK;
}

static void prof_begin(cg_sumcheck* sc) {
    if (!(sc->flags & CG_SC_PROFILE)) return;
    sc->ev.resize(2 * (size_t)sc->num_vars);
    for (auto& e : sc->ev) cudaEventCreate(&e);
}
static void prof_mark(cg_sumcheck* sc, uint32_t j, int end) {
    if (sc->flags & CG_SC_PROFILE) cudaEventRecord(sc->ev[2 * j + end], sc->stream);
}
static void prof_end(cg_sumcheck* sc) {
    if (!(sc->flags & CG_SC_PROFILE)) return;
    cudaStreamSynchronize(sc->stream);
    std::lock_guard<std::mutex> g(sc->ctx->mu);
    if (!sc->profile_append) sc->ctx->profile_ms.clear();
    for (uint32_t j = 0; j < sc->num_vars; j++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, sc->ev[2 * j], sc->ev[2 * j + 1]);
        sc->ctx->profile_ms.push_back(ms);
    }
    for (auto& e : sc->ev) cudaEventDestroy(e);
    sc->ev.clear();
}
CG_EXPORT int cg_profile_last(cg_ctx* c, float* ms_out, uint32_t cap, uint32_t* n) {
    if (!c || !n) return CG_ERR_INVALID;
    std::lock_guard<std::mutex> g(c->mu);
    *n = (uint32_t)c->profile_ms.size();
    for (uint32_t i = 0; i < *n && i < cap && ms_out; i++) ms_out[i] = c->profile_ms[i];
    return CG_OK;
}

static int sc_run_host(cg_sumcheck* sc, cg_challenge_cb cb, void* user, uint64_t* h_rounds, uint64_t* h_final, uint64_t* h_chal, uint32_t round_base = 0) {
    bool have_final = false;
    if (!sc->ctx->host_wait_ok) sc->flags |= CG_SC_NO_TAIL | CG_SC_NO_MID;   // launches block the host (profiler): one launch per round
    prof_begin(sc);
    for (uint32_t j = 0; j < sc->num_vars; j++) {
        uint64_t* msg = h_rounds + (size_t)j * sc->degree * 2;
        prof_mark(sc, j, 0);
        CHK(veq_prepare(sc));
        if (persist_eligible(sc) || mid_eligible(sc) || tail_eligible(sc)) {
            // the device runs every remaining round on its own (persistent split-eq rounds, cooperative mid kernel, then the
            // shared-memory tail kernel, all enqueued now); the transcript stays on the host and answers through the mailbox
            cg_ctx* c = sc->ctx;
            TailMailbox* mb = sc_mailbox(sc);
            if (j == 0) mailbox_reset(mb);
            uint32_t upto = j;   // rounds [j, upto) are owned by enqueued kernels
            bool done = false;
            if (persist_eligible(sc)) {
                CHK(launch_veq_persist(sc, nullptr, &upto));
                CHK(veq_prepare(sc));   // leave split mode: the eq state is materialised behind the persistent kernel
            }
            if (mid_eligible(sc)) CHK(launch_mid(sc, nullptr, &upto));
            if (tail_eligible(sc)) {
                CHK(launch_tail(sc, nullptr, sc->d_msgs, sc->d_chal));
                upto = sc->num_vars + sc->extra_rounds;
                done = true;
            }
            int rc = CG_OK;
            for (uint32_t jj = j; jj < upto && rc == CG_OK; jj++) {
                uint64_t spins = 0;
                uint64_t* m = h_rounds + (size_t)jj * sc->degree * 2;
                while (!mailbox_take(mb, 2 * sc->degree, (uint64_t)jj + 1, m)) {
                    if ((++spins & 0xFFFFF) == 0) {
                        const cudaError_t q = cudaStreamQuery(sc->stream);
                        if (q == cudaErrorNotReady) continue;
                        // the stream has drained: the message must be there; re-read for a short grace period before giving up
                        bool seen = false;
                        for (int grace = 0; grace < 2000000 && !seen; grace++) { __sync_synchronize(); seen = mailbox_take(mb, 2 * sc->degree, (uint64_t)jj + 1, m); }
                        if (seen) break;
                        rc = set_err(c, CG_ERR_CUDA, std::string("persistent kernel ended before posting its round message (round ") + std::to_string(jj) +
                                                         " of " + std::to_string(sc->num_vars) + ", owned up to " + std::to_string(upto) + ", seq_r=" +
                                                         std::to_string((unsigned long long)mb->seq_r) + " abort=" + std::to_string((int)mb->abort) +
                                                         ", stream: " + cudaGetErrorString(q) + ")");
                        break;
                    }
                }
                if (rc != CG_OK) break;
                uint64_t r[2] = {0, 0};
                cb(user, round_base + jj, m, sc->degree, r);
                if (h_chal) { h_chal[2 * jj] = r[0]; h_chal[2 * jj + 1] = r[1]; }
                mailbox_reply(mb, r, (uint64_t)jj + 1);
            }
            if (rc != CG_OK) { mb->abort = 1; __sync_synchronize(); }
            prof_mark(sc, j, 1);
            for (uint32_t jj = j + 1; jj < sc->num_vars && jj < upto; jj++) { prof_mark(sc, jj, 0); prof_mark(sc, jj, 1); }
            if (done && rc == CG_OK && sc->n_mles <= CG_COMM_GATHER_SLOTS && !(sc->flags & CG_SC_PROFILE)) {
                // the tail kernel posts the final evaluations into the mailbox: no stream synchronisation, no D2H copy
                // (the sumcheck's buffers are released in stream order, so the kernel may still be exiting)
                uint64_t fin_tmp[2 * CG_COMM_GATHER_SLOTS];
                uint64_t spins = 0;
                bool got = false;
                const uint64_t fseq = (uint64_t)sc->num_vars + sc->extra_rounds + 1;
                while (!(got = mailbox_take_final(mb, sc->n_mles, fseq, fin_tmp))) {
                    if ((++spins & 0xFFFFF) == 0 && cudaStreamQuery(sc->stream) != cudaErrorNotReady) {
                        for (int grace = 0; grace < 2000000 && !got; grace++) { __sync_synchronize(); got = mailbox_take_final(mb, sc->n_mles, fseq, fin_tmp); }
                        break;
                    }
                }
                if (got) {
                    if (h_final) memcpy(h_final, fin_tmp, sizeof(ext_t) * sc->n_mles);
                    sc_mark_done(sc);
                    have_final = true;
                    break;
                }
            }
            if (done || rc != CG_OK) {
                CU(c, cudaStreamSynchronize(sc->stream));
                CHK(rc);
                sc_mark_done(sc);
                break;
            }
            j = upto - 1;   // tail not possible (flags / shape): continue with per-round launches from round `upto`
            continue;
        }
        {   // one launch; its last block posts the message into the host mailbox (no D2H copy + stream sync)
            cg_ctx* c = sc->ctx;
            TailMailbox* mb = sc_mailbox(sc);
            if (j == 0) mailbox_reset(mb);
            RoundOut ro = make_ro(sc);
            ro.d_out = sc->d_msgs;
            ro.d_tr_state = nullptr;
            ro.d_r_out = nullptr;
            ro.mail = mb;
            ro.mail_seq = (uint64_t)j + 1;
            CHK(sc_enqueue_round(sc, ro));
            prof_mark(sc, j, 1);
            uint64_t spins = 0;
            while (!mailbox_take(mb, 2 * sc->degree, (uint64_t)j + 1, msg)) {
                if ((++spins & 0xFFFFF) == 0) {
                    cudaError_t q = cudaStreamQuery(sc->stream);
                    if (q != cudaErrorNotReady) {
                        bool seen = false;
                        for (int grace = 0; grace < 2000000 && !seen; grace++) { __sync_synchronize(); seen = mailbox_take(mb, 2 * sc->degree, (uint64_t)j + 1, msg); }
                        if (!seen) return set_err(c, CG_ERR_CUDA, std::string("round kernel ended without posting its message: ") + cudaGetErrorString(q));
                        break;
                    }
                }
            }
        }
        uint64_t r[2] = {0, 0};
        cb(user, round_base + j, msg, sc->degree, r);
        if (h_chal) { h_chal[2 * j] = r[0]; h_chal[2 * j + 1] = r[1]; }
        CHK(cg_sumcheck_bind(sc, r));
    }
    prof_end(sc);
    if (h_final && !have_final) CHK(cg_sumcheck_final_evals(sc, h_final));
    return CG_OK;
}

// ---- mixed sizes: the cross-chip batched main sumcheck (prove_batched_main_constraints,
// ceno_zkvm/src/scheme/cpu/mod.rs:1052-1390; GPU call site ceno_zkvm/src/scheme/gpu/mod.rs:2968-2981).
// An MLE f with k' < k variables stands for F(x) = f(x_0..x_{k'-1}) * prod_{j>=k'} x_j ("frontload"): pinned by the
// verifier's final claim (ceno_zkvm/src/scheme/verifier.rs:180-238) as restated in
// ceno_recursion_v2/src/main/mod.rs:3414-3448.  Every term lives in one chip, i.e. its factors share one k'.  So
//   rounds j <  k': the term's round polynomial is the chip's own sumcheck round over its 2^(k'-j-1) pairs;
//   rounds j >= k': it is  scalar * prod_i c_i * X^(#factors),  c_i = f_i(r_<k') * prod_{k'<=l<j} r_l.
// The library runs one uniform sub-sumcheck per size class in lockstep (all device kernels as usual) and adds the
// closed-form contributions of the collapsed classes on the host (a handful of field multiplications per round,
// next to the transcript).  Reported final evaluations are the raw f_i(r_<k') (cpu/mod.rs:1346-1358).
static bool mles_mixed(const cg_mle_desc* mles, uint32_t n_mles, uint32_t num_vars) {
    for (uint32_t i = 0; i < n_mles; i++) if (mles && mles[i].num_vars != num_vars) return true;
    return false;
}
static int sc_prove_mixed(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, const uint64_t* coeff, const uint32_t* off,
                          const uint32_t* idx, uint32_t n_terms, uint32_t num_vars, uint32_t degree, uint32_t flags,
                          cg_challenge_cb cb, void* user, uint64_t* h_rounds, uint64_t* h_final, uint64_t* h_chal, cg_stream s) {
    if (degree == 0 || degree > CG_MAX_DEGREE) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_prove: degree must be in 1..8");
    struct Class {
        uint32_t kv = 0;
        std::vector<uint32_t> mle_ids;                  // global MLE index of local MLE l
        std::vector<cg_mle_desc> descs;
        std::vector<uint64_t> t_coeff;
        std::vector<uint32_t> t_off{0}, t_idx;          // local term tables
        cg_sumcheck* sc = nullptr;
        std::vector<ext_t> cval;                        // collapsed: c_i per local MLE
        bool collapsed = false;
    };
    std::map<uint32_t, Class> classes;
    std::vector<uint32_t> local_of(n_mles, 0);
    for (uint32_t i = 0; i < n_mles; i++) {
        if (mles[i].num_vars > num_vars) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_prove: an MLE has more variables than the sumcheck");
        if (mles[i].is_ext == CG_MLE_EQ) return set_err(c, CG_ERR_UNSUPPORTED, "mixed-size sumcheck: virtual eq MLEs are not supported");
        Class& cl = classes[mles[i].num_vars];
        cl.kv = mles[i].num_vars;
        local_of[i] = (uint32_t)cl.mle_ids.size();
        cl.mle_ids.push_back(i);
        cl.descs.push_back(mles[i]);
    }
    for (uint32_t t = 0; t < n_terms; t++) {
        if (off[t + 1] <= off[t]) return set_err(c, CG_ERR_UNSUPPORTED, "mixed-size sumcheck: constant terms are not supported");
        if (off[t + 1] - off[t] > degree) return set_err(c, CG_ERR_INVALID, "term has more factors than `degree`");
        for (uint32_t q = off[t]; q < off[t + 1]; q++)
            if (idx[q] >= n_mles) return set_err(c, CG_ERR_INVALID, "term references an MLE index out of range");
        const uint32_t kv = mles[idx[off[t]]].num_vars;
        for (uint32_t q = off[t]; q < off[t + 1]; q++)
            if (mles[idx[q]].num_vars != kv)
                return set_err(c, CG_ERR_UNSUPPORTED, "mixed-size sumcheck: the factors of a term must share one num_vars (one chip)");
        Class& cl = classes[kv];
        cl.t_coeff.push_back(coeff[2 * t] % GL_P);
        cl.t_coeff.push_back(coeff[2 * t + 1] % GL_P);
        for (uint32_t q = off[t]; q < off[t + 1]; q++) cl.t_idx.push_back(local_of[idx[q]]);
        cl.t_off.push_back((uint32_t)cl.t_idx.size());
    }
    int rc = CG_OK;
    auto cleanup = [&]() { for (auto& kv : classes) if (kv.second.sc) cg_sumcheck_destroy(kv.second.sc); };
    for (auto& kv : classes) {
        Class& cl = kv.second;
        const uint32_t nt = (uint32_t)cl.t_off.size() - 1;
        rc = sc_create_terms(c, cl.descs.data(), (uint32_t)cl.descs.size(), nt ? cl.t_coeff.data() : nullptr, nt ? cl.t_off.data() : nullptr,
                             nt ? cl.t_idx.data() : nullptr, nt, cl.kv, degree, flags & ~CG_SC_PROFILE, s, nullptr, &cl.sc);
        if (rc != CG_OK) { cleanup(); return rc; }
    }
    std::vector<uint64_t> tmp(2 * (size_t)degree), fin(2 * (size_t)(n_mles ? n_mles : 1));
    auto collapse = [&](Class& cl) -> int {   // all of the class's variables are bound: fetch the raw evaluations
        std::vector<uint64_t> f(2 * cl.descs.size());
        CHK(cg_sumcheck_final_evals(cl.sc, f.data()));
        cl.cval.resize(cl.descs.size());
        for (size_t l = 0; l < cl.descs.size(); l++) {
            cl.cval[l] = ext_t{f[2 * l], f[2 * l + 1]};
            fin[2 * cl.mle_ids[l]] = f[2 * l];
            fin[2 * cl.mle_ids[l] + 1] = f[2 * l + 1];
        }
        cl.collapsed = true;
        return CG_OK;
    };
    for (auto& kv : classes) if (kv.second.kv == 0 && rc == CG_OK) rc = collapse(kv.second);
    for (uint32_t j = 0; j < num_vars && rc == CG_OK; j++) {
        std::vector<ext_t> msg(degree, ext_t{0, 0});
        for (auto& kv : classes) {
            Class& cl = kv.second;
            if (!cl.collapsed) {
                rc = cg_sumcheck_round_eval(cl.sc, tmp.data());
                if (rc != CG_OK) break;
                for (uint32_t x = 0; x < degree; x++) msg[x] = hx_add(msg[x], ext_t{tmp[2 * x], tmp[2 * x + 1]});
            } else {
                const uint32_t nt = (uint32_t)cl.t_off.size() - 1;
                for (uint32_t t = 0; t < nt; t++) {
                    ext_t v{cl.t_coeff[2 * t], cl.t_coeff[2 * t + 1]};
                    const uint32_t d = cl.t_off[t + 1] - cl.t_off[t];
                    for (uint32_t q = cl.t_off[t]; q < cl.t_off[t + 1]; q++) v = hx_mul(v, cl.cval[cl.t_idx[q]]);
                    for (uint32_t x = 0; x < degree; x++) {
                        uint64_t pw = 1;
                        for (uint32_t e = 0; e < d; e++) pw = hx_mulmod(pw, x + 1);
                        msg[x] = hx_add(msg[x], ext_t{hx_mulmod(v.c0, pw), hx_mulmod(v.c1, pw)});
                    }
                }
            }
        }
        if (rc != CG_OK) break;
        uint64_t* m = h_rounds + (size_t)j * degree * 2;
        for (uint32_t x = 0; x < degree; x++) { m[2 * x] = msg[x].c0; m[2 * x + 1] = msg[x].c1; }
        uint64_t r[2] = {0, 0};
        cb(user, j, m, degree, r);
        if (h_chal) { h_chal[2 * j] = r[0]; h_chal[2 * j + 1] = r[1]; }
        const ext_t re{r[0] % GL_P, r[1] % GL_P};
        for (auto& kv : classes) {
            Class& cl = kv.second;
            if (cl.collapsed) { for (auto& cv : cl.cval) cv = hx_mul(cv, re); continue; }
            rc = cg_sumcheck_bind(cl.sc, r);
            if (rc == CG_OK && cl.kv == j + 1) rc = collapse(cl);
            if (rc != CG_OK) break;
        }
    }
    if (rc == CG_OK && h_final) memcpy(h_final, fin.data(), sizeof(uint64_t) * 2 * n_mles);
    cleanup();
    return rc;
}

CG_EXPORT int cg_sumcheck_prove(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, const uint64_t* coeff,
                                const uint32_t* off, const uint32_t* idx, uint32_t n_terms, uint32_t num_vars,
                                uint32_t degree, uint32_t flags, cg_challenge_cb cb, void* user, uint64_t* h_rounds,
                                uint64_t* h_final, uint64_t* h_chal, cg_stream s) {
    if (!cb || (!h_rounds && num_vars)) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_prove: null callback/output");
    if (c && mles_mixed(mles, n_mles, num_vars))
        return sc_prove_mixed(c, mles, n_mles, coeff, off, idx, n_terms, num_vars, degree, flags, cb, user, h_rounds, h_final, h_chal, s);
    cg_sumcheck* sc = nullptr;
    CHK(cg_sumcheck_create(c, mles, n_mles, coeff, off, idx, n_terms, num_vars, degree, flags, s, &sc));
    int rc = sc_run_host(sc, cb, user, h_rounds, h_final, h_chal);
    cg_sumcheck_destroy(sc);
    return rc;
}

// device-resident challenger: enqueue every round back to back, read everything once at the end
static int sc_run_device(cg_sumcheck* sc, uint64_t* h_state, uint64_t* h_rounds, uint64_t* h_final, uint64_t* h_chal) {
    cg_ctx* c = sc->ctx;
    void* p = nullptr;
    CHK(sc_alloc(sc, 256, &p));
    sc->d_tr_state = (uint64_t*)p;
    CU(c, cudaMemcpyAsync(sc->d_tr_state, h_state, 8, cudaMemcpyHostToDevice, sc->stream));
    prof_begin(sc);
    for (uint32_t j = 0; j < sc->num_vars; j++) {
        prof_mark(sc, j, 0);
        CHK(veq_prepare(sc));
        if (persist_eligible(sc)) {   // every remaining streaming round of the split-eq form in one cooperative launch
            uint32_t upto = j;
            CHK(launch_veq_persist(sc, sc->d_tr_state, &upto));
            prof_mark(sc, j, 1);
            for (uint32_t jj = j + 1; jj < upto; jj++) { prof_mark(sc, jj, 0); prof_mark(sc, jj, 1); }
            j = upto - 1;
            continue;
        }
        if (mid_eligible(sc)) {    // one cooperative launch for the latency-bound rounds before the tail
            uint32_t upto = j;
            CHK(launch_mid(sc, sc->d_tr_state, &upto));
            prof_mark(sc, j, 1);
            for (uint32_t jj = j + 1; jj < upto; jj++) { prof_mark(sc, jj, 0); prof_mark(sc, jj, 1); }
            j = upto - 1;
            continue;
        }
        if (tail_eligible(sc)) {   // one persistent launch for every remaining round
            CHK(launch_tail(sc, sc->d_tr_state, sc->d_msgs, sc->d_chal));
            prof_mark(sc, j, 1);
            for (uint32_t jj = j + 1; jj < sc->num_vars; jj++) { prof_mark(sc, jj, 0); prof_mark(sc, jj, 1); }
            sc_mark_done(sc);
            break;
        }
        RoundOut ro = make_ro(sc);
        ro.d_out = sc->d_msgs + (size_t)j * sc->degree;
        ro.d_tr_state = sc->d_tr_state;
        ro.d_r_out = sc->d_chal + j;
        CHK(sc_enqueue_round(sc, ro));
        prof_mark(sc, j, 1);
        CHK(sc_apply_pending_fold_only(sc));
        sc->pending = true;
        sc->pending_r_ptr = sc->d_chal + j;
        CHK(sc_bind_common(sc));
    }
    if (sc->num_vars) {
        const uint32_t nr = sc->num_vars + (sc->extra_done ? sc->extra_rounds : 0);
        CU(c, cudaMemcpyAsync(h_rounds, sc->d_msgs, sizeof(ext_t) * nr * sc->degree, cudaMemcpyDeviceToHost, sc->stream));
        if (h_chal) CU(c, cudaMemcpyAsync(h_chal, sc->d_chal, sizeof(ext_t) * nr, cudaMemcpyDeviceToHost, sc->stream));
    }
    CU(c, cudaMemcpyAsync(h_state, sc->d_tr_state, 8, cudaMemcpyDeviceToHost, sc->stream));
    CU(c, cudaStreamSynchronize(sc->stream));
    prof_end(sc);
    if (h_final) CHK(cg_sumcheck_final_evals(sc, h_final));
    return CG_OK;
}
CG_EXPORT int cg_sumcheck_prove_standin_device(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, const uint64_t* coeff,
                                               const uint32_t* off, const uint32_t* idx, uint32_t n_terms, uint32_t num_vars,
                                               uint32_t degree, uint32_t flags, uint64_t* h_state, uint64_t* h_rounds,
                                               uint64_t* h_final, uint64_t* h_chal, cg_stream s) {
    if (!h_state || (!h_rounds && num_vars)) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_prove_standin_device: null output");
    if (c && mles_mixed(mles, n_mles, num_vars))   // mixed sizes: the class lockstep runs on the host; same stand-in sponge, same proof
        return sc_prove_mixed(c, mles, n_mles, coeff, off, idx, n_terms, num_vars, degree, flags, cg_standin_challenge_cb, h_state,
                              h_rounds, h_final, h_chal, s);
    cg_sumcheck* sc = nullptr;
    CHK(cg_sumcheck_create(c, mles, n_mles, coeff, off, idx, n_terms, num_vars, degree, flags, s, &sc));
    int rc = sc_run_device(sc, h_state, h_rounds, h_final, h_chal);
    cg_sumcheck_destroy(sc);
    return rc;
}


// the rounds of a sharded sumcheck whose state `sc` was created with (comm, extra_rounds = log2 nranks); consumes sc
static int sc_run_sharded(cg_ctx* c, cg_comm* cm, cg_sumcheck* sc, uint32_t n_mles, const uint64_t* coeff, const uint32_t* off, const uint32_t* idx,
                          uint32_t n_terms, uint32_t num_vars_global, uint32_t degree, uint32_t flags, cg_challenge_cb cb, void* user,
                          uint64_t* h_standin_state, uint64_t* h_rounds, uint64_t* h_final, uint64_t* h_chal, cg_stream s) {
    int g = 0;
    while ((1 << g) < cm->nranks) g++;
    const uint32_t k_local = num_vars_global - g;
    cudaStream_t st = S(c, s);
    std::vector<uint64_t> fin_local(2 * (size_t)(n_mles ? n_mles : 1));
    int rc = h_standin_state ? sc_run_device(sc, h_standin_state, h_rounds, fin_local.data(), h_chal)
                             : sc_run_host(sc, cb, user, h_rounds, fin_local.data(), h_chal);
    if (rc == CG_OK) rc = comm_check(cm, st);
    if (rc != CG_OK || g == 0 || sc->extra_done) {
        if (rc == CG_OK && h_final) memcpy(h_final, fin_local.data(), sizeof(uint64_t) * 2 * n_mles);
        cg_sumcheck_destroy(sc);
        return rc;
    }
    // all-gather the m final local evaluations (they sit in sc->d_final) into every mailbox
    const int par = (int)(cm->gather_calls++ & 1);
    CommDev cd;
    comm_dev(cm, cd, 1);
    comm_allgather_kernel<<<1, 64, 0, st>>>(sc->d_final, (int)n_mles, cd, par);
    LAUNCHED(c);
    if (cudaGetLastError() != cudaSuccess) rc = set_err(c, CG_ERR_CUDA, "comm_allgather_kernel launch failed");
    if (k_local == 0 && rc == CG_OK) {
        // zero local rounds: d_final was never written by a fold; the single local values are the inputs
        rc = set_err(c, CG_ERR_UNSUPPORTED, "sharded prove needs at least one local variable per rank");
    }
    cg_sumcheck* sc2 = nullptr;
    if (rc == CG_OK) {
        std::vector<cg_mle_desc> gd(n_mles);
        for (uint32_t i = 0; i < n_mles; i++) gd[i] = cg_mle_desc{&cm->mine->gather.v[par][i][0], (uint64_t)cm->nranks, (uint32_t)g, 1u};
        rc = cg_sumcheck_create(c, gd.data(), n_mles, coeff, off, idx, n_terms, (uint32_t)g, degree, flags, s, &sc2);
        if (rc == CG_OK) sc2->profile_append = true;
    }
    if (rc == CG_OK) {
        uint64_t* r2 = h_rounds + (size_t)k_local * degree * 2;
        uint64_t* c2 = h_chal ? h_chal + (size_t)k_local * 2 : nullptr;
        rc = h_standin_state ? sc_run_device(sc2, h_standin_state, r2, h_final, c2) : sc_run_host(sc2, cb, user, r2, h_final, c2, k_local);
    }
    if (rc == CG_OK) rc = comm_check(cm, st);
    if (sc2) cg_sumcheck_destroy(sc2);
    cg_sumcheck_destroy(sc);
    return rc;
}

// Sharded prove: this rank holds the slice [rank 2^(k-g), (rank+1) 2^(k-g)) of every MLE
// (descs carry num_vars = k - g).  Rounds 0..k-g-1 run locally with the in-kernel NVLink exchange of
// the partial sums; the final local evaluations are all-gathered into every rank's mailbox and the
// last g rounds run replicated on the N gathered elements.  Outputs are identical on every rank and
// identical to the single-device proof.  h_standin_state != NULL selects the device challenger.
CG_EXPORT int cg_sumcheck_prove_sharded(cg_ctx* c, cg_comm* cm, const cg_mle_desc* mles, uint32_t n_mles, const uint64_t* coeff,
                                        const uint32_t* off, const uint32_t* idx, uint32_t n_terms, uint32_t num_vars_global,
                                        uint32_t degree, uint32_t flags, cg_challenge_cb cb, void* user, uint64_t* h_standin_state,
                                        uint64_t* h_rounds, uint64_t* h_final, uint64_t* h_chal, cg_stream s) {
    if (!c || !cm || (!cb && !h_standin_state) || !h_rounds) return set_err(c, CG_ERR_INVALID, "cg_sumcheck_prove_sharded: null argument");
    int g = 0;
    while ((1 << g) < cm->nranks) g++;
    if (num_vars_global < (uint32_t)g) return set_err(c, CG_ERR_INVALID, "fewer variables than log2(ranks)");
    if (n_mles > CG_COMM_GATHER_MLES) return set_err(c, CG_ERR_UNSUPPORTED, "sharded prove supports at most 64 MLEs");

    const uint32_t k_local = num_vars_global - g;
    cudaStream_t st = S(c, s);
    // a virtual eq MLE carries the GLOBAL point (num_vars_global ext): on this rank's slice eq(w, .) is
    // eq(w_top, rank) * eq(w_low, .), so the local prover gets w_low and the constant factor as its initial prefix
    ext_t veq_scale{1, 0};
    uint32_t n_veq = 0;
    for (uint32_t i = 0; i < n_mles; i++) {
        if (!mles || mles[i].is_ext != CG_MLE_EQ) continue;
        if (++n_veq > 1) return set_err(c, CG_ERR_UNSUPPORTED, "sharded prove: at most one virtual eq MLE");
        if (!mles[i].dptr) return set_err(c, CG_ERR_INVALID, "sharded prove: virtual eq MLE without a point");
        const uint64_t* pt = (const uint64_t*)mles[i].dptr;
        unsigned __int128 P = GL_P;
        uint64_t a0 = 1, a1 = 0;
        for (int b = 0; b < g; b++) {
            uint64_t w0 = pt[2 * (k_local + b)] % GL_P, w1 = pt[2 * (k_local + b) + 1] % GL_P;
            if (!((cm->rank >> b) & 1)) { w0 = (uint64_t)(((unsigned __int128)1 + P - w0) % P); w1 = (uint64_t)((P - w1) % P); }
            const uint64_t n0 = (uint64_t)((((unsigned __int128)a0 * w0) % P + ((((unsigned __int128)a1 * w1) % P) * 7) % P) % P);
            const uint64_t n1 = (uint64_t)((((unsigned __int128)a0 * w1) % P + ((unsigned __int128)a1 * w0) % P) % P);
            a0 = n0; a1 = n1;
        }
        veq_scale = ext_t{a0, a1};
    }
    cg_sumcheck* sc = nullptr;
    // the persistent tail kernel continues through the replicated rounds when it can (extra_rounds = g)
    CHK(sc_create_terms(c, mles, n_mles, coeff, off, idx, n_terms, k_local, degree, flags, s, n_veq ? &veq_scale : nullptr, &sc, cm, (uint32_t)g));
    return sc_run_sharded(c, cm, sc, n_mles, coeff, off, idx, n_terms, num_vars_global, degree, flags, cb, user, h_standin_state, h_rounds, h_final, h_chal, s);
}

// ------------------------------------------------------------------- fold / evaluate helpers
CG_EXPORT int cg_fix_variable(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, const uint64_t r[2],
                              uint64_t* const* d_out, cg_stream s) {
    if (!c || !mles || !r || !d_out) return CG_ERR_INVALID;
    if (n_mles == 0) return CG_OK;
    cudaStream_t st = S(c, s);
    CU(c, cudaSetDevice(c->device));
    const uint32_t nv = mles[0].num_vars;
    if (nv == 0) return set_err(c, CG_ERR_INVALID, "cg_fix_variable: MLE has no variables");
    std::vector<FoldSlot> slots(n_mles);
    for (uint32_t i = 0; i < n_mles; i++) {
        if (mles[i].is_ext > CG_MLE_EXT) return set_err(c, CG_ERR_INVALID, "cg_fix_variable: dense MLEs only");
        if (mles[i].num_vars != nv) return set_err(c, CG_ERR_INVALID, "cg_fix_variable: all MLEs must have the same num_vars");
        if (mles[i].len != (1ULL << nv)) return set_err(c, CG_ERR_UNSUPPORTED, "cg_fix_variable: len must equal 2^num_vars");
        if (((uintptr_t)mles[i].dptr & 31) || ((uintptr_t)d_out[i] & 15)) return set_err(c, CG_ERR_INVALID, "cg_fix_variable: pointers must be 32-byte (in) / 16-byte (out) aligned");
        slots[i] = FoldSlot{mles[i].dptr, (ext_t*)d_out[i], mles[i].is_ext ? 1u : 0u, 1u};
    }
    void* d_slots = nullptr;
    CHK(upload_small(c, slots.data(), sizeof(FoldSlot) * n_mles, &d_slots, st));
    const uint64_t n_out = 1ULL << (nv - 1);
    ext_t rv{r[0] >= GL_P ? r[0] - GL_P : r[0], r[1] >= GL_P ? r[1] - GL_P : r[1]};
    fold_kernel<<<dim3(grid_for(c, n_out), n_mles), CG_THREADS, 0, st>>>((const FoldSlot*)d_slots, n_out, rv, nullptr);
    LAUNCHED(c);
    tmp_free(d_slots, st);
    CU(c, cudaGetLastError());
    return CG_OK;
}

CG_EXPORT int cg_mle_evaluate(cg_ctx* c, const cg_mle_desc* mle, const uint64_t* h_point, uint64_t h_out[2], cg_stream s) {
    if (!c || !mle || !h_out) return CG_ERR_INVALID;
    cg_sumcheck* sc = nullptr;
    CHK(cg_sumcheck_create(c, mle, 1, nullptr, nullptr, nullptr, 0, mle->num_vars, 1, CG_SC_FORCE_GENERIC, s, &sc));
    int rc = CG_OK;
    for (uint32_t j = 0; j < mle->num_vars && rc == CG_OK; j++) rc = cg_sumcheck_bind(sc, h_point + 2 * j);
    if (rc == CG_OK) rc = cg_sumcheck_final_evals(sc, h_out);
    cg_sumcheck_destroy(sc);
    return rc;
}

CG_EXPORT int cg_wit_infer_by_monomial_expr(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, const uint64_t* coeff,
                                            const uint32_t* off, const uint32_t* idx, uint32_t n_terms, uint32_t num_vars,
                                            uint64_t* d_out, cg_stream s) {
    if (!c || !d_out || (n_terms && (!coeff || !off || !idx))) return CG_ERR_INVALID;
    cudaStream_t st = S(c, s);
    CU(c, cudaSetDevice(c->device));
    const uint64_t n = 1ULL << num_vars;
    std::vector<MleSlot> slots(n_mles ? n_mles : 1);
    for (uint32_t i = 0; i < n_mles; i++) {
        if (mles[i].is_ext > CG_MLE_EXT) return set_err(c, CG_ERR_INVALID, "wit_infer: dense MLEs only");
        if (mles[i].num_vars != num_vars || mles[i].len != n) return set_err(c, CG_ERR_UNSUPPORTED, "wit_infer: every MLE must be full length 2^num_vars");
        slots[i] = MleSlot{mles[i].dptr, mles[i].is_ext ? 1u : 0u, 1u};
    }
    const uint32_t n_idx = n_terms ? off[n_terms] : 0;
    for (uint32_t q = 0; q < n_idx; q++) if (idx[q] >= n_mles) return set_err(c, CG_ERR_INVALID, "wit_infer: MLE index out of range");
    std::vector<uint64_t> cc(2 * (size_t)(n_terms ? n_terms : 1), 0);
    for (size_t i = 0; i < 2 * (size_t)n_terms; i++) cc[i] = coeff[i] >= GL_P ? coeff[i] - GL_P : coeff[i];
    std::vector<uint32_t> offv(n_terms + 1, 0), idxv(n_idx ? n_idx : 1, 0);
    if (n_terms) memcpy(offv.data(), off, sizeof(uint32_t) * (n_terms + 1));
    if (n_idx) memcpy(idxv.data(), idx, sizeof(uint32_t) * n_idx);
    void *d_slots = nullptr, *d_c = nullptr, *d_o = nullptr, *d_i = nullptr;
    int rc = upload_small(c, slots.data(), sizeof(MleSlot) * slots.size(), &d_slots, st);
    if (rc == CG_OK) rc = upload_small(c, cc.data(), sizeof(uint64_t) * cc.size(), &d_c, st);
    if (rc == CG_OK) rc = upload_small(c, offv.data(), sizeof(uint32_t) * offv.size(), &d_o, st);
    if (rc == CG_OK) rc = upload_small(c, idxv.data(), sizeof(uint32_t) * idxv.size(), &d_i, st);
    if (rc == CG_OK) {
        InferArgs a{(const MleSlot*)d_slots, (const ext_t*)d_c, (const uint32_t*)d_o, (const uint32_t*)d_i, n_terms, n, (ext_t*)d_out};
        wit_infer_kernel<<<grid_for(c, n, 8), CG_THREADS, 0, st>>>(a);
        LAUNCHED(c);
        if (cudaGetLastError() != cudaSuccess) rc = set_err(c, CG_ERR_CUDA, "wit_infer_kernel launch failed");
    }
    tmp_free(d_slots, st); tmp_free(d_c, st); tmp_free(d_o, st); tmp_free(d_i, st);
    return rc;
}

// ------------------------------------------------------------------------------------- tower
struct TowerSpecState {
    bool is_logup = false;
    uint32_t num_vars = 0;
    uint32_t layers = 0;             // witness.len(): product num_vars, logup num_vars + 1
    std::vector<const ext_t*> layer; // layer[l] -> base of [a|b] or [p1|p2|q1|q2] with arrays of len[l] ext
    std::vector<uint64_t> len;       // array length of layer l on this rank: 2^l, or 2^l / nranks for a distributed layer
    const ext_t* leaves[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ones = false;               // logup numerators implicit ones
    bool virt = false;               // the leaf layer is described, not stored (cg_tower_build_virtual)
    const VirtLeaf* d_virt[4] = {nullptr, nullptr, nullptr, nullptr};   // device descriptions of the leaf arrays
    uint32_t virt_l2m = 0;           // log2 of the padded record count (leaf index = row << l2m | record)
    uint32_t virt_nrec[4] = {0, 0, 0, 0};   // per leaf array: record count and padding value of its description
    ext_t virt_def[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
};
struct cg_tower {
    cg_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr;
    std::vector<TowerSpecState> specs;
    std::vector<void*> owned;
    uint32_t max_round = 0;          // max_round_index
    uint32_t n_prod = 0, n_logup = 0;
    // sharded tower (cg_tower_build_sharded): layers >= dist_from are sliced over the ranks of `comm` (rank r holds the slice r of
    // BOTH halves of a layer), smaller layers are replicated
    cg_comm* comm = nullptr;
    uint32_t g = 0, dist_from = UINT32_MAX;
};
static bool layer_distributed(const cg_tower* tw, uint32_t l) { return tw->comm && l >= tw->dist_from; }
CG_EXPORT int cg_tower_destroy(cg_tower* tw) {
    if (!tw) return CG_ERR_INVALID;
    for (void* p : tw->owned) tmp_free(p, tw->stream);
    delete tw;
    return CG_OK;
}
// array z of layer l of spec sp
static const ext_t* tower_arr(const TowerSpecState& sp, uint32_t l, uint32_t z) {
    if (l + 1 == sp.layers && !(sp.is_logup && sp.ones && z < 2)) {
        if (!sp.is_logup) return sp.leaves[z];
        return sp.leaves[z];
    }
    return sp.layer[l] + (uint64_t)z * sp.len[l];
}
// upper layers of one spec (infer_tower_product_witness / infer_tower_logup_witness, ceno_zkvm/src/scheme/utils.rs:488-659):
// sp.leaves (or sp.d_virt for a virtual leaf layer), sp.layers, sp.ones are set; allocates and fills sp.layer[]
static int tower_build_spec(cg_tower* tw, TowerSpecState& sp) {
    cg_ctx* c = tw->ctx;
    cudaStream_t st = tw->stream;
    cg_comm* cm = tw->comm;
    const int N = cm ? cm->nranks : 1, rank = cm ? cm->rank : 0;
    sp.layer.assign(sp.layers, nullptr);
    sp.len.assign(sp.layers, 0);
    const uint32_t arrs = sp.is_logup ? 4 : 2;
    const uint64_t top = sp.layers - 1;   // leaf layer index; arrays there have 2^top ext (globally)
    for (uint32_t l = 0; l <= top; l++) sp.len[l] = layer_distributed(tw, l) ? ((1ULL << l) / N) : (1ULL << l);
    if (cm && !layer_distributed(tw, (uint32_t)top)) return set_err(c, CG_ERR_UNSUPPORTED, "sharded tower: the leaf layer is too small to slice (build it on one rank)");
    // storage: layers written by peers (distributed ones and the first replicated one) live in the peer arena, at the same
    // offset on every rank; the rest in one pooled block
    std::vector<size_t> arena_off(sp.layers, SIZE_MAX);
    uint64_t pooled = 0;
    const bool ones_arrays = sp.ones && !sp.virt;   // materialised all-one numerators (a virtual spec reads them as "no p arrays")
    for (uint32_t l = 0; l < top; l++) {
        const bool peer_written = cm && layer_distributed(tw, l + 1);
        if (peer_written) CHK(arena_alloc(cm, sizeof(ext_t) * arrs * sp.len[l], &arena_off[l]));
        else pooled += (uint64_t)arrs * sp.len[l];
    }
    if (ones_arrays) pooled += 2 * sp.len[top];
    void* blk = nullptr;
    if (pooled) {
        CHK(tmp_alloc(c, sizeof(ext_t) * pooled, &blk, st));
        tw->owned.push_back(blk);
    }
    ext_t* base = (ext_t*)blk;
    for (uint32_t l = 0; l < top; l++) {
        if (arena_off[l] != SIZE_MAX) sp.layer[l] = (const ext_t*)(cm->arena[rank] + arena_off[l]);
        else { sp.layer[l] = base; base += (uint64_t)arrs * sp.len[l]; }
    }
    if (ones_arrays) {   // input-layer numerators materialised as ones (utils.rs:556-577)
        sp.layer[top] = base;   // only arrays 0,1 live here; q1,q2 stay in the caller's buffers
        fill_ext_kernel<<<grid_for(c, 2 * sp.len[top]), CG_THREADS, 0, st>>>(base, 2 * sp.len[top], ext_t{1, 0});
        LAUNCHED(c);
    }
    for (int32_t l = (int32_t)top - 1; l >= 0; l--) {
        const uint64_t n = 2 * sp.len[l + 1] / 2 * 1;   // results of this launch = length of this rank's layer l+1 arrays
        const uint64_t n_res = sp.len[l + 1];
        (void)n;
        const bool from_leaves = ((uint32_t)l + 1 == top);
        // destinations: array `zlo` takes results of the low half of the combined index, `zhi` of the high half
        auto dst_for = [&](uint32_t z_first) {   // z_first: 0 for (a|b) / (p1|p2), 2 for (q1|q2)
            TowerDst t;
            memset(&t, 0, sizeof(t));
            if (!cm || !layer_distributed(tw, l + 1)) {          // one device / replicated source: both halves are local and contiguous
                t.d[0] = (ext_t*)sp.layer[l] + (uint64_t)z_first * sp.len[l];
                t.n_dst = 1;
            } else if (layer_distributed(tw, l)) {               // shuffle: my results are two slices of ONE half of layer l
                const uint32_t part = rank < N / 2 ? 0 : 1;
                const size_t off = arena_off[l] + sizeof(ext_t) * (uint64_t)(z_first + part) * sp.len[l];
                t.d[0] = (ext_t*)(cm->arena[(2 * rank) % N] + off);
                t.d[1] = (ext_t*)(cm->arena[(2 * rank + 1) % N] + off);
                t.split = 1;
                t.n_dst = 2;
            } else {                                             // all-gather into the first replicated layer
                const uint32_t part = rank < N / 2 ? 0 : 1;
                const size_t off = arena_off[l] + sizeof(ext_t) * ((uint64_t)(z_first + part) * sp.len[l] + (uint64_t)(rank % (N / 2)) * n_res);
                for (int p = 0; p < N; p++) t.d[p] = (ext_t*)(cm->arena[p] + off);
                t.n_dst = N;
            }
            return t;
        };
        const TowerDst d0 = dst_for(0), d2 = dst_for(2);
        if (from_leaves && sp.virt) {
            if (!sp.is_logup) tower_prod_layer_virt_kernel<<<grid_for(c, n_res, 8), CG_THREADS, 0, st>>>(sp.d_virt[0], sp.d_virt[1], n_res, d0);
            else tower_logup_layer_virt_kernel<<<grid_for(c, n_res, 8), CG_THREADS, 0, st>>>(sp.d_virt[0], sp.d_virt[1], sp.d_virt[2], sp.d_virt[3], n_res, d0, d2);
        } else if (!sp.is_logup) {
            tower_prod_layer_kernel<<<grid_for(c, n_res, 8), CG_THREADS, 0, st>>>(tower_arr(sp, l + 1, 0), tower_arr(sp, l + 1, 1), n_res, d0, from_leaves);
        } else {
            const bool implicit_ones = from_leaves && sp.ones;
            tower_logup_layer_kernel<<<grid_for(c, n_res, 8), CG_THREADS, 0, st>>>(
                implicit_ones ? nullptr : tower_arr(sp, l + 1, 0), implicit_ones ? nullptr : tower_arr(sp, l + 1, 1),
                tower_arr(sp, l + 1, 2), tower_arr(sp, l + 1, 3), n_res, d0, d2, from_leaves);
        }
        LAUNCHED(c);
        if (cudaGetLastError() != cudaSuccess) return set_err(c, CG_ERR_CUDA, "tower layer kernel launch failed");
        if (cm && layer_distributed(tw, l + 1)) CHK(comm_barrier(cm, st));   // the partners' stores into my layer l are complete
    }
    if (sp.layers - 1 > tw->max_round) tw->max_round = sp.layers - 1;
    if (sp.is_logup) tw->n_logup++; else tw->n_prod++;
    return CG_OK;
}
#define CG_TOWER_DIST_MIN_LOCAL_LOG 12   // a layer is sliced over the ranks while every rank keeps >= 2^12 entries per array
static void tower_set_comm(cg_tower* tw, cg_comm* cm) {
    if (!cm || cm->nranks <= 1) return;
    tw->comm = cm;
    tw->g = 0;
    while ((1 << tw->g) < cm->nranks) tw->g++;
    tw->dist_from = tw->g + CG_TOWER_DIST_MIN_LOCAL_LOG;
    cm->arena_off = 0;   // one sharded tower at a time owns the arena ...
    comm_barrier(cm, tw->stream);   // ... and no rank may still be reading the previous tower's layers out of it
}
static int tower_build_impl(cg_ctx* c, cg_comm* cm, const cg_tower_spec* specs, uint32_t n_specs, cg_stream s, cg_tower** out) {
    if (!c || !specs || !out || n_specs == 0) return set_err(c, CG_ERR_INVALID, "cg_tower_build: bad argument");
    cudaStream_t st = S(c, s);
    CU(c, cudaSetDevice(c->device));
    cg_tower* tw = new cg_tower();
    tw->ctx = c;
    tw->stream = st;
    tower_set_comm(tw, cm);
    int rc = CG_OK;
    // reference order: product specs first, then logup specs (cpu/mod.rs:405-413)
    for (int pass = 0; pass < 2 && rc == CG_OK; pass++)
        for (uint32_t i = 0; i < n_specs && rc == CG_OK; i++) {
            const cg_tower_spec& in = specs[i];
            if ((in.is_logup != 0) != (pass == 1)) continue;
            TowerSpecState sp;
            sp.is_logup = in.is_logup != 0;
            sp.num_vars = in.num_vars;
            if (in.num_vars == 0 || in.num_vars > 32) { rc = set_err(c, CG_ERR_INVALID, "tower spec: num_vars out of range"); break; }
            sp.layers = sp.is_logup ? in.num_vars + 1 : in.num_vars;
            for (int z = 0; z < 4; z++) sp.leaves[z] = (const ext_t*)in.leaves[z];
            sp.ones = sp.is_logup && !in.leaves[0];
            if (sp.is_logup ? (!in.leaves[2] || !in.leaves[3] || (!in.leaves[0] != !in.leaves[1])) : (!in.leaves[0] || !in.leaves[1])) {
                rc = set_err(c, CG_ERR_INVALID, "tower spec: missing leaf pointer");
                break;
            }
            rc = tower_build_spec(tw, sp);
            if (rc == CG_OK) tw->specs.push_back(std::move(sp));
        }
    if (rc != CG_OK) { cg_tower_destroy(tw); return rc; }
    *out = tw;
    return CG_OK;
}
CG_EXPORT int cg_tower_build(cg_ctx* c, const cg_tower_spec* specs, uint32_t n_specs, cg_stream s, cg_tower** out) {
    return tower_build_impl(c, nullptr, specs, n_specs, s, out);
}
CG_EXPORT int cg_tower_build_sharded(cg_ctx* c, cg_comm* cm, const cg_tower_spec* specs, uint32_t n_specs, cg_stream s, cg_tower** out) {
    if (!cm) return set_err(c, CG_ERR_INVALID, "cg_tower_build_sharded: null comm");
    return tower_build_impl(c, cm, specs, n_specs, s, out);
}
static uint32_t ceil_log2_u64(uint64_t x) { uint32_t l = 0; while ((1ULL << l) < x) l++; return l; }
CG_EXPORT uint64_t cg_tower_interleave_out_len(uint32_t n_mles, uint64_t num_instances, uint32_t num_limbs) {
    uint64_t np2 = 1;
    while (np2 < num_instances) np2 <<= 1;
    if (np2 < 2) np2 = 2;   // next_pow2_instance_padding: minimum 2
    const uint32_t l2i = ceil_log2_u64(np2), l2m = ceil_log2_u64(n_mles), l2l = ceil_log2_u64(num_limbs);
    return 1ULL << (l2m + (l2i > l2l ? l2i - l2l : 0));
}
CG_EXPORT int cg_tower_interleave(cg_ctx* c, const cg_mle_desc* mles, uint32_t n_mles, uint64_t num_instances, uint32_t num_limbs,
                                  const uint64_t default_ext[2], uint64_t* d_out, cg_stream s) {
    if (!c || !mles || !n_mles || !default_ext || !d_out) return set_err(c, CG_ERR_INVALID, "cg_tower_interleave: null argument");
    if (!num_limbs || (num_limbs & (num_limbs - 1))) return set_err(c, CG_ERR_INVALID, "cg_tower_interleave: num_limbs must be a power of two (utils.rs:408)");
    uint64_t np2 = 1;
    while (np2 < num_instances) np2 <<= 1;
    if (np2 < 2) np2 = 2;
    const uint64_t mle_len = mles[0].len;
    std::vector<const void*> ptrs(n_mles);
    std::vector<uint32_t> ext(n_mles);
    for (uint32_t i = 0; i < n_mles; i++) {
        if (mles[i].is_ext > CG_MLE_EXT || !mles[i].dptr) return set_err(c, CG_ERR_INVALID, "cg_tower_interleave: dense MLEs only");
        if (mles[i].len != mle_len) return set_err(c, CG_ERR_INVALID, "cg_tower_interleave: every MLE must have the same length");
        if (mles[i].len > np2) return set_err(c, CG_ERR_INVALID, "cg_tower_interleave: MLE longer than the padded instance count (utils.rs:411-414)");
        if ((uintptr_t)mles[i].dptr & (mles[i].is_ext ? 15 : 7)) return set_err(c, CG_ERR_INVALID, "cg_tower_interleave: misaligned MLE pointer");
        ptrs[i] = mles[i].dptr;
        ext[i] = mles[i].is_ext;
    }
    cudaStream_t st = S(c, s);
    CU(c, cudaSetDevice(c->device));
    void *d_ptrs = nullptr, *d_ext = nullptr;
    CHK(upload_small(c, ptrs.data(), sizeof(void*) * n_mles, &d_ptrs, st));
    int rc = upload_small(c, ext.data(), sizeof(uint32_t) * n_mles, &d_ext, st);
    if (rc == CG_OK) {
        InterleaveArgs a;
        memset(&a, 0, sizeof(a));
        a.ptrs = (const void* const*)d_ptrs;
        a.is_ext = (const uint32_t*)d_ext;
        a.n_mles = n_mles;
        a.l2m = ceil_log2_u64(n_mles);
        a.out_len = cg_tower_interleave_out_len(n_mles, num_instances, num_limbs);
        a.per_fanin_len = mle_len / num_limbs ? mle_len / num_limbs : 1;
        a.num_instances = num_instances;
        a.mle_len = mle_len;
        a.def = ext_t{default_ext[0] % GL_P, default_ext[1] % GL_P};
        a.out = (ext_t*)d_out;
        const uint64_t n_inst_out = a.out_len >> a.l2m, per_instance = 1ULL << a.l2m;
        const dim3 grid((unsigned)((n_inst_out + 31) / 32), (unsigned)((per_instance + 31) / 32), num_limbs);
        if (grid.y > 65535 || grid.z > 65535) rc = set_err(c, CG_ERR_UNSUPPORTED, "cg_tower_interleave: too many records / limbs for one launch");
        else {
            tower_interleave_kernel<<<grid, dim3(32, 8), 0, st>>>(a);
            LAUNCHED(c);
            if (cudaGetLastError() != cudaSuccess) rc = set_err(c, CG_ERR_CUDA, "tower_interleave_kernel launch failed");
        }
    }
    tmp_free(d_ptrs, st);
    tmp_free(d_ext, st);
    return rc;
}
// ---- virtual leaf layers: the reference's GpuVirtualInterleavedExt (ceno_zkvm/src/scheme/gpu/mod.rs:2195-2268, builders
// build_prod_tower_from_virtual_ext_batch / build_logup_tower_from_virtual_ext_batch :2365-2402).  A spec is given by its RECORD
// MLEs; the interleaved fan-in leaves (2^ceil_log2(R) times the rows — 2^33 ext for keccak's 1094 lookup records at 2^22 rows)
// exist only as descriptions read by the first build level and by rounds 0 / 1 of the leaf-layer sumcheck.
static int make_virt_limbs(cg_tower* tw, const cg_tower_vgroup& g, const VirtLeaf** d_out /*[2]*/, uint64_t* out_len) {
    cg_ctx* c = tw->ctx;
    cudaStream_t st = tw->stream;
    VirtLeaf h[2];
    memset(h, 0, sizeof(h));
    uint64_t np2 = 1;
    while (np2 < g.num_instances) np2 <<= 1;
    if (np2 < 2) np2 = 2;
    const uint32_t l2m = g.n_records ? ceil_log2_u64(g.n_records) : 0;
    const uint64_t per_fanin_len = np2 / 2 ? np2 / 2 : 1;
    *out_len = cg_tower_interleave_out_len(g.n_records ? g.n_records : 1, g.num_instances, 2);
    const uint64_t n_inst_out = *out_len >> l2m;
    void *d_ptrs = nullptr, *d_ext = nullptr;
    uint64_t mle_len = 0;
    if (g.n_records) {
        std::vector<const void*> ptrs(g.n_records);
        std::vector<uint32_t> ext(g.n_records);
        mle_len = g.records[0].len;
        for (uint32_t i = 0; i < g.n_records; i++) {
            const cg_mle_desc& m = g.records[i];
            if (m.is_ext > CG_MLE_EXT || !m.dptr) return set_err(c, CG_ERR_INVALID, "cg_tower_build_virtual: dense record MLEs only");
            if (m.len != mle_len || m.len > np2) return set_err(c, CG_ERR_INVALID, "cg_tower_build_virtual: records must share one length <= the padded instance count (utils.rs:411-414)");
            if ((uintptr_t)m.dptr & (m.is_ext ? 15 : 7)) return set_err(c, CG_ERR_INVALID, "cg_tower_build_virtual: misaligned record pointer");
            ptrs[i] = m.dptr;
            ext[i] = m.is_ext;
        }
        CHK(upload_small(c, ptrs.data(), sizeof(void*) * g.n_records, &d_ptrs, st));
        tw->owned.push_back(d_ptrs);
        CHK(upload_small(c, ext.data(), sizeof(uint32_t) * g.n_records, &d_ext, st));
        tw->owned.push_back(d_ext);
    }
    for (int limb = 0; limb < 2; limb++) {
        VirtLeaf& v = h[limb];
        v.ptrs = (const void* const*)d_ptrs;
        v.is_ext = (const uint32_t*)d_ext;
        v.n_records = g.n_records;
        v.l2m = l2m;
        v.row_offset = per_fanin_len * limb;
        v.def = ext_t{g.default_ext[0] % GL_P, g.default_ext[1] % GL_P};
        // row counts exactly as interleaving_mles_to_mles takes them (utils.rs:433-456; tower_interleave_kernel)
        const uint64_t start = v.row_offset;
        if (g.n_records && start < g.num_instances) {
            const uint64_t valid = std::min<uint64_t>(per_fanin_len, g.num_instances - start);
            uint64_t ce = valid, cb = per_fanin_len;
            if (start + ce > mle_len) ce = 0;
            if (start + cb > mle_len) cb = 0;
            v.cnt_ext = std::min(ce, n_inst_out);
            v.cnt_base = std::min(cb, n_inst_out);
        }
    }
    void* d_v = nullptr;
    CHK(upload_small(c, h, sizeof(h), &d_v, st));
    tw->owned.push_back(d_v);
    d_out[0] = (const VirtLeaf*)d_v;
    d_out[1] = (const VirtLeaf*)d_v + 1;
    return CG_OK;
}
static int tower_build_virtual_impl(cg_ctx* c, cg_comm* cm, const cg_tower_vspec* specs, uint32_t n_specs, cg_stream s, cg_tower** out);
CG_EXPORT int cg_tower_build_virtual(cg_ctx* c, const cg_tower_vspec* specs, uint32_t n_specs, cg_stream s, cg_tower** out) {
    return tower_build_virtual_impl(c, nullptr, specs, n_specs, s, out);
}
CG_EXPORT int cg_tower_build_virtual_sharded(cg_ctx* c, cg_comm* cm, const cg_tower_vspec* specs, uint32_t n_specs, cg_stream s, cg_tower** out) {
    if (!cm) return set_err(c, CG_ERR_INVALID, "cg_tower_build_virtual_sharded: null comm");
    return tower_build_virtual_impl(c, cm, specs, n_specs, s, out);
}
static int tower_build_virtual_impl(cg_ctx* c, cg_comm* cm, const cg_tower_vspec* specs, uint32_t n_specs, cg_stream s, cg_tower** out) {
    if (!c || !specs || !out || n_specs == 0) return set_err(c, CG_ERR_INVALID, "cg_tower_build_virtual: bad argument");
    cudaStream_t st = S(c, s);
    CU(c, cudaSetDevice(c->device));
    uint32_t n_prod = 0, n_logup = 0;
    for (uint32_t i = 0; i < n_specs; i++) (specs[i].is_logup ? n_logup : n_prod)++;
    if (n_prod > CG_TOWER_MAX_PROD || n_logup > CG_TOWER_MAX_LOGUP)
        return set_err(c, CG_ERR_UNSUPPORTED, "cg_tower_build_virtual: at most 8 product and 4 logup specs (materialise the leaves with cg_tower_interleave beyond that)");
    cg_tower* tw = new cg_tower();
    tw->ctx = c;
    tw->stream = st;
    tower_set_comm(tw, cm);
    const uint32_t gbits = tw->g;   // sharded: the local leaf arrays are 1 / 2^g of the global ones
    int rc = CG_OK;
    for (int pass = 0; pass < 2 && rc == CG_OK; pass++)
        for (uint32_t i = 0; i < n_specs && rc == CG_OK; i++) {
            const cg_tower_vspec& in = specs[i];
            if ((in.is_logup != 0) != (pass == 1)) continue;
            if (!in.q.n_records || !in.q.records) { rc = set_err(c, CG_ERR_INVALID, "cg_tower_build_virtual: spec without records"); break; }
            TowerSpecState sp;
            sp.is_logup = in.is_logup != 0;
            sp.virt = true;
            sp.virt_l2m = ceil_log2_u64(in.q.n_records);
            uint64_t len_q = 0, len_p = 0;
            const VirtLeaf* dq[2];
            rc = make_virt_limbs(tw, in.q, dq, &len_q);
            if (rc != CG_OK) break;
            const ext_t dq_def{in.q.default_ext[0] % GL_P, in.q.default_ext[1] % GL_P};
            if (!sp.is_logup) {
                sp.d_virt[0] = dq[0]; sp.d_virt[1] = dq[1];
                sp.virt_nrec[0] = sp.virt_nrec[1] = in.q.n_records;
                sp.virt_def[0] = sp.virt_def[1] = dq_def;
                sp.num_vars = ceil_log2_u64(len_q) + 1 + gbits;
                sp.layers = sp.num_vars;
            } else {
                sp.d_virt[2] = dq[0]; sp.d_virt[3] = dq[1];
                sp.virt_nrec[2] = sp.virt_nrec[3] = in.q.n_records;
                sp.virt_def[2] = sp.virt_def[3] = dq_def;
                cg_tower_vgroup pg = in.p;
                if (!pg.n_records) {   // numerators all one (utils.rs:556-577): a description with no records and default 1
                    pg = in.q;
                    pg.n_records = 0;
                    pg.records = nullptr;
                    pg.default_ext[0] = 1; pg.default_ext[1] = 0;
                }
                const VirtLeaf* dp[2];
                rc = make_virt_limbs(tw, pg, dp, &len_p);
                if (rc != CG_OK) break;
                if (in.p.n_records && len_p != len_q) { rc = set_err(c, CG_ERR_INVALID, "cg_tower_build_virtual: numerator and denominator groups differ in shape"); break; }
                sp.d_virt[0] = dp[0]; sp.d_virt[1] = dp[1];
                sp.virt_nrec[0] = sp.virt_nrec[1] = pg.n_records;
                sp.virt_def[0] = sp.virt_def[1] = ext_t{pg.default_ext[0] % GL_P, pg.default_ext[1] % GL_P};
                sp.num_vars = ceil_log2_u64(len_q) + gbits;
                sp.layers = sp.num_vars + 1;
            }
            if (sp.num_vars == 0 || sp.num_vars > 34) { rc = set_err(c, CG_ERR_INVALID, "cg_tower_build_virtual: num_vars out of range"); break; }
            rc = tower_build_spec(tw, sp);
            if (rc == CG_OK) tw->specs.push_back(std::move(sp));
        }
    if (rc != CG_OK) { cg_tower_destroy(tw); return rc; }
    *out = tw;
    return CG_OK;
}
CG_EXPORT int cg_tower_output_evals(cg_tower* tw, uint32_t spec, uint64_t* h_out) {
    if (!tw || spec >= tw->specs.size() || !h_out) return CG_ERR_INVALID;
    const TowerSpecState& sp = tw->specs[spec];
    const uint32_t arrs = sp.is_logup ? 4 : 2;
    for (uint32_t z = 0; z < arrs; z++)
        CU(tw->ctx, cudaMemcpyAsync(h_out + 2 * z, tower_arr(sp, 0, z), sizeof(ext_t), cudaMemcpyDeviceToHost, tw->stream));
    CU(tw->ctx, cudaStreamSynchronize(tw->stream));
    for (uint32_t z = 0; z < 2 * arrs; z++) if (h_out[z] >= GL_P) h_out[z] -= GL_P;
    return CG_OK;
}
CG_EXPORT uint64_t cg_tower_proof_len(const cg_tower* tw) {
    if (!tw) return 0;
    uint64_t w = 0;
    for (uint32_t r = 1; r <= tw->max_round; r++) {
        w += (uint64_t)r * 3 * 2;
        for (const auto& sp : tw->specs) if (r < sp.layers) w += sp.is_logup ? 8 : 4;
    }
    return w;
}
CG_EXPORT uint32_t cg_tower_point_len(const cg_tower* tw) { return tw ? tw->max_round + 1 : 0; }

static void alpha_pows(const cg_transcript_vt* tr, uint32_t n, std::vector<ext_t>& out) {
    // get_challenge_pows: label "combine subset evals", one alpha, [1, a, a^2, ...]
    uint64_t a[2];
    tr->sample(tr->user, "combine subset evals", a);
    out.resize(n);
    unsigned __int128 P = GL_P;
    uint64_t p0 = 1, p1 = 0;
    for (uint32_t i = 0; i < n; i++) {
        out[i] = ext_t{p0, p1};
        // (p0 + p1 X)(a0 + a1 X), X^2 = 7 — transcript-side scalar work stays on the host like the reference
        unsigned __int128 c0 = ((unsigned __int128)p0 * a[0]) % P + (((unsigned __int128)p1 * a[1]) % P) * 7 % P;
        unsigned __int128 c1 = ((unsigned __int128)p0 * a[1]) % P + ((unsigned __int128)p1 * a[0]) % P;
        p0 = (uint64_t)(c0 % P);
        p1 = (uint64_t)(c1 % P);
    }
}

CG_EXPORT int cg_tower_create_proof(cg_tower* tw, const cg_transcript_vt* tr, uint64_t* h_proof, uint64_t* h_point) {
    if (!tw || !tr || !h_proof || !h_point) return CG_ERR_INVALID;
    cg_ctx* c = tw->ctx;
    CU(c, cudaSetDevice(c->device));
    const uint32_t n_alpha = tw->n_prod + 2 * tw->n_logup;
    std::vector<ext_t> alpha;
    alpha_pows(tr, n_alpha, alpha);
    std::vector<uint64_t> rt(2 * (size_t)(tw->max_round + 2), 0);
    uint32_t rt_len = 1;
    tr->sample(tr->user, "product_sum", rt.data());
    uint64_t w = 0;
    // one eq buffer and one fold workspace for every layer's sumcheck (sized for the largest layer)
    ScWorkspace lent;
    void* d_eq_all = nullptr;
    // Layers large enough for split-eq rounds hand eq over as its point (no table is built, streamed or folded before the
    // cluster tail takes over); CG_TOWER_VEQ=0 keeps the table everywhere (A/B switch).
    struct LayerPlan { uint32_t nv; size_t slots; bool any_virt; uint32_t vl2m; bool veq; };
    auto layer_plan = [&](uint32_t round) {
        LayerPlan lp{};
        const bool dist = layer_distributed(tw, round);
        lp.nv = dist ? round - tw->g : round;
        lp.slots = 1;
        uint32_t np = 0, nl = 0;
        for (const auto& sp : tw->specs) {
            if (round >= sp.layers) continue;
            lp.slots += sp.is_logup ? 4 : 2;
            (sp.is_logup ? nl : np)++;
            if (sp.virt && round + 1 == sp.layers && !lp.any_virt) { lp.any_virt = true; lp.vl2m = sp.virt_l2m; }
        }
        // Below 2^20 points a layer's time is launch latency, and the split form costs two more launches (tables, hand-over
        // to the tail) than building the table.  Read per call: one process can compare both paths (tests force 2^10).
        const char* e = getenv("CG_TOWER_VEQ");
        const char* em = getenv("CG_TOWER_VEQ_MIN_NV");
        const int on = (e ? atoi(e) : 1) && lp.nv >= (uint32_t)(em ? atoi(em) : 20);
        const bool fits = np <= CG_TOWER_MAX_PROD && nl <= CG_TOWER_MAX_LOGUP;
        const uint32_t J = (on && fits && lp.nv <= 32) ? veq_split_rounds(c, lp.nv, lp.slots, dist && tw->comm->nranks > 1, dist ? tw->g : 0) : 0;
        lp.veq = J >= 1 && (!lp.any_virt || (J >= 2 && lp.vl2m >= 3));
        return lp;
    };
    {
        size_t max_ws = 0, max_eq = sizeof(ext_t) * 2;
        for (uint32_t round = 1; round <= tw->max_round; round++) {
            const LayerPlan lp = layer_plan(round);
            max_ws = std::max(max_ws, lp.slots * (size_t)ws_per(1ULL << lp.nv) * sizeof(ext_t));
            // (below 2^20 the plan may change with the number of live sumchecks on the context: always keep room for the table)
            if (!lp.veq || lp.nv < 20) max_eq = std::max(max_eq, sizeof(ext_t) << lp.nv);
        }
        CHK(tmp_alloc(c, max_eq, &d_eq_all, tw->stream));
        if (tmp_alloc(c, max_ws, &lent.ptr, tw->stream) == CG_OK) lent.bytes = max_ws;   // on failure the layers allocate their own
    }
    struct LendGuard {
        const ScWorkspace* prev;
        explicit LendGuard(const ScWorkspace* w) : prev(tl_lent_ws) { tl_lent_ws = w; }
        ~LendGuard() { tl_lent_ws = prev; }
    } lend_guard(&lent);
    struct FreeGuard {
        void *a, *b;
        cudaStream_t st;
        ~FreeGuard() { tmp_free(a, st); tmp_free(b, st); }
    } free_guard{d_eq_all, lent.ptr, tw->stream};
    ext_t next_claim{0, 0};      // the claimed sum of the next layer's sumcheck, as the verifier derives it from this layer's evaluations
    bool have_claim = false;
    static const bool trace = getenv("CG_TOWER_TRACE") != nullptr;   // host-side time per phase, summed over the layers
    double t_eq = 0, t_create = 0, t_run = 0, t_fin = 0;
    auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (uint32_t round = 1; round <= tw->max_round; round++) {
        const uint32_t nv_glob = rt_len;
        const bool dist = layer_distributed(tw, round);   // this layer's arrays are rank slices (top g index bits = rank)
        const uint32_t nv = dist ? nv_glob - tw->g : nv_glob;
        const uint64_t n = 1ULL << nv;
        void* d_eq = d_eq_all;
        const double t0 = now();
        const LayerPlan lp = layer_plan(round);
        ext_t sc_{1, 0};
        if (dist)   // eq(rt, .) on my slice = eq(rt_top, rank) * eq(rt_low, .)
            for (uint32_t b = 0; b < tw->g; b++) {
                ext_t w{rt[2 * (nv + b)] % GL_P, rt[2 * (nv + b) + 1] % GL_P};
                if (!((tw->comm->rank >> b) & 1)) w = ext_t{hx_submod(1, w.c0), hx_submod(0, w.c1)};
                sc_ = hx_mul(sc_, w);
            }
        int rc = CG_OK;
        if (!lp.veq) {
            rc = cg_build_eq(c, rt.data(), nv, (uint64_t*)d_eq, 0, n, tw->stream);
            if (rc == CG_OK && dist) {
                scale_ext_kernel<<<grid_for(c, n, 8), CG_THREADS, 0, tw->stream>>>((ext_t*)d_eq, n, sc_);
                LAUNCHED(c);
            }
        }
        const double t1 = now();
        // MLE list in the reference's lift order: eq, then live product specs, then live logup specs
        std::vector<cg_mle_desc> mles;
        std::vector<const VirtLeaf*> virt{nullptr};
        std::vector<cg_sumcheck::VirtHost> virt_h(1);
        if (lp.veq) mles.push_back(cg_mle_desc{rt.data(), n, nv, CG_MLE_EQ});   // the point itself (host memory)
        else mles.push_back(cg_mle_desc{d_eq, n, nv, 1});
        TowerLayout tl;
        tl.on = true;
        tl.eq = 0;
        std::vector<uint32_t> first(tw->specs.size(), 0);
        uint32_t pi = 0, li = 0;
        for (size_t si = 0; si < tw->specs.size(); si++) {
            const TowerSpecState& sp = tw->specs[si];
            const uint32_t my = sp.is_logup ? li++ : pi++;
            if (round >= sp.layers) continue;
            first[si] = (uint32_t)mles.size();
            const uint32_t arrs = sp.is_logup ? 4 : 2;
            for (uint32_t z = 0; z < arrs; z++) {
                const bool vleaf = sp.virt && round + 1 == sp.layers;   // the leaf layer of a virtual spec: placeholder pointer + description
                mles.push_back(cg_mle_desc{vleaf ? (const void*)d_eq : (const void*)tower_arr(sp, round, z), n, nv, 1});
                virt.push_back(vleaf ? sp.d_virt[z] : nullptr);
                cg_sumcheck::VirtHost vh;
                if (vleaf) { vh.n_records = sp.virt_nrec[z]; vh.l2m = sp.virt_l2m; vh.def = sp.virt_def[z]; }
                virt_h.push_back(vh);
                (sp.is_logup ? tl.lk : tl.prod).push_back((uint32_t)mles.size() - 1);
            }
            if (sp.is_logup) { tl.lk_an.push_back(alpha[tw->n_prod + 2 * my]); tl.lk_ad.push_back(alpha[tw->n_prod + 2 * my + 1]); }
            else tl.prod_alpha.push_back(alpha[my]);
        }
        tl.alpha_one = false;
        cg_sumcheck* sc = nullptr;
        const bool fits = tl.prod_alpha.size() <= CG_TOWER_MAX_PROD && tl.lk_an.size() <= CG_TOWER_MAX_LOGUP;
        std::vector<uint64_t> coeff;
        std::vector<uint32_t> off{0}, idx;
        if (rc == CG_OK) {
            // monomial terms of the layer expression (cpu/mod.rs:441-485) — used by the generic kernel
            // when the spec count exceeds the specialised kernel's table
            for (size_t p = 0; p < tl.prod_alpha.size(); p++) {
                coeff.push_back(tl.prod_alpha[p].c0); coeff.push_back(tl.prod_alpha[p].c1);
                idx.insert(idx.end(), {0u, tl.prod[2 * p], tl.prod[2 * p + 1]}); off.push_back((uint32_t)idx.size());
            }
            for (size_t l = 0; l < tl.lk_an.size(); l++) {
                const uint32_t p1 = tl.lk[4 * l], p2 = tl.lk[4 * l + 1], q1 = tl.lk[4 * l + 2], q2 = tl.lk[4 * l + 3];
                coeff.push_back(tl.lk_an[l].c0); coeff.push_back(tl.lk_an[l].c1);
                idx.insert(idx.end(), {0u, p1, q2}); off.push_back((uint32_t)idx.size());
                coeff.push_back(tl.lk_an[l].c0); coeff.push_back(tl.lk_an[l].c1);
                idx.insert(idx.end(), {0u, p2, q1}); off.push_back((uint32_t)idx.size());
                coeff.push_back(tl.lk_ad[l].c0); coeff.push_back(tl.lk_ad[l].c1);
                idx.insert(idx.end(), {0u, q1, q2}); off.push_back((uint32_t)idx.size());
            }
            const uint32_t fl = CG_SC_FORCE_GENERIC | (lp.veq ? CG_SC_INT_DEFER_VEQ : 0u);
            if (dist) rc = sc_create_terms(c, mles.data(), (uint32_t)mles.size(), coeff.data(), off.data(), idx.data(), (uint32_t)off.size() - 1, nv, 3,
                                           fl, tw->stream, lp.veq ? &sc_ : nullptr, &sc, tw->comm, tw->g);
            else rc = sc_create_terms(c, mles.data(), (uint32_t)mles.size(), coeff.data(), off.data(), idx.data(),
                                      (uint32_t)off.size() - 1, nv, 3, fl, tw->stream, nullptr, &sc);
        }
        std::vector<uint64_t> fin(2 * mles.size()), chal(2 * (size_t)nv_glob);
        const double t2 = now();
        if (rc == CG_OK) {
            if (fits) sc->tl = tl;
            bool any_virt = false;
            for (const VirtLeaf* v : virt) any_virt |= v != nullptr;
            if (any_virt) {
                if (!fits) rc = set_err(c, CG_ERR_UNSUPPORTED, "virtual tower leaves need the specialised tower kernels (<= 8 product, <= 4 logup specs)");
                sc->virt = virt;
                sc->virt_h = virt_h;
            }
            if (rc == CG_OK && lp.veq) {   // (lp.veq implies fits)
                sc->virt_l2m = lp.vl2m;
                sc->veq.have_claim = have_claim && !getenv("CG_TOWER_VEQ_NOCLAIM");   // (testing: three-sum round 0)
                sc->veq.claim0 = next_claim;
                rc = veq_setup_split(sc);
            }
        }
        if (rc == CG_OK) {
            tr->sumcheck_begin(tr->user, nv_glob, 3);
            if (dist) {
                rc = sc_run_sharded(c, tw->comm, sc, (uint32_t)mles.size(), coeff.data(), off.data(), idx.data(), (uint32_t)off.size() - 1, nv_glob, 3,
                                    CG_SC_FORCE_GENERIC, tr->round_challenge, tr->user, nullptr, h_proof + w, fin.data(), chal.data(), tw->stream);
                sc = nullptr;   // consumed
            } else {
                rc = sc_run_host(sc, tr->round_challenge, tr->user, h_proof + w, fin.data(), chal.data());
            }
        }
        const double t3 = now();
        if (sc) cg_sumcheck_destroy(sc);
        if (rc != CG_OK) return rc;
        if (trace) {
            t_eq += t1 - t0; t_create += t2 - t1; t_run += t3 - t2; t_fin += now() - t3;
            fprintf(stderr, "[tower] layer %2u nv=%2u eq %.0f create %.0f run %.0f destroy %.0f us\n", round, nv, t1 - t0, t2 - t1, t3 - t2, now() - t3);
        }
        w += (uint64_t)nv_glob * 3 * 2;
        for (int pass = 0; pass < 2; pass++)
            for (size_t si = 0; si < tw->specs.size(); si++) {
                const TowerSpecState& sp = tw->specs[si];
                if ((sp.is_logup ? 1 : 0) != pass || round >= sp.layers) continue;
                const uint32_t arrs = sp.is_logup ? 4 : 2;
                tr->append_exts(tr->user, fin.data() + 2 * first[si], arrs);
                memcpy(h_proof + w, fin.data() + 2 * first[si], sizeof(ext_t) * arrs);
                w += 2 * arrs;
            }
        uint64_t rm[2];
        tr->sample(tr->user, "merge", rm);
        memcpy(rt.data(), chal.data(), sizeof(uint64_t) * 2 * nv_glob);
        rt[2 * nv_glob] = rm[0];
        rt[2 * nv_glob + 1] = rm[1];
        rt_len = nv_glob + 1;
        alpha_pows(tr, n_alpha, alpha);
        // next claim = sum over the specs still alive of alpha * v(point, merge), v(point, y) = (1 - y) first half + y second half
        {
            const ext_t rme{rm[0] % GL_P, rm[1] % GL_P}, omr{hx_submod(1, rme.c0), hx_submod(0, rme.c1)};
            next_claim = ext_t{0, 0};
            uint32_t pi2 = 0, li2 = 0;
            for (size_t si = 0; si < tw->specs.size(); si++) {
                const TowerSpecState& sp = tw->specs[si];
                const uint32_t my = sp.is_logup ? li2++ : pi2++;
                if (round + 1 >= sp.layers) continue;
                auto F = [&](uint32_t z) { return ext_t{fin[2 * (first[si] + z)] % GL_P, fin[2 * (first[si] + z) + 1] % GL_P}; };
                auto merged = [&](uint32_t z) { return hx_add(hx_mul(F(z), omr), hx_mul(F(z + 1), rme)); };
                if (!sp.is_logup) next_claim = hx_add(next_claim, hx_mul(alpha[my], merged(0)));
                else {
                    next_claim = hx_add(next_claim, hx_mul(alpha[tw->n_prod + 2 * my], merged(0)));
                    next_claim = hx_add(next_claim, hx_mul(alpha[tw->n_prod + 2 * my + 1], merged(2)));
                }
            }
            have_claim = true;
        }
    }
    memcpy(h_point, rt.data(), sizeof(uint64_t) * 2 * rt_len);
    if (trace) fprintf(stderr, "[tower] total: eq %.0f create %.0f run %.0f destroy %.0f us\n", t_eq, t_create, t_run, t_fin);
    return CG_OK;
}

// --------------------------------------------------------------------- Poseidon2 / Merkle (a9)
static_assert(sizeof(P2Params) == sizeof(cg_poseidon2_params), "params layout");
CG_EXPORT int cg_poseidon2_set_params(cg_ctx* c, const cg_poseidon2_params* p) {
    if (!c || !p) return CG_ERR_INVALID;
    if (p->mds_variant > 1) return set_err(c, CG_ERR_INVALID, "cg_poseidon2_set_params: mds_variant must be 0 or 1");
    CU(c, cudaSetDevice(c->device));
    if (!c->d_p2) CU(c, cudaMalloc((void**)&c->d_p2, sizeof(P2Params)));
    P2Params h;
    memcpy(&h, p, sizeof(h));
    uint64_t* w = reinterpret_cast<uint64_t*>(&h);
    for (size_t i = 0; i < (8 * 8 + 22 + 8); i++) w[i] = w[i] >= GL_P ? w[i] - GL_P : w[i];
    CU(c, cudaDeviceSynchronize());   // hash kernels on the lanes' non-blocking streams read d_p2: replace it only when the device is idle
    CU(c, cudaMemcpy(c->d_p2, &h, sizeof(h), cudaMemcpyHostToDevice));
    return CG_OK;
}
CG_EXPORT int cg_poseidon2_permute(cg_ctx* c, uint64_t* d_states, uint64_t n, cg_stream s) {
    if (!c || !d_states) return CG_ERR_INVALID;
    if (!c->d_p2) return set_err(c, CG_ERR_STATE, "cg_poseidon2_set_params has not been called (constants are upstream-only: the caller supplies them)");
    if (n == 0) return CG_OK;
    p2_permute_kernel<<<(unsigned)((n + 127) / 128), 128, 0, S(c, s)>>>(c->d_p2, d_states, n);
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    return CG_OK;
}
CG_EXPORT int cg_merkle_commit(cg_ctx* c, const uint64_t* d_matrix, uint64_t width, uint64_t height, int col_major,
                               uint64_t* d_tree, uint64_t h_root[4], cg_stream s) {
    if (!c || !d_matrix || !d_tree || width == 0) return CG_ERR_INVALID;
    if (!c->d_p2) return set_err(c, CG_ERR_STATE, "cg_poseidon2_set_params has not been called (constants are upstream-only: the caller supplies them)");
    if (height == 0 || (height & (height - 1))) return set_err(c, CG_ERR_INVALID, "cg_merkle_commit: height must be a power of two");
    if (((uintptr_t)d_tree & 31) || ((uintptr_t)d_matrix & 7)) return set_err(c, CG_ERR_INVALID, "cg_merkle_commit: d_tree must be 32-byte aligned (digests are stored as 256-bit words)");
    cudaStream_t st = S(c, s);
    CU(c, cudaSetDevice(c->device));
    p2_leaf_kernel<<<(unsigned)((height + 127) / 128), 128, 0, st>>>(c->d_p2, d_matrix, width, height, col_major, d_tree);
    LAUNCHED(c);
    uint64_t off = 0, n = height;
    while (n > 1) {
        p2_compress_kernel<<<(unsigned)((n / 2 + 127) / 128), 128, 0, st>>>(c->d_p2, d_tree + 4 * off, n / 2, d_tree + 4 * (off + n));
        LAUNCHED(c);
        off += n;
        n /= 2;
    }
    CU(c, cudaGetLastError());
    if (h_root) {
        CU(c, cudaMemcpyAsync(h_root, d_tree + 4 * off, 32, cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
    }
    return CG_OK;
}

// --------------------------------------------------------------------- rotation pre-passes (f-3)
CG_EXPORT int cg_rotation_next_base_mle(cg_ctx* c, const cg_mle_desc* mle, uint32_t cyclic_group_log2, uint64_t* d_out, cg_stream s) {
    if (!c || !mle || !d_out) return CG_ERR_INVALID;
    if (cyclic_group_log2 != 5 && cyclic_group_log2 != 6) return set_err(c, CG_ERR_UNSUPPORTED, "BooleanHypercube supports 5 or 6 variables (booleanhypercube.rs:119-121)");
    if (mle->is_ext) return set_err(c, CG_ERR_INVALID, "rotation_next_base_mle takes a base-field MLE (get_base_field_vec, utils.rs:35)");
    const uint64_t n = mle->len;
    if (n % (1ULL << cyclic_group_log2)) return set_err(c, CG_ERR_INVALID, "rotation: length must be a multiple of the cyclic group size");
    CU(c, cudaSetDevice(c->device));
    rotation_next_base_kernel<<<grid_for(c, n, 8), CG_THREADS, 0, S(c, s)>>>((const uint64_t*)mle->dptr, d_out, n, cyclic_group_log2);
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    return CG_OK;
}
CG_EXPORT int cg_rotation_selector(cg_ctx* c, const uint64_t* d_eq_ext, uint64_t total_len, uint32_t cyclic_subgroup_size,
                                   uint32_t cyclic_group_log2, uint64_t* d_out_ext, cg_stream s) {
    if (!c || !d_eq_ext || !d_out_ext) return CG_ERR_INVALID;
    if (cyclic_group_log2 != 5 && cyclic_group_log2 != 6) return set_err(c, CG_ERR_UNSUPPORTED, "BooleanHypercube supports 5 or 6 variables (booleanhypercube.rs:119-121)");
    if (cyclic_subgroup_size > (1u << cyclic_group_log2)) return set_err(c, CG_ERR_INVALID, "cyclic_subgroup_size > cyclic_group_size (utils.rs:62)");
    if (total_len % (1ULL << cyclic_group_log2)) return set_err(c, CG_ERR_INVALID, "rotation: length must be a multiple of the cyclic group size");
    uint64_t keep = 0, cur = 1;
    for (uint32_t i = 0; i < cyclic_subgroup_size; i++) {
        keep |= 1ULL << cur;
        cur <<= 1;
        if (cur >> cyclic_group_log2) cur ^= (cyclic_group_log2 == 5 ? 0x25u : 0x43u);
    }
    CU(c, cudaSetDevice(c->device));
    rotation_selector_kernel<<<grid_for(c, total_len, 8), CG_THREADS, 0, S(c, s)>>>((const ext_t*)d_eq_ext, (ext_t*)d_out_ext, total_len, cyclic_group_log2, keep);
    LAUNCHED(c);
    CU(c, cudaGetLastError());
    return CG_OK;
}

#include "sched.cuh"
#include "ntt_host.cuh"
#include "ecc_host.cuh"
#include "basefold_host.cuh"
