/*
 * ceno_b200.h — C ABI of the B200-native GKR-sumcheck device backend.
 *
 * This is the drop-in boundary for the ONE hot path of scroll-tech/ceno named by
 * BASELINE.json: per-round sumcheck evaluation, fix_variable fold, build_eq_x_r (+ selector
 * masks), the tower (grand-product / logUp) prover built from them, and the Merkle leaf hash.
 * The reference funnels that path into ~10 calls on its private `cuda_hal` crate
 * (SURVEY.md §2.3 / §B); each entry point below names the reference interface it replaces.
 * INTEGRATION.md shows the Rust `extern "C"` binding a maintainer would add.
 *
 * Conventions
 *  - Field: Goldilocks p = 2^64 - 2^32 + 1.  A base element is one canonical u64; an
 *    extension element (GoldilocksExt2, X^2 = 7) is two consecutive u64 limbs [c0, c1] — the
 *    host in-memory layout of the field type, which the reference transmutes to/from the
 *    device representation (gkr_iop/src/gpu/mod.rs:311-324, gkr_iop/src/gkr/layer/gpu/mod.rs:252-280).
 *    Inputs may be any u64 < 2^64 (p3's Goldilocks is not always canonical); outputs are canonical.
 *  - MLE index b = sum_i b_i 2^i; sumcheck round j binds variable j (LSB first), i.e. round 0
 *    folds adjacent pairs (2b, 2b+1)  (gkr_iop/src/utils.rs:209-232).
 *  - Every call returns 0 on success or a CG_ERR_* code; cg_last_error(ctx) gives the text.
 *    Nothing aborts (reference: Result<_, HalError> mapped to ZKVMError::BackendError,
 *    ceno_zkvm/src/scheme/gpu/mod.rs:347-351).
 *  - All device work is issued on the caller's stream (reference binds one CUDA stream per OS
 *    thread and forbids default-stream fallback, gkr_iop/src/gpu/mod.rs:79-154); the library is
 *    re-entrant per (ctx, stream).  `stream` is a cudaStream_t passed as void*; NULL = the
 *    context's own non-blocking stream.
 *  - The library never frees or writes caller MLE buffers: like the reference's
 *    prove_generic_sumcheck_gpu it takes shared inputs and folds into pooled workspace.
 *  - There is no CPU fallback: cg_init fails with CG_ERR_NO_DEVICE when no sm_100 GPU is present.
 */
#ifndef CENO_B200_H
#define CENO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CG_OK 0
#define CG_ERR_CUDA 1        /* a CUDA runtime call failed */
#define CG_ERR_INVALID 2     /* bad argument */
#define CG_ERR_UNSUPPORTED 3 /* valid in the reference but not implemented here (see message) */
#define CG_ERR_OOM 4
#define CG_ERR_NO_DEVICE 5
#define CG_ERR_STATE 6       /* call order violated (e.g. bind before round_eval) */

typedef struct cg_ctx cg_ctx;
typedef struct cg_sumcheck cg_sumcheck;
typedef void* cg_stream; /* cudaStream_t */

/* Device MLE descriptor.  Mirrors what MultilinearExtensionGpu carries
 * (gkr_iop/src/gpu/mod.rs:157-161, 331-344): a pointer into a (possibly larger, column-major)
 * device allocation, the occupied prefix `len` (<= 2^num_vars; the implicit tail is zero,
 * SURVEY §A9) and whether elements are base (8 B) or ext (16 B).  dptr must be 16-byte aligned
 * (32-byte for best bandwidth). */
typedef struct cg_mle_desc {
    const void* dptr;
    uint64_t len;
    uint32_t num_vars;
    uint32_t is_ext; /* CG_MLE_BASE, CG_MLE_EXT or CG_MLE_EQ */
} cg_mle_desc;
#define CG_MLE_BASE 0u
#define CG_MLE_EXT 1u
/* Virtual eq MLE, accepted by the cg_sumcheck_* entry points: the polynomial is eq(w, .) for the point w
 * and no table exists.  dptr = HOST pointer to w (num_vars ext = 2*num_vars u64, read during the call),
 * len is ignored.  It behaves exactly like the table cg_build_eq(w) would produce (same round messages,
 * same final evaluation eq(w, r)); for the shape eq*A*B (one degree-3 product, coefficient 1) the large
 * rounds then run the split-eq kernel, which never streams or folds an eq table.  This is what the
 * reference's virtual device MLEs are for (GpuVirtualInterleavedExt, ceno_zkvm/src/scheme/gpu/mod.rs:2195-2268):
 * hand the device a description instead of 2^k elements.  In cg_sumcheck_prove_sharded the point is the GLOBAL one
 * (num_vars_global ext) although num_vars is the local count: the rank's constant factor eq(w_top, rank) is derived
 * by the library.  Not accepted together with cg_sumcheck_attach_comm (step API). */
#define CG_MLE_EQ 2u

/* ---- lifecycle / memory: replaces cuda_hal context + mem_pool
 * (gkr_iop/src/gpu/mod.rs:53-66 get_cuda_hal; alloc_*_on_device / alloc_*_from_host / to_cpu_vec,
 *  gkr_iop/src/gpu/mod.rs:260-320, 593; mem_pool stats ceno_zkvm/src/scheme/gpu/mod.rs:226-269) */
int cg_init(int device_id, cg_ctx** ctx);
int cg_destroy(cg_ctx* ctx);
const char* cg_last_error(cg_ctx* ctx);
const char* cg_version(void);
int cg_device_info(cg_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes, size_t* total_bytes);
int cg_alloc(cg_ctx* ctx, size_t bytes, void** dptr);   /* pooled; 256-byte aligned */
/* cg_free returns the block to the pool IMMEDIATELY: the caller must have synchronised every stream that still reads or
 * writes it (the library's own calls may return while their kernels run).  cg_free_async is the stream-ordered form:
 * the block becomes reusable only after the work enqueued on `s` so far has finished (alloc_*_on_device buffers dropped
 * by a lane thread, gkr_iop/src/gpu/mod.rs:79-154). */
int cg_free(cg_ctx* ctx, void* dptr);
int cg_free_async(cg_ctx* ctx, void* dptr, cg_stream s);
int cg_pool_stats(cg_ctx* ctx, size_t* used_bytes, size_t* reserved_bytes);
int cg_pool_trim(cg_ctx* ctx);                          /* cudaFree every cached block */
int cg_h2d(cg_ctx* ctx, void* dst, const void* src, size_t bytes, cg_stream s); /* async on s */
int cg_d2h(cg_ctx* ctx, void* dst, const void* src, size_t bytes, cg_stream s); /* async on s */
int cg_d2d(cg_ctx* ctx, void* dst, const void* src, size_t bytes, cg_stream s); /* dtod_copy_sync, scheme/gpu/mod.rs:3274 */
int cg_stream_sync(cg_ctx* ctx, cg_stream s);
int cg_host_alloc_pinned(cg_ctx* ctx, size_t bytes, void** hptr);
int cg_host_free_pinned(cg_ctx* ctx, void* hptr);
/* number of kernels this context has launched (bench.py's gpu_launches) */
uint64_t cg_launch_count(cg_ctx* ctx);

/* ---- kernel (iii-eq): build_eq_x_r_vec with prefix masking.
 * Replaces build_mle_as_ceno(hal.inner, gpu_points, &mut out, offset, num_instances, stream)
 * (gkr_iop/src/gkr/layer/gpu/utils.rs:172-180, 211-219) and the CPU build_eq_x_r_vec call sites
 * (gkr_iop/src/selector.rs:140,152; ceno_zkvm/src/scheme/cpu/mod.rs:121,417).
 * out[b] = prod_i (b_i r_i + (1-b_i)(1-r_i)) for offset <= b < offset+num_instances, else 0.
 * Pass offset = 0, num_instances = 2^k for the unmasked table.  h_point_ext: k ext elements on
 * the HOST (k*2 u64).  d_out_ext: 2^k ext elements on the device, 32-byte aligned (the same holds for the outputs of
 * cg_selector_compute / cg_ecc_quark_selectors and for cg_merkle_commit's d_tree: they are written with 256-bit stores;
 * a misaligned pointer is rejected with CG_ERR_INVALID). */
int cg_build_eq(cg_ctx* ctx, const uint64_t* h_point_ext, uint32_t k, uint64_t* d_out_ext,
                uint64_t offset, uint64_t num_instances, cg_stream s);

/* SelectorType::compute (gkr_iop/src/selector.rs:131-245); replaces build_eq_x_r_with_sel_gpu /
 * ordered_sparse_selector_gpu (gkr_iop/src/gkr/layer/gpu/utils.rs:121-229).
 * kind: 0 Whole, 1 Prefix, 2 OrderedSparse{indices (sorted, host), inner_vars}, 3 QuarkBinaryTreeLessThan. */
#define CG_SEL_WHOLE 0
#define CG_SEL_PREFIX 1
#define CG_SEL_ORDERED_SPARSE 2
#define CG_SEL_QUARK_LT 3
int cg_selector_compute(cg_ctx* ctx, int kind, const uint64_t* h_point_ext, uint32_t num_vars,
                        uint64_t offset, uint64_t num_instances, const uint64_t* h_indices,
                        uint32_t n_indices, uint32_t inner_vars, uint64_t* d_out_ext, cg_stream s);

/* ---- kernel (ii): fix_variable, one variable (LSB).  MultilinearExtension::fix_variables*
 * (external multilinear_extensions; invoked inside IOPProverState).
 * d_out[b] = f[2b] + r (f[2b+1] - f[2b]), b < 2^(num_vars-1); output is always ext.
 * Out of place (d_out must not overlap the input). */
int cg_fix_variable(cg_ctx* ctx, const cg_mle_desc* mles, uint32_t n_mles, const uint64_t r_ext[2],
                    uint64_t* const* d_out_ext, cg_stream s);

/* MultilinearExtension::evaluate(point) for a device MLE (used for final-evaluation checks,
 * eval_cols_at_point_gpu in ceno_zkvm/src/scheme/gpu/mod.rs:24-40). */
int cg_mle_evaluate(cg_ctx* ctx, const cg_mle_desc* mle, const uint64_t* h_point_ext, uint64_t h_out_ext[2], cg_stream s);

/* ---- kernel (i) + round loop: IOPProverState (external sumcheck crate; call sites
 * gkr_iop/src/gkr/layer/cpu/mod.rs:87-91, 217-237; ceno_zkvm/src/scheme/cpu/mod.rs:273-279, 490-498).
 * Replaces CudaHalBB31::prove_generic_sumcheck_gpu(mles, mle_size_info, term_coefficients,
 * mle_indices_per_term, max_num_var, max_degree, plan, transcript, stream)
 * (gkr_iop/src/gkr/layer/gpu/mod.rs:259-271).
 *
 * P(x) = sum_t coeff_t * prod_{i in term t} mle_i(x) over {0,1}^num_vars.
 *   term_coeff_ext : n_terms ext (HOST)      (extract_mle_relationships_from_monomial_terms)
 *   term_offsets   : n_terms+1 u32 (HOST), CSR offsets into term_mle_idx
 *   term_mle_idx   : indices into `mles`
 * Round message = [p(1) .. p(degree)] (p(0) is not sent).
 * Mixed sizes — the cross-chip batched main sumcheck, prove_batched_main_constraints
 * (ceno_zkvm/src/scheme/cpu/mod.rs:1052-1390, GPU call ceno_zkvm/src/scheme/gpu/mod.rs:2968-2981): an MLE with
 * k' < num_vars variables stands for F(x) = f(x_0..x_{k'-1}) * prod_{j>=k'} x_j ("frontload"; verifier.rs:180-238,
 * restated in ceno_recursion_v2/src/main/mod.rs:3414-3448).  cg_sumcheck_prove / _standin_device accept such lists
 * when the factors of every term share one num_vars (one chip) and no term is a bare constant; the final evaluation
 * reported for a small MLE is the raw f(r_0..r_{k'-1}).  The step API (cg_sumcheck_create ...) and the sharded
 * prove take uniform sizes only. */
#define CG_SC_DEFAULT 0u
#define CG_SC_FORCE_GENERIC 1u /* disable shape-specialised kernels (testing) */
#define CG_SC_NO_FUSE 2u       /* separate fold and eval launches (testing / profiling) */
#define CG_SC_NO_TAIL 8u       /* one launch per round to the end (no persistent shared-memory tail kernel) */
#define CG_SC_NO_PLAN 16u      /* evaluate monomial terms one by one (no grouping by shared ext factors) */
#define CG_SC_NO_MID 32u       /* no cooperative persistent kernel for the mid-size rounds */
#define CG_SC_NO_PERSIST 128u  /* one launch per split-eq round (no persistent cooperative round kernel; testing / per-launch profiling) */
#define CG_SC_NO_DERIVE 64u    /* split-eq rounds accumulate all three bilinear sums (no claim-derived q(0); testing) */
#define CG_SC_PROFILE 4u       /* time each round's kernels with CUDA events on the launching stream */
int cg_sumcheck_create(cg_ctx* ctx, const cg_mle_desc* mles, uint32_t n_mles,
                       const uint64_t* term_coeff_ext, const uint32_t* term_offsets,
                       const uint32_t* term_mle_idx, uint32_t n_terms, uint32_t num_vars,
                       uint32_t degree, uint32_t flags, cg_stream s, cg_sumcheck** out);
/* Evaluate the current round's message into h_out_ext (degree ext = degree*2 u64).  Blocks until
 * the values are on the host.  A pending challenge from cg_sumcheck_bind is applied first (the
 * fold of round j-1 is fused with the evaluation of round j). */
int cg_sumcheck_round_eval(cg_sumcheck* sc, uint64_t* h_out_ext);
/* Bind the current round's variable to r (the transcript's challenge). */
int cg_sumcheck_bind(cg_sumcheck* sc, const uint64_t r_ext[2]);
/* After num_vars binds: get_mle_flatten_final_evaluations() — one ext per MLE, input order. */
int cg_sumcheck_final_evals(cg_sumcheck* sc, uint64_t* h_out_ext);
uint32_t cg_sumcheck_round(const cg_sumcheck* sc);
/* current (partially folded) contents of MLE i: device pointer to 2^(num_vars-round) ext
 * (or the caller's original buffer before the first fold).  Valid until the next call on sc. */
int cg_sumcheck_peek(cg_sumcheck* sc, uint32_t mle, const void** dptr, uint64_t* len, uint32_t* is_ext);
int cg_sumcheck_destroy(cg_sumcheck* sc);

/* Whole loop with the transcript behind a host callback — what the reference does by handing
 * &mut BasicTranscript to the device crate (gkr_iop/src/gkr/layer/gpu/mod.rs:252-253, 268).
 * cb receives the round's evaluations and must return the challenge (absorb evals, absorb
 * b"Internal round", sample).  Outputs on the HOST:
 *   h_round_evals num_vars*degree ext, h_final_evals n_mles ext, h_challenges num_vars ext. */
typedef void (*cg_challenge_cb)(void* user, uint32_t round, const uint64_t* round_evals_ext,
                                uint32_t degree, uint64_t out_r_ext[2]);
int cg_sumcheck_prove(cg_ctx* ctx, const cg_mle_desc* mles, uint32_t n_mles,
                      const uint64_t* term_coeff_ext, const uint32_t* term_offsets,
                      const uint32_t* term_mle_idx, uint32_t n_terms, uint32_t num_vars,
                      uint32_t degree, uint32_t flags, cg_challenge_cb cb, void* user,
                      uint64_t* h_round_evals, uint64_t* h_final_evals, uint64_t* h_challenges,
                      cg_stream s);

/* Same loop with a DEVICE-RESIDENT challenger: no host round trip between rounds; every round's
 * kernels are enqueued back to back and the messages are read once at the end.  The challenger is
 * the documented stand-in sponge (splitmix64, `cg_standin_*` below) — the reference's Poseidon2
 * constants are upstream-only; a Poseidon2 challenger slots in behind the same
 * device hook once they are available.  h_state_inout: the 8-byte stand-in transcript state. */
int cg_sumcheck_prove_standin_device(cg_ctx* ctx, const cg_mle_desc* mles, uint32_t n_mles,
                                     const uint64_t* term_coeff_ext, const uint32_t* term_offsets,
                                     const uint32_t* term_mle_idx, uint32_t n_terms, uint32_t num_vars,
                                     uint32_t degree, uint32_t flags, uint64_t* h_state_inout,
                                     uint64_t* h_round_evals, uint64_t* h_final_evals,
                                     uint64_t* h_challenges, cg_stream s);

/* ---- multi-GPU (one process per GPU).  The reference has no multi-GPU path ("Distributed Sumcheck —
 * TODO", docs/src/optimizations.md:3-5; single device id gkr_iop/src/gpu/mod.rs:55-56); this is the
 * hypercube slicing of SURVEY §8e.  Each rank creates a mailbox and publishes its 64-byte CUDA-IPC
 * handle; the host framework all-gathers the handles once (torch.distributed, MPI, a file ...) and
 * every rank connects.  After that the ranks exchange per-round partial sums directly over NVLink
 * peer memory from inside the round kernels — no host or NCCL call per round.  The caller must
 * barrier between connect and first use, and before destroy. */
typedef struct cg_comm cg_comm;
int cg_comm_create(cg_ctx* ctx, int rank, int nranks, cg_comm** out, uint8_t handle_out[64]);
int cg_comm_connect(cg_comm* comm, const uint8_t* all_handles /* nranks * 64 bytes, rank order */);
int cg_comm_destroy(cg_comm* comm);
/* step API: combine this sumcheck's round messages across the ranks of `comm` (call before round 0) */
int cg_sumcheck_attach_comm(cg_sumcheck* sc, cg_comm* comm);
/* Whole sharded proof: `mles` are this rank's slices (num_vars = num_vars_global - log2 nranks),
 * outputs are the GLOBAL proof (num_vars_global rounds), identical on every rank.  Exactly one of
 * (cb, h_standin_state) selects the host transcript or the device-resident challenger. */
int cg_sumcheck_prove_sharded(cg_ctx* ctx, cg_comm* comm, const cg_mle_desc* mles, uint32_t n_mles,
                              const uint64_t* term_coeff_ext, const uint32_t* term_offsets,
                              const uint32_t* term_mle_idx, uint32_t n_terms, uint32_t num_vars_global,
                              uint32_t degree, uint32_t flags, cg_challenge_cb cb, void* user,
                              uint64_t* h_standin_state, uint64_t* h_round_evals, uint64_t* h_final_evals,
                              uint64_t* h_challenges, cg_stream s);

/* Per-round device time (ms) of the last cg_sumcheck_prove* call made with CG_SC_PROFILE on this
 * context: n receives the round count, up to `cap` values are written.  (The reference wraps the
 * same phases in tracing spans / NVTX ranges, ceno_zkvm/src/scheme/prover.rs:92-180.) */
int cg_profile_last(cg_ctx* ctx, float* ms_out, uint32_t cap, uint32_t* n);

/* Stand-in transcript on the host (same algorithm as the device challenger; NOT Poseidon2). */
void cg_standin_init(uint64_t* state, const uint8_t* label, uint64_t len);
void cg_standin_append_message(uint64_t* state, const uint8_t* msg, uint64_t len);
void cg_standin_append_ext(uint64_t* state, const uint64_t* ext, uint64_t n);
void cg_standin_sample(uint64_t* state, const char* label, uint64_t out_ext[2]);
/* a ready-made cg_challenge_cb over a stand-in state (user = uint64_t* state) */
void cg_standin_challenge_cb(void* user, uint32_t round, const uint64_t* evals, uint32_t degree, uint64_t out_r[2]);

/* ---- tower prover: CpuTowerProver::create_proof (ceno_zkvm/src/scheme/cpu/mod.rs:346-554);
 * replaces cuda_hal.tower.create_proof(hal, TowerInput{prod_specs, logup_specs}, NUM_FANIN,
 * transcript, stream) (ceno_zkvm/src/scheme/gpu/mod.rs:336-353) and the tower builders
 * build_prod_tower_from_virtual_ext_batch / build_logup_tower_from_virtual_ext_batch (:2365-2402).
 *
 * A spec is described by its LAST layer (the leaves); the library builds all upper layers on the
 * device (infer_tower_product_witness / infer_tower_logup_witness,
 * ceno_zkvm/src/scheme/utils.rs:488-659).
 *   product spec: leaves[0..2) = two ext MLEs of 2^(num_vars-1) elements (low half, high half).
 *   logup spec  : leaves[0..4) = p1, p2, q1, q2 of 2^num_vars ext; p1 = p2 = NULL means
 *                 numerators are all one (utils.rs:556-577). */
typedef struct cg_tower_spec {
    const uint64_t* leaves[4]; /* device pointers */
    uint32_t num_vars;         /* product: layers = num_vars; logup: layers = num_vars + 1 */
    uint32_t is_logup;
} cg_tower_spec;
/* interleaving_mles_to_mles (ceno_zkvm/src/scheme/utils.rs:402-462; the reference GPU path keeps this virtual,
 * GpuVirtualInterleavedExt, ceno_zkvm/src/scheme/gpu/mod.rs:2195-2268): the chip's R record MLEs (one value per
 * instance, base or ext, all of length mles[0].len <= next_pow2(num_instances)) become `num_limbs` (= fan-in, 2)
 * tower leaves of cg_tower_interleave_out_len() ext each, written consecutively to d_out_ext:
 *   out[limb][s * 2^ceil_log2(R) + i] = mle_i[limb * per_fanin_len + s],  everything else = default
 * (1 for read/write records, the challenge alpha for lookup records). */
uint64_t cg_tower_interleave_out_len(uint32_t n_mles, uint64_t num_instances, uint32_t num_limbs);
int cg_tower_interleave(cg_ctx* ctx, const cg_mle_desc* mles, uint32_t n_mles, uint64_t num_instances, uint32_t num_limbs,
                        const uint64_t default_ext[2], uint64_t* d_out_ext, cg_stream s);
typedef struct cg_tower cg_tower;
/* Virtual leaf layers — the reference's GpuVirtualInterleavedExt (ceno_zkvm/src/scheme/gpu/mod.rs:2195-2268; builders
 * build_prod_tower_from_virtual_ext_batch / build_logup_tower_from_virtual_ext_batch :2365-2402): a spec is given by its RECORD
 * MLEs (one value per instance, base or ext, equal length <= next_pow2(num_instances)); the interleaved fan-in leaves
 * (cg_tower_interleave's output: 2^ceil_log2(R) x rows, e.g. 2^33 ext for keccak's 1094 lookup records at 2^22 rows) are never
 * materialised — the first build level and rounds 0 / 1 of the leaf-layer sumcheck read the records through the description.
 * Same proof bits as cg_tower_interleave + cg_tower_build.  At most 8 product + 4 logup specs (the specialised tower kernels).
 *   product spec: q = the read (or write) records, default 1;   logup spec: q = denominators (default alpha),
 *   p = numerators (default 1) or p.n_records = 0 for all-one numerators (utils.rs:556-577). */
typedef struct cg_tower_vgroup {
    const cg_mle_desc* records;
    uint32_t n_records;
    uint32_t reserved;
    uint64_t num_instances;
    uint64_t default_ext[2];
} cg_tower_vgroup;
typedef struct cg_tower_vspec {
    cg_tower_vgroup q, p;
    uint32_t is_logup;
    uint32_t reserved;
} cg_tower_vspec;
int cg_tower_build_virtual(cg_ctx* ctx, const cg_tower_vspec* specs, uint32_t n_specs, cg_stream s, cg_tower** out);
/* SHARDED towers (BASELINE config #4: a chip's rows sliced over the GPUs of one box; new — the reference is single-device).
 * Rank r passes its slice r of BOTH fan-in halves of every leaf array (cg_tower_build_sharded: leaves[] point to the local
 * slices, num_vars is the GLOBAL one) or its slice of the record MLEs with the two fan-in row blocks back to back
 * (cg_tower_build_virtual_sharded: exactly a chip of rows / nranks rows).  A layer stays sliced while every rank keeps
 * >= 2^12 entries per array; the layer kernels store their products straight into the owning partner ranks' peer-mapped
 * buffers over NVLink (results [0, n/2) to rank 2r mod N, [n/2, n) to rank 2r+1 mod N: the rank that needs a slice of both
 * halves of the next layer is not the one that computed it), small layers are all-gathered and replicated.
 * cg_tower_create_proof then runs the big layers' sumchecks sharded (in-kernel exchange of the round sums, early all-gather
 * into the replicated cluster tail) and the small ones replicated; the proof is identical on every rank and equal to the
 * single-device proof.  Needs a peer arena (cg_comm_arena_create / _connect, same size on every rank, >= the sliced layers). */
int cg_comm_arena_create(cg_comm* comm, size_t bytes, uint8_t handle_out[64]);
int cg_comm_arena_connect(cg_comm* comm, const uint8_t* all_handles /* nranks * 64 bytes, rank order */);
int cg_tower_build_sharded(cg_ctx* ctx, cg_comm* comm, const cg_tower_spec* specs, uint32_t n_specs, cg_stream s, cg_tower** out);
int cg_tower_build_virtual_sharded(cg_ctx* ctx, cg_comm* comm, const cg_tower_vspec* specs, uint32_t n_specs, cg_stream s, cg_tower** out);
int cg_tower_build(cg_ctx* ctx, const cg_tower_spec* specs, uint32_t n_specs, cg_stream s, cg_tower** out);
/* get_output_evals (ceno_zkvm/src/scheme/gpu/mod.rs:369-420): layer-0 values of spec i:
 * 2 ext for a product spec, 4 for a logup spec. */
int cg_tower_output_evals(cg_tower* tw, uint32_t spec, uint64_t* h_out_ext);
/* Transcript hooks for the tower (SURVEY §A2 order).  sample(label) must absorb the label and
 * return one ext challenge; append_exts absorbs final evaluations; sumcheck_begin absorbs
 * (num_vars, degree) exactly as IOPProverState::prove does before its first round. */
typedef struct cg_transcript_vt {
    void* user;
    void (*sample)(void* user, const char* label, uint64_t out_ext[2]);
    void (*append_exts)(void* user, const uint64_t* ext, uint64_t n);
    void (*sumcheck_begin)(void* user, uint64_t num_vars, uint64_t degree);
    cg_challenge_cb round_challenge;
} cg_transcript_vt;
/* Returns the proof flattened the way TowerProofs stores it per round (round = 1..max):
 *   round*3 ext sumcheck messages, then 2 ext per live product spec, 4 ext per live logup spec.
 * h_proof must hold cg_tower_proof_len() u64; h_point receives the final point (max_round+1 ext). */
uint64_t cg_tower_proof_len(const cg_tower* tw);
uint32_t cg_tower_point_len(const cg_tower* tw);
int cg_tower_create_proof(cg_tower* tw, const cg_transcript_vt* tr, uint64_t* h_proof, uint64_t* h_point);
int cg_tower_destroy(cg_tower* tw);
/* ready-made vtable functions over a stand-in state (user = uint64_t* state) */
void cg_standin_vt(uint64_t* state, cg_transcript_vt* out);

/* ---- point-wise layer-output inference: wit_infer_by_monomial_expr
 * (gkr_iop/src/gpu/mod.rs:599-609; CPU gkr_iop/src/cpu/mod.rs:119-176):
 * out[b] = sum_t coeff_t prod_{i in t} mle_i[b], ext output of 2^num_vars elements. */
int cg_wit_infer_by_monomial_expr(cg_ctx* ctx, const cg_mle_desc* mles, uint32_t n_mles,
                                  const uint64_t* term_coeff_ext, const uint32_t* term_offsets,
                                  const uint32_t* term_mle_idx, uint32_t n_terms, uint32_t num_vars,
                                  uint64_t* d_out_ext, cg_stream s);

/* ---- rotation pre-passes: rotation_next_base_mle / rotation_selector
 * (gkr_iop/src/utils.rs:19-76); replace rotation_next_base_mle_gpu / rotation_selector_gpu
 * (gkr_iop/src/gkr/layer/gpu/utils.rs:231-336).  The BooleanHypercube cyclic order (5 or 6 variables,
 * gkr_iop/src/gkr/booleanhypercube.rs) is applied inside every chunk of 2^cyclic_group_log2 elements. */
int cg_rotation_next_base_mle(cg_ctx* ctx, const cg_mle_desc* base_mle, uint32_t cyclic_group_log2, uint64_t* d_out_base, cg_stream s);
int cg_rotation_selector(cg_ctx* ctx, const uint64_t* d_eq_ext, uint64_t total_len, uint32_t cyclic_subgroup_size,
                         uint32_t cyclic_group_log2, uint64_t* d_out_ext, cg_stream s);

/* ---- f-3: EC-sum Quark pre-passes (CpuEccProver::create_ecc_proof, ceno_zkvm/src/scheme/cpu/mod.rs:72-316; the
 * zerocheck itself is a degree-3 cg_sumcheck_prove over the monomial terms of the septic-extension constraints).
 * cg_ecc_quark_selectors: for out_rt (n ext, host) writes three ext MLEs of 2^n entries:
 *   sel_add    = SelectorType::QuarkBinaryTreeLessThan.compute(out_rt, {offset 0, num_instances, n})   (:100-107)
 *   sel_export = one-hot at index 2^n - 2 with value eq_eval(out_rt, (0,1,..,1))                       (:109-117)
 *   sel_bypass = eq(out_rt, .) zeroed wherever sel_add != 0 and at the last index                      (:119-133)
 * cg_split_even_odd: filter_bj (:138-152): even[i][b] = mle_i[2b], odd[i][b] = mle_i[2b+1] (base MLEs; x[b,0] / x[b,1]).
 * x[1,b] = as_view_slice(2, 1) is the second half of the same buffer: a pointer offset, no call needed. */
int cg_ecc_quark_selectors(cg_ctx* ctx, const uint64_t* h_out_rt_ext, uint32_t num_vars, uint64_t num_instances,
                           uint64_t* d_sel_add_ext, uint64_t* d_sel_bypass_ext, uint64_t* d_sel_export_ext, cg_stream s);
int cg_split_even_odd(cg_ctx* ctx, const cg_mle_desc* mles, uint32_t n_mles, uint64_t* const* d_even, uint64_t* const* d_odd, cg_stream s);
/* Host-only: the monomial term table of the EC-sum Quark zerocheck (cpu/mod.rs:153-262: add / bypass / export constraint
 * families under their selectors, septic products expanded by z^7 = 2z + 5) over the MLE order
 * [sel_add, sel_bypass, sel_export, s(7), x0(7), y0(7), x1(7), y1(7), x3(7), y3(7)], in the layout cg_sumcheck_* take
 * (coeff: 2 u64 per term, off: n_terms + 1 prefix offsets, idx: factor lists).  alpha_pows: 49 ext; final_x / final_y: the 7 + 7
 * base limbs of the exported sum.  Call with NULL outputs to get the sizes (260 terms, 717 factors for generic alphas). */
int cg_ecc_quark_terms(const uint64_t* alpha_pows_ext, const uint64_t* final_x, const uint64_t* final_y, uint64_t* coeff_out,
                       uint32_t* off_out, uint32_t* idx_out, uint32_t cap_terms, uint32_t cap_idx, uint32_t* n_terms, uint32_t* n_idx);

/* ---- kernel (iii-commit): Merkle commitment over Poseidon2-Goldilocks (TraceCommitter::commit_traces ->
 * PCS::batch_commit, ceno_zkvm/src/scheme/cpu/mod.rs:559-584; GPU basefold.batch_commit_*,
 * ceno_zkvm/src/scheme/gpu/mod.rs:1062-1509).  PARITY UNPINNED: Poseidon2 round constants, the internal diagonal
 * and Basefold's code/leaf arrangement are defined only in un-vendored crates, so the
 * constants are supplied by the caller (the Rust side has them) and the layout offered is the plain
 * Plonky3 one: leaf = PaddingFreeSponge<8, rate 4, out 4> over one matrix row, node = TruncatedPermutation.
 * RS-encoding of the columns is not included. */
typedef struct cg_poseidon2_params {
    uint64_t ext_rc[8][8];   /* external round constants: rounds 0-3 initial, 4-7 terminal */
    uint64_t int_rc[22];     /* internal round constants (lane 0) */
    uint64_t diag[8];        /* internal layer: state[i] = state[i] * diag[i] + sum(state) */
    uint32_t mds_variant;    /* 0: circ(2,3,1,1)  1: Horizen-Labs M4 */
    uint32_t pad;
} cg_poseidon2_params;
int cg_poseidon2_set_params(cg_ctx* ctx, const cg_poseidon2_params* params);
int cg_poseidon2_permute(cg_ctx* ctx, uint64_t* d_states /* n x 8 */, uint64_t n, cg_stream s);
/* d_matrix: height x width base elements (height a power of two), column-major (col_major != 0, what the
 * reference keeps on the device after matrix_transpose) or row-major.  d_tree receives 2*height-1 digests of
 * 4 u64 (leaf level first, root last); h_root (optional) the root. */
int cg_merkle_commit(cg_ctx* ctx, const uint64_t* d_matrix, uint64_t width, uint64_t height, int col_major,
                     uint64_t* d_tree, uint64_t h_root[4], cg_stream s);

/* ---- a9 / f-2: Reed-Solomon encoding of witness columns = batched radix-2 NTT over Goldilocks (the encode step of
 * PCS::batch_commit, EXTERNAL mpcs::Basefold over p3-dft; call site ceno_zkvm/src/scheme/cpu/mod.rs:559-584, GPU
 * basefold.batch_commit_*, ceno_zkvm/src/scheme/gpu/mod.rs:1062-1509).  The transform is p3's:
 * X[k] = sum_j x[j] w^(jk), w = two_adic_generator(log_n) = g^(2^(32-log_n)), g = 7^((p-1)/2^32).
 * PARITY UNPINNED for Basefold's arrangement (rate, basecode size, leaf order live in un-vendored mpcs):
 * rate_log and the output order are parameters.  log_n (+ rate_log) <= 27.
 * cg_ntt: in place, n_cols columns of 2^log_n elements, column c at d_data + c*col_stride (elements).
 *   CG_NTT_INVERSE: inverse transform (includes 1/n).
 *   CG_NTT_BITREV : forward writes / inverse reads bit-reversed order (the fast path: no permutation pass; it is the
 *                   order in which a folding prover pairs adjacent entries).
 *   CG_NTT_EXT    : elements are ext ([c0,c1]); both limb arrays are transformed.
 * cg_rs_encode: every column of the column-major message matrix (width x 2^log_n base elements, an MLE's evaluation
 *   vector taken as coefficients) is zero-padded to 2^(log_n+rate_log) and transformed into d_code (width x 2^(log_n+rate_log),
 *   column-major: directly what cg_merkle_commit hashes row-wise). */
#define CG_NTT_INVERSE 1u
#define CG_NTT_BITREV 2u
#define CG_NTT_EXT 4u
int cg_ntt(cg_ctx* ctx, uint64_t* d_data, uint32_t log_n, uint64_t n_cols, uint64_t col_stride, uint32_t flags, cg_stream s);
int cg_rs_encode(cg_ctx* ctx, const uint64_t* d_msg, uint64_t width, uint32_t log_n, uint32_t rate_log, uint64_t* d_code,
                 uint32_t flags, cg_stream s);

/* ---- f-2 / a9: Basefold PCS — commit and batch_open.
 * Replaces TraceCommitter::commit_traces -> PCS::batch_commit (ceno_zkvm/src/scheme/cpu/mod.rs:559-584; GPU
 * basefold.batch_commit_*, ceno_zkvm/src/scheme/gpu/mod.rs:1062-1509) and OpeningProver::open -> PCS::batch_open
 * (ceno_zkvm/src/scheme/cpu/mod.rs:1415-1457; GPU ceno_zkvm/src/scheme/gpu/mod.rs:3324-3413).  The protocol itself is in the
 * un-vendored `mpcs` crate; what is implemented is exactly what the in-tree verifier restatement accepts
 * (ceno_recursion_v2/src/pcs/mod.rs:1111-1317 replay_basefold, :7494-7727 query checks, :7765-7781 fold rule, :444-592
 * final claim, :138-145 basecode_log == 0):
 *   commitment  : every column (an MLE's 2^num_vars evaluations = the message coefficients) RS-encoded at rate 2^-rate_log,
 *                 codeword rows in bit-reversed order, leaf = PaddingFreeSponge over the row of all columns, 2-to-1 compression;
 *   batch_open  : batch coefficients 1, a, a^2, ... (label "batch coeffs"); per round: degree-2 sumcheck message
 *                 [p(1), p(2)], label "commit round", challenge, Merkle commitment of the running codeword as (even, odd) ext
 *                 pairs (observed after the challenge), fold lo=(a+b)/2, hi=(a-b) g^-bitrev(i)/2, lo + r (hi-lo); smaller
 *                 codewords join at their height; final message = one ext per opening; proof of work; label
 *                 "query indices"; per query one opened row + path per commitment and one sibling + path per round.
 * PARITY UNPINNED for rate_log, the number of queries, proof-of-work bits (parameters), the Poseidon2 constants and the
 * duplex challenger (behind the vtable), and mixed-height commitments (one matrix per commitment, like the restatement).
 * Proof layout (u64 words, cg_basefold_proof_len of them), R = max num_vars, Q = n_queries:
 *   sumcheck R x [p(1).c0 c1 p(2).c0 c1] | commits R x 4 | final message n_openings x 2 | pow witness 1 |
 *   Q x { index 1 | per opening: row (width) , path (num_vars + rate_log) x 4 | per round r: sibling 2, path (R + rate_log - r - 1) x 4 } */
typedef struct cg_basefold_params {
    uint32_t rate_log;   /* BasefoldSpec::get_rate_log()        (upstream; 1 in the tests) */
    uint32_t n_queries;  /* BasefoldSpec::get_number_queries()  (upstream) */
    uint32_t pow_bits;   /* proof-of-work bits before the queries (pcs/mod.rs:1255-1259); 0 = none */
    uint32_t reserved;
} cg_basefold_params;
typedef struct cg_pcs_commitment cg_pcs_commitment;   /* PCS::CommitmentWithWitness: codeword matrix + Merkle tree on the device */
/* d_msg: width x 2^num_vars base elements, column-major (what the reference keeps on the device after matrix_transpose); it
 * is read again by cg_basefold_batch_open and must stay alive until then.  Needs cg_poseidon2_set_params. */
int cg_basefold_commit(cg_ctx* ctx, const uint64_t* d_msg, uint64_t width, uint32_t num_vars, const cg_basefold_params* params,
                       cg_stream s, cg_pcs_commitment** out);
int cg_basefold_commitment_root(const cg_pcs_commitment* cm, uint64_t h_root[4]);
/* device pointers of the codeword matrix (width x 2^(num_vars+rate_log), column-major, bit-reversed rows) and the tree */
int cg_basefold_commitment_codeword(const cg_pcs_commitment* cm, const uint64_t** d_code, const uint64_t** d_tree);
int cg_basefold_commitment_free(cg_pcs_commitment* cm);
/* Transcript events of the PCS (the reference hands &mut impl Transcript<E> to batch_open): */
typedef struct cg_pcs_transcript_vt {
    void* user;
    void (*observe_label)(void* user, const char* label);
    void (*sample_ext)(void* user, uint64_t out_ext[2]);
    void (*observe_exts)(void* user, const uint64_t* ext, uint64_t n);
    void (*observe_base)(void* user, const uint64_t* base, uint64_t n);   /* Merkle digests */
    uint64_t (*sample_bits)(void* user, uint32_t bits);
    uint64_t (*grind)(void* user, uint32_t bits);                          /* returns the proof-of-work witness */
} cg_pcs_transcript_vt;
void cg_standin_pcs_vt(uint64_t* state, cg_pcs_transcript_vt* out);       /* over a stand-in state (NOT Poseidon2) */
typedef struct cg_basefold_opening {
    const cg_pcs_commitment* commit;
    const uint64_t* h_point_ext;   /* num_vars ext (host) */
    const uint64_t* h_evals_ext;   /* width ext (host): the claimed evaluations of the columns at the point */
} cg_basefold_opening;
uint64_t cg_basefold_proof_len(const cg_basefold_opening* ops, uint32_t n_openings, const cg_basefold_params* params);
int cg_basefold_batch_open(cg_ctx* ctx, const cg_basefold_opening* ops, uint32_t n_openings, const cg_basefold_params* params,
                           const cg_pcs_transcript_vt* tr, uint64_t* h_proof, uint64_t proof_cap_words, cg_stream s);

/* ---- f-4: chip-level concurrency — ChipScheduler::execute (ceno_zkvm/src/scheme/scheduler.rs:109-400,
 * docs/src/concurrent-chip-proving.md).  Greedy backfilling over 1..8 lanes (0 = the reference's default, 4), one OS
 * thread + one non-default stream per lane: tasks are sorted by estimated memory (descending), the first pending task
 * whose booking fits the remaining budget is launched while a lane is free, the scheduler blocks on completions when
 * nothing fits, and reports "Deadlock: Remaining tasks are too big for the memory pool" (CG_ERR_OOM) when nothing
 * fits and nothing runs.  A lane (and its booking) is released only after the lane's stream has drained.  The
 * callback does the chip's work (transcript fork + cg_tower_* / cg_sumcheck_* calls) on the stream it is given and
 * returns a CG_* status; the first failure is returned after in-flight tasks have finished, tasks never started get
 * status CG_ERR_STATE.  results[] (n_tasks entries) comes back ordered by task_id.  ctx may be NULL for host-only
 * use (then stream is NULL in the callback and mem_budget_bytes is required); mem_budget_bytes = 0 means the
 * device memory free now plus the pool's idle blocks. */
#define CG_SCHED_DEFAULT_LANES 4u
#define CG_SCHED_MAX_LANES 8u
typedef struct cg_sched_task {
    uint32_t task_id;                /* result ordering */
    uint32_t reserved;
    uint64_t estimated_memory_bytes; /* sort key */
    uint64_t booked_memory_bytes;    /* what the scheduler reserves (0 = estimated_memory_bytes) */
} cg_sched_task;
typedef struct cg_sched_result {
    uint32_t task_id;
    uint32_t lane_id;
    int32_t status;
    uint32_t launch_seq;             /* order in which the scheduler admitted the task */
    uint64_t booked_total_at_launch; /* bytes booked right after admission (never above the budget) */
    double queue_delay_ms, host_execution_ms, event_wait_ms;
} cg_sched_result;
typedef int (*cg_sched_fn)(void* user, uint32_t task_index, uint32_t task_id, uint32_t lane_id, cg_stream stream);
int cg_sched_execute(cg_ctx* ctx, const cg_sched_task* tasks, uint32_t n_tasks, uint32_t lanes, uint64_t mem_budget_bytes,
                     cg_sched_fn fn, void* user, cg_sched_result* results);
/* a non-blocking stream for a caller-owned lane thread (get_pool_stream / bind_thread_stream, gkr_iop/src/gpu/mod.rs:79-154) */
int cg_stream_create(cg_ctx* ctx, cg_stream* out);
int cg_stream_destroy(cg_ctx* ctx, cg_stream s);

#ifdef __cplusplus
}
#endif
#endif /* CENO_B200_H */
