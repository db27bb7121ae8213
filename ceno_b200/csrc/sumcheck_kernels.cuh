// sumcheck_kernels.cuh — the hot kernels of the GKR-sumcheck path, hand-written for sm_100a.
//
//   (i)  round evaluation  : [p(1)..p(d)] = sum over the remaining hypercube pairs
//   (ii) fix_variable fold : f'[b] = f[2b] + r (f[2b+1] - f[2b])
//   fused (ii)+(i)         : one pass per round — reads 4 consecutive elements per MLE, writes the
//                            folded pair, accumulates the next round's evaluations from registers.
//   (iii) build_eq_x_r     : two-level outer product L[b_lo] * H[b_hi] with the selector mask fused.
//
// Memory plan (HBM-bound integer streaming; no tensor-core path): every thread issues 256-bit
// global loads (LDG.E.ENL2.256) of consecutive elements, a warp covers one contiguous 2 KB span per
// MLE, stores are 256-bit; grids are sized in multiples of the SM count with a grid-stride loop;
// reductions are warp-shuffle -> shared -> one partial per block -> last-block finish (ticket).
// Field arithmetic never leaves registers (gl64.cuh).
//
// Reference semantics: IOPProverState::prove round loop (external sumcheck crate; SURVEY.md §8a1,
// §A1-A3), call sites gkr_iop/src/gkr/layer/cpu/mod.rs:217-237, ceno_zkvm/src/scheme/cpu/mod.rs:490-498.
#pragma once
#include "gl64.cuh"

#define CG_THREADS 256
#define CG_MAX_DEGREE 8
#define CG_MAX_BLOCKS 2048
#define CG_TOWER_MAX_PROD 8
#define CG_TOWER_MAX_LOGUP 4

// ---------------------------------------------------------------------------------------------
// Multi-GPU exchange over NVLink peer memory: every rank owns a mailbox buffer that its
// peers map through CUDA IPC.  The last block of a round kernel stores its partial round sums into
// slot [ring][my_rank] of EVERY peer's mailbox (plain P2P stores + system fence + sequence flag),
// waits for the N flags in its own mailbox and adds the N partials mod p.  All ranks obtain the same
// message, so the transcript runs replicated and nothing is broadcast.  A fast rank can be at most one
// exchange ahead of a slow one (it needs the slow rank's next partial to finish), so a ring of 4 is safe.
#define CG_MAX_RANKS 8
#define CG_COMM_RING 4
#define CG_COMM_GATHER_MLES 64
struct __align__(32) CommSlot {       // payload words w[0 .. 2D), validation flag at w[2D]; 32-byte chunks
    uint64_t w[2 * CG_MAX_DEGREE + 8];
};
#define CG_TAIL_MAX_N 4096      // register staging of the in-place fold is sized for this
#define CG_TAIL_START_N 2048    // enter the tail once each (global) MLE has <= this many elements (one SM: ~2 items/thread)
#define CG_COMM_GATHER_SLOTS 33 // 1 + 2*8 + 4*4 MLE slots of the tower layout
struct CommGather {
    ext_t v[2][CG_COMM_GATHER_MLES][CG_MAX_RANKS];                 // fallback path: final local evaluations
    uint64_t seq[2][CG_MAX_RANKS];
    ext_t big[2][1u << 18];                                        // tail entry (CG_GATHER_EXT): every rank's folded slice, [slot][n0]
    uint64_t big_seq[2][CG_MAX_RANKS];
};
struct CommBuf {
    CommSlot slots[CG_COMM_RING][CG_MAX_RANKS];
    CommGather gather;
    uint64_t bar[CG_MAX_RANKS];        // rank barrier: bar[p] = last barrier sequence rank p has reached
};
struct CommDev {
    int rank, nranks;               // nranks <= 1: no exchange
    uint64_t seq;                   // sequence number of this launch's (first) exchange
    CommBuf* peers[CG_MAX_RANKS];   // peers[rank] is the local buffer
    int* d_error;
    unsigned long long timeout_cycles;
    unsigned long long* dbg;        // optional: [seq % 1024][4] = {wait_ns(lane 0), wait_ns(lane 1), polls(lane 1), rank}
};
// Slot validation without a writer-side fence: the flag word is seq mixed with a checksum of the
// payload, so a reader that sees the flag before (part of) the payload simply keeps polling.
GL_DEV uint64_t comm_checksum(uint64_t seq, const uint64_t* v, int n) {
    uint64_t h = seq * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL;
    for (int i = 0; i < n; i++) h = (h ^ v[i]) * 0xD6E8FEB86659FD93ULL + (uint64_t)i;
    return h ^ (h >> 29);
}
// executed by a whole converged warp; the local partial is in lane 0, the combined sum returns in lane 0.
// The message (payload + flag) is staged in shared memory and leaves as ONE warp-wide instruction of
// 32-byte stores (lane = peer x chunk): peer stores cost ~2 us of issue stall each, so the number of
// store instructions — not bytes — is what matters.
template <int D>
GL_DEV void comm_exchange(ext_t (&res)[D], const CommDev& cm, uint64_t seq, uint64_t* s_msg /* >= 2D+4 words, 32-B aligned */) {
    const int lane = threadIdx.x & 31;
    constexpr int CHUNKS = (2 * D + 1 + 3) / 4;
    if (lane == 0) {
        uint64_t v[2 * D];
#pragma unroll
        for (int x = 0; x < D; x++) { v[2 * x] = res[x].c0; v[2 * x + 1] = res[x].c1; s_msg[2 * x] = res[x].c0; s_msg[2 * x + 1] = res[x].c1; }
        s_msg[2 * D] = comm_checksum(seq, v, 2 * D);
#pragma unroll
        for (int i = 2 * D + 1; i < 4 * CHUNKS; i++) s_msg[i] = 0;
    }
    __syncwarp();
    const int ring = (int)(seq % CG_COMM_RING);
    for (int item = lane; item < cm.nranks * CHUNKS; item += 32) {
        const int p = item / CHUNKS, ch = item % CHUNKS;
        const uint64_t* src = s_msg + 4 * ch;
        uint64_t* dst = &cm.peers[p]->slots[ring][cm.rank].w[4 * ch];
        asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(src[0]), "l"(src[1]), "l"(src[2]), "l"(src[3]) : "memory");
    }
    ext_t got[D];
#pragma unroll
    for (int x = 0; x < D; x++) got[x] = ext_zero();
    if (lane < cm.nranks) {
        const CommSlot* src = &cm.peers[cm.rank]->slots[ring][lane];
        const long long t0 = clock64();
        unsigned long long g0 = 0, polls = 0;
        if (cm.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        uint64_t w[4 * CHUNKS];
        while (true) {
#pragma unroll
            for (int ch = 0; ch < CHUNKS; ch++)
                asm volatile("ld.volatile.global.v4.u64 {%0, %1, %2, %3}, [%4];"
                             : "=l"(w[4 * ch]), "=l"(w[4 * ch + 1]), "=l"(w[4 * ch + 2]), "=l"(w[4 * ch + 3]) : "l"(&src->w[4 * ch]) : "memory");
            polls++;
            if (w[2 * D] == comm_checksum(seq, w, 2 * D)) {   // stale or torn contents fail and are re-read
#pragma unroll
                for (int x = 0; x < D; x++) got[x] = ext_make(gl_canon(w[2 * x]), gl_canon(w[2 * x + 1]));
                if (cm.dbg && lane < 2) {
                    unsigned long long g1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
                    cm.dbg[(seq % 1024) * 4 + lane] = g1 - g0;
                    if (lane == 1) { cm.dbg[(seq % 1024) * 4 + 2] = polls; cm.dbg[(seq % 1024) * 4 + 3] = g0; }
                }
                break;
            }
            if ((unsigned long long)(clock64() - t0) > cm.timeout_cycles) { *cm.d_error = 2; break; }
        }
    }
#pragma unroll
    for (int x = 0; x < D; x++) res[x] = warp_reduce_ext(got[x]);
}
// rank barrier over the mailboxes (sharded tower build: every rank's peer stores of a level are complete before the next
// level reads them).  Launched after the kernels whose stores it orders; the system fence publishes them.
__global__ void comm_barrier_kernel(const __grid_constant__ CommDev cm, uint64_t seq) {
    __threadfence_system();
    if ((int)threadIdx.x < cm.nranks) {
        *(volatile uint64_t*)&cm.peers[threadIdx.x]->bar[cm.rank] = seq;
        volatile uint64_t* f = &cm.peers[cm.rank]->bar[threadIdx.x];
        const long long t0 = clock64();
        while (*f < seq) {
            if ((unsigned long long)(clock64() - t0) > cm.timeout_cycles) { *cm.d_error = 2; break; }
        }
        __threadfence_system();
    }
}
// all-gather of the m final local evaluations into every rank's gather area [par][mle][rank]
__global__ void comm_allgather_kernel(const ext_t* __restrict__ d_final, int m, const __grid_constant__ CommDev cm, int par) {
    const int tid = threadIdx.x;
    for (int p = 0; p < cm.nranks; p++)
        for (int i = tid; i < m; i += blockDim.x) cm.peers[p]->gather.v[par][i][cm.rank] = d_final[i];
    __threadfence_system();
    __syncthreads();
    if (tid < cm.nranks) {
        *(volatile uint64_t*)&cm.peers[tid]->gather.seq[par][cm.rank] = cm.seq;
        volatile uint64_t* f = &cm.peers[cm.rank]->gather.seq[par][tid];
        const long long t0 = clock64();
        while (*f != cm.seq) {
            if ((unsigned long long)(clock64() - t0) > cm.timeout_cycles) { *cm.d_error = 2; break; }
        }
        __threadfence_system();
    }
}

// Host mailbox in mapped pinned memory: the device posts a round message, the host answers with the
// challenge (tail / mid kernels) or simply launches the next round (per-round kernels) — no D2H copy + stream sync.
// Latency protocol (one PCIe round trip per round, no system fences on the critical path):
//   device -> host: the message words, then flag = (round + 1) ^ mix(words).  The host accepts the message when the flag
//                   matches the words it read, so payload and flag need no ordering fence between them;
//   host -> device: {r0, r1, seq_r, abort} is ONE 32-byte line the device polls with a single 256-bit load: when seq_r
//                   matches, the challenge arrived in the same load (the host writes r before seq_r, x86 store order).
struct __align__(64) TailMailbox {
    uint64_t msg[2 * 8];        // device -> host
    volatile uint64_t seq_msg;  // (round index + 1) ^ cg_mb_mix(msg words)
    uint64_t pad0[7];
    uint64_t r[2];              // host -> device: the reply line (32-byte aligned)
    volatile uint64_t seq_r;    // round index + 1 when r[] is valid
    volatile uint64_t abort;    // give up (callback failed)
    // device -> host, once, after the last round: the final evaluations in the caller's MLE order, validated by
    // fin_flag = seq ^ (order-independent checksum) — the host needs neither a D2H copy nor a stream synchronisation
    uint64_t fin[2 * 33];       // CG_COMM_GATHER_SLOTS ext
    volatile uint64_t fin_flag;
};
GL_HD uint64_t cg_mb_word(uint64_t v, uint32_t pos) {
    uint64_t h = (v ^ (0x9E3779B97F4A7C15ULL * (pos + 1))) * 0xBF58476D1CE4E5B9ULL;
    return h ^ (h >> 29);
}
GL_HD uint64_t cg_mb_mix(const uint64_t* m, uint32_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (uint32_t i = 0; i < n; i++) { h = (h ^ m[i]) * 0xBF58476D1CE4E5B9ULL; h ^= h >> 31; }
    return h;
}
template <int D>
GL_DEV void mailbox_post(TailMailbox* mb, const ext_t (&res)[D], uint64_t seq) {
    uint64_t w[2 * D];
#pragma unroll
    for (int x = 0; x < D; x++) { w[2 * x] = res[x].c0; w[2 * x + 1] = res[x].c1; }
    volatile uint64_t* m = mb->msg;
#pragma unroll
    for (int x = 0; x < 2 * D; x++) m[x] = w[x];
    mb->seq_msg = seq ^ cg_mb_mix(w, 2 * D);
}
// 1: challenge ready (r set, canonical), -1: host asked to abort, 0: not yet
GL_DEV int mailbox_poll(const TailMailbox* mb, uint64_t seq, ext_t& r) {
    unsigned long long r0, r1, sq, ab;
    asm volatile("ld.volatile.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r0), "=l"(r1), "=l"(sq), "=l"(ab) : "l"(mb->r) : "memory");
    if (sq == seq) { r.c0 = gl_canon(r0); r.c1 = gl_canon(r1); return 1; }
    return ab ? -1 : 0;
}

// ---------------------------------------------------------------------------------------------
// round output / finish
struct RoundOut {
    ext_t* partials;          // [CG_MAX_BLOCKS * CG_MAX_DEGREE]
    unsigned int* ticket;     // zero on entry, zero on exit
    ext_t* d_out;             // device: degree ext (this round's message)
    // device-resident stand-in challenger (optional): absorbs the message, squeezes r
    uint64_t* d_tr_state;     // nullptr -> host transcript
    ext_t* d_r_out;           // where to put the challenge for the next launch
    CommDev comm;             // multi-GPU: combine the partial sums of all ranks (nranks <= 1: off)
    TailMailbox* mail;        // host transcript: also post the message to the host mailbox with sequence mail_seq
    uint64_t mail_seq;
};

struct NoXf {
    template <int D>
    GL_DEV void pre(ext_t (&)[D]) const {}
    template <int D>
    GL_DEV void post(ext_t (&)[D]) const {}
};
// xf: transforms applied by lane 0 of the last block to the combined local sums: pre() before the multi-GPU exchange
// (must be linear: the ranks' results are added), post() after it (the split-eq kernels turn their bilinear sums and
// the running claim into the round message there).
template <int D, class XF = NoXf>
GL_DEV void block_finish(ext_t (&acc)[D], const RoundOut& out, const XF xf = XF()) {
    __shared__ ext_t s_part[CG_THREADS / 32][D];   // blockDim.x <= CG_THREADS
    __shared__ __align__(32) uint64_t s_msg[2 * D + 8];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
#pragma unroll
    for (int x = 0; x < D; x++) {
        ext_t v = warp_reduce_ext(acc[x]);
        if (lane == 0) s_part[warp][x] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int x = 0; x < D; x++) {
            ext_t v = lane < n_warps ? s_part[lane][x] : ext_zero();
            v = warp_reduce_ext(v);
            if (lane == 0) out.partials[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * D + x] = v;
        }
        if (lane == 0) {
            __threadfence();
            unsigned t = atomicAdd(out.ticket, 1u);
            s_last = (t == gridDim.x * gridDim.y - 1);
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: sum the per-block partials
#pragma unroll
    for (int x = 0; x < D; x++) {
        ext_t v = ext_zero();
        for (unsigned b = threadIdx.x; b < gridDim.x * gridDim.y; b += blockDim.x) {
            const ulonglong2 p = __ldcg(reinterpret_cast<const ulonglong2*>(&out.partials[(size_t)b * D + x]));
            v = ext_add(v, ext_make(p.x, p.y));
        }
        v = warp_reduce_ext(v);
        if (lane == 0) s_part[warp][x] = v;
    }
    __syncthreads();
    if (warp == 0) {
        ext_t res[D];
#pragma unroll
        for (int x = 0; x < D; x++) {
            ext_t v = lane < n_warps ? s_part[lane][x] : ext_zero();
            res[x] = warp_reduce_ext(v);
        }
        if (lane == 0) xf.pre(res);
        if (out.comm.nranks > 1) comm_exchange<D>(res, out.comm, out.comm.seq, s_msg);
        if (lane == 0) {
            xf.post(res);
#pragma unroll
            for (int x = 0; x < D; x++) out.d_out[x] = res[x];
            if (out.mail) mailbox_post<D>(out.mail, res, out.mail_seq);
            if (out.d_tr_state) {   // stand-in challenger: absorb evals, label "Internal round", squeeze
                uint64_t h = *out.d_tr_state;
                const ext_t r = cg_tr_round<D>(h, res);
                *out.d_tr_state = h;
                *out.d_r_out = r;
            }
            *out.ticket = 0;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pair loader: FOLD = read 4 consecutive ext, fold by r (one reduction per limb, x + d*r accumulated
// unreduced), write the pair; else read 2.  Results are canonical.
template <bool FOLD, bool CANON>
GL_DEV void load_pair(const ext_t* __restrict__ in, ext_t* __restrict__ out, uint64_t item,
                      const extmul_t& r, ext_t& lo, ext_t& hi) {
    if (FOLD) {
        ext_t x0, x1, x2, x3;
        ld_ext2(in + 4 * item, x0, x1);
        ld_ext2(in + 4 * item + 2, x2, x3);
        if (CANON) { x0 = ext_canon(x0); x1 = ext_canon(x1); x2 = ext_canon(x2); x3 = ext_canon(x3); }
        lo = ext_fma_prep(x0, ext_sub(x1, x0), r);
        hi = ext_fma_prep(x2, ext_sub(x3, x2), r);
        st_ext2(out + 2 * item, lo, hi);
    } else {
        ld_ext2(in + 2 * item, lo, hi);
        if (CANON) { lo = ext_canon(lo); hi = ext_canon(hi); }
    }
}

// ---------------------------------------------------------------------------------------------
// Tower-layer round kernel (degree 3):
//   p(X) = sum_b eq(X,b) * ( sum_i alpha_i a_i(X,b) b_i(X,b)
//                          + sum_j [ an_j (p1_j q2_j + p2_j q1_j) + ad_j q1_j q2_j ](X,b) )
// = the expression CpuTowerProver::create_proof builds per layer
//   (ceno_zkvm/src/scheme/cpu/mod.rs:417-485).  T3 (eq*A*B) is n_prod = 1, n_logup = 0.
//
// Evaluation points by subtraction only: nd = lo - hi, f(1) = hi, f(2) = f(1) - nd, f(3) = f(2) - nd.
// Inner sums and the per-thread round sums are kept as unreduced 160-bit accumulators; each is
// reduced once (inner: once per item and point; round sums: once per thread).
// Virtual tower leaves (the reference's GpuVirtualInterleavedExt, ceno_zkvm/src/scheme/gpu/mod.rs:2195-2268): one fan-in limb of
// interleaving_mles_to_mles (ceno_zkvm/src/scheme/utils.rs:402-462) described instead of materialised —
//   leaf[s * 2^l2m + i] = record_i[row_offset + s]  (i < n_records, s < the record kind's row count), else `def`.
// The kernels that read a tower's leaf layer (first level of the build, round 0 / 1 of the leaf-layer sumcheck) take the
// description; nothing of the leaf layer's size is ever written for the records themselves.
struct VirtLeaf {
    const void* const* ptrs;    // device array of n_records record pointers
    const uint32_t* is_ext;     // device array: record i is ext (16 B) or base (8 B)
    uint32_t n_records, l2m;    // 2^l2m = records per instance after padding
    uint64_t row_offset;        // limb * per_fanin_len
    uint64_t cnt_ext, cnt_base; // rows an ext / a base record contributes to this limb (utils.rs:433-456)
    ext_t def;                  // padding value: 1 for read / write records, the challenge alpha for lookups
};
GL_DEV ext_t virt_leaf(const VirtLeaf& v, uint64_t x) {
    const uint32_t i = (uint32_t)(x & ((1ULL << v.l2m) - 1));
    const uint64_t s = x >> v.l2m;
    if (i < v.n_records) {
        const uint32_t e = v.is_ext[i];
        if (s < (e ? v.cnt_ext : v.cnt_base)) {
            if (e) return ext_canon(ld_ext(reinterpret_cast<const ext_t*>(v.ptrs[i]) + v.row_offset + s));
            return ext_make(gl_canon(reinterpret_cast<const uint64_t*>(v.ptrs[i])[v.row_offset + s]), 0);
        }
    }
    return v.def;
}
#define CG_TOWER_SLOTS (1 + 2 * CG_TOWER_MAX_PROD + 4 * CG_TOWER_MAX_LOGUP)
struct TowerArgs {
    const ext_t* eq_in;
    ext_t* eq_out;
    const ext_t* prod_in[CG_TOWER_MAX_PROD][2];
    ext_t* prod_out[CG_TOWER_MAX_PROD][2];
    ext_t alpha_prod[CG_TOWER_MAX_PROD];
    const ext_t* lk_in[CG_TOWER_MAX_LOGUP][4];
    ext_t* lk_out[CG_TOWER_MAX_LOGUP][4];
    ext_t alpha_num[CG_TOWER_MAX_LOGUP];
    ext_t alpha_den[CG_TOWER_MAX_LOGUP];
    int n_prod, n_logup;
    int alpha_one;            // every alpha_prod is 1: skip the alpha multiply
    const VirtLeaf* virt[CG_TOWER_SLOTS];   // VIRT kernels: slot (1 + 2p + z | 1 + 2 n_prod + 4l + z) read through a description (null: plain array)
    uint32_t pone_mask;       // split-eq VIRT launches: logup spec l's numerators are the constant one (no records: utils.rs:556-577)
    uint64_t n_pairs;         // pairs evaluated this launch (after the fold, if FOLD)
    ext_t r;                  // fold challenge (FOLD only) ...
    const ext_t* r_ptr;       // ... or read it from device memory (device challenger)
    RoundOut out;
};

// H[t] += u * e for the three points; e given as (value at t=1, nd)
GL_DEV void accumulate_point(ecacc& H, ext_t u, ext_t e) {
    eacc T;
    eacc_mul(T, u, e, gl_mul7_weak(e.c1));
    ecacc_add(H, T);   // long-lived sums stay compact (5 limbs); the 4 products accumulate in aligned limb pairs
}

// One hypercube pair of the tower-layer polynomial: H[t] += eq(t) * inner(t), t = 1, 2, 3.
// L supplies the (lo, hi) pair of every MLE: from global memory with the fused fold (GlobalLoader) or
// from shared memory (SmemLoader, tail kernel).
// SIMPLE = exactly one product spec with alpha = 1 and no logup spec (the T3 shape): no spec loops.
template <bool SIMPLE, class L>
GL_DEV void tower_item(const TowerArgs& a, L& ld, uint64_t item, ecacc (&H)[3]) {
    ext_t u[3];   // inner value at t = 1, 2, 3 (weak)
    if (SIMPLE) {
        ext_t alo, av, blo, bv;
        ld.prod(0, 0, item, alo, av);
        ld.prod(0, 1, item, blo, bv);
        const ext_t and_ = ext_sub(alo, av), bnd = ext_sub(blo, bv);
#pragma unroll
        for (int t = 0; t < 3; t++) {
            u[t] = ext_mul_weak(av, bv);
            if (t < 2) { av = ext_sub(av, and_); bv = ext_sub(bv, bnd); }
        }
    } else {
        ecacc in[3];
        ecacc_zero(in[0]); ecacc_zero(in[1]); ecacc_zero(in[2]);
        for (int p = 0; p < a.n_prod; p++) {
            ext_t alo, av, blo, bv;
            ld.prod(p, 0, item, alo, av);
            ld.prod(p, 1, item, blo, bv);
            if (!a.alpha_one) {   // fold alpha into a (2 muls) instead of into the 3 products
                const extmul_t al = extmul_prep(a.alpha_prod[p]);
                alo = ext_mul_prep(alo, al);
                av = ext_mul_prep(av, al);
            }
            const ext_t and_ = ext_sub(alo, av), bnd = ext_sub(blo, bv);
#pragma unroll
            for (int t = 0; t < 3; t++) {
                eacc T;
                eacc_mul(T, av, bv, gl_mul7_weak(bv.c1));
                ecacc_add(in[t], T);
                if (t < 2) { av = ext_sub(av, and_); bv = ext_sub(bv, bnd); }
            }
        }
        for (int l = 0; l < a.n_logup; l++) {
            ext_t p1lo, p1, p2lo, p2, q1lo, q1, q2lo, q2;
            ld.lk(l, 0, item, p1lo, p1);
            ld.lk(l, 1, item, p2lo, p2);
            ld.lk(l, 2, item, q1lo, q1);
            ld.lk(l, 3, item, q2lo, q2);
            const extmul_t an = extmul_prep(a.alpha_num[l]), adn = extmul_prep(a.alpha_den[l]);
            const ext_t p1n = ext_sub(p1lo, p1), p2n = ext_sub(p2lo, p2);
            const ext_t q1n = ext_sub(q1lo, q1), q2n = ext_sub(q2lo, q2);
#pragma unroll
            for (int t = 0; t < 3; t++) {
                const uint64_t q1_7 = gl_mul7_weak(q1.c1), q2_7 = gl_mul7_weak(q2.c1);
                eacc N;   // p1 q2 + p2 q1, one reduction
                eacc_zero(N);
                eacc_mac(N, p1, q2, q2_7);
                eacc_mac(N, p2, q1, q1_7);
                eacc D;   // q1 q2
                eacc_zero(D);
                eacc_mac(D, q1, q2, q2_7);
                eacc T;
                eacc_zero(T);
                eacc_mac_prep(T, eacc_weak(N), an);
                eacc_mac_prep(T, eacc_weak(D), adn);
                ecacc_add(in[t], T);
                if (t < 2) { p1 = ext_sub(p1, p1n); p2 = ext_sub(p2, p2n); q1 = ext_sub(q1, q1n); q2 = ext_sub(q2, q2n); }
            }
        }
        u[0] = ext_make(cacc_weak(in[0].A0), cacc_weak(in[0].A1));
        u[1] = ext_make(cacc_weak(in[1].A0), cacc_weak(in[1].A1));
        u[2] = ext_make(cacc_weak(in[2].A0), cacc_weak(in[2].A1));
    }
    ext_t elo, ev;
    ld.eq(item, elo, ev);
    const ext_t end_ = ext_sub(elo, ev);
    accumulate_point(H[0], u[0], ev);
    ev = ext_sub(ev, end_);
    accumulate_point(H[1], u[1], ev);
    ev = ext_sub(ev, end_);
    accumulate_point(H[2], u[2], ev);
}

// The same pair evaluated at ONE point X = t + 1 (t = 0, 1, 2) into a single round sum: the latency-bound small rounds of
// the tail kernel give every pair to three threads (one per point), which cuts the dependent chain of an item to a third.
template <bool SIMPLE, class L>
GL_DEV void tower_item_pt(const TowerArgs& a, L& ld, uint64_t item, int t, ecacc& H) {
    auto at = [&](ext_t lo, ext_t hi) -> ext_t {   // f(t + 1) = hi - t (lo - hi)
        const ext_t nd = ext_sub(lo, hi);
        ext_t v = hi;
        if (t >= 1) v = ext_sub(v, nd);
        if (t >= 2) v = ext_sub(v, nd);
        return v;
    };
    ext_t u;
    if (SIMPLE) {
        ext_t alo, ahi, blo, bhi;
        ld.prod(0, 0, item, alo, ahi);
        ld.prod(0, 1, item, blo, bhi);
        u = ext_mul_weak(at(alo, ahi), at(blo, bhi));
    } else {
        ecacc in;
        ecacc_zero(in);
        for (int p = 0; p < a.n_prod; p++) {
            ext_t alo, ahi, blo, bhi;
            ld.prod(p, 0, item, alo, ahi);
            ld.prod(p, 1, item, blo, bhi);
            ext_t av = at(alo, ahi);
            const ext_t bv = at(blo, bhi);
            if (!a.alpha_one) av = ext_mul_prep(av, extmul_prep(a.alpha_prod[p]));
            eacc T;
            eacc_mul(T, av, bv, gl_mul7_weak(bv.c1));
            ecacc_add(in, T);
        }
        for (int l = 0; l < a.n_logup; l++) {
            ext_t lo, hi;
            ld.lk(l, 0, item, lo, hi); const ext_t p1 = at(lo, hi);
            ld.lk(l, 1, item, lo, hi); const ext_t p2 = at(lo, hi);
            ld.lk(l, 2, item, lo, hi); const ext_t q1 = at(lo, hi);
            ld.lk(l, 3, item, lo, hi); const ext_t q2 = at(lo, hi);
            const extmul_t an = extmul_prep(a.alpha_num[l]), adn = extmul_prep(a.alpha_den[l]);
            const uint64_t q1_7 = gl_mul7_weak(q1.c1), q2_7 = gl_mul7_weak(q2.c1);
            eacc N;
            eacc_zero(N);
            eacc_mac(N, p1, q2, q2_7);
            eacc_mac(N, p2, q1, q1_7);
            eacc D;
            eacc_zero(D);
            eacc_mac(D, q1, q2, q2_7);
            eacc T;
            eacc_zero(T);
            eacc_mac_prep(T, eacc_weak(N), an);
            eacc_mac_prep(T, eacc_weak(D), adn);
            ecacc_add(in, T);
        }
        u = ext_make(cacc_weak(in.A0), cacc_weak(in.A1));
    }
    ext_t elo, ehi;
    ld.eq(item, elo, ehi);
    accumulate_point(H, u, at(elo, ehi));
}

// the same pair read through a leaf description (values come back canonical)
template <bool FOLD>
GL_DEV void load_pair_virt(const VirtLeaf& v, ext_t* __restrict__ out, uint64_t item, const extmul_t& r, ext_t& lo, ext_t& hi) {
    if (FOLD) {
        const ext_t x0 = virt_leaf(v, 4 * item), x1 = virt_leaf(v, 4 * item + 1), x2 = virt_leaf(v, 4 * item + 2), x3 = virt_leaf(v, 4 * item + 3);
        lo = ext_fma_prep(x0, ext_sub(x1, x0), r);
        hi = ext_fma_prep(x2, ext_sub(x3, x2), r);
        st_ext2(out + 2 * item, lo, hi);
    } else {
        lo = virt_leaf(v, 2 * item);
        hi = virt_leaf(v, 2 * item + 1);
    }
}
template <bool FOLD, bool CANON, bool VIRT = false>
struct GlobalLoader {
    const TowerArgs& a;
    extmul_t rm;
    GL_DEV void eq(uint64_t item, ext_t& lo, ext_t& hi) { load_pair<FOLD, CANON>(a.eq_in, a.eq_out, item, rm, lo, hi); }
    GL_DEV void prod(int p, int z, uint64_t item, ext_t& lo, ext_t& hi) {
        if (VIRT) {
            const VirtLeaf* v = a.virt[1 + 2 * p + z];
            if (v) { load_pair_virt<FOLD>(*v, a.prod_out[p][z], item, rm, lo, hi); return; }
        }
        load_pair<FOLD, CANON>(a.prod_in[p][z], a.prod_out[p][z], item, rm, lo, hi);
    }
    // a slot whose values are one known constant: the fold's output is the constant
    GL_DEV void store_const(ext_t* out, uint64_t item, ext_t v) { if (FOLD) st_ext2(out + 2 * item, v, v); }
    GL_DEV void lk(int l, int z, uint64_t item, ext_t& lo, ext_t& hi) {
        if (VIRT) {
            const VirtLeaf* v = a.virt[1 + 2 * a.n_prod + 4 * l + z];
            if (v) { load_pair_virt<FOLD>(*v, a.lk_out[l][z], item, rm, lo, hi); return; }
        }
        load_pair<FOLD, CANON>(a.lk_in[l][z], a.lk_out[l][z], item, rm, lo, hi);
    }
};

// THREADS x MINB = launch shape (registers per thread are capped at 65536 / (THREADS * MINB)).
// VIRT: some slots are virtual tower leaves (only the launches that read a sumcheck's ORIGINAL inputs can be).
template <bool FOLD, bool CANON, bool SIMPLE, int THREADS, int MINB, bool VIRT = false>
__global__ void __launch_bounds__(THREADS, MINB) tower_round_kernel(const __grid_constant__ TowerArgs a) {
    GlobalLoader<FOLD, CANON, VIRT> ld{a, {0, 0, 0}};
    if (FOLD) ld.rm = extmul_prep(a.r_ptr ? ld_ext(a.r_ptr) : a.r);
    ecacc H[3];
    ecacc_zero(H[0]); ecacc_zero(H[1]); ecacc_zero(H[2]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; item < a.n_pairs; item += stride)
        tower_item<SIMPLE>(a, ld, item, H);
    ext_t acc[3] = {ecacc_canon(H[0]), ecacc_canon(H[1]), ecacc_canon(H[2])};
    block_finish<3>(acc, a.out);
}

// ---------------------------------------------------------------------------------------------
// Split-eq rounds for a VIRTUAL eq MLE (cg_mle_desc kind CG_MLE_EQ: the caller hands over the point w
// instead of the 2^k table).  With  eq(w, (r_0..r_{j-1}, X, x)) = P_j * eq(w_j, X) * E_j(x),
//   P_j = prod_{i<j} eq(w_i, r_i),   E_j(x) = eq(w_{j+1..k-1}, x) = L_j[x & 255] * H_j[x >> 8],
// the round polynomial is  p_j(X) = P_j eq(w_j, X) q_j(X),  q_j(X) = sum_x E_j(x) A(X, x) B(X, x)  (degree 2).
// The kernel never streams or folds an eq table: per pair it weights A by the row weight H_j[x >> 8]
// (uniform over the block), accumulates the three bilinear sums
//   S00 = sum A'lo Blo,  S11 = sum A'hi Bhi,  Sx = sum (A'lo Bhi + A'hi Blo)
// UNREDUCED and without a single modular subtraction, applies the thread's fixed L_j[tid] once at the
// end, and the last block turns (S00, S11, Sx) into [p_j(1), p_j(2), p_j(3)]:
//   q(0) = S00, q(1) = S11, X^2 coefficient c2 = S00 + S11 - Sx.
// Field arithmetic is exact, so the message is bit-identical to the one computed from a materialised
// eq table (IOPProverState semantics); this is the eq handling of the reference's
// tower/zerocheck sumchecks (ceno_zkvm/src/scheme/cpu/mod.rs:417-485) without the eq stream.
#define CG_VEQ_MAX_ROUNDS 16
#define CG_VEQ_LO_BITS 8
struct VeqFin {
    const ext_t* w;        // device: the point (num_vars ext), followed by ...
    const ext_t* inv1mw;   // ... 1 / (1 - w_j) for every variable (host-computed; the claim-derived rounds need it)
    ext_t* prefix;         // device scalar: P_{j-1} on entry when `fold`, P_j otherwise (WITHOUT the rank factor)
    ext_t* qstate;         // device: q_{j-1}(X) = qstate[0] + qstate[1] X + qstate[2] X^2 of the previous round (written every round)
    ext_t scale;           // sharded: eq(w_top, rank), the constant factor of eq on this rank's slice (else 1)
    uint32_t round;        // j
    int fold;              // this launch folds by r_{j-1}: P_j = P_{j-1} eq(w_{j-1}, r_{j-1}) is stored back
    int derive;            // 1: res = (S11, c2, -) and q(0) follows from the running claim; 0: res = (S00, S11, Sx)
    int sharded;
    ext_t r;
    const ext_t* r_ptr;
    // linear part, before the ranks' sums are added: the rank's constant eq factor
    GL_DEV void pre(ext_t (&res)[3]) const {
        if (!sharded) return;
        const extmul_t sm = extmul_prep(scale);
#pragma unroll
        for (int x = 0; x < 3; x++) res[x] = ext_mul_prep(res[x], sm);
    }
    // Claim-derived rounds (j >= 1): the verifier's relation p_j(0) + p_j(1) = claim_j, divided by P_j, reads
    //   (1 - w_j) q_j(0) + w_j q_j(1) = q_{j-1}(r_{j-1}),
    // so the kernel only accumulates q_j(1) = S11 and the X^2 coefficient c2 = sum E (A_lo - A_hi)(B_lo - B_hi) — two
    // products per pair instead of four — and q_j(0) is solved from the claim.  Exact field arithmetic: the same bits.
    GL_DEV void post(ext_t (&res)[3]) const {
        ext_t P = ext_canon(*prefix);
        const ext_t one = ext_one();
        ext_t rr = ext_zero();
        if (fold) {
            rr = ext_canon(r_ptr ? ld_ext(r_ptr) : r);
            const ext_t wp = ext_canon(w[round - 1]);
            P = ext_mul(P, ext_add(ext_mul(ext_sub(one, wp), ext_sub(one, rr)), ext_mul(wp, rr)));
            *prefix = P;
        }
        const ext_t wj = ext_canon(w[round]);
        ext_t q0, q1, c2;
        if (derive) {
            q1 = res[0];
            c2 = res[1];
            const ext_t claim = ext_add(qstate[0], ext_mul(rr, ext_add(qstate[1], ext_mul(rr, qstate[2]))));
            q0 = ext_mul(ext_sub(claim, ext_mul(wj, q1)), ext_canon(inv1mw[round]));
        } else {
            q0 = res[0];
            q1 = res[1];
            c2 = ext_sub(ext_add(q0, q1), res[2]);
        }
        const ext_t c1 = ext_sub(ext_sub(q1, q0), c2);
        qstate[0] = q0; qstate[1] = c1; qstate[2] = c2;
        const ext_t q2 = ext_add(ext_add(q0, ext_mul_base(c1, 2)), ext_mul_base(c2, 4));
        const ext_t q3 = ext_add(ext_add(q0, ext_mul_base(c1, 3)), ext_mul_base(c2, 9));
        // eq(w_j, t) = 1 - w_j + t (2 w_j - 1):  t=1: w_j,  t=2: 3 w_j - 1,  t=3: 5 w_j - 2
        const ext_t e2 = ext_sub(ext_mul_base(wj, 3), one);
        const ext_t e3 = ext_sub(ext_mul_base(wj, 5), ext_make(2, 0));
        res[0] = ext_mul(ext_mul(P, wj), q1);
        res[1] = ext_mul(ext_mul(P, e2), q2);
        res[2] = ext_mul(ext_mul(P, e3), q3);
    }
};
struct VeqArgs {
    const ext_t* in[2];       // A, B: the state this launch reads
    ext_t* out[2];            // FOLD: where the folded state goes
    const ulonglong4* L;      // 256 entries {c0, c1, 7 c1, 0}: eq over variables [j+1, j+9)
    const ulonglong4* H;      // n_rows entries, same format: eq over variables [j+9, k)
    uint64_t n_rows;          // pairs / 256
    ext_t r;                  // fold challenge (FOLD) ...
    const ext_t* r_ptr;       // ... or its device location
    VeqFin fin;
    RoundOut out_;
};
GL_DEV ulonglong4 ld_tab(const ulonglong4* p) {   // read-only table entry, one 256-bit load
    ulonglong4 v;
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.x), "=l"(v.y), "=l"(v.z), "=l"(v.w) : "l"(p));
    return v;
}
// one pair: weight A by W, accumulate the three bilinear sums (operands may be any u64 — no canonical form needed)
GL_DEV void veq_item(ext_t alo, ext_t ahi, ext_t blo, ext_t bhi, const extmul_t& W, eacc& S00, eacc& S11, eacc& Sx) {
    const ext_t wl = ext_mul_prep_weak(alo, W), wh = ext_mul_prep_weak(ahi, W);
    const uint64_t b7l = gl_mul7_weak(blo.c1), b7h = gl_mul7_weak(bhi.c1);
    eacc_mac(S00, wl, blo, b7l);
    eacc_mac(S11, wh, bhi, b7h);
    eacc_mac(Sx, wl, bhi, b7h);
    eacc_mac(Sx, wh, blo, b7l);
}
// claim-derived rounds: q(1) and the X^2 coefficient only (canonical operands: the fold's outputs)
GL_DEV void veq_item2(ext_t alo, ext_t ahi, ext_t blo, ext_t bhi, const extmul_t& W, eacc& S11, eacc& C2) {
    const ext_t da = ext_sub(alo, ahi), db = ext_sub(blo, bhi);
    const ext_t wh = ext_mul_prep_weak(ahi, W), wd = ext_mul_prep_weak(da, W);
    eacc_mac(S11, wh, bhi, gl_mul7_weak(bhi.c1));
    eacc_mac(C2, wd, db, gl_mul7_weak(db.c1));
}
template <bool FOLD, bool CANON, int MINB>
__global__ void __launch_bounds__(256, MINB) veq_round_kernel(const __grid_constant__ VeqArgs a) {
    extmul_t rm = {0, 0, 0};
    if (FOLD) rm = extmul_prep(a.r_ptr ? ld_ext(a.r_ptr) : a.r);
    // balanced contiguous row range of this block; a row is 256 consecutive pairs (one per thread)
    const uint64_t r0 = a.n_rows * blockIdx.x / gridDim.x, r1 = a.n_rows * (blockIdx.x + 1) / gridDim.x;
    eacc S00, S11, Sx;
    eacc_zero(S00); eacc_zero(S11); eacc_zero(Sx);
    for (uint64_t row = r0; row < r1; row++) {
        const ulonglong4 hv = ld_tab(a.H + row);   // uniform across the block
        extmul_t W; W.c0 = hv.x; W.c1 = hv.y; W.c1_7 = hv.z;
        const uint64_t item = row * 256 + threadIdx.x;
        ext_t alo, ahi, blo, bhi;
        load_pair<FOLD, CANON && FOLD>(a.in[0], a.out[0], item, rm, alo, ahi);
        load_pair<FOLD, CANON && FOLD>(a.in[1], a.out[1], item, rm, blo, bhi);
        veq_item(alo, ahi, blo, bhi, W, S00, S11, Sx);
    }
    const ulonglong4 lv = ld_tab(a.L + threadIdx.x);   // this thread's fixed low-variable weight, applied once
    extmul_t Lm; Lm.c0 = lv.x; Lm.c1 = lv.y; Lm.c1_7 = lv.z;
    ext_t acc[3] = {ext_mul_prep(eacc_weak(S00), Lm), ext_mul_prep(eacc_weak(S11), Lm), ext_mul_prep(eacc_weak(Sx), Lm)};
    block_finish<3, VeqFin>(acc, a.out_, a.fin);
}
// ---- split-eq rounds for the GENERAL tower-layer polynomial (any number of product / logup specs with their alphas, virtual
// leaves): the eq factor of CpuTowerProver::create_proof's layer sumcheck (ceno_zkvm/src/scheme/cpu/mod.rs:417-485) handed
// over as its point.  With E(x) = F[fixed part of x] * U[uniform part of x]:
//   q(X) = sum_x E(x) g(X, x),  g = sum_p alpha_p a_p b_p + sum_l [ an_l (p1 q2 + p2 q1) + ad_l q1 q2 ]   (degree 2 in X)
// the alphas are folded into the uniform weights on the host side of the launch (table U holds S = n_prod + 2 n_logup
// prepared entries alpha_s * U per index), one operand of every product is weighted, the thread-fixed part F is applied once
// per thread, and (claim-derived rounds) only q(1) and the X^2 coefficient are accumulated:
//   product spec:  q(1) += (W a_hi) b_hi,   c2 += (W da) db                                 (d. = lo - hi)
//   logup spec:    q(1) += (Wn p1 + Wd q1)_hi q2_hi + (Wn p2)_hi q1_hi,   c2 likewise on the differences
// i.e. 4 / 10 extension multiplications per pair instead of 12 / 21 with a streamed eq table evaluated at three points.
// Two index mappings: rows of 256 pairs with the lanes over the LOW item bits (F = low table, U = one row weight per block
// step), or — launches that read virtual leaves, whose neighbouring items are different record arrays — lanes over the HIGH
// item bits (the record rows, contiguous in memory) and a loop over the low bits (U = low table, F = high table).
#define CG_TVEQ_PAD_CHUNKS 8
struct TVeqArgs {
    TowerArgs t;               // eq_in / eq_out / out unused
    const ulonglong4* U;       // uniform part: S prepared entries {c0, c1, 7 c1, -} per index, alpha folded in
    const ulonglong4* F;       // thread-fixed part
    uint32_t S;
    uint32_t lo_bits;          // lanes-over-high mapping: item = (hi << lo_bits) | lo, U indexed by lo, F by hi
    // ... whose items lo >= pad_lo lie in the record padding of EVERY slot (all values default, lo == hi): their constant
    // contribution to q(0) = q(1), summed per chunk of 256 items with the U weights, comes from the host (pad_sum)
    uint32_t pad_lo;
    ext_t pad_sum[CG_TVEQ_PAD_CHUNKS];
    VeqFin fin;
    RoundOut out_;
};
GL_DEV extmul_t tab_mul(const ulonglong4* p) {
    const ulonglong4 v = ld_tab(p);
    extmul_t m; m.c0 = v.x; m.c1 = v.y; m.c1_7 = v.z;
    return m;
}
// one pair of every spec; loaders return canonical values
template <bool DERIVE, class L>
GL_DEV void tveq_item(const TowerArgs& a, L& ld, uint64_t item, const ulonglong4* __restrict__ U, eacc& S0, eacc& S1, eacc& C2) {
    for (int p = 0; p < a.n_prod; p++) {
        const extmul_t W = tab_mul(U + p);
        ext_t alo, ahi, blo, bhi;
        ld.prod(p, 0, item, alo, ahi);
        ld.prod(p, 1, item, blo, bhi);
        const ext_t db = ext_sub(blo, bhi);
        if (DERIVE) {
            const ext_t wh = ext_mul_prep_weak(ahi, W), wd = ext_mul_prep_weak(ext_sub(alo, ahi), W);
            eacc_mac(S1, wh, bhi, gl_mul7_weak(bhi.c1));
            eacc_mac(C2, wd, db, gl_mul7_weak(db.c1));
        } else {
            const ext_t wl = ext_mul_prep(alo, W), wh = ext_mul_prep(ahi, W);   // canonical: their difference is taken
            eacc_mac(S0, wl, blo, gl_mul7_weak(blo.c1));
            eacc_mac(S1, wh, bhi, gl_mul7_weak(bhi.c1));
            eacc_mac(C2, ext_sub(wl, wh), db, gl_mul7_weak(db.c1));
        }
    }
    for (int l = 0; l < a.n_logup; l++) {
        const extmul_t Wn = tab_mul(U + a.n_prod + 2 * l), Wd = tab_mul(U + a.n_prod + 2 * l + 1);
        ext_t p1lo, p1hi, p2lo, p2hi, q1lo, q1hi, q2lo, q2hi;
        if ((a.pone_mask >> l) & 1) {   // p1 = p2 = 1:  g = an (q1 + q2) + ad q1 q2 = (an + ad q1) q2 + an q1, differences of p vanish
            ld.store_const(a.lk_out[l][0], item, ext_one());
            ld.store_const(a.lk_out[l][1], item, ext_one());
            ld.lk(l, 2, item, q1lo, q1hi);
            ld.lk(l, 3, item, q2lo, q2hi);
            const ext_t wn = ext_make(Wn.c0, Wn.c1), dq1 = ext_sub(q1lo, q1hi), dq2 = ext_sub(q2lo, q2hi);
            if (!DERIVE) {
                eacc_mac(S0, ext_add(ext_mul_prep(q1lo, Wd), wn), q2lo, gl_mul7_weak(q2lo.c1));
                eacc_mac(S0, wn, q1lo, gl_mul7_weak(q1lo.c1));
            }
            eacc_mac(S1, ext_add(ext_mul_prep(q1hi, Wd), wn), q2hi, gl_mul7_weak(q2hi.c1));
            eacc_mac(S1, wn, q1hi, gl_mul7_weak(q1hi.c1));
            eacc_mac(C2, ext_mul_prep_weak(dq1, Wd), dq2, gl_mul7_weak(dq2.c1));
            continue;
        }
        ld.lk(l, 0, item, p1lo, p1hi);
        ld.lk(l, 1, item, p2lo, p2hi);
        ld.lk(l, 2, item, q1lo, q1hi);
        ld.lk(l, 3, item, q2lo, q2hi);
        const ext_t dq1 = ext_sub(q1lo, q1hi), dq2 = ext_sub(q2lo, q2hi);
        if (DERIVE) {
            eacc E;
            eacc_mul_prep(E, p1hi, Wn);
            eacc_mac_prep(E, q1hi, Wd);
            eacc_mac(S1, eacc_weak(E), q2hi, gl_mul7_weak(q2hi.c1));
            eacc_mac(S1, ext_mul_prep_weak(p2hi, Wn), q1hi, gl_mul7_weak(q1hi.c1));
            eacc_mul_prep(E, ext_sub(p1lo, p1hi), Wn);
            eacc_mac_prep(E, dq1, Wd);
            eacc_mac(C2, eacc_weak(E), dq2, gl_mul7_weak(dq2.c1));
            eacc_mac(C2, ext_mul_prep_weak(ext_sub(p2lo, p2hi), Wn), dq1, gl_mul7_weak(dq1.c1));
        } else {
            eacc E;
            eacc_mul_prep(E, p1lo, Wn);
            eacc_mac_prep(E, q1lo, Wd);
            const ext_t ul = eacc_canon(E), vl = ext_mul_prep(p2lo, Wn);
            eacc_mul_prep(E, p1hi, Wn);
            eacc_mac_prep(E, q1hi, Wd);
            const ext_t uh = eacc_canon(E), vh = ext_mul_prep(p2hi, Wn);
            eacc_mac(S0, ul, q2lo, gl_mul7_weak(q2lo.c1));
            eacc_mac(S0, vl, q1lo, gl_mul7_weak(q1lo.c1));
            eacc_mac(S1, uh, q2hi, gl_mul7_weak(q2hi.c1));
            eacc_mac(S1, vh, q1hi, gl_mul7_weak(q1hi.c1));
            eacc_mac(C2, ext_sub(ul, uh), dq2, gl_mul7_weak(dq2.c1));
            eacc_mac(C2, ext_sub(vl, vh), dq1, gl_mul7_weak(dq1.c1));
        }
    }
}
// VIRT selects the lanes-over-high mapping as well as the leaf loaders
template <bool FOLD, bool CANON, bool DERIVE, bool VIRT>
__global__ void __launch_bounds__(256, 2) tveq_round_kernel(const __grid_constant__ TVeqArgs a) {
    GlobalLoader<FOLD, CANON, VIRT> ld{a.t, {0, 0, 0}};
    if (FOLD) ld.rm = extmul_prep(a.t.r_ptr ? ld_ext(a.t.r_ptr) : a.t.r);
    const uint32_t S = a.S;
    ext_t s0 = ext_zero(), s1 = ext_zero(), c2 = ext_zero();
    eacc S0, S1, C2;
    if (!VIRT) {
        const uint64_t n_rows = a.t.n_pairs >> CG_VEQ_LO_BITS;
        const uint64_t r0 = n_rows * blockIdx.x / gridDim.x, r1 = n_rows * (blockIdx.x + 1) / gridDim.x;
        eacc_zero(S0); eacc_zero(S1); eacc_zero(C2);
        for (uint64_t row = r0; row < r1; row++)
            tveq_item<DERIVE>(a.t, ld, row * 256 + threadIdx.x, a.U + row * S, S0, S1, C2);
        const extmul_t Fm = tab_mul(a.F + threadIdx.x);
        if (!DERIVE) s0 = ext_mul_prep(eacc_weak(S0), Fm);
        s1 = ext_mul_prep(eacc_weak(S1), Fm);
        c2 = ext_mul_prep(eacc_weak(C2), Fm);
    } else {
        const uint64_t n_lo = 1ULL << a.lo_bits, n_hi = a.t.n_pairs >> a.lo_bits;
        const uint64_t chunk = n_lo < 256 ? n_lo : 256;
        // A unit = (chunk c, row hi) is shared by FOUR neighbouring lanes (items lo = q, q + 4, ...): a warp covers 8 consecutive
        // rows, so every record read is a full 128-byte line and the four lanes' folded outputs (32 B each) form one line too.
        const uint64_t units = n_hi * (n_lo / chunk), stride = ((uint64_t)gridDim.x * blockDim.x) >> 2;
        const uint32_t q = threadIdx.x & 3;
        for (uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2; g < units; g += stride) {
            const uint64_t c = g / n_hi, hi = g - c * n_hi;
            eacc_zero(S0); eacc_zero(S1); eacc_zero(C2);
            const uint64_t lo0 = c * chunk, lo1 = lo0 + chunk;
            const uint64_t live = a.pad_lo < lo0 ? lo0 : (a.pad_lo < lo1 ? a.pad_lo : lo1);
            for (uint64_t lo = lo0 + q; lo < lo1; lo += 4) {
                const uint64_t item = (hi << a.lo_bits) + lo;
                if (lo < live) tveq_item<DERIVE>(a.t, ld, item, a.U + lo * S, S0, S1, C2);
                else if (FOLD) {   // record padding: every slot holds its default on both sides of the pair
                    for (int p = 0; p < a.t.n_prod; p++)
                        for (int z = 0; z < 2; z++) ld.store_const(a.t.prod_out[p][z], item, a.t.virt[1 + 2 * p + z]->def);
                    for (int l = 0; l < a.t.n_logup; l++)
                        for (int z = 0; z < 4; z++) ld.store_const(a.t.lk_out[l][z], item, a.t.virt[1 + 2 * a.t.n_prod + 4 * l + z]->def);
                }
            }
            const extmul_t Fm = tab_mul(a.F + hi);
            ext_t t0 = ext_zero(), t1 = eacc_canon(S1);
            if (!DERIVE) t0 = eacc_canon(S0);
            if (live < lo1 && q == 0) {   // the padding's constant contribution, once per unit
                t1 = ext_add(t1, a.pad_sum[c]);
                if (!DERIVE) t0 = ext_add(t0, a.pad_sum[c]);
            }
            if (!DERIVE) s0 = ext_add(s0, ext_mul_prep(t0, Fm));
            s1 = ext_add(s1, ext_mul_prep(t1, Fm));
            c2 = ext_add(c2, ext_mul_prep(eacc_weak(C2), Fm));
        }
    }
    // VeqFin::post takes (q(1), c2, -) when deriving, (q(0), q(1), q(0) + q(1) - c2) otherwise
    ext_t acc[3];
    if (DERIVE) { acc[0] = s1; acc[1] = c2; acc[2] = ext_zero(); }
    else { acc[0] = s0; acc[1] = s1; acc[2] = ext_sub(ext_add(s0, s1), c2); }
    block_finish<3, VeqFin>(acc, a.out_, a.fin);
}
// ---- the same round with TMA staging: rows (256 pairs = 16 KB per MLE when folding, 8 KB in round 0) are
// brought into a shared-memory ring by 1-D bulk copies (cp.async.bulk, mbarrier complete_tx), so DRAM latency is
// hidden by the ring depth instead of by resident warps (the register-heavy lazy accumulators allow only 16
// warps per SM; ncu showed long-scoreboard stalls as the top issue-slot loss of the LDG version).
// Bank conflicts: a thread owns 64 contiguous bytes (4 ext), so a plain read order would be 4-way conflicted.
// Thread t reads its 16-byte chunks in the order i ^ s, s = (t >> 1) & 3 — conflict-free — and the permutation
// costs nothing per pair: s & 1 swaps (x0, x1) and (x2, x3), absorbed by folding with 1 - r
// (x1 + (1-r)(x0 - x1) = x0 + r (x1 - x0)); s & 2 swaps the lo/hi pair, absorbed by swapping S00 and S11
// once at the end (Sx is symmetric) and by the store address.
GL_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
GL_DEV void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
GL_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
GL_DEV void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
GL_DEV bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
GL_DEV void mbar_wait(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
GL_DEV void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
GL_DEV ext_t lds_ext(uint32_t addr) {
    ext_t v;
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v.c0), "=l"(v.c1) : "r"(addr));
    return v;
}
// MINB = resident blocks per SM (2: 128 registers per thread; 3: 80 registers, a few spills, 24 warps per SM);
// the ring depth follows from the shared memory left per block.
template <bool FOLD, int MINB = 2>
struct VeqTmaCfg {
    static constexpr uint32_t ROWB = FOLD ? 16384u : 8192u;   // bytes per MLE per row
    static constexpr uint32_t STAGEB = 2 * ROWB + 128;        // A row | B row | H entry (32 B) + pad
    static constexpr int STAGES = MINB == 2 ? (FOLD ? 3 : 4) : (FOLD ? 2 : 4);
    static constexpr uint32_t SMEM = STAGES * STAGEB + 2 * STAGES * 8;
};
// DERIVE (FOLD only): claim-derived round — accumulate q(1) and the X^2 coefficient only (VeqFin::post solves q(0) from the
// running claim).  Its chunk order swaps within a pair only (absorbed by folding with 1 - r), so lo / hi are never exchanged
// (a 2-way instead of a conflict-free shared-memory read: ~1 % of the row time).
template <bool FOLD, bool CANON, int MINB = 2, bool DERIVE = false>
__global__ void __launch_bounds__(256, MINB) veq_tma_kernel(const __grid_constant__ VeqArgs a) {
    static_assert(FOLD || !DERIVE, "round 0 has no claim to derive from");
    using Cfg = VeqTmaCfg<FOLD, MINB>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr uint32_t ROWB = Cfg::ROWB, STAGEB = Cfg::STAGEB;
    extern __shared__ __align__(128) unsigned char veq_smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t sbase = smem_u32(veq_smem);
    const uint32_t bar_full = sbase + STAGES * STAGEB, bar_empty = bar_full + STAGES * 8;
    const uint64_t r0 = a.n_rows * blockIdx.x / gridDim.x, r1 = a.n_rows * (blockIdx.x + 1) / gridDim.x;
    const uint32_t nrows = (uint32_t)(r1 - r0);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint32_t i) {   // thread 0: bring row r0 + i into stage i % STAGES
        const uint32_t s = i % STAGES;
        const uint64_t row = r0 + i;
        const uint32_t dst = sbase + s * STAGEB, bar = bar_full + 8 * s;
        mbar_expect_tx(bar, 2 * ROWB + 32);
        tma_load_1d(dst, reinterpret_cast<const unsigned char*>(a.in[0]) + row * ROWB, ROWB, bar);
        tma_load_1d(dst + ROWB, reinterpret_cast<const unsigned char*>(a.in[1]) + row * ROWB, ROWB, bar);
        tma_load_1d(dst + 2 * ROWB, a.H + row, 32, bar);
    };
    if (tid == 0)
        for (uint32_t i = 0; i < (uint32_t)STAGES && i < nrows; i++) issue(i);
    // per-thread chunk permutation (see above)
    const uint32_t sx = FOLD ? (DERIVE ? ((tid >> 1) & 1) : ((tid >> 1) & 3)) : ((tid >> 2) & 1);
    const uint32_t hs = FOLD ? (sx >> 1) : sx;            // lo/hi pair swapped for this thread
    const uint32_t toff = FOLD ? tid * 64 : tid * 32;
    extmul_t rm = {0, 0, 0};
    if (FOLD) {
        ext_t r = ext_canon(a.r_ptr ? ld_ext(a.r_ptr) : a.r);
        if (sx & 1) r = ext_sub(ext_one(), r);
        rm = extmul_prep(r);
    }
    eacc Sf, Ss, Sx;   // first*first, second*second, cross
    eacc_zero(Sf); eacc_zero(Ss); eacc_zero(Sx);
    uint32_t s = 0, ph = 0;
#pragma unroll 1   // measured: unrolling by 2 costs registers (spills) and 8 % of the round time
    for (uint32_t i = 0; i < nrows; i++) {
        if (tid == 0 && i >= 1 && i - 1 + STAGES < nrows) {   // refill the stage released one row ago
            const uint32_t ps = (i - 1) % STAGES, pph = ((i - 1) / STAGES) & 1;
            mbar_wait(bar_empty + 8 * ps, pph);
            issue(i - 1 + STAGES);
        }
        mbar_wait(bar_full + 8 * s, ph);
        const uint32_t st = sbase + s * STAGEB;
        ext_t af, as_, bf, bs;
        const uint64_t item = (r0 + i) * 256 + tid;
        if (FOLD) {
            ext_t y0 = lds_ext(st + toff + ((0 ^ sx) << 4)), y1 = lds_ext(st + toff + ((1 ^ sx) << 4));
            ext_t y2 = lds_ext(st + toff + ((2 ^ sx) << 4)), y3 = lds_ext(st + toff + ((3 ^ sx) << 4));
            ext_t z0 = lds_ext(st + ROWB + toff + ((0 ^ sx) << 4)), z1 = lds_ext(st + ROWB + toff + ((1 ^ sx) << 4));
            ext_t z2 = lds_ext(st + ROWB + toff + ((2 ^ sx) << 4)), z3 = lds_ext(st + ROWB + toff + ((3 ^ sx) << 4));
            const ext_t hw0 = lds_ext(st + 2 * ROWB), hw1 = lds_ext(st + 2 * ROWB + 16);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);
            if (CANON) {
                y0 = ext_canon(y0); y1 = ext_canon(y1); y2 = ext_canon(y2); y3 = ext_canon(y3);
                z0 = ext_canon(z0); z1 = ext_canon(z1); z2 = ext_canon(z2); z3 = ext_canon(z3);
            }
            af = ext_fma_prep(y0, ext_sub(y1, y0), rm);
            as_ = ext_fma_prep(y2, ext_sub(y3, y2), rm);
            bf = ext_fma_prep(z0, ext_sub(z1, z0), rm);
            bs = ext_fma_prep(z2, ext_sub(z3, z2), rm);
            st_ext(a.out[0] + 2 * item + hs, af);
            st_ext(a.out[0] + 2 * item + (1 - hs), as_);
            st_ext(a.out[1] + 2 * item + hs, bf);
            st_ext(a.out[1] + 2 * item + (1 - hs), bs);
            extmul_t W; W.c0 = hw0.c0; W.c1 = hw0.c1; W.c1_7 = hw1.c0;
            if (DERIVE) veq_item2(af, as_, bf, bs, W, Ss, Sx);   // Ss = q(1) part, Sx = X^2-coefficient part
            else veq_item(af, as_, bf, bs, W, Sf, Ss, Sx);
        } else {
            af = lds_ext(st + toff + ((0 ^ sx) << 4));
            as_ = lds_ext(st + toff + ((1 ^ sx) << 4));
            bf = lds_ext(st + ROWB + toff + ((0 ^ sx) << 4));
            bs = lds_ext(st + ROWB + toff + ((1 ^ sx) << 4));
            const ext_t hw0 = lds_ext(st + 2 * ROWB), hw1 = lds_ext(st + 2 * ROWB + 16);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);
            extmul_t W; W.c0 = hw0.c0; W.c1 = hw0.c1; W.c1_7 = hw1.c0;
            veq_item(af, as_, bf, bs, W, Sf, Ss, Sx);
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    const ulonglong4 lv = ld_tab(a.L + tid);
    extmul_t Lm; Lm.c0 = lv.x; Lm.c1 = lv.y; Lm.c1_7 = lv.z;
    if (DERIVE) {
        ext_t acc[3] = {ext_mul_prep(eacc_weak(Ss), Lm), ext_mul_prep(eacc_weak(Sx), Lm), ext_zero()};
        block_finish<3, VeqFin>(acc, a.out_, a.fin);
    } else {
        const ext_t vf = ext_mul_prep(eacc_weak(Sf), Lm), vs = ext_mul_prep(eacc_weak(Ss), Lm);
        ext_t acc[3] = {hs ? vs : vf, hs ? vf : vs, ext_mul_prep(eacc_weak(Sx), Lm)};
        block_finish<3, VeqFin>(acc, a.out_, a.fin);
    }
}

// ---- persistent split-eq rounds: ONE cooperative launch runs the claim-derived rounds j0 .. j1-1 (every streaming round
// after the first).  A separate launch per round costs ~18 us of fixed time (launch latency, ring fill, finish chain) that
// does not shrink with the data — 15 % of a T3-24 step and most of an 8-GPU one.  Here every block keeps its ring and
// barriers, draws a ticket after its rows, the block that draws the round's last ticket combines the partials, runs the
// multi-GPU exchange, solves q(0) from the claim, obtains the challenge (device challenger or host mailbox) and releases
// the round flag the other blocks poll — the hand-off of tower_mid_kernel around the row pipeline of veq_tma_kernel.
struct VeqPersistArgs {
    const ext_t* bufA[CG_VEQ_MAX_ROUNDS + 2];   // state of A after f folds, f = j0 - 1 + i   (i = 0: the first round's input)
    const ext_t* bufB[CG_VEQ_MAX_ROUNDS + 2];
    const ulonglong4* L;                        // [J][256]
    const ulonglong4* H;                        // concatenated, H_j at h_off[j]
    uint64_t h_off[CG_VEQ_MAX_ROUNDS + 1];
    uint32_t k, j0, j1;
    int canon_first;                            // the first round reads caller-provided buffers
    ext_t r;                                    // challenge of the first round's fold ...
    const ext_t* r_ptr;                         // ... or its device location
    VeqFin fin;                                 // w, inv1mw, prefix, qstate, scale, sharded (round / fold / r are set per round)
    ext_t* d_msgs;                              // [round * 3]
    ext_t* d_chal;                              // [round]
    uint64_t* d_tr_state;                       // device challenger, or nullptr -> host mailbox
    TailMailbox* mail;
    int* d_error;
    unsigned long long timeout_cycles;
    CommDev comm;                               // exchange of round j uses sequence comm.seq + (j - j0)
    ext_t* partials;                            // [gridDim.x * 3]
    unsigned int* ticket;                       // zero on entry
    volatile unsigned int* round_flag;          // index of the last finished round + 1
};
GL_DEV void mbar_inval(uint32_t bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__global__ void __launch_bounds__(256, 2) veq_persist_kernel(const __grid_constant__ VeqPersistArgs a) {
    using Cfg = VeqTmaCfg<true, 2>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr uint32_t ROWB = Cfg::ROWB, STAGEB = Cfg::STAGEB;
    extern __shared__ __align__(128) unsigned char veq_smem[];
    __shared__ ext_t s_part[8][3];
    __shared__ __align__(32) uint64_t s_msg[2 * 3 + 8];
    __shared__ ext_t s_r;
    __shared__ int s_flag;   // 1: this block drew the round's last ticket, 2: abort
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sbase = smem_u32(veq_smem);
    const uint32_t bar_full = sbase + STAGES * STAGEB, bar_empty = bar_full + STAGES * 8;
    const uint32_t sx = (tid >> 1) & 1;          // claim-derived chunk order: swap within a pair only
    const uint32_t toff = tid * 64;
    ext_t r_cur = ext_canon(a.r_ptr ? ld_ext(a.r_ptr) : a.r);
    for (uint32_t j = a.j0; j < a.j1; j++) {
        const uint32_t step = j - a.j0;
        const ext_t* inA = a.bufA[step];
        const ext_t* inB = a.bufB[step];
        ext_t* outA = const_cast<ext_t*>(a.bufA[step + 1]);
        ext_t* outB = const_cast<ext_t*>(a.bufB[step + 1]);
        const ulonglong4* Hj = a.H + a.h_off[j];
        const uint64_t n_rows = 1ULL << (a.k - j - 1 - CG_VEQ_LO_BITS);
        const uint64_t r0 = n_rows * blockIdx.x / gridDim.x, r1 = n_rows * (blockIdx.x + 1) / gridDim.x;
        const uint32_t nrows = (uint32_t)(r1 - r0);
        const bool canon = a.canon_first && step == 0;
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; s++) {
                if (step) { mbar_inval(bar_full + 8 * s); mbar_inval(bar_empty + 8 * s); }
                mbar_init(bar_full + 8 * s, 1);
                mbar_init(bar_empty + 8 * s, 8);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            s_flag = 0;
        }
        __syncthreads();
        auto issue = [&](uint32_t i) {
            const uint32_t s = i % STAGES;
            const uint64_t row = r0 + i;
            const uint32_t dst = sbase + s * STAGEB, bar = bar_full + 8 * s;
            mbar_expect_tx(bar, 2 * ROWB + 32);
            tma_load_1d(dst, reinterpret_cast<const unsigned char*>(inA) + row * ROWB, ROWB, bar);
            tma_load_1d(dst + ROWB, reinterpret_cast<const unsigned char*>(inB) + row * ROWB, ROWB, bar);
            tma_load_1d(dst + 2 * ROWB, Hj + row, 32, bar);
        };
        if (tid == 0)
            for (uint32_t i = 0; i < (uint32_t)STAGES && i < nrows; i++) issue(i);
        ext_t rr = r_cur;
        if (sx & 1) rr = ext_sub(ext_one(), rr);
        const extmul_t rm = extmul_prep(rr);
        eacc Ss, Sx;
        eacc_zero(Ss); eacc_zero(Sx);
        uint32_t s = 0, ph = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < nrows; i++) {
            if (tid == 0 && i >= 1 && i - 1 + STAGES < nrows) {
                const uint32_t ps = (i - 1) % STAGES, pph = ((i - 1) / STAGES) & 1;
                mbar_wait(bar_empty + 8 * ps, pph);
                issue(i - 1 + STAGES);
            }
            mbar_wait(bar_full + 8 * s, ph);
            const uint32_t st = sbase + s * STAGEB;
            const uint64_t item = (r0 + i) * 256 + tid;
            ext_t y0 = lds_ext(st + toff + ((0 ^ sx) << 4)), y1 = lds_ext(st + toff + ((1 ^ sx) << 4));
            ext_t y2 = lds_ext(st + toff + ((2 ^ sx) << 4)), y3 = lds_ext(st + toff + ((3 ^ sx) << 4));
            ext_t z0 = lds_ext(st + ROWB + toff + ((0 ^ sx) << 4)), z1 = lds_ext(st + ROWB + toff + ((1 ^ sx) << 4));
            ext_t z2 = lds_ext(st + ROWB + toff + ((2 ^ sx) << 4)), z3 = lds_ext(st + ROWB + toff + ((3 ^ sx) << 4));
            const ext_t hw0 = lds_ext(st + 2 * ROWB), hw1 = lds_ext(st + 2 * ROWB + 16);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);
            if (canon) {
                y0 = ext_canon(y0); y1 = ext_canon(y1); y2 = ext_canon(y2); y3 = ext_canon(y3);
                z0 = ext_canon(z0); z1 = ext_canon(z1); z2 = ext_canon(z2); z3 = ext_canon(z3);
            }
            const ext_t af = ext_fma_prep(y0, ext_sub(y1, y0), rm), as_ = ext_fma_prep(y2, ext_sub(y3, y2), rm);
            const ext_t bf = ext_fma_prep(z0, ext_sub(z1, z0), rm), bs = ext_fma_prep(z2, ext_sub(z3, z2), rm);
            st_ext(outA + 2 * item, af);
            st_ext(outA + 2 * item + 1, as_);
            st_ext(outB + 2 * item, bf);
            st_ext(outB + 2 * item + 1, bs);
            extmul_t W; W.c0 = hw0.c0; W.c1 = hw0.c1; W.c1_7 = hw1.c0;
            veq_item2(af, as_, bf, bs, W, Ss, Sx);
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");   // this round's stores are the next round's bulk-copy sources
        const ulonglong4 lv = ld_tab(a.L + ((size_t)j << CG_VEQ_LO_BITS) + tid);
        extmul_t Lm; Lm.c0 = lv.x; Lm.c1 = lv.y; Lm.c1_7 = lv.z;
        ext_t acc[3] = {ext_mul_prep(eacc_weak(Ss), Lm), ext_mul_prep(eacc_weak(Sx), Lm), ext_zero()};
#pragma unroll
        for (int x = 0; x < 2; x++) {
            const ext_t v = warp_reduce_ext(acc[x]);
            if (lane == 0) s_part[warp][x] = v;
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int x = 0; x < 2; x++) {
                const ext_t v = warp_reduce_ext(lane < 8 ? s_part[lane][x] : ext_zero());
                if (lane == 0) a.partials[(size_t)blockIdx.x * 3 + x] = v;
            }
            if (lane == 0) {
                __threadfence();   // partial + this block's folded outputs are visible before the ticket
                const unsigned tk = atomicAdd(a.ticket, 1u);
                if (tk == gridDim.x * (step + 1) - 1) s_flag = 1;
            }
        }
        __syncthreads();
        if (s_flag == 1) {   // last block of the round: combine, exchange, solve q(0), challenge, release
            __threadfence();
            ext_t res[3];
#pragma unroll
            for (int x = 0; x < 2; x++) {
                ext_t v = ext_zero();
                for (unsigned b = tid; b < gridDim.x; b += blockDim.x) {
                    const ulonglong2 p = __ldcg(reinterpret_cast<const ulonglong2*>(&a.partials[(size_t)b * 3 + x]));
                    v = ext_add(v, ext_make(p.x, p.y));
                }
                v = warp_reduce_ext(v);
                if (lane == 0) s_part[warp][x] = v;
            }
            __syncthreads();
            if (warp == 0) {
#pragma unroll
                for (int x = 0; x < 2; x++) res[x] = warp_reduce_ext(lane < 8 ? s_part[lane][x] : ext_zero());
                res[2] = ext_zero();
                VeqFin fin = a.fin;
                fin.round = j;
                fin.fold = 1;
                fin.derive = 1;
                fin.r = r_cur;
                fin.r_ptr = nullptr;
                if (lane == 0) fin.pre(res);
                if (a.comm.nranks > 1) comm_exchange<3>(res, a.comm, a.comm.seq + step, s_msg);
                if (lane == 0) {
                    fin.post(res);
                    ext_t rn = ext_zero();
#pragma unroll
                    for (int x = 0; x < 3; x++) a.d_msgs[(size_t)j * 3 + x] = res[x];
                    bool abort = false;
                    if (a.d_tr_state) {
                        uint64_t h = *a.d_tr_state;
                        rn = cg_tr_round<3>(h, res);
                        *a.d_tr_state = h;
                    } else {
                        TailMailbox* mb = a.mail;
                        mailbox_post<3>(mb, res, (uint64_t)j + 1);
                        const long long t0 = clock64();
                        while (true) {
                            const int st = mailbox_poll(mb, (uint64_t)j + 1, rn);
                            if (st == 1) break;
                            if (st < 0 || (unsigned long long)(clock64() - t0) > a.timeout_cycles) { abort = true; *a.d_error = 1; break; }
                        }
                    }
                    a.d_chal[j] = rn;
                    __threadfence();
                    *a.round_flag = abort ? 0xFFFFFFFFu : (j + 1);   // release
                }
            }
        }
        if (tid == 0) {   // every block: wait for the round to be released, pick up the challenge
            const long long t0 = clock64();
            unsigned fl;
            while ((fl = *a.round_flag) < j + 1) {
                if ((unsigned long long)(clock64() - t0) > 2 * a.timeout_cycles) { fl = 0xFFFFFFFFu; break; }
            }
            if (fl == 0xFFFFFFFFu) s_flag = 2;
            else {
                const ulonglong2 p = __ldcg(reinterpret_cast<const ulonglong2*>(&a.d_chal[j]));
                s_r = ext_make(p.x, p.y);
            }
        }
        __syncthreads();
        if (s_flag == 2) return;
        __threadfence();   // acquire: the other blocks' folded outputs (the gpu-scope fence also invalidates L1)
        r_cur = s_r;
    }
}

// all tables of the split rounds in one launch: entry = direct product over its variables
#define CG_TVEQ_MAX_SLOTS (CG_TOWER_MAX_PROD + 2 * CG_TOWER_MAX_LOGUP)
struct VeqTabArgs {
    const ext_t* w;
    uint32_t k, J;
    ulonglong4* L;                              // [J][256]
    ulonglong4* H;                              // concatenated, H_j at h_off[j]
    uint64_t h_off[CG_VEQ_MAX_ROUNDS + 1];
    ulonglong4* UA;                             // general tower layouts: S alpha-folded copies of every H entry, [entry][slot]
    uint32_t S;
    ext_t alpha[CG_TVEQ_MAX_SLOTS];
};
__global__ void __launch_bounds__(CG_THREADS) veq_tables_kernel(const __grid_constant__ VeqTabArgs a) {
    const uint64_t n_lo = (uint64_t)a.J << CG_VEQ_LO_BITS, total = n_lo + a.h_off[a.J];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        uint32_t v0, nb;
        uint64_t x;
        ulonglong4* dst;
        if (e < n_lo) {
            const uint32_t j = (uint32_t)(e >> CG_VEQ_LO_BITS);
            x = e & ((1u << CG_VEQ_LO_BITS) - 1);
            v0 = j + 1; nb = CG_VEQ_LO_BITS;
            dst = a.L + e;
        } else {
            const uint64_t y = e - n_lo;
            uint32_t j = 0;
            while (y >= a.h_off[j + 1]) j++;
            x = y - a.h_off[j];
            v0 = j + 1 + CG_VEQ_LO_BITS; nb = a.k - v0;
            dst = a.H + y;
        }
        ext_t v = ext_one();
        for (uint32_t i = 0; i < nb; i++) {
            const ext_t wi = ext_canon(a.w[v0 + i]);
            v = ext_mul(v, ((x >> i) & 1) ? wi : ext_sub(ext_one(), wi));
        }
        *dst = make_ulonglong4(v.c0, v.c1, gl_canon(gl_mul7_weak(v.c1)), 0ULL);
        if (e >= n_lo)
            for (uint32_t sl = 0; sl < a.S; sl++) {
                const ext_t av = ext_mul(ext_canon(a.alpha[sl]), v);
                a.UA[(e - n_lo) * a.S + sl] = make_ulonglong4(av.c0, av.c1, gl_canon(gl_mul7_weak(av.c1)), 0ULL);
            }
    }
}
// one eq table over the variables [v0, v0 + nb) of the point, optionally as S alpha-folded copies per entry (the tables of the
// lanes-over-high mapping)
struct VeqRangeArgs {
    const ext_t* w;
    uint32_t v0, nb, S;
    ext_t alpha[CG_TVEQ_MAX_SLOTS];
    ulonglong4* out;
};
__global__ void __launch_bounds__(CG_THREADS) veq_range_table_kernel(const __grid_constant__ VeqRangeArgs a) {
    const uint64_t total = 1ULL << a.nb, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        ext_t v = ext_one();
        for (uint32_t i = 0; i < a.nb; i++) {
            const ext_t wi = ext_canon(a.w[a.v0 + i]);
            v = ext_mul(v, ((x >> i) & 1) ? wi : ext_sub(ext_one(), wi));
        }
        if (a.S == 0) { a.out[x] = make_ulonglong4(v.c0, v.c1, gl_canon(gl_mul7_weak(v.c1)), 0ULL); continue; }
        for (uint32_t sl = 0; sl < a.S; sl++) {
            const ext_t av = ext_mul(ext_canon(a.alpha[sl]), v);
            a.out[x * a.S + sl] = make_ulonglong4(av.c0, av.c1, gl_canon(gl_mul7_weak(av.c1)), 0ULL);
        }
    }
}
// leave split mode: out[x] = P * L[x & 255] * H[x >> 8]  =  the eq state a materialised table would hold
// after the same folds (variables [f, k), prefix P = prod_{i<f} eq(w_i, r_i))
__global__ void __launch_bounds__(CG_THREADS) veq_materialise_kernel(const ulonglong4* __restrict__ L, const ulonglong4* __restrict__ H,
                                                                      const ext_t* __restrict__ prefix, ext_t scale, uint64_t n, ext_t* __restrict__ out) {
    const extmul_t P = extmul_prep(ext_mul(ext_canon(ld_ext(prefix)), scale));   // scale: the rank's constant eq factor (sharded), else 1
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t pair = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; pair < n / 2; pair += stride) {
        const uint64_t b = 2 * pair;
        const ulonglong4 hv = ld_tab(H + (b >> CG_VEQ_LO_BITS));
        const extmul_t ph = extmul_prep(ext_mul_prep(ext_make(hv.x, hv.y), P));
        const ulonglong4 l0 = ld_tab(L + (b & 255)), l1 = ld_tab(L + (b & 255) + 1);
        st_ext2(out + b, ext_mul_prep(ext_make(l0.x, l0.y), ph), ext_mul_prep(ext_make(l1.x, l1.y), ph));
    }
}

// ---------------------------------------------------------------------------------------------
// Tail kernel (tower_ctail_kernel below): once every MLE of the layer fits in shared memory, ALL remaining rounds
// (evaluate, challenge, fold ... final evaluations) run in a single persistent launch: no per-round
// launch or host synchronisation.  The challenge comes from the device-resident challenger or, for
// the reference's host-side transcript, from a host mailbox in mapped pinned memory (the kernel posts
// the round message, the host answers with the challenge; ~one PCIe round trip per round).
struct TailArgs {
    TowerArgs t;                // *_in = state to load; *_out unused; r / r_ptr = entry fold challenge
    int entry_fold;             // 1: arrays hold 2*n0 elements, fold by r while loading
    int canon;                  // caller-provided buffers: canonicalise on load
    uint32_t n0;                // elements per MLE once loaded (power of two)
    uint32_t first_round, num_rounds;   // rounds first_round .. num_rounds-1 are run here
    ext_t* d_msgs;              // [num_rounds * 3]
    ext_t* d_chal;              // [num_rounds]
    ext_t* d_final;             // one ext per MLE, in the caller's MLE order
    uint16_t final_idx[1 + 2 * CG_TOWER_MAX_PROD + 4 * CG_TOWER_MAX_LOGUP];   // slot -> MLE index
    uint64_t* d_tr_state;       // device challenger state, or nullptr -> mailbox
    TailMailbox* mail;
    int* d_error;               // set to 1 on mailbox timeout/abort
    unsigned long long timeout_cycles;
    CommDev comm;               // multi-GPU: this rank loads n_loc = n0 / nranks elements per MLE (its slice of
    int gather_par;             // the global folded array) and all-gathers the other slices over NVLink on
    uint64_t gather_seq;        // entry; every later round runs replicated, with no further exchange
};
struct SmemLoader {
    const ext_t* sm;
    uint32_t n0;
    int n_prod;
    GL_DEV void get(int slot, uint64_t item, ext_t& lo, ext_t& hi) {
        const ext_t* p = sm + (size_t)slot * n0 + 2 * item;
        lo = p[0];
        hi = p[1];
    }
    GL_DEV void eq(uint64_t item, ext_t& lo, ext_t& hi) { get(0, item, lo, hi); }
    GL_DEV void prod(int p, int z, uint64_t item, ext_t& lo, ext_t& hi) { get(1 + 2 * p + z, item, lo, hi); }
    GL_DEV void lk(int l, int z, uint64_t item, ext_t& lo, ext_t& hi) { get(1 + 2 * n_prod + 4 * l + z, item, lo, hi); }
};
#define CG_TAIL_THREADS 512
GL_DEV const ext_t* tail_slot_ptr(const TowerArgs& t, int slot) {
    if (slot == 0) return t.eq_in;
    slot -= 1;
    if (slot < 2 * t.n_prod) return t.prod_in[slot >> 1][slot & 1];
    slot -= 2 * t.n_prod;
    return t.lk_in[slot >> 2][slot & 3];
}
// ---------------------------------------------------------------------------------------------
// Cluster tail kernel.  The single-CTA tail above starts once an MLE has <= 2048 elements; everything between the
// streaming rounds and that point used to be one grid-wide hand-off per round (ticket -> last block -> flag: ~14 us
// of L2 round trips per round).  Here ONE thread-block cluster (<= 16 CTAs, one per SM) takes over as soon as the
// layer fits the cluster's combined shared memory (16 x ~200 KB: 2^16 elements per MLE for the T3 shape):
//   * CTA c keeps the contiguous slice [c n/C, (c+1) n/C) of every MLE in its own shared memory (LSB-first binding:
//     the fold is local);
//   * per round every CTA evaluates its pairs, stores its partial [p(1), p(2), p(3)] into EVERY CTA's shared memory
//     through DSMEM (st.shared::cluster), and after one barrier.cluster all CTAs add the C partials themselves, so the
//     message — and with the device challenger the challenge — is available everywhere with no second hop
//     (host transcript: CTA 0 posts the message, every CTA polls the host's reply line);
//   * once an MLE is down to `nt` elements the slices are gathered into CTA 0 (DSMEM stores + one cluster barrier),
//     the other CTAs exit and CTA 0 finishes the remaining rounds with block barriers only.
// Folds compact the arrays (slot s moves from s*n to s*n/2), so the live data always sits at the bottom of the
// region and the top half is free for the gather.  cluster size 1 = the plain single-CTA tail.
// Sharded prove: every rank computes its slice of the entry state, stores it into every peer's gather buffer over
// NVLink, and all ranks then run the whole tail replicated (no further exchange).
#define CG_CT_THREADS 512
#define CG_CT_MAX_C 16
#define CG_CT_MAX_NLOC 4096      // elements per MLE per CTA (register staging of the fold: 4 pairs per thread)
#define CG_CT_GATHER_N 512       // the slices are gathered into CTA 0 once an MLE is down to this many elements (cluster-wide)
#define CG_CT_SPLIT_PAIRS 256    // rounds with at most this many pairs per CTA split every pair over three threads (one per point)
#define CG_GATHER_EXT (1u << 18) // ext elements of one gather buffer (per parity): n_slots * n0 must fit
struct CTailArgs {
    TailArgs t;                 // as for the single-CTA tail; t.n0 = elements per MLE at entry (after the entry fold)
    uint32_t n_loc0;            // t.n0 / cluster size
    uint32_t nt;                // gather into CTA 0 once an MLE has <= nt elements (nt <= n_loc0 / 2, nt >= cluster size)
    ext_t* gbuf[CG_MAX_RANKS];  // sharded: every rank's gather buffer of this parity ([slot][n0] ext), gbuf[rank] is local
    uint64_t* gflag[CG_MAX_RANKS];   // sharded: every rank's arrival flags of this parity ([CG_MAX_RANKS])
    long long* dbg;             // CG_TAIL_DEBUG: CTA 0 records clock64() at [round][0..4] = start, evaluated, exchanged, challenged, folded
};
#define CT_STAMP(ph) do { if (ca.dbg && cr == 0 && tid == 0) ca.dbg[(size_t)(j - a.first_round) * 8 + (ph)] = clock64(); } while (0)
GL_DEV uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
GL_DEV uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
GL_DEV void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
GL_DEV uint32_t dsmem_addr(const void* local, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(cta));
    return r;
}
GL_DEV void dsmem_st_ext(uint32_t addr, ext_t v) {
    asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(addr), "l"(v.c0), "l"(v.c1) : "memory");
}
// res[x] = sum_{i < n} part[i][x] (x = 0, 1, 2; n <= 32 canonical ext partials in shared memory), valid in every lane of the
// calling warp.  Lane 2x + l adds limb l of output x as a plain 128-bit integer sum (n loads, two carry adds each) and
// reduces once; six shuffles hand the results round.  Replaces three dependent warp reductions on the round's critical path.
GL_DEV void warp_sum3_smem(const ext_t (*part)[3], int n, ext_t (&res)[3]) {
    const int lane = threadIdx.x & 31;
    uint64_t lo = 0;
    uint32_t hi = 0;
    if (lane < 6) {
        const int x = lane >> 1, l = lane & 1;
        for (int i = 0; i < n; i++) {
            const uint64_t v = l ? part[i][x].c1 : part[i][x].c0;
            lo += v;
            hi += (lo < v) ? 1u : 0u;
        }
    }
    const uint64_t r = gl_canon(gl_reduce_limbs((uint32_t)lo, (uint32_t)(lo >> 32), hi, 0, 0));
#pragma unroll
    for (int x = 0; x < 3; x++) res[x] = ext_make(__shfl_sync(0xffffffffu, r, 2 * x), __shfl_sync(0xffffffffu, r, 2 * x + 1));
}
template <bool SIMPLE>
__global__ void __launch_bounds__(CG_CT_THREADS, 1) tower_ctail_kernel(const __grid_constant__ CTailArgs ca) {
    extern __shared__ __align__(16) ext_t ct_sm[];
    __shared__ ext_t s_red[CG_CT_THREADS / 32][3];
    __shared__ __align__(16) ext_t s_parts[2][CG_CT_MAX_C][3];
    __shared__ ext_t s_r;
    __shared__ __align__(16) uint64_t s_bc[4];   // host transcript, CTAs != 0: {r.c0, r.c1, (round + 1) ^ mix(r)} written by CTA 0
    __shared__ int s_abort;
    const TailArgs& a = ca.t;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t C = cluster_nctarank(), cr = cluster_ctarank();
    const int n_slots = 1 + 2 * a.t.n_prod + 4 * a.t.n_logup;
    const uint32_t n_loc0 = ca.n_loc0;
    if (tid == 0) { s_abort = 0; s_bc[0] = s_bc[1] = s_bc[2] = 0; }
    // ---- entry: this CTA's slice [cr * n_loc0, (cr + 1) * n_loc0) of every slot (with the entry fold)
    {
        const CommDev& cm = a.comm;
        const bool sharded = cm.nranks > 1;
        const extmul_t rm = extmul_prep(a.entry_fold ? (a.t.r_ptr ? ld_ext(a.t.r_ptr) : a.t.r) : ext_zero());
        auto entry_value = [&](const ext_t* src, uint32_t b) -> ext_t {
            if (a.entry_fold) {
                ext_t lo, hi;
                ld_ext2(src + 2 * (uint64_t)b, lo, hi);
                if (a.canon) { lo = ext_canon(lo); hi = ext_canon(hi); }
                return ext_fma_prep(lo, ext_sub(hi, lo), rm);
            }
            ext_t v = ld_ext(src + b);
            if (a.canon) v = ext_canon(v);
            return v;
        };
        if (!sharded) {
            for (int slot = 0; slot < n_slots; slot++) {
                const ext_t* src = tail_slot_ptr(a.t, slot);
                for (uint32_t b = tid; b < n_loc0; b += CG_CT_THREADS) ct_sm[(size_t)slot * n_loc0 + b] = entry_value(src, cr * n_loc0 + b);
            }
        } else {
            // my rank's slice has n_sl = n0 / nranks elements per slot; this CTA computes the part [cr n_sl / C, ..) of it
            // and stores it into every rank's gather buffer (its own included)
            const uint32_t n_sl = a.n0 / (uint32_t)cm.nranks, part = n_sl / C ? n_sl / C : 1;
            const uint32_t b0 = cr * part, b1 = (b0 + part <= n_sl) ? b0 + part : (b0 < n_sl ? n_sl : b0);
            for (int slot = 0; slot < n_slots; slot++) {
                const ext_t* src = tail_slot_ptr(a.t, slot);
                for (uint32_t b = b0 + tid; b < b1; b += CG_CT_THREADS) {
                    const ext_t v = entry_value(src, b);
                    const size_t g = (size_t)slot * a.n0 + (size_t)cm.rank * n_sl + b;
                    for (int p = 0; p < cm.nranks; p++) st_ext(ca.gbuf[p] + g, v);
                }
            }
            __threadfence_system();
            cluster_sync_all();
            if (cr == 0 && tid < cm.nranks) {
                *(volatile uint64_t*)&ca.gflag[tid][cm.rank] = a.gather_seq;
                volatile uint64_t* f = &ca.gflag[cm.rank][tid];
                const long long t0 = clock64();
                while (*f != a.gather_seq) {
                    if ((unsigned long long)(clock64() - t0) > cm.timeout_cycles) { *cm.d_error = 2; break; }
                }
                __threadfence_system();
            }
            cluster_sync_all();
            const ext_t* gb = ca.gbuf[cm.rank];
            for (int slot = 0; slot < n_slots; slot++)
                for (uint32_t b = tid; b < n_loc0; b += CG_CT_THREADS) {
                    const volatile uint64_t* src = (const volatile uint64_t*)(gb + (size_t)slot * a.n0 + (size_t)cr * n_loc0 + b);
                    ct_sm[(size_t)slot * n_loc0 + b] = ext_make(src[0], src[1]);
                }
        }
    }
    __syncthreads();
    ext_t* base = ct_sm;
    bool dist = C > 1;
    uint32_t n_cur = a.n0;                  // elements per MLE, cluster-wide
    uint32_t n_my = n_loc0;                 // elements per MLE held here (= stride between slots: the arrays stay compact)
    const bool writer = (cr == 0);          // the CTA that publishes messages / challenges / final evaluations
    uint64_t h = a.d_tr_state ? *a.d_tr_state : 0;
    int par = 0;
    for (uint32_t j = a.first_round; j < a.num_rounds; j++) {
        if (dist && n_cur <= ca.nt) {
            // ---- gather the slices into CTA 0 (top half of its region), everyone else leaves
            ext_t* gdst = ct_sm + ((size_t)n_slots * n_loc0 >> 1);
            const uint32_t g0 = dsmem_addr(gdst, 0);
            for (int slot = 0; slot < n_slots; slot++)
                for (uint32_t b = tid; b < n_my; b += CG_CT_THREADS)
                    dsmem_st_ext(g0 + (uint32_t)(((size_t)slot * n_cur + (size_t)cr * n_my + b) * sizeof(ext_t)), base[(size_t)slot * n_my + b]);
            cluster_sync_all();
            if (cr != 0) return;
            base = gdst;
            n_my = n_cur;
            dist = false;
        }
        const uint32_t pairs = n_my >> 1;
        SmemLoader ld{base, n_my, a.t.n_prod};
        CT_STAMP(0);
        if (pairs > CG_CT_SPLIT_PAIRS) {   // throughput mode: one thread evaluates a pair at all three points
            ecacc H[3];
            ecacc_zero(H[0]); ecacc_zero(H[1]); ecacc_zero(H[2]);
            for (uint32_t item = tid; item < pairs; item += CG_CT_THREADS) tower_item<SIMPLE>(a.t, ld, item, H);
            ext_t acc[3] = {ecacc_canon(H[0]), ecacc_canon(H[1]), ecacc_canon(H[2])};
#pragma unroll
            for (int x = 0; x < 3; x++) {
                const ext_t v = warp_reduce_ext(acc[x]);
                if (lane == 0) s_red[warp][x] = v;
            }
        } else {                           // latency mode: three groups of five warps, group t evaluates point t + 1
            const int tg = warp / 5;       // warp 15 idles
            ecacc H1;
            ecacc_zero(H1);
            if (tg < 3)
                for (uint32_t item = tid - 160 * tg; item < pairs; item += 160) tower_item_pt<SIMPLE>(a.t, ld, item, tg, H1);
            const ext_t v = warp_reduce_ext(ecacc_canon(H1));
            if (lane == 0) {
#pragma unroll
                for (int x = 0; x < 3; x++) s_red[warp][x] = (x == tg) ? v : ext_zero();
            }
        }
        __syncthreads();
        CT_STAMP(1);
        if (dist) {
            if (warp == 0) {
                ext_t part[3];
                warp_sum3_smem(s_red, CG_CT_THREADS / 32, part);
                if (lane < (int)C) {   // lane l delivers this CTA's partial to CTA l
                    const uint32_t dst = dsmem_addr(&s_parts[par][cr][0], (uint32_t)lane);
#pragma unroll
                    for (int x = 0; x < 3; x++) dsmem_st_ext(dst + 16u * x, part[x]);
                }
            }
            cluster_sync_all();
        }
        CT_STAMP(2);
        if (warp == 0) {
            ext_t res[3];
            if (dist) warp_sum3_smem(s_parts[par], (int)C, res);
            else warp_sum3_smem(s_red, CG_CT_THREADS / 32, res);
            if (lane == 0) {
                ext_t r;
                CT_STAMP(5);
                if (writer) {
#pragma unroll
                    for (int x = 0; x < 3; x++) a.d_msgs[(size_t)j * 3 + x] = res[x];
                }
                if (a.d_tr_state) {   // every CTA runs the (deterministic) challenger on the same message
                    r = cg_tr_round<3>(h, res);
                } else if (writer) {
                    TailMailbox* mb = a.mail;
                    mailbox_post<3>(mb, res, (uint64_t)j + 1);
                    const long long t0 = clock64();
                    r = ext_zero();
                    while (true) {
                        const int st = mailbox_poll(mb, (uint64_t)j + 1, r);
                        if (st == 1) break;
                        if (st < 0 || (unsigned long long)(clock64() - t0) > a.timeout_cycles) { s_abort = 1; *a.d_error = 1; break; }
                    }
                } else {   // CTA 0 forwards the host's challenge into s_bc (below): spin on local shared memory, not on PCIe
                    const volatile uint64_t* bc = s_bc;
                    const long long t0 = clock64();
                    r = ext_zero();
                    while (true) {
                        const uint64_t c0 = bc[0], c1 = bc[1], fl = bc[2];
                        const uint64_t w2[2] = {c0, c1};
                        if (fl == (((uint64_t)j + 1) ^ cg_mb_mix(w2, 2))) { r = ext_make(c0, c1); break; }
                        if (fl == ~0ULL || (unsigned long long)(clock64() - t0) > 2 * a.timeout_cycles) { s_abort = 1; break; }
                    }
                }
                CT_STAMP(6);
                if (writer) a.d_chal[j] = r;
                s_r = r;
            }
            if (dist && writer && !a.d_tr_state) {   // host transcript: lane l hands the challenge (or the abort) to CTA l
                __syncwarp();
                const int ab = s_abort;
                const ext_t rr = s_r;
                if (lane >= 1 && lane < (int)C) {
                    const uint64_t w2[2] = {rr.c0, rr.c1};
                    const uint32_t dst = dsmem_addr(s_bc, (uint32_t)lane);
                    dsmem_st_ext(dst, rr);
                    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(dst + 16u), "l"(ab ? ~0ULL : (((uint64_t)j + 1) ^ cg_mb_mix(w2, 2))) : "memory");
                }
            }
        }
        __syncthreads();
        CT_STAMP(3);
        if (s_abort) return;
        // ---- fold every MLE in shared memory, compacting: slot s moves from s * n_my to s * n_my / 2
        const extmul_t rm = extmul_prep(s_r);
        constexpr int PER = (CG_CT_MAX_NLOC / 2 + CG_CT_THREADS - 1) / CG_CT_THREADS;
        if ((uint32_t)n_slots * pairs <= (uint32_t)(PER * CG_CT_THREADS)) {   // small round: all slots in one staged pass (one barrier)
            ext_t v[PER];
            const uint32_t total = (uint32_t)n_slots * pairs, plog = 31 - __clz(pairs);
#pragma unroll
            for (int c = 0; c < PER; c++) {
                const uint32_t i = tid + c * CG_CT_THREADS;
                if (i < total) {
                    const uint32_t slot = i >> plog, b = i & (pairs - 1);
                    const ext_t* src = base + (size_t)slot * n_my;
                    const ext_t lo = src[2 * b], hi = src[2 * b + 1];
                    v[c] = ext_fma_prep(lo, ext_sub(hi, lo), rm);
                }
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < PER; c++) {
                const uint32_t i = tid + c * CG_CT_THREADS;
                if (i < total) base[i] = v[c];   // slot * pairs + b == i: the compact layout
            }
        } else {
            for (int slot = 0; slot < n_slots; slot++) {
                const ext_t* src = base + (size_t)slot * n_my;
                ext_t* dst = base + (size_t)slot * pairs;
                ext_t v[PER];
#pragma unroll
                for (int c = 0; c < PER; c++) {
                    const uint32_t b = tid + c * CG_CT_THREADS;
                    if (b < pairs) {
                        const ext_t lo = src[2 * b], hi = src[2 * b + 1];
                        v[c] = ext_fma_prep(lo, ext_sub(hi, lo), rm);
                    }
                }
                __syncthreads();
#pragma unroll
                for (int c = 0; c < PER; c++) {
                    const uint32_t b = tid + c * CG_CT_THREADS;
                    if (b < pairs) dst[b] = v[c];
                }
            }
        }
        __syncthreads();
        CT_STAMP(4);
        n_cur >>= 1;
        n_my = pairs;
        par ^= 1;
    }
    if (!writer) return;
    if (a.d_tr_state && tid == 0) *a.d_tr_state = h;
    for (int slot = tid; slot < n_slots; slot += CG_CT_THREADS) a.d_final[a.final_idx[slot]] = base[slot];
    if (a.mail && tid == 0) {   // host transcript: also post the final evaluations into the mapped mailbox
        volatile uint64_t* f = a.mail->fin;
        uint64_t hh = 0;
        for (int slot = 0; slot < n_slots; slot++) {
            const ext_t v = base[slot];
            const uint32_t w = 2u * a.final_idx[slot];
            f[w] = v.c0;
            f[w + 1] = v.c1;
            hh ^= cg_mb_word(v.c0, w) ^ cg_mb_word(v.c1, w + 1);
        }
        a.mail->fin_flag = ((uint64_t)a.num_rounds + 1) ^ hh;
    }
}

// ---------------------------------------------------------------------------------------------
// Generic monomial-term round evaluation: P = sum_t c_t prod_{i in S_t} f_i, base or ext MLEs
// (the table extract_mle_relationships_from_monomial_terms hands to prove_generic_sumcheck_gpu,
//  gkr_iop/src/gkr/layer/gpu/mod.rs:204-215).  Products of fewer than D factors are evaluated at all
// D points too — bit-identical to extrapolation because field arithmetic is exact.
struct MleSlot {
    const void* ptr;
    uint32_t is_ext;
    uint32_t canon;   // caller-provided buffer: canonicalise on load
};
struct GenericArgs {
    const MleSlot* mles;
    const ext_t* coeff;
    const uint32_t* off;
    const uint32_t* idx;
    uint32_t n_terms;
    uint64_t n_pairs;
    RoundOut out;
};

template <int D>
__global__ void __launch_bounds__(CG_THREADS) generic_round_kernel(const __grid_constant__ GenericArgs a) {
    ext_t acc[D];
#pragma unroll
    for (int x = 0; x < D; x++) acc[x] = ext_zero();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; item < a.n_pairs; item += stride) {
        for (uint32_t t = 0; t < a.n_terms; t++) {
            ext_t prod[D];
            const ext_t c = a.coeff[t];
#pragma unroll
            for (int x = 0; x < D; x++) prod[x] = c;
            const uint32_t q0 = a.off[t], q1 = a.off[t + 1];
            for (uint32_t q = q0; q < q1; q++) {
                const MleSlot s = a.mles[a.idx[q]];
                if (s.is_ext) {
                    ext_t lo = ld_ext(reinterpret_cast<const ext_t*>(s.ptr) + 2 * item);
                    ext_t hi = ld_ext(reinterpret_cast<const ext_t*>(s.ptr) + 2 * item + 1);
                    if (s.canon) { lo = ext_canon(lo); hi = ext_canon(hi); }
                    const ext_t d = ext_sub(hi, lo);
                    ext_t v = hi;
#pragma unroll
                    for (int x = 0; x < D; x++) { prod[x] = ext_mul(prod[x], v); v = ext_add(v, d); }
                } else {
                    const ulonglong2 pr = *(reinterpret_cast<const ulonglong2*>(s.ptr) + item);
                    uint64_t lo = pr.x, hi = pr.y;
                    if (s.canon) { lo = gl_canon(lo); hi = gl_canon(hi); }
                    const uint64_t d = gl_sub(hi, lo);
                    uint64_t v = hi;
#pragma unroll
                    for (int x = 0; x < D; x++) { prod[x] = ext_mul_base(prod[x], v); v = gl_add(v, d); }
                }
            }
#pragma unroll
            for (int x = 0; x < D; x++) acc[x] = ext_add(acc[x], prod[x]);
        }
    }
    block_finish<D>(acc, a.out);
}

// ---------------------------------------------------------------------------------------------
// Mid kernel: the latency-bound rounds between the big streaming rounds and the shared-memory tail
// (2^17 .. 2^12 pairs) run in ONE cooperative persistent launch.  Every block evaluates (and folds) its
// grid-stride share of the round, publishes a partial, and takes a ticket; the block that draws the last
// ticket of the round combines the partials, does the multi-GPU exchange, obtains the challenge (device
// challenger or host mailbox), stores it and releases the round flag the other blocks spin on.  One
// grid-wide hand-off per round replaces launch + finish + host round trip.
struct MidArgs {
    TowerArgs t;                    // spec counts and alphas; r / r_ptr = challenge of the pending entry fold
    const ext_t* buf_odd[CG_COMM_GATHER_SLOTS];    // workspace of each slot: states f odd ...
    const ext_t* buf_even[CG_COMM_GATHER_SLOTS];   // ... and f even (f >= 1; the kernel never touches caller data)
    uint32_t f0;                    // state read by the first round (it is folded into state f0+1)
    uint32_t log_n_in;              // log2(elements per MLE) of state f0
    uint32_t first_round, end_round;
    ext_t* d_msgs;                  // [round * 3]
    ext_t* d_chal;                  // [round]
    uint64_t* d_tr_state;           // device challenger, or nullptr -> host mailbox
    TailMailbox* mail;
    int* d_error;
    unsigned long long timeout_cycles;
    CommDev comm;                   // exchange of round j uses sequence comm.seq + (j - first_round)
    ext_t* partials;                // [gridDim.x * 3]
    unsigned int* ticket;           // monotonically increasing across rounds (zero on entry, reset on exit)
    volatile unsigned int* round_flag;   // = index of the last finished round + 1
};
struct MidLoader {
    const MidArgs& a;
    uint32_t f_in;
    extmul_t rm;
    GL_DEV const ext_t* in(int slot) const { return (f_in & 1) ? a.buf_odd[slot] : a.buf_even[slot]; }
    GL_DEV ext_t* out(int slot) const { return const_cast<ext_t*>((f_in & 1) ? a.buf_even[slot] : a.buf_odd[slot]); }
    GL_DEV void eq(uint64_t item, ext_t& lo, ext_t& hi) { load_pair<true, false>(in(0), out(0), item, rm, lo, hi); }
    GL_DEV void prod(int p, int z, uint64_t item, ext_t& lo, ext_t& hi) { const int s = 1 + 2 * p + z; load_pair<true, false>(in(s), out(s), item, rm, lo, hi); }
    GL_DEV void lk(int l, int z, uint64_t item, ext_t& lo, ext_t& hi) { const int s = 1 + 2 * a.t.n_prod + 4 * l + z; load_pair<true, false>(in(s), out(s), item, rm, lo, hi); }
};
template <bool SIMPLE>
__global__ void __launch_bounds__(CG_THREADS, 2) tower_mid_kernel(const __grid_constant__ MidArgs a) {
    __shared__ ext_t s_part[CG_THREADS / 32][3];
    __shared__ __align__(32) uint64_t s_msg[2 * 3 + 8];
    __shared__ ext_t s_r;
    __shared__ int s_flag;   // 1: this block drew the last ticket, 2: abort
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ext_t r = a.t.r_ptr ? ext_canon(ld_ext(a.t.r_ptr)) : a.t.r;
    for (uint32_t j = a.first_round; j < a.end_round; j++) {
        const uint32_t step = j - a.first_round;
        MidLoader ld{a, a.f0 + step, extmul_prep(r)};
        const uint64_t n_pairs = 1ULL << (a.log_n_in - step - 2);
        ecacc H[3];
        ecacc_zero(H[0]); ecacc_zero(H[1]); ecacc_zero(H[2]);
        const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
        for (uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_pairs; item += stride)
            tower_item<SIMPLE>(a.t, ld, item, H);
        ext_t acc[3] = {ecacc_canon(H[0]), ecacc_canon(H[1]), ecacc_canon(H[2])};
#pragma unroll
        for (int x = 0; x < 3; x++) {
            const ext_t v = warp_reduce_ext(acc[x]);
            if (lane == 0) s_part[warp][x] = v;
        }
        if (threadIdx.x == 0) s_flag = 0;
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int x = 0; x < 3; x++) {
                const ext_t v = warp_reduce_ext(lane < (CG_THREADS / 32) ? s_part[lane][x] : ext_zero());
                if (lane == 0) a.partials[(size_t)blockIdx.x * 3 + x] = v;
            }
            if (lane == 0) {
                __threadfence();   // partial + this block's folded outputs are visible before the ticket
                const unsigned tk = atomicAdd(a.ticket, 1u);
                if (tk == gridDim.x * (step + 1) - 1) s_flag = 1;
            }
        }
        __syncthreads();
        if (s_flag == 1) {   // last block of the round: combine, exchange, challenge, release
            __threadfence();
            ext_t res[3];
#pragma unroll
            for (int x = 0; x < 3; x++) {
                ext_t v = ext_zero();
                for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
                    const ulonglong2 p = __ldcg(reinterpret_cast<const ulonglong2*>(&a.partials[(size_t)b * 3 + x]));
                    v = ext_add(v, ext_make(p.x, p.y));
                }
                v = warp_reduce_ext(v);
                if (lane == 0) s_part[warp][x] = v;
            }
            __syncthreads();
            if (warp == 0) {
#pragma unroll
                for (int x = 0; x < 3; x++) res[x] = warp_reduce_ext(lane < (CG_THREADS / 32) ? s_part[lane][x] : ext_zero());
                if (a.comm.nranks > 1) comm_exchange<3>(res, a.comm, a.comm.seq + step, s_msg);
                if (lane == 0) {
                    ext_t rn = ext_zero();
#pragma unroll
                    for (int x = 0; x < 3; x++) a.d_msgs[(size_t)j * 3 + x] = res[x];
                    bool abort = false;
                    if (a.d_tr_state) {
                        uint64_t h = *a.d_tr_state;
                        rn = cg_tr_round<3>(h, res);
                        *a.d_tr_state = h;
                    } else {
                        TailMailbox* mb = a.mail;
                        mailbox_post<3>(mb, res, (uint64_t)j + 1);
                        const long long t0 = clock64();
                        while (true) {
                            const int st = mailbox_poll(mb, (uint64_t)j + 1, rn);
                            if (st == 1) break;
                            if (st < 0 || (unsigned long long)(clock64() - t0) > a.timeout_cycles) { abort = true; *a.d_error = 1; break; }
                        }
                    }
                    a.d_chal[j] = rn;
                    __threadfence();
                    *a.round_flag = abort ? 0xFFFFFFFFu : (j + 1);   // release
                }
            }
        }
        // every block: wait for the round to be released, pick up the challenge
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            unsigned fl;
            while ((fl = *a.round_flag) < j + 1) {
                if ((unsigned long long)(clock64() - t0) > 2 * a.timeout_cycles) { fl = 0xFFFFFFFFu; break; }
            }
            if (fl == 0xFFFFFFFFu) s_flag = 2;
            else {
                const ulonglong2 p = __ldcg(reinterpret_cast<const ulonglong2*>(&a.d_chal[j]));
                s_r = ext_make(p.x, p.y);
            }
        }
        __syncthreads();
        if (s_flag == 2) return;
        __threadfence();   // acquire: the other blocks' folded outputs (gpu-scope fence also invalidates L1)
        r = s_r;
    }
}

// ---------------------------------------------------------------------------------------------
// Grouped monomial-term evaluation (the zerocheck shape, gkr_iop/src/gkr/layer/zerocheck_layer.rs:86-207):
//   P = sum_g [ prod_{e in E_g} e(x) ] * [ sum_{t in g} c_t prod_{w in W_t} w(x) ]
// Terms are grouped on the host by their factors that are ext-field MLEs in round 0 (selectors / eq) — the
// factoring the reference's CommonTermPlan performs (gkr_iop/src/gkr/layer/gpu/mod.rs:186-230).  The inner sum
// is accumulated UNREDUCED (one reduction per group and point); in round 0 the witness products W_t are
// base-field products (1 multiply per factor and point instead of 4); the group's last ext factor is
// multiplied straight into the thread's compact round-sum accumulators.
struct GroupedArgs {
    const MleSlot* mles;          // state of every MLE this round
    const uint32_t* g_term_off;   // [G+1] term range of group g
    const uint32_t* g_ext_off;    // [G+1] range into g_ext_idx
    const uint32_t* g_ext_idx;    // shared ext factors
    const uint64_t* t_coeff;      // [T][3] = c0, c1, 7*c1 (canonical)
    const uint32_t* t_off;        // [T+1] range into t_idx
    const uint32_t* t_idx;        // residual (witness) factors
    uint32_t n_groups;
    uint64_t n_pairs;
    RoundOut out;
};
template <int D, bool PHASE0>
__global__ void __launch_bounds__(CG_THREADS, 2) grouped_round_kernel(const __grid_constant__ GroupedArgs a) {
    ecacc H[D];
#pragma unroll
    for (int x = 0; x < D; x++) ecacc_zero(H[x]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; item < a.n_pairs; item += stride) {
        for (uint32_t g = blockIdx.y; g < a.n_groups; g += gridDim.y) {   // small rounds: groups spread over blockIdx.y
            ecacc inner[D];
#pragma unroll
            for (int x = 0; x < D; x++) ecacc_zero(inner[x]);
            const uint32_t t0 = a.g_term_off[g], t1 = a.g_term_off[g + 1];
            for (uint32_t t = t0; t < t1; t++) {
                const uint64_t c0 = a.t_coeff[3 * t], c1 = a.t_coeff[3 * t + 1], c1_7 = a.t_coeff[3 * t + 2];
                const uint32_t q0 = a.t_off[t], q1 = a.t_off[t + 1];
                if (q0 == q1) {   // constant residual: W = 1
#pragma unroll
                    for (int x = 0; x < D; x++) { acc_t T; acc_set64(T, c0); cacc_add(inner[x].A0, T); acc_set64(T, c1); cacc_add(inner[x].A1, T); }
                    continue;
                }
                if (PHASE0) {     // witnesses are base-field: the product stays in the base field
                    uint64_t w[D];
                    for (uint32_t q = q0; q < q1; q++) {
                        const MleSlot sl = a.mles[a.t_idx[q]];
                        const ulonglong2 pr = *(reinterpret_cast<const ulonglong2*>(sl.ptr) + item);
                        const uint64_t lo = gl_canon(pr.x);
                        uint64_t v = gl_canon(pr.y);
                        const uint64_t nd = gl_sub(lo, v);
#pragma unroll
                        for (int x = 0; x < D; x++) {
                            w[x] = (q == q0) ? v : gl_mul_weak(w[x], v);
                            if (x + 1 < D) v = gl_sub(v, nd);
                        }
                    }
#pragma unroll
                    for (int x = 0; x < D; x++) {
                        acc_t T;
                        acc_zero(T); acc_mac(T, c0, w[x]); cacc_add(inner[x].A0, T);
                        acc_zero(T); acc_mac(T, c1, w[x]); cacc_add(inner[x].A1, T);
                    }
                } else {
                    ext_t w[D];
                    for (uint32_t q = q0; q < q1; q++) {
                        const MleSlot sl = a.mles[a.t_idx[q]];
                        ext_t lo, v;
                        ld_ext2(reinterpret_cast<const ext_t*>(sl.ptr) + 2 * item, lo, v);
                        const ext_t nd = ext_sub(lo, v);
#pragma unroll
                        for (int x = 0; x < D; x++) {
                            w[x] = (q == q0) ? v : ext_mul_weak(w[x], v);
                            if (x + 1 < D) v = ext_sub(v, nd);
                        }
                    }
                    extmul_t cm; cm.c0 = c0; cm.c1 = c1; cm.c1_7 = c1_7;
#pragma unroll
                    for (int x = 0; x < D; x++) {
                        eacc T;
                        eacc_zero(T);
                        eacc_mac_prep(T, w[x], cm);
                        ecacc_add(inner[x], T);
                    }
                }
            }
            const uint32_t e0 = a.g_ext_off[g], e1 = a.g_ext_off[g + 1];
            if (e0 == e1) {   // no shared ext factor: the inner sums go straight into the round sums
#pragma unroll
                for (int x = 0; x < D; x++) {
                    acc_t T;
                    acc_set64(T, cacc_weak(inner[x].A0)); cacc_add(H[x].A0, T);
                    acc_set64(T, cacc_weak(inner[x].A1)); cacc_add(H[x].A1, T);
                }
                continue;
            }
            ext_t u[D];
#pragma unroll
            for (int x = 0; x < D; x++) u[x] = ext_make(cacc_weak(inner[x].A0), cacc_weak(inner[x].A1));
            for (uint32_t q = e0; q < e1; q++) {
                const MleSlot sl = a.mles[a.g_ext_idx[q]];
                ext_t lo, v;
                ld_ext2(reinterpret_cast<const ext_t*>(sl.ptr) + 2 * item, lo, v);
                if (PHASE0) { lo = ext_canon(lo); v = ext_canon(v); }
                const ext_t nd = ext_sub(lo, v);
                const bool last = (q + 1 == e1);
#pragma unroll
                for (int x = 0; x < D; x++) {
                    if (last) {
                        eacc T;
                        eacc_zero(T);
                        eacc_mac(T, u[x], v, gl_mul7_weak(v.c1));
                        ecacc_add(H[x], T);
                    } else {
                        u[x] = ext_mul_weak(u[x], v);
                    }
                    if (x + 1 < D) v = ext_sub(v, nd);
                }
            }
        }
    }
    ext_t acc[D];
#pragma unroll
    for (int x = 0; x < D; x++) acc[x] = ecacc_canon(H[x]);
    block_finish<D>(acc, a.out);
}

// ---------------------------------------------------------------------------------------------
// fold of many MLEs in one launch: grid.y = MLE.  Output always ext.
struct FoldSlot {
    const void* in;
    ext_t* out;
    uint32_t is_ext;
    uint32_t canon;
};
__global__ void __launch_bounds__(CG_THREADS) fold_kernel(const FoldSlot* __restrict__ slots, uint64_t n_out,
                                                           ext_t r_val, const ext_t* r_ptr) {
    const FoldSlot s = slots[blockIdx.y];
    const extmul_t rm = extmul_prep(r_ptr ? ld_ext(r_ptr) : r_val);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_out; b += stride) {
        if (s.is_ext) {
            ext_t lo, hi;
            ld_ext2(reinterpret_cast<const ext_t*>(s.in) + 2 * b, lo, hi);
            if (s.canon) { lo = ext_canon(lo); hi = ext_canon(hi); }
            st_ext(s.out + b, ext_add(lo, ext_mul_prep(ext_sub(hi, lo), rm)));
        } else {
            const ulonglong2 pr = *(reinterpret_cast<const ulonglong2*>(s.in) + b);
            uint64_t lo = pr.x, hi = pr.y;
            if (s.canon) { lo = gl_canon(lo); hi = gl_canon(hi); }
            const uint64_t d = gl_sub(hi, lo);
            st_ext(s.out + b, ext_make(gl_add(lo, gl_mul(d, rm.c0)), gl_mul(d, rm.c1)));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// build_eq_x_r.  Small table: one block, doubling in shared memory (k <= 12).
// eq[b + 2^i] = eq[b] r_i ; eq[b] -= eq[b + 2^i]   (any order gives the same bits — exact field).
#define CG_EQ_SMALL_K 12
__global__ void __launch_bounds__(1024) eq_small_kernel(const ext_t* __restrict__ point, uint32_t k, ext_t* __restrict__ out) {
    extern __shared__ ext_t s_eq[];
    if (threadIdx.x == 0) s_eq[0] = ext_one();
    __syncthreads();
    for (uint32_t i = 0; i < k; i++) {
        const ext_t ri = ext_canon(point[i]);
        const uint32_t n = 1u << i;
        for (uint32_t b = threadIdx.x; b < n; b += blockDim.x) {
            const ext_t lo = s_eq[b];
            const ext_t hi = ext_mul(lo, ri);
            s_eq[b + n] = hi;
            s_eq[b] = ext_sub(lo, hi);
        }
        __syncthreads();
    }
    for (uint32_t b = threadIdx.x; b < (1u << k); b += blockDim.x) out[b] = s_eq[b];
}
// Large table: out[hi * 2^lo_k + lo] = L[lo] * H[hi], prefix mask [start, end) fused.
// One ext product per output is ~65 instructions for 16 bytes written, so the kernel is issue-bound, not write-bound: a thread
// keeps ONE H entry (with 7 c1 prepared once) for 8 consecutive outputs and the four independent product pairs overlap.
#define CG_EQ_PER_THREAD 8
__global__ void __launch_bounds__(CG_THREADS) eq_outer_kernel(const ext_t* __restrict__ L, const ext_t* __restrict__ H,
                                                               uint32_t lo_k, uint64_t n, uint64_t start, uint64_t end,
                                                               ext_t* __restrict__ out) {
    const uint64_t lo_mask = (1ULL << lo_k) - 1;
    constexpr uint64_t CHUNK = (uint64_t)CG_THREADS * CG_EQ_PER_THREAD;   // 2048 consecutive outputs per block step: one H entry (lo_k >= 11)
    for (uint64_t c0 = (uint64_t)blockIdx.x * CHUNK; c0 < n; c0 += (uint64_t)gridDim.x * CHUNK) {
        const extmul_t hm = extmul_prep(ld_ext(H + (c0 >> lo_k)));
#pragma unroll
        for (int q = 0; q < CG_EQ_PER_THREAD / 2; q++) {
            const uint64_t b = c0 + (uint64_t)q * (2 * CG_THREADS) + 2 * threadIdx.x;   // a warp covers 1 KB contiguous per step
            ext_t l0, l1;
            ld_ext2(L + (b & lo_mask), l0, l1);
            ext_t v0 = ext_mul_prep(l0, hm), v1 = ext_mul_prep(l1, hm);
            if (b < start || b >= end) v0 = ext_zero();
            if (b + 1 < start || b + 1 >= end) v1 = ext_zero();
            st_ext2(out + b, v0, v1);
        }
    }
}
// generic mask passes for the selector variants (gkr_iop/src/selector.rs:141-243)
__global__ void prefix_mask_kernel(ext_t* __restrict__ v, uint64_t n, uint64_t start, uint64_t end) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride)
        if (b < start || b >= end) v[b] = ext_zero();
}
// OrderedSparse: keep[i] bitmap over the inner 2^inner_vars indices (device), zero chunks >= num_instances
__global__ void sparse_mask_kernel(ext_t* __restrict__ v, uint64_t n, uint32_t inner_vars, uint64_t num_instances,
                                   const uint8_t* __restrict__ keep) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t mask = (1ULL << inner_vars) - 1;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride) {
        const uint64_t chunk = b >> inner_vars;
        if (chunk >= num_instances || !keep[b & mask]) v[b] = ext_zero();
    }
}
// QuarkBinaryTreeLessThan: region i = [n - n/2^i ... ) of length n/2^(i+1), keep first seq[i] entries
struct QuarkArgs {
    uint64_t seq[64];
    uint32_t num_vars;
};
__global__ void quark_mask_kernel(ext_t* __restrict__ v, uint64_t n, const __grid_constant__ QuarkArgs q) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride) {
        if (b == n - 1) { v[b] = ext_zero(); continue; }
        // region index i: b in [n - n>>i, n - n>>(i+1))  <=>  i = number of leading ones of b (k-bit)
        const uint64_t inv = (~b) & (n - 1);
        const uint32_t i = q.num_vars - 1 - (63 - __clzll(inv));   // inv != 0 because b != n-1
        const uint64_t region_start = n - (n >> i);
        const uint64_t keep = i < q.num_vars ? q.seq[i] : 0;
        if (b - region_start >= keep) v[b] = ext_zero();
    }
}

// EC-sum Quark pre-passes (CpuEccProver::create_ecc_proof, ceno_zkvm/src/scheme/cpu/mod.rs:100-133, 138-162).
// On entry sel_add = the QuarkBinaryTreeLessThan selector (masked eq), sel_bypass = the unmasked eq(out_rt, .) table.
//   sel_bypass[b] = 0 where sel_add[b] != 0 and at the last index;  sel_export = one-hot at n - 2 carrying
//   eq_eval(out_rt, (0,1,...,1)), which IS the eq table entry at n - 2 (index bit i = variable i).
__global__ void ecc_selectors_kernel(const ext_t* __restrict__ sel_add, ext_t* __restrict__ sel_bypass, ext_t* __restrict__ sel_export, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride) {
        const ext_t a = ld_ext(sel_add + b), e = ld_ext(sel_bypass + b);
        const bool add_on = (a.c0 | a.c1) != 0;
        st_ext(sel_export + b, (n >= 2 && b == n - 2) ? e : ext_zero());
        if (add_on || b == n - 1) st_ext(sel_bypass + b, ext_zero());
    }
}
// filter_bj (cpu/mod.rs:138-152): even[b] = v[2b], odd[b] = v[2b+1] for base-field vectors; one 16-byte load per pair
struct SplitArgs {
    const uint64_t* const* in;     // device array of n_mles pointers
    uint64_t* const* even;
    uint64_t* const* odd;
    uint64_t n_out;                // pairs per MLE
};
__global__ void __launch_bounds__(256) split_even_odd_kernel(const __grid_constant__ SplitArgs a) {
    const uint64_t* __restrict__ in = a.in[blockIdx.y];
    uint64_t* __restrict__ ev = a.even[blockIdx.y];
    uint64_t* __restrict__ od = a.odd[blockIdx.y];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.n_out; b += stride) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(in + 2 * b);
        ev[b] = gl_canon(v.x);
        od[b] = gl_canon(v.y);
    }
}

// ---------------------------------------------------------------------------------------------
// interleaving_mles_to_mles (ceno_zkvm/src/scheme/utils.rs:402-462): R record MLEs (one value per instance)
// become `num_limbs` tower leaves, out[limb][s * 2^ceil_log2(R) + i] = mle_i[limb * per_fanin_len + s], padded with
// `def`.  A 32 x 32 shared-memory tile transpose: reads run along the instances of one MLE, writes along the
// records of one instance, both coalesced.
struct InterleaveArgs {
    const void* const* ptrs;     // device array of n_mles pointers
    const uint32_t* is_ext;      // device array
    uint32_t n_mles, l2m;        // per_instance = 2^l2m
    uint64_t out_len, per_fanin_len, num_instances, mle_len;
    ext_t def;
    ext_t* out;                  // [num_limbs][out_len]
};
__global__ void __launch_bounds__(256) tower_interleave_kernel(const __grid_constant__ InterleaveArgs a) {
    __shared__ ext_t tile[32][33];
    const uint32_t tx = threadIdx.x, ty = threadIdx.y;
    const uint64_t fi = blockIdx.z;
    const uint64_t per_instance = 1ULL << a.l2m, n_inst_out = a.out_len >> a.l2m;
    const uint64_t start = a.per_fanin_len * fi;
    const uint64_t valid = start < a.num_instances ? (a.per_fanin_len < a.num_instances - start ? a.per_fanin_len : a.num_instances - start) : 0;
    const uint64_t s0 = (uint64_t)blockIdx.x * 32, i0 = (uint64_t)blockIdx.y * 32;
    for (uint32_t ii = ty; ii < 32; ii += 8) {
        const uint64_t i = i0 + ii, sidx = s0 + tx;
        ext_t v = a.def;
        if (i < a.n_mles && start < a.num_instances) {
            const uint32_t e = a.is_ext[i];
            uint64_t cnt = e ? valid : a.per_fanin_len;     // Ext arm: valid_instances_len, Base arm: per_fanin_len (utils.rs:433-456)
            if (start + cnt > a.mle_len) cnt = 0;           // `.get(range)` out of range -> empty
            if (cnt > n_inst_out) cnt = n_inst_out;
            if (sidx < cnt) {
                if (e) v = ext_canon(ld_ext(reinterpret_cast<const ext_t*>(a.ptrs[i]) + start + sidx));
                else v = ext_make(gl_canon(reinterpret_cast<const uint64_t*>(a.ptrs[i])[start + sidx]), 0);
            }
        }
        tile[ii][tx] = v;
    }
    __syncthreads();
    for (uint32_t ss = ty; ss < 32; ss += 8) {
        const uint64_t sidx = s0 + ss, i = i0 + tx;
        if (sidx < n_inst_out && i < per_instance) st_ext(a.out + fi * a.out_len + sidx * per_instance + i, tile[tx][ss]);
    }
}

// ---------------------------------------------------------------------------------------------
// tower witness layers (infer_tower_product_witness / infer_tower_logup_witness,
// ceno_zkvm/src/scheme/utils.rs:488-659).  Layer l buffer = [a | b] (product) or [p1|p2|q1|q2]
// (logup), each 2^l ext; layer l = pointwise combination of layer l+1's arrays over 2^(l+1) points.
// Where the n results of a launch go.  One device: d[0].  SHARDED tower (rows sliced over the ranks): the
// combined array is the next layer's (low half | high half), and the rank that must hold a slice of both halves is not the
// rank that computed it, so the layer kernel stores straight into the owners' peer-mapped buffers over NVLink —
//   split != 0: results [0, n/2) -> d[0], [n/2, n) -> d[1]   (perfect shuffle: ranks 2r and 2r+1 mod N);
//   split == 0: every result -> d[0 .. n_dst)                (all-gather into the first replicated layer).
struct TowerDst {
    ext_t* d[CG_MAX_RANKS];
    int n_dst;
    int split;
};
GL_DEV void tower_store(const TowerDst& t, uint64_t x, uint64_t n, ext_t v) {
    if (t.split) {
        const uint64_t h = n >> 1;
        st_ext((x < h ? t.d[0] : t.d[1]) + (x < h ? x : x - h), v);
    } else {
#pragma unroll 1
        for (int p = 0; p < t.n_dst; p++) st_ext(t.d[p] + x, v);
    }
}
__global__ void __launch_bounds__(CG_THREADS) tower_prod_layer_kernel(const ext_t* __restrict__ a, const ext_t* __restrict__ b,
                                                                       uint64_t n, const __grid_constant__ TowerDst out, int canon) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) {
        ext_t u = ld_ext(a + x), v = ld_ext(b + x);
        if (canon) { u = ext_canon(u); v = ext_canon(v); }
        tower_store(out, x, n, ext_mul(u, v));
    }
}
// p_out[x] = q1 p2 + q2 p1 (or q1 + q2 when numerators are implicit ones), q_out[x] = q1 q2
__global__ void __launch_bounds__(CG_THREADS) tower_logup_layer_kernel(const ext_t* __restrict__ p1, const ext_t* __restrict__ p2,
                                                                        const ext_t* __restrict__ q1, const ext_t* __restrict__ q2,
                                                                        uint64_t n, const __grid_constant__ TowerDst p_out,
                                                                        const __grid_constant__ TowerDst q_out, int canon) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) {
        ext_t a1 = ld_ext(q1 + x), a2 = ld_ext(q2 + x);
        if (canon) { a1 = ext_canon(a1); a2 = ext_canon(a2); }
        ext_t p;
        if (p1) {
            ext_t u1 = ld_ext(p1 + x), u2 = ld_ext(p2 + x);
            if (canon) { u1 = ext_canon(u1); u2 = ext_canon(u2); }
            p = ext_add(ext_mul(a1, u2), ext_mul(a2, u1));
        } else {
            p = ext_add(a1, a2);
        }
        tower_store(p_out, x, n, p);
        tower_store(q_out, x, n, ext_mul(a1, a2));
    }
}
// first level above VIRTUAL leaves (description instead of arrays; p1 / p2 null = numerators are all one).
// Iteration order: neighbouring leaves are DIFFERENT record arrays (leaf = row << l2m | record), so consecutive lanes take
// consecutive rows of one record pair (contiguous 16-byte reads) and every thread writes the pair's two results as one
// 32-byte store.
GL_DEV void tower_store2(const TowerDst& t, uint64_t x, uint64_t n, ext_t v0, ext_t v1) {   // x even
    if (t.split) {
        const uint64_t h = n >> 1;
        st_ext2((x < h ? t.d[0] : t.d[1]) + (x < h ? x : x - h), v0, v1);
    } else {
#pragma unroll 1
        for (int p = 0; p < t.n_dst; p++) st_ext2(t.d[p] + x, v0, v1);
    }
}
GL_DEV uint64_t virt_pair_index(uint64_t g, uint64_t n, uint32_t l2m) {   // unit g of n / 2 -> even leaf index
    const uint32_t rows_log = 63 - __clzll((long long)n) - l2m;
    if (l2m < 3 || rows_log < 3) return 2 * g;
    // four neighbouring lanes: four consecutive leaf pairs of one row (one 128-byte line of results); the warp's eight lane
    // quads: eight consecutive rows (one 128-byte line of every record read)
    const uint64_t q = g & 3, u = g >> 2;
    const uint64_t s = u & ((1ULL << rows_log) - 1), i8 = u >> rows_log;
    return (s << l2m) | (8 * i8 + 2 * q);
}
__global__ void __launch_bounds__(CG_THREADS) tower_prod_layer_virt_kernel(const VirtLeaf* __restrict__ a, const VirtLeaf* __restrict__ b,
                                                                            uint64_t n, const __grid_constant__ TowerDst out) {
    const VirtLeaf va = *a, vb = *b;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if (n < 2) {
        if (blockIdx.x == 0 && threadIdx.x == 0) tower_store(out, 0, n, ext_mul(virt_leaf(va, 0), virt_leaf(vb, 0)));
        return;
    }
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n / 2; g += stride) {
        const uint64_t x = virt_pair_index(g, n, va.l2m);
        tower_store2(out, x, n, ext_mul(virt_leaf(va, x), virt_leaf(vb, x)), ext_mul(virt_leaf(va, x + 1), virt_leaf(vb, x + 1)));
    }
}
__global__ void __launch_bounds__(CG_THREADS) tower_logup_layer_virt_kernel(const VirtLeaf* __restrict__ p1, const VirtLeaf* __restrict__ p2,
                                                                             const VirtLeaf* __restrict__ q1, const VirtLeaf* __restrict__ q2,
                                                                             uint64_t n, const __grid_constant__ TowerDst p_out,
                                                                             const __grid_constant__ TowerDst q_out) {
    const VirtLeaf v1 = *q1, v2 = *q2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // numerators described as "no records, default one" (utils.rs:556-577) need no multiplication
    const bool ones = !p1 || (p1->n_records == 0 && p2->n_records == 0 && p1->def.c0 == 1 && p1->def.c1 == 0 && p2->def.c0 == 1 && p2->def.c1 == 0);
    auto one = [&](uint64_t x, ext_t& p, ext_t& q) {
        const ext_t a1 = virt_leaf(v1, x), a2 = virt_leaf(v2, x);
        if (!ones) p = ext_add(ext_mul(a1, virt_leaf(*p2, x)), ext_mul(a2, virt_leaf(*p1, x)));
        else p = ext_add(a1, a2);
        q = ext_mul(a1, a2);
    };
    if (n < 2) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            ext_t p, q;
            one(0, p, q);
            tower_store(p_out, 0, n, p);
            tower_store(q_out, 0, n, q);
        }
        return;
    }
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n / 2; g += stride) {
        const uint64_t x = virt_pair_index(g, n, v1.l2m);
        ext_t pa, qa, pb, qb;
        one(x, pa, qa);
        one(x + 1, pb, qb);
        tower_store2(p_out, x, n, pa, pb);
        tower_store2(q_out, x, n, qa, qb);
    }
}
// ---------------------------------------------------------------------------------------------
// Rotation pre-passes (gkr_iop/src/utils.rs:19-76; GPU call sites rotation_next_base_mle_gpu / rotation_selector_gpu,
// gkr_iop/src/gkr/layer/gpu/utils.rs:231-336).  The BooleanHypercube order is x -> x*X mod (X^5+X^2+1) resp.
// (X^6+X+1) (gkr_iop/src/gkr/booleanhypercube.rs:10-118): a one-step shift register, no table needed.
GL_DEV uint32_t bh_next(uint32_t x, uint32_t log2) {
    x <<= 1;
    if (x >> log2) x ^= (log2 == 5 ? 0x25u : 0x43u);
    return x;
}
// out[chunk + x] = in[chunk + next(x)] for x != 0, out[chunk] = in[chunk]   (base field)
__global__ void rotation_next_base_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint64_t n, uint32_t log2) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, mask = (1ULL << log2) - 1;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride) {
        const uint32_t x = (uint32_t)(b & mask);
        out[b] = gl_canon(in[(b & ~mask) | (x ? bh_next(x, log2) : 0u)]);
    }
}
// keep[x] bit set for the first `subgroup` elements of the group
__global__ void rotation_selector_kernel(const ext_t* __restrict__ eq, ext_t* __restrict__ out, uint64_t n, uint32_t log2, uint64_t keep) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, mask = (1ULL << log2) - 1;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride)
        st_ext(out + b, ((keep >> (b & mask)) & 1) ? ext_canon(ld_ext(eq + b)) : ext_zero());
}
__global__ void __launch_bounds__(CG_THREADS) scale_ext_kernel(ext_t* __restrict__ v, uint64_t n, ext_t c) {
    const extmul_t cm = extmul_prep(c);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride) st_ext(v + b, ext_mul_prep(ext_canon(ld_ext(v + b)), cm));
}
__global__ void fill_ext_kernel(ext_t* __restrict__ v, uint64_t n, ext_t val) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride) v[b] = val;
}

// ---------------------------------------------------------------------------------------------
// wit_infer_by_monomial_expr: out[b] = sum_t c_t prod f_i[b]   (gkr_iop/src/gpu/mod.rs:599-609)
struct InferArgs {
    const MleSlot* mles;
    const ext_t* coeff;
    const uint32_t* off;
    const uint32_t* idx;
    uint32_t n_terms;
    uint64_t n;
    ext_t* out;
};
__global__ void __launch_bounds__(CG_THREADS) wit_infer_kernel(const __grid_constant__ InferArgs a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.n; b += stride) {
        ext_t acc = ext_zero();
        for (uint32_t t = 0; t < a.n_terms; t++) {
            ext_t prod = a.coeff[t];
            for (uint32_t q = a.off[t]; q < a.off[t + 1]; q++) {
                const MleSlot s = a.mles[a.idx[q]];
                if (s.is_ext) prod = ext_mul(prod, ext_canon(ld_ext(reinterpret_cast<const ext_t*>(s.ptr) + b)));
                else prod = ext_mul_base(prod, reinterpret_cast<const uint64_t*>(s.ptr)[b]);
            }
            acc = ext_add(acc, prod);
        }
        st_ext(a.out + b, acc);
    }
}
