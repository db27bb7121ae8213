"""Multi-GPU sumcheck: the boolean hypercube sharded across the GPUs of one box.

Layout: with LSB-first binding the top g = log2(N) index bits are bound last, so rank q
owns the contiguous slice [q 2^(k-g), (q+1) 2^(k-g)) of every MLE.  Rounds 0 .. k-g-1 fold locally;
the only exchange per round is the d extension-field partial sums (48 B for d = 3), combined by
modular addition — NCCL has no F_p reduction, so the partials are all-gathered and summed (identical
on every rank, which then run the transcript redundantly; no broadcast).  After k-g rounds every rank
holds one element per MLE; those N*m elements are all-gathered and the last g rounds run replicated.

Two exchange engines:
  * host-orchestrated (this file): `exchange` = torch.distributed all_gather (NCCL on GPUs, gloo in
    the CPU tests).  The protocol logic is independent of the local prover, which is pluggable.
  * in-kernel over NVLink peer memory (cg_comm_* in the C ABI): the round kernel's last block stores
    its partial into every peer's mailbox and combines — no host round trip.
"""
import numpy as np

P = 0xFFFFFFFF00000001


def modsum_ext(parts):
    """Sum of ext-element arrays (u64 limbs) mod p, limb-wise."""
    acc = [0] * parts[0].size
    for p in parts:
        for i, v in enumerate(p.reshape(-1)):
            acc[i] = (acc[i] + int(v)) % P
    return np.array(acc, dtype=np.uint64)


def ext_mul_host(a, b):
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def eq_slice_scalar(w_high, q):
    """prod_i (q_i w_i + (1 - q_i)(1 - w_i)) over the top variables: the factor of eq on rank q's slice."""
    acc = (1, 0)
    for i in range(len(w_high) // 2):
        wi = (int(w_high[2 * i]), int(w_high[2 * i + 1]))
        f = wi if (q >> i) & 1 else ((1 - wi[0]) % P, (-wi[1]) % P)
        acc = ext_mul_host(acc, f)
    return acc


class TorchExchange:
    """all_gather of a small u64 array through torch.distributed (NCCL -> CUDA staging, gloo -> CPU)."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, device
        self.world = dist.get_world_size()

    def __call__(self, arr):
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).copy())
        if self.device is not None:
            t = t.to(self.device)
        outs = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t)
        return [o.cpu().numpy().view(np.uint64) for o in outs]


def sharded_prove(local, k_local, g, degree, transcript, exchange, make_tail):
    """Run the sharded protocol on this rank.

    local     : local prover over this rank's slice: round_eval() -> 2d u64, bind(r), final_evals() -> [m, 2]
    transcript: host transcript with sumcheck_begin/round (replicated on every rank)
    exchange  : arr -> list of every rank's arr (rank order)
    make_tail : arrays [m][2N u64] -> local-prover-like object over the gathered N-element MLEs
    Returns (round_evals [k, d, 2], final_evals [m, 2], challenges [k, 2]) — identical on every rank and
    identical to the single-device proof."""
    k = k_local + g
    transcript.append_message(int(k).to_bytes(8, "little"))
    transcript.append_message(int(degree).to_bytes(8, "little"))
    msgs, chals = [], []

    def one_round(prover, combine):
        part = prover.round_eval()
        msg = modsum_ext(exchange(part)) if combine else part
        transcript.append_field_element_exts(msg)
        r = transcript.sample_and_append_challenge(b"Internal round")
        prover.bind(r)
        msgs.append(msg.copy())
        chals.append(np.array(r, dtype=np.uint64))

    for _ in range(k_local):
        one_round(local, True)
    fin = np.ascontiguousarray(local.final_evals(), dtype=np.uint64)           # [m, 2]
    if g == 0:
        final = fin
    else:
        allfin = exchange(fin.reshape(-1))                                        # N x (m*2)
        m = fin.shape[0]
        arrays = [np.concatenate([allfin[q][2 * i:2 * i + 2] for q in range(len(allfin))]) for i in range(m)]
        tail = make_tail(arrays)
        for _ in range(g):
            one_round(tail, False)
        final = np.ascontiguousarray(tail.final_evals(), dtype=np.uint64)
    return (np.array(msgs, dtype=np.uint64).reshape(k, degree, 2), final.reshape(-1, 2), np.array(chals, dtype=np.uint64).reshape(k, 2))


class GpuLocalProver:
    """Local prover on this rank's GPU (IOPProverState step API over the C ABI)."""

    def __init__(self, dev, mles, terms, num_vars, degree):
        from .api import IOPProverState
        self.st = IOPProverState(dev, mles, terms, num_vars, degree)

    def round_eval(self):
        return self.st.round_eval()

    def bind(self, r):
        self.st.bind(r)

    def final_evals(self):
        return self.st.get_mle_flatten_final_evaluations()

    def close(self):
        self.st.close()


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config #5 on N GPUs: Basefold commit of a row-sharded witness batch.
# RS-encoding is per column and global over the rows, the leaf hash is per row and global over the columns, so the
# commit has two REAL exchange steps (NCCL all-to-all over NVLink — this is where a collective belongs) and one tiny one:
#   rows -> columns : every rank receives whole columns (a contiguous block of width/N of them) for the transform;
#   columns -> rows : every rank receives its contiguous range of codeword rows (bit-reversed order, the leaf order) of
#                     ALL columns — a column-major local matrix, exactly what cg_merkle_commit hashes;
#   roots           : all-gather of the N subtree roots (32 bytes each); every rank finishes the top log2 N levels.
# A contiguous aligned range of leaves is a subtree of the global tree, so the root equals the single-GPU commitment's.
def sharded_commit_layout(width, log_n, rate_log, world):
    """Shapes of the two exchanges (pure arithmetic; covered by the CPU tests)."""
    assert width % world == 0 and (1 << log_n) % world == 0
    rows_local, cols_local, code_rows = (1 << log_n) // world, width // world, 1 << (log_n + rate_log)
    return {"rows_local": rows_local, "cols_local": cols_local, "code_rows": code_rows, "code_rows_local": code_rows // world}


def rows_to_columns(t, world, torch, dist=None):
    """t: [width, rows_local] (this rank's rows of every column, column-major).  Returns [cols_local, rows] — this rank's
    whole columns [rank * cols_local, (rank + 1) * cols_local)."""
    width, rows_local = t.shape
    cols_local = width // world
    recv = torch.empty_like(t)
    if dist is None or world == 1:
        recv.copy_(t)
    else:
        dist.all_to_all_single(recv.view(-1), t.contiguous().view(-1))          # chunk d = columns block d -> rank d
    return recv.view(world, cols_local, rows_local).permute(1, 0, 2).contiguous().view(cols_local, world * rows_local)


def columns_to_rows(code, world, torch, dist=None):
    """code: [cols_local, code_rows] (whole encoded columns).  Returns [width, code_rows_local]: every column's rows
    [rank * code_rows_local, ...), column c = source_rank * cols_local + local index (the natural column order)."""
    cols_local, code_rows = code.shape
    send = code.view(cols_local, world, code_rows // world).permute(1, 0, 2).contiguous()
    recv = torch.empty_like(send)
    if dist is None or world == 1:
        recv.copy_(send)
    else:
        dist.all_to_all_single(recv.view(-1), send.view(-1))
    return recv.view(world * cols_local, code_rows // world)


def commit_sharded(dev, msg_local, log_n, rate_log, rank, world, torch, dist, stream=None):
    """msg_local: torch int64 CUDA tensor [width, 2^log_n / world] — this rank's contiguous row range of every witness
    column (canonical Goldilocks values).  Returns (root[4] np.uint64 — identical on every rank and equal to the
    single-GPU commitment root, code_local tensor, tree DeviceBuffer of the local subtree)."""
    import ctypes as C
    from . import api
    width = msg_local.shape[0]
    lay = sharded_commit_layout(width, log_n, rate_log, world)
    cols = rows_to_columns(msg_local, world, torch, dist)                                      # [cols_local, 2^log_n]
    code = torch.empty((lay["cols_local"], lay["code_rows"]), dtype=torch.int64, device=msg_local.device)
    torch.cuda.current_stream().synchronize()
    dev.check(dev.lib.cg_rs_encode(dev.ctx, C.c_void_p(cols.data_ptr()), lay["cols_local"], log_n, rate_log, C.c_void_p(code.data_ptr()),
                                   api.NTT_BITREV, C.c_void_p(stream) if stream else None))
    dev.check(dev.lib.cg_stream_sync(dev.ctx, C.c_void_p(stream) if stream else None))
    code_local = columns_to_rows(code, world, torch, dist)                                     # [width, code_rows_local], column-major
    torch.cuda.current_stream().synchronize()
    h_loc = lay["code_rows_local"]
    tree = dev.alloc(32 * (2 * h_loc - 1))
    sub = np.zeros(4, np.uint64)
    dev.check(dev.lib.cg_merkle_commit(dev.ctx, C.c_void_p(code_local.data_ptr()), width, h_loc, 1, C.c_void_p(tree.ptr),
                                       sub.ctypes.data_as(C.c_void_p), C.c_void_p(stream) if stream else None))
    # top of the tree: all-gather the subtree roots, compress pairwise on the device
    roots = torch.from_numpy(sub.view(np.int64).copy()).to(msg_local.device)
    if world > 1:
        allr = [torch.empty_like(roots) for _ in range(world)]
        dist.all_gather(allr, roots)
        level = torch.stack(allr)                                                              # [world, 4]
    else:
        level = roots.view(1, 4)
    while level.shape[0] > 1:
        states = level.view(-1, 8).contiguous()                                                # left || right
        torch.cuda.current_stream().synchronize()
        dev.check(dev.lib.cg_poseidon2_permute(dev.ctx, C.c_void_p(states.data_ptr()), states.shape[0], C.c_void_p(stream) if stream else None))
        dev.check(dev.lib.cg_stream_sync(dev.ctx, C.c_void_p(stream) if stream else None))
        level = states[:, :4].contiguous()
    root = level.view(-1).cpu().numpy().view(np.uint64).copy()
    return root, code_local, tree
