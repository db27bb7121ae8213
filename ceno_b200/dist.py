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
