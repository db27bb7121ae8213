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


def bench_sharded(args, rank, world, local_rank):
    """bench.py's N>1 arm: T3-k strong scaling — each rank owns a 1/N slice of the same 2^k instance."""
    import json
    import torch
    import torch.distributed as dist
    from . import api as cb
    from . import synth

    k, deg = args.k, 3
    g = world.bit_length() - 1
    assert 1 << g == world, "number of GPUs must be a power of two"
    k_local = k - g
    n_local = 1 << k_local
    dev = cb.Device(local_rank)
    seed_a, seed_b, seed_w = 0xC0FFEE ^ 1, 0xC0FFEE ^ 2, 0xE9
    w = synth.fill_ext(seed_w, k)
    # this rank's slices, pinned on the host (the e2e source) and resident on the device
    nbytes = 16 * n_local
    a_h, a_hp = dev.pinned(nbytes)
    b_h, b_hp = dev.pinned(nbytes)
    synth.fill_ext(seed_a, n_local, start=rank * n_local, out=a_h)
    synth.fill_ext(seed_b, n_local, start=rank * n_local, out=b_h)
    a_d, b_d = dev.alloc(nbytes), dev.alloc(nbytes)
    dev.h2d(a_d.ptr, a_hp, nbytes)
    dev.h2d(b_d.ptr, b_hp, nbytes)
    dev.sync()
    A = cb.MultilinearExtension(dev, a_d, k_local, True)
    B = cb.MultilinearExtension(dev, b_d, k_local, True)
    if getattr(args, "eq", "virtual") == "table":
        eq_lo = cb.build_eq_x_r_vec(dev, w[:2 * k_local])
        s = eq_slice_scalar(w[2 * k_local:], rank)
        EQ = cb.wit_infer_by_monomial_expr(dev, [eq_lo], [(list(s), [0])], k_local)      # eq slice = scalar * eq(w_low, .)
        eq_lo.free()
    else:   # eq handed over as its (global) point: split-eq rounds, no eq stream; the rank factor is derived by the library
        EQ = cb.EqPolynomial(dev, w, num_vars=k_local)
    terms = [([1, 0], [0, 1, 2])]

    def xchg(blob):
        outs = [None] * world
        dist.all_gather_object(outs, blob)
        return outs
    comm = cb.Comm(dev, rank, world, xchg, barrier=dist.barrier)
    stream = torch.cuda.Stream()
    sh = stream.cuda_stream

    def step(device_challenger=False):
        return cb.prove_sharded(dev, comm, [EQ, A, B], terms, k, deg, cb.StandInTranscript(b"bench"),
                                device_challenger=device_challenger, stream=sh)

    def timed(fn, steps):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            o = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)     # max over ranks
        dist.barrier()
        return float(t.item()), o

    def e2e_step():   # every rank uploads its own slices over its own PCIe link, then the sharded prove
        dev.h2d(a_d.ptr, a_hp, nbytes, sh)
        dev.h2d(b_d.ptr, b_hp, nbytes, sh)
        return step()

    clocks = None
    if rank == 0:
        try:
            import bench as _bench
            clocks = _bench.ClockSampler(local_rank)
            clocks.start()
        except Exception:  # noqa: BLE001
            clocks = None
    for _ in range(args.warmup):
        out = step()
        step(True)
    l0 = dev.launch_count()
    ms, out = timed(step, args.steps)
    launches = dev.launch_count() - l0
    ms_dev, out_dev = timed(lambda: step(True), args.steps)
    assert all(np.array_equal(x, y) for x, y in zip(out, out_dev))
    e2e_step()
    ms_e2e, out_e2e = timed(e2e_step, max(1, min(args.steps, 5)))
    assert all(np.array_equal(x, y) for x, y in zip(out, out_e2e))
    clk = clocks.stop() if clocks is not None else None
    if rank == 0:
        n = 1 << k
        ops = 99 * n
        line = {
            "metric": "sumcheck Gfield-ops/s", "value": ops / (ms * 1e-3) / 1e9, "unit": "Gfield-ops/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64 (Goldilocks, ext2)", "data": "synthetic",
            "config": {"workload": f"T3-{k}: eq(w,x)*A(x)*B(x), 2^{k}-point hypercube, degree 3, GoldilocksExt2, sliced 1/{world} per GPU",
                       "k": k, "degree": deg, "n_mles": 3, "parallelism": f"hypercube slices x{world}", "eq": getattr(args, "eq", "virtual"),
                       "exchange": "in-kernel: last block stores its 3 ext partials into every peer's NVLink-mapped mailbox, waits for the N flags, sums mod p; replicated host transcript; all-gather of the final local elements + replicated tail",
                       "l2": f"per-GPU inputs {3 * 16 * n_local >> 20} MiB"},
            "points_per_s": n / (ms * 1e-3), "rounds_per_s": k / (ms * 1e-3),
            "device_challenger": {"ms_per_step": ms_dev, "value": ops / (ms_dev * 1e-3) / 1e9, "unit": "Gfield-ops/s"},
            "e2e": {"value": ops / (ms_e2e * 1e-3) / 1e9, "unit": "Gfield-ops/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": world * 2 * nbytes + 16 * k, "d2h_bytes_per_step": 16 * (k * deg + 3 + k),
                    "note": f"every rank uploads its 1/{world} slices of A and B from pinned host memory over its own PCIe link, then the sharded prove; max over ranks"},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        print(json.dumps(line))
    comm.close()
    dev.close()
    dist.destroy_process_group()
