"""Chip-level proof flow over the device path — the shape of ZKVMProver::create_chip_proof
(reference ceno_zkvm/src/scheme/prover.rs:700-1000, cpu/mod.rs:586-700 build_tower_witness, :346-554 tower prover,
gkr_iop/src/gkr/layer/cpu/mod.rs:99-239 main zerocheck) for SYNTHETIC chips:

    committed witness columns (column-major device matrix, zero-copy MLE views)
      -> record MLEs by wit_infer_by_monomial_expr               (read / write / lookup expressions)
      -> interleaving_mles_to_mles -> tower build -> tower proof (2 product specs + 1 logup spec)
      -> main zerocheck: eq(rt_tower) prefix selector x alpha-combined record expressions, degree <= 3
      -> Basefold-style commitment of the witness matrix (RS-encode + Merkle), when Poseidon2 constants are set

Everything runs on the stream it is given, so chips can be proved concurrently on ChipScheduler lanes.  The circuits are
synthetic (random monomial expressions): the real ones come from ceno's circuit builder, which is out of scope."""
import random

import numpy as np

from . import api, synth
from .api import DeviceBuffer, MultilinearExtension

P = 0xFFFFFFFF00000001


class SyntheticChip:
    """num_vars rows = 2^num_vars, n_wit base witness columns, records given as monomial expressions over the columns."""

    def __init__(self, seed, num_vars, num_instances=None, n_wit=16, n_read=3, n_write=3, n_lk=6, name=None):
        self.seed, self.num_vars, self.n_wit = seed, num_vars, n_wit
        self.num_instances = (1 << num_vars) if num_instances is None else num_instances
        self.name = name or f"chip{seed}"
        rng = random.Random(seed)

        def expr():   # c0 + c1 w_a + c2 w_b w_c  (a record: rlc of witness columns)
            return [([rng.randrange(P), rng.randrange(P)], []), ([rng.randrange(P), rng.randrange(P)], [rng.randrange(n_wit)]),
                    ([rng.randrange(P), rng.randrange(P)], [rng.randrange(n_wit), rng.randrange(n_wit)])]

        self.read_exprs = [expr() for _ in range(n_read)]
        self.write_exprs = [expr() for _ in range(n_write)]
        self.lk_exprs = [expr() for _ in range(n_lk)]

    def witness(self):
        """column-major host matrix: column c at [c * 2^k, (c+1) * 2^k); rows past num_instances are zero padding."""
        n = 1 << self.num_vars
        m = synth.fill_base(0xC41B ^ self.seed, self.n_wit * n).reshape(self.n_wit, n)
        m[:, self.num_instances:] = 0
        return m.reshape(-1)

    def estimated_memory_bytes(self):
        n = 1 << self.num_vars
        recs = len(self.read_exprs) + len(self.write_exprs) + len(self.lk_exprs)
        return 8 * n * self.n_wit + 16 * n * recs * 3 + 16 * n * 4


def upload_witness(dev, chip):
    """-> (matrix DeviceBuffer, [column MLE views])  (owned_subrange views, ceno_zkvm/src/scheme/gpu/mod.rs:2088-2113)"""
    n = 1 << chip.num_vars
    buf = dev.to_device(chip.witness())
    cols = [MultilinearExtension(dev, DeviceBuffer(dev, buf.ptr + 8 * n * c, 8 * n, owner=False), chip.num_vars, False) for c in range(chip.n_wit)]
    return buf, cols


def create_chip_proof(dev, chip, cols, transcript, stream=None, alpha=(12345, 678), commit_matrix=None):
    """One chip: records -> towers -> tower proof -> main zerocheck (-> commitment).  Returns a dict of host arrays."""
    k, ninst = chip.num_vars, chip.num_instances
    sync = (lambda: dev.check(dev.lib.cg_stream_sync(dev.ctx, api.C.c_void_p(stream)))) if stream else dev.sync
    out, temps = {}, []
    import os
    import time
    trace = [] if os.environ.get("CHIP_TRACE") else None
    t_last = [time.perf_counter()]

    def mark(name):
        if trace is not None:
            sync()
            now = time.perf_counter()
            trace.append((name, round((now - t_last[0]) * 1e3, 3)))
            t_last[0] = now

    if commit_matrix is not None:
        code, tree, root = api.basefold_style_commit(dev, commit_matrix, chip.n_wit, k, 1, stream=stream)
        out["commit_root"] = root
        temps += [code, tree]
    mark("commit")
    # records
    groups = []
    for exprs in (chip.read_exprs, chip.write_exprs, chip.lk_exprs):
        groups.append([api.wit_infer_by_monomial_expr(dev, cols, e, k, stream=stream) for e in exprs])
    specs, keep = [], []
    for recs in groups[:2]:
        leaves, buf = api.interleaving_mles_to_mles(dev, recs, ninst, 2, [1, 0], stream=stream)
        keep.append(buf)
        specs.append(api.TowerProverSpec(leaves, leaves[0].num_vars + 1, False))
    leaves, buf = api.interleaving_mles_to_mles(dev, groups[2], ninst, 2, list(alpha), stream=stream)
    keep.append(buf)
    specs.append(api.TowerProverSpec([None, None, leaves[0], leaves[1]], leaves[0].num_vars, True))
    mark("records+interleave")
    tw = api.TowerProver(dev, specs, stream=stream)
    mark("tower_build")
    out["tower_proof"], out["tower_point"] = tw.create_proof(transcript)
    tw.close()
    mark("tower_prove")
    # main zerocheck at the tower's point: sel(rt[:k]) * sum_i alpha^i record_i(w)
    rt = out["tower_point"][:2 * k]
    if rt.size < 2 * k:                                                    # tiny chips: tower point shorter than k
        rt = np.concatenate([rt, transcript.sample_and_append_vec(b"pad", k - rt.size // 2)])
    sel = api.build_eq_x_r_vec(dev, rt, 0, ninst, stream=stream)
    apows = transcript.sample_and_append_challenge_pows(sum(len(g) for g in groups), b"combine subset evals")
    from .expr import Poly, ext
    poly, ai = Poly(), 0
    for exprs in (chip.read_exprs, chip.write_exprs, chip.lk_exprs):
        for e in exprs:
            rec = Poly()
            for c, ids in e:
                t = Poly.const(ext(c[0], c[1]))
                for i in ids:
                    t = t * Poly.var(1 + i)
                rec = rec + t
            poly = poly + rec * ext(int(apows[ai][0]), int(apows[ai][1]))
            ai += 1
    terms = (poly * Poly.var(0)).terms()
    rounds, evals, point = api.IOPProverState.prove(dev, [sel] + cols, terms, k, 3, transcript=transcript, stream=stream)
    out["main_proof"], out["main_evals"], out["main_point"] = rounds, evals, point
    sync()
    mark("main_sumcheck")
    if trace is not None:
        out["_trace"] = trace
    for g in groups:
        for m in g:
            m.free()
    for b in keep + temps:
        b.free()
    sel.free()
    return out
