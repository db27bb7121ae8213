"""world_size-2 (and 4) gloo tests of the sharded-sumcheck protocol (ceno_b200/dist.py) on CPU:
hypercube slices, per-round all_gather + modular sum of the partial round messages, replicated
transcript, gathered tail rounds.  The local prover here is backed by the ORACLE (test
infrastructure) so the host-side logic is exercised without a GPU; the result must equal the
monolithic oracle proof bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleLocalProver:
    def __init__(self, arrays, nv, terms):
        from oracle import oracle as orc
        self.orc, self.cur, self.nv, self.terms = orc, [np.array(a) for a in arrays], nv, terms

    def round_eval(self):
        nv = self.nv
        rounds, _, _ = self.orc.sumcheck_prove([(c, True, nv) for c in self.cur], self.terms, nv, 3,
                                               challenge_fn=lambda *_: np.array([0, 0], dtype=np.uint64))
        return rounds[0].reshape(-1).copy()

    def bind(self, r):
        self.cur = [self.orc.fix_variable(c, True, r) for c in self.cur]
        self.nv -= 1

    def final_evals(self):
        return np.array(self.cur, dtype=np.uint64).reshape(-1, 2)


def _worker(rank, world, k, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ceno_b200 import dist as cdist
    from ceno_b200.api import StandInTranscript
    from oracle import oracle as orc
    g = world.bit_length() - 1
    kl = k - g
    nl = 1 << kl
    w = orc.fill_ext(0xE9, k)
    eq = orc.build_eq_x_r_vec(w)
    a, b = orc.fill_ext(1, 1 << k), orc.fill_ext(2, 1 << k)
    sl = slice(2 * rank * nl, 2 * (rank + 1) * nl)
    # the eq slice equals scalar(q) * eq(w_low): check the helper the GPU path uses
    s = cdist.eq_slice_scalar(w[2 * kl:], rank)
    lo = orc.build_eq_x_r_vec(w[:2 * kl])
    for i in (0, 1, nl - 1):
        e = (int(lo[2 * i]), int(lo[2 * i + 1]))
        assert cdist.ext_mul_host(s, e) == (int(eq[sl][2 * i]), int(eq[sl][2 * i + 1]))
    terms = [([1, 0], [0, 1, 2])]
    local = OracleLocalProver([eq[sl], a[sl], b[sl]], kl, terms)
    out = cdist.sharded_prove(local, kl, g, 3, StandInTranscript(b"dist"), cdist.TorchExchange(None),
                              lambda arrays: OracleLocalProver(arrays, g, terms))
    want = orc.sumcheck_prove([(eq, True, k), (a, True, k), (b, True, k)], terms, k, 3, transcript=orc.Transcript(b"dist"))
    ok = all(np.array_equal(x, y) for x, y in zip(out, want))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,k", [(2, 6), (4, 7)])
def test_sharded_protocol_matches_monolithic_oracle(world, k):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, k, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def test_sharded_commit_exchange_layout_world_1_2_4():
    """BASELINE config #5 sharding (rows -> columns -> rows): composing the two exchanges of every rank reproduces the
    single-device column-major matrices (checked with the exchanges emulated in-process, no GPU, no process group)."""
    import torch
    from ceno_b200 import dist as cdist
    width, log_n, rate_log = 8, 5, 1
    full = torch.arange(width * (1 << log_n), dtype=torch.int64).view(width, 1 << log_n)
    for world in (1, 2, 4):
        lay = cdist.sharded_commit_layout(width, log_n, rate_log, world)
        locs = [full[:, r * lay["rows_local"]:(r + 1) * lay["rows_local"]].contiguous() for r in range(world)]
        # emulate all_to_all_single: chunk d of rank s goes to slot s of rank d
        sends = [t.view(world, -1) for t in locs]
        recvs = [torch.stack([sends[s][d] for s in range(world)]).view(width, lay["rows_local"]) for d in range(world)]
        cols = [r.view(world, lay["cols_local"], lay["rows_local"]).permute(1, 0, 2).contiguous().view(lay["cols_local"], -1) for r in recvs]
        for r in range(world):
            assert torch.equal(cols[r], full[r * lay["cols_local"]:(r + 1) * lay["cols_local"]])
        code = [torch.cat([c * 1000, c * 1000 + 1], dim=1) for c in cols]                     # stand-in "codeword" of the right shape
        sends2 = [c.view(lay["cols_local"], world, lay["code_rows_local"]).permute(1, 0, 2).contiguous() for c in code]
        for d in range(world):
            got = torch.stack([sends2[s][d] for s in range(world)]).view(width, lay["code_rows_local"])
            want = torch.cat(code, dim=0)[:, d * lay["code_rows_local"]:(d + 1) * lay["code_rows_local"]]
            assert torch.equal(got, want)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_tower_shuffle_rule(world):
    """The placement rule of the sharded tower build (cg_tower_build_sharded, ceno_b200/csrc/cabi.cu: tower_build_spec /
    TowerDst), simulated with numpy: rank r holds slice r of BOTH fan-in halves of a layer; its products are one contiguous
    piece of ONE half of the next layer up, whose two sub-slices belong to ranks 2r mod N and 2r+1 mod N.  After every level
    each rank must again hold slice r of both halves (so layer sumchecks and the next level stay local), and the global layer
    must equal the single-device layer (ceno_zkvm/src/scheme/utils.rs:606-652: low half | high half of the product array)."""
    P = 0xFFFFFFFF00000001
    rng = np.random.default_rng(5 + world)
    top = 4 + world.bit_length()                      # arrays of the leaf layer: 2^top elements each (globally)
    a = [int(x) for x in rng.integers(0, P, 1 << top, dtype=np.uint64)]
    b = [int(x) for x in rng.integers(0, P, 1 << top, dtype=np.uint64)]
    local = []                                        # local[r] = (slice r of a, slice r of b)
    m = (1 << top) // world
    for r in range(world):
        local.append((a[r * m:(r + 1) * m], b[r * m:(r + 1) * m]))
    ga, gb = a, b
    for level in range(top - 1, world.bit_length() - 2, -1):   # while every rank still keeps at least one element per array
        res = [ga[x] * gb[x] % P for x in range(len(ga))]  # single device: layer `level` = (low half | high half) of the products
        half = len(res) // 2
        ga, gb = res[:half], res[half:]
        m2 = len(local[0][0]) // 2                         # slice length of the new layer
        new = [([None] * m2, [None] * m2) for _ in range(world)]
        for r in range(world):
            mine = [x * y % P for x, y in zip(*local[r])]  # a contiguous piece of ONE half of the new layer
            part = 0 if r < world // 2 else 1
            for sub, dst in ((0, (2 * r) % world), (1, (2 * r + 1) % world)):
                new[dst][part][:] = mine[sub * m2:(sub + 1) * m2]
        local = new
        for r in range(world):
            assert local[r][0] == ga[r * m2:(r + 1) * m2] and local[r][1] == gb[r * m2:(r + 1) * m2], (level, r)
        if m2 == 1:
            break
