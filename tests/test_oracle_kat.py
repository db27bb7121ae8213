"""Pins the CPU oracle (oracle/ceno_oracle.c) against every known-answer relation the
reference tree holds for this path and against the independent
big-int restatement oracle/pyref.py.  CPU-only."""
import itertools
import random

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import pyref as pr

P = pr.P


def rnd_ext(rng):
    return (rng.randrange(P), rng.randrange(P))


def ext_arr(pairs):
    return pr.from_pairs(pairs)


def E(x):
    return (x % P, 0)


# ---------------------------------------------------------------- field
def test_gl_mul_fast_equals_slow_and_edge_cases():
    rng = random.Random(1)
    L = orc.lib()
    edge = [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0xFFFFFFFF00000000, 1 << 32, (1 << 32) - 1, (1 << 63) % P]
    for a, b in itertools.product(edge, edge):
        assert L.or_gl_mul(a, b) == a * b % P == L.or_gl_mul_slow(a, b)
        assert L.or_gl_add(a, b) == (a + b) % P
        assert L.or_gl_sub(a, b) == (a - b) % P
    for _ in range(20000):
        a, b = rng.randrange(P), rng.randrange(P)
        assert L.or_gl_mul(a, b) == a * b % P
    for a in edge[1:]:
        assert L.or_gl_mul(a, L.or_gl_inv(a)) == 1


def test_ext_mul_w7_and_inverse():
    # X^2 = 7 (p3-goldilocks 0.4.3 BinomiallyExtendable<2>)
    assert tuple(int(x) for x in orc.ext_mul([0, 1], [0, 1])) == (7, 0)
    rng = random.Random(2)
    for _ in range(2000):
        a, b = rnd_ext(rng), rnd_ext(rng)
        assert tuple(int(x) for x in orc.ext_mul(list(a), list(b))) == pr.emul(a, b)
    a = rnd_ext(rng)
    assert pr.emul(a, tuple(int(x) for x in orc.ext_inv(list(a)))) == pr.ONE


# ------------------------------------------------- eq table and evaluate
@pytest.mark.parametrize("k", [0, 1, 2, 5, 9])
def test_build_eq_matches_definition(k):
    rng = random.Random(10 + k)
    r = [rnd_ext(rng) for _ in range(k)]
    got = pr.to_pairs(orc.build_eq_x_r_vec(ext_arr(r)))
    assert got == pr.build_eq_x_r_vec(r)
    s = pr.ZERO
    for e in got:
        s = pr.eadd(s, e)
    assert s == pr.ONE  # sum_b eq(r,b) = 1


def test_mle_evaluate_is_lsb_first():
    # point[i] pairs with index bit i (gkr_iop/src/utils.rs:209-232): evaluating at a
    # boolean point returns evals[sum b_i 2^i].
    rng = random.Random(3)
    k = 4
    evals = [rnd_ext(rng) for _ in range(1 << k)]
    for b in range(1 << k):
        pt = [E((b >> i) & 1) for i in range(k)]
        assert tuple(int(x) for x in orc.mle_evaluate(ext_arr(evals), True, ext_arr(pt))) == evals[b]
    pt = [rnd_ext(rng) for _ in range(k)]
    assert tuple(int(x) for x in orc.mle_evaluate(ext_arr(evals), True, ext_arr(pt))) == pr.mle_evaluate(evals, pt)
    base = np.array([rng.randrange(P) for _ in range(1 << k)], dtype=np.uint64)
    assert tuple(int(x) for x in orc.mle_evaluate(base, False, ext_arr(pt))) == pr.mle_evaluate([E(int(x)) for x in base], pt)


R5 = [E(123), E(456), E(789), E(3210), E(9876)]  # gkr_iop/src/utils.rs:357-363


def test_kat_eval_stacked_wellform_address_vec():
    # gkr_iop/src/utils.rs:355-374
    for n in range(len(R5)):
        v = [pr.ZERO] + [E(x) for i in range(n + 1) for x in range(1 << i)]
        r = R5[: n + 1]
        want = tuple(int(x) for x in orc.mle_evaluate(ext_arr(v), True, ext_arr(r)))
        assert tuple(int(x) for x in orc.eval_stacked_wellform_address_vec(ext_arr(r))) == want == pr.mle_evaluate(v, r)


def test_kat_eval_stacked_constant_vec():
    # gkr_iop/src/utils.rs:376-395
    for n in range(len(R5)):
        v = [pr.ZERO] + [E(i) for i in range(n + 1) for _ in range(1 << i)]
        r = R5[: n + 1]
        want = tuple(int(x) for x in orc.mle_evaluate(ext_arr(v), True, ext_arr(r)))
        assert tuple(int(x) for x in orc.eval_stacked_constant_vec(ext_arr(r))) == want


def test_kat_inner_outer_repeated_incremental_vec():
    # gkr_iop/src/utils.rs:397-441
    for n in range(1, len(R5) + 1):
        for k in range(n + 1):
            r = R5[:n]
            inner = [E(i) for i in range(1 << (n - k)) for _ in range(1 << k)]
            got = tuple(int(x) for x in orc.eval_wellform_address_vec(0, 1, ext_arr(r[k:]) if k < n else np.zeros(0, np.uint64)))
            assert got == tuple(int(x) for x in orc.mle_evaluate(ext_arr(inner), True, ext_arr(r)))
            outer = [E(x) for _ in range(1 << (n - k)) for x in range(1 << k)]
            got = tuple(int(x) for x in orc.eval_wellform_address_vec(0, 1, ext_arr(r[:k]) if k else np.zeros(0, np.uint64)))
            assert got == tuple(int(x) for x in orc.mle_evaluate(ext_arr(outer), True, ext_arr(r)))


def test_eq_eval_and_less_or_equal_than():
    rng = random.Random(4)
    n = 5
    a = [rnd_ext(rng) for _ in range(n)]
    b = [rnd_ext(rng) for _ in range(n)]
    eqa, eqb = pr.build_eq_x_r_vec(a), pr.build_eq_x_r_vec(b)
    full = pr.ZERO
    for x, y in zip(eqa, eqb):
        full = pr.eadd(full, pr.emul(x, y))
    assert tuple(int(x) for x in orc.eq_eval(ext_arr(a), ext_arr(b))) == full
    for max_idx in [0, 1, 5, 0b10101, 30, 31]:
        want = pr.ZERO
        for i in range(max_idx + 1):
            want = pr.eadd(want, pr.emul(eqa[i], eqb[i]))
        assert tuple(int(x) for x in orc.eq_eval_less_or_equal_than(max_idx, ext_arr(a), ext_arr(b))) == want


# ------------------------------------------------------------ selectors
def test_kat_quark_lt_selector():
    # gkr_iop/src/selector.rs:396-435: n_points = 5 -> n_vars = 3
    rng = random.Random(5)
    out_rt = [rnd_ext(rng) for _ in range(3)]
    eq = pr.build_eq_x_r_vec(out_rt)
    sel = pr.to_pairs(orc.selector_compute(orc.SEL_QUARK_LT, ext_arr(out_rt), num_instances=5))
    assert sel == [eq[0], eq[1], pr.ZERO, pr.ZERO, eq[4], pr.ZERO, eq[6], pr.ZERO]


def test_selector_prefix_and_sparse_evaluate_relation():
    # compute() vs evaluate() closed forms (gkr_iop/src/selector.rs:262-303)
    rng = random.Random(6)
    nv = 5
    out_pt = [rnd_ext(rng) for _ in range(nv)]
    in_pt = [rnd_ext(rng) for _ in range(nv)]
    for offset, ninst in [(0, 32), (0, 7), (3, 11), (31, 1), (0, 0)]:
        sel = orc.selector_compute(orc.SEL_PREFIX, ext_arr(out_pt), offset=offset, num_instances=ninst)
        got = tuple(int(x) for x in orc.mle_evaluate(sel, True, ext_arr(in_pt)))
        end = offset + ninst
        if end == 0:
            want = pr.ZERO
        else:
            want = tuple(int(x) for x in orc.eq_eval_less_or_equal_than(end - 1, ext_arr(out_pt), ext_arr(in_pt)))
            if offset > 0:
                want = pr.esub(want, tuple(int(x) for x in orc.eq_eval_less_or_equal_than(offset - 1, ext_arr(out_pt), ext_arr(in_pt))))
        assert got == want
    # OrderedSparse: inner 3 vars, indices {1,4,6}, 3 instances of 4 chunks
    indices, inner = [1, 4, 6], 3
    ninst = 3
    sel = orc.selector_compute(orc.SEL_ORDERED_SPARSE, ext_arr(out_pt), num_instances=ninst, indices=indices, inner_vars=inner)
    got = tuple(int(x) for x in orc.mle_evaluate(sel, True, ext_arr(in_pt)))
    oe, ie = pr.build_eq_x_r_vec(out_pt[:inner]), pr.build_eq_x_r_vec(in_pt[:inner])
    ev = pr.ZERO
    for i in indices:
        ev = pr.eadd(ev, pr.emul(oe[i], ie[i]))
    s = tuple(int(x) for x in orc.eq_eval_less_or_equal_than(ninst - 1, ext_arr(out_pt[inner:]), ext_arr(in_pt[inner:])))
    assert got == pr.emul(ev, s)


# ------------------------------------------------------- tower witnesses
def test_kat_interleaving_mles_to_mles():
    # ceno_zkvm/src/scheme/utils.rs:968-1065 literal vectors
    def mle(*xs):
        return (ext_arr([E(x) for x in xs]), True)

    res = orc.interleaving_mles_to_mles([mle(1, 2), mle(3, 4), mle(5, 6), mle(7, 8)], 2, 2, [1, 0])
    assert pr.to_pairs(res[0]) == [E(1), E(3), E(5), E(7)]
    assert pr.to_pairs(res[1]) == [E(2), E(4), E(6), E(8)]
    res = orc.interleaving_mles_to_mles([mle(1, 2), mle(3, 4), mle(5, 6)], 2, 2, [0, 0])
    assert pr.to_pairs(res[0]) == [E(1), E(3), E(5), E(0)]
    assert pr.to_pairs(res[1]) == [E(2), E(4), E(6), E(0)]
    res = orc.interleaving_mles_to_mles([mle(1, 0), mle(3, 0), mle(5, 0)], 1, 2, [1, 0])
    assert pr.to_pairs(res[0]) == [E(1), E(3), E(5), E(1)]
    assert pr.to_pairs(res[1]) == [E(1)] * 4
    res = orc.interleaving_mles_to_mles([mle(2), mle(3)], 1, 2, [1, 0])
    assert pr.to_pairs(res[0]) == [E(2), E(3)]
    assert pr.to_pairs(res[1]) == [E(1), E(1)]


def test_kat_infer_tower_product_witness():
    # ceno_zkvm/src/scheme/utils.rs:934-966
    _, layers = orc.infer_tower_product_witness(2, ext_arr([E(1), E(2)]), ext_arr([E(3), E(4)]))
    assert len(layers) == 2
    left, right = pr.to_pairs(layers[0][0]), pr.to_pairs(layers[0][1])
    assert len(left) == 1 and len(right) == 1
    assert pr.emul(left[0], right[0]) == E(1 * 2 * 3 * 4)


def test_kat_infer_tower_logup_witness():
    # ceno_zkvm/src/scheme/utils.rs:1067-1195 literal vectors
    q1 = ext_arr([E(x) for x in (1, 2, 3, 4)])
    q2 = ext_arr([E(x) for x in (5, 6, 7, 8)])
    _, layers = orc.infer_tower_logup_witness(2, None, None, q1, q2)
    assert len(layers) == 3
    L = [[pr.to_pairs(a) for a in lay] for lay in layers]
    assert L[2][0] == [E(1)] * 4 and L[2][1] == [E(1)] * 4
    assert L[2][2] == [E(1), E(2), E(3), E(4)] and L[2][3] == [E(5), E(6), E(7), E(8)]
    assert L[1][0] == [E(1 + 5), E(2 + 6)]
    assert L[1][1] == [E(3 + 7), E(4 + 8)]
    assert L[1][2] == [E(5), E(2 * 6)]
    assert L[1][3] == [E(3 * 7), E(4 * 8)]
    assert L[0][0] == [E((1 + 5) * (3 * 7) + (3 + 7) * 5)]
    assert L[0][1] == [E((2 + 6) * (4 * 8) + (4 + 8) * (2 * 6))]
    assert L[0][2] == [E((3 * 7) * 5)]
    assert L[0][3] == [E((4 * 8) * (2 * 6))]


# ------------------------------------------------------------- sumcheck
def _standin_cb(seed):
    def cb(j, msg):
        t = orc.Transcript(b"cb%d" % seed)
        t.append_message(int(j).to_bytes(8, "little"))
        t.append_ext(np.array(msg, dtype=np.uint64).reshape(-1))
        return t.sample(b"r")
    return cb


@pytest.mark.parametrize("k,degree", [(1, 3), (3, 3), (6, 3), (5, 2), (4, 5)])
def test_sumcheck_matches_pyref_and_verifier_relations(k, degree):
    rng = random.Random(100 * k + degree)
    m = degree
    mles_p = [[rnd_ext(rng) for _ in range(1 << k)] for _ in range(m)]
    # one full-degree product term + a lower-degree term (extrapolated up)
    terms = [(rnd_ext(rng), list(range(m))), (rnd_ext(rng), [0])]
    cb = _standin_cb(k)
    rounds, fin, chal = orc.sumcheck_prove([(ext_arr(x), True, k) for x in mles_p],
                                           [(list(c), ids) for c, ids in terms], k, degree, challenge_fn=cb)
    pmsgs, pfin, pchal = pr.sumcheck_prove(mles_p, terms, k, degree,
                                           lambda j, msg: tuple(int(x) for x in cb(j, [v for e in msg for v in e])))
    assert [[tuple(int(x) for x in e) for e in r] for r in rounds] == pmsgs
    assert [tuple(int(x) for x in e) for e in fin] == pfin
    assert [tuple(int(x) for x in e) for e in chal] == pchal
    # verifier (ceno_recursion_v2/src/main/mod.rs:3513-3526): eval_0 = claim - evals[0]; claim' = interp(r)
    claim = pr.ZERO
    for b in range(1 << k):
        claim = pr.eadd(claim, pr.poly_eval(mles_p, terms, b))
    for j in range(k):
        e0 = pr.esub(claim, pmsgs[j][0])
        claim = tuple(int(x) for x in orc.extrapolate_uni_poly(list(e0), rounds[j].reshape(-1), chal[j]))
        assert claim == pr.lagrange_eval([e0] + pmsgs[j], pchal[j])
    # final check: sum_t c_t prod f_i(r) == last claim, f_i(r) by direct MLE evaluation
    direct = [pr.mle_evaluate(m_, pchal) for m_ in mles_p]
    assert direct == pfin
    acc = pr.ZERO
    for c, ids in terms:
        p = c
        for i in ids:
            p = pr.emul(p, direct[i])
        acc = pr.eadd(acc, p)
    assert acc == claim


@pytest.mark.parametrize("k,sizes,degree", [(5, [5, 3, 3, 0], 3), (6, [2, 2, 6, 6, 4], 4), (4, [4, 1], 2)])
def test_mixed_size_frontload_sumcheck_verifies(k, sizes, degree):
    """The cross-chip batched main sumcheck (ceno_zkvm/src/scheme/cpu/mod.rs:1332-1360): MLEs with k' < k variables.
    Pinned by the verifier's final claim as restated in ceno_recursion_v2/src/main/mod.rs:3414-3448:
    sum_t scalar_t prod_i (f_i(r_0..r_{k'-1}) * prod_{j>=k'} r_j), with the raw f_i(r_<k') as reported evaluations,
    and by the claimed sum being the sum of every chip's own hypercube sum."""
    rng = random.Random(7 * k + degree)
    mles_p = [[rnd_ext(rng) for _ in range(1 << kv)] for kv in sizes]
    by_size = {}
    for i, kv in enumerate(sizes):
        by_size.setdefault(kv, []).append(i)
    terms = []
    for kv, ids in by_size.items():     # every term stays inside one size class (one chip)
        terms.append((rnd_ext(rng), ids[:degree]))
        terms.append((rnd_ext(rng), [ids[0]]))
        if len(ids) >= 2:
            terms.append((rnd_ext(rng), [ids[-1], ids[0]]))
    cb = _standin_cb(k)
    rounds, fin, chal = orc.sumcheck_prove([(ext_arr(x), True, kv) for x, kv in zip(mles_p, sizes)],
                                           [(list(c), ids) for c, ids in terms], k, degree, challenge_fn=cb)
    pchal = [tuple(int(x) for x in e) for e in chal]
    pfin = [tuple(int(x) for x in e) for e in fin]
    # claimed sum = sum over chips of their own hypercube sums (the padding region contributes nothing)
    claim = pr.ZERO
    for c, ids in terms:
        kv = sizes[ids[0]]
        for b in range(1 << kv):
            p_ = c
            for i in ids:
                p_ = pr.emul(p_, mles_p[i][b])
            claim = pr.eadd(claim, p_)
    for j in range(k):
        msg = [tuple(int(x) for x in e) for e in rounds[j]]
        e0 = pr.esub(claim, msg[0])
        claim = pr.lagrange_eval([e0] + msg, pchal[j])
    # reported evaluations are the raw small-MLE evaluations at the first k' challenges
    assert pfin == [pr.mle_evaluate(m_, pchal[:kv]) for m_, kv in zip(mles_p, sizes)]
    acc = pr.ZERO
    for c, ids in terms:
        value = c
        for i in ids:
            value = pr.emul(value, pfin[i])
            for tail in pchal[sizes[i]:]:
                value = pr.emul(value, tail)
        acc = pr.eadd(acc, value)
    assert acc == claim


def test_sumcheck_base_field_mles_and_transcript_order():
    rng = random.Random(77)
    k = 5
    a = np.array([rng.randrange(P) for _ in range(1 << k)], dtype=np.uint64)
    b = [rnd_ext(rng) for _ in range(1 << k)]
    terms = [([1, 0], [0, 1])]
    t1 = orc.Transcript(b"order")
    rounds, fin, chal = orc.sumcheck_prove([(a, False, k), (ext_arr(b), True, k)], terms, k, 2, transcript=t1)
    # replay the protocol order of SURVEY §A2 by hand
    t2 = orc.Transcript(b"order")
    t2.append_message((k).to_bytes(8, "little"))
    t2.append_message((2).to_bytes(8, "little"))
    for j in range(k):
        t2.append_ext(rounds[j].reshape(-1))
        assert tuple(t2.sample(b"Internal round")) == tuple(chal[j])
    assert t1.state == t2.state
    pt = [tuple(int(x) for x in c) for c in chal]
    assert tuple(int(x) for x in fin[0]) == pr.mle_evaluate([E(int(x)) for x in a], pt)


def test_tower_proof_verifies():
    """Prover->verifier round trip in the style of test_tower_proof_various_prod_size
    (ceno_zkvm/src/scheme/tests.rs:447-500): replay the transcript, check every layer's
    sumcheck and the claim flow of TowerVerify (ceno_zkvm/src/scheme/verifier.rs:1543-1700)."""
    rng = random.Random(9)
    nv_prod, nv_lk = 4, 3
    f1 = [rnd_ext(rng) for _ in range(1 << (nv_prod - 1))]
    f2 = [rnd_ext(rng) for _ in range(1 << (nv_prod - 1))]
    pw, players = orc.infer_tower_product_witness(nv_prod, ext_arr(f1), ext_arr(f2))
    q1 = [rnd_ext(rng) for _ in range(1 << nv_lk)]
    q2 = [rnd_ext(rng) for _ in range(1 << nv_lk)]
    lw, llayers = orc.infer_tower_logup_witness(nv_lk, None, None, ext_arr(q1), ext_arr(q2))
    tr = orc.Transcript(b"tower")
    proof, point = orc.tower_create_proof([(pw, nv_prod)], [(lw, nv_lk + 1)], tr)
    # ---- verifier
    tv = orc.Transcript(b"tower")
    alpha = tuple(int(x) for x in tv.sample(b"combine subset evals"))
    apow = [pr.ONE, alpha, pr.emul(alpha, alpha)]
    rt = [tuple(int(x) for x in tv.sample(b"product_sum"))]
    po = [pr.to_pairs(x)[0] for x in players[0]]
    lo = [pr.to_pairs(x)[0] for x in llayers[0]]
    # initial claims: evaluate the 1-var MLE of the two output values at rt
    def mle1(v0, v1, r):
        return pr.eadd(v0, pr.emul(r, pr.esub(v1, v0)))
    prod_claim = mle1(po[0], po[1], rt[0])
    p_claim = mle1(lo[0], lo[1], rt[0])
    q_claim = mle1(lo[2], lo[3], rt[0])
    pos = 0
    pp = pr.to_pairs(proof)
    max_round = 3
    for rnd in range(1, max_round + 1):
        nv = rnd
        claim = pr.eadd(pr.emul(apow[0], prod_claim), pr.eadd(pr.emul(apow[1], p_claim), pr.emul(apow[2], q_claim)))
        tv.append_message(nv.to_bytes(8, "little"))
        tv.append_message((3).to_bytes(8, "little"))
        chal = []
        for j in range(nv):
            msg = pp[pos:pos + 3]
            pos += 3
            e0 = pr.esub(claim, msg[0])
            tv.append_ext(pr.from_pairs(msg))
            r = tuple(int(x) for x in tv.sample(b"Internal round"))
            chal.append(r)
            claim = pr.lagrange_eval([e0] + msg, r)
        pe = pp[pos:pos + 2]
        pos += 2
        tv.append_ext(pr.from_pairs(pe))
        le = pp[pos:pos + 4]
        pos += 4
        tv.append_ext(pr.from_pairs(le))
        eqv = tuple(int(x) for x in orc.eq_eval(ext_arr(rt), ext_arr(chal)))
        inner = pr.eadd(pr.emul(apow[0], pr.emul(pe[0], pe[1])),
                        pr.eadd(pr.emul(apow[1], pr.eadd(pr.emul(le[0], le[3]), pr.emul(le[1], le[2]))),
                                pr.emul(apow[2], pr.emul(le[2], le[3]))))
        assert pr.emul(eqv, inner) == claim
        rm = tuple(int(x) for x in tv.sample(b"merge"))
        rt = chal + [rm]
        prod_claim = mle1(pe[0], pe[1], rm)
        p_claim = mle1(le[0], le[1], rm)
        q_claim = mle1(le[2], le[3], rm)
        alpha = tuple(int(x) for x in tv.sample(b"combine subset evals"))
        apow = [pr.ONE, alpha, pr.emul(alpha, alpha)]
    assert pos == len(pp)
    assert pr.to_pairs(point) == rt
    assert tv.state == tr.state
    # leaf claims equal direct evaluation of the input layer at rt (tests.rs:490-499)
    full = [x for x in f1] + [x for x in f2]  # top variable selects the half
    assert pr.mle_evaluate(full, rt) == prod_claim
    assert pr.mle_evaluate(q1 + q2, rt) == q_claim


@pytest.mark.parametrize("k,shape", [(0, "t3"), (1, "t3"), (2, "t3"), (5, "t3"), (9, "t3"), (12, "t3"), (7, "generic"), (11, "generic")])
def test_chunked_cpu_baseline_variant_is_bit_identical(k, shape):
    """The timed CPU arm (reference decomposition, in-place chunk folds) against the simple restatement."""
    rng = random.Random(k)
    n = 1 << k
    if shape == "t3":
        mles = [(orc.fill_ext(10 + i, n), True, k) for i in range(3)]
        terms = [([1, 0] if k % 2 else [rng.randrange(P), rng.randrange(P)], [0, 1, 2])]
        degree = 3
    else:
        mles = [(orc.fill_base(20, n), False, k), (orc.fill_ext(21, n), True, k), (orc.fill_ext(22, n), True, k), (orc.fill_ext(23, n), True, k)]
        terms = [([rng.randrange(P), rng.randrange(P)], [0, 1, 2, 3]), ([3, 0], [1]), ([5, 7], [])]
        degree = 4
    want = orc.sumcheck_prove(mles, terms, k, degree, transcript=orc.Transcript(b"chunk"))
    for consume in (False, True):
        copies = [(m[0].copy(), m[1], m[2]) for m in mles]
        got = orc.sumcheck_prove_chunked(copies, terms, k, degree, orc.Transcript(b"chunk"), consume=consume)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)


@pytest.mark.parametrize("variant", [0, 1])
def test_poseidon2_and_merkle_vs_bigint_restatement(variant):
    """a9 building blocks, constants supplied by the caller (parity unpinned: SURVEY §C-2/3)."""
    prm = orc.p2_params(seed=3, mds_variant=variant)
    ext_rc = [[int(prm.ext_rc[r][i]) for i in range(8)] for r in range(8)]
    int_rc = [int(prm.int_rc[r]) for r in range(22)]
    diag = [int(prm.diag[i]) for i in range(8)]
    perm = lambda st: pr.poseidon2_permute(ext_rc, int_rc, diag, variant, st)
    rng = random.Random(variant)
    for st in ([0] * 8, list(range(8)), [P - 1] * 8, [rng.randrange(P) for _ in range(8)]):
        assert [int(x) for x in orc.poseidon2_permute(prm, np.array(st, dtype=np.uint64))] == perm(st)
    for width, height in [(1, 1), (3, 2), (4, 4), (9, 8), (17, 16)]:
        m = orc.fill_base(77 + width, width * height)
        tree, root = orc.merkle_commit(prm, m, width, height)
        rows = [[int(x) for x in m[i * width:(i + 1) * width]] for i in range(height)]
        assert [int(x) for x in root] == pr.merkle_root(perm, rows)
        assert [int(x) for x in tree[:4]] == pr.hash_row(perm, rows[0])


@pytest.mark.parametrize("nv", [5, 6])
def test_kat_rotation_next_base_mle_eval(nv):
    """gkr_iop/src/utils.rs:343-365 (test_rotation_next_base_mle_eval) + get_rotation_points
    (gkr_iop/src/gkr/booleanhypercube.rs:124-163): rotated(point) = (1 - r_top) poly(left) + r_top poly(right)."""
    rng = random.Random(nv)
    total_vars = nv + 2
    poly = np.arange(1 << total_vars, dtype=np.uint64)
    rotated = orc.rotation_next_base_mle(poly, nv)
    tab = [int(x) for x in orc.bh_table(nv)]
    assert len(set(tab[:-1])) == (1 << nv) - 1 and tab[0] == tab[-1] == 1       # booleanhypercube.rs:196-242
    pt = [rnd_ext(rng) for _ in range(total_vars)]
    one_minus = lambda e: ((1 - e[0]) % P, (-e[1]) % P)
    if nv == 5:
        left = [pr.ZERO] + pt[:4] + pt[5:]
        right = [pr.ONE, pt[0], one_minus(pt[1])] + pt[2:4] + pt[5:]
    else:
        left = [pr.ZERO] + pt[:5] + pt[6:]
        right = [pr.ONE, one_minus(pt[0]), pt[1]] + pt[2:5] + pt[6:]
    left, right = left[:total_vars], right[:total_vars]
    ev = lambda arr, p: tuple(int(x) for x in orc.mle_evaluate(arr, False, ext_arr(p)))
    top = pt[nv - 1]
    want = pr.eadd(pr.emul(one_minus(top), ev(poly, left)), pr.emul(top, ev(poly, right)))
    assert ev(rotated, pt) == want
    # rotation_selector keeps exactly the first `subgroup` group elements of every chunk (utils.rs:54-76)
    eq = orc.build_eq_x_r_vec(ext_arr(pt))
    sel = pr.to_pairs(orc.rotation_selector(eq, 23, nv))
    eqp = pr.to_pairs(eq)
    keep = set(tab[:23])
    for b in range(1 << total_vars):
        assert sel[b] == (eqp[b] if (b & ((1 << nv) - 1)) in keep else pr.ZERO)


# ------------------------------------------------------------------ NTT / RS-encode (a9, f-2)
def test_two_adic_generator_is_the_p3_goldilocks_constant():
    """p3-goldilocks: GENERATOR = 7, TWO_ADICITY = 32, two_adic_generator(32) = 7^((p-1)/2^32) = 1753635133440165772."""
    g = orc.two_adic_generator(32)
    assert g == pow(7, (P - 1) >> 32, P) == 1753635133440165772
    assert pow(g, 1 << 31, P) == P - 1                     # primitive 2^32-th root of unity
    for bits in (0, 1, 5, 12, 27):
        assert orc.two_adic_generator(bits) == pow(g, 1 << (32 - bits), P)


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 6])
def test_ntt_matches_the_dft_definition(log_n):
    n = 1 << log_n
    x = [int(v) for v in orc.fill_base(77 + log_n, n)]
    w = orc.two_adic_generator(log_n)
    want = [sum(x[j] * pow(w, j * k, P) for j in range(n)) % P for k in range(n)]
    got = orc.ntt(np.array(x, dtype=np.uint64), log_n)
    assert [int(v) for v in got] == want
    rev = [int(format(i, "0%db" % log_n)[::-1], 2) if log_n else 0 for i in range(n)]
    br = orc.ntt(np.array(x, dtype=np.uint64), log_n, bitrev=True)
    assert all(int(br[rev[k]]) == want[k] for k in range(n))
    assert [int(v) for v in orc.ntt(got, log_n, inverse=True)] == x
    assert [int(v) for v in orc.ntt(br, log_n, inverse=True, bitrev=True)] == x


def test_rs_encode_is_polynomial_evaluation_and_linear():
    """codeword[k] = f(w^k) for f = the message as coefficients; systematic properties: linearity, and the
    rate-1/2 code of a constant message is constant."""
    log_n, rate_log, width = 4, 1, 3
    n, m = 1 << log_n, 1 << (log_n + rate_log)
    msg = orc.fill_base(91, width * n)
    code = orc.rs_encode(msg, width, log_n, rate_log, bitrev=False)
    w = orc.two_adic_generator(log_n + rate_log)
    for c in range(width):
        coeffs = [int(v) for v in msg[c * n:(c + 1) * n]]
        for k in (0, 1, 7, m - 1):
            x = pow(w, k, P)
            assert int(code[c * m + k]) == sum(a * pow(x, j, P) for j, a in enumerate(coeffs)) % P
    a, b = orc.fill_base(92, n), orc.fill_base(93, n)
    s = np.array([(int(u) + int(v)) % P for u, v in zip(a, b)], dtype=np.uint64)
    ca, cb_, cs = (orc.rs_encode(v, 1, log_n, rate_log) for v in (a, b, s))
    assert [int(v) for v in cs] == [(int(u) + int(v)) % P for u, v in zip(ca, cb_)]
    const = np.zeros(n, np.uint64); const[0] = 5
    assert set(int(v) for v in orc.rs_encode(const, 1, log_n, rate_log)) == {5}


# ------------------------------------------------------------------ EC-sum Quark (f-3)
def test_septic_extension_arithmetic():
    """F[z]/(z^7 - 2z - 5) (ceno_zkvm/src/scheme/septic_curve.rs:30-39): z^7 = 2z + 5, inverses, associativity."""
    z = [0, 1, 0, 0, 0, 0, 0]
    acc = [1, 0, 0, 0, 0, 0, 0]
    for _ in range(7):
        acc = orc.septic_mul(acc, z)
    assert acc == [5, 2, 0, 0, 0, 0, 0]
    a, b, c = ([int(v) for v in orc.fill_base(s, 7)] for s in (1, 2, 3))
    assert orc.septic_mul(a, orc.septic_inv(a)) == [1, 0, 0, 0, 0, 0, 0]
    assert orc.septic_mul(orc.septic_mul(a, b), c) == orc.septic_mul(a, orc.septic_mul(b, c))


@pytest.mark.parametrize("n,num_instances", [(1, 2), (3, 8), (3, 5), (4, 11), (5, 32)])
def test_ecc_quark_zerocheck_vanishes_on_a_valid_witness_and_verifies(n, num_instances):
    """create_ecc_proof (ceno_zkvm/src/scheme/cpu/mod.rs:72-316) on a witness built with the affine-addition formulas:
    the claimed sum is 0 (it is a zerocheck), every round satisfies p(0) + p(1) = claim with p(0) derived by the
    verifier, and the final claim equals the expression at the returned evaluations (its sanity-check block :268-290)."""
    xs, ys, invs = orc.ecc_quark_make_witness(700 + n, n, num_instances)
    proof = orc.ecc_quark_create_proof(num_instances, xs, ys, invs, orc.Transcript(b"ecc-kat"))
    rounds, evals, rt, terms = proof["zerocheck_proof"], proof["evals"], proof["rt"], proof["terms"]
    assert evals.shape[0] == 3 + 7 * 7
    mles_p = [[(int(a), int(b)) for a, b in m.reshape(-1, 2)] if is_ext else [(int(a), 0) for a in m] for m, is_ext, _ in proof["mles"]]
    tl = [((int(c[0]), int(c[1])), ids) for c, ids in terms]
    # the polynomial vanishes at every hypercube point, not only in sum (zero constraints under their selectors)
    for b in range(1 << n):
        assert pr.poly_eval(mles_p, tl, b) == pr.ZERO
    claim = pr.ZERO
    for j in range(n):
        msg = [tuple(int(x) for x in e) for e in rounds[j]]
        e0 = pr.esub(claim, msg[0])
        claim = pr.lagrange_eval([e0] + msg, tuple(int(x) for x in rt[j]))
    fin = [tuple(int(x) for x in e) for e in evals]
    assert fin == [pr.mle_evaluate(m_, [tuple(int(x) for x in r) for r in rt]) for m_ in mles_p]
    acc = pr.ZERO
    for c, ids in tl:
        p = c
        for i in ids:
            p = pr.emul(p, fin[i])
        acc = pr.eadd(acc, p)
    assert acc == claim
    # sel_export(rt) = eq(out_rt, lsi) * eq(rt, lsi), lsi = (0,1,..,1)   (cpu/mod.rs:274-275)
    lsi = np.array([0, 0] + [1, 0] * (n - 1), dtype=np.uint64)
    want = orc.ext_mul(orc.eq_eval(proof["out_rt"], lsi), orc.eq_eval(rt.reshape(-1), lsi))
    assert tuple(int(x) for x in want) == fin[2]


def test_ecc_quark_detects_a_broken_witness():
    n, ni = 3, 8
    xs, ys, invs = orc.ecc_quark_make_witness(9, n, ni)
    xs[2][(1 << n) + 1] ^= np.uint64(1)                  # corrupt one coordinate limb of node (1,1)
    proof = orc.ecc_quark_create_proof(ni, xs, ys, invs, orc.Transcript(b"ecc-kat"))
    mles_p = [[(int(a), int(b)) for a, b in m.reshape(-1, 2)] if is_ext else [(int(a), 0) for a in m] for m, is_ext, _ in proof["mles"]]
    tl = [((int(c[0]), int(c[1])), ids) for c, ids in proof["terms"]]
    total = pr.ZERO
    for b in range(1 << n):
        total = pr.eadd(total, pr.poly_eval(mles_p, tl, b))
    assert total != pr.ZERO


# ------------------------------------------------------------------ prove_rotation (gkr_iop/src/gkr/layer/cpu/mod.rs:249-389)
@pytest.mark.parametrize("log2,subgroup,k,n_rot", [(5, 23, 7, 2), (5, 31, 5, 1), (6, 63, 8, 3)])
def test_kat_prove_rotation_is_accepted_by_verify_rotation(log2, subgroup, k, n_rot):
    """The restated prover's output satisfies the restated verifier (zerocheck_layer.rs:678-790): the degree-2 sumcheck of
    claim 0 chains to sel(origin) * sum_i alpha_i ((1 - r) left_i + r right_i - target_i), the right evaluation derived from
    the left one equals the direct evaluation at the right point (the reference's debug assertion, cpu/mod.rs:350-362), and a
    witness violating the rotation on the subgroup is rejected."""
    import random
    rng = random.Random(log2 * 100 + k)
    n = 1 << k
    src = [orc.fill_base(7000 + i, n) for i in range(n_rot)]
    wit = src + [orc.rotation_next_base_mle(s, log2) for s in src]
    exprs = [(i, n_rot + i) for i in range(n_rot)]
    rt = orc.fill_ext(7100 + k, k)
    rounds, evals, (left, right, origin) = orc.prove_rotation(k, subgroup, log2, wit, exprs, rt, orc.Transcript(b"rot"))
    l2, r2, o2 = orc.verify_rotation(k, n_rot, rounds, evals, subgroup, log2, rt, orc.Transcript(b"rot"))
    assert (l2, r2, o2) == (left, right, origin)
    for i, (s, _) in enumerate(exprs):
        direct = orc.mle_evaluate(wit[s], False, np.array(right, dtype=np.uint64).reshape(-1))
        assert (int(direct[0]), int(direct[1])) == evals[3 * i + 1]
    # break the relation at a selected position (group element X^1 = index 2 of the first chunk)
    bad = [w.copy() for w in wit]
    bad[n_rot][2] = (int(bad[n_rot][2]) + 1) % P
    rounds_b, evals_b, _ = orc.prove_rotation(k, subgroup, log2, bad, exprs, rt, orc.Transcript(b"rot"))
    with pytest.raises(ValueError):
        orc.verify_rotation(k, n_rot, rounds_b, evals_b, subgroup, log2, rt, orc.Transcript(b"rot"))
