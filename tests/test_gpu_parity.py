"""Parity tests proper: every kernel of the hot path, called through the C ABI, against the CPU
oracle on the same seeded inputs — bit-exact (integer work).  Needs a B200: `pytest -m gpu`."""
import random

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


@pytest.fixture(scope="module")
def dev():
    import ceno_b200 as cb
    d = cb.Device(0)
    yield d
    d.close()


def rnd_point(seed, k):
    return orc.fill_ext(0xE9 ^ (seed << 8), k)


def eq_np(a, b):
    return np.array_equal(np.asarray(a, dtype=np.uint64).reshape(-1), np.asarray(b, dtype=np.uint64).reshape(-1))


# ------------------------------------------------------------------ eq-build / selectors
@pytest.mark.parametrize("k", [0, 1, 2, 5, 11, 12, 13, 16, 20])
def test_build_eq_x_r_vec(dev, k):
    import ceno_b200 as cb
    r = rnd_point(k, k)
    got = cb.build_eq_x_r_vec(dev, r)
    assert eq_np(got.evaluations(), orc.build_eq_x_r_vec(r))
    got.free()


@pytest.mark.parametrize("k,offset,ninst", [(4, 3, 7), (12, 0, 1000), (14, 100, 9000), (14, 0, 0), (13, 8191, 1), (16, 12345, 40000)])
def test_selector_prefix(dev, k, offset, ninst):
    import ceno_b200 as cb
    r = rnd_point(100 + k, k)
    got = cb.SelectorType.compute(dev, cb.SelectorType.PREFIX, r, offset=offset, num_instances=ninst)
    assert eq_np(got.evaluations(), orc.selector_compute(orc.SEL_PREFIX, r, offset=offset, num_instances=ninst))
    got.free()


def test_selector_ordered_sparse_and_quark(dev):
    import ceno_b200 as cb
    r = rnd_point(7, 9)
    for ninst in [1, 5, 16]:
        got = cb.SelectorType.compute(dev, cb.SelectorType.ORDERED_SPARSE, r, num_instances=ninst, indices=[0, 3, 17, 31], inner_vars=5)
        assert eq_np(got.evaluations(), orc.selector_compute(orc.SEL_ORDERED_SPARSE, r, num_instances=ninst, indices=[0, 3, 17, 31], inner_vars=5))
        got.free()
    for nv, ninst in [(3, 5), (1, 2), (6, 33), (9, 512), (9, 300), (13, 5000)]:
        r = rnd_point(nv + ninst, nv)
        got = cb.SelectorType.compute(dev, cb.SelectorType.QUARK_LT, r, num_instances=ninst)
        assert eq_np(got.evaluations(), orc.selector_compute(orc.SEL_QUARK_LT, r, num_instances=ninst))
        got.free()


# --------------------------------------------------------------------------- fold / evaluate
@pytest.mark.parametrize("k,is_ext", [(1, True), (1, False), (2, True), (7, False), (12, True), (18, True), (18, False)])
def test_fix_variable_and_evaluate(dev, k, is_ext):
    import ceno_b200 as cb
    n = 1 << k
    data = orc.fill_ext(11 + k, n) if is_ext else orc.fill_base(11 + k, n)
    mle = (cb.MultilinearExtension.from_evaluations_ext_vec if is_ext else cb.MultilinearExtension.from_evaluations_vec)(dev, k, data)
    r = rnd_point(5, 1)
    folded = mle.fix_variable(r)
    assert eq_np(folded.evaluations(), orc.fix_variable(data, is_ext, r))
    pt = rnd_point(6, k)
    assert eq_np(mle.evaluate(pt), orc.mle_evaluate(data, is_ext, pt))
    folded.free(); mle.free()


def test_non_canonical_inputs_are_reduced(dev):
    import ceno_b200 as cb
    k = 6
    n = 1 << k
    canon = orc.fill_ext(99, n)
    raw = canon.copy()
    small = raw < np.uint64(0xFFFFFFFF)      # x + p still fits in u64
    raw[small] = raw[small] + np.uint64(P)   # non-canonical representative of the same element
    assert small.sum() == 0 or (raw[small] >= np.uint64(P)).all()
    raw[0] = np.uint64(P)                    # p itself == 0
    canon[0] = 0
    raw[3] = np.uint64(2**64 - 1)
    canon[3] = np.uint64(2**64 - 1 - P)
    a = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, raw)
    b = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, orc.fill_ext(98, n))
    c = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, orc.fill_ext(97, n))
    terms = [([1, 0], [0, 1, 2])]
    for flags in (0, cb.IOPProverState.FORCE_GENERIC):
        got = cb.IOPProverState.prove(dev, [a, b, c], terms, k, 3, transcript=cb.StandInTranscript(b"nc"), flags=flags)
        want = orc.sumcheck_prove([(canon, True, k), (orc.fill_ext(98, n), True, k), (orc.fill_ext(97, n), True, k)], terms, k, 3,
                                  transcript=orc.Transcript(b"nc"))
        for g, w in zip(got, want):
            assert eq_np(g, w)


# ------------------------------------------------------------------------------- sumcheck
def t3_inputs(k, seed=0):
    n = 1 << k
    w = orc.fill_ext(0xE9 + seed, k)
    eq = orc.build_eq_x_r_vec(w)
    a = orc.fill_ext(0xC0FFEE ^ (1 + 16 * seed), n)
    b = orc.fill_ext(0xC0FFEE ^ (2 + 16 * seed), n)
    return eq, a, b


@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 10, 13, 16, 20])
@pytest.mark.parametrize("mode", ["fused", "nofuse", "generic", "device", "notail", "device_notail", "nomid", "device_nomid"])
def test_t3_sumcheck_bit_exact(dev, k, mode):
    """BASELINE config #2 shape (eq*A*B, degree 3, all ext) — every kernel variant."""
    import ceno_b200 as cb
    if mode not in ("fused", "device") and k == 20:
        pytest.skip("large size covered by the fused path")
    eq, a, b = t3_inputs(k)
    terms = [([1, 0], [0, 1, 2])]
    want = orc.sumcheck_prove([(eq, True, k), (a, True, k), (b, True, k)], terms, k, 3, transcript=orc.Transcript(b"t3"))
    mles = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, x) for x in (eq, a, b)]
    flags = {"fused": 0, "device": 0, "nofuse": cb.IOPProverState.NO_FUSE, "generic": cb.IOPProverState.FORCE_GENERIC,
             "notail": cb.IOPProverState.NO_TAIL, "device_notail": cb.IOPProverState.NO_TAIL,
             "nomid": cb.IOPProverState.NO_MID, "device_nomid": cb.IOPProverState.NO_MID}[mode]
    tr = cb.StandInTranscript(b"t3")
    got = cb.IOPProverState.prove(dev, mles, terms, k, 3, transcript=tr, flags=flags, device_challenger=mode.startswith("device"))
    for g, w in zip(got, want):
        assert eq_np(g, w)
    # inputs are shared (Arc) in the reference: the prover must not have modified them
    assert eq_np(mles[1].evaluations(), a)
    for m in mles:
        m.free()


@pytest.mark.parametrize("k,mode,pos", [(20, "host", 0), (21, "device", 0), (22, "host", 1), (21, "host", 2), (20, "device_nomid", 0),
                                        (21, "notail", 0), (12, "host", 0), (19, "device", 0), (20, "nofuse", 0), (20, "generic", 1)])
def test_t3_virtual_eq_bit_exact(dev, k, mode, pos):
    """eq handed over as its point (CG_MLE_EQ): the split-eq rounds (k >= 20) and every fallback must
    give the proof of the materialised table, bit for bit."""
    import ceno_b200 as cb
    n = 1 << k
    w = orc.fill_ext(0xE9 + k, k)
    a = orc.fill_ext(0xC0FFEE ^ (1 + 16 * k), n)
    b = orc.fill_ext(0xC0FFEE ^ (2 + 16 * k), n)
    host = [(a, True, k), (b, True, k)]
    host.insert(pos, (orc.build_eq_x_r_vec(w), True, k))
    terms = [([1, 0], [0, 1, 2])]
    want = orc.sumcheck_prove(host, terms, k, 3, transcript=orc.Transcript(b"veq"))
    mles = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, x) for x in (a, b)]
    mles.insert(pos, cb.EqPolynomial(dev, w))
    flags = {"host": 0, "device": 0, "device_nomid": cb.IOPProverState.NO_MID, "notail": cb.IOPProverState.NO_TAIL,
             "nofuse": cb.IOPProverState.NO_FUSE, "generic": cb.IOPProverState.FORCE_GENERIC}[mode]
    got = cb.IOPProverState.prove(dev, mles, terms, k, 3, transcript=cb.StandInTranscript(b"veq"), flags=flags,
                                  device_challenger=mode.startswith("device"))
    for g, x in zip(got, want):
        assert eq_np(g, x)
    for m in mles:
        m.free()


def test_virtual_eq_step_api_other_shapes(dev):
    """Virtual eq through the step API (round_eval / bind / peek, bind without an evaluation), with a
    coefficient != 1 and inside a two-term expression: all fall back to exact table semantics."""
    import ceno_b200 as cb
    k = 20
    n = 1 << k
    w = orc.fill_ext(0x51, k)
    eq = orc.build_eq_x_r_vec(w)
    a, b = orc.fill_ext(0xAB1, n), orc.fill_ext(0xAB2, n)
    terms = [([1, 0], [0, 1, 2])]
    mles = [cb.EqPolynomial(dev, w)] + [cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, x) for x in (a, b)]
    st = cb.IOPProverState(dev, mles, terms, k, 3)
    cur = [eq, a, b]
    for j in range(4):     # two split rounds, then the switch to the materialised state inside the step API
        msg = st.round_eval()
        want, _, _ = orc.sumcheck_prove([(c, True, k - j) for c in cur], terms, k - j, 3, challenge_fn=lambda *_: np.array([1, 2], dtype=np.uint64))
        assert eq_np(msg, want[0])
        r = rnd_point(70 + j, 1)
        st.bind(r)
        cur = [orc.fix_variable(c, True, r) for c in cur]
    assert eq_np(st.peek(0), cur[0]) and eq_np(st.peek(2), cur[2])
    st.close()
    # peek / bind-without-eval while still in split mode
    st = cb.IOPProverState(dev, mles, terms, k, 3)
    cur = [eq, a, b]
    st.round_eval()
    r = rnd_point(90, 1)
    st.bind(r)
    cur = [orc.fix_variable(c, True, r) for c in cur]
    assert eq_np(st.peek(0), cur[0])
    msg = st.round_eval()
    want, _, _ = orc.sumcheck_prove([(c, True, k - 1) for c in cur], terms, k - 1, 3, challenge_fn=lambda *_: np.array([1, 2], dtype=np.uint64))
    assert eq_np(msg, want[0])
    st.close()
    # coefficient != 1 and a two-term expression
    for tms, deg in (([([5, 9], [0, 1, 2])], 3), ([([1, 0], [0, 1, 2]), ([3, 4], [0, 1])], 3)):
        want = orc.sumcheck_prove([(eq, True, k), (a, True, k), (b, True, k)], tms, k, deg, transcript=orc.Transcript(b"v2"))
        got = cb.IOPProverState.prove(dev, mles, tms, k, deg, transcript=cb.StandInTranscript(b"v2"))
        for g, x in zip(got, want):
            assert eq_np(g, x)
    for m in mles:
        m.free()


def test_t3_with_coefficient_and_python_transcript(dev):
    import ceno_b200 as cb
    k = 9
    eq, a, b = t3_inputs(k, seed=3)
    coeff = [12345678901234567, 76543210987654321]
    terms = [(coeff, [2, 0, 1])]

    def cb_fn(j, evals):
        t = orc.Transcript(b"py")
        t.append_message(int(j).to_bytes(8, "little"))
        t.append_ext(evals)
        return t.sample(b"x")
    want = orc.sumcheck_prove([(eq, True, k), (a, True, k), (b, True, k)], terms, k, 3, challenge_fn=cb_fn)
    mles = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, x) for x in (eq, a, b)]
    got = cb.IOPProverState.prove(dev, mles, terms, k, 3, challenge_fn=cb_fn)
    for g, w in zip(got, want):
        assert eq_np(g, w)


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 8])
def test_generic_terms_mixed_base_ext(dev, degree):
    """Zerocheck-like instance: base-field witness columns, ext selectors, many monomial terms of
    degree <= d, lower-degree terms included (gkr_iop/src/gkr/layer/zerocheck_layer.rs:86-207)."""
    import ceno_b200 as cb
    rng = random.Random(degree)
    k, m = 8, 7
    n = 1 << k
    host, mles = [], []
    for i in range(m):
        is_ext = i >= 4
        d = orc.fill_ext(500 + i, n) if is_ext else orc.fill_base(500 + i, n)
        host.append((d, is_ext, k))
        mles.append((cb.MultilinearExtension.from_evaluations_ext_vec if is_ext else cb.MultilinearExtension.from_evaluations_vec)(dev, k, d))
    terms = []
    for _ in range(12):
        nf = rng.randint(0, degree)
        terms.append(([rng.randrange(P), rng.randrange(P)], [rng.randrange(m) for _ in range(nf)]))
    terms.append(([rng.randrange(P), 0], list(range(min(degree, m)))))
    want = orc.sumcheck_prove(host, terms, k, degree, transcript=orc.Transcript(b"gen"))
    for dc, flags in ((False, 0), (True, 0), (False, cb.IOPProverState.NO_PLAN)):     # grouped plan and term-by-term kernel
        got = cb.IOPProverState.prove(dev, mles, terms, k, degree, transcript=cb.StandInTranscript(b"gen"), device_challenger=dc, flags=flags)
        for g, w in zip(got, want):
            assert eq_np(g, w)


def test_zerocheck_layer_shape(dev):
    """ZerocheckLayerProver::prove shape (gkr_iop/src/gkr/layer/cpu/mod.rs:99-239): selector eq MLEs from
    SelectorType::compute, base-field witness columns, alpha-weighted monomial terms sel_g * prod(witness)."""
    import ceno_b200 as cb
    rng = random.Random(42)
    k, n_wit = 10, 12
    n = 1 << k
    pt = rnd_point(77, k)
    sel_specs = [(cb.SelectorType.WHOLE, orc.SEL_WHOLE, {}), (cb.SelectorType.PREFIX, orc.SEL_PREFIX, dict(offset=3, num_instances=700)),
                 (cb.SelectorType.PREFIX, orc.SEL_PREFIX, dict(offset=0, num_instances=1000))]
    wit_h = [orc.fill_base(600 + i, n) for i in range(n_wit)]
    mles = [cb.MultilinearExtension.from_evaluations_vec(dev, k, w) for w in wit_h]
    host = [(w, False, k) for w in wit_h]
    for kind_d, kind_o, kw in sel_specs:
        mles.append(cb.SelectorType.compute(dev, kind_d, pt, **kw))
        host.append((orc.selector_compute(kind_o, pt, **kw), True, k))
    alpha = [rng.randrange(P), rng.randrange(P)]
    terms, apow = [], (1, 0)
    for e in range(30):                      # alpha^e * sel_g * w_a * w_b (* w_c)
        g = n_wit + rng.randrange(len(sel_specs))
        ws = [rng.randrange(n_wit) for _ in range(rng.randint(1, 3))]
        terms.append(([apow[0], apow[1]], sorted([g] + ws, reverse=True)))     # product sorted by descending witness id
        apow = ((apow[0] * alpha[0] + 7 * apow[1] * alpha[1]) % P, (apow[0] * alpha[1] + apow[1] * alpha[0]) % P)
    want = orc.sumcheck_prove(host, terms, k, 4, transcript=orc.Transcript(b"zc"))
    for flags in (0, cb.IOPProverState.NO_PLAN):
        got = cb.IOPProverState.prove(dev, mles, terms, k, 4, transcript=cb.StandInTranscript(b"zc"), flags=flags)
        for g, w in zip(got, want):
            assert eq_np(g, w)


@pytest.mark.parametrize("chip_vars", [[14, 11, 11, 7, 3, 0], [17, 17, 13], [5, 9]])
def test_batched_main_constraints_mixed_sizes(dev, chip_vars):
    """prove_batched_main_constraints shape (ceno_zkvm/src/scheme/cpu/mod.rs:1052-1390): one sumcheck over the
    union of all chips' monomial terms, chips of different num_vars ("frontload" embedding, pinned by
    tests/test_oracle_kat.py::test_mixed_size_frontload_sumcheck_verifies)."""
    import ceno_b200 as cb
    k = max(chip_vars)
    host, mles, terms = [], [], []
    rng = random.Random(sum(chip_vars))
    for c, kv in enumerate(chip_vars):
        n = 1 << kv
        base = len(host)
        sel = orc.build_eq_x_r_vec(rnd_point(300 + c, kv)) if kv else np.array([3, 4], dtype=np.uint64)
        host.append((sel, True, kv))
        for wi in range(3):
            host.append((orc.fill_base(1000 * c + wi, n), False, kv))
        al = lambda: [rng.randrange(P), rng.randrange(P)]
        terms += [(al(), [base, base + 1, base + 2]), (al(), [base, base + 3]), (al(), [base, base + 1, base + 2, base + 3]),
                  (al(), [base + 1])]
    want = orc.sumcheck_prove(host, terms, k, 4, transcript=orc.Transcript(b"bm"))
    for d, is_ext, kv in host:
        mles.append((cb.MultilinearExtension.from_evaluations_ext_vec if is_ext else cb.MultilinearExtension.from_evaluations_vec)(dev, kv, d))
    for dc in (False, True):
        got = cb.IOPProverState.prove(dev, mles, terms, k, 4, transcript=cb.StandInTranscript(b"bm"), device_challenger=dc)
        for g, w in zip(got, want):
            assert eq_np(g, w)
    for m in mles:
        m.free()


def test_step_api_and_peek(dev):
    import ceno_b200 as cb
    k = 6
    eq, a, b = t3_inputs(k, seed=5)
    mles = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, x) for x in (eq, a, b)]
    st = cb.IOPProverState(dev, mles, [([1, 0], [0, 1, 2])], k, 3)
    cur = [eq, a, b]
    for j in range(k):
        msg = st.round_eval()
        want, _, _ = orc.sumcheck_prove([(c, True, k - j) for c in cur], [([1, 0], [0, 1, 2])], k - j, 3, challenge_fn=lambda *_: np.array([1, 2], dtype=np.uint64))
        assert eq_np(msg, want[0])
        r = rnd_point(40 + j, 1)
        st.bind(r)
        cur = [orc.fix_variable(c, True, r) for c in cur]
        if j < k - 1:
            assert eq_np(st.peek(1), cur[1])
    assert eq_np(st.get_mle_flatten_final_evaluations(), np.concatenate(cur))
    st.close()


def test_occupied_prefix_shorter_than_hypercube(dev):
    """evaluations_len() < 2^num_vars: implicit zero tail."""
    import ceno_b200 as cb
    k = 7
    n = 1 << k
    occ = 77
    a_full = np.zeros(n, np.uint64)
    a_full[:occ] = orc.fill_base(1, occ)
    e = orc.fill_ext(2, n)
    a = cb.MultilinearExtension.from_evaluations_vec(dev, k, a_full[:occ])
    em = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, e)
    terms = [([5, 0], [0, 1])]
    want = orc.sumcheck_prove([(a_full, False, k), (e, True, k)], terms, k, 2, transcript=orc.Transcript(b"occ"))
    got = cb.IOPProverState.prove(dev, [a, em], terms, k, 2, transcript=cb.StandInTranscript(b"occ"))
    for g, w in zip(got, want):
        assert eq_np(g, w)


def test_errors_mirror_reference_behaviour(dev):
    import ceno_b200 as cb
    a = cb.MultilinearExtension.from_evaluations_ext_vec(dev, 3, orc.fill_ext(1, 8))
    b = cb.MultilinearExtension.from_evaluations_ext_vec(dev, 2, orc.fill_ext(2, 4))
    with pytest.raises(cb.CenoB200Error) as ei:      # mixed num_vars: served by prove(), not by the step API
        cb.IOPProverState(dev, [a, b], [([1, 0], [0, 1])], 3, 2)
    assert ei.value.code == 3
    with pytest.raises(cb.CenoB200Error) as ei:      # term with more factors than the stated degree
        cb.IOPProverState(dev, [a], [([1, 0], [0, 0, 0])], 3, 2)
    assert ei.value.code == 2
    st = cb.IOPProverState(dev, [a], [([1, 0], [0])], 3, 1)
    with pytest.raises(cb.CenoB200Error) as ei:      # final evals before all variables are bound
        st.get_mle_flatten_final_evaluations()
    assert ei.value.code == 6
    st.close()
    with pytest.raises(cb.CenoB200Error):            # Prefix selector: end > 2^num_vars (selector.rs:144-150)
        cb.SelectorType.compute(dev, cb.SelectorType.PREFIX, rnd_point(1, 3), offset=5, num_instances=4)


def test_zero_variable_sumcheck_is_noop(dev):
    import ceno_b200 as cb
    a = cb.MultilinearExtension.from_evaluations_ext_vec(dev, 0, np.array([5, 6], dtype=np.uint64))
    rounds, fin, chal = cb.IOPProverState.prove(dev, [a], [([1, 0], [0])], 0, 1, transcript=cb.StandInTranscript(b"z"))
    assert rounds.size == 0 and chal.size == 0 and eq_np(fin, [5, 6])


# ------------------------------------------------------------------------- wit_infer / tower
def test_wit_infer_by_monomial_expr(dev):
    import ceno_b200 as cb
    from oracle import pyref as pr
    k, n = 6, 64
    b0 = orc.fill_base(1, n)
    e1 = orc.fill_ext(2, n)
    e2 = orc.fill_ext(3, n)
    mles = [cb.MultilinearExtension.from_evaluations_vec(dev, k, b0), cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, e1),
            cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, e2)]
    terms = [([3, 4], [0, 1]), ([5, 0], [1, 2, 2]), ([7, 1], [])]
    got = pr.to_pairs(cb.wit_infer_by_monomial_expr(dev, mles, terms, k).evaluations())
    hp = [[(int(x), 0) for x in b0], pr.to_pairs(e1), pr.to_pairs(e2)]
    for b in range(n):
        assert got[b] == pr.poly_eval(hp, [(tuple(c), ids) for c, ids in terms], b)


def _tower_case(dev, prod_nvs, logup_nvs, with_p, seed):
    import ceno_b200 as cb
    specs, o_prod, o_lk = [], [], []
    s = seed
    for nv in prod_nvs:
        f1, f2 = orc.fill_ext(s, 1 << (nv - 1)), orc.fill_ext(s + 1, 1 << (nv - 1))
        s += 2
        pw, layers = orc.infer_tower_product_witness(nv, f1, f2)
        o_prod.append((pw, nv, layers))
        specs.append(cb.TowerProverSpec([cb.MultilinearExtension.from_evaluations_ext_vec(dev, nv - 1, f1),
                                         cb.MultilinearExtension.from_evaluations_ext_vec(dev, nv - 1, f2)], nv, False))
    for nv in logup_nvs:
        q1, q2 = orc.fill_ext(s, 1 << nv), orc.fill_ext(s + 1, 1 << nv)
        p1 = orc.fill_ext(s + 2, 1 << nv) if with_p else None
        p2 = orc.fill_ext(s + 3, 1 << nv) if with_p else None
        s += 4
        lw, layers = orc.infer_tower_logup_witness(nv, p1, p2, q1, q2)
        o_lk.append((lw, nv + 1, layers))
        mk = lambda x: cb.MultilinearExtension.from_evaluations_ext_vec(dev, nv, x) if x is not None else None
        specs.append(cb.TowerProverSpec([mk(p1), mk(p2), mk(q1), mk(q2)], nv, True))
    tw = cb.TowerProver(dev, specs)
    # output layer values (get_output_evals)
    for i, (_, _, layers) in enumerate(o_prod):
        assert eq_np(tw.output_evals(i)[:4], np.concatenate(layers[0]))
    for i, (_, _, layers) in enumerate(o_lk):
        assert eq_np(tw.output_evals(len(o_prod) + i), np.concatenate(layers[0]))
    t_o, t_d = orc.Transcript(b"tower"), cb.StandInTranscript(b"tower")
    want_proof, want_point = orc.tower_create_proof([(w, l) for w, l, _ in o_prod], [(w, l) for w, l, _ in o_lk], t_o)
    got_proof, got_point = tw.create_proof(t_d)
    assert eq_np(got_proof, want_proof)
    assert eq_np(got_point, want_point)
    assert int(t_d.state[0]) == t_o.state
    tw.close()


def test_interleaving_mles_to_mles(dev):
    """ceno_zkvm/src/scheme/utils.rs:402-462 on the device: the reference's literal vectors (utils.rs:968-1065) and
    seeded cases against the oracle (base and ext records, instance counts off the powers of two, R up to 70)."""
    import ceno_b200 as cb
    from ceno_b200 import api
    from oracle import pyref as pr

    def run(cols, is_ext, num_instances, num_limbs, default):
        host = [(c, e) for c, e in zip(cols, is_ext)]
        want = orc.interleaving_mles_to_mles(host, num_instances, num_limbs, default)
        mles = []
        for c, e in host:
            nv = max((c.size // (2 if e else 1)) - 1, 0).bit_length()
            mles.append((cb.MultilinearExtension.from_evaluations_ext_vec if e else cb.MultilinearExtension.from_evaluations_vec)(dev, nv, c))
        got, buf = api.interleaving_mles_to_mles(dev, mles, num_instances, num_limbs, default)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert eq_np(g.evaluations(), w)
        buf.free()
        for m in mles:
            m.free()
        return want

    E = lambda *xs: np.array([v for x in xs for v in (x, 0)], dtype=np.uint64)
    w = run([E(1, 2), E(3, 4), E(5, 6), E(7, 8)], [True] * 4, 2, 2, [1, 0])
    assert pr.to_pairs(w[0]) == [(1, 0), (3, 0), (5, 0), (7, 0)] and pr.to_pairs(w[1]) == [(2, 0), (4, 0), (6, 0), (8, 0)]
    run([E(1, 2), E(3, 4), E(5, 6)], [True] * 3, 2, 2, [0, 0])
    run([E(1, 0), E(3, 0), E(5, 0)], [True] * 3, 1, 2, [1, 0])
    run([E(2), E(3)], [True] * 2, 1, 2, [1, 0])
    rng = random.Random(77)
    for R, log_n, ninst, limbs in [(1, 6, 64, 2), (2, 7, 100, 2), (5, 10, 1000, 2), (17, 12, 4096, 2), (33, 9, 300, 2), (70, 8, 255, 2), (3, 11, 2048, 4),
                                   (9, 5, 17, 1)]:
        n = 1 << log_n
        ext_flags = [rng.random() < 0.5 for _ in range(R)]
        cols = [orc.fill_ext(5000 + 97 * R + i, n) if e else orc.fill_base(5000 + 97 * R + i, n) for i, e in enumerate(ext_flags)]
        run(cols, ext_flags, ninst, limbs, [rng.randrange(P), rng.randrange(P)])


def test_chip_tower_flow_records_to_proof(dev):
    """The tower half of create_chip_proof (ceno_zkvm/src/scheme/prover.rs:717, cpu/mod.rs:586-700): record MLEs ->
    interleaving_mles_to_mles (pad 1 for read/write, alpha for lookups) -> tower build -> tower proof, all on the
    device, against the same chain in the oracle."""
    import ceno_b200 as cb
    from ceno_b200 import api
    log_n, ninst = 10, 1000
    n = 1 << log_n
    alpha = [12345, 678]
    r_recs = [orc.fill_ext(9100 + i, n) for i in range(3)]        # read records  (R = 3 -> padded to 4 per instance)
    w_recs = [orc.fill_ext(9200 + i, n) for i in range(2)]        # write records
    lk_recs = [orc.fill_ext(9300 + i, n) for i in range(5)]       # lookup denominators
    specs_dev, o_prod, o_lk, keep = [], [], [], []
    for recs in (r_recs, w_recs):
        want = orc.interleaving_mles_to_mles([(c, True) for c in recs], ninst, 2, [1, 0])
        nv = (want[0].size // 2).bit_length()      # leaves of 2^(nv-1)
        pw, layers = orc.infer_tower_product_witness(nv, want[0], want[1])
        o_prod.append((pw, nv))
        mles = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, log_n, c) for c in recs]
        got, buf = api.interleaving_mles_to_mles(dev, mles, ninst, 2, [1, 0])
        keep.append(buf)
        specs_dev.append(cb.TowerProverSpec(got, nv, False))
    want = orc.interleaving_mles_to_mles([(c, True) for c in lk_recs], ninst, 2, alpha)
    nv = (want[0].size // 2).bit_length() - 1
    lw, layers = orc.infer_tower_logup_witness(nv, None, None, want[0], want[1])
    o_lk.append((lw, nv + 1))
    mles = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, log_n, c) for c in lk_recs]
    got, buf = api.interleaving_mles_to_mles(dev, mles, ninst, 2, alpha)
    keep.append(buf)
    specs_dev.append(cb.TowerProverSpec([None, None, got[0], got[1]], nv, True))
    tw = cb.TowerProver(dev, specs_dev)
    t_o, t_d = orc.Transcript(b"chip"), cb.StandInTranscript(b"chip")
    want_proof, want_point = orc.tower_create_proof(o_prod, o_lk, t_o)
    got_proof, got_point = tw.create_proof(t_d)
    assert eq_np(got_proof, want_proof) and eq_np(got_point, want_point)
    tw.close()


@pytest.mark.parametrize("log_n,ninst,n_r,n_w,n_lk,with_num,base_mix", [
    (12, 3000, 3, 2, 5, False, False),      # the opcode-chip shape: read / write / lookup groups, padded instances
    (12, 4096, 1, 4, 9, True, True),        # explicit numerators, base-field records mixed in, full instances
    (13, 5000, 37, 2, 70, False, False),    # many records per row (keccak-like: padded_ops 64 / 128)
    (7, 100, 3, 2, 1094, False, False),     # keccak's 1094 lookup records per row (lookup_keccakf.rs:97-101), scaled-down rows
    (10, 1000, 9, 12, 40, True, True),      # leaf layers of 2^14 .. 2^16 read by the split-eq rounds (lanes over the record rows)
    (9, 512, 16, 16, 300, False, True),
])
def test_tower_over_virtual_leaves_bit_exact(dev, monkeypatch, log_n, ninst, n_r, n_w, n_lk, with_num, base_mix):
    """cfg-4: the tower built and proven over VIRTUAL leaf layers (GpuVirtualInterleavedExt: records described, the interleaved
    fan-in leaves never stored) gives the oracle's proof of the materialised chain, bit for bit."""
    import ceno_b200 as cb
    monkeypatch.setenv("CG_TOWER_VEQ_MIN_NV", "10")   # split-eq rounds on these scaled-down layers too
    n = 1 << log_n
    alpha = [12345, 678]

    def recs(seed, cnt):
        out = []
        for i in range(cnt):
            is_ext = not (base_mix and i % 2 == 1)
            out.append((orc.fill_ext(seed + i, n) if is_ext else orc.fill_base(seed + i, n), is_ext))
        return out

    def dev_mles(rs):
        return [(cb.MultilinearExtension.from_evaluations_ext_vec if e else cb.MultilinearExtension.from_evaluations_vec)(dev, log_n, c) for c, e in rs]
    r_recs, w_recs, lk_recs = recs(9100, n_r), recs(9200, n_w), recs(9300, n_lk)
    num_recs = recs(9400, n_lk) if with_num else None
    o_prod, o_lk, vspecs, keep = [], [], [], []
    for rs in (r_recs, w_recs):
        want = orc.interleaving_mles_to_mles(rs, ninst, 2, [1, 0])
        nv = (want[0].size // 2).bit_length()
        o_prod.append((orc.infer_tower_product_witness(nv, want[0], want[1])[0], nv))
        ms = dev_mles(rs)
        keep += ms
        vspecs.append(cb.VirtualTowerSpec(ms, ninst, [1, 0], False))
    want = orc.interleaving_mles_to_mles(lk_recs, ninst, 2, alpha)
    nv = (want[0].size // 2).bit_length() - 1
    if with_num:
        wn = orc.interleaving_mles_to_mles(num_recs, ninst, 2, [1, 0])
        o_lk.append((orc.infer_tower_logup_witness(nv, wn[0], wn[1], want[0], want[1])[0], nv + 1))
    else:
        o_lk.append((orc.infer_tower_logup_witness(nv, None, None, want[0], want[1])[0], nv + 1))
    ms, nms = dev_mles(lk_recs), (dev_mles(num_recs) if with_num else None)
    keep += ms + (nms or [])
    vspecs.append(cb.VirtualTowerSpec(ms, ninst, alpha, True, numerators=nms))
    tw = cb.TowerProver.from_records(dev, vspecs)
    want_proof, want_point = orc.tower_create_proof(o_prod, o_lk, orc.Transcript(b"virt"))
    got_proof, got_point = tw.create_proof(cb.StandInTranscript(b"virt"))
    assert eq_np(got_proof, want_proof) and eq_np(got_point, want_point)
    tw.close()
    for m in keep:
        m.free()


@pytest.mark.parametrize("prod_nvs,logup_nvs,with_p", [
    ([4], [], False),                 # test_tower_proof_various_prod_size style (scheme/tests.rs:447-500)
    ([2], [], False), ([10], [], False),
    ([], [3], False), ([], [6], True),
    ([6, 6], [5], False),             # read + write + lookup towers of one chip
    ([9, 5], [8, 3], True),           # specs of different depth drop out of the early rounds
    ([3] * 9, [2] * 5, False),        # more specs than the specialised kernel's table -> generic kernel
])
def test_tower_proof_bit_exact(dev, prod_nvs, logup_nvs, with_p):
    _tower_case(dev, prod_nvs, logup_nvs, with_p, seed=1000 + 7 * len(prod_nvs) + len(logup_nvs))


def test_tower_large(dev):
    _tower_case(dev, [17, 17], [16], False, seed=77)


@pytest.mark.parametrize("prod_nvs,logup_nvs,with_p,env", [
    ([16, 15], [15], True, {}),                                   # layers >= 2^13 run split-eq rounds (tveq_round_kernel), claim-derived
    ([], [16, 14], False, {}),
    ([15] * 3, [14], False, {"CG_TOWER_VEQ_NOCLAIM": "1"}),       # round 0 without a claim: the three-sum variant
    ([16], [15], True, {"CG_TAIL_CLUSTER": "1"}),                 # small tail: more split rounds per layer
    ([16, 15], [15], True, {"CG_TOWER_VEQ": "0"}),                # A/B: the eq table everywhere
])
def test_tower_split_eq_layers_bit_exact(dev, monkeypatch, prod_nvs, logup_nvs, with_p, env):
    """The eq factor of a tower layer's sumcheck stays virtual on large layers (ceno_zkvm/src/scheme/cpu/mod.rs:417-485 without
    the eq table): same proof as the oracle's table formulation, bit for bit."""
    monkeypatch.setenv("CG_TOWER_VEQ_MIN_NV", "10")   # (default: layers of 2^20 points and more)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    _tower_case(dev, prod_nvs, logup_nvs, with_p, seed=4100 + len(prod_nvs) + 3 * len(logup_nvs))


@pytest.mark.parametrize("shapes", [[(3, 1)], [(6, 3)], [(5, 2), (5, 1)], [(7, 2), (4, 3), (6, 1)], [(2, 1), (9, 4)], [(14, 5), (12, 3)]])
def test_basefold_commit_and_batch_open_bit_exact(dev, shapes):
    """f-2 / a9: commitment roots, every prover message, the final message, the proof-of-work witness and every query opening
    equal the oracle's (oracle/basefold.py, pinned on the in-tree verifier restatement ceno_recursion_v2/src/pcs/mod.rs);
    the oracle's verifier accepts the device proof."""
    import random
    import ceno_b200 as cb
    from ceno_b200 import api
    from oracle import basefold as bf
    from oracle import pyref as pr
    rng = random.Random(4242 + len(shapes) + shapes[0][0])
    p2 = orc.p2_params(seed=5)
    api.poseidon2_set_params(dev, [[int(p2.ext_rc[r][i]) for i in range(8)] for r in range(8)], [int(x) for x in p2.int_rc], [int(x) for x in p2.diag], 0)
    params_o = bf.Params(rate_log=1, n_queries=5, pow_bits=3)
    params_d = api.BasefoldParams(rate_log=1, n_queries=5, pow_bits=3)
    oc, dc, points, evals, bufs = [], [], [], [], []
    for nv, width in shapes:
        cols = [orc.fill_base(rng.randrange(1 << 30), 1 << nv) for _ in range(width)]
        oc.append(bf.commit(p2, params_o, cols, nv))
        buf = dev.to_device(np.concatenate(cols))
        bufs.append(buf)
        dc.append(api.BasefoldCommitment(dev, buf, width, nv, params_d))
        assert [int(x) for x in dc[-1].root] == oc[-1]["root"]
        pt = [(rng.randrange(P), rng.randrange(P)) for _ in range(nv)]
        points.append(pt)
        evals.append([tuple(int(x) for x in orc.mle_evaluate(c, False, np.array(pt, dtype=np.uint64).reshape(-1))) for c in cols])
    want = bf.batch_open(p2, params_o, oc, points, evals, orc.Transcript(b"pcs"))
    got = api.basefold_batch_open(dev, dc, [np.array(p, dtype=np.uint64).reshape(-1) for p in points],
                                  [np.array(e, dtype=np.uint64).reshape(-1) for e in evals], params_d, cb.StandInTranscript(b"pcs"))
    for key in ("sumcheck", "commits", "final_message", "pow_witness"):
        assert got[key] == want[key] or [list(x) for x in got[key]] == [list(x) for x in want[key]], key
    assert len(got["queries"]) == len(want["queries"])
    for g, x in zip(got["queries"], want["queries"]):
        assert g["index"] == x["index"]
        for gi, xi in zip(g["inputs"], x["inputs"]):
            assert gi["opened"] == [int(v) for v in xi["opened"]] and gi["path"] == [list(d) for d in xi["path"]]
        for gc, xc in zip(g["commit_phase"], x["commit_phase"]):
            assert tuple(gc["sibling"]) == tuple(xc["sibling"]) and gc["path"] == [list(d) for d in xc["path"]]
    assert bf.batch_verify(p2, params_o, shapes, [c["root"] for c in oc], points, evals, got, orc.Transcript(b"pcs"))
    for c in dc:
        c.free()
    for b in bufs:
        b.free()


@pytest.mark.parametrize("log2,subgroup,k,n_rot", [(5, 23, 7, 2), (5, 31, 5, 1), (6, 63, 9, 3), (5, 23, 16, 2)])
def test_prove_rotation_bit_exact(dev, log2, subgroup, k, n_rot):
    """f-3: prove_rotation end to end (gkr_iop/src/gkr/layer/cpu/mod.rs:249-389) — rotated MLEs + selector pre-passes, the
    degree-2 sumcheck, left / right point evaluations — against the restatement (itself accepted by the restated
    verify_rotation, tests/test_oracle_kat.py)."""
    import ceno_b200 as cb
    from ceno_b200 import gkr
    n = 1 << k
    src = [orc.fill_base(7000 + i, n) for i in range(n_rot)]
    wit_h = src + [orc.rotation_next_base_mle(s, log2) for s in src]
    exprs = [(i, n_rot + i) for i in range(n_rot)]
    rt = orc.fill_ext(7100 + k, k)
    rounds, evals, (left, right, origin) = orc.prove_rotation(k, subgroup, log2, wit_h, exprs, rt, orc.Transcript(b"rot"))
    wit = [cb.MultilinearExtension.from_evaluations_vec(dev, k, w) for w in wit_h]
    lp, pts = gkr.prove_rotation(dev, k, subgroup, log2, wit, exprs, rt, [], cb.StandInTranscript(b"rot"))
    assert eq_np(lp.proof, rounds)
    assert eq_np(lp.evals, np.array(evals, dtype=np.uint64))
    assert eq_np(pts.left, np.array(left, dtype=np.uint64)) and eq_np(pts.right, np.array(right, dtype=np.uint64))
    assert eq_np(pts.origin, np.array(origin, dtype=np.uint64))
    for m in wit:
        m.free()


@pytest.mark.parametrize("case", ["default", "no_derive", "no_persist", "w_one", "w_zero"])
def test_split_eq_claim_derived_rounds(dev, case):
    """Split-eq rounds >= 1 accumulate q(1) and the X^2 coefficient only and solve q(0) from the running claim
    ((1 - w_j) q(0) + w_j q(1) = q_{j-1}(r_{j-1})); the flag CG_SC_NO_DERIVE and a point with w_j = 1 (no inverse of
    1 - w_j) take the three-sum rounds.  All must give the oracle's bits."""
    import ceno_b200 as cb
    k = 21
    n = 1 << k
    w = rnd_point(7700, k).copy()
    if case == "w_one":
        w[2 * 2], w[2 * 2 + 1] = 1, 0
    if case == "w_zero":
        w[2 * 1], w[2 * 1 + 1] = 0, 0
        w[2 * 3], w[2 * 3 + 1] = 0, 0
    a_h, b_h = orc.fill_ext(7701, n), orc.fill_ext(7702, n)
    a = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, a_h)
    b = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, b_h)
    terms = [([1, 0], [0, 1, 2])]
    want = orc.sumcheck_prove([(orc.build_eq_x_r_vec(w), True, k), (a_h, True, k), (b_h, True, k)], terms, k, 3, transcript=orc.Transcript(b"derive"))
    flags = 64 if case == "no_derive" else (128 if case == "no_persist" else 0)   # CG_SC_NO_DERIVE / CG_SC_NO_PERSIST
    for devch in (False, True):
        got = cb.IOPProverState.prove(dev, [cb.EqPolynomial(dev, w), a, b], terms, k, 3, transcript=cb.StandInTranscript(b"derive"),
                                      device_challenger=devch, flags=flags)
        for g, x in zip(got, want):
            assert eq_np(g, x), (case, devch)
    a.free(); b.free()


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
def test_cluster_tail_every_cluster_size(dev, cluster, monkeypatch):
    """The cluster tail kernel (DSMEM partial exchange, gather into CTA 0) at every cluster size, against the oracle:
    T3 with table and virtual eq (host transcript and device challenger) and a multi-spec tower."""
    import ceno_b200 as cb
    monkeypatch.setenv("CG_TAIL_CLUSTER", str(cluster))
    terms = [([1, 0], [0, 1, 2])]
    for k in (13, 17, 20):
        n = 1 << k
        w = rnd_point(4000 + k, k)
        a_h, b_h = orc.fill_ext(4100 + k, n), orc.fill_ext(4200 + k, n)
        a = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, a_h)
        b = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, b_h)
        eq = cb.build_eq_x_r_vec(dev, w)
        want = orc.sumcheck_prove([(orc.build_eq_x_r_vec(w), True, k), (a_h, True, k), (b_h, True, k)], terms, k, 3, transcript=orc.Transcript(b"ct"))
        for mles in ([eq, a, b], [cb.EqPolynomial(dev, w), a, b]):
            for devch in (False, True):
                got = cb.IOPProverState.prove(dev, mles, terms, k, 3, transcript=cb.StandInTranscript(b"ct"), device_challenger=devch)
                for g, x in zip(got, want):
                    assert eq_np(g, x), (cluster, k, devch)
        eq.free(); a.free(); b.free()
    _tower_case(dev, [15, 14], [14], False, seed=4321 + cluster)
    _tower_case(dev, [12] * 3, [11, 12], True, seed=99 + cluster)


# ------------------------------------------------ full-size, size-independent properties
def test_t3_k24_verifier_relations(dev):
    """BASELINE headline size (2^24 variables... points): too large for the oracle in seconds, so check
    the verifier's relations (ceno_recursion_v2/src/main/mod.rs:3513-3526): claim chaining through
    every round and the final product check with independently evaluated MLEs."""
    import ceno_b200 as cb
    from oracle import pyref as pr
    k = 24
    n = 1 << k
    w = orc.fill_ext(0xE9, k)
    eq = cb.build_eq_x_r_vec(dev, w)
    a_h, b_h = orc.fill_ext(0xC0FFEE ^ 1, n), orc.fill_ext(0xC0FFEE ^ 2, n)
    a = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, a_h)
    b = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, b_h)
    rounds, fin, chal = cb.IOPProverState.prove(dev, [eq, a, b], [([1, 0], [0, 1, 2])], k, 3, transcript=cb.StandInTranscript(b"k24"))
    rounds2, fin2, chal2 = cb.IOPProverState.prove(dev, [eq, a, b], [([1, 0], [0, 1, 2])], k, 3, transcript=cb.StandInTranscript(b"k24"),
                                                   device_challenger=True)
    assert eq_np(rounds, rounds2) and eq_np(fin, fin2) and eq_np(chal, chal2)
    # claim = sum_b eq(w,b) a(b) b(b) = (a*b)~(w): evaluate the pointwise product MLE at w on the device
    ab = cb.wit_infer_by_monomial_expr(dev, [a, b], [([1, 0], [0, 1])], k)
    claim = tuple(int(x) for x in ab.evaluate(w))
    for j in range(k):
        msg = [tuple(int(x) for x in e) for e in rounds[j]]
        e0 = pr.esub(claim, msg[0])
        claim = pr.lagrange_eval([e0] + msg, tuple(int(x) for x in chal[j]))
    pt = chal.reshape(-1)
    fe = [tuple(int(x) for x in f) for f in fin]
    assert pr.emul(fe[0], pr.emul(fe[1], fe[2])) == claim
    assert tuple(int(x) for x in orc.eq_eval(w, pt)) == fe[0]
    assert tuple(int(x) for x in a.evaluate(pt)) == fe[1]
    assert tuple(int(x) for x in orc.mle_evaluate(b_h, True, pt)) == fe[2]   # CPU check of one MLE (single fold chain)


def test_t3_k24_bit_exact_vs_oracle(dev):
    """The headline instance itself (bench.py's T3-24 seeds), every output bit against the oracle's proof:
    table eq and virtual eq, host transcript and device challenger."""
    import ceno_b200 as cb
    k = 24
    n = 1 << k
    w = orc.fill_ext(0xE9, k)
    a_h, b_h = orc.fill_ext(0xC0FFEE ^ 1, n), orc.fill_ext(0xC0FFEE ^ 2, n)
    a = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, a_h)
    b = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, b_h)
    eq = cb.build_eq_x_r_vec(dev, w)
    eqv = cb.EqPolynomial(dev, w)
    terms = [([1, 0], [0, 1, 2])]
    want = orc.sumcheck_prove_chunked([(orc.build_eq_x_r_vec(w), True, k), (a_h, True, k), (b_h, True, k)], terms, k, 3,
                                      orc.Transcript(b"k24"), consume=True)
    for mles in ([eqv, a, b], [eq, a, b]):
        for devch in (False, True):
            got = cb.IOPProverState.prove(dev, mles, terms, k, 3, transcript=cb.StandInTranscript(b"k24"), device_challenger=devch)
            for g, x in zip(got, want):
                assert eq_np(g, x), ("virtual" if mles[0] is eqv else "table", devch)
    eq.free(); a.free(); b.free()


@pytest.mark.parametrize("k", [3, 12, 16])
def test_t3_shape_with_an_unreferenced_mle(dev, k):
    """a zerocheck layer passes every column; an MLE that no term references must still be folded so that
    get_mle_flatten_final_evaluations returns its value (here: eq*A*B plus an unused ext and an unused base column)."""
    import ceno_b200 as cb
    n = 1 << k
    w = rnd_point(900 + k, k)
    hs = [orc.build_eq_x_r_vec(w), orc.fill_ext(901 + k, n), orc.fill_ext(902 + k, n), orc.fill_ext(903 + k, n), orc.fill_base(904 + k, n)]
    ext = [True, True, True, True, False]
    ms = [(cb.MultilinearExtension.from_evaluations_ext_vec if e else cb.MultilinearExtension.from_evaluations_vec)(dev, k, h) for h, e in zip(hs, ext)]
    terms = [([1, 0], [0, 1, 2])]
    want = orc.sumcheck_prove([(h, e, k) for h, e in zip(hs, ext)], terms, k, 3, transcript=orc.Transcript(b"unused"))
    for devch in (False, True):
        got = cb.IOPProverState.prove(dev, ms, terms, k, 3, transcript=cb.StandInTranscript(b"unused"), device_challenger=devch)
        for g, x in zip(got, want):
            assert eq_np(g, x)
    for m in ms:
        m.free()


def test_free_async_and_alignment_checks(dev):
    """cg_free_async defers reuse behind the stream; 256-bit-store outputs reject misaligned pointers."""
    import ctypes as C
    import ceno_b200 as cb
    lib = dev.lib
    buf = dev.alloc(1 << 20)
    ptr = buf.ptr
    assert lib.cg_free_async(dev.ctx, C.c_void_p(ptr), None) == 0
    buf.ptr = None
    dev.sync()
    again = dev.alloc(1 << 20)          # the block is reusable once the stream has drained
    again.free()
    out = dev.alloc(16 * 64 + 64)
    w = np.ascontiguousarray(rnd_point(5, 5))
    rc = lib.cg_build_eq(dev.ctx, w.ctypes.data_as(C.c_void_p), 5, C.c_void_p(out.ptr + 16), 0, 32, None)
    assert rc == 2, rc                  # CG_ERR_INVALID, not a sticky misaligned-address fault
    rc = lib.cg_build_eq(dev.ctx, w.ctypes.data_as(C.c_void_p), 5, C.c_void_p(out.ptr), 0, 32, None)
    assert rc == 0
    dev.sync()
    out.free()


# ------------------------------------------------------------------------------ multi-GPU
def _dist_gpu_worker(rank, world, k, port, q):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ceno_b200 as cb
    from ceno_b200 import dist as cdist
    g = world.bit_length() - 1
    kl, nl = k - g, 1 << (k - g)
    dev = cb.Device(rank)
    w = orc.fill_ext(0xE9, k)
    eq, a, b = orc.build_eq_x_r_vec(w), orc.fill_ext(1, 1 << k), orc.fill_ext(2, 1 << k)
    sl = slice(2 * rank * nl, 2 * (rank + 1) * nl)
    mles = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, kl, x[sl]) for x in (eq, a, b)]
    terms = [([1, 0], [0, 1, 2])]
    lp = cdist.GpuLocalProver(dev, mles, terms, kl, 3)
    out = cdist.sharded_prove(lp, kl, g, 3, cb.StandInTranscript(b"dist"), cdist.TorchExchange(torch.device("cuda", rank)),
                              lambda arrays: cdist.GpuLocalProver(dev, [cb.MultilinearExtension.from_evaluations_ext_vec(dev, g, x) for x in arrays], terms, g, 3))
    want = orc.sumcheck_prove([(eq, True, k), (a, True, k), (b, True, k)], terms, k, 3, transcript=orc.Transcript(b"dist"))
    ok = all(np.array_equal(x, y) for x, y in zip(out, want))
    # in-kernel NVLink mailbox exchange (cg_comm): host transcript, device challenger, and a generic-kernel run

    def xchg(blob):
        outs = [None] * world
        dist.all_gather_object(outs, blob)
        return outs
    comm = cb.Comm(dev, rank, world, xchg, barrier=dist.barrier)
    for dc, flags in ((False, 0), (True, 0), (False, cb.IOPProverState.FORCE_GENERIC), (True, cb.IOPProverState.NO_TAIL)):
        got = cb.prove_sharded(dev, comm, mles, terms, k, 3, cb.StandInTranscript(b"dist"), flags=flags, device_challenger=dc)
        ok = ok and all(np.array_equal(x, y) for x, y in zip(got, want))
    comm.close()
    q.put((rank, ok))
    dist.destroy_process_group()


def _dist_gpu_worker_veq(rank, world, k, port, q):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ceno_b200 as cb
    g = world.bit_length() - 1
    kl, nl = k - g, 1 << (k - g)
    dev = cb.Device(rank)
    w = orc.fill_ext(0xE9, k)
    a, b = orc.fill_ext(1, 1 << k), orc.fill_ext(2, 1 << k)
    sl = slice(2 * rank * nl, 2 * (rank + 1) * nl)
    mles = [cb.EqPolynomial(dev, w, num_vars=kl)] + [cb.MultilinearExtension.from_evaluations_ext_vec(dev, kl, x[sl]) for x in (a, b)]
    terms = [([1, 0], [0, 1, 2])]
    want = orc.sumcheck_prove([(orc.build_eq_x_r_vec(w), True, k), (a, True, k), (b, True, k)], terms, k, 3, transcript=orc.Transcript(b"dv"))

    def xchg(blob):
        outs = [None] * world
        dist.all_gather_object(outs, blob)
        return outs
    comm = cb.Comm(dev, rank, world, xchg, barrier=dist.barrier)
    ok = True
    for dc, flags in ((False, 0), (True, 0), (True, cb.IOPProverState.NO_MID), (False, cb.IOPProverState.FORCE_GENERIC)):
        got = cb.prove_sharded(dev, comm, mles, terms, k, 3, cb.StandInTranscript(b"dv"), flags=flags, device_challenger=dc)
        ok = ok and all(np.array_equal(x, y) for x, y in zip(got, want))
    comm.close()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_virtual_eq_multi_gpu_bit_exact():
    """The virtual eq under hypercube sharding: every rank passes the GLOBAL point, the library derives the
    rank factor eq(w_top, rank); split-eq rounds when the local slice has >= 20 variables, scaled table otherwise."""
    import torch
    import torch.multiprocessing as mp
    world = 1
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_gpu_worker_veq, args=(r, world, 22, 29653, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def _dist_commit_worker(rank, world, port, q):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ceno_b200 as cb
    from ceno_b200 import api
    from ceno_b200 import dist as cdist
    from oracle import basefold as bf
    dev = cb.Device(rank)
    p2 = orc.p2_params(seed=5)
    api.poseidon2_set_params(dev, [[int(p2.ext_rc[r][i]) for i in range(8)] for r in range(8)], [int(x) for x in p2.int_rc], [int(x) for x in p2.diag], 0)
    width, log_n, rate_log = 8, 10, 1
    cols = [orc.fill_base(8800 + c, 1 << log_n) for c in range(width)]
    rows_local = (1 << log_n) // world
    loc = np.stack([c[rank * rows_local:(rank + 1) * rows_local] for c in cols]).view(np.int64)
    t = torch.from_numpy(loc.copy()).cuda()
    root, code_local, tree = cdist.commit_sharded(dev, t, log_n, rate_log, rank, world, torch, dist)
    want = bf.commit(p2, bf.Params(rate_log=rate_log), cols, log_n)
    ok = [int(x) for x in root] == want["root"]
    # the local code rows are the oracle's codeword rows of this rank's range (bit-reversed order)
    hl = (1 << (log_n + rate_log)) // world
    got_rows = code_local.cpu().numpy().view(np.uint64).T
    ok = ok and [[int(v) for v in r] for r in got_rows] == want["rows"][rank * hl:(rank + 1) * hl]
    tree.free()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_commit_multi_gpu_bit_exact():
    """BASELINE config #5 sharding: row-sharded witness -> all-to-all -> per-column RS-encode -> all-to-all -> row-range leaf hash
    -> all-gather of subtree roots; the root equals the single-device / oracle commitment root."""
    import torch
    import torch.multiprocessing as mp
    world = 1
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_commit_worker, args=(r, world, 29667, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def _dist_tower_worker(rank, world, port, q):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ceno_b200 as cb
    dev = cb.Device(rank)

    def xchg(blob):
        outs = [None] * world
        dist.all_gather_object(outs, blob)
        return outs
    comm = cb.Comm(dev, rank, world, xchg, barrier=dist.barrier)
    comm.create_arena(256 << 20, xchg)
    ok = True
    # (a) materialised leaves: 2 product specs of different depth + 1 logup spec with explicit numerators
    prod_nvs, logup_nvs = [17, 16], [15]
    o_prod, o_lk, specs, keep = [], [], [], []
    for i, nv in enumerate(prod_nvs):
        half = 1 << (nv - 1)
        f1, f2 = orc.fill_ext(6100 + 2 * i, half), orc.fill_ext(6101 + 2 * i, half)
        o_prod.append((orc.infer_tower_product_witness(nv, f1, f2)[0], nv))
        loc = half // world
        ms = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, (nv - 1) - (world.bit_length() - 1), x[2 * rank * loc:2 * (rank + 1) * loc]) for x in (f1, f2)]
        keep += ms
        specs.append(cb.TowerProverSpec(ms, nv, False))
    for i, nv in enumerate(logup_nvs):
        n = 1 << nv
        arrs = [orc.fill_ext(6200 + 4 * i + z, n) for z in range(4)]
        o_lk.append((orc.infer_tower_logup_witness(nv, *arrs)[0], nv + 1))
        loc = n // world
        ms = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, nv - (world.bit_length() - 1), x[2 * rank * loc:2 * (rank + 1) * loc]) for x in arrs]
        keep += ms
        specs.append(cb.TowerProverSpec(ms, nv, True))
    want_proof, want_point = orc.tower_create_proof(o_prod, o_lk, orc.Transcript(b"shtower"))
    for veq_min in ("10", "64"):   # sliced layers with split-eq rounds (eq handed over as its point), then with eq tables
        os.environ["CG_TOWER_VEQ_MIN_NV"] = veq_min
        tw = cb.TowerProver.sharded(dev, comm, specs)
        got_proof, got_point = tw.create_proof(cb.StandInTranscript(b"shtower"))
        ok = ok and np.array_equal(got_proof, want_proof) and np.array_equal(got_point, want_point)
        tw.close()
    os.environ["CG_TOWER_VEQ_MIN_NV"] = "10"
    # (b) virtual leaves: every rank holds its rows of both fan-in blocks of 5 lookup records (a chip of rows / world rows)
    log_n, n_rec, alpha = 14, 5, [12345, 678]
    n = 1 << log_n
    recs = [orc.fill_ext(6300 + i, n) for i in range(n_rec)]
    want = orc.interleaving_mles_to_mles([(r, True) for r in recs], n, 2, alpha)
    nv = (want[0].size // 2).bit_length() - 1
    o_lk2 = [(orc.infer_tower_logup_witness(nv, None, None, want[0], want[1])[0], nv + 1)]
    want_proof2, want_point2 = orc.tower_create_proof([], o_lk2, orc.Transcript(b"shvirt"))
    hl = (n // 2) // world                      # rows of one fan-in block per rank
    loc_recs = []
    for r in recs:
        rr = r.reshape(-1, 2)
        loc_recs.append(np.concatenate([rr[rank * hl:(rank + 1) * hl], rr[n // 2 + rank * hl:n // 2 + (rank + 1) * hl]]).reshape(-1))
    ms = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, log_n - (world.bit_length() - 1), x) for x in loc_recs]
    tw2 = cb.TowerProver.from_records(dev, [cb.VirtualTowerSpec(ms, n // world, alpha, True)], comm=comm)
    got_proof2, got_point2 = tw2.create_proof(cb.StandInTranscript(b"shvirt"))
    ok = ok and np.array_equal(got_proof2, want_proof2) and np.array_equal(got_point2, want_point2)
    tw2.close()
    dist.barrier()
    comm.close()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_tower_multi_gpu_bit_exact():
    """cfg-4: the tower built and proven over rank slices (layer kernels store into the partner ranks' buffers over NVLink,
    big layers' sumchecks sharded, small layers replicated) gives the oracle's single-device proof — with materialised leaves
    and with virtual leaves."""
    import torch
    import torch.multiprocessing as mp
    world = 1
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_tower_worker, args=(r, world, 29681, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def test_sharded_sumcheck_multi_gpu_bit_exact():
    import torch
    import torch.multiprocessing as mp
    world = 1
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_gpu_worker, args=(r, world, 14, 29641, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


# ------------------------------------------------------------------- Poseidon2 / Merkle (a9)
@pytest.mark.parametrize("variant", [0, 1])
def test_poseidon2_and_merkle_commit(dev, variant):
    """Constants are caller-supplied placeholders (parity unpinned upstream); the kernels must match
    the oracle bit for bit on the same constants."""
    import ceno_b200 as cb
    from ceno_b200 import api
    prm = orc.p2_params(seed=5, mds_variant=variant)
    api.poseidon2_set_params(dev, [[int(prm.ext_rc[r][i]) for i in range(8)] for r in range(8)], [int(x) for x in prm.int_rc],
                             [int(x) for x in prm.diag], variant)
    st = orc.fill_base(31, 8 * 300)
    st[:8] = 0
    st[8:16] = np.uint64(P - 1)
    st[16:24] = np.uint64(2**64 - 1)     # non-canonical input lanes
    want = np.concatenate([orc.poseidon2_permute(prm, st[8 * i:8 * i + 8]) for i in range(300)])
    assert eq_np(api.poseidon2_permute(dev, st), want)
    for width, height in [(1, 1), (3, 2), (8, 64), (13, 1024), (64, 4096)]:
        m = orc.fill_base(500 + width, width * height)                   # row-major host matrix
        tree_w, root_w = orc.merkle_commit(prm, m, width, height)
        rm = dev.to_device(m)
        cm = dev.to_device(np.ascontiguousarray(m.reshape(height, width).T).reshape(-1))
        for buf, col_major in ((rm, False), (cm, True)):
            tree, root = api.merkle_commit(dev, buf, width, height, col_major=col_major)
            assert eq_np(root, root_w)
            assert eq_np(tree.to_host(), tree_w)
            tree.free()
        rm.free(); cm.free()


# ------------------------------------------------------------------- rotation pre-passes (f-3)
@pytest.mark.parametrize("log2", [5, 6])
def test_rotation_prepasses(dev, log2):
    import ceno_b200 as cb
    from ceno_b200 import api
    k = 12
    base = orc.fill_base(3 + log2, 1 << k)
    m = cb.MultilinearExtension.from_evaluations_vec(dev, k, base)
    assert eq_np(api.rotation_next_base_mle(dev, m, log2).evaluations(), orc.rotation_next_base_mle(base, log2))
    pt = rnd_point(9, k)
    eq = cb.build_eq_x_r_vec(dev, pt)
    for sub in (1, 23, (1 << log2) - 1, 1 << log2):
        assert eq_np(api.rotation_selector(dev, eq, sub, log2).evaluations(), orc.rotation_selector(orc.build_eq_x_r_vec(pt), sub, log2))
    with pytest.raises(cb.CenoB200Error):
        api.rotation_next_base_mle(dev, m, 4)      # BooleanHypercube::new asserts 5 or 6


# ------------------------------------------------------------------ f-4: concurrent chip proving on lanes
def test_chip_scheduler_concurrent_lanes_bit_exact(dev):
    """ChipScheduler::execute shape (ceno_zkvm/src/scheme/scheduler.rs:205-250): several chip-sized sumchecks of
    different sizes proved concurrently on 4 lanes (one stream + one OS thread each), every proof bit-exact and the
    library re-entrant per (ctx, stream)."""
    import ceno_b200 as cb
    ks = [16, 12, 18, 14, 17, 13, 15, 11]
    terms = [([1, 0], [0, 1, 2])]
    inputs = {i: t3_inputs(k, seed=40 + i) for i, k in enumerate(ks)}
    want = {i: orc.sumcheck_prove([(x, True, ks[i]) for x in inputs[i]], terms, ks[i], 3, transcript=orc.Transcript(b"chip%d" % i))
            for i in range(len(ks))}
    tasks = [cb.ChipTask(i, 3 * 16 * (1 << k) * 2, payload=i, circuit_name="chip%d" % i) for i, k in enumerate(ks)]

    def prove(task, lane, stream):
        i = task.payload
        assert stream                                               # a real, non-default lane stream
        mles = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, ks[i], x) for x in inputs[i]]
        got = cb.IOPProverState.prove(dev, mles, terms, ks[i], 3, transcript=cb.StandInTranscript(b"chip%d" % i), stream=stream)
        for m in mles:
            m.free()
        return got

    for lanes in (1, 4, 8):
        out, tel = cb.ChipScheduler(dev).execute(tasks, prove, lanes=lanes)
        for i in range(len(ks)):
            for g, w in zip(out[i], want[i]):
                assert eq_np(g, w), (lanes, i)
        assert {t["lane_id"] for t in tel} <= set(range(lanes))
        assert [t["task_id"] for t in tel] == list(range(len(ks)))
    # caller-owned lane stream (cg_stream_create): same result
    st = cb.Stream(dev)
    mles = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, ks[0], x) for x in inputs[0]]
    got = cb.IOPProverState.prove(dev, mles, terms, ks[0], 3, transcript=cb.StandInTranscript(b"chip0"), stream=st.handle)
    st.sync()
    for g, w in zip(got, want[0]):
        assert eq_np(g, w)
    st.close()


# ------------------------------------------------------------------ a9 / f-2: NTT, RS-encode, Basefold-style commit
@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 9, 12, 13, 16, 20, 21, 22])
def test_ntt_bit_exact(dev, log_n):
    from ceno_b200 import api
    n = 1 << log_n
    n_cols = 3 if log_n <= 16 else 1
    x = orc.fill_base(1000 + log_n, n_cols * n)
    if n >= 4:
        x[1] = np.uint64(2**64 - 1)                               # non-canonical input
        x[2] = np.uint64(P)
    for bitrev in (True, False):
        want = orc.ntt(x, log_n, n_cols, bitrev=bitrev)
        buf = dev.to_device(x)
        api.ntt(dev, buf, log_n, n_cols, bitrev=bitrev)
        assert eq_np(buf.to_host(), want), ("forward", bitrev)
        api.ntt(dev, buf, log_n, n_cols, inverse=True, bitrev=bitrev)
        assert eq_np(buf.to_host(), orc.ntt(want, log_n, n_cols, inverse=True, bitrev=bitrev)), ("inverse", bitrev)
        buf.free()


def test_ntt_ext_and_column_stride(dev):
    from ceno_b200 import api
    log_n, n_cols = 10, 4
    n, stride = 1 << log_n, (1 << log_n) + 24
    x = orc.fill_ext(2024, n_cols * stride)                      # ext elements, columns `stride` apart
    buf = dev.to_device(x)
    api.ntt(dev, buf, log_n, n_cols, col_stride=stride, bitrev=True, ext=True)
    got = buf.to_host().reshape(n_cols, stride, 2)
    xr = x.reshape(n_cols, stride, 2)
    for c in range(n_cols):
        for limb in range(2):
            assert eq_np(got[c, :n, limb], orc.ntt(np.ascontiguousarray(xr[c, :n, limb]), log_n, bitrev=True))
        assert eq_np(got[c, n:], xr[c, n:])                       # the gap between columns is untouched
    buf.free()


@pytest.mark.parametrize("width,log_n,rate_log", [(1, 0, 0), (1, 0, 1), (5, 3, 1), (7, 10, 1), (3, 12, 2), (2, 17, 1), (64, 12, 3)])
def test_rs_encode_bit_exact(dev, width, log_n, rate_log):
    from ceno_b200 import api
    msg = orc.fill_base(3000 + log_n, width << log_n)
    mb = dev.to_device(msg)
    for bitrev in (True, False):
        code = api.rs_encode(dev, mb, width, log_n, rate_log, bitrev=bitrev)
        assert eq_np(code.to_host(), orc.rs_encode(msg, width, log_n, rate_log, bitrev=bitrev))
        code.free()
    mb.free()


def test_basefold_style_commit_flow(dev):
    """RS-encode + row-wise Merkle hash of the codeword matrix (commit_traces shape), against the oracle composition."""
    from ceno_b200 import api
    prm = orc.p2_params(seed=9)
    api.poseidon2_set_params(dev, [[int(prm.ext_rc[r][i]) for i in range(8)] for r in range(8)], [int(x) for x in prm.int_rc],
                             [int(x) for x in prm.diag], 0)
    width, log_n, rate_log = 6, 10, 1
    msg = orc.fill_base(4242, width << log_n)
    mb = dev.to_device(msg)
    code, tree, root = api.basefold_style_commit(dev, mb, width, log_n, rate_log)
    h = 1 << (log_n + rate_log)
    code_w = orc.rs_encode(msg, width, log_n, rate_log, bitrev=True)
    rows = np.ascontiguousarray(code_w.reshape(width, h).T).reshape(-1)          # oracle hashes a row-major matrix
    tree_w, root_w = orc.merkle_commit(prm, rows, width, h)
    assert eq_np(code.to_host(), code_w) and eq_np(root, root_w) and eq_np(tree.to_host(), tree_w)
    for b in (mb, code, tree):
        b.free()


def test_ntt_large_roundtrip_and_linearity(dev):
    """BASELINE config #5 scale (2^26 elements: 4 columns x 2^24): inverse(forward(x)) == x bit for bit, and the
    transform of a sum is the sum of transforms on a sampled column."""
    from ceno_b200 import api
    log_n, n_cols = 24, 4
    x = orc.fill_base(555, n_cols << log_n)
    buf = dev.to_device(x)
    api.ntt(dev, buf, log_n, n_cols, bitrev=True)
    fx = buf.to_host()
    assert eq_np(fx[:1 << log_n], orc.ntt(x[:1 << log_n], log_n, bitrev=True))   # one full column against the oracle
    api.ntt(dev, buf, log_n, n_cols, inverse=True, bitrev=True)
    assert eq_np(buf.to_host(), x)
    buf.free()


# ------------------------------------------------------------------ f-3: EC-sum Quark prover
@pytest.mark.parametrize("n,num_instances", [(1, 2), (3, 5), (6, 64), (10, 777), (14, 1 << 14)])
def test_ecc_quark_prover_bit_exact(dev, n, num_instances):
    """EccQuarkProver::prove_ec_sum_quark shape (cpu/mod.rs:72-316): device pre-passes (selectors, even/odd split, views)
    + the degree-3 zerocheck over 260 monomial terms, against the oracle restatement; host term expansion against the
    oracle's independent one."""
    import ceno_b200 as cb
    from ceno_b200 import api
    if n <= 6:
        xs, ys, invs = orc.ecc_quark_make_witness(800 + n, n, num_instances)
    else:                                                   # parity does not need a satisfying witness
        xs, ys, invs = ([orc.fill_base(900 + 10 * g + i, 2 << n) for i in range(7)] for g in range(3))
    want = orc.ecc_quark_create_proof(num_instances, xs, ys, invs, orc.Transcript(b"ecc"))
    # pre-passes on their own
    sels = api.ecc_quark_selectors(dev, want["out_rt"], num_instances)
    for m, w in zip(sels, orc.ecc_quark_selectors(want["out_rt"], num_instances)):
        assert eq_np(m.evaluations(), w)
        m.free()
    dx = [cb.MultilinearExtension.from_evaluations_vec(dev, n + 1, v) for v in xs]
    dy = [cb.MultilinearExtension.from_evaluations_vec(dev, n + 1, v) for v in ys]
    di = [cb.MultilinearExtension.from_evaluations_vec(dev, n + 1, v) for v in invs]
    ev, od = api.split_even_odd(dev, dx)
    for i in range(7):
        assert eq_np(ev[i].evaluations(), xs[i][0::2]) and eq_np(od[i].evaluations(), xs[i][1::2])
    for m in ev + od:
        m.free()
    got = cb.EccQuarkProver.create_ecc_proof(dev, num_instances, dx, dy, di, cb.StandInTranscript(b"ecc"))
    assert sorted((tuple(c), tuple(i)) for c, i in api.EccQuarkProver.build_terms(
        orc_alpha_pows(b"ecc", n), want["sum"][0], want["sum"][1])) == sorted((tuple(c), tuple(i)) for c, i in want["terms"])
    for key in ("zerocheck_proof", "evals", "rt"):
        assert eq_np(got[key], want[key]), key
    assert got["sum"] == want["sum"]
    if n <= 6:                                              # a zerocheck: first-round claim p(0) + p(1) = 0
        assert eq_np(got["zerocheck_proof"], want["zerocheck_proof"])
    for m in dx + dy + di:
        m.free()


def orc_alpha_pows(label, n):
    t = orc.Transcript(label)
    for _ in range(n):
        t.sample(b"ecc")
    a = t.sample(b"ecc_alpha")
    a, cur, out = (int(a[0]), int(a[1])), (1, 0), []
    for _ in range(49):
        out.append(cur)
        cur = ((cur[0] * a[0] + 7 * cur[1] * a[1]) % P, (cur[0] * a[1] + cur[1] * a[0]) % P)
    return np.array(out, dtype=np.uint64)


# ------------------------------------------------------------------ chip flow (create_chip_proof shape) on scheduler lanes
def _chip_oracle(chip, alpha=(12345, 678)):
    """The same composition on the CPU oracle (records by direct evaluation, towers, main zerocheck)."""
    k, ninst, n = chip.num_vars, chip.num_instances, 1 << chip.num_vars
    wit = chip.witness().reshape(chip.n_wit, n)
    mles = [(np.ascontiguousarray(wit[c]), False, k) for c in range(chip.n_wit)]
    from oracle import pyref as pr

    def record(e):   # ext vector of the monomial expression, row by row (vectorised big-int is too slow: use the oracle's generic evaluator)
        out = np.zeros((n, 2), np.uint64)
        for row in range(n):
            acc = (0, 0)
            for c, ids in e:
                t = (int(c[0]), int(c[1]))
                for i in ids:
                    w = int(wit[i, row])
                    t = (t[0] * w % P, t[1] * w % P)
                acc = ((acc[0] + t[0]) % P, (acc[1] + t[1]) % P)
            out[row] = acc
        return out.reshape(-1)

    groups = [[record(e) for e in ex] for ex in (chip.read_exprs, chip.write_exprs, chip.lk_exprs)]
    o_prod, o_lk = [], []
    for recs in groups[:2]:
        leaves = orc.interleaving_mles_to_mles([(c, True) for c in recs], ninst, 2, [1, 0])
        nv = (leaves[0].size // 2).bit_length()
        o_prod.append((orc.infer_tower_product_witness(nv, leaves[0], leaves[1])[0], nv))
    leaves = orc.interleaving_mles_to_mles([(c, True) for c in groups[2]], ninst, 2, list(alpha))
    nv = (leaves[0].size // 2).bit_length() - 1
    o_lk.append((orc.infer_tower_logup_witness(nv, None, None, leaves[0], leaves[1])[0], nv + 1))
    t = orc.Transcript(chip.name.encode())
    proof, point = orc.tower_create_proof(o_prod, o_lk, t)
    rt = point[:2 * k]
    sel = orc.selector_compute(1, rt, 0, ninst)
    a = t.sample(b"combine subset evals")
    a, cur, pows = (int(a[0]), int(a[1])), (1, 0), []
    for _ in range(sum(len(g) for g in groups)):
        pows.append(cur)
        cur = ((cur[0] * a[0] + 7 * cur[1] * a[1]) % P, (cur[0] * a[1] + cur[1] * a[0]) % P)
    acc, ai = {}, 0
    for ex in (chip.read_exprs, chip.write_exprs, chip.lk_exprs):
        for e in ex:
            for c, ids in e:
                key = tuple(sorted([0] + [1 + i for i in ids]))
                cc = ((int(c[0]) * pows[ai][0] + 7 * int(c[1]) * pows[ai][1]) % P, (int(c[0]) * pows[ai][1] + int(c[1]) * pows[ai][0]) % P)
                v = acc.get(key, (0, 0))
                acc[key] = ((v[0] + cc[0]) % P, (v[1] + cc[1]) % P)
            ai += 1
    terms = [([c[0], c[1]], list(kk)) for kk, c in sorted(acc.items()) if c != (0, 0)]
    rounds, evals, pt = orc.sumcheck_prove([(sel, True, k)] + mles, terms, k, 3, transcript=t)
    return {"tower_proof": proof, "tower_point": point, "main_proof": rounds, "main_evals": evals, "main_point": pt}


def test_chip_proof_flow_on_lanes_bit_exact(dev):
    """Six synthetic chips proved through ChipScheduler lanes (records -> towers -> tower proof -> main zerocheck ->
    commitment): lane execution equals sequential execution, and the small chips equal the CPU oracle composition."""
    import ceno_b200 as cb
    from ceno_b200 import api, chip as chipmod
    prm = orc.p2_params(seed=3)
    api.poseidon2_set_params(dev, [[int(prm.ext_rc[r][i]) for i in range(8)] for r in range(8)], [int(x) for x in prm.int_rc],
                             [int(x) for x in prm.diag], 0)
    chips = [chipmod.SyntheticChip(s, k, ni, n_wit=8) for s, (k, ni) in enumerate([(9, 500), (7, 128), (12, 4000), (8, 200), (11, 2048), (6, 33)])]
    wits = [chipmod.upload_witness(dev, c) for c in chips]
    tasks = [cb.ChipTask(i, c.estimated_memory_bytes(), payload=i, circuit_name=c.name) for i, c in enumerate(chips)]

    def prove(task, lane, stream):
        i = task.payload
        return chipmod.create_chip_proof(dev, chips[i], wits[i][1], cb.StandInTranscript(chips[i].name.encode()), stream=stream, commit_matrix=wits[i][0])

    seq, _ = cb.ChipScheduler(dev).execute(tasks, prove, lanes=1)
    par, tel = cb.ChipScheduler(dev).execute(tasks, prove, lanes=4)
    assert len({t["lane_id"] for t in tel}) > 1
    for a, b in zip(seq, par):
        assert a.keys() == b.keys()
        for key in a:
            assert eq_np(a[key], b[key]), key
    for i in (1, 3, 5):                                   # small chips against the oracle composition
        want = _chip_oracle(chips[i])
        for key, w in want.items():
            assert eq_np(par[i][key], w), (i, key)
    for buf, _ in wits:
        buf.free()


def test_concurrent_towers_of_different_sizes_on_lanes(dev):
    """Regression for a lane race: tower proofs whose persistent tail kernels need DIFFERENT amounts of dynamic shared
    memory, proved concurrently on 8 lanes several times over, must equal the sequential proofs (the shared-memory opt-in
    is process-global function state and is now raised once per context)."""
    import ceno_b200 as cb
    shapes = [([4], [3]), ([9, 9], [8]), ([6], []), ([11, 10], [10]), ([], [7]), ([12], [11]), ([8, 3], [5]), ([10], [9])]
    towers = []
    for i, (pn, ln) in enumerate(shapes):
        specs, s = [], 5000 + 50 * i
        for nv in pn:
            specs.append(cb.TowerProverSpec([cb.MultilinearExtension.from_evaluations_ext_vec(dev, nv - 1, orc.fill_ext(s + z, 1 << (nv - 1))) for z in range(2)], nv, False))
            s += 2
        for nv in ln:
            specs.append(cb.TowerProverSpec([None, None] + [cb.MultilinearExtension.from_evaluations_ext_vec(dev, nv, orc.fill_ext(s + z, 1 << nv)) for z in range(2)], nv, True))
            s += 2
        towers.append(specs)
    tasks = [cb.ChipTask(i, 1 << 20, payload=i) for i in range(len(shapes))]

    def prove(task, lane, stream):
        tw = cb.TowerProver(dev, towers[task.payload], stream=stream)
        out = tw.create_proof(cb.StandInTranscript(b"tw%d" % task.payload))
        tw.close()
        return out

    want, _ = cb.ChipScheduler(dev).execute(tasks, prove, lanes=1)
    for rep in range(4):
        got, tel = cb.ChipScheduler(dev).execute(tasks, prove, lanes=8)
        for i, (g, w) in enumerate(zip(got, want)):
            assert eq_np(g[0], w[0]) and eq_np(g[1], w[1]), (rep, i)
    assert len({t["lane_id"] for t in tel}) > 1


# ------------------------------------------------------------------ gkr_iop layer / circuit API mirror (a5, a10)
def test_gkr_circuit_two_layers_bit_exact(dev):
    """GKRCircuit::prove (gkr_iop/src/gkr.rs:70-117) over a zerocheck layer with two selector groups (Prefix + Whole,
    ZerocheckLayerProver::prove cpu/mod.rs:99-239) followed by a linear layer (cpu/mod.rs:44-67): the device mirror
    against the same call / transcript order composed from the oracle's primitives."""
    import ceno_b200 as cb
    from ceno_b200 import gkr
    from ceno_b200.expr import Poly
    from oracle import pyref as pr
    k, ninst = 8, 200
    n = 1 << k
    W = [orc.fill_base(7000 + i, n) for i in range(3)]
    F0 = orc.fill_base(7010, n)
    V = [orc.fill_ext(7020 + i, n) for i in range(2)]
    w0, w1, w2, f0 = (Poly.var(i) for i in range(4))
    exprs = [w0 * w1 + f0, w2 * w2 * w0, w1 - w2, w0 + w1 * 3]
    groups = [gkr.OutGroup(cb.SelectorType.PREFIX, 0, [0, 1, 2]), gkr.OutGroup(cb.SelectorType.WHOLE, 1, [3])]
    l0 = gkr.Layer("out", gkr.ZEROCHECK, 3, 1, 2, exprs, groups, in_eval_positions=[4, 5, 6, 7, 8, 9])
    l1 = gkr.Layer("inner", gkr.LINEAR, 2, 0, 0, [], [gkr.OutGroup(None, 0, [4])], in_eval_positions=[10, 11])
    circuit = gkr.GKRCircuit([l0, l1], 12, [10, 11, 7])
    p0, p1 = rnd_point(71, k), rnd_point(72, k)
    zero = np.zeros(2, np.uint64)
    out_evals = [(p0, zero), (p0, zero), (p0, zero), (p1, zero)]
    challenges = orc.fill_ext(7030, 2)
    ctxs = [gkr.SelectorContext(0, ninst, k), gkr.SelectorContext(0, ninst, k)]
    mk_b = lambda v: cb.MultilinearExtension.from_evaluations_vec(dev, k, v)
    wit0 = [mk_b(v) for v in W] + [mk_b(F0), None, None]
    wit1 = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, v) for v in V]
    got = circuit.prove(dev, [wit0, wit1], out_evals, [], challenges, cb.StandInTranscript(b"gkr"), ctxs)

    # ---- the same flow from the oracle's primitives
    t = orc.Transcript(b"gkr")
    a = t.sample(b"combine subset evals")
    a, cur, al = (int(a[0]), int(a[1])), (1, 0), []
    for _ in range(len(exprs)):
        al.append(cur)
        cur = ((cur[0] * a[0] + 7 * cur[1] * a[1]) % P, (cur[0] * a[1] + cur[1] * a[0]) % P)
    sel0 = orc.selector_compute(1, p0, 0, ninst)
    sel1 = orc.selector_compute(0, p1)
    terms = l0.main_sumcheck_terms(np.array(al, dtype=np.uint64))
    mles = [(v, False, k) for v in W] + [(F0, False, k), (sel0, True, k), (sel1, True, k)]
    # the monomial table equals the layer polynomial sum_g sel_g * sum_j alpha_j expr_j at sampled points (direct evaluation)
    mp = [[(int(x), 0) for x in v] for v in W + [F0]] + [[(int(x), int(y)) for x, y in s.reshape(-1, 2)] for s in (sel0, sel1)]
    tl = [((int(c[0]), int(c[1])), ids) for c, ids in terms]
    for x in (0, 1, 57, ninst - 1, ninst, n - 1):
        v = [m[x][0] for m in mp[:4]]
        ev = [(v[0] * v[1] + v[3]) % P, v[2] * v[2] * v[0] % P, (v[1] - v[2]) % P, (v[0] + 3 * v[1]) % P]
        g0 = pr.ZERO
        for j in range(3):
            g0 = pr.eadd(g0, pr.emul(al[j], (ev[j], 0)))
        want_x = pr.eadd(pr.emul(mp[4][x], g0), pr.emul(mp[5][x], pr.emul(al[3], (ev[3], 0))))
        assert pr.poly_eval(mp, tl, x) == want_x
    rounds, evals, point = orc.sumcheck_prove(mles, terms, k, 4, transcript=t)
    t.append_ext(evals.reshape(-1))
    lin = np.array([orc.mle_evaluate(v, True, point.reshape(-1)) for v in V], dtype=np.uint64)
    t.append_ext(lin.reshape(-1))
    assert eq_np(got["gkr_proof"][0].proof, rounds) and eq_np(got["gkr_proof"][0].evals, evals)
    assert eq_np(got["rt"][0], point) and eq_np(got["rt"][1], point)          # the linear layer opens at the zerocheck's point
    assert got["gkr_proof"][1].proof.size == 0 and eq_np(got["gkr_proof"][1].evals, lin)
    vals = [o[0] for o in got["opening_evaluations"]]
    assert eq_np(vals[0], lin[0]) and eq_np(vals[1], lin[1]) and eq_np(vals[2], evals[3])
    for m in wit0[:4] + wit1:
        m.free()


# ------------------------------------------------------------------ randomized sweep over term tables / shapes / code paths
@pytest.mark.parametrize("seed", list(range(12)))
def test_random_sumcheck_instances_bit_exact(dev, seed):
    """Random monomial term tables (repeated factors, constants, lower-degree terms), base / ext / occupied-prefix MLEs,
    1..11 variables, degree 1..5, every kernel family (grouped plan, term-by-term, unfused, tower-shaped when it applies,
    host transcript or device challenger): bit-exact against the oracle."""
    import ceno_b200 as cb
    rng = random.Random(9000 + seed)
    for _ in range(4):
        k = rng.randint(1, 11)
        n = 1 << k
        degree = rng.randint(1, 5)
        m = rng.randint(1, 6)
        host, mles = [], []
        for i in range(m):
            is_ext = rng.random() < 0.5
            ln = n if rng.random() < 0.7 else rng.randint(1, n)                    # occupied prefix
            d = orc.fill_ext(rng.randrange(1 << 30), ln) if is_ext else orc.fill_base(rng.randrange(1 << 30), ln)
            full = np.zeros((2 if is_ext else 1) * n, np.uint64)
            full[:d.size] = d
            host.append((full, is_ext, k))
            mles.append((cb.MultilinearExtension.from_evaluations_ext_vec if is_ext else cb.MultilinearExtension.from_evaluations_vec)(dev, k, d))
        terms = []
        for _t in range(rng.randint(1, 9)):
            nf = rng.randint(0 if len(terms) else 1, degree)
            coeff = [rng.randrange(P), rng.choice([0, rng.randrange(P)])]
            terms.append((coeff, [rng.randrange(m) for _ in range(nf)]))
        label = b"rnd%d" % seed
        want = orc.sumcheck_prove(host, terms, k, degree, transcript=orc.Transcript(label))
        flags = rng.choice([0, 0, cb.IOPProverState.NO_PLAN, cb.IOPProverState.NO_FUSE, cb.IOPProverState.FORCE_GENERIC,
                            cb.IOPProverState.NO_TAIL, cb.IOPProverState.NO_MID])
        dc = rng.random() < 0.5
        got = cb.IOPProverState.prove(dev, mles, terms, k, degree, transcript=cb.StandInTranscript(label), device_challenger=dc, flags=flags)
        for g, w in zip(got, want):
            assert eq_np(g, w), (seed, k, degree, m, terms, flags, dc)
        for mm in mles:
            mm.free()


@pytest.mark.parametrize("seed", list(range(6)))
def test_random_tower_shapes_bit_exact(dev, seed):
    """Random tower shapes: 0..5 product specs and 0..3 logup specs of independent depths 1..11, numerators present or
    implicit ones — output evaluations, proof and point against the oracle (CpuTowerProver::create_proof semantics)."""
    rng = random.Random(4000 + seed)
    while True:
        prod = [rng.randint(1, 11) for _ in range(rng.randint(0, 5))]
        lk = [rng.randint(1, 10) for _ in range(rng.randint(0, 3))]
        if prod or lk:
            break
    _tower_case(dev, prod, lk, with_p=rng.random() < 0.5, seed=6000 + 100 * seed)
