"""Host-side bookkeeping (no GPU): the monomial expression engine (ceno_b200/expr.py), the EC-sum Quark term expansion
(EccQuarkProver.build_terms, reference ceno_zkvm/src/scheme/cpu/mod.rs:153-262) against the oracle's independent
expansion, and the zerocheck layer polynomial of ceno_b200/gkr.py (gkr_iop/src/gkr/layer/cpu/mod.rs:131-139)."""
import random

import numpy as np

from ceno_b200 import api, gkr
from ceno_b200.expr import Poly, SymbolicSepticExtension, ext, ext_mul
from oracle import oracle as orc
from oracle import pyref as pr

P = 0xFFFFFFFF00000001


def test_poly_algebra_and_term_table():
    x, y = Poly.var(0), Poly.var(1)
    p = (x + y) * (x - y) + Poly.const(5) * x
    assert p.t == {(0, 0): ext(1), (1, 1): ext(P - 1), (0,): ext(5)}          # x^2 - y^2 + 5x, cross terms cancel
    assert p.degree() == 2 and (p - p).t == {}
    assert [tuple(ids) for _, ids in p.terms()] == sorted(p.t)               # deterministic order
    a, b = (3, 4), (5, 6)
    assert ext_mul(a, b) == ((15 + 7 * 24) % P, (18 + 20) % P)                # (3+4X)(5+6X), X^2 = 7


def test_symbolic_septic_multiplication_matches_numeric():
    rng = random.Random(3)
    a = [rng.randrange(P) for _ in range(7)]
    b = [rng.randrange(P) for _ in range(7)]
    sa = SymbolicSepticExtension([Poly.var(i) for i in range(7)])
    sb = SymbolicSepticExtension([Poly.var(7 + i) for i in range(7)])
    prod = (sa * sb).to_exprs()
    vals = [(v, 0) for v in a + b]
    want = orc.septic_mul(a, b)
    for k in range(7):
        acc = (0, 0)
        for key, c in prod[k].t.items():
            t = c
            for i in key:
                t = ext_mul(t, vals[i])
            acc = ((acc[0] + t[0]) % P, (acc[1] + t[1]) % P)
        assert acc == (want[k], 0)


def test_ecc_quark_terms_match_the_oracle_expansion():
    rng = random.Random(11)
    alpha = np.array([[rng.randrange(P), rng.randrange(P)] for _ in range(49)], dtype=np.uint64)
    fx, fy = [rng.randrange(P) for _ in range(7)], [rng.randrange(P) for _ in range(7)]
    from ceno_b200 import build as cbuild
    cbuild.build()
    terms = api.EccQuarkProver.build_terms(alpha, fx, fy)                       # the library's C++ expansion (cg_ecc_quark_terms)
    got = sorted((tuple(c), tuple(i)) for c, i in terms)
    sym = sorted((tuple(c), tuple(i)) for c, i in api.EccQuarkProver.build_terms_symbolic(alpha, fx, fy))
    want = sorted((tuple(c), tuple(i)) for c, i in orc.ecc_quark_terms(alpha, fx, fy))
    assert got == want == sym and len(got) == 260
    assert max(len(i) for _, i in got) == 3
    assert [tuple(i) for _, i in terms] == sorted(tuple(i) for _, i in terms)     # deterministic (lexicographic) order
    assert [(tuple(c), tuple(i)) for c, i in terms] == [(tuple(c), tuple(i)) for c, i in api.EccQuarkProver.build_terms_symbolic(alpha, fx, fy)]


def test_zerocheck_layer_polynomial():
    """sum_g sel_g * sum_j alpha_j expr_{g,j}: a selector-less group contributes its expressions unmasked."""
    w0, w1 = Poly.var(0), Poly.var(1)
    layer = gkr.Layer("l", gkr.ZEROCHECK, 2, 0, 1, [w0 * w1, w0 + w1, w1 * 2],
                      [gkr.OutGroup(0, 0, [0, 1]), gkr.OutGroup(None, 0, [2])], in_eval_positions=[3, 4, 5])
    al = [(1, 0), (9, 2), (4, 4)]
    terms = layer.main_sumcheck_terms(np.array(al, dtype=np.uint64))
    rng = random.Random(5)
    vals = [(rng.randrange(P), 0), (rng.randrange(P), 0), (rng.randrange(P), rng.randrange(P))]      # w0, w1, sel (structural id 0 -> index 2)
    got = pr.ZERO
    for c, ids in terms:
        t = (int(c[0]), int(c[1]))
        for i in ids:
            t = pr.emul(t, vals[i])
        got = pr.eadd(got, t)
    e = [pr.emul(vals[0], vals[1]), pr.eadd(vals[0], vals[1]), pr.emul((2, 0), vals[1])]
    want = pr.eadd(pr.emul(vals[2], pr.eadd(pr.emul(al[0], e[0]), pr.emul(al[1], e[1]))), pr.emul(al[2], e[2]))
    assert got == want
    assert layer.max_expr_degree == 2
