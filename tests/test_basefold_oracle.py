"""Known-answer tests pinning the Basefold restatement (oracle/basefold.py): the prover's transcript is accepted by the
verifier restated from ceno_recursion_v2/src/pcs/mod.rs:1111-1317, 7494-7781, 444-592, every tampered part is rejected, and
the pieces agree with independent definitions (fold of a codeword pair == RS-encoding of fix_variable)."""
import copy
import random

import numpy as np
import pytest

from oracle import basefold as bf
from oracle import oracle as orc
from oracle import pyref as pr

P = pr.P


def _instance(shapes, seed, params):
    rng = random.Random(seed)
    p2 = orc.p2_params(seed=5)
    commits, points, evals = [], [], []
    for nv, width in shapes:
        cols = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(width)]
        commits.append(bf.commit(p2, params, cols, nv))
        pt = [(rng.randrange(P), rng.randrange(P)) for _ in range(nv)]
        points.append(pt)
        evals.append([pr.mle_evaluate([pr.efrom(x) for x in c], pt) for c in cols])
    return p2, commits, points, evals


@pytest.mark.parametrize("shapes", [[(3, 1)], [(4, 3)], [(4, 2), (4, 1)], [(5, 2), (3, 3), (4, 1)], [(2, 1), (5, 1)]])
def test_prover_transcript_is_accepted_by_the_restated_verifier(shapes):
    params = bf.Params(rate_log=1, n_queries=6, pow_bits=4)
    p2, commits, points, evals = _instance(shapes, 1234 + len(shapes), params)
    proof = bf.batch_open(p2, params, commits, points, evals, orc.Transcript(b"pcs"))
    roots = [c["root"] for c in commits]
    assert bf.batch_verify(p2, params, shapes, roots, points, evals, proof, orc.Transcript(b"pcs"))
    # shape facts the restatement checks: 2 evaluations per round, one commit per round, one final element per opening
    assert len(proof["sumcheck"]) == max(nv for nv, _ in shapes) == len(proof["commits"])
    assert all(len(row) == 1 for row in proof["final_message"])


def test_tampering_is_rejected():
    shapes = [(4, 2), (3, 1)]
    params = bf.Params(rate_log=1, n_queries=5, pow_bits=0)
    p2, commits, points, evals = _instance(shapes, 77, params)
    proof = bf.batch_open(p2, params, commits, points, evals, orc.Transcript(b"pcs"))
    roots = [c["root"] for c in commits]

    def rejects(mutate, ev=evals):
        bad = copy.deepcopy(proof)
        mutate(bad)
        with pytest.raises(bf.VerifyError):
            bf.batch_verify(p2, params, shapes, roots, points, ev, bad, orc.Transcript(b"pcs"))

    rejects(lambda p: p["sumcheck"].__setitem__(1, ((p["sumcheck"][1][0][0] ^ 1, p["sumcheck"][1][0][1]), p["sumcheck"][1][1])))
    rejects(lambda p: p["final_message"].__setitem__(0, [(p["final_message"][0][0][0] ^ 1, 0)]))
    rejects(lambda p: p["queries"][0]["inputs"][0]["opened"].__setitem__(0, (p["queries"][0]["inputs"][0]["opened"][0] + 1) % P))
    rejects(lambda p: p["queries"][2]["commit_phase"][1].__setitem__("sibling", (1, 2)))
    rejects(lambda p: p["commits"].__setitem__(0, [1, 2, 3, 4]))
    wrong = copy.deepcopy(evals)
    wrong[1][0] = (wrong[1][0][0] ^ 1, wrong[1][0][1])
    rejects(lambda p: None, ev=wrong)


def test_fold_of_codeword_is_the_codeword_of_the_folded_message():
    """The arrangement pinned by fold_codeword_pair / verifier_folding_coeff: folding adjacent entries of the bit-reversed
    RS codeword with r equals RS-encoding fix_variable(evals, r) — evaluations are the message coefficients, LSB first."""
    rng = random.Random(9)
    nv, rate = 5, 1
    f = [rng.randrange(P) for _ in range(1 << nv)]
    code = orc.rs_encode(np.array(f, dtype=np.uint64), 1, nv, rate, bitrev=True)
    r = (rng.randrange(P), rng.randrange(P))
    folded = [bf.fold_pair(pr.efrom(int(code[2 * i])), pr.efrom(int(code[2 * i + 1])), r, bf.folding_coeff(nv + rate, i)) for i in range(len(code) // 2)]
    g = pr.fix_variable([pr.efrom(x) for x in f], r)
    c0 = orc.rs_encode(np.array([x[0] for x in g], dtype=np.uint64), 1, nv - 1, rate, bitrev=True)
    c1 = orc.rs_encode(np.array([x[1] for x in g], dtype=np.uint64), 1, nv - 1, rate, bitrev=True)
    assert folded == [(int(a), int(b)) for a, b in zip(c0, c1)]
