"""The algebra behind the split-eq rounds of a general tower layer (ceno_b200/csrc/sumcheck_kernels.cuh: tveq_round_kernel,
VeqFin::post), restated with Python big ints and checked against the TABLE formulation of the same layer sumcheck
(CpuTowerProver::create_proof's layer polynomial, ceno_zkvm/src/scheme/cpu/mod.rs:417-485; oracle/pyref.py round_message).

Pinned here, on the CPU, independent of any device code:
  * p_j(X) = P_j eq(w_j, X) q_j(X) with q_j(X) = sum_x F[x_hi] U[x_lo] g(X, x) for ANY split of the remaining variables;
  * the alphas folded into the uniform weights, ONE weighted operand per product (product spec: W a; logup spec: u = Wn p1 + Wd q1,
    v = Wn p2, then u q2 + v q1);
  * claim-derived rounds: only q(1) and the X^2 coefficient are accumulated and q(0) follows from the running claim — in round 0
    from the claimed sum, afterwards from q_{j-1}(r_{j-1});
  * numerators that are the constant one: g = (an + ad q1) q2 + an q1, differences of p vanish;
  * record padding: pairs whose every slot holds its default on both sides contribute K * sum U[lo] (times the fixed weight) to
    q(0) = q(1) and nothing to the X^2 coefficient.
TEST INFRASTRUCTURE (pure Python, small sizes)."""
import random

from oracle import pyref as pr

P = pr.P


def rnd_ext(rng):
    return (rng.randrange(P), rng.randrange(P))


def eq_table(w):
    return pr.build_eq_x_r_vec(list(w))


def layer_terms(n_prod, n_logup, alpha_p, alpha_n, alpha_d):
    """MLE order: eq, then (a, b) per product spec, then (p1, p2, q1, q2) per logup spec — the monomial terms of the layer."""
    terms, idx = [], 1
    for p in range(n_prod):
        terms.append((alpha_p[p], [0, idx, idx + 1]))
        idx += 2
    for l in range(n_logup):
        p1, p2, q1, q2 = idx, idx + 1, idx + 2, idx + 3
        terms += [(alpha_n[l], [0, p1, q2]), (alpha_n[l], [0, p2, q1]), (alpha_d[l], [0, q1, q2])]
        idx += 4
    return terms


def split_round(state, n_prod, n_logup, alpha_p, alpha_n, alpha_d, w_rest, lo_bits, pone=(), pad_lo=None, defaults=None):
    """(q(1), c2) of one round from the CURRENT arrays (without eq), the way tveq_round_kernel accumulates them:
    item = (hi << lo_bits) | lo, weight = F[hi] * U_s[lo] with the spec's alpha folded into U."""
    n_pairs = len(state[0]) // 2
    n_lo = 1 << lo_bits
    U = eq_table(w_rest[:lo_bits])
    F = eq_table(w_rest[lo_bits:])
    assert len(U) * len(F) == n_pairs
    s1, c2 = pr.ZERO, pr.ZERO
    for hi in range(len(F)):
        t1, tc = pr.ZERO, pr.ZERO
        for lo in range(n_lo):
            if pad_lo is not None and lo >= pad_lo:
                continue
            item = (hi << lo_bits) | lo
            slot = 0
            for p in range(n_prod):
                W = pr.emul(alpha_p[p], U[lo])
                a, b = state[slot], state[slot + 1]
                alo, ahi, blo, bhi = a[2 * item], a[2 * item + 1], b[2 * item], b[2 * item + 1]
                t1 = pr.eadd(t1, pr.emul(pr.emul(W, ahi), bhi))
                tc = pr.eadd(tc, pr.emul(pr.emul(W, pr.esub(alo, ahi)), pr.esub(blo, bhi)))
                slot += 2
            for l in range(n_logup):
                Wn, Wd = pr.emul(alpha_n[l], U[lo]), pr.emul(alpha_d[l], U[lo])
                p1, p2, q1, q2 = (state[slot + z] for z in range(4))
                lo_v = [m[2 * item] for m in (p1, p2, q1, q2)]
                hi_v = [m[2 * item + 1] for m in (p1, p2, q1, q2)]
                d_v = [pr.esub(x, y) for x, y in zip(lo_v, hi_v)]
                if l in pone:      # p1 = p2 = 1
                    t1 = pr.eadd(t1, pr.eadd(pr.emul(pr.eadd(pr.emul(Wd, hi_v[2]), Wn), hi_v[3]), pr.emul(Wn, hi_v[2])))
                    tc = pr.eadd(tc, pr.emul(pr.emul(Wd, d_v[2]), d_v[3]))
                else:
                    u_h = pr.eadd(pr.emul(Wn, hi_v[0]), pr.emul(Wd, hi_v[2]))
                    t1 = pr.eadd(t1, pr.eadd(pr.emul(u_h, hi_v[3]), pr.emul(pr.emul(Wn, hi_v[1]), hi_v[2])))
                    u_d = pr.eadd(pr.emul(Wn, d_v[0]), pr.emul(Wd, d_v[2]))
                    tc = pr.eadd(tc, pr.eadd(pr.emul(u_d, d_v[3]), pr.emul(pr.emul(Wn, d_v[1]), d_v[2])))
                slot += 4
        if pad_lo is not None and pad_lo < n_lo:   # the padding's constant contribution, once per (hi): K * sum_{lo >= pad_lo} U[lo]
            K, slot = pr.ZERO, 0
            for p in range(n_prod):
                K = pr.eadd(K, pr.emul(alpha_p[p], pr.emul(defaults[slot], defaults[slot + 1])))
                slot += 2
            for l in range(n_logup):
                dp1, dp2, dq1, dq2 = (defaults[slot + z] for z in range(4))
                K = pr.eadd(K, pr.emul(alpha_n[l], pr.eadd(pr.emul(dp1, dq2), pr.emul(dp2, dq1))))
                K = pr.eadd(K, pr.emul(alpha_d[l], pr.emul(dq1, dq2)))
                slot += 4
            usum = pr.ZERO
            for lo in range(pad_lo, n_lo):
                usum = pr.eadd(usum, U[lo])
            t1 = pr.eadd(t1, pr.emul(K, usum))
        s1 = pr.eadd(s1, pr.emul(F[hi], t1))
        c2 = pr.eadd(c2, pr.emul(F[hi], tc))
    return s1, c2


def finish(q1, c2, claim, w_j, prefix):
    """VeqFin::post: q(0) from (1 - w_j) q(0) + w_j q(1) = claim, then [p(1), p(2), p(3)] = P eq(w_j, t) q(t)."""
    q0 = pr.emul(pr.esub(claim, pr.emul(w_j, q1)), pr.einv(pr.esub(pr.ONE, w_j)))
    c1 = pr.esub(pr.esub(q1, q0), c2)
    q = lambda t: pr.eadd(pr.eadd(q0, pr.emul(pr.efrom(t), c1)), pr.emul(pr.efrom(t * t), c2))   # noqa: E731
    e = lambda t: pr.eadd(pr.esub(pr.ONE, w_j), pr.emul(pr.efrom(t), pr.esub(pr.emul(pr.efrom(2), w_j), pr.ONE)))   # noqa: E731
    return [pr.emul(pr.emul(prefix, e(t)), q(t)) for t in (1, 2, 3)], (q0, c1, c2)


def run_case(rng, k, n_prod, n_logup, lo_bits_of_round, pone=(), padding=None):
    n = 1 << k
    w = [rnd_ext(rng) for _ in range(k)]
    alpha_p = [rnd_ext(rng) for _ in range(n_prod)]
    alpha_n = [rnd_ext(rng) for _ in range(n_logup)]
    alpha_d = [rnd_ext(rng) for _ in range(n_logup)]
    n_arr = 2 * n_prod + 4 * n_logup
    arrays = [[rnd_ext(rng) for _ in range(n)] for _ in range(n_arr)]
    for l in pone:
        arrays[2 * n_prod + 4 * l] = [pr.ONE] * n
        arrays[2 * n_prod + 4 * l + 1] = [pr.ONE] * n
    defaults = None
    if padding is not None:   # (l2m, n_records): leaf index = row << l2m | record; records >= n_records hold the slot's default
        l2m, n_rec = padding
        defaults = [rnd_ext(rng) for _ in range(n_arr)]
        for l in pone:
            defaults[2 * n_prod + 4 * l] = defaults[2 * n_prod + 4 * l + 1] = pr.ONE
        for z in range(n_arr):
            for x in range(n):
                if (x & ((1 << l2m) - 1)) >= n_rec:
                    arrays[z][x] = defaults[z]
    terms = layer_terms(n_prod, n_logup, alpha_p, alpha_n, alpha_d)
    table = [eq_table(w)] + [list(a) for a in arrays]
    claim = pr.ZERO
    for x in range(n):
        claim = pr.eadd(claim, pr.poly_eval(table, terms, x))
    state = [list(a) for a in arrays]
    prefix, qstate, r_prev = pr.ONE, None, None
    for j, lo_bits in enumerate(lo_bits_of_round):
        want = pr.round_message(table, terms, 3)
        if j > 0:      # fold by the previous challenge; the running claim is q_{j-1}(r_{j-1})
            state = [pr.fix_variable(a, r_prev) for a in state]
            prefix = pr.emul(prefix, pr.eadd(pr.emul(pr.esub(pr.ONE, w[j - 1]), pr.esub(pr.ONE, r_prev)), pr.emul(w[j - 1], r_prev)))
            claim_j = pr.eadd(qstate[0], pr.emul(r_prev, pr.eadd(qstate[1], pr.emul(r_prev, qstate[2]))))
        else:
            claim_j = claim
        pad_lo = None
        if padding is not None and j < 2:   # launches that read the leaves: pairs of 2^(j+1) leaves, all default from here on
            l2m, n_rec = padding
            assert lo_bits == l2m - 1 - j
            pad_lo = (n_rec + (2 << j) - 1) >> (j + 1)
        q1, c2 = split_round(state, n_prod, n_logup, alpha_p, alpha_n, alpha_d, w[j + 1:], lo_bits, pone if j < 2 else (), pad_lo, defaults)
        got, qstate = finish(q1, c2, claim_j, w[j], prefix)
        assert got == want, f"round {j}"
        r_prev = rnd_ext(rng)
        table = [pr.fix_variable(m, r_prev) for m in table]


def test_split_eq_general_layer_equals_table_formulation():
    rng = random.Random(20261017)
    run_case(rng, 7, 2, 1, [3, 2, 4, 0])          # two product specs + one logup spec, different splits per round
    run_case(rng, 6, 0, 2, [2, 2, 1])
    run_case(rng, 6, 3, 0, [5, 0, 2])


def test_split_eq_constant_numerators():
    rng = random.Random(77)
    run_case(rng, 6, 1, 1, [2, 3, 1], pone=(0,))


def test_split_eq_record_padding_closed_form():
    rng = random.Random(99)
    # leaf layer of 2^7 entries per array: 2^3 = 8 record slots per row, 5 records (3 slots of padding), 16 rows
    run_case(rng, 7, 0, 1, [2, 1, 3], pone=(0,), padding=(3, 5))
    run_case(rng, 7, 1, 1, [2, 1, 2], padding=(3, 3))
    run_case(rng, 8, 0, 1, [3, 2], padding=(4, 9))
