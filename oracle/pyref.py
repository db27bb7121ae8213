"""Pure-Python (big-int) restatement of the same path — for SMALL cases only.

TEST INFRASTRUCTURE.  A second, independent statement of the algebra (Python
ints with `% P`, no shared code with ceno_oracle.c) used to cross-check the C
oracle and to spell out the known-answer relations the reference tree holds:
  gkr_iop/src/utils.rs:332-441, gkr_iop/src/selector.rs:396-435,
  ceno_zkvm/src/scheme/utils.rs:934-1195, SURVEY.md §A.
An extension element is a tuple (c0, c1) in F_p[X]/(X^2 - 7).
"""
P = 0xFFFFFFFF00000001
W = 7


def eadd(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def esub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def emul(a, b):
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def efrom(x):
    return (x % P, 0)


ONE = (1, 0)
ZERO = (0, 0)


def einv(a):
    n = (a[0] * a[0] - W * a[1] * a[1]) % P
    ni = pow(n, P - 2, P)
    return (a[0] * ni % P, (-a[1]) * ni % P)


def build_eq_x_r_vec(r):
    """eq[b] = prod_i (b_i r_i + (1-b_i)(1-r_i)), direct per-index product."""
    k = len(r)
    out = []
    for b in range(1 << k):
        acc = ONE
        for i in range(k):
            acc = emul(acc, r[i] if (b >> i) & 1 else esub(ONE, r[i]))
        out.append(acc)
    return out


def mle_evaluate(evals, point):
    """sum_b evals[b] * eq(point, b)  — definition, not the folding algorithm."""
    eq = build_eq_x_r_vec(point)
    acc = ZERO
    for e, w in zip(evals, eq):
        acc = eadd(acc, emul(e, w))
    return acc


def fix_variable(evals, r):
    return [eadd(evals[2 * b], emul(r, esub(evals[2 * b + 1], evals[2 * b]))) for b in range(len(evals) // 2)]


def poly_eval(mles, terms, idx):
    """P at hypercube index idx."""
    acc = ZERO
    for c, ids in terms:
        p = c
        for i in ids:
            p = emul(p, mles[i][idx])
        acc = eadd(acc, p)
    return acc


def round_message(mles, terms, degree):
    """[p(1..degree)] by literally substituting X=t into each pair."""
    half = len(mles[0]) // 2
    msg = []
    for t in range(1, degree + 1):
        acc = ZERO
        for b in range(half):
            vals = [eadd(m[2 * b], emul(efrom(t), esub(m[2 * b + 1], m[2 * b]))) for m in mles]
            for c, ids in terms:
                p = c
                for i in ids:
                    p = emul(p, vals[i])
                acc = eadd(acc, p)
        msg.append(acc)
    return msg


def sumcheck_prove(mles, terms, num_vars, degree, challenge_fn):
    mles = [list(m) for m in mles]
    msgs, chals = [], []
    for j in range(num_vars):
        msg = round_message(mles, terms, degree)
        r = challenge_fn(j, msg)
        msgs.append(msg)
        chals.append(r)
        mles = [fix_variable(m, r) for m in mles]
    return msgs, [m[0] for m in mles], chals


def lagrange_eval(ys, r):
    """interpolate through (i, ys[i]) i=0..d, evaluate at ext r."""
    d = len(ys) - 1
    acc = ZERO
    for i in range(d + 1):
        num, den = ONE, 1
        for j in range(d + 1):
            if j != i:
                num = emul(num, esub(r, efrom(j)))
                den = den * (i - j) % P
        acc = eadd(acc, emul(ys[i], emul(num, efrom(pow(den, P - 2, P)))))
    return acc


def to_pairs(arr):
    a = [int(x) for x in arr]
    return [(a[2 * i], a[2 * i + 1]) for i in range(len(a) // 2)]


def from_pairs(pairs):
    import numpy as np
    out = np.zeros(2 * len(pairs), dtype=np.uint64)
    for i, (a, b) in enumerate(pairs):
        out[2 * i] = a
        out[2 * i + 1] = b
    return out


# ---------------------------------------------------------------- Poseidon2 width 8 (structure only; constants are parameters)
M4 = {0: [[2, 3, 1, 1], [1, 2, 3, 1], [1, 1, 2, 3], [3, 1, 1, 2]], 1: [[5, 7, 1, 3], [4, 6, 1, 1], [1, 3, 5, 7], [1, 1, 4, 6]]}


def _p2_external(s, variant):
    m = M4[variant]
    s = [sum(m[i][j] * s[4 * c + j] for j in range(4)) % P for c in range(2) for i in range(4)]
    sums = [(s[i] + s[4 + i]) % P for i in range(4)]
    return [(s[4 * c + i] + sums[i]) % P for c in range(2) for i in range(4)]


def poseidon2_permute(ext_rc, int_rc, diag, variant, state):
    s = _p2_external([x % P for x in state], variant)
    for r in range(4):
        s = _p2_external([pow((s[i] + ext_rc[r][i]) % P, 7, P) for i in range(8)], variant)
    for r in range(22):
        s[0] = pow((s[0] + int_rc[r]) % P, 7, P)
        tot = sum(s) % P
        s = [(s[i] * diag[i] + tot) % P for i in range(8)]
    for r in range(4, 8):
        s = _p2_external([pow((s[i] + ext_rc[r][i]) % P, 7, P) for i in range(8)], variant)
    return s


def hash_row(perm, row):
    st = [0] * 8
    for c in range(0, len(row), 4):
        chunk = row[c:c + 4]
        st[:len(chunk)] = [x % P for x in chunk]
        st = perm(st)
    return st[:4]


def merkle_root(perm, rows):
    level = [hash_row(perm, r) for r in rows]
    while len(level) > 1:
        level = [perm(level[2 * i] + level[2 * i + 1])[:4] for i in range(len(level) // 2)]
    return level[0]
