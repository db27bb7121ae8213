"""CPU restatement of the Basefold PCS path (TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline may import this): commit (RS-encode + Merkle), batch_open (prover) and the verifier.

The protocol lives in the un-vendored `mpcs` crate (call sites: TraceCommitter::commit_traces / OpeningProver::open,
ceno_zkvm/src/scheme/cpu/mod.rs:559-584, 1415-1457; GPU ceno_zkvm/src/scheme/gpu/mod.rs:1062-1509, 3324-3413), but its
VERIFIER is restated in-tree by the recursion circuit's preflight, which this file follows line by line:
  * transcript order, batch coefficients, initial claim with the 2^(max_num_var - num_var) scale, degree-2 sumcheck with
    two evaluations per round, `commit round` label, commit digests observed after the round's challenge, final message,
    `query indices` label and sample_bits:                     ceno_recursion_v2/src/pcs/mod.rs:1111-1317  (replay_basefold)
  * query checks — input-MMCS leaf = PaddingFreeSponge over the opened row, reduced openings added at the matching height,
    commit-phase leaf = hash of the (even, odd) ext pair, sibling openings, final codeword check:       :7494-7727
  * fold rule  lo = (a+b)/2, hi = (a-b) g_h^{-bitrev(idx)} / 2, lo + r (hi - lo):                          :7765-7781
  * final claim  sum_p prod_i eq(point_p[i], r_{..}) * final_message[p]:                                     :444-592
  * basecode_log == 0 (one final element per opening):                                                        :138-145
  * the final codeword is the bit-reversed DFT of the summed final messages, zero-padded by the rate:         :7729-7751
The MLE's evaluation vector is the coefficient vector of the RS message polynomial, so folding a codeword pair is
fix_variable (LSB first) of the evaluations.  PARITY UNPINNED for what the restatement leaves to upstream: rate_log,
number of queries and proof-of-work bits (parameters here), the Poseidon2 constants (placeholders), the duplex
challenger (stand-in sponge) and mixed-height MMCS commitments (one matrix per commitment here, as in the restatement,
which rejects mixed heights :7797-7803).
"""
import numpy as np

from . import oracle as orc
from . import pyref as pr

P = pr.P


class Params:
    def __init__(self, rate_log=1, n_queries=8, pow_bits=0):
        self.rate_log, self.n_queries, self.pow_bits = rate_log, n_queries, pow_bits


# ------------------------------------------------------------------ transcript events (stand-in sponge, restated order)
def observe_label(tr, label: bytes):
    tr.append_message(label)


def sample_ext(tr):
    o = np.zeros(2, np.uint64)
    orc.lib().or_tr_challenge(orc.C.byref(tr.t), orc._p(o))
    return (int(o[0]), int(o[1]))


def observe_exts(tr, vals):
    tr.append_ext(np.array([x for v in vals for x in v], dtype=np.uint64))


def observe_digest(tr, d):
    tr.append_ext(np.array([int(x) for x in d], dtype=np.uint64))      # 4 base words = the same absorb sequence as 2 ext


def sample_bits(tr, bits):
    return sample_ext(tr)[0] & ((1 << bits) - 1)


def grind(tr, bits):
    """Proof of work (check_witness, pcs/mod.rs:1255-1259): observe the witness, sample `bits` bits, they must be zero."""
    if bits == 0:
        return 0
    w = 0
    while True:
        probe = orc.Transcript(b"")
        probe.t.h = tr.t.h
        observe_digest(probe, [w, 0])
        if sample_bits(probe, bits) == 0:
            observe_digest(tr, [w, 0])
            assert sample_bits(tr, bits) == 0
            return w
        w += 1


def check_witness(tr, bits, w):
    if bits == 0:
        return True
    observe_digest(tr, [w, 0])
    return sample_bits(tr, bits) == 0


# ------------------------------------------------------------------ hashing
def _perm(p2):
    return lambda st: [int(x) for x in orc.poseidon2_permute(p2, np.array(st, dtype=np.uint64))]


def leaf_hash_row(p2, row):
    return pr.hash_row(_perm(p2), [int(x) for x in row])


def leaf_hash_pair(p2, a, b):
    return pr.hash_row(_perm(p2), [a[0], a[1], b[0], b[1]])            # poseidon2_hash_ext_pair (:8006-8010 region)


def compress(p2, left, right):
    return _perm(p2)(list(left) + list(right))[:4]


def merkle_levels(p2, leaves):
    levels = [leaves]
    while len(levels[-1]) > 1:
        cur = levels[-1]
        levels.append([compress(p2, cur[2 * i], cur[2 * i + 1]) for i in range(len(cur) // 2)])
    return levels


def merkle_path(levels, idx):
    path = []
    for lvl in levels[:-1]:
        path.append(lvl[idx ^ 1])
        idx >>= 1
    return path


def merkle_replay(p2, leaf, idx, path):
    cur = leaf
    for sib in path:
        cur = compress(p2, cur, sib) if (idx & 1) == 0 else compress(p2, sib, cur)
        idx >>= 1
    return cur


# ------------------------------------------------------------------ commit
def commit(p2, params, cols, nv):
    """cols: list of `width` base-field evaluation vectors of 2^nv entries (one committed matrix).  Returns the
    commitment-with-witness: codeword matrix (rows in bit-reversed order), Merkle levels, root."""
    width = len(cols)
    msg = np.concatenate([np.asarray(c, dtype=np.uint64) for c in cols])
    code = orc.rs_encode(msg, width, nv, params.rate_log, bitrev=True).reshape(width, -1)
    h = 1 << (nv + params.rate_log)
    rows = [[int(code[c][i]) for c in range(width)] for i in range(h)]
    levels = merkle_levels(p2, [leaf_hash_row(p2, r) for r in rows])
    return {"nv": nv, "width": width, "cols": [[int(x) for x in c] for c in cols], "rows": rows, "levels": levels, "root": levels[-1][0]}


def _bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def folding_coeff(log2_height, leaf_idx):
    """verifier_folding_coeff (:7765-7769): g_h^{-bitrev(leaf_idx, h-1)} / 2."""
    g_inv = pow(orc.two_adic_generator(log2_height), P - 2, P)
    return pow(g_inv, _bitrev(leaf_idx, log2_height - 1), P) * pow(2, P - 2, P) % P


def fold_pair(a, b, r, coeff):
    """fold_codeword_pair (:7771-7781)."""
    inv2 = pow(2, P - 2, P)
    s = pr.eadd(a, b)
    lo = (s[0] * inv2 % P, s[1] * inv2 % P)
    d = pr.esub(a, b)
    hi = (d[0] * coeff % P, d[1] * coeff % P)
    return pr.eadd(lo, pr.emul(r, pr.esub(hi, lo)))


def _pows(alpha, n):
    out, acc = [], (1, 0)
    for _ in range(n):
        out.append(acc)
        acc = pr.emul(acc, alpha)
    return out


# ------------------------------------------------------------------ prover
def batch_open(p2, params, commits, points, evals, tr):
    """commits[i]: commit() output; points[i]: nv_i ext; evals[i]: width_i ext (the claimed evaluations).
    Returns the proof as a dict of python ints (the layout cg_basefold_batch_open serialises)."""
    rate = params.rate_log
    total = sum(c["width"] for c in commits)
    observe_label(tr, b"batch coeffs")
    coeffs = _pows(sample_ext(tr), total)
    max_nv = max(c["nv"] for c in commits)
    num_rounds = max_nv
    off, cofs = 0, []
    for c in commits:
        cofs.append(coeffs[off:off + c["width"]])
        off += c["width"]
    # per opening: g = sum_j coeff_j f_j (ext evaluations), eq(point, .), claimed sum S = sum_j coeff_j eval_j
    g, eqs, S = [], [], []
    for c, pt, ev, cf in zip(commits, points, evals, cofs):
        n = 1 << c["nv"]
        g.append([pr.efrom(0)] * n)
        for j in range(c["width"]):
            g[-1] = [pr.eadd(g[-1][b], (cf[j][0] * c["cols"][j][b] % P, cf[j][1] * c["cols"][j][b] % P)) for b in range(n)]
        eqs.append(pr.build_eq_x_r_vec([tuple(int(x) for x in p) for p in pt]) if c["nv"] else [(1, 0)])
        acc = (0, 0)
        for j in range(c["width"]):
            acc = pr.eadd(acc, pr.emul(cf[j], tuple(int(x) for x in ev[j])))
        S.append(acc)

    def rlc_codeword(nv):
        h = 1 << (nv + rate)
        out = [(0, 0)] * h
        for c, cf in zip(commits, cofs):
            if c["nv"] != nv:
                continue
            for i in range(h):
                acc = out[i]
                for j in range(c["width"]):
                    acc = pr.eadd(acc, (cf[j][0] * c["rows"][i][j] % P, cf[j][1] * c["rows"][i][j] % P))
                out[i] = acc
        return out

    oracle = rlc_codeword(max_nv)
    msgs, roots, trees, oracles, chals = [], [], [], [], []
    for r in range(num_rounds):
        # ---- sumcheck message [p(1), p(2)]
        e1, e2 = (0, 0), (0, 0)
        for c, gp, ep, sp in zip(commits, g, eqs, S):
            join = max_nv - c["nv"]
            if r < join:                                   # not joined yet: constant in X, 2^(join - r - 1) copies of its sum
                k = pow(2, join - r - 1, P)
                v = (sp[0] * k % P, sp[1] * k % P)
                e1, e2 = pr.eadd(e1, v), pr.eadd(e2, v)
                continue
            for b in range(len(gp) // 2):
                g0, g1, q0, q1 = gp[2 * b], gp[2 * b + 1], ep[2 * b], ep[2 * b + 1]
                e1 = pr.eadd(e1, pr.emul(g1, q1))
                g2, q2 = pr.esub(pr.eadd(g1, g1), g0), pr.esub(pr.eadd(q1, q1), q0)
                e2 = pr.eadd(e2, pr.emul(g2, q2))
        msgs.append((e1, e2))
        observe_exts(tr, [e1, e2])
        observe_label(tr, b"commit round")
        ch = sample_ext(tr)
        chals.append(ch)
        # ---- commit the current oracle as (even, odd) pairs; the digest is observed after the challenge (:1222-1227)
        h = len(oracle)
        levels = merkle_levels(p2, [leaf_hash_pair(p2, oracle[2 * i], oracle[2 * i + 1]) for i in range(h // 2)])
        roots.append(levels[-1][0])
        trees.append(levels)
        oracles.append(oracle)
        observe_digest(tr, levels[-1][0])
        # ---- fold codeword and polynomials
        lh = max_nv + rate - r
        oracle = [fold_pair(oracle[2 * i], oracle[2 * i + 1], ch, folding_coeff(lh, i)) for i in range(h // 2)]
        nxt = max_nv - r - 1
        if any(c["nv"] == nxt for c in commits) and r + 1 < num_rounds:
            add = rlc_codeword(nxt)
            oracle = [pr.eadd(a, b) for a, b in zip(oracle, add)]
        for i, c in enumerate(commits):
            if r >= max_nv - c["nv"] and len(g[i]) > 1:
                g[i] = pr.fix_variable(g[i], ch)
                eqs[i] = pr.fix_variable(eqs[i], ch)
    final_message = [[gp[0]] for gp in g]                  # basecode_log = 0: one element per opening
    observe_exts(tr, [row[0] for row in final_message])
    pow_witness = grind(tr, params.pow_bits)
    observe_label(tr, b"query indices")
    query_bits = max_nv + rate
    queries = [sample_bits(tr, query_bits) for _ in range(params.n_queries)]
    qproofs = []
    for q in queries:
        inputs = []
        for c in commits:
            red = q >> (max_nv - c["nv"])
            inputs.append({"opened": list(c["rows"][red]), "path": merkle_path(c["levels"], red)})
        cps, idx = [], q
        for r in range(num_rounds):
            cps.append({"sibling": oracles[r][idx ^ 1], "path": merkle_path(trees[r], idx >> 1)})
            idx >>= 1
        qproofs.append({"index": q, "inputs": inputs, "commit_phase": cps})
    return {"sumcheck": msgs, "commits": roots, "final_message": final_message, "pow_witness": pow_witness, "queries": qproofs}


# ------------------------------------------------------------------ verifier
class VerifyError(Exception):
    pass


def batch_verify(p2, params, shapes, roots, points, evals, proof, tr):
    """shapes[i] = (nv_i, width_i); roots[i] = commitment root.  Follows replay_basefold + record_basefold_query_checks."""
    rate = params.rate_log
    total = sum(w for _, w in shapes)
    observe_label(tr, b"batch coeffs")
    coeffs = _pows(sample_ext(tr), total)
    max_nv = max(nv for nv, _ in shapes)
    num_rounds = max_nv
    if len(proof["sumcheck"]) != num_rounds or len(proof["commits"]) != num_rounds:
        raise VerifyError("basefold round count mismatch")
    expected, it = (0, 0), iter(coeffs)
    for (nv, w), ev in zip(shapes, evals):
        scale = pow(2, max_nv - nv, P)
        for j in range(w):
            cf = next(it)
            t = pr.emul(cf, tuple(int(x) for x in ev[j]))
            expected = pr.eadd(expected, (t[0] * scale % P, t[1] * scale % P))
    chals, claim = [], expected
    for r, (e1, e2) in enumerate(proof["sumcheck"]):
        observe_exts(tr, [e1, e2])
        observe_label(tr, b"commit round")
        ch = sample_ext(tr)
        chals.append(ch)
        claim = pr.lagrange_eval([pr.esub(claim, e1), e1, e2], ch)
        observe_digest(tr, proof["commits"][r])
    observe_exts(tr, [row[0] for row in proof["final_message"]])
    if not check_witness(tr, params.pow_bits, proof["pow_witness"]):
        raise VerifyError("basefold pow witness check failed")
    observe_label(tr, b"query indices")
    query_bits = max_nv + rate
    queries = [sample_bits(tr, query_bits) for _ in range(params.n_queries)]
    # final codeword: bit-reversed DFT of the summed final messages, zero-padded by the rate (:7729-7751)
    fsum = (0, 0)
    for row in proof["final_message"]:
        if len(row) != 1:
            raise VerifyError("basefold final message width does not match basecode size")
        fsum = pr.eadd(fsum, row[0])
    final_codeword = [fsum] * (1 << rate)                 # the DFT of (c, 0, ..., 0) is constant
    if len(proof["queries"]) != len(queries):
        raise VerifyError("basefold query opening count mismatch")
    for q, qp in zip(queries, proof["queries"]):
        reduced = {}
        it = iter(coeffs)
        for (nv, w), root, inp in zip(shapes, roots, qp["inputs"]):
            red = q >> (max_nv - nv)
            if len(inp["opened"]) != w:
                raise VerifyError("basefold opened-value width mismatch")
            if merkle_replay(p2, leaf_hash_row(p2, inp["opened"]), red, inp["path"]) != list(root):
                raise VerifyError("base input MMCS root replay mismatch")
            lh = nv + rate
            acc = reduced.get(lh, (0, 0))
            for v in inp["opened"]:
                cf = next(it)
                acc = pr.eadd(acc, (cf[0] * v % P, cf[1] * v % P))
            reduced[lh] = acc
        idx, folded, lh = q, (0, 0), max_nv + rate
        for r in range(num_rounds):
            ro = reduced.pop(lh, (0, 0))
            leafs = [qp["commit_phase"][r]["sibling"]] * 2
            leafs[idx & 1] = pr.eadd(folded, ro)
            leaf_idx = idx >> 1
            if merkle_replay(p2, leaf_hash_pair(p2, leafs[0], leafs[1]), leaf_idx, qp["commit_phase"][r]["path"]) != list(proof["commits"][r]):
                raise VerifyError("commit phase Merkle root mismatch")
            folded = fold_pair(leafs[0], leafs[1], chals[r], folding_coeff(lh, leaf_idx))
            lh -= 1
            idx >>= 1
        if reduced:
            raise VerifyError("basefold unused reduced openings remain")
        if final_codeword[idx] != folded:
            raise VerifyError("basefold final codeword reconstruction mismatch")
    # final claim (:444-592)
    acc = (0, 0)
    for (nv, _), pt, row in zip(shapes, points, proof["final_message"]):
        cf = (1, 0)
        for i in range(nv):
            x, y = tuple(int(v) for v in pt[i]), chals[num_rounds - nv + i]
            xy = pr.emul(x, y)
            cf = pr.emul(cf, pr.eadd(pr.esub(pr.esub(pr.eadd(xy, xy), x), y), (1, 0)))
        acc = pr.eadd(acc, pr.emul(cf, row[0]))
    if acc != claim:
        raise VerifyError("basefold final claim mismatch")
    return True
