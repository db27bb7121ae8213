"""ctypes binding of the CPU oracle (oracle/ceno_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
`ceno_b200` never imports this module.

Arrays are numpy uint64; an extension element is two consecutive u64 limbs
[c0, c1] (the host in-memory layout of GoldilocksExt2 the reference transmutes,
gkr_iop/src/gpu/mod.rs:311-324).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libceno_oracle.so")
P = 0xFFFFFFFF00000001

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "ceno_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class OrMle(C.Structure):
    _fields_ = [("data", C.c_void_p), ("num_vars", C.c_uint32), ("is_ext", C.c_uint32)]


class OrTranscript(C.Structure):
    _fields_ = [("h", C.c_uint64)]


CHALLENGE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, u64p, C.c_uint32, u64p)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.or_gl_add.restype = C.c_uint64
        _lib.or_gl_sub.restype = C.c_uint64
        _lib.or_gl_mul.restype = C.c_uint64
        _lib.or_gl_mul_slow.restype = C.c_uint64
        _lib.or_gl_inv.restype = C.c_uint64
        for f in ("or_gl_add", "or_gl_sub", "or_gl_mul", "or_gl_mul_slow"):
            getattr(_lib, f).argtypes = [C.c_uint64, C.c_uint64]
        _lib.or_gl_inv.argtypes = [C.c_uint64]
        _lib.or_interleave_out_len.restype = C.c_uint64
        _lib.or_interleave_out_len.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32]
        _lib.or_tower_create_proof.restype = C.c_int64
        _lib.or_num_threads.restype = C.c_int
        _lib.or_two_adic_generator.restype = C.c_uint64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def num_threads():
    return lib().or_num_threads()


def set_num_threads(n):
    """OpenMP thread count of the C restatement (torchrun exports OMP_NUM_THREADS=1 to its ranks)."""
    lib().or_set_num_threads(C.c_int(int(n)))


# ------------------------------------------------------------------ field
def gl_mul(a, b):
    return lib().or_gl_mul(a, b)


def ext_mul(a, b):
    a, b = _u64(a), _u64(b)
    o = np.zeros(2, np.uint64)
    lib().or_ext_mul(_p(a), _p(b), _p(o))
    return o


def ext_inv(a):
    a = _u64(a)
    o = np.zeros(2, np.uint64)
    lib().or_ext_inv(_p(a), _p(o))
    return o


# ----------------------------------------------------------------- inputs
def fill_ext(seed, n):
    out = np.empty(2 * n, np.uint64)
    lib().or_fill_ext(C.c_uint64(seed), C.c_uint64(n), _p(out))
    return out


def fill_base(seed, n):
    out = np.empty(n, np.uint64)
    lib().or_fill_base(C.c_uint64(seed), C.c_uint64(n), _p(out))
    return out


# ------------------------------------------------------------- transcript
class Transcript:
    """Stand-in transcript (splitmix64); same call order as BasicTranscript."""

    def __init__(self, label=b"test"):
        self.t = OrTranscript()
        buf = (C.c_uint8 * len(label)).from_buffer_copy(label) if label else None
        lib().or_tr_init(C.byref(self.t), buf, C.c_uint64(len(label)))

    def append_message(self, msg: bytes):
        buf = (C.c_uint8 * len(msg)).from_buffer_copy(msg)
        lib().or_tr_append_message(C.byref(self.t), buf, C.c_uint64(len(msg)))

    def append_ext(self, e):
        e = _u64(e)
        lib().or_tr_append_ext(C.byref(self.t), _p(e), C.c_uint64(e.size // 2))

    def sample(self, label: bytes):
        o = np.zeros(2, np.uint64)
        lib().or_tr_sample(C.byref(self.t), C.c_char_p(label), _p(o))
        return o

    @property
    def state(self):
        return int(self.t.h)


# ----------------------------------------------------------------- eq/MLE
def build_eq_x_r_vec(r):
    r = _u64(r)
    k = r.size // 2
    out = np.empty(2 << k, np.uint64)
    lib().or_build_eq_x_r_vec(_p(r), C.c_uint32(k), _p(out))
    return out


def eq_eval(a, b):
    a, b = _u64(a), _u64(b)
    o = np.zeros(2, np.uint64)
    lib().or_eq_eval(_p(a), _p(b), C.c_uint32(a.size // 2), _p(o))
    return o


def eq_eval_less_or_equal_than(max_idx, a, b):
    a, b = _u64(a), _u64(b)
    o = np.zeros(2, np.uint64)
    lib().or_eq_eval_less_or_equal_than(C.c_uint64(max_idx), _p(a), C.c_uint32(a.size // 2), _p(b), C.c_uint32(b.size // 2), _p(o))
    return o


def mle_evaluate(evals, is_ext, point):
    evals, point = _u64(evals), _u64(point)
    nv = point.size // 2
    assert evals.size == (2 if is_ext else 1) << nv
    o = np.zeros(2, np.uint64)
    lib().or_mle_evaluate(_p(evals), C.c_uint32(1 if is_ext else 0), C.c_uint32(nv), _p(point), _p(o))
    return o


def fix_variable(evals, is_ext, r):
    evals, r = _u64(evals), _u64(r)
    n = evals.size // (2 if is_ext else 1)
    out = np.empty(n, np.uint64)  # n/2 ext
    lib().or_fix_variable(_p(evals), C.c_uint32(1 if is_ext else 0), C.c_uint64(n), _p(r), _p(out))
    return out


def eval_wellform_address_vec(offset, scaled, r, descending=False):
    r = _u64(r)
    o = np.zeros(2, np.uint64)
    lib().or_eval_wellform_address_vec(C.c_uint64(offset), C.c_uint64(scaled), _p(r), C.c_uint32(r.size // 2), C.c_int(int(descending)), _p(o))
    return o


def eval_stacked_wellform_address_vec(r):
    r = _u64(r)
    o = np.zeros(2, np.uint64)
    lib().or_eval_stacked_wellform_address_vec(_p(r), C.c_uint32(r.size // 2), _p(o))
    return o


def eval_stacked_constant_vec(r):
    r = _u64(r)
    o = np.zeros(2, np.uint64)
    lib().or_eval_stacked_constant_vec(_p(r), C.c_uint32(r.size // 2), _p(o))
    return o


SEL_WHOLE, SEL_PREFIX, SEL_ORDERED_SPARSE, SEL_QUARK_LT = 0, 1, 2, 3


def selector_compute(kind, point, offset=0, num_instances=0, indices=(), inner_vars=0):
    point = _u64(point)
    nv = point.size // 2
    idx = _u64(np.array(list(indices), dtype=np.uint64))
    out = np.empty(2 << nv, np.uint64)
    rc = lib().or_selector_compute(C.c_int(kind), _p(point), C.c_uint32(nv), C.c_uint64(offset), C.c_uint64(num_instances),
                                   _p(idx), C.c_uint32(idx.size), C.c_uint32(inner_vars), _p(out))
    if rc:
        raise ValueError(f"or_selector_compute rc={rc}")
    return out


# --------------------------------------------------------------- sumcheck
def _mk_mles(mles):
    """mles: list of (np.uint64 array, is_ext, num_vars)."""
    arr = (OrMle * len(mles))()
    keep = []
    for i, (d, is_ext, nv) in enumerate(mles):
        d = _u64(d)
        keep.append(d)
        arr[i].data = d.ctypes.data
        arr[i].num_vars = nv
        arr[i].is_ext = 1 if is_ext else 0
    return arr, keep


def _mk_terms(terms):
    """terms: list of (coeff_ext(2,), [mle idx...])."""
    coeff = np.zeros(2 * len(terms), np.uint64)
    off = np.zeros(len(terms) + 1, np.uint32)
    idx = []
    for t, (c, ids) in enumerate(terms):
        coeff[2 * t:2 * t + 2] = _u64(c)
        off[t] = len(idx)
        idx.extend(ids)
    off[len(terms)] = len(idx)
    return coeff, off, np.array(idx, dtype=np.uint32)


def sumcheck_prove_chunked(mles, terms, num_vars, degree, transcript, consume=False):
    """CPU-baseline variant (reference decomposition: per-thread chunks folded in place + single-thread tail).
    With consume=True the ext input arrays are overwritten (Either::Right semantics)."""
    arr, keep = _mk_mles(mles)
    coeff, off, idx = _mk_terms(terms)
    rounds = np.zeros(max(num_vars, 1) * degree * 2, np.uint64)
    fin = np.zeros(len(mles) * 2, np.uint64)
    chal = np.zeros(max(num_vars, 1) * 2, np.uint64)
    rc = lib().or_sumcheck_prove_chunked_standin(arr, C.c_uint32(len(mles)), _p(coeff), _p(off), _p(idx), C.c_uint32(len(terms)),
                                                 C.c_uint32(num_vars), C.c_uint32(degree), C.byref(transcript.t), _p(rounds), _p(fin),
                                                 _p(chal), C.c_int(1 if consume else 0))
    if rc:
        raise ValueError(f"or_sumcheck_prove_chunked rc={rc}")
    return (rounds[:num_vars * degree * 2].reshape(num_vars, degree, 2), fin.reshape(-1, 2), chal[:num_vars * 2].reshape(num_vars, 2))


def sumcheck_prove(mles, terms, num_vars, degree, transcript=None, challenge_fn=None):
    """Returns (round_evals[num_vars, degree, 2], final_evals[n_mles, 2], challenges[num_vars, 2]).

    With `transcript` (a stand-in Transcript) the protocol order of SURVEY §A2 is
    used; with `challenge_fn(round, evals)->ext` an arbitrary host source."""
    arr, keep = _mk_mles(mles)
    coeff, off, idx = _mk_terms(terms)
    rounds = np.zeros(max(num_vars, 1) * degree * 2, np.uint64)
    fin = np.zeros(len(mles) * 2, np.uint64)
    chal = np.zeros(max(num_vars, 1) * 2, np.uint64)
    L = lib()
    if transcript is not None:
        rc = L.or_sumcheck_prove_standin(arr, C.c_uint32(len(mles)), _p(coeff), _p(off), _p(idx), C.c_uint32(len(terms)),
                                         C.c_uint32(num_vars), C.c_uint32(degree), C.byref(transcript.t), _p(rounds), _p(fin), _p(chal))
    else:
        def _cb(user, rnd, evals, deg, out):
            e = np.ctypeslib.as_array(evals, shape=(deg * 2,)).copy()
            r = _u64(challenge_fn(rnd, e))
            out[0] = int(r[0])
            out[1] = int(r[1])
        cb = CHALLENGE_FN(_cb)
        rc = L.or_sumcheck_prove(arr, C.c_uint32(len(mles)), _p(coeff), _p(off), _p(idx), C.c_uint32(len(terms)),
                                 C.c_uint32(num_vars), C.c_uint32(degree), cb, None, _p(rounds), _p(fin), _p(chal))
    if rc:
        raise ValueError(f"or_sumcheck_prove rc={rc}")
    return (rounds[:num_vars * degree * 2].reshape(num_vars, degree, 2), fin.reshape(-1, 2), chal[:num_vars * 2].reshape(num_vars, 2))


def extrapolate_uni_poly(eval0, evals, r):
    eval0, evals, r = _u64(eval0), _u64(evals), _u64(r)
    o = np.zeros(2, np.uint64)
    lib().or_extrapolate_uni_poly(_p(eval0), _p(evals), C.c_uint32(evals.size // 2), _p(r), _p(o))
    return o


# ------------------------------------------------------------------ tower
def interleaving_mles_to_mles(mles, num_instances, num_limbs, default):
    """mles: list of (array, is_ext); returns list of num_limbs ext arrays."""
    n = len(mles)
    arrs = [_u64(m[0]) for m in mles]
    mle_len = arrs[0].size // (2 if mles[0][1] else 1)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    is_ext = np.array([1 if m[1] else 0 for m in mles], dtype=np.uint32)
    out_len = lib().or_interleave_out_len(n, num_instances, num_limbs)
    out = np.empty(num_limbs * out_len * 2, np.uint64)
    d = _u64(default)
    lib().or_interleaving_mles_to_mles(ptrs, _p(is_ext), C.c_uint32(n), C.c_uint64(mle_len), C.c_uint64(num_instances),
                                       C.c_uint32(num_limbs), _p(d), _p(out))
    return [out[i * out_len * 2:(i + 1) * out_len * 2].copy() for i in range(num_limbs)]


def infer_tower_product_witness(num_vars, f1, f2):
    """Returns packed witness (see ceno_oracle.c) and a list of layers [[a, b], ...]."""
    f1, f2 = _u64(f1), _u64(f2)
    total = 2 * ((1 << num_vars) - 1)  # sum_{l<num_vars} 2*2^l ext
    out = np.empty(total * 2, np.uint64)
    lib().or_infer_tower_product_witness(C.c_uint32(num_vars), _p(f1), _p(f2), _p(out))
    layers = []
    for l in range(num_vars):
        n = 1 << l
        off = 2 * (n - 1)
        layers.append([out[2 * off:2 * (off + n)], out[2 * (off + n):2 * (off + 2 * n)]])
    return out, layers


def infer_tower_logup_witness(nv, p1, p2, q1, q2):
    q1, q2 = _u64(q1), _u64(q2)
    if p1 is not None:
        p1, p2 = _u64(p1), _u64(p2)
    total = 4 * ((2 << nv) - 1)
    out = np.empty(total * 2, np.uint64)
    lib().or_infer_tower_logup_witness(C.c_uint32(nv), _p(p1) if p1 is not None else None, _p(p2) if p2 is not None else None,
                                       _p(q1), _p(q2), _p(out))
    layers = []
    for l in range(nv + 1):
        n = 1 << l
        off = 4 * (n - 1)
        layers.append([out[2 * (off + z * n):2 * (off + (z + 1) * n)] for z in range(4)])
    return out, layers


def tower_create_proof(prod_wits, logup_wits, transcript):
    """prod_wits: list of (packed, n_layers); logup_wits likewise.  Returns (proof u64 array, point ext array)."""
    npd, nl = len(prod_wits), len(logup_wits)
    pk = [_u64(w[0]) for w in prod_wits]
    lk = [_u64(w[0]) for w in logup_wits]
    pl = np.array([w[1] for w in prod_wits], dtype=np.uint32)
    ll = np.array([w[1] for w in logup_wits], dtype=np.uint32)
    pp = (C.c_void_p * max(npd, 1))(*[a.ctypes.data for a in pk])
    lp = (C.c_void_p * max(nl, 1))(*[a.ctypes.data for a in lk])
    maxr = max([int(x) for x in pl] + [int(x) for x in ll]) - 1
    cap = sum((r * 3 + 2 * npd + 4 * nl) * 2 for r in range(1, maxr + 1)) + 16
    proof = np.zeros(cap, np.uint64)
    point = np.zeros(2 * (maxr + 2), np.uint64)
    plen = C.c_uint32(0)
    w = lib().or_tower_create_proof(C.c_uint32(npd), _p(pl), pp, C.c_uint32(nl), _p(ll), lp, C.byref(transcript.t),
                                    _p(proof), _p(point), C.byref(plen))
    if w < 0:
        raise ValueError(f"or_tower_create_proof rc={w}")
    return proof[:w].copy(), point[:2 * plen.value].copy()


# -------------------------------------------------------------- Poseidon2 / Merkle (a9, parity unpinned)
class P2Params(C.Structure):
    _fields_ = [("ext_rc", (C.c_uint64 * 8) * 8), ("int_rc", C.c_uint64 * 22), ("diag", C.c_uint64 * 8),
                ("mds_variant", C.c_uint32), ("pad", C.c_uint32)]


def p2_params(seed=1, mds_variant=0):
    """Deterministic PLACEHOLDER constants (the real ones are upstream-only)."""
    p = P2Params()
    vals = fill_base(0x9052 + seed, 8 * 8 + 22 + 8)
    k = 0
    for r in range(8):
        for i in range(8):
            p.ext_rc[r][i] = int(vals[k]); k += 1
    for r in range(22):
        p.int_rc[r] = int(vals[k]); k += 1
    for i in range(8):
        p.diag[i] = int(vals[k]); k += 1
    p.mds_variant = mds_variant
    return p


def poseidon2_permute(params, state):
    st = _u64(state).copy()
    lib().or_poseidon2_permute(C.byref(params), _p(st))
    return st


def merkle_commit(params, matrix, width, height):
    matrix = _u64(matrix)
    tree = np.zeros(4 * (2 * height - 1), np.uint64)
    root = np.zeros(4, np.uint64)
    rc = lib().or_merkle_commit(C.byref(params), _p(matrix), C.c_uint64(width), C.c_uint64(height), _p(tree), _p(root))
    if rc:
        raise ValueError(f"or_merkle_commit rc={rc}")
    return tree, root


# ------------------------------------------------------------------------------- rotation (f-3)
def bh_table(nv):
    tab = np.zeros(1 << nv, np.uint64)
    if lib().or_bh_table(C.c_uint32(nv), _p(tab)):
        raise ValueError("BooleanHypercube supports 5 or 6 variables")
    return tab


def rotation_next_base_mle(evals, log2):
    evals = _u64(evals)
    out = np.zeros_like(evals)
    if lib().or_rotation_next_base_mle(_p(evals), C.c_uint64(evals.size), C.c_uint32(log2), _p(out)):
        raise ValueError("rotation_next_base_mle: bad group size")
    return out


def rotation_selector(eq, subgroup_size, log2):
    eq = _u64(eq)
    out = np.zeros_like(eq)
    if lib().or_rotation_selector(_p(eq), C.c_uint64(eq.size // 2), C.c_uint32(subgroup_size), C.c_uint32(log2), _p(out)):
        raise ValueError("rotation_selector: bad arguments")
    return out


def get_rotation_points(point, log2):
    """BooleanHypercube::get_rotation_points (gkr_iop/src/gkr/booleanhypercube.rs:124-163); point: list of ext tuples."""
    from . import pyref as pr
    om = lambda x: pr.esub(pr.ONE, x)
    if log2 == 5:
        left = [pr.ZERO] + point[:4] + point[5:]
        right = [pr.ONE, point[0], om(point[1])] + point[2:4] + point[5:]
    elif log2 == 6:
        left = [pr.ZERO] + point[:5] + point[6:]
        right = [pr.ONE, om(point[0]), point[1]] + point[2:5] + point[6:]
    else:
        raise ValueError("BooleanHypercube supports 5 or 6 variables")
    return left[:len(point)], right[:len(point)]


def _alpha_pows(transcript, n):
    from . import pyref as pr
    a = transcript.sample(b"combine subset evals")
    a = (int(a[0]), int(a[1]))
    out, cur = [], pr.ONE
    for _ in range(n):
        out.append(cur)
        cur = pr.emul(cur, a)
    return out


def prove_rotation(max_num_variables, subgroup_size, log2, wit, raw_rotation_exprs, rt, transcript):
    """prove_rotation (gkr_iop/src/gkr/layer/cpu/mod.rs:249-389) restated over this oracle's pieces.
    wit: list of base-field evaluation vectors; raw_rotation_exprs: [(source_wit_id, target_wit_id)]; rt: k ext (u64 pairs).
    Returns (round_evals, evals [3n ext tuples], (left, right, origin) points)."""
    from . import pyref as pr
    n = len(raw_rotation_exprs)
    eq = build_eq_x_r_vec(rt)
    rotated = [rotation_next_base_mle(wit[src], log2) for src, _ in raw_rotation_exprs]
    selector = rotation_selector(eq, subgroup_size, log2)
    alphas = _alpha_pows(transcript, n)
    mles, terms = [], []
    for i, (_, tgt) in enumerate(raw_rotation_exprs):
        mles += [(rotated[i], False, max_num_variables), (wit[tgt], False, max_num_variables)]
        a = alphas[i]
        terms.append(([a[0], a[1]], [2 * n, 2 * i]))
        terms.append(([(-a[0]) % pr.P, (-a[1]) % pr.P], [2 * n, 2 * i + 1]))
    mles.append((selector, True, max_num_variables))
    rounds, fin, chal = sumcheck_prove(mles, terms, max_num_variables, 2, transcript=transcript)
    origin = [(int(c[0]), int(c[1])) for c in chal]
    left, right = get_rotation_points(origin, log2)
    r = origin[log2 - 1]
    evals = []
    for i, (src, _) in enumerate(raw_rotation_exprs):
        rot_e, tgt_e = (int(fin[2 * i][0]), int(fin[2 * i][1])), (int(fin[2 * i + 1][0]), int(fin[2 * i + 1][1]))
        le = mle_evaluate(wit[src], False, np.array(left, dtype=np.uint64).reshape(-1))
        le = (int(le[0]), int(le[1]))
        re = pr.emul(pr.esub(rot_e, pr.emul(pr.esub(pr.ONE, r), le)), pr.einv(r))
        evals += [le, re, tgt_e]
    transcript.append_ext(np.array(evals, dtype=np.uint64).reshape(-1))
    return rounds, evals, (left, right, origin)


def verify_rotation(max_num_variables, n, proof_rounds, evals, subgroup_size, log2, rt, transcript):
    """verify_rotation (gkr_iop/src/gkr/layer/zerocheck_layer.rs:678-790): sumcheck of claim 0 and degree 2, then
    sel(origin) * sum_i alpha_i ((1 - r) left_i + r right_i - target_i) == the sumcheck's final claim, with
    sel(origin) = rotation_selector_eval (gkr_iop/src/utils.rs:78-102).  Returns (left, right, origin) or raises."""
    from . import pyref as pr
    assert len(evals) == 3 * n
    alphas = _alpha_pows(transcript, n)
    transcript.append_message(int(max_num_variables).to_bytes(8, "little"))
    transcript.append_message(int(2).to_bytes(8, "little"))
    claim, origin = pr.ZERO, []
    for j in range(max_num_variables):
        msg = [(int(e[0]), int(e[1])) for e in proof_rounds[j]]
        transcript.append_ext(np.array(msg, dtype=np.uint64).reshape(-1))
        r = transcript.sample(b"Internal round")
        r = (int(r[0]), int(r[1]))
        claim = pr.lagrange_eval([pr.esub(claim, msg[0])] + msg, r)
        origin.append(r)
    transcript.append_ext(np.array(evals, dtype=np.uint64).reshape(-1))
    rtl = [(int(a), int(b)) for a, b in np.asarray(rt, dtype=np.uint64).reshape(-1, 2)]
    group = [int(x) for x in bh_table(log2)][:subgroup_size]
    oe, ie = pr.build_eq_x_r_vec(rtl[:log2]), pr.build_eq_x_r_vec(origin[:log2])
    sel = pr.ZERO
    for b in group:
        sel = pr.eadd(sel, pr.emul(oe[b], ie[b]))
    for x, y in zip(rtl[log2:], origin[log2:]):
        xy = pr.emul(x, y)
        sel = pr.emul(sel, pr.eadd(pr.esub(pr.esub(pr.eadd(xy, xy), x), y), pr.ONE))
    r = origin[log2 - 1]
    got = pr.ZERO
    for i in range(n):
        le, re, te = evals[3 * i], evals[3 * i + 1], evals[3 * i + 2]
        rot = pr.eadd(pr.emul(pr.esub(pr.ONE, r), le), pr.emul(r, re))
        got = pr.eadd(got, pr.emul(alphas[i], pr.esub(rot, te)))
    got = pr.emul(got, sel)
    if got != claim:
        raise ValueError("rotation verify failed: claim mismatch")
    left, right = get_rotation_points(origin, log2)
    return left, right, origin


# ------------------------------------------------------------------------------- NTT / RS-encode (a9, f-2)
def two_adic_generator(bits):
    return int(lib().or_two_adic_generator(C.c_uint32(bits)))


def ntt(data, log_n, n_cols=1, inverse=False, bitrev=False):
    """data: n_cols columns of 2^log_n base elements (column-major).  Forward: X[k] = sum_j x[j] w^(jk),
    w = two_adic_generator(log_n); bitrev: forward output / inverse input in bit-reversed order."""
    a = _u64(data).copy()
    assert a.size == n_cols << log_n
    if lib().or_ntt(_p(a), C.c_uint32(log_n), C.c_uint64(n_cols), C.c_int(int(inverse)), C.c_int(int(bitrev))):
        raise ValueError("or_ntt: bad size")
    return a


def rs_encode(msg, width, log_n, rate_log, bitrev=True):
    msg = _u64(msg)
    out = np.zeros(width << (log_n + rate_log), np.uint64)
    if lib().or_rs_encode(_p(msg), C.c_uint64(width), C.c_uint32(log_n), C.c_uint32(rate_log), _p(out), C.c_int(int(bitrev))):
        raise ValueError("or_rs_encode: bad size")
    return out


# ------------------------------------------------------------------------------- EC-sum Quark (f-3)
# CpuEccProver::create_ecc_proof (ceno_zkvm/src/scheme/cpu/mod.rs:72-316) restated: selectors, even/odd split,
# the septic-extension constraints expanded to monomials, then the generic sumcheck.  Written independently of
# ceno_b200/expr.py (explicit multiplication table instead of a symbolic polynomial class).
SEPTIC_D = 7


def _septic_table():
    """z^i * z^j = sum c z^k with z^7 = 2z + 5 (ceno_zkvm/src/scheme/septic_curve.rs:689-701)."""
    return [[([(i + j, 1)] if i + j < 7 else [(i + j - 7, 5), (i + j - 6, 2)]) for j in range(7)] for i in range(7)]


def septic_mul(a, b):
    out = [0] * 7
    tab = _septic_table()
    for i in range(7):
        for j in range(7):
            for k, c in tab[i][j]:
                out[k] = (out[k] + c * a[i] * b[j]) % P
    return out


def septic_inv(a):
    """Inverse in F_p[z]/(z^7 - 2z - 5) by solving (mult-by-a) u = 1 (Gaussian elimination mod p; test-size only)."""
    cols = []
    for j in range(7):
        e = [0] * 7
        e[j] = 1
        cols.append(septic_mul(a, e))
    m = [[cols[j][i] for j in range(7)] + [1 if i == 0 else 0] for i in range(7)]
    for c in range(7):
        piv = next(r for r in range(c, 7) if m[r][c])
        m[c], m[piv] = m[piv], m[c]
        inv = pow(m[c][c], P - 2, P)
        m[c] = [v * inv % P for v in m[c]]
        for r in range(7):
            if r != c and m[r][c]:
                f = m[r][c]
                m[r] = [(v - f * w) % P for v, w in zip(m[r], m[c])]
    return [m[i][7] for i in range(7)]


def ecc_quark_make_witness(seed, n, num_instances):
    """xs, ys, invs: 7 base arrays of 2^(n+1) each laid out like the reference's EC-sum chip: leaves in the first half,
    node (1,b) at 2^n + b = children 2b (+) 2b+1 where the Quark selector is on, a copy of child 2b elsewhere; the slope
    s[1,b] in invs.  The affine-addition identities hold for ANY pairs with x0 != x1, so random 'points' give a witness on
    which all seven constraint families vanish."""
    N = 1 << n
    rnd = fill_base(seed, 14 * num_instances).reshape(num_instances, 14)
    X = [[0] * 7 for _ in range(2 * N)]
    Y = [[0] * 7 for _ in range(2 * N)]
    S = [[0] * 7 for _ in range(2 * N)]
    for i in range(num_instances):
        X[i] = [int(v) for v in rnd[i, :7]]
        Y[i] = [int(v) for v in rnd[i, 7:]]
    on = selector_compute(3, fill_ext(seed + 1, n), 0, num_instances).reshape(-1, 2)
    sub = lambda a, b: [(u - v) % P for u, v in zip(a, b)]
    for b in range(N - 1):
        l, r = 2 * b, 2 * b + 1
        if on[b].any():
            s = septic_mul(sub(Y[l], Y[r]), septic_inv(sub(X[l], X[r])))
            x3 = sub(sub(septic_mul(s, s), X[l]), X[r])
            y3 = sub(septic_mul(s, sub(X[l], x3)), Y[l])
            X[N + b], Y[N + b], S[N + b] = x3, y3, s
        else:
            X[N + b], Y[N + b] = list(X[l]), list(Y[l])
    col = lambda M, i: np.array([row[i] for row in M], dtype=np.uint64)
    return [col(X, i) for i in range(7)], [col(Y, i) for i in range(7)], [col(S, i) for i in range(7)]


def ecc_quark_selectors(out_rt, num_instances):
    n = _u64(out_rt).size // 2
    sel_add = selector_compute(3, out_rt, 0, num_instances).reshape(-1, 2)
    eq = build_eq_x_r_vec(out_rt).reshape(-1, 2)
    sel_export = np.zeros_like(eq)
    sel_export[(1 << n) - 2] = eq_eval(out_rt, np.array([0, 0] + [1, 0] * (n - 1), dtype=np.uint64))
    sel_bypass = eq.copy()
    sel_bypass[sel_add.any(axis=1)] = 0
    sel_bypass[-1] = 0
    return sel_add.reshape(-1), sel_bypass.reshape(-1), sel_export.reshape(-1)


def ecc_quark_terms(alpha_pows, fx, fy):
    """MLE order [sel_add, sel_bypass, sel_export, s, x0, y0, x1, y1, x3, y3] (7 each after the selectors)."""
    tab = _septic_table()
    V = lambda g, i: 3 + 7 * g + i
    Sg, X0, Y0, X1, Y1, X3, Y3 = range(7)
    acc = {}

    def add_term(coef, factors):
        key = tuple(sorted(factors))
        c = acc.get(key, (0, 0))
        acc[key] = ((c[0] + coef[0]) % P, (c[1] + coef[1]) % P)

    def scale(a, c):
        return (a[0] * c % P, a[1] * c % P)

    al = [(int(a[0]), int(a[1])) for a in np.asarray(alpha_pows, dtype=np.uint64).reshape(-1, 2)]
    ai = 0
    SEL_ADD, SEL_BYP, SEL_EXP = 0, 1, 2
    # 1) s (x0 - x1) - (y0 - y1)
    for k in range(7):
        a = al[ai + k]
        for i in range(7):
            for j in range(7):
                for kk, c in tab[i][j]:
                    if kk == k:
                        add_term(scale(a, c), [SEL_ADD, V(Sg, i), V(X0, j)])
                        add_term(scale(a, P - c), [SEL_ADD, V(Sg, i), V(X1, j)])
        add_term(scale(a, P - 1), [SEL_ADD, V(Y0, k)])
        add_term(a, [SEL_ADD, V(Y1, k)])
    ai += 7
    # 2) s^2 - x0 - x1 - x3
    for k in range(7):
        a = al[ai + k]
        for i in range(7):
            for j in range(7):
                for kk, c in tab[i][j]:
                    if kk == k:
                        add_term(scale(a, c), [SEL_ADD, V(Sg, i), V(Sg, j)])
        for g in (X0, X1, X3):
            add_term(scale(a, P - 1), [SEL_ADD, V(g, k)])
    ai += 7
    # 3) s (x0 - x3) - (y0 + y3)
    for k in range(7):
        a = al[ai + k]
        for i in range(7):
            for j in range(7):
                for kk, c in tab[i][j]:
                    if kk == k:
                        add_term(scale(a, c), [SEL_ADD, V(Sg, i), V(X0, j)])
                        add_term(scale(a, P - c), [SEL_ADD, V(Sg, i), V(X3, j)])
        add_term(scale(a, P - 1), [SEL_ADD, V(Y0, k)])
        add_term(scale(a, P - 1), [SEL_ADD, V(Y3, k)])
    ai += 7
    # bypass: x3 - x0, y3 - y0
    for g3, g0 in ((X3, X0), (Y3, Y0)):
        for k in range(7):
            a = al[ai + k]
            add_term(a, [SEL_BYP, V(g3, k)])
            add_term(scale(a, P - 1), [SEL_BYP, V(g0, k)])
        ai += 7
    # export: x3 - final_x, y3 - final_y
    for g3, fin in ((X3, fx), (Y3, fy)):
        for k in range(7):
            a = al[ai + k]
            add_term(a, [SEL_EXP, V(g3, k)])
            add_term(scale(a, (P - int(fin[k])) % P), [SEL_EXP])
        ai += 7
    return [([c[0], c[1]], list(k)) for k, c in sorted(acc.items()) if c != (0, 0)]


def ecc_quark_create_proof(num_instances, xs, ys, invs, transcript):
    n = int(np.log2(xs[0].size)) - 1
    N = 1 << n
    out_rt = np.concatenate([transcript.sample(b"ecc") for _ in range(n)])
    a = transcript.sample(b"ecc_alpha")
    alpha, cur, one = (int(a[0]), int(a[1])), (1, 0), None
    pows = []
    for _ in range(49):
        pows.append(cur)
        cur = ((cur[0] * alpha[0] + 7 * cur[1] * alpha[1]) % P, (cur[0] * alpha[1] + cur[1] * alpha[0]) % P)
    sel = ecc_quark_selectors(out_rt, num_instances)
    ev = lambda v: [np.ascontiguousarray(a[0::2]) for a in v]
    od = lambda v: [np.ascontiguousarray(a[1::2]) for a in v]
    hi = lambda v: [np.ascontiguousarray(a[N:]) for a in v]
    last = N - 2
    fx, fy = [int(a[N + last]) for a in xs], [int(a[N + last]) for a in ys]
    terms = ecc_quark_terms(np.array(pows, dtype=np.uint64), fx, fy)
    groups = hi(invs) + ev(xs) + ev(ys) + od(xs) + od(ys) + hi(xs) + hi(ys)
    mles = [(s, True, n) for s in sel] + [(g, False, n) for g in groups]
    rounds, evals, rt = sumcheck_prove(mles, terms, n, 3, transcript=transcript)
    return {"zerocheck_proof": rounds, "num_instances": num_instances, "evals": evals, "rt": rt, "sum": (fx, fy),
            "out_rt": out_rt, "mles": mles, "terms": terms}
