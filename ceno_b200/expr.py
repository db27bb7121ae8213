"""Host-side monomial-form expressions over device MLEs — the role VirtualPolynomialsBuilder / Expression /
monomialize play on the Rust side (EXTERNAL multilinear_extensions; used by CpuEccProver::create_ecc_proof,
ceno_zkvm/src/scheme/cpu/mod.rs:96-262, and by build_static_expression, gkr_iop/src/gkr/layer/zerocheck_layer.rs:86-207).
Only bookkeeping happens here: a polynomial is {sorted tuple of MLE indices: ext coefficient}; the device evaluates it
(cg_sumcheck_* take exactly this term table)."""
P = 0xFFFFFFFF00000001
W = 7   # GoldilocksExt2 = F_p[X]/(X^2 - 7)


def ext(a, b=0):
    return (int(a) % P, int(b) % P)


def ext_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def ext_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def ext_mul(a, b):
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


class Poly:
    __slots__ = ("t",)

    def __init__(self, t=None):
        self.t = t or {}

    @staticmethod
    def var(i):
        return Poly({(i,): ext(1)})

    @staticmethod
    def const(c):
        c = c if isinstance(c, tuple) else ext(c)
        return Poly({(): c} if c != (0, 0) else {})

    def _acc(self, key, c):
        v = ext_add(self.t.get(key, (0, 0)), c)
        if v == (0, 0):
            self.t.pop(key, None)
        else:
            self.t[key] = v

    def __add__(self, o):
        r = Poly(dict(self.t))
        for k, c in o.t.items():
            r._acc(k, c)
        return r

    def __neg__(self):
        return Poly({k: ext_neg(c) for k, c in self.t.items()})

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, o):
        if not isinstance(o, Poly):
            o = Poly.const(o)
        r = Poly()
        for ka, ca in self.t.items():
            for kb, cb in o.t.items():
                r._acc(tuple(sorted(ka + kb)), ext_mul(ca, cb))
        return r

    def degree(self):
        return max((len(k) for k in self.t), default=0)

    def terms(self):
        """[(coeff [c0, c1], [mle indices])] in a deterministic order — the table cg_sumcheck_* take."""
        return [([c[0], c[1]], list(k)) for k, c in sorted(self.t.items())]


class SymbolicSepticExtension:
    """Seven expressions = one element of F[z]/(z^7 - 2z - 5) (ceno_zkvm/src/scheme/septic_curve.rs:681-705)."""

    def __init__(self, limbs):
        assert len(limbs) == 7, "exprs length must be 7"
        self.l = list(limbs)

    def __add__(self, o):
        return SymbolicSepticExtension([a + b for a, b in zip(self.l, o.l)])

    def __sub__(self, o):
        return SymbolicSepticExtension([a - b for a, b in zip(self.l, o.l)])

    def __mul__(self, o):
        res = [Poly() for _ in range(7)]
        for i in range(7):
            for j in range(7):
                term = self.l[i] * o.l[j]
                idx = i + j
                if idx < 7:
                    res[idx] = res[idx] + term
                else:                       # z^7 = 2z + 5
                    idx -= 7
                    res[idx] = res[idx] + term * 5
                    res[idx + 1] = res[idx + 1] + term * 2
        return SymbolicSepticExtension(res)

    def to_exprs(self):
        return list(self.l)


def terms_from_layer_json(layer, challenges):
    """Import a real circuit's term table (BASELINE config #3): `Layer::main_sumcheck_expression_monomial_terms`
    (gkr_iop/src/gkr/layer/zerocheck_layer.rs:86-207) dumped with the schema `ceno_b200/layer/1` (tests/golden/SCHEMA.md).
    `challenges`: ext values by challenge id (ids 0, 1 = the global challenges, 2.. = the alpha powers).
    Returns (terms in the cg_sumcheck_* layout, n_mles, degree): the scalar of every monomial is resolved against
    `challenges`, equal products are merged (exact field addition: term order never changes a bit)."""
    if layer.get("schema") != "ceno_b200/layer/1":
        raise ValueError("unknown layer schema")
    n_mles = layer["n_witin"] + layer["n_fixed"] + layer["n_structural_witin"]
    poly = Poly()
    for t in layer["monomial_terms"]:
        sc = t["scalar"]
        if isinstance(sc, dict):                       # Expression::Challenge(id, pow, scalar, offset)
            c = challenges[sc["challenge"]]
            c = (int(c[0]), int(c[1]))
            v = ext(1)
            for _ in range(sc.get("pow", 1)):
                v = ext_mul(v, c)
            if "scalar" in sc:
                v = ext_mul(v, ext(*sc["scalar"]))
            if "offset" in sc:
                v = ext_add(v, ext(*sc["offset"]))
        else:
            v = ext(*sc)
        prod = tuple(sorted(int(i) for i in t["product"]))
        if any(i >= n_mles for i in prod):
            raise ValueError("monomial references a witness id outside witin ++ fixed ++ structural")
        poly._acc(prod, v)
    return poly.terms(), n_mles, max(layer.get("max_expr_degree", 0) + 1, poly.degree())
