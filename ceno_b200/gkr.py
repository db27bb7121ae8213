"""Host mirror of the gkr_iop layer / circuit API over the device path:

  Layer, LayerWitness, LayerProof          gkr_iop/src/gkr/layer.rs:69-126, gkr_iop/src/gkr/layer/sumcheck_layer.rs:23-35
  ZerocheckLayer.prove                     ZerocheckLayerProver::prove, gkr_iop/src/gkr/layer/cpu/mod.rs:99-239 (GPU: layer/gpu/mod.rs:186-290)
  LinearLayer.prove                        LinearLayerProver::prove, gkr_iop/src/gkr/layer/cpu/mod.rs:44-67
  GKRCircuit.prove                         gkr_iop/src/gkr.rs:70-117 (layers in order, claims threaded through `running_evals`)

The reference's `Layer` is driven by symbolic `Expression`s built by its circuit builder (out of scope).  What
crosses the device boundary is their monomial form, so a layer here carries its expressions already monomialised
(`ceno_b200.expr.Poly` over the witness order  witin ++ fixed ++ structural)  and evaluation references as plain positions
in the chip-wide `running_evals` table.  Everything numerical — selector eq tables, the main sumcheck, MLE evaluations —
runs on the GPU through the C ABI; this file only orders the calls and the transcript operations like the reference."""
import numpy as np

from . import api
from .expr import Poly, ext

ZEROCHECK, LINEAR = "zerocheck", "linear"


class SelectorContext:
    """gkr_iop/src/selector.rs:28-41"""

    def __init__(self, offset, num_instances, num_vars):
        self.offset, self.num_instances, self.num_vars = offset, num_instances, num_vars


class OutGroup:
    """One entry of Layer.out_sel_and_eval_exprs: a selector (kind + the structural witness id it stands for) and the
    positions in `running_evals` holding the claimed evaluations of the group's expressions (their shared point is
    the group's out point, Layer::extract_claim_and_point)."""

    def __init__(self, selector_kind, selector_wit_id, eval_positions, sparse_indices=(), inner_vars=0):
        self.selector_kind, self.selector_wit_id, self.eval_positions = selector_kind, selector_wit_id, list(eval_positions)
        self.sparse_indices, self.inner_vars = tuple(sparse_indices), inner_vars


class Layer:
    def __init__(self, name, ty, n_witin, n_fixed, n_structural_witin, exprs, out_groups, in_eval_positions, max_expr_degree=None):
        self.name, self.ty = name, ty
        self.n_witin, self.n_fixed, self.n_structural_witin = n_witin, n_fixed, n_structural_witin
        self.exprs = list(exprs)                      # Poly per output expression, grouped in out_groups order
        self.out_sel_and_eval_exprs = list(out_groups)
        self.in_eval_expr = list(in_eval_positions)   # where the evaluations of this layer's witnesses go
        self.max_expr_degree = max((p.degree() for p in self.exprs), default=0) if max_expr_degree is None else max_expr_degree
        assert sum(len(g.eval_positions) for g in out_groups) == len(self.exprs) or ty == LINEAR

    def main_sumcheck_terms(self, alpha_pows):
        """sum_g sel_g(x) * sum_j alpha_{offset(g,j)} expr_{g,j}(x) in monomial form (cpu/mod.rs:131-139; the reference
        builds it once at keygen, zerocheck_layer.rs:86-207, with the alphas as challenge ids 2..)."""
        base = self.n_witin + self.n_fixed
        total, i = Poly(), 0
        for g in self.out_sel_and_eval_exprs:
            inner = Poly()
            for _ in g.eval_positions:
                a = alpha_pows[i]
                inner = inner + self.exprs[i] * ext(int(a[0]), int(a[1]))
                i += 1
            total = total + (inner * Poly.var(base + g.selector_wit_id) if g.selector_kind is not None else inner)
        return total.terms()


class LayerProof:
    """LayerProof { main: SumcheckLayerProof { proof, evals } }"""

    def __init__(self, proof, evals):
        self.proof, self.evals = proof, evals


class ZerocheckLayer:
    @staticmethod
    def prove(dev, layer, wit, out_points, pub_io_evals, challenges, transcript, selector_ctxs, stream=None):
        assert len(challenges) == 2 * 2
        assert len(layer.out_sel_and_eval_exprs) == len(out_points) == len(selector_ctxs)
        alpha_pows = transcript.sample_and_append_challenge_pows(len(layer.exprs), b"combine subset evals")
        base = layer.n_witin + layer.n_fixed
        all_witins, owned = list(wit[:base + layer.n_structural_witin]), []
        done = set()
        for g, point, ctx in zip(layer.out_sel_and_eval_exprs, out_points, selector_ctxs):
            if g.selector_kind is None or g.selector_wit_id in done:     # first group wins, like the reference's merge by wit id
                continue
            done.add(g.selector_wit_id)
            eq = api.SelectorType.compute(dev, g.selector_kind, point, ctx.offset, ctx.num_instances, g.sparse_indices, g.inner_vars)
            all_witins[base + g.selector_wit_id] = eq
            owned.append(eq)
        num_vars = all_witins[0].num_vars
        terms = layer.main_sumcheck_terms(alpha_pows)
        rounds, evals, point = api.IOPProverState.prove(dev, all_witins, terms, num_vars, layer.max_expr_degree + 1, transcript=transcript, stream=stream)
        transcript.append_field_element_exts(evals.reshape(-1))
        for m in owned:
            m.free()
        return LayerProof(rounds, evals), point.reshape(-1)


class RotationPoints:
    """gkr_iop/src/gkr/layer/sumcheck_layer.rs: RotationPoints { left, right, origin }"""

    def __init__(self, left, right, origin):
        self.left, self.right, self.origin = left, right, origin


def get_rotation_points(point, cyclic_group_log2):
    """BooleanHypercube::get_rotation_points (gkr_iop/src/gkr/booleanhypercube.rs:124-163).  point: [k, 2] u64."""
    P = 0xFFFFFFFF00000001
    pt = [(int(a), int(b)) for a, b in np.asarray(point, dtype=np.uint64).reshape(-1, 2)]
    one_minus = lambda x: ((1 - x[0]) % P, (-x[1]) % P)
    L = cyclic_group_log2
    if L == 5:      # left (0, r0, r1, r2, r3, r5, ..), right (1, r0, 1 - r1, r2, r3, r5, ..)
        left = [(0, 0)] + pt[:4] + pt[5:]
        right = [(1, 0), pt[0], one_minus(pt[1])] + pt[2:4] + pt[5:]
    elif L == 6:    # left (0, r0, .., r4, r6, ..), right (1, 1 - r0, r1, .., r4, r6, ..)
        left = [(0, 0)] + pt[:5] + pt[6:]
        right = [(1, 0), one_minus(pt[0]), pt[1]] + pt[2:5] + pt[6:]
    else:
        raise ValueError("BooleanHypercube supports 5 or 6 variables")
    k = len(pt)
    return np.array(left[:k], dtype=np.uint64), np.array(right[:k], dtype=np.uint64)


def prove_rotation(dev, max_num_variables, rotation_cyclic_subgroup_size, rotation_cyclic_group_log2, wit, raw_rotation_exprs, rt,
                   global_challenges, transcript, stream=None):
    """prove_rotation (gkr_iop/src/gkr/layer/cpu/mod.rs:249-389): for every (source, target) witness pair prove
    rotated(source) == target on the cyclic subgroup through
        0 = sum_b sel(b) * sum_i alpha^i (rotated_i(b) - target_i(b)),
    a degree-2 sumcheck over [rotated_1, target_1, ..., selector] (the expression of zerocheck_layer.rs:88-104), then turn the
    evaluation of each rotated MLE into evaluations of its source at the left / right rotation points.
    raw_rotation_exprs: [(source_wit_id, target_wit_id)].  Returns (LayerProof(rounds, evals [3n, 2]), RotationPoints)."""
    from .expr import ext_mul
    P = 0xFFFFFFFF00000001
    n = len(raw_rotation_exprs)
    eq = api.build_eq_x_r_vec(dev, rt, stream=stream)
    rotated = [api.rotation_next_base_mle(dev, wit[src], rotation_cyclic_group_log2) for src, _ in raw_rotation_exprs]
    selector = api.rotation_selector(dev, eq, rotation_cyclic_subgroup_size, rotation_cyclic_group_log2)
    alphas = transcript.sample_and_append_challenge_pows(n, b"combine subset evals")     # challenge ids 2.. (global challenges are ids 0, 1)
    mles, terms = [], []
    for i, (_, tgt) in enumerate(raw_rotation_exprs):
        mles += [rotated[i], wit[tgt]]
        a = (int(alphas[i][0]), int(alphas[i][1]))
        terms.append(([a[0], a[1]], [2 * n, 2 * i]))
        terms.append(([(-a[0]) % P, (-a[1]) % P], [2 * n, 2 * i + 1]))
    mles.append(selector)
    rounds, evals, origin = api.IOPProverState.prove(dev, mles, terms, max_num_variables, 2, transcript=transcript, stream=stream)
    origin = origin.reshape(-1, 2)
    left, right = get_rotation_points(origin, rotation_cyclic_group_log2)
    r = (int(origin[rotation_cyclic_group_log2 - 1][0]), int(origin[rotation_cyclic_group_log2 - 1][1]))
    nrm = (r[0] * r[0] - 7 * r[1] * r[1]) % P
    ni = pow(nrm, P - 2, P)
    r_inv = (r[0] * ni % P, (-r[1]) * ni % P)
    one_minus_r = ((1 - r[0]) % P, (-r[1]) % P)
    out = []
    for i, (src, _) in enumerate(raw_rotation_exprs):
        rotated_eval, target_eval = (int(evals[2 * i][0]), int(evals[2 * i][1])), (int(evals[2 * i + 1][0]), int(evals[2 * i + 1][1]))
        le = wit[src].evaluate(left.reshape(-1))
        le = (int(le[0]), int(le[1]))
        t = ext_mul(one_minus_r, le)
        re = ext_mul(((rotated_eval[0] - t[0]) % P, (rotated_eval[1] - t[1]) % P), r_inv)     # get_rotation_right_eval_from_left
        out += [le, re, target_eval]
    out = np.array(out, dtype=np.uint64).reshape(-1, 2)
    transcript.append_field_element_exts(out.reshape(-1))
    for m in rotated + [selector, eq]:
        m.free()
    return LayerProof(rounds, out), RotationPoints(left, right, origin)


class LinearLayer:
    @staticmethod
    def prove(dev, layer, wit, out_point, transcript):
        evals = np.array([m.evaluate(out_point) for m in wit], dtype=np.uint64).reshape(-1, 2)
        transcript.append_field_element_exts(evals.reshape(-1))
        return LayerProof(np.zeros((0, 0, 2), np.uint64), evals)


class GKRCircuit:
    def __init__(self, layers, n_evaluations, final_out_evals):
        self.layers, self.n_evaluations, self.final_out_evals = list(layers), n_evaluations, list(final_out_evals)

    def prove(self, dev, circuit_wit, out_evals, pub_io_evals, challenges, transcript, selector_ctxs, stream=None):
        """-> dict(gkr_proof=[LayerProof], opening_evaluations=[(value, point, poly)], rt=[point per layer]).
        out_evals: [(point, eval)] for the output layer; running_evals is the chip-wide table the layers read and fill."""
        running = list(out_evals) + [(None, None)] * (self.n_evaluations - len(out_evals))
        proofs, rts = [], []
        for layer, wit in zip(self.layers, circuit_wit):
            out_points = [running[g.eval_positions[0]][0] for g in layer.out_sel_and_eval_exprs]     # extract_claim_and_point
            if layer.ty == ZEROCHECK:
                lp, point = ZerocheckLayer.prove(dev, layer, wit, out_points, pub_io_evals, challenges, transcript, selector_ctxs, stream=stream)
            else:
                assert len(out_points) == 1
                point = out_points[0]
                lp = LinearLayer.prove(dev, layer, wit, point, transcript)
            for pos, ev in zip(layer.in_eval_expr, lp.evals):                                           # update_claims
                running[pos] = (point, ev)
            proofs.append(lp)
            rts.append(point)
        opening = [(running[p][1], running[p][0], i) for i, p in enumerate(self.final_out_evals)]
        return {"gkr_proof": proofs, "opening_evaluations": opening, "rt": rts}
