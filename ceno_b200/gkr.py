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
