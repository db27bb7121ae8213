"""ceno_b200 — B200-native (sm_100a) device backend for the GKR-sumcheck hot path of scroll-tech/ceno.

The product is the C ABI in include/ceno_b200.h implemented by ceno_b200/csrc (hand-written CUDA);
this package is the thin host-side mirror of the reference's interface used by tests and bench.py.
"""
from .api import (CenoB200Error, ChipScheduler, EccQuarkProver, ChipTask, Comm, Device, DeviceBuffer, EqPolynomial, IOPProverState, MultilinearExtension, SelectorType,  # noqa: F401
                  StandInTranscript, Stream, TowerProver, TowerProverSpec, build_eq_x_r_vec, prove_sharded, wit_infer_by_monomial_expr,
                  BasefoldCommitment, BasefoldParams, basefold_batch_open, poseidon2_set_params, VirtualTowerSpec)
from . import chip, expr, gkr  # noqa: E402,F401
