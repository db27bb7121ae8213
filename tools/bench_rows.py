"""Secondary measurements for the other SURVEY §8 rows (not the headline bench):
  Z     : zerocheck-shaped generic sumcheck (SURVEY §8d instance "Z": 64 base witness MLEs + 4 ext selectors,
          200 monomial terms of degree <= 4, k = 20)  — a5 / the uniform-size part of a7
  TOWER : CpuTowerProver::create_proof shape: 2 product specs + 1 logup spec, 2^21-point leaves (a6 + a8)
  EQ    : build_eq_x_r, k = 24 (a3)
  C-26  : Merkle commit of 64 columns x 2^20 rows = 2^26 base elements, placeholder Poseidon2 constants (a9)
usage: python tools/bench_rows.py [--cpu]   (--cpu also times the oracle on the host for Z and TOWER)"""
import ctypes as C
import json
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceno_b200 as cb
from ceno_b200 import _lib, api, synth

P = 0xFFFFFFFF00000001


def keccak_tower_row(dev, rows_log_local, n_rec=1094, comm=None, xchg=None, barrier=None, world=1, rank=0):
    """BASELINE config #4's tower: the keccak-f chip's 1094 lookup records per row (ceno_zkvm/src/precompiles/lookup_keccakf.rs:97-101)
    over VIRTUAL leaves, 2^rows_log_local rows on this GPU (times `world` GPUs when sharded: every rank holds its rows of both
    fan-in blocks).  Returns build / prove wall-clock ms between device synchronisations (rank barriers when sharded)."""
    n_loc = 1 << rows_log_local
    g = world.bit_length() - 1
    l2m = (n_rec - 1).bit_length()
    if comm is not None:
        comm.create_arena(16 * 4 * ((1 << (l2m + rows_log_local - 1)) + (1 << 22)), xchg)
    big = dev.alloc(16 * n_loc * n_rec)
    chunk = synth.fill_ext(77 + rank, n_loc)
    pin, pinp = dev.pinned(chunk.nbytes)
    pin[:] = chunk
    for i in range(n_rec):                       # same values in every record (content does not change the cost)
        dev.h2d(big.ptr + 16 * n_loc * i, pinp, chunk.nbytes)
    dev.sync()
    recs = [cb.MultilinearExtension(dev, cb.DeviceBuffer(dev, big.ptr + 16 * n_loc * i, 16 * n_loc, owner=False), rows_log_local, True) for i in range(n_rec)]
    sync = (lambda: (dev.sync(), barrier())) if barrier else dev.sync

    def run():
        sync()
        t0 = time.perf_counter()
        tw = cb.TowerProver.from_records(dev, [cb.VirtualTowerSpec(recs, n_loc, [12345, 678], True)], comm=comm)
        sync()
        t1 = time.perf_counter()
        proof, _ = tw.create_proof(cb.StandInTranscript(b"keccak"))
        sync()
        t2 = time.perf_counter()
        tw.close()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3, proof
    run()
    b_ms, p_ms, proof = run()
    big.free()
    dev.lib.cg_host_free_pinned(dev.ctx, pinp)
    return {"workload": "BASELINE #4 shape: keccak-f lookup tower, 1094 ext records per row, virtual leaves", "n_gpus": world,
            "rows_log": rows_log_local + g, "rows_per_gpu_log": rows_log_local, "records": n_rec, "record_bytes_per_gpu": 16 * n_loc * n_rec,
            "virtual_leaf_ext_elements": 4 << (l2m + rows_log_local + g - 1), "tower_layers": l2m + rows_log_local + g,
            "build_ms": b_ms, "prove_ms": p_ms, "proof_digest": int(np.bitwise_xor.reduce(proof)),
            "timing": "host wall clock between device synchronisations" + (" and rank barriers (build includes the NVLink layer shuffles)" if comm is not None else "")}


def run_rows(dev, peak_gbs=None, cpu=False):
    """All secondary rows on `dev`; returns the dict.  peak_gbs: measured HBM peak for the `frac_of_hbm_peak` fields."""
    out = {}



    def timeit(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        dev.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dev.sync()
        return (time.perf_counter() - t0) / reps * 1e3


    # ---------------------------------------------------------------- Z
    k, nb, ne, nt, deg = 20, 64, 4, 200, 4
    n = 1 << k
    rng = random.Random(1234)
    mles = [cb.MultilinearExtension.from_evaluations_vec(dev, k, synth.fill_base(100 + i, n)) for i in range(nb)]
    mles += [cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, synth.fill_ext(900 + i, n)) for i in range(ne)]
    terms = []
    for t in range(nt):
        sel = nb + rng.randrange(ne)                       # one ext selector ...
        wit = [rng.randrange(nb) for _ in range(rng.randint(1, deg - 1))]   # ... times 1..3 base witnesses
        terms.append(([rng.randrange(P), rng.randrange(P)], [sel] + wit))
    z_ms = timeit(lambda: cb.IOPProverState.prove(dev, mles, terms, k, deg, transcript=cb.StandInTranscript(b"z")), reps=3, warm=1)
    factors = sum(len(t[1]) for t in terms)
    out["Z"] = {"k": k, "base_mles": nb, "ext_mles": ne, "terms": nt, "degree": deg, "ms": z_ms, "points_per_s": n / (z_ms * 1e-3),
                "term_factor_evals_per_s": factors * n / (z_ms * 1e-3)}
    if cpu:
        from oracle import oracle as orc
        host = [(synth.fill_base(100 + i, n), False, k) for i in range(nb)] + [(synth.fill_ext(900 + i, n), True, k) for i in range(ne)]
        t0 = time.perf_counter()
        orc.sumcheck_prove_chunked(host, terms, k, deg, orc.Transcript(b"z"))
        out["Z"]["cpu_ms"] = (time.perf_counter() - t0) * 1e3
        out["Z"]["cpu_threads"] = orc.num_threads()
    for m in mles:
        m.free()

    # ---------------------------------------------------------------- TOWER
    nvp, nvl = 22, 21
    specs = []
    for s in range(2):
        specs.append(cb.TowerProverSpec([cb.MultilinearExtension.from_evaluations_ext_vec(dev, nvp - 1, synth.fill_ext(50 + 2 * s + z, 1 << (nvp - 1))) for z in range(2)], nvp, False))
    specs.append(cb.TowerProverSpec([None, None] + [cb.MultilinearExtension.from_evaluations_ext_vec(dev, nvl, synth.fill_ext(60 + z, 1 << nvl)) for z in range(2)], nvl, True))


    def tower():
        tw = cb.TowerProver(dev, specs)
        tw.create_proof(cb.StandInTranscript(b"tower"))
        tw.close()


    def tower_build_only():
        tw = cb.TowerProver(dev, specs)
        dev.sync()
        tw.close()


    t_all, t_build = timeit(tower, reps=3, warm=1), timeit(tower_build_only, reps=3, warm=1)
    leaf_elems = 2 * (1 << nvp) + 4 * (1 << nvl)
    out["TOWER"] = {"specs": "2 product (2^21-point halves) + 1 logup (2^21 points)", "leaf_ext_elements": leaf_elems, "build_ms": t_build,
                    "build_plus_prove_ms": t_all, "layers": nvp - 1}
    for sp in specs:
        for m in sp.leaves:
            if m is not None:
                m.free()

    # ---------------------------------------------------------------- EQ
    w = synth.fill_ext(0xE9, 24)
    eqb = dev.alloc(16 << 24)
    eq_ms = timeit(lambda: cb.build_eq_x_r_vec(dev, w, out=eqb), reps=10)
    out["EQ"] = {"k": 24, "ms": eq_ms, "GBps_written": (16 << 24) / (eq_ms * 1e-3) / 1e9}
    eqb.free()

    # ---------------------------------------------------------------- C-26
    width, height = 64, 1 << 20
    vals = synth.fill_base(0x9052, 8 * 8 + 22 + 8)
    api.poseidon2_set_params(dev, vals[:64].reshape(8, 8), vals[64:86], vals[86:94], 0)
    mat = dev.to_device(synth.fill_base(4242, width * height))       # column-major: column c at [c*height, (c+1)*height)
    tree = dev.alloc(32 * (2 * height - 1))


    def commit():
        root = np.zeros(4, np.uint64)
        dev.check(dev.lib.cg_merkle_commit(dev.ctx, C.c_void_p(mat.ptr), width, height, 1, C.c_void_p(tree.ptr), root.ctypes.data_as(C.c_void_p), None))


    c_ms = timeit(commit, reps=5)
    perms = height * (width // 4) + height - 1
    out["C-26"] = {"width": width, "height": height, "elements": width * height, "ms": c_ms, "permutations": perms,
                   "Mperm_per_s": perms / (c_ms * 1e-3) / 1e6, "GBps_read": 8 * width * height / (c_ms * 1e-3) / 1e9,
                   "note": "placeholder constants; leaf hash + Merkle only (no RS encode)"}
    tree.free()

    # ---------------------------------------------------------------- RS-26 / COMMIT-26 (BASELINE config #5: 2^26-element batch, eq-build + fold + commit)
    log_n, rate_log = 20, 1
    code = dev.alloc(8 * (width << (log_n + rate_log)))


    def encode():
        dev.check(dev.lib.cg_rs_encode(dev.ctx, C.c_void_p(mat.ptr), width, log_n, rate_log, C.c_void_p(code.ptr), api.NTT_BITREV, None))


    e_ms = timeit(encode, reps=5)
    n_code = width << (log_n + rate_log)
    # algorithmic bytes: pass 1 reads the message (zero padding is implicit) and writes the code, pass 2 reads + writes the code
    alg = 8 * (width << log_n) + 3 * 8 * n_code
    out["RS-26"] = {"width": width, "log_n": log_n, "rate_log": rate_log, "ms": e_ms, "passes": 2, "algorithmic_bytes": alg,
                    "GBps": alg / (e_ms * 1e-3) / 1e9, "butterflies_per_s": (n_code // 2) * (log_n + rate_log) / (e_ms * 1e-3)}
    tree2 = dev.alloc(32 * (2 * (height << rate_log) - 1))


    def full_commit():
        encode()
        root = np.zeros(4, np.uint64)
        dev.check(dev.lib.cg_merkle_commit(dev.ctx, C.c_void_p(code.ptr), width, height << rate_log, 1, C.c_void_p(tree2.ptr), root.ctypes.data_as(C.c_void_p), None))


    fc_ms = timeit(full_commit, reps=3)
    out["COMMIT-26"] = {"ms": fc_ms, "note": "RS-encode (rate 1/2, bit-reversed rows) + Poseidon2 leaf hash + Merkle over the 64 x 2^21 codeword matrix; placeholder constants"}
    # the other two legs of config #5 on the same 64 x 2^20 batch: eq-build (k = 20) and one fix_variable of all 64 columns
    w20 = synth.fill_ext(0xE9, log_n)
    eq20 = dev.alloc(16 << log_n)
    eq20_ms = timeit(lambda: cb.build_eq_x_r_vec(dev, w20, out=eq20), reps=10)
    descs = (_lib.CgMleDesc * width)(*[_lib.CgMleDesc(mat.ptr + 8 * height * cidx, height, log_n, 0) for cidx in range(width)])
    fold_out = dev.alloc(16 * (height // 2) * width)
    outs = (C.c_void_p * width)(*[fold_out.ptr + 16 * (height // 2) * cidx for cidx in range(width)])
    r_fold = synth.fill_ext(0xF01D, 1)


    def fold_all():
        dev.check(dev.lib.cg_fix_variable(dev.ctx, descs, width, r_fold.ctypes.data_as(C.c_void_p), outs, None))


    f_ms = timeit(fold_all, reps=10)
    fold_bytes = width * (8 * height + 16 * (height // 2))
    out["BATCH-26"] = {"config": "BASELINE #5: 2^26-element MLE batch = 64 base columns x 2^20 rows", "eq_build_ms": eq20_ms,
                       "fold_ms": f_ms, "fold_GBps": fold_bytes / (f_ms * 1e-3) / 1e9, "rs_encode_ms": e_ms, "commit_total_ms": fc_ms,
                       "total_ms": eq20_ms + f_ms + fc_ms}

    # ---------------------------------------------------------------- FOLD-24: fix_variable of the T3-24 MLEs alone (a2)
    k24 = 24
    fa = dev.alloc(3 * (16 << k24))
    fo = dev.alloc(3 * (16 << (k24 - 1)))
    d3 = (_lib.CgMleDesc * 3)(*[_lib.CgMleDesc(fa.ptr + (16 << k24) * i, 1 << k24, k24, 1) for i in range(3)])
    o3 = (C.c_void_p * 3)(*[fo.ptr + (16 << (k24 - 1)) * i for i in range(3)])
    dev.check(dev.lib.cg_fix_variable(dev.ctx, d3, 3, r_fold.ctypes.data_as(C.c_void_p), o3, None))
    f24_ms = timeit(lambda: dev.check(dev.lib.cg_fix_variable(dev.ctx, d3, 3, r_fold.ctypes.data_as(C.c_void_p), o3, None)), reps=10)
    f24_bytes = 3 * (16 << k24) + 3 * (16 << (k24 - 1))
    out["FOLD-24"] = {"k": k24, "mles": 3, "ms": f24_ms, "algorithmic_bytes": f24_bytes, "GBps": f24_bytes / (f24_ms * 1e-3) / 1e9}
    fa.free(); fo.free()
    if peak_gbs:
        out["EQ"]["frac_of_hbm_peak"] = out["EQ"]["GBps_written"] / peak_gbs
        out["FOLD-24"]["frac_of_hbm_peak"] = out["FOLD-24"]["GBps"] / peak_gbs
        out["RS-26"]["frac_of_hbm_peak"] = out["RS-26"]["GBps"] / peak_gbs
        out["BATCH-26"]["fold_frac_of_hbm_peak"] = out["BATCH-26"]["fold_GBps"] / peak_gbs
        out["C-26"]["frac_of_hbm_peak"] = out["C-26"]["GBps_read"] / peak_gbs
        out["C-26"]["bound"] = "integer issue (Poseidon2: ~14 k instructions per permutation), not HBM"
    for b in (mat, code, tree2, eq20, fold_out):
        b.free()
    # ---------------------------------------------------------------- KECCAK: config #4's lookup tower, one GPU's 2^17 rows
    try:
        out["KECCAK-17"] = keccak_tower_row(dev, 17)
    except Exception as e:  # noqa: BLE001
        out["KECCAK-17"] = {"error": repr(e)}
    return out


if __name__ == "__main__":
    _dev = cb.Device(0)
    print(json.dumps(run_rows(_dev, cpu="--cpu" in sys.argv)))
    _dev.close()
