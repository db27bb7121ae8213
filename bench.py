#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (BASELINE.json): a degree-3 sumcheck
P = eq(w,x)*A(x)*B(x) over GoldilocksExt2 on a 2^24-variable instance ("T3-24"), rounds + folds,
reported as Goldilocks field-ops/s (99 base-field ops per pair per round), plus
hypercube points/s and rounds/s, the HBM-roofline fraction of the dominant kernel measured live
with CUDA events, and a CPU baseline (the oracle port of the reference algorithm) on the host cores.

  python bench.py --gpus N --steps K --warmup W        (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                 CPU arm: the reference algorithm's port

A "step" is one full sumcheck (k rounds of evaluate+fold) over the resident instance.
`value`  : inputs resident in HBM, transcript on the host behind the C-ABI callback (the
           reference's flow: it hands &mut BasicTranscript to the device crate).
`e2e`    : same call with HOST buffers — H2D of A and B from pinned memory, prove, results back on
           the host — inside the timed region.
--eq virtual (default): eq(w, .) is handed over as its point (cg_mle_desc kind CG_MLE_EQ) and the large
           rounds run the split-eq kernel (no eq stream, no eq fold); --eq table: eq is a resident
           2^k table built by cg_build_eq (the `table_eq` sub-object always reports that variant too).
           Both produce the same proof bit for bit.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OPS_PER_PAIR = 99            # 12 ext add + 6 ext mul (eval) + 3*(sub+mul+add) (fold) in base-field ops
SEED_A, SEED_B, SEED_W = 0xC0FFEE ^ 1, 0xC0FFEE ^ 2, 0xE9


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the GPU work of the bench."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index=0):
        self.index, self.samples, self.reasons, self.stop_flag, self.t, self.max_mhz = index, [], set(), False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # older bindings
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(mhz), int(util)))
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")], "samples": 0}
        self.stop_flag = True
        self.t.join(timeout=2)
        sm = [m for m, _ in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm), "sm_mhz_min": min(sm) if sm else None}


def workload_config(k, n_gpus):
    """`config` of the JSON line: identical in the GPU arm and the `--impl reference` arm (same workload, same sharding request)."""
    return {"workload": f"T3-{k}: eq(w,x)*A(x)*B(x), 2^{k}-point hypercube, degree 3, GoldilocksExt2 (3 ext MLEs x {(16 << k) >> 20} MiB)",
            "k": k, "degree": 3, "n_mles": 3, "n_gpus": n_gpus,
            "l2": f"inputs {(3 * 16 << k) >> 20} MiB per job > L2 126 MB (no flush needed)"}


# ----------------------------------------------------------------------------------- CPU arm
def cpu_run(k, threads=None):
    """One T3-k step on the host cores with the oracle port: eq-build + sumcheck (rounds + folds)."""
    from oracle import oracle as orc
    n = 1 << k
    w = orc.fill_ext(SEED_W, k)
    a, b = orc.fill_ext(SEED_A, n), orc.fill_ext(SEED_B, n)
    eq = orc.build_eq_x_r_vec(w)
    t0 = time.perf_counter()
    # reference decomposition: per-thread chunks folded in place (inputs consumed), single-thread tail
    proof = orc.sumcheck_prove_chunked([(eq, True, k), (a, True, k), (b, True, k)], [([1, 0], [0, 1, 2])], k, 3, orc.Transcript(b"bench"), consume=True)
    return time.perf_counter() - t0, proof


def assert_same_proof(got, want, what):
    """Bit-exact comparison of (round_evals, final_evals, challenges) against the oracle's proof of the same instance."""
    names = ("round polynomials", "final evaluations", "challenges")
    for g, x, nm in zip(got, want, names):
        if not np.array_equal(np.asarray(g, dtype=np.uint64).reshape(-1), np.asarray(x, dtype=np.uint64).reshape(-1)):
            raise AssertionError(f"PARITY FAILURE ({what}): {nm} differ from the oracle")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    k = args.cpu_k
    for _ in range(max(args.warmup, 1) if args.warmup else 0):
        cpu_run(min(k, 18))
    ts = [cpu_run(k)[0] for _ in range(args.steps)]
    t = float(np.mean(ts))
    val = OPS_PER_PAIR * (1 << k) / t / 1e9
    cores = orc.num_threads()
    line = {
        "impl": "reference", "metric": "sumcheck Gfield-ops/s", "value": val, "unit": "Gfield-ops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64 (Goldilocks, ext2)", "data": "synthetic",
        "config": workload_config(args.k, args.gpus),
        "config_detail": {"sample": f"CPU arm times T3-{k} per step", "parallelism": f"OpenMP {cores} threads on the host"},
        "points_per_s": (1 << k) / t, "rounds_per_s": k / t,
        "cpu_baseline": {"value": val, "unit": "Gfield-ops/s", "cores": cores, "kind": "port",
                         "sample": f"T3-{k} full sumcheck (2^{k} points, {k} rounds) per step; oracle port with the reference's decomposition, OpenMP {cores} threads; the reference is Rust (Rayon) and cannot be built here"},
        "e2e": {"value": val, "unit": "Gfield-ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------- GPU arm
def gpu_arm(args):
    import torch
    import ceno_b200 as cb
    from ceno_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return gpu_arm_sharded(args, rank, world, local)

    dev = cb.Device(local)
    lib = dev.lib
    k, deg = args.k, 3
    n = 1 << k
    stream = torch.cuda.Stream()
    sh = stream.cuda_stream

    # ---- synthetic inputs, pinned on the host (e2e source) and resident on the device
    nbytes = 16 * n
    a_h, a_hp = dev.pinned(nbytes)
    b_h, b_hp = dev.pinned(nbytes)
    synth.fill_ext(SEED_A, n, out=a_h)
    synth.fill_ext(SEED_B, n, out=b_h)
    w = synth.fill_ext(SEED_W, k)
    a_d, b_d, eq_d = dev.alloc(nbytes), dev.alloc(nbytes), dev.alloc(nbytes)
    dev.h2d(a_d.ptr, a_hp, nbytes, sh)
    dev.h2d(b_d.ptr, b_hp, nbytes, sh)
    A = cb.MultilinearExtension(dev, a_d, k, True)
    B = cb.MultilinearExtension(dev, b_d, k, True)
    EQ = cb.build_eq_x_r_vec(dev, w, stream=sh, out=eq_d)
    stream.synchronize()
    virt = args.eq == "virtual"
    EQV = cb.EqPolynomial(dev, w)
    terms = [([1, 0], [0, 1, 2])]
    mles_table, mles = [EQ, A, B], ([EQV, A, B] if virt else [EQ, A, B])

    def step(device_challenger=False, flags=0, which=None):
        return cb.IOPProverState.prove(dev, mles if which is None else which, terms, k, deg, transcript=cb.StandInTranscript(b"bench"),
                                       flags=flags, device_challenger=device_challenger, stream=sh)

    def e2e_step():
        dev.h2d(a_d.ptr, a_hp, nbytes, sh)
        dev.h2d(b_d.ptr, b_hp, nbytes, sh)
        if not virt:
            cb.build_eq_x_r_vec(dev, w, stream=sh, out=eq_d)
        return step()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, out

    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(args.warmup):
        ref_out = step()
        step(True)
    l0 = dev.launch_count()
    ms, out = timed(step, args.steps)
    launches = dev.launch_count() - l0
    ms_dev, out_dev = timed(lambda: step(True), args.steps)
    ms_tab, out_tab = timed(lambda: step(which=mles_table), args.steps)          # resident eq table, streamed and folded
    ms_tab_dev, _ = timed(lambda: step(True, which=mles_table), args.steps)
    for g, d in zip(out, out_tab):
        assert np.array_equal(g, d), "virtual-eq and table-eq proofs disagree"
    # EQ-k (build_eq_x_r alone), timed separately
    ms_eq, _ = timed(lambda: cb.build_eq_x_r_vec(dev, w, stream=sh, out=eq_d), max(args.steps, 5))
    for g, d in zip(out, out_dev):
        assert np.array_equal(g, d), "host-transcript and device-challenger runs disagree"

    # ---- per-kernel roofline, measured live with CUDA events on the launching stream (CG_SC_PROFILE)
    prof = []
    for _ in range(max(3, min(args.steps, 10))):
        step(True, flags=4)
        prof.append(dev.profile_last())
    prof = np.mean(np.array(prof), axis=0)       # ms per round
    # Dominant kernel: the fused fix_variable + next-round evaluation of the streaming rounds.
    #   virtual eq: veq_tma_kernel<FOLD=1>, rounds 1 .. J-1 (J = k - 18), streams m = 2 MLEs (A, B);
    #   table eq  : tower_round_kernel<FOLD=1>, rounds 1 .. 5, streams m = 3 MLEs.
    # Algorithmic bytes of a fused launch over inputs of n_in elements: read m*16*n_in, write half.
    s = 16
    m = 2 if virt else 3
    # first round NOT run by the dominant kernel: the tail launch shows up as the first round whose time jumps back up
    # (the streaming rounds shrink geometrically; a persistent launch puts all of them into round 1's events)
    last = min(6, k)
    if virt:
        last = next((j for j in range(2, k) if prof[j] > 2 * prof[j - 1] + 0.01), k - 16)
    rounds = list(range(1, max(last, 2)))
    fused_bytes = [1.5 * m * s * (1 << (k - j + 1)) for j in rounds]
    fused_ms = [float(prof[j]) for j in rounds]
    peak, peak_src = read_peaks()
    ach_all = sum(fused_bytes) / (sum(fused_ms) * 1e-3) / 1e9
    # one persistent launch runs all these rounds (veq_persist_kernel): the events bracket the launch, later rounds read ~0
    persistent = len(fused_ms) > 1 and fused_ms[1] < 0.01
    if persistent:   # rounds 2.. are empty event pairs (2.7 us of event latency each, no kernel): the launch is round 1's bracket
        fused_ms = [fused_ms[0]] + [0.0] * (len(fused_ms) - 1)
        ach_all = sum(fused_bytes) / (fused_ms[0] * 1e-3) / 1e9
    ach_top = (sum(fused_bytes) if persistent else fused_bytes[0]) / (fused_ms[0] * 1e-3) / 1e9
    r0_bytes = m * s * n
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the same launches from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        tk = tj.get(("veq_persist_kernel" if persistent else "veq_tma_kernel<FOLD=1>") if virt else "tower_round_kernel<FOLD=1>")
        if tk and tk.get("k") == k:
            traffic = float(np.mean(tk["bytes_per_launch"]))
    except Exception:
        pass
    roofline = {
        "bound": "hbm",
        "kernel": ("veq_persist_kernel / veq_tma_kernel<FOLD=1> (split-eq, claim-derived: fused fix_variable + next-round evaluation, TMA-staged; streams A and B only; "
                   "one persistent cooperative launch for rounds 1..k-17 unless CG_VEQ_PERSIST=0)" if virt
                   else "tower_round_kernel<FOLD=1> (fused fix_variable + next-round evaluation; streams eq, A, B)"),
        "achieved": ach_all, "peak": peak, "unit": "GB/s", "frac": ach_all / peak, "peak_source": peak_src,
        "traffic": traffic,
        "bytes_definition": (f"1.5*m*16*n_in per round with m = {m} streamed MLEs (SURVEY §8d), summed over rounds 1..{rounds[-1]} — ONE persistent launch per step" if persistent else
                             f"per launch 1.5*m*16*n_in with m = {m} streamed MLEs (SURVEY §8d); achieved/traffic are averages over the {len(rounds)} launches of this kernel per step"),
        "achieved_per_launch_avg_bytes": float(sum(fused_bytes)) if persistent else float(np.mean(fused_bytes)),
        "launches_per_step": 1 if persistent else len(rounds), "algorithmic_bytes_per_step": sum(fused_bytes), "ms_per_step_in_kernel": sum(fused_ms),
        "top_launch": {"round": 1, "bytes": sum(fused_bytes) if persistent else fused_bytes[0], "ms": fused_ms[0], "achieved": ach_top, "frac": ach_top / peak,
                       "persistent": persistent},
        "round0_eval": {"bytes": r0_bytes, "ms": float(prof[0]), "achieved": r0_bytes / (float(prof[0]) * 1e-3) / 1e9},
        "eq_build": {"bytes": 16 * n, "ms": ms_eq, "achieved": 16 * n / (ms_eq * 1e-3) / 1e9},
        "round_ms": [round(float(x), 5) for x in prof],
    }
    if virt:   # what SURVEY §8d's 3-MLE table formulation would have to move in the same time (the eq stream this kernel avoids)
        roofline["table_formulation_equiv"] = {"bytes_per_step": 1.5 * sum(fused_bytes), "achieved": 1.5 * ach_all, "frac": 1.5 * ach_all / peak,
                                               "note": "SURVEY §8d counts m = 3 (eq materialised); the split-eq kernel never reads or writes it"}

    # ---- e2e: host buffers, copies inside the timed region
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    ms_e2e, out_e2e = timed(e2e_step, max(1, min(args.steps, 5)))
    for g, d in zip(out, out_e2e):
        assert np.array_equal(g, d)
    clk = clocks.stop()

    # ---- CPU baseline on this box's host cores (bounded sample)
    from oracle import oracle as orc
    cpu_run(min(args.cpu_k, 16))
    t_cpu, cpu_proof = cpu_run(args.cpu_k)
    parity_checked = None
    if args.cpu_k == k:   # the oracle proved the very instance the GPU timed: compare every output bit for bit
        for nm, o in (("host transcript", out), ("device challenger", out_dev), ("table eq", out_tab), ("e2e from host buffers", out_e2e)):
            assert_same_proof(o, cpu_proof, f"T3-{k}, {nm}")
        parity_checked = f"oracle k={k}: round polynomials, final evaluations and challenges bit-exact for the virtual-eq, table-eq, device-challenger and e2e runs"
    cpu_val = OPS_PER_PAIR * (1 << args.cpu_k) / t_cpu / 1e9
    cores = orc.num_threads()

    # ---- the other SURVEY §8 rows on the same box (secondary keys; never allowed to break the headline line)
    rows = None
    if args.rows:
        try:
            for b in (a_d, b_d, eq_d):
                b.free()
            dev.lib.cg_pool_trim(dev.ctx)
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_rows
            rows = bench_rows.run_rows(dev, peak_gbs=peak)
        except Exception as e:  # noqa: BLE001
            rows = {"error": repr(e)}

    ops = OPS_PER_PAIR * n
    value = ops / (ms * 1e-3) / 1e9
    line = {
        "metric": "sumcheck Gfield-ops/s", "value": value, "unit": "Gfield-ops/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64 (Goldilocks, ext2)", "data": "synthetic",
        "config": workload_config(k, 1),
        "config_detail": {"parallelism": "1 GPU",
                          "eq": ("virtual: eq(w,.) passed as its point (CG_MLE_EQ), split-eq kernels" if virt else "table: resident 2^k ext table"),
                          "transcript": "stand-in sponge on the host behind cg_challenge_cb (reference flow); Poseidon2 constants are upstream-only"},
        "points_per_s": n / (ms * 1e-3), "rounds_per_s": k / (ms * 1e-3),
        "device_challenger": {"ms_per_step": ms_dev, "value": ops / (ms_dev * 1e-3) / 1e9, "unit": "Gfield-ops/s",
                              "note": "same kernels, stand-in challenger on the device: no host round trip per round"},
        "table_eq": {"ms_per_step": ms_tab, "value": ops / (ms_tab * 1e-3) / 1e9, "unit": "Gfield-ops/s", "device_challenger_ms": ms_tab_dev,
                     "note": "same instance with eq as a resident 2^k table (cg_build_eq) streamed and folded like A and B; identical proof"},
        "e2e": {"value": ops / (ms_e2e * 1e-3) / 1e9, "unit": "Gfield-ops/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * nbytes + 16 * k, "d2h_bytes_per_step": 16 * (k * deg + 3 + k),
                "note": "H2D of A,B from pinned host memory" + ("" if virt else " + eq-build") + " + prove through cg_sumcheck_prove, results on the host"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_val, "unit": "Gfield-ops/s", "cores": cores, "kind": "port", "ms": t_cpu * 1e3,
                         "sample": f"T3-{args.cpu_k} full sumcheck ({'the whole workload' if args.cpu_k == k else f'1/{1 << (k - args.cpu_k)} of its points'}), one run; oracle port with the reference's decomposition (per-thread chunks folded in place, single-thread tail), OpenMP {cores} threads"},
        "clocks": clk,
        "parity_checked": parity_checked,
        "rows": rows,
    }
    print(json.dumps(line))
    dev.close()


def gpu_arm_sharded(args, rank, world, local_rank):
    """bench.py's N>1 arm: T3-k strong scaling — each rank owns a 1/N slice of the same 2^k instance — plus one
    weak-scaling measurement (T3-(k + log2 N): every GPU holds a 2^k slice).  Rank 0 checks the proof bit for bit against
    the oracle's proof of the same instance and reports the roofline of its dominant kernel and the CPU baseline."""
    import torch
    import torch.distributed as dist
    from ceno_b200 import api as cb
    from ceno_b200 import synth
    from ceno_b200.dist import eq_slice_scalar

    deg = 3
    g = world.bit_length() - 1
    assert 1 << g == world, "number of GPUs must be a power of two"
    dev = cb.Device(local_rank)
    seed_a, seed_b, seed_w = SEED_A, SEED_B, SEED_W
    terms = [([1, 0], [0, 1, 2])]

    def xchg(blob):
        outs = [None] * world
        dist.all_gather_object(outs, blob)
        return outs
    comm = cb.Comm(dev, rank, world, xchg, barrier=dist.barrier)
    stream = torch.cuda.Stream()
    sh = stream.cuda_stream

    def timed(fn, steps):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            o = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)     # max over ranks
        dist.barrier()
        return float(t.item()), o

    def instance(k):
        """This rank's slices of T3-k, pinned on the host (the e2e source) and resident on the device."""
        k_local = k - g
        n_local = 1 << k_local
        nbytes = 16 * n_local
        w = synth.fill_ext(seed_w, k)
        a_h, a_hp = dev.pinned(nbytes)
        b_h, b_hp = dev.pinned(nbytes)
        synth.fill_ext(seed_a, n_local, start=rank * n_local, out=a_h)
        synth.fill_ext(seed_b, n_local, start=rank * n_local, out=b_h)
        a_d, b_d = dev.alloc(nbytes), dev.alloc(nbytes)
        dev.h2d(a_d.ptr, a_hp, nbytes)
        dev.h2d(b_d.ptr, b_hp, nbytes)
        dev.sync()
        A = cb.MultilinearExtension(dev, a_d, k_local, True)
        B = cb.MultilinearExtension(dev, b_d, k_local, True)
        if getattr(args, "eq", "virtual") == "table":
            eq_lo = cb.build_eq_x_r_vec(dev, w[:2 * k_local])
            s = eq_slice_scalar(w[2 * k_local:], rank)
            EQ = cb.wit_infer_by_monomial_expr(dev, [eq_lo], [(list(s), [0])], k_local)      # eq slice = scalar * eq(w_low, .)
            eq_lo.free()
        else:   # eq handed over as its (global) point: split-eq rounds, no eq stream; the rank factor is derived by the library
            EQ = cb.EqPolynomial(dev, w, num_vars=k_local)

        def step(device_challenger=False, flags=0):
            return cb.prove_sharded(dev, comm, [EQ, A, B], terms, k, deg, cb.StandInTranscript(b"bench"),
                                    device_challenger=device_challenger, stream=sh, flags=flags)

        def e2e_step():   # every rank uploads its own slices over its own PCIe link, then the sharded prove
            dev.h2d(a_d.ptr, a_hp, nbytes, sh)
            dev.h2d(b_d.ptr, b_hp, nbytes, sh)
            return step()

        def free():
            a_d.free(); b_d.free()
            if hasattr(EQ, "free"):
                EQ.free()
            dev.lib.cg_host_free_pinned(dev.ctx, a_hp)
            dev.lib.cg_host_free_pinned(dev.ctx, b_hp)
        return step, e2e_step, nbytes, free

    k = args.k
    k_local = k - g
    step, e2e_step, nbytes, free_inst = instance(k)
    clocks = None
    if rank == 0:
        clocks = ClockSampler(local_rank)
        clocks.start()
    for _ in range(args.warmup):
        out = step()
        step(True)
    l0 = dev.launch_count()
    ms, out = timed(step, args.steps)
    launches = dev.launch_count() - l0
    ms_dev, out_dev = timed(lambda: step(True), args.steps)
    assert all(np.array_equal(x, y) for x, y in zip(out, out_dev))
    e2e_step()
    ms_e2e, out_e2e = timed(e2e_step, max(1, min(args.steps, 5)))
    assert all(np.array_equal(x, y) for x, y in zip(out, out_e2e))
    # per-round device time of this rank's kernels (CUDA events on the launching stream)
    prof = []
    for _ in range(max(3, min(args.steps, 10))):
        step(True, flags=4)
        prof.append(dev.profile_last())
        dist.barrier()
    prof = np.mean(np.array(prof), axis=0)
    clk = clocks.stop() if clocks is not None else None
    free_inst()

    # ---- weak scaling: every GPU holds a 2^k slice of a T3-(k + log2 N) instance
    weak = None
    if getattr(args, "weak", True):
        kw = k + g
        wstep, _, _, wfree = instance(kw)
        for _ in range(min(args.warmup, 2)):
            wstep()
        ms_w, out_w = timed(wstep, max(1, min(args.steps, 5)))
        weak = {"k": kw, "per_gpu_slice_log2": k, "ms_per_step": ms_w, "value": 99 * (1 << kw) / (ms_w * 1e-3) / 1e9, "unit": "Gfield-ops/s",
                "points_per_s": (1 << kw) / (ms_w * 1e-3), "scaling": "weak",
                "note": f"T3-{kw} sliced over {world} GPUs (per-GPU work equal to the 1-GPU T3-{k} run), host transcript"}
        wfree()

    # ---- BASELINE config #5 on N GPUs: 64 base columns x 2^20 rows, row-sharded -> Basefold commitment root
    config5 = None
    if getattr(args, "config5", True):
        try:
            import ctypes as C
            from ceno_b200 import dist as cdist
            c5_w, c5_log, c5_rate = 64, 20, 1
            vals = synth.fill_base(0x9052, 8 * 8 + 22 + 8)
            cb.poseidon2_set_params(dev, vals[:64].reshape(8, 8), vals[64:86], vals[86:94], 0)
            rows_local = (1 << c5_log) // world
            loc = np.stack([synth.fill_base(4242 + c, rows_local, start=rank * rows_local) for c in range(c5_w)])
            t_loc = torch.from_numpy(loc.view(np.int64)).cuda()

            def c5_step():
                root, code_local, tree = cdist.commit_sharded(dev, t_loc, c5_log, c5_rate, rank, world, torch, dist)
                tree.free()
                return root
            root_n = c5_step()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                root_n = c5_step()
            e1.record()
            torch.cuda.synchronize()
            tt = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            config5 = {"workload": "BASELINE #5: 2^26-element batch = 64 base columns x 2^20 rows, row-sharded; RS-encode (rate 1/2) + Poseidon2 Merkle commit",
                       "n_gpus": world, "ms": float(tt.item()), "root": [int(x) for x in root_n],
                       "exchange": "NCCL all-to-all rows->columns (RS-encode per column), all-to-all columns->rows (leaf hash per row range), all-gather of subtree roots",
                       "note": "placeholder Poseidon2 constants; max over ranks, CUDA events"}
            if rank == 0:   # parity: the single-device commitment of the same matrix on rank 0
                full = np.concatenate([synth.fill_base(4242 + c, 1 << c5_log) for c in range(c5_w)])
                buf = dev.to_device(full)
                cm = cb.BasefoldCommitment(dev, buf, c5_w, c5_log, cb.BasefoldParams(rate_log=c5_rate))
                config5["root_equals_single_gpu_commit"] = bool([int(x) for x in cm.root] == config5["root"])
                cm.free(); buf.free()
        except Exception as e:  # noqa: BLE001
            config5 = {"error": repr(e)}

    # ---- BASELINE config #4 shape on N GPUs: the keccak lookup tower (1094 records per row, virtual leaves), 2^17 rows per GPU
    config4 = None
    if getattr(args, "config4", True):
        try:
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
            import bench_rows
            config4 = bench_rows.keccak_tower_row(dev, 17, comm=comm, xchg=xchg, barrier=dist.barrier, world=world, rank=rank)
            digests = [None] * world
            dist.all_gather_object(digests, config4.pop("proof_digest"))
            config4["proof_identical_on_all_ranks"] = len(set(digests)) == 1
        except Exception as e:  # noqa: BLE001
            config4 = {"error": repr(e)}

    if rank == 0:
        n = 1 << k
        ops = 99 * n
        # ---- parity: the oracle's proof of the very same instance (rank 0's host cores)
        from oracle import oracle as orc
        if hasattr(orc, "set_num_threads"):
            orc.set_num_threads(_physical_cores())
        cpu_run(min(args.cpu_k, 16))
        t_cpu, cpu_proof = cpu_run(args.cpu_k)
        parity_checked = None
        if args.cpu_k == k:
            for nm, o in (("host transcript", out), ("device challenger", out_dev), ("e2e from host buffers", out_e2e)):
                assert_same_proof(o, cpu_proof, f"T3-{k} on {world} GPUs, {nm}")
            parity_checked = f"oracle k={k}: the {world}-GPU proof (host transcript, device challenger, e2e) is bit-exact"
        cores = orc.num_threads()
        cpu_val = 99 * (1 << args.cpu_k) / t_cpu / 1e9
        # ---- roofline of rank 0's dominant kernel: the fused fix_variable + evaluation of the streaming rounds
        virt = getattr(args, "eq", "virtual") == "virtual"
        m = 2 if virt else 3
        peak, peak_src = read_peaks()
        n_split = max(2, k_local - 13)                     # rounds 1 .. n_split-1 run the streaming kernel on the local slice
        rounds = [j for j in range(1, n_split) if j < len(prof)]
        fused_bytes = [1.5 * m * 16 * (1 << (k_local - j + 1)) for j in rounds]
        fused_ms = [float(prof[j]) for j in rounds]
        ach = sum(fused_bytes) / (sum(fused_ms) * 1e-3) / 1e9 if rounds else 0.0
        roofline = {"bound": "hbm", "kernel": "veq_tma_kernel<FOLD=1> on rank 0's 1/%d slice (its launches include the in-kernel NVLink exchange of the round's partial sums)" % world,
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "traffic": None,
                    "bytes_definition": f"per launch 1.5*m*16*n_in over the rank's slice, m = {m} streamed MLEs",
                    "launches_per_step": len(rounds), "ms_per_step_in_kernel": sum(fused_ms), "round_ms": [round(float(x), 5) for x in prof]}
        line = {
            "metric": "sumcheck Gfield-ops/s", "value": ops / (ms * 1e-3) / 1e9, "unit": "Gfield-ops/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64 (Goldilocks, ext2)", "data": "synthetic",
            "config": workload_config(k, world),
            "config_detail": {"parallelism": f"hypercube slices x{world}: rank q owns [q 2^{k_local}, (q+1) 2^{k_local}) of every MLE", "eq": getattr(args, "eq", "virtual"),
                              "exchange": "in-kernel: the round kernel's last block stores its partial sums into every peer's NVLink-mapped mailbox, waits for the N flags, "
                                          "adds mod p; once the global state fits one thread-block cluster (2^16 elements per MLE) the slices are all-gathered over NVLink "
                                          "and the remaining rounds run replicated (no further exchange); replicated host transcript",
                              "per_gpu_inputs_MiB": 3 * 16 * (1 << k_local) >> 20},
            "points_per_s": n / (ms * 1e-3), "rounds_per_s": k / (ms * 1e-3),
            "device_challenger": {"ms_per_step": ms_dev, "value": ops / (ms_dev * 1e-3) / 1e9, "unit": "Gfield-ops/s"},
            "e2e": {"value": ops / (ms_e2e * 1e-3) / 1e9, "unit": "Gfield-ops/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": world * 2 * nbytes + 16 * k, "d2h_bytes_per_step": 16 * (k * deg + 3 + k),
                    "note": f"every rank uploads its 1/{world} slices of A and B from pinned host memory over its own PCIe link, then the sharded prove; max over ranks"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_val, "unit": "Gfield-ops/s", "cores": cores, "kind": "port", "ms": t_cpu * 1e3,
                             "sample": f"T3-{args.cpu_k} full sumcheck, one run on rank 0's host; oracle port with the reference's decomposition, OpenMP {cores} threads"},
            "weak_scaling": weak,
            "config5": config5,
            "config4": config4,
            "clocks": clk,
            "parity_checked": parity_checked,
        }
        print(json.dumps(line))
    dist.barrier()
    comm.close()
    dev.close()
    dist.destroy_process_group()


def _physical_cores():
    """Distinct physical cores in this process's affinity mask (SMT siblings counted once)."""
    try:
        seen = set()
        for cpu in os.sched_getaffinity(0):
            with open(f"/sys/devices/system/cpu/cpu{cpu}/topology/thread_siblings_list") as f:
                seen.add(f.read().strip())
        return max(1, len(seen))
    except Exception:
        return max(1, (os.cpu_count() or 2) // 2)


def _cpu_env():
    """The CPU arm runs one thread per physical core, spread and pinned (measured on the 2x32-core host: 128 SMT
    threads are ~10x slower than 64 pinned cores for this memory-bound integer loop).  Must run before libgomp loads."""
    if "--impl" in sys.argv and os.environ.get("TORCHELASTIC_RUN_ID"):
        os.environ["OMP_NUM_THREADS"] = str(_physical_cores())   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core
    os.environ.setdefault("OMP_NUM_THREADS", str(_physical_cores()))
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    os.environ.setdefault("OMP_PLACES", "cores")


def main():
    # N>1 ranks must NOT bind: every rank would pin its main thread to the same first core, and the ranks' host
    # transcript loops (which answer each other's in-kernel exchange) would time-slice one core (measured: 52 ms/step)
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 or "--impl" in sys.argv:
        _cpu_env()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--k", type=int, default=24, help="log2 hypercube size of the T3 instance")
    ap.add_argument("--eq", default="virtual", choices=["virtual", "table"], help="how eq(w,.) is given to the prover (see the module docstring)")
    ap.add_argument("--cpu-k", type=int, default=24, help="log2 size of the bounded CPU sample")
    ap.add_argument("--no-rows", dest="rows", action="store_false", help="N=1: skip the secondary rows (Z, TOWER, EQ-24, FOLD-24, C-26, RS-26, BATCH-26)")
    ap.add_argument("--no-weak", dest="weak", action="store_false", help="N>1: skip the weak-scaling measurement (T3-(k + log2 N))")
    ap.add_argument("--no-config4", dest="config4", action="store_false", help="N>1: skip the sharded keccak lookup tower (BASELINE config #4 shape, 2^17 rows per GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
