"""ctypes loader for the C ABI (include/ceno_b200.h).  Fails loudly when the CUDA library is
missing — there is no CPU or eager fallback anywhere in this package."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("CENO_B200_LIB") or os.path.join(HERE, "lib", "libceno_b200.so")   # override: kernel A/B experiments only

CG_OK = 0
ERR_NAMES = {1: "CG_ERR_CUDA", 2: "CG_ERR_INVALID", 3: "CG_ERR_UNSUPPORTED", 4: "CG_ERR_OOM", 5: "CG_ERR_NO_DEVICE", 6: "CG_ERR_STATE"}


class CgMleDesc(C.Structure):
    _fields_ = [("dptr", C.c_void_p), ("len", C.c_uint64), ("num_vars", C.c_uint32), ("is_ext", C.c_uint32)]


class CgPoseidon2Params(C.Structure):
    _fields_ = [("ext_rc", (C.c_uint64 * 8) * 8), ("int_rc", C.c_uint64 * 22), ("diag", C.c_uint64 * 8),
                ("mds_variant", C.c_uint32), ("pad", C.c_uint32)]


class CgSchedTask(C.Structure):
    _fields_ = [("task_id", C.c_uint32), ("reserved", C.c_uint32), ("estimated_memory_bytes", C.c_uint64), ("booked_memory_bytes", C.c_uint64)]


class CgSchedResult(C.Structure):
    _fields_ = [("task_id", C.c_uint32), ("lane_id", C.c_uint32), ("status", C.c_int32), ("launch_seq", C.c_uint32),
                ("booked_total_at_launch", C.c_uint64), ("queue_delay_ms", C.c_double), ("host_execution_ms", C.c_double),
                ("event_wait_ms", C.c_double)]


SCHED_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p)


class CgBasefoldParams(C.Structure):
    _fields_ = [("rate_log", C.c_uint32), ("n_queries", C.c_uint32), ("pow_bits", C.c_uint32), ("reserved", C.c_uint32)]


class CgBasefoldOpening(C.Structure):
    _fields_ = [("commit", C.c_void_p), ("h_point_ext", C.c_void_p), ("h_evals_ext", C.c_void_p)]


class CgPcsTranscriptVt(C.Structure):
    _fields_ = [("user", C.c_void_p), ("observe_label", C.c_void_p), ("sample_ext", C.c_void_p), ("observe_exts", C.c_void_p),
                ("observe_base", C.c_void_p), ("sample_bits", C.c_void_p), ("grind", C.c_void_p)]


class CgTowerSpec(C.Structure):
    _fields_ = [("leaves", C.c_void_p * 4), ("num_vars", C.c_uint32), ("is_logup", C.c_uint32)]


CHALLENGE_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(C.c_uint64))
SAMPLE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64))
APPEND_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64)
BEGIN_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint64)


class CgTowerVGroup(C.Structure):
    _fields_ = [("records", C.c_void_p), ("n_records", C.c_uint32), ("reserved", C.c_uint32), ("num_instances", C.c_uint64), ("default_ext", C.c_uint64 * 2)]


class CgTowerVSpec(C.Structure):
    _fields_ = [("q", CgTowerVGroup), ("p", CgTowerVGroup), ("is_logup", C.c_uint32), ("reserved", C.c_uint32)]


class CgTranscriptVt(C.Structure):
    _fields_ = [("user", C.c_void_p), ("sample", SAMPLE_FN), ("append_exts", APPEND_FN),
                ("sumcheck_begin", BEGIN_FN), ("round_challenge", CHALLENGE_CB)]


# every symbol include/ceno_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "cg_init", "cg_destroy", "cg_last_error", "cg_version", "cg_device_info", "cg_alloc", "cg_free", "cg_free_async", "cg_pool_stats",
    "cg_pool_trim", "cg_h2d", "cg_d2h", "cg_d2d", "cg_stream_sync", "cg_host_alloc_pinned", "cg_host_free_pinned",
    "cg_launch_count", "cg_build_eq", "cg_selector_compute", "cg_fix_variable", "cg_mle_evaluate", "cg_sumcheck_create",
    "cg_sumcheck_round_eval", "cg_sumcheck_bind", "cg_sumcheck_final_evals", "cg_sumcheck_round", "cg_sumcheck_peek",
    "cg_sumcheck_destroy", "cg_sumcheck_prove", "cg_sumcheck_prove_standin_device", "cg_standin_init",
    "cg_standin_append_message", "cg_standin_append_ext", "cg_standin_sample", "cg_standin_challenge_cb", "cg_tower_interleave_out_len", "cg_tower_interleave", "cg_tower_build", "cg_tower_build_virtual", "cg_tower_build_sharded", "cg_tower_build_virtual_sharded",
    "cg_comm_arena_create", "cg_comm_arena_connect",
    "cg_tower_output_evals", "cg_tower_proof_len", "cg_tower_point_len", "cg_tower_create_proof", "cg_tower_destroy",
    "cg_standin_vt", "cg_wit_infer_by_monomial_expr", "cg_profile_last",
    "cg_comm_create", "cg_comm_connect", "cg_comm_destroy", "cg_sumcheck_attach_comm", "cg_sumcheck_prove_sharded",
    "cg_poseidon2_set_params", "cg_poseidon2_permute", "cg_merkle_commit",
    "cg_rotation_next_base_mle", "cg_rotation_selector",
    "cg_sched_execute", "cg_stream_create", "cg_stream_destroy", "cg_ntt", "cg_rs_encode",
    "cg_ecc_quark_selectors", "cg_split_even_odd", "cg_ecc_quark_terms",
    "cg_basefold_commit", "cg_basefold_commitment_root", "cg_basefold_commitment_codeword", "cg_basefold_commitment_free",
    "cg_standin_pcs_vt", "cg_basefold_proof_len", "cg_basefold_batch_open",
]

_lib = None


def load():
    """Load libceno_b200.so.  Raises if it has not been built (run `python -m ceno_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise RuntimeError(f"{SO} is missing: build it with `python -m ceno_b200.build` (nvcc, sm_100a). "
                           "ceno_b200 has no CPU fallback.")
    lib = C.CDLL(SO)
    vp, u64, u32, i32, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_size_t
    P = C.POINTER
    sig = {
        "cg_init": (i32, [i32, P(vp)]),
        "cg_destroy": (i32, [vp]),
        "cg_last_error": (C.c_char_p, [vp]),
        "cg_version": (C.c_char_p, []),
        "cg_device_info": (i32, [vp, P(i32), P(i32), P(i32), P(sz), P(sz)]),
        "cg_alloc": (i32, [vp, sz, P(vp)]),
        "cg_free": (i32, [vp, vp]),
        "cg_free_async": (i32, [vp, vp, vp]),
        "cg_pool_stats": (i32, [vp, P(sz), P(sz)]),
        "cg_pool_trim": (i32, [vp]),
        "cg_h2d": (i32, [vp, vp, vp, sz, vp]),
        "cg_d2h": (i32, [vp, vp, vp, sz, vp]),
        "cg_d2d": (i32, [vp, vp, vp, sz, vp]),
        "cg_stream_sync": (i32, [vp, vp]),
        "cg_host_alloc_pinned": (i32, [vp, sz, P(vp)]),
        "cg_host_free_pinned": (i32, [vp, vp]),
        "cg_launch_count": (u64, [vp]),
        "cg_build_eq": (i32, [vp, vp, u32, vp, u64, u64, vp]),
        "cg_selector_compute": (i32, [vp, i32, vp, u32, u64, u64, vp, u32, u32, vp, vp]),
        "cg_fix_variable": (i32, [vp, P(CgMleDesc), u32, vp, P(vp), vp]),
        "cg_mle_evaluate": (i32, [vp, P(CgMleDesc), vp, vp, vp]),
        "cg_sumcheck_create": (i32, [vp, P(CgMleDesc), u32, vp, vp, vp, u32, u32, u32, u32, vp, P(vp)]),
        "cg_sumcheck_round_eval": (i32, [vp, vp]),
        "cg_sumcheck_bind": (i32, [vp, vp]),
        "cg_sumcheck_final_evals": (i32, [vp, vp]),
        "cg_sumcheck_round": (u32, [vp]),
        "cg_sumcheck_peek": (i32, [vp, u32, P(vp), P(u64), P(u32)]),
        "cg_sumcheck_destroy": (i32, [vp]),
        "cg_sumcheck_prove": (i32, [vp, P(CgMleDesc), u32, vp, vp, vp, u32, u32, u32, u32, CHALLENGE_CB, vp, vp, vp, vp, vp]),
        "cg_sumcheck_prove_standin_device": (i32, [vp, P(CgMleDesc), u32, vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, vp, vp]),
        "cg_standin_init": (None, [vp, vp, u64]),
        "cg_standin_append_message": (None, [vp, vp, u64]),
        "cg_standin_append_ext": (None, [vp, vp, u64]),
        "cg_standin_sample": (None, [vp, C.c_char_p, vp]),
        "cg_standin_challenge_cb": (None, [vp, u32, vp, u32, vp]),
        "cg_tower_interleave_out_len": (u64, [u32, u64, u32]),
        "cg_tower_interleave": (i32, [vp, P(CgMleDesc), u32, u64, u32, vp, vp, vp]),
        "cg_tower_build": (i32, [vp, P(CgTowerSpec), u32, vp, P(vp)]),
        "cg_tower_build_virtual": (i32, [vp, P(CgTowerVSpec), u32, vp, P(vp)]),
        "cg_tower_build_sharded": (i32, [vp, vp, P(CgTowerSpec), u32, vp, P(vp)]),
        "cg_tower_build_virtual_sharded": (i32, [vp, vp, P(CgTowerVSpec), u32, vp, P(vp)]),
        "cg_comm_arena_create": (i32, [vp, sz, vp]),
        "cg_comm_arena_connect": (i32, [vp, vp]),
        "cg_tower_output_evals": (i32, [vp, u32, vp]),
        "cg_tower_proof_len": (u64, [vp]),
        "cg_tower_point_len": (u32, [vp]),
        "cg_tower_create_proof": (i32, [vp, P(CgTranscriptVt), vp, vp]),
        "cg_tower_destroy": (i32, [vp]),
        "cg_standin_vt": (None, [vp, P(CgTranscriptVt)]),
        "cg_wit_infer_by_monomial_expr": (i32, [vp, P(CgMleDesc), u32, vp, vp, vp, u32, u32, vp, vp]),
        "cg_profile_last": (i32, [vp, vp, u32, P(u32)]),
        "cg_comm_create": (i32, [vp, i32, i32, P(vp), vp]),
        "cg_comm_connect": (i32, [vp, vp]),
        "cg_comm_destroy": (i32, [vp]),
        "cg_sumcheck_attach_comm": (i32, [vp, vp]),
        "cg_poseidon2_set_params": (i32, [vp, vp]),
        "cg_poseidon2_permute": (i32, [vp, vp, u64, vp]),
        "cg_merkle_commit": (i32, [vp, vp, u64, u64, i32, vp, vp, vp]),
        "cg_rotation_next_base_mle": (i32, [vp, P(CgMleDesc), u32, vp, vp]),
        "cg_rotation_selector": (i32, [vp, vp, u64, u32, u32, vp, vp]),
        "cg_sched_execute": (i32, [vp, P(CgSchedTask), u32, u32, u64, SCHED_FN, vp, P(CgSchedResult)]),
        "cg_ntt": (i32, [vp, vp, u32, u64, u64, u32, vp]),
        "cg_rs_encode": (i32, [vp, vp, u64, u32, u32, vp, u32, vp]),
        "cg_ecc_quark_selectors": (i32, [vp, vp, u32, u64, vp, vp, vp, vp]),
        "cg_split_even_odd": (i32, [vp, P(CgMleDesc), u32, P(vp), P(vp), vp]),
        "cg_ecc_quark_terms": (i32, [vp, vp, vp, vp, vp, vp, u32, u32, P(u32), P(u32)]),
        "cg_basefold_commit": (i32, [vp, vp, u64, u32, P(CgBasefoldParams), vp, P(vp)]),
        "cg_basefold_commitment_root": (i32, [vp, vp]),
        "cg_basefold_commitment_codeword": (i32, [vp, P(vp), P(vp)]),
        "cg_basefold_commitment_free": (i32, [vp]),
        "cg_standin_pcs_vt": (None, [vp, P(CgPcsTranscriptVt)]),
        "cg_basefold_proof_len": (u64, [P(CgBasefoldOpening), u32, P(CgBasefoldParams)]),
        "cg_basefold_batch_open": (i32, [vp, P(CgBasefoldOpening), u32, P(CgBasefoldParams), P(CgPcsTranscriptVt), vp, u64, vp]),
        "cg_stream_create": (i32, [vp, P(vp)]),
        "cg_stream_destroy": (i32, [vp, vp]),
        "cg_sumcheck_prove_sharded": (i32, [vp, vp, P(CgMleDesc), u32, vp, vp, vp, u32, u32, u32, u32, CHALLENGE_CB, vp, vp, vp, vp, vp, vp]),
    }
    for name in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype, fn.argtypes = sig[name]
    _lib = lib
    return lib
