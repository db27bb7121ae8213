"""Host-side mirror (Python) of the reference interface for the hot path, over the C ABI.

Names follow the reference so the parity tests read like its own tests:
  MultilinearExtension        multilinear_extensions::mle::MultilinearExtension (device-resident;
                              the role MultilinearExtensionGpu plays, gkr_iop/src/gpu/mod.rs:157-161)
  build_eq_x_r_vec            multilinear_extensions::virtual_poly::build_eq_x_r_vec
  SelectorType.compute        gkr_iop/src/selector.rs:131-245
  IOPProverState.prove        sumcheck::structs::IOPProverState (call sites gkr_iop/src/gkr/layer/cpu/mod.rs:217-237)
  TowerProver.create_proof    ceno_zkvm/src/scheme/cpu/mod.rs:346-554 (CpuTowerProver) / hal.rs:173-207
  StandInTranscript           stand-in for transcript::BasicTranscript (NOT Poseidon2)

Everything here only marshals arguments; all arithmetic happens in libceno_b200.so on the GPU.
Arrays are numpy uint64; an extension element is two consecutive limbs [c0, c1].
"""
import ctypes as C

import numpy as np

from . import _lib

P = 0xFFFFFFFF00000001


class CenoB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_lib.ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class Device:
    """A cg_ctx: one GPU, its stream and pooled allocator (cuda_hal context, gkr_iop/src/gpu/mod.rs:53-66)."""

    def __init__(self, device_id=0):
        self.lib = _lib.load()
        self.ctx = C.c_void_p()
        rc = self.lib.cg_init(device_id, C.byref(self.ctx))
        if rc != _lib.CG_OK:
            raise CenoB200Error(rc, "cg_init failed: an sm_100 (B200) GPU is required; there is no CPU fallback")
        self.device_id = device_id

    def check(self, rc):
        if rc != _lib.CG_OK:
            raise CenoB200Error(rc, self.lib.cg_last_error(self.ctx).decode())

    def close(self):
        if self.ctx:
            self.lib.cg_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def info(self):
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        fr, to = C.c_size_t(), C.c_size_t()
        self.check(self.lib.cg_device_info(self.ctx, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(fr), C.byref(to)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "free": fr.value, "total": to.value}

    def launch_count(self):
        return int(self.lib.cg_launch_count(self.ctx))

    def sync(self):
        self.check(self.lib.cg_stream_sync(self.ctx, None))

    # -- memory
    def alloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.lib.cg_alloc(self.ctx, nbytes, C.byref(p)))
        return DeviceBuffer(self, p.value, nbytes)

    def to_device(self, arr):
        arr = _u64(arr)
        buf = self.alloc(arr.nbytes)
        self.check(self.lib.cg_h2d(self.ctx, buf.ptr, _vp(arr), arr.nbytes, None))
        self.sync()
        return buf

    def h2d(self, dst_ptr, host_ptr, nbytes, stream=None):
        self.check(self.lib.cg_h2d(self.ctx, C.c_void_p(dst_ptr), C.c_void_p(host_ptr), nbytes, C.c_void_p(stream) if stream else None))

    def pinned(self, nbytes):
        """Pinned host buffer as a numpy uint64 view (kept alive by the Device)."""
        p = C.c_void_p()
        self.check(self.lib.cg_host_alloc_pinned(self.ctx, nbytes, C.byref(p)))
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(nbytes // 8,))
        return arr, p.value

    def profile_last(self):
        n = C.c_uint32()
        buf = np.zeros(64, np.float32)
        self.check(self.lib.cg_profile_last(self.ctx, _vp(buf), 64, C.byref(n)))
        return buf[:n.value].copy()

    def pool_stats(self):
        u, r = C.c_size_t(), C.c_size_t()
        self.check(self.lib.cg_pool_stats(self.ctx, C.byref(u), C.byref(r)))
        return u.value, r.value


class DeviceBuffer:
    def __init__(self, dev, ptr, nbytes, owner=True):
        self.dev, self.ptr, self.nbytes, self.owner = dev, ptr, nbytes, owner

    def to_host(self, nbytes=None, offset=0):
        nbytes = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(nbytes // 8, np.uint64)
        self.dev.check(self.dev.lib.cg_d2h(self.dev.ctx, _vp(out), C.c_void_p(self.ptr + offset), nbytes, None))
        self.dev.sync()
        return out

    def free(self, stream=None):
        """Return the block to the pool.  Without `stream` the caller must have synchronised every stream that still uses
        the buffer (cg_free); with `stream` the release is ordered behind the work enqueued on it (cg_free_async)."""
        if self.owner and self.ptr:
            if stream is not None:
                self.dev.check(self.dev.lib.cg_free_async(self.dev.ctx, C.c_void_p(self.ptr), C.c_void_p(stream) if stream else None))
            else:
                self.dev.lib.cg_free(self.dev.ctx, C.c_void_p(self.ptr))
            self.ptr = 0


class MultilinearExtension:
    """Dense device MLE: 2^num_vars evaluations, base (8 B) or ext (16 B)."""

    def __init__(self, dev, buf, num_vars, is_ext, length=None):
        self.dev, self.buf, self.num_vars, self.is_ext = dev, buf, num_vars, bool(is_ext)
        self.len = (1 << num_vars) if length is None else length

    @classmethod
    def from_evaluations_vec(cls, dev, num_vars, evals):          # base field
        evals = _u64(evals)
        assert evals.size <= (1 << num_vars)
        return cls(dev, dev.to_device(evals), num_vars, False, evals.size)

    @classmethod
    def from_evaluations_ext_vec(cls, dev, num_vars, evals):      # extension field, limbs interleaved
        evals = _u64(evals)
        assert evals.size % 2 == 0 and evals.size // 2 <= (1 << num_vars)
        return cls(dev, dev.to_device(evals), num_vars, True, evals.size // 2)

    @classmethod
    def from_evaluation_vec_smart(cls, dev, num_vars, evals, is_ext=False):
        """MultilinearExtension::from_evaluation_vec_smart: base or ext evaluations, chosen by the element type."""
        return cls.from_evaluations_ext_vec(dev, num_vars, evals) if is_ext else cls.from_evaluations_vec(dev, num_vars, evals)

    def as_view_slice(self, num_chunks, chunk_index):
        """Zero-copy view of chunk `chunk_index` of `num_chunks` equal contiguous chunks (as_view_slice; the device form is
        owned_subrange, ceno_zkvm/src/scheme/gpu/mod.rs:2088-2113): log2(num_chunks) fewer variables, the TOP index bits fixed."""
        assert num_chunks & (num_chunks - 1) == 0 and self.len == (1 << self.num_vars) and 0 <= chunk_index < num_chunks
        sub = self.num_vars - (num_chunks.bit_length() - 1)
        nbytes = (16 if self.is_ext else 8) << sub
        view = DeviceBuffer(self.dev, self.buf.ptr + nbytes * chunk_index, nbytes, owner=False)
        return MultilinearExtension(self.dev, view, sub, self.is_ext)

    def get_base_field_vec(self):
        assert not self.is_ext
        return self.evaluations()

    def get_ext_field_vec(self):
        assert self.is_ext
        return self.evaluations()

    def desc(self):
        return _lib.CgMleDesc(self.buf.ptr, self.len, self.num_vars, 1 if self.is_ext else 0)

    def evaluations(self):
        return self.buf.to_host(self.len * (16 if self.is_ext else 8))

    def evaluate(self, point):
        point = _u64(point)
        assert point.size == 2 * self.num_vars
        out = np.zeros(2, np.uint64)
        d = self.desc()
        self.dev.check(self.dev.lib.cg_mle_evaluate(self.dev.ctx, C.byref(d), _vp(point), _vp(out), None))
        return out

    def fix_variable(self, r):
        """Out-of-place LSB fold; returns a new ext MLE with one variable fewer."""
        r = _u64(r)
        out = self.dev.alloc(16 << (self.num_vars - 1))
        d = (_lib.CgMleDesc * 1)(self.desc())
        outs = (C.c_void_p * 1)(out.ptr)
        self.dev.check(self.dev.lib.cg_fix_variable(self.dev.ctx, d, 1, _vp(r), outs, None))
        self.dev.sync()
        return MultilinearExtension(self.dev, out, self.num_vars - 1, True)

    def free(self, stream=None):
        self.buf.free(stream)


class EqPolynomial:
    """Virtual eq(w, .) MLE (cg_mle_desc kind CG_MLE_EQ): the prover receives the point instead of the
    2^k table build_eq_x_r_vec(w) would produce, with identical round messages and final evaluation.
    The role of the reference's virtual device MLEs (ceno_zkvm/src/scheme/gpu/mod.rs:2195-2268)."""

    is_ext = True

    def __init__(self, dev, w, num_vars=None):
        """num_vars < len(w): this rank's slice in prove_sharded (w stays the GLOBAL point)."""
        self.dev = dev
        self.w = _u64(w).copy()
        self.num_vars = self.w.size // 2 if num_vars is None else num_vars
        self.len = 1 << self.num_vars

    def desc(self):
        return _lib.CgMleDesc(self.w.ctypes.data, 0, self.num_vars, 2)

    def free(self):
        pass


def build_eq_x_r_vec(dev, r, offset=0, num_instances=None, stream=None, out=None):
    """eq table of point r (k ext) as a device ext MLE; optional prefix mask (build_mle_as_ceno).
    With `stream` (a cudaStream_t handle) the call is asynchronous on that stream."""
    r = _u64(r)
    k = r.size // 2
    n = 1 << k
    num_instances = n - offset if num_instances is None else num_instances
    out = dev.alloc(16 * n) if out is None else out
    dev.check(dev.lib.cg_build_eq(dev.ctx, _vp(r), k, C.c_void_p(out.ptr), offset, num_instances, C.c_void_p(stream) if stream else None))
    if not stream:
        dev.sync()
    return MultilinearExtension(dev, out, k, True)


class SelectorType:
    WHOLE, PREFIX, ORDERED_SPARSE, QUARK_LT = 0, 1, 2, 3

    @staticmethod
    def compute(dev, kind, out_point, offset=0, num_instances=0, indices=(), inner_vars=0):
        out_point = _u64(out_point)
        nv = out_point.size // 2
        idx = _u64(np.array(list(indices), dtype=np.uint64))
        out = dev.alloc(16 << nv)
        dev.check(dev.lib.cg_selector_compute(dev.ctx, kind, _vp(out_point), nv, offset, num_instances, _vp(idx), idx.size,
                                              inner_vars, C.c_void_p(out.ptr), None))
        dev.sync()
        return MultilinearExtension(dev, out, nv, True)


class StandInTranscript:
    """Stand-in sponge (splitmix64) with BasicTranscript's call order. Not Poseidon2."""

    def __init__(self, label=b"test"):
        self.lib = _lib.load()
        self.state = np.zeros(1, np.uint64)
        buf = (C.c_uint8 * max(len(label), 1)).from_buffer_copy(label or b"\0")
        self.lib.cg_standin_init(_vp(self.state), buf, len(label))

    def append_message(self, msg: bytes):
        buf = (C.c_uint8 * max(len(msg), 1)).from_buffer_copy(msg or b"\0")
        self.lib.cg_standin_append_message(_vp(self.state), buf, len(msg))

    def append_field_element_exts(self, e):
        e = _u64(e)
        self.lib.cg_standin_append_ext(_vp(self.state), _vp(e), e.size // 2)

    def sample_and_append_challenge(self, label: bytes):
        out = np.zeros(2, np.uint64)
        self.lib.cg_standin_sample(_vp(self.state), C.c_char_p(label), _vp(out))
        return out

    def sample_and_append_vec(self, label: bytes, n):
        """n challenges under one label (Transcript::sample_and_append_vec); stand-in semantics: n successive samples."""
        return np.concatenate([self.sample_and_append_challenge(label) for _ in range(n)]) if n else np.zeros(0, np.uint64)

    def sample_and_append_challenge_pows(self, size, label: bytes):
        """[1, a, a^2, ...] for ONE sampled a (Transcript::sample_and_append_challenge_pows)."""
        from .expr import ext_mul
        a = self.sample_and_append_challenge(label)
        a = (int(a[0]), int(a[1]))
        out, cur = [], (1, 0)
        for _ in range(size):
            out.append(cur)
            cur = ext_mul(cur, a)
        return np.array(out, dtype=np.uint64)

    def vtable(self):
        vt = _lib.CgTranscriptVt()
        self.lib.cg_standin_vt(_vp(self.state), C.byref(vt))
        return vt


def _terms(terms):
    coeff = np.zeros(2 * max(len(terms), 1), np.uint64)
    off = np.zeros(len(terms) + 1, np.uint32)
    idx = []
    for t, (c, ids) in enumerate(terms):
        coeff[2 * t:2 * t + 2] = _u64(c)
        off[t] = len(idx)
        idx.extend(ids)
    off[len(terms)] = len(idx)
    return coeff, off, np.array(idx if idx else [0], dtype=np.uint32)


class IOPProverState:
    """Device sumcheck prover for P(x) = sum_t c_t prod_{i in S_t} mle_i(x).

    `prove(...)` mirrors IOPProverState::prove(virtual_polys, transcript) -> (proof, state);
    the step API (round_eval / bind) is what a multi-GPU or Rust-transcript host drives."""

    FORCE_GENERIC, NO_FUSE, PROFILE, NO_TAIL, NO_PLAN, NO_MID = 1, 2, 4, 8, 16, 32

    def __init__(self, dev, mles, terms, num_vars, degree, flags=0):
        self.dev, self.mles, self.num_vars, self.degree = dev, mles, num_vars, degree
        descs = (_lib.CgMleDesc * max(len(mles), 1))(*[m.desc() for m in mles])
        coeff, off, idx = _terms(terms)
        self.h = C.c_void_p()
        dev.check(dev.lib.cg_sumcheck_create(dev.ctx, descs, len(mles), _vp(coeff), _vp(off), _vp(idx), len(terms),
                                             num_vars, degree, flags, None, C.byref(self.h)))
        self._keep = (descs, coeff, off, idx)

    def round_eval(self):
        out = np.zeros(2 * self.degree, np.uint64)
        self.dev.check(self.dev.lib.cg_sumcheck_round_eval(self.h, _vp(out)))
        return out

    def bind(self, r):
        r = _u64(r)
        self.dev.check(self.dev.lib.cg_sumcheck_bind(self.h, _vp(r)))

    def get_mle_flatten_final_evaluations(self):
        out = np.zeros(2 * max(len(self.mles), 1), np.uint64)
        self.dev.check(self.dev.lib.cg_sumcheck_final_evals(self.h, _vp(out)))
        return out[:2 * len(self.mles)].reshape(-1, 2)

    def peek(self, i):
        p, n, e = C.c_void_p(), C.c_uint64(), C.c_uint32()
        self.dev.check(self.dev.lib.cg_sumcheck_peek(self.h, i, C.byref(p), C.byref(n), C.byref(e)))
        buf = DeviceBuffer(self.dev, p.value, n.value * (16 if e.value else 8), owner=False)
        return buf.to_host()

    def close(self):
        if self.h:
            self.dev.lib.cg_sumcheck_destroy(self.h)
            self.h = C.c_void_p()

    @staticmethod
    def prove(dev, mles, terms, num_vars, degree, transcript=None, challenge_fn=None, flags=0, device_challenger=False, stream=None):
        """Returns (round_evals[num_vars,degree,2], final_evals[n_mles,2], challenges[num_vars,2]).

        transcript: a StandInTranscript driven in the reference order; challenge_fn(round, evals)->ext:
        an arbitrary host source (this is where a real BasicTranscript plugs in)."""
        lib = dev.lib
        descs = (_lib.CgMleDesc * max(len(mles), 1))(*[m.desc() for m in mles])
        coeff, off, idx = _terms(terms)
        rounds = np.zeros(2 * degree * max(num_vars, 1), np.uint64)
        fin = np.zeros(2 * max(len(mles), 1), np.uint64)
        chal = np.zeros(2 * max(num_vars, 1), np.uint64)
        st = C.c_void_p(stream) if stream else None
        if transcript is not None:
            transcript.append_message(int(num_vars).to_bytes(8, "little"))
            transcript.append_message(int(degree).to_bytes(8, "little"))
            if device_challenger:
                dev.check(lib.cg_sumcheck_prove_standin_device(dev.ctx, descs, len(mles), _vp(coeff), _vp(off), _vp(idx), len(terms),
                                                               num_vars, degree, flags, _vp(transcript.state), _vp(rounds), _vp(fin),
                                                               _vp(chal), st))
            else:
                cb = C.cast(lib.cg_standin_challenge_cb, _lib.CHALLENGE_CB)
                dev.check(lib.cg_sumcheck_prove(dev.ctx, descs, len(mles), _vp(coeff), _vp(off), _vp(idx), len(terms), num_vars,
                                                degree, flags, cb, _vp(transcript.state), _vp(rounds), _vp(fin), _vp(chal), st))
        else:
            def _cb(user, rnd, evals, deg, out):
                e = np.ctypeslib.as_array(evals, shape=(deg * 2,)).copy()
                r = _u64(challenge_fn(rnd, e))
                out[0], out[1] = int(r[0]), int(r[1])
            cb = _lib.CHALLENGE_CB(_cb)
            dev.check(lib.cg_sumcheck_prove(dev.ctx, descs, len(mles), _vp(coeff), _vp(off), _vp(idx), len(terms), num_vars,
                                            degree, flags, cb, None, _vp(rounds), _vp(fin), _vp(chal), st))
        return (rounds[:2 * degree * num_vars].reshape(num_vars, degree, 2), fin[:2 * len(mles)].reshape(-1, 2),
                chal[:2 * num_vars].reshape(num_vars, 2))


class Comm:
    """NVLink mailbox of one rank (cg_comm).  `exchange_handles` is any all-gather of 64-byte blobs
    (e.g. torch.distributed.all_gather_object); the library itself never calls a collective."""

    def __init__(self, dev, rank, nranks, exchange_handles, barrier=None):
        self.dev, self.rank, self.nranks, self.barrier = dev, rank, nranks, barrier
        self.h = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        dev.check(dev.lib.cg_comm_create(dev.ctx, rank, nranks, C.byref(self.h), handle))
        blobs = exchange_handles(bytes(handle))
        assert len(blobs) == nranks and all(len(b) == 64 for b in blobs)
        allh = (C.c_uint8 * (64 * nranks)).from_buffer_copy(b"".join(blobs))
        dev.check(dev.lib.cg_comm_connect(self.h, allh))
        if barrier:
            barrier()

    def create_arena(self, nbytes, exchange_handles):
        """Peer-mapped arena for sharded towers (cg_comm_arena_create / _connect): the same size on every rank."""
        handle = (C.c_uint8 * 64)()
        self.dev.check(self.dev.lib.cg_comm_arena_create(self.h, nbytes, handle))
        blobs = exchange_handles(bytes(handle))
        allh = (C.c_uint8 * (64 * self.nranks)).from_buffer_copy(b"".join(blobs))
        self.dev.check(self.dev.lib.cg_comm_arena_connect(self.h, allh))
        if self.barrier:
            self.barrier()

    def close(self):
        if self.h:
            if self.barrier:
                self.barrier()
            self.dev.lib.cg_comm_destroy(self.h)
            self.h = C.c_void_p()


def prove_sharded(dev, comm, mles, terms, num_vars_global, degree, transcript, flags=0, device_challenger=False, stream=None):
    """IOPProverState::prove over MLEs sharded across the ranks of `comm` (this rank passes its slices).
    Returns the global (round_evals, final_evals, challenges), identical on every rank."""
    lib = dev.lib
    descs = (_lib.CgMleDesc * max(len(mles), 1))(*[m.desc() for m in mles])
    coeff, off, idx = _terms(terms)
    k = num_vars_global
    rounds = np.zeros(2 * degree * max(k, 1), np.uint64)
    fin = np.zeros(2 * max(len(mles), 1), np.uint64)
    chal = np.zeros(2 * max(k, 1), np.uint64)
    transcript.append_message(int(k).to_bytes(8, "little"))
    transcript.append_message(int(degree).to_bytes(8, "little"))
    st = C.c_void_p(stream) if stream else None
    cb = C.cast(lib.cg_standin_challenge_cb, _lib.CHALLENGE_CB)
    dev.check(lib.cg_sumcheck_prove_sharded(dev.ctx, comm.h, descs, len(mles), _vp(coeff), _vp(off), _vp(idx), len(terms), k, degree, flags,
                                            cb, _vp(transcript.state), _vp(transcript.state) if device_challenger else None,
                                            _vp(rounds), _vp(fin), _vp(chal), st))
    return rounds[:2 * degree * k].reshape(k, degree, 2), fin[:2 * len(mles)].reshape(-1, 2), chal[:2 * k].reshape(k, 2)


def wit_infer_by_monomial_expr(dev, mles, terms, num_vars, stream=None):
    descs = (_lib.CgMleDesc * max(len(mles), 1))(*[m.desc() for m in mles])
    coeff, off, idx = _terms(terms)
    out = dev.alloc(16 << num_vars)
    dev.check(dev.lib.cg_wit_infer_by_monomial_expr(dev.ctx, descs, len(mles), _vp(coeff), _vp(off), _vp(idx), len(terms), num_vars,
                                                    C.c_void_p(out.ptr), C.c_void_p(stream) if stream else None))
    if not stream:
        dev.sync()
    return MultilinearExtension(dev, out, num_vars, True)


def interleaving_mles_to_mles(dev, mles, num_instances, num_limbs, default, stream=None):
    """ceno_zkvm/src/scheme/utils.rs:402-462 on the device: returns `num_limbs` ext MLEs (tower leaves)."""
    descs = (_lib.CgMleDesc * len(mles))(*[m.desc() for m in mles])
    out_len = int(dev.lib.cg_tower_interleave_out_len(len(mles), num_instances, num_limbs))
    out = dev.alloc(16 * out_len * num_limbs)
    d = _u64(default)
    dev.check(dev.lib.cg_tower_interleave(dev.ctx, descs, len(mles), num_instances, num_limbs, _vp(d), C.c_void_p(out.ptr),
                                          C.c_void_p(stream) if stream else None))
    if not stream:
        dev.sync()
    nv = out_len.bit_length() - 1
    return [MultilinearExtension(dev, DeviceBuffer(dev, out.ptr + 16 * out_len * i, 16 * out_len, owner=(i == 0)), nv, True) for i in range(num_limbs)], out


class TowerProverSpec:
    """Leaves of one tower: product (a, b of 2^(num_vars-1)) or logup (p1, p2, q1, q2 of 2^num_vars;
    p1 = p2 = None -> numerators are ones)."""

    def __init__(self, leaves, num_vars, is_logup):
        self.leaves, self.num_vars, self.is_logup = leaves, num_vars, is_logup


class VirtualTowerSpec:
    """A tower spec given by its RECORD MLEs (GpuVirtualInterleavedExt, ceno_zkvm/src/scheme/gpu/mod.rs:2195-2268): the
    interleaved fan-in leaves are never materialised.  Product: records = the read / write set (default 1).  Logup:
    records = denominators (default = the challenge alpha), numerators = None for all-one numerators."""

    def __init__(self, records, num_instances, default, is_logup, numerators=None, numerator_default=(1, 0)):
        self.records, self.num_instances, self.default, self.is_logup = list(records), num_instances, default, is_logup
        self.numerators, self.numerator_default = (list(numerators) if numerators else []), numerator_default


class TowerProver:
    @classmethod
    def sharded(cls, dev, comm, specs, stream=None):
        """cg_tower_build_sharded: `specs` carry this rank's slice of both fan-in halves of every leaf array and the GLOBAL
        num_vars; create_proof returns the global proof on every rank."""
        self = cls.__new__(cls)
        self.dev = dev
        arr = (_lib.CgTowerSpec * len(specs))()
        for i, s in enumerate(specs):
            for z in range(4):
                m = s.leaves[z] if z < len(s.leaves) else None
                arr[i].leaves[z] = m.buf.ptr if m is not None else None
            arr[i].num_vars = s.num_vars
            arr[i].is_logup = 1 if s.is_logup else 0
        self.h = C.c_void_p()
        dev.check(dev.lib.cg_tower_build_sharded(dev.ctx, comm.h, arr, len(specs), C.c_void_p(stream) if stream else None, C.byref(self.h)))
        self.specs = specs
        return self

    @classmethod
    def from_records(cls, dev, vspecs, stream=None, comm=None):
        """cg_tower_build_virtual (comm=None) / cg_tower_build_virtual_sharded: towers over virtual leaf layers."""
        self = cls.__new__(cls)
        self.dev = dev
        arr = (_lib.CgTowerVSpec * len(vspecs))()
        keep = []

        def group(g, recs, ninst, default):
            descs = (_lib.CgMleDesc * max(len(recs), 1))(*[m.desc() for m in recs])
            keep.append(descs)
            g.records = C.cast(descs, C.c_void_p) if recs else None
            g.n_records, g.num_instances = len(recs), ninst
            g.default_ext[0], g.default_ext[1] = int(default[0]), int(default[1])
        for i, v in enumerate(vspecs):
            group(arr[i].q, v.records, v.num_instances, v.default)
            group(arr[i].p, v.numerators, v.num_instances, v.numerator_default)
            arr[i].is_logup = 1 if v.is_logup else 0
        self.h = C.c_void_p()
        if comm is not None:
            dev.check(dev.lib.cg_tower_build_virtual_sharded(dev.ctx, comm.h, arr, len(vspecs), C.c_void_p(stream) if stream else None, C.byref(self.h)))
        else:
            dev.check(dev.lib.cg_tower_build_virtual(dev.ctx, arr, len(vspecs), C.c_void_p(stream) if stream else None, C.byref(self.h)))
        self.specs, self._keep = vspecs, keep
        return self

    def __init__(self, dev, specs, stream=None):
        self.dev = dev
        arr = (_lib.CgTowerSpec * len(specs))()
        for i, s in enumerate(specs):
            for z in range(4):
                m = s.leaves[z] if z < len(s.leaves) else None
                arr[i].leaves[z] = m.buf.ptr if m is not None else None
            arr[i].num_vars = s.num_vars
            arr[i].is_logup = 1 if s.is_logup else 0
        self.h = C.c_void_p()
        dev.check(dev.lib.cg_tower_build(dev.ctx, arr, len(specs), C.c_void_p(stream) if stream else None, C.byref(self.h)))
        self.specs = specs

    def output_evals(self, i):
        out = np.zeros(8, np.uint64)
        self.dev.check(self.dev.lib.cg_tower_output_evals(self.h, i, _vp(out)))
        return out

    def create_proof(self, transcript):
        """(proof u64 array, point ext array) — flattened TowerProofs (see include/ceno_b200.h)."""
        lib = self.dev.lib
        proof = np.zeros(max(int(lib.cg_tower_proof_len(self.h)), 1), np.uint64)
        point = np.zeros(2 * int(lib.cg_tower_point_len(self.h)), np.uint64)
        vt = transcript.vtable()
        self.dev.check(lib.cg_tower_create_proof(self.h, C.byref(vt), _vp(proof), _vp(point)))
        return proof[:int(lib.cg_tower_proof_len(self.h))], point

    def close(self):
        if self.h:
            self.dev.lib.cg_tower_destroy(self.h)
            self.h = C.c_void_p()


# ------------------------------------------------------------------- Poseidon2 / Merkle (a9; constants supplied by the caller)
def poseidon2_set_params(dev, ext_rc, int_rc, diag, mds_variant=0):
    p = _lib.CgPoseidon2Params()
    for r in range(8):
        for i in range(8):
            p.ext_rc[r][i] = int(ext_rc[r][i])
    for r in range(22):
        p.int_rc[r] = int(int_rc[r])
    for i in range(8):
        p.diag[i] = int(diag[i])
    p.mds_variant = mds_variant
    dev.check(dev.lib.cg_poseidon2_set_params(dev.ctx, C.byref(p)))


def poseidon2_permute(dev, states):
    states = _u64(states)
    buf = dev.to_device(states)
    dev.check(dev.lib.cg_poseidon2_permute(dev.ctx, C.c_void_p(buf.ptr), states.size // 8, None))
    out = buf.to_host()
    buf.free()
    return out


def merkle_commit(dev, matrix_buf, width, height, col_major=True, stream=None):
    """matrix_buf: DeviceBuffer of height*width base elements.  Returns (tree DeviceBuffer, root[4])."""
    tree = dev.alloc(32 * (2 * height - 1))
    root = np.zeros(4, np.uint64)
    dev.check(dev.lib.cg_merkle_commit(dev.ctx, C.c_void_p(matrix_buf.ptr), width, height, 1 if col_major else 0, C.c_void_p(tree.ptr), _vp(root),
                                       C.c_void_p(stream) if stream else None))
    return tree, root


# ------------------------------------------------------------------- rotation pre-passes (gkr_iop/src/utils.rs:19-76)
def rotation_next_base_mle(dev, mle, cyclic_group_log2_size):
    out = dev.alloc(8 * mle.len)
    d = mle.desc()
    dev.check(dev.lib.cg_rotation_next_base_mle(dev.ctx, C.byref(d), cyclic_group_log2_size, C.c_void_p(out.ptr), None))
    dev.sync()
    return MultilinearExtension(dev, out, mle.num_vars, False, mle.len)


def rotation_selector(dev, eq, cyclic_subgroup_size, cyclic_group_log2_size):
    out = dev.alloc(16 * eq.len)
    dev.check(dev.lib.cg_rotation_selector(dev.ctx, C.c_void_p(eq.buf.ptr), eq.len, cyclic_subgroup_size, cyclic_group_log2_size,
                                           C.c_void_p(out.ptr), None))
    dev.sync()
    return MultilinearExtension(dev, out, eq.num_vars, True)


class ChipTask:
    """One chip proof for the lane scheduler (ChipTask, ceno_zkvm/src/scheme/scheduler.rs:113-145): `payload` is
    whatever the execute callback needs; `booked_memory_bytes` defaults to the estimate."""

    def __init__(self, task_id, estimated_memory_bytes, payload=None, booked_memory_bytes=0, circuit_name=""):
        self.task_id, self.estimated_memory_bytes, self.booked_memory_bytes = task_id, estimated_memory_bytes, booked_memory_bytes
        self.payload, self.circuit_name = payload, circuit_name


class ChipScheduler:
    """ChipScheduler::execute (ceno_zkvm/src/scheme/scheduler.rs:205-250): run `execute_task(task, lane_id, stream)` for
    every task on up to `lanes` (1..8, default 4) concurrent lanes with greedy memory backfilling.  Returns
    (outputs ordered by task_id, telemetry ordered by task_id).  `dev=None` schedules host-only work (no streams).
    Exceptions raised by a task are re-raised after the in-flight tasks have drained, like the reference returns
    the first task error."""

    DEFAULT_LANES = 4

    def __init__(self, dev=None):
        self.dev = dev
        self.lib = dev.lib if dev is not None else _lib.load()

    def execute(self, tasks, execute_task, lanes=0, mem_budget_bytes=0):
        n = len(tasks)
        arr = (_lib.CgSchedTask * max(n, 1))()
        for i, t in enumerate(tasks):
            arr[i].task_id, arr[i].estimated_memory_bytes, arr[i].booked_memory_bytes = t.task_id, t.estimated_memory_bytes, t.booked_memory_bytes
        res = (_lib.CgSchedResult * max(n, 1))()
        outputs, errors = {}, {}

        def tramp(_user, index, task_id, lane_id, stream):
            try:
                outputs[task_id] = execute_task(tasks[index], lane_id, stream)
                return _lib.CG_OK
            except CenoB200Error as e:
                errors[task_id] = e
                return e.code
            except Exception as e:   # noqa: BLE001 — a task failure must not unwind through the C frames
                errors[task_id] = e
                return 6

        cb = _lib.SCHED_FN(tramp)
        ctx = self.dev.ctx if self.dev is not None else None
        rc = self.lib.cg_sched_execute(ctx, arr, n, lanes, mem_budget_bytes, cb, None, res)
        telemetry = [{"task_id": r.task_id, "lane_id": r.lane_id, "status": r.status, "launch_seq": r.launch_seq,
                      "booked_total_at_launch": r.booked_total_at_launch, "queue_delay_ms": r.queue_delay_ms,
                      "host_execution_ms": r.host_execution_ms, "event_wait_ms": r.event_wait_ms} for r in res[:n]]
        if errors:
            first = min(errors, key=lambda tid: next(t["launch_seq"] for t in telemetry if t["task_id"] == tid))
            raise errors[first]
        if rc != _lib.CG_OK:
            msg = self.lib.cg_last_error(ctx).decode() if ctx else "cg_sched_execute failed (deadlock: remaining tasks are too big for the memory pool, or bad arguments)"
            raise CenoB200Error(rc, msg)
        return [outputs[t["task_id"]] for t in telemetry], telemetry


class Stream:
    """A lane stream (one non-blocking CUDA stream per OS thread, gkr_iop/src/gpu/mod.rs:79-154)."""

    def __init__(self, dev):
        self.dev = dev
        h = C.c_void_p()
        dev.check(dev.lib.cg_stream_create(dev.ctx, C.byref(h)))
        self.handle = h.value

    def sync(self):
        self.dev.check(self.dev.lib.cg_stream_sync(self.dev.ctx, C.c_void_p(self.handle)))

    def close(self):
        if self.handle:
            self.dev.lib.cg_stream_destroy(self.dev.ctx, C.c_void_p(self.handle))
            self.handle = None


# ------------------------------------------------------------------- NTT / RS-encode (a9, f-2)
NTT_INVERSE, NTT_BITREV, NTT_EXT = 1, 2, 4


def ntt(dev, buf, log_n, n_cols=1, col_stride=None, inverse=False, bitrev=False, ext=False, stream=None):
    """In-place batched NTT of `n_cols` columns held in DeviceBuffer `buf` (p3 Radix2 semantics, see cg_ntt)."""
    flags = (NTT_INVERSE if inverse else 0) | (NTT_BITREV if bitrev else 0) | (NTT_EXT if ext else 0)
    dev.check(dev.lib.cg_ntt(dev.ctx, C.c_void_p(buf.ptr), log_n, n_cols, col_stride if col_stride is not None else (1 << log_n), flags,
                             C.c_void_p(stream) if stream else None))
    if not stream:
        dev.sync()


def rs_encode(dev, msg_buf, width, log_n, rate_log, bitrev=True, stream=None):
    """Reed-Solomon encode every column of a column-major message matrix; returns the code matrix DeviceBuffer
    (width x 2^(log_n+rate_log), column-major)."""
    code = dev.alloc(8 * (width << (log_n + rate_log)))
    dev.check(dev.lib.cg_rs_encode(dev.ctx, C.c_void_p(msg_buf.ptr), width, log_n, rate_log, C.c_void_p(code.ptr), NTT_BITREV if bitrev else 0,
                                   C.c_void_p(stream) if stream else None))
    if not stream:
        dev.sync()
    return code


def basefold_style_commit(dev, msg_buf, width, log_n, rate_log=1, stream=None):
    """TraceCommitter::commit_traces shape (ceno_zkvm/src/scheme/cpu/mod.rs:559-584): RS-encode the witness columns,
    Merkle-hash the codeword matrix row-wise.  Returns (code DeviceBuffer, tree DeviceBuffer, root[4]).  Arrangement
    parity is unpinned: bit-reversed codeword rows, one leaf per row across all columns."""
    code = rs_encode(dev, msg_buf, width, log_n, rate_log, bitrev=True, stream=stream)
    tree, root = merkle_commit(dev, code, width, 1 << (log_n + rate_log), col_major=True, stream=stream)
    return code, tree, root


# ------------------------------------------------------------------- f-2 / a9: Basefold commit + batch_open
class BasefoldParams:
    """BasefoldSpec parameters (rate_log / number of queries / proof-of-work bits are upstream-only: parameters here)."""

    def __init__(self, rate_log=1, n_queries=8, pow_bits=0):
        self.rate_log, self.n_queries, self.pow_bits = rate_log, n_queries, pow_bits

    def c(self):
        return _lib.CgBasefoldParams(self.rate_log, self.n_queries, self.pow_bits, 0)


class BasefoldCommitment:
    """PCS::CommitmentWithWitness on the device (TraceCommitter::commit_traces, ceno_zkvm/src/scheme/cpu/mod.rs:559-584):
    the RS-encoded columns (bit-reversed rows) and their Merkle tree.  `msg_buf` (width x 2^num_vars base elements,
    column-major) must stay alive until the opening."""

    def __init__(self, dev, msg_buf, width, num_vars, params, stream=None):
        self.dev, self.msg, self.width, self.num_vars, self.params = dev, msg_buf, width, num_vars, params
        self.h = C.c_void_p()
        cp = params.c()
        dev.check(dev.lib.cg_basefold_commit(dev.ctx, C.c_void_p(msg_buf.ptr), width, num_vars, C.byref(cp), C.c_void_p(stream) if stream else None,
                                             C.byref(self.h)))
        root = np.zeros(4, np.uint64)
        dev.check(dev.lib.cg_basefold_commitment_root(self.h, _vp(root)))
        self.root = root

    def codeword(self):
        code = C.c_void_p()
        self.dev.check(self.dev.lib.cg_basefold_commitment_codeword(self.h, C.byref(code), None))
        n = self.width << (self.num_vars + self.params.rate_log)
        return DeviceBuffer(self.dev, code.value, 8 * n, owner=False).to_host().reshape(self.width, -1)

    def free(self):
        if self.h:
            self.dev.lib.cg_basefold_commitment_free(self.h)
            self.h = C.c_void_p()


def basefold_batch_open(dev, commits, points, evals, params, transcript, stream=None):
    """OpeningProver::open -> PCS::batch_open (ceno_zkvm/src/scheme/cpu/mod.rs:1415-1457): one opening (point, evals) per
    commitment.  `transcript` is a StandInTranscript.  Returns the proof parsed into the structure of BasefoldProof:
    {sumcheck: [(p1, p2)], commits: [digest], final_message: [[ext]], pow_witness, queries: [{index, inputs, commit_phase}]}."""
    n = len(commits)
    keep = []
    ops = (_lib.CgBasefoldOpening * n)()
    for i, (cm, pt, ev) in enumerate(zip(commits, points, evals)):
        pt, ev = _u64(pt), _u64(ev)
        keep += [pt, ev]
        ops[i].commit, ops[i].h_point_ext, ops[i].h_evals_ext = cm.h.value, pt.ctypes.data, ev.ctypes.data
    cp = params.c()
    words = dev.lib.cg_basefold_proof_len(ops, n, C.byref(cp))
    assert words > 0
    proof = np.zeros(words, np.uint64)
    vt = _lib.CgPcsTranscriptVt()
    dev.lib.cg_standin_pcs_vt(_vp(transcript.state), C.byref(vt))
    dev.check(dev.lib.cg_basefold_batch_open(dev.ctx, ops, n, C.byref(cp), C.byref(vt), _vp(proof), words, C.c_void_p(stream) if stream else None))
    # ---- parse
    rate = params.rate_log
    max_nv = max(c.num_vars for c in commits)
    w = [0]

    def take(k):
        out = [int(x) for x in proof[w[0]:w[0] + k]]
        w[0] += k
        return out

    def ext():
        return tuple(take(2))
    out = {"sumcheck": [(ext(), ext()) for _ in range(max_nv)], "commits": [take(4) for _ in range(max_nv)],
           "final_message": [[ext()] for _ in range(n)], "pow_witness": take(1)[0], "queries": []}
    for _ in range(params.n_queries):
        q = {"index": take(1)[0], "inputs": [], "commit_phase": []}
        for c in commits:
            q["inputs"].append({"opened": take(c.width), "path": [take(4) for _ in range(c.num_vars + rate)]})
        for r in range(max_nv):
            q["commit_phase"].append({"sibling": ext(), "path": [take(4) for _ in range(max_nv + rate - r - 1)]})
        out["queries"].append(q)
    assert w[0] == words
    return out


# ------------------------------------------------------------------- f-3: EC-sum Quark prover (cpu/mod.rs:72-316)
SEPTIC_EXTENSION_DEGREE = 7


def ecc_quark_selectors(dev, out_rt, num_instances):
    """(sel_add, sel_bypass, sel_export) as device ext MLEs (cg_ecc_quark_selectors)."""
    out_rt = _u64(out_rt)
    n = out_rt.size // 2
    bufs = [dev.alloc(16 << n) for _ in range(3)]
    dev.check(dev.lib.cg_ecc_quark_selectors(dev.ctx, _vp(out_rt), n, num_instances, *[C.c_void_p(b.ptr) for b in bufs], None))
    dev.sync()
    return [MultilinearExtension(dev, b, n, True) for b in bufs]


def split_even_odd(dev, mles):
    """filter_bj: ([v[2b]], [v[2b+1]]) for base MLEs (cg_split_even_odd)."""
    nv = mles[0].num_vars
    ev = [dev.alloc(8 << (nv - 1)) for _ in mles]
    od = [dev.alloc(8 << (nv - 1)) for _ in mles]
    descs = (_lib.CgMleDesc * len(mles))(*[m.desc() for m in mles])
    pe = (C.c_void_p * len(mles))(*[b.ptr for b in ev])
    po = (C.c_void_p * len(mles))(*[b.ptr for b in od])
    dev.check(dev.lib.cg_split_even_odd(dev.ctx, descs, len(mles), pe, po, None))
    dev.sync()
    return ([MultilinearExtension(dev, b, nv - 1, False) for b in ev], [MultilinearExtension(dev, b, nv - 1, False) for b in od])


class EccQuarkProver:
    """CpuEccProver::create_ecc_proof (ceno_zkvm/src/scheme/cpu/mod.rs:72-316; trait EccQuarkProver, hal.rs:164-171):
    accumulate 2^n EC points (affine, septic extension coordinates) in one Quark-style layer.  xs / ys / invs are 7
    base-field device MLEs each with n+1 variables.  Returns the EccQuarkProof fields as a dict."""

    @staticmethod
    def build_terms(alpha_pows, final_sum_x, final_sum_y):
        """The zerocheck expression in monomial form over the MLE order
        [sel_add, sel_bypass, sel_export, s(7), x0(7), y0(7), x1(7), y1(7), x3(7), y3(7)]  (:153-262), expanded by the
        library (cg_ecc_quark_terms, host-side C++)."""
        lib = _lib.load()
        al = _u64(np.asarray(alpha_pows, dtype=np.uint64).reshape(-1))
        fx, fy = _u64(np.array([int(v) for v in final_sum_x], dtype=np.uint64)), _u64(np.array([int(v) for v in final_sum_y], dtype=np.uint64))
        assert al.size == 2 * 49 and fx.size == 7 and fy.size == 7
        nt, ni = C.c_uint32(), C.c_uint32()
        rc = lib.cg_ecc_quark_terms(_vp(al), _vp(fx), _vp(fy), None, None, None, 0, 0, C.byref(nt), C.byref(ni))
        if rc != _lib.CG_OK:
            raise CenoB200Error(rc, "cg_ecc_quark_terms failed")
        coeff, off, idx = np.zeros(2 * nt.value, np.uint64), np.zeros(nt.value + 1, np.uint32), np.zeros(max(ni.value, 1), np.uint32)
        rc = lib.cg_ecc_quark_terms(_vp(al), _vp(fx), _vp(fy), _vp(coeff), _vp(off), _vp(idx), nt.value, ni.value, C.byref(nt), C.byref(ni))
        if rc != _lib.CG_OK:
            raise CenoB200Error(rc, "cg_ecc_quark_terms failed")
        return [([int(coeff[2 * t]), int(coeff[2 * t + 1])], [int(i) for i in idx[off[t]:off[t + 1]]]) for t in range(nt.value)]

    @staticmethod
    def build_terms_symbolic(alpha_pows, final_sum_x, final_sum_y):
        """The same table through the symbolic engine of expr.py (cross-check of the C++ expansion in the host tests)."""
        from .expr import Poly, SymbolicSepticExtension as Sep, ext
        D = SEPTIC_EXTENSION_DEGREE
        sel_add, sel_bypass, sel_export = Poly.var(0), Poly.var(1), Poly.var(2)
        grp = lambda g: Sep([Poly.var(3 + D * g + i) for i in range(D)])
        s, x0, y0, x1, y1, x3, y3 = (grp(g) for g in range(7))
        al = iter([ext(int(a[0]), int(a[1])) for a in alpha_pows])
        comb = lambda exprs: sum((e * next(al) for e in exprs), Poly())
        add = Poly()
        add = add + comb((s * (x0 - x1) - (y0 - y1)).to_exprs())            # slope
        add = add + comb(((s * s) - x0 - x1 - x3).to_exprs())               # x3 = s^2 - x0 - x1
        add = add + comb((s * (x0 - x3) - (y0 + y3)).to_exprs())            # y3 = s (x0 - x3) - y0
        byp = comb((x3 - x0).to_exprs()) + comb((y3 - y0).to_exprs())
        exp = comb([x - Poly.const(int(f)) for x, f in zip(x3.to_exprs() + y3.to_exprs(), list(final_sum_x) + list(final_sum_y))])
        return (add * sel_add + byp * sel_bypass + exp * sel_export).terms()

    @staticmethod
    def create_ecc_proof(dev, num_instances, xs, ys, invs, transcript, flags=0):
        D = SEPTIC_EXTENSION_DEGREE
        assert len(xs) == D and len(ys) == D and len(invs) == D
        n = xs[0].num_vars - 1
        out_rt = transcript.sample_and_append_vec(b"ecc", n)
        alpha_pows = transcript.sample_and_append_challenge_pows(D * 3 + D * 2 + D * 2, b"ecc_alpha")
        sels = ecc_quark_selectors(dev, out_rt, num_instances)
        x0, x1 = split_even_odd(dev, xs)
        y0, y1 = split_even_odd(dev, ys)
        x3, y3, s = ([m.as_view_slice(2, 1) for m in grp] for grp in (xs, ys, invs))   # x[1,b], y[1,b], s[1,b]
        last = (1 << n) - 2                                            # the final sum sits at [1,..,1,0]
        fx = [int(m.buf.to_host(8, 8 * last)[0]) for m in x3]
        fy = [int(m.buf.to_host(8, 8 * last)[0]) for m in y3]
        terms = EccQuarkProver.build_terms(alpha_pows, fx, fy)
        mles = sels + s + x0 + y0 + x1 + y1 + x3 + y3
        rounds, evals, rt = IOPProverState.prove(dev, mles, terms, n, 3, transcript=transcript, flags=flags)
        for m in sels + x0 + x1 + y0 + y1:
            m.free()
        assert evals.shape[0] == 3 + D * 7
        return {"zerocheck_proof": rounds, "num_instances": num_instances, "evals": evals, "rt": rt, "sum": (fx, fy)}
