"""CPU check of the device field layer's algebra (ceno_b200/csrc/gl64.cuh compiled for the host):
weak 160-bit reduction, canonicalisation, sub/add, lazy dot products — against Python big ints."""
import ctypes as C
import os
import random
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
P = 0xFFFFFFFF00000001
M64 = (1 << 64) - 1


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(HERE, "host", "libfield_host.so")
    src = os.path.join(HERE, "host", "field_host.cpp")
    hdr = os.path.join(HERE, "..", "ceno_b200", "csrc", "gl64.cuh")
    hdr2 = os.path.join(HERE, "..", "ceno_b200", "csrc", "poseidon2.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(hdr2)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-o", so, src])
    L = C.CDLL(so)
    for f in ("h_reduce_weak", "h_canon", "h_sub", "h_add", "h_mul", "h_mul7_weak"):
        getattr(L, f).restype = C.c_uint64
    L.h_reduce_weak.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    L.h_canon.argtypes = [C.c_uint64]
    L.h_mul7_weak.argtypes = [C.c_uint64]
    for f in ("h_sub", "h_add", "h_mul"):
        getattr(L, f).argtypes = [C.c_uint64, C.c_uint64]
    return L


EDGE = [0, 1, 2, 7, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000, P - 1, P - 2, P, P + 1, M64, M64 - 1, 1 << 63, (1 << 63) - 1,
        0xFFFFFFFE00000001, 0x00000001FFFFFFFF, 0xFFFFFFFFFFFFFFFF - 0xFFFFFFFF]


def test_reduce_weak_and_canon(lib):
    rng = random.Random(1)
    cases = [(a, b, c) for a in EDGE for b in EDGE for c in (0, 1, 2, 3, 1000, (1 << 31) - 1)]
    cases += [(rng.randrange(1 << 64), rng.randrange(1 << 64), rng.randrange(1 << 20)) for _ in range(50000)]
    for s0, s1, s2 in cases:
        x = s0 + (s1 << 64) + (s2 << 128)
        w = lib.h_reduce_weak(s0, s1, s2)
        assert w % P == x % P, (hex(s0), hex(s1), s2)
        assert lib.h_canon(w) == x % P


def test_sub_add_mul_canonical(lib):
    rng = random.Random(2)
    canon_edge = [e for e in EDGE if e < P]
    pairs = [(a, b) for a in canon_edge for b in canon_edge] + [(rng.randrange(P), rng.randrange(P)) for _ in range(50000)]
    for a, b in pairs:
        assert lib.h_sub(a, b) == (a - b) % P
        assert lib.h_add(a, b) == (a + b) % P
    for a, b in [(a, b) for a in EDGE for b in EDGE] + [(rng.randrange(1 << 64), rng.randrange(1 << 64)) for _ in range(50000)]:
        assert lib.h_mul(a, b) == a * b % P          # any u64 operands
    for a in EDGE + [rng.randrange(1 << 64) for _ in range(10000)]:
        assert lib.h_mul7_weak(a) % P == 7 * a % P


def test_aligned_limb_accumulators(lib):
    """acc_t (E/M limb columns) + cacc_t (compact) against big-int dot products, worst-case carries included."""
    lib.h_cacc_dot.restype = C.c_uint64
    rng = random.Random(7)
    for n, per in [(1, 1), (2, 2), (5, 2), (64, 4), (1000, 7), (4096, 4096)]:
        for mode in ("max", "rand", "edge"):
            if mode == "max":
                a = [M64] * n
                b = [M64] * n
            elif mode == "rand":
                a = [rng.randrange(1 << 64) for _ in range(n)]
                b = [rng.randrange(1 << 64) for _ in range(n)]
            else:
                a = [rng.choice(EDGE) for _ in range(n)]
                b = [rng.choice(EDGE) for _ in range(n)]
            got = lib.h_cacc_dot((C.c_uint64 * n)(*a), (C.c_uint64 * n)(*b), n, per)
            assert got == sum(x * y for x, y in zip(a, b)) % P, (n, per, mode)


def emul(a, b):
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def test_ext_mul_fma_and_lazy_dot(lib):
    rng = random.Random(3)
    A2 = C.c_uint64 * 2

    def call2(fn, *args):
        o = A2()
        fn(*[A2(*x) for x in args], o)
        return (o[0], o[1])
    pool = EDGE + [rng.randrange(1 << 64) for _ in range(40)]
    for _ in range(20000):
        a = (rng.choice(pool), rng.choice(pool))
        b = (rng.choice(pool), rng.choice(pool))
        assert call2(lib.h_ext_mul, a, b) == emul(a, b)
        x = (rng.randrange(P), rng.randrange(P))
        got = call2(lib.h_ext_fma, x, a, b)
        m = emul(a, b)
        assert got == ((x[0] + m[0]) % P, (x[1] + m[1]) % P)
    for n in (1, 2, 3, 17, 200):
        a = [rng.choice(pool) if rng.random() < 0.3 else M64 for _ in range(2 * n)]   # worst-case carries
        b = [rng.choice(pool) if rng.random() < 0.3 else M64 for _ in range(2 * n)]
        o = A2()
        lib.h_eacc_dot((C.c_uint64 * (2 * n))(*a), (C.c_uint64 * (2 * n))(*b), n, o)
        acc = (0, 0)
        for i in range(n):
            m = emul((a[2 * i], a[2 * i + 1]), (b[2 * i], b[2 * i + 1]))
            acc = ((acc[0] + m[0]) % P, (acc[1] + m[1]) % P)
        assert (o[0], o[1]) == acc


def test_poseidon2_device_code_on_host_matches_oracle(lib):
    """ceno_b200/csrc/poseidon2.cuh compiled for the host vs the oracle's restatement (same placeholder constants)."""
    import numpy as np
    from oracle import oracle as orc
    for variant in (0, 1):
        prm = orc.p2_params(seed=9, mds_variant=variant)
        rng = random.Random(variant)
        for st in ([0] * 8, [P - 1] * 8, [rng.randrange(P) for _ in range(8)], [rng.randrange(P) for _ in range(8)]):
            arr = (C.c_uint64 * 8)(*st)
            lib.h_p2_permute(C.byref(prm), arr)
            assert list(arr) == [int(x) for x in orc.poseidon2_permute(prm, np.array(st, dtype=np.uint64))]
