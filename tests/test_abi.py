"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header
declares, and fails loudly (no fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

from ceno_b200 import _lib
from ceno_b200 import build as cbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    cbuild.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ceno_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", src)) - {"cg_challenge_cb"})


def test_header_and_loader_agree(lib):
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    raw = C.CDLL(_lib.SO)
    for name in header_symbols():
        assert hasattr(raw, name), name


def test_version_and_no_torch_types_in_abi(lib):
    assert b"sm_100a" in lib.cg_version()
    hdr = open(os.path.join(ROOT, "include", "ceno_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)   # declarations only
    assert "torch" not in code and "at::" not in code and "Tensor" not in code


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure; the product path must never reach it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ceno_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle is test infrastructure", ""), os.path.join(dirpath, f)


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    assert lib.cg_init(0, C.byref(ctx)) == 5  # CG_ERR_NO_DEVICE
    from ceno_b200 import CenoB200Error, Device
    with pytest.raises(CenoB200Error):
        Device(0)


def test_standin_transcript_host_matches_oracle():
    # the stand-in sponge is restated independently in the library (host+device) and in the oracle
    import numpy as np
    from ceno_b200 import StandInTranscript
    from oracle import oracle as orc
    a, b = StandInTranscript(b"abc"), orc.Transcript(b"abc")
    a.append_message(b"hello world!!"); b.append_message(b"hello world!!")
    e = np.array([1, 2, 3, 4], dtype=np.uint64)
    a.append_field_element_exts(e); b.append_ext(e)
    assert tuple(a.sample_and_append_challenge(b"Internal round")) == tuple(b.sample(b"Internal round"))
    assert int(a.state[0]) == b.state


def _build_c_smoke():
    import subprocess
    from oracle import oracle as orc
    orc.build()
    exe = os.path.join(ROOT, "tests", "c", "abi_smoke")
    src = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
    libdir, odir = os.path.join(ROOT, "ceno_b200", "lib"), os.path.join(ROOT, "oracle")
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-o", exe, src, "-L" + libdir, "-lceno_b200", "-L" + odir, "-lceno_oracle",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + odir])
    return exe


def test_c_consumer_links_against_the_abi(lib):
    """A plain-C program (what a Rust extern "C" block sees) compiles and links against the header + .so.
    Without a GPU it must report that cg_init refuses (exit 77) — no CPU fallback."""
    import subprocess
    import torch
    exe = _build_c_smoke()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_consumer_bit_exact_on_gpu(lib):
    import subprocess
    exe = _build_c_smoke()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "bit-exact" in r.stdout, r.stdout + r.stderr


def test_integration_doc_binds_only_declared_symbols():
    """Every `pub fn cg_*` of the Rust extern block in INTEGRATION.md and every cg_* call named in DESIGN.md is a symbol the
    header declares (the documents cannot drift from the ABI)."""
    import re
    declared = set(header_symbols())
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    integ = open(os.path.join(root, "INTEGRATION.md")).read()
    rust = set(re.findall(r"pub fn (cg_[a-z0-9_]+)", integ))
    assert rust and rust <= declared, sorted(rust - declared)
    design = open(os.path.join(root, "DESIGN.md")).read()
    named = {n for n in re.findall(r"`(cg_[a-z0-9_]+)`", design) if not n.endswith("_")}
    structs = {"cg_mle_desc", "cg_challenge_cb", "cg_transcript_vt", "cg_tower_spec", "cg_sched_task", "cg_sched_result", "cg_stream", "cg_ctx",
               "cg_pcs_transcript_vt", "cg_tower_vspec", "cg_tower_vgroup", "cg_basefold_params", "cg_basefold_opening", "cg_pcs_commitment"}
    assert named - structs <= declared, sorted(named - structs - declared)


def test_rust_bindings_are_generated_from_the_header():
    """bindings/rust/ceno_b200_sys/src/lib.rs (what a ceno maintainer links against) is generated from the header: it must be up to
    date and declare exactly the header's functions."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_bindings.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rs = open(os.path.join(ROOT, "bindings", "rust", "ceno_b200_sys", "src", "lib.rs")).read()
    assert sorted(re.findall(r"pub fn (cg_[a-z0-9_]+)\(", rs)) == header_symbols()
    for st in ("cg_mle_desc", "cg_tower_spec", "cg_tower_vspec", "cg_transcript_vt", "cg_pcs_transcript_vt", "cg_basefold_opening", "cg_sched_task"):
        assert f"pub struct {st} " in rs
