"""CPU-side checks of the drop-in boundary: libabr.so loads, exports every
symbol include/abr.h declares, and fails loudly without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from aboria_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _lib


def test_header_symbols_are_exported():
    _lib = _ensure_built()
    header = open(os.path.join(ROOT, "include", "abr.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(abr_[a-z0-9_]+)\s*\(", header))
    declared -= {"abr_launch_fn"}
    assert len(declared) >= 20
    L = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    # the ctypes table covers the same set
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_version_and_loud_failure_without_gpu():
    _lib = _ensure_built()
    L = _lib.lib()
    assert b"sm_100a" in L.abr_version()
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = L.abr_create(C.byref(h), 0, None)
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in L.abr_last_error_string(None)
    import aboria_b200

    with pytest.raises(aboria_b200.AbrError):
        aboria_b200.Particles(3, 10)


def test_product_does_not_import_oracle():
    # the product package and the CUDA sources must not reference oracle/
    bad = []
    for base in ("aboria_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(^|\W)(from|import)\s+oracle|oracle/|liboria_oracle|libaboria_oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_sass_is_sm100a_and_unfused():
    """the shipped cubin targets sm_100a and the distance predicate is not
    contracted (no DFMA between the three DMULs and the compare is acceptable,
    but the library must be built with -fmad=false)."""
    mk = open(os.path.join(ROOT, "aboria_b200", "csrc", "Makefile")).read()
    assert "-fmad=false" in mk and "compute_100a" in mk
