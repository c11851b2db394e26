"""The C-ABI library loads and exports every symbol include/opengjk_b200.h declares; struct mirrors have the
reference's layouts.  No compute calls (no GPU here)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "opengjk_b200.h")).read()
    plain = set(re.findall(r"\b(ogjk_[a-z_]+)\s*\(", text))
    plain = {s for s in plain if "##" not in s}
    templ = set(re.findall(r"ogjk_##P##_([a-z_]+)\s*\(", text))
    syms = {s for s in plain if not s.startswith("ogjk_f")}
    for p in ("f32", "f64"):
        syms |= {f"ogjk_{p}_{t}" for t in templ}
    return syms


def test_library_exports_declared_symbols(pkg):
    lib = pkg.load_library()
    syms = declared_symbols()
    assert len(syms) >= 2 * 22 + 7
    missing = [s for s in sorted(syms) if not hasattr(lib, s)]
    assert not missing, missing
    assert b"sm_100a" in lib.ogjk_version()


def test_struct_layouts_match_reference(pkg):
    """SURVEY.md Appendix B (offsetof probes of the reference structs)."""
    p32, p64 = pkg.polytope_dtype(np.float32), pkg.polytope_dtype(np.float64)
    s32, s64 = pkg.simplex_dtype(np.float32), pkg.simplex_dtype(np.float64)
    assert (p32.itemsize, p64.itemsize, s32.itemsize, s64.itemsize) == (32, 48, 108, 184)
    assert p32.fields["coord"][1] == 24 and p64.fields["coord"][1] == 40
    assert s32.fields["witnesses"][1] == 84 and s64.fields["witnesses"][1] == 136
    assert pkg.PAIR_DTYPE.itemsize == 8


def test_no_device_fails_loudly(pkg):
    """Without a usable GPU a compute call must raise, never fall back to the CPU."""
    lib = pkg.load_library()
    if lib.ogjk_device_count() > 0:
        return
    eng = pkg.Engine(np.float32)
    a, b = pkg.workloads.random_pairs(4, 8, 2.0, seed=1)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    try:
        eng.compute_minimum_distance(bd1, bd2)
    except pkg.OgjkError as e:
        assert "cuda" in str(e).lower()
    else:
        raise AssertionError("compute call succeeded without a GPU")


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under opengjk-gpu_b200/ or include/ may reference it."""
    for base in ("opengjk-gpu_b200", "include"):
        for dirpath, _d, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "pyoracle" not in text and "ogjk_oracle" not in text and "oracle/" not in text, (dirpath, f)
