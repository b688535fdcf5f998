"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/qpalette.h declares,
and the Python binding table mirrors the header (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "qpalette.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from qpalette import _cabi
    h = ctypes.CDLL(_cabi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(h, s), f"{s} declared in include/qpalette.h but not exported"
    assert sorted(_cabi.SIGNATURES) == syms, "qpalette/_cabi.py SIGNATURES out of sync with include/qpalette.h"
    assert _cabi.lib().qp_version() >= 100


def test_no_oracle_in_product_path():
    """the product package must never import the oracle (a CPU fallback would void every parity claim)."""
    pkg = os.path.join(ROOT, "q-palette_b200")
    for dp, _dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) and f != "host_emul.cpp":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "qp_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, (dp, f)
