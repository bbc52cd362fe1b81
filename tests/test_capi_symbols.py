"""CPU: the C-ABI library loads without a GPU and exports every symbol include/ovdet_b200.h declares
(no compute calls here); the ctypes table mirrors the header; the product never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ovdet_b200.h")
PKG = os.path.join(ROOT, "open-vocabulary-3d-object-detection_b200")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ovdet_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    so = g.build()
    lib = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ovdet_b200.h but not exported"
    lib.ovdet_version.restype = ctypes.c_int
    assert lib.ovdet_version() >= 1
    lib.ovdet_last_error.restype = ctypes.c_char_p
    assert lib.ovdet_last_error() is not None


def test_ctypes_table_matches_header():
    import ovdet_b200  # noqa: F401
    from ovdet_b200 import _capi
    assert sorted(_capi.SIGNATURES.keys()) == declared_symbols()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _capi.SIGNATURES.items():
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, ctypes table {len(args)}"


def test_invalid_arguments_fail_loudly_without_gpu():
    import ovdet_b200  # noqa: F401
    from ovdet_b200 import _capi
    L = _capi.lib()
    rc = L.ovdet_giou3d_f32(None, None, None, 1, 1, 1, 0, 0, None, None)
    assert rc == -1 and b"null" in L.ovdet_last_error()
    with pytest.raises(_capi.OvdetError):
        _capi.check(rc)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), os.path.join(dirpath, f)
                assert "ovdet_oracle" not in txt


def test_flag_words_match_header():
    """Every `#define OVDET_<NAME> <int>` of the header has the same value as `_capi.<NAME>` (the host shims pass these)."""
    import re
    from ovdet_b200 import _capi as C
    text = open(os.path.join(ROOT, "include", "ovdet_b200.h")).read()
    seen = 0
    for name, val in re.findall(r"#define OVDET_([A-Z0-9_]+) \(?(-?(?:0x[0-9a-fA-F]+|\d+))u?\)?", text):
        if name in ("B200_H", "OK") or name.startswith("ERR_"):
            continue
        assert hasattr(C, name), "include/ovdet_b200.h defines OVDET_%s but _capi.py has no %s" % (name, name)
        assert getattr(C, name) == int(val, 0), (name, getattr(C, name), val)
        seen += 1
    assert seen >= 20
