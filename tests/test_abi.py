"""The C-ABI library builds, loads and exports every symbol include/parcop_b200.h declares.
No compute calls: this runs without a GPU."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "parcop_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from pyranda_b200 import build, _lib
    so = build.build()
    L = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    # the Python binding declares exactly the header's entry points
    assert sorted(_lib.EXPORTS) == names


def test_version_and_error_strings():
    from pyranda_b200 import _lib
    L = _lib.load()
    assert b"sm_100a" in L.pb_version()
    assert L.pb_launch_count() >= 0


def test_bad_arguments_are_refused_without_a_gpu():
    from pyranda_b200 import _lib
    L = _lib.load()
    h = ctypes.c_void_p()
    args = [64, 64, 64, 1, 1, 1, 0, 0, 0, 0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0]
    rc = L.pb_plan_create(ctypes.byref(h), 64, 64, 63, 1, 1, 2, 0, 0, 0, 0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0,
                          b"NONE", b"NONE", b"NONE", b"NONE", b"NONE", b"NONE", -1)
    assert rc == -1 and b"divisible" in L.pb_last_error()
    rc = L.pb_plan_create(ctypes.byref(h), *args, b"WALL", b"NONE", b"NONE", b"NONE", b"NONE", b"NONE", -1)
    assert rc == -1 and b"SYMM" in L.pb_last_error()
    rc = L.pb_plan_create(ctypes.byref(h), 64, 64, 64, 2, 1, 1, 0, 0, 0, 0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0,
                          b"NONE", b"NONE", b"NONE", b"NONE", b"NONE", b"NONE", -1)
    assert rc == -2


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import / load anything under oracle/."""
    pkg = os.path.join(ROOT, "pyranda_b200")
    bad = re.compile(r"^\s*(import\s+oracle|from\s+oracle|from\s+\.+\s*oracle)|libparcop_oracle|oracle[/\\.]parcop|[\"']oracle[\"']", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), "%s reaches into oracle/" % f
