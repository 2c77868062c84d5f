"""CPU-side checks of the boundary: libskb.so loads and exports every symbol include/skb.h declares; the
Python extension imports, Hit behaves like the reference's (hit.rs:27-74), and the product refuses to run
without a CUDA device instead of falling back to anything."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "skb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(skb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from pyskani_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == names


def test_version_string():
    from pyskani_b200 import capi
    assert b"sm_100a" in capi.lib().skb_version()


def test_null_arguments_are_rejected_without_touching_a_device():
    from pyskani_b200 import capi
    L = capi.lib()
    assert L.skb_ctx_create(0, None) == capi.SKB_ERR_ARG
    assert L.skb_db_create(None, None) == capi.SKB_ERR_ARG
    assert L.skb_sketch_info(None, None) == capi.SKB_ERR_ARG
    assert L.skb_db_size(None) == 0


def test_package_surface_matches_reference():
    import pyskani_b200 as pyskani
    assert pyskani.__all__ == ["Sketch", "Database", "Hit", "SKANI_VERSION"]   # reference __init__.py:8-13
    assert pyskani.SKANI_VERSION == "0.3.0"
    for name in ("load", "open", "sketch", "query", "save", "flush", "path", "compression", "marker_compression",
                 "__enter__", "__exit__"):
        assert hasattr(pyskani.Database, name)
    for name in ("identity", "query_name", "query_fraction", "reference_name", "reference_fraction"):
        assert hasattr(pyskani.Hit, name)
    for name in ("name", "c", "amino_acid"):
        assert hasattr(pyskani.Sketch, name)


def test_hit_constructor_validation():
    import pyskani_b200 as pyskani
    h = pyskani.Hit(0.5, "q", 0.25, "r", 0.75)
    assert (h.identity, h.query_name, h.query_fraction, h.reference_name, h.reference_fraction) == (0.5, "q", 0.25, "r", 0.75)
    assert repr(h) == "Hit(identity=0.5, query_name='q', query_fraction=0.25, reference_name='r', reference_fraction=0.75)"
    for bad in ((-0.1, 0.5, 0.5), (1.1, 0.5, 0.5), (0.5, -1.0, 0.5), (0.5, 0.5, 2.0)):
        with pytest.raises(ValueError):
            pyskani.Hit(bad[0], "q", bad[1], "r", bad[2])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import pyskani_b200 as pyskani
    from pyskani_b200 import capi
    with pytest.raises(RuntimeError):
        pyskani.Database()
    with pytest.raises(capi.SkbError):
        capi.Context(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pyskani_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "liboracle" not in text and "skani_oracle" not in text.replace("oracle/skani_oracle.cpp", ""), f
