"""The C-ABI library loads without a GPU and exports every symbol include/minife_b200.h declares (not gpu)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import minife_b200 as mf
from minife_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "minife_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mfe_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported(product_lib):
    names = declared_symbols()
    assert len(names) >= 30
    L = C.CDLL(mf.library_path())
    for n in names:
        assert hasattr(L, n), "libminife_b200.so does not export %s" % n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_box_partition_matches_oracle(product_lib, oracle_built):
    from oracle_lib import Oracle
    for n, npr in (((4, 4, 4), 1), ((5, 3, 7), 3), ((64, 64, 64), 8), ((400, 400, 400), 8), ((17, 3, 5), 6), ((256, 256, 512), 2), ((9, 9, 9), 7)):
        boxes = mf.box_partition(0, npr, 2, mf.Box([0, n[0], 0, n[1], 0, n[2]]))
        assert np.array_equal(np.stack([b.ranges for b in boxes]), Oracle.box_partition(*n, npr))


def test_no_cpu_fallback(product_lib):
    """Without a CUDA device every compute entry point must fail loudly."""
    if product_lib.mfe_device_count() > 0:
        pytest.skip("a GPU is present")
    g = mf.Box([0, 4, 0, 4, 0, 4])
    with pytest.raises(mf.LibraryError, match="no CUDA device"):
        mf.simple_mesh_description(g, g)


def test_product_does_not_touch_the_oracle():
    """Nothing under minife_b200/ may import, load or link oracle/ or the test harness."""
    bad = re.compile(r"oracle|cpu_harness|libminife_ref|libminife_oracle")
    for d, _, files in os.walk(os.path.join(ROOT, "minife_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                for i, line in enumerate(open(os.path.join(d, f), errors="replace")):
                    assert not bad.search(line) or "oracle" in line and ("//" in line or "#" in line or '"""' in line) and "import" not in line and "include" not in line, (f, i + 1, line)
