"""Large single-GPU parity (-m gpu), kept in a file of its own that sorts last: 400^3 (BASELINE configs[2] at one rank, 27 GB on
the device) against the record of the UNMODIFIED reference at that size (tests/golden/big_400x400x400_np1.npz, made by
tests/golden/make_big_golden.py from ref/src over the MPI shim at np = 1: CRC-32 of the arrays after every phase and the four
per-iteration histories of the 200-iteration solve).  Everything is compared bit for bit."""
import os
import zlib

import numpy as np
import pytest

import minife_b200 as mf
from oracle_lib import EPS

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big_400x400x400_np1.npz")
ORACLE_512 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_512x512x512_np1.npz")


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.int64)


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden file missing")
def test_400_cubed_equals_the_reference_record(product_lib):
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a 60 GB device")
    g = np.load(GOLDEN)
    n = tuple(int(v) for v in g["n"])
    assert n == (400, 400, 400) and int(g["np"]) == 1
    gbox = mf.Box([0, n[0], 0, n[1], 0, n[2]])
    mesh = mf.simple_mesh_description(gbox, gbox)
    try:
        A = mf.CSRMatrix()
        mf.generate_matrix_structure(mesh, A)
        nrows, nnz = A.num_rows(), A.num_nonzeros()
        assert (nrows, nnz) == tuple(int(v) for v in g["sizes_0"])
        assert crc(A.row_offsets) == g["crc_row_offsets_0"]
        b = mf.Vector(0, nrows)
        x = mf.Vector(0, nrows)
        mf.assemble_FE_data(mesh, A, b)
        assert crc(b.coefs) == g["crc_b_prebc_0"], "b after assembly"
        mf.impose_dirichlet(0.0, A, b, bc_rows=0)
        mf.impose_dirichlet(1.0, A, b, bc_rows=1)
        assert crc(b.coefs) == g["crc_b_0"], "b after the Dirichlet conditions"
        mf.make_local_matrix(A)
        iters, normr, times, h = mf.cg_solve(A, b, x, mf.matvec_std(), 200, EPS, history=True)
        assert iters == len(g["normr"]) - 1 == 200
        k = iters
        for name in ("normr", "rtrans", "pap", "alpha"):
            lo = 1 if name in ("pap", "alpha") else 0          # the reference leaves p.Ap and alpha of "iteration 0" unset
            assert np.array_equal(bits(h[name][lo:k + 1]), bits(g[name][lo:k + 1])), "%s history" % name
        assert crc(x.coefs) == g["crc_x_0"], "solution"
    finally:
        mesh.close()


def test_512_cubed_size_independent_properties(product_lib):
    """BASELINE configs[4] (512^3, 135 M rows, 57 GB on the device).  The reference cannot run at this size in the authoring
    container (65 GB, and 3.6e9 nonzeros overflow its `int` offsets); the CPU oracle built with 64-bit row offsets can, and
    its record (tests/golden/oracle_512x512x512_np1.npz) is compared bit for bit: b after assembly and after the Dirichlet
    conditions, the four per-iteration histories, the solution.  Beside it the size-independent checks: closed-form row /
    nonzero counts, row lengths between 8 and 27, and the order-reproducing dot of the solution with itself equal -- bit for
    bit -- to the literal serial loop on the device (dot mode 1), the reference's summation order
    (ref/src/Vector_functions.hpp:194-226)."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a 100 GB device")
    n = 512
    gbox = mf.Box([0, n, 0, n, 0, n])
    mesh = mf.simple_mesh_description(gbox, gbox)
    try:
        A = mf.CSRMatrix()
        mf.generate_matrix_structure(mesh, A)
        N = (n + 1) ** 3
        assert A.num_rows() == N and A.num_nonzeros() == (3 * (n + 1) - 2) ** 3
        lens = np.diff(A.row_offsets64)
        assert lens.min() == 8 and lens.max() == 27 and int(lens.sum()) == A.num_nonzeros()
        del lens
        b = mf.Vector(0, N)
        x = mf.Vector(0, N)
        g = np.load(ORACLE_512) if os.path.exists(ORACLE_512) else None
        mf.assemble_FE_data(mesh, A, b)
        if g is not None:
            assert (N, A.num_nonzeros()) == tuple(int(v) for v in g["sizes_0"])
            assert crc(b.coefs) == g["crc_b_prebc_0"], "b after assembly"
        mf.impose_dirichlet(0.0, A, b, bc_rows=0)
        mf.impose_dirichlet(1.0, A, b, bc_rows=1)
        if g is not None:
            assert crc(b.coefs) == g["crc_b_0"], "b after the Dirichlet conditions"
        mf.make_local_matrix(A)
        iters, normr, times, h = mf.cg_solve(A, b, x, mf.matvec_std(), 200, EPS, history=True)
        assert iters == 200
        assert len(h["normr"]) == 201 and np.all(np.isfinite(h["normr"])) and 0.0 < h["normr"][200] < h["normr"][0]
        if g is not None:
            # the record of the CPU oracle at this size (64-bit row offsets; tests/golden/make_oracle_golden.py --wide, pinned to
            # the reference's own 400^3 record): all four histories and the solution, bit for bit
            for name in ("normr", "rtrans", "pap", "alpha"):
                lo = 1 if name in ("pap", "alpha") else 0
                assert np.array_equal(bits(h[name][lo:201]), bits(g[name][lo:201])), "%s history" % name
            assert crc(x.coefs) == g["crc_x_0"], "solution"
        d_exact = mf.dot(x, x, ctx=mesh)
        mesh.set_dot_mode(1)
        try:
            d_serial = mf.dot(x, x, ctx=mesh)
        finally:
            mesh.set_dot_mode(0)
        assert np.float64(d_exact).view(np.int64) == np.float64(d_serial).view(np.int64), (d_exact, d_serial)
        assert d_exact > 0.0
    finally:
        mesh.close()
