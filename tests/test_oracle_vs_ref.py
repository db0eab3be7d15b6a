"""Pins the CPU oracle (oracle/minife_oracle.c) -- not gpu.

1. against the golden fixtures under tests/golden/ (generated from the UNMODIFIED reference by
   tests/golden/make_golden.py; they travel with the repo), and
2. where oracle/_ref exists (the reference itself compiled by oracle/Makefile), directly against it, serial and
   multi-rank through the MPI shim.
Everything is compared bit for bit."""
import glob
import os
import zlib

import numpy as np
import pytest

from oracle_lib import Oracle, Ref, ref_dump_mpi, have_ref, have_ref_dump, EPS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def beq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "serial_*.npz"))))
def test_serial_golden(path, oracle_built):
    g = np.load(path)
    n = tuple(int(v) for v in g["n"])
    o = Oracle(*n)
    o.generate_matrix_structure()
    rows, ro, cols, _ = o.csr()
    assert o.nrows() == int(g["nrows"]) and o.nnz() == int(g["nnz"])
    assert crc(ro) == g["crc_row_offsets"] and crc(cols) == g["crc_cols"]
    o.assemble_FE_data()
    assert crc(o.csr()[3]) == g["crc_coefs_prebc"] and crc(o.b()) == g["crc_b_prebc"]
    if "coefs_prebc" in g:
        assert beq(o.csr()[3], g["coefs_prebc"]) and beq(o.b(), g["b_prebc"]) and beq(ro, g["row_offsets"]) and beq(cols, g["cols"])
    o.impose_dirichlet(0); o.impose_dirichlet(1)
    assert crc(o.csr()[3]) == g["crc_coefs"] and crc(o.b()) == g["crc_b"]
    o.make_local_matrix()
    h = o.cg_solve(200, EPS)
    assert h["num_iters"] == int(g["num_iters"])
    assert beq(h["normr"][1:], g["normr"][1:]) and beq(h["pap"][1:], g["pap"][1:]) and beq(h["alpha"][1:], g["alpha"][1:])
    assert crc(o.x()) == g["crc_x"]


@pytest.mark.parametrize("n", [32, 64])
def test_history_golden(n, oracle_built):
    g = np.load(os.path.join(GOLD, "history_%d.npz" % n))
    h = Oracle(n, n, n).setup().cg_solve(200, EPS)
    assert h["num_iters"] == int(g["num_iters"])
    assert beq(h["normr"][1:], g["normr"][1:]) and beq(h["pap"][1:], g["pap"][1:])
    assert h["final_normr"] == float(g["final_normr"])


def test_reference_known_answers():
    """values the reference prints (6 significant digits)."""
    g32, g64 = np.load(os.path.join(GOLD, "history_32.npz")), np.load(os.path.join(GOLD, "history_64.npz"))
    assert int(g32["num_iters"]) == 106 and abs(float(g32["final_normr"]) - 1.2517e-16) < 5e-21
    assert abs(g32["normr"][0] - 33.0132) < 5e-5 and abs(g32["normr"][20] - 0.0213628) < 5e-8
    assert int(g64["num_iters"]) == 200 and abs(g64["normr"][0] - 65.0072) < 5e-5
    assert abs(g64["normr"][20] - 0.0409638) < 5e-8 and abs(g64["normr"][60] - 0.00330043) < 5e-9 and abs(g64["normr"][100] - 8.66762e-06) < 5e-12
    assert abs(float(g64["final_normr"]) - 4.15881e-15) < 5e-21
    p = os.path.join(GOLD, "history_256.npz")
    if os.path.exists(p):
        g = np.load(p)
        assert int(g["num_iters"]) == 200 and abs(g["normr"][0] - 257.002) < 5e-4 and abs(float(g["final_normr"]) - 0.0019344) < 5e-8


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "mpi_*.npz"))))
def test_multirank_golden(path, oracle_built):
    g = np.load(path)
    n, npr = tuple(int(v) for v in g["n"]), int(g["np"])
    # *_imb*.npz: the boxes are the ones the reference's add_imbalance left on the ranks (not an RCB partition)
    o = Oracle(*n, npr, boxes=np.stack([g["box_%d" % r] for r in range(npr)]) if "imbalance" in g else None)
    o.generate_matrix_structure()
    for r in range(npr):
        rows, ro, cols, _ = o.csr(r)
        assert beq(o.box(r), g["box_%d" % r]) and beq(rows, g["rows_%d" % r]) and beq(ro, g["row_offsets_%d" % r]) and beq(cols, g["cols_global_%d" % r])
    o.assemble_FE_data(); o.impose_dirichlet(0); o.impose_dirichlet(1)
    for r in range(npr):
        assert crc(o.csr(r)[3]) == g["crc_coefs_%d" % r] and crc(o.b(r)) == g["crc_b_%d" % r]
    o.make_local_matrix()
    for r in range(npr):
        assert beq(o.csr(r)[2], g["cols_local_%d" % r]) and o.num_cols(r) == int(g["num_cols_%d" % r][0])
        for m in ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length"):
            assert beq(o.get_map(m, r), g["%s_%d" % (m, r)]), (r, m)
    h = o.cg_solve(200, EPS)
    assert beq(h["normr"], g["normr"])
    for r in range(npr):
        assert crc(o.x(r)) == g["crc_x_%d" % r]


def test_survey_appendix_b_values(oracle_built):
    """Hand-checkable facts from SURVEY.md Appendix B (reference run through the MPI shim)."""
    o = Oracle(2, 2, 2, 2).setup()
    assert o.get_map("external_index", 0).tolist() == [9, 10, 12, 13, 11, 14, 15, 16, 17]
    assert o.get_map("elements_to_send", 0).tolist() == [0, 1, 3, 4, 2, 5, 6, 7, 8]
    o = Oracle(8, 8, 8, 8).setup()
    assert o.get_map("neighbors", 4).tolist() == [0, 1, 5, 2, 6, 3, 7]
    assert o.get_map("recv_length", 0).tolist() == [16, 16, 4, 16, 4, 4, 1]
    assert [o.nrows(r) for r in range(8)] == [64, 80, 80, 100, 80, 100, 100, 125]
    o = Oracle(5, 3, 7, 3).setup()
    assert [o.nrows(r) for r in range(3)] == [72, 60, 60] and [o.num_cols(r) for r in range(3)] == [96, 108, 108]


def test_structure_counts(oracle_built):
    """ref/basic/utest_cases.hpp:151,183,187 (API independent): 1 element -> 8 rows / 64 nnz; 2^3 -> 27 rows / 343 nnz,
    centre row has 27 entries; nnz = (3(n+1)-2)^3."""
    o = Oracle(1, 1, 1); o.generate_matrix_structure()
    assert o.nrows() == 8 and o.nnz() == 64
    o = Oracle(2, 2, 2); o.generate_matrix_structure()
    ro = o.csr()[1]
    assert o.nrows() == 27 and o.nnz() == 343 and ro[14] - ro[13] == 27
    o = Oracle(32, 32, 32); o.generate_matrix_structure()
    assert o.nnz() == 912673 == (3 * 33 - 2) ** 3


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libminife_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("n", [(1, 1, 2), (4, 4, 4), (5, 3, 7), (17, 3, 5), (12, 12, 12)])
def test_oracle_equals_reference_serial(n, oracle_built):
    assert beq(Oracle.gradient_values(), Ref.gradient_values())
    o, r = Oracle(*n), Ref(*n)
    for e in range(min(n[0] * n[1] * n[2], 40)):
        a, b = Oracle.element(*n, e), r.element(e)
        assert beq(a[1], b[1]) and beq(a[2], b[2]) and beq(a[3], b[3])
    o.generate_matrix_structure(); r.generate_matrix_structure()
    assert all(beq(a, b) for a, b in zip(o.csr(), r.csr()))
    assert beq(o.bc_rows(0), r.bc_rows(0)) and beq(o.bc_rows(1), r.bc_rows(1))
    o.assemble_FE_data(); r.assemble_FE_data()
    assert all(beq(a, b) for a, b in zip(o.csr(), r.csr())) and beq(o.b(), r.b())
    for w in (0, 1):
        o.impose_dirichlet(w); r.impose_dirichlet(w)
        assert all(beq(a, b) for a, b in zip(o.csr(), r.csr())) and beq(o.b(), r.b())
    o.make_local_matrix(); r.make_local_matrix()
    rng = np.random.default_rng(1)
    x = rng.standard_normal(o.nrows())
    assert beq(o.matvec(x), r.matvec(x))
    assert beq(Oracle.waxpby(0.3, x, -1.7, x[::-1].copy()), Ref.waxpby(0.3, x, -1.7, x[::-1].copy()))
    assert Oracle.dot(x, x[::-1].copy()) == Ref.dot(x, x[::-1].copy())
    ho, hr = o.cg_solve(), r.cg_solve()
    assert ho["num_iters"] == hr["num_iters"]
    for k in ("normr", "rtrans", "pap", "alpha"):
        assert beq(ho[k][1:], hr[k][1:]), k
    assert beq(o.x(), r.x())
    # the reference's own cg_solve (not the traced restatement in ref_entry.cpp) ends in the same state
    r2 = Ref(*n).setup()
    h2 = r2.cg_solve(trace=False)
    assert h2["num_iters"] == ho["num_iters"] and h2["final_normr"] == ho["final_normr"] and beq(r2.x(), o.x())


@pytest.mark.skipif(not have_ref_dump(), reason="oracle/_ref/ref_dump_mpi.x not built (needs /root/reference)")
@pytest.mark.parametrize("n,npr", [((4, 4, 4), 2), ((12, 10, 9), 5), ((9, 9, 9), 8)])
def test_oracle_equals_reference_multirank(n, npr, oracle_built):
    d = ref_dump_mpi(*n, npr)
    o = Oracle(*n, npr).setup()
    for r in range(npr):
        rows, ro, cols, coefs = o.csr(r)
        assert beq(rows, d["rows"][r]) and beq(ro, d["row_offsets"][r]) and beq(cols, d["cols_local"][r]) and beq(coefs, d["coefs"][r])
        assert beq(o.b(r), d["b"][r])
        for m in ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length"):
            assert beq(o.get_map(m, r), d[m][r]), (r, m)
    h = o.cg_solve()
    assert beq(h["normr"], d["normr_hist"][0])
    for r in range(npr):
        assert beq(o.x(r), d["x"][r])


@pytest.mark.parametrize("n,npr", [((8, 8, 8), 8), ((20, 30, 40), 2), ((12, 10, 9), 4), ((5, 3, 7), 3)])
def test_oracle_rearranged_matvec_equals_reference_overlap_golden(n, npr, oracle_built):
    """mv_overlap_comm_comp=1: the golden files hold what the UNMODIFIED reference computes with
    rearrange_matrix_local_external + matvec_overlap (ref_dump_mpi.x, REF_DUMP_OVERLAP=1); the oracle re-sorts its rows
    instead (mfo_rearrange_matrix_local_external) and must reproduce history and solution bit for bit."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpiov_%dx%dx%d_np%d.npz" % (n + (npr,))))
    o = Oracle(*n, npr).setup()
    o.rearrange_matrix_local_external()
    h = o.cg_solve()
    assert beq(h["normr"], g["normr"])
    for r in range(npr):
        assert beq(o.x(r), g["x_%d" % r])


@pytest.mark.skipif(not have_ref_dump(), reason="oracle/_ref/ref_dump_mpi.x not built (needs /root/reference)")
def test_oracle_rearranged_matvec_equals_reference_overlap_live(oracle_built):
    n, npr = (7, 6, 9), 4
    d = ref_dump_mpi(*n, npr, overlap=True)
    o = Oracle(*n, npr).setup()
    o.rearrange_matrix_local_external()
    h = o.cg_solve()
    assert beq(h["normr"], d["normr_hist"][0])
    for r in range(npr):
        assert beq(o.x(r), d["x"][r])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("n,whole", [((10, 10, 10), False), ((6, 5, 7), True), ((20, 20, 20), False)])
def test_verify_solution_vs_reference(n, whole, oracle_built):
    """ref/fem/verify_solution.hpp only reports a return code; bisecting the tolerance against the UNMODIFIED
    reference pins the oracle's max error (and with it analytic_soln) to the last bit."""
    o = Oracle(*n).setup(); o.cg_solve(200, EPS)
    r = Ref(*n).setup(); r.cg_solve(200, EPS, trace=False)
    assert beq(o.x(), r.x())
    rc, gmax, info = o.verify_solution(0.06, whole, info=True)
    assert rc == r.verify_solution(0.06, whole)          # coarse whole-domain runs legitimately fail the 0.06 test
    assert gmax > 0 and info[0] == gmax
    # the reference returns 1 exactly when max_err > tolerance: flips between nextafter(gmax, 0) and gmax
    assert r.verify_solution(float(np.nextafter(gmax, 0.0)), whole) == 1
    assert r.verify_solution(gmax, whole) == 0
    assert o.verify_solution(float(np.nextafter(gmax, 0.0)), whole) == 1 and o.verify_solution(gmax, whole) == 0


def test_verify_solution_multirank_equals_serial(oracle_built):
    """the verdict does not depend on the partition (MPI_Allreduce MAX, verify_solution.hpp:148-154) when x is the same"""
    o1 = Oracle(8, 8, 8).setup(); o1.cg_solve(200, EPS)
    _, g1, i1 = o1.verify_solution(0.06, True, info=True)
    o4 = Oracle(8, 8, 8, 4).setup()
    x = o1.x()
    for rk in range(4):
        bx = np.zeros(6, np.int32); o4.L.mfo_get_box(o4.h, rk, bx)
        hi = [int(bx[2 * a + 1]) + (1 if bx[2 * a + 1] == 8 else 0) for a in range(3)]      # owned nodes: +1 on the global high side
        ids = [ix + 9 * (iy + 9 * iz) for iz in range(bx[4], hi[2]) for iy in range(bx[2], hi[1]) for ix in range(bx[0], hi[0])]
        o4.L.mfo_set_x(o4.h, rk, np.ascontiguousarray(x[ids]))
    _, g4, i4 = o4.verify_solution(0.06, True, info=True)
    assert g1 == g4 and np.array_equal(i1, i4)
