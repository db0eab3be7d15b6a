"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes wrapper minife_b200), against the
CPU oracle on the same inputs.  Bars (BASELINE.json north_star): CSR structure and index maps bit-exact; assembled
values <= 1e-12 relative (we get bit-exact); CG residual history <= 1e-8 relative at the same iteration count (we
get bit-exact, because matvec / waxpby / dot reproduce the reference's rounding sequence)."""
import numpy as np
import pytest

import minife_b200 as mf
from oracle_lib import Oracle, EPS

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.int64)


def assert_bits(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    same = bits(a) == bits(b)
    if not same.all():
        i = int(np.argmin(same))
        raise AssertionError("%s: %d of %d differ, first at %d: %r vs %r" % (what, (~same).sum(), same.size, i, a.flat[i], b.flat[i]))


class Problem:
    """One rank of an nx*ny*nz problem on cuda:0 through the reference-shaped API."""

    def __init__(self, nx, ny, nz, nranks=1, rank=0, boxes=None):
        self.gbox = mf.Box([0, nx, 0, ny, 0, nz])
        self.boxes = mf.box_partition(0, nranks, 2, self.gbox) if boxes is None else [mf.Box(b) for b in boxes]
        self.mesh = mf.simple_mesh_description(self.gbox, self.boxes[rank], rank=rank, nranks=nranks, all_boxes=self.boxes)
        self.A = mf.CSRMatrix()
        self.b = None

    def structure(self):
        assert mf.generate_matrix_structure(self.mesh, self.A) == 0
        n = self.A.num_rows()
        self.b = mf.Vector(int(self.A.rows[0]), n)
        self.x = mf.Vector(int(self.A.rows[0]), n)
        return self

    def assemble(self):
        mf.assemble_FE_data(self.mesh, self.A, self.b, mf.Parameters())
        return self

    def bc(self):
        mf.impose_dirichlet(0.0, self.A, self.b, bc_rows=0)
        mf.impose_dirichlet(1.0, self.A, self.b, bc_rows=1)
        return self

    def local(self):
        mf.make_local_matrix(self.A)
        return self

    def setup(self):
        return self.structure().assemble().bc().local()


SIZES = [(1, 1, 2), (2, 2, 2), (4, 4, 4), (5, 3, 7), (17, 3, 5), (20, 30, 40), (32, 32, 32), (70, 9, 33)]


@pytest.mark.parametrize("n", SIZES)
def test_phases_single_rank(n, oracle_built, product_lib):
    o = Oracle(*n)
    p = Problem(*n)
    o.generate_matrix_structure(); p.structure()
    rows, ro, cols, coefs = o.csr()
    assert np.array_equal(p.A.rows, rows)
    assert np.array_equal(p.A.row_offsets, ro)
    assert np.array_equal(p.A.row_offsets64, ro.astype(np.int64))
    assert np.array_equal(p.A.packed_cols, cols)
    assert_bits(p.A.packed_coefs, coefs, "zero coefs")
    o.assemble_FE_data(); p.assemble()
    assert_bits(p.A.packed_coefs, o.csr()[3], "A_prebc")
    assert_bits(p.b.coefs, o.b(), "b_prebc")
    o.impose_dirichlet(0); mf.impose_dirichlet(0.0, p.A, p.b, bc_rows=0)
    assert_bits(p.A.packed_coefs, o.csr()[3], "A after bc0")
    assert_bits(p.b.coefs, o.b(), "b after bc0")
    o.impose_dirichlet(1); mf.impose_dirichlet(1.0, p.A, p.b, bc_rows=1)
    assert_bits(p.A.packed_coefs, o.csr()[3], "A")
    assert_bits(p.b.coefs, o.b(), "b")
    o.make_local_matrix(); p.local()
    assert p.A.num_cols == o.num_cols()
    assert np.array_equal(p.A.packed_cols, o.csr()[2])


@pytest.mark.parametrize("n,npr", [((2, 2, 2), 2), ((2, 2, 2), 8), ((8, 8, 8), 8), ((6, 6, 6), 4), ((5, 3, 7), 3), ((12, 10, 9), 5), ((24, 24, 24), 8)])
def test_phases_every_rank_of_a_partition(n, npr, oracle_built, product_lib):
    """Each rank's view is built on cuda:0 in turn (no communication is needed before the solve)."""
    o = Oracle(*n, npr).setup()
    op = Oracle(*n, npr); op.generate_matrix_structure(); op.assemble_FE_data()
    for r in range(npr):
        p = Problem(*n, nranks=npr, rank=r).structure()
        srows, sro, scols, _ = op.csr(r)
        assert np.array_equal(p.A.rows, srows) and np.array_equal(p.A.row_offsets, sro), (r, "structure")
        og = Oracle(*n, npr); og.generate_matrix_structure()
        assert np.array_equal(p.A.packed_cols, og.csr(r)[2]), (r, "global cols")
        p.assemble()
        assert_bits(p.A.packed_coefs, op.csr(r)[3], "rank %d A_prebc" % r)
        assert_bits(p.b.coefs, op.b(r), "rank %d b_prebc" % r)
        p.bc().local()
        rows, ro, cols, coefs = o.csr(r)
        assert_bits(p.A.packed_coefs, coefs, "rank %d A" % r)
        assert_bits(p.b.coefs, o.b(r), "rank %d b" % r)
        assert np.array_equal(p.A.packed_cols, cols), (r, "local cols")
        assert p.A.num_cols == o.num_cols(r)
        for m in ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length"):
            assert np.array_equal(getattr(p.A, m), o.get_map(m, r)), (r, m)
        p.mesh.close()


@pytest.mark.parametrize("n", [(4, 4, 4), (17, 3, 5), (20, 30, 40), (48, 48, 48)])
def test_matvec_bit_exact(n, oracle_built, product_lib):
    o = Oracle(*n).setup()
    p = Problem(*n).setup()
    rng = np.random.default_rng(7)
    for trial in range(3):
        x = rng.standard_normal(o.nrows()) * 10.0 ** rng.integers(-3, 4, o.nrows())
        y = np.full(o.nrows(), 123.0)
        mf.matvec(p.A, x, y)
        assert_bits(y, o.matvec(x), "matvec trial %d" % trial)


def test_matvec_every_configuration_bit_exact(oracle_built, product_lib):
    """All SpMV kernel configurations (warps per CTA, stages, gather batches) give the reference's bits; 40x33x21 has
    x-lines of 41 nodes so tiles straddle lines, faces and the last partial tile."""
    n = (40, 33, 21)
    o = Oracle(*n).setup()
    p = Problem(*n).setup()
    rng = np.random.default_rng(11)
    x = rng.standard_normal(o.nrows()) * 10.0 ** rng.integers(-3, 4, o.nrows())
    want = o.matvec(x)
    for cfg in range(5):
        p.mesh.set_option("matvec_config", cfg)
        y = np.full(o.nrows(), 123.0)
        mf.matvec(p.A, x, y)
        assert_bits(y, want, "matvec config %d" % cfg)


def test_matvec_imported_csr_ragged(product_lib):
    """Hand matrix with ragged rows, empty rows and a row longer than a stencil row (general CSR path)."""
    rng = np.random.default_rng(3)
    nrows, ncols = 1000, 1200
    lens = rng.integers(0, 40, nrows)
    lens[5] = 0; lens[17] = 2000; lens[500] = 1800; lens[999] = 0
    ro = np.zeros(nrows + 1, np.int32); ro[1:] = np.cumsum(lens)
    cols = rng.integers(0, ncols, ro[-1]).astype(np.int32)
    coefs = rng.standard_normal(ro[-1])
    x = rng.standard_normal(ncols)
    p = Problem(10, 10, 9)       # any context; its CSR is replaced
    mf.lib().mfe_import_csr(p.mesh._h, nrows, ncols, ro, cols, coefs)
    y = np.zeros(nrows)
    assert mf.lib().mfe_matvec(p.mesh._h, x, y) == 0, mf.lib().mfe_last_error()
    ref = np.zeros(nrows)
    for r in range(nrows):
        s = 0.0
        for i in range(ro[r], ro[r + 1]):
            s = s + coefs[i] * x[cols[i]]
        ref[r] = s
    assert_bits(y, ref, "ragged matvec")


@pytest.mark.parametrize("n", [1, 2, 3, 255, 4096, 100003])
def test_waxpby_bit_exact(n, oracle_built, product_lib):
    p = Problem(2, 2, 2)
    rng = np.random.default_rng(n)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    for a, b in ((1.0, 0.0), (1.0, -1.0), (1.0, 0.3712), (-2.5e-3, 1.0), (0.1, 0.7)):
        w = np.zeros(n)
        mf.waxpby(a, x, b, y, w, ctx=p.mesh)
        assert_bits(w, Oracle.waxpby(a, x, b, y), "waxpby %g %g" % (a, b))


DOT_CASES = {
    "uniform_sq": lambda r, n: (lambda x: (x, x))(r.random(n)),
    "mostly_positive": lambda r, n: (r.random(n), r.random(n) - 0.02),
    "ones": lambda r, n: (np.ones(n), np.ones(n)),
    "ties": lambda r, n: (lambda x: (x, x))(r.integers(0, 4, n) * 0.5),
    "leading_zeros": lambda r, n: (lambda x: (x, x))(np.concatenate([np.zeros(n // 2), r.random(n - n // 2)])),
    "all_zero": lambda r, n: (np.zeros(n), np.zeros(n)),
    "wide_range": lambda r, n: (lambda x: (x, x))(np.abs(r.standard_normal(n)) * 10.0 ** r.integers(-8, 8, n)),
    "negative_sum": lambda r, n: (-r.random(n), r.random(n)),
    "mixed_sign": lambda r, n: (r.standard_normal(n), r.standard_normal(n)),      # whole-vector replay path
    "with_inf": lambda r, n: (lambda x: (x, x))(np.where(np.arange(n) == n // 3, np.inf, r.random(n))),
}


@pytest.mark.parametrize("case", sorted(DOT_CASES))
@pytest.mark.parametrize("n", [1, 17, 4095, 4097, 70001, 1234567])
def test_dot_reproduces_serial_sum(case, n, oracle_built, product_lib):
    p = Problem(2, 2, 2)
    x, y = DOT_CASES[case](np.random.default_rng(n), n)
    x, y = np.ascontiguousarray(x, np.float64), np.ascontiguousarray(y, np.float64)
    want = Oracle.dot(x, y)
    got = mf.dot(x, y, ctx=p.mesh)
    assert bits(np.array([got]))[0] == bits(np.array([want]))[0] or (np.isnan(got) and np.isnan(want)), (got, want)
    p.mesh.set_dot_mode(mf.DOT_SERIAL)
    got = mf.dot(x, y, ctx=p.mesh)
    assert bits(np.array([got]))[0] == bits(np.array([want]))[0] or (np.isnan(got) and np.isnan(want)), ("serial", got, want)
    p.mesh.set_dot_mode(mf.DOT_TREE)
    got = mf.dot(x, y, ctx=p.mesh)
    if np.isfinite(want):
        assert abs(got - want) <= 1e-9 * np.abs(x * y).sum() + 1e-300, ("tree", got, want)


@pytest.mark.parametrize("n", [(1, 1, 2), (4, 4, 4), (17, 3, 5), (20, 30, 40), (32, 32, 32), (64, 64, 64)])
def test_cg_history_bit_exact(n, oracle_built, product_lib):
    """Free-running CG: same iteration count, bit-identical residual / p.Ap / alpha history and solution
    (tolerance demanded by BASELINE.json: 1e-8 relative)."""
    o = Oracle(*n).setup()
    want = o.cg_solve(200, EPS)
    p = Problem(*n).setup()
    iters, normr, times, h = mf.cg_solve(p.A, p.b, p.x, mf.matvec_std(), 200, EPS, history=True)
    assert iters == want["num_iters"], (iters, want["num_iters"])
    assert_bits(h["normr"], want["normr"], "normr history")
    assert_bits(h["pap"][1:], want["pap"][1:], "p.Ap history")
    assert_bits(h["alpha"][1:], want["alpha"][1:], "alpha history")
    assert_bits([normr], [want["final_normr"]], "final normr")
    assert_bits(p.x.coefs, o.x(), "solution x")
    rel = np.abs(h["normr"] - want["normr"]) / want["normr"].clip(1e-300)
    assert rel.max() <= 1e-8
    assert times[mf.TOTAL] > 0 and times[mf.MATVEC] > 0


def test_cg_32_stops_at_106_like_the_reference(oracle_built, product_lib):
    """tol = eps on the absolute residual stops 32^3 after 106 iterations (1.2517e-16)."""
    p = Problem(32, 32, 32).setup()
    iters, normr, _ = mf.cg_solve(p.A, p.b, p.x, mf.matvec_std(), 200, EPS)
    assert iters == 106
    assert abs(normr - 1.2517e-16) < 1e-20


def test_cg_dot_modes(oracle_built, product_lib):
    """MFE_DOT_SERIAL is bit-identical too; MFE_DOT_TREE is not reproducible but converges."""
    n = (24, 24, 24)
    want = Oracle(*n).setup().cg_solve(200, EPS)
    p = Problem(*n).setup()
    p.mesh.set_dot_mode(mf.DOT_SERIAL)
    iters, normr, _, h = mf.cg_solve(p.A, p.b, p.x, None, 200, EPS, history=True)
    assert iters == want["num_iters"]
    assert_bits(h["normr"], want["normr"], "serial-dot history")
    p.mesh.set_dot_mode(mf.DOT_TREE)
    p.x.coefs[:] = 0
    iters, normr, _ = mf.cg_solve(p.A, p.b, p.x, None, 200, EPS)
    assert abs(iters - want["num_iters"]) <= 3 and normr < 1e-14


def test_cg_max_iter_and_tolerance_semantics(oracle_built, product_lib):
    n = (12, 12, 12)
    for max_iter, tol in ((0, EPS), (1, EPS), (5, EPS), (200, 1e-3), (200, 1e9)):
        o = Oracle(*n).setup()
        want = o.cg_solve(max_iter, tol)
        p = Problem(*n).setup()
        iters, normr, _ = mf.cg_solve(p.A, p.b, p.x, None, max_iter, tol)
        assert iters == want["num_iters"], (max_iter, tol)
        assert_bits([normr], [want["final_normr"]], "normr %s %s" % (max_iter, tol))
        assert_bits(p.x.coefs, o.x(), "x %s %s" % (max_iter, tol))


def test_cg_nonzero_initial_guess(oracle_built, product_lib):
    n = (10, 11, 12)
    o = Oracle(*n).setup()
    x0 = np.random.default_rng(5).random(o.nrows())
    o.L.mfo_set_x(o.h, 0, x0)
    want = o.cg_solve(50, EPS)
    p = Problem(*n).setup()
    p.x.coefs[:] = x0
    iters, normr, _, h = mf.cg_solve(p.A, p.b, p.x, None, 50, EPS, history=True)
    assert iters == want["num_iters"]
    assert_bits(h["normr"], want["normr"], "history")
    assert_bits(p.x.coefs, o.x(), "x")


def test_errors_are_loud(product_lib):
    p = Problem(4, 4, 4).structure()
    with pytest.raises(mf.LibraryError):
        mf.cg_solve(p.A, p.b, p.x)          # make_local_matrix not called (cg_solve.hpp:99-104)
    p.assemble().bc().local()
    with pytest.raises(mf.LibraryError):
        mf.impose_dirichlet(0.0, p.A, p.b, bc_rows=0)    # columns are local now


def test_size_independent_properties_128(product_lib):
    """At a size the oracle does not finish in seconds: structure counts, symmetry-free invariants, and the
    dot-mode cross-check (exact == serial bits) on real CG vectors."""
    n = 128
    p = Problem(n, n, n).setup()
    N = (n + 1) ** 3
    assert p.A.num_rows() == N and p.A.num_nonzeros() == (3 * (n + 1) - 2) ** 3
    ro = p.A.row_offsets64
    lens = np.diff(ro)
    assert lens.min() == 8 and lens.max() == 27 and ro[-1] == p.A.num_nonzeros()
    coefs, cols = p.A.packed_coefs, p.A.packed_cols
    # interior rows sum to ~0 (constant vector in the null space of the diffusion operator before BCs touch them)
    ones = np.ones(N); y = np.zeros(N)
    mf.matvec(p.A, ones, y)
    idx = np.arange(N); ix, iy, iz = idx % (n + 1), (idx // (n + 1)) % (n + 1), idx // (n + 1) ** 2
    interior = (ix > 1) & (ix < n - 1) & (iy > 1) & (iy < n - 1) & (iz > 1) & (iz < n - 1)
    assert np.abs(y[interior]).max() < 1e-12
    bc = (ix == 0) | (iy == 0) | (iz == 0) | (ix == n) | (iy == n) | (iz == n)
    assert np.array_equal(y[bc], np.ones(bc.sum()))          # identity rows
    iters, normr, _, h = mf.cg_solve(p.A, p.b, p.x, None, 200, EPS, history=True)
    assert iters == 200
    # golden values of the reference at 128^3 (printed with 6 digits)
    assert abs(h["normr"][0] - 129.004) < 5e-4 and abs(h["normr"][20] - 0.0529612) < 5e-8
    assert abs(h["normr"][100] - 0.0036414) < 5e-8 and abs(normr - 1.13069e-05) < 5e-11


@pytest.mark.parametrize("n", [128, 256])
def test_cg_history_equals_reference_golden_at_full_size(n, product_lib):
    """BASELINE config 2 (256^3): the 200-iteration history of the UNMODIFIED reference (tests/golden/history_<n>.npz,
    generated by tests/golden/make_golden.py from ref/src) is reproduced bit for bit -- residual norms, p.Ap and alpha
    of every iteration -- and so is the solution (CRC)."""
    import os, zlib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "history_%d.npz" % n))
    p = Problem(n, n, n).setup()
    iters, normr, times, h = mf.cg_solve(p.A, p.b, p.x, mf.matvec_std(), 200, EPS, history=True)
    assert iters == int(g["num_iters"]) == 200
    assert_bits(h["normr"][1:], g["normr"][1:], "normr history vs reference")
    assert_bits(h["pap"][1:], g["pap"][1:], "p.Ap history vs reference")
    assert_bits(h["alpha"][1:], g["alpha"][1:], "alpha history vs reference")
    assert_bits([normr], [float(g["final_normr"])], "final residual")
    assert np.uint32(zlib.crc32(p.x.coefs.tobytes())) == g["crc_x"]
    rel = np.abs(h["normr"][1:] - g["normr"][1:]) / g["normr"][1:]
    assert rel.max() <= 1e-8      # the tolerance BASELINE.json states; we are at 0


@pytest.mark.parametrize("n,whole", [((10, 10, 10), False), ((6, 5, 7), True), ((32, 32, 32), False), ((12, 12, 12), True)])
def test_verify_solution_on_device(n, whole, oracle_built, product_lib):
    """ref/fem/verify_solution.hpp:55-174 evaluated on the device-resident x: same verdict, same worst node, max error
    within 1e-12 relative of the oracle's (CUDA sin/sinh differ from glibc's in the last bits, so the 90601-term
    series is not bit-identical; the solution it is compared with is)."""
    o = Oracle(*n).setup(); o.cg_solve(200, EPS)
    want_rc, want_max, want = o.verify_solution(0.06, whole, info=True)
    p = Problem(*n).setup()
    mf.cg_solve(p.A, p.b, p.x, mf.matvec_std(), 200, EPS)
    rc, gmax, info = mf.verify_solution(p.mesh, None, 0.06, whole, info=True)
    assert rc == want_rc
    assert abs(gmax - want_max) <= 1e-12 * max(want_max, 1e-30) + 1e-15, (gmax, want_max)
    assert info["coords"] == tuple(want[3:6]), (info, want)
    assert_bits([info["computed"]], [want[1]], "computed value at the worst node")
    assert abs(info["analytic"] - want[2]) <= 1e-12
    # tolerance semantics: return code 1 iff max error > tolerance
    assert mf.verify_solution(p.mesh, None, gmax * 0.5, whole) == 1
    assert mf.verify_solution(p.mesh, None, gmax * 2.0, whole) == 0
    # a wrong solution is caught
    bad = p.x.coefs.copy(); bad[:] = 0.25
    assert mf.verify_solution(p.mesh, bad, 0.06, False if n[0] >= 10 else whole) == 1


def test_fused_reporting_mode_and_time_kernels(oracle_built, product_lib):
    """MINIFE_FUSED (ref/src/cg_solve.hpp:172-176): bit-identical history, the SpMV + p.Ap interval lands in cg_times[MATVECDOT];
    time_kernels (ref/src/time_kernels.hpp:97-124): p = x, x = A p, xdotp = sum of max_iter equal dots -- against the oracle."""
    n = (20, 30, 40)
    o = Oracle(*n).setup()
    want = o.cg_solve(200, EPS)
    p = Problem(*n).setup()
    p.mesh.set_option("fused", 1)
    iters, normr, times, h = mf.cg_solve(p.A, p.b, p.x, mf.matvec_std(), 200, EPS, history=True)
    assert iters == want["num_iters"]
    assert_bits(h["normr"], want["normr"], "normr history (fused)")
    assert times[mf.MATVECDOT] > 0 and times[mf.MATVEC] > 0 and times[mf.MATVECDOT] > times[mf.MATVEC]
    p.mesh.set_option("fused", 0)
    x0 = np.cos(0.01 * np.arange(p.A.num_rows())) + 1.5
    p.x.coefs[:] = x0
    xdotp, kt = mf.time_kernels(p.A, p.b, p.x, mf.matvec_std(), 7)
    ax = o.matvec(x0)
    assert_bits(p.x.coefs, ax, "x = A p after time_kernels")
    d, acc = Oracle.dot(ax, x0), 0.0
    for _ in range(7):
        acc += d
    assert_bits([xdotp], [acc], "xdotp")
    assert kt[mf.WAXPY] > 0 and kt[mf.MATVEC] > 0 and kt[mf.DOT] > 0


@pytest.mark.parametrize("n,npr", [((8, 8, 8), 8), ((6, 6, 6), 4), ((5, 3, 7), 3), ((40, 24, 24), 2), ((70, 40, 36), 8), ((33, 80, 17), 4)])
def test_two_pass_matvec_of_every_rank_bit_exact(n, npr, oracle_built, product_lib):
    """The overlapped SpMV of a multi-rank run -- interior pass over the flagged tile table (boundary tiles skipped, warps
    rotated), then the listed boundary tiles -- on every rank's matrix, with the halo supplied by the caller (test hook
    test_no_exchange) instead of the neighbours: y must equal the oracle's matvec of that rank bit for bit, in a single
    pass and in two passes, and so must the CG-loop variant that accumulates the p.Ap prediction sums."""
    o = Oracle(*n, npr).setup()
    rng = np.random.default_rng(3)
    for r in range(npr):
        p = Problem(*n, nranks=npr, rank=r).setup()
        ncols, nrows = p.A.num_cols, p.A.num_rows()
        x = rng.standard_normal(ncols)
        want = o.matvec(x, r)
        p.mesh.set_option("test_no_exchange", 1)
        for overlap in (0, 1, 2):          # 1: the fused one-launch kernel (externals read through its halo pointer), 2: two launches
            p.mesh.set_option("overlap", overlap)
            for cfg in (0, 2, 4):
                p.mesh.set_option("matvec_config", cfg)
                y = np.zeros(nrows)
                mf.matvec(p.A, x.copy(), y)
                assert_bits(y, want, "rank %d overlap %d cfg %d" % (r, overlap, cfg))
        p.mesh.close()


@pytest.mark.parametrize("n,npr", [((8, 8, 8), 8), ((20, 30, 40), 2), ((12, 10, 9), 4), ((5, 3, 7), 3)])
def test_rearranged_matrix_sums_rows_like_matvec_overlap(n, npr, oracle_built, product_lib):
    """mv_overlap_comm_comp=1 (driver.hpp:278-282): rearrange_matrix_local_external + matvec_overlap
    (SparseMatrix_functions.hpp:319-408, 624-675) add a row's external entries AFTER its local ones, which changes the
    rounding of the boundary rows.  mfe_rearrange_matrix_local_external re-sorts the rows on the device; the SpMV of every
    rank (halo supplied by the caller) must then equal the oracle's re-arranged matvec bit for bit -- the oracle's
    re-arrangement itself is pinned against the UNMODIFIED reference's overlap run (tests/golden/mpiov_*.npz,
    test_oracle_vs_ref.py) -- and must differ from the storage-order result somewhere."""
    o = Oracle(*n, npr).setup()
    plain = [o.matvec(np.random.default_rng(5 + r).standard_normal(o.num_cols(r)), r) for r in range(npr)]
    o.rearrange_matrix_local_external()
    differs = False
    for r in range(npr):
        p = Problem(*n, nranks=npr, rank=r).setup()
        x = np.random.default_rng(5 + r).standard_normal(p.A.num_cols)
        mf.rearrange_matrix_local_external(p.A)
        p.mesh.set_option("test_no_exchange", 1)
        want = o.matvec(x, r)
        differs = differs or not np.array_equal(bits(want), bits(plain[r]))
        for overlap in (0, 1, 2):
            p.mesh.set_option("overlap", overlap)
            y = np.zeros(p.A.num_rows())
            mf.matvec_overlap()(p.A, x.copy(), y)
            assert_bits(y, want, "rank %d overlap %d" % (r, overlap))
        rows, ro, cols, coefs = o.csr(r)
        assert np.array_equal(p.A.packed_cols, cols)
        assert_bits(p.A.packed_coefs, coefs, "re-sorted coefficients of rank %d" % r)
        p.mesh.close()
    assert differs, "the overlap order should change some row sums"


import glob as _glob
import os as _os
_IMB = sorted(_glob.glob(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "mpi_*_imb*.npz")))


@pytest.mark.parametrize("path", _IMB)
def test_boxes_changed_by_add_imbalance(path, oracle_built, product_lib):
    """the reference's add_imbalance (ref/utils/imbalance.hpp:167-293, driver.hpp:137-139) shifts whole box
    planes, so the ranks no longer form an RCB partition.  The golden files hold what the UNMODIFIED reference computes on
    such boxes (through the MPI shim); every rank's structure, index maps, values and two-pass SpMV must match."""
    g = np.load(path)
    n, npr = tuple(int(v) for v in g["n"]), int(g["np"])
    boxes = np.stack([g["box_%d" % r] for r in range(npr)])
    o = Oracle(*n, npr, boxes=boxes).setup()
    rng = np.random.default_rng(7)
    for r in range(npr):
        p = Problem(*n, nranks=npr, rank=r, boxes=boxes).structure()
        assert np.array_equal(p.A.rows, g["rows_%d" % r]) and np.array_equal(p.A.row_offsets, g["row_offsets_%d" % r])
        assert np.array_equal(p.A.packed_cols, g["cols_global_%d" % r])
        p.assemble().bc().local()
        assert np.array_equal(p.A.packed_cols, g["cols_local_%d" % r]) and p.A.num_cols == int(g["num_cols_%d" % r][0])
        for m in ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length"):
            assert np.array_equal(getattr(p.A, m), g["%s_%d" % (m, r)]), (r, m)
        assert_bits(p.A.packed_coefs, o.csr(r)[3], "rank %d A" % r)
        assert_bits(p.b.coefs, o.b(r), "rank %d b" % r)
        x = rng.standard_normal(p.A.num_cols)
        p.mesh.set_option("test_no_exchange", 1)
        for overlap in (0, 1, 2):
            p.mesh.set_option("overlap", overlap)
            y = np.zeros(p.A.num_rows())
            mf.matvec(p.A, x.copy(), y)
            assert_bits(y, o.matvec(x, r), "rank %d matvec overlap %d" % (r, overlap))
        p.mesh.close()


@pytest.mark.parametrize("n,npr", [((5, 3, 7), 1), ((20, 30, 40), 1), ((33, 17, 40), 1), ((40, 24, 24), 2), ((33, 50, 17), 4), ((18, 18, 18), 8)])
def test_fused_assembly_kernel_bit_exact(n, npr, oracle_built, product_lib):
    """assembly_mode 1: k_assemble_fused keeps the element matrices in shared memory (two-plane ring, per-axis Jacobian
    tables) instead of the scratch slab in global memory; A and b before and after the Dirichlet conditions must be the
    oracle's bits on every rank (ref/src/perform_element_loop.hpp:40-104, SparseMatrix_functions.hpp:183-255)."""
    o_pre = Oracle(*n, npr).setup(upto="assembly")
    o = Oracle(*n, npr).setup()
    for r in range(npr):
        p = Problem(*n, nranks=npr, rank=r).structure()
        p.mesh.set_option("assembly_mode", 1)
        p.assemble()
        assert_bits(p.A.packed_coefs, o_pre.csr(r)[3], "rank %d A before BC" % r)
        assert_bits(p.b.coefs, o_pre.b(r), "rank %d b before BC" % r)
        p.bc()
        assert_bits(p.A.packed_coefs, o.csr(r)[3], "rank %d A" % r)          # make_local_matrix changes columns only
        assert_bits(p.b.coefs, o.b(r), "rank %d b" % r)
        p.mesh.close()


def test_matvec_forms_beta_times_the_stale_y(oracle_built, product_lib):
    """ref/src/SparseMatrix_functions.hpp:565-568: every row starts from beta*y[row] with beta = 0.  Value-neutral for a finite
    positive y, but a negative or -0.0 stale y makes the seed -0.0 (which survives a row whose products are all -0.0) and a
    non-finite one makes the row NaN.  Mesh matrix against the oracle, and a hand matrix against the literal loop."""
    n = (6, 5, 4)
    o = Oracle(*n).setup()
    p = Problem(*n).setup()
    nrows = p.A.num_rows()
    rng = np.random.default_rng(11)
    x = rng.standard_normal(nrows)
    x[rng.random(nrows) < 0.5] = -0.0                      # many rows with only zero products
    y0 = rng.standard_normal(nrows)
    y0[::7] = -0.0; y0[3::11] = np.nan; y0[5::13] = np.inf; y0[6::17] = -np.inf
    for xx in (x, np.full(nrows, -0.0), np.zeros(nrows)):
        want = o.matvec(xx, 0, y0=y0)
        y = y0.copy()
        mf.matvec(p.A, xx.copy(), y)
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(y), nan)
        assert_bits(y[~nan], want[~nan], "matvec with stale y")
    p.mesh.close()
    # all-positive coefficients times x = -0.0: every product is -0.0, so the sign of the result is the sign of 0*y[row]
    nr = 40
    lens = np.full(nr, 3); lens[[5, 6, 17]] = 0          # and rows without entries: the result IS 0*y[row]
    ro = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    nz = int(ro[-1])
    cols = ((np.arange(nz) * 7) % nr).astype(np.int32)
    coefs = np.abs(rng.standard_normal(nz)) + 0.5
    h = Problem(2, 2, 2).structure()
    assert mf.lib().mfe_import_csr(h.mesh._h, nr, nr, ro, cols, coefs) == 0
    xz = np.full(nr, -0.0)
    y0 = np.where(np.arange(nr) % 3 == 0, -2.5, np.where(np.arange(nr) % 3 == 1, -0.0, 4.0))
    want = np.empty(nr)
    for r in range(nr):
        s_ = np.float64(0.0) * np.float64(y0[r])
        for i in range(ro[r], ro[r + 1]):
            s_ = s_ + np.float64(coefs[i]) * np.float64(xz[cols[i]])
        want[r] = s_
    assert np.signbit(want).any() and not np.signbit(want).all()
    y = y0.copy()
    assert mf.lib().mfe_matvec(h.mesh._h, xz.copy(), y) == 0
    assert_bits(y, want, "hand matrix, -0.0 products")
    h.mesh.close()


def test_pll_matvec_known_answers(product_lib):
    """The hand matrices of ref/basic/utest_cases.hpp:518-617 (pll_matvec2: 4x4 over 2 ranks) and 619-735 (pll_matvec3: 6x6
    over 3 ranks) with their exact answers -- API-independent data.  Every rank's rows are imported in
    local-index form (owned columns first, externals appended in first-encounter order, make_local_matrix.hpp:87-180) and
    multiplied with the vector the halo exchange would have produced."""
    cases = [
        (np.array([[1, 1, 0, 0], [2, 1, -1, 1], [0, -2, 1, 0], [0, 2, 0, 1]], float), [1, 2, 3, 4], [3, 5, -1, 8], 2),
        (np.array([[1, 0, 0, -1, 0, 0], [0, 1, 0, 0, 0, -1], [2, 0, 1, 0, -1, 0], [0, 0, 0, 1, 0, 0], [0, 2, 0, 0, 1, 0], [0, 0, 0, 2, 0, 1]], float),
         [1, 2, 3, 4, 5, 6], [-3, -4, 0, 4, 9, 14], 3),
    ]
    for A, x, want, nranks in cases:
        x, want = np.array(x, float), np.array(want, float)
        per = len(x) // nranks
        for r in range(nranks):
            lo, hi = r * per, (r + 1) * per
            ext = []                                    # global ids of the externals, first-encounter order
            ro, cols, coefs = [0], [], []
            for i in range(lo, hi):
                for j in np.nonzero(A[i])[0]:
                    if lo <= j < hi:
                        cols.append(j - lo)
                    else:
                        if j not in ext:
                            ext.append(int(j))
                        cols.append(per + ext.index(int(j)))
                    coefs.append(A[i, j])
                ro.append(len(cols))
            xl = np.concatenate([x[lo:hi], x[ext]]) if ext else x[lo:hi].copy()
            p = Problem(4, 4, 4)
            mf.lib().mfe_import_csr(p.mesh._h, per, len(xl), np.array(ro, np.int32), np.array(cols, np.int32), np.array(coefs, float))
            y = np.zeros(per)
            assert mf.lib().mfe_matvec(p.mesh._h, np.ascontiguousarray(xl), y) == 0, mf.lib().mfe_last_error()
            assert_bits(y, want[lo:hi], "pll_matvec%d rank %d" % (nranks, r))
            p.mesh.close()
