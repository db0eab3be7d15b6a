"""Host-side / shared logic of the CUDA path, run on the CPU (not gpu).

tests/cpu_harness.cpp compiles the MFE_HD headers the kernels are made of (mesh.h, hex8.h, assembly.h,
exact_dot.h, halo_maps.h) and executes them thread by thread; this checks them against the oracle bit for bit:
analytic structure, element arithmetic order, row-owner gather, Dirichlet rows, halo index maps without messages,
and the order-reproducing dot."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_lib import Oracle

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
MAPS = ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length")


def P(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def H(harness):
    h = harness
    h.hs_create.restype = C.c_void_p; h.hs_create.argtypes = [C.c_int] * 5
    for f in ("hs_destroy", "hs_structure", "hs_make_local"):
        getattr(h, f).argtypes = [C.c_void_p]
    h.hs_assemble.argtypes = [C.c_void_p, C.c_int]; h.hs_assemble_ring.argtypes = [C.c_void_p, C.c_int, C.c_int]; h.hs_dirichlet.argtypes = [C.c_void_p, C.c_int, C.c_double]
    for f in ("hs_nrows", "hs_nnz", "hs_num_cols"):
        getattr(h, f).restype = C.c_longlong; getattr(h, f).argtypes = [C.c_void_p]
    h.hs_get_csr.argtypes = [C.c_void_p] * 6
    h.hs_get_map.restype = C.c_longlong; h.hs_get_map.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    h.hs_element.argtypes = [C.c_int] * 6 + [f64p, f64p]
    h.hs_element_tables.argtypes = [C.c_int] * 6 + [f64p, f64p]
    h.hs_tables.argtypes = [f64p, f64p]
    h.hs_exact_dot.restype = C.c_double; h.hs_exact_dot.argtypes = [C.c_longlong, f64p, f64p, C.c_void_p]
    h.hs_serial_dot.restype = C.c_double; h.hs_serial_dot.argtypes = [C.c_longlong, f64p, f64p]
    return h


def get(H, h):
    n, nnz = H.hs_nrows(h), H.hs_nnz(h)
    rows, ro, cols, co, b = np.zeros(n, np.int32), np.zeros(n + 1, np.int64), np.zeros(nnz, np.int32), np.zeros(nnz), np.zeros(n)
    H.hs_get_csr(h, P(rows), P(ro), P(cols), P(co), P(b))
    return rows, ro, cols, co, b


def beq(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.int64), b.view(np.int64))


def test_element_tables_and_matrices(H, oracle_built):
    g, p = np.zeros(192), np.zeros(64)
    H.hs_tables(g, p)
    assert beq(g, Oracle.gradient_values())
    for n in ((1, 1, 1), (4, 5, 6), (256, 256, 256), (400, 400, 400)):
        tot = n[0] * n[1] * n[2]
        for e in sorted({0, 1 % tot, (tot // 2 + 3) % tot, tot - 1}):
            ex, ey, ez = e % n[0], (e // n[0]) % n[1], e // (n[0] * n[1])
            K, f = np.zeros(36), np.zeros(8)
            H.hs_element(*n, ex, ey, ez, K, f)
            _, _, Ko, fo = Oracle.element(*n, e)
            assert beq(K, Ko) and beq(f, fo), (n, e)
            # the fused assembly kernel's formulation: Jacobian columns from per-axis tables (hex8_jacobian_column)
            Kt, ft = np.zeros(36), np.zeros(8)
            H.hs_element_tables(*n, ex, ey, ez, Kt, ft)
            assert beq(Kt, Ko) and beq(ft, fo), ("tables", n, e)
    # unit cube and the h = 1/256 origin element of the reference
    K, f = np.zeros(36), np.zeros(8)
    H.hs_element(1, 1, 1, 0, 0, 0, K, f)
    near = lambda a, b: abs(a - b) <= 2e-15 * abs(b)      # BASELINE.md prints %.16g
    assert near(K[0], 0.3333333332238529) and near(K[1], 5.47402399016933e-11) and near(K[2], -0.08333333333333337) and near(f[0], 0.12500000000000011)
    H.hs_element(256, 256, 256, 0, 0, 0, K, f)
    assert near(K[0], 0.0013020833329056756) and near(K[1], 2.1382906211598945e-13) and near(K[6], -0.00032552083354716254) and near(f[0], 7.4505805969238347e-09)


CASES = [((4, 4, 4), 1, 100), ((5, 3, 7), 1, 2), ((2, 2, 2), 8, 1), ((8, 8, 8), 8, 3), ((6, 6, 6), 4, 100), ((5, 3, 7), 3, 1),
         ((12, 10, 9), 5, 4), ((16, 16, 16), 2, 5), ((3, 9, 4), 7, 2), ((20, 6, 11), 6, 3), ((1, 1, 2), 1, 1), ((1, 1, 2), 2, 1),
         ((17, 3, 5), 2, 2), ((20, 30, 40), 2, 7)]      # the reference CI's own np = 2 cases


IMBALANCED = sorted(__import__("glob").glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpi_*_imb*.npz")))


@pytest.mark.parametrize("n,npr,planes", CASES + [(p, 0, 3) for p in IMBALANCED])
def test_structure_assembly_dirichlet_maps(H, oracle_built, n, npr, planes):
    boxes = None
    if isinstance(n, str):       # boxes left by the reference's add_imbalance (golden file): not an RCB partition
        g = np.load(n)
        n, npr = tuple(int(v) for v in g["n"]), int(g["np"])
        boxes = np.ascontiguousarray(np.stack([g["box_%d" % r] for r in range(npr)]), np.int32)
        H.hs_create_boxes.restype = C.c_void_p
        H.hs_create_boxes.argtypes = [C.c_int] * 5 + [C.c_void_p]
    o = Oracle(*n, npr, boxes=boxes); o.generate_matrix_structure()
    hs = [H.hs_create(*n, npr, r) if boxes is None else H.hs_create_boxes(*n, npr, r, boxes.ctypes.data_as(C.c_void_p)) for r in range(npr)]
    try:
        for r, h in enumerate(hs):
            H.hs_structure(h); rows, ro, cols, co, b = get(H, h); R = o.csr(r)
            assert np.array_equal(rows, R[0]) and np.array_equal(ro, R[1]) and np.array_equal(cols, R[2]), (r, "structure")
        o.assemble_FE_data()
        for r, h in enumerate(hs):
            # the fused kernel's formulation first (two-plane ring, Jacobian tables; small and odd column sizes so that the
            # seams between columns and the clipped windows are exercised), then the slab + gather form overwrites it
            H.hs_assemble_ring(h, 3 + planes % 3, 2 + planes % 4); rows, ro, cols, co, b = get(H, h)
            assert beq(co, o.csr(r)[3]) and beq(b, o.b(r)), (r, "fused-ring assembly")
            H.hs_assemble(h, planes); rows, ro, cols, co, b = get(H, h)
            assert beq(co, o.csr(r)[3]) and beq(b, o.b(r)), (r, "assembly")
        for w, val in ((0, 0.0), (1, 1.0)):
            o.impose_dirichlet(w)
            for r, h in enumerate(hs):
                H.hs_dirichlet(h, w, val); rows, ro, cols, co, b = get(H, h)
                assert beq(co, o.csr(r)[3]) and beq(b, o.b(r)), (r, "dirichlet", w)
        o.make_local_matrix()
        for r, h in enumerate(hs):
            H.hs_make_local(h); rows, ro, cols, co, b = get(H, h)
            assert np.array_equal(cols, o.csr(r)[2]) and H.hs_num_cols(h) == o.num_cols(r), (r, "local cols")
            for m in MAPS:
                cnt = H.hs_get_map(h, m.encode(), None); a = np.zeros(cnt, np.int32)
                if cnt:
                    H.hs_get_map(h, m.encode(), P(a))
                assert np.array_equal(a, o.get_map(m, r)), (r, m)
    finally:
        for h in hs:
            H.hs_destroy(h)


def _big_goldens():
    import glob, os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(os.path.basename(f) for f in glob.glob(os.path.join(d, "big_*.npz")) + glob.glob(os.path.join(d, "oracle_*.npz")) if "_np1." not in f)


@pytest.mark.parametrize("name", _big_goldens())
def test_index_maps_at_full_size_equal_the_reference_record(name, H):
    """make_local_matrix's index maps (halo_maps.h: enumerated in encounter order, no sort) at the sizes of BASELINE's
    multi-rank configurations, on every rank, against the record of the UNMODIFIED reference at that np
    (tests/golden/make_big_golden.py: lengths and CRC-32 of external_index / external_local_index / elements_to_send,
    neighbors, recv_length, send_length, num_cols); at 512^3, which the reference cannot hold here, against the record of
    the oracle (oracle_*.npz, tests/golden/make_oracle_golden.py).  Host logic only: no matrix is built."""
    import os, zlib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    n, npr = tuple(int(v) for v in g["n"]), int(g["np"])
    crc = lambda a: np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))
    for r in range(npr):
        h = H.hs_create(n[0], n[1], n[2], npr, r)
        try:
            assert H.hs_nrows(h) == int(g["sizes_%d" % r][0])
            H.hs_make_local(h)
            m = {}
            for k in MAPS + ("sorted_global", "sorted_local"):
                cnt = H.hs_get_map(h, k.encode(), None); a = np.zeros(cnt, np.int32)
                if cnt:
                    H.hs_get_map(h, k.encode(), P(a))
                m[k] = a
            assert H.hs_num_cols(h) == int(g["num_cols_%d" % r]), (r, "num_cols")
            for k in ("neighbors", "recv_length", "send_length"):
                assert np.array_equal(m[k], g["%s_%d" % (k, r)]), (r, k)
            for k in ("external_index", "external_local_index", "elements_to_send"):
                assert len(m[k]) == int(g["len_%s_%d" % (k, r)]) and crc(m[k]) == g["crc_%s_%d" % (k, r)], (r, k)
            # the column lookup table: the same pairs by ascending global row
            order = np.argsort(m["external_index"], kind="stable")
            assert np.array_equal(m["sorted_global"], m["external_index"][order]), (r, "sorted_global")
            assert np.array_equal(m["sorted_local"], m["external_local_index"][order]), (r, "sorted_local")
        finally:
            H.hs_destroy(h)


def _dot_cases():
    rng = np.random.default_rng(11)
    for n in (0, 1, 15, 16, 17, 255, 4095, 4096, 4097, 100000, 1234567):
        x = rng.random(n); yield "uniform_%d" % n, x, x
    for n in (1000, 65536, 300007):
        yield "normal_mixed_%d" % n, rng.standard_normal(n), rng.standard_normal(n)
        x = np.abs(rng.standard_normal(n)) * 10.0 ** rng.integers(-8, 8, n); yield "wide_%d" % n, x, x
        yield "ones_%d" % n, np.ones(n), np.ones(n)
        x = np.arange(n, dtype=np.float64); yield "integers_%d" % n, x, x
        yield "tenth_%d" % n, np.full(n, 0.1), np.full(n, 0.1)
        x = rng.integers(0, 4, n).astype(np.float64) * 0.5; yield "ties_%d" % n, x, x
        x = np.zeros(n); x[n // 2:] = rng.random(n - n // 2); yield "leading_zeros_%d" % n, x, x
        yield "zeros_%d" % n, np.zeros(n), np.zeros(n)
        yield "mostly_positive_%d" % n, rng.random(n), rng.random(n) - 0.02
        x = rng.random(n); x[n // 3] = 1e200; yield "outlier_%d" % n, x, x
        x = rng.random(n); x[n // 3] = np.inf; yield "inf_%d" % n, x, x
        x = rng.random(n) * 1e-160; yield "subnormal_%d" % n, x, x
        yield "negative_%d" % n, -rng.random(n), rng.random(n)
    n = 8192 * 16 * 3 + 5
    yield "supers", rng.random(n) + 0.01, rng.random(n)
    n = 262144 * 16 + 12345
    yield "hypers", rng.random(n) + 0.01, rng.random(n)


def test_exact_dot_reproduces_the_serial_loop(H):
    for name, x, y in _dot_cases():
        x, y = np.ascontiguousarray(x, np.float64), np.ascontiguousarray(y, np.float64)
        st = np.zeros(4, np.int64)
        got = H.hs_exact_dot(len(x), x, y, P(st))
        want = H.hs_serial_dot(len(x), x, y)
        assert got == want or (np.isnan(got) and np.isnan(want)), (name, got, want)
        assert st[3] == 0, (name, "fp and integer map formulations disagree")
        assert st[2] < 1000000, (name, "gap decomposition does not tile the gap")
        want_np = 0.0
        if len(x) <= 5000:
            for a, b in zip(x, y):
                want_np = want_np + a * b
            assert got == want_np or (np.isnan(got) and np.isnan(want_np)), name


def test_exact_dot_on_cg_vectors(H, oracle_built):
    """r.r and p.Ap of a real CG run (the data the kernel sees)."""
    o = Oracle(24, 24, 24).setup()
    rows, ro, cols, coefs = o.csr()
    b = o.b()
    r = b.copy(); p = r.copy()
    for it in range(6):
        Ap = o.matvec(p)
        for a, c in ((r, r), (Ap, p)):
            st = np.zeros(4, np.int64)
            assert H.hs_exact_dot(len(a), a, c, P(st)) == Oracle.dot(a, c)
            assert st[1] == 0 and st[0] < 64, "unexpected number of dirty leaves on CG data: %d" % st[0]
        alpha = Oracle.dot(r, r) / Oracle.dot(Ap, p)
        rn = Oracle.waxpby(1.0, r, -alpha, Ap)
        beta = Oracle.dot(rn, rn) / Oracle.dot(r, r)
        p = Oracle.waxpby(1.0, rn, beta, p); r = rn


def test_exact_dot_randomised_stress(H):
    """Randomised classes the CG never produces but a caller of mfe_dot may: wide dynamic range, sign flips, sums that
    return to (or sit exactly on) a power of two, runs of zeros between exact values, sparse spikes."""
    rng = np.random.default_rng(20261017)
    trials = 0
    for n in (17, 513, 4097, 33000, 131072 + 5):
        for kind in range(8):
            if kind == 0:      # wide dynamic range, positive
                x = rng.random(n) * 10.0 ** rng.integers(-30, 30, n); y = x
            elif kind == 1:    # random signs, wide range
                x = rng.standard_normal(n) * 10.0 ** rng.integers(-12, 12, n); y = rng.standard_normal(n)
            elif kind == 2:    # exact small integers: sums hit powers of two exactly, with zero runs in between
                x = (rng.random(n) < 0.02).astype(np.float64); y = x
            elif kind == 3:    # cancellation back to exactly zero and up again
                x = np.tile([1.0, -1.0, 0.5, -0.5], n // 4 + 1)[:n]; y = np.ones(n)
            elif kind == 4:    # halves and quarters: every addition is an exact tie candidate
                x = rng.integers(-3, 4, n) * 0.25; y = np.ones(n) * (1.0 + 2.0 ** -52)
            elif kind == 5:    # slow drift around a binade boundary
                x = np.full(n, 2.0 ** -30); x[0] = 1.0 - 2.0 ** -20; y = np.where(rng.random(n) < 0.5, 1.0, -1.0); y[0] = 1.0
            elif kind == 6:    # sparse spikes over a tiny background
                x = rng.random(n) * 1e-12; x[rng.integers(0, n, 5)] = 1e6; y = x
            else:              # negative running sum
                x = -np.abs(rng.standard_normal(n)); y = np.abs(rng.standard_normal(n)) + 0.01
            x, y = np.ascontiguousarray(x, np.float64), np.ascontiguousarray(y, np.float64)
            st = np.zeros(4, np.int64)
            got = H.hs_exact_dot(n, x, y, P(st))
            want = H.hs_serial_dot(n, x, y)
            assert got == want or (np.isnan(got) and np.isnan(want)), (n, kind, got, want)
            assert st[3] == 0, (n, kind, "fp and integer map formulations disagree")
            trials += 1
    assert trials == 40
