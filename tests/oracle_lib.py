"""ctypes bindings of the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  -- oracle/libminife_oracle.so, the plain-C restatement (oracle/minife_oracle.c).
* ``Ref``     -- oracle/_ref/libminife_ref.so, the UNMODIFIED reference headers bound by
                 oracle/ref_entry.cpp (built in the authoring container where /root/reference exists;
                 the prebuilt .so travels to the GPU box).
* ``ref_dump_mpi`` -- runs oracle/_ref/ref_dump_mpi.x on N ranks over oracle/mpi_shim and loads its dumps.

Nothing under minife_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libminife_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libminife_ref.so")
REF_DUMP = os.path.join(ORACLE_DIR, "_ref", "ref_dump_mpi.x")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
EPS = float(np.finfo(np.float64).eps)


def build_oracle():
    """(Re)build the C restatement; cheap (<2 s)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def have_ref():
    return os.path.exists(REF_SO)


def have_ref_dump():
    return os.path.exists(REF_DUMP)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """All ranks of one nx*ny*nz problem, in-process (oracle/minife_oracle.h)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build_oracle()
            L = C.CDLL(ORACLE_SO)
            L.mfo_world_create.restype = C.c_void_p
            L.mfo_world_create.argtypes = [C.c_int] * 4
            L.mfo_world_create_boxes.restype = C.c_void_p
            L.mfo_world_create_boxes.argtypes = [C.c_int] * 4 + [C.c_void_p]
            L.mfo_world_destroy.argtypes = [C.c_void_p]
            for n in ("mfo_generate_matrix_structure", "mfo_assemble_FE_data", "mfo_make_local_matrix"):
                getattr(L, n).argtypes = [C.c_void_p]
            L.mfo_impose_dirichlet.argtypes = [C.c_void_p, C.c_int]
            for n in ("mfo_nrows", "mfo_nnz", "mfo_num_cols", "mfo_first_row"):
                getattr(L, n).restype = C.c_longlong
                getattr(L, n).argtypes = [C.c_void_p, C.c_int]
            L.mfo_get_box.argtypes = [C.c_void_p, C.c_int, _i32p]
            L.mfo_get_csr.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
            L.mfo_get_b.argtypes = [C.c_void_p, C.c_int, _f64p]
            L.mfo_get_x.argtypes = [C.c_void_p, C.c_int, _f64p]
            L.mfo_set_x.argtypes = [C.c_void_p, C.c_int, _f64p]
            L.mfo_bc_rows.restype = C.c_longlong
            L.mfo_bc_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            L.mfo_map_id_to_row.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.mfo_get_map.restype = C.c_longlong
            L.mfo_get_map.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]
            L.mfo_matvec.argtypes = [C.c_void_p, C.c_int, _f64p, _f64p]
            L.mfo_waxpby.argtypes = [C.c_longlong, C.c_double, _f64p, C.c_double, _f64p, _f64p]
            L.mfo_dot.restype = C.c_double
            L.mfo_dot.argtypes = [C.c_longlong, _f64p, _f64p]
            L.mfo_cg_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double)] + [C.c_void_p] * 4
            L.mfo_verify_solution.argtypes = [C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_double), _f64p]
            L.mfo_box_partition.argtypes = [C.c_int] * 4 + [_i32p]
            L.mfo_gradient_values.argtypes = [_f64p]
            L.mfo_element.argtypes = [C.c_int] * 4 + [_i32p, _f64p, _f64p, _f64p]
            cls._lib = L
        return cls._lib

    def __init__(self, nx, ny, nz, nranks=1, boxes=None):
        """boxes: optional (nranks, 6) int array of element boxes (e.g. after the reference's add_imbalance)"""
        self.L = self.lib()
        self.nx, self.ny, self.nz, self.np = nx, ny, nz, nranks
        if boxes is None:
            self.h = self.L.mfo_world_create(nx, ny, nz, nranks)
        else:
            bx = np.ascontiguousarray(boxes, np.int32).reshape(nranks, 6)
            self.h = self.L.mfo_world_create_boxes(nx, ny, nz, nranks, bx.ctypes.data_as(C.c_void_p))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.mfo_world_destroy(self.h)
            self.h = None

    # phases
    def generate_matrix_structure(self):
        return self.L.mfo_generate_matrix_structure(self.h)

    def assemble_FE_data(self):
        self.L.mfo_assemble_FE_data(self.h)

    def impose_dirichlet(self, which):
        self.L.mfo_impose_dirichlet(self.h, which)

    def make_local_matrix(self):
        self.L.mfo_make_local_matrix(self.h)

    def setup(self, upto="local"):
        """structure -> assembly -> BC(0), BC(1) -> make_local (driver.hpp:181-230)."""
        self.generate_matrix_structure()
        if upto == "structure":
            return self
        self.assemble_FE_data()
        if upto == "assembly":
            return self
        self.impose_dirichlet(0)
        self.impose_dirichlet(1)
        if upto == "bc":
            return self
        self.make_local_matrix()
        return self

    def cg_solve(self, max_iter=200, tol=EPS):
        n = max_iter + 1
        hist = {k: np.full(n, np.nan) for k in ("normr", "rtrans", "pap", "alpha")}
        it, nr = C.c_int(0), C.c_double(0)
        rc = self.L.mfo_cg_solve(self.h, max_iter, tol, C.byref(it), C.byref(nr), _ptr(hist["normr"]),
                                 _ptr(hist["rtrans"]), _ptr(hist["pap"]), _ptr(hist["alpha"]))
        k = it.value
        return dict(rc=rc, num_iters=k, final_normr=nr.value, **{kk: v[:k + 1] for kk, v in hist.items()})

    def verify_solution(self, tol=0.06, whole_domain=False, info=False):
        g, loc = C.c_double(0), np.zeros(6)
        rc = self.L.mfo_verify_solution(self.h, tol, int(whole_domain), C.byref(g), loc)
        return (rc, g.value, loc) if info else rc

    # queries
    def nrows(self, r=0): return int(self.L.mfo_nrows(self.h, r))
    def nnz(self, r=0): return int(self.L.mfo_nnz(self.h, r))
    def num_cols(self, r=0): return int(self.L.mfo_num_cols(self.h, r))
    def first_row(self, r=0): return int(self.L.mfo_first_row(self.h, r))

    def box(self, r=0):
        b = np.zeros(6, np.int32)
        self.L.mfo_get_box(self.h, r, b)
        return b

    def csr(self, r=0):
        n, nnz = self.nrows(r), self.nnz(r)
        rows, ro = np.zeros(n, np.int32), np.zeros(n + 1, np.int32)
        cols, coefs = np.zeros(nnz, np.int32), np.zeros(nnz, np.float64)
        self.L.mfo_get_csr(self.h, r, _ptr(rows), _ptr(ro), _ptr(cols), _ptr(coefs))
        return rows, ro, cols, coefs

    def b(self, r=0):
        o = np.zeros(self.nrows(r)); self.L.mfo_get_b(self.h, r, o); return o

    def x(self, r=0):
        o = np.zeros(self.nrows(r)); self.L.mfo_get_x(self.h, r, o); return o

    def bc_rows(self, which, r=0):
        n = self.L.mfo_bc_rows(self.h, r, which, None)
        o = np.zeros(n, np.int32)
        self.L.mfo_bc_rows(self.h, r, which, _ptr(o))
        return o

    def get_map(self, name, r=0):
        n = self.L.mfo_get_map(self.h, r, name.encode(), None)
        o = np.zeros(max(n, 0), np.int32)
        if n > 0:
            self.L.mfo_get_map(self.h, r, name.encode(), _ptr(o))
        return o

    def matvec(self, x, r=0, y0=None):
        """y0: the stale contents of y (the reference forms beta*y[row] with beta = 0, SparseMatrix_functions.hpp:565)"""
        y = np.zeros(self.nrows(r)) if y0 is None else np.array(y0, np.float64)
        self.L.mfo_matvec(self.h, r, np.ascontiguousarray(x, np.float64), y)
        return y

    def rearrange_matrix_local_external(self):
        """ref/src/SparseMatrix_functions.hpp:319-408: from here on matvec / cg_solve sum the rows as matvec_overlap does"""
        self.L.mfo_rearrange_matrix_local_external.argtypes = [C.c_void_p]
        self.L.mfo_rearrange_matrix_local_external(self.h)
        return self

    def time_kernels(self, f, max_iter):
        """ref/src/time_kernels.hpp:97-124 on every rank, with x_global[row] = f(row) on entry: p = 1*x + 0*x (+ the halo
        exchange), max_iter x matvec(A, p, x), xdotp = sum over max_iter x dot(x, p) (rank-ordered global sums).
        Returns ([x of every rank afterwards], xdotp)."""
        xs, ps = [], []
        for r in range(self.np):
            nrows, first = self.nrows(r), int(self.csr(r)[0][0])
            x = np.asarray(f(first + np.arange(nrows)), np.float64)
            p = np.zeros(nrows + len(self.get_map("external_index", r)))
            p[:nrows] = self.waxpby(1.0, x, 0.0, x)
            p[self.get_map("external_local_index", r)] = f(self.get_map("external_index", r).astype(np.float64))
            xs.append(x); ps.append(p)
        for _ in range(max_iter):
            xs = [self.matvec(ps[r], r, y0=xs[r]) for r in range(self.np)]
        acc = 0.0
        for _ in range(max_iter):
            g = None
            for r in range(self.np):
                l = self.dot(xs[r], ps[r][:len(xs[r])])
                g = l if g is None else g + l
            acc = acc + g
        return xs, acc

    @classmethod
    def waxpby(cls, alpha, x, beta, y):
        w = np.zeros_like(x)
        cls.lib().mfo_waxpby(len(x), alpha, np.ascontiguousarray(x), beta, np.ascontiguousarray(y), w)
        return w

    @classmethod
    def dot(cls, x, y):
        return float(cls.lib().mfo_dot(len(x), np.ascontiguousarray(x), np.ascontiguousarray(y)))

    @classmethod
    def box_partition(cls, nx, ny, nz, np_):
        o = np.zeros(6 * np_, np.int32)
        cls.lib().mfo_box_partition(nx, ny, nz, np_, o)
        return o.reshape(np_, 6)

    @classmethod
    def gradient_values(cls):
        o = np.zeros(192); cls.lib().mfo_gradient_values(o); return o

    @classmethod
    def element(cls, nx, ny, nz, elem_id):
        ids, c, K, f = np.zeros(8, np.int32), np.zeros(24), np.zeros(36), np.zeros(8)
        cls.lib().mfo_element(nx, ny, nz, elem_id, ids, c, K, f)
        return ids, c, K, f


class Ref:
    """Serial (np=1) view of the unmodified reference (oracle/ref_entry.cpp)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REF_SO)
            L.ref_create.restype = C.c_void_p
            L.ref_create.argtypes = [C.c_int] * 3
            L.ref_destroy.argtypes = [C.c_void_p]
            L.ref_generate_matrix_structure.argtypes = [C.c_void_p]
            L.ref_assemble_FE_data.argtypes = [C.c_void_p]
            L.ref_impose_dirichlet.argtypes = [C.c_void_p, C.c_int]
            L.ref_make_local_matrix.argtypes = [C.c_void_p]
            for n in ("ref_nrows", "ref_nnz", "ref_num_cols", "ref_first_row"):
                getattr(L, n).restype = C.c_longlong
                getattr(L, n).argtypes = [C.c_void_p]
            L.ref_get_csr.argtypes = [C.c_void_p] * 5
            L.ref_get_b.argtypes = [C.c_void_p, _f64p]
            L.ref_get_x.argtypes = [C.c_void_p, _f64p]
            L.ref_set_x.argtypes = [C.c_void_p, _f64p]
            L.ref_bc_rows.restype = C.c_longlong
            L.ref_bc_rows.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
            L.ref_map_id_to_row.argtypes = [C.c_void_p, C.c_int]
            L.ref_element.argtypes = [C.c_void_p, C.c_int, _i32p, _f64p, _f64p, _f64p]
            L.ref_gradient_values.argtypes = [_f64p]
            L.ref_matvec.argtypes = [C.c_void_p, _f64p, _f64p]
            L.ref_waxpby.argtypes = [C.c_int, C.c_double, _f64p, C.c_double, _f64p, _f64p]
            L.ref_dot.restype = C.c_double
            L.ref_dot.argtypes = [C.c_int, _f64p, _f64p]
            L.ref_cg_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                       C.c_void_p, C.c_void_p, C.c_int]
            L.ref_cg_trace.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double)] + [C.c_void_p] * 4
            L.ref_verify_solution.argtypes = [C.c_void_p, C.c_double, C.c_int]
            L.ref_box_partition.argtypes = [C.c_int] * 4 + [_i32p]
            cls._lib = L
        return cls._lib

    def __init__(self, nx, ny, nz):
        self.L = self.lib()
        self.nx, self.ny, self.nz, self.np = nx, ny, nz, 1
        self.h = self.L.ref_create(nx, ny, nz)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_destroy(self.h)
            self.h = None

    def generate_matrix_structure(self): return self.L.ref_generate_matrix_structure(self.h)
    def assemble_FE_data(self): self.L.ref_assemble_FE_data(self.h)
    def impose_dirichlet(self, which): self.L.ref_impose_dirichlet(self.h, which)
    def make_local_matrix(self): self.L.ref_make_local_matrix(self.h)
    setup = Oracle.setup

    def cg_solve(self, max_iter=200, tol=EPS, trace=True):
        n = max_iter + 1
        hist = {k: np.full(n, np.nan) for k in ("normr", "rtrans", "pap", "alpha")}
        it, nr = C.c_int(0), C.c_double(0)
        if trace:
            rc = self.L.ref_cg_trace(self.h, max_iter, tol, C.byref(it), C.byref(nr), _ptr(hist["normr"]),
                                     _ptr(hist["rtrans"]), _ptr(hist["pap"]), _ptr(hist["alpha"]))
        else:
            rc = self.L.ref_cg_solve(self.h, max_iter, tol, C.byref(it), C.byref(nr), None, _ptr(hist["normr"]), n)
        k = it.value
        return dict(rc=rc, num_iters=k, final_normr=nr.value, **{kk: v[:k + 1] for kk, v in hist.items()})

    def nrows(self, r=0): return int(self.L.ref_nrows(self.h))
    def nnz(self, r=0): return int(self.L.ref_nnz(self.h))
    def num_cols(self, r=0): return int(self.L.ref_num_cols(self.h))
    def first_row(self, r=0): return int(self.L.ref_first_row(self.h))

    def csr(self, r=0):
        n, nnz = self.nrows(), self.nnz()
        rows, ro = np.zeros(n, np.int32), np.zeros(n + 1, np.int32)
        cols, coefs = np.zeros(nnz, np.int32), np.zeros(nnz, np.float64)
        self.L.ref_get_csr(self.h, _ptr(rows), _ptr(ro), _ptr(cols), _ptr(coefs))
        return rows, ro, cols, coefs

    def b(self, r=0):
        o = np.zeros(self.nrows()); self.L.ref_get_b(self.h, o); return o

    def x(self, r=0):
        o = np.zeros(self.nrows()); self.L.ref_get_x(self.h, o); return o

    def bc_rows(self, which, r=0):
        n = self.L.ref_bc_rows(self.h, which, None)
        o = np.zeros(n, np.int32)
        self.L.ref_bc_rows(self.h, which, _ptr(o))
        return o

    def matvec(self, x, r=0, y0=None):
        y = np.zeros(self.nrows()) if y0 is None else np.array(y0, np.float64)
        self.L.ref_matvec(self.h, np.ascontiguousarray(x, np.float64), y)
        return y

    def element(self, elem_id):
        ids, c, K, f = np.zeros(8, np.int32), np.zeros(24), np.zeros(36), np.zeros(8)
        self.L.ref_element(self.h, elem_id, ids, c, K, f)
        return ids, c, K, f

    def verify_solution(self, tol=0.06, whole_domain=False):
        return self.L.ref_verify_solution(self.h, tol, int(whole_domain))

    @classmethod
    def waxpby(cls, alpha, x, beta, y):
        w = np.zeros_like(x)
        cls.lib().ref_waxpby(len(x), alpha, np.ascontiguousarray(x), beta, np.ascontiguousarray(y), w)
        return w

    @classmethod
    def dot(cls, x, y):
        return float(cls.lib().ref_dot(len(x), np.ascontiguousarray(x), np.ascontiguousarray(y)))

    @classmethod
    def box_partition(cls, nx, ny, nz, np_):
        o = np.zeros(6 * np_, np.int32)
        cls.lib().ref_box_partition(nx, ny, nz, np_, o)
        return o.reshape(np_, 6)

    @classmethod
    def gradient_values(cls):
        o = np.zeros(192); cls.lib().ref_gradient_values(o); return o


_DUMP_INT = ("rows", "row_offsets", "cols_global", "cols_local", "box", "num_cols", "external_index",
             "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length")
_DUMP_F64 = ("coefs_prebc", "b_prebc", "coefs", "b", "normr_hist", "pap_hist", "x")


def ref_dump_mpi(nx, ny, nz, nranks, max_iter=200, timeout=600, imbalance=0, overlap=False):
    """Run the unmodified reference on `nranks` ranks (MPI shim) and return {name: [per-rank arrays]}.
    imbalance > 0: the reference's add_imbalance(load_imbalance = imbalance percent) changes the boxes first.
    overlap: the driver's mv_overlap_comm_comp=1 branch (rearrange_matrix_local_external + matvec_overlap)."""
    sys.path.insert(0, os.path.join(ORACLE_DIR, "mpi_shim"))
    import launch  # noqa
    out = {}
    os.environ["REF_DUMP_OVERLAP"] = "1" if overlap else "0"
    os.environ["REF_DUMP_LEAN"] = "0"
    with tempfile.TemporaryDirectory() as d:
        rc = launch.launch(nranks, [REF_DUMP, str(nx), str(ny), str(nz), d, str(max_iter)] + ([str(imbalance)] if imbalance else []),
                           timeout=timeout, stdout=subprocess.DEVNULL)
        if rc != 0:
            raise RuntimeError("ref_dump_mpi.x failed rc=%d" % rc)
        for name in _DUMP_INT + _DUMP_F64:
            dt, ext = (np.int32, "i32") if name in _DUMP_INT else (np.float64, "f64")
            per = []
            for r in range(nranks):
                fn = os.path.join(d, "%s.%d.%d.%s" % (name, nranks, r, ext))
                per.append(np.fromfile(fn, dtype=dt) if os.path.exists(fn) else None)
            out[name] = per
    return out
