"""Host-side mirror of the miniFE operator API (see package docstring).  Each callable cites the reference
interface it stands for; all arithmetic happens in libminife_b200.so on the GPU."""
import ctypes as C

import numpy as np

from ._lib import lib, check, LibraryError

# ref/utils/mytimer.hpp:37-44
WAXPY, DOT, MATVEC, MATVECDOT, TOTAL, NUM_TIMERS = 0, 1, 2, 3, 4, 5
DOT_EXACT, DOT_SERIAL, DOT_TREE = 0, 1, 2
_MAPS = ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length")


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Box:
    """ref/utils/Box.hpp:42-52 -- ``ranges[6]``; ``box[axis][0|1]``."""

    def __init__(self, ranges=(0, 0, 0, 0, 0, 0)):
        self.ranges = np.array(ranges, dtype=np.int32).reshape(6).copy()

    def __getitem__(self, axis):
        return self.ranges[2 * axis:2 * axis + 2]

    def __repr__(self):
        return "Box(%s)" % (self.ranges.tolist(),)


class Parameters:
    """ref/utils/Parameters.hpp:35-59 (the fields the hot path reads)."""

    def __init__(self, nx=10, ny=None, nz=None, **kw):
        self.nx = nx
        self.ny = nx if ny is None else ny
        self.nz = self.ny if nz is None else nz
        self.mv_overlap_comm_comp = 0
        self.verify_solution = 0
        self.__dict__.update(kw)


def box_partition(ip, up, axis, box, p_box=None):
    """ref/utils/BoxPartition.cpp:79-122, called as ``box_partition(0, numprocs, 2, global_box, local_boxes)``
    (ref/src/main.cpp:112).  Returns the list of per-rank Boxes."""
    if ip != 0 or axis != 2:
        raise ValueError("only the reference's own call box_partition(0, np, 2, ...) is supported")
    r = box.ranges
    if r[0] != 0 or r[2] != 0 or r[4] != 0:
        raise ValueError("global box must start at 0")
    out = np.zeros(6 * up, np.int32)
    check(lib().mfe_box_partition(int(r[1]), int(r[3]), int(r[5]), up, out), "box_partition")
    boxes = [Box(out[6 * p:6 * p + 6]) for p in range(up)]
    if p_box is not None:
        p_box[:] = boxes
    return boxes


class simple_mesh_description:
    """ref/src/simple_mesh_description.hpp:38-244.  Owns the device context of this rank (the reference keeps the
    id->row std::map and the bc_rows std::sets here; on the device they are closed-form)."""

    def __init__(self, global_box, local_box, rank=0, nranks=1, all_boxes=None, device=0):
        self.global_box, self.local_box = global_box, local_box
        self.rank, self.nranks, self.device = rank, nranks, device
        if all_boxes is None:
            if nranks != 1:
                raise ValueError("all_boxes (from box_partition) is required when nranks > 1")
            all_boxes = [local_box]
        flat = np.concatenate([b.ranges for b in all_boxes]).astype(np.int32)
        if not np.array_equal(flat[6 * rank:6 * rank + 6], local_box.ranges):
            raise ValueError("local_box is not all_boxes[rank]")
        self._h = lib().mfe_create(rank, nranks, device, np.ascontiguousarray(global_box.ranges), np.ascontiguousarray(flat))
        if not self._h:
            raise LibraryError("mfe_create failed: %s" % lib().mfe_last_error().decode(errors="replace"))

    def comm_init(self, unique_id):
        """Bind this rank to the NCCL communicator (stands for MPI_COMM_WORLD)."""
        check(lib().mfe_comm_init(self._h, unique_id), "mfe_comm_init")

    def close(self):
        if getattr(self, "_h", None):
            lib().mfe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def phase_ms(self, phase):
        return float(lib().mfe_phase_ms(self._h, phase.encode()))

    def launch_count(self):
        return int(lib().mfe_launch_count(self._h))

    def set_dot_mode(self, mode):
        check(lib().mfe_set_dot_mode(self._h, mode), "mfe_set_dot_mode")

    def set_option(self, name, value):
        check(lib().mfe_set_option(self._h, name.encode(), int(value)), "mfe_set_option(%s)" % name)

    def set_op_timing(self, enabled):
        check(lib().mfe_set_op_timing(self._h, int(enabled)), "mfe_set_op_timing")

    def time_kernel(self, what, reps=10):
        ms = C.c_double(0)
        check(lib().mfe_time_kernel(self._h, what.encode(), reps, C.byref(ms)), "mfe_time_kernel(%s)" % what)
        return ms.value


def comm_unique_id():
    buf = C.create_string_buffer(128)
    check(lib().mfe_comm_unique_id(buf), "mfe_comm_unique_id")
    return buf.raw


class CSRMatrix:
    """ref/src/CSRMatrix.hpp:40-129.  The arrays live on the device; the host-visible fields of the reference
    struct are materialised on access (``rows``, ``row_offsets``, ``packed_cols``, ``packed_coefs`` and the
    index maps of make_local_matrix)."""

    def __init__(self):
        self.has_local_indices = False
        self.num_cols = 0
        self._mesh = None

    def _ctx(self):
        if self._mesh is None or not self._mesh._h:
            raise LibraryError("CSRMatrix is not bound to a mesh: call generate_matrix_structure(mesh, A) first")
        return self._mesh._h

    def num_rows(self):
        return int(lib().mfe_nrows(self._ctx()))

    def num_nonzeros(self):
        return int(lib().mfe_nnz(self._ctx()))

    def _export(self, rows=False, offs=False, cols=False, coefs=False, offs64=False):
        n, nnz = self.num_rows(), self.num_nonzeros()
        r = np.zeros(n, np.int32) if rows else None
        o = np.zeros(n + 1, np.int32) if offs else None
        o64 = np.zeros(n + 1, np.int64) if offs64 else None
        c = np.zeros(nnz, np.int32) if cols else None
        v = np.zeros(nnz, np.float64) if coefs else None
        check(lib().mfe_export_csr(self._ctx(), _vp(r), _vp(o), _vp(o64), _vp(c), _vp(v)), "mfe_export_csr")
        return r, o, o64, c, v

    rows = property(lambda self: self._export(rows=True)[0])
    row_offsets = property(lambda self: self._export(offs=True)[1])
    row_offsets64 = property(lambda self: self._export(offs64=True)[2])
    packed_cols = property(lambda self: self._export(cols=True)[3])
    packed_coefs = property(lambda self: self._export(coefs=True)[4])

    def _map(self, name):
        n = lib().mfe_get_map(self._ctx(), name.encode(), None)
        out = np.zeros(max(int(n), 0), np.int32)
        if n > 0:
            lib().mfe_get_map(self._ctx(), name.encode(), _vp(out))
        return out


for _m in _MAPS:
    setattr(CSRMatrix, _m, property(lambda self, _m=_m: self._map(_m)))


class Vector:
    """ref/src/Vector.hpp:39-61: ``startIndex``, ``local_size``, ``coefs`` (zero-filled host array)."""

    def __init__(self, startIndex, local_sz):
        self.startIndex = int(startIndex)
        self.local_size = int(local_sz)
        self.coefs = np.zeros(self.local_size, np.float64)


def generate_matrix_structure(mesh, A):
    """ref/src/generate_matrix_structure.hpp:47-149; returns 0 like the reference."""
    nrows, first, nnz = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    check(lib().mfe_generate_matrix_structure(mesh._h, C.byref(nrows), C.byref(first), C.byref(nnz)), "generate_matrix_structure")
    A._mesh = mesh
    A.has_local_indices = False
    A.num_cols = 0
    return 0


def assemble_FE_data(mesh, A, b, params=None):
    """ref/src/assemble_FE_data.hpp:38-73.  ``b`` (a Vector) receives the assembled right-hand side."""
    if A._mesh is not mesh:
        raise LibraryError("A was not generated from this mesh")
    check(lib().mfe_assemble_FE_data(mesh._h), "assemble_FE_data")
    if b is not None:
        check(lib().mfe_export_vector(mesh._h, b"b", b.coefs), "export b")


def impose_dirichlet(prescribed_value, A, b, global_nx=None, global_ny=None, global_nz=None, bc_rows=0):
    """ref/src/SparseMatrix_functions.hpp:431-480.  ``bc_rows``: 0 stands for ``mesh.bc_rows_0``, 1 for
    ``mesh.bc_rows_1`` (the sets themselves are analytic on the device)."""
    h = A._ctx()
    check(lib().mfe_impose_dirichlet(h, float(prescribed_value), int(bc_rows)), "impose_dirichlet")
    if b is not None:
        check(lib().mfe_export_vector(h, b"b", b.coefs), "export b")


def make_local_matrix(A):
    """ref/src/make_local_matrix.hpp:42-442."""
    nc = C.c_int64(0)
    check(lib().mfe_make_local_matrix(A._ctx(), C.byref(nc)), "make_local_matrix")
    A.num_cols = int(nc.value)
    A.has_local_indices = True


def rearrange_matrix_local_external(A):
    """ref/src/SparseMatrix_functions.hpp:319-408 (driver option mv_overlap_comm_comp=1, driver.hpp:278-282)."""
    check(lib().mfe_rearrange_matrix_local_external(A._ctx()), "rearrange_matrix_local_external")


class matvec_overlap:
    """ref/src/SparseMatrix_functions.hpp:624-675: after rearrange_matrix_local_external the device SpMV sums every row in
    this functor's order (local entries, then externals); the halo exchange is overlapped on the device either way."""

    def __call__(self, A, x, y):
        matvec_std()(A, x, y)


class matvec_std:
    """ref/src/SparseMatrix_functions.hpp:546-578: ``matvec_std()(A, x, y)`` (x has num_cols entries)."""

    def __call__(self, A, x, y):
        xs = x.coefs if isinstance(x, Vector) else x
        ys = y.coefs if isinstance(y, Vector) else y
        check(lib().mfe_matvec(A._ctx(), xs, ys), "matvec")


def matvec(A, x, y):
    """ref/src/SparseMatrix_functions.hpp:616-622."""
    matvec_std()(A, x, y)


def exchange_externals(A, x):
    """ref/src/exchange_externals.hpp:44-161."""
    xs = x.coefs if isinstance(x, Vector) else x
    check(lib().mfe_exchange_externals(A._ctx(), xs), "exchange_externals")


def _any_ctx(*objs):
    for o in objs:
        h = getattr(o, "_ctx_handle", None)
        if h:
            return h
    raise LibraryError("waxpby/dot need a device context: pass ctx=mesh")


def waxpby(alpha, x, beta, y, w, ctx=None):
    """ref/src/Vector_functions.hpp:127-150: ``w = alpha*x + beta*y`` over ``len(x)`` entries."""
    xs, ys, ws = (v.coefs if isinstance(v, Vector) else v for v in (x, y, w))
    check(lib().mfe_waxpby(ctx._h, len(xs), float(alpha), xs, float(beta), ys, ws), "waxpby")


def dot(x, y, ctx=None):
    """ref/src/Vector_functions.hpp:194-226 (serial-order sum + rank-ordered allreduce)."""
    xs, ys = (v.coefs if isinstance(v, Vector) else v for v in (x, y))
    r = C.c_double(0)
    check(lib().mfe_dot(ctx._h, len(xs), xs, ys, C.byref(r)), "dot")
    return r.value


def cg_solve(A, b, x, matvec_op=None, max_iter=200, tolerance=float(np.finfo(np.float64).eps), history=False,
             upload_b=False):
    """ref/src/cg_solve.hpp:69-229.  Returns ``(num_iters, normr, cg_times[NUM_TIMERS])`` (the reference's three
    output arguments); with ``history=True`` a dict of per-iteration scalars is appended.  ``x.coefs`` is the
    initial guess on entry and the solution on exit (host buffer, copied both ways inside the call).  ``b`` stays
    device resident from assembly unless ``upload_b``."""
    h = A._ctx()
    L = lib()
    if upload_b:
        check(L.mfe_import_vector(h, b"b", b.coefs), "import b")
    it, nr = C.c_int(0), C.c_double(0)
    times = np.zeros(NUM_TIMERS)
    hist = {k: np.full(max_iter + 1, np.nan) for k in ("normr", "rtrans", "pap", "alpha")} if history else None
    rc = L.mfe_cg_solve(h, _vp(x.coefs) if x is not None else None, int(max_iter), float(tolerance), C.byref(it), C.byref(nr),
                        _vp(times), *([_vp(hist[k]) for k in ("normr", "rtrans", "pap", "alpha")] if history else [None] * 4))
    if rc not in (0, 2):          # 2 = numerical breakdown: outputs are valid, like the reference's early return
        check(rc, "cg_solve")
    if history:
        k = it.value
        return it.value, nr.value, times, dict(rc=rc, **{kk: v[:k + 1] for kk, v in hist.items()})
    return it.value, nr.value, times


def verify_solution(mesh, x=None, tolerance=0.06, verify_whole_domain=False, info=False):
    """ref/fem/verify_solution.hpp:55-174 on the DEVICE-resident solution (the x left by cg_solve; pass a Vector /
    array to upload a different one first).  Returns the reference's return code (0 = within tolerance); with
    ``info=True`` also ``(global_max_err, dict(err, computed, analytic, coords))`` of this rank's worst node."""
    L = lib()
    if x is not None:
        xs = x.coefs if isinstance(x, Vector) else x
        check(L.mfe_import_vector(mesh._h, b"x", np.ascontiguousarray(xs, np.float64)), "import x")
    rc, gmax, loc = C.c_int(0), C.c_double(0), np.zeros(6)
    check(L.mfe_verify_solution(mesh._h, float(tolerance), int(bool(verify_whole_domain)), C.byref(rc), C.byref(gmax), _vp(loc)), "verify_solution")
    if info:
        return rc.value, gmax.value, dict(err=loc[0], computed=loc[1], analytic=loc[2], coords=tuple(loc[3:6]))
    return rc.value


def time_kernels(A, b, x, matvec_op=None, max_iter=500):
    """ref/src/time_kernels.hpp:42-134 (MINIFE_KERNELS != 0): ``max_iter`` x waxpby(1,x,0,x,p), ``max_iter`` x
    matvec(A,p,x), ``xdotp`` = sum of ``max_iter`` x dot(x,p).  Returns ``(xdotp, kern_times[NUM_TIMERS])``; ``x.coefs``
    holds A*p afterwards, as in the reference."""
    r, times = C.c_double(0), np.zeros(NUM_TIMERS)
    check(lib().mfe_time_kernels(A._ctx(), _vp(x.coefs) if x is not None else None, int(max_iter), C.byref(r), _vp(times)), "time_kernels")
    return r.value, times
