"""Loader of libminife_b200.so (the C ABI of include/minife_b200.h).  Fails loudly: there is no fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class LibraryError(RuntimeError):
    pass


def library_path():
    return os.path.join(_HERE, "libminife_b200.so")


_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_i64 = C.c_int64
_pi64 = C.POINTER(C.c_int64)

# every symbol include/minife_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mfe_abi_version": (C.c_int, []),
    "mfe_last_error": (C.c_char_p, []),
    "mfe_device_count": (C.c_int, []),
    "mfe_box_partition": (C.c_int, [C.c_int] * 4 + [_i32p]),
    "mfe_create": (_vp, [C.c_int, C.c_int, C.c_int, _i32p, _i32p]),
    "mfe_destroy": (None, [_vp]),
    "mfe_comm_unique_id": (C.c_int, [C.c_char_p]),
    "mfe_comm_init": (C.c_int, [_vp, C.c_char_p]),
    "mfe_generate_matrix_structure": (C.c_int, [_vp, _pi64, _pi64, _pi64]),
    "mfe_assemble_FE_data": (C.c_int, [_vp]),
    "mfe_impose_dirichlet": (C.c_int, [_vp, C.c_double, C.c_int]),
    "mfe_make_local_matrix": (C.c_int, [_vp, _pi64]),
    "mfe_rearrange_matrix_local_external": (C.c_int, [_vp]),
    "mfe_cg_solve": (C.c_int, [_vp, _vp, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double), _vp, _vp, _vp, _vp, _vp]),
    "mfe_time_kernels": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_double), _vp]),
    "mfe_verify_solution": (C.c_int, [_vp, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), _vp]),
    "mfe_set_dot_mode": (C.c_int, [_vp, C.c_int]),
    "mfe_set_op_timing": (C.c_int, [_vp, C.c_int]),
    "mfe_set_option": (C.c_int, [_vp, C.c_char_p, _i64]),
    "mfe_matvec": (C.c_int, [_vp, _f64p, _f64p]),
    "mfe_waxpby": (C.c_int, [_vp, _i64, C.c_double, _f64p, C.c_double, _f64p, _f64p]),
    "mfe_dot": (C.c_int, [_vp, _i64, _f64p, _f64p, C.POINTER(C.c_double)]),
    "mfe_exchange_externals": (C.c_int, [_vp, _f64p]),
    "mfe_nrows": (_i64, [_vp]),
    "mfe_first_row": (_i64, [_vp]),
    "mfe_nnz": (_i64, [_vp]),
    "mfe_num_cols": (_i64, [_vp]),
    "mfe_get_map": (_i64, [_vp, C.c_char_p, _vp]),
    "mfe_export_csr": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mfe_export_vector": (C.c_int, [_vp, C.c_char_p, _f64p]),
    "mfe_import_vector": (C.c_int, [_vp, C.c_char_p, _f64p]),
    "mfe_import_csr": (C.c_int, [_vp, _i64, _i64, _i32p, _i32p, _f64p]),
    "mfe_time_kernel": (C.c_int, [_vp, C.c_char_p, C.c_int, C.POINTER(C.c_double)]),
    "mfe_dot_stats": (_i64, [_vp, _vp, _i64, C.c_int]),
    "mfe_launch_count": (_i64, [_vp]),
    "mfe_phase_ms": (C.c_double, [_vp, C.c_char_p]),
}

_lib = None


def lib():
    """The loaded library with typed signatures.  Raises LibraryError when the .so is missing or lacks a symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise LibraryError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C minife_b200/csrc`).  minife_b200 has no CPU fallback." % path)
    try:
        L = C.CDLL(path)
    except OSError as e:
        raise LibraryError("cannot load %s: %s" % (path, e))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            raise LibraryError("%s does not export %s (stale build?)" % (path, name))
        fn.restype = res
        fn.argtypes = args
    if L.mfe_abi_version() != 1:
        raise LibraryError("ABI version mismatch in %s" % path)
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise LibraryError("%s failed: %s" % (what, lib().mfe_last_error().decode(errors="replace")))
