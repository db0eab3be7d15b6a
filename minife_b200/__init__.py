"""minife_b200 -- host-side mirror (Python) of the miniFE operator API over libminife_b200.so.

The product is the CUDA library behind the C ABI in ``include/minife_b200.h``; this module only marshals.
Names and argument meaning follow the reference's free functions (``ref/src/driver.hpp:124-404`` calls them in
this order): ``box_partition`` -> ``simple_mesh_description`` -> ``generate_matrix_structure`` ->
``assemble_FE_data`` -> ``impose_dirichlet`` x2 -> ``make_local_matrix`` -> ``cg_solve`` (with ``matvec_std``,
``waxpby``, ``dot``).  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from ._lib import lib, LibraryError, library_path  # noqa: F401
from .api import (  # noqa: F401
    Box, Parameters, simple_mesh_description, CSRMatrix, Vector, box_partition,
    generate_matrix_structure, assemble_FE_data, impose_dirichlet, make_local_matrix,
    rearrange_matrix_local_external, matvec_overlap,
    cg_solve, matvec_std, matvec, waxpby, dot, exchange_externals, comm_unique_id, verify_solution, time_kernels,
    WAXPY, DOT, MATVEC, MATVECDOT, TOTAL, NUM_TIMERS, DOT_EXACT, DOT_SERIAL, DOT_TREE,
)

__all__ = [
    "lib", "LibraryError", "library_path", "Box", "Parameters", "simple_mesh_description", "CSRMatrix", "Vector",
    "box_partition", "generate_matrix_structure", "assemble_FE_data", "impose_dirichlet", "make_local_matrix",
    "rearrange_matrix_local_external", "matvec_overlap", "cg_solve", "matvec_std", "matvec", "waxpby", "dot", "exchange_externals", "comm_unique_id", "verify_solution", "time_kernels",
]
