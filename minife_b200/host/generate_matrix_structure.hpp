// Shadows ref/src/generate_matrix_structure.hpp (same name, namespace and signature) -> mfe_generate_matrix_structure.
#ifndef _generate_matrix_structure_hpp_
#define _generate_matrix_structure_hpp_

#include <b200_bridge.hpp>
#include <simple_mesh_description.hpp>

namespace miniFE {

// ref/src/generate_matrix_structure.hpp:47-149.  Returns 0, or 1 when any rank failed (the reference allreduces
// an exception flag; the driver ignores the value, driver.hpp:181-183).
template<typename MatrixType>
int
generate_matrix_structure(const simple_mesh_description<typename MatrixType::GlobalOrdinalType>& mesh,
                          MatrixType& A)
{
  typedef typename MatrixType::GlobalOrdinalType GlobalOrdinal;
  typedef typename MatrixType::LocalOrdinalType LocalOrdinal;
  b200::Entry &e = b200::table()[&A];
  if (e.ctx) mfe_destroy(e.ctx);
  e = b200::Entry();
  e.ctx = b200::create_context(mesh.global_box, mesh.local_box);
  e.bc_rows_0 = &mesh.bc_rows_0;
  e.bc_rows_1 = &mesh.bc_rows_1;
  int64_t nrows = 0, first = 0, nnz = 0;
  int threw_exc = mfe_generate_matrix_structure(e.ctx, &nrows, &first, &nnz) != 0 ? 1 : 0;
  if (threw_exc)
    std::cout << "proc threw an exception in generate_matrix_structure, probably due to running out of memory: " << mfe_last_error() << std::endl;
  else {
    // host-visible fields the unchanged driver reads: rows (driver.hpp:185-186) and row_offsets (num_nonzeros(),
    // compute_matrix_stats.hpp:62).  packed_cols / packed_coefs stay on the device (write_matrix exports on demand).
    A.rows.resize(nrows);
    A.row_offsets.resize(nrows + 1);
    if (nnz <= 0x7fffffffLL) {
      if (mfe_export_csr(e.ctx, (int *)(sizeof(GlobalOrdinal) == 4 ? &A.rows[0] : 0), (int *)(sizeof(LocalOrdinal) == 4 ? &A.row_offsets[0] : 0), 0, 0, 0) != 0)
        b200::die("mfe_export_csr");
    }
    if (sizeof(GlobalOrdinal) != 4) for (int64_t i = 0; i < nrows; ++i) A.rows[i] = (GlobalOrdinal)(first + i);
    if (sizeof(LocalOrdinal) != 4 || nnz > 0x7fffffffLL) {
      std::vector<int64_t> ro(nrows + 1);
      if (mfe_export_csr(e.ctx, 0, 0, &ro[0], 0, 0) != 0) b200::die("mfe_export_csr");
      for (int64_t i = 0; i <= nrows; ++i) A.row_offsets[i] = (LocalOrdinal)ro[i];   // wraps for int at 512^3, as the reference's own int would
    }
    A.has_local_indices = false;
  }
#ifdef HAVE_MPI
  int global_throw = 0;
  MPI_Allreduce(&threw_exc, &global_throw, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  threw_exc = global_throw;
#endif
  return threw_exc ? 1 : 0;
}

}//namespace miniFE

#endif
