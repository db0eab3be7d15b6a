// Shadows ref/src/generate_matrix_structure.hpp (same name, namespace and signature) -> mfe_generate_matrix_structure.
#ifndef _generate_matrix_structure_hpp_
#define _generate_matrix_structure_hpp_

#include <b200_bridge.hpp>
#include <simple_mesh_description.hpp>
#include <CSRMatrix.hpp>
#include <ELLMatrix.hpp>

namespace miniFE {

namespace b200 {

// Host-visible fields the unchanged driver reads after generate_matrix_structure: rows (driver.hpp:185-186) and what
// num_nonzeros() needs (compute_matrix_stats.hpp:62).  Columns / coefficients stay on the device (write_matrix exports
// them on demand).  One overload per matrix type the driver can be built with (driver.hpp:170-176).
template<typename Scalar, typename LocalOrdinal, typename GlobalOrdinal>
inline void publish_structure(CSRMatrix<Scalar,LocalOrdinal,GlobalOrdinal>& A, mfe_ctx *ctx, int64_t nrows, int64_t first, int64_t nnz)
{
  A.rows.resize(nrows);
  A.row_offsets.resize(nrows + 1);
  if (nnz <= 0x7fffffffLL) {
    if (mfe_export_csr(ctx, (int *)(sizeof(GlobalOrdinal) == 4 ? &A.rows[0] : 0), (int *)(sizeof(LocalOrdinal) == 4 ? &A.row_offsets[0] : 0), 0, 0, 0) != 0)
      die("mfe_export_csr");
  }
  if (sizeof(GlobalOrdinal) != 4) for (int64_t i = 0; i < nrows; ++i) A.rows[i] = (GlobalOrdinal)(first + i);
  if (sizeof(LocalOrdinal) != 4 || nnz > 0x7fffffffLL) {
    std::vector<int64_t> ro(nrows + 1);
    if (mfe_export_csr(ctx, 0, 0, &ro[0], 0, 0) != 0) die("mfe_export_csr");
    for (int64_t i = 0; i <= nrows; ++i) A.row_offsets[i] = (LocalOrdinal)ro[i];   // wraps for int at 512^3, as the reference's own int would
    if (sizeof(GlobalOrdinal) == 4 && mfe_export_csr(ctx, (int *)&A.rows[0], 0, 0, 0, 0) != 0) die("mfe_export_csr");
  }
}

// ELLMatrix (ref/src/ELLMatrix.hpp:40-134): generate_matrix_structure reserves 27 entries per row
// (generate_matrix_structure.hpp:81); num_nonzeros() = rows * 27.  The device keeps the compact CSR arrays and is told
// to apply the ELL matvec's padding terms (mfe_set_option "ell_width").
template<typename Scalar, typename LocalOrdinal, typename GlobalOrdinal>
inline void publish_structure(ELLMatrix<Scalar,LocalOrdinal,GlobalOrdinal>& A, mfe_ctx *ctx, int64_t nrows, int64_t first, int64_t /*nnz*/)
{
  A.rows.resize(nrows);
  if (sizeof(GlobalOrdinal) == 4) { if (nrows > 0 && mfe_export_csr(ctx, (int *)&A.rows[0], 0, 0, 0, 0) != 0) die("mfe_export_csr"); }
  else for (int64_t i = 0; i < nrows; ++i) A.rows[i] = (GlobalOrdinal)(first + i);
  A.num_cols_per_row = 27;
  if (mfe_set_option(ctx, "ell_width", 27) != 0) die("mfe_set_option");
}

}  // namespace b200

// ref/src/generate_matrix_structure.hpp:47-149.  Returns 0, or 1 when any rank failed (the reference allreduces
// an exception flag; the driver ignores the value, driver.hpp:181-183).
template<typename MatrixType>
int
generate_matrix_structure(const simple_mesh_description<typename MatrixType::GlobalOrdinalType>& mesh,
                          MatrixType& A)
{
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
    b200::publish_structure(A, e.ctx, nrows, first, nnz);
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
