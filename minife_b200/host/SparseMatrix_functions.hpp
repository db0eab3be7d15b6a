// Shadows ref/src/SparseMatrix_functions.hpp: the operators the driver calls on the assembled matrix --
// impose_dirichlet, matvec_std / matvec_overlap / matvec, rearrange_matrix_local_external, write_matrix.
#ifndef _SparseMatrix_functions_hpp_
#define _SparseMatrix_functions_hpp_

#include <cstddef>
#include <vector>
#include <set>
#include <sstream>
#include <fstream>

#include <b200_bridge.hpp>
#include <Vector.hpp>
#include <Vector_functions.hpp>
#include <exchange_externals.hpp>
#include <mytimer.hpp>

namespace miniFE {

// ref/src/SparseMatrix_functions.hpp:98-140: "<nrows> <nnz>" (rank 0 only), then "<row> <col> <value>" per stored entry,
// rows in order.  The matrix lives on the device: its CSR arrays are exported here, on demand.
template<typename MatrixType>
void write_matrix(const std::string& filename,
                  MatrixType& mat)
{
  typedef typename MatrixType::GlobalOrdinalType GlobalOrdinalType;
  typedef typename MatrixType::ScalarType ScalarType;
  b200::Entry &e = b200::entry_of(&mat, "write_matrix");
  const int64_t nrows = mfe_nrows(e.ctx), nnz = mfe_nnz(e.ctx);
  std::vector<int64_t> begin(nrows + 1);
  std::vector<int> cols(nnz);
  std::vector<double> coefs(nnz);
  if (mfe_export_csr(e.ctx, 0, 0, &begin[0], nnz ? &cols[0] : 0, nnz ? &coefs[0] : 0) != 0) b200::die("mfe_export_csr");
  struct Lines {
    MatrixType *A; const std::vector<int64_t> *begin; const std::vector<int> *cols; const std::vector<double> *coefs;
    void operator()(std::ostream &os, bool is_rank0) const
    {
      // CSRMatrix: num_nonzeros() = row_offsets[nrows]; ELLMatrix: rows * num_cols_per_row (padding counted)
      if (is_rank0) os << A->rows.size() << " " << (size_t)A->num_nonzeros() << std::endl;
      for (size_t r = 0; r < A->rows.size(); ++r)
        for (int64_t k = (*begin)[r]; k < (*begin)[r + 1]; ++k)
          os << A->rows[r] << " " << (GlobalOrdinalType)(*cols)[k] << " " << (ScalarType)(*coefs)[k] << std::endl;
    }
  } lines = { &mat, &begin, &cols, &coefs };
  b200::RankFile(filename).in_rank_order(lines);
}

// ref/src/SparseMatrix_functions.hpp:295-317 (reporting helper used by compute_matrix_stats): size of the index maps
template<typename MatrixType>
double
parallel_memory_overhead_MB(const MatrixType& A)
{
  typedef typename MatrixType::GlobalOrdinalType GlobalOrdinal;
  typedef typename MatrixType::LocalOrdinalType LocalOrdinal;
  double mem_MB = 0;
#ifdef HAVE_MPI
  double invMB = 1.0/(1024*1024);
  mem_MB = invMB*(A.external_index.size() + A.external_local_index.size() + A.elements_to_send.size())*sizeof(GlobalOrdinal);
  mem_MB += invMB*A.neighbors.size()*sizeof(int);
  mem_MB += invMB*(A.recv_length.size() + A.send_length.size())*sizeof(LocalOrdinal);
  double tmp = mem_MB;
  MPI_Allreduce(&tmp, &mem_MB, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
#endif
  return mem_MB;
}

// ref/src/SparseMatrix_functions.hpp:319-408: on the device every row is re-sorted in place (owned columns first, externals
// after them), which makes the SpMV sum the rows in matvec_overlap's order (mfe_rearrange_matrix_local_external); the host
// struct keeps the arrays the driver has already reported and dumped.
template<typename MatrixType>
void rearrange_matrix_local_external(MatrixType& A)
{
  if (mfe_rearrange_matrix_local_external(b200::entry_of(&A, "rearrange_matrix_local_external").ctx) != 0) b200::die("rearrange_matrix_local_external");
}

// ref/src/SparseMatrix_functions.hpp:431-480.  `bc_rows` must be one of the two sets of the mesh the matrix was
// generated from (that is what driver.hpp:215-218 passes); on the device the sets are closed-form.
template<typename MatrixType,
         typename VectorType>
void
impose_dirichlet(typename MatrixType::ScalarType prescribed_value,
                    MatrixType& A,
                    VectorType& b,
                    int /*global_nx*/,
                    int /*global_ny*/,
                    int /*global_nz*/,
                    const std::set<typename MatrixType::GlobalOrdinalType>& bc_rows)
{
  b200::Entry &e = b200::entry_of(&A, "impose_dirichlet");
  int which = -1;
  if ((const void *)&bc_rows == e.bc_rows_0) which = 0;
  else if ((const void *)&bc_rows == e.bc_rows_1) which = 1;
  if (which < 0) {
    std::cerr << "minife_b200::impose_dirichlet: bc_rows must be mesh.bc_rows_0 or mesh.bc_rows_1 of the generating mesh" << std::endl;
    std::exit(1);
  }
  if (mfe_impose_dirichlet(e.ctx, prescribed_value, which) != 0) b200::die("mfe_impose_dirichlet");
  e.b_vec = &b;
#ifdef MINIFE_DEBUG
  if (!b.coefs.empty() && mfe_export_vector(e.ctx, "b", &b.coefs[0]) != 0) b200::die("mfe_export_vector");
#endif
}

static timer_type exchtime = 0;

// ref/src/SparseMatrix_functions.hpp:546-578: y = A*x including exchange_externals, on host vectors
template<typename MatrixType,
         typename VectorType>
struct matvec_std {
void operator()(MatrixType& A,
            VectorType& x,
            VectorType& y)
{
  b200::Entry &e = b200::entry_of(&A, "matvec_std");
  if (mfe_matvec(e.ctx, &x.coefs[0], &y.coefs[0]) != 0) b200::die("mfe_matvec");
}
};

// ref/src/SparseMatrix_functions.hpp:616-622
template<typename MatrixType,
         typename VectorType>
void matvec(MatrixType& A, VectorType& x, VectorType& y)
{
  matvec_std<MatrixType,VectorType> mv;
  mv(A, x, y);
}

// ref/src/SparseMatrix_functions.hpp:624-675
template<typename MatrixType,
         typename VectorType>
struct matvec_overlap {
void operator()(MatrixType& A,
                    VectorType& x,
                    VectorType& y)
{
  matvec_std<MatrixType,VectorType> mv;
  mv(A, x, y);
}
};

// ref/src/SparseMatrix_functions.hpp:491-536 (MINIFE_FUSED)
template<typename MatrixType,
         typename VectorType>
typename TypeTraits<typename MatrixType::ScalarType>::magnitude_type
matvec_and_dot(MatrixType& A,
               VectorType& x,
               VectorType& y)
{
  matvec(A, x, y);
  VectorType xr(x.startIndex, y.local_size);
  for (size_t i = 0; i < (size_t)y.local_size; ++i) xr.coefs[i] = x.coefs[i];
  return dot(xr, y);
}

}//namespace miniFE

#endif
