// Shadows ref/src/make_local_matrix.hpp -> mfe_make_local_matrix (index maps without messages + device rewrite).
#ifndef _make_local_matrix_hpp_
#define _make_local_matrix_hpp_

#include <b200_bridge.hpp>

namespace miniFE {

namespace b200 {
template<typename T>
inline void fetch_map(mfe_ctx *ctx, const char *name, std::vector<T> &out)
{
  int64_t n = mfe_get_map(ctx, name, 0);
  std::vector<int> tmp(n > 0 ? n : 0);
  if (n > 0) mfe_get_map(ctx, name, &tmp[0]);
  out.assign(tmp.begin(), tmp.end());
}
}

// ref/src/make_local_matrix.hpp:42-442.  Fills the reference's index-map fields (CSRMatrix.hpp:73-80) so that
// compute_matrix_stats / debugging code keep working.
template<typename MatrixType>
void
make_local_matrix(MatrixType& A)
{
  b200::Entry &e = b200::entry_of(&A, "make_local_matrix");
  int64_t num_cols = 0;
  if (mfe_make_local_matrix(e.ctx, &num_cols) != 0) b200::die("mfe_make_local_matrix");
  A.num_cols = (typename MatrixType::LocalOrdinalType)num_cols;
#ifdef HAVE_MPI
  b200::fetch_map(e.ctx, "external_index", A.external_index);
  b200::fetch_map(e.ctx, "external_local_index", A.external_local_index);
  b200::fetch_map(e.ctx, "elements_to_send", A.elements_to_send);
  b200::fetch_map(e.ctx, "neighbors", A.neighbors);
  b200::fetch_map(e.ctx, "recv_length", A.recv_length);
  b200::fetch_map(e.ctx, "send_length", A.send_length);
  A.send_buffer.assign(A.elements_to_send.size(), 0);
  A.request.resize(A.neighbors.size());
#endif
  A.has_local_indices = true;
}

}//namespace miniFE

#endif
