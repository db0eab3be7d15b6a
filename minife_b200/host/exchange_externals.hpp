// Shadows ref/src/exchange_externals.hpp -> mfe_exchange_externals (pack kernel + grouped NCCL send/recv).
#ifndef _exchange_externals_hpp_
#define _exchange_externals_hpp_

#include <b200_bridge.hpp>

namespace miniFE {

// ref/src/exchange_externals.hpp:44-161 on a host vector of num_cols entries
template<typename MatrixType,
         typename VectorType>
void
exchange_externals(MatrixType& A,
                   VectorType& x)
{
  b200::Entry &e = b200::entry_of(&A, "exchange_externals");
  if (mfe_exchange_externals(e.ctx, &x.coefs[0]) != 0) b200::die("mfe_exchange_externals");
}

// ref/src/exchange_externals.hpp:167-264 (begin/finish split used by matvec_overlap): the device path overlaps
// internally, so begin does the whole exchange and finish is a no-op
template<typename MatrixType, typename VectorType>
void begin_exchange_externals(MatrixType& A, VectorType& x) { exchange_externals(A, x); }
inline void finish_exchange_externals(int /*num_neighbors*/) {}

}//namespace miniFE

#endif
