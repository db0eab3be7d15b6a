// Shadows ref/src/assemble_FE_data.hpp -> mfe_assemble_FE_data (element loop + scatter on the device).
#ifndef _assemble_FE_data_hpp_
#define _assemble_FE_data_hpp_

#include <b200_bridge.hpp>
#include <simple_mesh_description.hpp>
#include <Parameters.hpp>

namespace miniFE {

// ref/src/assemble_FE_data.hpp:38-73.  b stays device resident; its host coefs are refreshed so that code reading
// b.coefs (write_vector under MINIFE_DEBUG) sees the assembled values.
template<typename MatrixType,
         typename VectorType>
void
assemble_FE_data(const simple_mesh_description<typename MatrixType::GlobalOrdinalType>& mesh,
                 MatrixType& A,
                 VectorType& b,
                 Parameters& /*params*/)
{
  b200::Entry &e = b200::entry_of(&A, "assemble_FE_data");
  if (mfe_assemble_FE_data(e.ctx) != 0) b200::die("mfe_assemble_FE_data");
  e.b_vec = &b;
#ifdef MINIFE_DEBUG
  if (!b.coefs.empty() && mfe_export_vector(e.ctx, "b", &b.coefs[0]) != 0) b200::die("mfe_export_vector");
#endif
}

}//namespace miniFE

#endif
