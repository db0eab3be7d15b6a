// Shadows ref/src/Vector_functions.hpp: waxpby, fused_waxpby, dot, write_vector, sum_into_vector.
#ifndef _Vector_functions_hpp_
#define _Vector_functions_hpp_

#include <vector>
#include <sstream>
#include <fstream>
#include <b200_bridge.hpp>
#include <TypeTraits.hpp>
#include <Vector.hpp>

namespace miniFE {

namespace b200 {
// waxpby / dot are called on free-standing Vectors; any live context can run them (they do not touch its matrix)
inline mfe_ctx *any_context()
{
  for (std::map<const void *, Entry>::iterator it = table().begin(); it != table().end(); ++it)
    if (it->second.ctx) return it->second.ctx;
  std::cerr << "minife_b200: waxpby/dot need a device context (call generate_matrix_structure first)" << std::endl;
  std::exit(1);
}
}

// ref/src/Vector_functions.hpp:49-82: "<local_size>" (rank 0 only), then "<global row> <value>" per entry.  The values
// of the assembled b and of the solved x live on the device; they are fetched here.
template<typename VectorType>
void write_vector(const std::string& filename,
                  const VectorType& vec)
{
  typedef typename VectorType::ScalarType Scalar;
  std::vector<Scalar> values(vec.coefs);
  if (b200::Entry *e = b200::entry_of_vector(&vec)) {
    if (!values.empty()) mfe_export_vector(e->ctx, e->b_vec == &vec ? "b" : "x", &values[0]);
  }
  struct Lines {
    const VectorType *v; const std::vector<Scalar> *vals;
    void operator()(std::ostream &os, bool is_rank0) const
    {
      if (is_rank0) os << v->local_size << std::endl;
      size_t row = v->startIndex;
      for (size_t k = 0; k < (size_t)v->local_size; ++k, ++row) os << row << " " << (*vals)[k] << std::endl;
    }
  } lines = { &vec, &values };
  b200::RankFile(filename).in_rank_order(lines);
}

// ref/src/Vector_functions.hpp:84-103 (host helper, kept for API completeness)
template<typename VectorType>
void sum_into_vector(size_t num_indices,
                     const typename VectorType::GlobalOrdinalType* indices,
                     const typename VectorType::ScalarType* coefs,
                     VectorType& vec)
{
  typedef typename VectorType::GlobalOrdinalType GO;
  const GO lo = vec.startIndex, n = (GO)vec.local_size;
  for (size_t k = 0; k < num_indices; ++k) {
    const GO local = indices[k] - lo;
    if (local >= 0 && local < n) vec.coefs[local] += coefs[k];     // entries of other ranks are ignored
  }
}

// ref/src/Vector_functions.hpp:127-150: w = alpha*x + beta*y over x.coefs.size() entries
template<typename VectorType>
void
  waxpby(typename VectorType::ScalarType alpha, const VectorType& x,
         typename VectorType::ScalarType beta, const VectorType& y,
         VectorType& w)
{
  if (x.coefs.empty()) return;
  if (mfe_waxpby(b200::any_context(), (int64_t)x.coefs.size(), alpha, &x.coefs[0], beta, &y.coefs[0], &w.coefs[0]) != 0) b200::die("mfe_waxpby");
}

// ref/src/Vector_functions.hpp:154-185
template<typename VectorType>
void
  fused_waxpby(typename VectorType::ScalarType alpha, const VectorType& x,
         typename VectorType::ScalarType beta, const VectorType& y,
         VectorType& w,
         typename VectorType::ScalarType alpha2, const VectorType& x2,
         typename VectorType::ScalarType beta2, const VectorType& y2,
         VectorType& w2)
{
  waxpby(alpha, x, beta, y, w);
  waxpby(alpha2, x2, beta2, y2, w2);
}

// ref/src/Vector_functions.hpp:194-226: serial-order sum (bit-identical) + rank-ordered global sum
template<typename Vector>
typename TypeTraits<typename Vector::ScalarType>::magnitude_type
  dot(const Vector& x,
      const Vector& y)
{
  double result = 0;
  const double zero = 0;
  if (mfe_dot(b200::any_context(), (int64_t)x.coefs.size(), x.coefs.empty() ? &zero : &x.coefs[0], y.coefs.empty() ? &zero : &y.coefs[0], &result) != 0)
    b200::die("mfe_dot");
  return result;
}

}//namespace miniFE

#endif
