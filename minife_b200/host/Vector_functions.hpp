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

// ref/src/Vector_functions.hpp:49-82 (same text format: "<row> <value>" per line, one file per rank)
template<typename VectorType>
void write_vector(const std::string& filename,
                  const VectorType& vec)
{
  int numprocs = 1, myproc = 0;
#ifdef HAVE_MPI
  MPI_Comm_size(MPI_COMM_WORLD, &numprocs);
  MPI_Comm_rank(MPI_COMM_WORLD, &myproc);
#endif
  std::vector<typename VectorType::ScalarType> host(vec.coefs);
  b200::Entry *e = b200::entry_of_vector(&vec);
  if (e && !host.empty()) mfe_export_vector(e->ctx, e->b_vec == &vec ? "b" : "x", &host[0]);   // device copy is the truth
  std::ostringstream osstr;
  osstr << filename << "." << numprocs << "." << myproc;
  std::string full_name = osstr.str();
  std::ofstream ofs(full_name.c_str());
  typedef typename VectorType::ScalarType ScalarType;
  const std::vector<ScalarType>& coefs = host;
  for(int p=0; p<numprocs; ++p) {
    if (p == myproc) {
      if (p == 0) {
        ofs << vec.local_size << std::endl;
      }
      typename VectorType::GlobalOrdinalType first = vec.startIndex;
      for(size_t i=0; i<vec.local_size; ++i) {
        ofs << first+i << " " << coefs[i] << std::endl;
      }
    }
#ifdef HAVE_MPI
    MPI_Barrier(MPI_COMM_WORLD);
#endif
  }
}

// ref/src/Vector_functions.hpp:84-103 (host helper, kept for API completeness)
template<typename VectorType>
void sum_into_vector(size_t num_indices,
                     const typename VectorType::GlobalOrdinalType* indices,
                     const typename VectorType::ScalarType* coefs,
                     VectorType& vec)
{
  typename VectorType::GlobalOrdinalType first = vec.startIndex, last = first + vec.local_size - 1;
  for(size_t i=0; i<num_indices; ++i) {
    if (indices[i] < first || indices[i] > last) continue;
    vec.coefs[indices[i] - first] += coefs[i];
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
