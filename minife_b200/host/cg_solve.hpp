// Shadows ref/src/cg_solve.hpp -> mfe_cg_solve (whole iteration on the device, scalars included).
#ifndef _cg_solve_hpp_
#define _cg_solve_hpp_

#include <cmath>
#include <limits>

#include <b200_bridge.hpp>
#include <Vector_functions.hpp>
#include <mytimer.hpp>
#include <outstream.hpp>

namespace miniFE {

// ref/src/cg_solve.hpp:49-67
template<typename VectorType>
bool breakdown(typename VectorType::ScalarType inner,
               const VectorType& v,
               const VectorType& w)
{
  typedef typename TypeTraits<typename VectorType::ScalarType>::magnitude_type magnitude;
  magnitude vnorm = std::sqrt(dot(v,v));
  magnitude wnorm = std::sqrt(dot(w,w));
  return std::abs(inner) <= 100*vnorm*wnorm*std::numeric_limits<magnitude>::epsilon();
}

// ref/src/cg_solve.hpp:69-229.  `matvec` (matvec_std or matvec_overlap) only selects the reference's code path; on
// the device both are the same SpMV.  x.coefs: initial guess in, solution out.  Prints what the reference prints.
template<typename OperatorType,
         typename VectorType,
         typename Matvec>
void
cg_solve(OperatorType& A,
         const VectorType& b,
         VectorType& x,
         Matvec /*matvec*/,
         typename OperatorType::LocalOrdinalType max_iter,
         typename TypeTraits<typename OperatorType::ScalarType>::magnitude_type& tolerance,
         typename OperatorType::LocalOrdinalType& num_iters,
         typename TypeTraits<typename OperatorType::ScalarType>::magnitude_type& normr,
         timer_type* my_cg_times)
{
  int myproc = 0;
#ifdef HAVE_MPI
  MPI_Comm_rank(MPI_COMM_WORLD, &myproc);
#endif
  if (!A.has_local_indices) {
    std::cerr << "miniFE::cg_solve ERROR, A.has_local_indices is false, needs to be true. This probably means "
       << "miniFE::make_local_matrix(A) was not called prior to calling miniFE::cg_solve."
       << std::endl;
    return;
  }
  b200::Entry &e = b200::entry_of(&A, "cg_solve");
  if (e.b_vec != &b) {      // a right-hand side that was not assembled on the device: upload it
    if (!b.coefs.empty() && mfe_import_vector(e.ctx, "b", &b.coefs[0]) != 0) b200::die("mfe_import_vector");
  }
  e.x_vec = &x;
#ifdef MINIFE_FUSED
  // cg_solve.hpp:172-176, 203-206: matvec_and_dot / fused_waxpby -- same arithmetic on the device, reported as MATVECDOT
  if (mfe_set_option(e.ctx, "fused", 1) != 0) b200::die("mfe_set_option");
#endif
  std::vector<double> hist(max_iter + 1);
  double times[MFE_NUM_TIMERS] = {0, 0, 0, 0, 0};
  int iters = 0;
  double rnorm = 0;
  int rc = mfe_cg_solve(e.ctx, x.coefs.empty() ? 0 : &x.coefs[0], max_iter, tolerance, &iters, &rnorm, times, &hist[0], 0, 0, 0);
  if (rc == 1) b200::die("mfe_cg_solve");
  if (myproc == 0) {
    typename OperatorType::LocalOrdinalType print_freq = max_iter/10;
    if (print_freq>50) print_freq = 50;
    if (print_freq<1)  print_freq = 1;
    std::cout << "Initial Residual = "<< hist[0] << std::endl;
    for(int k=1; k<=iters + (rc == 2 ? 1 : 0) && k<=max_iter; ++k) {
      if (k%print_freq==0 || k==max_iter) std::cout << "Iteration = "<<k<<"   Residual = "<<hist[k]<<std::endl;
    }
  }
  num_iters = iters;
  normr = rnorm;
  my_cg_times[WAXPY] = times[MFE_WAXPY];
  my_cg_times[DOT] = times[MFE_DOT];
  my_cg_times[MATVEC] = times[MFE_MATVEC];
  my_cg_times[MATVECDOT] = times[MFE_MATVECDOT];
  my_cg_times[TOTAL] = times[MFE_TOTAL];
}

}//namespace miniFE

#endif
