// Shadows ref/src/time_kernels.hpp (MINIFE_KERNELS != 0, driver.hpp:257-270) -> mfe_time_kernels: the three timing
// loops run on the device-resident vectors.
#ifndef _time_kernels_hpp_
#define _time_kernels_hpp_

#include <iostream>

#include <b200_bridge.hpp>
#include <mytimer.hpp>

namespace miniFE {

template<typename OperatorType,
         typename VectorType,
         typename Matvec>
void
time_kernels(OperatorType& A,
             const VectorType& b,
             VectorType& x,
             Matvec /*matvec*/,
             typename OperatorType::LocalOrdinalType max_iter,
             typename OperatorType::ScalarType& xdotp,
             timer_type* my_kern_times)
{
  if (!A.has_local_indices) {
    std::cerr << "miniFE::time_kernels ERROR, A.has_local_indices is false, needs to be true. This probably means "
       << "miniFE::make_local_matrix(A) was not called prior to calling miniFE::time_kernels."
       << std::endl;
    return;
  }
  b200::Entry &e = b200::entry_of(&A, "time_kernels");
  (void)b;
  e.x_vec = &x;
  double times[MFE_NUM_TIMERS] = {0, 0, 0, 0, 0};
  double r = 0;
  if (mfe_time_kernels(e.ctx, x.coefs.empty() ? 0 : &x.coefs[0], max_iter, &r, times) != 0) b200::die("mfe_time_kernels");
  xdotp = r;
  my_kern_times[WAXPY] = times[MFE_WAXPY];
  my_kern_times[DOT] = times[MFE_DOT];
  my_kern_times[MATVEC] = times[MFE_MATVEC];
  my_kern_times[TOTAL] = 0;
}

}//namespace miniFE

#endif
