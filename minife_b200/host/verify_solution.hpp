// Shadows ref/fem/verify_solution.hpp -> mfe_verify_solution: the 301x301-term analytic series is evaluated on the
// device against the device-resident x of the last cg_solve (no host pass over the solution).
#ifndef _verify_solution_hpp_
#define _verify_solution_hpp_

#include <iostream>

#include <b200_bridge.hpp>
#include <simple_mesh_description.hpp>

namespace miniFE {

// ref/fem/verify_solution.hpp:47-53
template<typename Scalar>
struct err_info {
  Scalar err;
  Scalar computed;
  Scalar analytic;
  Scalar coords[3];
};

// ref/fem/verify_solution.hpp:55-174: same signature, same messages, same return code
template<typename VectorType>
int
verify_solution(const simple_mesh_description<typename VectorType::GlobalOrdinalType>& /*mesh*/,
                const VectorType& x, double tolerance, bool verify_whole_domain = false)
{
  b200::Entry *e = b200::entry_of_vector(&x);
  if (!e) {
    std::cerr << "minife_b200: verify_solution: x was not produced by cg_solve on the device" << std::endl;
    std::exit(1);
  }
  // x.coefs is the caller's copy of the device x (cg_solve copies it back); a caller that changed it afterwards
  // gets the changed values verified, as in the reference
  if (!x.coefs.empty() && mfe_import_vector(e->ctx, "x", &x.coefs[0]) != 0) b200::die("mfe_import_vector");
  int return_code = 0;
  double global_max_err = 0, info[6] = {0, 0, 0, 0, 0, 0};
  if (mfe_verify_solution(e->ctx, tolerance, verify_whole_domain ? 1 : 0, &return_code, &global_max_err, info) != 0)
    b200::die("mfe_verify_solution");
  if (info[0] == global_max_err) {     // verify_solution.hpp:156-166
    if (info[0] > tolerance) {
      std::cout << "max absolute error is "<<info[0]<<":"<<std::endl;
      std::cout << "   at position ("<<info[3]<<","<<info[4]<<","<<info[5]<<"), "<<std::endl;
      std::cout << "   computed solution: "<<info[1]<<",  analytic solution: "<<info[2]<<std::endl;
    }
    else {
      std::cout << "solution matches analytic solution to within "<<tolerance<<" or better."<<std::endl;
    }
  }
  return return_code;
}

}//namespace miniFE

#endif
