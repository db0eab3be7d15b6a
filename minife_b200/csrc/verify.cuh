// verify.cuh -- ref/fem/verify_solution.hpp:55-174 + ref/fem/analytic_soln.hpp:69-118 on the device-resident x.
// The analytic solution is a double series (301 x 301 terms of sinh/sin per point); one CTA per verified point,
// thread t owns p = t, t + 128, t + 256 and walks q in the reference's order with the reference's NaN rule.
// The tables l(p,q) and sinh(l) do not depend on the point and are built once.  CUDA's sin/sinh differ from glibc's
// by an ulp or two, so the analytic value (and the reported max error) agree with the reference to ~1e-15, not
// bit for bit; verify_solution is a tolerance test (0.06 in the driver), its verdict is identical.
#pragma once
#include "kernels.cuh"

namespace mfe {

enum { VERIFY_TERMS = 300, VERIFY_NT = VERIFY_TERMS + 1, VERIFY_THREADS = 128 };

__global__ void k_verify_tables(double *__restrict__ l_tab, double *__restrict__ sinh_tab)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= VERIFY_NT * VERIFY_NT) return;
  const int p = i / VERIFY_NT, q = i % VERIFY_NT;
  const double PI = 3.141592653589793238462;
  const double PI_SQR = __dmul_rn(PI, PI);
  // analytic_soln.hpp:71-74  sqrt((2p+1)(2p+1) PI^2 + (2q+1)(2q+1) PI^2), integer products first
  const double l = sqrt(__dadd_rn(__dmul_rn((double)((2 * p + 1) * (2 * p + 1)), PI_SQR), __dmul_rn((double)((2 * q + 1) * (2 * q + 1)), PI_SQR)));
  l_tab[i] = l;
  sinh_tab[i] = sinh(l);
}

struct VerifyAxes { int n[3]; };   // number of verified indices per axis; the verified points are their product

// err_out[i] = |analytic - computed| of verified point i (x fastest), analytic_out[i] likewise
__global__ void __launch_bounds__(VERIFY_THREADS) k_verify(MeshDesc m, VerifyAxes va, const int *__restrict__ vix, const int *__restrict__ viy,
                                                           const int *__restrict__ viz, const double *__restrict__ xsol,
                                                           const double *__restrict__ l_tab, const double *__restrict__ sinh_tab,
                                                           double *__restrict__ err_out, double *__restrict__ analytic_out)
{
  __shared__ double s_sin_qz[VERIFY_NT];
  __shared__ double s_part[VERIFY_THREADS];
  const int64_t pt = blockIdx.x;
  const int ix = vix[pt % va.n[0]], iy = viy[(pt / va.n[0]) % va.n[1]], iz = viz[pt / ((int64_t)va.n[0] * va.n[1])];
  // get_coords (ref/utils/box_utils.hpp:68-82) with the global NODE counts
  const int xdiv = m.gnx > 1 ? m.gnx - 1 : 1, ydiv = m.gny > 1 ? m.gny - 1 : 1, zdiv = m.gnz > 1 ? m.gnz - 1 : 1;
  const double x = __ddiv_rn((double)ix, (double)xdiv), y = __ddiv_rn((double)iy, (double)ydiv), z = __ddiv_rn((double)iz, (double)zdiv);
  const int64_t lr = (int64_t)(ix - m.nb[0]) + (int64_t)m.lnx * ((iy - m.nb[2]) + (int64_t)m.lny * (iz - m.nb[4]));
  const double PI = 3.141592653589793238462;
  double analytic;
  bool series = false;
  if (x == 1.0) analytic = 1.0;                                   // verify_solution.hpp:118-131
  else if (x == 0.0 || y == 0.0 || z == 0.0) analytic = 0.0;
  else if (y == 1.0 || z == 1.0) analytic = 0.0;
  else { analytic = 0.0; series = true; }
  if (series) {                                                   // uniform per CTA
    for (int q = threadIdx.x; q < VERIFY_NT; q += VERIFY_THREADS)
      s_sin_qz[q] = __ddiv_rn(sin(__dmul_rn(__dmul_rn((double)(2 * q + 1), PI), z)), (double)(2 * q + 1));
    __syncthreads();
    double sum = 0.0;
    for (int p = threadIdx.x; p < VERIFY_NT; p += VERIFY_THREADS) {
      const double sin_py = __ddiv_rn(sin(__dmul_rn(__dmul_rn((double)(2 * p + 1), PI), y)), (double)(2 * p + 1));
      const double *lt = l_tab + p * VERIFY_NT, *st = sinh_tab + p * VERIFY_NT;
      for (int q = 0; q < VERIFY_NT; ++q) {
        const double sinh1 = sinh(__dmul_rn(lt[q], x));
        const double tmp = __dmul_rn(__dmul_rn(sinh1, sin_py), __ddiv_rn(s_sin_qz[q], st[q]));
        if (tmp == tmp) sum = __dadd_rn(sum, tmp); else break;    // analytic_soln.hpp:98-108
      }
    }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < VERIFY_THREADS; ++i) t = __dadd_rn(t, s_part[i]);
      const double PI_SQR = __dmul_rn(PI, PI);
      analytic = __dmul_rn(__ddiv_rn(16.0, PI_SQR), t);
    }
  }
  if (threadIdx.x == 0) {
    err_out[pt] = fabs(__dadd_rn(analytic, -xsol[lr]));
    analytic_out[pt] = analytic;
  }
}

}  // namespace mfe
