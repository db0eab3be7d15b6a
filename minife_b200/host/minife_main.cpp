// minife_main.cpp -- standalone single-process driver over the C ABI (no reference sources needed):
//   minife_b200_standalone.x nx=256 [ny=.. nz=..] [max_iters=200] [verify=1]
// Runs the reference's phase sequence (ref/src/driver.hpp:159-289) on GPU 0 and prints a YAML-style report with the
// reference's own flop model (driver.hpp:322-341).  For multi-GPU runs use bench.py (torch.distributed launcher).
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <minife_b200.h>

static int arg_int(int argc, char **argv, const char *name, int dflt)
{
  const size_t l = strlen(name);
  for (int i = 1; i < argc; ++i)
    if (!strncmp(argv[i], name, l) && argv[i][l] == '=') return atoi(argv[i] + l + 1);
  return dflt;
}
#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, mfe_last_error()); return 1; } } while (0)

int main(int argc, char **argv)
{
  const int nx = arg_int(argc, argv, "nx", 10), ny = arg_int(argc, argv, "ny", nx), nz = arg_int(argc, argv, "nz", ny);
  const int max_iters = arg_int(argc, argv, "max_iters", 200);
  int gbox[6] = { 0, nx, 0, ny, 0, nz }, boxes[6];
  CHECK(mfe_box_partition(nx, ny, nz, 1, boxes));
  mfe_ctx *ctx = mfe_create(0, 1, 0, gbox, boxes);
  if (!ctx) { fprintf(stderr, "mfe_create failed: %s\n", mfe_last_error()); return 1; }
  int64_t nrows = 0, first = 0, nnz = 0, ncols = 0;
  CHECK(mfe_generate_matrix_structure(ctx, &nrows, &first, &nnz));
  CHECK(mfe_assemble_FE_data(ctx));
  CHECK(mfe_impose_dirichlet(ctx, 0.0, 0));
  CHECK(mfe_impose_dirichlet(ctx, 1.0, 1));
  CHECK(mfe_make_local_matrix(ctx, &ncols));
  std::vector<double> x(nrows, 0.0), hist(max_iters + 1);
  int iters = 0; double normr = 0, t[MFE_NUM_TIMERS];
  double tol = DBL_EPSILON;
  int rc = mfe_cg_solve(ctx, x.data(), max_iters, tol, &iters, &normr, t, hist.data(), 0, 0, 0);
  if (rc == 1) { fprintf(stderr, "mfe_cg_solve failed: %s\n", mfe_last_error()); return 1; }
  const double n = (double)nx * ny * nz;
  const double mv = 2.0 * nnz * (iters + 1), dt = 2.0 * n * 2 * iters, wx = 3.0 * n * (3 * iters + 2), tot = mv + dt + wx;
  printf("Mini-Application Name: miniFE (minife_b200, sm_100a)\n");
  printf("Global Run Parameters:\n  dimensions:\n    nx: %d\n    ny: %d\n    nz: %d\n", nx, ny, nz);
  printf("Matrix structure generation:\n  Mat-struc-gen Time: %g\n", mfe_phase_ms(ctx, "structure") * 1e-3);
  printf("FE assembly:\n  FE assembly Time: %g\n", mfe_phase_ms(ctx, "assembly") * 1e-3);
  printf("Dirichlet BC:\n  Time: %g\n", (mfe_phase_ms(ctx, "dirichlet0") + mfe_phase_ms(ctx, "dirichlet1")) * 1e-3);
  printf("Matrix attributes:\n  Global Nrows: %lld\n  Global NNZ: %lld\n", (long long)nrows, (long long)nnz);
  printf("CG solve:\n  Iterations: %d\n  Final Resid Norm: %g\n", iters, normr);
  printf("  WAXPY Time: %g\n  WAXPY Flops: %g\n  WAXPY Mflops: %g\n", t[MFE_WAXPY], wx, 1e-6 * wx / t[MFE_WAXPY]);
  printf("  DOT Time: %g\n  DOT Flops: %g\n  DOT Mflops: %g\n", t[MFE_DOT], dt, 1e-6 * dt / t[MFE_DOT]);
  printf("  MATVEC Time: %g\n  MATVEC Flops: %g\n  MATVEC Mflops: %g\n", t[MFE_MATVEC], mv, 1e-6 * mv / t[MFE_MATVEC]);
  printf("  Total:\n    Total CG Time: %g\n    Total CG Flops: %g\n    Total CG Mflops: %g\n  Time per iteration: %g\n", t[MFE_TOTAL], tot,
         1e-6 * tot / t[MFE_TOTAL], t[MFE_TOTAL] / (iters > 0 ? iters : 1));
  if (rc == 2) printf("  numerical breakdown: yes\n");
  mfe_destroy(ctx);
  return 0;
}
