// hex8.h -- hex8 element diffusion matrix + source vector with the reference's operation order, one IEEE
// rounding per operation (no FMA contraction): ref/fem/Hex8.hpp:43-113 (shape fns / gradients), 177-257
// (Jacobian, inverse, determinant), 259-361 (diffusionMatrix_symm), 363-410 (sourceVector),
// ref/fem/gauss_pts.hpp:37-62, ref/fem/Hex8_ElemData.hpp:37-61, ref/src/Hex8_box_utils.hpp:59-99.
// Bit-exactness against the reference matters: 6 of the 27 stencil entries per row are ~5e-11 "noise"
// produced by the truncated Gauss point literal, and they change under FMA contraction.
#pragma once
#include "mesh.h"

#ifdef __CUDA_ARCH__
#define MFE_MUL(a, b) __dmul_rn((a), (b))
#define MFE_ADD(a, b) __dadd_rn((a), (b))
#define MFE_SUB(a, b) __dsub_rn((a), (b))
#define MFE_DIV(a, b) __ddiv_rn((a), (b))
#else
// host build: compiled without -march / -ffast-math, -ffp-contract=off
#define MFE_MUL(a, b) ((a) * (b))
#define MFE_ADD(a, b) ((a) + (b))
#define MFE_SUB(a, b) ((a) - (b))
#define MFE_DIV(a, b) ((a) / (b))
#endif

namespace mfe {

#define MFE_GAUSS_X2 0.577350269   /* gauss_pts.hpp:41, deliberately the truncated literal */

// Hex8::gradients (Hex8.hpp:67-113) at point x[3] -> g[24]
MFE_HD void hex8_gradients(const double *x, double *g)
{
  const double u = MFE_SUB(1.0, x[0]), v = MFE_SUB(1.0, x[1]), w = MFE_SUB(1.0, x[2]);
  const double up1 = MFE_ADD(1.0, x[0]), vp1 = MFE_ADD(1.0, x[1]), wp1 = MFE_ADD(1.0, x[2]);
#define G3(c, a, b) MFE_MUL(MFE_MUL((c), (a)), (b))
  g[0] = G3(-0.125, v, w);     g[1] = G3(-0.125, u, w);      g[2] = G3(-0.125, u, v);
  g[3] = G3(0.125, v, w);      g[4] = G3(-0.125, up1, w);    g[5] = G3(-0.125, up1, v);
  g[6] = G3(0.125, vp1, w);    g[7] = G3(0.125, up1, w);     g[8] = G3(-0.125, up1, vp1);
  g[9] = G3(-0.125, vp1, w);   g[10] = G3(0.125, u, w);      g[11] = G3(-0.125, u, vp1);
  g[12] = G3(-0.125, v, wp1);  g[13] = G3(-0.125, u, wp1);   g[14] = G3(0.125, u, v);
  g[15] = G3(0.125, v, wp1);   g[16] = G3(-0.125, up1, wp1); g[17] = G3(0.125, up1, v);
  g[18] = G3(0.125, vp1, wp1); g[19] = G3(0.125, up1, wp1);  g[20] = G3(0.125, up1, vp1);
  g[21] = G3(-0.125, vp1, wp1); g[22] = G3(0.125, u, wp1);   g[23] = G3(0.125, u, vp1);
#undef G3
}

// Hex8::shape_fns (Hex8.hpp:43-65): 0.125 * a * b * c, left to right
MFE_HD void hex8_shape_fns(const double *x, double *p)
{
  const double u = MFE_SUB(1.0, x[0]), v = MFE_SUB(1.0, x[1]), w = MFE_SUB(1.0, x[2]);
  const double up1 = MFE_ADD(1.0, x[0]), vp1 = MFE_ADD(1.0, x[1]), wp1 = MFE_ADD(1.0, x[2]);
#define S4(a, b, c) MFE_MUL(MFE_MUL(MFE_MUL(0.125, (a)), (b)), (c))
  p[0] = S4(u, v, w);     p[1] = S4(up1, v, w);     p[2] = S4(up1, vp1, w);   p[3] = S4(u, vp1, w);
  p[4] = S4(u, v, wp1);   p[5] = S4(up1, v, wp1);   p[6] = S4(up1, vp1, wp1); p[7] = S4(u, vp1, wp1);
#undef S4
}

// compute_gradient_values (Hex8_ElemData.hpp:37-61): 8 Gauss points (ig,jg,kg) x 24 gradients;
// psi[64]: shape functions at the same points (sourceVector, Hex8.hpp:386-396)
MFE_HD void hex8_tables(double *grad192, double *psi64)
{
  const double gp[2] = { -MFE_GAUSS_X2, MFE_GAUSS_X2 };
  int o = 0;
  for (int ig = 0; ig < 2; ++ig)
    for (int jg = 0; jg < 2; ++jg)
      for (int kg = 0; kg < 2; ++kg) {
        double pt[3] = { gp[ig], gp[jg], gp[kg] };
        hex8_gradients(pt, grad192 + 24 * o);
        hex8_shape_fns(pt, psi64 + 8 * o);
        ++o;
      }
}

// The Jacobian of a box element separates by axis: J[a][b] = sum_i grad_i[a] * coord_i[b] (Hex8.hpp:205-232) reads the
// b-coordinates of the eight nodes only, and those take just two values, c0 (low face) and c1 = c0 + h (high face), in a
// fixed pattern per axis.  So column b of J at Gauss point g is a function of (g, c0, c1) alone -- of the element's index
// along axis b -- and the 144 multiply-adds per Gauss point can come from three small tables (hex8_jacobian_column),
// bit-identical to accumulating them in place: the same products are added in the same order i = 0..7 onto 0.0.
// pattern of the high-face coordinate over the local nodes (get_hex8_node_coords_3d, Hex8_box_utils.hpp:59-99)
MFE_HD bool hex8_node_high(int axis, int i)
{
  return axis == 0 ? (i == 1 || i == 2 || i == 5 || i == 6) : axis == 1 ? (i == 2 || i == 3 || i == 6 || i == 7) : i >= 4;
}
// out[3] = J[0][axis], J[1][axis], J[2][axis] at Gauss point g for an element whose faces along `axis` are at c0 and c1
MFE_HD void hex8_jacobian_column(const double *gv, int g, int axis, double c0, double c1, double *out)
{
  const double *gvp = gv + 24 * g;
  double j0 = 0, j1 = 0, j2 = 0;
  for (int i = 0; i < 8; ++i) {
    const double c = hex8_node_high(axis, i) ? c1 : c0;
    j0 = MFE_ADD(j0, MFE_MUL(gvp[3 * i], c));
    j1 = MFE_ADD(j1, MFE_MUL(gvp[3 * i + 1], c));
    j2 = MFE_ADD(j2, MFE_MUL(gvp[3 * i + 2], c));
  }
  out[0] = j0; out[1] = j1; out[2] = j2;
}

// One Gauss point of one element, given its Jacobian: inverse + determinant (Hex8.hpp:234-257), then the contributions to
// K[36] (diffusionMatrix_symm, Hex8.hpp:300-341) and f[8] (sourceVector, Hex8.hpp:386-406).
MFE_HD void hex8_gauss_point(const double *gvp, const double *psig, double J00, double J01, double J02, double J10, double J11, double J12,
                             double J20, double J21, double J22, double *K, double *f)
{
  const double term0 = MFE_SUB(MFE_MUL(J22, J11), MFE_MUL(J21, J12));
  const double term1 = MFE_SUB(MFE_MUL(J22, J01), MFE_MUL(J21, J02));
  const double term2 = MFE_SUB(MFE_MUL(J12, J01), MFE_MUL(J11, J02));
  const double detJ = MFE_ADD(MFE_SUB(MFE_MUL(J00, term0), MFE_MUL(J10, term1)), MFE_MUL(J20, term2));
  const double inv_detJ = MFE_DIV(1.0, detJ);
  double iJ[9];
  iJ[0] = MFE_MUL(term0, inv_detJ);
  iJ[1] = MFE_MUL(-term1, inv_detJ);
  iJ[2] = MFE_MUL(term2, inv_detJ);
  iJ[3] = MFE_MUL(-(MFE_SUB(MFE_MUL(J22, J10), MFE_MUL(J20, J12))), inv_detJ);
  iJ[4] = MFE_MUL(MFE_SUB(MFE_MUL(J22, J00), MFE_MUL(J20, J02)), inv_detJ);
  iJ[5] = MFE_MUL(-(MFE_SUB(MFE_MUL(J12, J00), MFE_MUL(J10, J02))), inv_detJ);
  iJ[6] = MFE_MUL(MFE_SUB(MFE_MUL(J21, J10), MFE_MUL(J20, J11)), inv_detJ);
  iJ[7] = MFE_MUL(-(MFE_SUB(MFE_MUL(J21, J00), MFE_MUL(J20, J01))), inv_detJ);
  iJ[8] = MFE_MUL(MFE_SUB(MFE_MUL(J11, J00), MFE_MUL(J10, J01)), inv_detJ);

  // diffusionMatrix_symm (Hex8.hpp:300-341): k*detJ*wi*wj*wk with k = wi = wj = wk = 1.0
  const double kd = MFE_MUL(MFE_MUL(1.0, detJ), 1.0);
  double dx[8], dy[8], dz[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double g0 = gvp[3 * i], g1 = gvp[3 * i + 1], g2 = gvp[3 * i + 2];
    dx[i] = MFE_ADD(MFE_ADD(MFE_MUL(g0, iJ[0]), MFE_MUL(g1, iJ[1])), MFE_MUL(g2, iJ[2]));
    dy[i] = MFE_ADD(MFE_ADD(MFE_MUL(g0, iJ[3]), MFE_MUL(g1, iJ[4])), MFE_MUL(g2, iJ[5]));
    dz[i] = MFE_ADD(MFE_ADD(MFE_MUL(g0, iJ[6]), MFE_MUL(g1, iJ[7])), MFE_MUL(g2, iJ[8]));
  }
  int off = 0;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
#pragma unroll
    for (int n = m; n < 8; ++n) {
      const double s = MFE_ADD(MFE_ADD(MFE_MUL(dx[m], dx[n]), MFE_MUL(dy[m], dy[n])), MFE_MUL(dz[m], dz[n]));
      K[off] = MFE_ADD(K[off], MFE_MUL(kd, s));
      ++off;
    }
  }
  // sourceVector (Hex8.hpp:386-406): term = Q*detJ*wi*wj*wk, same detJ expression (gradients_and_detJ, 115-175)
  const double term = MFE_MUL(MFE_MUL(MFE_MUL(MFE_MUL(1.0, detJ), 1.0), 1.0), 1.0);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = MFE_ADD(f[i], MFE_MUL(psig[i], term));
}

// One element whose local node 0 sits at (x,y,z) with edge lengths (hx,hy,hz): K[36] upper triangle row-major,
// f[8].  gv = grad192, psi = psi64 from hex8_tables.  Gauss weights are all 1.0 and k = Q = 1.0, so the
// reference's k*detJ*wi*wj*wk and Q*detJ*wi*wj*wk are exactly detJ; the multiplications are kept anyway.
MFE_HD void hex8_element(const double *gv, const double *psi, double x, double y, double z,
                         double hx, double hy, double hz, double *K, double *f)
{
  // get_hex8_node_coords_3d (Hex8_box_utils.hpp:59-99)
  const double x1 = MFE_ADD(x, hx), y1 = MFE_ADD(y, hy), z1 = MFE_ADD(z, hz);
  const double cx[8] = { x, x1, x1, x, x, x1, x1, x };
  const double cy[8] = { y, y, y1, y1, y, y, y1, y1 };
  const double cz[8] = { z, z, z, z, z1, z1, z1, z1 };
#pragma unroll
  for (int i = 0; i < 36; ++i) K[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = 0.0;

#pragma unroll 1
  for (int g = 0; g < 8; ++g) {
    const double *gvp = gv + 24 * g;
    // gradients_and_invJ_and_detJ (Hex8.hpp:177-257)
    double J00 = 0, J01 = 0, J02 = 0, J10 = 0, J11 = 0, J12 = 0, J20 = 0, J21 = 0, J22 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double g0 = gvp[3 * i], g1 = gvp[3 * i + 1], g2 = gvp[3 * i + 2];
      J00 = MFE_ADD(J00, MFE_MUL(g0, cx[i])); J01 = MFE_ADD(J01, MFE_MUL(g0, cy[i])); J02 = MFE_ADD(J02, MFE_MUL(g0, cz[i]));
      J10 = MFE_ADD(J10, MFE_MUL(g1, cx[i])); J11 = MFE_ADD(J11, MFE_MUL(g1, cy[i])); J12 = MFE_ADD(J12, MFE_MUL(g1, cz[i]));
      J20 = MFE_ADD(J20, MFE_MUL(g2, cx[i])); J21 = MFE_ADD(J21, MFE_MUL(g2, cy[i])); J22 = MFE_ADD(J22, MFE_MUL(g2, cz[i]));
    }
    hex8_gauss_point(gvp, psi + 8 * g, J00, J01, J02, J10, J11, J12, J20, J21, J22, K, f);
  }
}

// The same element from the per-axis Jacobian tables: jx/jy/jz point at this element's 24 entries ([g][a] = J[a][axis] at
// Gauss point g) of the tables that hex8_jacobian_column filled.
MFE_HD void hex8_element_tables(const double *gv, const double *psi, const double *jx, const double *jy, const double *jz, double *K, double *f)
{
#pragma unroll
  for (int i = 0; i < 36; ++i) K[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = 0.0;
#pragma unroll 1
  for (int g = 0; g < 8; ++g)
    hex8_gauss_point(gv + 24 * g, psi + 8 * g, jx[3 * g], jy[3 * g], jz[3 * g], jx[3 * g + 1], jy[3 * g + 1], jz[3 * g + 1],
                     jx[3 * g + 2], jy[3 * g + 2], jz[3 * g + 2], K, f);
}

// exodus local node number of the corner (px,py,pz) in {0,1}^3 (Hex8_box_utils.hpp:41-57)
MFE_HD int hex8_local_node(int px, int py, int pz) { return pz * 4 + (py ? 3 - px : px); }
// index of K(m,n) in the 36-entry upper triangle, row-major
MFE_HD int hex8_sym_index(int m, int n)
{
  int a = m < n ? m : n, b = m < n ? n : m;
  return a * 8 - (a * (a - 1)) / 2 + (b - a);
}

}  // namespace mfe
