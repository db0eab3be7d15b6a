/* oracle/minife_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See minife_oracle.h.
 *
 * Plain-C restatement of miniFE ref/ (paths below are relative to /root/reference).
 * PARITY PINNED against the unmodified reference (oracle/_ref) by tests/test_oracle_vs_ref.py and
 * against tests/golden/ fixtures generated from it.
 */
#include "minife_oracle.h"
#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* Row offsets are `int` as in the reference (MINIFE_LOCAL_ORDINAL=int).  -DMFO_WIDE_OFFSETS makes them 64-bit so that one
 * rank can hold 512^3 (3.6e9 nonzeros) -- the route SURVEY 0.6 names for the reference itself (MINIFE_LOCAL_ORDINAL=
 * 'long long').  Only the type of the offsets and of the loop counters that walk the packed arrays changes: column
 * indices stay int, no arithmetic on values is touched.  tests/golden/make_oracle_golden.py --check pins the wide build to
 * the reference's own records. */
#ifdef MFO_WIDE_OFFSETS
typedef long long mfo_off;
#else
typedef int mfo_off;
#endif

typedef struct {
  int box[6];      /* element box  [x0,x1,y0,y1,z0,z1) */
  int nbox[6];     /* owned-node box */
  int nrows, first_row;
  int *rows, *cols;
  mfo_off *row_offsets;
  double *coefs, *b, *x;
  int *bc[2]; int nbc[2];
  int num_cols, has_local;
  int *external_index, *external_local_index, *elements_to_send, *neighbors, *recv_length, *send_length;
  int n_external, n_send, n_neighbors;
  int *new_external;   /* externals in local-index order (make_local_matrix.hpp:309-312) */
} mfo_rank;

struct mfo_world {
  int nx, ny, nz, np;
  int gnx, gny, gnz;     /* global node counts */
  int global_nrows;
  mfo_rank *r;
};

/* ---- ref/utils/BoxPartition.cpp:79-122 ------------------------------------------------------- */
static void box_partition_rec(int ip, int up, int axis, const int *box, int *p_box)
{
  int np = up - ip;
  if (np == 1) { memcpy(p_box + 6 * ip, box, 6 * sizeof(int)); return; }
  {
    int n = box[2 * axis + 1] - box[2 * axis];
    int np_low = np / 2, np_upp = np - np_low;
    int n_upp = (int)(((double)n) * (((double)np_upp) / ((double)np)));
    int n_low = n - n_upp;
    int next_axis = (axis + 2) % 3;
    if (np_low) {
      int d[6]; memcpy(d, box, sizeof d);
      d[2 * axis + 1] = d[2 * axis] + n_low;
      box_partition_rec(ip, ip + np_low, next_axis, d, p_box);
    }
    if (np_upp) {
      int d[6]; memcpy(d, box, sizeof d);
      ip += np_low;
      d[2 * axis] += n_low;
      d[2 * axis + 1] = d[2 * axis] + n_upp;
      box_partition_rec(ip, ip + np_upp, next_axis, d, p_box);
    }
  }
}
void mfo_box_partition(int nx, int ny, int nz, int np, int *out)
{
  int g[6] = { 0, nx, 0, ny, 0, nz };
  box_partition_rec(0, np, 2, g, out);   /* ref/src/main.cpp:112 */
}

/* ---- ref/utils/box_utils.hpp:99-111 get_id ---------------------------------------------------- */
static int get_id(int nx, int ny, int nz, int x, int y, int z)
{
  if (x < 0 || y < 0 || z < 0) return -1;
  if (x >= nx || y >= ny || z >= nz) return -1;
  return x + nx * y + nx * ny * z;
}

/* ---- ref/utils/box_utils.hpp:233-315 create_map_id_to_row + ref/utils/utils.hpp:165-199 -------
 * row = (exclusive prefix sum of owned-node counts of lower ranks) + x-fastest index in the owner's node box */
static int map_id_to_row(const mfo_world *w, int id)
{
  int nxy = w->gnx * w->gny, z = id / nxy, y = (id % nxy) / w->gnx, x = id % w->gnx, p;
  for (p = 0; p < w->np; ++p) {
    const int *nb = w->r[p].nbox;
    if (x >= nb[0] && x < nb[1] && y >= nb[2] && y < nb[3] && z >= nb[4] && z < nb[5]) {
      int lx = nb[1] - nb[0], ly = nb[3] - nb[2];
      return w->r[p].first_row + (x - nb[0]) + lx * ((y - nb[2]) + ly * (z - nb[4]));
    }
  }
  return -99;
}
int mfo_map_id_to_row(mfo_world *w, int rank, int id) { (void)rank; return map_id_to_row(w, id); }

static int cmp_int(const void *a, const void *b) { int x = *(const int *)a, y = *(const int *)b; return (x > y) - (x < y); }
static int sort_unique(int *v, int n)
{
  int i, m = 0;
  qsort(v, n, sizeof(int), cmp_int);
  for (i = 0; i < n; ++i) if (m == 0 || v[m - 1] != v[i]) v[m++] = v[i];
  return m;
}
static int in_set(const int *v, int n, int key) { return bsearch(&key, v, n, sizeof(int), cmp_int) != NULL; }

/* ---- ref/src/simple_mesh_description.hpp:41-226 (bc_rows_0 / bc_rows_1 incl. the ghost layer) -- */
static void build_bc_sets(mfo_world *w, int rank)
{
  mfo_rank *R = &w->r[rank];
  const int *nb = R->nbox, *lb = R->box;
  int mx = w->gnx, my = w->gny, mz = w->gnz, x1 = mx - 1, y1 = my - 1, z1 = mz - 1;
  int lox = nb[0] - (nb[0] > 0), hix = nb[1] + (nb[1] < mx);
  int loy = nb[2] - (nb[2] > 0), hiy = nb[3] + (nb[3] < my);
  int loz = nb[4] - (nb[4] > 0), hiz = nb[5] + (nb[5] < mz);
  int cap = 2 * ((hiy - loy) * (hiz - loz) + (hix - lox) * (hiz - loz) + (hix - lox) * (hiy - loy)) + 8;
  int *s0 = (int *)malloc(sizeof(int) * cap), *s1 = (int *)malloc(sizeof(int) * cap), n0 = 0, n1 = 0, ix, iy, iz;
  if (lb[0] == 0)      for (iz = loz; iz < hiz; ++iz) for (iy = loy; iy < hiy; ++iy) s0[n0++] = map_id_to_row(w, get_id(mx, my, mz, 0, iy, iz));
  if (lb[2] == 0)      for (iz = loz; iz < hiz; ++iz) for (ix = lox; ix < hix; ++ix) s0[n0++] = map_id_to_row(w, get_id(mx, my, mz, ix, 0, iz));
  if (lb[4] == 0)      for (iy = loy; iy < hiy; ++iy) for (ix = lox; ix < hix; ++ix) s0[n0++] = map_id_to_row(w, get_id(mx, my, mz, ix, iy, 0));
  if (lb[1] == w->nx)  for (iy = loy; iy < hiy; ++iy) for (iz = loz; iz < hiz; ++iz) s1[n1++] = map_id_to_row(w, get_id(mx, my, mz, x1, iy, iz));
  if (lb[3] == w->ny)  for (ix = lox; ix < hix; ++ix) for (iz = loz; iz < hiz; ++iz) s0[n0++] = map_id_to_row(w, get_id(mx, my, mz, ix, y1, iz));
  if (lb[5] == w->nz)  for (ix = lox; ix < hix; ++ix) for (iy = loy; iy < hiy; ++iy) s0[n0++] = map_id_to_row(w, get_id(mx, my, mz, ix, iy, z1));
  R->nbc[0] = sort_unique(s0, n0); R->bc[0] = s0;
  R->nbc[1] = sort_unique(s1, n1); R->bc[1] = s1;
}

mfo_world *mfo_world_create(int nx, int ny, int nz, int np) { return mfo_world_create_boxes(nx, ny, nz, np, NULL); }

/* boxes6np == NULL: box_partition (ref/src/main.cpp:112); otherwise the given element boxes, e.g. the ones
 * add_imbalance (ref/utils/imbalance.hpp:167-293, called at ref/src/driver.hpp:137-139) leaves on the ranks */
mfo_world *mfo_world_create_boxes(int nx, int ny, int nz, int np, const int *boxes6np)
{
  mfo_world *w = (mfo_world *)calloc(1, sizeof *w);
  int *boxes = (int *)malloc(sizeof(int) * 6 * np), p, a, off = 0;
  w->nx = nx; w->ny = ny; w->nz = nz; w->np = np;
  w->gnx = nx + 1; w->gny = ny + 1; w->gnz = nz + 1;
  w->global_nrows = w->gnx * w->gny * w->gnz;
  w->r = (mfo_rank *)calloc(np, sizeof(mfo_rank));
  if (boxes6np) memcpy(boxes, boxes6np, sizeof(int) * 6 * np);
  else mfo_box_partition(nx, ny, nz, np, boxes);
  {
    int g[3]; g[0] = nx; g[1] = ny; g[2] = nz;
    for (p = 0; p < np; ++p) {
      mfo_rank *R = &w->r[p];
      memcpy(R->box, boxes + 6 * p, sizeof R->box);
      memcpy(R->nbox, R->box, sizeof R->box);
      /* simple_mesh_description.hpp:49-53: +1 node on the global high side of a non-empty box */
      for (a = 0; a < 3; ++a) if (R->box[2 * a + 1] > R->box[2 * a] && R->box[2 * a + 1] == g[a]) R->nbox[2 * a + 1] += 1;
      R->nrows = (R->nbox[1] - R->nbox[0]) * (R->nbox[3] - R->nbox[2]) * (R->nbox[5] - R->nbox[4]);
      R->first_row = off;            /* box_utils.hpp:254-264 */
      off += R->nrows;
    }
  }
  free(boxes);
  for (p = 0; p < np; ++p) build_bc_sets(w, p);
  return w;
}

void mfo_world_destroy(mfo_world *w)
{
  int p;
  if (!w) return;
  for (p = 0; p < w->np; ++p) {
    mfo_rank *R = &w->r[p];
    free(R->rows); free(R->row_offsets); free(R->cols); free(R->coefs); free(R->b); free(R->x);
    free(R->bc[0]); free(R->bc[1]);
    free(R->external_index); free(R->external_local_index); free(R->elements_to_send);
    free(R->neighbors); free(R->recv_length); free(R->send_length); free(R->new_external);
  }
  free(w->r); free(w);
}

/* ---- ref/src/generate_matrix_structure.hpp:47-149 + MatrixInitOp.hpp:127-156 ------------------- */
int mfo_generate_matrix_structure(mfo_world *w)
{
  int p;
  for (p = 0; p < w->np; ++p) {
    mfo_rank *R = &w->r[p];
    const int *nb = R->nbox;
    int nrows = R->nrows, roff = 0, ix, iy, iz, sx, sy, sz;
    long long nnz = 0;
    R->rows = (int *)malloc(sizeof(int) * (nrows + 1));
    R->row_offsets = (mfo_off *)malloc(sizeof(mfo_off) * (nrows + 1));
    R->cols = (int *)malloc(sizeof(int) * ((size_t)nrows * 27 + 1));
    R->coefs = (double *)malloc(sizeof(double) * ((size_t)nrows * 27 + 1));
    for (iz = nb[4]; iz < nb[5]; ++iz) for (iy = nb[2]; iy < nb[3]; ++iy) for (ix = nb[0]; ix < nb[1]; ++ix) {
      int len = 0, *rc = R->cols + nnz, i, j;
      R->rows[roff] = map_id_to_row(w, get_id(w->gnx, w->gny, w->gnz, ix, iy, iz));
      R->row_offsets[roff++] = (mfo_off)nnz;
      for (sz = -1; sz <= 1; ++sz) for (sy = -1; sy <= 1; ++sy) for (sx = -1; sx <= 1; ++sx) {
        int col_id = get_id(w->gnx, w->gny, w->gnz, ix + sx, iy + sy, iz + sz);
        if (col_id >= 0 && col_id < w->global_nrows) { rc[len] = map_id_to_row(w, col_id); R->coefs[nnz + len] = 0; ++len; }
      }
      /* sort_if_needed (MatrixInitOp.hpp:40-54): ascending */
      for (i = 1; i < len; ++i) { int v = rc[i]; for (j = i; j > 0 && rc[j - 1] > v; --j) rc[j] = rc[j - 1]; rc[j] = v; }
      nnz += len;
    }
    R->row_offsets[roff] = (mfo_off)nnz;
    R->b = (double *)calloc(nrows + 1, sizeof(double));   /* Vector ctor zero-fills, Vector.hpp:44-52 */
    R->x = (double *)calloc(nrows + 1, sizeof(double));
    R->num_cols = 0; R->has_local = 0;
  }
  return 0;
}

/* ---- ref/fem/gauss_pts.hpp:37-62, Hex8.hpp:43-113, Hex8_ElemData.hpp:37-61 --------------------- */
static void gauss_pts2(double *pts, double *wts)
{
  const double x2 = 0.577350269;
  pts[0] = -x2; wts[0] = 1.0; pts[1] = x2; wts[1] = 1.0;
}
static void shape_fns(const double *x, double *v)
{
  const double u = 1.0 - x[0], vv = 1.0 - x[1], w = 1.0 - x[2];
  const double up1 = 1.0 + x[0], vp1 = 1.0 + x[1], wp1 = 1.0 + x[2];
  v[0] = 0.125 * u * vv * w;     v[1] = 0.125 * up1 * vv * w;
  v[2] = 0.125 * up1 * vp1 * w;  v[3] = 0.125 * u * vp1 * w;
  v[4] = 0.125 * u * vv * wp1;   v[5] = 0.125 * up1 * vv * wp1;
  v[6] = 0.125 * up1 * vp1 * wp1; v[7] = 0.125 * u * vp1 * wp1;
}
static void gradients(const double *x, double *g)
{
  const double u = 1.0 - x[0], v = 1.0 - x[1], w = 1.0 - x[2];
  const double up1 = 1.0 + x[0], vp1 = 1.0 + x[1], wp1 = 1.0 + x[2];
  g[0] = -0.125 * v * w;     g[1] = -0.125 * u * w;     g[2] = -0.125 * u * v;
  g[3] = 0.125 * v * w;      g[4] = -0.125 * up1 * w;   g[5] = -0.125 * up1 * v;
  g[6] = 0.125 * vp1 * w;    g[7] = 0.125 * up1 * w;    g[8] = -0.125 * up1 * vp1;
  g[9] = -0.125 * vp1 * w;   g[10] = 0.125 * u * w;     g[11] = -0.125 * u * vp1;
  g[12] = -0.125 * v * wp1;  g[13] = -0.125 * u * wp1;  g[14] = 0.125 * u * v;
  g[15] = 0.125 * v * wp1;   g[16] = -0.125 * up1 * wp1; g[17] = 0.125 * up1 * v;
  g[18] = 0.125 * vp1 * wp1; g[19] = 0.125 * up1 * wp1; g[20] = 0.125 * up1 * vp1;
  g[21] = -0.125 * vp1 * wp1; g[22] = 0.125 * u * wp1;  g[23] = 0.125 * u * vp1;
}
void mfo_gradient_values(double *out)
{
  double gp[2], gw[2], pt[3]; int ig, jg, kg;
  gauss_pts2(gp, gw);
  for (ig = 0; ig < 2; ++ig) { pt[0] = gp[ig];
    for (jg = 0; jg < 2; ++jg) { pt[1] = gp[jg];
      for (kg = 0; kg < 2; ++kg) { pt[2] = gp[kg]; gradients(pt, out); out += 24; } } }
}

/* ---- ref/fem/Hex8.hpp:177-257 (J, detJ, invJ) -------------------------------------------------- */
static void jac(const double *c, const double *gv, double *invJ, double *detJ_out)
{
  double J00 = 0, J01 = 0, J02 = 0, J10 = 0, J11 = 0, J12 = 0, J20 = 0, J21 = 0, J22 = 0;
  double term0, term1, term2, detJ, inv_detJ; int i;
  for (i = 0; i < 8; ++i) {
    const double *g = gv + 3 * i, *x = c + 3 * i;
    J00 += g[0] * x[0]; J01 += g[0] * x[1]; J02 += g[0] * x[2];
    J10 += g[1] * x[0]; J11 += g[1] * x[1]; J12 += g[1] * x[2];
    J20 += g[2] * x[0]; J21 += g[2] * x[1]; J22 += g[2] * x[2];
  }
  term0 = J22 * J11 - J21 * J12;
  term1 = J22 * J01 - J21 * J02;
  term2 = J12 * J01 - J11 * J02;
  detJ = J00 * term0 - J10 * term1 + J20 * term2;
  *detJ_out = detJ;
  if (!invJ) return;
  inv_detJ = 1.0 / detJ;
  invJ[0] = term0 * inv_detJ;  invJ[1] = -term1 * inv_detJ;  invJ[2] = term2 * inv_detJ;
  invJ[3] = -(J22 * J10 - J20 * J12) * inv_detJ;
  invJ[4] = (J22 * J00 - J20 * J02) * inv_detJ;
  invJ[5] = -(J12 * J00 - J10 * J02) * inv_detJ;
  invJ[6] = (J21 * J10 - J20 * J11) * inv_detJ;
  invJ[7] = -(J21 * J00 - J20 * J01) * inv_detJ;
  invJ[8] = (J11 * J00 - J10 * J01) * inv_detJ;
}

/* ---- ref/fem/Hex8.hpp:259-361 diffusionMatrix_symm, 363-410 sourceVector ----------------------- */
static void diffusion_matrix_symm(const double *c, const double *grad_vals, double *K)
{
  double gp[2], gw[2], invJ[9], detJ = 0.0, dx[8], dy[8], dz[8];
  const double k = 1.0;
  int ig, jg, kg, i, m, n, gvo = 0;
  for (i = 0; i < 36; ++i) K[i] = 0;
  gauss_pts2(gp, gw);
  for (ig = 0; ig < 2; ++ig) { double wi = gw[ig];
    for (jg = 0; jg < 2; ++jg) { double wi_wj = wi * gw[jg];
      for (kg = 0; kg < 2; ++kg) {
        double wijk = wi_wj * gw[kg], kd; const double *gv = grad_vals + gvo; int off = 0;
        gvo += 24;
        jac(c, gv, invJ, &detJ);
        kd = k * detJ * wijk;
        for (i = 0; i < 8; ++i) {
          double g0 = gv[3 * i], g1 = gv[3 * i + 1], g2 = gv[3 * i + 2];
          dx[i] = g0 * invJ[0] + g1 * invJ[1] + g2 * invJ[2];
          dy[i] = g0 * invJ[3] + g1 * invJ[4] + g2 * invJ[5];
          dz[i] = g0 * invJ[6] + g1 * invJ[7] + g2 * invJ[8];
        }
        for (m = 0; m < 8; ++m) {
          K[off++] += kd * ((dx[m] * dx[m]) + (dy[m] * dy[m]) + (dz[m] * dz[m]));
          for (n = m + 1; n < 8; ++n) K[off++] += kd * ((dx[m] * dx[n]) + (dy[m] * dy[n]) + (dz[m] * dz[n]));
        }
      } } }
}
static void source_vector(const double *c, const double *grad_vals, double *f)
{
  double gp[2], gw[2], psi[8], pt[3], detJ; const double Q = 1.0;
  int ig, jg, kg, i, gvo = 0;
  for (i = 0; i < 8; ++i) f[i] = 0;
  gauss_pts2(gp, gw);
  for (ig = 0; ig < 2; ++ig) { double wi = gw[ig]; pt[0] = gp[ig];
    for (jg = 0; jg < 2; ++jg) { double wj = gw[jg]; pt[1] = gp[jg];
      for (kg = 0; kg < 2; ++kg) { double wk = gw[kg], term; pt[2] = gp[kg];
        shape_fns(pt, psi);
        jac(c, grad_vals + gvo, NULL, &detJ); gvo += 24;
        term = Q * detJ * wi * wj * wk;
        for (i = 0; i < 8; ++i) f[i] += psi[i] * term;
      } } }
}

/* ---- ref/src/Hex8_box_utils.hpp:41-159 (node ids in exodus order; coords x0 = 1.0*ix/n, +h) ----- */
static void elem_nodes_and_coords(int nx, int ny, int nz, int elemID, int *ids, double *c)
{
  int gnx = nx + 1, gny = ny + 1, gnz = nz + 1;
  int ez = elemID / (nx * ny), ey = (elemID % (nx * ny)) / nx, ex = elemID % nx;   /* get_int_coords, box_utils.hpp:56-66 */
  int node0 = get_id(gnx, gny, gnz, ex, ey, ez);
  int xdiv = gnx > 1 ? gnx - 1 : 1, ydiv = gny > 1 ? gny - 1 : 1, zdiv = gnz > 1 ? gnz - 1 : 1;
  double z = (1.0 * (node0 / (gnx * gny))) / zdiv;          /* get_coords, box_utils.hpp:68-82 */
  double y = 1.0 * ((node0 % (gnx * gny)) / gnx) / ydiv;
  double x = 1.0 * (node0 % gnx) / xdiv;
  double hx = 1.0 / nx, hy = 1.0 / ny, hz = 1.0 / nz;
  ids[0] = node0;             ids[1] = node0 + 1;
  ids[2] = node0 + gnx + 1;   ids[3] = node0 + gnx;
  ids[4] = node0 + gnx * gny; ids[5] = node0 + 1 + gnx * gny;
  ids[6] = node0 + gnx + gnx * gny + 1; ids[7] = node0 + gnx + gnx * gny;
  c[0] = x;       c[1] = y;       c[2] = z;
  c[3] = x + hx;  c[4] = y;       c[5] = z;
  c[6] = x + hx;  c[7] = y + hy;  c[8] = z;
  c[9] = x;       c[10] = y + hy; c[11] = z;
  c[12] = x;      c[13] = y;      c[14] = z + hz;
  c[15] = x + hx; c[16] = y;      c[17] = z + hz;
  c[18] = x + hx; c[19] = y + hy; c[20] = z + hz;
  c[21] = x;      c[22] = y + hy; c[23] = z + hz;
}

void mfo_element(int nx, int ny, int nz, int elemID, int *ids8, double *c24, double *K36, double *f8)
{
  double gv[192], c[24], K[36], f[8]; int ids[8];
  mfo_gradient_values(gv);
  elem_nodes_and_coords(nx, ny, nz, elemID, ids, c);
  diffusion_matrix_symm(c, gv, K);
  source_vector(c, gv, f);
  if (ids8) memcpy(ids8, ids, sizeof ids);
  if (c24) memcpy(c24, c, sizeof c);
  if (K36) memcpy(K36, K, sizeof K);
  if (f8) memcpy(f8, f, sizeof f);
}

/* ---- ref/src/SparseMatrix_functions.hpp:142-160 sum_into_row (lower_bound in the sorted row) ---- */
static void sum_into_row(mfo_rank *R, int local_row, int col, double v)
{
  mfo_off lo = R->row_offsets[local_row], hi = R->row_offsets[local_row + 1], end = hi;
  while (lo < hi) { mfo_off mid = lo + (hi - lo) / 2; if (R->cols[mid] < col) lo = mid + 1; else hi = mid; }
  if (lo < end && R->cols[lo] == col) R->coefs[lo] += v;
}

/* ---- assemble_FE_data.hpp:38-73, perform_element_loop.hpp:40-104,
 *      SparseMatrix_functions.hpp:183-255 (sum_in_symm_elem_matrix), Vector_functions.hpp:84-103 --- */
void mfo_assemble_FE_data(mfo_world *w)
{
  double gv[192]; int p;
  mfo_gradient_values(gv);
  for (p = 0; p < w->np; ++p) {
    mfo_rank *R = &w->r[p];
    int eb[6], ex, ey, ez;
    memcpy(eb, R->box, sizeof eb);
    if ((eb[1] - eb[0]) * (eb[3] - eb[2]) * (eb[5] - eb[4]) < 1) continue;
    if (eb[0] > 0) eb[0] -= 1; if (eb[2] > 0) eb[2] -= 1; if (eb[4] > 0) eb[4] -= 1;
    if (eb[1] < w->nx) eb[1] += 1; if (eb[3] < w->ny) eb[3] += 1; if (eb[5] < w->nz) eb[5] += 1;
    for (ez = eb[4]; ez < eb[5]; ++ez) for (ey = eb[2]; ey < eb[3]; ++ey) for (ex = eb[0]; ex < eb[1]; ++ex) {
      int elemID = get_id(w->nx, w->ny, w->nz, ex, ey, ez), ids[8], i, j, row_offset = 0;
      double c[24], K[36], f[8];
      elem_nodes_and_coords(w->nx, w->ny, w->nz, elemID, ids, c);
      for (i = 0; i < 8; ++i) ids[i] = map_id_to_row(w, ids[i]);
      diffusion_matrix_symm(c, gv, K);
      source_vector(c, gv, f);
      for (i = 0; i < 8; ++i) {
        int row = ids[i], lr = row - R->first_row, row_len = 8 - i, offset;
        const double *row_coefs = K + row_offset;
        row_offset += row_len;
        if (lr < 0 || lr >= R->nrows) continue;           /* get_row_pointers: not local -> skipped */
        for (j = 0; j < row_len; ++j) sum_into_row(R, lr, ids[i + j], row_coefs[j]);
        offset = i;
        for (j = 0; j < i; ++j) { sum_into_row(R, lr, ids[j], K[offset]); offset += 8 - (j + 1); }
      }
      for (i = 0; i < 8; ++i) {
        int lr = ids[i] - R->first_row;
        if (lr < 0 || lr >= R->nrows) continue;
        R->b[lr] += f[i];
      }
    }
  }
}

/* ---- ref/src/SparseMatrix_functions.hpp:411-480 impose_dirichlet ------------------------------- */
void mfo_impose_dirichlet(mfo_world *w, int which)
{
  int p; double val = which == 0 ? 0.0 : 1.0;
  for (p = 0; p < w->np; ++p) {
    mfo_rank *R = &w->r[p];
    const int *bc = R->bc[which]; int nbc = R->nbc[which], k, i; mfo_off j;
    int first = R->nrows > 0 ? R->rows[0] : 0, last = R->nrows > 0 ? R->rows[R->nrows - 1] : -1;
    for (k = 0; k < nbc; ++k) {
      int row = bc[k];
      if (row >= first && row <= last) {
        int lr = row - first;
        R->b[lr] = val;
        for (j = R->row_offsets[lr]; j < R->row_offsets[lr + 1]; ++j) R->coefs[j] = (R->cols[j] == row) ? 1 : 0;
      }
    }
    for (i = 0; i < R->nrows; ++i) {
      double sum = 0;
      if (in_set(bc, nbc, R->rows[i])) continue;
      for (j = R->row_offsets[i]; j < R->row_offsets[i + 1]; ++j)
        if (in_set(bc, nbc, R->cols[j])) { sum += R->coefs[j]; R->coefs[j] = 0; }
      R->b[i] -= sum * val;
    }
  }
}

/* ---- ref/src/make_local_matrix.hpp:42-442 (messages replaced by reads of the peer's struct) ----- */
void mfo_make_local_matrix(mfo_world *w)
{
  int np = w->np, p, q;
  if (np < 2) {
    for (p = 0; p < np; ++p) { w->r[p].num_cols = w->r[p].nrows; w->r[p].has_local = 1; }
    return;
  }
  {
    int *ext_slot = (int *)malloc(sizeof(int) * (size_t)w->global_nrows);   /* std::map externals (87-107) */
    int **ext_proc_new = (int **)calloc(np, sizeof(int *));
    for (p = 0; p < np; ++p) {
      mfo_rank *R = &w->r[p];
      int start_row = R->nrows > 0 ? R->rows[0] : -1, stop_row = R->nrows > 0 ? R->rows[R->nrows - 1] : -1;
      mfo_off nnz = R->row_offsets[R->nrows], j; int i, num_external = 0, count, *eproc, *newproc;
      R->external_index = (int *)malloc(sizeof(int) * ((size_t)nnz + 1));
      memset(ext_slot, 0xff, sizeof(int) * (size_t)w->global_nrows);
      for (j = 0; j < nnz; ++j) {
        int cur = R->cols[j];
        if (start_row <= cur && cur <= stop_row) R->cols[j] -= start_row;
        else {
          if (ext_slot[cur] < 0) { ext_slot[cur] = num_external; R->external_index[num_external++] = cur; }
          R->cols[j] = -(R->cols[j] + 1);
        }
      }
      R->n_external = num_external;
      eproc = (int *)malloc(sizeof(int) * (num_external + 1));
      for (i = 0; i < num_external; ++i) {        /* 135-143: highest rank whose start_row <= id */
        int cur = R->external_index[i];
        for (q = np - 1; q >= 0; --q) {
          int off = w->r[q].nrows > 0 ? w->r[q].first_row : -1;
          if (off <= cur && off >= 0) { eproc[i] = q; break; }
        }
      }
      count = R->nrows;                            /* 153-166 */
      R->external_local_index = (int *)malloc(sizeof(int) * (num_external + 1));
      for (i = 0; i < num_external; ++i) R->external_local_index[i] = -1;
      for (i = 0; i < num_external; ++i) {
        if (R->external_local_index[i] == -1) {
          R->external_local_index[i] = count++;
          for (j = i + 1; j < num_external; ++j) if (eproc[j] == eproc[i]) R->external_local_index[j] = count++;
        }
      }
      for (j = 0; j < nnz; ++j) if (R->cols[j] < 0) {   /* 168-180 */
        int cur = -R->cols[j] - 1;
        R->cols[j] = R->external_local_index[ext_slot[cur]];
      }
      newproc = (int *)malloc(sizeof(int) * (num_external + 1));
      R->new_external = (int *)malloc(sizeof(int) * (num_external + 1));
      for (i = 0; i < num_external; ++i) {
        newproc[R->external_local_index[i] - R->nrows] = eproc[i];
        R->new_external[R->external_local_index[i] - R->nrows] = R->external_index[i];
      }
      ext_proc_new[p] = newproc;
      /* recv_list (236-242) and recv_length (344-363) */
      R->neighbors = (int *)malloc(sizeof(int) * (np + 1));
      R->recv_length = (int *)calloc(np + 1, sizeof(int));
      R->send_length = (int *)calloc(np + 1, sizeof(int));
      R->n_neighbors = 0;
      for (i = 0; i < num_external; ++i) {
        if (i == 0 || newproc[i - 1] != newproc[i]) R->neighbors[R->n_neighbors++] = newproc[i];
        R->recv_length[R->n_neighbors - 1] += 1;
      }
      free(eproc);
      R->num_cols = R->nrows + num_external;
      R->has_local = 1;
    }
    /* pure send-neighbours appended (291-301); never happens for the symmetric stencil */
    for (p = 0; p < np; ++p) for (q = 0; q < np; ++q) if (q != p) {
      mfo_rank *Q = &w->r[q], *R = &w->r[p]; int i, has = 0, found = 0;
      for (i = 0; i < Q->n_neighbors; ++i) if (Q->neighbors[i] == p && Q->recv_length[i] > 0) has = 1;
      for (i = 0; i < R->n_neighbors; ++i) if (R->neighbors[i] == q) found = 1;
      if (has && !found) { R->neighbors[R->n_neighbors] = q; R->recv_length[R->n_neighbors] = 0; R->n_neighbors++; }
    }
    /* send_length (371-377) and elements_to_send (386-429) */
    for (p = 0; p < np; ++p) {
      mfo_rank *R = &w->r[p]; int i, total = 0, k = 0;
      for (i = 0; i < R->n_neighbors; ++i) {
        mfo_rank *Q = &w->r[R->neighbors[i]]; int t;
        for (t = 0; t < Q->n_neighbors; ++t) if (Q->neighbors[t] == p) R->send_length[i] = Q->recv_length[t];
        total += R->send_length[i];
      }
      R->n_send = total;
      R->elements_to_send = (int *)malloc(sizeof(int) * (total + 1));
      for (i = 0; i < R->n_neighbors; ++i) {
        mfo_rank *Q = &w->r[R->neighbors[i]]; int t, start = 0;
        for (t = 0; t < Q->n_neighbors; ++t) {
          if (Q->neighbors[t] == p) { int e; for (e = 0; e < Q->recv_length[t]; ++e) R->elements_to_send[k++] = Q->new_external[start + e] - R->first_row; }
          start += Q->recv_length[t];
        }
      }
    }
    for (p = 0; p < np; ++p) free(ext_proc_new[p]);
    free(ext_proc_new); free(ext_slot);
  }
}

/* ---- ref/src/exchange_externals.hpp:44-161 ---------------------------------------------------- */
void mfo_exchange_externals(mfo_world *w, double **x)
{
  int p;
  if (w->np < 2) return;
  for (p = 0; p < w->np; ++p) {          /* receiver side: externals of p arrive grouped by neighbour, in order */
    mfo_rank *R = &w->r[p]; int i, dst = R->nrows;
    for (i = 0; i < R->n_neighbors; ++i) {
      int q = R->neighbors[i], t, k = 0, e; mfo_rank *Q = &w->r[q];
      for (t = 0; t < Q->n_neighbors; ++t) {
        if (Q->neighbors[t] == p) for (e = 0; e < Q->send_length[t]; ++e) x[p][dst++] = x[q][Q->elements_to_send[k + e]];
        k += Q->send_length[t];
      }
    }
  }
}

/* ---- ref/src/SparseMatrix_functions.hpp:546-578 matvec_std (local part) ------------------------ */
void mfo_matvec(mfo_world *w, int rank, const double *x, double *y)
{
  mfo_rank *R = &w->r[rank]; int row; mfo_off i; const double beta = 0;
  for (row = 0; row < R->nrows; ++row) {
    double sum = beta * y[row];
    for (i = R->row_offsets[row]; i < R->row_offsets[row + 1]; ++i) sum += R->coefs[i] * x[R->cols[i]];
    y[row] = sum;
  }
}
/* ---- ref/src/SparseMatrix_functions.hpp:319-408 rearrange_matrix_local_external + 624-675 matvec_overlap -------
 * The reference sorts every row by (local) column index (sort_with_companions, 347), which puts the owned columns
 * (< nrows) first and the externals after them, then moves the external entries of all rows to the back of the packed
 * arrays (row_offsets_external).  matvec_overlap sums a row's local entries onto 0*y (first loop, 649-657), stores, and
 * adds its external entries onto 1*y (second loop, 664-672): per row that is ONE left-to-right pass over
 * [locals ascending | externals ascending], because 1*y is exact.  So the oracle keeps one CSR and only re-sorts each
 * row; mfo_matvec then returns matvec_overlap's bits. */
void mfo_rearrange_matrix_local_external(mfo_world *w)
{
  int p;
  for (p = 0; p < w->np; ++p) {
    mfo_rank *R = &w->r[p]; int row;
    for (row = 0; row < R->nrows; ++row) {
      mfo_off b = R->row_offsets[row], e = R->row_offsets[row + 1], i, j;
      for (i = b + 1; i < e; ++i) {           /* insertion sort of (cols, coefs) by column */
        int c = R->cols[i]; double v = R->coefs[i];
        for (j = i; j > b && R->cols[j - 1] > c; --j) { R->cols[j] = R->cols[j - 1]; R->coefs[j] = R->coefs[j - 1]; }
        R->cols[j] = c; R->coefs[j] = v;
      }
    }
  }
}
/* ---- ref/src/Vector_functions.hpp:127-150, 194-226 --------------------------------------------- */
void mfo_waxpby(long long n, double alpha, const double *x, double beta, const double *y, double *wv)
{
  long long i;
  for (i = 0; i < n; ++i) wv[i] = alpha * x[i] + beta * y[i];
}
double mfo_dot(long long n, const double *x, const double *y)
{
  double result = 0; long long i;
  for (i = 0; i < n; ++i) result += x[i] * y[i];
  return result;
}
static double world_dot(mfo_world *w, double **x, double **y)
{
  double g = 0; int p;   /* MPI_Allreduce(SUM) in rank order, Vector_functions.hpp:218-222 */
  for (p = 0; p < w->np; ++p) { double l = mfo_dot(w->r[p].nrows, x[p], y[p]); g = (p == 0) ? l : g + l; }
  return g;
}
static void world_matvec(mfo_world *w, double **x, double **y)
{
  int p;
  mfo_exchange_externals(w, x);
  for (p = 0; p < w->np; ++p) mfo_matvec(w, p, x[p], y[p]);
}
static void world_waxpby(mfo_world *w, double a, double **x, double b, double **y, double **o)
{
  int p;
  for (p = 0; p < w->np; ++p) mfo_waxpby(w->r[p].nrows, a, x[p], b, y[p], o[p]);
}

/* ---- ref/src/cg_solve.hpp:49-229 ---------------------------------------------------------------- */
int mfo_cg_solve(mfo_world *w, int max_iter, double tol, int *num_iters, double *normr_out,
                 double *normr_hist, double *rtrans_hist, double *pap_hist, double *alpha_hist)
{
  int np = w->np, p, k, rc = 0, iters = 0;
  double **r = (double **)malloc(sizeof(double *) * np), **pp = (double **)malloc(sizeof(double *) * np);
  double **Ap = (double **)malloc(sizeof(double *) * np), **x = (double **)malloc(sizeof(double *) * np);
  double **b = (double **)malloc(sizeof(double *) * np);
  double normr = 0, rtrans = 0, oldrtrans = 0, brkdown_tol = DBL_EPSILON;
  const double one = 1.0, zero = 0.0;
  for (p = 0; p < np; ++p) {
    mfo_rank *R = &w->r[p];
    if (!R->has_local) { rc = -1; }
    r[p] = (double *)calloc(R->nrows + 1, sizeof(double));
    pp[p] = (double *)calloc(R->num_cols + 1, sizeof(double));
    Ap[p] = (double *)calloc(R->nrows + 1, sizeof(double));
    x[p] = R->x; b[p] = R->b;
  }
  if (rc == 0) {
    world_waxpby(w, one, x, zero, x, pp);
    world_matvec(w, pp, Ap);
    world_waxpby(w, one, b, -one, Ap, r);
    rtrans = world_dot(w, r, r);
    normr = sqrt(rtrans);
    if (normr_hist) normr_hist[0] = normr;
    if (rtrans_hist) rtrans_hist[0] = rtrans;
    for (k = 1; k <= max_iter && normr > tol; ++k) {
      double alpha, p_ap_dot;
      if (k == 1) world_waxpby(w, one, r, zero, r, pp);
      else {
        double beta;
        oldrtrans = rtrans;
        rtrans = world_dot(w, r, r);
        beta = rtrans / oldrtrans;
        world_waxpby(w, one, r, beta, pp, pp);
      }
      normr = sqrt(rtrans);
      if (normr_hist) normr_hist[k] = normr;
      if (rtrans_hist) rtrans_hist[k] = rtrans;
      world_matvec(w, pp, Ap);
      p_ap_dot = world_dot(w, Ap, pp);
      if (pap_hist) pap_hist[k] = p_ap_dot;
      if (p_ap_dot < brkdown_tol) {
        int brk = p_ap_dot < 0;
        if (!brk) {   /* breakdown(), cg_solve.hpp:49-67 */
          double vnorm = sqrt(world_dot(w, Ap, Ap)), wnorm = sqrt(world_dot(w, pp, pp));
          brk = fabs(p_ap_dot) <= 100 * vnorm * wnorm * DBL_EPSILON;
        }
        if (brk) { rc = 1; break; }
        brkdown_tol = 0.1 * p_ap_dot;
      }
      alpha = rtrans / p_ap_dot;
      if (alpha_hist) alpha_hist[k] = alpha;
      world_waxpby(w, one, x, alpha, pp, x);
      world_waxpby(w, one, r, -alpha, Ap, r);
      iters = k;
    }
  }
  *num_iters = iters; *normr_out = normr;
  for (p = 0; p < np; ++p) { free(r[p]); free(pp[p]); free(Ap[p]); }
  free(r); free(pp); free(Ap); free(x); free(b);
  return rc;
}

/* ---- queries ------------------------------------------------------------------------------------ */
long long mfo_nrows(mfo_world *w, int p) { return w->r[p].nrows; }
long long mfo_nnz(mfo_world *w, int p) { return w->r[p].row_offsets ? w->r[p].row_offsets[w->r[p].nrows] : 0; }
long long mfo_num_cols(mfo_world *w, int p) { return w->r[p].num_cols; }
long long mfo_first_row(mfo_world *w, int p) { return w->r[p].nrows > 0 ? w->r[p].first_row : -1; }
void mfo_get_box(mfo_world *w, int p, int *b) { memcpy(b, w->r[p].box, 6 * sizeof(int)); }
void mfo_get_csr(mfo_world *w, int p, int *rows, int *ro, int *cols, double *coefs)
{
  mfo_rank *R = &w->r[p]; size_t nnz = (size_t)R->row_offsets[R->nrows]; int i;
  if (rows) memcpy(rows, R->rows, sizeof(int) * R->nrows);
  if (ro) for (i = 0; i <= R->nrows; ++i) ro[i] = (int)R->row_offsets[i];   /* (truncates in a wide build beyond 2^31) */
  if (cols) memcpy(cols, R->cols, sizeof(int) * nnz);
  if (coefs) memcpy(coefs, R->coefs, sizeof(double) * nnz);
}
void mfo_get_b(mfo_world *w, int p, double *o) { memcpy(o, w->r[p].b, sizeof(double) * w->r[p].nrows); }
void mfo_get_x(mfo_world *w, int p, double *o) { memcpy(o, w->r[p].x, sizeof(double) * w->r[p].nrows); }
void mfo_set_x(mfo_world *w, int p, const double *in) { memcpy(w->r[p].x, in, sizeof(double) * w->r[p].nrows); }
long long mfo_bc_rows(mfo_world *w, int p, int which, int *out)
{
  if (out) memcpy(out, w->r[p].bc[which], sizeof(int) * w->r[p].nbc[which]);
  return w->r[p].nbc[which];
}
long long mfo_get_map(mfo_world *w, int p, const char *name, int *out)
{
  mfo_rank *R = &w->r[p]; const int *src = NULL; long long n = 0;
  if (!strcmp(name, "external_index")) { src = R->external_index; n = R->n_external; }
  else if (!strcmp(name, "external_local_index")) { src = R->external_local_index; n = R->n_external; }
  else if (!strcmp(name, "elements_to_send")) { src = R->elements_to_send; n = R->n_send; }
  else if (!strcmp(name, "neighbors")) { src = R->neighbors; n = R->n_neighbors; }
  else if (!strcmp(name, "recv_length")) { src = R->recv_length; n = R->n_neighbors; }
  else if (!strcmp(name, "send_length")) { src = R->send_length; n = R->n_neighbors; }
  else return -1;
  if (out && src && n > 0) memcpy(out, src, sizeof(int) * n);
  return n;
}

/* ---- ref/fem/analytic_soln.hpp:56-118 ------------------------------------------------------------ */
static const double AS_PI = 3.141592653589793238462;
static double as_fcn_l(int p, int q)
{
  const double PI_SQR = AS_PI * AS_PI;
  return sqrt((2 * p + 1) * (2 * p + 1) * PI_SQR + (2 * q + 1) * (2 * q + 1) * PI_SQR);     /* :71-74 */
}
static double as_fcn(int n, double u) { return (2 * n + 1) * AS_PI * u; }                     /* :76-79 */
static double as_soln(double x, double y, double z, int max_p, int max_q)                     /* :81-116 */
{
  const double PI_SQR = AS_PI * AS_PI, term0 = 16.0 / PI_SQR;
  double sum = 0;
  int p, q;
  for (p = 0; p <= max_p; ++p) {
    const double p21y = as_fcn(p, y);
    const double sin_py = sin(p21y) / (2 * p + 1);
    for (q = 0; q <= max_q; ++q) {
      const double q21z = as_fcn(q, z);
      const double sin_qz = sin(q21z) / (2 * q + 1);
      const double l = as_fcn_l(p, q);
      const double sinh1 = sinh(l * x);
      const double sinh2 = sinh(l);
      const double tmp = (sinh1 * sin_py) * (sin_qz / sinh2);
      if (tmp == tmp) sum += tmp;
      else break;
    }
  }
  return term0 * sum;
}

/* ---- ref/fem/verify_solution.hpp:55-174 ----------------------------------------------------------
 * Verifies the x left by mfo_cg_solve (or mfo_set_x).  info6 (may be NULL) = {max err, computed, analytic,
 * x, y, z} of the rank holding the global maximum (lowest such rank).  Returns the reference's return code. */
int mfo_verify_solution(mfo_world *w, double tolerance, int verify_whole_domain, double *global_max_err, double *info6)
{
  const int num_terms = 300;
  double gmax = 0.0, best[6] = { 0, 0, 0, 0, 0, 0 };
  int p;
  for (p = 0; p < w->np; ++p) {
    const mfo_rank *R = &w->r[p];
    double loc[6] = { 0, 0, 0, 0, 0, 0 };
    int roffset = 0, ix, iy, iz;
    const int xdiv = w->gnx > 1 ? w->gnx - 1 : 1, ydiv = w->gny > 1 ? w->gny - 1 : 1, zdiv = w->gnz > 1 ? w->gnz - 1 : 1;
    for (iz = R->nbox[4]; iz < R->nbox[5]; ++iz)
      for (iy = R->nbox[2]; iy < R->nbox[3]; ++iy)
        for (ix = R->nbox[0]; ix < R->nbox[1]; ++ix, ++roffset) {
          /* get_coords, ref/utils/box_utils.hpp:68-82, of id = get_id(ix,iy,iz) */
          const double z = (1.0 * iz) / zdiv, y = 1.0 * iy / ydiv, x = 1.0 * ix / xdiv;
          double analytic, err;
          if (!(verify_whole_domain || (fabs(x - 0.5) < 0.05 && fabs(y - 0.5) < 0.05 && fabs(z - 0.5) < 0.05))) continue;
          if (x == 1.0) analytic = 1;
          else if (x == 0.0 || y == 0.0 || z == 0.0) analytic = 0;
          else if (y == 1.0 || z == 1.0) analytic = 0;
          else analytic = as_soln(x, y, z, num_terms, num_terms);
          err = fabs(analytic - R->x[roffset]);
          if (err > loc[0]) { loc[0] = err; loc[1] = R->x[roffset]; loc[2] = analytic; loc[3] = x; loc[4] = y; loc[5] = z; }
        }
    if (loc[0] > gmax) { gmax = loc[0]; memcpy(best, loc, sizeof best); }
  }
  if (global_max_err) *global_max_err = gmax;
  if (info6) memcpy(info6, best, sizeof best);
  return gmax > tolerance ? 1 : 0;
}
