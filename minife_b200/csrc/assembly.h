// assembly.h -- per-row building blocks of the assembly and Dirichlet kernels, shared with the CPU test harness.
#pragma once
#include "mesh.h"
#include "hex8.h"

namespace mfe {

struct ElemSlab {
  int ex0, ey0, ez0;     // first element of the slab (global element coords)
  int nex, ney, nez;     // slab extents
  int64_t stride;        // >= nex*ney*nez
};

// Row-owner gather: ref/src/SparseMatrix_functions.hpp:183-255 + Vector_functions.hpp:84-103.
// The reference scatters element by element (ascending element id) with +=; every matrix entry therefore
// receives its <= 8 contributions in ascending element order onto 0.0.  The owner of a row adds the same
// contributions in the same order -- atomic-free and bit-identical.
// KF: SoA scratch, KF[k*stride + e], k < 36 the symmetric element matrix, 36..43 the source vector.
// acc[s]: value of the stencil slot s = (sz+1)*9 + (sy+1)*3 + (sx+1).
MFE_HD void gather_row(const MeshDesc &m, const ElemSlab &sl, const double *KF, int ix, int iy, int iz, double *acc, double &bsum)
{
#pragma unroll
  for (int s = 0; s < 27; ++s) acc[s] = 0.0;
  bsum = 0.0;
#pragma unroll
  for (int dz = -1; dz <= 0; ++dz)
#pragma unroll
    for (int dy = -1; dy <= 0; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 0; ++dx) {
        const int ex = ix + dx, ey = iy + dy, ez = iz + dz;
        if (ex < 0 || ey < 0 || ez < 0 || ex >= m.nx || ey >= m.ny || ez >= m.nz) continue;
        const int64_t e = (int64_t)(ex - sl.ex0) + (int64_t)sl.nex * ((ey - sl.ey0) + (int64_t)sl.ney * (ez - sl.ez0));
        const int li = hex8_local_node(-dx, -dy, -dz);     // position of the row node inside this element
#pragma unroll
        for (int qz = 0; qz <= 1; ++qz)
#pragma unroll
          for (int qy = 0; qy <= 1; ++qy)
#pragma unroll
            for (int qx = 0; qx <= 1; ++qx) {
              const int lj = hex8_local_node(qx, qy, qz);
              const int s = (dz + qz + 1) * 9 + (dy + qy + 1) * 3 + (dx + qx + 1);
              const double v = KF[(int64_t)hex8_sym_index(li, lj) * sl.stride + e];
              acc[s] = MFE_ADD(acc[s], v);
            }
        bsum = MFE_ADD(bsum, KF[(int64_t)(36 + li) * sl.stride + e]);
      }
}

// The same gather out of a two-plane ring of element matrices in shared memory (k_assemble_fused): element plane ez lives in
// slot ez & 1; inside a plane the elements of the CTA's window are numbered e2 = (ex - wx0) + wnx * (ey - wy0);
// KF[(slot * 44 + k) * wstride + e2].
MFE_HD void gather_row_ring(const MeshDesc &m, const double *KF, int wx0, int wy0, int wnx, int wstride, int ix, int iy, int iz, double *acc, double &bsum)
{
#pragma unroll
  for (int s = 0; s < 27; ++s) acc[s] = 0.0;
  bsum = 0.0;
#pragma unroll
  for (int dz = -1; dz <= 0; ++dz)
#pragma unroll
    for (int dy = -1; dy <= 0; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 0; ++dx) {
        const int ex = ix + dx, ey = iy + dy, ez = iz + dz;
        if (ex < 0 || ey < 0 || ez < 0 || ex >= m.nx || ey >= m.ny || ez >= m.nz) continue;
        const double *E = KF + (size_t)((ez & 1) * 44) * wstride + ((ex - wx0) + wnx * (ey - wy0));
        const int li = hex8_local_node(-dx, -dy, -dz);     // position of the row node inside this element
#pragma unroll
        for (int qz = 0; qz <= 1; ++qz)
#pragma unroll
          for (int qy = 0; qy <= 1; ++qy)
#pragma unroll
            for (int qx = 0; qx <= 1; ++qx) {
              const int lj = hex8_local_node(qx, qy, qz);
              const int s = (dz + qz + 1) * 9 + (dy + qy + 1) * 3 + (dx + qx + 1);
              acc[s] = MFE_ADD(acc[s], E[(size_t)hex8_sym_index(li, lj) * wstride]);
            }
        bsum = MFE_ADD(bsum, E[(size_t)(36 + li) * wstride]);
      }
}

// write the row's values at their sorted-column positions
MFE_HD void store_row(const MeshDesc &m, int ix, int iy, int iz, int64_t off, const double *acc, double *coefs)
{
  if (neighbourhood_is_local(m, ix, iy, iz)) {
    int j = 0;
#pragma unroll
    for (int s = 0; s < 27; ++s) {
      const int sx = s % 3 - 1, sy = (s / 3) % 3 - 1, sz = s / 9 - 1;
      const int jx = ix + sx, jy = iy + sy, jz = iz + sz;
      if (jx < 0 || jy < 0 || jz < 0 || jx >= m.gnx || jy >= m.gny || jz >= m.gnz) continue;
      coefs[off + j] = acc[s];
      ++j;
    }
  } else {
    // columns owned by several ranks: position = rank of the column's global row among the row's columns
    int c[27];
    unsigned valid = 0;
#pragma unroll
    for (int s = 0; s < 27; ++s) {
      const int sx = s % 3 - 1, sy = (s / 3) % 3 - 1, sz = s / 9 - 1;
      const int jx = ix + sx, jy = iy + sy, jz = iz + sz;
      if (jx < 0 || jy < 0 || jz < 0 || jx >= m.gnx || jy >= m.gny || jz >= m.gnz) { c[s] = 0x7fffffff; continue; }
      c[s] = node_to_row(m, jx, jy, jz);
      valid |= 1u << s;
    }
#pragma unroll
    for (int s = 0; s < 27; ++s) {
      if (!((valid >> s) & 1u)) continue;
      const int cs = c[s];
      int pos = 0;
      for (int q = 0; q < 27; ++q) pos += (((valid >> q) & 1u) && c[q] < cs) ? 1 : 0;
      coefs[off + pos] = acc[s];
    }
  }
}

// rows that one impose_dirichlet call can change: within one node of a constrained face of this rank
MFE_HD bool dirichlet_touches(const MeshDesc &m, int which, int ix, int iy, int iz)
{
  if (which == 0)
    return (m.face_lo[0] && ix <= 1) || (m.face_lo[1] && iy <= 1) || (m.face_lo[2] && iz <= 1) ||
           (m.face_hi[1] && iy >= m.gny - 2) || (m.face_hi[2] && iz >= m.gnz - 2);
  return m.face_hi[0] && ix >= m.gnx - 2;
}

// ref/src/SparseMatrix_functions.hpp:411-480 for one owned row (columns still GLOBAL rows)
MFE_HD void dirichlet_row(const MeshDesc &m, int which, double val, int64_t lr, int ix, int iy, int iz, int64_t off, int len,
                          const int *cols, double *coefs, double *b)
{
  const int row = m.first_row + (int)lr;
  const bool row_bc = which == 0 ? in_bc0(m, ix, iy, iz) : in_bc1(m, ix, iy, iz);
  if (row_bc) {                                     // first loop: zero_row_and_put_1_on_diagonal, b = value
    b[lr] = val;
    for (int j = 0; j < len; ++j) coefs[off + j] = (cols[off + j] == row) ? 1.0 : 0.0;
    return;
  }
  double sum = 0.0;                                 // second loop
  for (int j = 0; j < len; ++j) {
    int jx, jy, jz;
    row_to_node(m, cols[off + j], jx, jy, jz);
    const bool col_bc = which == 0 ? in_bc0(m, jx, jy, jz) : in_bc1(m, jx, jy, jz);
    if (col_bc) { sum = MFE_ADD(sum, coefs[off + j]); coefs[off + j] = 0.0; }
  }
  b[lr] = MFE_SUB(b[lr], MFE_MUL(sum, val));
}

}  // namespace mfe
