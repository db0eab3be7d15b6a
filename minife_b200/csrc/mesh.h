// mesh.h -- analytic replacement of miniFE's simple_mesh_description (ref/src/simple_mesh_description.hpp:41-226):
// node id <-> global row, BC-set membership and CSR row offsets are closed-form functions of the box table,
// so neither std::map (map_ids_to_rows) nor std::set (bc_rows_*) nor a prefix scan is needed on the device.
// Shared by host and device code (MFE_HD).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MFE_HD __host__ __device__ __forceinline__
#else
#define MFE_HD inline
#endif

namespace mfe {

// One rank's owned-node box [lo,hi) per axis and the global row number of its first node
// (ref/utils/box_utils.hpp:243-264: exclusive prefix sum of owned-node counts).
struct RankBox {
  int nb[6];
  int first_row;
  int nrows;
};

struct MeshDesc {
  int nx, ny, nz;        // global element counts
  int gnx, gny, gnz;     // global node counts (= +1)
  int nranks, rank;
  int eb[6];             // my element box (ref Box::ranges order: x0,x1,y0,y1,z0,z1)
  int nb[6];             // my owned-node box
  int lnx, lny, lnz;     // extents of nb
  int first_row;
  int nrows;
  int face_lo[3];        // my ELEMENT box touches the global low / high face of the axis
  int face_hi[3];        //   (the `if we're on the x=0 face` tests, simple_mesh_description.hpp:78,102,...)
  const RankBox *boxes;  // nranks entries (device pointer in kernels, host pointer on the host)
};

// ref/utils/box_utils.hpp:99-111 get_id restricted to valid coordinates; ref/utils/utils.hpp:165-199
// find_row_for_id: row = owner's first_row + x-fastest index inside the owner's node box.
MFE_HD int node_to_row(const MeshDesc &m, int ix, int iy, int iz)
{
  if (ix >= m.nb[0] && ix < m.nb[1] && iy >= m.nb[2] && iy < m.nb[3] && iz >= m.nb[4] && iz < m.nb[5])
    return m.first_row + (ix - m.nb[0]) + m.lnx * ((iy - m.nb[2]) + m.lny * (iz - m.nb[4]));
  for (int p = 0; p < m.nranks; ++p) {
    const RankBox &b = m.boxes[p];
    if (ix >= b.nb[0] && ix < b.nb[1] && iy >= b.nb[2] && iy < b.nb[3] && iz >= b.nb[4] && iz < b.nb[5])
      return b.first_row + (ix - b.nb[0]) + (b.nb[1] - b.nb[0]) * ((iy - b.nb[2]) + (b.nb[3] - b.nb[2]) * (iz - b.nb[4]));
  }
  return -99;   // the value find_row_for_id reports for an unmappable id
}

// inverse map: global row -> node coordinates (owner = highest rank whose first_row <= row,
// ref/src/make_local_matrix.hpp:135-143)
MFE_HD int row_to_node(const MeshDesc &m, int row, int &ix, int &iy, int &iz)
{
  int owner = m.rank;
  if (row < m.first_row || row >= m.first_row + m.nrows) {
    owner = -1;
    for (int p = m.nranks - 1; p >= 0; --p)
      if (m.boxes[p].nrows > 0 && m.boxes[p].first_row <= row) { owner = p; break; }
    if (owner < 0) { ix = iy = iz = -1; return -1; }
  }
  const RankBox &b = m.boxes[owner];
  int l = row - b.first_row, lx = b.nb[1] - b.nb[0], ly = b.nb[3] - b.nb[2];
  ix = b.nb[0] + l % lx;
  iy = b.nb[2] + (l / lx) % ly;
  iz = b.nb[4] + l / (lx * ly);
  return owner;
}

// Membership of a node in THIS rank's bc_rows_0 / bc_rows_1 (simple_mesh_description.hpp:78-224).  A face
// contributes only when the rank's element box touches it; the reference's ghost-layer ranges always cover
// the columns of owned rows, so no range test is needed for those.
MFE_HD bool in_bc0(const MeshDesc &m, int ix, int iy, int iz)
{
  return (m.face_lo[0] && ix == 0) || (m.face_lo[1] && iy == 0) || (m.face_lo[2] && iz == 0) ||
         (m.face_hi[1] && iy == m.gny - 1) || (m.face_hi[2] && iz == m.gnz - 1);
}
MFE_HD bool in_bc1(const MeshDesc &m, int ix, int /*iy*/, int /*iz*/)
{
  return m.face_hi[0] && ix == m.gnx - 1;
}

// number of in-range stencil neighbours along one axis (generate_matrix_structure.hpp:109-121)
MFE_HD int axis_count(int i, int gn) { return 3 - (i == 0) - (i == gn - 1); }
// sum of axis_count over [lo, i)
MFE_HD int64_t axis_prefix(int lo, int i, int gn)
{
  if (i <= lo) return 0;
  int64_t s = 3LL * (i - lo);
  if (lo == 0) s -= 1;              // node 0 lies in [lo,i)
  if (i - 1 >= gn - 1) s -= 1;      // node gn-1 lies in [lo,i)
  return s;
}
// CSR offset of the owned node (ix,iy,iz): nodes are numbered x fastest, a row holds cx*cy*cz entries
MFE_HD int64_t row_offset_of(const MeshDesc &m, int ix, int iy, int iz)
{
  const int64_t Sx = axis_prefix(m.nb[0], m.nb[1], m.gnx), Sy = axis_prefix(m.nb[2], m.nb[3], m.gny);
  return Sx * Sy * axis_prefix(m.nb[4], iz, m.gnz) +
         Sx * axis_prefix(m.nb[2], iy, m.gny) * axis_count(iz, m.gnz) +
         axis_prefix(m.nb[0], ix, m.gnx) * axis_count(iy, m.gny) * axis_count(iz, m.gnz);
}
MFE_HD int64_t total_nnz_of(const MeshDesc &m)
{
  return axis_prefix(m.nb[0], m.nb[1], m.gnx) * axis_prefix(m.nb[2], m.nb[3], m.gny) * axis_prefix(m.nb[4], m.nb[5], m.gnz);
}

MFE_HD void local_row_to_node(const MeshDesc &m, int64_t lr, int &ix, int &iy, int &iz)
{
  ix = m.nb[0] + (int)(lr % m.lnx);
  iy = m.nb[2] + (int)((lr / m.lnx) % m.lny);
  iz = m.nb[4] + (int)(lr / ((int64_t)m.lnx * m.lny));
}

// true when every in-range stencil neighbour of the node is owned by this rank: the natural (sz,sy,sx)
// order of the columns is then already ascending (MatrixInitOp.hpp:40-54 sort_if_needed is a no-op)
MFE_HD bool neighbourhood_is_local(const MeshDesc &m, int ix, int iy, int iz)
{
  return (ix - 1 < 0 || ix - 1 >= m.nb[0]) && (ix + 1 >= m.gnx || ix + 1 < m.nb[1]) &&
         (iy - 1 < 0 || iy - 1 >= m.nb[2]) && (iy + 1 >= m.gny || iy + 1 < m.nb[3]) &&
         (iz - 1 < 0 || iz - 1 >= m.nb[4]) && (iz + 1 >= m.gnz || iz + 1 < m.nb[5]);
}

// Columns of one row: fills cols[] (global rows, ascending) and sidx[] (stencil slot (sz+1)*9+(sy+1)*3+(sx+1)
// stored at each sorted position); returns the row length.
MFE_HD int row_columns(const MeshDesc &m, int ix, int iy, int iz, int *cols, unsigned char *sidx)
{
  int len = 0;
  for (int sz = -1; sz <= 1; ++sz)
    for (int sy = -1; sy <= 1; ++sy)
      for (int sx = -1; sx <= 1; ++sx) {
        int jx = ix + sx, jy = iy + sy, jz = iz + sz;
        if (jx < 0 || jy < 0 || jz < 0 || jx >= m.gnx || jy >= m.gny || jz >= m.gnz) continue;
        cols[len] = node_to_row(m, jx, jy, jz);
        sidx[len] = (unsigned char)((sz + 1) * 9 + (sy + 1) * 3 + (sx + 1));
        ++len;
      }
  if (!neighbourhood_is_local(m, ix, iy, iz)) {
    for (int i = 1; i < len; ++i) {          // insertion sort, ascending global row
      int c = cols[i]; unsigned char s = sidx[i]; int j = i;
      for (; j > 0 && cols[j - 1] > c; --j) { cols[j] = cols[j - 1]; sidx[j] = sidx[j - 1]; }
      cols[j] = c; sidx[j] = s;
    }
  }
  return len;
}

}  // namespace mfe
