// halo_maps.h -- host-side index maps of make_local_matrix (ref/src/make_local_matrix.hpp:42-442), computed
// without messages and without sorting.  The reference discovers externals by scanning rows (ascending) and their
// sorted columns, numbering each external at first encounter.  For the 27-point stencil that order is closed-form:
// an external node E is first met in the smallest adjacent owned row, first(E) = row(max(lo, E-1) per axis), and
// inside one row externals appear in ascending global row.  So external_index = externals ordered by
// (first(E), globalrow(E)) -- which can be ENUMERATED in that order:
//   * my own externals: walk my boundary rows in ascending order; the externals first met in row r are the nodes E of
//     its stencil with max(lo, E-1) == r on every axis (at most 27, usually one), ordered by global row;
//   * what a neighbour p receives from me (-> my elements_to_send): p's externals inside my node box form a
//     rectangular region; on each axis the nodes of that region fall into runs of equal first-row coordinate (one
//     run of up to three nodes at p's low side, single nodes after it), so the region is walked run by run;
//   * where my values land in p's external tail (p groups its externals by owner in first-encounter order of the owners,
//     make_local_matrix.hpp:153-166): the regions of p's other neighbours are boxes too, each first met at its lowest
//     corner, so the offset is a sum of box volumes.
// Everything is O(number of externals), most of it straight fills of consecutive rows.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>
#include "mesh.h"

namespace mfe {

struct LocalMaps {
  std::vector<int> external_index, external_local_index, elements_to_send, neighbors, recv_length, send_length;
  // not a reference field: where neighbour i stores what I send it, as an offset into ITS external tail (its
  // recv_length prefix up to me) -- lets the halo exchange store straight into the peer's memory over NVLink
  std::vector<int> send_dst_offset;
  // not reference fields either: the externals by ascending global row with their local column, the lookup table
  // with which the columns are rewritten (the reference uses a std::map, make_local_matrix.hpp:98-117)
  std::vector<int> sorted_global, sorted_local;
  int64_t num_cols = 0;
};

// The externals of rank p owned by rank q: the nodes of q's node box inside the one-node shell around p's node box.
// Boxes of a partition are disjoint, so this is simply (p's box grown by one) intersected with q's box.
struct ShellRegion {
  int q;
  int lo[3], hi[3];
  int64_t volume() const { return (int64_t)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]); }
  bool holds(int x, int y, int z) const { return x >= lo[0] && x < hi[0] && y >= lo[1] && y < hi[1] && z >= lo[2] && z < hi[2]; }
  // x-fastest ordinal inside the region (ascending global row: the region lies in one owner's box)
  int64_t ordinal(int x, int y, int z) const { return (x - lo[0]) + (int64_t)(hi[0] - lo[0]) * ((y - lo[1]) + (int64_t)(hi[1] - lo[1]) * (z - lo[2])); }
};

inline bool shell_region(const MeshDesc &gm, int p, int q, ShellRegion &R)
{
  const RankBox &B = gm.boxes[p], &Q = gm.boxes[q];
  if (p == q || B.nrows <= 0 || Q.nrows <= 0) return false;
  R.q = q;
  for (int a = 0; a < 3; ++a) {
    R.lo[a] = std::max(Q.nb[2 * a], B.nb[2 * a] - 1);
    R.hi[a] = std::min(Q.nb[2 * a + 1], B.nb[2 * a + 1] + 1);
    if (R.lo[a] >= R.hi[a]) return false;
  }
  return true;
}

// global row of node (x,y,z) of box b (ref/utils/utils.hpp:165-199: owner's first row + x-fastest index in its box)
inline int row_in_box(const RankBox &b, int x, int y, int z)
{
  return b.first_row + (x - b.nb[0]) + (b.nb[1] - b.nb[0]) * ((y - b.nb[2]) + (b.nb[3] - b.nb[2]) * (z - b.nb[4]));
}

// local row of p in which its scan meets external (x,y,z) first: the smallest adjacent owned row
inline int64_t first_row_meeting(const RankBox &B, int x, int y, int z)
{
  const int fx = std::max(B.nb[0], x - 1), fy = std::max(B.nb[2], y - 1), fz = std::max(B.nb[4], z - 1);
  return (int64_t)(fx - B.nb[0]) + (int64_t)(B.nb[1] - B.nb[0]) * ((fy - B.nb[2]) + (int64_t)(B.nb[3] - B.nb[2]) * (fz - B.nb[4]));
}

// (first, grow) packed into one comparable word: both are below 2^31 (int32 ordinals)
inline uint64_t ext_key(int64_t first, int grow) { return ((uint64_t)first << 32) | (uint32_t)grow; }

// On one axis, the stencil nodes E of owned coordinate r for which r is the first-meeting coordinate (max(lo, E-1) == r):
// lo-1, lo, lo+1 when r is the low end of the box, r+1 otherwise -- as far as the mesh reaches.  out[i]: E lies outside the box.
struct AxisStencil {
  int n = 0;
  int e[3];
  bool out[3];
  int nout = 0, first_out = 0;
  bool any_out = false;
  AxisStencil(int r, int lo, int hi, int gn)
  {
    const int e0 = r == lo ? std::max(lo - 1, 0) : r + 1, e1 = std::min(r == lo ? lo + 2 : r + 2, gn);
    for (int v = e0; v < e1; ++v) {
      e[n] = v; out[n] = v < lo || v >= hi;
      if (out[n] && !nout++) first_out = n;
      ++n;
    }
    any_out = nout > 0;
  }
};

// Visit the externals of rank gm.rank in the reference's first-encounter order.  `regions`: its shell regions, one per
// neighbouring owner.  fn(region index, x, y, z, len) for every run of externals (x..x+len-1, y, z) that are consecutive in
// that order (and lie in one region, so their global rows are consecutive too).
template <typename F>
inline void for_each_external_in_order(const MeshDesc &gm, const std::vector<ShellRegion> &regions, F fn)
{
  const RankBox &B = gm.boxes[gm.rank];
  if (B.nrows <= 0 || regions.empty()) return;
  const int lo[3] = { B.nb[0], B.nb[2], B.nb[4] }, hi[3] = { B.nb[1], B.nb[3], B.nb[5] }, gn[3] = { gm.gnx, gm.gny, gm.gnz };
  struct Met { int grow, region, x, y, z; };
  Met met[27];
  size_t last = 0;                       // neighbouring externals mostly share their owner: try the last region first
  auto region_of = [&](int ex, int ey, int ez) -> size_t {
    if (regions[last].holds(ex, ey, ez)) return last;
    for (size_t ri = 0; ri < regions.size(); ++ri) if (regions[ri].holds(ex, ey, ez)) return last = ri;
    return regions.size();               // cannot happen for a partition that covers the mesh
  };
  // the externals first met in row (x,y,z), in ascending global row
  auto row = [&](const AxisStencil &sx, const AxisStencil &sy, const AxisStencil &sz) {
    int nmet = 0;
    for (int iz = 0; iz < sz.n; ++iz)
      for (int iy = 0; iy < sy.n; ++iy)
        for (int ix = 0; ix < sx.n; ++ix) {
          if (!(sz.out[iz] || sy.out[iy] || sx.out[ix])) continue;
          const int ex = sx.e[ix], ey = sy.e[iy], ez = sz.e[iz];
          const size_t ri = region_of(ex, ey, ez);
          if (ri >= regions.size()) continue;
          met[nmet++] = Met{ row_in_box(gm.boxes[regions[ri].q], ex, ey, ez), (int)ri, ex, ey, ez };
        }
    for (int i = 1; i < nmet; ++i) {      // (already ascending within one owner)
      const Met m = met[i];
      int j = i;
      for (; j > 0 && met[j - 1].grow > m.grow; --j) met[j] = met[j - 1];
      met[j] = m;
    }
    for (int i = 0; i < nmet; ++i) fn(met[i].region, met[i].x, met[i].y, met[i].z, 1);
  };
  const AxisStencil sx_lo(lo[0], lo[0], hi[0], gn[0]), sx_hi(hi[0] - 1, lo[0], hi[0], gn[0]);
  const bool two_ends = hi[0] - 1 > lo[0];
  const bool x_ends_out = sx_lo.any_out || (two_ends && sx_hi.any_out);
  std::vector<AxisStencil> sys;
  sys.reserve((size_t)(hi[1] - lo[1]));
  for (int y = lo[1]; y < hi[1]; ++y) sys.emplace_back(y, lo[1], hi[1], gn[1]);
  // one end row of an x line whose (y,z) stencil lies inside the box: externals only through its x stencil
  auto end_row = [&](const AxisStencil &sx, const AxisStencil &sy, const AxisStencil &sz) {
    if (!sx.any_out) return;
    if (sx.nout == 1 && sy.n == 1 && sz.n == 1) {
      const size_t ri = region_of(sx.e[sx.first_out], sy.e[0], sz.e[0]);
      if (ri < regions.size()) fn((int)ri, sx.e[sx.first_out], sy.e[0], sz.e[0], 1);
    } else row(sx, sy, sz);
  };
  for (int z = lo[2]; z < hi[2]; ++z) {
    const AxisStencil sz(z, lo[2], hi[2], gn[2]);
    if (!sz.n) continue;
    // a plane away from the z faces of a box without x neighbours meets externals only in its first and last y line
    const int ystep = (sz.any_out || x_ends_out) ? 1 : std::max(hi[1] - 1 - lo[1], 1);
    for (int y = lo[1]; y < hi[1]; y += ystep) {
      const AxisStencil &sy = sys[y - lo[1]];
      if (!sy.n) continue;
      // the (y,z) stencil nodes that lie outside the box in y or z: every row of this x line meets them
      int nout = 0, oy = 0, oz = 0;
      if (sz.any_out || sy.any_out)
        for (int iz = 0; iz < sz.n; ++iz)
          for (int iy = 0; iy < sy.n; ++iy)
            if (sz.out[iz] || sy.out[iy]) { ++nout; oy = sy.e[iy]; oz = sz.e[iz]; }
      if (nout == 0) {
        end_row(sx_lo, sy, sz);
        if (two_ends) end_row(sx_hi, sy, sz);
        continue;
      }
      if (sx_lo.n) row(sx_lo, sy, sz);
      if (nout == 1) {
        // rows lo+1 .. hi-2 each meet exactly (x+1, oy, oz): runs of consecutive externals, cut where the owner changes
        for (int ex = lo[0] + 2; ex < hi[0]; ) {
          const size_t ri = region_of(ex, oy, oz);
          if (ri >= regions.size()) { ++ex; continue; }
          const int end = std::min(hi[0], regions[ri].hi[0]);
          fn((int)ri, ex, oy, oz, end - ex);
          ex = end;
        }
      } else {
        for (int x = lo[0] + 1; x < hi[0] - 1; ++x) row(AxisStencil(x, lo[0], hi[0], gn[0]), sy, sz);
      }
      if (two_ends && sx_hi.n) row(sx_hi, sy, sz);
    }
  }
}

// Where p first meets one of its shell regions: the region's lowest corner (first_row_meeting and the global row both
// grow with every coordinate inside a region).  p numbers its groups of externals in ascending order of this key.
inline uint64_t region_first_key(const MeshDesc &gm, int p, const ShellRegion &R)
{
  return ext_key(first_row_meeting(gm.boxes[p], R.lo[0], R.lo[1], R.lo[2]), row_in_box(gm.boxes[R.q], R.lo[0], R.lo[1], R.lo[2]));
}

// What rank p receives from rank `me`: p's externals owned by `me` in p's encounter order, as local rows of `me`, appended
// to `rows` (-> my elements_to_send for p, make_local_matrix.hpp:386-429), and the offset of that group inside p's external tail.
inline void peer_wants_from(const MeshDesc &gm, int p, int me, std::vector<int> &rows, int &dst_offset)
{
  dst_offset = 0;
  ShellRegion R;
  if (!shell_region(gm, p, me, R)) return;
  const RankBox &B = gm.boxes[p], &M = gm.boxes[me];
  // runs of equal first-meeting coordinate on each axis: [R.lo, lo_p + 2) collapses onto p's low end, then single nodes
  struct Run { int b, e; };
  std::vector<Run> run[3];
  for (int a = 0; a < 3; ++a) {
    int e0 = R.lo[a];
    const int t = std::min(R.hi[a], B.nb[2 * a] + 2);
    if (e0 < t) { run[a].push_back(Run{ e0, t }); e0 = t; }
    for (; e0 < R.hi[a]; ++e0) run[a].push_back(Run{ e0, e0 + 1 });
  }
  size_t at = rows.size();
  rows.resize(at + (size_t)R.volume());
  int *out = rows.data();
  for (const Run &rz : run[2])
    for (const Run &ry : run[1])
      for (const Run &rx : run[0])
        for (int z = rz.b; z < rz.e; ++z)
          for (int y = ry.b; y < ry.e; ++y) {
            int r = row_in_box(M, rx.b, y, z) - M.first_row;
            for (int x = rx.b; x < rx.e; ++x) out[at++] = r++;
          }
  // p's groups come in the order in which p first meets each owner
  const uint64_t my_key = region_first_key(gm, p, R);
  int64_t off = 0;
  for (int q = 0; q < gm.nranks; ++q) {
    ShellRegion S;
    if (q != me && shell_region(gm, p, q, S) && region_first_key(gm, p, S) < my_key) off += S.volume();
  }
  dst_offset = (int)off;
}

// all CSRMatrix index-map fields (ref/src/CSRMatrix.hpp:73-78) of rank gm.rank.  `gm.boxes` must be a HOST table.
inline void compute_local_maps(const MeshDesc &gm, LocalMaps &L)
{
  L = LocalMaps();
  const int rank = gm.rank;
  const int nrows = gm.boxes[rank].nrows;
  // The ranks whose node box touches my one-node shell are exactly my neighbours (the stencil is symmetric), in rank order
  std::vector<ShellRegion> regions;
  for (int q = 0; q < gm.nranks; ++q) {
    ShellRegion R;
    if (shell_region(gm, rank, q, R)) regions.push_back(R);
  }
  const int nreg = (int)regions.size();
  // neighbours in the order in which my scan first meets them; their externals take the local columns nrows onwards,
  // grouped by owner in that order, encounter order inside a group (make_local_matrix.hpp:153-166)
  std::vector<int> by_first(nreg);
  std::vector<uint64_t> key(nreg);
  for (int i = 0; i < nreg; ++i) { by_first[i] = i; key[i] = region_first_key(gm, rank, regions[i]); }
  std::sort(by_first.begin(), by_first.end(), [&](int a, int b) { return key[a] < key[b]; });
  std::vector<int> next(nreg, 0);                        // region -> next local column of its group
  std::vector<int64_t> region_base(nreg + 1, 0);         // region -> its place in the ascending-global-row table (rank order)
  for (int i = 0; i < nreg; ++i) region_base[i + 1] = region_base[i] + regions[i].volume();
  const int n_ext = (int)region_base[nreg];
  {
    int at = nrows;
    for (int g = 0; g < nreg; ++g) {
      const ShellRegion &R = regions[by_first[g]];
      L.neighbors.push_back(R.q);
      L.recv_length.push_back((int)R.volume());
      next[by_first[g]] = at; at += (int)R.volume();
    }
  }
  L.external_index.resize(n_ext); L.external_local_index.resize(n_ext);
  L.sorted_global.resize(n_ext); L.sorted_local.resize(n_ext);
  int *ext_index = L.external_index.data(), *ext_local = L.external_local_index.data(), *sorted_local = L.sorted_local.data();
  int count = 0;
  for_each_external_in_order(gm, regions, [&](int ri, int x, int y, int z, int len) {
    if (count + len > n_ext) return;                     // cannot happen: the regions are exactly the shell
    const int grow = row_in_box(gm.boxes[regions[ri].q], x, y, z);
    int *sl = sorted_local + region_base[ri] + regions[ri].ordinal(x, y, z);
    int lc = next[ri];
    for (int i = 0; i < len; ++i) { ext_index[count + i] = grow + i; ext_local[count + i] = lc + i; sl[i] = lc + i; }
    next[ri] = lc + len; count += len;
  });
  // ascending global row: regions in rank order (= first_row order), nodes x-fastest inside
  {
    int *sg = L.sorted_global.data();
    for (int i = 0; i < nreg; ++i) {
      const ShellRegion &R = regions[i];
      for (int z = R.lo[2]; z < R.hi[2]; ++z)
        for (int y = R.lo[1]; y < R.hi[1]; ++y) {
          int grow = row_in_box(gm.boxes[R.q], R.lo[0], y, z);
          for (int x = R.lo[0]; x < R.hi[0]; ++x) *sg++ = grow++;
        }
    }
  }
  // what each neighbour needs from me is its shell region inside my box (make_local_matrix.hpp:386-429), in ITS order
  const size_t nn = L.neighbors.size();
  L.send_length.assign(nn, 0);
  L.send_dst_offset.assign(nn, 0);
  L.elements_to_send.reserve((size_t)n_ext);
  for (size_t i = 0; i < nn; ++i) {
    const size_t before = L.elements_to_send.size();
    peer_wants_from(gm, L.neighbors[i], rank, L.elements_to_send, L.send_dst_offset[i]);
    L.send_length[i] = (int)(L.elements_to_send.size() - before);
  }
  L.num_cols = (int64_t)nrows + n_ext;
}

}  // namespace mfe
