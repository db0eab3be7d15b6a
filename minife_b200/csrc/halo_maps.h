// halo_maps.h -- host-side index maps of make_local_matrix (ref/src/make_local_matrix.hpp:42-442), computed
// without messages.  The reference discovers externals by scanning rows (ascending) and their sorted columns,
// numbering each external at first encounter.  For the 27-point stencil that order is closed-form: an external
// node E is first met in the smallest adjacent owned row, first(E) = row(max(lo, E-1) per axis), and inside one
// row externals appear in ascending global row.  So external_index = externals sorted by (first(E), globalrow(E)).
// The peers' lists (-> elements_to_send) follow from the same rule applied to the peer's box.
#pragma once
#include <algorithm>
#include <system_error>
#include <thread>
#include <vector>
#include "mesh.h"

#ifndef MFE_HALO_THREADS
#define MFE_HALO_THREADS 1   /* helper threads for the neighbours' shells (measured: more than one loses to allocator contention) */
#endif

namespace mfe {

struct ExtEntry { int64_t first; int grow; int owner; };

struct LocalMaps {
  std::vector<int> external_index, external_local_index, elements_to_send, neighbors, recv_length, send_length;
  // not a reference field: where neighbour i stores what I send it, as an offset into ITS external tail (its
  // recv_length prefix up to me) -- lets the halo exchange store straight into the peer's memory over NVLink
  std::vector<int> send_dst_offset;
  int64_t num_cols = 0;
};

// Visit the externals of rank p (nodes of the one-node shell around p's node box, owned by someone else) in no particular
// order.  `gm.boxes` must be a HOST table.  fn(first, grow, owner): first = the smallest adjacent owned local row (the row in
// which the reference's scan meets the external first), grow = its global row, owner = its rank.
template <typename F>
inline void for_each_external(const MeshDesc &gm, int p, F fn)
{
  const RankBox &B = gm.boxes[p];
  if (B.nrows == 0) return;
  MeshDesc m = gm;
  m.rank = p;
  for (int i = 0; i < 6; ++i) m.nb[i] = B.nb[i];
  m.lnx = B.nb[1] - B.nb[0]; m.lny = B.nb[3] - B.nb[2]; m.lnz = B.nb[5] - B.nb[4];
  m.first_row = B.first_row; m.nrows = B.nrows;
  const int lo[3] = { std::max(B.nb[0] - 1, 0), std::max(B.nb[2] - 1, 0), std::max(B.nb[4] - 1, 0) };
  const int hi[3] = { std::min(B.nb[1] + 1, gm.gnx), std::min(B.nb[3] + 1, gm.gny), std::min(B.nb[5] + 1, gm.gnz) };
  int last = p;                         // neighbouring shell nodes mostly share their owner: try the last one first
  auto add = [&](int x, int y, int z) {
    const int fx = std::max(B.nb[0], x - 1), fy = std::max(B.nb[2], y - 1), fz = std::max(B.nb[4], z - 1);
    const int64_t first = (int64_t)(fx - B.nb[0]) + (int64_t)m.lnx * ((fy - B.nb[2]) + (int64_t)m.lny * (fz - B.nb[4]));
    // owner = the rank whose node box holds (x,y,z) (box_utils.hpp:233-315); global row = its first row + x-fastest index
    auto inside = [&](const RankBox &b) { return x >= b.nb[0] && x < b.nb[1] && y >= b.nb[2] && y < b.nb[3] && z >= b.nb[4] && z < b.nb[5]; };
    int q = last;
    if (!inside(gm.boxes[q])) { for (q = 0; q < gm.nranks; ++q) if (inside(gm.boxes[q])) break; }
    if (q >= gm.nranks) return;         // cannot happen for an in-range node of a partition that covers the mesh
    last = q;
    const RankBox &b = gm.boxes[q];
    const int grow = b.first_row + (x - b.nb[0]) + (b.nb[1] - b.nb[0]) * ((y - b.nb[2]) + (b.nb[3] - b.nb[2]) * (z - b.nb[4]));
    fn(first, grow, q);
  };
  for (int z = lo[2]; z < hi[2]; ++z)
    for (int y = lo[1]; y < hi[1]; ++y) {
      const bool yz_inside = y >= B.nb[2] && y < B.nb[3] && z >= B.nb[4] && z < B.nb[5];
      if (!yz_inside) { for (int x = lo[0]; x < hi[0]; ++x) add(x, y, z); }
      else {
        if (lo[0] < B.nb[0]) add(B.nb[0] - 1, y, z);
        if (hi[0] > B.nb[1]) add(B.nb[1], y, z);
      }
    }
}

inline bool ext_before(const ExtEntry &a, const ExtEntry &b) { return a.first != b.first ? a.first < b.first : a.grow < b.grow; }
// (first, grow) packed into one sortable word: both are below 2^31 (int32 ordinals)
inline uint64_t ext_key(int64_t first, int grow) { return ((uint64_t)first << 32) | (uint32_t)grow; }
inline int owner_of_row(const MeshDesc &gm, int grow)
{
  int owner = 0;
  for (int p = gm.nranks - 1; p >= 0; --p)
    if (gm.boxes[p].nrows > 0 && gm.boxes[p].first_row <= grow) { owner = p; break; }     // make_local_matrix.hpp:135-143
  return owner;
}
inline int64_t shell_bound(const RankBox &B)
{
  const int64_t lx = B.nb[1] - B.nb[0], ly = B.nb[3] - B.nb[2], lz = B.nb[5] - B.nb[4];
  return (lx + 2) * (ly + 2) * (lz + 2) - lx * ly * lz;
}

// externals of rank p in the reference's first-encounter order
inline void externals_of(const MeshDesc &gm, int p, std::vector<ExtEntry> &out)
{
  std::vector<uint64_t> keys;
  keys.reserve((size_t)shell_bound(gm.boxes[p]));
  for_each_external(gm, p, [&](int64_t first, int grow, int) { keys.push_back(ext_key(first, grow)); });
  std::sort(keys.begin(), keys.end());
  out.resize(keys.size());
  int last = 0;                         // consecutive externals mostly share their owner
  for (size_t i = 0; i < keys.size(); ++i) {
    const int grow = (int)(uint32_t)(keys[i] & 0xffffffffULL);
    const RankBox &b = gm.boxes[last];
    if (!(b.nrows > 0 && grow >= b.first_row && grow < b.first_row + b.nrows)) last = owner_of_row(gm, grow);
    out[i] = ExtEntry{ (int64_t)(keys[i] >> 32), grow, last };
  }
}

// What rank p receives from rank `me`, without sorting p's whole shell: its externals owned by `me` in p's encounter order
// (-> my elements_to_send for p) and the offset of that group inside p's external tail.  p groups its externals by owner
// in the order in which it first meets each owner (make_local_matrix.hpp:153-166), so the offset is the number of
// externals whose owner is first met before `me` is: one linear pass over the shell for the per-owner counts and
// first-encounter keys, one sort of my (small) group.
inline void peer_wants_from(const MeshDesc &gm, int p, int me, std::vector<ExtEntry> &mine, int &dst_offset)
{
  std::vector<uint64_t> keys;
  keys.reserve((size_t)shell_bound(gm.boxes[p]) / 2 + 64);
  std::vector<int> cnt(gm.nranks, 0);
  std::vector<uint64_t> firstkey(gm.nranks, ~0ULL);
  for_each_external(gm, p, [&](int64_t first, int grow, int owner) {
    const uint64_t k = ext_key(first, grow);
    ++cnt[owner];
    if (k < firstkey[owner]) firstkey[owner] = k;
    if (owner == me) keys.push_back(k);
  });
  std::sort(keys.begin(), keys.end());
  mine.resize(keys.size());
  for (size_t i = 0; i < keys.size(); ++i) mine[i] = ExtEntry{ (int64_t)(keys[i] >> 32), (int)(uint32_t)(keys[i] & 0xffffffffULL), me };
  dst_offset = 0;
  if (cnt[me] == 0) return;
  for (int q = 0; q < gm.nranks; ++q)
    if (q != me && cnt[q] > 0 && firstkey[q] < firstkey[me]) dst_offset += cnt[q];
}

// group by owner in first-encounter order of owners, encounter order inside a group (make_local_matrix.hpp:153-166)
inline void group_externals(const std::vector<ExtEntry> &ext, std::vector<int> &owners, std::vector<std::vector<int>> &groups)
{
  owners.clear(); groups.clear();
  std::vector<int> slot;                 // owner rank -> group index (or -1)
  for (size_t i = 0; i < ext.size(); ++i) {
    const int o = ext[i].owner;
    if (o >= (int)slot.size()) slot.resize(o + 1, -1);
    if (slot[o] < 0) { slot[o] = (int)owners.size(); owners.push_back(o); groups.emplace_back(); groups.back().reserve(ext.size() / 4 + 16); }
    groups[slot[o]].push_back((int)i);
  }
}

// all CSRMatrix index-map fields (ref/src/CSRMatrix.hpp:73-78) of rank gm.rank
inline void compute_local_maps(const MeshDesc &gm, LocalMaps &L)
{
  L = LocalMaps();
  const int rank = gm.rank;
  const int nrows = gm.boxes[rank].nrows, first_row = gm.boxes[rank].first_row;
  // The ranks whose node box touches my one-node shell are exactly my neighbours (the stencil is symmetric); what each of
  // them needs from me is one pass over ITS shell (make_local_matrix.hpp:386-429), independent of everything else: those
  // passes run on a few host threads while this thread orders my own externals.
  const RankBox &B = gm.boxes[rank];
  std::vector<int> cand;
  for (int q = 0; q < gm.nranks; ++q) {
    const RankBox &Q = gm.boxes[q];
    if (q == rank || Q.nrows == 0) continue;
    bool touch = true;
    for (int a = 0; a < 3; ++a) touch = touch && Q.nb[2 * a] < B.nb[2 * a + 1] + 1 && Q.nb[2 * a + 1] > B.nb[2 * a] - 1;
    if (touch) cand.push_back(q);
  }
  const size_t nc = cand.size();
  std::vector<std::vector<ExtEntry>> mine(nc);
  std::vector<int> poff(nc, 0);
  struct Pool {                          // joins on every way out of this function (an exception below must not reach ~thread)
    std::vector<std::thread> th;
    void join() { for (std::thread &t : th) if (t.joinable()) t.join(); }
    ~Pool() { join(); }
  } pool;
  {
    size_t nthreads = std::min<size_t>(nc, MFE_HALO_THREADS);
    try {
      for (size_t t = 0; t < nthreads; ++t)
        pool.th.emplace_back([&, t, nthreads]() { for (size_t i = t; i < nc; i += nthreads) peer_wants_from(gm, cand[i], rank, mine[i], poff[i]); });     // (nthreads by value: the threads outlive this block)
    } catch (const std::system_error &) {  // no thread to be had: do the passes here
      pool.join(); pool.th.clear(); nthreads = 0;
    }
    if (nthreads == 0) for (size_t i = 0; i < nc; ++i) peer_wants_from(gm, cand[i], rank, mine[i], poff[i]);
  }
  std::vector<ExtEntry> ext;
  externals_of(gm, rank, ext);
  std::vector<int> owners; std::vector<std::vector<int>> groups;
  group_externals(ext, owners, groups);
  const int n_ext = (int)ext.size();
  L.external_index.resize(n_ext); L.external_local_index.assign(n_ext, -1);
  for (int i = 0; i < n_ext; ++i) L.external_index[i] = ext[i].grow;
  int count = nrows;
  for (size_t g = 0; g < owners.size(); ++g) {
    L.neighbors.push_back(owners[g]);
    L.recv_length.push_back((int)groups[g].size());
    for (int i : groups[g]) L.external_local_index[i] = count++;
  }
  pool.join();
  const size_t nn = L.neighbors.size();
  L.send_length.assign(nn, 0);
  L.send_dst_offset.assign(nn, 0);
  for (size_t i = 0; i < nn; ++i) {
    size_t ci = 0;
    while (ci < nc && cand[ci] != L.neighbors[i]) ++ci;
    if (ci == nc) continue;               // cannot happen: a neighbour's box touches my shell
    L.send_length[i] = (int)mine[ci].size();
    L.send_dst_offset[i] = poff[ci];
    for (const ExtEntry &e : mine[ci]) L.elements_to_send.push_back(e.grow - first_row);
  }
  L.num_cols = (int64_t)nrows + n_ext;
}

}  // namespace mfe
