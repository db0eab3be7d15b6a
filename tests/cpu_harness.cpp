// tests/cpu_harness.cpp -- TEST INFRASTRUCTURE.  Runs the host+device-shared (MFE_HD) building blocks of the CUDA
// kernels on the CPU, thread by thread, so that their logic (analytic mesh maps, element arithmetic order,
// row-owner gather, Dirichlet rows, halo index maps, order-reproducing dot) can be checked against the oracle
// without a GPU.  It is NOT a product path: nothing in minife_b200/ links or loads it, and the library has
// no CPU fallback.  Built by tests/conftest.py:  g++ -O2 -ffp-contract=off -shared -fPIC.
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "../minife_b200/csrc/mesh.h"
#include "../minife_b200/csrc/hex8.h"
#include "../minife_b200/csrc/assembly.h"
#include "../minife_b200/csrc/exact_dot.h"
#include "../minife_b200/csrc/halo_maps.h"

using namespace mfe;

struct Harness {
  MeshDesc m;
  std::vector<RankBox> boxes;
  std::vector<int> rows, cols;
  std::vector<int64_t> roff;
  std::vector<double> coefs, b;
  LocalMaps maps;
};

static void rcb(int ip, int up, int axis, const int *box, int *out)
{
  const int np = up - ip;
  if (np == 1) { for (int i = 0; i < 6; ++i) out[6 * ip + i] = box[i]; return; }
  const int n = box[2 * axis + 1] - box[2 * axis];
  const int np_low = np / 2, np_upp = np - np_low;
  const int n_upp = (int)(((double)n) * (((double)np_upp) / ((double)np)));
  const int n_low = n - n_upp, next = (axis + 2) % 3;
  int lowb[6], uppb[6];
  for (int i = 0; i < 6; ++i) lowb[i] = uppb[i] = box[i];
  lowb[2 * axis + 1] = box[2 * axis] + n_low;
  uppb[2 * axis] = box[2 * axis] + n_low;
  uppb[2 * axis + 1] = uppb[2 * axis] + n_upp;
  if (np_low) rcb(ip, ip + np_low, next, lowb, out);
  rcb(ip + np_low, up, next, uppb, out);
}

extern "C" {

void *hs_create_boxes(int nx, int ny, int nz, int np, int rank, const int *boxes6np);
void *hs_create(int nx, int ny, int nz, int np, int rank) { return hs_create_boxes(nx, ny, nz, np, rank, nullptr); }
// boxes6np == nullptr: box_partition; otherwise caller-supplied element boxes (e.g. after add_imbalance)
void *hs_create_boxes(int nx, int ny, int nz, int np, int rank, const int *boxes6np)
{
  Harness *h = new Harness;
  std::vector<int> all(6 * np);
  const int g[6] = { 0, nx, 0, ny, 0, nz };
  if (boxes6np) all.assign(boxes6np, boxes6np + 6 * np);
  else rcb(0, np, 2, g, all.data());
  MeshDesc &m = h->m;
  m.nx = nx; m.ny = ny; m.nz = nz; m.gnx = nx + 1; m.gny = ny + 1; m.gnz = nz + 1;
  m.nranks = np; m.rank = rank;
  const int gg[3] = { nx, ny, nz };
  h->boxes.resize(np);
  int off = 0;
  for (int p = 0; p < np; ++p) {
    RankBox &b = h->boxes[p];
    const int *eb = &all[6 * p];
    for (int a = 0; a < 3; ++a) {
      b.nb[2 * a] = eb[2 * a];
      b.nb[2 * a + 1] = eb[2 * a + 1] + ((eb[2 * a + 1] > eb[2 * a] && eb[2 * a + 1] == gg[a]) ? 1 : 0);
    }
    b.nrows = (b.nb[1] - b.nb[0]) * (b.nb[3] - b.nb[2]) * (b.nb[5] - b.nb[4]);
    b.first_row = off; off += b.nrows;
  }
  const int *eb = &all[6 * rank];
  for (int i = 0; i < 6; ++i) { m.eb[i] = eb[i]; m.nb[i] = h->boxes[rank].nb[i]; }
  m.lnx = m.nb[1] - m.nb[0]; m.lny = m.nb[3] - m.nb[2]; m.lnz = m.nb[5] - m.nb[4];
  m.first_row = h->boxes[rank].first_row; m.nrows = h->boxes[rank].nrows;
  for (int a = 0; a < 3; ++a) { m.face_lo[a] = eb[2 * a] == 0; m.face_hi[a] = eb[2 * a + 1] == gg[a]; }
  m.boxes = h->boxes.data();
  return h;
}
void hs_destroy(void *p) { delete (Harness *)p; }
long long hs_nrows(void *p) { return ((Harness *)p)->m.nrows; }
long long hs_nnz(void *p) { return total_nnz_of(((Harness *)p)->m); }

// k_structure, one "thread" per row
void hs_structure(void *p)
{
  Harness *h = (Harness *)p; const MeshDesc &m = h->m;
  const int64_t nnz = total_nnz_of(m);
  h->rows.assign(m.nrows, 0); h->roff.assign(m.nrows + 1, 0); h->cols.assign(nnz, -7); h->coefs.assign(nnz, 0.0); h->b.assign(m.nrows, 0.0);
  for (int64_t lr = 0; lr < m.nrows; ++lr) {
    int ix, iy, iz; local_row_to_node(m, lr, ix, iy, iz);
    const int64_t off = row_offset_of(m, ix, iy, iz);
    h->rows[lr] = m.first_row + (int)lr; h->roff[lr] = off;
    int c[27]; unsigned char s[27];
    const int len = row_columns(m, ix, iy, iz, c, s);
    for (int j = 0; j < len; ++j) h->cols[off + j] = c[j];
    if (lr == m.nrows - 1) h->roff[m.nrows] = off + len;
  }
}

// k_elements + k_gather_rows, slabs of `planes` node planes
void hs_assemble(void *p, int planes)
{
  Harness *h = (Harness *)p; const MeshDesc &m = h->m;
  double grad[192], psi[64];
  hex8_tables(grad, psi);
  int eb[6]; for (int i = 0; i < 6; ++i) eb[i] = m.eb[i];
  if (eb[0] > 0) eb[0]--; if (eb[2] > 0) eb[2]--; if (eb[4] > 0) eb[4]--;
  if (eb[1] < m.nx) eb[1]++; if (eb[3] < m.ny) eb[3]++; if (eb[5] < m.nz) eb[5]++;
  ElemSlab sl; sl.ex0 = eb[0]; sl.ey0 = eb[2]; sl.nex = eb[1] - eb[0]; sl.ney = eb[3] - eb[2];
  const int64_t eplane = (int64_t)sl.nex * sl.ney;
  if (planes < 1) planes = 1;
  sl.stride = eplane * (planes + 1);
  std::vector<double> KF(sl.stride * 44);
  for (int iz0 = m.nb[4]; iz0 < m.nb[5]; iz0 += planes) {
    const int iz1 = std::min(iz0 + planes, m.nb[5]);
    const int ez0 = std::max(iz0 - 1, 0), ez1 = std::min(iz1, m.nz);
    sl.ez0 = ez0; sl.nez = ez1 - ez0;
    const int64_t ne = eplane * sl.nez;
    for (int64_t e = 0; e < ne; ++e) {
      const int ex = sl.ex0 + (int)(e % sl.nex), ey = sl.ey0 + (int)((e / sl.nex) % sl.ney), ez = sl.ez0 + (int)(e / eplane);
      const double x = (double)ex / (double)m.nx, y = (double)ey / (double)m.ny, z = (double)ez / (double)m.nz;
      const double hx = 1.0 / (double)m.nx, hy = 1.0 / (double)m.ny, hz = 1.0 / (double)m.nz;
      double K[36], f[8];
      hex8_element(grad, psi, x, y, z, hx, hy, hz, K, f);
      for (int k = 0; k < 36; ++k) KF[(int64_t)k * sl.stride + e] = K[k];
      for (int k = 0; k < 8; ++k) KF[(int64_t)(36 + k) * sl.stride + e] = f[k];
    }
    const int64_t plane = (int64_t)m.lnx * m.lny;
    for (int64_t idx = 0; idx < plane * (iz1 - iz0); ++idx) {
      const int64_t lr = (int64_t)(iz0 - m.nb[4]) * plane + idx;
      int ix, iy, iz; local_row_to_node(m, lr, ix, iy, iz);
      double acc[27], bsum;
      gather_row(m, sl, KF.data(), ix, iy, iz, acc, bsum);
      h->b[lr] = bsum;
      store_row(m, ix, iy, iz, h->roff[lr], acc, h->coefs.data());
    }
  }
}

// k_assemble_fused thread by thread: columns of BX x BY nodes marched through the node planes with a two-plane ring of
// element matrices (slot = element plane & 1), elements from the per-axis Jacobian tables, rows gathered out of the ring
void hs_assemble_ring(void *p, int BX, int BY)
{
  Harness *h = (Harness *)p; const MeshDesc &m = h->m;
  double grad[192], psi[64];
  hex8_tables(grad, psi);
  const int nmax = std::max(m.nx, std::max(m.ny, m.nz));
  std::vector<double> tab((size_t)3 * nmax * 24);
  const int n[3] = { m.nx, m.ny, m.nz };
  for (int axis = 0; axis < 3; ++axis)
    for (int e = 0; e < n[axis]; ++e) {
      const double c0 = (double)e / (double)n[axis], hh = 1.0 / (double)n[axis], c1 = c0 + hh;
      for (int g = 0; g < 8; ++g) hex8_jacobian_column(grad, g, axis, c0, c1, &tab[((size_t)axis * nmax + e) * 24 + 3 * g]);
    }
  const int WX = BX + 1, WY = BY + 1, NE = WX * WY;
  std::vector<double> KF((size_t)2 * 44 * NE);
  for (int by0 = m.nb[2]; by0 < m.nb[3]; by0 += BY)
    for (int bx0 = m.nb[0]; bx0 < m.nb[1]; bx0 += BX) {
      const int wx0 = bx0 - 1, wy0 = by0 - 1;
      auto element_plane = [&](int ez) {
        if (ez < 0 || ez >= m.nz) return;
        for (int t = 0; t < NE; ++t) {
          const int ex = wx0 + t % WX, ey = wy0 + t / WX;
          if (ex < 0 || ey < 0 || ex >= m.nx || ey >= m.ny || ex >= m.nb[1] || ey >= m.nb[3]) continue;
          double K[36], f[8];
          hex8_element_tables(grad, psi, &tab[((size_t)0 * nmax + ex) * 24], &tab[((size_t)1 * nmax + ey) * 24], &tab[((size_t)2 * nmax + ez) * 24], K, f);
          double *E = &KF[(size_t)((ez & 1) * 44) * NE + t];
          for (int k = 0; k < 36; ++k) E[(size_t)k * NE] = K[k];
          for (int k = 0; k < 8; ++k) E[(size_t)(36 + k) * NE] = f[k];
        }
      };
      element_plane(m.nb[4] - 1);
      for (int iz = m.nb[4]; iz < m.nb[5]; ++iz) {
        element_plane(iz);
        for (int t = 0; t < BX * BY; ++t) {
          const int ix = bx0 + t % BX, iy = by0 + t / BX;
          if (ix >= m.nb[1] || iy >= m.nb[3]) continue;
          double acc[27], bsum;
          gather_row_ring(m, KF.data(), wx0, wy0, WX, NE, ix, iy, iz, acc, bsum);
          const int64_t lr = (int64_t)(ix - m.nb[0]) + (int64_t)m.lnx * ((iy - m.nb[2]) + (int64_t)m.lny * (iz - m.nb[4]));
          h->b[lr] = bsum;
          store_row(m, ix, iy, iz, h->roff[lr], acc, h->coefs.data());
        }
      }
    }
}

void hs_dirichlet(void *p, int which, double val)
{
  Harness *h = (Harness *)p; const MeshDesc &m = h->m;
  for (int64_t lr = 0; lr < m.nrows; ++lr) {
    int ix, iy, iz; local_row_to_node(m, lr, ix, iy, iz);
    if (!dirichlet_touches(m, which, ix, iy, iz)) continue;
    dirichlet_row(m, which, val, lr, ix, iy, iz, h->roff[lr], (int)(h->roff[lr + 1] - h->roff[lr]), h->cols.data(), h->coefs.data(), h->b.data());
  }
}

// compute_local_maps + k_localize_cols
void hs_make_local(void *p)
{
  Harness *h = (Harness *)p; const MeshDesc &m = h->m;
  if (m.nranks == 1) { h->maps = LocalMaps(); h->maps.num_cols = m.nrows; return; }
  compute_local_maps(m, h->maps);
  // the lookup k_localize_cols does: binary search of the global row in sorted_global
  const std::vector<int> &sg = h->maps.sorted_global, &sl = h->maps.sorted_local;
  const int start = m.first_row, stop = m.first_row + m.nrows - 1;
  for (size_t i = 0; i < h->cols.size(); ++i) {
    const int c = h->cols[i];
    if (c >= start && c <= stop) { h->cols[i] = c - start; continue; }
    auto it = std::lower_bound(sg.begin(), sg.end(), c);
    h->cols[i] = (it != sg.end() && *it == c) ? sl[it - sg.begin()] : -1;
  }
}

void hs_get_csr(void *p, int *rows, long long *roff, int *cols, double *coefs, double *b)
{
  Harness *h = (Harness *)p;
  if (rows) memcpy(rows, h->rows.data(), h->rows.size() * sizeof(int));
  if (roff) for (size_t i = 0; i < h->roff.size(); ++i) roff[i] = h->roff[i];
  if (cols) memcpy(cols, h->cols.data(), h->cols.size() * sizeof(int));
  if (coefs) memcpy(coefs, h->coefs.data(), h->coefs.size() * sizeof(double));
  if (b) memcpy(b, h->b.data(), h->b.size() * sizeof(double));
}
long long hs_num_cols(void *p) { return ((Harness *)p)->maps.num_cols; }
long long hs_get_map(void *p, const char *name, int *out)
{
  Harness *h = (Harness *)p; const std::vector<int> *v = nullptr;
  if (!strcmp(name, "external_index")) v = &h->maps.external_index;
  else if (!strcmp(name, "external_local_index")) v = &h->maps.external_local_index;
  else if (!strcmp(name, "elements_to_send")) v = &h->maps.elements_to_send;
  else if (!strcmp(name, "neighbors")) v = &h->maps.neighbors;
  else if (!strcmp(name, "recv_length")) v = &h->maps.recv_length;
  else if (!strcmp(name, "send_length")) v = &h->maps.send_length;
  else if (!strcmp(name, "send_dst_offset")) v = &h->maps.send_dst_offset;
  else if (!strcmp(name, "sorted_global")) v = &h->maps.sorted_global;
  else if (!strcmp(name, "sorted_local")) v = &h->maps.sorted_local;
  else return -1;
  if (out && !v->empty()) memcpy(out, v->data(), v->size() * sizeof(int));
  return (long long)v->size();
}

void hs_element(int nx, int ny, int nz, int ex, int ey, int ez, double *K36, double *f8)
{
  double grad[192], psi[64];
  hex8_tables(grad, psi);
  hex8_element(grad, psi, (double)ex / (double)nx, (double)ey / (double)ny, (double)ez / (double)nz,
               1.0 / (double)nx, 1.0 / (double)ny, 1.0 / (double)nz, K36, f8);
}
void hs_tables(double *grad192, double *psi64) { hex8_tables(grad192, psi64); }
// the same element through the per-axis Jacobian tables of the fused assembly kernel (hex8_jacobian_column / hex8_element_tables)
void hs_element_tables(int nx, int ny, int nz, int ex, int ey, int ez, double *K36, double *f8)
{
  double grad[192], psi[64], jt[3][24];
  hex8_tables(grad, psi);
  const int n[3] = { nx, ny, nz }, e[3] = { ex, ey, ez };
  for (int axis = 0; axis < 3; ++axis) {
    const double c0 = (double)e[axis] / (double)n[axis], h = 1.0 / (double)n[axis], c1 = c0 + h;
    for (int g = 0; g < 8; ++g) hex8_jacobian_column(grad, g, axis, c0, c1, jt[axis] + 3 * g);
  }
  hex8_element_tables(grad, psi, jt[0], jt[1], jt[2], K36, f8);
}

// The exact-dot pipeline (k_chunk_sums -> k_dot_maps -> k_dot_walk) executed sequentially with the same per-thread
// arithmetic and the same storage conventions (piece maps always stored, leaf maps only for dirty pieces).
// stats[0] = dirty leaves, stats[1] = 1 if the whole-vector replay path ran, stats[2] = gaps whose collapsed map
// failed its check, stats[3] = leaves where the fp and integer map formulations disagree.
double hs_exact_dot(long long n, const double *x, const double *y, long long *stats)
{
  const int npieces = (int)((n + DOT_PIECE - 1) / DOT_PIECE);
  const int nchunks = (npieces + DOT_PIECES_PER_CHUNK - 1) / DOT_PIECES_PER_CHUNK;
  std::vector<double> chunk_sums(nchunks + 1, 0.0), piece_sums((size_t)nchunks * DOT_PIECES_PER_CHUNK + 1, 0.0);
  for (int c = 0; c < nchunks; ++c) {      // k_chunk_sums: one warp per piece, lane l strides the piece, shfl_down tree
    double tot = 0.0;
    for (int wq = 0; wq < DOT_PIECES_PER_CHUNK; ++wq) {
      const long long base = ((long long)c * DOT_PIECES_PER_CHUNK + wq) * DOT_PIECE;
      double v[32];
      for (int l = 0; l < 32; ++l) {
        double acc = 0.0;
        for (int it = 0; it < DOT_LEAF; ++it) { const long long i = base + it * 32 + l; if (i < n) acc += x[i] * y[i]; }
        v[l] = acc;
      }
      for (int o = 16; o > 0; o >>= 1) for (int l = 0; l < 32; ++l) v[l] += (l + o < 32) ? v[l + o] : 0.0;   // lane 0 exact
      if (base < n) piece_sums[(size_t)c * DOT_PIECES_PER_CHUNK + wq] = v[0];
      tot += v[0];
    }
    chunk_sums[c] = tot;
  }

  if (stats) stats[3] = 0;
  const int nsupers = (nchunks + 31) / 32, nhypers = (nsupers + 31) / 32;
  std::vector<DotMap> cta_maps(nchunks + 1, dot_invalid()), leaf_maps((size_t)(nchunks + 1) * DOT_CTA_THREADS, dot_invalid()),
      warp_maps((size_t)(nchunks + 1) * 8, dot_invalid()), super_maps(nsupers + 1), hyper_maps(nhypers + 1);
  std::vector<int> dirty_list;
  long long dirty_count = 0;
  // k_dot_maps: CTA b owns pieces [b*PP, (b+1)*PP); NT threads build the table of predicted running sums
  const int NW = 16, NT = 32 * NW, SMS = 148;
  long long grid = std::min<long long>(SMS, npieces);
  grid = std::max<long long>(grid, (npieces + 2048 - 1) / 2048);
  const int PP = npieces ? (int)((npieces + grid - 1) / grid) : 1;
  grid = npieces ? (npieces + PP - 1) / PP : 0;
  for (int b = 0; b < (int)grid; ++b) {
    const int p0 = b * PP, np = std::min(PP, npieces - p0), c0 = p0 / DOT_PIECES_PER_CHUNK;
    std::vector<double> table(np + 1);
    {
      double warp_part[NW], warp_run[NW], inc_t[NT], run_t[NT];
      const int R = (np + NT - 1) / NT;
      for (int wq = 0; wq < NW; ++wq) {
        double part[32], v[32];
        for (int l = 0; l < 32; ++l) {
          const int tid = wq * 32 + l;
          double pt = 0.0;
          for (int i = tid; i < c0; i += NT) pt += chunk_sums[i];
          if (tid < p0 - c0 * DOT_PIECES_PER_CHUNK) pt += piece_sums[(size_t)c0 * DOT_PIECES_PER_CHUNK + tid];
          part[l] = pt;
          double run = 0.0;
          for (int q = 0; q < R; ++q) if (tid * R + q < np) run += piece_sums[(size_t)p0 + tid * R + q];
          run_t[tid] = run; v[l] = run;
        }
        for (int o = 16; o > 0; o >>= 1) { double u[32]; for (int l = 0; l < 32; ++l) u[l] = part[l ^ o]; for (int l = 0; l < 32; ++l) part[l] += u[l]; }
        warp_part[wq] = part[0];
        for (int o = 1; o < 32; o <<= 1) { double u[32]; for (int l = 0; l < 32; ++l) u[l] = l >= o ? v[l - o] : 0.0; for (int l = 0; l < 32; ++l) if (l >= o) v[l] += u[l]; }
        for (int l = 0; l < 32; ++l) inc_t[wq * 32 + l] = v[l];
        warp_run[wq] = v[31];
      }
      double B0 = 0.0;
      for (int q = 0; q < NW; ++q) B0 += warp_part[q];
      for (int tid = 0; tid < NT; ++tid) {
        double woff = 0.0;
        for (int q = 0; q < tid / 32; ++q) woff += warp_run[q];
        double acc = woff + (inc_t[tid] - run_t[tid]);
        for (int q = 0; q < R; ++q) if (tid * R + q < np) { table[tid * R + q] = B0 + acc; acc += piece_sums[(size_t)p0 + tid * R + q]; }
        if (tid == NT - 1) table[np] = B0 + acc;
      }
    }
    for (int j = 0; j < np; ++j) {          // one warp per piece
      const int piece = p0 + j;
      double prod[32][DOT_LEAF], ordered[32][DOT_LEAF], ls[32], inc[32], pres[32];
      int cnt[32], hyps[33];
      for (int l = 0; l < 32; ++l) {
        const long long lb = (long long)piece * DOT_PIECE + (long long)l * DOT_LEAF;
        cnt[l] = (int)(lb >= n ? 0 : (n - lb < DOT_LEAF ? n - lb : DOT_LEAF));
        for (int i = 0; i < DOT_LEAF; ++i) ordered[l][i] = lb + i < n ? x[lb + i] * y[lb + i] : 0.0;
        const int rot = cnt[l] == DOT_LEAF ? (l & (DOT_LEAF - 1)) : 0;      // full leaves are visited in the rotated order
        double sacc = 0.0;
        for (int i = 0; i < DOT_LEAF; ++i) { prod[l][i] = ordered[l][i ^ rot]; sacc += prod[l][i]; }
        ls[l] = sacc; inc[l] = sacc;
      }
      for (int o = 1; o < 32; o <<= 1) { double u[32]; for (int l = 0; l < 32; ++l) u[l] = l >= o ? inc[l - o] : 0.0; for (int l = 0; l < 32; ++l) if (l >= o) inc[l] += u[l]; }
      for (int l = 0; l < 32; ++l) { pres[l] = table[j] + (inc[l] - ls[l]); hyps[l] = dot_hyp_of(pres[l]); }
      hyps[32] = dot_hyp_of(table[j + 1]);
      DotMap lm[32]; bool dl[32]; bool any = false;
      for (int l = 0; l < 32; ++l) {
        const int hyp = hyps[l];
        bool risky = false;
        DotMap m = dot_identity();
        if (cnt[l] == DOT_LEAF && hyp >= 0 && hyp != HYP_ZERO) {
          double sa, slo, shi;
          if (dot_leaf16_simple(prod[l], hyp, pres[l], sa, slo, shi)) { m.a0 = (long long)sa; m.a1 = m.a0; m.lo = (long long)slo; m.hi = (long long)shi; m.hyp = hyp; m.pad = 0; }
          else { m = dot_leaf_map16_fp(ordered[l], hyp); risky = dot_leaf_risky(m, hyp, pres[l]); }
        } else if (cnt[l] == DOT_LEAF) m = dot_leaf_map16(prod[l], hyp, pres[l]);
        else if (cnt[l] > 0) { const long long lb = (long long)piece * DOT_PIECE + (long long)l * DOT_LEAF; m = dot_leaf_map(x, y, lb, lb + cnt[l], hyp); }
        lm[l] = m;
        if (cnt[l] > 0) {   // the integer formulation must agree (offsets exactly; the fp bounds may only be wider: they are the sums of the
                            // negative / positive offsets when the leaf is far from a binade boundary) wherever usable
          const DotMap li = dot_leaf_map_products(ordered[l], cnt[l], hyp);
          const bool same = li.hyp == m.hyp && li.a0 == m.a0 && li.a1 == m.a1 && m.lo <= li.lo && m.hi >= li.hi;
          const bool int_unusable = li.hyp >= 0 && li.hyp != HYP_ZERO && (li.hi >= (1LL << 52) - 16 || li.lo <= -(1LL << 52) + 16 || li.hi - li.lo >= (1LL << 52) - 2);
          if (!same && !(m.hyp == HYP_INVALID && int_unusable) && stats) { stats[3]++; if (getenv("HS_DEBUG")) fprintf(stderr, "piece %d leaf %d cnt %d hyp %d: int(%d %lld %lld %lld %lld) fp(%d %lld %lld %lld %lld)\n", piece, l, cnt[l], hyp, li.hyp, li.a0, li.a1, li.lo, li.hi, m.hyp, m.a0, m.a1, m.lo, m.hi); }
        }
        dl[l] = cnt[l] > 0 && (m.hyp == HYP_INVALID || hyps[l] != hyps[l + 1] || risky);
        any = any || dl[l];
      }
      DotMap wm = dot_identity();
      for (int l = 0; l < 32; ++l) wm = dot_compose(wm, lm[l]);
      if (any) {
        for (int l = 0; l < 32; ++l) {
          leaf_maps[(size_t)piece * 32 + l] = lm[l];
          if (dl[l]) { if (dirty_count < DOT_MAX_DIRTY) dirty_list.push_back(piece * 32 + l); ++dirty_count; }
        }
      }
      wm.pad = any ? 1 : 0;
      warp_maps[piece] = wm;
    }
  }
  for (int c = 0; c < nchunks; ++c) {
    DotMap cm = dot_identity();
    for (int q = 0; q < DOT_PIECES_PER_CHUNK; ++q) if (c * DOT_PIECES_PER_CHUNK + q < npieces) { DotMap u = warp_maps[(size_t)c * DOT_PIECES_PER_CHUNK + q]; u.pad = 0; cm = dot_compose(cm, u); }
    cm.pad = 1;
    cta_maps[c] = cm;
  }
  for (int sg = 0; sg < nsupers; ++sg) { DotMap m = dot_identity(); for (int i = sg * 32; i < std::min(nchunks, sg * 32 + 32); ++i) { DotMap u = cta_maps[i]; u.pad = 0; m = dot_compose(m, u); } m.pad = 0; super_maps[sg] = m; }
  for (int hg = 0; hg < nhypers; ++hg) { DotMap m = dot_identity(); for (int i = hg * 32; i < std::min(nsupers, hg * 32 + 32); ++i) m = dot_compose(m, super_maps[i]); hyper_maps[hg] = m; }
  if (stats) { stats[0] = dirty_count; stats[1] = 0; stats[2] = 0; }
  if (dirty_count > DOT_MAX_DIRTY) { if (stats) stats[1] = 1; return dot_replay(x, y, 0, n, 0.0); }
  // k_dot_walk
  const int D = (int)dirty_count; const long long nleaves = (long long)npieces * 32;
  std::vector<int> sorted(dirty_list);
  std::sort(sorted.begin(), sorted.end());
  DotWork w{};
  w.cta_maps = cta_maps.data(); w.leaf_maps = leaf_maps.data(); w.warp_maps = warp_maps.data(); w.super_maps = super_maps.data(); w.hyper_maps = hyper_maps.data();
  double s = 0.0;
  for (int g = 0; g <= D; ++g) {
    const long long lo = g == 0 ? 0 : (long long)sorted[g - 1] + 1, hi = g == D ? nleaves : (long long)sorted[g];
    WalkPiece pc[DOT_GAP_PIECES];
    dot_gap_decompose(lo, hi, pc);
    DotMap gm = dot_identity();
    long long covered = lo;
    for (int q = 0; q < DOT_GAP_PIECES; ++q) {
      if (pc[q].count > 0 && (long long)pc[q].begin * dot_level_size(pc[q].level) != covered && stats) stats[2] += 1000000;   // decomposition must tile [lo,hi)
      for (int i = 0; i < pc[q].count; ++i) gm = dot_compose(gm, dot_level_maps(w, pc[q].level)[pc[q].begin + i]);
      covered += (long long)pc[q].count * dot_level_size(pc[q].level);
    }
    if (covered != (lo < hi ? hi : lo) && stats) stats[2] += 1000000;
    if (lo < hi && !dot_apply(gm, s)) { if (stats) stats[2]++; s = dot_walk_gap_slow(lo, hi, w, x, y, n, s); }
    if (g < D) { const long long b = (long long)sorted[g] * DOT_LEAF, e = b + DOT_LEAF < n ? b + DOT_LEAF : n; s = dot_replay(x, y, b, e, s); }
  }
  return s;
}

double hs_serial_dot(long long n, const double *x, const double *y) { return dot_replay(x, y, 0, n, 0.0); }

}  // extern "C"
