// kernels.cuh -- sm_100a device code of the miniFE hot path (structure, assembly, Dirichlet, CG).
// Compile with -fmad=false: every floating-point result must carry exactly one IEEE rounding per operation
// of the reference expression (the arithmetic-critical sites additionally use __dmul_rn/__dadd_rn).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mesh.h"
#include "hex8.h"
#include "exact_dot.h"
#include "assembly.h"

namespace mfe {

// ------------------------------------------------------------------------------------------------------
// CG scalar state, device resident (no host round trip per iteration): ref/src/cg_solve.hpp:113-215
// ------------------------------------------------------------------------------------------------------
struct CgScalars {
  double rtrans, oldrtrans, p_ap_dot, alpha, neg_alpha, beta, normr, brkdown_tol, tol;
  double one, zero, neg_one;
  double dot_local;      // result of the last local dot (before the rank-ordered sum)
  int stop_iter;         // kernels launched for iteration k do nothing when k > stop_iter
  int status;            // 0 running, 1 converged, 2 breakdown (error), 3 paused: host must run breakdown()
  int pause_iter;
  int max_iter;
};
enum { CG_RUNNING = 0, CG_CONVERGED = 1, CG_BREAKDOWN = 2, CG_PAUSED = 3 };
enum { EPI_NONE = 0, EPI_INIT_RR = 1, EPI_RR = 2, EPI_PAP = 3 };

__device__ __forceinline__ bool cg_gate_closed(const CgScalars *g, int k) { return g != nullptr && k > g->stop_iter; }

// Scalar bookkeeping after a dot product (dot = the global, rank-ordered value).
// hist arrays are device arrays of max_iter+1 doubles (may be null).
__device__ inline void cg_epilogue(CgScalars *s, int epi, int k, double dot, double *normr_hist, double *rtrans_hist,
                                   double *pap_hist, double *alpha_hist)
{
  if (epi == EPI_INIT_RR) {                    // cg_solve.hpp:134-138
    s->rtrans = dot;
    s->normr = sqrt(dot);
    if (normr_hist) normr_hist[0] = s->normr;
    if (rtrans_hist) rtrans_hist[0] = dot;
    if (!(s->normr > s->tol)) { s->stop_iter = 0; s->status = CG_CONVERGED; }   // loop test, cg_solve.hpp:152
    else if (s->max_iter >= 1) {               // iteration 1 recomputes normr = sqrt(rtrans) unchanged (cg_solve.hpp:154,163)
      if (normr_hist) normr_hist[1] = s->normr;
      if (rtrans_hist) rtrans_hist[1] = dot;
    }
  } else if (epi == EPI_RR) {                  // cg_solve.hpp:157-163 (k >= 2)
    s->oldrtrans = s->rtrans;
    s->rtrans = dot;
    s->beta = __ddiv_rn(s->rtrans, s->oldrtrans);
    s->normr = sqrt(s->rtrans);
    if (normr_hist) normr_hist[k] = s->normr;
    if (rtrans_hist) rtrans_hist[k] = s->rtrans;
    // iteration k runs to its end; iteration k+1 is entered only if normr > tol
    if (!(s->normr > s->tol) && s->stop_iter > k) { s->stop_iter = k; s->status = CG_CONVERGED; }
  } else if (epi == EPI_PAP) {                 // cg_solve.hpp:179-201
    s->p_ap_dot = dot;
    if (pap_hist) pap_hist[k] = dot;
    if (dot < s->brkdown_tol) {
      if (dot < 0) { s->status = CG_BREAKDOWN; s->stop_iter = k - 1; return; }
      s->status = CG_PAUSED; s->pause_iter = k; s->stop_iter = k - 1;   // host evaluates breakdown()
      return;
    }
    s->alpha = __ddiv_rn(s->rtrans, dot);
    s->neg_alpha = -s->alpha;
    if (alpha_hist) alpha_hist[k] = s->alpha;
  }
}

// ------------------------------------------------------------------------------------------------------
// Matrix structure: ref/src/generate_matrix_structure.hpp:98-132 + MatrixInitOp.hpp:127-156.
// One thread per owned node; offsets are closed-form (mesh.h), columns sorted ascending.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_structure(MeshDesc m, int *__restrict__ rows, uint32_t *__restrict__ roff,
                                                   int *__restrict__ cols)
{
  const int64_t lr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lr >= m.nrows) return;
  int ix, iy, iz;
  local_row_to_node(m, lr, ix, iy, iz);
  const int64_t off = row_offset_of(m, ix, iy, iz);
  rows[lr] = m.first_row + (int)lr;
  roff[lr] = (uint32_t)off;
  int c[27]; unsigned char s[27];
  const int len = row_columns(m, ix, iy, iz, c, s);
  for (int j = 0; j < len; ++j) cols[off + j] = c[j];
  if (lr == m.nrows - 1) roff[m.nrows] = (uint32_t)(off + len);
}

// per-tile 64-bit nnz offsets for the SpMV (tile = TR consecutive rows), structured mesh
__global__ void k_tile_base_mesh(MeshDesc m, int tr, int64_t ntiles, int64_t *__restrict__ tile_base)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  const int64_t lr = t * tr;
  if (lr >= m.nrows) { tile_base[t] = total_nnz_of(m); return; }
  int ix, iy, iz;
  local_row_to_node(m, lr, ix, iy, iz);
  tile_base[t] = row_offset_of(m, ix, iy, iz);
}
// same from explicit 32-bit offsets (imported CSR, nnz < 2^31)
__global__ void k_tile_base_csr(const uint32_t *__restrict__ roff, int64_t nrows, int tr, int64_t ntiles, int64_t *__restrict__ tile_base)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  const int64_t lr = t * tr < nrows ? t * tr : nrows;
  tile_base[t] = (int64_t)roff[lr];
}

// ------------------------------------------------------------------------------------------------------
// Assembly, step 1: element diffusion matrices and source vectors for a slab of element planes.
// ref/src/perform_element_loop.hpp:80-100 (get_elem_nodes_and_coords + compute_element_matrix_and_vector).
// Output SoA: KF[k * stride + e], k < 36 -> K, 36..43 -> f;  e = x-fastest index inside the slab.
// ------------------------------------------------------------------------------------------------------
__constant__ double c_grad[192];
__constant__ double c_psi[64];

__global__ void __launch_bounds__(128) k_elements(MeshDesc m, ElemSlab sl, double *__restrict__ KF)
{
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t ne = (int64_t)sl.nex * sl.ney * sl.nez;
  if (e >= ne) return;
  const int ex = sl.ex0 + (int)(e % sl.nex);
  const int ey = sl.ey0 + (int)((e / sl.nex) % sl.ney);
  const int ez = sl.ez0 + (int)(e / ((int64_t)sl.nex * sl.ney));
  // get_coords (ref/utils/box_utils.hpp:68-82): 1.0*i/n, one division; h = 1.0/n
  const double x = __ddiv_rn((double)ex, (double)m.nx), y = __ddiv_rn((double)ey, (double)m.ny), z = __ddiv_rn((double)ez, (double)m.nz);
  const double hx = __ddiv_rn(1.0, (double)m.nx), hy = __ddiv_rn(1.0, (double)m.ny), hz = __ddiv_rn(1.0, (double)m.nz);
  double K[36], f[8];
  hex8_element(c_grad, c_psi, x, y, z, hx, hy, hz, K, f);
#pragma unroll
  for (int k = 0; k < 36; ++k) KF[(int64_t)k * sl.stride + e] = K[k];
#pragma unroll
  for (int k = 0; k < 8; ++k) KF[(int64_t)(36 + k) * sl.stride + e] = f[k];
}

// ------------------------------------------------------------------------------------------------------
// Assembly, step 2: row-owner gather.  ref/src/SparseMatrix_functions.hpp:183-255 + Vector_functions.hpp:84-103.
// The reference scatters element by element (ascending element id) with +=; every matrix entry therefore
// receives its <= 8 contributions in ascending element order onto 0.0.  Here the thread that owns a row adds
// the same contributions in the same order -- atomic-free and bit-identical.
// Rows of node planes [iz0, iz1) are processed; the slab must hold element planes [iz0-1, iz1) ∩ mesh.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gather_rows(MeshDesc m, ElemSlab sl, const double *__restrict__ KF, int iz0, int iz1,
                                                     const uint32_t *__restrict__ roff, const int64_t *__restrict__ tile_base, int tr,
                                                     double *__restrict__ coefs, double *__restrict__ b)
{
  const int64_t plane = (int64_t)m.lnx * m.lny;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= plane * (iz1 - iz0)) return;
  const int64_t lr = (int64_t)(iz0 - m.nb[4]) * plane + idx;
  int ix, iy, iz;
  local_row_to_node(m, lr, ix, iy, iz);

  double acc[27], bsum;
  gather_row(m, sl, KF, ix, iy, iz, acc, bsum);
  // nnz offset of this row: 64-bit tile base + 32-bit in-tile difference
  const int64_t t = lr / tr;
  const int64_t off = tile_base[t] + (int64_t)(uint32_t)(roff[lr] - roff[t * tr]);
  b[lr] = bsum;
  store_row(m, ix, iy, iz, off, acc, coefs);
}

// ------------------------------------------------------------------------------------------------------
// Dirichlet: ref/src/SparseMatrix_functions.hpp:411-480, one call == one launch.
// Only rows within one node of a constrained face can change (an empty sum leaves b bit-identical), so
// every other thread returns at once.  Columns must still be GLOBAL rows (call before make_local_matrix).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dirichlet(MeshDesc m, int which, double val, const uint32_t *__restrict__ roff,
                                                   const int64_t *__restrict__ tile_base, int tr, const int *__restrict__ cols,
                                                   double *__restrict__ coefs, double *__restrict__ b)
{
  const int64_t lr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lr >= m.nrows) return;
  int ix, iy, iz;
  local_row_to_node(m, lr, ix, iy, iz);
  if (!dirichlet_touches(m, which, ix, iy, iz)) return;
  const int64_t t = lr / tr;
  const int64_t off = tile_base[t] + (int64_t)(uint32_t)(roff[lr] - roff[t * tr]);
  const int len = (int)(uint32_t)(roff[lr + 1] - roff[lr]);
  dirichlet_row(m, which, val, lr, ix, iy, iz, off, len, cols, coefs, b);
}

// ------------------------------------------------------------------------------------------------------
// make_local_matrix, device part: ref/src/make_local_matrix.hpp:87-107,168-180 -- owned columns become
// col - start_row, external columns become their external_local_index (binary search in the sorted table).
// ------------------------------------------------------------------------------------------------------
__global__ void k_localize_cols(int *__restrict__ cols, int64_t nnz, int start_row, int stop_row,
                                const int *__restrict__ ext_sorted_global, const int *__restrict__ ext_sorted_local, int n_ext)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
    const int c = cols[i];
    if (c >= start_row && c <= stop_row) { cols[i] = c - start_row; continue; }
    int lo = 0, hi = n_ext;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (ext_sorted_global[mid] < c) lo = mid + 1; else hi = mid; }
    cols[i] = (lo < n_ext && ext_sorted_global[lo] == c) ? ext_sorted_local[lo] : -1;
  }
}

// halo pack: ref/src/exchange_externals.hpp:116-124
__global__ void k_pack(const double *__restrict__ x, const int *__restrict__ elements_to_send, int n, double *__restrict__ send_buf,
                       const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) send_buf[i] = x[elements_to_send[i]];
}

// ------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  while (!mbar_try_wait(bar, parity)) { }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void group_sync(int id, int nthreads)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------------
// SpMV  y = A x : ref/src/SparseMatrix_functions.hpp:546-578 (matvec_std, local part).
// CSR-stream design: the nnz range of TR consecutive rows is contiguous in packed_coefs / packed_cols, so it
// is staged into shared memory by two 1-D bulk async copies (TMA engine; full 128-byte HBM bursts whatever the
// row lengths), then thread r walks ITS row in storage order: sum = fl(sum + fl(coef*x[col])) -- the reference's
// exact sequence.  G independent groups of TR threads per CTA each own S smem stages, so G*S tiles are in
// flight per SM.  x is gathered through L1/L2 (consecutive rows read consecutive x -> coalesced).
// Optionally emits per-tile plain sums of x[row]*y[row] (binade prediction for the exact dot).
// ------------------------------------------------------------------------------------------------------
template <int TR>
struct MatvecCfg {
  static constexpr int CAP = TR * 27 + 8;                       // nnz capacity of one stage
  static constexpr int SLOT_BYTES = ((CAP * 12 + 127) / 128) * 128;
};

enum { MV_STREAM_HINT = 1, MV_PREFETCH = 2, MV_X_KEEP = 4 };

__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ double ldg_keep(const double *p, uint64_t policy)
{
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int TR, int G, int S>
__global__ void __launch_bounds__(TR *G, 1)
k_matvec(const double *__restrict__ coefs, const int *__restrict__ cols, const uint32_t *__restrict__ roff,
         const int64_t *__restrict__ tile_base, const double *__restrict__ x, double *__restrict__ y, int64_t nrows,
         int64_t ntiles, int64_t nnz_padded, int64_t xlen, int flags, int zmask, double *__restrict__ tile_sums, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  constexpr int CAP = MatvecCfg<TR>::CAP;
  constexpr int SLOT = MatvecCfg<TR>::SLOT_BYTES;
  constexpr int NW = TR / 32;
  constexpr int U = 27;                               // one pass covers a whole stencil row
  extern __shared__ __align__(128) unsigned char smem[];
  const int group = threadIdx.x / TR, tg = threadIdx.x % TR;
  unsigned char *gbase = smem + (size_t)group * S * SLOT;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)G * S * SLOT) + group * S;
  double *wsum = reinterpret_cast<double *>(smem + (size_t)G * S * SLOT + G * S * 8) + group * 2 * NW;   // [2][NW], by tile parity

  const int64_t gg = (int64_t)blockIdx.x * G + group, stride = (int64_t)gridDim.x * G;
  uint64_t pol_stream = 0, pol_keep = 0;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  const bool hint = (flags & MV_STREAM_HINT) != 0, pref = (flags & MV_PREFETCH) != 0, xkeep = (flags & MV_X_KEEP) != 0;
  const int64_t pf_delta = stride * TR;               // rows between this tile and the group's next tile

  auto issue = [&](int64_t t, int s) {
    const int64_t a0 = tile_base[t] & ~3LL;
    int64_t a1 = (tile_base[t + 1] + 3) & ~3LL;
    if (a1 > nnz_padded) a1 = nnz_padded;
    const int64_t n = a1 - a0;
    if (n <= 0 || n > CAP) return;                      // oversized tile: consumers read global memory directly
    double *sc = reinterpret_cast<double *>(gbase + (size_t)s * SLOT);
    int *si = reinterpret_cast<int *>(gbase + (size_t)s * SLOT + CAP * 8);
    mbar_expect_tx(&bars[s], (uint32_t)(n * 12));
    if (hint) {
      bulk_g2s_hint(sc, coefs + a0, (uint32_t)(n * 8), &bars[s], pol_stream);
      bulk_g2s_hint(si, cols + a0, (uint32_t)(n * 4), &bars[s], pol_stream);
    } else {
      bulk_g2s(sc, coefs + a0, (uint32_t)(n * 8), &bars[s]);
      bulk_g2s(si, cols + a0, (uint32_t)(n * 4), &bars[s]);
    }
  };

  if (tg == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tg == 0)
    for (int s = 0; s < S; ++s) { const int64_t t = gg + s * stride; if (t < ntiles) issue(t, s); }

  // software-pipelined row offsets: (o0,o1) of this thread's row in the NEXT tile are loaded one tile ahead
  int64_t t = gg;
  uint32_t phase_bits = 0;            // expected mbarrier parity per stage (flips each time a stage completes)
  uint32_t o0n = 0, o1n = 0; int64_t tb0n = 0, tb1n = 0;
  if (t < ntiles) {
    const int64_t r = t * TR + tg;
    if (r < nrows) { o0n = roff[r]; o1n = roff[r + 1]; }
    tb0n = tile_base[t]; tb1n = tile_base[t + 1];
  }
  for (int64_t j = 0; t < ntiles; ++j, t += stride) {
    const int s = (int)(j % S);
    const uint32_t parity = (phase_bits >> s) & 1u;
    const int64_t r = t * TR + tg;
    const uint32_t o0 = o0n, o1 = o1n;
    const int64_t tb0 = tb0n, tb1 = tb1n;
    const int64_t tn = t + stride;
    if (tn < ntiles) {
      const int64_t rn = tn * TR + tg;
      if (rn < nrows) { o0n = roff[rn]; o1n = roff[rn + 1]; }
      tb0n = tile_base[tn]; tb1n = tile_base[tn + 1];
    }
    const int64_t a0 = tb0 & ~3LL;
    int64_t a1 = (tb1 + 3) & ~3LL;
    if (a1 > nnz_padded) a1 = nnz_padded;
    const int64_t ntile = a1 - a0;
    const bool staged = ntile > 0 && ntile <= CAP;

    double sum = 0.0, prod = 0.0;     // beta*y with beta = 0 (SparseMatrix_functions.hpp:565-568)
    const double xr = (tile_sums && r < nrows) ? __ldg(x + r) : 0.0;
    double *ws = wsum + (j & 1) * NW;
    if (staged) {
      mbar_wait(&bars[s], parity);
      phase_bits ^= 1u << s;
      if (r < nrows) {
        const double *sc = reinterpret_cast<const double *>(gbase + (size_t)s * SLOT);
        const int *si = reinterpret_cast<const int *>(gbase + (size_t)s * SLOT + CAP * 8);
        const int len = (int)(o1 - o0);
        // first entry of this row inside the stage: 32-bit difference to the tile's first row + tile skew
        const int base = (int)(uint32_t)(o0 - (uint32_t)tb0) + (int)(tb0 - a0);
        if (pref && len > 0) {
          // The x entries first touched by the group's NEXT tile are, for a banded matrix, this row's largest
          // columns shifted by the row distance: pull them into L2 now (a hint only: any in-range address is safe).
          const int p0 = len - (len + 2) / 3;
          for (int i = p0; i < len; ++i) {
            int64_t c = (int64_t)si[base + i] + pf_delta;
            if (c < xlen) prefetch_l2(x + c);
          }
        }
        for (int i0 = 0; i0 < len; i0 += U) {
          double cv[U], xv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const bool ok = i0 + u < len;
            const int ii = ok ? base + i0 + u : base;
            cv[u] = sc[ii];
            xv[u] = xkeep ? ldg_keep(x + si[ii], pol_keep) : __ldg(x + si[ii]);
          }
          // All U gathers must be in flight before the dependent add chain starts (the SM issues in order: an add
          // placed between the loads would serialise them).  Fold a bit of every gathered value into a token that
          // is masked to zero at run time (zmask == 0) and seed the chain with it: +0.0 for i0 == 0, exact.
          unsigned tok = 0;
#pragma unroll
          for (int u = 0; u < U; ++u) tok |= (unsigned)__double2hiint(xv[u]);
          sum = __dadd_rn(sum, __hiloint2double((int)(tok & (unsigned)zmask), 0));
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (i0 + u < len) sum = __dadd_rn(sum, __dmul_rn(cv[u], xv[u]));
        }
      }
    } else if (r < nrows) {
      const int64_t b0 = tb0 + (int64_t)(uint32_t)(o0 - (uint32_t)tb0);
      const int64_t len = (int64_t)(uint32_t)(o1 - o0);
      for (int64_t i = 0; i < len; ++i) sum = __dadd_rn(sum, __dmul_rn(coefs[b0 + i], __ldg(x + cols[b0 + i])));
    }
    if (r < nrows) {
      y[r] = sum;
      prod = xr * sum;
    }
    if (tile_sums) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) prod += __shfl_down_sync(0xffffffffu, prod, o);
      if ((tg & 31) == 0) ws[tg >> 5] = prod;
    }
    // every thread of the group is done with stage s -> it can be refilled
    if (TR == 32) __syncwarp(); else group_sync(1 + group, TR);
    if (tg == 0) {
      if (tile_sums) { double ts = 0.0; for (int w = 0; w < NW; ++w) ts += ws[w]; tile_sums[t] = ts; }
      const int64_t t2 = t + (int64_t)S * stride;
      if (t2 < ntiles) issue(t2, s);
    }
    // ws of this parity is next written two tiles later, i.e. after the next tile's barrier: no extra sync
  }
}

// ------------------------------------------------------------------------------------------------------
// SpMV v2 (default): one WARP per tile of 32 rows, G warps per CTA, S smem stages per warp, no runtime flags.
// Same staging as k_matvec (two 1-D bulk copies per tile, matrix stream marked evict_first), but the consumer is
// specialised: when every row of the tile holds exactly 27 entries (interior of the hex8 mesh: ~87 % of the
// tiles at 256^3) the row walk has no predication at all -- per entry LDS.64 coef, LDS col, IMAD.WIDE, LDG x,
// DMUL, DADD -- and any other staged tile uses predicated adds (the select stays off the dependent chain).
// line_tiles > 0 permutes the tile order so that the G warps of a CTA walk G ADJACENT x-lines of the mesh in step:
// the x entries of neighbouring lines are then shared through L1 instead of being re-fetched from L2.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dadd_if(double &sum, double prod, int u, int len)
{
  asm volatile("{\n.reg .pred p;\nsetp.lt.s32 p, %2, %3;\n@p add.rn.f64 %0, %0, %1;\n}\n" : "+d"(sum) : "d"(prod), "r"(u), "r"(len));
}

template <int G, int S, int H>
__global__ void __launch_bounds__(32 * G, 1)
k_matvec2(const double *__restrict__ coefs, const int *__restrict__ cols, const uint32_t *__restrict__ roff,
          const int64_t *__restrict__ tile_base, const double *__restrict__ x, double *__restrict__ y, int64_t nrows,
          int64_t ntiles, int64_t nnz_padded, int line_tiles, int zmask, double *__restrict__ tile_sums, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  constexpr int TR = 32;
  constexpr int CAP = MatvecCfg<TR>::CAP;
  constexpr int SLOT = MatvecCfg<TR>::SLOT_BYTES;
  constexpr int U = 27;
  extern __shared__ __align__(128) unsigned char smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *gbase = smem + (size_t)w * S * SLOT;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)G * S * SLOT) + w * S;
  uint64_t pol_stream = 0;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));

  // tile sequence of this warp: superstep q = blockIdx.x, blockIdx.x + gridDim.x, ...; tile_of(q) < 0: nothing there
  const int64_t LT = line_tiles > 0 ? line_tiles : 1;
  const int64_t band = (int64_t)G * LT;
  const int64_t nsuper = line_tiles > 0 ? ((ntiles + band - 1) / band) * LT : (ntiles + G - 1) / G;
  auto tile_of = [&](int64_t q) -> int64_t {
    const int64_t t = line_tiles > 0 ? (q / LT) * band + (int64_t)w * LT + (q % LT) : q * G + w;
    return t < ntiles ? t : -1;
  };
  auto next_valid = [&](int64_t &q) -> int64_t {      // advance q to the next superstep holding a tile for this warp
    for (q += gridDim.x; q < nsuper; q += gridDim.x) { const int64_t t = tile_of(q); if (t >= 0) return t; }
    return -1;
  };
  auto issue = [&](int64_t t, int s) {
    const int64_t a0 = tile_base[t] & ~3LL;
    int64_t a1 = (tile_base[t + 1] + 3) & ~3LL;
    if (a1 > nnz_padded) a1 = nnz_padded;
    const int64_t n = a1 - a0;
    if (n <= 0 || n > CAP) return;                      // oversized tile: consumers read global memory directly
    double *sc = reinterpret_cast<double *>(gbase + (size_t)s * SLOT);
    int *si = reinterpret_cast<int *>(gbase + (size_t)s * SLOT + CAP * 8);
    mbar_expect_tx(&bars[s], (uint32_t)(n * 12));
    bulk_g2s_hint(sc, coefs + a0, (uint32_t)(n * 8), &bars[s], pol_stream);
    bulk_g2s_hint(si, cols + a0, (uint32_t)(n * 4), &bars[s], pol_stream);
  };

  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  int64_t qc = (int64_t)blockIdx.x - gridDim.x;          // consume cursor (superstep), advanced by next_valid
  int64_t t = next_valid(qc);
  int64_t qi = qc, ti = t;                               // issue cursor, runs S tiles ahead
  if (lane == 0)
    for (int s = 0; s < S && ti >= 0; ++s) { issue(ti, s); ti = next_valid(qi); }
  // (lanes other than 0 do not track the issue cursor)

  uint32_t phase_bits = 0;
  uint32_t o0n = 0, o1n = 0; int64_t tb0n = 0, tb1n = 0;
  if (t >= 0) {
    const int64_t r = t * TR + lane;
    if (r < nrows) { o0n = roff[r]; o1n = roff[r + 1]; }
    tb0n = tile_base[t]; tb1n = tile_base[t + 1];
  }
  for (int64_t j = 0; t >= 0; ++j) {
    const int s = (int)(j % S);
    const uint32_t parity = (phase_bits >> s) & 1u;
    const int64_t r = t * TR + lane;
    const uint32_t o0 = o0n, o1 = o1n;
    const int64_t tb0 = tb0n, tb1 = tb1n;
    int64_t qn = qc;
    const int64_t tn = next_valid(qn);
    if (tn >= 0) {
      const int64_t rn = tn * TR + lane;
      o0n = 0; o1n = 0;
      if (rn < nrows) { o0n = roff[rn]; o1n = roff[rn + 1]; }
      tb0n = tile_base[tn]; tb1n = tile_base[tn + 1];
    }
    const int64_t a0 = tb0 & ~3LL;
    int64_t a1 = (tb1 + 3) & ~3LL;
    if (a1 > nnz_padded) a1 = nnz_padded;
    const int64_t ntile = a1 - a0;
    const bool staged = ntile > 0 && ntile <= CAP;
    const bool live = r < nrows;
    const int len = live ? (int)(o1 - o0) : 0;
    const bool full = staged && __all_sync(0xffffffffu, len == U);

    double sum = 0.0;                 // beta*y with beta = 0 (SparseMatrix_functions.hpp:565-568)
    const double xr = (tile_sums && live) ? __ldg(x + r) : 0.0;
    if (staged) {
      mbar_wait(&bars[s], parity);
      phase_bits ^= 1u << s;
      const double *sc = reinterpret_cast<const double *>(gbase + (size_t)s * SLOT);
      const int *si = reinterpret_cast<const int *>(gbase + (size_t)s * SLOT + CAP * 8);
      // first entry of this row inside the stage: 32-bit difference to the tile's first row + tile skew
      const int base = (int)(uint32_t)(o0 - (uint32_t)tb0) + (int)(tb0 - a0);
      if (full) {
        // H batches of B entries: all gathers of a batch are in flight before its dependent add chain starts (see
        // k_matvec: run-time-zero token).  H > 1 trades one extra gather latency per tile for fewer registers,
        // i.e. more resident warps.
        constexpr int B = (U + H - 1) / H;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          double cv[B], xv[B];
#pragma unroll
          for (int u = 0; u < B; ++u) if (h * B + u < U) { cv[u] = sc[base + h * B + u]; xv[u] = __ldg(x + si[base + h * B + u]); }
          unsigned tok = 0;
#pragma unroll
          for (int u = 0; u < B; ++u) if (h * B + u < U) tok |= (unsigned)__double2hiint(xv[u]);
          sum = __dadd_rn(sum, __hiloint2double((int)(tok & (unsigned)zmask), 0));
#pragma unroll
          for (int u = 0; u < B; ++u) if (h * B + u < U) sum = __dadd_rn(sum, __dmul_rn(cv[u], xv[u]));
        }
      } else if (live) {
        constexpr int B = (U + H - 1) / H;
        for (int i0 = 0; i0 < len; i0 += B) {
          double cv[B], xv[B];
#pragma unroll
          for (int u = 0; u < B; ++u) {
            const int ii = (i0 + u < len) ? base + i0 + u : base;
            cv[u] = sc[ii];
            xv[u] = __ldg(x + si[ii]);
          }
          unsigned tok = 0;
#pragma unroll
          for (int u = 0; u < B; ++u) tok |= (unsigned)__double2hiint(xv[u]);
          sum = __dadd_rn(sum, __hiloint2double((int)(tok & (unsigned)zmask), 0));
#pragma unroll
          for (int u = 0; u < B; ++u) dadd_if(sum, __dmul_rn(cv[u], xv[u]), i0 + u, len);
        }
      }
    } else if (live) {
      const int64_t b0 = tb0 + (int64_t)(uint32_t)(o0 - (uint32_t)tb0);
      for (int64_t i = 0; i < len; ++i) sum = __dadd_rn(sum, __dmul_rn(coefs[b0 + i], __ldg(x + cols[b0 + i])));
    }
    double prod = 0.0;
    if (live) { y[r] = sum; prod = xr * sum; }
    if (tile_sums) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) prod += __shfl_down_sync(0xffffffffu, prod, o);
    }
    __syncwarp();                     // every lane is done with stage s -> it can be refilled
    if (lane == 0) {
      if (tile_sums) tile_sums[t] = prod;
      if (ti >= 0) { issue(ti, s); ti = next_valid(qi); }
    }
    qc = qn; t = tn;
  }
}

// Streaming probe (measurement only): the staging half of k_matvec2 with the consumer removed -- every warp waits
// for its tile and re-arms the stage at once.  Its GB/s is the ceiling of the TMA stream at this occupancy.
template <int G, int S>
__global__ void __launch_bounds__(32 * G, 1)
k_stream_probe(const double *__restrict__ coefs, const int *__restrict__ cols, const int64_t *__restrict__ tile_base, int64_t ntiles,
               int64_t nnz_padded, double *__restrict__ sink)
{
  constexpr int CAP = MatvecCfg<32>::CAP;
  constexpr int SLOT = MatvecCfg<32>::SLOT_BYTES;
  extern __shared__ __align__(128) unsigned char smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *gbase = smem + (size_t)w * S * SLOT;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)G * S * SLOT) + w * S;
  uint64_t pol_stream = 0;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  auto issue = [&](int64_t t, int s) -> bool {
    const int64_t a0 = tile_base[t] & ~3LL;
    int64_t a1 = (tile_base[t + 1] + 3) & ~3LL;
    if (a1 > nnz_padded) a1 = nnz_padded;
    const int64_t n = a1 - a0;
    if (n <= 0 || n > CAP) return false;
    mbar_expect_tx(&bars[s], (uint32_t)(n * 12));
    bulk_g2s_hint(gbase + (size_t)s * SLOT, coefs + a0, (uint32_t)(n * 8), &bars[s], pol_stream);
    bulk_g2s_hint(gbase + (size_t)s * SLOT + CAP * 8, cols + a0, (uint32_t)(n * 4), &bars[s], pol_stream);
    return true;
  };
  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lane != 0) return;
  const int64_t stride = (int64_t)gridDim.x * G;
  int64_t t = (int64_t)blockIdx.x * G + w, ti = t;
  unsigned armed = 0, phase_bits = 0;
  for (int s = 0; s < S && ti < ntiles; ++s, ti += stride) if (issue(ti, s)) armed |= 1u << s;
  double acc = 0.0;
  for (int64_t j = 0; t < ntiles; ++j, t += stride) {
    const int s = (int)(j % S);
    if ((armed >> s) & 1u) {
      mbar_wait(&bars[s], (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
      acc += *reinterpret_cast<const double *>(gbase + (size_t)s * SLOT);
    }
    armed &= ~(1u << s);
    if (ti < ntiles) { if (issue(ti, s)) armed |= 1u << s; ti += stride; }
  }
  if (acc == 12345.678) *sink = acc;
}

// plain thread-per-row CSR SpMV in the reference order (validation twin of k_matvec; also any CSR import)
__global__ void k_matvec_simple(const double *__restrict__ coefs, const int *__restrict__ cols, const uint32_t *__restrict__ roff,
                                const int64_t *__restrict__ tile_base, int tr, const double *__restrict__ x, double *__restrict__ y,
                                int64_t nrows)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int64_t t = r / tr;
  const int64_t off = tile_base[t] + (int64_t)(uint32_t)(roff[r] - roff[t * tr]);
  const int len = (int)(uint32_t)(roff[r + 1] - roff[r]);
  double sum = 0.0;
  for (int i = 0; i < len; ++i) sum = __dadd_rn(sum, __dmul_rn(coefs[off + i], __ldg(x + cols[off + i])));
  y[r] = sum;
}

// ------------------------------------------------------------------------------------------------------
// waxpby: ref/src/Vector_functions.hpp:127-150   w = alpha*x + beta*y  (two products, one sum, no FMA)
// Scalars come from device memory (CG state) or by value.
// ------------------------------------------------------------------------------------------------------
struct ScalarArg { const double *ptr; double val; };
__device__ __forceinline__ double scalar_of(const ScalarArg &a) { return a.ptr ? *a.ptr : a.val; }

__global__ void __launch_bounds__(256) k_waxpby(int64_t n, ScalarArg alpha, const double *__restrict__ x, ScalarArg beta,
                                                const double *__restrict__ y, double *__restrict__ w, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  const double a = scalar_of(alpha), bt = scalar_of(beta);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n2 = n >> 1;
  const double2 *x2 = reinterpret_cast<const double2 *>(x), *y2 = reinterpret_cast<const double2 *>(y);
  double2 *w2 = reinterpret_cast<double2 *>(w);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const double2 xv = x2[i], yv = y2[i];
    double2 r;
    r.x = __dadd_rn(__dmul_rn(a, xv.x), __dmul_rn(bt, yv.x));
    r.y = __dadd_rn(__dmul_rn(a, xv.y), __dmul_rn(bt, yv.y));
    w2[i] = r;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) w[n - 1] = __dadd_rn(__dmul_rn(a, x[n - 1]), __dmul_rn(bt, y[n - 1]));
}

__device__ __forceinline__ double block_sum_256(double v, double *sh)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;    // valid on thread 0
}

// cg_solve.hpp:211-212 fused:  x = 1*x + alpha*p ;  r = 1*r + (-alpha)*Ap ;  chunk sums of r*r (binade prediction
// for the next exact dot).  One CTA per DOT_CHUNK elements.
__global__ void __launch_bounds__(DOT_CTA_THREADS) k_xr_update(int64_t n, const CgScalars *sc, const double *__restrict__ p,
                                                               const double *__restrict__ Ap, double *__restrict__ x,
                                                               double *__restrict__ r, double *__restrict__ chunk_sums,
                                                               const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  __shared__ double sh[DOT_CTA_THREADS / 32];
  const double alpha = sc->alpha, nalpha = sc->neg_alpha, one = 1.0;
  const int64_t base = (int64_t)blockIdx.x * DOT_CHUNK;
  double acc = 0.0;
#pragma unroll 4
  for (int it = 0; it < DOT_LEAF; ++it) {
    const int64_t i = base + (int64_t)it * DOT_CTA_THREADS + threadIdx.x;
    if (i < n) {
      x[i] = __dadd_rn(__dmul_rn(one, x[i]), __dmul_rn(alpha, p[i]));
      const double rn = __dadd_rn(__dmul_rn(one, r[i]), __dmul_rn(nalpha, Ap[i]));
      r[i] = rn;
      acc += rn * rn;
    }
  }
  const double t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) chunk_sums[blockIdx.x] = t;
}

// plain per-chunk sums of x*y (binade prediction when no producer kernel supplied them)
__global__ void __launch_bounds__(DOT_CTA_THREADS) k_chunk_sums(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                                                double *__restrict__ chunk_sums, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  __shared__ double sh[DOT_CTA_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * DOT_CHUNK;
  double acc = 0.0;
#pragma unroll 4
  for (int it = 0; it < DOT_LEAF; ++it) {
    const int64_t i = base + (int64_t)it * DOT_CTA_THREADS + threadIdx.x;
    if (i < n) acc += x[i] * y[i];
  }
  const double t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) chunk_sums[blockIdx.x] = t;
}

// SpMV tile sums -> chunk sums (per = DOT_CHUNK / rows-per-tile partial sums per chunk), one warp per chunk
__global__ void __launch_bounds__(256) k_tiles_to_chunks(const double *__restrict__ tile_sums, int per, int64_t ntiles, int nchunks,
                                                         double *__restrict__ chunk_sums, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= nchunks) return;
  double v = 0.0;
  for (int q = lane; q < per; q += 32) { const int64_t i = (int64_t)c * per + q; if (i < ntiles) v += tile_sums[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) chunk_sums[c] = v;
}

// Exclusive prefix over chunks: sums[] has `per` consecutive partial sums per chunk (per = 1 for chunk sums,
// DOT_CHUNK/TR for SpMV tile sums).  chunk_prefix[nchunks] = total (the MFE_DOT_TREE result).  One CTA.
__global__ void __launch_bounds__(1024) k_chunk_prefix(const double *__restrict__ sums, int per, int64_t nsums, int nchunks,
                                                       double *__restrict__ shp /* scratch, nchunks */,
                                                       double *__restrict__ chunk_prefix, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  __shared__ double wtot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (per == 1) {
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) shp[c] = sums[c];
  } else {
    for (int c = warp; c < nchunks; c += nw) {     // one warp reduces the partial sums of one chunk
      double v = 0.0;
      for (int q = lane; q < per; q += 32) { const int64_t i = (int64_t)c * per + q; if (i < nsums) v += sums[i]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) shp[c] = v;
    }
  }
  __syncthreads();
  // block scan: each thread owns a contiguous run of chunks
  const int runlen = (nchunks + blockDim.x - 1) / blockDim.x;
  const int c0 = threadIdx.x * runlen, c1 = min(nchunks, c0 + runlen);
  double run = 0.0;
  for (int c = c0; c < c1; ++c) run += shp[c];
  double inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    double v = lane < nw ? wtot[lane] : 0.0, iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += u; }
    wtot[lane] = iv - v;                           // exclusive warp offsets
  }
  __syncthreads();
  double pre = wtot[warp] + (inc - run);
  for (int c = c0; c < c1; ++c) { const double v = shp[c]; chunk_prefix[c] = pre; pre += v; }
  if (c1 == nchunks && c0 < nchunks) chunk_prefix[nchunks] = pre;
  if (nchunks == 0 && threadIdx.x == 0) chunk_prefix[0] = 0.0;
}

// ------------------------------------------------------------------------------------------------------
// Exact dot, map phase (see exact_dot.h).  One CTA per chunk of DOT_CHUNK terms, one leaf per thread.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ DotMap shfl_up_map(const DotMap &m, int o)
{
  DotMap r;
  r.a0 = __shfl_up_sync(0xffffffffu, m.a0, o); r.a1 = __shfl_up_sync(0xffffffffu, m.a1, o);
  r.lo = __shfl_up_sync(0xffffffffu, m.lo, o); r.hi = __shfl_up_sync(0xffffffffu, m.hi, o);
  r.hyp = __shfl_up_sync(0xffffffffu, m.hyp, o); r.pad = 0;
  return r;
}
// ordered warp reduction: lane 31 ends with the composition of lanes 0..31 in lane order
__device__ __forceinline__ DotMap warp_compose(DotMap m, int lane)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const DotMap u = shfl_up_map(m, o);
    if (lane >= o) m = dot_compose(u, m);
  }
  return m;
}

// level 2 -> 3 -> 4 of the map hierarchy, called by warp 0 of a chunk with the chunk map on lane 31.
__device__ __forceinline__ void dot_publish_chunk_map(const DotWork &w, const DotMap &cmap, int lane)
{
  // level 2 -> 3 -> 4: the last chunk of a group of 32 to arrive collapses the group (arrival tickets reset
  // themselves, so the workspace is reusable without a memset)
  const int c = blockIdx.x, nchunks = gridDim.x;
  const int sgrp = c >> 5, ssize = min(32, nchunks - (sgrp << 5));
  int last = 0;
  if (lane == 31) {
    w.cta_maps[c] = cmap;
    __threadfence();
    last = atomicAdd(&w.super_cnt[sgrp], 1) == ssize - 1;
  }
  last = __shfl_sync(0xffffffffu, last, 31);
  if (!last) return;
  __threadfence();
  DotMap sm = dot_identity();
  if (lane < ssize) {
    const DotMap *src = w.cta_maps + ((int64_t)sgrp << 5) + lane;
    sm.a0 = __ldcg(&src->a0); sm.a1 = __ldcg(&src->a1); sm.lo = __ldcg(&src->lo); sm.hi = __ldcg(&src->hi); sm.hyp = __ldcg(&src->hyp); sm.pad = 0;
  }
  sm = warp_compose(sm, lane);
  const int nsupers = (nchunks + 31) >> 5;
  const int hgrp = sgrp >> 5, hsize = min(32, nsupers - (hgrp << 5));
  if (lane == 31) {
    w.super_maps[sgrp] = sm;
    w.super_cnt[sgrp] = 0;
    __threadfence();
    last = atomicAdd(&w.hyper_cnt[hgrp], 1) == hsize - 1;
  }
  last = __shfl_sync(0xffffffffu, last, 31);
  if (!last) return;
  __threadfence();
  DotMap hm = dot_identity();
  if (lane < hsize) {
    const DotMap *src = w.super_maps + ((int64_t)hgrp << 5) + lane;
    hm.a0 = __ldcg(&src->a0); hm.a1 = __ldcg(&src->a1); hm.lo = __ldcg(&src->lo); hm.hi = __ldcg(&src->hi); hm.hyp = __ldcg(&src->hyp); hm.pad = 0;
  }
  hm = warp_compose(hm, lane);
  if (lane == 31) { w.hyper_maps[hgrp] = hm; w.hyper_cnt[hgrp] = 0; }
}

__global__ void __launch_bounds__(DOT_CTA_THREADS) k_dot_maps(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                                              DotWork w, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  __shared__ double prod[DOT_CTA_THREADS * (DOT_LEAF + 1)];     // leaf t at [t*(LEAF+1) ...): conflict-free
  __shared__ double wtot[DOT_CTA_THREADS / 32];
  __shared__ int hyps[DOT_CTA_THREADS + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * DOT_CHUNK;
  if (blockIdx.x == 0 && tid == 0) { /* dirty_count is reset by the walker of the previous dot */ }
  // coalesced load, products rounded once (fl(x*y)) exactly as the serial loop forms them
#pragma unroll 4
  for (int it = 0; it < DOT_LEAF; ++it) {
    const int e = it * DOT_CTA_THREADS + tid;
    const int64_t i = base + e;
    const double v = i < n ? __dmul_rn(x[i], y[i]) : 0.0;
    prod[(e / DOT_LEAF) * (DOT_LEAF + 1) + (e % DOT_LEAF)] = v;
  }
  __syncthreads();
  const double *mine = prod + tid * (DOT_LEAF + 1);
  const int64_t lbeg = base + (int64_t)tid * DOT_LEAF;
  const int cnt = (int)(lbeg >= n ? 0 : (n - lbeg < DOT_LEAF ? n - lbeg : DOT_LEAF));
  double ls = 0.0;
#pragma unroll
  for (int i = 0; i < DOT_LEAF; ++i) ls += mine[i];
  // exclusive block scan of the leaf sums -> predicted running sum at the start of every leaf
  double inc = ls;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  double woff = 0.0;
  for (int q = 0; q < warp; ++q) woff += wtot[q];
  const double pre = w.chunk_prefix[blockIdx.x] + (woff + (inc - ls));
  const int hyp = dot_hyp_of(pre);
  hyps[tid] = hyp;
  if (tid == 0) hyps[DOT_CTA_THREADS] = dot_hyp_of(w.chunk_prefix[blockIdx.x + 1]);
  __syncthreads();
  const int hyp_end = hyps[tid + 1];
  DotMap m = dot_leaf_map_products_fast(mine, cnt, hyp, pre);
  const bool dirty = cnt > 0 && (m.hyp == HYP_INVALID || hyp != hyp_end);
  const int any_dirty = __syncthreads_or(dirty ? 1 : 0);
  DotMap cmap;                                      // chunk map (valid on warp 0, lane 31)
  if (any_dirty) {
    // dirty chunk: keep the leaf maps and the 8 warp maps so that the walker can stitch around the dirty leaves
    w.leaf_maps[(int64_t)blockIdx.x * DOT_CTA_THREADS + tid] = m;
    if (dirty) {
      const int slot = atomicAdd(w.dirty_count, 1);
      if (slot < DOT_MAX_DIRTY) w.dirty_list[slot] = blockIdx.x * DOT_CTA_THREADS + tid;
    }
    const DotMap wm = warp_compose(m, lane);
    if (lane == 31) w.warp_maps[(int64_t)blockIdx.x * (DOT_CTA_THREADS / 32) + warp] = wm;
    cmap = dot_invalid();
    cmap.pad = 1;
  } else {
    // clean chunk: collapse the 256 leaf maps.  They are parked in shared memory (the products are dead now) in
    // a transposed order so that warp 0 reads them conflict-free: lane l composes leaves 8l..8l+7, then a scan.
    DotMap *lm = reinterpret_cast<DotMap *>(prod);
    lm[(tid & 7) * 32 + (tid >> 3)] = m;
    __syncthreads();
    if (warp == 0) {
      cmap = lm[lane];
#pragma unroll
      for (int q = 1; q < 8; ++q) cmap = dot_compose(cmap, lm[q * 32 + lane]);
      cmap = warp_compose(cmap, lane);
    }
  }
  if (warp != 0) return;
  dot_publish_chunk_map(w, cmap, lane);
}

#ifndef DOT_MAPS_MIN_CTAS
#define DOT_MAPS_MIN_CTAS 3
#endif
// Exact dot, map phase v2 (default).  Same maps as k_dot_maps, restructured for memory-level parallelism:
// every thread fetches ITS leaf (16 consecutive terms = 128 contiguous bytes per operand) with 128-bit loads that
// are all issued before the first use -- the 4 KB a warp touches are consumed whole through L1 -- the products
// stay in registers (no shared-memory transposition), the leaf maps are collapsed per warp by ordered shuffles,
// and only two block barriers remain (prefix of the leaf sums; the eight warp maps).
template <bool SAME>
__global__ void __launch_bounds__(DOT_CTA_THREADS, DOT_MAPS_MIN_CTAS) k_dot_maps2(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                                               DotWork w, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  __shared__ double wtot[DOT_CTA_THREADS / 32];
  __shared__ DotMap wmaps[DOT_CTA_THREADS / 32];
  __shared__ int wdirty[DOT_CTA_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * DOT_CHUNK;
  const int64_t lbeg = base + (int64_t)tid * DOT_LEAF;
  const int cnt = (int)(lbeg >= n ? 0 : (n - lbeg < DOT_LEAF ? n - lbeg : DOT_LEAF));
  const bool vec = cnt == DOT_LEAF && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0;
  double prod[DOT_LEAF];
  if (vec) {
    const double2 *x2 = reinterpret_cast<const double2 *>(x + lbeg), *y2 = reinterpret_cast<const double2 *>(y + lbeg);
    // products rounded once (fl(x*y)) exactly as the serial loop forms them
    if (SAME) {
      double2 xa[DOT_LEAF / 2];
#pragma unroll
      for (int q = 0; q < DOT_LEAF / 2; ++q) xa[q] = __ldg(x2 + q);
#pragma unroll
      for (int q = 0; q < DOT_LEAF / 2; ++q) { prod[2 * q] = __dmul_rn(xa[q].x, xa[q].x); prod[2 * q + 1] = __dmul_rn(xa[q].y, xa[q].y); }
    } else {
      double2 xa[DOT_LEAF / 2], ya[DOT_LEAF / 2];
#pragma unroll
      for (int q = 0; q < DOT_LEAF / 2; ++q) { xa[q] = __ldg(x2 + q); ya[q] = __ldg(y2 + q); }
#pragma unroll
      for (int q = 0; q < DOT_LEAF / 2; ++q) { prod[2 * q] = __dmul_rn(xa[q].x, ya[q].x); prod[2 * q + 1] = __dmul_rn(xa[q].y, ya[q].y); }
    }
  } else {
#pragma unroll
    for (int i = 0; i < DOT_LEAF; ++i) prod[i] = i < cnt ? __dmul_rn(__ldg(x + lbeg + i), __ldg(y + lbeg + i)) : 0.0;
  }
  double ls = 0.0;
#pragma unroll
  for (int i = 0; i < DOT_LEAF; ++i) ls += prod[i];
  // exclusive block scan of the leaf sums -> predicted running sum at the start of every leaf
  double inc = ls;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
  if (lane == 31) wtot[warp] = inc;
  const double cp0 = w.chunk_prefix[blockIdx.x], cp1 = w.chunk_prefix[blockIdx.x + 1];
  __syncthreads();
  double woff = 0.0, wmine = 0.0;
#pragma unroll
  for (int q = 0; q < DOT_CTA_THREADS / 32; ++q) { const double v = wtot[q]; if (q < warp) woff += v; if (q == warp) wmine = v; }
  const double pre = cp0 + (woff + (inc - ls));
  const int hyp = dot_hyp_of(pre);
  // predicted binade at the start of the NEXT leaf (same expression the next thread evaluates for itself)
  int hyp_end = __shfl_down_sync(0xffffffffu, hyp, 1);
  if (lane == 31) hyp_end = warp == DOT_CTA_THREADS / 32 - 1 ? dot_hyp_of(cp1) : dot_hyp_of(cp0 + ((woff + wmine) + 0.0));
  // leaf map: the common "simple" leaf (no tie, one sign) stays in floating point (a, lo, hi as exact integers in
  // doubles); anything else gets the general int64 DotMap
  bool simple = false;
  double sa = 0.0, slo = 0.0, shi = 0.0;
  DotMap m = dot_identity();
  if (cnt == DOT_LEAF && hyp >= 0 && hyp != HYP_ZERO) {
    simple = dot_leaf16_simple(prod, hyp, pre, sa, slo, shi);
    if (!simple) m = dot_leaf_map16_fp(prod, hyp);
  } else if (cnt == DOT_LEAF) {
    m = dot_leaf_map16(prod, hyp, pre);                    // hyp < 0 or the running sum is exactly +0.0
  } else if (cnt > 0) {
    m = dot_leaf_map(x, y, lbeg, lbeg + cnt, hyp);         // ragged last leaf of the vector
  }
  const bool dirty = cnt > 0 && ((!simple && m.hyp == HYP_INVALID) || hyp != hyp_end);
  const int wd = __any_sync(0xffffffffu, dirty) ? 1 : 0;
  // ordered collapse of the warp's 32 leaves onto lane 31
  const int hyp0 = __shfl_sync(0xffffffffu, hyp, 0);
  const bool all_simple = __all_sync(0xffffffffu, simple && hyp == hyp0);
  DotMap wm;
  if (all_simple) {
    // simple maps compose in floating point: a = fa + ga, lo = min(flo, fa + glo), hi = max(fhi, fa + ghi); every
    // value is an exact integer while below 2^53, and a map that ever exceeds 2^52 is unusable (checked below)
    double ca = sa, clo = slo, chi = shi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double fa = __shfl_up_sync(0xffffffffu, ca, o), flo = __shfl_up_sync(0xffffffffu, clo, o), fhi = __shfl_up_sync(0xffffffffu, chi, o);
      if (lane >= o) {
        const double nlo = fa + clo, nhi = fa + chi;
        clo = flo < nlo ? flo : nlo;
        chi = fhi > nhi ? fhi : nhi;
        ca = fa + ca;
      }
    }
    if (clo > -4503599627370496.0 && chi < 4503599627370496.0) {
      wm.a0 = (long long)ca; wm.a1 = wm.a0; wm.lo = (long long)clo; wm.hi = (long long)chi; wm.hyp = hyp0; wm.pad = 0;
    } else wm = dot_invalid();
  } else {
    if (simple) { m.a0 = (long long)sa; m.a1 = m.a0; m.lo = (long long)slo; m.hi = (long long)shi; m.hyp = hyp; m.pad = 0; }
    wm = warp_compose(m, lane);
  }
  if (lane == 31) { wmaps[warp] = wm; wdirty[warp] = wd; }
  __syncthreads();
  int any_dirty = 0;
#pragma unroll
  for (int q = 0; q < DOT_CTA_THREADS / 32; ++q) any_dirty |= wdirty[q];
  DotMap cmap;                                      // chunk map (valid on warp 0, lane 31)
  if (any_dirty) {
    // dirty chunk: keep the leaf maps and the 8 warp maps so that the walker can stitch around the dirty leaves
    if (simple) { m.a0 = (long long)sa; m.a1 = m.a0; m.lo = (long long)slo; m.hi = (long long)shi; m.hyp = hyp; m.pad = 0; }
    w.leaf_maps[(int64_t)blockIdx.x * DOT_CTA_THREADS + tid] = m;
    if (dirty) {
      const int slot = atomicAdd(w.dirty_count, 1);
      if (slot < DOT_MAX_DIRTY) w.dirty_list[slot] = blockIdx.x * DOT_CTA_THREADS + tid;
    }
    if (lane == 31) w.warp_maps[(int64_t)blockIdx.x * (DOT_CTA_THREADS / 32) + warp] = wm;
    cmap = dot_invalid();
    cmap.pad = 1;
  } else if (warp == 0) {
    // the 8 warp maps sit on lanes 24..31; three ordered steps leave their composition on lane 31
    constexpr int NWM = DOT_CTA_THREADS / 32;
    cmap = lane >= 32 - NWM ? wmaps[lane - (32 - NWM)] : dot_identity();
#pragma unroll
    for (int o = 1; o < NWM; o <<= 1) {
      const DotMap u = shfl_up_map(cmap, o);
      if (lane >= o) cmap = dot_compose(u, cmap);
    }
  }
  if (warp != 0) return;
  dot_publish_chunk_map(w, cmap, lane);
}

// ------------------------------------------------------------------------------------------------------
// Exact dot, walk phase: one CTA.  Clean stretches between dirty leaves are collapsed by warps (ordered
// compositions), then thread 0 applies  [replay dirty leaf][head leaves][full chunks][tail leaves] ...  in order.
// Any map that fails its check is replayed from global memory in plain floating point, so the result is the
// serial loop's for every input.
// ------------------------------------------------------------------------------------------------------
enum { WALK_THREADS = 512 };

__device__ __forceinline__ DotMap load_map_cg(const DotMap *src)
{
  DotMap m;
  m.a0 = __ldcg(&src->a0); m.a1 = __ldcg(&src->a1); m.lo = __ldcg(&src->lo); m.hi = __ldcg(&src->hi); m.hyp = __ldcg(&src->hyp); m.pad = __ldcg(&src->pad);
  return m;
}

// One CTA.  (1) sort the dirty leaves, fetch their products; (2) one warp per gap: the <= 9 runs of the gap are
// loaded at once (one map per lane and run) and collapsed by ordered shuffle scans into ONE map per gap;
// (3) thread 0 walks  [gap 0][dirty leaf 0][gap 1][dirty leaf 1] ... [gap D]  -- ~2 short steps per dirty leaf.
__global__ void __launch_bounds__(WALK_THREADS) k_dot_walk(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                                          DotWork w, int nchunks, CgScalars *sc, int epi, int k,
                                                          double *normr_hist, double *rtrans_hist, double *pap_hist, double *alpha_hist,
                                                          const CgScalars *gate)
{
  if (cg_gate_closed(gate, k)) return;
  __shared__ int dirty[DOT_MAX_DIRTY];
  __shared__ int sorted[DOT_MAX_DIRTY];
  __shared__ DotMap gmaps[DOT_MAX_DIRTY + 1];
  __shared__ double replay_prod[DOT_MAX_DIRTY * DOT_LEAF];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = *w.dirty_count;
  const long long nleaves = (long long)nchunks * DOT_CTA_THREADS;
  double s = 0.0;
  if (D > DOT_MAX_DIRTY) {                     // pathological input: the literal loop
    if (tid == 0) s = dot_replay(x, y, 0, n, 0.0);
  } else {
    for (int i = tid; i < D; i += WALK_THREADS) dirty[i] = w.dirty_list[i];
    __syncthreads();
    for (int i = tid; i < D; i += WALK_THREADS) {        // rank sort (leaf indices are distinct)
      const int v = dirty[i]; int rk = 0;
      for (int j = 0; j < D; ++j) rk += dirty[j] < v;
      sorted[rk] = v;
    }
    __syncthreads();
    // products of the dirty leaves, fetched by all threads at once (the sequential walk then reads smem only)
    for (int q = tid; q < D * DOT_LEAF; q += WALK_THREADS) {
      const int64_t i = (int64_t)sorted[q / DOT_LEAF] * DOT_LEAF + (q % DOT_LEAF);
      replay_prod[q] = i < n ? __dmul_rn(x[i], y[i]) : 0.0;
    }
    for (int g = warp; g <= D; g += WALK_THREADS / 32) {
      const long long lo = g == 0 ? 0 : (long long)sorted[g - 1] + 1, hi = g == D ? nleaves : (long long)sorted[g];
      WalkPiece pc[DOT_GAP_PIECES];
      dot_gap_decompose(lo, hi, pc);
      // the maps of the gap in application order form one flat list; lane l takes a contiguous slice of it
      int total = 0;
#pragma unroll
      for (int q = 0; q < DOT_GAP_PIECES; ++q) total += pc[q].count;
      const int per = (total + 31) / 32, f0 = lane * per, f1 = min(total, f0 + per);
      DotMap gm = dot_identity();
      for (int base = f0; base < f1; base += 8) {
        DotMap buf[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          int f = base + u;
          if (f < f1) {
            int q = 0;
#pragma unroll
            for (int t = 0; t < DOT_GAP_PIECES - 1; ++t) if (q == t && f >= pc[t].count) { f -= pc[t].count; q = t + 1; }
            buf[u] = load_map_cg(dot_level_maps(w, pc[q].level) + pc[q].begin + f);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) if (base + u < f1) gm = dot_compose(gm, buf[u]);
      }
      gm = warp_compose(gm, lane);
      if (lane == 31) gmaps[g] = gm;
    }
    __syncthreads();
    if (tid == 0) {
      for (int g = 0; g <= D; ++g) {
        const long long lo = g == 0 ? 0 : (long long)sorted[g - 1] + 1, hi = g == D ? nleaves : (long long)sorted[g];
        if (lo < hi && !dot_apply(gmaps[g], s)) s = dot_walk_gap_slow(lo, hi, w, x, y, n, s);
        if (g < D) {
          const int64_t b = (int64_t)sorted[g] * DOT_LEAF;
          const int cnt = (int)(b + DOT_LEAF <= n ? DOT_LEAF : n - b);
          double pr[DOT_LEAF];
#pragma unroll
          for (int i = 0; i < DOT_LEAF; ++i) pr[i] = replay_prod[g * DOT_LEAF + i];
#pragma unroll
          for (int i = 0; i < DOT_LEAF; ++i) if (i < cnt) s = __dadd_rn(s, pr[i]);
        }
      }
    }
  }
  if (tid == 0) {
    *w.result = s;
    *w.dirty_count = 0;                         // ready for the next dot on this workspace
    if (sc) {
      sc->dot_local = s;
      if (epi != EPI_NONE) cg_epilogue(sc, epi, k, s, normr_hist, rtrans_hist, pap_hist, alpha_hist);
    }
  }
}

// the literal serial loop on one device thread (MFE_DOT_SERIAL; validation of the exact dot)
__global__ void k_dot_serial(int64_t n, const double *__restrict__ x, const double *__restrict__ y, double *result, CgScalars *sc,
                             int epi, int k, double *normr_hist, double *rtrans_hist, double *pap_hist, double *alpha_hist,
                             const CgScalars *gate)
{
  if (cg_gate_closed(gate, k)) return;
  const double s = dot_replay(x, y, 0, n, 0.0);
  *result = s;
  if (sc) { sc->dot_local = s; if (epi != EPI_NONE) cg_epilogue(sc, epi, k, s, normr_hist, rtrans_hist, pap_hist, alpha_hist); }
}
// MFE_DOT_TREE: the total left by k_chunk_prefix
__global__ void k_dot_tree_finish(const double *chunk_prefix, int nchunks, double *result, CgScalars *sc, int epi, int k,
                                  double *normr_hist, double *rtrans_hist, double *pap_hist, double *alpha_hist, const CgScalars *gate)
{
  if (cg_gate_closed(gate, k)) return;
  const double s = chunk_prefix[nchunks];
  *result = s;
  if (sc) { sc->dot_local = s; if (epi != EPI_NONE) cg_epilogue(sc, epi, k, s, normr_hist, rtrans_hist, pap_hist, alpha_hist); }
}
// multi-rank: rank-ordered sum of the gathered local dots ((r0+r1)+r2)+..., then the scalar bookkeeping
__global__ void k_dot_allreduce_finish(const double *gathered, int nranks, double *result, CgScalars *sc, int epi, int k,
                                       double *normr_hist, double *rtrans_hist, double *pap_hist, double *alpha_hist, const CgScalars *gate)
{
  if (cg_gate_closed(gate, k)) return;
  double g = gathered[0];
  for (int p = 1; p < nranks; ++p) g = __dadd_rn(g, gathered[p]);
  *result = g;
  if (sc && epi != EPI_NONE) cg_epilogue(sc, epi, k, g, normr_hist, rtrans_hist, pap_hist, alpha_hist);
}

__global__ void k_fill(double *p, int64_t n, double v)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

}  // namespace mfe
