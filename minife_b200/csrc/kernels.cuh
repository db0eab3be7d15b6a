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
  int status;            // 0 running, 1 converged, 2 breakdown (error), 3 paused: host must run breakdown(), 4 peer lost
  int pause_iter;
  int max_iter;
};
enum { CG_RUNNING = 0, CG_CONVERGED = 1, CG_BREAKDOWN = 2, CG_PAUSED = 3, CG_COMM_ERROR = 4 };
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
  // the columns of a warp's 32 consecutive rows are one contiguous run of packed_cols: staged in shared memory, then
  // written with consecutive lanes on consecutive addresses (see k_gather_rows)
  __shared__ int sbuf[4][32 * 27];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int64_t lr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = lr < m.nrows;
  int64_t off = 0;
  int len = 0;
  if (active) {
    int ix, iy, iz;
    local_row_to_node(m, lr, ix, iy, iz);
    off = row_offset_of(m, ix, iy, iz);
    rows[lr] = m.first_row + (int)lr;
    roff[lr] = (uint32_t)off;
    int c[27]; unsigned char s[27];
    len = row_columns(m, ix, iy, iz, c, s);
    const int rel = (int)(off - __shfl_sync(__activemask(), off, 0));
    for (int j = 0; j < len; ++j) sbuf[wq][rel + j] = c[j];
    if (lr == m.nrows - 1) roff[m.nrows] = (uint32_t)(off + len);
  }
  __syncwarp();
  const int64_t woff = __shfl_sync(0xffffffffu, off, 0);
  int total = active ? (int)(off - woff) + len : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total = max(total, __shfl_xor_sync(0xffffffffu, total, o));
  for (int i = lane; i < total; i += 32) cols[woff + i] = sbuf[wq][i];
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

// tab[axis]: per element index along the axis, 8 Gauss points x 3 entries of the Jacobian column (hex8_jacobian_column)
__global__ void k_jacobian_tables(MeshDesc m, double *__restrict__ tx, double *__restrict__ ty, double *__restrict__ tz)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nmax = max(m.nx, max(m.ny, m.nz));
  if (t >= 3 * 8 * nmax) return;
  const int axis = t / (8 * nmax), e = (t / 8) % nmax, g = t % 8;
  const int n = axis == 0 ? m.nx : axis == 1 ? m.ny : m.nz;
  if (e >= n) return;
  // get_coords (ref/utils/box_utils.hpp:68-82): 1.0*i/n, one division; h = 1.0/n; high face = low + h
  const double c0 = __ddiv_rn((double)e, (double)n), h = __ddiv_rn(1.0, (double)n), c1 = __dadd_rn(c0, h);
  double *out = (axis == 0 ? tx : axis == 1 ? ty : tz) + ((size_t)e * 8 + g) * 3;
  hex8_jacobian_column(c_grad, g, axis, c0, c1, out);
}

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

// The same with the Jacobian columns taken from the per-axis tables (k_jacobian_tables, hex8_jacobian_column): 1152 of the
// ~4800 multiply-adds per element disappear, bit-identical results (the kernel is FP64-pipe bound).
__global__ void __launch_bounds__(128) k_elements_tables(MeshDesc m, ElemSlab sl, const double *__restrict__ tx, const double *__restrict__ ty,
                                                         const double *__restrict__ tz, double *__restrict__ KF)
{
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t ne = (int64_t)sl.nex * sl.ney * sl.nez;
  if (e >= ne) return;
  const int ex = sl.ex0 + (int)(e % sl.nex);
  const int ey = sl.ey0 + (int)((e / sl.nex) % sl.ney);
  const int ez = sl.ez0 + (int)(e / ((int64_t)sl.nex * sl.ney));
  double K[36], f[8];
  hex8_element_tables(c_grad, c_psi, tx + (size_t)ex * 24, ty + (size_t)ey * 24, tz + (size_t)ez * 24, K, f);
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
  // The 32 rows of a warp are consecutive, so their coefficients are one contiguous run of the CSR array: they are staged in
  // shared memory at their final relative positions and written out with consecutive lanes on consecutive addresses (a
  // thread storing its own 27 values writes with a 216-byte stride between lanes: 4x the store transactions).
  __shared__ double sbuf[4][32 * 27];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int64_t plane = (int64_t)m.lnx * m.lny;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx < plane * (iz1 - iz0);
  int64_t off = 0;
  int len = 0;
  if (active) {
    const int64_t lr = (int64_t)(iz0 - m.nb[4]) * plane + idx;
    int ix, iy, iz;
    local_row_to_node(m, lr, ix, iy, iz);
    double acc[27], bsum;
    gather_row(m, sl, KF, ix, iy, iz, acc, bsum);
    // nnz offset of this row: 64-bit tile base + 32-bit in-tile difference
    const int64_t t = lr / tr;
    off = tile_base[t] + (int64_t)(uint32_t)(roff[lr] - roff[t * tr]);
    len = (int)(uint32_t)(roff[lr + 1] - roff[lr]);
    b[lr] = bsum;
    const int64_t woff = __shfl_sync(__activemask(), off, 0);
    store_row(m, ix, iy, iz, off - woff, acc, sbuf[wq]);
  }
  __syncwarp();
  const int64_t woff = __shfl_sync(0xffffffffu, off, 0);          // lane 0 is active whenever any lane of the warp is
  int total = active ? (int)(off - woff) + len : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total = max(total, __shfl_xor_sync(0xffffffffu, total, o));
  for (int i = lane; i < total; i += 32) coefs[woff + i] = sbuf[wq][i];
}

// ------------------------------------------------------------------------------------------------------
// Fused assembly: element matrices never leave the SM.  A work unit is a column of ASM_BX x ASM_BY owned nodes over a range
// of node planes; the CTA marches through the planes with a two-plane ring of element matrices in shared memory:
//   compute element plane iz (one element per thread, per-axis Jacobian tables: hex8_jacobian_column) -> barrier ->
//   every row of node plane iz gathers its <= 8 elements from planes iz-1 and iz in ascending element id (the order of the
//   reference's scatter, SparseMatrix_functions.hpp:183-255) and writes its 27 coefficients and b -> barrier.
// The window of elements is one wider than the node column in x and y, so elements on the seams between columns are
// computed twice (16*17 elements for 15*16 rows: +13 %); nothing is written to global memory but A and b.
// The element computation is FP64-pipe bound (~3.7 k mul/add per element, no FMA: one rounding per reference operation).
// ------------------------------------------------------------------------------------------------------
enum { ASM_BX = 15, ASM_BY = 16, ASM_WX = ASM_BX + 1, ASM_WY = ASM_BY + 1, ASM_NE = ASM_WX * ASM_WY, ASM_THREADS = 288 };
static_assert(ASM_NE <= ASM_THREADS && ASM_BX * ASM_BY <= ASM_THREADS, "one element and one row per thread");
#define ASM_SMEM_BYTES (2 * 44 * ASM_NE * 8)

struct AsmUnits { int ncx, ncy, ncz, zlen; };     // columns in x / y, z segments, node planes per segment

__global__ void __launch_bounds__(ASM_THREADS, 1)
k_assemble_fused(MeshDesc m, AsmUnits u, const double *__restrict__ tx, const double *__restrict__ ty, const double *__restrict__ tz,
                 const uint32_t *__restrict__ roff, const int64_t *__restrict__ tile_base, int tr, double *__restrict__ coefs,
                 double *__restrict__ b, unsigned int *__restrict__ next_unit)
{
  extern __shared__ __align__(16) double KF[];          // [2][44][ASM_NE]
  __shared__ int unit_s;
  const int tid = threadIdx.x;
  const int nunits = u.ncx * u.ncy * u.ncz;
  for (;;) {
    __syncthreads();                                     // everybody is done with the previous unit (ring and unit_s)
    if (tid == 0) unit_s = (int)atomicAdd(next_unit, 1u);
    __syncthreads();
    const int unit = unit_s;
    if (unit >= nunits) return;
    const int cx = unit % u.ncx, cy = (unit / u.ncx) % u.ncy, cz = unit / (u.ncx * u.ncy);
    const int bx0 = m.nb[0] + cx * ASM_BX, by0 = m.nb[2] + cy * ASM_BY;
    const int z0 = m.nb[4] + cz * u.zlen, z1 = min(z0 + u.zlen, m.nb[5]);
    const int wx0 = bx0 - 1, wy0 = by0 - 1;              // element window (may start at -1: such elements do not exist)
    // my element of every plane, my row of every plane
    const int ex = wx0 + tid % ASM_WX, ey = wy0 + tid / ASM_WX;
    const bool elem_ok = tid < ASM_NE && ex >= 0 && ey >= 0 && ex < m.nx && ey < m.ny && ex < m.nb[1] && ey < m.nb[3];
    const int ix = bx0 + tid % ASM_BX, iy = by0 + tid / ASM_BX;
    const bool row_ok = tid < ASM_BX * ASM_BY && ix < m.nb[1] && iy < m.nb[3];
    auto element_plane = [&](int ez) {
      if (!elem_ok || ez < 0 || ez >= m.nz) return;
      double K[36], f[8];
      hex8_element_tables(c_grad, c_psi, tx + (size_t)ex * 24, ty + (size_t)ey * 24, tz + (size_t)ez * 24, K, f);
      double *E = KF + (size_t)((ez & 1) * 44) * ASM_NE + tid;
#pragma unroll
      for (int k = 0; k < 36; ++k) E[k * ASM_NE] = K[k];
#pragma unroll
      for (int k = 0; k < 8; ++k) E[(36 + k) * ASM_NE] = f[k];
    };
    element_plane(z0 - 1);
    for (int iz = z0; iz < z1; ++iz) {
      element_plane(iz);
      __syncthreads();
      if (row_ok) {
        double acc[27], bsum;
        gather_row_ring(m, KF, wx0, wy0, ASM_WX, ASM_NE, ix, iy, iz, acc, bsum);
        const int64_t lr = (int64_t)(ix - m.nb[0]) + (int64_t)m.lnx * ((iy - m.nb[2]) + (int64_t)m.lny * (iz - m.nb[4]));
        const int64_t t = lr / tr;
        const int64_t off = tile_base[t] + (int64_t)(uint32_t)(roff[lr] - roff[t * tr]);
        b[lr] = bsum;
        store_row(m, ix, iy, iz, off, acc, coefs);
      }
      __syncthreads();                                   // plane iz-1 may be overwritten by plane iz+1
    }
  }
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

// rearrange_matrix_local_external: ref/src/SparseMatrix_functions.hpp:319-408.  The reference sorts every row by local column
// index (owned columns first, externals after them) and matvec_overlap (624-675) then sums the local entries onto 0*y and
// the external ones onto 1*y -- per row ONE left-to-right pass over the re-sorted row, since 1*y is exact.  So re-sorting
// the rows in place is all it takes for k_matvec to return matvec_overlap's bits (only rows with externals change).
__global__ void k_sort_rows_local(int *__restrict__ cols, double *__restrict__ coefs, const uint32_t *__restrict__ roff,
                                  const int64_t *__restrict__ tile_base, int tr, int64_t nrows)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int64_t t = r / tr;
  const int64_t off = tile_base[t] + (int64_t)(uint32_t)(roff[r] - roff[t * tr]);
  const int len = (int)(uint32_t)(roff[r + 1] - roff[r]);
  bool sorted = true;
  for (int i = 1; i < len; ++i) sorted = sorted && cols[off + i - 1] <= cols[off + i];
  if (sorted) return;
  for (int i = 1; i < len; ++i) {            // insertion sort of (cols, coefs) by column
    const int c = cols[off + i]; const double v = coefs[off + i];
    int j = i;
    for (; j > 0 && cols[off + j - 1] > c; --j) { cols[off + j] = cols[off + j - 1]; coefs[off + j] = coefs[off + j - 1]; }
    cols[off + j] = c; coefs[off + j] = v;
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
// waxpby: ref/src/Vector_functions.hpp:127-150   w = alpha*x + beta*y  (two products, one sum, no FMA)
// Scalars come from device memory (CG state) or by value.
// ------------------------------------------------------------------------------------------------------
struct ScalarArg { const double *ptr; double val; };
__device__ __forceinline__ double scalar_of(const ScalarArg &a) { return a.ptr ? *a.ptr : a.val; }

// cg_solve calls this with w aliasing y (p = r + beta*p) and with x aliasing y (p = 1*x + 0*x): no __restrict__; every
// element is read and then written by the same thread.
__global__ void __launch_bounds__(256) k_waxpby(int64_t n, ScalarArg alpha, const double *x, ScalarArg beta,
                                                const double *y, double *w, const CgScalars *gate, int k)
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


__global__ void k_fill(double *p, int64_t n, double v)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

}  // namespace mfe
