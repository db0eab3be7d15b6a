// api.cu -- context, host-side orchestration and the extern "C" surface declared in include/minife_b200.h.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/minife_b200.h"
#include "kernels.cuh"
#include "spmv.cuh"
#include "dot_kernels.cuh"
#include "halo_maps.h"
#include "verify.cuh"

using namespace mfe;

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const char *fmt, ...)
{
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define CKN(call)                                                                                    \
  do {                                                                                               \
    int r_ = (call);                                                                                 \
    if (r_ != 0) return fail("%s:%d %s: NCCL error %d (%s)", __FILE__, __LINE__, #call, r_, nccl_errstr(r_)); \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// NCCL, bound lazily (dlopen) so that the library loads on a box without NCCL / without a GPU
// ---------------------------------------------------------------------------------------------------------
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_t_;
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
  int (*CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_t_) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t_, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
enum { NCCL_FLOAT64 = 8, NCCL_INT32 = 2, NCCL_INT8 = 0 };   // ncclDataType_t values (nccl.h)
static const char *nccl_errstr(int r) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"; }
static int nccl_load()
{
  if (g_nccl.h) return 0;
  const char *names[] = { "libnccl.so.2", "libnccl.so" };
  for (const char *n : names) { g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
  if (!g_nccl.h) return fail("cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                  \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name);                                      \
  if (!g_nccl.field) return fail("libnccl: missing symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
  SYM(AllGather, "ncclAllGather") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------
struct DotWs {                  // workspace of the exact dot (sized for nrows)
  DotWork w{};                  // w.chunk_sums / w.piece_sums are set per launch to one of the two pairs below
  double *chunk_sums_a = nullptr, *piece_sums_a = nullptr;   // written by k_xr_update (r.r) and k_chunk_sums
  double *chunk_sums_b = nullptr, *piece_sums_b = nullptr;   // accumulated by the SpMV (p.Ap), zeroed by k_xr_update
};

// NVLink peer-memory communicator (peer.cuh): one mailbox per rank, opened by every other rank through CUDA IPC
struct PeerComm {
  bool ok = false;                 // every rank exported its mailbox and opened all the others
  void *mailbox = nullptr;         // my mailbox (cudaMalloc)
  std::vector<void *> base;        // base[p]: rank p's mailbox in my address space (base[rank] = mailbox)
  std::vector<int64_t> halo_cap;   // per rank: capacity (doubles) of one halo buffer
  size_t off_flags = 0, off_halo = 0;
  unsigned long long halo_epoch = 0, dot_epoch = 0;
  unsigned int *d_ticket = nullptr;
  int *d_err = nullptr;
};

// CUDA events are created once per context and reused by every solve (no create/destroy inside the timed region, and
// nothing to leak on an error return)
struct EventPool {
  std::vector<cudaEvent_t> ev; size_t used = 0;
  cudaEvent_t get() { if (used == ev.size()) { cudaEvent_t e = nullptr; cudaEventCreate(&e); ev.push_back(e); } return ev[used++]; }
  void reset() { used = 0; }
  void destroy() { for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e); ev.clear(); used = 0; }
};

struct mfe_ctx {
  int rank = 0, nranks = 1, device = 0;
  MeshDesc m{};
  std::vector<RankBox> boxes;
  RankBox *d_boxes = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_x = nullptr;    // lower priority: the x update of iteration k runs here, beside the latency-bound dot kernels
  cudaEvent_t ev_alpha = nullptr, ev_xdone = nullptr;
  EventPool events;
  int assembly_mode = 0;              // 0: element slab + row-owner gather kernels (default), 1: fused, element matrices in shared memory (measured slower)
  int x_overlap = 1;                  // 1: x += alpha*p on stream_x (off the critical path), 0: fused into the r update (one stream)
  int sm_count = 148;
  int64_t launches = 0;

  // CSR (device)
  int64_t nrows = 0, nnz = 0, nnz_padded = 0, num_cols = 0;
  int *d_rows = nullptr;
  uint32_t *d_roff = nullptr;
  int *d_cols = nullptr;
  double *d_coefs = nullptr;
  int64_t ntiles = 0;
  int64_t *d_tile_base = nullptr;
  bool has_structure = false, has_local = false, mesh_csr = true;

  // vectors (device)
  double *d_b = nullptr, *d_x = nullptr, *d_r = nullptr, *d_p = nullptr, *d_Ap = nullptr;
  double *d_tmp1 = nullptr, *d_tmp2 = nullptr;        // scratch for the host-buffer entry points
  int64_t vec_cap = 0;

  // index maps (host, as the reference's CSRMatrix fields) + device copies used by the halo exchange
  std::vector<int> external_index, external_local_index, elements_to_send, neighbors, recv_length, send_length;
  int *d_elements_to_send = nullptr;
  double *d_send_buf = nullptr;

  // CG state
  CgScalars *d_sc = nullptr;
  double *d_hist = nullptr;      // 4 * (max_iter+1)
  int hist_cap = 0;
  DotWs dot;
  double *d_gather = nullptr;    // nranks doubles (allgather of local dots)
  int dot_mode = MFE_DOT_EXACT;
  int op_timing = 1;
  int mv_cfg = 0;                // SpMV configuration (see matvec_configs): 16 warps x 1 stage, 2 gather batches per row
  int tr = MV_TILE_ROWS;         // rows per SpMV tile
  int dot_maps_warps = 0;        // warps per CTA of k_dot_maps: 0 = default (32 for x.x, 24 for x.y), else 16, 24 or 32

  // comm
  ncclComm_t_ comm = nullptr;
  PeerComm peer;
  int comm_mode = 1;             // 1: NVLink peer stores when available (default), 0: NCCL send/recv + allgather
  int overlap = 1;               // peer mode: interior SpMV tiles run while the halo is in flight
  int dbg_no_ext = 0;
  int mv_ring = 1;
  int test_no_exchange = 0;      // test hook: skip the halo exchange (the caller supplies the externals), keep the two SpMV passes
  int ell_width = 0;             // > 0: the matrix stands for the reference's ELLMatrix of that row width (padding terms in the SpMV)
  int fused = 0;                 // MINIFE_FUSED reporting: SpMV + dot(Ap,p) are timed as MATVECDOT (cg_solve.hpp:172-176)
  int64_t peer_timeout_ms = 30000;
  bool peer_fits = true;         // EVERY rank's neighbour count fits the by-value tables of the peer kernels (agreed over NCCL)
  bool peer_fits_local = true;
  std::vector<int> send_dst_offset;
  int *d_tiles_boundary = nullptr;          // SpMV tiles that reference external columns (second pass, after the halo)
  int64_t *d_tile_base_flagged = nullptr;   // tile_base with bit 62 set on those tiles (first pass skips them)
  int64_t n_interior = 0, n_boundary = 0;

  // phase timings (ms)
  double t_structure = 0, t_assembly = 0, t_assembly_elem = 0, t_assembly_gather = 0, t_dir0 = 0, t_dir1 = 0, t_make_local = 0, t_cg = 0;
};

static int free_dev(void *p) { if (p) cudaFree(p); return 0; }
static void peer_teardown(mfe_ctx *c);

template <typename T>
static int dev_alloc(T **p, int64_t n)
{
  *p = nullptr;
  cudaError_t e = cudaMalloc((void **)p, (size_t)std::max<int64_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) return fail("cudaMalloc(%lld bytes): %s", (long long)(n * (int64_t)sizeof(T)), cudaGetErrorString(e));
  return 0;
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)std::max<int64_t>(1, (n + block - 1) / block); }

// Launch of a CG-loop kernel on the context stream (counts the launch for mfe_launch_count).
template <typename... P, typename... A>
static cudaError_t launch_on(mfe_ctx *c, cudaStream_t st, void (*kern)(P...), dim3 grid, dim3 block, size_t smem, A &&...args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  c->launches++;
  return cudaLaunchKernelEx(&cfg, kern, P(std::forward<A>(args))...);
}
template <typename... P, typename... A>
static cudaError_t launch_cg(mfe_ctx *c, void (*kern)(P...), dim3 grid, dim3 block, size_t smem, A &&...args)
{
  return launch_on(c, c->stream, kern, grid, block, smem, std::forward<A>(args)...);
}

// ---------------------------------------------------------------------------------------------------------
// box_partition: ref/utils/BoxPartition.cpp:79-122 (recursive coordinate bisection, z then y then x)
// ---------------------------------------------------------------------------------------------------------
static void rcb(int ip, int up, int axis, const int *box, int *out)
{
  const int np = up - ip;
  if (np == 1) { for (int i = 0; i < 6; ++i) out[6 * ip + i] = box[i]; return; }
  const int n = box[2 * axis + 1] - box[2 * axis];
  const int np_low = np / 2, np_upp = np - np_low;
  const int n_upp = (int)(((double)n) * (((double)np_upp) / ((double)np)));
  const int n_low = n - n_upp;
  const int next = (axis + 2) % 3;
  int lowb[6], uppb[6];
  for (int i = 0; i < 6; ++i) lowb[i] = uppb[i] = box[i];
  lowb[2 * axis + 1] = box[2 * axis] + n_low;
  uppb[2 * axis] = box[2 * axis] + n_low;
  uppb[2 * axis + 1] = uppb[2 * axis] + n_upp;
  if (np_low) rcb(ip, ip + np_low, next, lowb, out);
  rcb(ip + np_low, up, next, uppb, out);
}

extern "C" int mfe_abi_version(void) { return MFE_ABI_VERSION; }
extern "C" const char *mfe_last_error(void) { return g_err.c_str(); }
extern "C" int mfe_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
extern "C" int mfe_box_partition(int nx, int ny, int nz, int nranks, int *boxes)
{
  if (nx < 1 || ny < 1 || nz < 1 || nranks < 1 || !boxes) return fail("mfe_box_partition: bad arguments");
  const int g[6] = { 0, nx, 0, ny, 0, nz };
  rcb(0, nranks, 2, g, boxes);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------------------------------------
static void fill_mesh(mfe_ctx *c, const int gb[6], const int *all)
{
  MeshDesc &m = c->m;
  m.nx = gb[1]; m.ny = gb[3]; m.nz = gb[5];
  m.gnx = m.nx + 1; m.gny = m.ny + 1; m.gnz = m.nz + 1;
  m.nranks = c->nranks; m.rank = c->rank;
  const int g[3] = { m.nx, m.ny, m.nz };
  c->boxes.resize(c->nranks);
  int off = 0;
  for (int p = 0; p < c->nranks; ++p) {
    RankBox &b = c->boxes[p];
    const int *eb = all + 6 * p;
    for (int a = 0; a < 3; ++a) {      // simple_mesh_description.hpp:49-53
      b.nb[2 * a] = eb[2 * a];
      b.nb[2 * a + 1] = eb[2 * a + 1] + ((eb[2 * a + 1] > eb[2 * a] && eb[2 * a + 1] == g[a]) ? 1 : 0);
    }
    b.nrows = (b.nb[1] - b.nb[0]) * (b.nb[3] - b.nb[2]) * (b.nb[5] - b.nb[4]);
    b.first_row = off;                 // box_utils.hpp:254-264
    off += b.nrows;
  }
  const int *eb = all + 6 * c->rank;
  for (int i = 0; i < 6; ++i) { m.eb[i] = eb[i]; m.nb[i] = c->boxes[c->rank].nb[i]; }
  m.lnx = m.nb[1] - m.nb[0]; m.lny = m.nb[3] - m.nb[2]; m.lnz = m.nb[5] - m.nb[4];
  m.first_row = c->boxes[c->rank].first_row;
  m.nrows = c->boxes[c->rank].nrows;
  for (int a = 0; a < 3; ++a) { m.face_lo[a] = eb[2 * a] == 0; m.face_hi[a] = eb[2 * a + 1] == g[a]; }
  m.boxes = nullptr;
}

extern "C" mfe_ctx *mfe_create(int rank, int nranks, int device, const int global_box[6], const int *all_boxes)
{
  if (nranks < 1 || rank < 0 || rank >= nranks || !global_box || !all_boxes) { fail("mfe_create: bad arguments"); return nullptr; }
  if (global_box[0] != 0 || global_box[2] != 0 || global_box[4] != 0 || global_box[1] < 1 || global_box[3] < 1 || global_box[5] < 1) {
    fail("mfe_create: global box must be [0,nx)x[0,ny)x[0,nz)"); return nullptr;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1) { cudaGetLastError(); fail("mfe_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e)); return nullptr; }
  if (device < 0 || device >= ndev) { fail("mfe_create: device %d out of range (%d devices)", device, ndev); return nullptr; }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { fail("cudaSetDevice: %s", cudaGetErrorString(e)); return nullptr; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { fail("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return nullptr; }
  if (prop.major < 10) { fail("mfe_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return nullptr; }
  mfe_ctx *c = new mfe_ctx;
  c->rank = rank; c->nranks = nranks; c->device = device; c->sm_count = prop.multiProcessorCount;
  fill_mesh(c, global_box, all_boxes);
  if (c->m.nrows < 1) { fail("mfe_create: rank %d owns no rows (ref/src/main.cpp:116-130 rejects such runs)", rank); delete c; return nullptr; }
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if ((e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
      (e = cudaStreamCreateWithPriority(&c->stream_x, cudaStreamNonBlocking, prio_lo)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&c->ev_alpha, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&c->ev_xdone, cudaEventDisableTiming)) != cudaSuccess) { fail("cudaStreamCreate: %s", cudaGetErrorString(e)); delete c; return nullptr; }
  if (dev_alloc(&c->d_boxes, nranks)) { delete c; return nullptr; }
  cudaMemcpy(c->d_boxes, c->boxes.data(), sizeof(RankBox) * nranks, cudaMemcpyHostToDevice);
  c->m.boxes = c->d_boxes;
  // element tables (Hex8_ElemData.hpp:37-61), computed on the host with the same single-rounding arithmetic
  double grad[192], psi[64];
  hex8_tables(grad, psi);
  cudaMemcpyToSymbol(c_grad, grad, sizeof grad);
  cudaMemcpyToSymbol(c_psi, psi, sizeof psi);
  if (dev_alloc(&c->d_sc, 1) || dev_alloc(&c->d_gather, nranks)) { delete c; return nullptr; }
  cudaMemset(c->d_sc, 0, sizeof(CgScalars));
  if ((e = cudaGetLastError()) != cudaSuccess) { fail("mfe_create: %s", cudaGetErrorString(e)); delete c; return nullptr; }
  return c;
}

static void free_matrix(mfe_ctx *c)
{
  free_dev(c->d_rows); free_dev(c->d_roff); free_dev(c->d_cols); free_dev(c->d_coefs); free_dev(c->d_tile_base);
  c->d_rows = nullptr; c->d_roff = nullptr; c->d_cols = nullptr; c->d_coefs = nullptr; c->d_tile_base = nullptr;
}
static void free_vectors(mfe_ctx *c)
{
  free_dev(c->d_b); free_dev(c->d_x); free_dev(c->d_r); free_dev(c->d_p); free_dev(c->d_Ap); free_dev(c->d_tmp1); free_dev(c->d_tmp2);
  free_dev(c->dot.chunk_sums_a); free_dev(c->dot.piece_sums_a); free_dev(c->dot.chunk_sums_b); free_dev(c->dot.piece_sums_b);
  free_dev(c->dot.w.chunk_prefix); free_dev(c->dot.w.cta_maps); free_dev(c->dot.w.chunk_cnt);
  free_dev(c->dot.w.leaf_maps); free_dev(c->dot.w.warp_maps); free_dev(c->dot.w.super_maps); free_dev(c->dot.w.hyper_maps);
  free_dev(c->dot.w.super_cnt); free_dev(c->dot.w.hyper_cnt); free_dev(c->dot.w.dirty_count); free_dev(c->dot.w.dirty_list); free_dev(c->dot.w.result); free_dev(c->dot.w.stats);
  c->d_b = c->d_x = c->d_r = c->d_p = c->d_Ap = c->d_tmp1 = c->d_tmp2 = nullptr;
  c->dot = DotWs();
  c->vec_cap = 0;
}

extern "C" void mfe_destroy(mfe_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream_x) cudaStreamSynchronize(c->stream_x);
  if (c->stream) cudaStreamSynchronize(c->stream);
  peer_teardown(c);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  free_dev(c->d_tile_base_flagged); free_dev(c->d_tiles_boundary);
  free_matrix(c); free_vectors(c);
  free_dev(c->d_boxes); free_dev(c->d_sc); free_dev(c->d_hist); free_dev(c->d_gather);
  free_dev(c->d_elements_to_send); free_dev(c->d_send_buf);
  c->events.destroy();
  if (c->ev_alpha) cudaEventDestroy(c->ev_alpha);
  if (c->ev_xdone) cudaEventDestroy(c->ev_xdone);
  if (c->stream_x) cudaStreamDestroy(c->stream_x);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int mfe_comm_unique_id(char id128[128])
{
  if (nccl_load()) return 1;
  ncclUniqueId_t id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return fail("ncclGetUniqueId: %s", nccl_errstr(r));
  memcpy(id128, id.internal, 128);
  return 0;
}
// Mailbox layout (peer.cuh): dot slots [2][nranks] | halo flags [nranks] | halo data [2][halo_cap(rank)]
static int64_t halo_cap_of(const RankBox &b)
{
  const int64_t lx = b.nb[1] - b.nb[0], ly = b.nb[3] - b.nb[2], lz = b.nb[5] - b.nb[4];
  return (lx + 2) * (ly + 2) * (lz + 2) - lx * ly * lz + 8;      // the one-node shell bounds the externals
}
static DotSlot *peer_dot_slots(mfe_ctx *c, int p) { return reinterpret_cast<DotSlot *>(c->peer.base[p]); }
static unsigned long long *peer_halo_flags(mfe_ctx *c, int p) { return reinterpret_cast<unsigned long long *>((char *)c->peer.base[p] + c->peer.off_flags); }
static double *peer_halo_buf(mfe_ctx *c, int p, unsigned long long epoch)
{
  return reinterpret_cast<double *>((char *)c->peer.base[p] + c->peer.off_halo) + (epoch & 1ULL) * c->peer.halo_cap[p];
}

static void peer_teardown(mfe_ctx *c)
{
  PeerComm &P = c->peer;
  for (size_t p = 0; p < P.base.size(); ++p)
    if ((int)p != c->rank && P.base[p]) cudaIpcCloseMemHandle(P.base[p]);
  P.base.clear();
  if (P.ok && c->comm && c->d_gather) {
    // exported memory must outlive every importer's mapping: all ranks close first (mfe_destroy is collective
    // for nranks > 1, like MPI_Finalize), then free
    g_nccl.AllGather(c->dot.w.result ? (const void *)c->dot.w.result : (const void *)c->d_sc, c->d_gather, 1, NCCL_FLOAT64, c->comm, c->stream);
    cudaStreamSynchronize(c->stream);
  }
  free_dev(P.mailbox); free_dev(P.d_ticket); free_dev(P.d_err);
  P.mailbox = nullptr; P.d_ticket = nullptr; P.d_err = nullptr; P.ok = false;
  cudaGetLastError();
}

// Collective over the NCCL communicator: export my mailbox, gather every rank's IPC handle, open them.  Any failure
// on any rank (no IPC in this container, no peer access, ...) leaves peer.ok = false on ALL ranks -> NCCL path.
static int peer_setup(mfe_ctx *c)
{
  PeerComm &P = c->peer;
  const int np = c->nranks;
  P.ok = false;
  if (np > PEER_MAX_RANKS || getenv("MFE_NO_PEER")) return 0;      // the same decision on every rank
  P.halo_cap.resize(np);
  for (int p = 0; p < np; ++p) P.halo_cap[p] = halo_cap_of(c->boxes[p]);
  P.off_flags = ((size_t)2 * np * sizeof(DotSlot) + 255) & ~(size_t)255;
  P.off_halo = (P.off_flags + (size_t)np * 8 + 255) & ~(size_t)255;
  const size_t bytes = P.off_halo + (size_t)2 * P.halo_cap[c->rank] * 8;
  int local_ok = 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  if (cudaMalloc(&P.mailbox, bytes) != cudaSuccess || cudaMalloc((void **)&P.d_ticket, 4) != cudaSuccess || cudaMalloc((void **)&P.d_err, 4) != cudaSuccess) local_ok = 0;
  if (local_ok) {
    cudaMemsetAsync(P.mailbox, 0, bytes, c->stream);
    cudaMemsetAsync(P.d_ticket, 0, 4, c->stream);
    cudaMemsetAsync(P.d_err, 0, 4, c->stream);
    if (cudaIpcGetMemHandle(&mine, P.mailbox) != cudaSuccess) local_ok = 0;
  }
  cudaGetLastError();
  // gather {handle, ok} of every rank (the mailboxes are zeroed before anybody can learn their address)
  struct Rec { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  static_assert(sizeof(Rec) == 80, "Rec layout");
  Rec rec; rec.h = mine; rec.ok = local_ok; rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
  Rec *d_one = nullptr, *d_all = nullptr;
  if (dev_alloc(&d_one, 1) || dev_alloc(&d_all, np)) return 1;
  std::vector<Rec> all(np);
  CK(cudaMemcpyAsync(d_one, &rec, sizeof rec, cudaMemcpyHostToDevice, c->stream));
  CKN(g_nccl.AllGather(d_one, d_all, sizeof(Rec), NCCL_INT8, c->comm, c->stream));
  CK(cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * np, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  bool everybody = true;
  for (int p = 0; p < np; ++p) everybody = everybody && all[p].ok == 1;
  P.base.assign(np, nullptr);
  int open_ok = everybody ? 1 : 0;
  if (everybody) {
    for (int p = 0; p < np; ++p) {
      if (p == c->rank) { P.base[p] = P.mailbox; continue; }
      if (cudaIpcOpenMemHandle(&P.base[p], all[p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { P.base[p] = nullptr; open_ok = 0; cudaGetLastError(); }
    }
  }
  // second round: did every rank open every mailbox?
  rec.ok = open_ok;
  CK(cudaMemcpyAsync(d_one, &rec, sizeof rec, cudaMemcpyHostToDevice, c->stream));
  CKN(g_nccl.AllGather(d_one, d_all, sizeof(Rec), NCCL_INT8, c->comm, c->stream));
  CK(cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * np, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_one); cudaFree(d_all);
  bool all_open = true;
  for (int p = 0; p < np; ++p) all_open = all_open && all[p].ok == 1;
  if (!all_open) { peer_teardown(c); return 0; }
  P.ok = true;
  return 0;
}

// peer transport or NCCL is a decision all ranks must take alike: gather every rank's "my neighbour tables fit" flag.
// Collective (called by every rank from mfe_make_local_matrix, and from mfe_comm_init when the maps already exist).
static int peer_agree(mfe_ctx *c)
{
  c->peer_fits = c->peer_fits_local;
  if (c->nranks == 1 || !c->comm) return 0;
  std::vector<double> all(c->nranks);
  const double mine = c->peer_fits_local ? 1.0 : 0.0;
  double *d_one = nullptr;
  if (dev_alloc(&d_one, 1)) return 1;
  CK(cudaMemcpyAsync(d_one, &mine, 8, cudaMemcpyHostToDevice, c->stream));
  CKN(g_nccl.AllGather(d_one, c->d_gather, 1, NCCL_FLOAT64, c->comm, c->stream));
  CK(cudaMemcpyAsync(all.data(), c->d_gather, 8 * c->nranks, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_one);
  for (double v : all) if (v == 0.0) c->peer_fits = false;
  return 0;
}

extern "C" int mfe_comm_init(mfe_ctx *c, const char id128[128])
{
  if (!c) return fail("null context");
  if (c->nranks == 1) return 0;
  if (nccl_load()) return 1;
  CK(cudaSetDevice(c->device));
  ncclUniqueId_t id;
  memcpy(id.internal, id128, 128);
  CKN(g_nccl.CommInitRank(&c->comm, c->nranks, id, c->rank));
  if (peer_setup(c)) return 1;
  return c->has_local ? peer_agree(c) : 0;
}

// vectors + dot workspace sized for (nrows, ncols)
static int ensure_vectors(mfe_ctx *c, int64_t nrows, int64_t ncols)
{
  const int64_t need = std::max(nrows, ncols) + 8;
  if (c->vec_cap >= need) return 0;
  std::vector<double> keep_b, keep_x;
  const bool keep = c->d_b != nullptr && c->vec_cap > 0;
  double *ob = c->d_b, *ox = c->d_x;
  c->d_b = c->d_x = nullptr;
  const int64_t oldcap = c->vec_cap;
  free_vectors(c);
  if (dev_alloc(&c->d_b, need) || dev_alloc(&c->d_x, need) || dev_alloc(&c->d_r, need) || dev_alloc(&c->d_p, need) ||
      dev_alloc(&c->d_Ap, need) || dev_alloc(&c->d_tmp1, need) || dev_alloc(&c->d_tmp2, need)) return 1;
  CK(cudaMemsetAsync(c->d_b, 0, need * 8, c->stream)); CK(cudaMemsetAsync(c->d_x, 0, need * 8, c->stream));
  CK(cudaMemsetAsync(c->d_r, 0, need * 8, c->stream)); CK(cudaMemsetAsync(c->d_p, 0, need * 8, c->stream));
  CK(cudaMemsetAsync(c->d_Ap, 0, need * 8, c->stream));
  if (keep) {
    CK(cudaMemcpyAsync(c->d_b, ob, std::min(oldcap, need) * 8, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_x, ox, std::min(oldcap, need) * 8, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  free_dev(ob); free_dev(ox);
  const int64_t nchunks = (need + DOT_CHUNK - 1) / DOT_CHUNK, npieces = nchunks * DOT_PIECES_PER_CHUNK;
  if (dev_alloc(&c->dot.chunk_sums_a, nchunks + 1) || dev_alloc(&c->dot.piece_sums_a, npieces + 1) || dev_alloc(&c->dot.chunk_sums_b, nchunks + 1) ||
      dev_alloc(&c->dot.piece_sums_b, npieces + 1) || dev_alloc(&c->dot.w.chunk_prefix, nchunks + 2) ||
      dev_alloc(&c->dot.w.cta_maps, nchunks + 1) || dev_alloc(&c->dot.w.leaf_maps, (nchunks + 1) * DOT_CTA_THREADS) ||
      dev_alloc(&c->dot.w.warp_maps, npieces + 8) || dev_alloc(&c->dot.w.super_maps, nchunks / 32 + 2) ||
      dev_alloc(&c->dot.w.hyper_maps, nchunks / 1024 + 2) || dev_alloc(&c->dot.w.chunk_cnt, nchunks + 1) || dev_alloc(&c->dot.w.super_cnt, nchunks / 32 + 2) ||
      dev_alloc(&c->dot.w.hyper_cnt, nchunks / 1024 + 2) || dev_alloc(&c->dot.w.dirty_count, 1) || dev_alloc(&c->dot.w.dirty_list, DOT_MAX_DIRTY) ||
      dev_alloc(&c->dot.w.result, 1) || dev_alloc(&c->dot.w.stats, DOT_STATS_WORDS)) return 1;
  CK(cudaMemsetAsync(c->dot.w.stats, 0, 8 * DOT_STATS_WORDS, c->stream));
  CK(cudaMemsetAsync(c->dot.w.dirty_count, 0, sizeof(int), c->stream));
  CK(cudaMemsetAsync(c->dot.w.chunk_cnt, 0, sizeof(int) * (nchunks + 1), c->stream));
  CK(cudaMemsetAsync(c->dot.w.super_cnt, 0, sizeof(int) * (nchunks / 32 + 2), c->stream));
  CK(cudaMemsetAsync(c->dot.w.hyper_cnt, 0, sizeof(int) * (nchunks / 1024 + 2), c->stream));
  CK(cudaMemsetAsync(c->dot.chunk_sums_b, 0, 8 * (nchunks + 1), c->stream));
  CK(cudaMemsetAsync(c->dot.piece_sums_b, 0, 8 * (npieces + 1), c->stream));
  c->dot.w.chunk_sums = c->dot.chunk_sums_a; c->dot.w.piece_sums = c->dot.piece_sums_a;
  c->vec_cap = need;
  return 0;
}

// cudaFuncSetAttribute once per (kernel instantiation, device); one process may drive several GPUs from several threads
static std::mutex g_attr_mutex;
template <typename K>
static int set_smem_attr_once(K kern, bool (&done)[64], int device, size_t smem)
{
  std::lock_guard<std::mutex> lk(g_attr_mutex);
  if (done[device & 63]) return 0;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  done[device & 63] = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// phases
// ---------------------------------------------------------------------------------------------------------
struct PhaseTimer {
  cudaEvent_t a = nullptr, b = nullptr; cudaStream_t s;
  explicit PhaseTimer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
  double stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
  ~PhaseTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
};

static int build_tile_base(mfe_ctx *c)
{
  c->ntiles = (c->nrows + c->tr - 1) / c->tr;
  free_dev(c->d_tile_base);
  if (dev_alloc(&c->d_tile_base, c->ntiles + 2)) return 1;
  if (c->mesh_csr) k_tile_base_mesh<<<grid_for(c->ntiles + 1, 256), 256, 0, c->stream>>>(c->m, c->tr, c->ntiles, c->d_tile_base);
  else k_tile_base_csr<<<grid_for(c->ntiles + 1, 256), 256, 0, c->stream>>>(c->d_roff, c->nrows, c->tr, c->ntiles, c->d_tile_base);
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mfe_generate_matrix_structure(mfe_ctx *c, int64_t *nrows, int64_t *first_row, int64_t *nnz)
{
  if (!c) return fail("null context");
  CK(cudaSetDevice(c->device));
  PhaseTimer tm(c->stream);
  free_matrix(c);
  c->mesh_csr = true; c->has_local = false;
  c->nrows = c->m.nrows;
  c->nnz = total_nnz_of(c->m);
  c->nnz_padded = ((c->nnz + 3) & ~3LL) + 4;
  c->num_cols = 0;
  if (dev_alloc(&c->d_rows, c->nrows) || dev_alloc(&c->d_roff, c->nrows + 8) || dev_alloc(&c->d_cols, c->nnz_padded) ||
      dev_alloc(&c->d_coefs, c->nnz_padded)) return 1;
  CK(cudaMemsetAsync(c->d_cols + c->nnz, 0, (c->nnz_padded - c->nnz) * sizeof(int), c->stream));      // the padding tail; k_structure writes every real entry
  CK(cudaMemsetAsync(c->d_coefs, 0, c->nnz_padded * sizeof(double), c->stream));   // MatrixInitOp.hpp:147 coefs = 0
  k_structure<<<grid_for(c->nrows, 128), 128, 0, c->stream>>>(c->m, c->d_rows, c->d_roff, c->d_cols);
  c->launches++;
  CK(cudaGetLastError());
  if (build_tile_base(c)) return 1;
  // sized for the external tail make_local_matrix will add (bounded by the one-node shell), so nothing is reallocated later
  if (ensure_vectors(c, c->nrows, c->nrows + (c->nranks > 1 ? halo_cap_of(c->boxes[c->rank]) : 0))) return 1;
  // Vector b, x are zero-filled on construction (Vector.hpp:44-52, driver.hpp:188-189)
  CK(cudaMemsetAsync(c->d_b, 0, c->vec_cap * 8, c->stream));
  CK(cudaMemsetAsync(c->d_x, 0, c->vec_cap * 8, c->stream));
  c->t_structure = tm.stop();
  CK(cudaGetLastError());
  c->has_structure = true;
  if (nrows) *nrows = c->nrows;
  if (first_row) *first_row = c->m.first_row;
  if (nnz) *nnz = c->nnz;
  return 0;
}

// round-1 scheme: element matrices into an SoA scratch slab in global memory (k_elements), then the row-owner gather
// (k_gather_rows); kept as assembly_mode 0 for comparison
static int assemble_two_kernels(mfe_ctx *c)
{
  PhaseTimer tm(c->stream);
  const MeshDesc &m = c->m;
  // element box + ghost layer (assemble_FE_data.hpp:59-70)
  int eb[6]; for (int i = 0; i < 6; ++i) eb[i] = m.eb[i];
  if (eb[0] > 0) eb[0]--; if (eb[2] > 0) eb[2]--; if (eb[4] > 0) eb[4]--;
  if (eb[1] < m.nx) eb[1]++; if (eb[3] < m.ny) eb[3]++; if (eb[5] < m.nz) eb[5]++;
  const int nex = eb[1] - eb[0], ney = eb[3] - eb[2];
  const int64_t eplane = (int64_t)nex * ney;
  // node planes are processed in slabs so that the element scratch stays bounded (<= ~2 GiB)
  int planes = (int)std::max<int64_t>(1, (int64_t)(2LL << 30) / (eplane * 44 * 8) - 1);
  planes = std::min(planes, m.lnz);
  ElemSlab sl; sl.ex0 = eb[0]; sl.ey0 = eb[2]; sl.nex = nex; sl.ney = ney; sl.stride = eplane * (planes + 1);
  double *KF = nullptr, *tab = nullptr;
  const int nmax = std::max(m.nx, std::max(m.ny, m.nz));
  if (dev_alloc(&KF, sl.stride * 44) || dev_alloc(&tab, (int64_t)3 * nmax * 24)) return 1;
  double *tx = tab, *ty = tab + (size_t)nmax * 24, *tz = tab + (size_t)2 * nmax * 24;
  k_jacobian_tables<<<grid_for(3 * 8 * nmax, 128), 128, 0, c->stream>>>(m, tx, ty, tz);
  c->launches++;
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  double t_el = 0, t_ga = 0;
  for (int iz0 = m.nb[4]; iz0 < m.nb[5]; iz0 += planes) {
    const int iz1 = std::min(iz0 + planes, m.nb[5]);
    const int ez0 = std::max(iz0 - 1, 0), ez1 = std::min(iz1, m.nz);     // element planes touching node planes [iz0,iz1)
    sl.ez0 = ez0; sl.nez = ez1 - ez0;
    cudaEventRecord(e0, c->stream);
    const int64_t ne = eplane * sl.nez;
    if (ne > 0) { k_elements_tables<<<grid_for(ne, 128), 128, 0, c->stream>>>(m, sl, tx, ty, tz, KF); c->launches++; }
    cudaEventRecord(e1, c->stream);
    const int64_t nr = (int64_t)m.lnx * m.lny * (iz1 - iz0);
    k_gather_rows<<<grid_for(nr, 128), 128, 0, c->stream>>>(m, sl, KF, iz0, iz1, c->d_roff, c->d_tile_base, c->tr, c->d_coefs, c->d_b);
    c->launches++;
    cudaEventRecord(e2, c->stream);
    cudaEventSynchronize(e2);
    float a = 0, b = 0; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
    t_el += a; t_ga += b;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  c->t_assembly = tm.stop();
  c->t_assembly_elem = t_el; c->t_assembly_gather = t_ga;
  cudaFree(KF); cudaFree(tab);
  CK(cudaGetLastError());
  return 0;
}

// element matrices stay in shared memory: k_jacobian_tables + k_assemble_fused (kernels.cuh)
static int assemble_fused(mfe_ctx *c)
{
  PhaseTimer tm(c->stream);
  const MeshDesc &m = c->m;
  const int nmax = std::max(m.nx, std::max(m.ny, m.nz));
  double *tab = nullptr; unsigned int *d_next = nullptr;
  if (dev_alloc(&tab, (int64_t)3 * nmax * 24) || dev_alloc(&d_next, 1)) return 1;
  CK(cudaMemsetAsync(d_next, 0, sizeof(unsigned int), c->stream));
  double *tx = tab, *ty = tab + (size_t)nmax * 24, *tz = tab + (size_t)2 * nmax * 24;
  k_jacobian_tables<<<grid_for(3 * 8 * nmax, 128), 128, 0, c->stream>>>(m, tx, ty, tz);
  static bool attr_set[64] = { false };
  if (set_smem_attr_once(k_assemble_fused, attr_set, c->device, ASM_SMEM_BYTES)) return 1;
  AsmUnits u;
  u.ncx = (m.lnx + ASM_BX - 1) / ASM_BX; u.ncy = (m.lny + ASM_BY - 1) / ASM_BY;
  // z segments: enough work units to balance the SMs (~6 per SM), each long enough to amortise its extra element plane
  const int64_t cols = (int64_t)u.ncx * u.ncy;
  int ncz = (int)std::max<int64_t>(1, ((int64_t)c->sm_count * 6 + cols - 1) / cols);
  ncz = std::min(ncz, std::max(1, m.lnz / 16));
  u.zlen = (m.lnz + ncz - 1) / ncz;
  u.ncz = (m.lnz + u.zlen - 1) / u.zlen;
  const int grid = (int)std::min<int64_t>(c->sm_count, cols * u.ncz);
  k_assemble_fused<<<grid, ASM_THREADS, ASM_SMEM_BYTES, c->stream>>>(m, u, tx, ty, tz, c->d_roff, c->d_tile_base, c->tr, c->d_coefs, c->d_b, d_next);
  c->launches += 2;
  c->t_assembly = tm.stop();
  c->t_assembly_elem = c->t_assembly; c->t_assembly_gather = 0;
  cudaFree(tab); cudaFree(d_next);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mfe_assemble_FE_data(mfe_ctx *c)
{
  if (!c || !c->has_structure || !c->mesh_csr) return fail("mfe_assemble_FE_data: call mfe_generate_matrix_structure first");
  if (c->has_local) return fail("mfe_assemble_FE_data: matrix already has local indices");
  CK(cudaSetDevice(c->device));
  return c->assembly_mode == 0 ? assemble_two_kernels(c) : assemble_fused(c);
}

extern "C" int mfe_impose_dirichlet(mfe_ctx *c, double val, int which)
{
  if (!c || !c->has_structure || !c->mesh_csr) return fail("mfe_impose_dirichlet: no mesh matrix");
  if (c->has_local) return fail("mfe_impose_dirichlet: must be called before mfe_make_local_matrix (columns must be global rows)");
  if (which != 0 && which != 1) return fail("mfe_impose_dirichlet: which_set must be 0 or 1");
  CK(cudaSetDevice(c->device));
  PhaseTimer tm(c->stream);
  k_dirichlet<<<grid_for(c->nrows, 128), 128, 0, c->stream>>>(c->m, which, val, c->d_roff, c->d_tile_base, c->tr, c->d_cols, c->d_coefs, c->d_b);
  c->launches++;
  (which == 0 ? c->t_dir0 : c->t_dir1) = tm.stop();
  CK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// make_local_matrix: ref/src/make_local_matrix.hpp:42-442.
// The reference discovers externals by scanning rows (ascending) and their sorted columns, numbering each
// external at first encounter.  For the 27-point stencil that order is closed-form: an external node E is
// first met in the smallest adjacent owned row, first(E) = row(max(lo, E-1) per axis), and inside one row
// externals appear in ascending global row.  So external_index = externals sorted by (first(E), globalrow(E)).
// The peers' lists (-> elements_to_send) follow from the same rule applied to the peer's box: no messages.
// ---------------------------------------------------------------------------------------------------------
// SpMV tiles that touch an external column (boundary) and those that do not (interior): the interior pass needs
// no halo, so it runs while the neighbours' stores are in flight (the reference's matvec_overlap idea,
// SparseMatrix_functions.hpp:624-675, without re-ordering the matrix: whole rows stay intact and bit-exact).
static int classify_tiles(mfe_ctx *c)
{
  free_dev(c->d_tile_base_flagged); free_dev(c->d_tiles_boundary);
  c->d_tile_base_flagged = nullptr; c->d_tiles_boundary = nullptr; c->n_interior = c->n_boundary = 0;
  if (c->ntiles <= 0 || c->ntiles > 0x7fffffffLL) return 0;
  unsigned char *d_flag = nullptr;
  int *d_count = nullptr;
  // the list can hold every tile (its true length is only known on the device); 4 bytes per 32 rows
  if (dev_alloc(&d_flag, c->ntiles) || dev_alloc(&c->d_tile_base_flagged, c->ntiles + 2) || dev_alloc(&c->d_tiles_boundary, c->ntiles + 1) || dev_alloc(&d_count, 1)) return 1;
  k_classify_tiles<<<grid_for((c->ntiles + 1) * 32, 256), 256, 0, c->stream>>>(c->d_cols, c->d_tile_base, c->ntiles, (int)c->nrows, d_flag, c->d_tile_base_flagged);
  k_compact_tiles<<<1, 1024, 0, c->stream>>>(d_flag, c->ntiles, c->d_tiles_boundary, d_count);
  c->launches += 2;
  int nb = 0;
  CK(cudaMemcpyAsync(&nb, d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_flag); cudaFree(d_count);
  c->n_boundary = nb; c->n_interior = c->ntiles - nb;
  return 0;
}

extern "C" int mfe_make_local_matrix(mfe_ctx *c, int64_t *num_cols)
{
  if (!c || !c->has_structure) return fail("mfe_make_local_matrix: no matrix");
  CK(cudaSetDevice(c->device));
  PhaseTimer tm(c->stream);
  if (c->has_local) { if (num_cols) *num_cols = c->num_cols; return 0; }      // already local: keep the index maps as they are
  c->external_index.clear(); c->external_local_index.clear(); c->elements_to_send.clear();
  c->neighbors.clear(); c->recv_length.clear(); c->send_length.clear(); c->send_dst_offset.clear();
  if (c->nranks == 1) {              // make_local_matrix.hpp:51-55
    c->num_cols = c->nrows; c->has_local = true;
    c->t_make_local = tm.stop();
    if (num_cols) *num_cols = c->num_cols;
    return 0;
  }
  MeshDesc hm = c->m;                    // host view of the box table
  hm.boxes = c->boxes.data();
  LocalMaps L;
  compute_local_maps(hm, L);
  c->external_index = std::move(L.external_index); c->external_local_index = std::move(L.external_local_index);
  c->elements_to_send = std::move(L.elements_to_send); c->neighbors = std::move(L.neighbors);
  c->recv_length = std::move(L.recv_length); c->send_length = std::move(L.send_length);
  c->send_dst_offset = std::move(L.send_dst_offset);
  c->num_cols = L.num_cols;
  // the peer kernels take the neighbour tables by value (PEER_MAX_NEIGHBORS entries): more neighbours than that
  // (unaligned boxes at np >= 28, add_imbalance) go over NCCL, which has no such limit
  c->peer_fits_local = c->neighbors.size() <= PEER_MAX_NEIGHBORS;
  if (peer_agree(c)) return 1;
  const int n_ext = (int)c->external_index.size();
  // device: rewrite the columns (owned: col - start_row; external: external_local_index, looked up by global row)
  int *d_sg = nullptr, *d_sl = nullptr;
  if (dev_alloc(&d_sg, n_ext + 1) || dev_alloc(&d_sl, n_ext + 1)) return 1;
  if (n_ext > 0) {
    CK(cudaMemcpyAsync(d_sg, L.sorted_global.data(), sizeof(int) * n_ext, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_sl, L.sorted_local.data(), sizeof(int) * n_ext, cudaMemcpyHostToDevice, c->stream));
  }
  const int start_row = c->m.first_row, stop_row = c->m.first_row + (int)c->nrows - 1;
  k_localize_cols<<<c->sm_count * 8, 256, 0, c->stream>>>(c->d_cols, c->nnz, start_row, stop_row, d_sg, d_sl, n_ext);
  c->launches++;
  free_dev(c->d_elements_to_send); free_dev(c->d_send_buf);
  const int nsend = (int)c->elements_to_send.size();
  if (dev_alloc(&c->d_elements_to_send, nsend + 1) || dev_alloc(&c->d_send_buf, nsend + 1)) return 1;
  CK(cudaMemcpyAsync(c->d_elements_to_send, c->elements_to_send.data(), sizeof(int) * nsend, cudaMemcpyHostToDevice, c->stream));
  if (ensure_vectors(c, c->nrows, c->num_cols)) return 1;
  if (classify_tiles(c)) return 1;
  c->t_make_local = tm.stop();
  cudaFree(d_sg); cudaFree(d_sl);
  CK(cudaGetLastError());
  c->has_local = true;
  if (num_cols) *num_cols = c->num_cols;
  return 0;
}

// ref/src/SparseMatrix_functions.hpp:319-408 (driver.hpp:278-282, mv_overlap_comm_comp=1)
extern "C" int mfe_rearrange_matrix_local_external(mfe_ctx *c)
{
  if (!c || !c->has_structure || !c->has_local) return fail("mfe_rearrange_matrix_local_external: matrix needs local indices (mfe_make_local_matrix)");
  CK(cudaSetDevice(c->device));
  k_sort_rows_local<<<grid_for(c->nrows, 128), 128, 0, c->stream>>>(c->d_cols, c->d_coefs, c->d_roff, c->d_tile_base, c->tr, c->nrows);
  c->launches++;
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// kernels launchers
// ---------------------------------------------------------------------------------------------------------
// SpMV configurations: warps per CTA (= tiles in flight per SM), smem stages per warp, gather batches per row
struct MvCfg { int g, s, h; };
static const MvCfg matvec_configs[] = { {16, 1, 2}, {16, 1, 1}, {10, 2, 1}, {20, 1, 2}, {21, 1, 3} };
enum { N_MV_CFG = sizeof(matvec_configs) / sizeof(matvec_configs[0]) };

template <int G, int S, int H, int MODE>
static int launch_matvec_mode(mfe_ctx *c, const MvArgs &a)
{
  static bool attr_set[64] = { false };
  const size_t smem = (size_t)G * S * MatvecCfg::SLOT_BYTES + (size_t)G * S * 8 + 8 * 8 + 8 * 4 + 8 * 4 + 64;
  auto kern = k_matvec<G, S, H, MODE>;
  if (set_smem_attr_once(kern, attr_set, c->device, smem)) return 1;
  const int64_t nwork = MODE == MV_LIST ? a.nlist : c->ntiles;
  if (nwork <= 0) return 0;
  const int grid = (int)std::min<int64_t>(c->sm_count, std::max<int64_t>(1, (nwork + G - 1) / G));
  CK(launch_cg(c, kern, dim3(grid), dim3(32 * G), smem, a));
  return 0;
}

template <int G, int S, int H>
static int launch_matvec_cfg(mfe_ctx *c, const MvArgs &a, int mode)
{
  switch (mode) {
    case MV_ALL: return launch_matvec_mode<G, S, H, MV_ALL>(c, a);
    case MV_SKIP: return launch_matvec_mode<G, S, H, MV_SKIP>(c, a);
    case MV_LIST: return launch_matvec_mode<G, S, H, MV_LIST>(c, a);
    case MV_FUSED: return launch_matvec_mode<G, S, H, MV_FUSED>(c, a);
  }
  return fail("bad matvec mode %d", mode);
}

// one SpMV launch over this rank's matrix in the given mode (no halo exchange here except inside MV_FUSED)
static int launch_matvec_local(mfe_ctx *c, const MvArgs &a_in, int mode)
{
  // p.Ap prediction sums: the single-GPU pass reduces each round (= one 512-row piece) in shared memory; the interior pass
  // of a multi-rank SpMV uses one global atomic per tile instead -- measured on the 400x400x200 share of a 2-rank 400^3 run:
  // the shared-memory ring couples the 16 warps of a CTA to within 8 rounds of each other and cost the fused kernel 35 %
  // (2.17 vs 1.60 ms) where the tile costs vary more
  MvArgs a = a_in;
  a.use_ring = (c->mv_ring && mode == MV_ALL) ? 1 : 0;
  switch (c->mv_cfg) {
    case 0: return launch_matvec_cfg<16, 1, 2>(c, a, mode);
    case 1: return launch_matvec_cfg<16, 1, 1>(c, a, mode);
    case 2: return launch_matvec_cfg<10, 2, 1>(c, a, mode);
    case 3: return launch_matvec_cfg<20, 1, 2>(c, a, mode);
    case 4: return launch_matvec_cfg<21, 1, 3>(c, a, mode);
  }
  return fail("bad matvec configuration %d", c->mv_cfg);
}

static MvArgs mv_args(mfe_ctx *c, const double *x, double *y, bool with_sums, const CgScalars *gate, int k, bool seed_y)
{
  MvArgs a{};
  a.coefs = c->d_coefs; a.cols = c->d_cols; a.roff = c->d_roff;
  a.tile_base = c->d_tile_base; a.tile_base_flagged = c->d_tile_base_flagged;
  a.x = x; a.y = y; a.nrows = c->nrows; a.ntiles = c->ntiles; a.nnz_padded = c->nnz_padded; a.zmask = 0;
  a.piece_sums = with_sums ? c->dot.piece_sums_b : nullptr; a.chunk_sums = with_sums ? c->dot.chunk_sums_b : nullptr;
  a.gate = gate; a.k = k; a.tile_list = c->d_tiles_boundary; a.nlist = c->n_boundary; a.ell_width = c->ell_width; a.seed_y = seed_y ? 1 : 0; a.use_ring = c->mv_ring;
  return a;
}

// halo exchange on the context stream: ref/src/exchange_externals.hpp:44-161
static bool use_peer(const mfe_ctx *c) { return c->nranks > 1 && c->comm_mode == 1 && c->peer.ok && c->peer_fits; }

// NCCL path: pack + grouped send/recv per neighbour (complete when the stream reaches the next kernel)
static int exchange_nccl(mfe_ctx *c, double *x, const CgScalars *gate, int k)
{
  const int nsend = (int)c->elements_to_send.size();
  if (nsend > 0) CK(launch_cg(c, k_pack, dim3(grid_for(nsend, 256)), dim3(256), 0, x, c->d_elements_to_send, nsend, c->d_send_buf, gate, k));
  CKN(g_nccl.GroupStart());
  int64_t ro = c->nrows, so = 0;
  for (size_t i = 0; i < c->neighbors.size(); ++i) {
    if (c->recv_length[i] > 0) CKN(g_nccl.Recv(x + ro, (size_t)c->recv_length[i], NCCL_FLOAT64, c->neighbors[i], c->comm, c->stream));
    if (c->send_length[i] > 0) CKN(g_nccl.Send(c->d_send_buf + so, (size_t)c->send_length[i], NCCL_FLOAT64, c->neighbors[i], c->comm, c->stream));
    ro += c->recv_length[i]; so += c->send_length[i];
  }
  CKN(g_nccl.GroupEnd());
  return 0;
}

// peer mode: where my boundary values go (PeerHalo) and whose flags I wait for (PeerWait) in halo epoch `epoch`
static void peer_halo_args(mfe_ctx *c, unsigned long long epoch, PeerHalo &ph, PeerWait &pw)
{
  ph = PeerHalo{}; pw = PeerWait{};
  int n = 0, so = 0;
  for (size_t i = 0; i < c->neighbors.size(); ++i) {
    if (c->send_length[i] <= 0) continue;
    const int p = c->neighbors[i];
    ph.start[n] = so;
    ph.dst[n] = peer_halo_buf(c, p, epoch) + c->send_dst_offset[i];
    ph.flag[n] = peer_halo_flags(c, p) + c->rank;
    so += c->send_length[i]; ++n;
  }
  ph.start[n] = so; ph.n = n;
  n = 0;
  for (size_t i = 0; i < c->neighbors.size(); ++i)
    if (c->recv_length[i] > 0) pw.flag[n++] = peer_halo_flags(c, c->rank) + c->neighbors[i];
  pw.n = n;
}

// first half: peer mode stores my boundary values into the neighbours' halo buffers; NCCL mode does everything
static int exchange_begin(mfe_ctx *c, double *x, const CgScalars *gate, int k)
{
  if (c->nranks == 1 || c->test_no_exchange) return 0;
  if (!c->comm) return fail("exchange_externals: nranks > 1 but mfe_comm_init was not called");
  if (!use_peer(c)) return exchange_nccl(c, x, gate, k);
  const unsigned long long epoch = ++c->peer.halo_epoch;
  const int nsend = (int)c->elements_to_send.size();
  if (nsend == 0) return 0;
  PeerHalo ph; PeerWait pw;
  peer_halo_args(c, epoch, ph, pw);
  CK(launch_cg(c, k_pack_peer, dim3(grid_for(nsend, 256)), dim3(256), 0, x, c->d_elements_to_send, nsend, ph, epoch, c->peer.d_ticket, gate, k));
  return 0;
}
// second half (peer mode): wait for the neighbours' stores of this epoch and fill the external tail of x
static int exchange_end(mfe_ctx *c, double *x, const CgScalars *gate, int k)
{
  if (!use_peer(c) || c->test_no_exchange) return 0;
  const int n_ext = (int)(c->num_cols - c->nrows);
  if (n_ext <= 0) return 0;
  const unsigned long long epoch = c->peer.halo_epoch;
  PeerHalo ph; PeerWait pw;
  peer_halo_args(c, epoch, ph, pw);
  const int grid = std::max(1, std::min(32, (n_ext + 2047) / 2048));
  CK(launch_cg(c, k_halo_wait_copy, dim3(grid), dim3(256), 0, pw, epoch, peer_halo_buf(c, c->rank, epoch), x + c->nrows, n_ext,
               (unsigned long long)c->peer_timeout_ms * 1000000ULL, c->peer.d_err, gate ? c->d_sc : nullptr, gate, k));
  return 0;
}
static int exchange(mfe_ctx *c, double *x, const CgScalars *gate, int k)
{
  if (exchange_begin(c, x, gate, k)) return 1;
  return exchange_end(c, x, gate, k);
}

// y = A x including the halo exchange (matvec_std::operator(), SparseMatrix_functions.hpp:546-578).
// overlap 1 (default): one fused launch (MV_FUSED); overlap 2: pack / interior pass / wait + copy / boundary pass as four
// launches (the round-1 scheme, kept for comparison); overlap 0 or NCCL: exchange, then one pass over all tiles.
// fill_tail: the caller reads the external tail of x afterwards (API calls; the CG loop never does).
static int launch_matvec(mfe_ctx *c, double *x, double *y, bool with_sums, const CgScalars *gate, int k, bool seed_y, bool fill_tail)
{
  MvArgs a = mv_args(c, x, y, with_sums, gate, k, seed_y);
  const bool two_sets = c->n_interior > 0 && c->n_boundary > 0;
  if ((use_peer(c) || (c->test_no_exchange && c->nranks > 1)) && c->overlap == 1 && two_sets) {
    a.halo = x + c->nrows;                                  // test hook: the caller supplied the externals in the tail of x
    a.dbg_no_ext = c->test_no_exchange ? c->dbg_no_ext : 0;
    if (use_peer(c) && !c->test_no_exchange) {
      if (!c->comm) return fail("matvec: nranks > 1 but mfe_comm_init was not called");
      const unsigned long long epoch = ++c->peer.halo_epoch;
      peer_halo_args(c, epoch, a.ph, a.pw);
      a.halo = peer_halo_buf(c, c->rank, epoch);
      a.elements_to_send = c->d_elements_to_send; a.nsend = (int)c->elements_to_send.size();
      a.epoch = epoch; a.timeout_ns = (unsigned long long)c->peer_timeout_ms * 1000000ULL;
      a.ticket = c->peer.d_ticket; a.err = c->peer.d_err; a.sc = gate ? c->d_sc : nullptr;
    }
    if (launch_matvec_local(c, a, MV_FUSED)) return 1;
    return fill_tail ? exchange_end(c, x, gate, k) : 0;
  }
  if (exchange_begin(c, x, gate, k)) return 1;
  if ((use_peer(c) || c->test_no_exchange) && c->overlap == 2 && two_sets) {
    if (launch_matvec_local(c, a, MV_SKIP)) return 1;       // skips the tiles that need the halo
    if (exchange_end(c, x, gate, k)) return 1;
    return launch_matvec_local(c, a, MV_LIST);
  }
  if (exchange_end(c, x, gate, k)) return 1;
  return launch_matvec_local(c, a, MV_ALL);
}

static void launch_waxpby(mfe_ctx *c, int64_t n, ScalarArg a, const double *x, ScalarArg b, const double *y, double *w, const CgScalars *gate, int k,
                          cudaStream_t st = nullptr)
{
  const int64_t n2 = std::max<int64_t>(1, n / 2);
  const int grid = (int)std::min<int64_t>((n2 + 255) / 256, (int64_t)c->sm_count * 16);
  launch_on(c, st ? st : c->stream, k_waxpby, dim3(grid), dim3(256), 0, n, a, x, b, y, w, gate, k);
}

template <bool SAME, int NW>
static int launch_dot_maps_cfg(mfe_ctx *c, int64_t n, const double *x, const double *y, const DotWork &w, const CgScalars *gate, int k)
{
  static bool attr_set[64] = { false };
  const size_t smem = DotMapsCfg<SAME, NW>::SMEM_BYTES;
  auto kern = k_dot_maps<SAME, NW>;
  if (set_smem_attr_once(kern, attr_set, c->device, smem)) return 1;
  const int64_t npieces = (n + DOT_PIECE - 1) / DOT_PIECE;
  if (npieces == 0) return 0;
  // one persistent CTA per SM owning a contiguous range of pieces (more CTAs only when a range would overflow the table)
  int64_t grid = std::min<int64_t>(c->sm_count, npieces);
  grid = std::max<int64_t>(grid, (npieces + DOT_MAPS_MAX_PP - 1) / DOT_MAPS_MAX_PP);
  const int PP = (int)((npieces + grid - 1) / grid);
  grid = (npieces + PP - 1) / PP;
  CK(launch_cg(c, kern, dim3((unsigned)grid), dim3(32 * NW), smem, n, x, y, w, (int)npieces, PP, gate, k));
  return 0;
}

static int launch_dot_maps(mfe_ctx *c, int64_t n, const double *x, const double *y, const DotWork &w, const CgScalars *gate, int k)
{
  if ((((uintptr_t)x) | ((uintptr_t)y)) & 15) return fail("dot: operands must be 16-byte aligned");
  // x.x dots stage 4 KB per warp: 32 warps (64 registers) fit and are ~15 us faster per dot than 24 (the kernel is bound by
  // dependent-issue latency, so warps count); x.y dots stage 8 KB per warp: 24
  int nw = c->dot_maps_warps ? c->dot_maps_warps : (x == y ? 32 : 24);
  if (nw == 32 && x != y) nw = 24;
  if (x == y) return nw == 32 ? launch_dot_maps_cfg<true, 32>(c, n, x, y, w, gate, k) : nw == 24 ? launch_dot_maps_cfg<true, 24>(c, n, x, y, w, gate, k) : launch_dot_maps_cfg<true, 16>(c, n, x, y, w, gate, k);
  return nw == 24 ? launch_dot_maps_cfg<false, 24>(c, n, x, y, w, gate, k) : launch_dot_maps_cfg<false, 16>(c, n, x, y, w, gate, k);
}

// local dot of x,y (n terms) + rank-ordered global sum + CG scalar epilogue.
// sums_kind: where the prediction sums come from: 0 = nowhere yet (k_chunk_sums here), 1 = the r update (pair a),
//            2 = the SpMV (pair b)
static int launch_dot(mfe_ctx *c, int64_t n, const double *x, const double *y, int sums_kind, int epi, int k, bool gated, double *hist)
{
  const CgScalars *gate = gated ? c->d_sc : nullptr;
  CgScalars *sc = c->d_sc;
  double *nh = hist, *rh = hist ? hist + c->hist_cap : nullptr, *ph = hist ? hist + 2 * c->hist_cap : nullptr, *ah = hist ? hist + 3 * c->hist_cap : nullptr;
  const int nchunks = (int)((n + DOT_CHUNK - 1) / DOT_CHUNK);
  const bool multi = c->nranks > 1;
  const bool peer = use_peer(c);
  if (multi && !c->comm) return fail("dot: nranks > 1 but mfe_comm_init was not called");
  const int epi_local = multi ? EPI_NONE : epi;
  PeerDot pd{};                                   // nranks == 0: no peer all-reduce inside the kernel
  if (peer) {
    pd.nranks = c->nranks; pd.rank = c->rank; pd.epoch = ++c->peer.dot_epoch;
    pd.timeout_ns = (unsigned long long)c->peer_timeout_ms * 1000000ULL; pd.err = c->peer.d_err;
    for (int p = 0; p < c->nranks; ++p) pd.box[p] = peer_dot_slots(c, p);
  }
  DotWork w = c->dot.w;
  w.chunk_sums = sums_kind == 2 ? c->dot.chunk_sums_b : c->dot.chunk_sums_a;
  w.piece_sums = sums_kind == 2 ? c->dot.piece_sums_b : c->dot.piece_sums_a;
  bool finished = false;                          // global sum + epilogue already done by the local kernel
  if (c->dot_mode == MFE_DOT_SERIAL) {
    CK(launch_cg(c, k_dot_serial, dim3(1), dim3(1), 0, n, x, y, w.result, sc, epi_local, k, nh, rh, ph, ah, gate));
  } else {
    if (sums_kind == 0 && nchunks > 0) CK(launch_cg(c, k_chunk_sums, dim3(nchunks), dim3(DOT_CTA_THREADS), 0, n, x, y, w.piece_sums, w.chunk_sums, gate, k));
    if (c->dot_mode == MFE_DOT_TREE) {
      CK(launch_cg(c, k_chunk_prefix, dim3(1), dim3(1024), 0, w.chunk_sums, nchunks, w.chunk_prefix, gate, k));
      CK(launch_cg(c, k_dot_tree_finish, dim3(1), dim3(1), 0, w.chunk_prefix, nchunks, w.result, sc, epi_local, k, nh, rh, ph, ah, gate));
    } else {
      WalkArgs wa{};
      wa.sc = sc; wa.epi = peer ? epi : epi_local; wa.k = k; wa.normr_hist = nh; wa.rtrans_hist = rh; wa.pap_hist = ph; wa.alpha_hist = ah; wa.pd = pd;
      if (launch_dot_maps(c, n, x, y, w, gate, k)) return 1;
      CK(launch_cg(c, k_dot_walk, dim3(1), dim3(WALK_THREADS), 0, n, x, y, w, gate, wa));
      finished = peer;
    }
  }
  if (multi && peer && !finished) {
    CK(launch_cg(c, k_dot_peer_finish, dim3(1), dim3(32), 0, w.result, sc, epi, k, nh, rh, ph, ah, gate, pd));
  } else if (multi && !peer) {
    CKN(g_nccl.AllGather(w.result, c->d_gather, 1, NCCL_FLOAT64, c->comm, c->stream));
    CK(launch_cg(c, k_dot_allreduce_finish, dim3(1), dim3(1), 0, c->d_gather, c->nranks, w.result, sc, epi, k, nh, rh, ph, ah, gate));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// cg_solve: ref/src/cg_solve.hpp:69-229
// ---------------------------------------------------------------------------------------------------------
struct OpTimer {          // CUDA-event pairs per op class from the context's pool, resolved after the solve (no host sync inside)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[4];     // WAXPY, DOT, MATVEC, MATVECDOT
  mfe_ctx *c; int mode;          // 0: none, 1: every op class, 2: MATVEC only (two events per iteration)
  OpTimer(mfe_ctx *ctx, int m) : c(ctx), mode(m) {}
  bool on(int cls) const { return mode == 1 || (mode == 2 && (cls == MFE_MATVEC || cls == MFE_MATVECDOT)); }
  void begin(int cls, cudaStream_t st = nullptr) { if (!on(cls)) return; cudaEvent_t a = c->events.get(), b = c->events.get(); cudaEventRecord(a, st ? st : c->stream); ev[cls].push_back({ a, b }); }
  void end(int cls, cudaStream_t st = nullptr) { if (!on(cls)) return; cudaEventRecord(ev[cls].back().second, st ? st : c->stream); }
  double total_s(int cls) {
    double t = 0;
    for (auto &p : ev[cls]) { float ms = 0; if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) t += ms; }
    ev[cls].clear();
    return t * 1e-3;
  }
};

// a peer never delivered (bounded spins in peer.cuh): report instead of hanging
static int peer_error(mfe_ctx *c)
{
  if (!c->peer.ok || !c->peer.d_err) return 0;
  int e = 0;
  if (cudaMemcpy(&e, c->peer.d_err, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return fail("peer_error: %s", cudaGetErrorString(cudaGetLastError()));
  if (e == 0) return 0;
  cudaMemset(c->peer.d_err, 0, 4);
  return fail("NVLink peer %s timed out after %lld ms on rank %d (a neighbour rank died or fell behind)", e == 1 ? "halo exchange" : "dot all-reduce",
              (long long)c->peer_timeout_ms, c->rank);
}

static int read_scalars(mfe_ctx *c, CgScalars *h)
{
  CK(cudaMemcpyAsync(h, c->d_sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int mfe_cg_solve(mfe_ctx *c, double *x_host, int max_iter, double tol, int *num_iters, double *normr, double cg_times[MFE_NUM_TIMERS],
                            double *normr_hist, double *rtrans_hist, double *pap_hist, double *alpha_hist)
{
  if (!c || !c->has_structure) return fail("mfe_cg_solve: no matrix");
  if (!c->has_local) {   // cg_solve.hpp:99-104
    fprintf(stderr, "miniFE::cg_solve ERROR, A.has_local_indices is false, needs to be true. This probably means "
                    "miniFE::make_local_matrix(A) was not called prior to calling miniFE::cg_solve.\n");
    return fail("mfe_cg_solve: make_local_matrix was not called");
  }
  if (max_iter < 0) return fail("mfe_cg_solve: max_iter < 0");
  CK(cudaSetDevice(c->device));
  if (ensure_vectors(c, c->nrows, c->num_cols)) return 1;
  const int64_t n = c->nrows;
  if (c->hist_cap < max_iter + 1) {
    free_dev(c->d_hist);
    c->hist_cap = max_iter + 1;
    if (dev_alloc(&c->d_hist, 4 * (int64_t)c->hist_cap)) return 1;
  }
  c->events.reset();
  cudaEvent_t t0 = c->events.get(), t1 = c->events.get();
  cudaEventRecord(t0, c->stream);
  if (x_host) CK(cudaMemcpyAsync(c->d_x, x_host, n * 8, cudaMemcpyHostToDevice, c->stream));
  k_fill<<<64, 256, 0, c->stream>>>(c->d_hist, 4 * (int64_t)c->hist_cap, std::numeric_limits<double>::quiet_NaN());
  c->launches++;
  CgScalars h{};
  h.tol = tol; h.brkdown_tol = std::numeric_limits<double>::epsilon();
  h.one = 1.0; h.zero = 0.0; h.neg_one = -1.0;
  h.stop_iter = max_iter; h.status = CG_RUNNING; h.max_iter = max_iter;
  CK(cudaMemcpyAsync(c->d_sc, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
  // r, p (num_cols), Ap start zero-filled (cg_solve.hpp:109-111)
  CK(cudaMemsetAsync(c->d_r, 0, n * 8, c->stream));
  CK(cudaMemsetAsync(c->d_p, 0, c->num_cols * 8, c->stream));
  CK(cudaMemsetAsync(c->d_Ap, 0, n * 8, c->stream));
  {   // the SpMV accumulates its p.Ap prediction sums into pair b; the r update re-zeroes them every iteration
    const int64_t nch = (n + DOT_CHUNK - 1) / DOT_CHUNK;
    CK(cudaMemsetAsync(c->dot.chunk_sums_b, 0, 8 * (nch + 1), c->stream));
    CK(cudaMemsetAsync(c->dot.piece_sums_b, 0, 8 * (nch * DOT_PIECES_PER_CHUNK + 1), c->stream));
  }

  OpTimer ot(c, c->op_timing);
  const ScalarArg one{ nullptr, 1.0 }, zero{ nullptr, 0.0 }, neg_one{ nullptr, -1.0 };
  CgScalars *sc = c->d_sc;
  double *hist = c->d_hist;
  // x += alpha*p is the only update of the iteration that nothing waits for (cg_solve.hpp:211 vs 212): with x_overlap it
  // runs on the low-priority stream beside the latency-bound dot kernels; p must not change before it has been read
  const bool xov = c->x_overlap != 0;
  bool x_pending = false;
  auto join_x = [&]() { if (x_pending) { cudaStreamWaitEvent(c->stream, c->ev_xdone, 0); x_pending = false; } };
  // x and r updates of iteration k (fused_waxpby, cg_solve.hpp:211-212): the r update also leaves the r.r prediction sums
  auto launch_updates = [&](int k) -> int {
    ot.begin(MFE_WAXPY);
    if (xov) {
      CK(launch_cg(c, k_xr_update<false>, dim3(grid_for(n, DOT_CHUNK)), dim3(DOT_CTA_THREADS), 0, n, sc, c->d_p, c->d_Ap, c->d_x, c->d_r, c->dot.piece_sums_a,
                   c->dot.chunk_sums_a, c->dot.piece_sums_b, c->dot.chunk_sums_b, sc, k));
      ot.end(MFE_WAXPY);
      cudaEventRecord(c->ev_alpha, c->stream);             // alpha of iteration k is final on the main stream from here on
      cudaStreamWaitEvent(c->stream_x, c->ev_alpha, 0);
      ot.begin(MFE_WAXPY, c->stream_x);
      launch_waxpby(c, n, one, c->d_x, ScalarArg{ &sc->alpha, 0.0 }, c->d_p, c->d_x, sc, k, c->stream_x);
      ot.end(MFE_WAXPY, c->stream_x);
      cudaEventRecord(c->ev_xdone, c->stream_x);
      x_pending = true;
    } else {
      CK(launch_cg(c, k_xr_update<true>, dim3(grid_for(n, DOT_CHUNK)), dim3(DOT_CTA_THREADS), 0, n, sc, c->d_p, c->d_Ap, c->d_x, c->d_r, c->dot.piece_sums_a,
                   c->dot.chunk_sums_a, c->dot.piece_sums_b, c->dot.chunk_sums_b, sc, k));
      ot.end(MFE_WAXPY);
    }
    return 0;
  };

  ot.begin(MFE_WAXPY); launch_waxpby(c, n, one, c->d_x, zero, c->d_x, c->d_p, nullptr, 0); ot.end(MFE_WAXPY);           // p = x        (:124)
  ot.begin(MFE_MATVEC); if (launch_matvec(c, c->d_p, c->d_Ap, false, nullptr, 0, false, false)) return 1; ot.end(MFE_MATVEC);         // Ap = A p     (:129)
  ot.begin(MFE_WAXPY); launch_waxpby(c, n, one, c->d_b, neg_one, c->d_Ap, c->d_r, nullptr, 0); ot.end(MFE_WAXPY);       // r = b - Ap   (:132)
  ot.begin(MFE_DOT); if (launch_dot(c, n, c->d_r, c->d_r, 0, EPI_INIT_RR, 0, false, hist)) return 1; ot.end(MFE_DOT);   // rtrans       (:134)

  int rc = 0;
  int k = 1;
  int resume_phase = 0;      // 0: start of iteration k, 1: after the p_ap dot of iteration k (breakdown pause resolved)
  while (k <= max_iter) {
    const int k_batch_end = std::min(max_iter, k + 24);
    for (; k <= k_batch_end; ++k) {
      if (resume_phase == 0) {
        if (k == 1) {
          ot.begin(MFE_WAXPY); launch_waxpby(c, n, one, c->d_r, zero, c->d_r, c->d_p, sc, k); ot.end(MFE_WAXPY);        // p = r (:154)
        } else {
          ot.begin(MFE_DOT); if (launch_dot(c, n, c->d_r, c->d_r, 1, EPI_RR, k, true, hist)) return 1; ot.end(MFE_DOT); // (:158)
          join_x();                                                                                                      // p is about to change
          ot.begin(MFE_WAXPY); launch_waxpby(c, n, one, c->d_r, ScalarArg{ &sc->beta, 0.0 }, c->d_p, c->d_p, sc, k); ot.end(MFE_WAXPY);  // (:160)
        }
        // MINIFE_FUSED (cg_solve.hpp:172-176): matvec_and_dot = the same SpMV and the same serial-order dot (x[row]*sum
        // commutes), timed as one MATVECDOT interval; the SpMV always feeds the dot its prediction sums, fused or not
        const int mv_cls = c->fused ? MFE_MATVECDOT : MFE_MATVEC, pap_cls = c->fused ? MFE_MATVECDOT : MFE_DOT;
        ot.begin(mv_cls); if (launch_matvec(c, c->d_p, c->d_Ap, c->dot_mode != MFE_DOT_SERIAL, sc, k, false, false)) return 1; if (!c->fused) ot.end(mv_cls);   // (:177)
        if (!c->fused) ot.begin(pap_cls);
        if (launch_dot(c, n, c->d_Ap, c->d_p, c->dot_mode != MFE_DOT_SERIAL ? 2 : 0, EPI_PAP, k, true, hist)) return 1; ot.end(pap_cls);  // (:179)
      }
      resume_phase = 0;
      if (launch_updates(k)) return 1;                                                                                   // (:211-212)
    }
    join_x();
    if (read_scalars(c, &h)) return 1;
    if (h.status == CG_PAUSED) {
      // cg_solve.hpp:186-200: p_ap_dot < brkdown_tol and >= 0  ->  breakdown(p_ap_dot, Ap, p)  (cg_solve.hpp:49-67)
      const int kp = h.pause_iter;
      double vv = 0, ww = 0;
      if (launch_dot(c, n, c->d_Ap, c->d_Ap, 0, EPI_NONE, 0, false, nullptr)) return 1;
      CK(cudaMemcpyAsync(&vv, c->dot.w.result, 8, cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream));
      if (launch_dot(c, n, c->d_p, c->d_p, 0, EPI_NONE, 0, false, nullptr)) return 1;
      CK(cudaMemcpyAsync(&ww, c->dot.w.result, 8, cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream));
      const double vnorm = std::sqrt(vv), wnorm = std::sqrt(ww);
      const bool brk = std::fabs(h.p_ap_dot) <= 100 * vnorm * wnorm * std::numeric_limits<double>::epsilon();
      if (brk) { h.status = CG_BREAKDOWN; h.stop_iter = kp - 1; CK(cudaMemcpyAsync(c->d_sc, &h, sizeof h, cudaMemcpyHostToDevice, c->stream)); }
      else {
        h.brkdown_tol = 0.1 * h.p_ap_dot;
        h.alpha = h.rtrans / h.p_ap_dot; h.neg_alpha = -h.alpha;
        h.status = CG_RUNNING; h.stop_iter = max_iter;
        // the loop test that the convergence epilogue would have applied for iteration kp
        if (!(h.normr > h.tol)) { h.stop_iter = kp; h.status = CG_CONVERGED; }
        CK(cudaMemcpyAsync(c->d_sc, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_hist + 3 * c->hist_cap + kp, &h.alpha, 8, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        k = kp; resume_phase = 1;
        continue;
      }
    }
    if (h.status == CG_BREAKDOWN) {
      fprintf(stderr, "miniFE::cg_solve ERROR, numerical breakdown!\n");
      rc = 2;
      break;
    }
    if (h.status == CG_COMM_ERROR) break;
    if (h.status == CG_CONVERGED || h.stop_iter < k) break;
  }
  join_x();
  if (read_scalars(c, &h)) return 1;
  cudaEventRecord(t1, c->stream);
  if (x_host) CK(cudaMemcpyAsync(x_host, c->d_x, n * 8, cudaMemcpyDeviceToHost, c->stream));
  std::vector<double> hh(4 * (size_t)c->hist_cap);
  CK(cudaMemcpyAsync(hh.data(), c->d_hist, hh.size() * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  float ms = 0; cudaEventElapsedTime(&ms, t0, t1);
  c->t_cg = ms;
  const int iters = std::min(h.stop_iter, max_iter);
  if (num_iters) *num_iters = iters;
  if (normr) *normr = h.normr;
  if (cg_times) {
    cg_times[MFE_WAXPY] = ot.total_s(MFE_WAXPY); cg_times[MFE_DOT] = ot.total_s(MFE_DOT); cg_times[MFE_MATVEC] = ot.total_s(MFE_MATVEC);
    cg_times[MFE_MATVECDOT] = ot.total_s(MFE_MATVECDOT); cg_times[MFE_TOTAL] = ms * 1e-3;
  }
  const size_t cap = c->hist_cap, cnt = (size_t)max_iter + 1;
  if (normr_hist) memcpy(normr_hist, hh.data(), cnt * 8);
  if (rtrans_hist) memcpy(rtrans_hist, hh.data() + cap, cnt * 8);
  if (pap_hist) memcpy(pap_hist, hh.data() + 2 * cap, cnt * 8);
  if (alpha_hist) memcpy(alpha_hist, hh.data() + 3 * cap, cnt * 8);
  CK(cudaGetLastError());
  if (peer_error(c)) return 1;
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// time_kernels: ref/src/time_kernels.hpp:42-134 (MINIFE_KERNELS != 0, driver.hpp:257-270): max_iter x waxpby(1,x,0,x,p),
// max_iter x matvec(A,p,x), xdotp = sum of max_iter x dot(x,p) -- the same sequences on the device-resident vectors
// ---------------------------------------------------------------------------------------------------------
extern "C" int mfe_time_kernels(mfe_ctx *c, double *x_host, int max_iter, double *xdotp, double kern_times[MFE_NUM_TIMERS])
{
  if (!c || !c->has_structure) return fail("mfe_time_kernels: no matrix");
  if (!c->has_local) {
    fprintf(stderr, "miniFE::time_kernels ERROR, A.has_local_indices is false, needs to be true. This probably means "
                    "miniFE::make_local_matrix(A) was not called prior to calling miniFE::time_kernels.\n");
    return fail("mfe_time_kernels: make_local_matrix was not called");
  }
  if (max_iter < 0) return fail("mfe_time_kernels: max_iter < 0");
  CK(cudaSetDevice(c->device));
  if (ensure_vectors(c, c->nrows, c->num_cols)) return 1;
  const int64_t n = c->nrows;
  if (x_host) CK(cudaMemcpyAsync(c->d_x, x_host, n * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(c->d_p, 0, c->num_cols * 8, c->stream));           // Vector p(0, ncols): zero-filled (time_kernels.hpp:75)
  double *d_vals = nullptr;
  if (dev_alloc(&d_vals, std::max(max_iter, 1))) return 1;
  const ScalarArg one{ nullptr, 1.0 }, zero{ nullptr, 0.0 };
  c->events.reset();
  cudaEvent_t ev[4]; for (auto &e : ev) e = c->events.get();
  cudaEventRecord(ev[0], c->stream);
  for (int i = 0; i < max_iter; ++i) launch_waxpby(c, n, one, c->d_x, zero, c->d_x, c->d_p, nullptr, 0);
  cudaEventRecord(ev[1], c->stream);
  for (int i = 0; i < max_iter; ++i) if (launch_matvec(c, c->d_p, c->d_x, false, nullptr, 0, true, false)) return 1;
  cudaEventRecord(ev[2], c->stream);
  for (int i = 0; i < max_iter; ++i) {
    if (launch_dot(c, n, c->d_x, c->d_p, 0, EPI_NONE, 0, false, nullptr)) return 1;
    CK(cudaMemcpyAsync(d_vals + i, c->dot.w.result, 8, cudaMemcpyDeviceToDevice, c->stream));
  }
  cudaEventRecord(ev[3], c->stream);
  std::vector<double> vals(std::max(max_iter, 1));
  CK(cudaMemcpyAsync(vals.data(), d_vals, 8 * (size_t)std::max(max_iter, 1), cudaMemcpyDeviceToHost, c->stream));
  if (x_host) CK(cudaMemcpyAsync(x_host, c->d_x, n * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  double acc = 0;                                  // xdotp += dot(x, p), time_kernels.hpp:119-122
  for (int i = 0; i < max_iter; ++i) acc += vals[i];
  if (xdotp) *xdotp = acc;
  float a = 0, b = 0, d = 0;
  cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]); cudaEventElapsedTime(&d, ev[2], ev[3]);
  cudaFree(d_vals);
  if (kern_times) { kern_times[MFE_WAXPY] = a * 1e-3; kern_times[MFE_MATVEC] = b * 1e-3; kern_times[MFE_DOT] = d * 1e-3; kern_times[MFE_MATVECDOT] = 0; kern_times[MFE_TOTAL] = 0; }
  CK(cudaGetLastError());
  return peer_error(c);
}

// ---------------------------------------------------------------------------------------------------------
// verify_solution: ref/fem/verify_solution.hpp:55-174 on the device-resident x
// ---------------------------------------------------------------------------------------------------------
extern "C" int mfe_verify_solution(mfe_ctx *c, double tolerance, int verify_whole_domain, int *return_code, double *global_max_err, double local_info[6])
{
  if (!c || !c->has_structure || !c->d_x) return fail("mfe_verify_solution: no solution vector (run mfe_generate_matrix_structure / mfe_cg_solve first)");
  CK(cudaSetDevice(c->device));
  const MeshDesc &m = c->m;
  // verified nodes: the condition is separable per axis (verify_solution.hpp:90-95), so they are a product of index sets
  std::vector<int> v[3];
  const int gn[3] = { m.gnx, m.gny, m.gnz };
  for (int a = 0; a < 3; ++a) {
    const int div = gn[a] > 1 ? gn[a] - 1 : 1;
    for (int i = m.nb[2 * a]; i < m.nb[2 * a + 1]; ++i) {
      const double coord = (1.0 * i) / div;
      if (verify_whole_domain || std::abs(coord - 0.5) < 0.05) v[a].push_back(i);
    }
  }
  const int64_t npts = (int64_t)v[0].size() * (int64_t)v[1].size() * (int64_t)v[2].size();
  double info[6] = { 0, 0, 0, 0, 0, 0 };      // err, computed, analytic, x, y, z  (err_info, verify_solution.hpp:47-53)
  if (npts > 0x7fffffffLL) return fail("mfe_verify_solution: %lld points", (long long)npts);
  if (npts > 0) {
    int *d_v[3] = { nullptr, nullptr, nullptr };
    double *d_l = nullptr, *d_s = nullptr, *d_err = nullptr, *d_an = nullptr;
    for (int a = 0; a < 3; ++a) {
      if (dev_alloc(&d_v[a], (int64_t)v[a].size())) return 1;
      CK(cudaMemcpyAsync(d_v[a], v[a].data(), sizeof(int) * v[a].size(), cudaMemcpyHostToDevice, c->stream));
    }
    if (dev_alloc(&d_l, VERIFY_NT * VERIFY_NT) || dev_alloc(&d_s, VERIFY_NT * VERIFY_NT) || dev_alloc(&d_err, npts) || dev_alloc(&d_an, npts)) return 1;
    k_verify_tables<<<grid_for(VERIFY_NT * VERIFY_NT, 256), 256, 0, c->stream>>>(d_l, d_s);
    VerifyAxes va; for (int a = 0; a < 3; ++a) va.n[a] = (int)v[a].size();
    k_verify<<<(unsigned)npts, VERIFY_THREADS, 0, c->stream>>>(m, va, d_v[0], d_v[1], d_v[2], c->d_x, d_l, d_s, d_err, d_an);
    c->launches += 2;
    std::vector<double> err(npts);
    CK(cudaMemcpyAsync(err.data(), d_err, 8 * npts, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int64_t best = -1;                            // first strictly larger error in row order (verify_solution.hpp:139-146)
    for (int64_t i = 0; i < npts; ++i) if (err[i] > info[0]) { info[0] = err[i]; best = i; }
    if (best >= 0) {
      const int ix = v[0][best % v[0].size()], iy = v[1][(best / v[0].size()) % v[1].size()], iz = v[2][best / (v[0].size() * v[1].size())];
      const int64_t lr = (int64_t)(ix - m.nb[0]) + (int64_t)m.lnx * ((iy - m.nb[2]) + (int64_t)m.lny * (iz - m.nb[4]));
      CK(cudaMemcpy(&info[1], c->d_x + lr, 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&info[2], d_an + best, 8, cudaMemcpyDeviceToHost));
      info[3] = (1.0 * ix) / (gn[0] > 1 ? gn[0] - 1 : 1); info[4] = (1.0 * iy) / (gn[1] > 1 ? gn[1] - 1 : 1); info[5] = (1.0 * iz) / (gn[2] > 1 ? gn[2] - 1 : 1);
    }
    for (int a = 0; a < 3; ++a) cudaFree(d_v[a]);
    cudaFree(d_l); cudaFree(d_s); cudaFree(d_err); cudaFree(d_an);
  }
  double gmax = info[0];
  if (c->nranks > 1) {                            // MPI_Allreduce(MAX), verify_solution.hpp:148-154
    if (!c->comm) return fail("mfe_verify_solution: nranks > 1 but mfe_comm_init was not called");
    std::vector<double> all(c->nranks);
    CK(cudaMemcpyAsync(c->d_tmp1, &info[0], 8, cudaMemcpyHostToDevice, c->stream));
    CKN(g_nccl.AllGather(c->d_tmp1, c->d_gather, 1, NCCL_FLOAT64, c->comm, c->stream));
    CK(cudaMemcpyAsync(all.data(), c->d_gather, 8 * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (double e : all) gmax = std::max(gmax, e);
  }
  if (local_info) memcpy(local_info, info, sizeof info);
  if (global_max_err) *global_max_err = gmax;
  if (return_code) *return_code = gmax > tolerance ? 1 : 0;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mfe_set_dot_mode(mfe_ctx *c, int mode)
{
  if (!c || mode < 0 || mode > 2) return fail("mfe_set_dot_mode: bad arguments");
  c->dot_mode = mode;
  return 0;
}
extern "C" int mfe_set_op_timing(mfe_ctx *c, int on) { if (!c) return fail("null context"); c->op_timing = on; return 0; }

extern "C" int mfe_set_option(mfe_ctx *c, const char *name, int64_t value)
{
  if (!c || !name) return fail("mfe_set_option: bad arguments");
  if (!strcmp(name, "matvec_config")) {
    if (value < 0 || value >= N_MV_CFG) return fail("mfe_set_option: matvec_config must be in [0,%d)", (int)N_MV_CFG);
    c->mv_cfg = (int)value;
    return 0;
  }
  if (!strcmp(name, "dot_maps_warps")) {
    if (value != 0 && value != 16 && value != 24 && value != 32) return fail("mfe_set_option: dot_maps_warps must be 0, 16, 24 or 32 (32: x.x dots only)");
    c->dot_maps_warps = (int)value;
    return 0;
  }
  if (!strcmp(name, "comm_mode")) {           // 1: NVLink peer stores (when every rank could open every mailbox), 0: NCCL
    if (value != 0 && value != 1) return fail("mfe_set_option: comm_mode must be 0 (NCCL) or 1 (peer)");
    c->comm_mode = (int)value;
    return 0;
  }
  if (!strcmp(name, "overlap")) {            // 0 none, 1 one fused launch (default), 2 interior / boundary passes as separate launches
    if (value < 0 || value > 2) return fail("mfe_set_option: overlap must be 0, 1 or 2");
    c->overlap = (int)value;
    return 0;
  }
  if (!strcmp(name, "fused")) { c->fused = value != 0; return 0; }
  if (!strcmp(name, "x_overlap")) { c->x_overlap = value != 0; return 0; }
  if (!strcmp(name, "assembly_mode")) { c->assembly_mode = value != 0; return 0; }
  if (!strcmp(name, "test_no_exchange")) { c->test_no_exchange = value != 0; return 0; }
  if (!strcmp(name, "dbg_no_ext")) { c->dbg_no_ext = (int)value; return 0; }
  if (!strcmp(name, "mv_ring")) { c->mv_ring = value != 0; return 0; }
  if (!strcmp(name, "ell_width")) {          // ref/src/ELLMatrix.hpp: 0 = CSR semantics, 27 = the width generate_matrix_structure reserves
    if (value < 0 || value > 64) return fail("mfe_set_option: ell_width must be in [0,64]");
    c->ell_width = (int)value;
    return 0;
  }
  if (!strcmp(name, "peer_timeout_ms")) { if (value < 1) return fail("mfe_set_option: peer_timeout_ms < 1"); c->peer_timeout_ms = value; return 0; }
  return fail("mfe_set_option: unknown option '%s'", name);
}

// ---------------------------------------------------------------------------------------------------------
// host-buffer kernel entry points
// ---------------------------------------------------------------------------------------------------------
extern "C" int mfe_matvec(mfe_ctx *c, double *x_host, double *y_host)
{
  if (!c || !c->has_structure || !c->has_local) return fail("mfe_matvec: matrix needs local indices (mfe_make_local_matrix)");
  if (!x_host || !y_host) return fail("mfe_matvec: null buffer");
  CK(cudaSetDevice(c->device));
  if (ensure_vectors(c, c->nrows, c->num_cols)) return 1;
  CK(cudaMemcpyAsync(c->d_tmp1, x_host, c->num_cols * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_tmp2, y_host, c->nrows * 8, cudaMemcpyHostToDevice, c->stream));      // beta*y[row] reads the caller's y
  if (launch_matvec(c, c->d_tmp1, c->d_tmp2, false, nullptr, 0, true, true)) return 1;
  CK(cudaMemcpyAsync(y_host, c->d_tmp2, c->nrows * 8, cudaMemcpyDeviceToHost, c->stream));
  if (c->nranks > 1 && !c->test_no_exchange) CK(cudaMemcpyAsync(x_host + c->nrows, c->d_tmp1 + c->nrows, (c->num_cols - c->nrows) * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  return peer_error(c);
}

extern "C" int mfe_exchange_externals(mfe_ctx *c, double *x_host)
{
  if (!c || !c->has_local) return fail("mfe_exchange_externals: matrix needs local indices");
  CK(cudaSetDevice(c->device));
  if (ensure_vectors(c, c->nrows, c->num_cols)) return 1;
  CK(cudaMemcpyAsync(c->d_tmp1, x_host, c->num_cols * 8, cudaMemcpyHostToDevice, c->stream));
  if (exchange(c, c->d_tmp1, nullptr, 0)) return 1;
  CK(cudaMemcpyAsync(x_host, c->d_tmp1, c->num_cols * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return peer_error(c);
}

extern "C" int mfe_waxpby(mfe_ctx *c, int64_t n, double alpha, const double *x_host, double beta, const double *y_host, double *w_host)
{
  if (!c || n < 0) return fail("mfe_waxpby: bad arguments");
  CK(cudaSetDevice(c->device));
  if (ensure_vectors(c, n, n)) return 1;
  double *dw = c->d_r;   // scratch: r is rebuilt by every cg_solve
  CK(cudaMemcpyAsync(c->d_tmp1, x_host, n * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_tmp2, y_host, n * 8, cudaMemcpyHostToDevice, c->stream));
  launch_waxpby(c, n, ScalarArg{ nullptr, alpha }, c->d_tmp1, ScalarArg{ nullptr, beta }, c->d_tmp2, dw, nullptr, 0);
  CK(cudaMemcpyAsync(w_host, dw, n * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  return 0;
}

extern "C" int mfe_dot(mfe_ctx *c, int64_t n, const double *x_host, const double *y_host, double *result)
{
  if (!c || n < 0 || !result) return fail("mfe_dot: bad arguments");
  CK(cudaSetDevice(c->device));
  if (ensure_vectors(c, n, n)) return 1;
  CK(cudaMemcpyAsync(c->d_tmp1, x_host, n * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_tmp2, y_host, n * 8, cudaMemcpyHostToDevice, c->stream));
  if (launch_dot(c, n, c->d_tmp1, c->d_tmp2, 0, EPI_NONE, 0, false, nullptr)) return 1;
  CK(cudaMemcpyAsync(result, c->dot.w.result, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  return peer_error(c);
}

// ---------------------------------------------------------------------------------------------------------
// queries / import / export
// ---------------------------------------------------------------------------------------------------------
extern "C" int64_t mfe_nrows(const mfe_ctx *c) { return c ? c->nrows : -1; }
extern "C" int64_t mfe_first_row(const mfe_ctx *c) { return c ? (c->nrows > 0 ? c->m.first_row : -1) : -1; }
extern "C" int64_t mfe_nnz(const mfe_ctx *c) { return c ? c->nnz : -1; }
extern "C" int64_t mfe_num_cols(const mfe_ctx *c) { return c ? c->num_cols : -1; }
extern "C" int64_t mfe_launch_count(const mfe_ctx *c) { return c ? c->launches : -1; }

extern "C" int64_t mfe_get_map(const mfe_ctx *c, const char *name, int *out)
{
  if (!c || !name) return -1;
  const std::vector<int> *v = nullptr;
  if (!strcmp(name, "external_index")) v = &c->external_index;
  else if (!strcmp(name, "external_local_index")) v = &c->external_local_index;
  else if (!strcmp(name, "elements_to_send")) v = &c->elements_to_send;
  else if (!strcmp(name, "neighbors")) v = &c->neighbors;
  else if (!strcmp(name, "recv_length")) v = &c->recv_length;
  else if (!strcmp(name, "send_length")) v = &c->send_length;
  else if (!strcmp(name, "send_dst_offset")) v = &c->send_dst_offset;
  else return -1;
  if (out && !v->empty()) memcpy(out, v->data(), v->size() * sizeof(int));
  return (int64_t)v->size();
}

extern "C" int mfe_export_csr(mfe_ctx *c, int *rows, int *row_offsets, int64_t *row_offsets64, int *cols, double *coefs)
{
  if (!c || !c->has_structure) return fail("mfe_export_csr: no matrix");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (rows) CK(cudaMemcpy(rows, c->d_rows, c->nrows * sizeof(int), cudaMemcpyDeviceToHost));
  if (row_offsets) {
    if (c->nnz > 0x7fffffffLL) return fail("mfe_export_csr: nnz=%lld overflows 32-bit row_offsets; use row_offsets64", (long long)c->nnz);
    CK(cudaMemcpy(row_offsets, c->d_roff, (c->nrows + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  }
  if (row_offsets64) {
    std::vector<uint32_t> lo(c->nrows + 1);
    std::vector<int64_t> tb(c->ntiles + 1);
    CK(cudaMemcpy(lo.data(), c->d_roff, (c->nrows + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tb.data(), c->d_tile_base, (c->ntiles + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    for (int64_t r = 0; r <= c->nrows; ++r) {
      const int64_t t = std::min(r / c->tr, c->ntiles - 1);
      row_offsets64[r] = tb[t] + (int64_t)(uint32_t)(lo[r] - lo[t * c->tr]);
    }
  }
  if (cols) CK(cudaMemcpy(cols, c->d_cols, c->nnz * sizeof(int), cudaMemcpyDeviceToHost));
  if (coefs) CK(cudaMemcpy(coefs, c->d_coefs, c->nnz * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

static double *vec_by_name(mfe_ctx *c, const char *name)
{
  if (!name) return nullptr;
  if (!strcmp(name, "b")) return c->d_b;
  if (!strcmp(name, "x")) return c->d_x;
  return nullptr;
}
extern "C" int mfe_export_vector(mfe_ctx *c, const char *name, double *out)
{
  if (!c || !out) return fail("mfe_export_vector: bad arguments");
  double *d = vec_by_name(c, name);
  if (!d) return fail("mfe_export_vector: unknown vector '%s'", name ? name : "(null)");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(out, d, c->nrows * 8, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int mfe_import_vector(mfe_ctx *c, const char *name, const double *in)
{
  if (!c || !in) return fail("mfe_import_vector: bad arguments");
  double *d = vec_by_name(c, name);
  if (!d) return fail("mfe_import_vector: unknown vector '%s'", name ? name : "(null)");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(d, in, c->nrows * 8, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int mfe_import_csr(mfe_ctx *c, int64_t nrows, int64_t num_cols, const int *row_offsets, const int *cols, const double *coefs)
{
  if (!c || nrows < 1 || num_cols < nrows || !row_offsets || !cols || !coefs) return fail("mfe_import_csr: bad arguments");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  free_matrix(c);
  c->mesh_csr = false;
  c->nrows = nrows; c->nnz = row_offsets[nrows]; c->num_cols = num_cols;
  c->nnz_padded = ((c->nnz + 3) & ~3LL) + 4;
  if (dev_alloc(&c->d_rows, nrows) || dev_alloc(&c->d_roff, nrows + 8) || dev_alloc(&c->d_cols, c->nnz_padded) || dev_alloc(&c->d_coefs, c->nnz_padded)) return 1;
  std::vector<int> rows(nrows);
  for (int64_t i = 0; i < nrows; ++i) rows[i] = c->m.first_row + (int)i;
  CK(cudaMemset(c->d_cols, 0, c->nnz_padded * sizeof(int)));
  CK(cudaMemset(c->d_coefs, 0, c->nnz_padded * sizeof(double)));
  CK(cudaMemcpy(c->d_rows, rows.data(), nrows * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_roff, row_offsets, (nrows + 1) * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_cols, cols, c->nnz * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_coefs, coefs, c->nnz * sizeof(double), cudaMemcpyHostToDevice));
  if (build_tile_base(c)) return 1;
  if (ensure_vectors(c, nrows, num_cols)) return 1;
  if (c->nranks > 1 && classify_tiles(c)) return 1;
  CK(cudaStreamSynchronize(c->stream));
  c->has_structure = true; c->has_local = true;
  return 0;
}

// exact-dot diagnostics (k_dot_walk): out[0..15] = counters, then DOT_STATS_TRACE pairs (slow gaps << 32 | dirty leaves, ns)
extern "C" int64_t mfe_dot_stats(mfe_ctx *c, uint64_t *out, int64_t cap, int reset)
{
  if (!c || !c->dot.w.stats) return -1;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  const int64_t nw = std::min<int64_t>(cap, DOT_STATS_WORDS);
  if (out && nw > 0 && cudaMemcpy(out, c->dot.w.stats, 8 * nw, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (reset) cudaMemset(c->dot.w.stats, 0, 8 * DOT_STATS_WORDS);
  return DOT_STATS_WORDS;
}

extern "C" double mfe_phase_ms(const mfe_ctx *c, const char *phase)
{
  if (!c || !phase) return -1;
  if (!strcmp(phase, "structure")) return c->t_structure;
  if (!strcmp(phase, "assembly")) return c->t_assembly;
  if (!strcmp(phase, "assembly_elem")) return c->t_assembly_elem;
  if (!strcmp(phase, "assembly_gather")) return c->t_assembly_gather;
  if (!strcmp(phase, "dirichlet0")) return c->t_dir0;
  if (!strcmp(phase, "dirichlet1")) return c->t_dir1;
  if (!strcmp(phase, "make_local")) return c->t_make_local;
  if (!strcmp(phase, "cg")) return c->t_cg;
  return -1;
}

// ---------------------------------------------------------------------------------------------------------
// timed repetitions of single kernels on the device-resident CG state (bench.py roofline, ncu captures)
// ---------------------------------------------------------------------------------------------------------
extern "C" int mfe_time_kernel(mfe_ctx *c, const char *what, int reps, double *ms_per_launch)
{
  if (!c || !what || reps < 1 || !ms_per_launch) return fail("mfe_time_kernel: bad arguments");
  if (!c->has_local) return fail("mfe_time_kernel: matrix needs local indices");
  CK(cudaSetDevice(c->device));
  if (ensure_vectors(c, c->nrows, c->num_cols)) return 1;
  const int64_t n = c->nrows;
  const ScalarArg one{ nullptr, 1.0 }, half{ nullptr, 0.5 };
  CgScalars h{}; h.alpha = 1e-3; h.neg_alpha = -1e-3; h.beta = 0.5; h.stop_iter = 1 << 30; h.tol = 0; h.one = 1; h.neg_one = -1;
  h.rtrans = 1; h.oldrtrans = 1; h.brkdown_tol = -1e300;
  CK(cudaMemcpyAsync(c->d_sc, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto body = [&](int rep) -> int {
    if (!strcmp(what, "matvec")) return launch_matvec_local(c, mv_args(c, c->d_p, c->d_Ap, false, nullptr, 0, false), MV_ALL);
    if (!strcmp(what, "matvec_dot")) return launch_matvec_local(c, mv_args(c, c->d_p, c->d_Ap, true, nullptr, 0, false), MV_ALL);
    if (!strcmp(what, "matvec_op")) return launch_matvec(c, c->d_p, c->d_Ap, true, nullptr, 0, false, false);   // incl. exchange / overlap mode (honours test_no_exchange)
    if (!strcmp(what, "matvec_simple")) { k_matvec_simple<<<grid_for(n, 128), 128, 0, c->stream>>>(c->d_coefs, c->d_cols, c->d_roff, c->d_tile_base, c->tr, c->d_p, c->d_Ap, n, c->ell_width); c->launches++; return 0; }
    if (!strcmp(what, "stream_probe")) {
      constexpr int PG = 16, PS = 1;
      const size_t smem = (size_t)PG * PS * MatvecCfg::SLOT_BYTES + (size_t)PG * PS * 8 + 64;
      auto kern = k_stream_probe<PG, PS>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return fail("stream_probe: smem attribute");
      if (c->tr != 32) return fail("stream_probe needs a 32-row tile configuration");
      kern<<<c->sm_count, 32 * PG, smem, c->stream>>>(c->d_coefs, c->d_cols, c->d_tile_base, c->ntiles, c->nnz_padded, c->d_tmp1);
      c->launches++; return 0;
    }
    if (!strcmp(what, "stream_probe21")) {
      constexpr int PG = 21, PS = 1;
      const size_t smem = (size_t)PG * PS * MatvecCfg::SLOT_BYTES + (size_t)PG * PS * 8 + 64;
      auto kern = k_stream_probe<PG, PS>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return fail("stream_probe: smem attribute");
      if (c->tr != 32) return fail("stream_probe needs a 32-row tile configuration");
      kern<<<c->sm_count, 32 * PG, smem, c->stream>>>(c->d_coefs, c->d_cols, c->d_tile_base, c->ntiles, c->nnz_padded, c->d_tmp1);
      c->launches++; return 0;
    }
    if (!strcmp(what, "waxpby")) { launch_waxpby(c, n, one, c->d_r, half, c->d_p, c->d_tmp1, nullptr, 0); return 0; }
    if (!strcmp(what, "dot_rr")) return launch_dot(c, n, c->d_r, c->d_r, 0, EPI_NONE, 0, false, nullptr);
    if (!strcmp(what, "dot_pap")) return launch_dot(c, n, c->d_Ap, c->d_p, 0, EPI_NONE, 0, false, nullptr);
    if (!strcmp(what, "xr_update")) { CK(launch_cg(c, k_xr_update<true>, dim3(grid_for(n, DOT_CHUNK)), dim3(DOT_CTA_THREADS), 0, n, c->d_sc, c->d_p, c->d_Ap, c->d_tmp1, c->d_tmp2, c->dot.piece_sums_a, c->dot.chunk_sums_a, nullptr, nullptr, nullptr, 0)); return 0; }
    if (!strcmp(what, "r_update")) { CK(launch_cg(c, k_xr_update<false>, dim3(grid_for(n, DOT_CHUNK)), dim3(DOT_CTA_THREADS), 0, n, c->d_sc, c->d_p, c->d_Ap, c->d_tmp1, c->d_tmp2, c->dot.piece_sums_a, c->dot.chunk_sums_a, nullptr, nullptr, nullptr, 0)); return 0; }
    (void)rep;
    return fail("mfe_time_kernel: unknown kernel '%s'", what);
  };
  if (body(-1)) return 1;                       // warm-up
  CK(cudaStreamSynchronize(c->stream));
  cudaEventRecord(a, c->stream);
  for (int i = 0; i < reps; ++i) if (body(i)) return 1;
  cudaEventRecord(b, c->stream);
  CK(cudaEventSynchronize(b));
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a); cudaEventDestroy(b);
  *ms_per_launch = ms / reps;
  CK(cudaGetLastError());
  return 0;
}
