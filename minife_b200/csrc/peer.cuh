// peer.cuh -- halo exchange and dot-product all-reduce over NVLink peer memory (one process per GPU).
//
// Replaces, on one NVSwitch box, the MPI traffic of ref/src/exchange_externals.hpp:44-161 (Irecv/Send/Wait per
// neighbour) and of ref/src/Vector_functions.hpp:218-222 (MPI_Allreduce of one double) by plain stores into the
// neighbour's memory:
//   * every rank owns a MAILBOX (one cudaMalloc, exported with cudaIpcGetMemHandle, opened by every other rank):
//       dot slots   [2][nranks] {value, tag}   -- tag = epoch of the dot that wrote the value
//       halo flags  [nranks]    epoch of the last halo that rank p delivered completely
//       halo data   [2][halo_cap] doubles      -- the external tail of the exchanged vector, double-buffered by epoch
//   * k_pack_peer gathers x[elements_to_send[i]] and stores it straight into the right place of each neighbour's
//     halo buffer; the last CTA to finish (ticket) publishes the epoch into the neighbours' flags;
//   * k_halo_wait_copy spins until all of MY neighbours have published the epoch, then copies the halo buffer into
//     the external tail of x.  Between the two the SpMV runs on the rows that need no external (overlap);
//   * peer_dot_allreduce: every rank stores its local dot into slot [epoch & 1][rank] of every rank, waits for the
//     nranks - 1 tags, and sums in rank order ((r0 + r1) + r2) ... -- the order of the reference through the MPI
//     shim -- so every rank gets the same bits without a collective call.
// Double buffering by epoch parity is enough: a rank can only be one epoch ahead of a neighbour, because finishing
// epoch e needs the neighbour's epoch-e data, which the neighbour sends after it finished reading epoch e - 1.
// All spins are bounded (wall-clock timeout): a lost peer ends in an error code, not in a hung GPU.
#pragma once
#include "kernels.cuh"

namespace mfe {

enum { PEER_MAX_RANKS = 16, PEER_MAX_NEIGHBORS = 26 };

struct DotSlot { double value; unsigned long long tag; };

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_f64(double *p, double v)
{
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p)
{
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *p == want; false on timeout
__device__ __forceinline__ bool peer_wait_eq(const unsigned long long *p, unsigned long long want, unsigned long long timeout_ns)
{
  if (ld_acquire_sys_u64(p) == want) return true;
  const unsigned long long t0 = global_ns();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i) if (ld_acquire_sys_u64(p) == want) return true;
    if (global_ns() - t0 > timeout_ns) return false;
    __nanosleep(64);
  }
}
// spin until *p >= want (monotone epoch counters: a neighbour may already have published a later epoch); false on timeout
__device__ __forceinline__ bool peer_wait_ge(const unsigned long long *p, unsigned long long want, unsigned long long timeout_ns)
{
  if (ld_acquire_sys_u64(p) >= want) return true;
  const unsigned long long t0 = global_ns();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i) if (ld_acquire_sys_u64(p) >= want) return true;
    if (global_ns() - t0 > timeout_ns) return false;
    __nanosleep(64);
  }
}

// ------------------------------------------------------------------------------------------------------
// halo
// ------------------------------------------------------------------------------------------------------
struct PeerHalo {                               // by-value kernel argument
  int n;                                        // neighbours
  int start[PEER_MAX_NEIGHBORS + 1];            // prefix of send_length: entries [start[i], start[i+1]) go to neighbour i
  double *dst[PEER_MAX_NEIGHBORS];              // neighbour i's halo buffer of THIS epoch + my offset inside its tail
  unsigned long long *flag[PEER_MAX_NEIGHBORS]; // neighbour i's halo flag for me
};

// ref/src/exchange_externals.hpp:116-137: gather + "send" (stores into the neighbours' memory)
__global__ void __launch_bounds__(256) k_pack_peer(const double *__restrict__ x, const int *__restrict__ elements_to_send, int nsend,
                                                   PeerHalo ph, unsigned long long epoch, unsigned int *ticket, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nsend) {
    int s = 0;
#pragma unroll 1
    while (s + 1 < ph.n && i >= ph.start[s + 1]) ++s;
    ph.dst[s][i - ph.start[s]] = x[elements_to_send[i]];
  }
  __threadfence_system();                       // my stores are ordered before the ticket below, system wide
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();                       // the other CTAs' stores (observed through the ticket) before the flags
  if (threadIdx.x < ph.n) st_release_sys_u64(ph.flag[threadIdx.x], epoch);
  if (threadIdx.x == 0) *ticket = 0u;
}

struct PeerWait {
  int n;
  const unsigned long long *flag[PEER_MAX_NEIGHBORS];   // MY flags of the neighbours I receive from
};

// ref/src/exchange_externals.hpp:90-101 + 147-153 (Irecv ... Wait): wait for the neighbours' epoch, then fill the
// external tail of x.  Few CTAs; every CTA checks all flags itself (no inter-CTA dependency).
__global__ void __launch_bounds__(256) k_halo_wait_copy(PeerWait pw, unsigned long long epoch, const double *halo, double *__restrict__ x_tail,
                                                        int n_ext, unsigned long long timeout_ns, int *err, CgScalars *sc, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if (threadIdx.x < pw.n) {
    // epochs are monotone and a neighbour may already be one epoch ahead (its data of that epoch went into the other
    // buffer): >=, not ==
    if (!peer_wait_ge(pw.flag[threadIdx.x], epoch, timeout_ns)) bad = 1;
  }
  __syncthreads();
  if (bad) {                                    // lost peer: close the CG gates at once (the host polls the status word)
    if (threadIdx.x == 0) { atomicExch(err, 1); if (sc) { sc->status = CG_COMM_ERROR; sc->stop_iter = -1; } }
    return;
  }
  __threadfence_system();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ext; i += gridDim.x * blockDim.x) x_tail[i] = __ldcg(halo + i);
}

// ------------------------------------------------------------------------------------------------------
// dot all-reduce
// ------------------------------------------------------------------------------------------------------
struct PeerDot {                                // by-value kernel argument; nranks == 0: single rank / NCCL path
  int nranks, rank;
  unsigned long long epoch;
  unsigned long long timeout_ns;
  DotSlot *box[PEER_MAX_RANKS];                 // dot slots of every rank's mailbox (box[rank] = mine)
  int *err;
};

// Called by ALL threads of a CTA (>= nranks threads); `local` must be valid on thread 0.  Returns the global sum
// (rank order) on thread 0; `ok` (thread 0) is false when a peer never answered.
__device__ __forceinline__ double peer_dot_allreduce(const PeerDot &pd, double local, bool &ok)
{
  __shared__ double sh_vals[PEER_MAX_RANKS];
  __shared__ int sh_bad;
  const int tid = threadIdx.x;
  if (tid == 0) { sh_vals[pd.rank] = local; sh_bad = 0; }
  __syncthreads();
  const int slot = (int)(pd.epoch & 1ULL) * pd.nranks;
  if (tid < pd.nranks && tid != pd.rank) {
    DotSlot *d = pd.box[tid] + slot + pd.rank;            // my slot in rank tid's mailbox
    st_relaxed_sys_f64(&d->value, sh_vals[pd.rank]);
    __threadfence_system();
    st_release_sys_u64(&d->tag, pd.epoch);
    const DotSlot *m = pd.box[pd.rank] + slot + tid;      // rank tid's slot in mine
    if (peer_wait_eq(&m->tag, pd.epoch, pd.timeout_ns)) sh_vals[tid] = ld_relaxed_sys_f64(&m->value);
    else { sh_bad = 1; sh_vals[tid] = 0.0; }
  }
  __syncthreads();
  double g = 0.0;
  ok = true;
  if (tid == 0) {
    g = sh_vals[0];
    for (int p = 1; p < pd.nranks; ++p) g = __dadd_rn(g, sh_vals[p]);
    if (sh_bad) { atomicExch(pd.err, 2); ok = false; }
  }
  return g;
}

// tiles (MV_TILE_ROWS consecutive rows) that reference an external column (col >= nrows): one warp per tile.
// Also writes the flagged copy of tile_base the interior SpMV pass uses (bit 62 = skip this tile).
__global__ void k_classify_tiles(const int *__restrict__ cols, const int64_t *__restrict__ tile_base, int64_t ntiles, int nrows,
                                 unsigned char *__restrict__ is_boundary, int64_t *__restrict__ tile_base_flagged)
{
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t > ntiles) return;
  if (t == ntiles) { if (lane == 0) tile_base_flagged[t] = tile_base[t]; return; }
  bool ext = false;
  for (int64_t i = tile_base[t] + lane; i < tile_base[t + 1]; i += 32) ext = ext || cols[i] >= nrows;
  const unsigned any = __ballot_sync(0xffffffffu, ext);
  if (lane == 0) { is_boundary[t] = any != 0; tile_base_flagged[t] = tile_base[t] | (any ? (1LL << 62) : 0LL); }
}

// ascending list of the flagged tiles (one CTA; ntiles is ~0.5 M at 256^3 per rank): every thread counts the flags of its
// contiguous run of tiles, a block-wide exclusive scan places the runs, every thread writes its indices
__global__ void __launch_bounds__(1024) k_compact_tiles(const unsigned char *__restrict__ is_boundary, int64_t ntiles, int *__restrict__ list, int *__restrict__ count)
{
  __shared__ int wtot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t run = (ntiles + blockDim.x - 1) / blockDim.x, t0 = (int64_t)tid * run, t1 = t0 + run < ntiles ? t0 + run : ntiles;
  int mine = 0;
  for (int64_t t = t0; t < t1; ++t) mine += is_boundary[t] != 0;
  int inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int v = wtot[lane], iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += u; }
    wtot[lane] = iv - v;
    if (lane == 31) *count = iv;
  }
  __syncthreads();
  int pos = wtot[warp] + inc - mine;
  for (int64_t t = t0; t < t1; ++t) if (is_boundary[t]) list[pos++] = (int)t;
}

}  // namespace mfe
