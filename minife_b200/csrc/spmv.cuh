// spmv.cuh -- CSR SpMV  y = A x  for sm_100a: ref/src/SparseMatrix_functions.hpp:546-578 (matvec_std, local part).
//
// CSR-stream design: the nnz range of 32 consecutive rows (a tile) is contiguous in packed_coefs / packed_cols, so
// it is staged into shared memory by two 1-D bulk async copies (TMA engine, SASS UBLKCP; full 128-byte HBM bursts
// whatever the row lengths; the matrix stream is marked L2 evict_first so that x stays cached), then lane r walks
// ITS row in storage order: sum = fl(sum + fl(coef*x[col])) -- the reference's exact sequence, no FMA.  x is
// gathered through L1/L2 (consecutive rows read consecutive x -> coalesced).
//
// One WARP per tile, G warps per CTA, S smem stages per warp, one persistent CTA per SM.  With S = 1 a warp arms
// its stage, waits, consumes, re-arms: the bytes in flight are the stages of the warps that are waiting, and
// 16 warps x 10.5 KB saturate HBM (measured 7.0 TB/s algorithmic at 256^3 = 98 % of the bare TMA stream,
// k_stream_probe).  The consumer is specialised: when every row of the tile holds exactly 27 entries (interior of
// the hex8 mesh, ~87 % of the tiles at 256^3) the row walk has no predication at all -- per entry LDS.64 coef,
// LDS col, IMAD.WIDE, LDG x, DMUL, DADD -- and any other staged tile uses predicated adds (the select stays off
// the dependent chain).  H gather batches per row trade one extra gather latency for fewer registers.
//
// Optionally accumulates plain sums of x[row]*y[row] per 512-row piece and per 4096-row chunk (binade prediction
// for the exact dot, exact_dot.h -- those values steer speed only, never results).
#pragma once
#include "kernels.cuh"
#include "peer.cuh"

namespace mfe {

enum { MV_TILE_ROWS = 32 };
#define MV_SKIP_BIT (1LL << 62)
struct MatvecCfg {
  static constexpr int CAP = MV_TILE_ROWS * 27 + 8;                // nnz capacity of one stage
  static constexpr int SLOT_BYTES = ((CAP * 12 + 127) / 128) * 128;
};

__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

// predicated add: sum += prod when u < len.  Kept as a predicated DADD (not a select) so that the dependent
// chain of a ragged row is one instruction per entry.
__device__ __forceinline__ void dadd_if(double &sum, double prod, int u, int len)
{
  asm volatile("{\n.reg .pred p;\nsetp.lt.s32 p, %2, %3;\n@p add.rn.f64 %0, %0, %1;\n}\n" : "+d"(sum) : "d"(prod), "r"(u), "r"(len));
}

// MODE: MV_ALL = every tile in natural order (single GPU, or no overlap); MV_SKIP = natural order, tiles whose
// tile_base carries MV_SKIP_BIT are skipped (interior pass); MV_LIST = work item i is tile tile_list[i] (boundary pass);
// MV_FUSED = the whole multi-rank SpMV in ONE launch: pack my boundary values into the neighbours' halo buffers
// (ref/src/exchange_externals.hpp:116-137), the MV_SKIP pass while those stores fly, then every warp waits for the
// neighbours' epoch flags itself and runs its share of the MV_LIST pass, gathering the externals straight from the halo
// buffer of this epoch (no copy into the tail of x, no second and third launch).
enum { MV_ALL = 0, MV_SKIP = 1, MV_LIST = 2, MV_FUSED = 3 };

struct MvArgs {                      // by-value kernel argument
  const double *coefs; const int *cols; const uint32_t *roff;
  const int64_t *tile_base;          // 64-bit nnz offset of every tile (ntiles + 1)
  const int64_t *tile_base_flagged;  // same with MV_SKIP_BIT on the tiles that reference externals (MV_SKIP / MV_FUSED)
  const double *x; double *y;
  int64_t nrows, ntiles, nnz_padded;
  int zmask;                         // always 0 at run time (see the gather batches)
  double *piece_sums, *chunk_sums;   // p.Ap prediction sums (null: none)
  const CgScalars *gate; int k;
  const int *tile_list; int64_t nlist;   // MV_LIST / MV_FUSED: the boundary tiles
  int ell_width;
  int use_ring;                      // 1: per-round piece sums reduced in shared memory (ring of 8 rounds), 0: one global atomic per tile
  int dbg_no_ext;                    // measurement hook (test_no_exchange only): 2 = leave after the interior pass
  int seed_y;                        // 1: the caller's y is meaningful (API call): beta*y[row] is formed for every row
  // MV_FUSED only
  const double *halo;                // externals of this epoch: column c >= nrows reads halo[c - nrows]
  const int *elements_to_send; int nsend;
  PeerHalo ph; PeerWait pw;
  unsigned long long epoch, timeout_ns;
  unsigned int *ticket; int *err; CgScalars *sc;
};

// x[col]; EXT: columns >= nrows are externals in the halo buffer, which the neighbours write while this kernel runs
// (L2 loads: L1 may still hold the buffer's lines from two epochs ago)
template <bool EXT>
__device__ __forceinline__ double mv_x(const MvArgs &a, int col)
{
  if (EXT && col >= (int)a.nrows) return __ldcg(a.halo + (col - (int)a.nrows));
  return __ldg(a.x + col);
}

struct MvSmem {
  unsigned char *gbase;    // this warp's stages
  uint64_t *bars;          // this warp's mbarriers
  double *racc; int *rcnt; volatile int *rtag;   // ring of round sums (G == 16)
};

// One pass over the work items of this CTA.  PASS is MV_ALL, MV_SKIP or MV_LIST; EXT: externals come from a.halo.
// Returns false when the gate was closed (nothing was computed).
template <int G, int S, int H, int PASS, bool EXT>
__device__ __forceinline__ bool mv_pass(const MvArgs &a, const MvSmem &sm, uint32_t &phase_bits, uint64_t pol_stream, bool check_gate)
{
  constexpr int TR = MV_TILE_ROWS;
  constexpr int CAP = MatvecCfg::CAP;
  constexpr int SLOT = MatvecCfg::SLOT_BYTES;
  constexpr int U = 27;                               // one pass covers a whole stencil row
  constexpr int B = (U + H - 1) / H;                  // entries per gather batch
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t *__restrict__ tile_base = PASS == MV_SKIP ? a.tile_base_flagged : a.tile_base;
  const int64_t nwork = PASS == MV_LIST ? a.nlist : a.ntiles;
  const int64_t nrows = a.nrows, ntiles = a.ntiles, nnz_padded = a.nnz_padded;
  const double *__restrict__ coefs = a.coefs; const int *__restrict__ cols = a.cols; const uint32_t *__restrict__ roff = a.roff;
  double *__restrict__ piece_sums = a.piece_sums, *__restrict__ chunk_sums = a.chunk_sums;
  // Round jj of this CTA covers the work items first .. first + G - 1, first = (blockIdx.x + jj * gridDim.x) * G; warp w
  // takes item first + w, which is tile first + w (MV_ALL), tile_list[first + w] (MV_LIST), or -- MV_SKIP -- tile
  // first + (w + jj) % G: the rotation spreads the skipped tiles over the warps (the boundary tiles of an x-split box
  // recur with period 8 in the tile index, which a fixed warp <-> residue assignment would turn into idle warps).
  auto item = [&](int64_t jj, int64_t &t) -> bool {
    const int64_t first = ((int64_t)blockIdx.x + jj * (int64_t)gridDim.x) * G;
    if (PASS == MV_LIST) {
      const int64_t i = first + w;
      if (i >= nwork) return false;
      t = (int64_t)a.tile_list[i];
      return true;
    }
    t = first + (PASS == MV_SKIP ? (int64_t)((w + (int)(jj % G)) % G) : (int64_t)w);
    return t < ntiles;
  };
  constexpr int64_t SKIP = PASS == MV_SKIP ? MV_SKIP_BIT : 0;

  auto issue = [&](int64_t t, int s) -> bool {
    const int64_t b0 = tile_base[t];
    if (b0 & SKIP) return false;                        // skipped in this pass: nothing to stage
    const int64_t a0 = b0 & ~3LL;
    int64_t a1 = ((tile_base[t + 1] & ~SKIP) + 3) & ~3LL;
    if (a1 > nnz_padded) a1 = nnz_padded;
    const int64_t n = a1 - a0;
    if (n <= 0 || n > CAP) return false;                // oversized tile: consumers read global memory directly
    double *sc = reinterpret_cast<double *>(sm.gbase + (size_t)s * SLOT);
    int *si = reinterpret_cast<int *>(sm.gbase + (size_t)s * SLOT + CAP * 8);
    mbar_expect_tx(&sm.bars[s], (uint32_t)(n * 12));
    bulk_g2s_hint(sc, coefs + a0, (uint32_t)(n * 8), &sm.bars[s], pol_stream);
    bulk_g2s_hint(si, cols + a0, (uint32_t)(n * 4), &sm.bars[s], pol_stream);
    return true;
  };

  unsigned armed0 = 0;
  if (lane == 0)
    for (int s = 0; s < S; ++s) { int64_t ti; if (item(s, ti) && issue(ti, s)) armed0 |= 1u << s; }
  if (check_gate && cg_gate_closed(a.gate, a.k)) {      // iteration switched off: let the copies in flight land, then leave
    if (lane == 0)
      for (int s = 0; s < S; ++s) if ((armed0 >> s) & 1u) mbar_wait(&sm.bars[s], (phase_bits >> s) & 1u);
    return false;
  }

  // software-pipelined row offsets: (o0,o1) of this lane's row in the NEXT tile are loaded one tile ahead
  uint32_t o0n = 0, o1n = 0; int64_t tb0n = 0, tb1n = 0;
  int64_t tnext = 0;
  bool vnext = item(0, tnext);
  if (vnext) {
    const int64_t r = tnext * TR + lane;
    if (r < nrows) { o0n = roff[r]; o1n = roff[r + 1]; }
    tb0n = tile_base[tnext]; tb1n = tile_base[tnext + 1];
  }
  // a warp without an item in round j has none in any later round (only the last round of the CTA can be partial,
  // and in MV_SKIP the warp may still own a tile of it -- item() decides)
  for (int64_t j = 0; ((int64_t)blockIdx.x + j * (int64_t)gridDim.x) * G < nwork; ++j) {
    const int s = (int)(j % S);
    const uint32_t parity = (phase_bits >> s) & 1u;
    const bool valid = vnext;
    const int64_t t = tnext;
    const int64_t r = t * TR + lane;
    const uint32_t o0 = o0n, o1 = o1n;
    const bool skipped = (tb0n & SKIP) != 0;
    const int64_t tb0 = tb0n & ~SKIP, tb1 = tb1n & ~SKIP;
    vnext = item(j + 1, tnext);
    if (vnext) {
      const int64_t rn = tnext * TR + lane;
      o0n = 0; o1n = 0;
      if (rn < nrows) { o0n = roff[rn]; o1n = roff[rn + 1]; }
      tb0n = tile_base[tnext]; tb1n = tile_base[tnext + 1];
    }
    if (!valid) continue;
    const int64_t a0 = tb0 & ~3LL;
    int64_t a1 = (tb1 + 3) & ~3LL;
    if (a1 > nnz_padded) a1 = nnz_padded;
    const int64_t ntile = a1 - a0;
    const bool staged = !skipped && ntile > 0 && ntile <= CAP;
    const bool live = !skipped && r < nrows;
    const int len = live ? (int)(o1 - o0) : 0;
    const bool full = staged && __all_sync(0xffffffffu, len == U);

    // The reference starts every row from beta*y[row] with beta = 0 (SparseMatrix_functions.hpp:565-568): +0.0 or -0.0 for
    // a finite y[row].  Here the entries are summed onto -0.0, the additive identity ((-0) + t = t for EVERY t, zeros of
    // either sign included), and the seed is added last: fl(0*y[row] + S) equals the reference's ((0*y + t1) + t2) + ...
    // bit for bit -- a nonzero S absorbs a zero seed, a zero S is -0.0 exactly when every term was -0.0, and a non-finite
    // y[row] makes both NaN.  So y[row] only has to be read when S is -0.0 (or always, for an API call: a.seed_y).
    double sum = -0.0;
    const double xr = (piece_sums && live) ? __ldg(a.x + r) : 0.0;
    if (staged) {
      mbar_wait(&sm.bars[s], parity);
      phase_bits ^= 1u << s;
      const double *sc = reinterpret_cast<const double *>(sm.gbase + (size_t)s * SLOT);
      const int *si = reinterpret_cast<const int *>(sm.gbase + (size_t)s * SLOT + CAP * 8);
      // first entry of this row inside the stage: 32-bit difference to the tile's first row + tile skew
      const int base = (int)(uint32_t)(o0 - (uint32_t)tb0) + (int)(tb0 - a0);
      if (full) {
        // All gathers of a batch must be in flight before its dependent add chain starts (the SM issues in order:
        // an add placed between the loads would serialise them).  Fold a bit of every gathered value into a token
        // that is masked to zero at run time (zmask == 0) and seed the chain with it: + (-0.0), the exact identity.
#pragma unroll
        for (int h = 0; h < H; ++h) {
          double cv[B], xv[B];
#pragma unroll
          for (int u = 0; u < B; ++u) if (h * B + u < U) { cv[u] = sc[base + h * B + u]; xv[u] = mv_x<EXT>(a, si[base + h * B + u]); }
          unsigned tok = 0;
#pragma unroll
          for (int u = 0; u < B; ++u) if (h * B + u < U) tok |= (unsigned)__double2hiint(xv[u]);
          sum = __dadd_rn(sum, __hiloint2double((int)((tok & (unsigned)a.zmask) | 0x80000000u), 0));
#pragma unroll
          for (int u = 0; u < B; ++u) if (h * B + u < U) sum = __dadd_rn(sum, __dmul_rn(cv[u], xv[u]));
        }
      } else if (live) {
        for (int i0 = 0; i0 < len; i0 += B) {
          double cv[B], xv[B];
#pragma unroll
          for (int u = 0; u < B; ++u) {
            const int ii = (i0 + u < len) ? base + i0 + u : base;
            cv[u] = sc[ii];
            xv[u] = mv_x<EXT>(a, si[ii]);
          }
          unsigned tok = 0;
#pragma unroll
          for (int u = 0; u < B; ++u) tok |= (unsigned)__double2hiint(xv[u]);
          sum = __dadd_rn(sum, __hiloint2double((int)((tok & (unsigned)a.zmask) | 0x80000000u), 0));
#pragma unroll
          for (int u = 0; u < B; ++u) dadd_if(sum, __dmul_rn(cv[u], xv[u]), i0 + u, len);
        }
        if (len < a.ell_width) {
          const double pad = __dmul_rn(0.0, __ldg(a.x));
          for (int u = len; u < a.ell_width; ++u) sum = __dadd_rn(sum, pad);
        }
      }
    } else if (live) {                // oversized tile: straight from global memory
      const int64_t b0 = tb0 + (int64_t)(uint32_t)(o0 - (uint32_t)tb0);
      for (int64_t i = 0; i < len; ++i) sum = __dadd_rn(sum, __dmul_rn(coefs[b0 + i], mv_x<EXT>(a, cols[b0 + i])));
      if (len < a.ell_width) {
        const double pad = __dmul_rn(0.0, __ldg(a.x));
        for (int u = len; u < a.ell_width; ++u) sum = __dadd_rn(sum, pad);
      }
    }
    if (live && (a.seed_y || __double_as_longlong(sum) == (long long)0x8000000000000000ULL)) sum = __dadd_rn(__dmul_rn(0.0, a.y[r]), sum);
    double prod = 0.0;
    if (live) { a.y[r] = sum; prod = xr * sum; }
    if (piece_sums) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) prod += __shfl_down_sync(0xffffffffu, prod, o);
    }
    __syncwarp();                     // every lane is done with stage s -> it can be refilled
    if (lane == 0) {
      int64_t t2;
      if (item(j + S, t2)) issue(t2, s);
      if (piece_sums) {               // rows [32 t, 32 t + 32) lie in piece t / 16 and in chunk t / 128
        if (G == 16 && PASS != MV_LIST && a.use_ring) {
          // the 16 tiles of one round of this CTA are exactly one piece: reduce them in shared memory (ring of 8
          // rounds; the warp that completes a round publishes it and hands the slot to round j + 8).  A warp that is
          // 8 rounds ahead of a straggler waits for the slot (the straggler never waits for a later round, so this
          // cannot deadlock): the sums must be exact per piece -- a sum that leaks into the wrong piece shifts the
          // predicted running sum by ~1e-6 relative, the dot then meets binade crossings where it expects none and
          // falls into its slow path (measured: 50-80 ms per p.Ap dot at 512^3).
          const int64_t jr = j;
          const int slot = (int)(jr & 7);
          const int64_t first = ((int64_t)blockIdx.x + j * (int64_t)gridDim.x) * G;   // first tile of the round
          const int expect = (int)(ntiles - first < G ? ntiles - first : G);
          while (sm.rtag[slot] != (int)jr) { }
          __threadfence_block();
          atomicAdd(&sm.racc[slot], prod);
          __threadfence_block();
          if (atomicAdd(&sm.rcnt[slot], 1) == expect - 1) {
            __threadfence_block();
            const double v = __longlong_as_double((long long)atomicExch(reinterpret_cast<unsigned long long *>(&sm.racc[slot]), 0ULL));
            sm.rcnt[slot] = 0;
            __threadfence_block();
            sm.rtag[slot] = (int)jr + 8;
            // MV_ALL owns the piece (plain store); the interior pass shares it with the boundary pass, which may run
            // earlier on another CTA of a fused launch (both add into sums that the previous k_xr_update zeroed)
            if (PASS == MV_ALL) piece_sums[first >> 4] = v; else atomicAdd(&piece_sums[first >> 4], v);
            atomicAdd(&chunk_sums[first >> 7], v);
          }
        } else {
          atomicAdd(&piece_sums[t >> 4], prod);
          atomicAdd(&chunk_sums[t >> 7], prod);
        }
      }
    }
  }
  return true;
}

template <int G, int S, int H, int MODE>
__global__ void __launch_bounds__(32 * G, 1)
k_matvec(const __grid_constant__ MvArgs a)
{
  // ell_width > 0: the matrix stands for the reference's ELLMatrix (ref/src/ELLMatrix.hpp), whose matvec
  // (SparseMatrix_functions.hpp:580-613) walks ell_width entries in every row: a row with fewer real entries ends with
  // padding entries (coef 0.0, col 0), i.e. with (ell_width - len) additions of 0.0 * x[0].  Reproduced literally (they
  // decide the sign of a zero result and propagate a non-finite x[0]), on the ragged path only.
  constexpr int SLOT = MatvecCfg::SLOT_BYTES;
  extern __shared__ __align__(128) unsigned char smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  MvSmem sm;
  sm.gbase = smem + (size_t)w * S * SLOT;
  sm.bars = reinterpret_cast<uint64_t *>(smem + (size_t)G * S * SLOT) + w * S;
  sm.racc = reinterpret_cast<double *>(smem + (size_t)G * S * SLOT + (size_t)G * S * 8);     // [8] round sums (G == 16)
  sm.rcnt = reinterpret_cast<int *>(sm.racc + 8);                                             // [8] round arrivals
  sm.rtag = sm.rcnt + 8;                                                                      // [8] round each slot collects
  int *flag_last = const_cast<int *>(sm.rtag) + 8;
  uint64_t pol_stream = 0;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&sm.bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 8) { sm.racc[threadIdx.x] = 0.0; sm.rcnt[threadIdx.x] = 0; sm.rtag[threadIdx.x] = (int)threadIdx.x; }
  __syncthreads();
  uint32_t phase_bits = 0;            // expected mbarrier parity per stage (flips each time a stage completes)

  if (MODE != MV_FUSED) {
    mv_pass<G, S, H, MODE, false>(a, sm, phase_bits, pol_stream, true);
    return;
  }
  // ---- MV_FUSED ----
  if (cg_gate_closed(a.gate, a.k)) return;             // uniform over the grid: nobody packs, nobody waits
  // (1) pack: ref/src/exchange_externals.hpp:116-137 as stores into the neighbours' halo buffers of this epoch; the last
  //     CTA to finish (ticket) publishes the epoch into the neighbours' flags
  if (a.ph.n > 0) {
    const int nthr = (int)(gridDim.x * blockDim.x);
    for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < a.nsend; i += nthr) {
      int s = 0;
#pragma unroll 1
      while (s + 1 < a.ph.n && i >= a.ph.start[s + 1]) ++s;
      a.ph.dst[s][i - a.ph.start[s]] = a.x[a.elements_to_send[i]];
    }
    __threadfence_system();                             // my stores are ordered before the ticket below, system wide
    __syncthreads();
    if (threadIdx.x == 0) *flag_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (*flag_last) {
      __threadfence_system();                           // the other CTAs' stores (observed through the ticket) before the flags
      if ((int)threadIdx.x < a.ph.n) st_release_sys_u64(a.ph.flag[threadIdx.x], a.epoch);
      if (threadIdx.x == 0) *a.ticket = 0u;
    }
  }
  // (2) the tiles that need no external
  mv_pass<G, S, H, MV_SKIP, false>(a, sm, phase_bits, pol_stream, false);
  if (a.nlist <= 0 || a.dbg_no_ext == 2) return;
  // (3) Irecv ... Wait (exchange_externals.hpp:90-101, 147-153): every warp checks the neighbours' flags itself
  if (a.pw.n > 0) {
    bool ok = true;
    if (lane < a.pw.n) ok = peer_wait_ge(a.pw.flag[lane], a.epoch, a.timeout_ns);
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    if (!ok) {
      if (lane == 0) { atomicExch(a.err, 1); if (a.sc) { a.sc->status = CG_COMM_ERROR; a.sc->stop_iter = -1; } }
      return;
    }
  }
  // (4) the boundary tiles (their prediction sums go through atomics, not through the ring of round sums)
  mv_pass<G, S, H, MV_LIST, true>(a, sm, phase_bits, pol_stream, false);
}

// Streaming probe (measurement only): the staging half of k_matvec with the consumer removed -- every warp waits
// for its tile and re-arms the stage at once.  Its GB/s is the ceiling of the TMA stream at this occupancy.
template <int G, int S>
__global__ void __launch_bounds__(32 * G, 1)
k_stream_probe(const double *__restrict__ coefs, const int *__restrict__ cols, const int64_t *__restrict__ tile_base, int64_t ntiles,
               int64_t nnz_padded, double *__restrict__ sink)
{
  constexpr int CAP = MatvecCfg::CAP;
  constexpr int SLOT = MatvecCfg::SLOT_BYTES;
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

// plain thread-per-row CSR SpMV in the reference order (validation twin of k_matvec)
__global__ void k_matvec_simple(const double *__restrict__ coefs, const int *__restrict__ cols, const uint32_t *__restrict__ roff,
                                const int64_t *__restrict__ tile_base, int tr, const double *__restrict__ x, double *__restrict__ y,
                                int64_t nrows, int ell_width = 0)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int64_t t = r / tr;
  const int64_t off = tile_base[t] + (int64_t)(uint32_t)(roff[r] - roff[t * tr]);
  const int len = (int)(uint32_t)(roff[r + 1] - roff[r]);
  double sum = 0.0;
  for (int i = 0; i < len; ++i) sum = __dadd_rn(sum, __dmul_rn(coefs[off + i], __ldg(x + cols[off + i])));
  for (int i = len; i < ell_width; ++i) sum = __dadd_rn(sum, __dmul_rn(0.0, __ldg(x)));
  y[r] = sum;
}

}  // namespace mfe
