// dot_kernels.cuh -- device side of the order-reproducing dot product (exact_dot.h) and of the fused CG vector
// updates that feed it.  ref/src/Vector_functions.hpp:194-226 (dot), ref/src/cg_solve.hpp:157-212.
//
// Pipeline of one exact dot over n terms (piece = 512 terms = one warp, chunk = 8 pieces):
//   producer   plain sums of the products per piece and per chunk -- written by the kernel that produced the
//              operands (k_xr_update for r.r, the SpMV for p.Ap) or by k_chunk_sums; they only PREDICT binades
//   k_dot_maps persistent TMA-staged pass over the operands: one DotMap per leaf, collapsed per piece, chunk,
//              super and hyper (arrival tickets); leaves where the predicted binade changes are listed as dirty
//   k_dot_walk one CTA: clean gaps between dirty leaves collapse to one map each (one warp per gap), thread 0
//              applies  [gap][dirty leaf][gap] ...  in order and runs the CG scalar epilogue
#pragma once
#include "kernels.cuh"
#include "peer.cuh"

namespace mfe {

// ------------------------------------------------------------------------------------------------------
// producers of the prediction sums
// ------------------------------------------------------------------------------------------------------
// cg_solve.hpp:211-212 fused:  x = 1*x + alpha*p ;  r = 1*r + (-alpha)*Ap ;  plain sums of r*r per piece / chunk
// (binade prediction for the next exact dot).  One CTA per chunk, one warp per piece (a warp instruction touches
// 256 contiguous bytes per operand).  Also zeroes the p.Ap prediction sums that the next SpMV accumulates into.
// WITH_X = false: the r half only; x += alpha*p then runs as a k_waxpby on the low-priority stream (api.cu, x_overlap).
template <bool WITH_X>
__global__ void __launch_bounds__(DOT_CTA_THREADS) k_xr_update(int64_t n, const CgScalars *sc, const double *__restrict__ p,
                                                               const double *__restrict__ Ap, double *__restrict__ x,
                                                               double *__restrict__ r, double *__restrict__ piece_sums,
                                                               double *__restrict__ chunk_sums, double *__restrict__ zero_piece_sums,
                                                               double *__restrict__ zero_chunk_sums, const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  __shared__ double sh[DOT_PIECES_PER_CHUNK];
  const double alpha = sc->alpha, nalpha = sc->neg_alpha, one = 1.0;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int64_t piece = (int64_t)blockIdx.x * DOT_PIECES_PER_CHUNK + wq;
  const int64_t base = piece * DOT_PIECE;
  double acc = 0.0;
#pragma unroll 8
  for (int it = 0; it < DOT_LEAF; ++it) {
    const int64_t i = base + it * 32 + lane;
    if (i < n) {
      if (WITH_X) x[i] = __dadd_rn(__dmul_rn(one, x[i]), __dmul_rn(alpha, p[i]));
      const double rn = __dadd_rn(__dmul_rn(one, r[i]), __dmul_rn(nalpha, Ap[i]));
      r[i] = rn;
      acc += rn * rn;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    sh[wq] = acc;
    if (base < n) { piece_sums[piece] = acc; if (zero_piece_sums) zero_piece_sums[piece] = 0.0; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < DOT_PIECES_PER_CHUNK; ++q) t += sh[q];
    chunk_sums[blockIdx.x] = t;
    if (zero_chunk_sums) zero_chunk_sums[blockIdx.x] = 0.0;
  }
}

// plain per-piece / per-chunk sums of x*y (when no producer kernel supplied them)
__global__ void __launch_bounds__(DOT_CTA_THREADS) k_chunk_sums(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                                                double *__restrict__ piece_sums, double *__restrict__ chunk_sums,
                                                                const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  __shared__ double sh[DOT_PIECES_PER_CHUNK];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int64_t piece = (int64_t)blockIdx.x * DOT_PIECES_PER_CHUNK + wq;
  const int64_t base = piece * DOT_PIECE;
  double acc = 0.0;
#pragma unroll 8
  for (int it = 0; it < DOT_LEAF; ++it) {
    const int64_t i = base + it * 32 + lane;
    if (i < n) acc += x[i] * y[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) { sh[wq] = acc; if (base < n) piece_sums[piece] = acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < DOT_PIECES_PER_CHUNK; ++q) t += sh[q];
    chunk_sums[blockIdx.x] = t;
  }
}

// Exclusive prefix over chunk sums; chunk_prefix[nchunks] = total (the MFE_DOT_TREE result).  One CTA.
__global__ void __launch_bounds__(1024) k_chunk_prefix(const double *__restrict__ sums, int nchunks, double *__restrict__ chunk_prefix,
                                                       const CgScalars *gate, int k)
{
  if (cg_gate_closed(gate, k)) return;
  __shared__ double wtot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // block scan: each thread owns a contiguous run of chunks
  const int runlen = (nchunks + blockDim.x - 1) / blockDim.x;
  const int c0 = threadIdx.x * runlen, c1 = min(nchunks, c0 + runlen);
  double run = 0.0;
  for (int c = c0; c < c1; ++c) run += sums[c];
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
  for (int c = c0; c < c1; ++c) { const double v = sums[c]; chunk_prefix[c] = pre; pre += v; }
  if (c1 == nchunks && c0 < nchunks) chunk_prefix[nchunks] = pre;
  if (nchunks == 0 && threadIdx.x == 0) chunk_prefix[0] = 0.0;
}

__device__ __forceinline__ double lds_f64(uint32_t addr)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
template <int OFF>
__device__ __forceinline__ double lds_f64_off(uint32_t addr)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
  return v;
}

// ------------------------------------------------------------------------------------------------------
// ordered map compositions inside a warp
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ DotMap shfl_up_map(const DotMap &m, int o)
{
  DotMap r;
  r.a0 = __shfl_up_sync(0xffffffffu, m.a0, o); r.a1 = __shfl_up_sync(0xffffffffu, m.a1, o);
  r.lo = __shfl_up_sync(0xffffffffu, m.lo, o); r.hi = __shfl_up_sync(0xffffffffu, m.hi, o);
  r.hyp = __shfl_up_sync(0xffffffffu, m.hyp, o); r.pad = 0;
  return r;
}
// ordered warp reduction: lane 31 ends with the composition of lanes 32-WIDTH..31 in lane order
template <int WIDTH = 32>
__device__ __forceinline__ DotMap warp_compose(DotMap m, int lane)
{
#pragma unroll
  for (int o = 1; o < WIDTH; o <<= 1) {
    const DotMap u = shfl_up_map(m, o);
    if (lane >= o) m = dot_compose(u, m);
  }
  return m;
}
__device__ __forceinline__ DotMap load_map_cg(const DotMap *src)
{
  DotMap m;
  m.a0 = __ldcg(&src->a0); m.a1 = __ldcg(&src->a1); m.lo = __ldcg(&src->lo); m.hi = __ldcg(&src->hi); m.hyp = __ldcg(&src->hyp); m.pad = __ldcg(&src->pad);
  return m;
}

// level 2 -> 3 -> 4 of the map hierarchy, called by a whole warp with the map of chunk c on lane 31: the last chunk
// of a group of 32 to arrive collapses the group (arrival tickets reset themselves, so the workspace is reusable
// without a memset).
__device__ __forceinline__ void dot_publish_chunk_map(const DotWork &w, const DotMap &cmap, int lane, int c, int nchunks)
{
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
  if (lane < ssize) { sm = load_map_cg(w.cta_maps + ((int64_t)sgrp << 5) + lane); sm.pad = 0; }
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
  if (lane < hsize) { hm = load_map_cg(w.super_maps + ((int64_t)hgrp << 5) + lane); hm.pad = 0; }
  hm = warp_compose(hm, lane);
  if (lane == 31) { w.hyper_maps[hgrp] = hm; w.hyper_cnt[hgrp] = 0; }
}

// ------------------------------------------------------------------------------------------------------
// Exact dot, walk phase: one CTA.  (1) sort the dirty leaves, fetch their products; (2) one warp per gap: the
// <= 9 runs of the gap are loaded at once (one map per lane and run) and collapsed by ordered shuffle scans into
// ONE map per gap; (3) thread 0 walks  [gap 0][dirty leaf 0][gap 1][dirty leaf 1] ... [gap D]  -- ~2 short steps
// per dirty leaf.  Any map that fails its check is replayed from global memory in plain floating point, so the
// result is the serial loop's for every input.
// (Folding this walk into the tail of k_dot_maps -- last CTA by arrival ticket -- was built and measured in round 2: the walk
// then runs under the maps kernel's 80-register cap and takes 45-52 us instead of 21-26; it stays a launch of its own.)
// ------------------------------------------------------------------------------------------------------
enum { WALK_THREADS = 512 };

struct WalkArgs {                 // by-value kernel argument: what happens with the local sum
  CgScalars *sc; int epi, k;
  double *normr_hist, *rtrans_hist, *pap_hist, *alpha_hist;
  PeerDot pd;
};
struct WalkSmem {
  int dirty[DOT_MAX_DIRTY];
  int sorted[DOT_MAX_DIRTY];
  DotMap gmaps[DOT_MAX_DIRTY + 1];
  double replay_prod[DOT_MAX_DIRTY * DOT_LEAF];
};

// called by all NT threads of one CTA
template <int NT>
__device__ __noinline__ void dot_walk_body(int64_t n, const double *__restrict__ x, const double *__restrict__ y, const DotWork &w,
                                          const WalkArgs &wa, WalkSmem &ws)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long t_begin = (w.stats && tid == 0) ? global_ns() : 0ULL;
  int slow_gaps = 0;
  const int D = __ldcg(w.dirty_count);
  const long long nleaves = ((n + DOT_PIECE - 1) / DOT_PIECE) * 32;     // whole pieces: the maps that exist
  double s = 0.0;
  if (D > DOT_MAX_DIRTY) {                     // pathological input: the literal loop
    if (tid == 0) s = dot_replay(x, y, 0, n, 0.0);
  } else {
    for (int i = tid; i < D; i += NT) ws.dirty[i] = __ldcg(w.dirty_list + i);
    __syncthreads();
    for (int i = tid; i < D; i += NT) {        // rank sort (leaf indices are distinct)
      const int v = ws.dirty[i]; int rk = 0;
      for (int j = 0; j < D; ++j) rk += ws.dirty[j] < v;
      ws.sorted[rk] = v;
    }
    __syncthreads();
    // products of the dirty leaves, fetched by all threads at once (the sequential walk then reads smem only)
    for (int q = tid; q < D * DOT_LEAF; q += NT) {
      const int64_t i = (int64_t)ws.sorted[q / DOT_LEAF] * DOT_LEAF + (q % DOT_LEAF);
      ws.replay_prod[q] = i < n ? __dmul_rn(x[i], y[i]) : 0.0;
    }
    const unsigned long long t_p1 = (w.stats && tid == 0) ? global_ns() : 0ULL;
    for (int g = warp; g <= D; g += NT / 32) {
      const long long lo = g == 0 ? 0 : (long long)ws.sorted[g - 1] + 1, hi = g == D ? nleaves : (long long)ws.sorted[g];
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
      if (lane == 31) ws.gmaps[g] = gm;
    }
    __syncthreads();
    if (w.stats && tid == 0) { const unsigned long long t = global_ns(); w.stats[11] += t_p1 - t_begin; w.stats[12] += t - t_p1; w.stats[13] = t; }
    if (tid == 0) {
      // The sequential part: [gap 0][dirty leaf 0][gap 1] ... [gap D].  Software-pipelined by hand: the map of the next gap
      // and the 16 products of the next dirty leaf are fetched from shared memory BEFORE the dependent chain of the current
      // step (apply the gap map, then 16 ordered adds) starts, so one step costs little more than that chain.
      DotMap mnext = ws.gmaps[0];
      double pnext[DOT_LEAF];
#pragma unroll
      for (int i = 0; i < DOT_LEAF; ++i) pnext[i] = D > 0 ? ws.replay_prod[i] : 0.0;
      for (int g = 0; g <= D; ++g) {
        const DotMap m = mnext;
        double pr[DOT_LEAF];
#pragma unroll
        for (int i = 0; i < DOT_LEAF; ++i) pr[i] = pnext[i];
        const long long lo = g == 0 ? 0 : (long long)ws.sorted[g - 1] + 1, hi = g == D ? nleaves : (long long)ws.sorted[g];
        const int64_t b = g < D ? (int64_t)ws.sorted[g] * DOT_LEAF : 0;
        if (g < D) {
          mnext = ws.gmaps[g + 1];
          if (g + 1 < D) {
#pragma unroll
            for (int i = 0; i < DOT_LEAF; ++i) pnext[i] = ws.replay_prod[(g + 1) * DOT_LEAF + i];
          }
        }
        if (lo < hi && !dot_apply(m, s)) { s = dot_walk_gap_slow(lo, hi, w, x, y, n, s); ++slow_gaps; }
        if (g < D) {
          if (b + DOT_LEAF <= n) {
#pragma unroll
            for (int i = 0; i < DOT_LEAF; ++i) s = __dadd_rn(s, pr[i]);
          } else {
            const int cnt = (int)(n - b);
#pragma unroll
            for (int i = 0; i < DOT_LEAF; ++i) if (i < cnt) s = __dadd_rn(s, pr[i]);
          }
        }
      }
    }
  }
  CgScalars *sc = wa.sc;
  if (tid == 0) {
    *w.dirty_count = 0;                         // ready for the next dot on this workspace
    if (sc) sc->dot_local = s;
    if (w.stats) {
      // diagnostics: [0] dots, [1] sum of dirty leaves, [2] max, [3] whole-vector replays, [4] gaps that took the slow path,
      // [5] total ns in the walk, [6] max ns, [7] dot index of the max; then a ring of (dirty leaves, ns) per dot
      const unsigned long long t_now = global_ns(), dt = t_now - t_begin, idx = w.stats[0];
      if (D <= DOT_MAX_DIRTY) w.stats[14] += t_now - w.stats[13];      // [11] prologue (list, sort, products) [12] gap maps [14] sequential part, ns
      w.stats[0] = idx + 1; w.stats[1] += (unsigned long long)D; if ((unsigned long long)D > w.stats[2]) w.stats[2] = D;
      if (D > DOT_MAX_DIRTY) w.stats[3] += 1;
      w.stats[4] += (unsigned long long)slow_gaps; w.stats[5] += dt;
      if (dt > w.stats[6]) { w.stats[6] = dt; w.stats[7] = idx; }
      w.stats[DOT_STATS_HEAD + 2 * (idx % DOT_STATS_TRACE)] = ((unsigned long long)slow_gaps << 32) | (unsigned)D;
      w.stats[DOT_STATS_HEAD + 2 * (idx % DOT_STATS_TRACE) + 1] = dt;
    }
  }
  bool ok = true;
  const unsigned long long t_ar = (w.stats && tid == 0 && wa.pd.nranks > 1) ? global_ns() : 0ULL;
  if (wa.pd.nranks > 1) s = peer_dot_allreduce(wa.pd, s, ok);     // rank-ordered global sum over NVLink peer memory
  if (tid == 0) {
    if (w.stats && wa.pd.nranks > 1) {          // [9] total / [10] max ns spent in the all-reduce (mostly: waiting for the slowest rank)
      const unsigned long long dt = global_ns() - t_ar;
      w.stats[9] += dt; if (dt > w.stats[10]) w.stats[10] = dt;
    }
    *w.result = s;
    if (sc) {
      if (!ok) { sc->status = CG_COMM_ERROR; sc->stop_iter = -1; }
      else if (wa.epi != EPI_NONE) cg_epilogue(sc, wa.epi, wa.k, s, wa.normr_hist, wa.rtrans_hist, wa.pap_hist, wa.alpha_hist);
    }
  }
}

__global__ void __launch_bounds__(WALK_THREADS) k_dot_walk(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                                          const __grid_constant__ DotWork w, const CgScalars *gate, const __grid_constant__ WalkArgs wa)
{
  if (cg_gate_closed(gate, wa.k)) return;
  __shared__ WalkSmem ws;
  dot_walk_body<WALK_THREADS>(n, x, y, w, wa, ws);
}

// ------------------------------------------------------------------------------------------------------
// Exact dot, map phase.  Persistent CTAs of NW independent warps; CTA b owns the contiguous pieces
// [b*PP, (b+1)*PP).  Every warp arms its own shared-memory stage with 1-D bulk async copies (TMA engine) of the
// 4 KB per operand of its next piece, waits on its mbarrier, and turns the piece into one DotMap:
//   * lane l owns leaf l (terms 16 l .. 16 l + 15) and reads it in the rotated order i ^ (l & 15): bank-conflict
//     free, and legal because the common "simple" leaf formulation is order independent (exact_dot.h);
//   * the predicted running sum at the piece start comes from a per-CTA table (prefix of the producer's piece
//     sums), the prediction inside the piece from a warp scan of the leaf sums;
//   * simple leaves compose in floating point (a, lo, hi as exact integers in doubles); a piece with any other
//     leaf takes the general int64 composition;
//   * the last piece of a chunk to arrive (ticket) collapses the chunk and feeds the super / hyper tickets.
// The bytes in flight are the stages of the waiting warps, as in k_matvec.
// ------------------------------------------------------------------------------------------------------
enum { DOT_MAPS_MAX_PP = 2048 };     // pieces per CTA (bounds the prefix table in shared memory)

template <bool SAME, int NW>
struct DotMapsCfg {
  static constexpr int STAGE_BYTES = (SAME ? 1 : 2) * DOT_PIECE * 8;
  static constexpr int MAX_CHUNKS = DOT_MAPS_MAX_PP / DOT_PIECES_PER_CHUNK + 2;      // chunks a CTA's range can touch
  static constexpr size_t MAIN_BYTES = (size_t)NW * STAGE_BYTES + (size_t)NW * 8 /*mbarriers*/ + (size_t)(DOT_MAPS_MAX_PP + 1) * 8 /*table*/ +
                                       (size_t)2 * NW * 8 /*reductions*/ + (size_t)MAX_CHUNKS * 4 /*arrival counters*/ + 64;
  static constexpr size_t SMEM_BYTES = MAIN_BYTES;
};

// The slow path of one piece: a leaf-by-leaf treatment with the prediction refined inside the piece (binade changes,
// exact ties, leaves near a binade end, ragged tail, zero / non-finite running sums).  Returns the piece map on lane 31
// and stores the leaf maps + dirty leaves when there are any.  Called by a whole warp.
// (forced inline: as a real call its DotMap result and the by-reference flag live in local memory, and the hot loop then
// stalls on local loads -- measured: long-scoreboard stalls 4x, kernel +25 %)
template <bool SAME>
__device__ __forceinline__ DotMap dot_piece_slow(int64_t n, const double *__restrict__ x, const double *__restrict__ y, const DotWork &w, int piece,
                                              const double *stage, double pre_begin, double pre_end, bool &piece_dirty_out)
{
  const int lane = threadIdx.x & 31;
  const int64_t lbeg = (int64_t)piece * DOT_PIECE + lane * DOT_LEAF;
  const int cnt = (int)(lbeg >= n ? 0 : (n - lbeg < DOT_LEAF ? n - lbeg : DOT_LEAF));
  const double *xs = stage + lane * DOT_LEAF, *ys = SAME ? xs : xs + DOT_PIECE;
  if (w.stats && lane == 0) atomicAdd(&w.stats[8], 1ULL);          // diagnostics: pieces that left the fast path
  // products rounded once (fl(x*y)) exactly as the serial loop forms them
  double prod[DOT_LEAF];
#pragma unroll
  for (int i = 0; i < DOT_LEAF; ++i) prod[i] = i < cnt ? __dmul_rn(xs[i], ys[i]) : 0.0;
  double ls = 0.0;
#pragma unroll
  for (int i = 0; i < DOT_LEAF; ++i) ls += prod[i];
  double inc = ls;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
  const double pre = pre_begin + (inc - ls);
  const int hyp = dot_hyp_of(pre);
  int hyp_end = __shfl_down_sync(0xffffffffu, hyp, 1);       // predicted binade at the start of the next leaf
  if (lane == 31) hyp_end = dot_hyp_of(pre_end);

  bool risky = false;
  DotMap m = dot_identity();
  if (cnt == DOT_LEAF) {
    m = dot_leaf_map16(prod, hyp, pre);                      // simple form when far from the binade ends, else ordered with exact bounds
    if (hyp >= 0 && hyp != HYP_ZERO) risky = dot_leaf_risky(m, hyp, pre);
  } else if (cnt > 0) {
    m = dot_leaf_map(x, y, lbeg, lbeg + cnt, hyp);           // ragged last leaf of the vector
  }
  const bool dirty = cnt > 0 && (m.hyp == HYP_INVALID || hyp != hyp_end || risky);
  const bool piece_dirty = __any_sync(0xffffffffu, dirty);
  const DotMap wm = warp_compose(m, lane);
  if (piece_dirty) {
    // keep the 32 leaf maps so that the walker can stitch around the dirty leaves of this piece
    w.leaf_maps[(int64_t)piece * 32 + lane] = m;
    if (dirty) {
      const int slot = atomicAdd(w.dirty_count, 1);
      if (slot < DOT_MAX_DIRTY) w.dirty_list[slot] = piece * 32 + lane;
    }
  }
  piece_dirty_out = piece_dirty;
  return wm;
}

template <bool SAME, int NW>
__global__ void __launch_bounds__(32 * NW, 1)
k_dot_maps(int64_t n, const double *__restrict__ x, const double *__restrict__ y, const __grid_constant__ DotWork w, int npieces, int PP,
           const CgScalars *gate, int k)
{
  constexpr int STAGE = DotMapsCfg<SAME, NW>::STAGE_BYTES;
  constexpr int NT = 32 * NW;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  double *stage = reinterpret_cast<double *>(smem + (size_t)wq * STAGE);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t)NW * STAGE) + wq;
  double *table = reinterpret_cast<double *>(smem + (size_t)NW * STAGE + (size_t)NW * 8);
  double *red = table + (DOT_MAPS_MAX_PP + 1);                  // [2][NW]
  int *ccnt = reinterpret_cast<int *>(red + 2 * NW);            // arrivals per chunk of this CTA's range
  const int nchunks = (npieces + DOT_PIECES_PER_CHUNK - 1) / DOT_PIECES_PER_CHUNK;
  const int p0 = blockIdx.x * PP;
  const int np = min(PP, npieces - p0);                         // pieces of this CTA (the host sizes the grid so that np >= 1)
  if (cg_gate_closed(gate, k)) return;

  auto issue = [&](int piece) {
    const int64_t b = (int64_t)piece * DOT_PIECE;
    const int64_t cntp = n - b < DOT_PIECE ? n - b : DOT_PIECE;
    const uint32_t bytes = (uint32_t)(((cntp + 1) & ~1LL) * 8);  // multiple of 16; the vectors carry >= 8 doubles of slack
    mbar_expect_tx(bar, SAME ? bytes : 2 * bytes);
    bulk_g2s(stage, x + b, bytes, bar);
    if (!SAME) bulk_g2s(stage + DOT_PIECE, y + b, bytes, bar);
  };
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < DotMapsCfg<SAME, NW>::MAX_CHUNKS; i += NT) ccnt[i] = 0;
  __syncthreads();
  if (lane == 0 && wq < np) issue(p0 + wq);                     // first piece of every warp is in flight during the prologue

  // ---- prologue: predicted running sum at the start of every piece of this CTA ----
  {
    const int c0 = p0 / DOT_PIECES_PER_CHUNK;
    double part = 0.0;
    for (int i = tid; i < c0; i += NT) part += w.chunk_sums[i];
    if (tid < p0 - c0 * DOT_PIECES_PER_CHUNK) part += w.piece_sums[c0 * DOT_PIECES_PER_CHUNK + tid];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    // runs of R consecutive pieces per thread
    const int R = (np + NT - 1) / NT, j0 = tid * R;
    double run = 0.0;
    for (int q = 0; q < R; ++q) if (j0 + q < np) run += w.piece_sums[p0 + j0 + q];
    double inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) red[NW + wq] = inc;
    if (lane == 0) red[wq] = part;
    __syncthreads();
    double B0 = 0.0, woff = 0.0;
#pragma unroll
    for (int q = 0; q < NW; ++q) { B0 += red[q]; if (q < wq) woff += red[NW + q]; }
    double acc = woff + (inc - run);
    for (int q = 0; q < R; ++q) if (j0 + q < np) { table[j0 + q] = B0 + acc; acc += w.piece_sums[p0 + j0 + q]; }
    if (tid == NT - 1) table[np] = B0 + acc;
    __syncthreads();
  }

  uint32_t parity = 0;
  for (int j = wq; j < np; j += NW) {
    const int piece = p0 + j;
    const double pre0 = table[j], pre1 = table[j + 1];
    const int hyp0 = dot_hyp_of(pre0);
    const int kk = 1075 - (hyp0 & 0x7ff);                       // the terms scale by 2^kk into units of the running sum's ulp
    // Fast path: the predicted running sum keeps one binade over the whole piece.  Every leaf is then taken in the
    // order-independent "simple" form (exact_dot.h, dot_leaf16_simple) under that one hypothesis, and the piece map gets
    // CONSERVATIVE bounds: every partial offset of the piece lies in [sum of the negative offsets, sum of the positive
    // ones].  Taken only when those bounds, applied to the predicted significand, stay MFE_DOT_FAR_MARGIN ulps inside
    // the binade (so dot_apply will accept the map); anything else -- a binade change, an exact tie, a huge term, the
    // ragged tail -- goes leaf by leaf through dot_piece_slow.
    const bool whole = (int64_t)(piece + 1) * DOT_PIECE <= n;
    bool fast = whole && hyp0 >= 0 && hyp0 != HYP_ZERO && hyp0 == dot_hyp_of(pre1) && kk <= 1023;
    mbar_wait(bar, parity);
    parity ^= 1u;
    DotMap wm;
    bool piece_dirty = false;
    if (whole && hyp0 == HYP_ZERO && dot_hyp_of(pre1) == HYP_ZERO) {
      // The running sum is predicted to be exactly +0.0 before and after the piece (a Dirichlet face: the z = 0 plane of a
      // rank is ~130 consecutive pieces of zeros, all in the range of ONE CTA -- through the leaf-by-leaf path they made that
      // CTA the straggler of the kernel: +65 us per dot on the ranks at x = 0, z = 0 of a 2x2x2 run).  If every product is
      // +-0.0 the sum stays +0.0: the zero map (exact_dot.h, HYP_ZERO).
      const uint32_t lbase = smem_u32(stage + lane * DOT_LEAF), rotb = (uint32_t)(lane & (DOT_LEAF - 1)) * 8u;
      bool allz = true;
#pragma unroll
      for (int i = 0; i < DOT_LEAF; ++i) {
        const uint32_t ad = lbase | ((uint32_t)(i * 8) ^ rotb);
        const double xv = lds_f64(ad);
        allz = allz && dot_is_zero(SAME ? __dmul_rn(xv, xv) : __dmul_rn(xv, lds_f64_off<DOT_PIECE * 8>(ad)));
      }
      if (__all_sync(0xffffffffu, allz)) { wm = dot_zero_map(); fast = true; }
    } else if (fast) {
      const double c = ((hyp0 >> 11) & 1) ? -dot_pow2(kk) : dot_pow2(kk);
      double A[4] = { 0.0, 0.0, 0.0, 0.0 }, AB[4] = { 0.0, 0.0, 0.0, 0.0 };
      unsigned dm[4] = { 0, 0, 0, 0 }, sg = 0;
      // byte address of element i ^ rot of this lane's leaf (bank-conflict free; the leaf base is 128-byte aligned, so OR = ADD)
      const uint32_t lbase = smem_u32(stage + lane * DOT_LEAF), rotb = (uint32_t)(lane & (DOT_LEAF - 1)) * 8u;
      // loads in batches ahead of the arithmetic (the shared-memory loads are volatile asm: the compiler keeps them where
      // they are written, and a load placed next to its use would expose its latency sixteen times); two batches of 8 when
      // the register budget is 64 (NW = 32), one batch of 16 otherwise
      constexpr int NB = NW > 24 ? 8 : DOT_LEAF;
#pragma unroll
      for (int i0 = 0; i0 < DOT_LEAF; i0 += NB) {
        double pr[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          const uint32_t ad = lbase | ((uint32_t)((i0 + i) * 8) ^ rotb);
          const double xv = lds_f64(ad);
          pr[i] = SAME ? __dmul_rn(xv, xv) : __dmul_rn(xv, lds_f64_off<DOT_PIECE * 8>(ad));
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) MFE_DOT_SIMPLE_TERM((i0 + i) & 3, __dmul_rn(pr[i], c))
      }
      double asum = __dadd_rn(__dadd_rn(A[0], A[1]), __dadd_rn(A[2], A[3]));
      double absum = __dadd_rn(__dadd_rn(AB[0], AB[1]), __dadd_rn(AB[2], AB[3]));
      unsigned dmax = max(max(dm[0], dm[1]), max(dm[2], dm[3]));
      // piece totals (exact integers: accepted only below 2^51, where every partial sum is exact whatever the order)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        asum = __dadd_rn(asum, __shfl_xor_sync(0xffffffffu, asum, o));
        absum = __dadd_rn(absum, __shfl_xor_sync(0xffffffffu, absum, o));
        dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        sg |= __shfl_xor_sync(0xffffffffu, sg, o);
      }
      const double pos = (absum + asum) * 0.5, ng = (asum - absum) * 0.5;
      const double Sp = __dmul_rn(pre0, c);
      const bool nonneg = !(sg >> 31);
      fast = dmax < 0x3fe00000u && absum < 2251799813685248.0 &&
             (nonneg || (Sp + ng) - MFE_DOT_FAR_MARGIN > 4503599627370496.0) && (Sp + pos) + MFE_DOT_FAR_MARGIN < 9007199254740992.0;
      wm.a0 = (long long)asum; wm.a1 = wm.a0; wm.lo = nonneg ? 1LL : (long long)ng; wm.hi = (long long)pos; wm.hyp = hyp0; wm.pad = 0;
    }
    if (!fast) wm = dot_piece_slow<SAME>(n, x, y, w, piece, stage, pre0, pre1, piece_dirty);      // (also a zero-hypothesis piece with a nonzero term)
    __syncwarp();
    if (lane == 0 && j + NW < np) issue(piece + NW);           // the stage is free: next piece in flight while we publish

    // publish the piece map.  The pieces of a chunk belong to this CTA (except at the two ends of its range), so
    // arrivals are counted in shared memory with a CTA-scope fence; only the warp that completes a chunk pays the
    // device-scope fence + ticket of the chunk level.
    const int c = piece / DOT_PIECES_PER_CHUNK;
    const int cfirst = c * DOT_PIECES_PER_CHUNK;
    const int nin = min(DOT_PIECES_PER_CHUNK, npieces - cfirst);                         // pieces the chunk has
    const int mine = min(p0 + np, cfirst + nin) - max(p0, cfirst);                       // ... of which in this CTA's range
    int last = 0;
    if (lane == 31) {
      wm.pad = piece_dirty ? 1 : 0;
      w.warp_maps[piece] = wm;
      __threadfence_block();
      last = atomicAdd(&ccnt[c - p0 / DOT_PIECES_PER_CHUNK], 1) == mine - 1;
      if (last && mine != nin) {                     // chunk shared with a neighbouring CTA: device-wide ticket
        __threadfence();
        last = atomicAdd(&w.chunk_cnt[c], mine) + mine == nin;
        if (last) w.chunk_cnt[c] = 0;
      }
    }
    last = __shfl_sync(0xffffffffu, last, 31);
    if (last) {
      __threadfence();
      // the (up to) 8 piece maps sit on lanes 24..31; three ordered steps leave their composition on lane 31
      DotMap cm = dot_identity();
      const int q = lane - (32 - DOT_PIECES_PER_CHUNK);
      if (q >= 0 && q < nin) { cm = load_map_cg(w.warp_maps + cfirst + q); cm.pad = 0; }
      cm = warp_compose<DOT_PIECES_PER_CHUNK>(cm, lane);
      if (lane == 31) cm.pad = 1;
      dot_publish_chunk_map(w, cm, lane, c, nchunks);
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
// multi-rank over peer memory, for the dot modes that do not end in k_dot_walk: `result` holds the local dot
__global__ void __launch_bounds__(32) k_dot_peer_finish(double *result, CgScalars *sc, int epi, int k, double *normr_hist, double *rtrans_hist,
                                                        double *pap_hist, double *alpha_hist, const CgScalars *gate, PeerDot pd)
{
  if (cg_gate_closed(gate, k)) return;
  bool ok = true;
  const double g = peer_dot_allreduce(pd, threadIdx.x == 0 ? *result : 0.0, ok);
  if (threadIdx.x == 0) {
    *result = g;
    if (sc) {
      if (!ok) { sc->status = CG_COMM_ERROR; sc->stop_iter = -1; }
      else if (epi != EPI_NONE) cg_epilogue(sc, epi, k, g, normr_hist, rtrans_hist, pap_hist, alpha_hist);
    }
  }
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

}  // namespace mfe
