// exact_dot.h -- order-reproducing parallel dot product.
//
// miniFE's dot (ref/src/Vector_functions.hpp:194-226) is the serial loop  s = fl(s + fl(x[i]*y[i])).  The CG
// residual history is chaotically sensitive to that summation order (SURVEY.md section 0 finding 4), so the
// device dot must return the SAME BITS as the serial loop while running in parallel.
//
// Idea: while the running sum s stays inside one binade [2^e, 2^(e+1)) its significand is an integer
// S in [2^52, 2^53) in units of U = ulp = 2^(e-52), and adding a term p (round-to-nearest-even) is
//     S  ->  S + t,   t = rne(p/U)  with ties resolved by the parity of S,
// i.e. every term is a map  S -> S + (S even ? t0 : t1)  on integers.  Such maps are closed under
// composition (the state is the pair (a0,a1) of net offsets for an even / odd entry), composition is
// associative, so a run of terms collapses to one DotMap by a parallel ordered reduction -- provided the sum
// never leaves the binade inside the run.  That proviso is tracked conservatively with (lo,hi), the extreme
// intermediate offsets, and checked against the actual S when the map is applied.
//
// The binade ("hypothesis") for each leaf of LEAF consecutive terms is predicted from a conventional
// floating-point prefix sum (accurate to ~1e-16 relative, more than enough to name the binade away from its
// boundaries).  Leaves where the predicted binade changes, or whose map is unusable, are "dirty": they are
// replayed term by term in plain floating point by the sequential walker.  Correctness never depends on the
// prediction: a wrong guess only makes a map fail its (hyp, lo, hi) check, which falls back to replay.
#pragma once
#include "mesh.h"
#include <stdint.h>
#include <math.h>

#ifdef __CUDA_ARCH__
#define MFE_DMUL(a, b) __dmul_rn((a), (b))
#define MFE_DADD(a, b) __dadd_rn((a), (b))
#define MFE_D2U(x) ((uint64_t)__double_as_longlong(x))
#define MFE_U2D(x) __longlong_as_double((long long)(x))
#else
#include <string.h>
#define MFE_DMUL(a, b) ((a) * (b))
#define MFE_DADD(a, b) ((a) + (b))
static inline uint64_t mfe_d2u_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double mfe_u2d_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define MFE_D2U(x) mfe_d2u_(x)
#define MFE_U2D(x) mfe_u2d_(x)
#endif

namespace mfe {

enum { DOT_LEAF = 16 };                 // terms per leaf (one thread)
enum { DOT_CTA_THREADS = 256 };         // leaves per chunk
enum { DOT_CHUNK = DOT_LEAF * DOT_CTA_THREADS };   // 4096 terms per chunk (one CTA)
enum { DOT_MAX_DIRTY = 256 };           // more dirty leaves than this -> whole-vector serial replay

enum { HYP_INVALID = -1,   // map unusable: replay
       HYP_ANY = -2,       // identity (empty run): composes with anything
       HYP_NONE = -3,      // value has no binade (-0, subnormal, inf, nan): never matches a map
       HYP_ZERO = 0x1000 };// running sum is exactly +0.0 and every term of the run is +-0.0: the sum stays +0.0

struct DotMap {
  long long a0, a1;   // net significand offset when the run is entered with S even / odd
  long long lo, hi;   // min / max offset after any term of the run (over both parities)
  int hyp;            // (sign << 11) | biased exponent of the running sum this map assumes
  int pad;
};

#define MFE_DOT_BIG (1LL << 60)
#define MFE_DOT_LIMIT (1LL << 54)

MFE_HD DotMap dot_identity()
{
  DotMap m; m.a0 = 0; m.a1 = 0; m.lo = MFE_DOT_BIG; m.hi = -MFE_DOT_BIG; m.hyp = HYP_ANY; m.pad = 0; return m;
}
MFE_HD DotMap dot_invalid()
{
  DotMap m; m.a0 = 0; m.a1 = 0; m.lo = 0; m.hi = 0; m.hyp = HYP_INVALID; m.pad = 0; return m;
}

// hypothesis of a value: sign and biased exponent of a normal nonzero finite double
MFE_HD int dot_hyp_of(double s)
{
  const uint64_t u = MFE_D2U(s);
  const int e = (int)((u >> 52) & 0x7ff);
  if (u == 0) return HYP_ZERO;
  if (e == 0 || e == 0x7ff) return HYP_NONE;
  return (int)((u >> 63) << 11) | e;
}

// Offsets (t0 for S even, t1 for S odd) that adding the term p makes to the significand of a running sum
// with hypothesis hyp.  Returns false when the term cannot be expressed (inf/nan or far larger than the sum).
MFE_HD bool dot_term_offsets(double p, int hyp, long long &t0, long long &t1)
{
  const uint64_t u = MFE_D2U(p);
  const int ep = (int)((u >> 52) & 0x7ff);
  uint64_t mant = u & 0x000fffffffffffffULL;
  if (ep == 0x7ff) return false;
  if (ep != 0) mant |= (1ULL << 52);
  if (mant == 0) { t0 = 0; t1 = 0; return true; }
  const int es = hyp & 0x7ff;
  const int d = (ep ? ep : 1) - es;          // p = mant * 2^d units
  long long m0, m1;
  if (d > 1) return false;                   // |p| >= 2^(e+2): the sum certainly leaves the binade
  if (d >= 0) { m0 = m1 = (long long)(mant << d); }
  else {
    const int sh = -d;
    if (sh >= 64) { m0 = m1 = 0; }
    else {
      const uint64_t q = mant >> sh;
      const uint64_t rem = mant & ((1ULL << sh) - 1ULL);
      const uint64_t half = 1ULL << (sh - 1);
      if (rem > half) { m0 = m1 = (long long)(q + 1); }
      else if (rem < half) { m0 = m1 = (long long)q; }
      else {                                  // exact tie: the result must be even
        m0 = (long long)(q + (q & 1ULL));          // S even: S +- (q + r) even  <=>  r = q & 1
        m1 = (long long)(q + ((q & 1ULL) ^ 1ULL)); // S odd
      }
    }
  }
  const bool negative = ((u >> 63) != 0) != (((hyp >> 11) & 1) != 0);   // relative to the sign of the sum
  t0 = negative ? -m0 : m0;
  t1 = negative ? -m1 : m1;
  return true;
}

// append one term to a map (the map of a run followed by the term)
MFE_HD void dot_append(DotMap &m, long long t0, long long t1)
{
  m.a0 += (m.a0 & 1) ? t1 : t0;                 // entered even: parity now = parity(a0)
  m.a1 += ((m.a1 + 1) & 1) ? t1 : t0;           // entered odd:  parity now = parity(1 + a1)
  const long long mn = m.a0 < m.a1 ? m.a0 : m.a1, mx = m.a0 < m.a1 ? m.a1 : m.a0;
  if (mn < m.lo) m.lo = mn;
  if (mx > m.hi) m.hi = mx;
}

// ordered composition: run f followed by run g
MFE_HD DotMap dot_compose(const DotMap &f, const DotMap &g)
{
  if (f.hyp == HYP_ANY) return g;
  if (g.hyp == HYP_ANY) return f;
  if (f.hyp == HYP_INVALID || g.hyp == HYP_INVALID || f.hyp != g.hyp) return dot_invalid();
  DotMap h;
  h.hyp = f.hyp; h.pad = 0;
  h.a0 = f.a0 + ((f.a0 & 1) ? g.a1 : g.a0);
  h.a1 = f.a1 + (((f.a1 + 1) & 1) ? g.a1 : g.a0);
  const long long mn = f.a0 < f.a1 ? f.a0 : f.a1, mx = f.a0 < f.a1 ? f.a1 : f.a0;
  h.lo = f.lo < mn + g.lo ? f.lo : mn + g.lo;
  h.hi = f.hi > mx + g.hi ? f.hi : mx + g.hi;
  if (h.lo < -MFE_DOT_LIMIT || h.hi > MFE_DOT_LIMIT) return dot_invalid();   // cannot be valid for any S; avoids overflow
  return h;
}

// Map of the terms x[i]*y[i], i in [begin,end), under hypothesis hyp.
MFE_HD DotMap dot_zero_map()
{
  DotMap m; m.a0 = 0; m.a1 = 0; m.lo = 0; m.hi = 0; m.hyp = HYP_ZERO; m.pad = 0; return m;
}
MFE_HD bool dot_is_zero(double p) { return (MFE_D2U(p) << 1) == 0; }

MFE_HD DotMap dot_leaf_map(const double *x, const double *y, long long begin, long long end, int hyp)
{
  if (end <= begin) return dot_identity();
  if (hyp < 0) return dot_invalid();
  if (hyp == HYP_ZERO) {
    for (long long i = begin; i < end; ++i) if (!dot_is_zero(MFE_DMUL(x[i], y[i]))) return dot_invalid();
    return dot_zero_map();
  }
  DotMap m = dot_identity();
  m.hyp = hyp;
  for (long long i = begin; i < end; ++i) {
    long long t0, t1;
    if (!dot_term_offsets(MFE_DMUL(x[i], y[i]), hyp, t0, t1)) return dot_invalid();
    dot_append(m, t0, t1);
  }
  return m;
}
// same from precomputed products
MFE_HD DotMap dot_leaf_map_products(const double *prod, int n, int hyp)
{
  if (n <= 0) return dot_identity();
  if (hyp < 0) return dot_invalid();
  if (hyp == HYP_ZERO) {
    for (int i = 0; i < n; ++i) if (!dot_is_zero(prod[i])) return dot_invalid();
    return dot_zero_map();
  }
  DotMap m = dot_identity();
  m.hyp = hyp;
  for (int i = 0; i < n; ++i) {
    long long t0, t1;
    if (!dot_term_offsets(prod[i], hyp, t0, t1)) return dot_invalid();
    dot_append(m, t0, t1);
  }
  return m;
}

// Floating-point formulation of dot_leaf_map_products (same maps, ~4x fewer instructions on the GPU):
// v = p/U is formed exactly by two power-of-two scalings, t = rint(v) is the round-to-nearest-even offset, and an
// exact tie (|v - trunc(v)| == 0.5) offers the even candidate r and the odd candidate 2v - r, chosen by the parity
// of the offset accumulated so far.  Offsets are exact integers in double while |a| <= 2^53; anything larger can
// never pass dot_apply's range check, so inexactness there is harmless.
MFE_HD double dot_pow2(int k)   // 2^k for k in [-1022, 1023]
{
  return MFE_U2D((uint64_t)(k + 1023) << 52);
}
MFE_HD bool dot_is_odd(double a) { return (((long long)a) & 1LL) != 0; }

MFE_HD DotMap dot_leaf_map_products_fp(const double *prod, int n, int hyp)
{
  if (n <= 0) return dot_identity();
  if (hyp < 0) return dot_invalid();
  if (hyp == HYP_ZERO) {
    for (int i = 0; i < n; ++i) if (!dot_is_zero(prod[i])) return dot_invalid();
    return dot_zero_map();
  }
  const int k = 1075 - (hyp & 0x7ff);                 // v = p * 2^k, k in [-971, 1074]
  const int k1 = k / 2, k2 = k - k1;
  const double c1 = dot_pow2(k1);
  const double c2 = ((hyp >> 11) & 1) ? -dot_pow2(k2) : dot_pow2(k2);   // offsets relative to the sign of the sum
  double A0 = 0.0, A1 = 0.0, lo = 1.152921504606846976e18, hi = -1.152921504606846976e18;   // +-2^60
  for (int i = 0; i < n; ++i) {
    const double v = MFE_DMUL(MFE_DMUL(prod[i], c1), c2);
    if (!(fabs(v) < 4503599627370496.0)) return dot_invalid();        // |v| >= 2^52, inf or nan: leaves the binade
    const double r = rint(v);
    const double fr = v - trunc(v);
    if (fabs(fr) == 0.5) {
      const double o = MFE_DADD(MFE_DADD(v, v), -r);                    // the odd neighbour
      A0 += dot_is_odd(A0) ? o : r;                                    // entered even: keep S + A0 even
      A1 += dot_is_odd(A1) ? r : o;                                    // entered odd:  keep S + A1 even (A1 odd)
    } else { A0 += r; A1 += r; }
    const double mn = A0 < A1 ? A0 : A1, mx = A0 < A1 ? A1 : A0;
    lo = mn < lo ? mn : lo;
    hi = mx > hi ? mx : hi;
  }
  // dot_apply needs 2^52 < S + lo and S + hi < 2^53 with 2^52 <= S < 2^53: a usable map has |lo|,|hi| < 2^52
  if (!(lo > -4503599627370496.0 && hi < 4503599627370496.0)) return dot_invalid();
  DotMap m;
  m.a0 = (long long)A0; m.a1 = (long long)A1; m.lo = (long long)lo; m.hi = (long long)hi; m.hyp = hyp; m.pad = 0;
  return m;
}

// Apply a map to the running sum; false (s untouched) when the map does not provably reproduce the serial loop.
MFE_HD bool dot_apply(const DotMap &m, double &s)
{
  if (m.hyp == HYP_ANY) return true;
  if (m.hyp < 0 || m.hyp != dot_hyp_of(s)) return false;
  if (m.hyp == HYP_ZERO) return true;            // +0.0 + (+-0.0) = +0.0
  const uint64_t u = MFE_D2U(s);
  const long long S = (long long)((u & 0x000fffffffffffffULL) | (1ULL << 52));
  // every intermediate significand must stay in [2^52+1, 2^53-1]: then each exact pre-rounding value lies
  // strictly inside the binade, where rounding is to the integer grid assumed by dot_term_offsets
  if (S + m.lo < (1LL << 52) + 1 || S + m.hi > (1LL << 53) - 1) return false;
  const long long Sn = S + ((S & 1) ? m.a1 : m.a0);
  s = MFE_U2D((u & 0xfff0000000000000ULL) | ((uint64_t)Sn & 0x000fffffffffffffULL));
  return true;
}

// the literal loop of the reference over [begin,end)
MFE_HD double dot_replay(const double *x, const double *y, long long begin, long long end, double s)
{
  for (long long i = begin; i < end; ++i) s = MFE_DADD(s, MFE_DMUL(x[i], y[i]));
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// workspace + the sequential part of the walk (shared by k_dot_walk and the CPU test harness)
// ---------------------------------------------------------------------------------------------------------
struct DotWork {
  double *chunk_sums;      // nchunks plain partial sums (binade prediction)
  double *chunk_prefix;    // nchunks + 1 exclusive prefix of the above
  DotMap *cta_maps;        // nchunks
  DotMap *leaf_maps;       // nchunks * DOT_CTA_THREADS (written for dirty chunks only)
  int *dirty_count;        // 1
  int *dirty_list;         // DOT_MAX_DIRTY leaf indices
  double *result;          // 1
};

enum { WALK_MAX_PIECES = 3 * (DOT_MAX_DIRTY + 1) };
struct WalkPiece { int kind; int begin; int end; };      // kind 0: leaf maps [begin,end), 1: chunk maps [begin,end)

// The clean stretch between dirty leaf g-1 and dirty leaf g (sorted ascending; g = 0..D) splits into
// [head leaves of a dirty chunk][full clean chunks][tail leaves of the next dirty chunk].
MFE_HD void dot_gap_pieces(int g, int D, const int *sorted, int nleaves, WalkPiece *out3)
{
  const int lo = g == 0 ? 0 : sorted[g - 1] + 1, hi = g == D ? nleaves : sorted[g];
  WalkPiece p0 = { 0, 0, 0 }, p1 = { 1, 0, 0 }, p2 = { 0, 0, 0 };
  if (lo < hi) {
    const int cfirst = (lo + DOT_CTA_THREADS - 1) / DOT_CTA_THREADS;    // first full chunk
    const int clast = hi / DOT_CTA_THREADS;                              // one past the last full chunk
    if (cfirst > clast) { p0.begin = lo; p0.end = hi; }                  // inside one chunk
    else {
      p0.begin = lo; p0.end = cfirst * DOT_CTA_THREADS;                  // head leaves (may be empty)
      p1.begin = cfirst; p1.end = clast;
      p2.begin = clast * DOT_CTA_THREADS; p2.end = hi;                   // tail leaves (may be empty)
    }
  }
  out3[0] = p0; out3[1] = p1; out3[2] = p2;
}

// a piece whose collapsed map failed its check: apply its members one by one, replaying the failing ones
MFE_HD double dot_walk_piece_slow(const WalkPiece &pc, const DotWork &w, const double *x, const double *y, long long n, double s)
{
  if (pc.kind == 0) {
    for (int l = pc.begin; l < pc.end; ++l) {
      const DotMap m = w.leaf_maps[l];
      if (!dot_apply(m, s)) {
        const long long b = (long long)l * DOT_LEAF, e = b + DOT_LEAF < n ? b + DOT_LEAF : n;
        s = dot_replay(x, y, b, e, s);
      }
    }
  } else {
    for (int c = pc.begin; c < pc.end; ++c) {
      const DotMap m = w.cta_maps[c];
      if (!dot_apply(m, s)) {
        const long long b = (long long)c * DOT_CHUNK, e = b + DOT_CHUNK < n ? b + DOT_CHUNK : n;
        s = dot_replay(x, y, b, e, s);
      }
    }
  }
  return s;
}

// the sequential walk: [pieces of gap 0][dirty leaf 0][pieces of gap 1][dirty leaf 1] ... [pieces of gap D]
MFE_HD double dot_walk_final(int D, const int *sorted, const WalkPiece *pieces, const DotMap *pmaps, const DotWork &w,
                             const double *x, const double *y, long long n)
{
  double s = 0.0;
  for (int g = 0; g <= D; ++g) {
    for (int q = 0; q < 3; ++q) {
      const WalkPiece pc = pieces[3 * g + q];
      if (pc.end <= pc.begin) continue;
      if (!dot_apply(pmaps[3 * g + q], s)) s = dot_walk_piece_slow(pc, w, x, y, n, s);
    }
    if (g < D) {
      const long long b = (long long)sorted[g] * DOT_LEAF, e = b + DOT_LEAF < n ? b + DOT_LEAF : n;
      s = dot_replay(x, y, b, e, s);
    }
  }
  return s;
}

}  // namespace mfe
