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
// A run whose terms are all >= 0 relative to the sign of the sum can never dip below its entry value, whatever S
// is -- in particular when S sits EXACTLY on the lower binade boundary (S = 2^52: the running sum is a power of
// two, which miniFE's first dots produce: r = b holds exact 1.0s followed by long runs of zeros).  Such runs carry
// lo = 1, the tightest bound that dot_apply's test  S + lo >= 2^52 + 1  accepts for every S; composition keeps it.
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
enum { DOT_CTA_THREADS = 256 };         // leaves per chunk (and threads of the per-chunk producer kernels)
enum { DOT_CHUNK = DOT_LEAF * DOT_CTA_THREADS };   // 4096 terms per chunk
enum { DOT_PIECE = DOT_LEAF * 32 };                // 512 terms per piece (one warp: 32 leaves)
enum { DOT_PIECES_PER_CHUNK = DOT_CHUNK / DOT_PIECE };
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

// the term does not pull the running sum towards zero: it is zero or has the sign of the sum
MFE_HD bool dot_term_nonneg(double p, int hyp)
{
  const uint64_t u = MFE_D2U(p);
  return (u << 1) == 0 || ((u >> 63) != 0) == (((hyp >> 11) & 1) != 0);
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
MFE_HD unsigned dot_hi32(double v) { return (unsigned)(MFE_D2U(v) >> 32); }

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
  bool nonneg = true;
  for (long long i = begin; i < end; ++i) {
    long long t0, t1;
    const double p = MFE_DMUL(x[i], y[i]);
    if (!dot_term_offsets(p, hyp, t0, t1)) return dot_invalid();
    nonneg = nonneg && dot_term_nonneg(p, hyp);
    dot_append(m, t0, t1);
  }
  if (nonneg) m.lo = 1;
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
  bool nonneg = true;
  for (int i = 0; i < n; ++i) {
    long long t0, t1;
    if (!dot_term_offsets(prod[i], hyp, t0, t1)) return dot_invalid();
    nonneg = nonneg && dot_term_nonneg(prod[i], hyp);
    dot_append(m, t0, t1);
  }
  if (nonneg) m.lo = 1;
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

MFE_HD double dot_fmin(double a, double b) { return a < b ? a : b; }
MFE_HD double dot_fmax(double a, double b) { return a > b ? a : b; }

// One term of the fp formulation.  vmax collects max|v| (range check deferred to the end of the leaf), nt counts
// exact ties: after nt ties the odd-entry trajectory differs from the even-entry one by at most nt, so lo/hi are
// tracked on the even-entry offset A0 only and widened by nt at the end (conservative).
#define MFE_DOT_FP_TERM(p)                                                                                  \
  {                                                                                                         \
    const double v = MFE_DMUL(MFE_DMUL((p), c1), c2);                                                       \
    const double av = fabs(v);                                                                              \
    vmax = av > vmax ? av : vmax;                                                                           \
    sg |= dot_hi32(v) & (dot_is_zero(v) ? 0u : 0x80000000u);   /* a NEGATIVE scaled term (v = -0.0 is harmless) */ \
    const double r = rint(v);                                                                               \
    if (fabs(v - r) == 0.5) {                              /* exact tie: even candidate r, odd 2v - r */  \
      const double o = MFE_DADD(MFE_DADD(v, v), -r);                                                        \
      A0 += dot_is_odd(A0) ? o : r;                                                                         \
      A1 += dot_is_odd(A1) ? r : o;                                                                         \
      nt += 1.0;                                                                                            \
    } else { A0 += r; A1 += r; }                                                                            \
    lo = A0 < lo ? A0 : lo;                                                                                 \
    hi = A0 > hi ? A0 : hi;                                                                                 \
  }

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
  double A0 = 0.0, A1 = 0.0, lo = 1.152921504606846976e18, hi = -1.152921504606846976e18, vmax = 0.0, nt = 0.0;
  unsigned sg = 0;
  if (n == DOT_LEAF) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < DOT_LEAF; ++i) MFE_DOT_FP_TERM(prod[i])
  } else {
    for (int i = 0; i < n; ++i) MFE_DOT_FP_TERM(prod[i])
  }
  lo -= nt; hi += nt;
  if (!(sg >> 31)) lo = 1.0;                          // no term pulls towards zero: the run never dips below its entry value
  // |v| >= 2^52 (or inf): the term alone leaves the binade.  nan: A0 is nan.  dot_apply needs 2^52 < S + lo and
  // S + hi < 2^53 with 2^52 <= S < 2^53, so a usable map has |lo|,|hi| < 2^52.
  if (!(vmax < 4503599627370496.0) || !(A0 == A0) || !(A1 == A1) || !(lo > -4503599627370496.0 && hi < 4503599627370496.0))
    return dot_invalid();
  DotMap m;
  m.a0 = (long long)A0; m.a1 = (long long)A1; m.lo = (long long)lo; m.hi = (long long)hi; m.hyp = hyp; m.pad = 0;
  return m;
}

// Fast path of the fp formulation for the common leaf: no exact tie, all offsets of one sign.  Then a0 == a1, the
// partial offsets are monotone, and lo/hi are the first and last partial offsets -- no per-term min/max needed
// (FP64 min/max costs ~6 instructions on sm_100).  Anything else falls back to dot_leaf_map_products_fp.

// Full leaf (DOT_LEAF products), everything unrolled so that `prod` can live in registers on the device.
MFE_HD DotMap dot_leaf_map16_fp(const double *prod, int hyp)
{
  const int k = 1075 - (hyp & 0x7ff);                 // v = p * 2^k, k in [-971, 1074]
  const int k1 = k / 2, k2 = k - k1;
  const double c1 = dot_pow2(k1);
  const double c2 = ((hyp >> 11) & 1) ? -dot_pow2(k2) : dot_pow2(k2);
  double A0 = 0.0, A1 = 0.0, lo = 1.152921504606846976e18, hi = -1.152921504606846976e18, vmax = 0.0, nt = 0.0;
  unsigned sg = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 0; i < DOT_LEAF; ++i) MFE_DOT_FP_TERM(prod[i])
  lo -= nt; hi += nt;
  if (!(sg >> 31)) lo = 1.0;
  if (!(vmax < 4503599627370496.0) || !(A0 == A0) || !(A1 == A1) || !(lo > -4503599627370496.0 && hi < 4503599627370496.0))
    return dot_invalid();
  DotMap m;
  m.a0 = (long long)A0; m.a1 = (long long)A1; m.lo = (long long)lo; m.hi = (long long)hi; m.hyp = hyp; m.pad = 0;
  return m;
}

// The common leaf in "simple" form: no exact tie, so a0 == a1 = a = sum of the rounded offsets, whatever their signs.
// Its bounds are CONSERVATIVE: every partial offset lies in [neg, pos], the sums of the negative / positive offsets
// (order-independent, so the caller may visit the 16 products in any order).  Conservative bounds can only make
// dot_apply reject a map whose running sum comes within one leaf's worth of offsets of a binade boundary, so they
// are used only when the PREDICTED significand Sp = |pre| / ulp is farther than DOT_FAR_MARGIN ulps from both ends of
// the binade; otherwise the function returns false and the caller takes the exact formulation
// (dot_leaf_map16_fp).  A wrong prediction therefore costs time, never correctness.
// Returns true and (a, lo, hi) as exact integers held in doubles.  Requires hyp >= 0 and hyp != HYP_ZERO.
// |v| >= 2^51, inf and nan all surface in absum (= sum of the |rounded offsets|), so no per-term range check is needed.
#define MFE_DOT_FAR_MARGIN 16777216.0   /* 2^24 ulps ~ 4e-9 relative; the prediction is good to ~1e-13 */
// Rounding to the nearest integer (ties to even) is done with the 1.5 * 2^52 trick -- (v + M) - M, exact for
// |v| < 2^51 -- instead of rint(): two FP64 adds, no conversion-pipe instruction.  For |v| >= 2^51 the trick no
// longer rounds correctly, but then |r| >= 2^51 - 1 and the absum < 2^51 test below rejects the leaf.  The offsets
// are exact integers, so their sums are order independent and kept in four partial accumulators (short dependent
// chains; the GPU kernel is latency bound on exactly these chains).
#define MFE_DOT_RNE_MAGIC 6755399441055744.0   /* 1.5 * 2^52 */
#define MFE_DOT_SIMPLE_TERM(k, vexpr)                                                                      \
  {                                                                                                         \
    const double v = (vexpr);                                                                               \
    const double r = MFE_DADD(MFE_DADD(v, MFE_DOT_RNE_MAGIC), -MFE_DOT_RNE_MAGIC);                          \
    const double d = MFE_DADD(v, -r);                                                                       \
    A[k] = MFE_DADD(A[k], r);                                                                               \
    AB[k] = MFE_DADD(AB[k], fabs(r));                                                                       \
    const unsigned hd = dot_hi32(d) & 0x7fffffffu;                                                          \
    dm[k] = hd > dm[k] ? hd : dm[k];                                                                        \
    sg |= dot_hi32(v);                     /* sign bits of the scaled terms (-0.0 counts as negative: harmless) */ \
  }
MFE_HD bool dot_leaf16_simple(const double *prod, int hyp, double pre, double &a, double &lo, double &hi)
{
  const int k = 1075 - (hyp & 0x7ff);                 // v = p * 2^k, k in [-971, 1074]
  const bool neg = ((hyp >> 11) & 1) != 0;
  double A[4] = { 0.0, 0.0, 0.0, 0.0 }, AB[4] = { 0.0, 0.0, 0.0, 0.0 }, Sp;
  unsigned dm[4] = { 0, 0, 0, 0 };                    // max high word of |v - r|: 0x3fe00000 <=> an exact tie
  unsigned sg = 0;
  if (k <= 1023) {                                    // 2^k is a normal double: one exact scaling
    const double c = neg ? -dot_pow2(k) : dot_pow2(k);
    Sp = MFE_DMUL(pre, c);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < DOT_LEAF; ++i) MFE_DOT_SIMPLE_TERM(i & 3, MFE_DMUL(prod[i], c))
  } else {
    const int k1 = k / 2, k2 = k - k1;
    const double c1 = dot_pow2(k1);
    const double c2 = neg ? -dot_pow2(k2) : dot_pow2(k2);
    Sp = MFE_DMUL(MFE_DMUL(pre, c1), c2);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < DOT_LEAF; ++i) MFE_DOT_SIMPLE_TERM(i & 3, MFE_DMUL(MFE_DMUL(prod[i], c1), c2))
  }
  const double Asum = MFE_DADD(MFE_DADD(A[0], A[1]), MFE_DADD(A[2], A[3]));
  const double absum = MFE_DADD(MFE_DADD(AB[0], AB[1]), MFE_DADD(AB[2], AB[3]));
  unsigned dmax = dm[0] > dm[1] ? dm[0] : dm[1];
  const unsigned dmax2 = dm[2] > dm[3] ? dm[2] : dm[3];
  dmax = dmax > dmax2 ? dmax : dmax2;
  // absum < 2^51 bounds every |r|, |A|, pos and -neg, keeps the rounding trick exact, and rejects inf / nan
  if (!(dmax < 0x3fe00000u && absum < 2251799813685248.0)) return false;
  const double pos = (absum + Asum) * 0.5, ng = (Asum - absum) * 0.5;     // exact: absum +- A are even integers < 2^53
  // no term pulls towards zero: the low end of the binade cannot be crossed (lo = 1, see the header); this is what
  // keeps a running sum that sits exactly on a power of two (S = 2^52) out of the slow path
  const bool nonneg = !(sg >> 31);
  const bool far = (nonneg || (Sp + ng) - MFE_DOT_FAR_MARGIN > 4503599627370496.0) && (Sp + pos) + MFE_DOT_FAR_MARGIN < 9007199254740992.0;
  if (!far) return false;
  a = Asum; lo = nonneg ? 1.0 : ng; hi = pos;
  return true;
}

// A leaf whose exact bounds, applied to the PREDICTED significand of the running sum, come within
// MFE_DOT_RISK_MARGIN ulps of either end of the binade will probably fail dot_apply (the running sum dips across
// the boundary inside the leaf even though it starts and ends in the same binade -- mixed-sign data).  Such
// leaves are listed as dirty, so that the walker replays them instead of discovering the failure in a collapsed
// gap (whose recovery is slow).  Prediction only: a miss costs time, never correctness.
#define MFE_DOT_RISK_MARGIN 1048576.0   /* 2^20 ulps */
MFE_HD bool dot_leaf_risky(const DotMap &m, int hyp, double pre)
{
  if (m.hyp < 0 || m.hyp == HYP_ZERO) return false;
  const int k = 1075 - (hyp & 0x7ff);
  const int k1 = k / 2, k2 = k - k1;
  const double Sp = fabs(MFE_DMUL(MFE_DMUL(pre, dot_pow2(k1)), dot_pow2(k2)));     // |pre| / ulp, in [2^52, 2^53)
  const bool lo_safe = m.lo >= 1 || (Sp + (double)m.lo) - MFE_DOT_RISK_MARGIN > 4503599627370496.0;
  return !(lo_safe && (Sp + (double)m.hi) + MFE_DOT_RISK_MARGIN < 9007199254740992.0);
}

MFE_HD DotMap dot_leaf_map16(const double *prod, int hyp, double pre)
{
  if (hyp < 0) return dot_invalid();
  if (hyp == HYP_ZERO) {
    bool allz = true;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < DOT_LEAF; ++i) allz = allz && dot_is_zero(prod[i]);
    return allz ? dot_zero_map() : dot_invalid();
  }
  double a, lo, hi;
  if (!dot_leaf16_simple(prod, hyp, pre, a, lo, hi)) return dot_leaf_map16_fp(prod, hyp);
  DotMap m;
  m.a0 = (long long)a; m.a1 = m.a0; m.lo = (long long)lo; m.hi = (long long)hi; m.hyp = hyp; m.pad = 0;
  return m;
}

// pre = predicted running sum at the start of the leaf (only steers the choice of formulation)
MFE_HD DotMap dot_leaf_map_products_fast(const double *prod, int n, int hyp, double pre)
{
  if (n != DOT_LEAF) return dot_leaf_map_products_fp(prod, n, hyp);
  return dot_leaf_map16(prod, hyp, pre);
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
  long long i = begin;
#ifdef __CUDA_ARCH__
  for (; i + 16 <= end; i += 16) {        // 32 independent loads in flight, then the ordered adds
    double p[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) p[q] = MFE_DMUL(x[i + q], y[i + q]);
#pragma unroll
    for (int q = 0; q < 16; ++q) s = MFE_DADD(s, p[q]);
  }
#endif
  for (; i < end; ++i) s = MFE_DADD(s, MFE_DMUL(x[i], y[i]));
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// Map hierarchy + the sequential part of the walk (shared by the kernels and the CPU test harness).
//   level 0  leaf   DOT_LEAF terms            leaf_maps   (stored only for pieces that hold a dirty leaf)
//   level 1  piece  32 leaves (one warp)      warp_maps   (always stored; pad = 1 when its leaf maps are stored)
//   level 2  chunk  8 pieces = 256 leaves     cta_maps
//   level 3  super  32 chunks = 8192 leaves   super_maps
//   level 4  hyper  32 supers = 262144 leaves hyper_maps
// A clean stretch of leaves [lo,hi) between two dirty leaves decomposes, segment-tree style, into <= 9 runs of
// consecutive same-level maps (ascending levels, the hypers, descending levels), each run <= 31 long below the
// top level: one warp collapses a run with a single coalesced load and an ordered shuffle scan.
// ---------------------------------------------------------------------------------------------------------
enum { DOT_LEVELS = 5, DOT_GAP_PIECES = 9 };
enum { DOT_STATS_HEAD = 16, DOT_STATS_TRACE = 512, DOT_STATS_WORDS = DOT_STATS_HEAD + 2 * DOT_STATS_TRACE };
MFE_HD int dot_level_size(int level)   // in leaves
{
  return level == 0 ? 1 : level == 1 ? 32 : level == 2 ? 256 : level == 3 ? 8192 : 262144;
}

struct DotWork {
  double *chunk_sums;      // nchunks plain partial sums of the products (binade prediction; written by the producer kernel)
  double *piece_sums;      // npieces = 8 per chunk, same purpose
  double *chunk_prefix;    // nchunks + 1 exclusive prefix of chunk_sums (MFE_DOT_TREE only)
  DotMap *leaf_maps;       // level 0: nchunks * 256 (stored only for pieces that hold a dirty leaf)
  DotMap *warp_maps;       // level 1: npieces       (pad = 1: the 32 leaf maps of this piece are stored)
  DotMap *cta_maps;        // level 2: nchunks       (pad = 1: the level-1 maps below are stored -- always)
  DotMap *super_maps;      // level 3: ceil(nchunks / 32)
  DotMap *hyper_maps;      // level 4: ceil(nsupers / 32)
  int *chunk_cnt;          // arrival tickets (self-resetting): pieces of a chunk
  int *super_cnt;          //                                    chunks of a super
  int *hyper_cnt;          //                                    supers of a hyper
  int *dirty_count;        // 1
  int *dirty_list;         // DOT_MAX_DIRTY leaf indices
  double *result;          // 1
  unsigned long long *stats; // optional diagnostics (DOT_STATS_WORDS words): see k_dot_walk
};
MFE_HD const DotMap *dot_level_maps(const DotWork &w, int level)
{
  return level == 0 ? w.leaf_maps : level == 1 ? w.warp_maps : level == 2 ? w.cta_maps : level == 3 ? w.super_maps : w.hyper_maps;
}

struct WalkPiece { int level; int begin; int count; };   // `count` consecutive level-`level` maps starting at `begin`

MFE_HD int dot_level_shift(int level) { return level == 0 ? 0 : level == 1 ? 5 : level == 2 ? 8 : level == 3 ? 13 : 18; }

// clean leaves [lo,hi) -> out[DOT_GAP_PIECES] (unused entries have count 0), in application order.
// Closed form of: ascend taking level-k maps until `a` is aligned to level k+1 (or hi is reached), take the
// top-level maps, then descend taking as many level-k maps as still fit below hi.
MFE_HD void dot_gap_decompose(long long lo, long long hi, WalkPiece *out)
{
  long long a = lo < hi ? lo : hi;
  int np = 0;
  for (int k = 0; k < DOT_LEVELS - 1; ++k) {
    const int sh = dot_level_shift(k), shu = dot_level_shift(k + 1);
    const long long first = a >> sh;
    long long aligned = ((a + ((1LL << shu) - 1)) >> shu) << shu;      // next multiple of the upper level
    const long long fit = a + (((hi - a) >> sh) << sh);                // last position reachable with whole level-k maps
    const long long stop = aligned < fit ? aligned : fit;
    const long long cnt = (a & ((1LL << sh) - 1)) == 0 ? (stop - a) >> sh : 0;
    out[np].level = k; out[np].begin = (int)first; out[np].count = (int)(cnt > 0 ? cnt : 0); ++np;
    if (cnt > 0) a += cnt << sh;
  }
  {
    const int sh = dot_level_shift(DOT_LEVELS - 1);
    const long long ok = (a & ((1LL << sh) - 1)) == 0;
    const long long cnt = ok ? (hi - a) >> sh : 0;
    out[np].level = DOT_LEVELS - 1; out[np].begin = (int)(a >> sh); out[np].count = (int)(cnt > 0 ? cnt : 0); ++np;
    if (cnt > 0) a += cnt << sh;
  }
  for (int k = DOT_LEVELS - 2; k >= 0; --k) {
    const int sh = dot_level_shift(k);
    const long long ok = (a & ((1LL << sh) - 1)) == 0;
    const long long cnt = ok ? (hi - a) >> sh : 0;
    out[np].level = k; out[np].begin = (int)(a >> sh); out[np].count = (int)(cnt > 0 ? cnt : 0); ++np;
    if (cnt > 0) a += cnt << sh;
  }
}

// Slow but always-correct application of one map of the hierarchy: descend into the children when the map fails
// its check; replay the terms when there is nothing stored below.
MFE_HD double dot_apply_or_descend(const DotWork &w, int level, long long index, const double *x, const double *y, long long n, double s)
{
  // explicit stack instead of recursion (device friendly): (level, index) pairs, children pushed in reverse order
  int lv[DOT_LEVELS * 32]; long long ix[DOT_LEVELS * 32]; int sp = 0;
  lv[sp] = level; ix[sp] = index; ++sp;
  while (sp > 0) {
    --sp;
    const int L = lv[sp]; const long long I = ix[sp];
    const long long b = I * dot_level_size(L) * DOT_LEAF;
    if (b >= n) continue;
    const DotMap m = dot_level_maps(w, L)[I];
    if (dot_apply(m, s)) continue;
    const bool children_stored = L >= 3 || ((L == 2 || L == 1) && m.pad == 1);
    if (L == 0 || !children_stored) {
      long long e = b + (long long)dot_level_size(L) * DOT_LEAF; if (e > n) e = n;
      s = dot_replay(x, y, b, e, s);
      continue;
    }
    const int fan = L == 2 ? 8 : 32;
    for (int q = fan - 1; q >= 0; --q) { lv[sp] = L - 1; ix[sp] = I * fan + q; ++sp; }
  }
  return s;
}

// a gap whose collapsed map failed: piece by piece, map by map
MFE_HD double dot_walk_gap_slow(long long lo, long long hi, const DotWork &w, const double *x, const double *y, long long n, double s)
{
  WalkPiece pc[DOT_GAP_PIECES];
  dot_gap_decompose(lo, hi, pc);
  for (int q = 0; q < DOT_GAP_PIECES; ++q)
    for (int i = 0; i < pc[q].count; ++i) s = dot_apply_or_descend(w, pc[q].level, (long long)pc[q].begin + i, x, y, n, s);
  return s;
}

}  // namespace mfe
