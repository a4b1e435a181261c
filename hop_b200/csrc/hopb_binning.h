// Exact float32 binning against float64 edges (host + device).
//
// Replaces numpy.histogramdd's bin search (hop/density.py:129) and numpy.digitize in
// HoppingTrajectory._coord2hop (hop/trajectory.py:428), both of which promote the float32
// coordinate to float64 and do searchsorted(edges, x, side='right').
//
// Exactness argument.  For a float32 value x and a real edge e:  x >= e  <=>  x >= up32(e), where
// up32(e) is the smallest float32 that is >= e (x is itself a float32, so no float32 lies in
// [e, up32(e)) ).  up32 is monotone, hence
//     searchsorted(edges, float64(x), 'right') = #{k : edges[k] <= x} = #{k : up32(edges[k]) <= x}
// and the whole search runs on a float32 table with float32 compares -- bit-identical to the
// float64 search, with no float64 arithmetic on the device.  A float32 multiply gives the starting
// guess; two short loops walk to the exact answer, so the guess only affects speed.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define HOPB_HD __host__ __device__ __forceinline__
#else
#define HOPB_HD inline
#endif

#define HOPB_MAX_EDGES 4096  // per axis, n+1 edges (bench grids use <= ~600)

// Per-axis constants that travel by value in kernel arguments.
struct HopbAxis {
  int32_t n;        // number of bins; edges 0..n
  float lo;         // (float)edges[0]          -- starting guess only
  float inv;        // (float)(n/(edges[n]-edges[0]))
  float top_eq;     // (float)edges[n] if exactly representable, else NaN (histogramdd x == edges[-1])
  float top_round;  // float32(np.around(edges[n], decimals))   (trajectory.py:437-439)
  float p10;        // 10^|decimals|
  int32_t dec_neg;  // 1 if decimals < 0 (np.around divides then multiplies)
  int32_t off;      // offset of this axis' table in the packed float table
  int32_t off2;     // offset (in float2 units) of this axis' pair table, n+2 entries
  float hop_lo;     // [hop_lo, hop_hi]: the float32 values whose np.around() equals top_round
  float hop_hi;     //   (around32 is monotone, so the pre-image is an interval; empty: lo > hi)
};

struct HopbGridTab {
  HopbAxis ax[3];
  int32_t table_len;   // total floats in the packed up32 table
  int32_t table2_len;  // total float2 entries in the packed pair table
  int32_t fast_ok;     // 1 if the one-step corrected guess is proven exact for these edges
  float sure_eps;      // > 0: hopb_bin_sure() with this margin is proven exact for these edges; 0: off
  // optional periodic wrapping of the coordinates before they are binned (off by default: hop bins
  // AtomGroup.positions as they are and expects pre-wrapped trajectories, hop/density.py:21-28)
  int32_t wrap;        // 0: off
  float box_o[3];      // box origin
  float box_l[3];      // box lengths
  float box_inv[3];    // float32(1 / length)
};

struct HopbPair { float lo, hi; };  // pair table entry b: (up32(e[b-1]) or -inf, up32(e[b]) or +inf)

// Starting guess shared by host verification and device code: clamp in float, then floor.
HOPB_HD int hopb_bin_guess(float x, const HopbAxis& a) {
  float g = (x - a.lo) * a.inv;
  g = fminf(fmaxf(g, -1.0f), (float)a.n);  // NaN -> -1 -> bin 0 (dropped / outlier)
#if defined(__CUDA_ARCH__)
  return __float2int_rd(g) + 1;
#else
  return (int)floorf(g) + 1;
#endif
}

// searchsorted(edges, x, 'right') with one branch-free correction step.  Exact whenever
// HopbGridTab::fast_ok (the host checks, edge by edge, that the guess is never off by more than 1).
HOPB_HD int hopb_bin_right_fast(float x, const HopbAxis& a, const HopbPair* pairs) {
  int b = hopb_bin_guess(x, a);
  const HopbPair e = pairs[b];
  b += (x >= e.hi) ? 1 : 0;
  b -= (x < e.lo) ? 1 : 0;
  return b;
}

// The bin without any table: the guess g = (x - lo) * inv is monotone in x, so when its fractional
// part keeps a margin eps from 0 and 1 -- and the host has checked, at the two float32 neighbours of
// every edge, that points the guess would misplace always land inside that margin -- floor(g) IS the
// searchsorted result, and x is also clear of the last-edge rules.  Returns the bin (1..n), or 0 when
// the point needs the table (about 2 * eps of the points per axis, everything outside the grid, NaN).
HOPB_HD int hopb_bin_sure(float x, float lo, float inv, float nf, float eps) {
  const float g = (x - lo) * inv;
  const float fl = floorf(g);
  const float fr = g - fl;  // exact
  const bool sure = (g >= 0.0f) & (g < nf) & (fabsf(fr - 0.5f) <= 0.5f - eps);
  return sure ? (int)fl + 1 : 0;
}

// numpy.histogramdd bin (1..n counted) and _coord2hop bin (1..n are map cells) on the fast path
HOPB_HD int hopb_bin_hist_fast(float x, const HopbAxis& a, const HopbPair* pairs) {
  int b = hopb_bin_right_fast(x, a, pairs);
  if (x == a.top_eq) b = a.n;
  return b;
}
HOPB_HD int hopb_bin_hop_fast(float x, const HopbAxis& a, const HopbPair* pairs) {
  int b = hopb_bin_right_fast(x, a, pairs);
  if (x >= a.hop_lo && x <= a.hop_hi) b -= 1;
  return b;
}

// searchsorted(edges, x, 'right') in [0, n+1]; `eup` points at this axis' up32 table (n+1 floats).
HOPB_HD int hopb_bin_right(float x, const HopbAxis& a, const float* eup) {
  const int n = a.n;
  float g = (x - a.lo) * a.inv;
  int b;
  if (!(g > -1.0f)) b = 0;                 // also catches NaN
  else if (g >= (float)n) b = n + 1;
  else b = (int)g + 1;
  while (b <= n && x >= eup[b]) ++b;
  while (b > 0 && !(x >= eup[b - 1])) --b;  // NaN: every compare false -> b = 0 (dropped / outlier)
  return b;
}

// numpy.histogramdd: a point exactly on the last edge belongs to the last bin.
HOPB_HD int hopb_bin_hist(float x, const HopbAxis& a, const float* eup) {
  int b = hopb_bin_right(x, a, eup);
  if (x == a.top_eq) b = a.n;
  return b;  // counted iff 1 <= b <= n
}

#if defined(__CUDA_ARCH__)
#define HOPB_FMUL(a, b) __fmul_rn((a), (b))
#define HOPB_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define HOPB_FMUL(a, b) ((a) * (b))
#define HOPB_FDIV(a, b) ((a) / (b))
#endif

// Periodic image of x in [o, o + l): x - floor((x - o) * inv) * l, every step a separately rounded float32
// operation (numpy: x - np.floor((x - o) * inv) * l on float32 arrays gives the same bits).
HOPB_HD float hopb_wrap1(float x, float o, float l, float inv) {
#if defined(__CUDA_ARCH__)
  const float t = floorf(__fmul_rn(__fsub_rn(x, o), inv));
  return __fsub_rn(x, __fmul_rn(t, l));
#else
  volatile float d = x - o;
  volatile float q = d * inv;
  volatile float t = floorf(q);
  volatile float m = t * l;
  return x - m;
#endif
}

HOPB_HD void hopb_wrap_point(float& x, float& y, float& z, const HopbGridTab& tab) {
  if (tab.wrap) {
    x = hopb_wrap1(x, tab.box_o[0], tab.box_l[0], tab.box_inv[0]);
    y = hopb_wrap1(y, tab.box_o[1], tab.box_l[1], tab.box_inv[1]);
    z = hopb_wrap1(z, tab.box_o[2], tab.box_l[2], tab.box_inv[2]);
  }
}

// numpy.around(float32, decimals): multiply, rint, divide in float32 (divide, rint, multiply for
// negative decimals).
HOPB_HD float hopb_around32(float x, const HopbAxis& a) {
  if (a.dec_neg) return HOPB_FMUL(rintf(HOPB_FDIV(x, a.p10)), a.p10);
  return HOPB_FDIV(rintf(HOPB_FMUL(x, a.p10)), a.p10);
}

// _coord2hop index: digitize, then shift one bin left when the rounded coordinate equals the
// rounded top edge (float32 comparison, numpy 1.x casting).  Result in [-1, n+1];
// 1..n are map cells, everything else is an outlier.
HOPB_HD int hopb_bin_hop(float x, const HopbAxis& a, const float* eup) {
  int b = hopb_bin_right(x, a, eup);
  if (hopb_around32(x, a) == a.top_round) b -= 1;
  return b;
}
