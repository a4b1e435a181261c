// Host-side construction of the float32 binning tables from float64 edges (see hopb_binning.h).
#pragma once
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "hopb_binning.h"

// smallest float32 >= e
static inline float hopb_up32(double e) {
  float f = (float)e;  // round to nearest
  if ((double)f < e) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
  return f;
}

// Fills `tab` and appends the packed up32 tables of the three axes to `table`.
// Returns 0, or -1 for invalid edges (n < 1, too many edges, not strictly increasing).
// `decimals_override` (may be NULL) supplies the three `decimal` values of hop/trajectory.py:435 as
// computed by the caller's numpy, so that a log10 rounding difference can never matter.
// float32 successor / predecessor
static inline float hopb_next32(float f) { return std::nextafterf(f, std::numeric_limits<float>::infinity()); }
static inline float hopb_prev32(float f) { return std::nextafterf(f, -std::numeric_limits<float>::infinity()); }

// key that orders finite floats like integers (for bisection over float32 values)
static inline int64_t hopb_fkey(float f) {
  int32_t i; std::memcpy(&i, &f, 4);
  return i >= 0 ? (int64_t)i : -(int64_t)(i & 0x7fffffff);
}
static inline float hopb_fromkey(int64_t k) {
  int32_t i = k >= 0 ? (int32_t)k : (int32_t)((-k) | 0x80000000ll);
  float f; std::memcpy(&f, &i, 4);
  return f;
}

static inline int hopb_build_gridtab(const double* const edges[3], const int n[3],
                                     const int* decimals_override, HopbGridTab* tab,
                                     std::vector<float>* table, std::vector<HopbPair>* table2 = nullptr) {
  table->clear();
  if (table2) table2->clear();
  tab->fast_ok = 1;
  tab->wrap = 0;
  for (int d = 0; d < 3; ++d) { tab->box_o[d] = 0.0f; tab->box_l[d] = 0.0f; tab->box_inv[d] = 0.0f; }
  for (int d = 0; d < 3; ++d) {
    const int nb = n[d];
    if (nb < 1 || nb + 1 > HOPB_MAX_EDGES) return -1;
    const double* e = edges[d];
    double dmin = std::numeric_limits<double>::infinity();
    for (int k = 0; k < nb; ++k) {
      const double w = e[k + 1] - e[k];
      if (!(w > 0.0)) return -1;
      if (w < dmin) dmin = w;
    }
    HopbAxis& a = tab->ax[d];
    a.n = nb;
    a.lo = (float)e[0];
    a.inv = (float)((double)nb / (e[nb] - e[0]));
    const float top = (float)e[nb];
    a.top_eq = ((double)top == e[nb]) ? top : std::numeric_limits<float>::quiet_NaN();
    // decimal = int(-log10(dedges.min())) + 6   (hop/trajectory.py:435); int() truncates toward 0
    const int decimals = decimals_override ? decimals_override[d] : (int)(-std::log10(dmin)) + 6;
    const int ad = decimals < 0 ? -decimals : decimals;
    double p = 1.0;
    for (int i = 0; i < ad; ++i) p *= 10.0;
    a.p10 = (float)p;
    a.dec_neg = decimals < 0 ? 1 : 0;
    // numpy.around(float64 scalar, decimals) then the numpy-1.x cast to float32
    const double r = decimals < 0 ? std::rint(e[nb] / p) * p : std::rint(e[nb] * p) / p;
    a.top_round = (float)r;
    a.off = (int)table->size();
    for (int k = 0; k <= nb; ++k) table->push_back(hopb_up32(e[k]));
    const float* eup = table->data() + a.off;
    // pair table for the branch-free search
    a.off2 = table2 ? (int)table2->size() : 0;
    if (table2) {
      const float inf = std::numeric_limits<float>::infinity();
      for (int b = 0; b <= nb + 1; ++b) {
        HopbPair pr;
        pr.lo = b >= 1 ? eup[b - 1] : -inf;
        pr.hi = b <= nb ? eup[b] : inf;
        table2->push_back(pr);
      }
    }
    // one-step correction is exact iff the guess is within +-1 of the truth at both sides of every
    // edge (guess and truth are monotone in x)
    for (int k = 0; k <= nb; ++k) {
      const int g_at = hopb_bin_guess(eup[k], a);             // truth: k + 1
      const int g_below = hopb_bin_guess(hopb_prev32(eup[k]), a);  // truth: k
      if (g_at < k || g_at > k + 2 || g_below < k - 1 || g_below > k + 1) tab->fast_ok = 0;
    }
    // interval of float32 values x with around32(x) == top_round (around32 is monotone)
    {
      const float big = 3.0e38f;
      int64_t lo = hopb_fkey(-big), hi = hopb_fkey(big);
      // first x with around32(x) >= top_round
      int64_t l = lo, h = hi;
      while (l < h) { int64_t m = l + (h - l) / 2; if (hopb_around32(hopb_fromkey(m), a) >= a.top_round) h = m; else l = m + 1; }
      const int64_t first_ge = l;
      l = lo; h = hi;
      // first x with around32(x) > top_round
      while (l < h) { int64_t m = l + (h - l) / 2; if (hopb_around32(hopb_fromkey(m), a) > a.top_round) h = m; else l = m + 1; }
      const int64_t first_gt = l;
      if (first_gt > first_ge && hopb_around32(hopb_fromkey(first_ge), a) == a.top_round) {
        a.hop_lo = hopb_fromkey(first_ge);
        a.hop_hi = hopb_fromkey(first_gt - 1);
      } else {
        a.hop_lo = 1.0f; a.hop_hi = 0.0f;  // empty
      }
    }
  }
  tab->table_len = (int)table->size();
  tab->table2_len = table2 ? (int)table2->size() : 0;
  // Margin for the table-free bin (hopb_bin_sure): the smallest eps for which, on every axis,
  //   (A) g(up32(e_k)) > k - eps   -- points at/above edge k that the guess leaves in bin k have frac > 1 - eps,
  //   (B) g(prev32(up32(e_k))) < k + eps   -- points below edge k that the guess puts in bin k+1 have frac < eps,
  //   (C) g(first coordinate of a last-edge rule) > n - eps   -- those points are never "sure".
  // g is monotone in x, so checking the two float32 neighbours of every edge covers all x.
  tab->sure_eps = 0.0f;
  if (tab->fast_ok) {
    const float candidates[] = {1.0f / 4096, 1.0f / 1024, 1.0f / 256, 1.0f / 64};
    for (float eps : candidates) {
      bool ok = true;
      for (int d = 0; d < 3 && ok; ++d) {
        const HopbAxis& a = tab->ax[d];
        const float* eup = table->data() + a.off;
        auto g = [&](float x) { return (x - a.lo) * a.inv; };
        for (int k = 0; k <= a.n && ok; ++k) {
          if (!(g(eup[k]) > (float)k - eps)) ok = false;
          if (!(g(hopb_prev32(eup[k])) < (float)k + eps)) ok = false;
        }
        float zone = std::numeric_limits<float>::infinity();
        if (a.top_eq == a.top_eq) zone = a.top_eq;
        if (a.hop_lo <= a.hop_hi && a.hop_lo < zone) zone = a.hop_lo;
        if (std::isfinite(zone) && !(g(zone) > (float)a.n - eps)) ok = false;
      }
      if (ok) { tab->sure_eps = eps; break; }
    }
  }
  return 0;
}
