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
static inline int hopb_build_gridtab(const double* const edges[3], const int n[3],
                                     const int* decimals_override, HopbGridTab* tab,
                                     std::vector<float>* table) {
  table->clear();
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
  }
  tab->table_len = (int)table->size();
  return 0;
}
