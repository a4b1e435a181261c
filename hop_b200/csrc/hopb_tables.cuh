// Small per-grid tables shared by label.cu, hist.cu and sitemap_fused.cu: the count thresholds that stand
// for "density >= threshold" and the density values of small counts, both per bin-width class and both
// built by the very function that computes the density (hopb_density_value), so that they are exact.
#pragma once
#include "hopb_common.cuh"

namespace hopb_tables {

// smallest count whose density reaches the threshold, per width class; 2^32 = no count does
static __global__ void count_threshold_kernel(int n_classes, int ncy, int ncz, const double* __restrict__ wx,
                                       const double* __restrict__ wy, const double* __restrict__ wz, double n_frames,
                                       double factor, int apply_factor, double threshold,
                                       unsigned long long* __restrict__ cmin) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_classes) return;
  const double dz = wz[t % ncz], dy = wy[(t / ncz) % ncy], dx = wx[t / (ncz * ncy)];
  auto reaches = [&](unsigned long long c) {
    return hopb_density_value((uint32_t)c, n_frames, dx, dy, dz, factor, apply_factor) >= threshold;
  };
  unsigned long long lo = 0, hi = 1ull << 32;   // answer in [0, 2^32]; NaN threshold: never -> 2^32
  if (!reaches(0xFFFFFFFFull)) { cmin[t] = hi; return; }
  hi = 0xFFFFFFFFull;
  while (lo < hi) {
    const unsigned long long mid = lo + (hi - lo) / 2;
    if (reaches(mid)) hi = mid; else lo = mid + 1;
  }
  cmin[t] = lo;
}

// The same values from a table.  For fixed bin widths the density of a cell is a function of its count
// alone, and a linspace grid has a handful of distinct widths per axis (the width classes of
// hopb_grid_create), so T[class triple][count] built by hopb_density_value itself holds every value a
// cell with count < kDensityTableCounts can take -- bit-identical by construction.  The per-cell work
// drops from four double divisions (~200 instructions of the FP64 pipe) to one 8-byte load from an
// L1/L2-resident table; cells with larger counts take the division path.
constexpr int kDensityTableCounts = 1024;

static __global__ void density_table_kernel(int n_classes, int ncy, int ncz, const double* __restrict__ wx,
                                     const double* __restrict__ wy, const double* __restrict__ wz, double n_frames,
                                     double factor, int apply_factor, double* __restrict__ table) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_classes * kDensityTableCounts) return;
  const int cls = t / kDensityTableCounts, c = t - cls * kDensityTableCounts;
  const double dz = wz[cls % ncz], dy = wy[(cls / ncz) % ncy], dx = wx[cls / (ncz * ncy)];
  table[t] = hopb_density_value((uint32_t)c, n_frames, dx, dy, dz, factor, apply_factor);
}

}  // namespace hopb_tables
