// Stage 2: thresholded hydration-site labelling (K2) -- Density.map_sites and friends.
//
// Replaces hop/sitemap.py:518-521 (mask), 1496-1541 (_make_graph/_shell: a networkx graph built
// by a Python loop over cells x 7 forward offsets), 1543-1571 (_label_connected_graphs: connected
// components ranked by size), 1574-1587 (_draw_map_from_sites), 877-918 (site_insert_bulk) and
// 1589-1612 / 601-627 (_annotate_sites).
//
// Connectivity is the reference's 14-neighbourhood: offsets {0,1}^3 \ 0 and their negatives, no
// periodic wrap.  Algorithm: lock-free union-find on the cell grid.
//   1. rows: every z-run of above-threshold cells is pre-linked to its first cell (offset (0,0,1)
//      costs nothing more);
//   2. merge: the other six forward offsets, skipping every union that is implied by the z-runs
//      and by the union one cell further down the run -- for a solid blob only run heads do work;
//   3. flatten: parent = root = minimum C-order index of the component (links always point to the
//      smaller index), which is exactly the tie-break key of the canonical numbering;
//   4. component sizes (warp- and block-aggregated atomics), root list, 64-bit radix sort on
//      (size desc, root desc), labels 1..S, relabel.
// All per-cell passes are HBM streams over G cells (8 B density + 4 B parent).
#include "hopb_common.cuh"
#include "hopb_ccl.cuh"
#include "hopb_tables.cuh"

namespace {

using hopb_ccl::kFull;

// Every per-cell pass below is "one warp per (x, y) row, lanes along z": accesses are contiguous,
// no per-cell division is needed, and rows without a single above-threshold cell (most rows of a
// hydration-site mask) are skipped after one byte read.

// ---- 1. threshold + z-runs ----------------------------------------------------------------------
// The mask comes either from the density (density >= threshold) or straight from the counts: the
// density of a cell is a monotone function of its count for fixed bin widths, and the widths take
// only a few distinct values per axis, so "density >= threshold" is "count >= cmin[width class]" with
// cmin found by bisection over the very function that computes the density (count_threshold_kernel).
// Labelling from the counts takes the density kernel (four double divisions per cell) off the
// critical path of the pipeline.
struct CountMask {
  const uint32_t* counts;
  const unsigned long long* cmin;   // [ncx][ncy][ncz]
  const uint8_t* cls_x;
  const uint8_t* cls_y;
  const uint8_t* cls_z;
  int ny, ncy, ncz;
};

template <bool FROM_COUNTS>
__global__ void __launch_bounds__(256)
ccl_rows_kernel(const double* __restrict__ density, double threshold, CountMask cm, int64_t n_rows, int nz,
                int* __restrict__ parent, uint8_t* __restrict__ rowflag) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    const int64_t base = row * nz;
    const unsigned long long* cmin_row = nullptr;
    if (FROM_COUNTS) {
      const int j = (int)(row % cm.ny);
      const int i = (int)(row / cm.ny);
      cmin_row = cm.cmin + ((int)cm.cls_x[i] * cm.ncy + (int)cm.cls_y[j]) * cm.ncz;
    }
    int carry = -1;  // start (z) of the run that reaches the previous chunk's last cell, or -1
    unsigned any = 0;
    auto in_mask = [&](int z) -> bool {
      if (z >= nz) return false;
      if (FROM_COUNTS) return (unsigned long long)cm.counts[base + z] >= cmin_row[cm.cls_z[z]];
      return density[base + z] >= threshold;
    };
    auto chunk = [&](int z0, bool m) {
      const int z = z0 + lane;
      const unsigned bits = __ballot_sync(kFull, m);
      any |= bits;
      const unsigned below = ~bits & ((1u << lane) - 1);  // non-mask cells below this lane
      int start;
      if (below == 0) start = carry >= 0 ? carry : z0;
      else start = z0 + (32 - __clz(below));
      if (z < nz) parent[base + z] = m ? (int)(base + start) : -1;
      const int s31 = __shfl_sync(kFull, start, 31);
      carry = (bits >> 31) ? s31 : -1;
    };
    // two chunks of 32 cells per step: both loads are in flight before either chunk is linked (a warp
    // that walks its row with one 128-byte request at a time leaves the memory system idle)
    for (int z0 = 0; z0 < nz; z0 += 64) {
      const bool ma = in_mask(z0 + lane);
      const bool mb = in_mask(z0 + 32 + lane);
      chunk(z0, ma);
      if (z0 + 32 < nz) chunk(z0 + 32, mb);
    }
    if (lane == 0) rowflag[row] = any ? 1 : 0;
  }
}

// ---- 2.-4. merge, flatten, sizes: hopb_ccl.cuh -------------------------------------------------------
__global__ void __launch_bounds__(256)
ccl_merge_kernel(int* parent, const uint8_t* __restrict__ rowflag, int nx, int ny, int nz, int phase) {
  hopb_ccl::merge_rows(parent, rowflag, 1u, nx, ny, nz, phase, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5,
                       ((int64_t)gridDim.x * blockDim.x) >> 5, threadIdx.x & 31);
}

__global__ void __launch_bounds__(256)
ccl_flatten_kernel(int* parent, const uint8_t* __restrict__ rowflag, int64_t n_rows, int nz, int heads) {
  hopb_ccl::flatten_rows(parent, rowflag, 1u, n_rows, nz, heads, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5,
                         ((int64_t)gridDim.x * blockDim.x) >> 5, threadIdx.x & 31);
}

__global__ void __launch_bounds__(256)
ccl_size_kernel(const int* __restrict__ parent, const uint8_t* __restrict__ rowflag, int64_t n_rows, int nz,
                int* __restrict__ size) {
  hopb_ccl::size_rows(parent, rowflag, 1u, n_rows, nz, size, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5,
                      ((int64_t)gridDim.x * blockDim.x) >> 5, threadIdx.x & 31);
}

// roots with size >= minsite become sites; smaller components are dropped (label 0)
__global__ void __launch_bounds__(256)
ccl_roots_kernel(const int* __restrict__ parent, const uint8_t* __restrict__ rowflag, int* __restrict__ size,
                 int64_t n_rows, int nz, int64_t n_cells, int minsite, uint64_t* __restrict__ keys,
                 uint32_t* __restrict__ vals, uint32_t cap, uint32_t* __restrict__ n_roots) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    if (!rowflag[row]) continue;
    const int64_t base = row * nz;
    for (int z0 = 0; z0 < nz; z0 += 32) {
      const int z = z0 + lane;
      if (z >= nz) continue;
      const int64_t c = base + z;
      if (parent[c] != (int)c) continue;
      const int sz = size[c];
      if (sz < minsite) { size[c] = 0; continue; }
      const uint32_t slot = atomicAdd(n_roots, 1u);
      if (slot < cap) {
        // ascending sort on this key = size descending, then root index descending
        keys[slot] = ((uint64_t)(uint32_t)(n_cells - sz) << 32) | (uint32_t)(n_cells - 1 - c);
        vals[slot] = (uint32_t)c;
      }
    }
  }
}

// Ranking of the (at most 32768) roots without sort passes: every key is unique, so the rank of a
// key is the number of smaller keys.  The root count is read from device memory, so the whole
// labelling runs without a host round trip.  Keys are staged through shared memory in tiles.
constexpr int kRankTile = 1024;
__global__ void __launch_bounds__(256)
ccl_rank_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ n_dev,
                uint32_t cap, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  __shared__ uint64_t s_keys[kRankTile];
  const uint32_t n_all = *n_dev;
  const int n = (int)(n_all < cap ? n_all : cap);
  if ((int)(blockIdx.x * blockDim.x) >= n) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t k = i < n ? keys[i] : ~0ull;
  int rank = 0;
  for (int t0 = 0; t0 < n; t0 += kRankTile) {
    const int m = n - t0 < kRankTile ? n - t0 : kRankTile;
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x) s_keys[j] = keys[t0 + j];
    __syncthreads();
    for (int j = 0; j < m; ++j) rank += (s_keys[j] < k) ? 1 : 0;
  }
  if (i < n) {
    keys_out[rank] = k;
    vals_out[rank] = vals[i];
  }
}

// label of rank i is i+1; size[root] is reused to hold the label
__global__ void ccl_assign_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                  const uint32_t* __restrict__ n_dev, uint32_t cap, int64_t n_cells,
                                  int* __restrict__ size, int32_t* __restrict__ volumes, int32_t* __restrict__ n_sites_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_all = *n_dev;
  const int n_roots = (int)(n_all < cap ? n_all : cap);
  if (i == 0) {
    if (volumes) volumes[0] = 0;
    if (n_sites_out) *n_sites_out = (int32_t)n_all;  // > HOPB_MAX_SITES signals "too many sites"
  }
  if (i >= n_roots) return;
  const int sz = (int)(n_cells - (int64_t)(keys[i] >> 32));
  size[vals[i]] = i + 1;
  if (volumes) volumes[i + 1] = sz;
}

__global__ void __launch_bounds__(256)
ccl_relabel_kernel(const int* __restrict__ parent, const uint8_t* __restrict__ rowflag, const int* __restrict__ rootlabel,
                   int64_t n_rows, int nz, int32_t* __restrict__ labels) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    const int64_t base = row * nz;
    const bool any = rowflag[row] != 0;
    for (int z0 = 0; z0 < nz; z0 += 64) {            // two chunks per step: loads first, stores last
      const int za = z0 + lane, zb = za + 32;
      const int ra = (any && za < nz) ? parent[base + za] : -1;
      const int rb = (any && zb < nz) ? parent[base + zb] : -1;
      const int la = ra >= 0 ? rootlabel[ra] : 0;
      const int lb = rb >= 0 ? rootlabel[rb] : 0;
      if (za < nz) labels[base + za] = la;
      if (zb < nz) labels[base + zb] = lb;
    }
  }
}

// ---- bulk insertion / map painting ---------------------------------------------------------------------
__global__ void __launch_bounds__(256)
insert_bulk_kernel(int64_t n_cells, int32_t* __restrict__ own, const int32_t* __restrict__ bulk_labels, int32_t bulklabel,
                   uint8_t* __restrict__ member, int16_t* __restrict__ map16) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
    const int32_t o = own[c] > 0 ? own[c] + 1 : 0;
    const bool mb = bulk_labels && bulk_labels[c] == bulklabel;
    own[c] = o;
    member[c] = mb ? 1 : 0;
    map16[c] = (int16_t)(o > 0 ? o : (mb ? 1 : 0));  // later (higher) labels overwrite the bulk site
  }
}

__global__ void __launch_bounds__(256)
to_map16_kernel(int64_t n_cells, const int32_t* __restrict__ labels, int16_t* __restrict__ map16) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x)
    map16[c] = (int16_t)labels[c];
}

// ---- site properties ---------------------------------------------------------------------------------
// acc[label] = {count, sum rho, sum mx, sum my, sum mz} (pass 0) or {sum (rho-mean)^2} (pass 1).
// Large sites (the bulk site, a solid blob) hold most cells: every thread keeps a running partial
// for the label it saw last and for the bulk membership, and blocks reduce those before touching
// the global accumulators, so the number of same-address atomics stays ~#blocks, not #cells.
__device__ __forceinline__ void prop_flush(double* acc, int label, const double* v, int nv) {
  if (label <= 0) return;
  for (int q = 0; q < nv; ++q)
    if (v[q] != 0.0) atomicAdd(acc + (int64_t)label * 5 + q, v[q]);
}

__device__ __forceinline__ void prop_block_flush(double* acc, int label, double* v, int nv, double (*s_red)[5]) {
  // all threads of the block hold partials for the same label
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = 0; q < nv; ++q) {
    double x = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    if (lane == 0) s_red[warp][q] = x;
  }
  __syncthreads();
  if ((int)threadIdx.x < nv) {
    double x = 0;
    for (int w = 0; w < 8; ++w) x += s_red[w][threadIdx.x];
    if (x != 0.0 && label > 0) atomicAdd(acc + (int64_t)label * 5 + threadIdx.x, x);
  }
  __syncthreads();
}

template <int PASS>
__global__ void __launch_bounds__(256)
site_prop_kernel(const double* __restrict__ density, const int32_t* __restrict__ own, const uint8_t* __restrict__ member,
                 int64_t n_cells, int ny, int nz, const double* __restrict__ mx, const double* __restrict__ my,
                 const double* __restrict__ mz, const double* __restrict__ mean, double* __restrict__ acc /*[n_labels][5]*/) {
  __shared__ double s_red[8][5];
  constexpr int NV = PASS == 0 ? 5 : 1;
  double b[5] = {0, 0, 0, 0, 0};  // partial of the bulk membership (label 1)
  double a[5] = {0, 0, 0, 0, 0};  // partial of the run of `cur` labels
  int cur = 0;
  // four cells per thread and step, their loads issued together (the kernel runs with few CTAs per SM beside
  // other work: one dependent load chain per thread left it waiting on memory); the cells are then accumulated
  // in the order c, c + stride, ... as before, so the floating-point sums are unchanged
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c0 < n_cells; c0 += 4 * stride) {
    int o4[4];
    bool mb4[4];
    double rho4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t cu = c0 + u * stride;
      const bool valid = cu < n_cells;
      o4[u] = valid ? own[cu] : 0;
      mb4[u] = valid && member && member[cu];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) rho4[u] = (o4[u] > 0 || mb4[u]) ? density[c0 + u * stride] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
    const int64_t c = c0 + u * stride;
    const int o = o4[u];
    const bool mb = mb4[u];
    if (o <= 0 && !mb) continue;
    const double rho = rho4[u];
    int i = 0, j = 0, k = 0;
    if (PASS == 0) {   // 32-bit index arithmetic: the grid has fewer than 2^31 cells
      const uint32_t cu = (uint32_t)c, r = cu / (uint32_t)nz;
      k = (int)(cu - r * (uint32_t)nz);
      i = (int)(r / (uint32_t)ny);
      j = (int)(r - (uint32_t)i * (uint32_t)ny);
    }
    if (o > 0) {
      if (o != cur) {
        prop_flush(acc, cur, a, NV);
        for (int q = 0; q < 5; ++q) a[q] = 0;
        cur = o;
      }
      if (PASS == 0) { a[0] += 1.0; a[1] += rho; a[2] += mx[i]; a[3] += my[j]; a[4] += mz[k]; }
      else { const double d = rho - mean[o]; a[0] += d * d; }
    }
    if (mb) {
      if (PASS == 0) { b[0] += 1.0; b[1] += rho; b[2] += mx[i]; b[3] += my[j]; b[4] += mz[k]; }
      else { const double d = rho - mean[1]; b[0] += d * d; }
    }
    }
  }
  // own-label partials: one block-level flush when the whole block ended on the same label
  const int first = __shfl_sync(kFull, cur, 0);
  __shared__ int s_first;
  if (threadIdx.x == 0) s_first = first;
  __syncthreads();
  const int uniform = __syncthreads_and(cur == s_first || cur == 0);
  if (uniform) {
    if (cur == 0) for (int q = 0; q < 5; ++q) a[q] = 0;
    prop_block_flush(acc, s_first, a, NV, s_red);
  } else {
    prop_flush(acc, cur, a, NV);
  }
  if (member) prop_block_flush(acc, 1, b, NV, s_red);
}

__global__ void site_prop_mean_kernel(const double* __restrict__ acc, int n_labels, double* __restrict__ mean) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_labels) return;
  const double n = acc[(int64_t)l * 5];
  mean[l] = n > 0 ? acc[(int64_t)l * 5 + 1] / n : 0.0;
}

__global__ void site_prop_final_kernel(const double* __restrict__ acc0, const double* __restrict__ acc1, int n_labels,
                                       int64_t* __restrict__ out_count, double* __restrict__ out_mean,
                                       double* __restrict__ out_std, double* __restrict__ out_center) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_labels) return;
  const double n = acc0[(int64_t)l * 5];
  out_count[l] = (int64_t)n;
  if (n > 0) {
    out_mean[l] = acc0[(int64_t)l * 5 + 1] / n;
    out_std[l] = sqrt(acc1[(int64_t)l * 5] / n);
    for (int q = 0; q < 3; ++q) out_center[l * 3 + q] = acc0[(int64_t)l * 5 + 2 + q] / n;
  } else {
    out_mean[l] = 0; out_std[l] = 0;
    for (int q = 0; q < 3; ++q) out_center[l * 3 + q] = 0;
  }
}

// ---- cell lists ----------------------------------------------------------------------------------------
// entry flags: cell c contributes (own[c] > 0) + (member[c] != 0) entries
__global__ void __launch_bounds__(256)
cells_count_kernel(int64_t n_cells, const int32_t* __restrict__ own, const uint8_t* __restrict__ member, uint32_t* __restrict__ cnt) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x)
    cnt[c] = (own[c] > 0 ? 1u : 0u) + ((member && member[c]) ? 1u : 0u);
}

__global__ void __launch_bounds__(256)
cells_emit_kernel(int64_t n_cells, const int32_t* __restrict__ own, const uint8_t* __restrict__ member,
                  const uint32_t* __restrict__ pos, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
    uint32_t p = pos[c];
    if (member && member[c]) { keys[p] = 1; vals[p] = (uint32_t)c; ++p; }
    if (own[c] > 0) { keys[p] = (uint64_t)own[c]; vals[p] = (uint32_t)c; }
  }
}

__global__ void __launch_bounds__(256)
cells_offsets_kernel(const uint64_t* __restrict__ keys, int64_t n, int n_labels, int64_t* __restrict__ offsets) {
  // offsets[l] = first index with key >= l  (keys sorted ascending)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lo = i == 0 ? 0 : (int64_t)keys[i - 1] + 1;
    const int64_t hi = i == n ? n_labels : (int64_t)keys[i];
    for (int64_t l = lo; l <= hi; ++l) offsets[l] = i;
  }
}

__global__ void __launch_bounds__(256) copy_u32_to_i32_kernel(const uint32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int32_t)in[i];
}

}  // namespace

extern "C" {

static int label_sites_impl(hopb_handle* h, void* stream, int nx, int ny, int nz, const double* density, double threshold,
                            const hopb_grid* g, const uint32_t* counts, int64_t n_frames, double factor, int minsite,
                            int32_t* labels, int32_t* n_sites_dev, int32_t* volumes) {
  HOPB_REQUIRE(h, (density || counts) && labels && n_sites_dev, "null argument");
  HOPB_REQUIRE(h, nx > 0 && ny > 0 && nz > 0, "bad grid shape");
  HOPB_REQUIRE(h, minsite == 1 || minsite == 2, "MINsite must be 1 or 2 (hop/sitemap.py:453-456)");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_cells = (int64_t)nx * ny * nz;
  HOPB_REQUIRE(h, n_cells < ((int64_t)1 << 31), "grid too large");
  // scratch: parent[G], size[G], rowflag[rows] (slot 0); root keys/vals + counter (slot 1)
  const int64_t n_rows = (int64_t)nx * ny;
  int* parent = nullptr;
  int rc = hopb_scratch(h, 0, (size_t)n_cells * 2 * sizeof(int) + (size_t)n_rows + 16, (void**)&parent);
  if (rc) return rc;
  int* size = parent + n_cells;
  uint8_t* rowflag = reinterpret_cast<uint8_t*>(size + n_cells);
  const uint32_t cap = HOPB_MAX_SITES + 1;
  unsigned char* rs = nullptr;
  rc = hopb_scratch(h, 1, (size_t)cap * 2 * (sizeof(uint64_t) + sizeof(uint32_t)) + 64, (void**)&rs);
  if (rc) return rc;
  uint64_t* keys = (uint64_t*)rs;
  uint64_t* keys_alt = keys + cap;
  uint32_t* vals = (uint32_t*)(keys_alt + cap);
  uint32_t* vals_alt = vals + cap;
  uint32_t* n_roots_dev = vals_alt + cap;

  HOPB_CUDA(h, cudaMemsetAsync(size, 0, (size_t)n_cells * sizeof(int), s));
  HOPB_CUDA(h, cudaMemsetAsync(n_roots_dev, 0, sizeof(uint32_t), s));
  const unsigned grows = hopb_grid_for(n_rows * 32, 256, h->num_sms, 8);
  CountMask cm;
  memset(&cm, 0, sizeof(cm));
  if (counts) {
    const int n_classes = g->n_cls[0] * g->n_cls[1] * g->n_cls[2];
    unsigned long long* cmin = nullptr;   // slot 9: the count thresholds
    rc = hopb_scratch(h, 9, (size_t)n_classes * sizeof(unsigned long long), (void**)&cmin);
    if (rc) return rc;
    // the thresholds depend on host-known constants only: computed once and reused while the handle
    // keeps labelling the same grid at the same threshold (every step of a pipeline does).  Work on one
    // handle is ordered (one stream at a time), so the cached table is complete when it is read.
    const bool cached = h->cmin_ptr == (void*)cmin && h->cmin_grid == g->serial && h->cmin_frames == n_frames &&
                        h->cmin_factor == factor && h->cmin_threshold == threshold;   // NaN threshold: never cached
    if (!cached) {
      const double* w = g->d_cls_width;
      HOPB_K(hopb_tables::count_threshold_kernel)<<<(n_classes + 127) / 128, 128, 0, s>>>(n_classes, g->n_cls[1], g->n_cls[2], w, w + 16,
                                                                            w + 32, (double)n_frames, factor,
                                                                            factor != 1.0 ? 1 : 0, threshold, cmin);
      h->cmin_ptr = cmin; h->cmin_grid = g->serial; h->cmin_frames = n_frames; h->cmin_factor = factor; h->cmin_threshold = threshold;
    }
    cm.counts = counts; cm.cmin = cmin; cm.cls_x = g->d_cls[0]; cm.cls_y = g->d_cls[1]; cm.cls_z = g->d_cls[2];
    cm.ny = ny; cm.ncy = g->n_cls[1]; cm.ncz = g->n_cls[2];
    hopb_note_launch();
    ccl_rows_kernel<true><<<grows, 256, 0, s>>>(nullptr, threshold, cm, n_rows, nz, parent, rowflag);
  } else {
    hopb_note_launch();
    ccl_rows_kernel<false><<<grows, 256, 0, s>>>(density, threshold, cm, n_rows, nz, parent, rowflag);
  }
  HOPB_K(ccl_merge_kernel)<<<grows, 256, 0, s>>>(parent, rowflag, nx, ny, nz, 0);
  HOPB_K(ccl_flatten_kernel)<<<grows, 256, 0, s>>>(parent, rowflag, n_rows, nz, 1);
  HOPB_K(ccl_merge_kernel)<<<grows, 256, 0, s>>>(parent, rowflag, nx, ny, nz, 1);
  HOPB_K(ccl_flatten_kernel)<<<grows, 256, 0, s>>>(parent, rowflag, n_rows, nz, 1);
  HOPB_K(ccl_flatten_kernel)<<<grows, 256, 0, s>>>(parent, rowflag, n_rows, nz, 0);
  HOPB_K(ccl_size_kernel)<<<grows, 256, 0, s>>>(parent, rowflag, n_rows, nz, size);
  HOPB_K(ccl_roots_kernel)<<<grows, 256, 0, s>>>(parent, rowflag, size, n_rows, nz, n_cells, minsite, keys, vals, cap, n_roots_dev);
  HOPB_K(ccl_rank_kernel)<<<cap / 256, 256, 0, s>>>(keys, vals, n_roots_dev, cap, keys_alt, vals_alt);
  HOPB_K(ccl_assign_kernel)<<<cap / 256, 256, 0, s>>>(keys_alt, vals_alt, n_roots_dev, cap, n_cells, size, volumes, n_sites_dev);
  HOPB_K(ccl_relabel_kernel)<<<grows, 256, 0, s>>>(parent, rowflag, size, n_rows, nz, labels);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_label_sites_async(hopb_handle* h, void* stream, int nx, int ny, int nz, const double* density,
                           double threshold, int minsite, int32_t* labels, int32_t* n_sites_dev, int32_t* volumes) {
  HOPB_REQUIRE(h, density != nullptr, "null argument");
  return label_sites_impl(h, stream, nx, ny, nz, density, threshold, nullptr, nullptr, 0, 1.0, minsite, labels, n_sites_dev,
                          volumes);
}

int hopb_label_sites_counts_async(hopb_handle* h, void* stream, const hopb_grid* g, const uint32_t* counts,
                                  int64_t n_frames, double factor, double threshold, int minsite, int32_t* labels,
                                  int32_t* n_sites_dev, int32_t* volumes) {
  HOPB_REQUIRE(h, g && counts && n_frames > 0, "bad argument");
  HOPB_REQUIRE(h, g->n_cls[0] > 0 && g->n_cls[1] > 0 && g->n_cls[2] > 0,
               "grid has too many distinct bin widths for count thresholds (use hopb_label_sites_async)");
  return label_sites_impl(h, stream, g->n[0], g->n[1], g->n[2], nullptr, threshold, g, counts, n_frames, factor, minsite,
                          labels, n_sites_dev, volumes);
}

int hopb_label_sites(hopb_handle* h, void* stream, int nx, int ny, int nz, const double* density, double threshold,
                     int minsite, int32_t* labels, int32_t* n_sites, int32_t* volumes) {
  HOPB_REQUIRE(h, n_sites != nullptr, "null argument");
  *n_sites = 0;
  int32_t* n_dev = nullptr;
  int rc = hopb_scratch(h, 8, 64, (void**)&n_dev);
  if (rc) return rc;
  rc = hopb_label_sites_async(h, stream, nx, ny, nz, density, threshold, minsite, labels, n_dev, volumes);
  if (rc) return rc;
  int32_t n = 0;
  HOPB_CUDA(h, cudaMemcpyAsync(&n, n_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  HOPB_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  *n_sites = n;
  if (n > HOPB_MAX_SITES)
    return hopb_fail(h, HOPB_E_TOO_MANY_SITES,
                     "map_sites: %d sites at this threshold; hop's int16 site map holds at most %d", n, HOPB_MAX_SITES);
  return HOPB_OK;
}

int hopb_insert_bulk(hopb_handle* h, void* stream, int64_t n_cells, int32_t* own, const int32_t* bulk_labels,
                     int32_t bulklabel, uint8_t* member, int16_t* map16) {
  HOPB_REQUIRE(h, own && member && map16 && n_cells > 0, "null argument");
  HOPB_K(insert_bulk_kernel)<<<hopb_grid_for(n_cells, 256, h->num_sms, 8), 256, 0, (cudaStream_t)stream>>>(
      n_cells, own, bulk_labels, bulklabel, member, map16);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_labels_to_map16(hopb_handle* h, void* stream, int64_t n_cells, const int32_t* labels, int16_t* map16) {
  HOPB_REQUIRE(h, labels && map16 && n_cells > 0, "null argument");
  HOPB_K(to_map16_kernel)<<<hopb_grid_for(n_cells, 256, h->num_sms, 8), 256, 0, (cudaStream_t)stream>>>(n_cells, labels, map16);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_site_properties(hopb_handle* h, void* stream, const hopb_grid* g, const double* density, const int32_t* own,
                         const uint8_t* member, int n_labels, int64_t* out_count, double* out_mean, double* out_std,
                         double* out_center) {
  HOPB_REQUIRE(h, g && density && own && out_count && out_mean && out_std && out_center, "null argument");
  HOPB_REQUIRE(h, n_labels >= 1 && n_labels <= HOPB_MAX_SITES + 2, "bad n_labels");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_cells = (int64_t)g->n[0] * g->n[1] * g->n[2];
  double* acc = nullptr;
  int rc = hopb_scratch(h, 2, (size_t)n_labels * 11 * sizeof(double), (void**)&acc);
  if (rc) return rc;
  double* acc1 = acc + (size_t)n_labels * 5;
  double* mean = acc1 + (size_t)n_labels * 5;
  HOPB_CUDA(h, cudaMemsetAsync(acc, 0, (size_t)n_labels * 10 * sizeof(double), s));
  // two CTAs per SM: in the pipeline these reductions run on a side stream beside the labelling or the
  // histogram of the next run and must leave room for those kernels' CTAs on every SM
  unsigned gcells = hopb_grid_for(n_cells, 256, h->num_sms, 4);
  if (gcells > 2u * (unsigned)h->num_sms) gcells = 2u * (unsigned)h->num_sms;
  HOPB_K(site_prop_kernel<0>)<<<gcells, 256, 0, s>>>(density, own, member, n_cells, g->n[1], g->n[2], g->d_mid[0], g->d_mid[1],
                                            g->d_mid[2], nullptr, acc);
  HOPB_K(site_prop_mean_kernel)<<<(n_labels + 255) / 256, 256, 0, s>>>(acc, n_labels, mean);
  HOPB_K(site_prop_kernel<1>)<<<gcells, 256, 0, s>>>(density, own, member, n_cells, g->n[1], g->n[2], g->d_mid[0], g->d_mid[1],
                                            g->d_mid[2], mean, acc1);
  HOPB_K(site_prop_final_kernel)<<<(n_labels + 255) / 256, 256, 0, s>>>(acc, acc1, n_labels, out_count, out_mean, out_std, out_center);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_site_cells(hopb_handle* h, void* stream, int64_t n_grid_cells, const int32_t* own, const uint8_t* member,
                    int n_labels, int32_t* cells, int64_t capacity, int64_t* offsets, int64_t* n_cells_out) {
  HOPB_REQUIRE(h, own && offsets && n_cells_out && n_labels >= 1, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t* pos = nullptr;
  int rc = hopb_scratch(h, 3, ((size_t)n_grid_cells + 4) * sizeof(uint32_t), (void**)&pos);
  if (rc) return rc;
  uint32_t* total_dev = pos + n_grid_cells;
  const unsigned gcells = hopb_grid_for(n_grid_cells, 256, h->num_sms, 8);
  HOPB_K(cells_count_kernel)<<<gcells, 256, 0, s>>>(n_grid_cells, own, member, pos);
  rc = hopb_exclusive_scan_u32(h, s, pos, pos, n_grid_cells, total_dev);
  if (rc) return rc;
  uint32_t total = 0;
  HOPB_CUDA(h, cudaMemcpyAsync(&total, total_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  HOPB_CUDA(h, cudaStreamSynchronize(s));
  *n_cells_out = (int64_t)total;
  if ((int64_t)total > capacity || (total > 0 && !cells))
    return hopb_fail(h, HOPB_E_CAPACITY, "hopb_site_cells: need room for %u cells", total);
  unsigned char* kv = nullptr;
  rc = hopb_scratch(h, 4, ((size_t)total + 1) * 2 * (sizeof(uint64_t) + sizeof(uint32_t)), (void**)&kv);
  if (rc) return rc;
  uint64_t* keys = (uint64_t*)kv;
  uint64_t* keys_alt = keys + total + 1;
  uint32_t* vals = (uint32_t*)(keys_alt + total + 1);
  uint32_t* vals_alt = vals + total + 1;
  HOPB_K(cells_emit_kernel)<<<gcells, 256, 0, s>>>(n_grid_cells, own, member, pos, keys, vals);
  int in_alt = 0;
  rc = hopb_radix_sort_pairs(h, s, keys, keys_alt, vals, vals_alt, total, 0, hopb_bits_for((uint64_t)n_labels), &in_alt);
  if (rc) return rc;
  const uint64_t* kf = in_alt ? keys_alt : keys;
  const uint32_t* vf = in_alt ? vals_alt : vals;
  HOPB_K(cells_offsets_kernel)<<<hopb_grid_for((int64_t)total + 1, 256, h->num_sms, 8), 256, 0, s>>>(kf, (int64_t)total, n_labels, offsets);
  if (total > 0)
    HOPB_K(copy_u32_to_i32_kernel)<<<hopb_grid_for((int64_t)total, 256, h->num_sms, 8), 256, 0, s>>>(vf, cells, (int64_t)total);
  HOPB_LAUNCH_CHECK(h);
  HOPB_CUDA(h, cudaStreamSynchronize(s));
  return HOPB_OK;
}

}  // extern "C"
