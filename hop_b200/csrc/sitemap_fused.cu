// Stage 2 in ONE launch: histogram counts -> density, hydration-site labels, bulk site, painted int16 map,
// bricked map + block maps for stage 3 and the per-site annotation.
//
// Replaces, for the whole-trajectory pipeline, the chain hopb_density_from_counts ->
// 2 x hopb_label_sites_counts_async -> hopb_insert_bulk -> hopb_brick_map -> hopb_site_properties
// (~30 launches of 2-100 us that each re-walked the parent arrays; profiles/r1_stage2_full.md), i.e.
// DensityCollector.finish + Grid.make_density + convert_density (hop/density.py:133-137,
// hop/sitemap.py:196-216, 236-264), Density.map_sites for the solvent and the bulk threshold
// (hop/sitemap.py:480-521, 1496-1587), site_insert_bulk (hop/sitemap.py:877-918) and _annotate_sites
// (hop/sitemap.py:1589-1612, 601-627).
//
// One cooperative, persistent kernel; grid-wide barriers separate the phases:
//   P0 rows      one read of the counts: both threshold masks (count >= cmin[width class]), z-runs linked,
//                run heads appended to a compact list with their run lengths, density written (table
//                look-up = the reference's four divisions, bit for bit)
//   P1 merge     one lane per run head: the unions that start at a head (hopb_ccl.cuh)
//   P2 flatten   heads find their roots
//   P3 merge     one lane per run head: the unions that end at a head
//   P4 flatten + sizes   heads find their roots and add their run length to the root's
//   P5 roots     hydration sites -> key list; bulk mask -> its largest component (atomicMax)
//   P6 rank      canonical numbering: volume descending, then larger minimum C-order cell first
//   P7a paint    one warp per row: head from the mask bits, root = parent[head] -> own / member / map16 rows and
//                the per-site sums of the annotation
//   P8 annotate  count, mean, population std, centre per site
//   P7b bricks   one warp per eight bricks: map16 (L2) -> bricked map, coarse / fine / 2-bit block maps
// Only P0 and P7 walk the grid; everything between them works on the run heads (about two per row for a
// bulk mask with thermal holes), which is what took the stand-alone chain from ~25 grid passes to two.
// The labels are those of the stand-alone kernels (same union-find, same ranking key).  The
// annotation differs from hopb_site_properties in summation order only: the variance is accumulated
// in one pass around a per-site shift K = density of the site's root cell (exact for one-cell sites,
// benign cancellation otherwise).
#include <cooperative_groups.h>

#include "hopb_ccl.cuh"
#include "hopb_common.cuh"
#include "hopb_tables.cuh"

namespace cg = cooperative_groups;

namespace {

using hopb_ccl::kFull;
using hopb_tables::kDensityTableCounts;

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;   // = the 16 rows of a 4 x 4 brick column
constexpr int kAcc = 6;                 // count, sum d, sum d^2, sum mx, sum my, sum mz   (d = rho - K)

struct FusedArgs {
  const uint32_t* counts;
  int nx, ny, nz;
  const unsigned long long* cmin_s;   // [classes] count thresholds, solvent / bulk
  const unsigned long long* cmin_b;
  const uint8_t* cls_x; const uint8_t* cls_y; const uint8_t* cls_z;
  int ncy, ncz;
  const double* dtable;               // [classes][kDensityTableCounts]
  const double* wx; const double* wy; const double* wz;
  double n_frames, factor;
  int apply_factor;
  int minsite, with_bulk;
  // scratch
  int* parent_s; int* parent_b; int* size_s; int* size_b;
  uint8_t* rowflag;
  uint64_t* keys; uint32_t* vals; uint32_t cap;
  uint32_t* counters;                 // [0] hydration roots, [1] bulk roots, [2..3] best bulk key (u64), [4] / [5] run heads, [6] root of the bulk site
  int* heads_s; int* heads_b;         // run heads of the two masks (cell indices)
  int* minidx_s; int* minidx_b;       // at roots: the minimum cell of the component (the canonical tie-break key)
  uint32_t* mask_s; uint32_t* mask_b; // the masks as bits, [rows][wpr] words (bit z & 31 of word z >> 5)
  int wpr;
  double* acc; double* kshift; int n_labels_cap;
  // outputs
  double* density;
  int32_t* own; uint8_t* member; int16_t* map16;
  int16_t* brick; uint32_t* coarse; int coarse_words; uint8_t* fine; int64_t nb_pad; uint32_t* cell2;
  int32_t* n_sites_dev;               // [2]: hydration sites, components of the bulk mask
  int32_t* volumes;                   // may be NULL
  const double* mx; const double* my; const double* mz;
  int64_t* p_count; double* p_mean; double* p_std; double* p_center;   // may be NULL (no annotation)
  unsigned long long* trace;          // [16] phase time stamps (ns), debugging aid (HOPB_FUSED_TRACE=1), else NULL
};

// Per-warp staging of run heads in shared memory: the global list is reserved in batches, one atomic per
// ~100 heads instead of one per chunk.
constexpr int kStage = 128;
struct HeadStage {
  int* buf;          // [kStage] shared
  int n;             // staged (warp-uniform)
  int* list;         // global list
  uint32_t* count;   // its fill counter
};

__device__ __forceinline__ void stage_flush(HeadStage& st, int lane) {
  if (st.n == 0) return;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(st.count, (uint32_t)st.n);
  base = __shfl_sync(kFull, base, 0);
  __syncwarp();
  for (int i = lane; i < st.n; i += 32) st.list[base + i] = st.buf[i];
  __syncwarp();
  st.n = 0;
}

// z-run linking of one 32-cell chunk of one mask (see ccl_rows_kernel); `bits` is the ballot of the mask over the
// chunk.  size[head] receives the length of the run when the run ends (one plain store per run: the running
// length of the run that reaches the end of a chunk travels in rs.pend).  The two common chunks -- nothing in the
// mask (hydration sites are sparse), everything in the mask and the run goes on (solid bulk) -- take a few
// instructions.
struct RunState { int carry, pend; };   // z of the head of the run reaching the previous chunk's last cell (or -1), its length so far

__device__ __forceinline__ void link_chunk(unsigned bits, int z0, int lane, int64_t base, int* __restrict__ parent,
                                           int* __restrict__ size, int* __restrict__ minidx, uint32_t* __restrict__ mask_word,
                                           RunState& rs, HeadStage& st) {
  if (lane == 0) *mask_word = bits;
  if (bits == 0u) {
    if (rs.carry >= 0) {
      if (lane == 0) size[base + rs.carry] = rs.pend;
      rs.carry = -1;
    }
    return;
  }
  const int z = z0 + lane;
  if (bits == kFull && rs.carry >= 0) {        // the run goes on: nothing to record but its growing length
    rs.pend += 32;
    return;
  }
  const bool m = (bits >> lane) & 1u;
  const unsigned below = ~bits & ((1u << lane) - 1);
  int start;
  if (below == 0) start = rs.carry >= 0 ? rs.carry : z0;
  else start = z0 + (32 - __clz(below));
  bool head = m && start == z;
  if (head) parent[base + z] = (int)(base + z);       // only run heads carry union-find state (every other cell: mask bits -> its head)
  if (head) minidx[base + z] = (int)(base + z);
  const unsigned hbits = __ballot_sync(kFull, head);
  if (hbits) {
    const int nh = __popc(hbits);
    if (st.n + nh > kStage) stage_flush(st, lane);
    if (head) st.buf[st.n + __popc(hbits & ((1u << lane) - 1))] = (int)(base + z);
    st.n += nh;
  }
  // the run that came in from the previous chunk ends here (bits is not all ones on this path)
  if (rs.carry >= 0 && lane == 0) size[base + rs.carry] = rs.pend + (__ffs(~bits) - 1);
  // runs that begin and end inside the chunk
  if (m && lane < 31 && !((bits >> (lane + 1)) & 1u) && start >= z0) size[base + start] = z - start + 1;
  // the run that reaches the last cell goes on
  if (bits >> 31) {
    const int s31 = __shfl_sync(kFull, start, 31);
    rs.pend = s31 >= z0 ? z0 + 32 - s31 : rs.pend + 32;
    rs.carry = s31;
  } else {
    rs.carry = -1;
  }
}

// bit s of the result is set when nibble s of x is zero
__device__ __forceinline__ uint32_t zero_nibbles(uint32_t x) {
  x |= x >> 1;
  x |= x >> 2;
  uint32_t f = ~x & 0x11111111u;
  f = (f | (f >> 3)) & 0x03030303u;
  f = (f | (f >> 6)) & 0x000F000Fu;
  return (f | (f >> 12)) & 0xFFu;
}

struct MaskBits {
  const uint32_t* w;
  int nx, ny, nz, wpr;
  // false outside the grid, without a branch (the address is clamped, the answer masked)
  __device__ __forceinline__ bool operator()(int x, int y, int z) const {
    const bool in = ((unsigned)x < (unsigned)nx) & ((unsigned)y < (unsigned)ny) & ((unsigned)z < (unsigned)nz);
    const int xs = in ? x : 0, ys = in ? y : 0, zs = in ? z : 0;
    return in & (bool)((w[((int64_t)xs * ny + ys) * wpr + (zs >> 5)] >> (zs & 31)) & 1u);
  }
  // head of the z-run that the mask cell (x, y, z) belongs to: the cell after the nearest cell below it that is
  // outside the mask (the row start if there is none); a solid 201-cell row is at most seven word loads, L1-resident
  __device__ __forceinline__ int head(int x, int y, int z) const {
    const uint32_t* row = w + ((int64_t)x * ny + y) * wpr;
    int wi = z >> 5;
    uint32_t inv = ~row[wi] & ((1u << (z & 31)) - 1u);
    while (inv == 0u && wi > 0) inv = ~row[--wi];
    const int hz = inv ? (wi << 5) + (32 - __clz(inv)) : 0;
    return (int)(((int64_t)x * ny + y) * nz + hz);
  }
};

__device__ __forceinline__ void stamp(const FusedArgs& a, int k) {
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[k] = t;
  }
}

__global__ void __launch_bounds__(kThreads, 2) sitemap_fused_kernel(FusedArgs a) {
  cg::grid_group grid = cg::this_grid();
  stamp(a, 0);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * kThreads;
  const int nx = a.nx, ny = a.ny, nz = a.nz;
  const int64_t n_rows = (int64_t)nx * ny;
  const int64_t n_cells = n_rows * nz;
  const bool bulk = a.with_bulk != 0;

  // ---- P0: zero the small accumulators; rows ---------------------------------------------------------
  // (counters[4..5], the fill of the head lists, are zeroed by the launcher: they are in use from the start)
  __shared__ int s_stage[2][kWarps][kStage];
  HeadStage st_s{s_stage[0][warp], 0, a.heads_s, a.counters + 4};
  HeadStage st_b{s_stage[1][warp], 0, a.heads_b, a.counters + 5};
  for (int64_t i = tid; i < (int64_t)a.n_labels_cap * kAcc; i += nth) a.acc[i] = 0.0;
  for (int64_t i = tid; i < a.coarse_words; i += nth) a.coarse[i] = 0u;
  if (tid < 4) a.counters[tid] = 0u;
  if (tid == 0) a.counters[6] = 0xFFFFFFFFu;
  if (tid == 0 && a.fine) *reinterpret_cast<uint32_t*>(a.fine + 2 * a.nb_pad) = 0u;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    const int j = (int)(row % ny);
    const int i = (int)(row / ny);
    const int64_t base = row * nz;
    const int cxy = (int)a.cls_x[i] * a.ncy + (int)a.cls_y[j];
    const unsigned long long* cs_row = a.cmin_s + (int64_t)cxy * a.ncz;
    const unsigned long long* cb_row = a.cmin_b + (int64_t)cxy * a.ncz;
    const double* trow = a.dtable + (int64_t)cxy * a.ncz * kDensityTableCounts;
    RunState rs_s{-1, 0}, rs_b{-1, 0};
    unsigned any_s = 0, any_b = 0;
    uint32_t* mrow_s = a.mask_s + row * a.wpr;
    uint32_t* mrow_b = a.mask_b + row * a.wpr;
    constexpr int U = 4;     // chunks per step, loads first: a warp with one request in flight leaves the memory system idle
    for (int z0 = 0; z0 < nz; z0 += 32 * U) {
      uint32_t c[U];
      int cz[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int z = z0 + 32 * u + lane;
        c[u] = z < nz ? a.counts[base + z] : 0u;
        cz[u] = z < nz ? (int)a.cls_z[z] : 0;
      }
      // density: table look-up (measured as a filler of the merge phases instead: the streaming traffic slows
      // their dependent L2 round trips several-fold, 65 -> 435 us; it stays with the other streaming work)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int z = z0 + 32 * u + lane;
        if (z >= nz) continue;
        double v;
        if (c[u] < (uint32_t)kDensityTableCounts) v = __ldg(trow + cz[u] * kDensityTableCounts + (int)c[u]);
        else v = hopb_density_value(c[u], a.n_frames, a.wx[i], a.wy[j], a.wz[z], a.factor, a.apply_factor);
        a.density[base + z] = v;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int zc = z0 + 32 * u;
        if (zc >= nz) break;
        const bool ok = zc + lane < nz;
        const unsigned bs = __ballot_sync(kFull, ok && (unsigned long long)c[u] >= cs_row[cz[u]]);
        any_s |= bs;
        link_chunk(bs, zc, lane, base, a.parent_s, a.size_s, a.minidx_s, mrow_s + (zc >> 5), rs_s, st_s);
        if (bulk) {
          const unsigned bb = __ballot_sync(kFull, ok && (unsigned long long)c[u] >= cb_row[cz[u]]);
          any_b |= bb;
          link_chunk(bb, zc, lane, base, a.parent_b, a.size_b, a.minidx_b, mrow_b + (zc >> 5), rs_b, st_b);
        }
      }
    }
    if (lane == 0) {
      if (rs_s.carry >= 0) a.size_s[base + rs_s.carry] = rs_s.pend;
      if (rs_b.carry >= 0) a.size_b[base + rs_b.carry] = rs_b.pend;
    }
    if (lane == 0) a.rowflag[row] = (uint8_t)((any_s ? 1 : 0) | (any_b ? 2 : 0));
  }
  stage_flush(st_s, lane);
  stage_flush(st_b, lane);
  grid.sync();
  stamp(a, 1);

  // ---- P1 - P4: union-find over the run heads --------------------------------------------------------------
  const int64_t nh_s = a.counters[4], nh_b = bulk ? a.counters[5] : 0;
  const MaskBits Ms{a.mask_s, nx, ny, nz, a.wpr}, Mb{a.mask_b, nx, ny, nz, a.wpr};
  const hopb_ccl::UnionHashed UF;
  auto list_heads = [&](const int* list, int64_t nh, auto&& f) {
    for (int64_t i = tid; i < nh; i += nth) f(list[i]);
  };
  list_heads(a.heads_s, nh_s, [&](int hd) { hopb_ccl::merge_head_forward(a.parent_s, hd, nx, ny, nz, Ms, UF); });
  list_heads(a.heads_b, nh_b, [&](int hd) { hopb_ccl::merge_head_forward(a.parent_b, hd, nx, ny, nz, Mb, UF); });
  grid.sync();
  stamp(a, 2);
  {
    auto flat_s = [&](int hd) { a.parent_s[hd] = hopb_ccl::uf_find_h(a.parent_s, hd); };
    auto flat_b = [&](int hd) { a.parent_b[hd] = hopb_ccl::uf_find_h(a.parent_b, hd); };
    list_heads(a.heads_s, nh_s, flat_s);
    if (bulk) { list_heads(a.heads_b, nh_b, flat_b); }
  }
  grid.sync();
  stamp(a, 3);
  {
    auto back_s = [&](int hd) { hopb_ccl::merge_head_backward(a.parent_s, hd, nx, ny, nz, Ms, UF); };
    auto back_b = [&](int hd) { hopb_ccl::merge_head_backward(a.parent_b, hd, nx, ny, nz, Mb, UF); };
    list_heads(a.heads_s, nh_s, back_s);
    if (bulk) { list_heads(a.heads_b, nh_b, back_b); }
  }
  grid.sync();
  stamp(a, 4);
  // heads find their final roots; a head that is not a root hands its run length to the root (only roots
  // receive additions, and a root's own entry already holds its own run).  One giant component would make
  // that hundreds of thousands of atomics on one address, so every thread keeps a running (root, sum) and the
  // lanes of a warp that ended on the same root add once.
  auto sizes = [&](int* parent, int* size, int* minidx, const int* heads, int64_t nh) {
    int run_root = -1, run_sum = 0, run_min = 0x7fffffff;
    auto one = [&](int hd) {
      const int r = hopb_ccl::uf_find_h(parent, hd);
      parent[hd] = r;
      if (r == hd) return;
      const int len = size[hd];
      if (r == run_root) { run_sum += len; run_min = hd < run_min ? hd : run_min; }
      else {
        if (run_root >= 0) { atomicAdd(size + run_root, run_sum); atomicMin(minidx + run_root, run_min); }
        run_root = r; run_sum = len; run_min = hd;
      }
    };
    list_heads(heads, nh, one);
    const unsigned peers = __match_any_sync(kFull, run_root);
    int tot = 0, mn = 0x7fffffff;
    for (unsigned m = peers; m; m &= m - 1) {
      tot += __shfl_sync(peers, run_sum, __ffs(m) - 1);
      const int o = __shfl_sync(peers, run_min, __ffs(m) - 1);
      mn = o < mn ? o : mn;
    }
    if (run_root >= 0 && (peers & ((1u << lane) - 1)) == 0) { atomicAdd(size + run_root, tot); atomicMin(minidx + run_root, mn); }
  };
  sizes(a.parent_s, a.size_s, a.minidx_s, a.heads_s, nh_s);
  if (bulk) sizes(a.parent_b, a.size_b, a.minidx_b, a.heads_b, nh_b);
  grid.sync();
  stamp(a, 5);

  // ---- P5: roots -----------------------------------------------------------------------------------------
  unsigned long long* best_bulk = reinterpret_cast<unsigned long long*>(a.counters + 2);
  {
    auto root_s = [&](int c) {
      if (a.parent_s[c] != c) return;
      const int sz = a.size_s[c];
      if (sz < a.minsite) { a.size_s[c] = 0; return; }
      const uint32_t slot = atomicAdd(a.counters + 0, 1u);
      if (slot < a.cap) {
        a.keys[slot] = ((uint64_t)(uint32_t)(n_cells - sz) << 32) | (uint32_t)(n_cells - 1 - a.minidx_s[c]);
        a.vals[slot] = (uint32_t)c;
      }
    };
    auto root_b = [&](int c) {
      if (a.parent_b[c] != c) return;
      const int sz = a.size_b[c];
      if (sz < a.minsite) return;
      atomicAdd(a.counters + 1, 1u);
      // label 1 of the bulk map: the largest component, ties to the larger minimum cell
      atomicMax(best_bulk, ((unsigned long long)(uint32_t)sz << 32) | (uint32_t)a.minidx_b[c]);
    };
    list_heads(a.heads_s, nh_s, root_s);
    if (bulk) { list_heads(a.heads_b, nh_b, root_b); }
  }
  grid.sync();
  stamp(a, 6);

  // ---- P6: rank the hydration sites (every key is unique: rank = number of smaller keys) ---------------------
  const uint32_t n_all = a.counters[0];
  const bool too_many = n_all > a.cap - 1;          // more than HOPB_MAX_SITES: the caller raises
  const int n_roots = too_many ? 0 : (int)n_all;
  const int shift = bulk ? 1 : 0;                   // site_insert_bulk moves every hydration site up by one label
  {
    // one warp per root: the lanes share the comparisons (keys stay in L1 / L2)
    for (int64_t i = warp0; i < n_roots; i += nwarps) {
      const uint64_t k = a.keys[i];
      int rank = 0;
      for (int q = lane; q < n_roots; q += 32) rank += (a.keys[q] < k) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(kFull, rank, o);
      if (lane == 0) {
        const uint32_t root = a.vals[i];
        a.size_s[root] = rank + 1;
        if (a.volumes) a.volumes[rank + 1] = (int32_t)(n_cells - (int64_t)(k >> 32));
        if (a.p_count && rank + 1 + shift < a.n_labels_cap) a.kshift[rank + 1 + shift] = a.density[root];
      }
    }
    if (bulk && *best_bulk) {     // which root is it: the one whose (size, minimum cell) won
      const unsigned long long bb = *best_bulk;
      auto match = [&](int c) {
        if (a.parent_b[c] == c && (((unsigned long long)(uint32_t)a.size_b[c] << 32) | (uint32_t)a.minidx_b[c]) == bb) {
          a.counters[6] = (uint32_t)c;
          a.kshift[1] = a.density[c];
        }
      };
      list_heads(a.heads_b, nh_b, match);
    }
    if (tid == 0) {
      a.n_sites_dev[0] = (int32_t)n_all;
      a.n_sites_dev[1] = (int32_t)a.counters[1];
      if (a.volumes) a.volumes[0] = 0;
      a.kshift[0] = 0.0;
      if (!(bulk && *best_bulk)) a.kshift[1] = 0.0;
    }
  }
  grid.sync();
  stamp(a, 7);

  // ---- P7: paint -------------------------------------------------------------------------------------------
  // P7a: one warp per row, no block barriers (the earlier version staged a 4 x 4 brick column in shared memory
  // between two __syncthreads; a third of the kernel's stall samples sat at those barriers, waiting for the slowest
  // row of the column): own / member / map16 rows and the annotation sums.
  // P7b (behind a grid barrier): one warp per 8 consecutive bricks of a column re-reads their 16 x 32 map16 cells
  // from L2 and writes the bricked map and the block maps.
  const int bulk_root = bulk ? (int)a.counters[6] : -1;      // 0xFFFFFFFF = -1: no bulk site
  const int cnx = (nx + 3) >> 2, cny = (ny + 3) >> 2, cnz = (nz + 3) >> 2;
  // running sums of the bulk site (label 1), per lane
  double b_n = 0, b_d = 0, b_d2 = 0, b_mx = 0, b_my = 0, b_mz = 0;
  const bool props = a.p_count != nullptr;
  const double k_bulk = (bulk && props) ? a.kshift[1] : 0.0;
  uint32_t n_ones = 0;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    const int j = (int)(row % ny);
    const int i = (int)(row / ny);
    const int64_t base = row * nz;
    const unsigned fl = a.rowflag[row];
    if (fl == 0u) {          // nothing of either mask in this row
      for (int z = lane; z < nz; z += 32) {
        a.own[base + z] = 0;
        if (a.member) a.member[base + z] = 0;
        a.map16[base + z] = 0;
      }
      continue;
    }
    const double mxi = props ? a.mx[i] : 0.0, myj = props ? a.my[j] : 0.0;
    const uint32_t* mrow_s = a.mask_s + row * a.wpr;
    const uint32_t* mrow_b = a.mask_b + row * a.wpr;
    constexpr int U = 4;   // chunks per step: every level of the dependent look-ups is issued for all of them at once
    int carry_s = -1, carry_b = -1;
    for (int z0 = 0; z0 < nz; z0 += 32 * U) {
      int hs[U], hb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // the head of a cell's run from the mask word of its chunk (and the run carried in from the chunk below):
        // no per-cell parent entry is read -- or was ever written
        const int zc = z0 + 32 * u;
        const uint32_t ws = (zc < nz && (fl & 1u)) ? mrow_s[zc >> 5] : 0u;
        const uint32_t wb = (zc < nz && (fl & 2u)) ? mrow_b[zc >> 5] : 0u;
        const uint32_t lt = (1u << lane) - 1u;
        const uint32_t bs = ~ws & lt, bb = ~wb & lt;
        const int st_s = bs ? zc + (32 - __clz(bs)) : (carry_s >= 0 ? carry_s : zc);
        const int st_b = bb ? zc + (32 - __clz(bb)) : (carry_b >= 0 ? carry_b : zc);
        hs[u] = ((ws >> lane) & 1u) ? (int)(base + st_s) : -1;
        hb[u] = ((wb >> lane) & 1u) ? (int)(base + st_b) : -1;
        const uint32_t b31s = ~ws & 0x7FFFFFFFu, b31b = ~wb & 0x7FFFFFFFu;      // the run that reaches the chunk's last cell
        carry_s = (ws >> 31) ? (b31s ? zc + (32 - __clz(b31s)) : (carry_s >= 0 ? carry_s : zc)) : -1;
        carry_b = (wb >> 31) ? (b31b ? zc + (32 - __clz(b31b)) : (carry_b >= 0 ? carry_b : zc)) : -1;
      }
      int ps[U], pb[U];
      double rho[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ps[u] = hs[u] >= 0 ? a.parent_s[hs[u]] : -1;          // heads point at their roots since P4
        pb[u] = hb[u] >= 0 ? a.parent_b[hb[u]] : -1;
        // the density of every cell of a mask (a superset of the cells the annotation needs)
        rho[u] = (props && (hs[u] >= 0 || hb[u] >= 0)) ? a.density[base + z0 + 32 * u + lane] : 0.0;
      }
      int ls[U];
#pragma unroll
      for (int u = 0; u < U; ++u) ls[u] = (ps[u] >= 0 && !too_many) ? a.size_s[ps[u]] : 0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int z = z0 + 32 * u + lane;
        if (z0 + 32 * u >= nz) break;
        if (z >= nz) continue;
        const int o = ls[u] > 0 ? ls[u] + shift : 0;
        const bool mb = pb[u] >= 0 && pb[u] == bulk_root;
        a.own[base + z] = o;
        if (a.member) a.member[base + z] = mb ? 1 : 0;
        a.map16[base + z] = (int16_t)(o > 0 ? o : (mb ? 1 : 0));
        if (props && (o > 0 || mb)) {
          const double mzk = a.mz[z];
          if (o > 0 && o < a.n_labels_cap) {
            const double d = rho[u] - a.kshift[o];
            double* q = a.acc + (int64_t)o * kAcc;
            atomicAdd(q + 0, 1.0); atomicAdd(q + 1, d); atomicAdd(q + 2, d * d);
            atomicAdd(q + 3, mxi); atomicAdd(q + 4, myj); atomicAdd(q + 5, mzk);
          }
          if (mb) {
            const double d = rho[u] - k_bulk;
            b_n += 1.0; b_d += d; b_d2 += d * d; b_mx += mxi; b_my += myj; b_mz += mzk;
          }
        }
      }
    }
  }
  if (props && bulk) {
    double sb[kAcc] = {b_n, b_d, b_d2, b_mx, b_my, b_mz};
#pragma unroll
    for (int q = 0; q < kAcc; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sb[q] += __shfl_xor_sync(kFull, sb[q], o);
      if (lane == 0 && sb[0] != 0.0) atomicAdd(a.acc + 1 * kAcc + q, sb[q]);
    }
  }
  grid.sync();
  stamp(a, 8);

  // ---- P8: annotation (the sums are complete; it runs in front of the bricks instead of behind another barrier) ----
  for (int64_t l = tid; props && l < a.n_labels_cap; l += nth) {
    const double* q = a.acc + l * kAcc;
    const double n = q[0];
    a.p_count[l] = (int64_t)n;
    if (n > 0) {
      const double md = q[1] / n;
      double var = (q[2] - q[1] * md) / n;
      if (!(var > 0.0)) var = 0.0;
      a.p_mean[l] = a.kshift[l] + md;
      a.p_std[l] = sqrt(var);
      a.p_center[l * 3 + 0] = q[3] / n; a.p_center[l * 3 + 1] = q[4] / n; a.p_center[l * 3 + 2] = q[5] / n;
    } else {
      a.p_mean[l] = 0; a.p_std[l] = 0;
      a.p_center[l * 3 + 0] = 0; a.p_center[l * 3 + 1] = 0; a.p_center[l * 3 + 2] = 0;
    }
  }

  // P7b: bricks.  Lane l holds the cells 2l, 2l + 1 of the brick order; the 16 rows x 32 cells of eight bricks go
  // through a per-warp tile (rows of a column are nz cells apart in map16: 64-byte row segments, read once).
  {
    __shared__ int16_t s_tile[kWarps][16 * 32];
    int16_t* wt = s_tile[warp];
    const int n_slabs = (cnz + 7) >> 3;
    const int64_t n_tasks = (int64_t)cnx * cny * n_slabs;
    const int i1 = (lane >> 4) & 1, j1 = (lane >> 3) & 1, k1 = (lane >> 2) & 1, i0 = (lane >> 1) & 1, j0 = lane & 1;
    const int di = 2 * i1 + i0, dj = 2 * j1 + j0;
    for (int64_t task = warp0; task < n_tasks; task += nwarps) {
      const int sl = (int)(task % n_slabs);
      const int col = (int)(task / n_slabs);
      const int ci = col / cny, cj = col - ci * cny;
      const int zl = sl * 32 + lane;
      int16_t v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int i = ci * 4 + (r >> 2), j = cj * 4 + (r & 3);
        v[r] = (i < nx && j < ny && zl < nz) ? a.map16[((int64_t)i * ny + j) * nz + zl] : (int16_t)0;
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) wt[r * 32 + lane] = v[r];
      __syncwarp();
      const bool row_ok = (ci * 4 + di < nx) && (cj * 4 + dj < ny);
      const int16_t* trow = wt + (di * 4 + dj) * 32;
      for (int q = 0; q < 8; ++q) {
        const int ck = sl * 8 + q;
        if (ck >= cnz) break;
        const int64_t b = ((int64_t)ci * cny + cj) * cnz + ck;
        const int z = ck * 4 + 2 * k1;
        const int v0 = trow[q * 4 + 2 * k1], v1 = trow[q * 4 + 2 * k1 + 1];
        const bool ok0 = row_ok && z < nz, ok1 = row_ok && z + 1 < nz;
        reinterpret_cast<short2*>(a.brick + b * 64)[lane] = make_short2((short)v0, (short)v1);
        const int ref = __shfl_sync(kFull, v0, 0);      // cell (0, 0, 0) of a brick is always inside the grid
        const bool uni = __all_sync(kFull, (!ok0 || v0 == ref) && (!ok1 || v1 == ref));
        const unsigned not0 = __ballot_sync(kFull, (ok0 && v0 != 0) || (ok1 && v1 != 0));
        const unsigned not1 = __ballot_sync(kFull, (ok0 && v0 != 1) || (ok1 && v1 != 1));
        n_ones += ((ok0 && v0 == 1) ? 1u : 0u) + ((ok1 && v1 == 1) ? 1u : 0u);
        if (a.cell2) {
          const uint32_t c0 = v0 == 0 ? 0u : (v0 == 1 ? 1u : 3u), c1 = v1 == 0 ? 0u : (v1 == 1 ? 1u : 3u);
          uint32_t w = (c0 | (c1 << 2)) << (4 * (lane & 7));
          w |= __shfl_xor_sync(kFull, w, 1);
          w |= __shfl_xor_sync(kFull, w, 2);
          w |= __shfl_xor_sync(kFull, w, 4);
          if ((lane & 7) == 0) a.cell2[b * 4 + (lane >> 3)] = w;
        }
        if (lane == 0) {
          const uint32_t st = (uni && (ref == 0 || ref == 1)) ? (uint32_t)ref : 3u;
          if (st) atomicOr(a.coarse + (b >> 4), st << ((b & 15) * 2));
          if (a.fine) {
            // bit s: all cells of sub-brick s (= lanes 4s .. 4s + 3) are 0 / are 1
            a.fine[b] = (uint8_t)zero_nibbles(not0);
            a.fine[a.nb_pad + b] = (uint8_t)zero_nibbles(not1);
          }
        }
      }
    }
  }
  if (a.fine) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_ones += __shfl_xor_sync(kFull, n_ones, o);
    if (lane == 0 && n_ones) atomicAdd(reinterpret_cast<uint32_t*>(a.fine + 2 * a.nb_pad), n_ones);
  }
  stamp(a, 9);
}

}  // namespace

extern "C" {

int hopb_site_map_fused(hopb_handle* h, void* stream, const hopb_grid* g, const uint32_t* counts, int64_t n_frames,
                        double factor, double solvent_threshold, double bulk_threshold, int with_bulk, int minsite,
                        double* density, int32_t* own, uint8_t* member, int16_t* map16, int16_t* brick, uint32_t* coarse,
                        uint8_t* fine, uint32_t* cell2, int32_t* n_sites_dev, int32_t* volumes, int n_labels_cap,
                        int64_t* p_count, double* p_mean, double* p_std, double* p_center) {
  HOPB_REQUIRE(h, g && counts && density && own && map16 && brick && coarse && n_sites_dev && n_frames > 0, "null argument");
  HOPB_REQUIRE(h, !with_bulk || member, "member buffer missing");
  HOPB_REQUIRE(h, minsite == 1 || minsite == 2, "MINsite must be 1 or 2 (hop/sitemap.py:453-456)");
  const int n_classes = (g->n_cls[0] > 0 && g->n_cls[1] > 0 && g->n_cls[2] > 0) ? g->n_cls[0] * g->n_cls[1] * g->n_cls[2] : 0;
  HOPB_REQUIRE(h, n_classes > 0 && n_classes <= 512, "grid has too many distinct bin widths for the one-launch site map");
  HOPB_REQUIRE(h, !p_count || (p_mean && p_std && p_center && n_labels_cap >= 2 && n_labels_cap <= HOPB_MAX_SITES + 2),
               "bad annotation buffers");
  cudaStream_t s = (cudaStream_t)stream;
  const int nx = g->n[0], ny = g->n[1], nz = g->n[2];
  const int64_t n_rows = (int64_t)nx * ny, n_cells = n_rows * nz;
  const int cnx = (nx + 3) / 4, cny = (ny + 3) / 4, cnz = (nz + 3) / 4;
  const int64_t nb = (int64_t)cnx * cny * cnz;
  HOPB_REQUIRE(h, nb * 64 < ((int64_t)1 << 32) - 1, "grid too large for brick indices");
  const size_t smem = 0;      // static shared memory only (head staging, the per-warp brick tiles)
  (void)cnz;
  if (!p_count) n_labels_cap = 2;

  // tables: count thresholds of both masks and the density of small counts, cached per handle
  unsigned long long* cmin = nullptr;
  double* dtable = nullptr;
  int rc = hopb_scratch(h, 11, (size_t)n_classes * 2 * sizeof(unsigned long long), (void**)&cmin);
  if (rc) return rc;
  rc = hopb_scratch(h, 12, (size_t)n_classes * kDensityTableCounts * sizeof(double), (void**)&dtable);
  if (rc) return rc;
  const double key[5] = {(double)g->serial, (double)n_frames, factor, solvent_threshold, bulk_threshold};
  const bool cached = h->fused_cmin == (void*)cmin && h->fused_dtable == (void*)dtable && memcmp(h->fused_key, key, sizeof(key)) == 0;
  if (!cached) {   // NaN thresholds compare unequal bit-wise only if the NaNs differ; memcmp keeps identical NaNs cached (same result)
    const double* w = g->d_cls_width;
    const int af = factor != 1.0 ? 1 : 0;
    HOPB_K(hopb_tables::count_threshold_kernel)<<<(n_classes + 127) / 128, 128, 0, s>>>(
        n_classes, g->n_cls[1], g->n_cls[2], w, w + 16, w + 32, (double)n_frames, factor, af, solvent_threshold, cmin);
    HOPB_K(hopb_tables::count_threshold_kernel)<<<(n_classes + 127) / 128, 128, 0, s>>>(
        n_classes, g->n_cls[1], g->n_cls[2], w, w + 16, w + 32, (double)n_frames, factor, af, bulk_threshold, cmin + n_classes);
    const int n_entries = n_classes * kDensityTableCounts;
    HOPB_K(hopb_tables::density_table_kernel)<<<(n_entries + 255) / 256, 256, 0, s>>>(
        n_classes, g->n_cls[1], g->n_cls[2], w, w + 16, w + 32, (double)n_frames, factor, af, dtable);
    h->fused_cmin = cmin; h->fused_dtable = dtable; memcpy(h->fused_key, key, sizeof(key));
  }
  // scratch: parents and sizes of both masks, row flags; root keys, counters, annotation sums
  int* parent_s = nullptr;
  // + the two run-head lists: a row holds at most ceil(nz / 2) runs
  const int64_t head_cap = n_rows * ((nz + 1) / 2);
  const int wpr = (nz + 31) / 32;
  rc = hopb_scratch(h, 13, ((size_t)n_cells * 6 + (size_t)head_cap * 2 + (size_t)n_rows * wpr * 2) * sizeof(int) + (size_t)n_rows + 16,
                    (void**)&parent_s);
  if (rc) return rc;
  const uint32_t cap = HOPB_MAX_SITES + 1;
  unsigned char* misc = nullptr;
  const size_t misc_bytes = (size_t)cap * (sizeof(uint64_t) + sizeof(uint32_t)) + 64 +   /* counters: 8 words */
                            (size_t)n_labels_cap * (kAcc + 1) * sizeof(double);
  rc = hopb_scratch(h, 14, misc_bytes, (void**)&misc);
  if (rc) return rc;

  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.counts = counts; a.nx = nx; a.ny = ny; a.nz = nz;
  a.cmin_s = cmin; a.cmin_b = cmin + n_classes;
  a.cls_x = g->d_cls[0]; a.cls_y = g->d_cls[1]; a.cls_z = g->d_cls[2];
  a.ncy = g->n_cls[1]; a.ncz = g->n_cls[2];
  a.dtable = dtable; a.wx = g->d_width[0]; a.wy = g->d_width[1]; a.wz = g->d_width[2];
  a.n_frames = (double)n_frames; a.factor = factor; a.apply_factor = factor != 1.0 ? 1 : 0;
  a.minsite = minsite; a.with_bulk = with_bulk ? 1 : 0;
  a.parent_s = parent_s; a.parent_b = parent_s + n_cells; a.size_s = parent_s + 2 * n_cells; a.size_b = parent_s + 3 * n_cells;
  a.minidx_s = parent_s + 4 * n_cells; a.minidx_b = parent_s + 5 * n_cells;
  a.heads_s = parent_s + 6 * n_cells; a.heads_b = a.heads_s + head_cap;
  a.mask_s = reinterpret_cast<uint32_t*>(a.heads_b + head_cap); a.mask_b = a.mask_s + n_rows * wpr; a.wpr = wpr;
  a.rowflag = reinterpret_cast<uint8_t*>(a.mask_b + n_rows * wpr);
  a.keys = reinterpret_cast<uint64_t*>(misc);
  a.acc = reinterpret_cast<double*>(a.keys + cap);
  a.kshift = a.acc + (size_t)n_labels_cap * kAcc;
  a.vals = reinterpret_cast<uint32_t*>(a.kshift + n_labels_cap);
  a.counters = a.vals + cap;      // 8-byte aligned: cap * 4 is a multiple of 8 (cap = 32768)
  a.cap = cap; a.n_labels_cap = n_labels_cap;
  a.density = density; a.own = own; a.member = with_bulk ? member : nullptr; a.map16 = map16;
  a.brick = brick; a.coarse = coarse; a.coarse_words = (int)((nb + 15) / 16); a.fine = fine; a.nb_pad = (nb + 3) / 4 * 4;
  a.cell2 = cell2; a.n_sites_dev = n_sites_dev; a.volumes = volumes;
  a.mx = g->d_mid[0]; a.my = g->d_mid[1]; a.mz = g->d_mid[2];
  a.p_count = p_count; a.p_mean = p_mean; a.p_std = p_std; a.p_center = p_center;

  int per_sm = 0;
  HOPB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sitemap_fused_kernel, kThreads, smem));
  HOPB_REQUIRE(h, per_sm >= 1, "cooperative launch not possible");
  if (per_sm > 2) per_sm = 2;
  int64_t nblocks = (int64_t)per_sm * h->num_sms;
  const int64_t useful = (n_rows + kWarps - 1) / kWarps;     // more warps than rows do nothing
  if (nblocks > useful) nblocks = useful < 1 ? 1 : useful;
  static const bool trace = getenv("HOPB_FUSED_TRACE") && atoi(getenv("HOPB_FUSED_TRACE")) != 0;   // debugging aid
  unsigned long long* d_trace = nullptr;
  if (trace) {
    rc = hopb_scratch(h, 15, 16 * sizeof(unsigned long long), (void**)&d_trace);
    if (rc) return rc;
    HOPB_CUDA(h, cudaMemsetAsync(d_trace, 0, 16 * sizeof(unsigned long long), s));
  }
  a.trace = d_trace;
  HOPB_CUDA(h, cudaMemsetAsync(a.counters + 4, 0, 2 * sizeof(uint32_t), s));   // fill of the head lists
  void* kargs[] = {(void*)&a};
  hopb_note_launch();
  HOPB_CUDA(h, cudaLaunchCooperativeKernel((void*)sitemap_fused_kernel, dim3((unsigned)nblocks), dim3(kThreads), kargs, smem, s));
  if (trace) {
    unsigned long long t[16];
    HOPB_CUDA(h, cudaMemcpyAsync(t, d_trace, sizeof(t), cudaMemcpyDeviceToHost, s));
    HOPB_CUDA(h, cudaStreamSynchronize(s));
    fprintf(stderr, "sitemap_fused %dx%dx%d, %lld blocks: phases (us)", nx, ny, nz, (long long)nblocks);
    static const char* const names[] = {"", "P0", "P1", "P2", "P3", "P4", "P5", "P6", "P7a", "P8+P7b"};
    for (int k = 1; k <= 9; ++k) fprintf(stderr, " %s %.1f", names[k], t[k] > t[k - 1] ? (t[k] - t[k - 1]) * 1e-3 : 0.0);
    fprintf(stderr, "\n");
  }
  return HOPB_OK;
}

}  // extern "C"
