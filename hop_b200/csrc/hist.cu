// Stage 1: density histogram (K1) and count -> density normalisation.
//
// K1 replaces the per-frame numpy.histogramdd + `grid += h` of DensityCollector.collect
// (hop/density.py:125-131).  Frames do not matter for the histogram, so the kernel sees a flat
// stream of n_points float32 triples.
//
// Roofline: 12 B read per point (HBM) and one 32-bit reduction per in-grid point into a grid that
// is L2-resident for the bench shapes (201^3 u32 = 32 MB).  Water at 0.5 A occupies ~0.004 atoms
// per cell per frame, so neither a warp nor a CTA sees duplicate cells: privatisation has nothing
// to merge, the reductions go straight to L2 (`red.global.add.u32`, fire-and-forget) and L2 atomic
// throughput (measured 2.1e11 RED/s on B200 for a 32 MB grid, tools/ubench.cu) bounds the kernel
// next to the HBM stream.  Everything else is kept off the critical path: the bin search is the
// branch-free pair-table search of hopb_binning.h (~11 instructions per axis).
//
// Optionally the kernel also emits, per point, the brick index of its cell under the
// hopping-trajectory binning rules (4x4x4 bricks, 0xFFFFFFFF = outlier).  Stage 3 can then read
// 4 B per atom-frame instead of re-reading and re-binning 12 B of coordinates.
//
// Data movement: each warp owns 128 consecutive points = 96 float4.  Lanes load float4 j, j+32,
// j+64 (three fully coalesced 512 B requests), park them in a per-warp shared-memory slab, and read
// the points back at a stride of 3 words (conflict-free) so that lane l handles points l, l+32,
// l+64, l+96; the packed cells go out as four coalesced 128 B stores per warp.
//
// Grids beyond the L2 budget (the reduction rate falls once the target outgrows ~half of L2) are reduced in x-slab
// passes over the cached brick indices (hist_cells_kernel) -- for large calls through a grid of 16-bit counters packed
// two per word, which halves the bytes behind the reductions and with them the number of passes; a sum check makes
// that exact whatever the input (hist_flush16_kernel).
#include <cooperative_groups.h>

#include <cstdlib>

#include "hopb_common.cuh"
#include "hopb_tables.cuh"

namespace {

constexpr int kHistThreads = 256;
constexpr int kWarpsPerBlock = kHistThreads / 32;

// streaming load: no L1 allocation, evict-first in L2 so that the coordinate stream does not push
// the count grid (the target of the reductions) out of L2
__device__ __forceinline__ uint64_t make_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ float4 ld_stream(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

// Grids larger than L2 are histogrammed in x-slabs: pass 1 (this kernel) reduces only the points
// whose cell lies in slab [x_lo, x_hi) and leaves the brick index of every point in `cells`; the
// later passes (hist_cells_kernel) re-read those 4-byte indices, one slab each, so that the
// reductions of every pass hit an L2-resident part of the grid.
constexpr uint32_t kCellDone = 0x40000000u;  // histogram contribution already made (edge-zone points)
constexpr uint32_t kCellOut = 0xFFFFFFFFu;

struct AxisReg {
  float lo, inv, nf, top_eq, hop_lo, hop_hi, edge_zone;
  int n, cn;  // bins, bricks (ceil(n/4)) along this axis
  const HopbPair* pairs;
};

__device__ __forceinline__ AxisReg load_axis(const HopbAxis& a, const HopbPair* s_pairs) {
  AxisReg r;
  r.lo = a.lo; r.inv = a.inv; r.nf = (float)a.n; r.top_eq = a.top_eq; r.hop_lo = a.hop_lo; r.hop_hi = a.hop_hi;
  r.n = a.n; r.cn = (a.n + 3) >> 2; r.pairs = s_pairs + a.off2;
  // smallest x for which a last-edge rule can apply (NaN top_eq / empty interval drop out of fminf)
  r.edge_zone = fminf(a.top_eq, a.hop_lo <= a.hop_hi ? a.hop_lo : a.top_eq);
  if (!(r.edge_zone == r.edge_zone)) r.edge_zone = __int_as_float(0x7f800000);  // +inf: never
  return r;
}


// digitize (searchsorted right) in [0, n+1], branch-free
__device__ __forceinline__ int bin_fast(float x, const AxisReg& a) {
  float g = (x - a.lo) * a.inv;
  g = fminf(fmaxf(g, -1.0f), a.nf);
  int b = __float2int_rd(g) + 1;
  const HopbPair e = a.pairs[b];
  b += (x >= e.hi) ? 1 : 0;
  b -= (x < e.lo) ? 1 : 0;
  return b;
}

struct Slab { int x_lo, x_hi; uint32_t done_flag; };  // reduce cells with x_lo <= i < x_hi now

// counts: the selection's own histogram (may be NULL); counts_b: a second histogram that receives the
// point only when `to_b` (the bulk density: solvent not within the cutoff of the solute)
// PACK16: the reduction goes into a grid of 16-bit counters, two cells per 32-bit word (red_packed) -- half the bytes
// behind the scattered reductions, which is what bounds a pass once the grid outgrows L2 (see hist_flush16_kernel for
// how a counter that wraps is found out); n_red counts the reductions this thread has made.
__device__ __forceinline__ void red_packed(uint32_t* counts16, uint32_t cell) {
  atomicAdd(counts16 + (cell >> 1), 1u << ((cell & 1u) << 4));   // result unused -> RED.E.ADD
}

template <bool CELLS, bool FAST, bool PACK16 = false>
__device__ __forceinline__ uint32_t hist_point(float x, float y, float z, const HopbGridTab& tab, const AxisReg& ax,
                                               const AxisReg& ay, const AxisReg& az, const float* s_tab,
                                               uint32_t* counts, const Slab& slab, uint32_t* counts_b = nullptr,
                                               bool to_b = false, uint32_t* lin_out = nullptr, uint32_t* n_red = nullptr) {
  // lin_out != NULL (partitioned histogram): the linear cell of an ordinary point is handed back instead of being
  // reduced (0xFFFFFFFF: nothing to count, or counted here already -- the rare last-edge points)
  if (lin_out) *lin_out = 0xFFFFFFFFu;
  hopb_wrap_point(x, y, z, tab);   // periodic image first, when the grid asks for it (default: off)
  int bx, by, bz;
  bool sure = false;
  if (FAST && tab.sure_eps > 0.0f) {
    // 99.8 % of the points: all three bins follow from the guess alone (hopb_bin_sure), no table
    // lookups -- the kernel is bound by the load/store pipe, where three 8-byte shared-memory loads at
    // scattered addresses per point cost a third of what its reduction costs
    bx = hopb_bin_sure(x, ax.lo, ax.inv, ax.nf, tab.sure_eps);
    by = hopb_bin_sure(y, ay.lo, ay.inv, ay.nf, tab.sure_eps);
    bz = hopb_bin_sure(z, az.lo, az.inv, az.nf, tab.sure_eps);
    sure = (bx != 0) & (by != 0) & (bz != 0);
  }
  if (!sure) {
    if (FAST) {
      bx = bin_fast(x, ax); by = bin_fast(y, ay); bz = bin_fast(z, az);
    } else {
      bx = hopb_bin_right(x, tab.ax[0], s_tab + tab.ax[0].off);
      by = hopb_bin_right(y, tab.ax[1], s_tab + tab.ax[1].off);
      bz = hopb_bin_right(z, tab.ax[2], s_tab + tab.ax[2].off);
    }
  }
  // Common case: the histogram bin and the hopping-trajectory bin are both the digitize result.
  // They differ only within rounding distance of the last edge (numpy.histogramdd's `x ==
  // edges[-1]` rule, _coord2hop's rounded-equality rule); `edge_zone` is the smallest coordinate at
  // which either rule can fire, so one compare per axis routes those rare points to the slow branch
  // (a sure point is never in the edge zone: condition C of the host's proof).
  if (!sure && ((x >= ax.edge_zone) | (y >= ay.edge_zone) | (z >= az.edge_zone))) {
    const int hx = (x == ax.top_eq) ? ax.n : bx;
    const int hy = (y == ay.top_eq) ? ay.n : by;
    const int hz = (z == az.top_eq) ? az.n : bz;
    const int px = bx - (((x >= ax.hop_lo) & (x <= ax.hop_hi)) ? 1 : 0);
    const int py = by - (((y >= ay.hop_lo) & (y <= ay.hop_hi)) ? 1 : 0);
    const int pz = bz - (((z >= az.hop_lo) & (z <= az.hop_hi)) ? 1 : 0);
    const bool okh = ((unsigned)(hx - 1) < (unsigned)ax.n) & ((unsigned)(hy - 1) < (unsigned)ay.n) & ((unsigned)(hz - 1) < (unsigned)az.n);
    if (okh) {  // counted now, whatever the slab; the later passes skip it (kCellDone)
      const uint32_t cell = ((uint32_t)(hx - 1) * (uint32_t)ay.n + (uint32_t)(hy - 1)) * (uint32_t)az.n + (uint32_t)(hz - 1);
      if (PACK16) { red_packed(counts, cell); ++*n_red; }
      else if (counts) atomicAdd(counts + cell, 1u);
      if (to_b) atomicAdd(counts_b + cell, 1u);
    }
    if (!CELLS) return 0u;
    const bool in = ((unsigned)(px - 1) < (unsigned)ax.n) & ((unsigned)(py - 1) < (unsigned)ay.n) & ((unsigned)(pz - 1) < (unsigned)az.n);
    return in ? (hopb_brick_index(px - 1, py - 1, pz - 1, ay.cn, az.cn) | slab.done_flag) : kCellOut;
  }
  const bool ok = ((unsigned)(bx - 1) < (unsigned)ax.n) & ((unsigned)(by - 1) < (unsigned)ay.n) & ((unsigned)(bz - 1) < (unsigned)az.n);
  if (ok & (bx - 1 >= slab.x_lo) & (bx - 1 < slab.x_hi)) {
    const uint32_t cell = ((uint32_t)(bx - 1) * (uint32_t)ay.n + (uint32_t)(by - 1)) * (uint32_t)az.n + (uint32_t)(bz - 1);
    if (lin_out) *lin_out = cell;
    else if (PACK16) { red_packed(counts, cell); ++*n_red; }
    else if (counts) atomicAdd(counts + cell, 1u);  // result unused -> RED.E.ADD
    if (to_b) atomicAdd(counts_b + cell, 1u);
  }
  if (!CELLS) return 0u;
  return ok ? hopb_brick_index(bx - 1, by - 1, bz - 1, ay.cn, az.cn) : kCellOut;
}

// later passes of a slabbed histogram: 4 B per point in, one reduction per point of this slab.
// Slabs are brick-aligned in x, so "in this slab" is a range test on the brick index; only the
// points that pass it (1 / n_pass of them) are decoded into a cell.
struct CellSlab { uint32_t c_lo, c_hi; };  // brick-index range [c_lo, c_hi) of the slab

template <bool PACK16>
__device__ __forceinline__ void hist_cell(uint32_t c, const CellSlab& sl, uint32_t plane, int ny, int nz, int cnz,
                                          uint32_t* __restrict__ counts, uint32_t& n_red) {
  // outliers (0xFFFFFFFF) and already counted points (bit 30) are >= 2^30 > c_hi
  if (c < sl.c_lo || c >= sl.c_hi) return;
  const uint32_t cb = c >> 6;
  const uint32_t ci = cb / plane;
  int di, dj, dk;
  hopb_brick_decode(c, di, dj, dk);
  const uint32_t r = cb - ci * plane;
  const uint32_t cj = r / (uint32_t)cnz;
  const uint32_t ck = r - cj * (uint32_t)cnz;
  const uint32_t i = ci * 4 + (uint32_t)di, j = cj * 4 + (uint32_t)dj, k = ck * 4 + (uint32_t)dk;
  const uint32_t cell = (i * (uint32_t)ny + j) * (uint32_t)nz + k;
  if (PACK16) { red_packed(counts, cell); ++n_red; }
  else atomicAdd(counts + cell, 1u);
}

// sum of a per-thread count over the block, added to *total by one thread (PACK16 bookkeeping)
__device__ __forceinline__ void block_add_u64(uint32_t v, unsigned long long* total) {
  __shared__ uint32_t s_part[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t w = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (threadIdx.x == 0 && w) atomicAdd(total, (unsigned long long)w);
  }
}

// run_flag (may be NULL): the launch does nothing unless *run_flag != 0 -- the direct-reduction passes queued behind a
// packed pass only run when the packed pass had to be thrown away (hist_flush16_kernel)
template <bool PACK16>
__global__ void __launch_bounds__(256)
hist_cells_kernel(const uint32_t* __restrict__ cells, int64_t n_points, int64_t lead, CellSlab sl, int ny, int nz, int cny,
                  int cnz, uint32_t* __restrict__ counts, unsigned long long* n_red_total, const uint32_t* run_flag) {
  if (run_flag && *reinterpret_cast<const volatile uint32_t*>(run_flag) == 0u) return;
  uint32_t n_red = 0;
  const uint64_t pol = make_evict_first_policy();
  const uint32_t plane = (uint32_t)(cny * cnz);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  // [lead, lead + 4 * n4) is 16-byte aligned: four points per load
  const int64_t n4 = (n_points - lead) / 4;
  const uint4* c4 = reinterpret_cast<const uint4*>(cells + lead);
  for (int64_t q = tid; q < n4; q += nth) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(c4 + q), "l"(pol));
    hist_cell<PACK16>(v.x, sl, plane, ny, nz, cnz, counts, n_red);
    hist_cell<PACK16>(v.y, sl, plane, ny, nz, cnz, counts, n_red);
    hist_cell<PACK16>(v.z, sl, plane, ny, nz, cnz, counts, n_red);
    hist_cell<PACK16>(v.w, sl, plane, ny, nz, cnz, counts, n_red);
  }
  const int64_t tail0 = lead + 4 * n4;
  const int64_t n_rem = lead + (n_points - tail0);
  for (int64_t i = tid; i < n_rem; i += nth) {
    const int64_t p = i < lead ? i : tail0 + (i - lead);
    hist_cell<PACK16>(ld_stream_u32(cells + p, pol), sl, plane, ny, nz, cnz, counts, n_red);
  }
  if (PACK16) block_add_u64(n_red, n_red_total);
}

template <bool CELLS, bool FAST, bool PACK16 = false>
__global__ void __launch_bounds__(kHistThreads, 4)
hist_kernel(const float* __restrict__ xyz, int64_t n_points, int64_t lead, HopbGridTab tab,
            const float* __restrict__ table, const HopbPair* __restrict__ table2, uint32_t* __restrict__ counts,
            uint32_t* __restrict__ cells, Slab slab, unsigned long long* n_red_total = nullptr, const uint32_t* run_flag = nullptr) {
  if (run_flag && *reinterpret_cast<const volatile uint32_t*>(run_flag) == 0u) return;
  uint32_t n_red = 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* stage = reinterpret_cast<float4*>(smem_raw);                         // [warps][96] float4
  HopbPair* s_pairs = reinterpret_cast<HopbPair*>(stage + kWarpsPerBlock * 96);  // pair tables
  float* s_tab = reinterpret_cast<float*>(s_pairs + tab.table2_len);            // up32 tables (slow path)
  for (int i = threadIdx.x; i < tab.table2_len; i += blockDim.x) s_pairs[i] = table2[i];
  if (!FAST)
    for (int i = threadIdx.x; i < tab.table_len; i += blockDim.x) s_tab[i] = table[i];
  __syncthreads();
  const AxisReg ax = load_axis(tab.ax[0], s_pairs);
  const AxisReg ay = load_axis(tab.ax[1], s_pairs);
  const AxisReg az = load_axis(tab.ax[2], s_pairs);

  const uint64_t pol = make_evict_first_policy();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float4* wbuf = stage + warp * 96;
  const float* wf = reinterpret_cast<const float*>(wbuf);

  // points [lead, lead + 128*nblk) are 16-byte aligned groups of 128; the rest is done scalar
  const int64_t nblk = (n_points - lead) / 128;
  const float4* base = reinterpret_cast<const float4*>(xyz + 3 * lead);
  const int64_t total_warps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t b = (int64_t)blockIdx.x * kWarpsPerBlock + warp; b < nblk; b += total_warps) {
    const float4* src = base + b * 96;
    const float4 v0 = ld_stream(src + lane, pol);
    const float4 v1 = ld_stream(src + 32 + lane, pol);
    const float4 v2 = ld_stream(src + 64 + lane, pol);
    wbuf[lane] = v0;
    wbuf[32 + lane] = v1;
    wbuf[64 + lane] = v2;
    __syncwarp();
    uint32_t* cout = CELLS ? cells + lead + b * 128 + lane : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = 3 * (lane + 32 * j);
      const uint32_t c = hist_point<CELLS, FAST, PACK16>(wf[p], wf[p + 1], wf[p + 2], tab, ax, ay, az, s_tab, counts, slab, nullptr, false,
                                                         nullptr, &n_red);
      if (CELLS) cout[32 * j] = c;
    }
    __syncwarp();
  }
  // scalar remainder: leading unaligned points and the tail
  const int64_t tail0 = lead + nblk * 128;
  const int64_t n_rem = lead + (n_points - tail0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rem; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i < lead ? i : tail0 + (i - lead);
    const uint32_t c = hist_point<CELLS, FAST, PACK16>(xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], tab, ax, ay, az, s_tab, counts, slab,
                                                       nullptr, false, nullptr, &n_red);
    if (CELLS) cells[p] = c;
  }
  if (PACK16) block_add_u64(n_red, n_red_total);
}

// ---- 16-bit packed count grid for grids larger than L2 ---------------------------------------------------------------
// The scattered reductions run at the L2 rate only while their target stays well inside L2 (64 MiB budget: config 3's
// 297^3 grid is 105 MB of u32, two x-slab passes; config 5's 402^3 is 260 MB, five).  With 16-bit counters the same
// grids are 52 / 130 MB: one pass instead of two, three instead of five.  A launch covers a bounded number of frames, so a
// 16-bit counter does not wrap for any physical input -- but the result must not depend on that: every wrap takes 65535 or
// 65536 out of the sum of all counters (a carry out of a low half adds 1 to the high half; a high half's carry is lost)
// and nothing ever adds to it, so "sum of the counters == number of reductions made" holds exactly when nothing wrapped.
// This kernel (cooperative, two phases) checks that, then either adds the packed grid into the u32 grid or discards it
// and raises *redo, on which the direct-reduction passes queued behind it run (they return at once otherwise).  The
// packed grid is left zeroed either way.
struct Flush16Ctl { unsigned long long n_red, total; uint32_t redo, pad; };

__global__ void __launch_bounds__(512)
hist_flush16_kernel(uint32_t* __restrict__ packed, int64_t n_words, uint32_t* __restrict__ counts, int64_t n_cells, Flush16Ctl* ctl) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  uint32_t part = 0;      // a thread sees at most 2^32 / 2^17 words: no overflow for grids below 2^40 cells
  unsigned long long big = 0;
  for (int64_t i = tid; i < n_words; i += nth) {
    const uint32_t w = packed[i];
    part += (w & 0xFFFFu) + (w >> 16);
    if (part >= 0x80000000u) { big += part; part = 0; }
  }
  big += part;
  __shared__ unsigned long long s_part[16];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) big += __shfl_xor_sync(0xffffffffu, big, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = big;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_part[k];
    if (t) atomicAdd(&ctl->total, t);
  }
  grid.sync();
  const bool ok = *reinterpret_cast<volatile unsigned long long*>(&ctl->total) == *reinterpret_cast<volatile unsigned long long*>(&ctl->n_red);
  for (int64_t i = tid; i < n_words; i += nth) {
    const uint32_t w = packed[i];
    if (w == 0u) continue;
    packed[i] = 0u;
    if (ok) {
      const int64_t c = 2 * i;
      if (w & 0xFFFFu) counts[c] += w & 0xFFFFu;
      if ((w >> 16) && c + 1 < n_cells) counts[c + 1] += w >> 16;
    }
  }
  if (!ok && tid == 0) ctl->redo = 1u;
}

// ---- spatially partitioned, shared-memory-privatised histogram (north_star kernel 2) ---------------------------------
// Water at 0.5 A leaves 0.004 atoms per cell and frame: no tile of the coordinate stream sees a cell twice, so
// privatisation only pays once the points are grouped by where they go.  Pass A is the streaming kernel above with
// the reduction replaced by a CTA-local multisplit: the points of a tile (8192) are counted per bucket (= 16384
// consecutive cells of the grid) in shared memory, every bucket's run is reserved in global memory with ONE atomic
// per bucket and tile, the 14-bit bucket-local cell indices are staged in shared memory in bucket order and written
// out in runs.  Pass B gives every bucket a CTA: ATOMS.ADD into a 64 KB shared tile (1.3e12 /s over the GPU against
// 1.9e11 /s for scattered L2 reductions, profiles/r2_ubench_b200.txt), then counts[cell] += tile[cell], plain and
// coalesced -- a bucket owns its cells.  Per point the L2 sees 2 B written and 2 B read in runs instead of one
// scattered reduction.  Points a bucket has no room for (capacity = twice the mean) and the rare last-edge points
// are reduced directly, so the result never depends on the distribution.  tools/histpart.cu is the stand-alone
// prototype (1.55 vs 1.87 ms on the bench's shape).
constexpr int kPartShift = 14;                  // 16384 cells per bucket: a 64 KB shared tile in pass B
constexpr int kPartTile = 8192;                 // points per CTA and tile in pass A
constexpr int kPartThreads = 512;
constexpr int kPartMaxBuckets = 1024;

template <bool FAST>
__global__ void __launch_bounds__(kPartThreads, 2)
hist_part_kernel(const float* __restrict__ xyz, int64_t n_points, int64_t lead, HopbGridTab tab, const float* __restrict__ table,
                 const HopbPair* __restrict__ table2, uint32_t* __restrict__ counts, uint32_t* __restrict__ cells, int n_buckets,
                 uint32_t* __restrict__ gfill, uint16_t* __restrict__ buckets, uint32_t cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kWarps = kPartThreads / 32;
  float4* stage = reinterpret_cast<float4*>(smem_raw);                             // [warps][96] float4
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(stage + kWarps * 96);               // [n_buckets] counts, then offsets
  uint32_t* s_gbase = s_cnt + kPartMaxBuckets;                                      // [n_buckets]
  uint16_t* s_stage = reinterpret_cast<uint16_t*>(s_gbase + kPartMaxBuckets);       // [kPartTile] bucket-local cells in bucket order
  uint16_t* s_bkt = s_stage + kPartTile;                                            // [kPartTile] their buckets
  HopbPair* s_pairs = reinterpret_cast<HopbPair*>(s_bkt + kPartTile);
  float* s_tab = reinterpret_cast<float*>(s_pairs + tab.table2_len);
  __shared__ uint32_t s_warp[kWarps];
  __shared__ uint32_t s_total;
  for (int i = threadIdx.x; i < tab.table2_len; i += blockDim.x) s_pairs[i] = table2[i];
  if (!FAST)
    for (int i = threadIdx.x; i < tab.table_len; i += blockDim.x) s_tab[i] = table[i];
  __syncthreads();
  const AxisReg ax = load_axis(tab.ax[0], s_pairs);
  const AxisReg ay = load_axis(tab.ax[1], s_pairs);
  const AxisReg az = load_axis(tab.ax[2], s_pairs);
  Slab slab; slab.x_lo = 0; slab.x_hi = tab.ax[0].n; slab.done_flag = 0u;
  const uint64_t pol = make_evict_first_policy();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* wbuf = stage + warp * 96;
  const float* wf = reinterpret_cast<const float*>(wbuf);
  constexpr int kBlocksPerWarp = kPartTile / 128 / kWarps;      // 128-point groups per warp and tile (4)
  const int64_t nblk = (n_points - lead) / 128;
  const float4* base = reinterpret_cast<const float4*>(xyz + 3 * lead);
  const int64_t n_tiles = (nblk + kPartTile / 128 - 1) / (kPartTile / 128);
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    for (int b = threadIdx.x; b < n_buckets; b += blockDim.x) s_cnt[b] = 0;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    uint32_t lin[kBlocksPerWarp * 4], slot[kBlocksPerWarp * 4];
#pragma unroll
    for (int g = 0; g < kBlocksPerWarp; ++g) {
      const int64_t b = t * (kPartTile / 128) + (int64_t)g * kWarps + warp;
      if (b < nblk) {
        const float4* src = base + b * 96;
        const float4 v0 = ld_stream(src + lane, pol);
        const float4 v1 = ld_stream(src + 32 + lane, pol);
        const float4 v2 = ld_stream(src + 64 + lane, pol);
        wbuf[lane] = v0; wbuf[32 + lane] = v1; wbuf[64 + lane] = v2;
        __syncwarp();
        uint32_t* cout = cells + lead + b * 128 + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p = 3 * (lane + 32 * j);
          uint32_t l;
          const uint32_t c = hist_point<true, FAST>(wf[p], wf[p + 1], wf[p + 2], tab, ax, ay, az, s_tab, counts, slab, nullptr, false, &l);
          cout[32 * j] = c;
          lin[4 * g + j] = l;
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) lin[4 * g + j] = 0xFFFFFFFFu;
      }
    }
#pragma unroll
    for (int u = 0; u < kBlocksPerWarp * 4; ++u)
      if (lin[u] != 0xFFFFFFFFu) slot[u] = atomicAdd(&s_cnt[lin[u] >> kPartShift], 1u);
    __syncthreads();
    // exclusive scan of the bucket counts (two buckets per thread), one global reservation per non-empty bucket
    uint32_t c0 = 0, c1 = 0;
    const int b0 = 2 * threadIdx.x, b1 = b0 + 1;
    if (b0 < n_buckets) c0 = s_cnt[b0];
    if (b1 < n_buckets) c1 = s_cnt[b1];
    uint32_t incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < kWarps ? s_warp[lane] : 0, wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
      if (lane < kWarps) s_warp[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t ex = s_warp[warp] + incl - (c0 + c1);
    if (b0 < n_buckets) { s_cnt[b0] = ex; s_gbase[b0] = c0 ? atomicAdd(gfill + b0, c0) : 0u; }
    if (b1 < n_buckets) { s_cnt[b1] = ex + c0; s_gbase[b1] = c1 ? atomicAdd(gfill + b1, c1) : 0u; }
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for (int u = 0; u < kBlocksPerWarp * 4; ++u)
      if (lin[u] != 0xFFFFFFFFu) {
        const uint32_t b = lin[u] >> kPartShift, p = s_cnt[b] + slot[u];
        s_stage[p] = (uint16_t)(lin[u] & ((1u << kPartShift) - 1));
        s_bkt[p] = (uint16_t)b;
        ++total;
      }
    if (total) atomicAdd(&s_total, total);       // number of staged entries of the tile
    __syncthreads();
    const uint32_t n_st = s_total;
    for (uint32_t i = threadIdx.x; i < n_st; i += blockDim.x) {
      const uint32_t b = s_bkt[i];
      const uint32_t pos = s_gbase[b] + (i - s_cnt[b]);
      if (pos < cap) buckets[(size_t)b * cap + pos] = s_stage[i];
      else atomicAdd(counts + ((b << kPartShift) | s_stage[i]), 1u);      // bucket full: an ordinary reduction
    }
    __syncthreads();
  }
  // scalar remainder: leading unaligned points and the tail, reduced directly
  const int64_t tail0 = lead + nblk * 128;
  const int64_t n_rem = lead + (n_points - tail0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rem; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i < lead ? i : tail0 + (i - lead);
    cells[p] = hist_point<true, FAST>(xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], tab, ax, ay, az, s_tab, counts, slab);
  }
}

__global__ void __launch_bounds__(1024)
hist_tiles_kernel(const uint16_t* __restrict__ buckets, const uint32_t* __restrict__ gfill, uint32_t cap, int n_buckets,
                  uint32_t* __restrict__ counts, uint32_t n_cells) {
  extern __shared__ uint32_t s_tile[];
  constexpr int kCells = 1 << kPartShift;
  for (int b = blockIdx.x; b < n_buckets; b += gridDim.x) {
    for (int i = threadIdx.x; i < kCells; i += blockDim.x) s_tile[i] = 0;
    __syncthreads();
    uint32_t nb = gfill[b];
    if (nb > cap) nb = cap;
    const uint4* src = reinterpret_cast<const uint4*>(buckets + (size_t)b * cap);     // eight entries per load
    for (uint32_t i = threadIdx.x; i < (nb + 7) / 8; i += blockDim.x) {
      const uint4 v = src[i];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (8 * i + 2 * k < nb) atomicAdd(&s_tile[w[k] & 0xFFFFu], 1u);
        if (8 * i + 2 * k + 1 < nb) atomicAdd(&s_tile[w[k] >> 16], 1u);
      }
    }
    __syncthreads();
    const uint32_t c0 = (uint32_t)b << kPartShift;
    for (int i = threadIdx.x; i < kCells; i += blockDim.x)
      if (c0 + i < n_cells && s_tile[i]) counts[c0 + i] += s_tile[i];
    __syncthreads();
  }
}

// ---- bulk density: '<solvent> not within <cutoff> of <solute>' fused with the histogram -------------
// Replaces the per-frame kd-tree selection of notwithin_coordinates_factory (hop/density.py:82-87)
// followed by numpy.histogramdd.  One CTA per frame: the solute atoms of the frame are binned into a
// uniform cell list (cell edge >= cutoff) in shared memory, every solvent atom tests the 27 cells
// around it (first hit wins) and is then reduced into the solvent histogram and, when no solute atom
// is within the cutoff, into the bulk histogram -- both densities from one read of the coordinates.
// "Within" means d^2 <= cutoff^2 with d^2 = (dx*dx + dy*dy) + dz*dz evaluated in float32 without FMA.
constexpr int kBulkThreads = 256;
constexpr int kBulkMaxCells = 4096;

template <bool CELLS, bool FAST>
__global__ void __launch_bounds__(kBulkThreads)
hist_bulk_kernel(const float* __restrict__ solvent, const float* __restrict__ solute, int64_t Nw, int Np, float cutoff,
                 HopbGridTab tab, const float* __restrict__ table, const HopbPair* __restrict__ table2,
                 uint32_t* __restrict__ counts_solvent, uint32_t* __restrict__ counts_bulk, uint32_t* __restrict__ cells,
                 uint8_t* __restrict__ within_out, const uint8_t* __restrict__ within_in, int64_t solute_stride, int flags_only) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HopbPair* s_pairs = reinterpret_cast<HopbPair*>(smem_raw);
  float* s_tab = reinterpret_cast<float*>(s_pairs + tab.table2_len);
  int* s_start = reinterpret_cast<int*>(s_tab + (FAST ? 0 : tab.table_len));  // [kBulkMaxCells + 1]
  int* s_cursor = s_start + kBulkMaxCells + 1;                               // [kBulkMaxCells]
  float* s_atoms = reinterpret_cast<float*>(s_cursor + kBulkMaxCells);        // [Np][3], sorted by cell
  __shared__ float s_lo[3], s_hi[3];
  __shared__ float s_red[2][3][kBulkThreads / 32];
  __shared__ uint32_t s_warp[33];

  const int64_t frame = blockIdx.x;
  // the solute may arrive in tiles (selections too large for the shared-memory cell list): P is this tile's part of the
  // frame, within_in what the earlier tiles found; all tiles but the last only update the flags
  const float* P = solute + frame * solute_stride * 3;
  const float* W = solvent + frame * Nw * 3;
  for (int i = threadIdx.x; i < tab.table2_len; i += blockDim.x) s_pairs[i] = table2[i];
  if (!FAST)
    for (int i = threadIdx.x; i < tab.table_len; i += blockDim.x) s_tab[i] = table[i];

  // 1. bounding box of the solute in this frame
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int i = threadIdx.x; i < Np; i += blockDim.x)
#pragma unroll
    for (int d = 0; d < 3; ++d) { const float v = P[3 * i + d]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float a = lo[d], b = hi[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o)); }
    if (lane == 0) { s_red[0][d][warp] = a; s_red[1][d][warp] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float a = 3.0e38f, b = -3.0e38f;
    for (int w = 0; w < kBulkThreads / 32; ++w) { a = fminf(a, s_red[0][threadIdx.x][w]); b = fmaxf(b, s_red[1][threadIdx.x][w]); }
    s_lo[threadIdx.x] = a; s_hi[threadIdx.x] = b;
  }
  __syncthreads();
  // 2. cell grid: edge = cutoff, doubled until the grid fits the shared-memory table
  // a hair above the cutoff so that rounding in the cell coordinates can never separate a pair
  // at distance <= cutoff by more than one cell
  float cs = cutoff > 0.0f ? cutoff * 1.001f : 1.0f;
  int dim[3];
  while (true) {
    int64_t tot = 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) { dim[d] = Np > 0 ? (int)floorf((s_hi[d] - s_lo[d]) / cs) + 1 : 1; tot *= dim[d]; }
    if (tot <= kBulkMaxCells) break;
    cs *= 2.0f;
  }
  const float inv_cs = 1.0f / cs;
  const int n_cells = dim[0] * dim[1] * dim[2];
  for (int i = threadIdx.x; i <= n_cells; i += blockDim.x) s_start[i] = 0;
  __syncthreads();
  auto cell_of = [&](float x, float y, float z) {
    const int cx = min(max((int)floorf((x - s_lo[0]) * inv_cs), 0), dim[0] - 1);
    const int cy = min(max((int)floorf((y - s_lo[1]) * inv_cs), 0), dim[1] - 1);
    const int cz = min(max((int)floorf((z - s_lo[2]) * inv_cs), 0), dim[2] - 1);
    return (cx * dim[1] + cy) * dim[2] + cz;
  };
  for (int i = threadIdx.x; i < Np; i += blockDim.x) atomicAdd(&s_start[cell_of(P[3 * i], P[3 * i + 1], P[3 * i + 2])], 1);
  __syncthreads();
  // exclusive scan of the cell counts (n_cells <= 4096: 16 per thread)
  {
    const int per = (n_cells + kBulkThreads - 1) / kBulkThreads;
    const int b0 = threadIdx.x * per;
    uint32_t sum = 0;
    for (int k = 0; k < per; ++k) if (b0 + k < n_cells) sum += (uint32_t)s_start[b0 + k];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < kBulkThreads / 32 ? s_warp[lane] : 0, winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
      s_warp[lane] = winc - w;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + inc - sum;
    for (int k = 0; k < per; ++k)
      if (b0 + k < n_cells) { const uint32_t c = (uint32_t)s_start[b0 + k]; s_start[b0 + k] = (int)run; s_cursor[b0 + k] = (int)run; run += c; }
    if (threadIdx.x == kBulkThreads - 1) s_start[n_cells] = Np;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Np; i += blockDim.x) {
    const float x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
    const int slot = atomicAdd(&s_cursor[cell_of(x, y, z)], 1);
    s_atoms[3 * slot] = x; s_atoms[3 * slot + 1] = y; s_atoms[3 * slot + 2] = z;
  }
  __syncthreads();
  const AxisReg ax = load_axis(tab.ax[0], s_pairs);
  const AxisReg ay = load_axis(tab.ax[1], s_pairs);
  const AxisReg az = load_axis(tab.ax[2], s_pairs);
  Slab slab; slab.x_lo = 0; slab.x_hi = tab.ax[0].n; slab.done_flag = 0u;
  const float c2 = __fmul_rn(cutoff, cutoff);
  // 3. every solvent atom: neighbour search, then both histograms
  for (int64_t i = threadIdx.x; i < Nw; i += blockDim.x) {
    const float x = W[3 * i], y = W[3 * i + 1], z = W[3 * i + 2];
    bool within = false;
    if (Np > 0) {
      // cells whose box can hold an atom within the cutoff: +-1 around the atom's (unclamped) cell
      const int cx = (int)floorf((x - s_lo[0]) * inv_cs), cy = (int)floorf((y - s_lo[1]) * inv_cs), cz = (int)floorf((z - s_lo[2]) * inv_cs);
      const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dim[0] - 1);
      const int y0 = max(cy - 1, 0), y1 = min(cy + 1, dim[1] - 1);
      const int z0 = max(cz - 1, 0), z1 = min(cz + 1, dim[2] - 1);
      for (int a = x0; a <= x1 && !within && z0 <= z1; ++a)
        for (int b = y0; b <= y1 && !within; ++b) {
          const int row = (a * dim[1] + b) * dim[2];
          const int j0 = s_start[row + z0], j1 = s_start[row + z1 + 1];   // cells z0..z1 are contiguous
          for (int j = j0; j < j1; ++j) {
            const float dx = __fadd_rn(x, -s_atoms[3 * j]), dy = __fadd_rn(y, -s_atoms[3 * j + 1]), dz = __fadd_rn(z, -s_atoms[3 * j + 2]);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d2 <= c2) { within = true; break; }
          }
        }
    }
    if (within_in) within = within || within_in[frame * Nw + i] != 0;
    if (flags_only) { within_out[frame * Nw + i] = within ? 1 : 0; continue; }
    const uint32_t c = hist_point<CELLS, FAST>(x, y, z, tab, ax, ay, az, s_tab, counts_solvent, slab, counts_bulk, !within);
    if (CELLS) cells[frame * Nw + i] = c;
    if (within_out) within_out[frame * Nw + i] = within ? 1 : 0;
  }
}

// ---- several selections from one coordinate stream (BASELINE config 4: Na+ / K+ / Cl- in one box) -----------------
// hop runs one DensityCreator per selection, i.e. reads the trajectory once per species
// (scripts/hop-generate-many-densities.py:98-125; hop/density.py:271-304), each with a grid of its own (the extent
// of THAT selection in frame 0).  Here every frame is read once: the species of an atom picks the grid tables, the
// count grid and the brick numbering its point is binned with.  Species are few and small (10^2 .. 10^3 ions), so
// the stream is latency- rather than reduction-bound: every thread takes four points per step, loads first.
constexpr int kMaxSpecies = 8;
struct MultiGrids {
  HopbGridTab tab[kMaxSpecies];
  const HopbPair* table2[kMaxSpecies];
  uint32_t* counts[kMaxSpecies];
  int n;
};

__global__ void __launch_bounds__(kHistThreads, 4)
hist_multi_kernel(const float* __restrict__ xyz, int64_t n_points, int64_t lead, int64_t N, const uint8_t* __restrict__ species,
                  const MultiGrids* __restrict__ mg, uint32_t* __restrict__ cells) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* stage = reinterpret_cast<float4*>(smem_raw);                           // [warps][96] float4, as hist_kernel
  HopbPair* s_pairs = reinterpret_cast<HopbPair*>(stage + kWarpsPerBlock * 96);
  __shared__ int s_off[kMaxSpecies];
  __shared__ AxisReg s_ax[kMaxSpecies][3];
  int off = 0;
  for (int k = 0; k < mg->n; ++k) {
    for (int i = threadIdx.x; i < mg->tab[k].table2_len; i += blockDim.x) s_pairs[off + i] = mg->table2[k][i];
    if (threadIdx.x == 0) s_off[k] = off;
    off += mg->tab[k].table2_len;
  }
  __syncthreads();
  if (threadIdx.x < mg->n * 3) {
    const int k = threadIdx.x / 3, d = threadIdx.x % 3;
    s_ax[k][d] = load_axis(mg->tab[k].ax[d], s_pairs + s_off[k]);
  }
  __syncthreads();
  const uint64_t pol = make_evict_first_policy();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* wbuf = stage + warp * 96;
  const float* wf = reinterpret_cast<const float*>(wbuf);
  const uint32_t Nu = (uint32_t)N;
  auto point = [&](float x, float y, float z, uint32_t atom) -> uint32_t {
    const int k = species[atom];
    Slab slab; slab.x_lo = 0; slab.x_hi = mg->tab[k].ax[0].n; slab.done_flag = 0u;
    return hist_point<true, true>(x, y, z, mg->tab[k], s_ax[k][0], s_ax[k][1], s_ax[k][2], nullptr, mg->counts[k], slab);
  };
  const int64_t nblk = (n_points - lead) / 128;
  const float4* base = reinterpret_cast<const float4*>(xyz + 3 * lead);
  const int64_t total_warps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t b = (int64_t)blockIdx.x * kWarpsPerBlock + warp; b < nblk; b += total_warps) {
    const float4* src = base + b * 96;
    const float4 v0 = ld_stream(src + lane, pol);
    const float4 v1 = ld_stream(src + 32 + lane, pol);
    const float4 v2 = ld_stream(src + 64 + lane, pol);
    wbuf[lane] = v0; wbuf[32 + lane] = v1; wbuf[64 + lane] = v2;
    __syncwarp();
    const int64_t p0 = lead + b * 128;
    const uint32_t a0 = (uint32_t)(p0 % N);        // atom of the block's first point; the others follow modulo N
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = lane + 32 * j;
      const uint32_t c = point(wf[3 * q], wf[3 * q + 1], wf[3 * q + 2], (a0 + (uint32_t)q) % Nu);
      if (cells) cells[p0 + q] = c;
    }
    __syncwarp();
  }
  const int64_t tail0 = lead + nblk * 128;
  const int64_t n_rem = lead + (n_points - tail0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rem; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i < lead ? i : tail0 + (i - lead);
    const uint32_t c = point(xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], (uint32_t)(p % N));
    if (cells) cells[p] = c;
  }
}

// One warp per (i, j) row of the grid: no per-cell index division, the two row divisors are loaded
// once, loads and stores are contiguous.  The four divisions stay IEEE double divisions in the
// reference's order (bit-identical densities); they are what the kernel spends its time on.
__global__ void __launch_bounds__(256)
density_kernel(const uint32_t* __restrict__ counts, int64_t n_rows, int ny, int nz, double n_frames,
               const double* __restrict__ wx, const double* __restrict__ wy, const double* __restrict__ wz,
               double factor, int apply_factor, double* __restrict__ density) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < n_rows; row += n_warps) {
    const int j = (int)(row % ny);
    const int i = (int)(row / ny);
    const double dx = wx[i], dy = wy[j];
    const uint32_t* src = counts + row * nz;
    double* dst = density + row * nz;
    for (int k = lane; k < nz; k += 32) {
      dst[k] = hopb_density_value(src[k], n_frames, dx, dy, wz[k], factor, apply_factor);
    }
  }
}

using hopb_tables::kDensityTableCounts;

__global__ void __launch_bounds__(256)
density_lookup_kernel(const uint32_t* __restrict__ counts, int64_t n_rows, int ny, int nz, double n_frames,
                      const double* __restrict__ wx, const double* __restrict__ wy, const double* __restrict__ wz,
                      double factor, int apply_factor, const uint8_t* __restrict__ cls_x,
                      const uint8_t* __restrict__ cls_y, const uint8_t* __restrict__ cls_z, int ncy, int ncz,
                      const double* __restrict__ table, double* __restrict__ density) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < n_rows; row += n_warps) {
    const int j = (int)(row % ny);
    const int i = (int)(row / ny);
    const double* trow = table + (int64_t)((int)cls_x[i] * ncy + (int)cls_y[j]) * ncz * kDensityTableCounts;
    const uint32_t* src = counts + row * nz;
    double* dst = density + row * nz;
    // four chunks of 32 cells per step: counts first, then the table look-ups, then the stores
    for (int k0 = lane; k0 < nz; k0 += 128) {
      uint32_t c[4];
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c[u] = k0 + 32 * u < nz ? src[k0 + 32 * u] : 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + 32 * u;
        if (k >= nz) continue;
        if (c[u] < (uint32_t)kDensityTableCounts) v[u] = __ldg(trow + (int)cls_z[k] * kDensityTableCounts + (int)c[u]);
        else v[u] = hopb_density_value(c[u], n_frames, wx[i], wy[j], wz[k], factor, apply_factor);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (k0 + 32 * u < nz) dst[k0 + 32 * u] = v[u];
    }
  }
}

}  // namespace

extern "C" {

int hopb_hist_accumulate_cells(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz, int64_t n_points,
                               uint32_t* counts, uint32_t* cells) {
  HOPB_REQUIRE(h, g && xyz && counts && n_points >= 0, "bad argument");
  HOPB_REQUIRE(h, ((uintptr_t)xyz & 3) == 0, "xyz must be 4-byte aligned");
  if (cells) {
    const int64_t nbk = (int64_t)((g->n[0] + 3) / 4) * ((g->n[1] + 3) / 4) * ((g->n[2] + 3) / 4);
    HOPB_REQUIRE(h, nbk * 64 < ((int64_t)1 << 30), "grid too large for 30-bit brick indices");
  }
  if (n_points == 0) return HOPB_OK;
  // number of leading points until the stream is 16-byte aligned (12 B per point)
  int64_t lead = 0;
  while (lead < n_points && (((uintptr_t)xyz + 12 * (uintptr_t)lead) & 15) != 0) ++lead;
  const int64_t nblk = (n_points - lead) / 128;
  const bool fast = g->tab.fast_ok != 0;
  const size_t smem = kWarpsPerBlock * 96 * sizeof(float4) + (size_t)g->tab.table2_len * sizeof(HopbPair) +
                      (fast ? 0 : (size_t)g->tab.table_len * sizeof(float));
  // ~4 groups of 128 points per warp: many short-lived blocks balance better across the SMs than a
  // persistent grid (measured on B200: 1.76 ms vs 2.11 ms for 3e8 points)
  static const int per_warp = getenv("HOPB_HIST_PER_WARP") ? atoi(getenv("HOPB_HIST_PER_WARP")) : 4;  // tuning hook
  int64_t blocks = (nblk + (int64_t)kWarpsPerBlock * per_warp - 1) / ((int64_t)kWarpsPerBlock * per_warp);
  if (blocks < 1) blocks = 1;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned nb = (unsigned)blocks;
  // x-slabs: keep the part of the grid that receives reductions in one pass well inside L2
  const int64_t row_bytes = (int64_t)g->n[1] * g->n[2] * 4;
  int n_pass = 1;
  if (cells) {
    const int64_t budget = h->hist_l2_budget;
    n_pass = (int)((row_bytes * g->n[0] + budget - 1) / budget);
    if (n_pass < 1) n_pass = 1;
    if (n_pass > g->n[0]) n_pass = g->n[0];
  }
  // slabs are whole bricks (multiples of 4 cells) wide, so that the later passes can tell "in my slab"
  // from the brick index alone
  // The partitioned, shared-memory-privatised histogram: one call with many points per cell of a grid that fits the
  // bucket scheme (the pipeline's resident and streamed runs); small calls keep the direct reductions.
  // OFF by default: measured on B200 at the bench shape the two passes take 2.13 + 0.21 ms against 1.75 ms for the
  // direct reductions -- the multisplit costs ~170 thread instructions per point and pass A is issue-bound (67 % issue
  // active, 1.6e9 warp instructions; profiles/r2_hist_partition.md).  HOPB_HIST_PART=1 enables it for large calls,
  // =2 for every call (tests).
  const char* part_env = getenv("HOPB_HIST_PART");
  const int part_mode = part_env ? atoi(part_env) : 0;
  const int64_t n_cells_all = (int64_t)g->n[0] * g->n[1] * g->n[2];
  const int n_buckets = (int)((n_cells_all + (1 << kPartShift) - 1) >> kPartShift);
  if (part_mode && cells && n_pass == 1 && n_buckets <= kPartMaxBuckets && (n_points >= 8 * n_cells_all || part_mode == 2)) {
    const uint32_t cap = (uint32_t)((((2 * n_points) / n_buckets) + 65535) & ~(int64_t)7);     // twice the mean, 16-byte rows
    unsigned char* scr = nullptr;
    int rc = hopb_scratch(h, 18, (size_t)n_buckets * cap * sizeof(uint16_t) + (size_t)kPartMaxBuckets * sizeof(uint32_t) + 64, (void**)&scr);
    if (rc) return rc;
    uint32_t* gfill = reinterpret_cast<uint32_t*>(scr);
    uint16_t* buckets = reinterpret_cast<uint16_t*>(scr + (size_t)kPartMaxBuckets * sizeof(uint32_t));
    HOPB_CUDA(h, cudaMemsetAsync(gfill, 0, (size_t)n_buckets * sizeof(uint32_t), s));
    const size_t smem_a = (size_t)(kPartThreads / 32) * 96 * sizeof(float4) + (size_t)kPartMaxBuckets * 8 + (size_t)kPartTile * 4 +
                          (size_t)g->tab.table2_len * sizeof(HopbPair) + (fast ? 0 : (size_t)g->tab.table_len * sizeof(float));
    const size_t smem_b = (size_t)4 << kPartShift;
    const int64_t n_tiles = (nblk + kPartTile / 128 - 1) / (kPartTile / 128);
    int64_t grid_a = (int64_t)h->num_sms * 2;
    if (grid_a > n_tiles) grid_a = n_tiles < 1 ? 1 : n_tiles;
    hopb_note_launch();
    if (fast) {
      HOPB_CUDA(h, cudaFuncSetAttribute(hist_part_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
      hist_part_kernel<true><<<(unsigned)grid_a, kPartThreads, smem_a, s>>>(xyz, n_points, lead, g->tab, g->d_table, g->d_table2, counts, cells,
                                                                            n_buckets, gfill, buckets, cap);
    } else {
      HOPB_CUDA(h, cudaFuncSetAttribute(hist_part_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
      hist_part_kernel<false><<<(unsigned)grid_a, kPartThreads, smem_a, s>>>(xyz, n_points, lead, g->tab, g->d_table, g->d_table2, counts, cells,
                                                                             n_buckets, gfill, buckets, cap);
    }
    HOPB_CUDA(h, cudaFuncSetAttribute(hist_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    HOPB_K(hist_tiles_kernel)<<<(unsigned)n_buckets, 1024, smem_b, s>>>(buckets + 0, gfill, cap, n_buckets, counts, (uint32_t)n_cells_all);
    HOPB_LAUNCH_CHECK(h);
    return HOPB_OK;
  }
  // all x-slab passes of one histogram: into the u32 grid, or (packed) into a grid of 16-bit counters
  auto launch_passes = [&](bool packed, uint32_t* target, int passes, unsigned long long* n_red_total, const uint32_t* run_flag) {
    int slab_w = (g->n[0] + passes - 1) / passes;
    if (passes > 1) slab_w = (slab_w + 3) / 4 * 4;
    Slab slab;
    slab.x_lo = 0; slab.x_hi = passes == 1 ? g->n[0] : slab_w; slab.done_flag = passes == 1 ? 0u : kCellDone;
    hopb_note_launch();
    if (packed) {
      hist_kernel<true, true, true><<<nb, kHistThreads, smem, s>>>(xyz, n_points, lead, g->tab, g->d_table, g->d_table2, target, cells, slab,
                                                                    n_red_total, run_flag);
    } else if (cells) {
      if (fast) hist_kernel<true, true><<<nb, kHistThreads, smem, s>>>(xyz, n_points, lead, g->tab, g->d_table, g->d_table2, target, cells, slab, nullptr, run_flag);
      else hist_kernel<true, false><<<nb, kHistThreads, smem, s>>>(xyz, n_points, lead, g->tab, g->d_table, g->d_table2, target, cells, slab, nullptr, run_flag);
    } else {
      if (fast) hist_kernel<false, true><<<nb, kHistThreads, smem, s>>>(xyz, n_points, lead, g->tab, g->d_table, g->d_table2, target, cells, slab, nullptr, run_flag);
      else hist_kernel<false, false><<<nb, kHistThreads, smem, s>>>(xyz, n_points, lead, g->tab, g->d_table, g->d_table2, target, cells, slab, nullptr, run_flag);
    }
    if (passes > 1) {
      const int cny = (g->n[1] + 3) / 4, cnz = (g->n[2] + 3) / 4;
      const uint32_t plane64 = (uint32_t)(cny * cnz) * 64u;
      int64_t lead4 = 0;  // points until the index stream is 16-byte aligned
      while (lead4 < n_points && (((uintptr_t)cells + 4 * (uintptr_t)lead4) & 15) != 0) ++lead4;
      for (int p = 1; p < passes; ++p) {
        const int x_lo = p * slab_w;
        const int x_hi = (p + 1) * slab_w < g->n[0] ? (p + 1) * slab_w : g->n[0];
        if (x_lo >= x_hi) break;
        CellSlab cs;
        cs.c_lo = (uint32_t)(x_lo / 4) * plane64;
        cs.c_hi = (uint32_t)((x_hi + 3) / 4) * plane64;
        const unsigned gb = hopb_grid_for((n_points + 3) / 4, 256, h->num_sms, 8);
        if (packed) HOPB_K(hist_cells_kernel<true>)<<<gb, 256, 0, s>>>(cells, n_points, lead4, cs, g->n[1], g->n[2], cny, cnz, target, n_red_total, run_flag);
        else HOPB_K(hist_cells_kernel<false>)<<<gb, 256, 0, s>>>(cells, n_points, lead4, cs, g->n[1], g->n[2], cny, cnz, target, nullptr, run_flag);
      }
    }
  };
  // Grids that need more than one pass as u32 go through 16-bit packed counters first (hist_flush16_kernel): half the
  // bytes behind the reductions, hence fewer passes.  HOPB_HIST_PACK16=0 switches it off, =2 forces it (tests).
  const char* pack_env = getenv("HOPB_HIST_PACK16");
  const int pack_mode = pack_env ? atoi(pack_env) : 1;
  // (only for calls with at least a point per cell: the flush walks the whole grid, which a call of a few frames would
  // not earn back)
  if (cells && fast && pack_mode && ((n_pass > 1 && n_points >= n_cells_all) || pack_mode == 2) && n_cells_all < ((int64_t)1 << 32)) {
    const int64_t n_words = (n_cells_all + 1) / 2;
    int64_t budget16 = h->hist_l2_budget;
    if (const char* e = getenv("HOPB_HIST_PACK16_BUDGET_MB")) {      // tuning hook
      if (atoll(e) > 0) budget16 = atoll(e) << 20;
    }
    int passes16 = (int)((row_bytes / 2 * g->n[0] + budget16 - 1) / budget16);
    if (passes16 < 1) passes16 = 1;
    if (passes16 > g->n[0]) passes16 = g->n[0];
    unsigned char* scr = nullptr;
    const size_t had = h->scratch_bytes[19];
    int rc = hopb_scratch(h, 19, (size_t)n_words * 4 + sizeof(Flush16Ctl) + 64, (void**)&scr);
    if (rc) return rc;
    uint32_t* packed = reinterpret_cast<uint32_t*>(scr + 64);
    Flush16Ctl* ctl = reinterpret_cast<Flush16Ctl*>(scr);
    if (had != h->scratch_bytes[19])      // fresh memory: the packed grid starts zeroed (every flush leaves it zeroed)
      HOPB_CUDA(h, cudaMemsetAsync(scr, 0, h->scratch_bytes[19], s));
    HOPB_CUDA(h, cudaMemsetAsync(ctl, 0, sizeof(Flush16Ctl), s));
    launch_passes(true, packed, passes16, &ctl->n_red, nullptr);
    {
      int per_sm = 0;
      HOPB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hist_flush16_kernel, 512, 0));
      HOPB_REQUIRE(h, per_sm >= 1, "cooperative launch not possible");
      if (per_sm > 2) per_sm = 2;
      int64_t nblocks = (int64_t)per_sm * h->num_sms;
      const int64_t useful = (n_words + 511) / 512;
      if (nblocks > useful) nblocks = useful;
      uint32_t* counts_arg = counts;
      int64_t n_words_arg = n_words, n_cells_arg = n_cells_all;
      void* kargs[] = {(void*)&packed, (void*)&n_words_arg, (void*)&counts_arg, (void*)&n_cells_arg, (void*)&ctl};
      hopb_note_launch();
      HOPB_CUDA(h, cudaLaunchCooperativeKernel((void*)hist_flush16_kernel, dim3((unsigned)nblocks), dim3(512), kargs, 0, s));
    }
    launch_passes(false, counts, n_pass, nullptr, &ctl->redo);     // only does anything when the packed grid was discarded
    HOPB_LAUNCH_CHECK(h);
    return HOPB_OK;
  }
  launch_passes(false, counts, n_pass, nullptr, nullptr);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hist_bulk(hopb_handle* h, void* stream, const hopb_grid* g, const float* solvent, const float* solute,
                   int64_t F, int64_t Nw, int64_t Np, double cutoff, uint32_t* counts_solvent, uint32_t* counts_bulk,
                   uint32_t* cells, uint8_t* within) {
  HOPB_REQUIRE(h, g && solvent && counts_bulk && F >= 0 && Nw >= 0 && Np >= 0 && (Np == 0 || solute), "bad argument");
  HOPB_REQUIRE(h, cutoff >= 0.0, "cutoff must be >= 0");
  if (F == 0 || Nw == 0) return HOPB_OK;
  HOPB_REQUIRE(h, F < ((int64_t)1 << 31), "too many frames for one call");
  const bool fast = g->tab.fast_ok != 0;
  const size_t smem_fixed = (size_t)g->tab.table2_len * sizeof(HopbPair) + (fast ? 0 : (size_t)g->tab.table_len * sizeof(float)) +
                            (size_t)(2 * kBulkMaxCells + 1) * sizeof(int) + 16;
  // solute atoms per launch: what fits beside the tables in shared memory; larger selections go through in tiles, the
  // "within" flag of every solvent atom carried from tile to tile (the kd-tree of the reference has no such limit)
  const size_t kSmemBudget = 200 * 1024;
  HOPB_REQUIRE(h, smem_fixed + 3 * sizeof(float) * 64 <= kSmemBudget, "grid tables too large for the bulk-selection kernel");
  int64_t tile = (int64_t)((kSmemBudget - smem_fixed) / (3 * sizeof(float)));
  if (const char* e = getenv("HOPB_BULK_TILE")) {          // testing hook: smaller tiles
    const int64_t v = atoll(e);
    if (v > 0 && v < tile) tile = v;
  }
  const int64_t n_tiles = Np > tile ? (Np + tile - 1) / tile : 1;
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* flags = within;
  if (n_tiles > 1 && !flags) {
    int rc = hopb_scratch(h, 15, (size_t)F * Nw, (void**)&flags);
    if (rc) return rc;
  }
#define HOPB_BULK_LAUNCH(C, FA, P0, NP, WIN, WOUT, ONLY)                                                             \
  do {                                                                                                               \
    const size_t smem = smem_fixed + (size_t)(NP) * 3 * sizeof(float);                                               \
    if (smem > 48 * 1024)                                                                                            \
      HOPB_CUDA(h, cudaFuncSetAttribute(hist_bulk_kernel<C, FA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    hopb_note_launch();                                                                                              \
    hist_bulk_kernel<C, FA><<<(unsigned)F, kBulkThreads, smem, s>>>(solvent, (P0), Nw, (int)(NP), (float)cutoff, g->tab, \
                                                                   g->d_table, g->d_table2, counts_solvent,          \
                                                                   counts_bulk, cells, (WOUT), (WIN), Np, (ONLY));   \
  } while (0)
  for (int64_t t = 0; t < n_tiles; ++t) {
    const float* p0 = solute ? solute + t * tile * 3 : nullptr;
    const int64_t np = n_tiles == 1 ? Np : (t + 1 < n_tiles ? tile : Np - t * tile);
    const uint8_t* win = t > 0 ? flags : nullptr;
    const int only = t + 1 < n_tiles ? 1 : 0;
    uint8_t* wout = only ? flags : within;
    if (cells) { if (fast) HOPB_BULK_LAUNCH(true, true, p0, np, win, wout, only); else HOPB_BULK_LAUNCH(true, false, p0, np, win, wout, only); }
    else { if (fast) HOPB_BULK_LAUNCH(false, true, p0, np, win, wout, only); else HOPB_BULK_LAUNCH(false, false, p0, np, win, wout, only); }
  }
#undef HOPB_BULK_LAUNCH
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hist_accumulate_multi(hopb_handle* h, void* stream, int n_species, const hopb_grid* const* grids,
                               const uint8_t* species, const float* xyz, int64_t F, int64_t N, uint32_t* const* counts,
                               uint32_t* cells) {
  HOPB_REQUIRE(h, grids && species && xyz && counts && F >= 0 && N > 0, "bad argument");
  HOPB_REQUIRE(h, n_species >= 1 && n_species <= kMaxSpecies, "1 to 8 species");
  if (F == 0) return HOPB_OK;
  MultiGrids mg;
  memset(&mg, 0, sizeof(mg));
  mg.n = n_species;
  size_t smem = 0;
  for (int k = 0; k < n_species; ++k) {
    HOPB_REQUIRE(h, grids[k] && counts[k], "null grid or count array");
    HOPB_REQUIRE(h, grids[k]->tab.fast_ok != 0, "hopb_hist_accumulate_multi needs regular (linspace) edges");
    const int64_t nbk = (int64_t)((grids[k]->n[0] + 3) / 4) * ((grids[k]->n[1] + 3) / 4) * ((grids[k]->n[2] + 3) / 4);
    HOPB_REQUIRE(h, nbk * 64 < ((int64_t)1 << 30), "grid too large for 30-bit brick indices");
    mg.tab[k] = grids[k]->tab;
    mg.table2[k] = grids[k]->d_table2;
    mg.counts[k] = counts[k];
    smem += (size_t)grids[k]->tab.table2_len * sizeof(HopbPair);
  }
  smem += kWarpsPerBlock * 96 * sizeof(float4);
  HOPB_REQUIRE(h, smem <= 48 * 1024, "grids too large for the shared-memory tables of hopb_hist_accumulate_multi");
  HOPB_REQUIRE(h, ((uintptr_t)xyz & 3) == 0 && N < ((int64_t)1 << 31), "bad coordinate array");
  MultiGrids* d_mg = nullptr;
  int rc = hopb_scratch(h, 17, sizeof(MultiGrids), (void**)&d_mg);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  HOPB_CUDA(h, cudaMemcpyAsync(d_mg, &mg, sizeof(mg), cudaMemcpyHostToDevice, s));
  HOPB_CUDA(h, cudaStreamSynchronize(s));      // mg lives on this stack frame
  const int64_t n_points = F * N;
  int64_t lead = 0;
  while (lead < n_points && (((uintptr_t)xyz + 12 * (uintptr_t)lead) & 15) != 0) ++lead;
  const int64_t nblk = (n_points - lead) / 128;
  int64_t blocks = (nblk + (int64_t)kWarpsPerBlock * 4 - 1) / ((int64_t)kWarpsPerBlock * 4);
  if (blocks < 1) blocks = 1;
  HOPB_K(hist_multi_kernel)<<<(unsigned)blocks, kHistThreads, smem, s>>>(xyz, n_points, lead, N, species, d_mg, cells);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hist_accumulate(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz, int64_t n_points,
                         uint32_t* counts) {
  return hopb_hist_accumulate_cells(h, stream, g, xyz, n_points, counts, nullptr);
}

int hopb_density_from_counts_ex(hopb_handle* h, void* stream, const hopb_grid* g, const uint32_t* counts,
                                int64_t n_frames, double factor, double* density, int ctas_per_sm) {
  HOPB_REQUIRE(h, g && counts && density && n_frames > 0 && ctas_per_sm >= 0, "bad argument");
  const int64_t n_rows = (int64_t)g->n[0] * g->n[1];
  unsigned grid = hopb_grid_for(n_rows * 32, 256, h->num_sms, 16);
  if (ctas_per_sm > 0 && grid > (unsigned)(ctas_per_sm * h->num_sms)) grid = (unsigned)(ctas_per_sm * h->num_sms);
  const int n_classes = (g->n_cls[0] > 0 && g->n_cls[1] > 0 && g->n_cls[2] > 0) ? g->n_cls[0] * g->n_cls[1] * g->n_cls[2] : 0;
  static const bool no_table = getenv("HOPB_DENSITY_TABLE") && atoi(getenv("HOPB_DENSITY_TABLE")) == 0;  // tuning hook
  if (n_classes > 0 && n_classes <= 512 && !no_table && n_rows * g->n[2] >= (int64_t)n_classes * kDensityTableCounts) {
    double* table = nullptr;
    int rc = hopb_scratch(h, 10, (size_t)n_classes * kDensityTableCounts * sizeof(double), (void**)&table);
    if (rc != HOPB_OK) return rc;
    const double* w = g->d_cls_width;
    const int n_entries = n_classes * kDensityTableCounts;
    HOPB_K(hopb_tables::density_table_kernel)<<<(n_entries + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        n_classes, g->n_cls[1], g->n_cls[2], w, w + 16, w + 32, (double)n_frames, factor, factor != 1.0 ? 1 : 0, table);
    HOPB_K(density_lookup_kernel)<<<grid, 256, 0, (cudaStream_t)stream>>>(
        counts, n_rows, g->n[1], g->n[2], (double)n_frames, g->d_width[0], g->d_width[1], g->d_width[2], factor,
        factor != 1.0 ? 1 : 0, g->d_cls[0], g->d_cls[1], g->d_cls[2], g->n_cls[1], g->n_cls[2], table, density);
    HOPB_LAUNCH_CHECK(h);
    return HOPB_OK;
  }
  HOPB_K(density_kernel)<<<grid, 256, 0, (cudaStream_t)stream>>>(
      counts, n_rows, g->n[1], g->n[2], (double)n_frames, g->d_width[0], g->d_width[1], g->d_width[2], factor,
      factor != 1.0 ? 1 : 0, density);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_density_from_counts(hopb_handle* h, void* stream, const hopb_grid* g, const uint32_t* counts,
                             int64_t n_frames, double factor, double* density) {
  return hopb_density_from_counts_ex(h, stream, g, counts, n_frames, factor, density, 0);
}

}  // extern "C"
