// Stage 4 output: order the transition events the way the reference emits them and reduce them to
// the sparse transition-count matrix.
//
// TransportNetwork._analyze_hops appends to graph_properties[(s0, s)] while looping frames (outer)
// and atoms (inner) (hop/graph.py:212-247), so every per-edge list is ordered by (frame, iatom).
// The kernels and the fix-up produce events in arbitrary order; here they are sorted by
// (edge, frame, iatom) with a stable LSD radix sort -- first on frame*N + iatom, then on the edge
// key -- and run-length encoded into COO form: (s0, s) -> count = len(tau), the `N` that
// HoppingGraph._rate (hop/graph.py:1728,1748) and MCMC (hop/MCMC.py:112-118) consume.
//
// There are far fewer events than atom-frames (~1e-3), so the ~30 small passes this takes are
// launch-latency bound when issued one by one.  The whole finalisation is therefore ONE cooperative
// kernel: every pass is a phase separated by a grid-wide barrier.
#include <cooperative_groups.h>

#include "hopb_sort_device.cuh"

namespace cg = cooperative_groups;

namespace {

using namespace hopb_sort;

struct EvArgs {
  const hopb_event* ev;
  int64_t n, N;
  int n_labels, bits1, bits2, tiles_per_block;
  uint64_t* k[2];
  uint32_t* v[2];
  uint32_t* ghist;     // [256][grid]
  uint32_t* totals;    // [256]
  uint32_t* blockcnt;  // [grid + 1]
  hopb_event* sorted;
  int32_t* e_s0;
  int32_t* e_s;
  int64_t* e_start;
  int64_t* e_count;
  int32_t* matrix;     // optional dense count matrix, row stride matrix_stride (0: n_labels + 1, caller-zeroed)
  int64_t* n_edges_dev;
  // device-side arguments (hopb_events_finalize_dev): n and n_labels are read from device memory,
  // so the host need not wait for stage 3 before it launches this kernel
  const unsigned long long* n_dev;  // number of events (NULL: use n)
  int64_t cap;                      // capacity of the event buffer
  const int32_t* n_sites_dev;       // number of hydration sites; n_labels = *n_sites_dev + label_offset
  int label_offset, max_sites, matrix_stride;
  int64_t* result_dev;              // [4]: n_events, n_edges, n_labels, status
};

__device__ __forceinline__ int coop_sort(cg::grid_group& grid, const EvArgs& a, int cur, int bits, SortSmem& sm) {
  for (int shift = 0; shift < bits; shift += 8) {
    sort_hist_block(a.k[cur], a.n, shift, a.tiles_per_block, a.ghist, sm);
    grid.sync();
    sort_rowscan_rows(a.ghist, (int)gridDim.x, a.totals, sm);
    grid.sync();
    sort_scatter_block(a.k[cur], a.v[cur], a.k[cur ^ 1], a.v[cur ^ 1], a.n, shift, a.tiles_per_block, a.ghist, a.totals, sm);
    grid.sync();
    cur ^= 1;
  }
  return cur;
}

__global__ void __launch_bounds__(kSortThreads)
ev_finalize_kernel(EvArgs a) {
  __shared__ SortSmem sm;
  __shared__ uint32_t s_carry;
  cg::grid_group grid = cg::this_grid();
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  if (a.n_dev) {
    // every thread derives the same launch parameters from the device scalars
    const unsigned long long n_all = *a.n_dev;
    const int n_sites = *a.n_sites_dev;
    int status = 0;
    if (n_all > (unsigned long long)a.cap) status |= HOPB_FIN_OVERFLOW;   // stage 3 must be redone with a larger buffer
    if (n_sites > a.max_sites) status |= HOPB_FIN_TOO_MANY_SITES;
    a.n = status ? 0 : (int64_t)n_all;
    a.n_labels = n_sites + a.label_offset;
    const uint64_t W = (uint64_t)a.n_labels + 1;
    a.bits2 = hopb_bits_for(W * W);
    if (a.matrix && (int64_t)W > a.matrix_stride) { status |= HOPB_FIN_MATRIX_SKIPPED; a.matrix = nullptr; }
    const int64_t ntiles = (a.n + kSortTile - 1) / kSortTile;
    a.tiles_per_block = (int)((ntiles + gridDim.x - 1) / gridDim.x);
    if (a.tiles_per_block < 1) a.tiles_per_block = 1;
    if (gtid == 0) {
      a.result_dev[0] = (int64_t)n_all; a.result_dev[1] = 0; a.result_dev[2] = a.n_labels; a.result_dev[3] = status;
    }
    if (a.matrix) {  // the whole buffer: it is all-reduced across ranks as it is
      const int64_t cells = (int64_t)a.matrix_stride * a.matrix_stride;
      for (int64_t i = gtid; i < cells; i += gstride) a.matrix[i] = 0;
    }
    if (a.n == 0) return;   // uniform: every thread of the grid leaves before the first barrier
  }
  const int64_t n = a.n;

  // key 1: position in the reference's loop order
  for (int64_t i = gtid; i < n; i += gstride) {
    a.k[0][i] = (uint64_t)a.ev[i].frame * (uint64_t)a.N + (uint64_t)a.ev[i].iatom;
    a.v[0][i] = (uint32_t)i;
  }
  grid.sync();
  int cur = coop_sort(grid, a, 0, a.bits1, sm);
  // key 2: the edge
  const uint64_t W = (uint64_t)a.n_labels + 1;
  for (int64_t i = gtid; i < n; i += gstride) {
    const hopb_event e = a.ev[a.v[cur][i]];
    a.k[cur][i] = (uint64_t)(e.s0 + 1) * W + (uint64_t)(e.s + 1);
  }
  grid.sync();
  cur = coop_sort(grid, a, cur, a.bits2, sm);
  const uint64_t* keys = a.k[cur];
  const uint32_t* perm = a.v[cur];

  // gather into the final order; count edge heads per block (blocks own contiguous ranges)
  const int64_t begin = (int64_t)blockIdx.x * a.tiles_per_block * kSortTile;
  int64_t end = begin + (int64_t)a.tiles_per_block * kSortTile;
  if (end > n) end = n;
  uint32_t heads = 0;
  for (int64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    a.sorted[i] = a.ev[perm[i]];
    heads += (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
  }
  uint32_t total;
  block_exclusive_scan(heads, sm.warp, &total);
  if (threadIdx.x == 0) a.blockcnt[blockIdx.x] = total;
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    uint32_t run = 0;
    for (unsigned b = 0; b < gridDim.x; ++b) { const uint32_t c = a.blockcnt[b]; a.blockcnt[b] = run; run += c; }
    a.blockcnt[gridDim.x] = run;
    *a.n_edges_dev = (int64_t)run;
    if (a.result_dev) a.result_dev[1] = (int64_t)run;
  }
  grid.sync();
  // edge records: rank of every head = block base + ordered rank inside the block
  if (threadIdx.x == 0) s_carry = a.blockcnt[blockIdx.x];
  __syncthreads();
  for (int64_t i0 = begin; i0 < end; i0 += blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    const bool head = i < end && (i == 0 || keys[i] != keys[i - 1]);
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan(head ? 1u : 0u, sm.warp, &tot);
    if (head) {
      const uint32_t e = s_carry + ex;
      a.e_s0[e] = a.sorted[i].s0;
      a.e_s[e] = a.sorted[i].s;
      a.e_start[e] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += tot;
    __syncthreads();
  }
  grid.sync();
  const int64_t n_edges = (int64_t)a.blockcnt[gridDim.x];
  for (int64_t e = gtid; e < n_edges; e += gstride) {
    const int64_t c = (e + 1 < n_edges ? a.e_start[e + 1] : n) - a.e_start[e];
    a.e_count[e] = c;
    const int64_t stride = a.matrix_stride ? (int64_t)a.matrix_stride : (int64_t)W;
    if (a.matrix) a.matrix[(int64_t)(a.e_s0[e] + 1) * stride + (a.e_s[e] + 1)] = (int32_t)c;
  }
}

}  // namespace

extern "C" int hopb_events_finalize_ex(hopb_handle* h, void* stream, const hopb_event* events, uint64_t n_u, int64_t N,
                                       int64_t F_total, int n_labels, hopb_event* events_sorted, int32_t* edge_s0,
                                       int32_t* edge_s, int64_t* edge_start, int64_t* edge_count, int32_t* count_matrix,
                                       int64_t* n_edges) {
  HOPB_REQUIRE(h, n_edges != nullptr, "null argument");
  *n_edges = 0;
  const int64_t n = (int64_t)n_u;
  if (n == 0) return HOPB_OK;
  HOPB_REQUIRE(h, events && events_sorted && edge_s0 && edge_s && edge_start && edge_count, "null argument");
  HOPB_REQUIRE(h, n < ((int64_t)1 << 32) && N > 0 && F_total > 0 && n_labels >= 1, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;

  int per_sm = 0;
  HOPB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ev_finalize_kernel, kSortThreads, 0));
  int64_t max_blocks = (int64_t)per_sm * h->num_sms;
  if (max_blocks > kMaxSortBlocks) max_blocks = kMaxSortBlocks;
  HOPB_REQUIRE(h, max_blocks >= 1, "cooperative launch not possible");
  const int64_t ntiles = (n + kSortTile - 1) / kSortTile;
  const int tiles_per_block = (int)((ntiles + max_blocks - 1) / max_blocks);
  const int nblocks = (int)((ntiles + tiles_per_block - 1) / tiles_per_block);

  unsigned char* buf = nullptr;
  const size_t per = 2 * sizeof(uint64_t) + 2 * sizeof(uint32_t);
  const size_t bytes = (size_t)(n + 2) * per + ((size_t)256 * nblocks + 256 + nblocks + 8) * sizeof(uint32_t) + 64;
  int rc = hopb_scratch(h, 5, bytes, (void**)&buf);
  if (rc) return rc;
  EvArgs a;
  a.ev = events; a.n = n; a.N = N; a.n_labels = n_labels;
  a.bits1 = hopb_bits_for((uint64_t)F_total * (uint64_t)N);
  const uint64_t W = (uint64_t)n_labels + 1;
  a.bits2 = hopb_bits_for(W * W);
  a.tiles_per_block = tiles_per_block;
  a.k[0] = (uint64_t*)buf;
  a.k[1] = a.k[0] + (n + 2);
  a.v[0] = (uint32_t*)(a.k[1] + (n + 2));
  a.v[1] = a.v[0] + (n + 2);
  a.ghist = a.v[1] + (n + 2);
  a.totals = a.ghist + (size_t)256 * nblocks;
  a.blockcnt = a.totals + 256;
  int64_t* n_edges_dev = (int64_t*)(((uintptr_t)(a.blockcnt + nblocks + 2) + 7) & ~(uintptr_t)7);
  a.n_edges_dev = n_edges_dev;
  a.sorted = events_sorted; a.e_s0 = edge_s0; a.e_s = edge_s; a.e_start = edge_start; a.e_count = edge_count;
  a.matrix = count_matrix;
  a.n_dev = nullptr; a.cap = n; a.n_sites_dev = nullptr; a.label_offset = 0; a.max_sites = 0; a.matrix_stride = 0;
  a.result_dev = nullptr;
  void* kargs[] = {(void*)&a};
  hopb_note_launch();
  HOPB_CUDA(h, cudaLaunchCooperativeKernel((void*)ev_finalize_kernel, dim3((unsigned)nblocks), dim3(kSortThreads), kargs, 0, s));
  int64_t total = 0;
  HOPB_CUDA(h, cudaMemcpyAsync(&total, n_edges_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  HOPB_CUDA(h, cudaStreamSynchronize(s));
  *n_edges = total;
  return HOPB_OK;
}

extern "C" int hopb_events_finalize_dev(hopb_handle* h, void* stream, const hopb_event* events,
                                        const uint64_t* n_events_dev, uint64_t cap_u, int64_t N, int64_t F_total,
                                        const int32_t* n_sites_dev, int label_offset, hopb_event* events_sorted,
                                        int32_t* edge_s0, int32_t* edge_s, int64_t* edge_start, int64_t* edge_count,
                                        int32_t* count_matrix, int matrix_stride, int64_t* result_dev) {
  HOPB_REQUIRE(h, events && n_events_dev && n_sites_dev && events_sorted && edge_s0 && edge_s && edge_start && edge_count &&
                      result_dev, "null argument");
  const int64_t cap = (int64_t)cap_u;
  HOPB_REQUIRE(h, cap >= 1 && cap < ((int64_t)1 << 32) && N > 0 && F_total > 0 && label_offset >= 1, "bad argument");
  HOPB_REQUIRE(h, count_matrix == nullptr || matrix_stride >= 2, "bad matrix stride");
  cudaStream_t s = (cudaStream_t)stream;
  int per_sm = 0;
  HOPB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ev_finalize_kernel, kSortThreads, 0));
  HOPB_REQUIRE(h, per_sm >= 1, "cooperative launch not possible");
  // This variant runs on a side stream beside the next run's histogram.  Its CTAs spend most of their time spinning
  // at grid barriers, and with one CTA on every SM that costs the histogram issue slots everywhere: measured on B200
  // (bench shape, 2.3e5 events) 148 CTAs: S1 1.889 / S4 0.20 ms, step 3.133; 32 CTAs: S1 1.827 / S4 0.73 ms (still
  // hidden behind S1), step 3.071; 16 CTAs: 1.796 / 1.58; 8 CTAs: S4 outlasts the histogram and stalls stage 3.
  int nblocks = h->num_sms < kMaxSortBlocks ? h->num_sms : kMaxSortBlocks;
  int want = 32;
  if (const char* e = getenv("HOPB_EV_BLOCKS")) want = atoi(e);     // tuning hook
  if (want >= 1 && want < nblocks) nblocks = want;
  unsigned char* buf = nullptr;
  const size_t per = 2 * sizeof(uint64_t) + 2 * sizeof(uint32_t);
  const size_t bytes = (size_t)(cap + 2) * per + ((size_t)256 * nblocks + 256 + nblocks + 8) * sizeof(uint32_t) + 64;
  int rc = hopb_scratch(h, 5, bytes, (void**)&buf);
  if (rc) return rc;
  EvArgs a;
  a.ev = events; a.n = 0; a.N = N; a.n_labels = 0; a.bits2 = 0; a.tiles_per_block = 1;
  a.bits1 = hopb_bits_for((uint64_t)F_total * (uint64_t)N);
  a.k[0] = (uint64_t*)buf;
  a.k[1] = a.k[0] + (cap + 2);
  a.v[0] = (uint32_t*)(a.k[1] + (cap + 2));
  a.v[1] = a.v[0] + (cap + 2);
  a.ghist = a.v[1] + (cap + 2);
  a.totals = a.ghist + (size_t)256 * nblocks;
  a.blockcnt = a.totals + 256;
  a.n_edges_dev = (int64_t*)(((uintptr_t)(a.blockcnt + nblocks + 2) + 7) & ~(uintptr_t)7);
  a.sorted = events_sorted; a.e_s0 = edge_s0; a.e_s = edge_s; a.e_start = edge_start; a.e_count = edge_count;
  a.matrix = count_matrix;
  a.n_dev = (const unsigned long long*)n_events_dev; a.cap = cap; a.n_sites_dev = n_sites_dev;
  a.label_offset = label_offset; a.max_sites = HOPB_MAX_SITES; a.matrix_stride = count_matrix ? matrix_stride : 0;
  a.result_dev = result_dev;
  void* kargs[] = {(void*)&a};
  hopb_note_launch();
  HOPB_CUDA(h, cudaLaunchCooperativeKernel((void*)ev_finalize_kernel, dim3((unsigned)nblocks), dim3(kSortThreads), kargs, 0, s));
  return HOPB_OK;
}

extern "C" int hopb_events_finalize(hopb_handle* h, void* stream, const hopb_event* events, uint64_t n_u, int64_t N,
                                    int64_t F_total, int n_labels, hopb_event* events_sorted, int32_t* edge_s0,
                                    int32_t* edge_s, int64_t* edge_start, int64_t* edge_count, int64_t* n_edges) {
  return hopb_events_finalize_ex(h, stream, events, n_u, N, F_total, n_labels, events_sorted, edge_s0, edge_s, edge_start,
                                 edge_count, nullptr, n_edges);
}
