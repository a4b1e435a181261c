// Stage 4 output: order the transition events the way the reference emits them and reduce them to
// the sparse transition-count matrix.
//
// TransportNetwork._analyze_hops appends to graph_properties[(s0, s)] while looping frames (outer)
// and atoms (inner) (hop/graph.py:212-247), so every per-edge list is ordered by (frame, iatom).
// The kernels and the fix-up produce events in arbitrary order; here they are sorted by
// (edge, frame, iatom) with an LSD radix sort -- first on frame*N + iatom, then (stable) on the
// edge key -- and run-length encoded into COO form: (s0, s) -> count = len(tau), the `N` that
// HoppingGraph._rate (hop/graph.py:1728,1748) and MCMC (hop/MCMC.py:112-118) consume.
#include "hopb_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
ev_key1_kernel(const hopb_event* __restrict__ ev, int64_t n, int64_t N, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = (uint64_t)ev[i].frame * (uint64_t)N + (uint64_t)ev[i].iatom;
    vals[i] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(256)
ev_key2_kernel(const hopb_event* __restrict__ ev, const uint32_t* __restrict__ perm, int64_t n, int n_labels,
               uint64_t* __restrict__ keys) {
  const uint64_t W = (uint64_t)n_labels + 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const hopb_event e = ev[perm[i]];
    keys[i] = (uint64_t)(e.s0 + 1) * W + (uint64_t)(e.s + 1);
  }
}

__global__ void __launch_bounds__(256)
ev_gather_kernel(const hopb_event* __restrict__ ev, const uint32_t* __restrict__ perm, const uint64_t* __restrict__ keys,
                 int64_t n, hopb_event* __restrict__ out, uint32_t* __restrict__ head) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    out[i] = ev[perm[i]];
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
  }
}

__global__ void __launch_bounds__(256)
ev_edges_kernel(const hopb_event* __restrict__ sorted, const uint32_t* __restrict__ head, const uint32_t* __restrict__ rank,
                int64_t n, int32_t* __restrict__ edge_s0, int32_t* __restrict__ edge_s, int64_t* __restrict__ edge_start) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (head[i]) {
      const uint32_t e = rank[i];
      edge_s0[e] = sorted[i].s0;
      edge_s[e] = sorted[i].s;
      edge_start[e] = i;
    }
  }
}

__global__ void __launch_bounds__(256)
ev_counts_kernel(const int64_t* __restrict__ edge_start, int64_t n_edges, int64_t n, int64_t* __restrict__ edge_count) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x)
    edge_count[e] = (e + 1 < n_edges ? edge_start[e + 1] : n) - edge_start[e];
}

}  // namespace

extern "C" int hopb_events_finalize(hopb_handle* h, void* stream, const hopb_event* events, uint64_t n_u, int64_t N,
                                    int64_t F_total, int n_labels, hopb_event* events_sorted, int32_t* edge_s0,
                                    int32_t* edge_s, int64_t* edge_start, int64_t* edge_count, int64_t* n_edges) {
  HOPB_REQUIRE(h, n_edges != nullptr, "null argument");
  *n_edges = 0;
  const int64_t n = (int64_t)n_u;
  if (n == 0) return HOPB_OK;
  HOPB_REQUIRE(h, events && events_sorted && edge_s0 && edge_s && edge_start && edge_count, "null argument");
  HOPB_REQUIRE(h, n < ((int64_t)1 << 32) && N > 0 && F_total > 0 && n_labels >= 1, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* buf = nullptr;
  const size_t per = 2 * sizeof(uint64_t) + 4 * sizeof(uint32_t);
  int rc = hopb_scratch(h, 5, (size_t)(n + 2) * per + 64, (void**)&buf);
  if (rc) return rc;
  uint64_t* keys = (uint64_t*)buf;
  uint64_t* keys_alt = keys + (n + 2);
  uint32_t* vals = (uint32_t*)(keys_alt + (n + 2));
  uint32_t* vals_alt = vals + (n + 2);
  uint32_t* head = vals_alt + (n + 2);
  uint32_t* rank = head + (n + 2);
  uint32_t* total_dev = rank + n;  // inside the padding
  const unsigned g = hopb_grid_for(n, 256, h->num_sms, 8);

  HOPB_K(ev_key1_kernel)<<<g, 256, 0, s>>>(events, n, N, keys, vals);
  int alt = 0;
  rc = hopb_radix_sort_pairs(h, s, keys, keys_alt, vals, vals_alt, n, 0, hopb_bits_for((uint64_t)F_total * (uint64_t)N), &alt);
  if (rc) return rc;
  uint64_t* k = alt ? keys_alt : keys; uint64_t* ka = alt ? keys : keys_alt;
  uint32_t* v = alt ? vals_alt : vals; uint32_t* va = alt ? vals : vals_alt;
  HOPB_K(ev_key2_kernel)<<<g, 256, 0, s>>>(events, v, n, n_labels, k);
  int alt2 = 0;
  const uint64_t W = (uint64_t)n_labels + 1;
  rc = hopb_radix_sort_pairs(h, s, k, ka, v, va, n, 0, hopb_bits_for(W * W), &alt2);
  if (rc) return rc;
  const uint64_t* kf = alt2 ? ka : k;
  const uint32_t* vf = alt2 ? va : v;
  HOPB_K(ev_gather_kernel)<<<g, 256, 0, s>>>(events, vf, kf, n, events_sorted, head);
  rc = hopb_exclusive_scan_u32(h, s, head, rank, n, total_dev);
  if (rc) return rc;
  HOPB_K(ev_edges_kernel)<<<g, 256, 0, s>>>(events_sorted, head, rank, n, edge_s0, edge_s, edge_start);
  uint32_t total = 0;
  HOPB_CUDA(h, cudaMemcpyAsync(&total, total_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  HOPB_CUDA(h, cudaStreamSynchronize(s));
  *n_edges = (int64_t)total;
  HOPB_K(ev_counts_kernel)<<<hopb_grid_for((int64_t)total, 256, h->num_sms, 8), 256, 0, s>>>(edge_start, (int64_t)total, n, edge_count);
  HOPB_LAUNCH_CHECK(h);
  HOPB_CUDA(h, cudaStreamSynchronize(s));
  return HOPB_OK;
}
