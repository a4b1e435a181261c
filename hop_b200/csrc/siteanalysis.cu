// Per-site observables: what hop.siteanalysis.Collection.compute accumulates frame by
// frame in Python (hop/siteanalysis.py:1055-1107) from the coordinates and the hopping trajectory.
//
//   occupancy   number of atoms on every site in every frame -- Occupancy._compute_occupancy
//               (hop/siteanalysis.py:837-841) takes numpy.histogram of the frame's site labels
//   distance    distance of every atom from the centre of the site it is on (Distance, :753-776) or orbits
//               (Orbit, :801-802: the same with the y column of the hoptraj), histogrammed per site by
//               SiteHistogramArray.add (:531-544)
//
// Reference behaviours kept on purpose (they are what the reference run in tests/golden produces):
//  * SiteHistogramArray.add increments through a fancy index, `_histogram[rows, bins] += 1`: a (site, bin) pair
//    that several atoms hit in the same frame counts ONCE for that frame;
//  * labels index Python arrays directly, so the outlier label -1 reads the LAST entry of site_centers and of
//    site2indx -- an outlier is treated as if it sat on the highest-numbered site;
//  * the interstitial has centre NaN; digitize puts NaN into the upper outlier bin.
#include "hopb_common.cuh"

namespace {

// ---- occupancy: one CTA per frame, counters in shared memory -----------------------------------------------
__global__ void __launch_bounds__(512)
site_occupancy_kernel(const float* __restrict__ labels, int64_t F, int64_t N, int n_labels, int32_t* __restrict__ occ) {
  extern __shared__ uint32_t s_cnt[];
  for (int64_t f = blockIdx.x; f < F; f += gridDim.x) {
    for (int l = threadIdx.x; l < n_labels; l += blockDim.x) s_cnt[l] = 0u;
    __syncthreads();
    const float* row = labels + f * N;
    const unsigned full = 0xffffffffu;
    for (int64_t i0 = 0; i0 < N; i0 += blockDim.x) {
      const int64_t i = i0 + threadIdx.x;
      const int l = i < N ? (int)row[i] : -1;
      // most atoms sit in the interstitial or the bulk site: those two go through a ballot
      const unsigned b0 = __ballot_sync(full, l == 0), b1 = __ballot_sync(full, l == 1);
      if ((threadIdx.x & 31) == 0) {
        if (b0) atomicAdd(&s_cnt[0], (uint32_t)__popc(b0));
        if (b1 && n_labels > 1) atomicAdd(&s_cnt[1], (uint32_t)__popc(b1));
      }
      if (l >= 2 && l < n_labels) atomicAdd(&s_cnt[l], 1u);
    }
    __syncthreads();
    for (int l = threadIdx.x; l < n_labels; l += blockDim.x) occ[f * n_labels + l] = (int32_t)s_cnt[l];
    __syncthreads();
  }
}

// ---- distance / orbit histograms ----------------------------------------------------------------------------
struct DistArgs {
  const float* xyz;        // [F][N][3] coordinates of the selected atoms
  const float* labels;     // [F][Nh] hoptraj x (distance) or y (orbit)
  const int32_t* amap;     // [N] index of the selected atom's hoptraj column, NULL: identity
  int64_t N, Nh;
  int64_t f0, f1, fstep;   // frames f0, f0 + fstep, ... < f1
  const double* centers;   // [L][3], NaN for the interstitial
  const int32_t* site2indx;  // [L] row of every label (unselected sites: n_rows - 1)
  int L;
  const double* edges;     // [nbins + 1]
  int nbins;
  int n_rows;              // Nsites + 1
  uint32_t* bitmap;        // [gridDim.x][words]: (row, bin) pairs seen in the frame a CTA is working on (all zero between frames)
  int64_t words;
  unsigned long long* hist;  // [n_rows][nbins + 2]
};

__device__ __forceinline__ int64_t dist_key(const DistArgs& a, const double* s_edges, int64_t frame, int64_t i) {
  const int64_t col = a.amap ? a.amap[i] : i;
  int l = (int)a.labels[frame * a.Nh + col];
  if (l < 0) l += a.L;                 // Python's negative index: the outlier reads the last site's entries
  if (l < 0 || l >= a.L) return -1;    // (the reference would raise IndexError; cannot happen for a consistent density)
  const float* p = a.xyz + (frame * a.N + i) * 3;
  const double* c = a.centers + (int64_t)l * 3;
  const double dx = __dsub_rn(c[0], (double)p[0]), dy = __dsub_rn(c[1], (double)p[1]), dz = __dsub_rn(c[2], (double)p[2]);
  const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
  // numpy.digitize(d, edges): number of edges <= d; NaN sorts behind everything
  int lo = 0, hi = a.nbins + 1;
  if (d != d) lo = hi;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s_edges[mid] <= d) lo = mid + 1; else hi = mid;
  }
  return (int64_t)a.site2indx[l] * (a.nbins + 2) + lo;
}

__global__ void __launch_bounds__(256) site_distance_kernel(DistArgs a) {
  extern __shared__ double s_edges[];
  for (int k = threadIdx.x; k <= a.nbins; k += blockDim.x) s_edges[k] = a.edges[k];
  __syncthreads();
  uint32_t* bm = a.bitmap + (int64_t)blockIdx.x * a.words;
  const int64_t n_frames = (a.f1 - a.f0 + a.fstep - 1) / a.fstep;
  for (int64_t k = blockIdx.x; k < n_frames; k += gridDim.x) {
    const int64_t frame = a.f0 + k * a.fstep;
    for (int64_t i = threadIdx.x; i < a.N; i += blockDim.x) {
      const int64_t key = dist_key(a, s_edges, frame, i);
      if (key < 0) continue;
      uint32_t* w = bm + (key >> 5);
      const uint32_t bit = 1u << (key & 31);
      // first atom of the frame on this (site, bin): count it; the others of the frame do not (fancy-index +=)
      if (!(*reinterpret_cast<volatile uint32_t*>(w) & bit) && !(atomicOr(w, bit) & bit)) atomicAdd(a.hist + key, 1ull);
    }
    __syncthreads();
    for (int64_t i = threadIdx.x; i < a.N; i += blockDim.x) {      // leave the bitmap clean for the next frame
      const int64_t key = dist_key(a, s_edges, frame, i);
      if (key >= 0) bm[key >> 5] = 0u;
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" {

int hopb_site_occupancy(hopb_handle* h, void* stream, const float* labels, int64_t F, int64_t N, int n_labels, int32_t* occ) {
  HOPB_REQUIRE(h, labels && occ && F >= 0 && N >= 0, "bad argument");
  HOPB_REQUIRE(h, n_labels >= 1 && n_labels <= HOPB_MAX_SITES + 2, "bad n_labels");
  if (F == 0) return HOPB_OK;
  const size_t smem = (size_t)n_labels * sizeof(uint32_t);
  if (smem > 48 * 1024)
    HOPB_CUDA(h, cudaFuncSetAttribute(site_occupancy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = F < (int64_t)h->num_sms * 4 ? F : (int64_t)h->num_sms * 4;
  HOPB_K(site_occupancy_kernel)<<<(unsigned)blocks, 512, smem, (cudaStream_t)stream>>>(labels, F, N, n_labels, occ);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_site_distance_hist(hopb_handle* h, void* stream, const float* xyz, const float* labels, const int32_t* atom_map,
                            int64_t N, int64_t N_hop, int64_t frame_start, int64_t frame_stop, int64_t frame_step,
                            const double* centers, const int32_t* site2indx, int n_labels, const double* edges, int n_bins,
                            int n_rows, uint64_t* hist) {
  HOPB_REQUIRE(h, xyz && labels && centers && site2indx && edges && hist, "null argument");
  HOPB_REQUIRE(h, N >= 0 && N_hop >= 0 && frame_step >= 1 && frame_start >= 0 && frame_stop >= frame_start, "bad frame range");
  HOPB_REQUIRE(h, n_labels >= 1 && n_bins >= 1 && n_bins <= 4096 && n_rows >= 1, "bad histogram shape");
  if (frame_stop == frame_start || N == 0) return HOPB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_frames = (frame_stop - frame_start + frame_step - 1) / frame_step;
  int64_t blocks = n_frames < (int64_t)h->num_sms * 4 ? n_frames : (int64_t)h->num_sms * 4;
  const int64_t words = ((int64_t)n_rows * (n_bins + 2) + 31) / 32;
  uint32_t* bitmap = nullptr;
  int rc = hopb_scratch(h, 16, (size_t)blocks * words * sizeof(uint32_t), (void**)&bitmap);
  if (rc) return rc;
  HOPB_CUDA(h, cudaMemsetAsync(bitmap, 0, (size_t)blocks * words * sizeof(uint32_t), s));
  DistArgs a;
  a.xyz = xyz; a.labels = labels; a.amap = atom_map; a.N = N; a.Nh = N_hop;
  a.f0 = frame_start; a.f1 = frame_stop; a.fstep = frame_step;
  a.centers = centers; a.site2indx = site2indx; a.L = n_labels; a.edges = edges; a.nbins = n_bins; a.n_rows = n_rows;
  a.bitmap = bitmap; a.words = words; a.hist = reinterpret_cast<unsigned long long*>(hist);
  HOPB_K(site_distance_kernel)<<<(unsigned)blocks, 256, (size_t)(n_bins + 1) * sizeof(double), s>>>(a);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

}  // extern "C"
