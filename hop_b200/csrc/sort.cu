// Device-wide primitives: exclusive scan and a stable LSD radix sort of (u64 key, u32 value) pairs.
// They serve the site ranking (label.cu), the site cell lists and the event ordering (events.cu);
// all of those work on #sites / #events items, orders of magnitude fewer than atom-frames.

#include "hopb_sort_device.cuh"

namespace {

using namespace hopb_sort;

// ------------------------------------------------------------------------------------------------
// scan
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

// phase 0: tile sums; phase 1: final scan with tile offsets
template <int PHASE>
__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t s_warp[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int64_t idx = base + i;
    v[i] = idx < n ? in[idx] : 0u;
    sum += v[i];
  }
  uint32_t total;
  uint32_t off = block_exclusive_scan(sum, s_warp, &total);
  if (PHASE == 0) {
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
  } else {
    off += tile_sums ? tile_sums[blockIdx.x] : 0u;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      const int64_t idx = base + i;
      if (idx < n) out[idx] = off;
      off += v[i];
    }
  }
}

__global__ void store_total_kernel(const uint32_t* in_last, const uint32_t* out_last, uint32_t* total) {
  *total = *in_last + *out_last;
}

__global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int tiles_per_block, uint32_t* __restrict__ ghist) {
  __shared__ SortSmem sm;
  sort_hist_block(keys, n, shift, tiles_per_block, ghist, sm);
}

__global__ void __launch_bounds__(kSortThreads)
sort_rowscan_kernel(uint32_t* __restrict__ ghist, int nblocks, uint32_t* __restrict__ totals) {
  __shared__ SortSmem sm;
  sort_rowscan_rows(ghist, nblocks, totals, sm);
}

__global__ void __launch_bounds__(kSortThreads)
sort_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t* __restrict__ keys_out,
                    uint32_t* __restrict__ vals_out, int64_t n, int shift, int tiles_per_block,
                    const uint32_t* __restrict__ ghist, const uint32_t* __restrict__ totals) {
  __shared__ SortSmem sm;
  sort_scatter_block(keys, vals, keys_out, vals_out, n, shift, tiles_per_block, ghist, totals, sm);
}

}  // namespace

int hopb_exclusive_scan_u32(hopb_handle* h, cudaStream_t s, const uint32_t* in, uint32_t* out, int64_t n,
                            uint32_t* total_dev) {
  if (n <= 0) {
    if (total_dev) HOPB_CUDA(h, cudaMemsetAsync(total_dev, 0, sizeof(uint32_t), s));
    return HOPB_OK;
  }
  // levels of tile sums live in scratch slot 7
  std::vector<int64_t> sizes;
  int64_t m = n;
  size_t total_tiles = 0;
  while (m > kScanTile) {
    m = (m + kScanTile - 1) / kScanTile;
    sizes.push_back(m);
    total_tiles += (size_t)m;
  }
  uint32_t* pool = nullptr;
  // the last input element is needed for the grand total (in may alias out)
  int rc = hopb_scratch(h, 7, (total_tiles + 4) * sizeof(uint32_t), (void**)&pool);
  if (rc) return rc;
  uint32_t* in_last_copy = pool + total_tiles;
  if (total_dev)
    HOPB_CUDA(h, cudaMemcpyAsync(in_last_copy, in + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  // up-sweep
  std::vector<uint32_t*> level(sizes.size());
  {
    uint32_t* p = pool;
    const uint32_t* src = in;
    int64_t cur = n;
    for (size_t l = 0; l < sizes.size(); ++l) {
      level[l] = p;
      HOPB_K(scan_kernel<0>)<<<(unsigned)sizes[l], kScanThreads, 0, s>>>(src, nullptr, cur, p);
      src = p;
      cur = sizes[l];
      p += sizes[l];
    }
  }
  // down-sweep: scan the top level alone, then each level below with the offsets of the one above
  for (int l = (int)sizes.size() - 1; l >= 0; --l) {
    const int64_t cur = sizes[l];
    const unsigned blocks = (unsigned)((cur + kScanTile - 1) / kScanTile);
    uint32_t* offs = (l + 1 < (int)sizes.size()) ? level[l + 1] : nullptr;
    HOPB_K(scan_kernel<1>)<<<blocks, kScanThreads, 0, s>>>(level[l], level[l], cur, offs);
  }
  {
    const unsigned blocks = (unsigned)((n + kScanTile - 1) / kScanTile);
    HOPB_K(scan_kernel<1>)<<<blocks, kScanThreads, 0, s>>>(in, out, n, sizes.empty() ? nullptr : level[0]);
  }
  if (total_dev) HOPB_K(store_total_kernel)<<<1, 1, 0, s>>>(in_last_copy, out + (n - 1), total_dev);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_radix_sort_pairs(hopb_handle* h, cudaStream_t s, uint64_t* keys, uint64_t* keys_alt, uint32_t* vals,
                          uint32_t* vals_alt, int64_t n, int begin_bit, int end_bit, int* in_alt) {
  *in_alt = 0;
  if (n <= 1 || end_bit <= begin_bit) return HOPB_OK;
  if (n >= ((int64_t)1 << 32)) return hopb_fail(h, HOPB_E_INVALID, "radix sort: n too large");
  const int64_t ntiles = (n + kSortTile - 1) / kSortTile;
  int max_blocks = h->num_sms * 4;
  if (max_blocks > kMaxSortBlocks) max_blocks = kMaxSortBlocks;
  const int tiles_per_block = (int)((ntiles + max_blocks - 1) / max_blocks);
  const int nblocks = (int)((ntiles + tiles_per_block - 1) / tiles_per_block);
  uint32_t* ghist = nullptr;
  int rc = hopb_scratch(h, 6, ((size_t)256 * nblocks + 256) * sizeof(uint32_t), (void**)&ghist);
  if (rc) return rc;
  uint32_t* totals = ghist + (size_t)256 * nblocks;
  uint64_t* kin = keys; uint64_t* kout = keys_alt;
  uint32_t* vin = vals; uint32_t* vout = vals_alt;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    HOPB_K(sort_hist_kernel)<<<nblocks, kSortThreads, 0, s>>>(kin, n, shift, tiles_per_block, ghist);
    HOPB_K(sort_rowscan_kernel)<<<256, kSortThreads, 0, s>>>(ghist, nblocks, totals);
    HOPB_K(sort_scatter_kernel)<<<nblocks, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, shift, tiles_per_block, ghist, totals);
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
    *in_alt ^= 1;
  }
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}
