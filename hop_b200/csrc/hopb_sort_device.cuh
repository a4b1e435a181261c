// Device-side building blocks of the scan and the stable LSD radix sort (8-bit digits), shared by
// the multi-launch driver in sort.cu and the single cooperative kernel in events.cu.
#pragma once
#include "hopb_common.cuh"

namespace hopb_sort {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;                         // rows of 32 per warp per tile
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048
constexpr int kMaxSortBlocks = 1024;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(kFull, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  const uint32_t res = s_warp[warp] + inc - v;
  if (total) *total = s_warp[32];
  __syncthreads();
  return res;
}

__device__ __forceinline__ unsigned digit_of(uint64_t key, int shift) { return (unsigned)(key >> shift) & 255u; }

struct SortSmem {
  uint32_t off[256];
  uint32_t wcnt[kSortWarps][256];
  uint32_t warp[33];
};

// phase 1: digit histogram of this block's tiles -> ghist[digit][block]
__device__ __forceinline__ void sort_hist_block(const uint64_t* __restrict__ keys, int64_t n, int shift, int tiles_per_block,
                                                uint32_t* __restrict__ ghist, SortSmem& sm) {
  sm.off[threadIdx.x] = 0;
  __syncthreads();
  const int64_t begin = (int64_t)blockIdx.x * tiles_per_block * kSortTile;
  int64_t end = begin + (int64_t)tiles_per_block * kSortTile;
  if (end > n) end = n;
  for (int64_t i = begin + threadIdx.x; i < end; i += kSortThreads) atomicAdd(&sm.off[digit_of(keys[i], shift)], 1u);
  __syncthreads();
  ghist[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = sm.off[threadIdx.x];
  __syncthreads();
}

// phase 2: per digit, exclusive scan along the blocks; digit totals.  Rows are dealt to the blocks.
__device__ __forceinline__ void sort_rowscan_rows(uint32_t* __restrict__ ghist, int nblocks, uint32_t* __restrict__ totals,
                                                  SortSmem& sm) {
  for (int d = blockIdx.x; d < 256; d += gridDim.x) {
    uint32_t carry = 0;
    for (int b0 = 0; b0 < nblocks; b0 += blockDim.x) {
      const int b = b0 + threadIdx.x;
      const uint32_t v = b < nblocks ? ghist[(int64_t)d * nblocks + b] : 0u;
      uint32_t total;
      const uint32_t ex = block_exclusive_scan(v, sm.warp, &total);
      if (b < nblocks) ghist[(int64_t)d * nblocks + b] = carry + ex;
      carry += total;
    }
    if (threadIdx.x == 0) totals[d] = carry;
  }
}

// phase 3: stable scatter of this block's tiles
__device__ __forceinline__ void sort_scatter_block(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                   uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                                                   int shift, int tiles_per_block, const uint32_t* __restrict__ ghist,
                                                   const uint32_t* __restrict__ totals, SortSmem& sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    // digit base = exclusive scan of the 256 digit totals + this block's offset within the digit
    uint32_t t = totals[threadIdx.x];
    uint32_t ex = block_exclusive_scan(t, sm.warp, nullptr);
    sm.off[threadIdx.x] = ex + ghist[(int64_t)threadIdx.x * gridDim.x + blockIdx.x];
  }
  __syncthreads();
  const int64_t begin = (int64_t)blockIdx.x * tiles_per_block * kSortTile;
  for (int t = 0; t < tiles_per_block; ++t) {
    const int64_t tile0 = begin + (int64_t)t * kSortTile;
    if (tile0 >= n) break;
    const int64_t wbase = tile0 + (int64_t)warp * 32 * kSortItems;
    uint64_t k[kSortItems];
    uint32_t v[kSortItems];
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&sm.wcnt[0][0])[i] = 0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
      const int64_t idx = wbase + r * 32 + lane;
      const bool valid = idx < n;
      k[r] = valid ? keys[idx] : 0;
      v[r] = valid ? vals[idx] : 0;
      const unsigned act = __ballot_sync(kFull, valid);
      if (valid) {
        const unsigned dgt = digit_of(k[r], shift);
        const unsigned peers = __match_any_sync(act, dgt);
        if ((peers & ((1u << lane) - 1)) == 0) sm.wcnt[warp][dgt] += __popc(peers);
      }
      __syncwarp();
    }
    __syncthreads();
    {
      const int dgt = threadIdx.x;
      uint32_t run = sm.off[dgt];
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t c = sm.wcnt[w][dgt];
        sm.wcnt[w][dgt] = run;
        run += c;
      }
      sm.off[dgt] = run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
      const int64_t idx = wbase + r * 32 + lane;
      const bool valid = idx < n;
      const unsigned act = __ballot_sync(kFull, valid);
      if (valid) {
        const unsigned dgt = digit_of(k[r], shift);
        const unsigned peers = __match_any_sync(act, dgt);
        const unsigned below = peers & ((1u << lane) - 1);
        const uint32_t pos = sm.wcnt[warp][dgt] + __popc(below);
        keys_out[pos] = k[r];
        vals_out[pos] = v[r];
        __syncwarp(act);
        if (below == 0) sm.wcnt[warp][dgt] += __popc(peers);
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

}  // namespace hopb_sort
