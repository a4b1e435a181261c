// Prototype: spatially partitioned, shared-memory-privatised histogram (north_star kernel 2) against the
// one-RED-per-point kernel, on the bench's shape: 3e8 points, 201^3 = 8.1 M cells, points uniformly spread
// (0.004 points per cell and frame: no duplicates inside a tile of the stream).
//   pass A  stream 16 B per point (12 B coordinates + the 4 B cell index that stands for the bin search), write the
//           4 B brick index stage 3 needs, and scatter a 16-bit bucket-local cell index into the bucket of its cell
//           (CTA-local multisplit in shared memory: count, reserve once per bucket and tile, stage, write runs)
//   pass B  one CTA per bucket: ATOMS.ADD into a shared-memory tile, then counts[cell] += tile[cell] (the bucket owns
//           its cell range: plain, coalesced read-modify-write)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/histpart tools/histpart.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

__global__ void fill_cells(uint32_t* cells, int64_t n, uint32_t n_cells) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    cells[i] = hash32((uint32_t)i * 2654435761u + 99u) % n_cells;
}

// baseline: what hist_kernel does per point
__global__ void __launch_bounds__(256) hist_red(const float4* __restrict__ xyz4, const uint32_t* __restrict__ cells_in,
                                                uint32_t* __restrict__ cells_out, uint32_t* counts, int64_t n) {
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (int64_t q = i0; 4 * q + 3 < n; q += stride) {
    const float4 a = xyz4[3 * q], b = xyz4[3 * q + 1], c = xyz4[3 * q + 2];
    acc += a.x + b.y + c.z;
    const uint4 cc = reinterpret_cast<const uint4*>(cells_in)[q];
    atomicAdd(counts + cc.x, 1u); atomicAdd(counts + cc.y, 1u); atomicAdd(counts + cc.z, 1u); atomicAdd(counts + cc.w, 1u);
    reinterpret_cast<uint4*>(cells_out)[q] = cc;
  }
  if (acc == 123.456f) counts[0] = 1;
}

constexpr int kThreadsA = 512;
template <int SHIFT, int TILE>
__global__ void __launch_bounds__(kThreadsA) part_a(const float4* __restrict__ xyz4, const uint32_t* __restrict__ cells_in,
                                                    uint32_t* __restrict__ cells_out, int64_t n, int n_buckets,
                                                    uint32_t* __restrict__ gfill, uint16_t* __restrict__ buckets, uint32_t cap,
                                                    uint32_t* counts /* overflow: straight reductions */) {
  extern __shared__ unsigned char smem[];
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem);                 // [n_buckets] counts, then tile offsets
  uint32_t* s_gbase = s_cnt + n_buckets;                               // [n_buckets] start of this tile's run in the bucket
  uint16_t* s_stage = reinterpret_cast<uint16_t*>(s_gbase + n_buckets);  // [TILE]
  uint16_t* s_bkt = s_stage + TILE;                                    // [TILE] bucket of every staged entry
  __shared__ uint32_t s_warp[kThreadsA / 32];
  constexpr int PER = TILE / kThreadsA;                                // points per thread and tile (multiple of 4)
  const int64_t n_tiles = (n + TILE - 1) / TILE;
  float acc = 0.f;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    for (int b = threadIdx.x; b < n_buckets; b += kThreadsA) s_cnt[b] = 0;
    __syncthreads();
    uint32_t cell[PER], slot[PER];
    const int64_t q0 = t * (TILE / 4);
#pragma unroll
    for (int u = 0; u < PER / 4; ++u) {
      const int64_t q = q0 + u * kThreadsA + threadIdx.x;
      uint4 cc = make_uint4(~0u, ~0u, ~0u, ~0u);
      if (4 * q + 3 < n) {
        const float4 a = xyz4[3 * q], b = xyz4[3 * q + 1], c = xyz4[3 * q + 2];
        acc += a.x + b.y + c.z;
        cc = reinterpret_cast<const uint4*>(cells_in)[q];
        reinterpret_cast<uint4*>(cells_out)[q] = cc;
      }
      cell[4 * u] = cc.x; cell[4 * u + 1] = cc.y; cell[4 * u + 2] = cc.z; cell[4 * u + 3] = cc.w;
    }
#pragma unroll
    for (int u = 0; u < PER; ++u)
      if (cell[u] != ~0u) slot[u] = atomicAdd(&s_cnt[cell[u] >> SHIFT], 1u);
    __syncthreads();
    // exclusive scan of the bucket counts (n_buckets <= 1024: up to two per thread), global reservation
    uint32_t c0 = 0, c1 = 0;
    const int b0 = 2 * threadIdx.x, b1 = b0 + 1;
    if (b0 < n_buckets) c0 = s_cnt[b0];
    if (b1 < n_buckets) c1 = s_cnt[b1];
    uint32_t incl = c0 + c1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < kThreadsA / 32 ? s_warp[lane] : 0, wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
      if (lane < kThreadsA / 32) s_warp[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t ex = s_warp[warp] + incl - (c0 + c1);
    if (b0 < n_buckets) { s_cnt[b0] = ex; s_gbase[b0] = c0 ? atomicAdd(gfill + b0, c0) : 0u; }
    if (b1 < n_buckets) { s_cnt[b1] = ex + c0; s_gbase[b1] = c1 ? atomicAdd(gfill + b1, c1) : 0u; }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PER; ++u)
      if (cell[u] != ~0u) {
        const uint32_t b = cell[u] >> SHIFT, p = s_cnt[b] + slot[u];
        s_stage[p] = (uint16_t)(cell[u] & ((1u << SHIFT) - 1));
        s_bkt[p] = (uint16_t)b;
      }
    __syncthreads();
    const int64_t rem = n - t * (int64_t)TILE;
    const int n_valid = (int)(rem < TILE ? rem : TILE);
    for (int i = threadIdx.x; i < n_valid; i += kThreadsA) {
      const uint32_t b = s_bkt[i];
      const uint32_t pos = s_gbase[b] + ((uint32_t)i - s_cnt[b]);
      if (pos < cap) buckets[(size_t)b * cap + pos] = s_stage[i];
      else atomicAdd(counts + ((b << SHIFT) | s_stage[i]), 1u);      // bucket full: an ordinary reduction (always correct)
    }
    __syncthreads();
  }
  if (acc == 123.456f) counts[0] = 1;
}

template <int SHIFT>
__global__ void __launch_bounds__(1024) part_b(const uint16_t* __restrict__ buckets, const uint32_t* __restrict__ gfill, uint32_t cap,
                                               int n_buckets, uint32_t* __restrict__ counts, uint32_t n_cells) {
  extern __shared__ uint32_t s_tile[];
  constexpr int CELLS = 1 << SHIFT;
  for (int b = blockIdx.x; b < n_buckets; b += gridDim.x) {
    for (int i = threadIdx.x; i < CELLS; i += blockDim.x) s_tile[i] = 0;
    __syncthreads();
    uint32_t nb = gfill[b];
    if (nb > cap) nb = cap;
    const uint4* src = reinterpret_cast<const uint4*>(buckets + (size_t)b * cap);    // 8 entries per load
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
    const uint32_t c0 = (uint32_t)b << SHIFT;
    for (int i = threadIdx.x; i < CELLS; i += blockDim.x)
      if (c0 + i < n_cells && s_tile[i]) counts[c0 + i] += s_tile[i];
    __syncthreads();
  }
}

template <int SHIFT, int TILE>
void run(const float4* xyz, const uint32_t* cin, uint32_t* cout, int64_t n, uint32_t n_cells, uint32_t* counts, uint32_t* ref, int sms) {
  const int n_buckets = (int)((n_cells + (1u << SHIFT) - 1) >> SHIFT);
  const uint32_t cap = (uint32_t)(((2 * n) / n_buckets + 65535) & ~7ll);
  uint32_t* gfill; uint16_t* buckets;
  CK(cudaMalloc(&gfill, n_buckets * 4)); CK(cudaMalloc(&buckets, (size_t)n_buckets * cap * 2));
  const size_t smem_a = (size_t)n_buckets * 8 + TILE * 4;
  CK(cudaFuncSetAttribute(part_a<SHIFT, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  CK(cudaFuncSetAttribute(part_b<SHIFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (4 << SHIFT)));
  int occ_a = 0, occ_b = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, part_a<SHIFT, TILE>, kThreadsA, smem_a));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, part_b<SHIFT>, 1024, (4 << SHIFT)));
  cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(counts, 0, (size_t)n_cells * 4));
    CK(cudaMemset(gfill, 0, n_buckets * 4));
    CK(cudaEventRecord(e0));
    part_a<SHIFT, TILE><<<sms * occ_a, kThreadsA, smem_a>>>(xyz, cin, cout, n, n_buckets, gfill, buckets, cap, counts);
    CK(cudaEventRecord(e1));
    part_b<SHIFT><<<n_buckets < sms * occ_b * 4 ? n_buckets : sms * occ_b * 4, 1024, (4 << SHIFT)>>>(buckets, gfill, cap, n_buckets, counts, n_cells);
    CK(cudaEventRecord(e2));
    CK(cudaDeviceSynchronize());
    float ta, tb; CK(cudaEventElapsedTime(&ta, e0, e1)); CK(cudaEventElapsedTime(&tb, e1, e2));
    std::vector<uint32_t> a(1 << 16), b(1 << 16);
    CK(cudaMemcpy(a.data(), counts + 1000000, a.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), ref + 1000000, b.size() * 4, cudaMemcpyDeviceToHost));
    printf("buckets of %6d cells (%4d), tile %5d, CTAs/SM A %d B %d: pass A %.3f ms + pass B %.3f ms = %.3f ms  (counts %s)\n", 1 << SHIFT, n_buckets, TILE,
           occ_a, occ_b, ta, tb, ta + tb, a == b ? "agree" : "DIFFER");
  }
  CK(cudaFree(gfill)); CK(cudaFree(buckets));
}

int main() {
  const int64_t n = 300000000;
  const uint32_t n_cells = 8120601u;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float4* xyz; uint32_t *cin, *cout, *counts, *ref;
  CK(cudaMalloc(&xyz, n * 12)); CK(cudaMemset(xyz, 0, n * 12));
  CK(cudaMalloc(&cin, n * 4)); CK(cudaMalloc(&cout, n * 4));
  CK(cudaMalloc(&counts, (size_t)n_cells * 4)); CK(cudaMalloc(&ref, (size_t)n_cells * 4));
  fill_cells<<<sms * 8, 256>>>(cin, n, n_cells);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(ref, 0, (size_t)n_cells * 4));
    CK(cudaEventRecord(e0));
    hist_red<<<(int)((n / 4 + 256 * 4 - 1) / (256 * 4)), 256>>>(xyz, cin, cout, ref, n);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float t; CK(cudaEventElapsedTime(&t, e0, e1));
    printf("one RED per point: %.3f ms\n", t);
  }
  run<15, 8192>(xyz, cin, cout, n, n_cells, counts, ref, sms);
  run<15, 16384>(xyz, cin, cout, n, n_cells, counts, ref, sms);
  run<14, 16384>(xyz, cin, cout, n, n_cells, counts, ref, sms);
  run<14, 8192>(xyz, cin, cout, n, n_cells, counts, ref, sms);
  return 0;
}
