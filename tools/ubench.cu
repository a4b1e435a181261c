// Micro-benchmarks that bound the hot-path kernels on B200: scattered RED rate into an L2-resident
// grid, random 2-byte gather rate from an L2-resident map (LSU vs texture path), plain streaming.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

__global__ void red_kernel(uint32_t* grid, uint32_t n_cells, int64_t n, int per_thread) {
  int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = i0; i < n; i += stride) {
    uint32_t c = hash32((uint32_t)i) % n_cells;
    atomicAdd(grid + c, 1u);
  }
}

// scattered REDs while streaming 12 B per RED from HBM (the histogram kernel's real mix)
__global__ void red_stream_kernel(uint32_t* grid, uint32_t n_cells, const float4* __restrict__ in, int64_t n4) {
  int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = i0; i + 2 * stride < n4; i += 3 * stride) {
    float4 a = in[i], b = in[i + stride], c = in[i + 2 * stride];
    uint32_t h0 = hash32(__float_as_uint(a.x) + (uint32_t)i) % n_cells;
    uint32_t h1 = hash32(__float_as_uint(b.y) + (uint32_t)i * 3u + 1u) % n_cells;
    uint32_t h2 = hash32(__float_as_uint(c.z) + (uint32_t)i * 5u + 2u) % n_cells;
    uint32_t h3 = hash32(__float_as_uint(c.w) + (uint32_t)i * 7u + 3u) % n_cells;
    atomicAdd(grid + h0, 1u); atomicAdd(grid + h1, 1u); atomicAdd(grid + h2, 1u); atomicAdd(grid + h3, 1u);
  }
}

__global__ void gather_kernel(const int16_t* __restrict__ map, uint32_t n_cells, int64_t n, int* out) {
  int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int acc = 0;
  for (int64_t i = i0; i < n; i += stride) {
    uint32_t c = hash32((uint32_t)i) % n_cells;
    acc += __ldg(map + c);
  }
  if (acc == 0x7fffffff) *out = acc;
}

__global__ void gather_tex_kernel(cudaTextureObject_t tex, uint32_t n_cells, int64_t n, int* out) {
  int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int acc = 0;
  for (int64_t i = i0; i < n; i += stride) {
    uint32_t c = hash32((uint32_t)i) % n_cells;
    acc += tex1Dfetch<short>(tex, (int)c);
  }
  if (acc == 0x7fffffff) *out = acc;
}

// shared-memory random lookup of 2-bit entries (the coarse-map idea)
__global__ void lds_kernel(const uint32_t* __restrict__ coarse, int words, int64_t n, int* out) {
  extern __shared__ uint32_t s[];
  for (int i = threadIdx.x; i < words; i += blockDim.x) s[i] = coarse[i];
  __syncthreads();
  int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int acc = 0;
  for (int64_t i = i0; i < n; i += stride) {
    uint32_t c = hash32((uint32_t)i) % (uint32_t)(words * 16);
    acc += (s[c >> 4] >> ((c & 15) * 2)) & 3;
  }
  if (acc == 0x7fffffff) *out = acc;
}

__global__ void stream_kernel(const float4* __restrict__ in, int64_t n4, float* out) {
  int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float acc = 0;
  for (int64_t i = i0; i < n4; i += stride) {
    float4 v = in[i];
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 1234.5f) *out = acc;
}

// ---- shared-memory privatisation: what a spatially partitioned histogram would be made of -------------------
// atomicAdd on shared memory at random addresses of a `tile`-counter tile (u32 counters; PACK = 2 / 4 packs u16 / u8
// counters into the same words: the instruction is the same ATOMS.ADD, the tile holds 2x / 4x the cells)
template <int PACK>
__global__ void atoms_kernel(uint32_t* out, int tile_cells, int64_t n) {
  extern __shared__ uint32_t s[];
  const int words = tile_cells / PACK;
  for (int i = threadIdx.x; i < words; i += blockDim.x) s[i] = 0;
  __syncthreads();
  int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = i0; i < n; i += stride) {
    const uint32_t c = hash32((uint32_t)i) % (uint32_t)tile_cells;
    if (PACK == 1) atomicAdd(s + c, 1u);
    else if (PACK == 2) atomicAdd(s + (c >> 1), 1u << (16 * (c & 1)));
    else atomicAdd(s + (c >> 2), 1u << (8 * (c & 3)));
  }
  __syncthreads();
  uint32_t acc = 0;
  for (int i = threadIdx.x; i < words; i += blockDim.x) acc += s[i];
  if (acc == 0xdeadbeefu) out[0] = acc;
}

// the same into the shared memory of the CTAs of a cluster (DSMEM): the tile is spread over CL CTAs
template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) dsmem_atoms_kernel(uint32_t* out, int tile_cells_per_cta, int64_t n) {
  extern __shared__ uint32_t s[];
  for (int i = threadIdx.x; i < tile_cells_per_cta; i += blockDim.x) s[i] = 0;
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(s);
  int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = i0; i < n; i += stride) {
    const uint32_t c = hash32((uint32_t)i) % (uint32_t)(tile_cells_per_cta * CL);
    const uint32_t rank = c / (uint32_t)tile_cells_per_cta, off = c - rank * (uint32_t)tile_cells_per_cta;
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + 4u * off), "r"(rank));
    asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  uint32_t acc = 0;
  for (int i = threadIdx.x; i < tile_cells_per_cta; i += blockDim.x) acc += s[i];
  if (acc == 0xdeadbeefu) out[0] = acc;
}

// the partition pass such a histogram needs first: 4-byte cell indices scattered into `NB` spatial buckets
// (CTA-local counting in shared memory, one global reservation per bucket and CTA, then the scattered stores)
template <int NB>
__global__ void __launch_bounds__(1024) partition_kernel(const uint32_t* __restrict__ cells, int64_t n, uint32_t n_cells,
                                                        uint32_t* __restrict__ bucket_fill, uint32_t* __restrict__ out,
                                                        uint32_t bucket_cap) {
  __shared__ uint32_t s_cnt[NB], s_base[NB];
  constexpr int PER = 8;                                      // 8192 keys per CTA
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x * PER;
  for (int b = threadIdx.x; b < NB; b += blockDim.x) s_cnt[b] = 0;
  __syncthreads();
  uint32_t key[PER], slot[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int64_t i = i0 + (int64_t)u * blockDim.x + threadIdx.x;
    key[u] = i < n ? cells[i] : 0xFFFFFFFFu;
  }
#pragma unroll
  for (int u = 0; u < PER; ++u)
    if (key[u] != 0xFFFFFFFFu) slot[u] = atomicAdd(&s_cnt[(uint32_t)(((uint64_t)key[u] * NB) / n_cells)], 1u);
  __syncthreads();
  for (int b = threadIdx.x; b < NB; b += blockDim.x) s_base[b] = atomicAdd(bucket_fill + b, s_cnt[b]);
  __syncthreads();
#pragma unroll
  for (int u = 0; u < PER; ++u)
    if (key[u] != 0xFFFFFFFFu) {
      const uint32_t b = (uint32_t)(((uint64_t)key[u] * NB) / n_cells);
      const uint32_t p = s_base[b] + slot[u];
      if (p < bucket_cap) out[(size_t)b * bucket_cap + p] = key[u];
    }
}

__global__ void fill_random_cells(uint32_t* cells, int64_t n, uint32_t n_cells) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    cells[i] = hash32((uint32_t)i * 2654435761u + 7u) % n_cells;
}

template <class F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const int64_t n = 60000000;  // 30k atoms x 2000 frames
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int* out; CK(cudaMalloc(&out, 4));
  for (uint32_t n_cells : {46656u, 8120601u, 26198073u, 65000000u}) {
    uint32_t* grid; CK(cudaMalloc(&grid, (size_t)n_cells * 4)); CK(cudaMemset(grid, 0, (size_t)n_cells * 4));
    for (int bps : {4, 8, 16}) {
      float ms = timeit([&] { red_kernel<<<sms * bps, 256>>>(grid, n_cells, n, 0); });
      printf("RED  cells=%9u blocks/SM=%2d  %.3f ms  %.3e red/s\n", n_cells, bps, ms, n / (ms * 1e-3));
    }
    int16_t* map = (int16_t*)grid;
    float ms = timeit([&] { gather_kernel<<<sms * 8, 256>>>(map, n_cells, n, out); });
    printf("LDG.U16 gather cells=%9u  %.3f ms  %.3e gather/s\n", n_cells, ms, n / (ms * 1e-3));
    if (n_cells < (1u << 27)) {
      cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = map;
      rd.res.linear.desc = cudaCreateChannelDesc<short>(); rd.res.linear.sizeInBytes = (size_t)n_cells * 2;
      cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
      cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
      ms = timeit([&] { gather_tex_kernel<<<sms * 8, 256>>>(tex, n_cells, n, out); });
      printf("TEX    gather cells=%9u  %.3f ms  %.3e gather/s\n", n_cells, ms, n / (ms * 1e-3));
      cudaDestroyTextureObject(tex);
    }
    CK(cudaFree(grid));
  }
  {
    const int64_t n4 = n * 3 / 4;
    float4* in; CK(cudaMalloc(&in, n4 * 16)); CK(cudaMemset(in, 0, n4 * 16));
    for (uint32_t n_cells : {8120601u, 4060300u, 2000000u}) {
      uint32_t* grid; CK(cudaMalloc(&grid, (size_t)n_cells * 4)); CK(cudaMemset(grid, 0, (size_t)n_cells * 4));
      for (int bps : {4, 8}) {
        float ms = timeit([&] { red_stream_kernel<<<sms * bps, 256>>>(grid, n_cells, in, n4); });
        printf("RED+STREAM(12B/red) cells=%9u blocks/SM=%d  %.3f ms  %.3e red/s  %.1f GB/s\n", n_cells, bps, ms, n / (ms * 1e-3), n4 * 16 / (ms * 1e-3) / 1e9);
      }
      for (int bps : {8}) {
        float ms = timeit([&] { red_kernel<<<sms * bps, 256>>>(grid, n_cells, n, 0); });
        printf("RED only            cells=%9u blocks/SM=%d  %.3f ms  %.3e red/s\n", n_cells, bps, ms, n / (ms * 1e-3));
      }
      CK(cudaFree(grid));
    }
    CK(cudaFree(in));
  }
  {
    cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int words = 8192;  // 32 KB of 2-bit entries = 131072 coarse blocks
    uint32_t* coarse; CK(cudaMalloc(&coarse, words * 4)); CK(cudaMemset(coarse, 0x55, words * 4));
    for (int w : {8192, 24576}) {
      float ms = timeit([&] { lds_kernel<<<sms * 2, 1024, 24576 * 4>>>(coarse, w > words ? words : w, n, out); });
      printf("LDS 2-bit lookup words=%d  %.3f ms  %.3e lookup/s\n", w, ms, n / (ms * 1e-3));
    }
    CK(cudaFree(coarse));
  }
  {
    // shared-memory atomics: the per-SM rate bounds any privatised histogram
    uint32_t* o32 = (uint32_t*)out;
    CK(cudaFuncSetAttribute(atoms_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(atoms_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(atoms_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int64_t na = 4 * n;
    for (int tile : {1024, 16384, 49152}) {
      for (int threads : {256, 1024}) {
        const int ctas = threads == 1024 ? 1 : 4;     // per SM, so that 4 x 48 K-counter tiles never have to fit
        if (tile * 4 * ctas > 200 * 1024) continue;
        float ms = timeit([&] { atoms_kernel<1><<<sms * ctas, threads, tile * 4>>>(o32, tile, na); });
        printf("ATOMS.ADD u32 random in a %6d-counter tile, %d x %4d threads/SM  %.3f ms  %.3e atom/s\n", tile, ctas, threads, ms, na / (ms * 1e-3));
      }
    }
    {
      float ms = timeit([&] { atoms_kernel<2><<<sms, 1024, 49152 * 4>>>(o32, 49152 * 2, na); });
      printf("ATOMS.ADD packed u16, 98304-counter tile (192 KB), 1 x 1024 threads/SM  %.3f ms  %.3e atom/s\n", ms, na / (ms * 1e-3));
      ms = timeit([&] { atoms_kernel<4><<<sms, 1024, 49152 * 4>>>(o32, 49152 * 4, na); });
      printf("ATOMS.ADD packed u8, 196608-counter tile (192 KB), 1 x 1024 threads/SM  %.3f ms  %.3e atom/s\n", ms, na / (ms * 1e-3));
    }
    {
      CK(cudaFuncSetAttribute(dsmem_atoms_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      float ms = timeit([&] { dsmem_atoms_kernel<8><<<(sms / 8) * 8, 1024, 49152 * 4>>>(o32, 49152, na); });
      printf("DSMEM red.shared::cluster.add.u32, 8-CTA cluster, 8 x 49152-counter tile  %.3f ms  %.3e atom/s\n", ms, na / (ms * 1e-3));
      CK(cudaFuncSetAttribute(dsmem_atoms_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      CK(cudaFuncSetAttribute(dsmem_atoms_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      cudaError_t e = cudaSuccess;
      ms = timeit([&] { dsmem_atoms_kernel<16><<<(sms / 16) * 16, 1024, 49152 * 4>>>(o32, 49152, na); e = cudaGetLastError(); });
      if (e == cudaSuccess) printf("DSMEM red.shared::cluster.add.u32, 16-CTA cluster, 16 x 49152-counter tile  %.3f ms  %.3e atom/s\n", ms, na / (ms * 1e-3));
      else printf("DSMEM 16-CTA cluster: %s\n", cudaGetErrorString(e));
    }
    // the partition pass: 3e8 4-byte cell indices into 64 / 256 spatial buckets
    const int64_t np = 300000000;
    const uint32_t n_cells = 8120601u;
    uint32_t *cells, *fill, *pout;
    CK(cudaMalloc(&cells, np * 4)); CK(cudaMalloc(&fill, 4096)); CK(cudaMalloc(&pout, np * 4 + (64 << 20)));
    fill_random_cells<<<sms * 8, 256>>>(cells, np, n_cells);
    const unsigned blocks = (unsigned)((np + 8191) / 8192);
    {
      const uint32_t cap = (uint32_t)(np / 64 + (1 << 16));
      float ms = timeit([&] { cudaMemsetAsync(fill, 0, 4096); partition_kernel<64><<<blocks, 1024>>>(cells, np, n_cells, fill, pout, cap); });
      printf("PARTITION 3e8 cell indices into  64 buckets  %.3f ms  %.3e keys/s (4 B in + 4 B out per key: %.0f GB/s)\n", ms, np / (ms * 1e-3), 8.0 * np / (ms * 1e-3) / 1e9);
    }
    {
      const uint32_t cap = (uint32_t)(np / 256 + (1 << 16));
      float ms = timeit([&] { cudaMemsetAsync(fill, 0, 4096); partition_kernel<256><<<blocks, 1024>>>(cells, np, n_cells, fill, pout, cap); });
      printf("PARTITION 3e8 cell indices into 256 buckets  %.3f ms  %.3e keys/s (4 B in + 4 B out per key: %.0f GB/s)\n", ms, np / (ms * 1e-3), 8.0 * np / (ms * 1e-3) / 1e9);
    }
    CK(cudaFree(cells)); CK(cudaFree(fill)); CK(cudaFree(pout));
  }
  {
    const int64_t n4 = n * 3 / 4;  // 720 MB
    float4* in; CK(cudaMalloc(&in, n4 * 16)); CK(cudaMemset(in, 0, n4 * 16));
    float* o2 = (float*)out;
    for (int bps : {4, 8}) {
      float ms = timeit([&] { stream_kernel<<<sms * bps, 256>>>(in, n4, o2); });
      printf("STREAM read 720MB blocks/SM=%d  %.3f ms  %.1f GB/s\n", bps, ms, n4 * 16 / (ms * 1e-3) / 1e9);
    }
    CK(cudaFree(in));
  }
  return 0;
}
