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
