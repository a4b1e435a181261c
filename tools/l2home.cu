// Probe: does the home L2 partition (die) of an address matter for the scattered-reduction rate?
//
// The `ncu --set full` capture of hist_kernel shows the L2 tag pipeline as the busiest unit and half
// of the RED sectors as `lookup_miss` although the grid is L2-resident (profiles/r1_hist_hoptraj_full.md):
// B200's L2 is two partitions, one per die.  This tool
//   1. classifies every SM by the latency of a returning atomic to one fixed line (-> SM groups),
//   2. classifies every 128-byte line of a 32 MB buffer as near/far from one SM of each group by the
//      latency of a returning atomic (L2-warm), and reports the granularity of the mapping,
//   3. measures the RED rate of the whole GPU when every SM reduces only into lines that are near to
//      it, only into far lines, or into random lines.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2home tools/l2home.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// 1. the CTA that runs on SM `target`: mean latency of `reps` dependent returning atomics to line `line`
__global__ void sm_probe(uint32_t* buf, int64_t line, int reps, uint32_t target, float* out_lat, int* found) {
  if (threadIdx.x != 0 || smid() != target) return;
  if (atomicAdd(found, 1) != 0) return;
  uint32_t* p = buf + line * 32;
  uint32_t v = atomicAdd(p, 1u);
  long long t0 = clock64();
  for (int i = 0; i < reps; ++i) v = atomicAdd(p + (v >> 31), 1u);      // v < 2^31: a true dependence the compiler cannot fold
  long long t1 = clock64();
  out_lat[0] = (float)(t1 - t0) / reps + (v == 0xFFFFFFFFu ? 1.f : 0.f);
}

// 2. the CTA that runs on SM `target` measures one returning atomic per line (two passes: warm, measure)
__global__ void line_probe(uint32_t* buf, int64_t n_lines, uint32_t target, uint16_t* lat, int* found) {
  if (threadIdx.x != 0 || smid() != target) return;
  if (atomicAdd(found, 1) != 0) return;      // one CTA only
  __shared__ volatile uint32_t sink;
  uint32_t v = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int64_t i = 0; i < n_lines; ++i) {
      uint32_t* p = buf + i * 32 + (v >> 31);
      long long t0 = clock64();
      v = atomicAdd(p, 1u);
      sink = v;                      // the store issues when the result has arrived (in-order issue) ...
      long long t1 = clock64();      // ... and the clock is read after it
      if (pass == 1) lat[i] = (uint16_t)min((long long)65535, t1 - t0);
    }
}

// 3. RED throughput: group g of the SM uses list[g] (lines), `mode` 0 = near list, 1 = far list, 2 = random line
__global__ void red_rate(uint32_t* buf, const uint32_t* __restrict__ list0, uint32_t n0, const uint32_t* __restrict__ list1,
                         uint32_t n1, const uint8_t* __restrict__ sm_group, const uint32_t* __restrict__ all, uint32_t n_lines, int mode,
                         int64_t per_thread) {
  const uint32_t g = sm_group[smid()];
  const uint32_t* list = (mode == 0) == (g == 0) ? list0 : list1;   // near: own list; far: the other one
  const uint32_t n = (mode == 0) == (g == 0) ? n0 : n1;
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  for (int64_t i = 0; i < per_thread; ++i) {
    s = hash32(s + (uint32_t)i);
    const uint32_t line = mode == 2 ? all[s % n_lines] : list[s % n];
    atomicAdd(buf + (size_t)line * 32 + ((s >> 20) & 31), 1u);      // result unused -> RED
  }
}

// 4. emulation of the histogram with die-local reductions.  `coords` stands for the 12 B / point coordinate
// stream, cells_in holds the cell of every point (what the bin search would produce), cells_out is the 4 B
// brick index stage 1 writes for stage 3.  home: 1 bit per 128-byte line of the grid (1 = homed with SM group 1).
constexpr uint32_t kDone = 0x40000000u;

__device__ __forceinline__ uint32_t home_of(const uint32_t* __restrict__ home, uint32_t cell) {
  const uint32_t line = cell >> 5;
  return (__ldg(home + (line >> 5)) >> (line & 31)) & 1u;
}

// mode 0: today's kernel (every point reduced where it is read); mode 1: phase A (only points homed on this SM's die)
__global__ void __launch_bounds__(256)
hist_emul(const float4* __restrict__ coords, const uint32_t* __restrict__ cells_in, uint32_t* __restrict__ cells_out,
          uint32_t* grid, const uint32_t* __restrict__ home, const uint8_t* __restrict__ sm_group, int64_t n, int mode,
          uint32_t n_cells_copy = 0, uint32_t pair_xor = 0) {
  const uint32_t g = sm_group[smid()];
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (int64_t q = i0; 4 * q + 3 < n; q += stride) {          // four points per thread and step: 48 B of coordinates
    const float4 a = coords[3 * q], b = coords[3 * q + 1], c = coords[3 * q + 2];
    acc += a.x + b.y + c.z;
    const uint4 cc = reinterpret_cast<const uint4*>(cells_in)[q];
    uint32_t v[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (mode == 2) {
        // two copies of the grid (copy B starts n_cells_copy words further): reduce into a copy whose line is homed
        // on this SM's die when there is one
        const uint32_t hb = home_of(home, v[k] + n_cells_copy);
        const uint32_t ha = home_of(home, v[k]);
        atomicAdd(grid + v[k] + ((ha != g && hb == g) ? n_cells_copy : 0u), 1u);
      } else if (mode == 3) {
        // pairing inside a page: partner line = line ^ pair_mask (opposite home when the hash is linear)
        const uint32_t w = ((v[k] / pair_xor) * (2u * pair_xor)) | (v[k] % pair_xor);   // a zero bit inserted at pair_xor
        const uint32_t ha = home_of(home, w);
        atomicAdd(grid + (w ^ (ha != g ? pair_xor : 0u)), 1u);
      } else if (mode == 0 || home_of(home, v[k]) == g) { atomicAdd(grid + v[k], 1u); v[k] |= kDone; }
    }
    reinterpret_cast<uint4*>(cells_out)[q] = make_uint4(v[0], v[1], v[2], v[3]);
  }
  if (acc == 123.456f) grid[0] = 1;
}

// phase B: every die walks the whole index array (chunks handed out per die) and reduces what is left and its own
__global__ void __launch_bounds__(256)
hist_emul_b(const uint32_t* __restrict__ cells, uint32_t* grid, const uint32_t* __restrict__ home,
            const uint8_t* __restrict__ sm_group, int64_t n, int* next_chunk /*[2]*/, int chunk) {
  __shared__ int s_c;
  const uint32_t g = sm_group[smid()];
  const int64_t n_chunks = (n / 4 + chunk - 1) / chunk;
  while (true) {
    if (threadIdx.x == 0) s_c = atomicAdd(next_chunk + g, 1);
    __syncthreads();
    const int c = s_c;
    __syncthreads();
    if (c >= n_chunks) return;
    const int64_t q0 = (int64_t)c * chunk, q1 = min(q0 + chunk, n / 4);
    for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
      const uint4 cc = __ldg(reinterpret_cast<const uint4*>(cells) + q);
      const uint32_t v[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (!(v[k] & kDone) && home_of(home, v[k]) == g) atomicAdd(grid + v[k], 1u);
    }
  }
}

__global__ void fill_cells(uint32_t* cells, int64_t n, uint32_t n_cells) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    cells[i] = hash32((uint32_t)i * 2654435761u + 99u) % n_cells;
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int n_sm = prop.multiProcessorCount;
  const size_t bytes = (argc > 1 ? atoi(argv[1]) : 32) * (size_t)1 << 20;
  const int64_t n_lines = bytes / 128;
  uint32_t* buf;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMemset(buf, 0, bytes));
  printf("%s, %d SMs, buffer %zu MB = %lld lines\n", prop.name, n_sm, bytes >> 20, (long long)n_lines);

  // ---- 1. SM groups ---------------------------------------------------------------------------
  float* d_lat; int* d_found;
  CK(cudaMalloc(&d_lat, 4)); CK(cudaMalloc(&d_found, 4));
  std::vector<uint8_t> group(1024, 0);
  {
    std::vector<float> lat(n_sm, 0.f);
    for (int t = 0; t < n_sm; ++t) {
      CK(cudaMemset(d_found, 0, 4));
      CK(cudaMemset(d_lat, 0, 4));
      sm_probe<<<n_sm * 4, 512>>>(buf, 0, 500, (uint32_t)t, d_lat, d_found);
      CK(cudaMemcpy(&lat[t], d_lat, 4, cudaMemcpyDeviceToHost));
    }
    std::vector<float> sorted(lat); std::sort(sorted.begin(), sorted.end());
    float cut = sorted[0]; float best = 0;      // the largest gap in the sorted latencies splits the groups
    for (int i = 1; i < n_sm; ++i) if (sorted[i] - sorted[i - 1] > best) { best = sorted[i] - sorted[i - 1]; cut = 0.5f * (sorted[i] + sorted[i - 1]); }
    int n_far = 0;
    for (int i = 0; i < n_sm; ++i) { group[i] = lat[i] > cut; n_far += lat[i] > cut; }
    printf("1. atomic latency to line 0 per SM: min %.0f median %.0f max %.0f cycles; largest gap %.0f at %.0f -> %d near SMs, %d far SMs\n",
           sorted[0], sorted[n_sm / 2], sorted[n_sm - 1], best, cut, n_sm - n_far, n_far);
    printf("   smid:group ");
    for (int i = 0; i < n_sm; ++i) printf("%d", group[i]);
    printf("\n   latencies  ");
    for (int i = 0; i < n_sm; i += 8) printf("%.0f ", lat[i]);
    printf("\n");
  }
  uint8_t* d_group;
  CK(cudaMalloc(&d_group, 1024));
  CK(cudaMemcpy(d_group, group.data(), 1024, cudaMemcpyHostToDevice));

  // ---- 2. lines, seen from one SM of each group -------------------------------------------------
  uint16_t* d_llat;
  CK(cudaMalloc(&d_llat, 2 * n_lines));
  std::vector<std::vector<uint8_t>> far_bits(2, std::vector<uint8_t>(n_lines));
  for (int g = 0; g < 2; ++g) {
    uint32_t target = 0;
    while (target < (uint32_t)n_sm && group[target] != g) ++target;
    if (target == (uint32_t)n_sm) { printf("no SM in group %d\n", g); return 0; }
    CK(cudaMemset(d_found, 0, 4));
    line_probe<<<n_sm * 4, 512>>>(buf, n_lines, target, d_llat, d_found);
    CK(cudaDeviceSynchronize());
    std::vector<uint16_t> lat(n_lines);
    CK(cudaMemcpy(lat.data(), d_llat, 2 * n_lines, cudaMemcpyDeviceToHost));
    std::vector<uint16_t> sorted(lat); std::sort(sorted.begin(), sorted.end());
    // histogram in 20-cycle bins
    printf("2. SM %u (group %d): line latency percentiles 1%% %u 25%% %u 50%% %u 75%% %u 99%% %u\n", target, g,
           sorted[n_lines / 100], sorted[n_lines / 4], sorted[n_lines / 2], sorted[3 * n_lines / 4], sorted[99 * n_lines / 100]);
    int hist[64] = {0};
    for (int64_t i = 0; i < n_lines; ++i) hist[std::min(63, lat[i] / 25)]++;
    printf("   histogram (25-cycle bins from 0): ");
    for (int b = 0; b < 64; ++b) if (hist[b] > n_lines / 500) printf("[%d]=%.1f%% ", b * 25, 100.0 * hist[b] / n_lines);
    printf("\n");
    // split at the largest gap between the 10th and 90th percentile
    uint16_t lo = sorted[n_lines / 10], hi = sorted[9 * n_lines / 10];
    float cut = 0.5f * (lo + hi);
    int64_t n_far = 0, runs = 1;
    for (int64_t i = 0; i < n_lines; ++i) { far_bits[g][i] = lat[i] > cut; n_far += far_bits[g][i]; if (i && far_bits[g][i] != far_bits[g][i - 1]) ++runs; }
    printf("   cut %.0f: %.1f%% of the lines far; mean run length %.1f lines (%.0f bytes)\n", cut, 100.0 * n_far / n_lines,
           (double)n_lines / runs, 128.0 * n_lines / runs);
  }
  int64_t opposite = 0;
  for (int64_t i = 0; i < n_lines; ++i) opposite += far_bits[0][i] != far_bits[1][i];
  printf("   lines classified oppositely from the two SM groups: %.1f%%\n", 100.0 * opposite / n_lines);

  // ---- 3. RED rate: near only / far only / random --------------------------------------------------
  std::vector<uint32_t> l0, l1;     // l0: lines near to group 0 (far_bits[0] == 0)
  for (int64_t i = 0; i < n_lines; ++i) (far_bits[0][i] ? l1 : l0).push_back((uint32_t)i);
  if (l0.empty() || l1.empty()) { printf("degenerate classification\n"); return 0; }
  uint32_t *d_l0, *d_l1;
  CK(cudaMalloc(&d_l0, 4 * l0.size())); CK(cudaMalloc(&d_l1, 4 * l1.size()));
  CK(cudaMemcpy(d_l0, l0.data(), 4 * l0.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_l1, l1.data(), 4 * l1.size(), cudaMemcpyHostToDevice));
  std::vector<uint32_t> all(n_lines);
  for (int64_t i = 0; i < n_lines; ++i) all[i] = (uint32_t)i;
  uint32_t* d_all;
  CK(cudaMalloc(&d_all, 4 * n_lines));
  CK(cudaMemcpy(d_all, all.data(), 4 * n_lines, cudaMemcpyHostToDevice));
  const char* names[3] = {"near lines only", "far lines only ", "random lines   "};
  for (int rep = 0; rep < 2; ++rep)
    for (int mode = 0; mode < 3; ++mode) {
      const int blocks = n_sm * 8, threads = 256;
      const int64_t per_thread = 2000;
      cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
      red_rate<<<blocks, threads>>>(buf, d_l0, (uint32_t)l0.size(), d_l1, (uint32_t)l1.size(), d_group, d_all, (uint32_t)n_lines, mode, 200);
      CK(cudaEventRecord(a));
      red_rate<<<blocks, threads>>>(buf, d_l0, (uint32_t)l0.size(), d_l1, (uint32_t)l1.size(), d_group, d_all, (uint32_t)n_lines, mode, per_thread);
      CK(cudaEventRecord(b));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      printf("3. %s: %.3f ms for %.3g REDs -> %.3g RED/s\n", names[mode], ms, (double)blocks * threads * per_thread,
             (double)blocks * threads * per_thread / (ms * 1e-3));
    }
  // ---- 4. histogram emulation: one pass vs die-local two-phase ----------------------------------------
  {
    const int64_t n = 300000000;
    const uint32_t n_cells = (uint32_t)(bytes / 4);
    std::vector<uint32_t> home((n_lines + 31) / 32, 0u);
    for (int64_t i = 0; i < n_lines; ++i) if (far_bits[0][i]) home[i >> 5] |= 1u << (i & 31);   // far from group 0 = homed with group 1
    uint32_t *d_home, *d_cin, *d_cout; float4* d_xyz; int* d_next;
    CK(cudaMalloc(&d_home, 4 * home.size()));
    CK(cudaMemcpy(d_home, home.data(), 4 * home.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_cin, 4 * n)); CK(cudaMalloc(&d_cout, 4 * n)); CK(cudaMalloc(&d_xyz, 12 * n)); CK(cudaMalloc(&d_next, 8));
    CK(cudaMemset(d_xyz, 0, 12 * n));
    fill_cells<<<n_sm * 8, 256>>>(d_cin, n, n_cells);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    const int blocks = (int)((n / 4 + 256 * 4 - 1) / (256 * 4));     // ~4 steps per thread: many short-lived CTAs, like hist_kernel
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemset(buf, 0, bytes));
      CK(cudaEventRecord(e0));
      hist_emul<<<blocks, 256>>>(d_xyz, d_cin, d_cout, buf, d_home, d_group, n, 0);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float t0; CK(cudaEventElapsedTime(&t0, e0, e1));
      uint32_t first_a; CK(cudaMemcpy(&first_a, buf + 12345, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemset(buf, 0, bytes));
      CK(cudaMemset(d_next, 0, 8));
      CK(cudaEventRecord(e0));
      hist_emul<<<blocks, 256>>>(d_xyz, d_cin, d_cout, buf, d_home, d_group, n, 1);
      CK(cudaEventRecord(e1));
      hist_emul_b<<<n_sm * 8, 256>>>(d_cout, buf, d_home, d_group, n, d_next, 2048);
      CK(cudaEventRecord(e2));
      CK(cudaDeviceSynchronize());
      float ta, tb; CK(cudaEventElapsedTime(&ta, e0, e1)); CK(cudaEventElapsedTime(&tb, e1, e2));
      uint32_t first_b; CK(cudaMemcpy(&first_b, buf + 12345, 4, cudaMemcpyDeviceToHost));
      printf("4. %.3g points into the %zu MB grid: one pass %.3f ms; die-local phase A %.3f ms + phase B %.3f ms = %.3f ms (counts agree: %s)\n",
             (double)n, bytes >> 20, t0, ta, tb, ta + tb, first_a == first_b ? "yes" : "NO");
    }
  }
  // ---- 5. two copies of the grid, reductions into the copy that is homed on the reading SM's die ----------
  if (argc > 2) {
    FILE* f = fopen(argv[2], "wb");
    if (f) {
      fprintf(f, "%p %lld %d\n", (void*)buf, (long long)n_lines, n_sm);
      fwrite(group.data(), 1, n_sm, f);
      fwrite(far_bits[0].data(), 1, n_lines, f);
      fclose(f);
    }
    const int64_t n = 300000000;
    const uint32_t n_cells = (uint32_t)(bytes / 8);       // the logical grid is half the buffer
    const uint32_t pair_xor = argc > 3 ? (uint32_t)strtoul(argv[3], 0, 0) : 32u;
    std::vector<uint32_t> home((n_lines + 31) / 32, 0u);
    int64_t near_any = 0;
    for (int64_t i = 0; i < n_lines; ++i) if (far_bits[0][i]) home[i >> 5] |= 1u << (i & 31);
    for (int64_t i = 0; i < n_lines / 2; ++i) near_any += far_bits[0][i] != far_bits[0][i + n_lines / 2];
    int64_t pair_opp = 0;
    for (int64_t i = 0; i < n_lines; ++i) pair_opp += far_bits[0][i] != far_bits[0][i ^ (pair_xor >> 5 ? pair_xor >> 5 : 1)];
    printf("5. two copies: %.1f%% of the logical lines have opposite homes in copy A and B (both dies served near); "
           "line ^ %u has the opposite home for %.1f%% of the lines\n", 100.0 * near_any / (n_lines / 2), pair_xor >> 5,
           100.0 * pair_opp / n_lines);
    uint32_t *d_home, *d_cin, *d_cout; float4* d_xyz;
    CK(cudaMalloc(&d_home, 4 * home.size()));
    CK(cudaMemcpy(d_home, home.data(), 4 * home.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_cin, 4 * n)); CK(cudaMalloc(&d_cout, 4 * n)); CK(cudaMalloc(&d_xyz, 12 * n));
    CK(cudaMemset(d_xyz, 0, 12 * n));
    fill_cells<<<n_sm * 8, 256>>>(d_cin, n, n_cells);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = (int)((n / 4 + 256 * 4 - 1) / (256 * 4));
    for (int rep = 0; rep < 3; ++rep)
      for (int mode = 0; mode <= 3; ++mode) {
        if (mode == 1) continue;
        CK(cudaMemset(buf, 0, bytes));
        CK(cudaEventRecord(e0));
        hist_emul<<<blocks, 256>>>(d_xyz, d_cin, d_cout, buf, d_home, d_group, n, mode, n_cells, pair_xor);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float t0; CK(cudaEventElapsedTime(&t0, e0, e1));
        printf("5. %.3g points, logical grid %zu MB, %s: %.3f ms\n", (double)n, bytes >> 21,
               mode == 0 ? "one copy" : mode == 2 ? "two copies, near one chosen per line" : "interleaved pair (line ^ xor)", t0);
      }
  }
  return 0;
}
