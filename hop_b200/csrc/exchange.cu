// The exchange steps of the frame-sharded pipeline on NVSwitch multicast memory (NVLS): `multimem.ld_reduce` lets
// the switch add the N ranks' copies of a word on the way to the SM, `multimem.st` writes a word into every rank's
// copy.  The three messages of SURVEY.md 8e -- the u32 count grid (32 MB at 201^3), one 44-byte slab summary per
// atom and rank, the dense transition-count matrix -- are small enough that a library collective is mostly launch
// and stream hand-over latency (NCCL: 0.15 ms for the count grid on 8 GPUs, ~0.3 ms of a 3.1 ms step for all three,
// profiles/r1_nccl_collectives_8gpu.txt); here each is ONE kernel on the pipeline's own stream between two
// cross-rank barriers of the symmetric-memory handle (hop_b200/exchange.py).
//
// The buffers live in symmetric memory (torch.distributed._symmetric_memory: the same allocation mapped on every
// rank plus a multicast address for all copies); `mc` arguments are multicast addresses, valid for multimem.* only.
// Replaces nothing in hop (the reference is serial): it is the exchange of partial histograms / chunk states that
// sharding DensityCollector.collect (hop/density.py:125-131) and TransportNetwork._analyze_hops
// (hop/graph.py:212-254) over frames needs.
#include <cstdlib>

#include "hopb_common.cuh"

namespace {

constexpr int kXThreads = 512;

__device__ __forceinline__ uint64_t mm_ld_add_u64(const uint64_t* mc) {
  uint64_t v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(mc) : "memory");
  return v;
}

__device__ __forceinline__ void mm_st_u64(uint64_t* mc, uint64_t v) {
  asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" : : "l"(mc), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t mm_ld_add_u32(const uint32_t* mc) {
  uint32_t v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u32 %0, [%1];" : "=r"(v) : "l"(mc) : "memory");
  return v;
}

__device__ __forceinline__ void mm_st_u32(uint32_t* mc, uint32_t v) {
  asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" : : "l"(mc), "r"(v) : "memory");
}

// In-place sum over the ranks of a u32 array: rank r reduces the r-th slice through the switch and broadcasts it
// (two-shot: every word crosses each NVLink once in and once out).  Two counters travel per 64-bit add; a carry
// from the low into the high one would need a counter to pass 2^32, which the u32 grid cannot hold anyway.
__global__ void __launch_bounds__(kXThreads)
nvls_allreduce_u32_kernel(uint32_t* mc, int64_t n_words, int rank, int world) {
  const int64_t n_pairs = n_words >> 1;
  const int64_t per = (n_pairs + world - 1) / world;
  const int64_t lo = per * rank, hi = lo + per < n_pairs ? lo + per : n_pairs;
  uint64_t* mc2 = reinterpret_cast<uint64_t*>(mc);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;      // loads of a step in flight together: a round trip through the switch is microseconds
  for (int64_t i = lo + tid; i < hi; i += nth * U) {
    uint64_t v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * nth < hi) v[u] = mm_ld_add_u64(mc2 + i + u * nth);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * nth < hi) mm_st_u64(mc2 + i + u * nth, v[u]);
  }
  if ((n_words & 1) && rank == world - 1 && tid == 0) mm_st_u32(mc + n_words - 1, mm_ld_add_u32(mc + n_words - 1));
}

// The same sum with the reduction on the SM: rank r reads slice r of every rank's copy through peer mappings (16-byte
// loads, all ranks' loads of a step in flight together) and broadcasts the sum with one 16-byte multicast store.
// For a few ranks this beats the in-switch reduction, whose integer form moves at most 8 bytes per instruction.
struct PeerPtrs { const uint4* p[16]; };

template <int W>
__global__ void __launch_bounds__(kXThreads)
p2p_allreduce_u32_kernel(PeerPtrs peers, uint32_t* mc, int64_t n_words, int rank, int world_rt) {
  const int world = W > 0 ? W : world_rt;
  const int64_t n4 = n_words >> 2;
  const int64_t per = (n4 + world - 1) / world;
  const int64_t lo = per * rank, hi = lo + per < n4 ? lo + per : n4;
  uint4* mc4 = reinterpret_cast<uint4*>(mc);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = lo + tid; i < hi; i += nth) {
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    if (W > 0) {
      uint4 v[W > 0 ? W : 1];
#pragma unroll
      for (int r = 0; r < W; ++r) v[r] = __ldcv(peers.p[r] + i);   // written by other GPUs since this SM last saw the line
#pragma unroll
      for (int r = 0; r < W; ++r) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
    } else {
      for (int r = 0; r < world; ++r) { const uint4 v = __ldcv(peers.p[r] + i); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    }
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" : : "l"(mc4 + i), "r"(acc.x), "r"(acc.y), "r"(acc.z), "r"(acc.w) : "memory");
  }
  // the last n_words % 4 words: by the last rank, word by word
  if (rank == world - 1 && tid < (n_words & 3)) {
    const int64_t w = (n4 << 2) + tid;
    uint32_t acc = 0;
    for (int r = 0; r < world; ++r) acc += __ldcv(reinterpret_cast<const uint32_t*>(peers.p[r]) + w);
    mm_st_u32(mc + w, acc);
  }
}

// all-gather: rank r writes its n_bytes (a multiple of 16) into slot r of every rank's [world][n_bytes] buffer
__global__ void __launch_bounds__(kXThreads)
nvls_allgather_kernel(const uint4* __restrict__ src, uint4* mc_out, int64_t n16, int rank) {
  uint4* dst = mc_out + n16 * rank;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" : : "l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
}

}  // namespace

extern "C" {

int hopb_nvls_allreduce_u32(hopb_handle* h, void* stream, uint32_t* mc, int64_t n_words, int rank, int world) {
  HOPB_REQUIRE(h, mc && n_words > 0 && world >= 1 && rank >= 0 && rank < world, "bad argument");
  HOPB_REQUIRE(h, ((uintptr_t)mc & 7) == 0, "multicast address must be 8-byte aligned");
  const int64_t per = ((n_words >> 1) + world - 1) / world;
  static const int ctas_per_sm = getenv("HOPB_X_CTAS") ? atoi(getenv("HOPB_X_CTAS")) : 2;
  int64_t blocks = (per + kXThreads * 4 - 1) / (kXThreads * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > (int64_t)ctas_per_sm * h->num_sms) blocks = (int64_t)ctas_per_sm * h->num_sms;
  HOPB_K(nvls_allreduce_u32_kernel)<<<(unsigned)blocks, kXThreads, 0, (cudaStream_t)stream>>>(mc, n_words, rank, world);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_p2p_allreduce_u32(hopb_handle* h, void* stream, const void* const* peer_ptrs, uint32_t* mc, int64_t n_words, int rank,
                           int world) {
  HOPB_REQUIRE(h, peer_ptrs && mc && n_words > 0 && world >= 1 && world <= 16 && rank >= 0 && rank < world, "bad argument");
  PeerPtrs pp;
  for (int r = 0; r < 16; ++r) pp.p[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    HOPB_REQUIRE(h, peer_ptrs[r] && ((uintptr_t)peer_ptrs[r] & 15) == 0, "peer buffers must be 16-byte aligned");
    pp.p[r] = reinterpret_cast<const uint4*>(peer_ptrs[r]);
  }
  HOPB_REQUIRE(h, ((uintptr_t)mc & 15) == 0, "multicast address must be 16-byte aligned");
  const int64_t per = ((n_words >> 2) + world - 1) / world;
  static const int ctas_per_sm = getenv("HOPB_X_CTAS") ? atoi(getenv("HOPB_X_CTAS")) : 2;
  int64_t blocks = (per + kXThreads - 1) / kXThreads;
  if (blocks < 1) blocks = 1;
  if (blocks > (int64_t)ctas_per_sm * h->num_sms) blocks = (int64_t)ctas_per_sm * h->num_sms;
  cudaStream_t s = (cudaStream_t)stream;
  hopb_note_launch();
  switch (world) {
    case 2: p2p_allreduce_u32_kernel<2><<<(unsigned)blocks, kXThreads, 0, s>>>(pp, mc, n_words, rank, world); break;
    case 4: p2p_allreduce_u32_kernel<4><<<(unsigned)blocks, kXThreads, 0, s>>>(pp, mc, n_words, rank, world); break;
    case 8: p2p_allreduce_u32_kernel<8><<<(unsigned)blocks, kXThreads, 0, s>>>(pp, mc, n_words, rank, world); break;
    default: p2p_allreduce_u32_kernel<0><<<(unsigned)blocks, kXThreads, 0, s>>>(pp, mc, n_words, rank, world); break;
  }
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_nvls_allgather(hopb_handle* h, void* stream, const void* src, void* mc_out, int64_t n_bytes, int rank, int world) {
  HOPB_REQUIRE(h, src && mc_out && n_bytes > 0 && world >= 1 && rank >= 0 && rank < world, "bad argument");
  HOPB_REQUIRE(h, (n_bytes & 15) == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)mc_out & 15) == 0, "16-byte alignment");
  const int64_t n16 = n_bytes / 16;
  int64_t blocks = (n16 + kXThreads - 1) / kXThreads;
  if (blocks > 2 * h->num_sms) blocks = 2 * h->num_sms;
  HOPB_K(nvls_allgather_kernel)<<<(unsigned)blocks, kXThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src),
                                                                                     reinterpret_cast<uint4*>(mc_out), n16, rank);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

}  // extern "C"
