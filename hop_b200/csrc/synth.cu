// Synthetic solvated-box trajectory generator (K6, bench / test input only).
// Bit-identical to hop_b200/synth.py: integer hashing plus float32 add / mul / compare in a fixed
// order with FMA contraction ruled out by the explicit _rn intrinsics.  One thread per atom walks
// the frames sequentially (the dynamics are a random walk), writing frames >= frame_start.
#include "hopb_common.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ uint64_t hash64(uint64_t seed, uint64_t atom, uint64_t frame, uint64_t stream) {
  const uint64_t z = seed * 0xD6E8FEB86659FD93ull + atom * 0x9E3779B97F4A7C15ull + frame * 0xC2B2AE3D27D4EB4Full + stream;
  return mix64(mix64(z));
}

__device__ __forceinline__ float u21(uint64_t h, int k) {
  return __fmul_rn((float)((h >> (21 * k)) & 0x1FFFFFull), 4.76837158203125e-07f);  // 2^-21
}

__device__ __forceinline__ float gauss3(uint64_t h) {
  const float s = __fadd_rn(__fadd_rn(u21(h, 0), u21(h, 1)), u21(h, 2));
  return __fmul_rn(__fadd_rn(s, -1.5f), 2.0f);
}

__global__ void __launch_bounds__(128)
synth_kernel(hopb_synth_params p, const float* __restrict__ centres, int64_t frame_start, int64_t n_out, float* __restrict__ xyz,
             float4* __restrict__ walker, int resume) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t N = p.n_atoms;
  if (i >= N) return;
  const float L = p.box;
  const float c = __fmul_rn(L, 0.5f);
  const float lo = __fadd_rn(c, -p.cube_half);
  const float hi = __fadd_rn(c, p.cube_half);
  const float top = __fmul_rn(L, 0.99999904632568359375f);
  const bool bound = (i % p.site_stride) == 0;
  const uint64_t K = (uint64_t)p.n_sites;
  const uint64_t seed = p.seed;
  const float p_jump = __fmul_rn(0.5f, p.p_hop);

  float pos[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) pos[d] = __fmul_rn(u21(hash64(seed, i, 0, 10 + d), 0), L);
  const bool inside = pos[0] > lo && pos[0] < hi && pos[1] > lo && pos[1] < hi && pos[2] > lo && pos[2] < hi;
  if (inside) pos[0] = __fadd_rn(pos[0], pos[0] < c ? -p.cube_half : p.cube_half);
  long long site = bound ? (long long)((hash64(seed, i, 0, 20) >> 11) % K) : -1;
  if (N > 2) {
    if (i == 1) { pos[0] = pos[1] = pos[2] = 0.0f; }
    if (i == 2) { pos[0] = pos[1] = pos[2] = top; }
  }
  if (site >= 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d)
      pos[d] = __fadd_rn(centres[site * 3 + d], __fmul_rn(gauss3(hash64(seed, i, 0, 30 + d)), p.sigma_site));
  }
  const int64_t f_stop = frame_start + n_out;
  if (frame_start == 0 && n_out > 0) {
    float* o = xyz + i * 3;
    o[0] = pos[0]; o[1] = pos[1]; o[2] = pos[2];
  }
  // a trajectory generated chunk by chunk continues from the walker the previous call left behind (position and
  // site after frame frame_start - 1) instead of walking there from frame 0 again
  int64_t f_first = 1;
  if (resume && frame_start > 0) {
    const float4 w = walker[i];
    pos[0] = w.x; pos[1] = w.y; pos[2] = w.z;
    site = (long long)__float_as_int(w.w);
    f_first = frame_start;
  }
  for (int64_t f = f_first; f < f_stop; ++f) {
    const bool free_mover = site < 0;
    float nw[3];
    bool in_cube = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float step = __fmul_rn(gauss3(hash64(seed, i, f, 40 + d)), p.sigma_bulk);
      float v = __fadd_rn(pos[d], step);
      if (v < 0.0f) v = -v;
      if (v >= L) v = __fadd_rn(__fadd_rn(L, L), -v);
      v = fminf(v, top);
      v = fmaxf(v, 0.0f);
      nw[d] = v;
      in_cube = in_cube && (v > lo) && (v < hi);
    }
    const bool reject = in_cube && !bound;
    if (free_mover && !reject) { pos[0] = nw[0]; pos[1] = nw[1]; pos[2] = nw[2]; }
    if (bound) {
      const float u = u21(hash64(seed, i, f, 50), 0);
      const long long pick = (long long)((hash64(seed, i, f, 51) >> 11) % K);
      const bool on = site >= 0;
      const bool jump = on && (u < p_jump);
      const bool leave = on && !jump && (u < p.p_hop);
      const bool rebind = !on && (u < p.p_rebind);
      if (jump || rebind) site = pick;
      if (leave) site = -1;
      if (site >= 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d)
          pos[d] = __fadd_rn(centres[site * 3 + d], __fmul_rn(gauss3(hash64(seed, i, f, 30 + d)), p.sigma_site));
      }
    }
    if (f >= frame_start) {
      float* o = xyz + ((f - frame_start) * N + i) * 3;
      o[0] = pos[0]; o[1] = pos[1]; o[2] = pos[2];
    }
  }
  if (walker) walker[i] = make_float4(pos[0], pos[1], pos[2], __int_as_float((int)site));
}

}  // namespace

extern "C" int hopb_synth_generate_chunk(hopb_handle* h, void* stream, const hopb_synth_params* p, const float* centres,
                                         int64_t frame_start, int64_t n_out, float* xyz, float* walker, int resume) {
  HOPB_REQUIRE(h, p && centres && xyz, "null argument");
  HOPB_REQUIRE(h, p->n_atoms > 0 && p->n_sites > 0 && p->site_stride > 0 && frame_start >= 0 && n_out >= 0, "bad parameters");
  HOPB_REQUIRE(h, !resume || walker, "resume needs the walker buffer of the previous chunk");
  HOPB_REQUIRE(h, ((uintptr_t)walker & 15) == 0, "walker must be 16-byte aligned");
  if (n_out == 0) return HOPB_OK;
  HOPB_K(synth_kernel)<<<(unsigned)((p->n_atoms + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      *p, centres, frame_start, n_out, xyz, reinterpret_cast<float4*>(walker), resume);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

extern "C" int hopb_synth_generate(hopb_handle* h, void* stream, const hopb_synth_params* p, const float* centres,
                                   int64_t frame_start, int64_t n_out, float* xyz) {
  return hopb_synth_generate_chunk(h, stream, p, centres, frame_start, n_out, xyz, nullptr, 0);
}
