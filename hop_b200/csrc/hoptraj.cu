// Stages 3 + 4 fused (K3): coordinates -> site label (hoptraj x), orbit site (hoptraj y) and the
// per-atom transition state machine, one pass over the coordinates.
//
// Replaces HoppingTrajectory._coord2hop (hop/trajectory.py:403-468: numpy.digitize per axis, the
// rounded top-edge rule, a full copy of the map into a -1 padded buffer per frame and a per-atom
// Python list comprehension) and the per-frame, per-atom dict state machine of
// TransportNetwork._analyze_hops (hop/graph.py:212-254).
//
// Parallelisation: one thread per (atom, frame chunk).  A thread walks its chunk frame by frame;
// lanes are consecutive atoms, so the coordinate reads are 384 B contiguous per warp and frame, and
// the x / y writes are 128 B contiguous per warp, frame and plane -- the record layout of the
// hoptraj DCD ([frame][X plane | Y plane]).  The state a chunk inherits from the frames before it is
// not known when the chunk runs; hopb_machine.h shows that only the part of a chunk up to its
// second distinct site depends on it, so each (atom, chunk) leaves a 44-byte summary and
// hop_fixup_kernel stitches the chunks (and, across GPUs, the slabs) exactly.
//
// Roofline per atom-frame: 12 B read + 8 B written to HBM; the site-map gather (2 B out of a 32 B
// sector) is served by L2 (map16 of 201^3 cells = 16 MB).
#include "hopb_common.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kHopThreads = 128;
constexpr int kUnroll = 4;

struct EventSink {
  hopb_event* events;
  unsigned long long* n_events;
  unsigned long long cap;
};

__device__ __forceinline__ void emit_event(const EventSink& sink, int iatom, int s0, int s, int frame, int tau, int tbar) {
  const unsigned long long slot = atomicAdd(sink.n_events, 1ull);
  if (slot < sink.cap) {
    hopb_event e;
    e.frame = frame; e.iatom = iatom; e.s0 = s0; e.s = s; e.tau = tau; e.tbarrier = tbar;
    // 24-byte record, 8-byte aligned: three 8-byte stores
    int2* dst = reinterpret_cast<int2*>(sink.events + slot);
    dst[0] = make_int2(e.frame, e.iatom);
    dst[1] = make_int2(e.s0, e.s);
    dst[2] = make_int2(e.tau, e.tbarrier);
  }
}

__device__ __forceinline__ int lookup_label(float x, float y, float z, const HopbGridTab& tab, const float* ex,
                                            const float* ey, const float* ez, const int16_t* __restrict__ map16) {
  const int bx = hopb_bin_hop(x, tab.ax[0], ex);
  const int by = hopb_bin_hop(y, tab.ax[1], ey);
  const int bz = hopb_bin_hop(z, tab.ax[2], ez);
  const bool ok = (bx >= 1) & (bx <= tab.ax[0].n) & (by >= 1) & (by <= tab.ax[1].n) & (bz >= 1) & (bz <= tab.ax[2].n);
  if (!ok) return -1;  // the -1 shell of buffered_map (trajectory.py:176-177)
  const uint32_t cell = ((uint32_t)(bx - 1) * (uint32_t)tab.ax[1].n + (uint32_t)(by - 1)) * (uint32_t)tab.ax[2].n + (uint32_t)(bz - 1);
  return (int)__ldg(map16 + cell);
}

// SRC = 0: labels from coordinates + map; SRC = 1: labels from an existing hoptraj x column
template <int SRC>
__global__ void __launch_bounds__(kHopThreads)
hoptraj_kernel(const float* __restrict__ xyz, const float* __restrict__ labels_in, int64_t F, int64_t N, int64_t frame0,
               HopbGridTab tab, const float* __restrict__ table, const int16_t* __restrict__ map16,
               float* __restrict__ hop_x, float* __restrict__ hop_y, int64_t chunk_len, int32_t* __restrict__ summaries,
               EventSink sink) {
  extern __shared__ float s_tab[];
  if (SRC == 0) {
    for (int i = threadIdx.x; i < tab.table_len; i += blockDim.x) s_tab[i] = table[i];
    __syncthreads();
  }
  const float* ex = s_tab + tab.ax[0].off;
  const float* ey = s_tab + tab.ax[1].off;
  const float* ez = s_tab + tab.ax[2].off;

  const int64_t atom = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;
  if (atom >= N) return;
  const int64_t f_begin = (int64_t)chunk * chunk_len;
  int64_t f_end = f_begin + chunk_len;
  if (f_end > F) f_end = F;

  HopbSummary T = hopb_summary_empty();
  int ypre = 0, yorb = 0;
  bool yknown = false;
  const int iatom = (int)atom;
  auto emit = [&](int s0, int s, int frame, int tau, int tbar) { emit_event(sink, iatom, s0, s, frame, tau, tbar); };

  int64_t f = f_begin;
  for (; f + kUnroll <= f_end; f += kUnroll) {
    int lab[kUnroll];
    if (SRC == 0) {
      float cx[kUnroll], cy[kUnroll], cz[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const float* p = xyz + ((f + u) * N + atom) * 3;
        cx[u] = __ldg(p); cy[u] = __ldg(p + 1); cz[u] = __ldg(p + 2);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) lab[u] = lookup_label(cx[u], cy[u], cz[u], tab, ex, ey, ez, map16);
    } else {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) lab[u] = (int)__ldg(labels_in + (f + u) * N + atom);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int l = lab[u];
      hopb_push_frame(T, l, (int)(frame0 + f + u), emit);
      if (l > 0) { yorb = l; yknown = true; }
      if (!yknown) ++ypre;
      if (hop_x) hop_x[(f + u) * N + atom] = (float)l;
      if (hop_y) hop_y[(f + u) * N + atom] = (float)yorb;
    }
  }
  for (; f < f_end; ++f) {
    int l;
    if (SRC == 0) {
      const float* p = xyz + (f * N + atom) * 3;
      l = lookup_label(__ldg(p), __ldg(p + 1), __ldg(p + 2), tab, ex, ey, ez, map16);
    } else {
      l = (int)__ldg(labels_in + f * N + atom);
    }
    hopb_push_frame(T, l, (int)(frame0 + f), emit);
    if (l > 0) { yorb = l; yknown = true; }
    if (!yknown) ++ypre;
    if (hop_x) hop_x[f * N + atom] = (float)l;
    if (hop_y) hop_y[f * N + atom] = (float)yorb;
  }
  T.ypre = ypre;
  hopb_summary_store(summaries + ((int64_t)chunk * HOPB_SUMMARY_FIELDS) * N + atom, N, T);
}

__global__ void __launch_bounds__(256)
hop_combine_kernel(int64_t N, int n_chunks, int64_t chunk_len, int64_t F, const int32_t* __restrict__ summaries,
                   int32_t* __restrict__ slab) {
  const int64_t atom = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (atom >= N) return;
  HopbSummary R = hopb_summary_empty();
  int64_t done = 0;
  for (int c = 0; c < n_chunks; ++c) {
    const HopbSummary T = hopb_summary_load(summaries + ((int64_t)c * HOPB_SUMMARY_FIELDS) * N + atom, N);
    R = hopb_combine(R, T, (int)done);
    int64_t len = F - (int64_t)c * chunk_len;
    if (len > chunk_len) len = chunk_len;
    done += len;
  }
  hopb_summary_store(slab + atom, N, R);
}

__device__ __forceinline__ HopbState load_state(const int32_t* state, int64_t N, int64_t atom) {
  HopbState S;
  S.s0 = state[atom]; S.t0 = state[N + atom]; S.t1 = state[2 * N + atom]; S.on = state[3 * N + atom];
  return S;
}

__device__ __forceinline__ void store_state(int32_t* state, int64_t N, int64_t atom, const HopbState& S) {
  state[atom] = S.s0; state[N + atom] = S.t0; state[2 * N + atom] = S.t1; state[3 * N + atom] = S.on;
}

__global__ void __launch_bounds__(256)
hop_advance_kernel(int64_t N, int n, const int32_t* __restrict__ summaries, int init, int32_t* __restrict__ state) {
  const int64_t atom = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (atom >= N) return;
  HopbState S = init ? hopb_state_virgin() : load_state(state, N, atom);
  for (int c = 0; c < n; ++c) {
    const HopbSummary T = hopb_summary_load(summaries + ((int64_t)c * HOPB_SUMMARY_FIELDS) * N + atom, N);
    hopb_apply(S, T, [](int, int, int, int, int) {});
  }
  store_state(state, N, atom, S);
}

__global__ void __launch_bounds__(256)
hop_fixup_kernel(int64_t N, int n_chunks, int64_t chunk_len, int64_t F, const int32_t* __restrict__ summaries,
                 int32_t* __restrict__ state, float* __restrict__ hop_y, EventSink sink) {
  const int64_t atom = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (atom >= N) return;
  HopbState S = load_state(state, N, atom);
  const int iatom = (int)atom;
  auto emit = [&](int s0, int s, int frame, int tau, int tbar) { emit_event(sink, iatom, s0, s, frame, tau, tbar); };
  for (int c = 0; c < n_chunks; ++c) {
    const HopbSummary T = hopb_summary_load(summaries + ((int64_t)c * HOPB_SUMMARY_FIELDS) * N + atom, N);
    // orbit site inherited by the leading frames of the chunk (the kernel wrote 0 there)
    const int yin = S.s0 > 0 ? S.s0 : 0;
    if (hop_y && yin != 0 && T.ypre > 0) {
      const int64_t f0 = (int64_t)c * chunk_len;
      const float v = (float)yin;
      for (int k = 0; k < T.ypre; ++k) hop_y[(f0 + k) * N + atom] = v;
    }
    hopb_apply(S, T, emit);
  }
  store_state(state, N, atom, S);
}

// graph_alltransitions: set of (label[t-1], label[t]) pairs
__global__ void __launch_bounds__(256)
all_transitions_kernel(const float* __restrict__ hop_x, int64_t F, int64_t N, int n_labels, uint32_t* __restrict__ bitmap) {
  const int64_t total = (F - 1) * N;
  const int64_t W = (int64_t)n_labels + 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)hop_x[i];
    const int b = (int)hop_x[i + N];
    if (a < -1 || b < -1 || a >= n_labels || b >= n_labels) continue;
    const int64_t bit = (int64_t)(a + 1) * W + (b + 1);
    const uint32_t m = 1u << (bit & 31);
    uint32_t* w = bitmap + (bit >> 5);
    if ((*reinterpret_cast<volatile uint32_t*>(w) & m) == 0) atomicOr(w, m);
  }
}

}  // namespace

extern "C" {

static int launch_hoptraj(hopb_handle* h, void* stream, int src, const hopb_grid* g, const float* xyz,
                          const float* labels_in, int64_t F, int64_t N, int64_t frame0, const int16_t* map16,
                          float* hop_x, float* hop_y, int n_chunks, int64_t chunk_len, int32_t* summaries,
                          hopb_event* events, uint64_t* n_events, uint64_t cap) {
  HOPB_REQUIRE(h, F > 0 && N > 0 && summaries && n_events, "bad argument");
  HOPB_REQUIRE(h, n_chunks >= 1 && chunk_len >= 1 && (int64_t)n_chunks * chunk_len >= F &&
                      (int64_t)(n_chunks - 1) * chunk_len < F, "chunks must tile the slab exactly");
  HOPB_REQUIRE(h, n_chunks <= 65535, "too many chunks");
  HOPB_REQUIRE(h, frame0 >= 0 && frame0 + F < ((int64_t)1 << 31) && N < ((int64_t)1 << 31), "frame/atom index overflow");
  HOPB_REQUIRE(h, cap == 0 || events, "events buffer missing");
  EventSink sink{events, (unsigned long long*)n_events, (unsigned long long)cap};
  dim3 grid((unsigned)((N + kHopThreads - 1) / kHopThreads), (unsigned)n_chunks);
  if (src == 0) {
    HOPB_REQUIRE(h, g && xyz && map16, "null argument");
    const size_t smem = (size_t)g->tab.table_len * sizeof(float);
    hoptraj_kernel<0><<<grid, kHopThreads, smem, (cudaStream_t)stream>>>(xyz, nullptr, F, N, frame0, g->tab, g->d_table, map16,
                                                                       hop_x, hop_y, chunk_len, summaries, sink);
  } else {
    HOPB_REQUIRE(h, labels_in, "null argument");
    HopbGridTab tab;
    memset(&tab, 0, sizeof(tab));
    hoptraj_kernel<1><<<grid, kHopThreads, 0, (cudaStream_t)stream>>>(nullptr, labels_in, F, N, frame0, tab, nullptr, nullptr,
                                                                    hop_x, hop_y, chunk_len, summaries, sink);
  }
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hoptraj(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz, int64_t F, int64_t N, int64_t frame0,
                 const int16_t* map16, float* hop_x, float* hop_y, int n_chunks, int64_t chunk_len, int32_t* summaries,
                 hopb_event* events, uint64_t* n_events, uint64_t cap) {
  return launch_hoptraj(h, stream, 0, g, xyz, nullptr, F, N, frame0, map16, hop_x, hop_y, n_chunks, chunk_len, summaries,
                        events, n_events, cap);
}

int hopb_hopscan_labels(hopb_handle* h, void* stream, const float* hop_x, int64_t F, int64_t N, int64_t frame0, int n_chunks,
                        int64_t chunk_len, int32_t* summaries, hopb_event* events, uint64_t* n_events, uint64_t cap) {
  return launch_hoptraj(h, stream, 1, nullptr, nullptr, hop_x, F, N, frame0, nullptr, nullptr, nullptr, n_chunks, chunk_len,
                        summaries, events, n_events, cap);
}

int hopb_hop_combine(hopb_handle* h, void* stream, int64_t N, int n_chunks, int64_t chunk_len, int64_t F,
                     const int32_t* summaries, int32_t* slab_summary) {
  HOPB_REQUIRE(h, N > 0 && n_chunks >= 1 && summaries && slab_summary, "bad argument");
  hop_combine_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(N, n_chunks, chunk_len, F, summaries, slab_summary);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hop_advance(hopb_handle* h, void* stream, int64_t N, int n_summaries, const int32_t* summaries, int init,
                     int32_t* state) {
  HOPB_REQUIRE(h, N > 0 && n_summaries >= 0 && state && (n_summaries == 0 || summaries), "bad argument");
  hop_advance_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(N, n_summaries, summaries, init, state);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hop_fixup(hopb_handle* h, void* stream, int64_t N, int n_chunks, int64_t chunk_len, int64_t F,
                   const int32_t* summaries, int32_t* state, float* hop_y, hopb_event* events, uint64_t* n_events,
                   uint64_t cap) {
  HOPB_REQUIRE(h, N > 0 && n_chunks >= 1 && summaries && state && n_events, "bad argument");
  HOPB_REQUIRE(h, cap == 0 || events, "events buffer missing");
  (void)F;
  EventSink sink{events, (unsigned long long*)n_events, (unsigned long long)cap};
  hop_fixup_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(N, n_chunks, chunk_len, F, summaries, state,
                                                                             hop_y, sink);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_all_transitions(hopb_handle* h, void* stream, const float* hop_x, int64_t F, int64_t N, int n_labels,
                         uint32_t* bitmap) {
  HOPB_REQUIRE(h, hop_x && bitmap && F >= 1 && N > 0 && n_labels >= 1, "bad argument");
  if (F < 2) return HOPB_OK;
  all_transitions_kernel<<<hopb_grid_for((F - 1) * N, 256, h->num_sms, 8), 256, 0, (cudaStream_t)stream>>>(hop_x, F, N, n_labels, bitmap);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

}  // extern "C"
