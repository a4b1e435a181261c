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
// (L2 evict-first hints on the streamed loads / stores were measured and made this kernel 5-25 %
// slower on B200; plain ld.global.nc / st.global it is.)
//
// Roofline per atom-frame: 12 B read + 8 B written to HBM.  The pipeline moves less: stage 1 leaves a
// 4-byte brick index per atom-frame which is read here instead of the coordinates (through TMA boxes, see
// kTmaStages), and the planes are written as int16 (hopb_hoptraj_i16: labels are int16 map values or -1) -- 4 + 4
// bytes; the site-map gather (2 B out of a 32 B sector) is served by shared-memory block maps and L2 (map16 of
// 201^3 cells = 16 MB).  Measured: profiles/r2_k3_decomposition.md.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "hopb_common.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kHopThreads = 256;
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

// Events are rare (~1e-3 per atom-frame) but every one used to pay a returning atomic on the one global counter --
// a ~1 us round trip on the emitting thread's critical path, 2.3e5 of them serialised on one L2 address per launch.
// Each warp now parks its events in a small shared-memory queue (a shared-memory atomic hands out the slot) and
// reserves global space for the whole queue with one atomic when the warp next reaches a common point.
constexpr int kEvQueue = 8;
struct EvStage {
  int* cnt;            // shared: events queued by this warp since the last flush (may exceed kEvQueue)
  hopb_event* q;       // shared: [kEvQueue]
};

__device__ __forceinline__ void emit_staged(const EventSink& sink, const EvStage& st, int iatom, int s0, int s, int frame, int tau,
                                            int tbar) {
  const int slot = atomicAdd(st.cnt, 1);
  if (slot < kEvQueue) {
    hopb_event e;
    e.frame = frame; e.iatom = iatom; e.s0 = s0; e.s = s; e.tau = tau; e.tbarrier = tbar;
    st.q[slot] = e;
  } else {
    emit_event(sink, iatom, s0, s, frame, tau, tbar);     // queue full: straight to the global buffer
  }
}

// all lanes of `lanes` call this together
__device__ __forceinline__ void flush_staged(const EventSink& sink, const EvStage& st, unsigned lanes) {
  __syncwarp(lanes);
  int n = *reinterpret_cast<volatile int*>(st.cnt);
  if (n == 0) return;
  n = n < kEvQueue ? n : kEvQueue;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(lanes) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(sink.n_events, (unsigned long long)n);
  base = __shfl_sync(lanes, base, leader);
  // 6 ints per event: the active lanes copy the queue word by word
  const int* src = reinterpret_cast<const int*>(st.q);
  const int rank = __popc(lanes & ((1u << lane) - 1)), width = __popc(lanes);
  for (int w = rank; w < n * 6; w += width) {
    const unsigned long long ev = base + (unsigned long long)(w / 6);
    if (ev < sink.cap) reinterpret_cast<int*>(sink.events + ev)[w % 6] = src[w];
  }
  __syncwarp(lanes);
  if (lane == leader) *st.cnt = 0;
  __syncwarp(lanes);
}

struct AxisReg {
  float lo, inv, nf, hop_lo, hop_hi;
  int n;
  const HopbPair* pairs;
};

// ---- bulk-copy (TMA) staging of the brick-index stream ----------------------------------------------------------------
// ncu put ~60 % of the kernel's stall samples on the first use of a loaded brick index and of a gathered label: with 64
// registers per thread a block of 8 frames is all a thread can keep in flight.  The TMA variant moves the indices of
// the next two blocks into shared memory with cp.async.bulk.tensor (one 32-atom x 8-frame box of the [frames][atoms]
// array per block and warp, issued by one lane, completion on a per-warp mbarrier): they cost no registers while in
// flight.  kTmaAhead additionally resolves the labels of block b+1 (its L2 gathers) before the state machine runs
// block b.
constexpr int kTmaStages = 2;
#ifndef HOPB_K3_TMA_PREFETCH
#define HOPB_K3_TMA_PREFETCH 0
#endif
constexpr int kTmaPrefetch = HOPB_K3_TMA_PREFETCH;   // boxes ahead of the load that are pulled into L2; measured with 4: 0.98 instead of 0.77 ms, hence off
#ifndef HOPB_K3_TMA_AHEAD
#define HOPB_K3_TMA_AHEAD 0
#endif
constexpr bool kTmaAhead = HOPB_K3_TMA_AHEAD != 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" : : "r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" : : "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" : : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// one box of the [frames][atoms] brick-index array (32 atoms x DEEP frames) -> shared memory; columns beyond the last
// atom arrive as zeros and count towards the transaction bytes
__device__ __forceinline__ void tma_box_prefetch_l2(const CUtensorMap* tmap, int c_atom, int c_frame) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" : : "l"(tmap), "r"(c_atom), "r"(c_frame) : "memory");
}

__device__ __forceinline__ void tma_box_g2s(void* dst_smem, const CUtensorMap* tmap, int c_atom, int c_frame, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               : : "r"(smem_u32(dst_smem)), "l"(tmap), "r"(c_atom), "r"(c_frame), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ AxisReg load_axis(const HopbAxis& a, const HopbPair* s_pairs) {
  AxisReg r;
  r.lo = a.lo; r.inv = a.inv; r.nf = (float)a.n; r.hop_lo = a.hop_lo; r.hop_hi = a.hop_hi;
  r.n = a.n; r.pairs = s_pairs + a.off2;
  return r;
}

// _coord2hop bin of one coordinate, 0-based; anything outside [0, n) is an outlier
__device__ __forceinline__ int hop_bin(float x, const AxisReg& a) {
  float g = (x - a.lo) * a.inv;
  g = fminf(fmaxf(g, -1.0f), a.nf);
  int b = __float2int_rd(g) + 1;
  const HopbPair e = a.pairs[b];
  b += (x >= e.hi) ? 1 : 0;
  b -= (x < e.lo) ? 1 : 0;
  b -= ((x >= a.hop_lo) & (x <= a.hop_hi)) ? 1 : 0;
  return b - 1;
}

template <bool FAST>
__device__ __forceinline__ uint32_t pack_cell(float x, float y, float z, const HopbGridTab& tab, const AxisReg& ax,
                                              const AxisReg& ay, const AxisReg& az, const float* s_tab, int cny, int cnz) {
  hopb_wrap_point(x, y, z, tab);
  int bx, by, bz;
  if (FAST) {
    bx = hop_bin(x, ax); by = hop_bin(y, ay); bz = hop_bin(z, az);
  } else {
    bx = hopb_bin_hop(x, tab.ax[0], s_tab + tab.ax[0].off) - 1;
    by = hopb_bin_hop(y, tab.ax[1], s_tab + tab.ax[1].off) - 1;
    bz = hopb_bin_hop(z, tab.ax[2], s_tab + tab.ax[2].off) - 1;
  }
  const bool ok = ((unsigned)bx < (unsigned)ax.n) & ((unsigned)by < (unsigned)ay.n) & ((unsigned)bz < (unsigned)az.n);
  return ok ? hopb_brick_index(bx, by, bz, cny, cnz) : 0xFFFFFFFFu;
}

// how a brick index becomes a label
enum { MAP_GATHER = 0,   // always gather from the bricked int16 map (L2)
       MAP_COARSE = 1,   // 2 bits per 4x4x4 brick in shared memory: 0 / 1 = uniformly interstitial / bulk, 3 = mixed
       MAP_FINE = 2,     // 1 bit per 2x2x2 sub-brick in shared memory: all 8 cells carry fine_value
       MAP_CELL2 = 3 };  // grids whose block maps do not fit: 2 bits per cell in L2 (0, 1, 3 = a hydration site)

struct MapView {
  const int16_t* brick;     // bricked site map
  const uint32_t* s_map;    // block map in shared memory (coarse or fine); MAP_CELL2: the 2-bit cell map in global memory
  int fine_value;
};

// Site label of a brick index.  The block map answers for the cells of uniform (sub-)bricks; only the
// others gather from the bricked int16 map in L2.  A gather costs one 32-byte sector per lane, and
// sectors -- not bytes -- are what the kernel is short of, so the fine map (2.6 % of the bulk cells
// of a 0.5 A water grid sit in a mixed 2x2x2 sub-brick, 20 % in a mixed 4x4x4 brick) is preferred
// whenever it fits into shared memory.  Selects and one predicated load: no divergence, and the
// load is the last thing that writes l so nothing waits for it before the label is used.
// PLAIN: the caller has checked that c is an ordinary brick index (no outlier, no bookkeeping bit) -- a warp vote per
// block of frames in the TMA path, which takes four instructions per frame out of the look-up.
template <int MAP, bool PLAIN = false>
__device__ __forceinline__ int label_of(uint32_t c, const MapView& m) {
  const bool out = PLAIN ? false : c == 0xFFFFFFFFu;  // the -1 shell of buffered_map (trajectory.py:176-177)
  const uint32_t cc = PLAIN ? c : (out ? 0u : (c & 0x3FFFFFFFu));  // bit 30: bookkeeping of the slabbed histogram
  int l;
  bool need;
  if (MAP == MAP_FINE) {
    const uint32_t sb = cc >> 3;
    const bool uni = (m.s_map[sb >> 5] >> (sb & 31u)) & 1u;
    l = m.fine_value;
    need = !uni;
  } else if (MAP == MAP_COARSE) {
    const uint32_t cb = cc >> 6;
    const uint32_t st = (m.s_map[cb >> 4] >> ((cb & 15u) * 2u)) & 3u;
    l = (int)st;
    need = st == 3u;
  } else if (MAP == MAP_CELL2) {
    // 16 cells per word, brick order: a 16x smaller gather target than the int16 map (16 MB instead
    // of 130 MB at 402^3), so it stays in L2; only atoms on hydration sites go on to the int16 map
    const uint32_t code = (__ldg(m.s_map + (cc >> 4)) >> ((cc & 15u) * 2u)) & 3u;
    l = (int)code;
    need = code == 3u;
  } else {
    l = 0;
    need = true;
  }
  l = out ? -1 : l;
  if (need && !out) l = (int)__ldg(m.brick + cc);
  return l;
}

// XY: both output planes are present (the product path), so the stores need no pointer tests;
// MAP: which block map sits in shared memory.
// Work items are (warp of 32 consecutive atoms, chunk), numbered chunk-major and dealt to the warps
// of the grid in order, so the CTA size (256 with a small block map, up to 1024 with a large one: one
// copy of the map per CTA) does not change which items exist.
// PT: element type of the two output planes -- float (the record type of the hoptraj DCD) or int16_t (the same labels
// at a quarter of the kernel's DRAM traffic less: labels are int16 map values or -1; the pipeline's default).
template <int SRC, bool FAST, bool XY, int MAP, bool STRIDED = false, int DEEP = 8, typename PT = float, bool TMA = false>
__global__ void __launch_bounds__(1024, 1)
hoptraj_kernel(const float* __restrict__ xyz, const float* __restrict__ labels_in, const uint32_t* __restrict__ cells,
               int64_t F, int64_t N, int64_t ld_arg, int64_t frame0, HopbGridTab tab, const float* __restrict__ table,
               const HopbPair* __restrict__ table2, const int16_t* __restrict__ brick,
               const uint32_t* __restrict__ blockmap, int map_words, int fine_value, uint32_t fine_auto_min,
               PT* __restrict__ hop_x,
               PT* __restrict__ hop_y, int n_chunks, int64_t chunk_len, int32_t* __restrict__ summaries,
               EventSink sink, const __grid_constant__ CUtensorMap cells_map) {
  // rows of the data arrays are N atoms apart, or ld_arg when the atoms are a column range of wider arrays (a
  // variant of its own: the common case keeps one value less alive at the 64-register limit)
  const int64_t ld = STRIDED ? ld_arg : N;
  __shared__ int s_ev_cnt[32];
  __shared__ hopb_event s_ev_q[32][kEvQueue];
  if (threadIdx.x < 32) s_ev_cnt[threadIdx.x] = 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HopbPair* s_pairs = reinterpret_cast<HopbPair*>(smem_raw);
  float* s_tab = reinterpret_cast<float*>(s_pairs + (SRC == 0 ? tab.table2_len : 0));
  uint32_t* s_map = reinterpret_cast<uint32_t*>(s_tab + ((SRC == 0 && !FAST) ? tab.table_len : 0));
  if (SRC == 0) {
    for (int i = threadIdx.x; i < tab.table2_len; i += blockDim.x) s_pairs[i] = table2[i];
    if (!FAST)
      for (int i = threadIdx.x; i < tab.table_len; i += blockDim.x) s_tab[i] = table[i];
  }
  if (SRC != 1 && MAP == MAP_FINE) {
    // two maps (sub-brick uniformly 0 / uniformly 1) and the number of bulk cells behind them
    if (fine_value < 0) fine_value = blockmap[2 * map_words] >= fine_auto_min ? 1 : 0;
    blockmap += fine_value * map_words;
  }
  if (SRC != 1 && (MAP == MAP_COARSE || MAP == MAP_FINE))
    for (int i = threadIdx.x; i < map_words; i += blockDim.x) s_map[i] = blockmap[i];
  // TMA staging: per warp kTmaStages x DEEP rows of 32 brick indices and one mbarrier per stage
  uint32_t* w_stage = nullptr;
  uint64_t* w_bar = nullptr;
  if (TMA) {
    const int words_before = (int)(reinterpret_cast<unsigned char*>(s_map + ((MAP == MAP_COARSE || MAP == MAP_FINE) ? map_words : 0)) - smem_raw);
    unsigned char* st0 = smem_raw + ((words_before + 127) & ~127);
    const int n_warps = blockDim.x >> 5;
    w_stage = reinterpret_cast<uint32_t*>(st0) + (threadIdx.x >> 5) * (kTmaStages * DEEP * 32);
    w_bar = reinterpret_cast<uint64_t*>(st0 + (size_t)n_warps * kTmaStages * DEEP * 128) + (threadIdx.x >> 5) * kTmaStages;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int q = 0; q < kTmaStages; ++q) mbar_init(w_bar + q, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" : : : "memory");
    }
  }
  __syncthreads();
  const AxisReg ax = load_axis(tab.ax[0], s_pairs);
  const AxisReg ay = load_axis(tab.ax[1], s_pairs);
  const AxisReg az = load_axis(tab.ax[2], s_pairs);
  MapView mv;
  mv.brick = brick; mv.fine_value = fine_value;
  mv.s_map = (SRC == 1 || MAP == MAP_GATHER) ? nullptr : (MAP == MAP_CELL2 ? blockmap : s_map);
  const int cny = (tab.ax[1].n + 3) >> 2, cnz = (tab.ax[2].n + 3) >> 2;

  const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t atom_warps = (N + 31) >> 5;
  const int chunk = (int)(item / atom_warps);
  const int64_t atom = (item - (int64_t)chunk * atom_warps) * 32 + (threadIdx.x & 31);
  if (chunk >= n_chunks || atom >= N) return;
  const unsigned lanes = __activemask();
  // frame numbers inside the slab fit 32 bits (checked by the launcher)
  const int f_begin = (int)((int64_t)chunk * chunk_len);
  const int f_end = (int)((int64_t)f_begin + chunk_len > F ? F : (int64_t)f_begin + chunk_len);

  HopbSummary T = hopb_summary_empty();
  int ypre = 0, yorb = 0;
  bool yknown = false;
  const int iatom = (int)atom;
  const EvStage evs{&s_ev_cnt[threadIdx.x >> 5], s_ev_q[threadIdx.x >> 5]};
  auto emit = [&](int s0, int s, int frame, int tau, int tbar) { emit_staged(sink, evs, iatom, s0, s, frame, tau, tbar); };

  // running pointers (one frame = N elements further)
  // data rows are `ld` elements apart (ld = N unless the atoms are a column range of a wider array)
  PT* px = hop_x ? hop_x + (int64_t)f_begin * ld + atom : nullptr;
  PT* py = hop_y ? hop_y + (int64_t)f_begin * ld + atom : nullptr;
  auto store = [&](int l) {
    if (XY) {
      *px = (PT)l; px += ld;
      *py = (PT)yorb; py += ld;
    } else {
      if (px) { *px = (PT)l; px += ld; }
      if (py) { *py = (PT)yorb; py += ld; }
    }
  };
  // label of frame f, whatever the source
  auto label_at = [&](int64_t fr) -> int {
    if (SRC == 1) return (int)__ldg(labels_in + fr * ld + atom);
    uint32_t c;
    if (SRC == 0) {
      const float* p = xyz + (fr * ld + atom) * 3;
      c = pack_cell<FAST>(__ldg(p), __ldg(p + 1), __ldg(p + 2), tab, ax, ay, az, s_tab, cny, cnz);
    } else {
      c = __ldg(cells + fr * ld + atom);
    }
    return label_of<MAP>(c, mv);
  };

  // Two regimes per (atom, chunk).  Until a first real site has been seen the generic summary
  // machine of hopb_machine.h runs (kinds EMPTY / PRE / SPEC); from then on the state is "current
  // site cur, on/off, last exit t1" in registers (regime A).  Frames go through in blocks: entering
  // another site is the only rare event (~1e-3 per atom-frame), so when no label of a block is a site
  // other than cur (A), or no label is a site at all and none has been seen yet (atoms that spend the
  // whole chunk in the interstitial, e.g. ions), the block is a few selects and two stores per
  // frame; otherwise it replays frame by frame.
  bool inA = false;
  int cur = 0, on = 1, t1 = 0;
  auto stepA = [&](int l, int t) {
    const bool same = l == cur, zero = l == 0;
    if (l > 0 && !same) {
      const int tx = on ? t : t1;
      if (T.kind == HOPB_SPEC) {          // second site of the chunk: the open event E2
        T.t1_2 = tx; T.A2 = l; T.f2 = t; T.kind = HOPB_FULL;
      } else {
        emit(cur, l, t, tx - T.t0, t - tx);
      }
      cur = l; T.t0 = t; on = 1; yorb = l;
    } else {
      t1 = (zero && on) ? t : t1;
      on = same ? 1 : (zero ? 0 : on);   // l == -1 (outlier): nothing changes
    }
    store(l);
  };
  auto generic = [&](int l, int t) {
    if (inA) { stepA(l, t); return; }
    hopb_push_frame(T, l, t, emit);
    if (l > 0) { yorb = l; yknown = true; }
    if (!yknown) ++ypre;
    store(l);
    if (T.kind >= HOPB_SPEC && yknown) {
      inA = true;
      cur = T.kind == HOPB_SPEC ? T.A1 : T.s0;
      on = T.on; t1 = T.t1;
      // these four never change again: out of the registers now (the kernel sits at the 64-register
      // limit of 1024 resident threads per SM)
      int32_t* sp = summaries + ((int64_t)chunk * HOPB_SUMMARY_FIELDS) * N + atom;
      sp[1 * N] = T.z; sp[2 * N] = T.A1; sp[3 * N] = T.f1; sp[10 * N] = ypre;
    }
  };
  const uint32_t Nu = (uint32_t)ld;
  // regime A, one block of K frames at global frame t; lab[] are its labels
  auto blockA = [&](const int* lab, auto kc, int, int t) {
    constexpr int K = decltype(kc)::value;
    bool rare = false;
#pragma unroll
    for (int u = 0; u < K; ++u) rare |= (lab[u] > 0) & (lab[u] != cur);
    if (!rare && XY) {
      const PT yf = (PT)yorb;
#pragma unroll
      for (int u = 0; u < K; ++u) {
        const int l = lab[u];
        const bool same = l == cur, zero = l == 0;
        t1 = (zero && on) ? t + u : t1;
        on = same ? 1 : (zero ? 0 : on);
        px[(uint64_t)Nu * (uint32_t)u] = (PT)l;
        py[(uint64_t)Nu * (uint32_t)u] = yf;
      }
      px += (uint64_t)Nu * K;
      py += (uint64_t)Nu * K;
    } else {
#pragma unroll
      for (int u = 0; u < K; ++u) stepA(lab[u], t + u);
    }
  };
  // before regime A, one block of K frames starting at slab frame fb: quick when there is nothing
  // but interstitial / outlier frames and no site has been seen so far (global frame 0 is special: an
  // outlier there is the pseudo-site -1, hop/graph.py:206-208)
  auto blockB = [&](const int* lab, auto kc, int fb, int t) {
    constexpr int K = decltype(kc)::value;
    bool quiet = (T.kind <= HOPB_PRE) & (t != 0);
#pragma unroll
    for (int u = 0; u < K; ++u) quiet &= lab[u] <= 0;
    if (quiet && XY) {
      const PT yf = (PT)yorb;
#pragma unroll
      for (int u = 0; u < K; ++u) {
        const int l = lab[u];
        const bool first0 = (l == 0) & (T.kind == HOPB_EMPTY);
        T.z = first0 ? t + u : T.z;
        T.kind = first0 ? (int)HOPB_PRE : T.kind;
        px[(uint64_t)Nu * (uint32_t)u] = (PT)l;
        py[(uint64_t)Nu * (uint32_t)u] = yf;
      }
      ypre += K;
      px += (uint64_t)Nu * K;
      py += (uint64_t)Nu * K;
    } else {
#pragma unroll 1
      for (int u = 0; u < K; ++u) generic(label_at(fb + u), t + u);
    }
  };
  // lanes of both kinds in one warp
  auto blockU = [&](const int* lab, auto kc, int fb, int t) {
    if (inA) blockA(lab, kc, fb, t); else blockB(lab, kc, fb, t);
  };

  int f = f_begin;
  // TMA pipeline state (see above): lab_cur = labels of the block at f; blocks are numbered from the chunk start
  int lab_cur[DEEP];
  int blk = 0;
  const int n_blk = (f_end - f_begin) / DEEP;
  const int lane_id = threadIdx.x & 31;
  const int leader = __ffs(lanes) - 1;
  const int64_t atom0 = atom - lane_id;
  auto tma_issue = [&](int b) {      // box of block b -> stage b % kTmaStages
    if (lane_id == leader) {
      uint64_t* bar = w_bar + (b % kTmaStages);
      mbar_expect_tx(bar, (uint32_t)(DEEP * 32 * 4));
      tma_box_g2s(w_stage + (b % kTmaStages) * (DEEP * 32), &cells_map, (int)atom0, f_begin + b * DEEP, bar);
      // the shared-memory ring holds two boxes (the block map takes the rest of the SM's shared memory); boxes further
      // ahead are pulled into L2 so that their load, when its turn comes, is an L2 hit
      if (kTmaPrefetch > 0 && b + kTmaPrefetch < n_blk) tma_box_prefetch_l2(&cells_map, (int)atom0, f_begin + (b + kTmaPrefetch) * DEEP);
    }
  };
  auto tma_consume = [&](int b, int* lab) {      // labels of block b (gathers in flight on return), its stage refilled
    mbar_wait(w_bar + (b % kTmaStages), (uint32_t)((b / kTmaStages) & 1));
    const uint32_t* st = w_stage + (b % kTmaStages) * (DEEP * 32) + lane_id;
    uint32_t c[DEEP];
    uint32_t any = 0u;
#pragma unroll
    for (int u = 0; u < DEEP; ++u) { c[u] = st[u * 32]; any |= c[u]; }
    if (!__any_sync(lanes, (any >> 30) != 0u)) {      // no outlier, no bookkeeping bit in this block of the warp
#pragma unroll
      for (int u = 0; u < DEEP; ++u) lab[u] = label_of<MAP, true>(c[u], mv);
    } else {
#pragma unroll
      for (int u = 0; u < DEEP; ++u) lab[u] = label_of<MAP>(c[u], mv);
    }
    __syncwarp(lanes);
    if (b + kTmaStages < n_blk) tma_issue(b + kTmaStages);
  };
  if (TMA && SRC == 2 && n_blk > 0) {
    if (kTmaPrefetch > 0 && lane_id == leader) {
      for (int b = 2; b < kTmaPrefetch && b < n_blk; ++b) tma_box_prefetch_l2(&cells_map, (int)atom0, f_begin + b * DEEP);
    }
    tma_issue(0);
    if (n_blk > 1) tma_issue(1);
    if (kTmaAhead) tma_consume(0, lab_cur);
  }
  // stream whole blocks of frames through `body` while `more()` (warp-uniform) holds and the chunk lasts
  auto stream_blocks = [&](auto&& body, auto&& more) {
    if (!more()) return;
    if (SRC == 0) {
      while (f + kUnroll <= f_end) {
        float cx[kUnroll], cy[kUnroll], cz[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const float* p = xyz + ((int64_t)(f + u) * ld + atom) * 3;
          cx[u] = __ldg(p); cy[u] = __ldg(p + 1); cz[u] = __ldg(p + 2);
        }
        int lab[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) lab[u] = label_of<MAP>(pack_cell<FAST>(cx[u], cy[u], cz[u], tab, ax, ay, az, s_tab, cny, cnz), mv);
        body(lab, std::integral_constant<int, kUnroll>(), f, (int)(frame0 + f));
        f += kUnroll;
        __syncwarp(lanes);
        if (*reinterpret_cast<volatile int*>(evs.cnt) >= kEvQueue / 2) flush_staged(sink, evs, lanes);
        if (!more()) break;
      }
    } else if (TMA && SRC == 2) {
      if (kTmaAhead) {
        // labels of block b + 1 are resolved (their gathers issued) before the state machine runs block b
        while (blk < n_blk) {
          int lab_next[DEEP];
          const bool nxt = blk + 1 < n_blk;
          if (nxt) tma_consume(blk + 1, lab_next);
          body(lab_cur, std::integral_constant<int, DEEP>(), f, (int)(frame0 + f));
          f += DEEP;
          ++blk;
          if (nxt) {
#pragma unroll
            for (int u = 0; u < DEEP; ++u) lab_cur[u] = lab_next[u];
          }
          __syncwarp(lanes);
          if (*reinterpret_cast<volatile int*>(evs.cnt) >= kEvQueue / 2) flush_staged(sink, evs, lanes);
          if (!more()) break;
        }
      } else {
        while (blk < n_blk) {
          int lab[DEEP];
          tma_consume(blk, lab);
          body(lab, std::integral_constant<int, DEEP>(), f, (int)(frame0 + f));
          f += DEEP;
          ++blk;
          __syncwarp(lanes);
          if (*reinterpret_cast<volatile int*>(evs.cnt) >= kEvQueue / 2) flush_staged(sink, evs, lanes);
          if (!more()) break;
        }
      }
    } else {
      // 4-byte inputs (brick indices or labels) in blocks of kDeep frames: as soon as the labels of a
      // block are resolved its input registers are reloaded with the next block, so those loads are
      // in flight while the state machine consumes the current labels.  (Going one block deeper --
      // gathers of block b+1 and loads of block b+2 in flight during block b -- was measured: slower.)
      constexpr int kDeep = DEEP;
      if (f + kDeep > f_end) return;
      const uint32_t* src = (SRC == 2 ? cells : reinterpret_cast<const uint32_t*>(labels_in)) + (int64_t)f * ld + atom;
      uint32_t v[kDeep];
#pragma unroll
      for (int u = 0; u < kDeep; ++u) v[u] = __ldg(src + (uint64_t)Nu * (uint32_t)u);
      src += (uint64_t)Nu * kDeep;
      while (f + kDeep <= f_end) {
        int lab[kDeep];
#pragma unroll
        for (int u = 0; u < kDeep; ++u) lab[u] = SRC == 2 ? label_of<MAP>(v[u], mv) : (int)__uint_as_float(v[u]);
        if (f + 2 * kDeep <= f_end) {
#pragma unroll
          for (int u = 0; u < kDeep; ++u) v[u] = __ldg(src + (uint64_t)Nu * (uint32_t)u);
          src += (uint64_t)Nu * kDeep;
        }
        body(lab, std::integral_constant<int, kDeep>(), f, (int)(frame0 + f));
        f += kDeep;
        __syncwarp(lanes);
        if (*reinterpret_cast<volatile int*>(evs.cnt) >= kEvQueue / 2) flush_staged(sink, evs, lanes);
        if (!more()) break;
      }
    }
  };
  // Three loops, chosen per warp: no lane has seen a site yet; some have; all have.  (One loop that
  // serves both kinds of lanes needs more than the 64 registers there are, so the common all-A case
  // gets its own lean loop; per-lane loops instead would make a mixed warp walk its chunk twice.)
  stream_blocks(blockB, [&]() { return !__any_sync(lanes, inA); });
  stream_blocks(blockU, [&]() { return !__all_sync(lanes, inA); });
  stream_blocks(blockA, [&]() { return true; });
  for (; f < f_end; ++f) generic(label_at(f), (int)(frame0 + f));
  flush_staged(sink, evs, lanes);
  int32_t* sp = summaries + ((int64_t)chunk * HOPB_SUMMARY_FIELDS) * N + atom;
  if (inA) {
    if (T.kind == HOPB_FULL) T.s0 = cur;
    sp[0 * N] = T.kind | (on << 2);
    sp[4 * N] = T.A2; sp[5 * N] = T.f2; sp[6 * N] = T.t1_2;
    sp[7 * N] = T.s0; sp[8 * N] = T.t0; sp[9 * N] = t1;
  } else {
    T.ypre = ypre;
    hopb_summary_store(sp, N, T);
  }
}

// Bricked copy of the site map + its block maps: one thread per 4x4x4 brick reads its (up to) 64
// cells from the C-ordered int16 map, writes them as one 128-byte line (sub-brick order, see
// hopb_brick_index) and classifies the brick -- coarse: 0 / 1 = every cell interstitial / bulk,
// 3 = mixed, 16 bricks per 32-bit word; fine: one bit per 2x2x2 sub-brick, set when all of its cells
// carry the same label v, one byte per brick, one map for v = 0 (fine[b]) and one for v = 1
// (fine[nb_pad + b]).  The number of cells with label 1 goes to the counter behind the two maps:
// stage 3 uses the v = 1 map when the bulk site covers at least 1/8 of the grid (then most atoms
// sit in it), else the v = 0 map.  Cells beyond the grid are never looked up and do not count.
__global__ void __launch_bounds__(256)
brick_map_kernel(const int16_t* __restrict__ map16, int nx, int ny, int nz, int cnx, int cny, int cnz,
                 int16_t* __restrict__ brick, uint32_t* __restrict__ coarse, uint8_t* __restrict__ fine, int64_t nb_pad,
                 uint4* __restrict__ cell2) {
  const int64_t n_blocks = (int64_t)cnx * cny * cnz;
  const int64_t n_round = (n_blocks + 31) / 32 * 32;
  const int lane = threadIdx.x & 31;
  uint32_t n_ones = 0;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_round; b += (int64_t)gridDim.x * blockDim.x) {
    uint32_t st = 0;
    if (b < n_blocks) {
      const int ck = (int)(b % cnz);
      const int64_t r = b / cnz;
      const int cj = (int)(r % cny);
      const int ci = (int)(r / cny);
      int first = -2;
      bool uniform = true;
      uint32_t not0 = 0, not1 = 0;  // bit s: sub-brick s has a cell that is not 0 / not 1
      uint32_t c2[4] = {0u, 0u, 0u, 0u};  // 2 bits per cell, brick order: 0, 1, 3 = any other label
      int16_t* out = brick + b * 64;
      for (int di = 0; di < 4; ++di) {
        const int i = ci * 4 + di;
        for (int dj = 0; dj < 4; ++dj) {
          const int j = cj * 4 + dj;
          const bool row_ok = i < nx && j < ny;
          const int16_t* row = map16 + ((int64_t)i * ny + j) * nz;
          int16_t v4[4];
          uint32_t part = 0;
#pragma unroll
          for (int dk = 0; dk < 4; ++dk) {
            const int k = ck * 4 + dk;
            int16_t v = 0;
            if (row_ok && k < nz) {
              v = row[k];
              if (first == -2) first = v;
              else if (v != first) uniform = false;
              const uint32_t sub = 1u << (((di & 2) << 1) | (dj & 2) | (dk >> 1));
              if (v != 0) not0 |= sub;
              if (v != 1) not1 |= sub; else ++n_ones;
            }
            v4[dk] = v;
            const uint32_t code = v == 0 ? 0u : (v == 1 ? 1u : 3u);
            part |= code << (2 * (((dk & 2) << 2) | ((di & 1) << 2) | ((dj & 1) << 1) | (dk & 1)));
          }
          // position inside the brick = (i1 j1 | k1 i0 j0 k0): word i1 j1, 2-bit slot k1 i0 j0 k0
          const int widx = (di & 2) | (dj >> 1);
          if (widx == 0) c2[0] |= part; else if (widx == 1) c2[1] |= part; else if (widx == 2) c2[2] |= part; else c2[3] |= part;
          // (di, dj) fixes i1 j1 . i0 j0 .; dk = 0,1 and dk = 2,3 are two runs of two cells 8 apart
          int16_t* o = out + (((di & 2) << 4) | ((dj & 2) << 3) | ((di & 1) << 2) | ((dj & 1) << 1));
          *reinterpret_cast<short2*>(o) = make_short2(v4[0], v4[1]);
          *reinterpret_cast<short2*>(o + 8) = make_short2(v4[2], v4[3]);
        }
      }
      st = (uniform && (first == 0 || first == 1)) ? (uint32_t)first : 3u;
      if (fine) {
        fine[b] = (uint8_t)(~not0 & 0xFFu);
        fine[nb_pad + b] = (uint8_t)(~not1 & 0xFFu);
      }
      if (cell2) cell2[b] = make_uint4(c2[0], c2[1], c2[2], c2[3]);
    }
    uint32_t w = st << ((lane & 15) * 2);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) w |= __shfl_xor_sync(kFull, w, o);
    if ((lane & 15) == 0 && b < n_blocks) coarse[b >> 4] = w;
  }
  if (fine) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_ones += __shfl_xor_sync(kFull, n_ones, o);
    if (lane == 0 && n_ones) atomicAdd(reinterpret_cast<uint32_t*>(fine + 2 * nb_pad), n_ones);
  }
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

template <typename PT>
__global__ void __launch_bounds__(256)
hop_fixup_kernel(int64_t N, int64_t ld, int n_chunks, int64_t chunk_len, int64_t F, const int32_t* __restrict__ summaries,
                 int32_t* __restrict__ state, PT* __restrict__ hop_y, EventSink sink) {
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
      const PT v = (PT)yin;
      for (int k = 0; k < T.ypre; ++k) hop_y[(f0 + k) * ld + atom] = v;
    }
    hopb_apply(S, T, emit);
  }
  store_state(state, N, atom, S);
}

// hoptraj planes as int16 for the trip to the host: labels are integers in [-1, 32767] (hop's site map is int16,
// hop/sitemap.py:518-519), so the conversion is lossless and halves the download; the float32 of the DCD record is
// restored on the host when a file is written.
__global__ void __launch_bounds__(256)
planes_to_int16_kernel(const float4* __restrict__ in, int64_t n4, const float* __restrict__ in_tail, int n_tail, short4* __restrict__ out,
                       int16_t* __restrict__ out_tail) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    out[i] = make_short4((short)v.x, (short)v.y, (short)v.z, (short)v.w);
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) out_tail[threadIdx.x] = (int16_t)in_tail[threadIdx.x];
}

// graph_alltransitions: set of (label[t-1], label[t]) pairs
template <typename PT>
__global__ void __launch_bounds__(256)
all_transitions_kernel(const PT* __restrict__ hop_x, int64_t F, int64_t N, int n_labels, uint32_t* __restrict__ bitmap) {
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

// Tensor map of the brick-index array [F][N] (rows ld elements apart) with a box of 32 atoms x 8 frames; the driver
// entry point is looked up once (no link against libcuda).  false: not available, the register-staged path runs.
static bool encode_cells_map(CUtensorMap* map, const uint32_t* cells, int64_t F, int64_t N, int64_t ld) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = []() -> encode_fn {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<encode_fn>(fn);
  }();
  if (!encode) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)F};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, 8};
  const cuuint32_t estr[2] = {1, 1};
  static const int promo = getenv("HOPB_K3_TMA_PROMO") ? atoi(getenv("HOPB_K3_TMA_PROMO")) : 128;   // tuning hook
  const CUtensorMapL2promotion l2 = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                    : promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(cells), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int launch_hoptraj(hopb_handle* h, void* stream, int src, const hopb_grid* g, const float* xyz,
                          const float* labels_in, const uint32_t* cells, int64_t F, int64_t N, int64_t ld, int64_t frame0,
                          const int16_t* map16, const uint32_t* coarse, const uint8_t* fine, int fine_value,
                          const uint32_t* cell2, void* hop_x, void* hop_y, int n_chunks, int64_t chunk_len,
                          int32_t* summaries, hopb_event* events, uint64_t* n_events, uint64_t cap, bool planes16 = false) {
  HOPB_REQUIRE(h, F > 0 && N > 0 && summaries && n_events, "bad argument");
  HOPB_REQUIRE(h, !planes16 || (hop_x && hop_y && src != 1), "int16 planes: both hop_x and hop_y are needed");
  HOPB_REQUIRE(h, ld >= N && ld < ((int64_t)1 << 31), "row stride must be >= the number of atoms");
  HOPB_REQUIRE(h, n_chunks >= 1 && chunk_len >= 1 && (int64_t)n_chunks * chunk_len >= F &&
                      (int64_t)(n_chunks - 1) * chunk_len < F, "chunks must tile the slab exactly");
  HOPB_REQUIRE(h, n_chunks <= 65535, "too many chunks");
  HOPB_REQUIRE(h, frame0 >= 0 && frame0 + F < ((int64_t)1 << 31) && N < ((int64_t)1 << 31), "frame/atom index overflow");
  HOPB_REQUIRE(h, cap == 0 || events, "events buffer missing");
  EventSink sink{events, (unsigned long long*)n_events, (unsigned long long)cap};
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t items = ((N + 31) / 32) * n_chunks;  // (warp of atoms, chunk)
  if (src == 1) {
    HOPB_REQUIRE(h, labels_in, "null argument");
    HopbGridTab tab;
    memset(&tab, 0, sizeof(tab));
    const unsigned grid = (unsigned)((items + kHopThreads / 32 - 1) / (kHopThreads / 32));
    CUtensorMap no_map;
    memset(&no_map, 0, sizeof(no_map));
    hopb_note_launch();
    if (hop_x && hop_y)
      hoptraj_kernel<1, true, true, MAP_GATHER><<<grid, kHopThreads, 0, s>>>(nullptr, labels_in, nullptr, F, N, ld, frame0, tab,
          nullptr, nullptr, nullptr, nullptr, 0, 0, 0u, (float*)hop_x, (float*)hop_y, n_chunks, chunk_len, summaries, sink, no_map);
    else
      hoptraj_kernel<1, true, false, MAP_GATHER><<<grid, kHopThreads, 0, s>>>(nullptr, labels_in, nullptr, F, N, ld, frame0, tab,
          nullptr, nullptr, nullptr, nullptr, 0, 0, 0u, (float*)hop_x, (float*)hop_y, n_chunks, chunk_len, summaries, sink, no_map);
    HOPB_LAUNCH_CHECK(h);
    return HOPB_OK;
  }
  HOPB_REQUIRE(h, g && map16 && (src == 0 ? xyz != nullptr : cells != nullptr), "null argument");
  HOPB_REQUIRE(h, fine_value >= -1 && fine_value <= 1, "fine_value must be 0, 1 or -1 (automatic)");
  const bool fast = g->tab.fast_ok != 0;
  const int64_t nb = (int64_t)((g->n[0] + 3) / 4) * ((g->n[1] + 3) / 4) * ((g->n[2] + 3) / 4);
  size_t smem = 0;
  if (src == 0) smem += (size_t)g->tab.table2_len * sizeof(HopbPair) + (fast ? 0 : (size_t)g->tab.table_len * sizeof(float));
  // Block map: the fine one when it fits beside the tables (one copy per CTA, so big maps mean big
  // CTAs: an SM holds 1024 threads of this kernel either way), else the coarse one, else none.
  const size_t kSmemSm = 220 * 1024;
  int mode = MAP_GATHER, map_words = 0, threads = kHopThreads;
  const uint32_t* blockmap = nullptr;
  const size_t fine_bytes = (size_t)((nb * 8 + 31) / 32) * 4, coarse_bytes = (size_t)((nb + 15) / 16) * 4;
  if (fine && smem + fine_bytes <= kSmemSm) {
    mode = MAP_FINE; map_words = (int)(fine_bytes / 4); blockmap = reinterpret_cast<const uint32_t*>(fine);
    smem += fine_bytes;
  } else if (coarse && smem + coarse_bytes <= kSmemSm) {
    mode = MAP_COARSE; map_words = (int)(coarse_bytes / 4); blockmap = coarse;
    smem += coarse_bytes;
  } else if (cell2) {
    mode = MAP_CELL2; blockmap = cell2;
  }
  const uint32_t fine_auto_min = (uint32_t)(((int64_t)g->n[0] * g->n[1] * g->n[2] + 7) / 8);
  // Bulk-copy staging of the brick indices (see kTmaStages): rows must be 16-byte multiples apart at a 16-byte aligned
  // address, both planes are written, and the staging area (2 KB + 16 B per warp) has to fit beside the block map.
  const char* tma_env = getenv("HOPB_K3_TMA");      // HOPB_K3_TMA=0: register-staged input (tests compare the two)
  const bool tma_off = tma_env && atoi(tma_env) == 0;
  bool tma = !tma_off && src == 2 && hop_x && hop_y && ld == N && (ld & 3) == 0 && ((uintptr_t)cells & 15) == 0 && F >= 8;
  const size_t kStatic = 32 * 4 + 32 * kEvQueue * sizeof(hopb_event) + 1024;     // event queues + the driver's reservation
  auto pick_threads = [&](bool with_tma, size_t& smem_out) -> int {
    for (int t = 256; t <= 1024; t *= 2) {
      const size_t stage = with_tma ? 128 + (size_t)(t / 32) * (kTmaStages * 8 * 128 + kTmaStages * 8) : 0;
      const size_t per_cta = smem + stage + kStatic;
      if (per_cta * (1024 / t) <= 227 * 1024 || (t == 1024 && !with_tma)) { smem_out = smem + stage; return t; }
    }
    return 0;
  };
  size_t smem_launch = smem;
  CUtensorMap cells_map;
  memset(&cells_map, 0, sizeof(cells_map));
  if (tma) tma = encode_cells_map(&cells_map, cells, F, N, ld);
  threads = tma ? pick_threads(true, smem_launch) : 0;
  if (threads == 0) { tma = false; threads = pick_threads(false, smem_launch); }
  smem = smem_launch;
  const unsigned grid = (unsigned)((items + threads / 32 - 1) / (threads / 32));
#define HOPB_HT_LAUNCH_T(SRC, FAST, XY, MAP, ST, PT, TM)                                                             \
  do {                                                                                                               \
    if (smem > 48 * 1024)                                                                                            \
      HOPB_CUDA(h, cudaFuncSetAttribute(hoptraj_kernel<SRC, FAST, XY, MAP, ST, 8, PT, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    hopb_note_launch();                                                                                              \
    hoptraj_kernel<SRC, FAST, XY, MAP, ST, 8, PT, TM><<<grid, threads, smem, s>>>(xyz, nullptr, cells, F, N, ld, frame0, g->tab, g->d_table, \
        g->d_table2, map16, blockmap, map_words, fine_value, fine_auto_min, (PT*)hop_x, (PT*)hop_y, n_chunks, chunk_len, \
        summaries, sink, cells_map);                                                                                 \
  } while (0)
#define HOPB_HT_LAUNCH_S(SRC, FAST, XY, MAP, ST, PT)                                                                 \
  do {                                                                                                               \
    if (SRC == 2 && XY && !ST && tma) HOPB_HT_LAUNCH_T(SRC, FAST, XY, MAP, ST, PT, (SRC == 2 && XY && !ST));         \
    else HOPB_HT_LAUNCH_T(SRC, FAST, XY, MAP, ST, PT, false);                                                        \
  } while (0)
#define HOPB_HT_LAUNCH_M(SRC, FAST, XY, MAP, PT)                                                                     \
  do {                                                                                                               \
    if (ld != N) HOPB_HT_LAUNCH_S(SRC, FAST, XY, MAP, true, PT); else HOPB_HT_LAUNCH_S(SRC, FAST, XY, MAP, false, PT); \
  } while (0)
#define HOPB_HT_LAUNCH_XY(SRC, FAST, XY, PT)                                                                         \
  do {                                                                                                               \
    if (mode == MAP_FINE) HOPB_HT_LAUNCH_M(SRC, FAST, XY, MAP_FINE, PT);                                             \
    else if (mode == MAP_COARSE) HOPB_HT_LAUNCH_M(SRC, FAST, XY, MAP_COARSE, PT);                                    \
    else if (mode == MAP_CELL2) HOPB_HT_LAUNCH_M(SRC, FAST, XY, MAP_CELL2, PT);                                      \
    else HOPB_HT_LAUNCH_M(SRC, FAST, XY, MAP_GATHER, PT);                                                            \
  } while (0)
#define HOPB_HT_LAUNCH(SRC, FAST)                                                                                    \
  do {                                                                                                               \
    if (planes16) HOPB_HT_LAUNCH_XY(SRC, FAST, true, int16_t);                                                       \
    else if (hop_x && hop_y) HOPB_HT_LAUNCH_XY(SRC, FAST, true, float);                                              \
    else HOPB_HT_LAUNCH_XY(SRC, FAST, false, float);                                                                 \
  } while (0)
  if (src == 0) { if (fast) HOPB_HT_LAUNCH(0, true); else HOPB_HT_LAUNCH(0, false); }
  else HOPB_HT_LAUNCH(2, true);
#undef HOPB_HT_LAUNCH_T
#undef HOPB_HT_LAUNCH
#undef HOPB_HT_LAUNCH_XY
#undef HOPB_HT_LAUNCH_M
#undef HOPB_HT_LAUNCH_S
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hoptraj_strided(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz, const uint32_t* cells, int64_t F,
                         int64_t N, int64_t ld, int64_t frame0, const int16_t* map16, const uint32_t* coarse, const uint8_t* fine,
                         int fine_value, const uint32_t* cell2, float* hop_x, float* hop_y, int n_chunks, int64_t chunk_len,
                         int32_t* summaries, hopb_event* events, uint64_t* n_events, uint64_t cap) {
  return launch_hoptraj(h, stream, cells ? 2 : 0, g, xyz, nullptr, cells, F, N, ld, frame0, map16, coarse, fine, fine_value,
                        cell2, hop_x, hop_y, n_chunks, chunk_len, summaries, events, n_events, cap);
}

int hopb_hoptraj_i16(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz, const uint32_t* cells, int64_t F,
                     int64_t N, int64_t ld, int64_t frame0, const int16_t* map16, const uint32_t* coarse, const uint8_t* fine,
                     int fine_value, const uint32_t* cell2, int16_t* hop_x, int16_t* hop_y, int n_chunks, int64_t chunk_len,
                     int32_t* summaries, hopb_event* events, uint64_t* n_events, uint64_t cap) {
  return launch_hoptraj(h, stream, cells ? 2 : 0, g, xyz, nullptr, cells, F, N, ld, frame0, map16, coarse, fine, fine_value,
                        cell2, hop_x, hop_y, n_chunks, chunk_len, summaries, events, n_events, cap, true);
}

int hopb_hoptraj(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz, const uint32_t* cells, int64_t F,
                 int64_t N, int64_t frame0, const int16_t* map16, const uint32_t* coarse, const uint8_t* fine,
                 int fine_value, const uint32_t* cell2, float* hop_x, float* hop_y, int n_chunks, int64_t chunk_len,
                 int32_t* summaries, hopb_event* events, uint64_t* n_events, uint64_t cap) {
  return hopb_hoptraj_strided(h, stream, g, xyz, cells, F, N, N, frame0, map16, coarse, fine, fine_value, cell2, hop_x, hop_y,
                              n_chunks, chunk_len, summaries, events, n_events, cap);
}

int hopb_hopscan_labels(hopb_handle* h, void* stream, const float* hop_x, int64_t F, int64_t N, int64_t frame0, int n_chunks,
                        int64_t chunk_len, int32_t* summaries, hopb_event* events, uint64_t* n_events, uint64_t cap) {
  return launch_hoptraj(h, stream, 1, nullptr, nullptr, hop_x, nullptr, F, N, N, frame0, nullptr, nullptr, nullptr, 0, nullptr,
                        (void*)nullptr, (void*)nullptr, n_chunks, chunk_len, summaries, events, n_events, cap);
}

int hopb_brick_map(hopb_handle* h, void* stream, int nx, int ny, int nz, const int16_t* map16, int16_t* brick,
                   uint32_t* coarse, uint8_t* fine, uint32_t* cell2) {
  HOPB_REQUIRE(h, map16 && brick && coarse && nx > 0 && ny > 0 && nz > 0, "bad argument");
  const int cnx = (nx + 3) / 4, cny = (ny + 3) / 4, cnz = (nz + 3) / 4;
  const int64_t nb = (int64_t)cnx * cny * cnz;
  HOPB_REQUIRE(h, nb * 64 < ((int64_t)1 << 32) - 1, "grid too large for brick indices");
  const int64_t nb_pad = (nb + 3) / 4 * 4;
  if (fine) HOPB_CUDA(h, cudaMemsetAsync(fine + 2 * nb_pad, 0, 16, (cudaStream_t)stream));
  HOPB_K(brick_map_kernel)<<<hopb_grid_for(nb, 256, h->num_sms, 8), 256, 0, (cudaStream_t)stream>>>(
      map16, nx, ny, nz, cnx, cny, cnz, brick, coarse, fine, nb_pad, reinterpret_cast<uint4*>(cell2));
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hop_combine(hopb_handle* h, void* stream, int64_t N, int n_chunks, int64_t chunk_len, int64_t F,
                     const int32_t* summaries, int32_t* slab_summary) {
  HOPB_REQUIRE(h, N > 0 && n_chunks >= 1 && summaries && slab_summary, "bad argument");
  HOPB_K(hop_combine_kernel)<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(N, n_chunks, chunk_len, F, summaries, slab_summary);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hop_advance(hopb_handle* h, void* stream, int64_t N, int n_summaries, const int32_t* summaries, int init,
                     int32_t* state) {
  HOPB_REQUIRE(h, N > 0 && n_summaries >= 0 && state && (n_summaries == 0 || summaries), "bad argument");
  HOPB_K(hop_advance_kernel)<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(N, n_summaries, summaries, init, state);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hop_fixup(hopb_handle* h, void* stream, int64_t N, int n_chunks, int64_t chunk_len, int64_t F,
                   const int32_t* summaries, int32_t* state, float* hop_y, hopb_event* events, uint64_t* n_events,
                   uint64_t cap) {
  return hopb_hop_fixup_strided(h, stream, N, N, n_chunks, chunk_len, F, summaries, state, hop_y, events, n_events, cap);
}

int hopb_hop_fixup_strided(hopb_handle* h, void* stream, int64_t N, int64_t ld, int n_chunks, int64_t chunk_len, int64_t F,
                           const int32_t* summaries, int32_t* state, float* hop_y, hopb_event* events, uint64_t* n_events,
                           uint64_t cap) {
  HOPB_REQUIRE(h, N > 0 && ld >= N && n_chunks >= 1 && summaries && state && n_events, "bad argument");
  HOPB_REQUIRE(h, cap == 0 || events, "events buffer missing");
  (void)F;
  EventSink sink{events, (unsigned long long*)n_events, (unsigned long long)cap};
  HOPB_K(hop_fixup_kernel<float>)<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(N, ld, n_chunks, chunk_len, F, summaries, state,
                                                                             hop_y, sink);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_hop_fixup_i16(hopb_handle* h, void* stream, int64_t N, int64_t ld, int n_chunks, int64_t chunk_len, int64_t F,
                       const int32_t* summaries, int32_t* state, int16_t* hop_y, hopb_event* events, uint64_t* n_events,
                       uint64_t cap) {
  HOPB_REQUIRE(h, N > 0 && ld >= N && n_chunks >= 1 && summaries && state && n_events, "bad argument");
  HOPB_REQUIRE(h, cap == 0 || events, "events buffer missing");
  (void)F;
  EventSink sink{events, (unsigned long long*)n_events, (unsigned long long)cap};
  HOPB_K(hop_fixup_kernel<int16_t>)<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(N, ld, n_chunks, chunk_len, F, summaries,
                                                                                                 state, hop_y, sink);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_labels_to_int16(hopb_handle* h, void* stream, const float* labels, int64_t n, int16_t* out) {
  HOPB_REQUIRE(h, labels && out && n >= 0, "bad argument");
  HOPB_REQUIRE(h, ((uintptr_t)labels & 15) == 0 && ((uintptr_t)out & 7) == 0, "labels must be 16-byte, out 8-byte aligned");
  if (n == 0) return HOPB_OK;
  const int64_t n4 = n / 4;
  HOPB_K(planes_to_int16_kernel)<<<hopb_grid_for(n4 > 0 ? n4 : 1, 256, h->num_sms, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(labels), n4, labels + 4 * n4, (int)(n - 4 * n4), reinterpret_cast<short4*>(out), out + 4 * n4);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_all_transitions(hopb_handle* h, void* stream, const float* hop_x, int64_t F, int64_t N, int n_labels,
                         uint32_t* bitmap) {
  HOPB_REQUIRE(h, hop_x && bitmap && F >= 1 && N > 0 && n_labels >= 1, "bad argument");
  if (F < 2) return HOPB_OK;
  HOPB_K(all_transitions_kernel<float>)<<<hopb_grid_for((F - 1) * N, 256, h->num_sms, 8), 256, 0, (cudaStream_t)stream>>>(hop_x, F, N, n_labels, bitmap);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

int hopb_all_transitions_i16(hopb_handle* h, void* stream, const int16_t* hop_x, int64_t F, int64_t N, int n_labels,
                             uint32_t* bitmap) {
  HOPB_REQUIRE(h, hop_x && bitmap && F >= 1 && N > 0 && n_labels >= 1, "bad argument");
  if (F < 2) return HOPB_OK;
  HOPB_K(all_transitions_kernel<int16_t>)<<<hopb_grid_for((F - 1) * N, 256, h->num_sms, 8), 256, 0, (cudaStream_t)stream>>>(hop_x, F, N, n_labels, bitmap);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

}  // extern "C"
