// CPU replay of the host/device headers used by the CUDA kernels -- TEST INFRASTRUCTURE ONLY.
// Built by tests/emul/build.py with g++ and loaded by the `-m "not gpu"` tests; hop_b200 never
// loads it.  It lets the chunk/slab stitching algebra (hopb_machine.h) and the float32 binning
// (hopb_binning.h) be checked against the oracle without a GPU.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../hop_b200/csrc/hopb_binning.h"
#include "../../hop_b200/csrc/hopb_gridtab_host.h"
#include "../../hop_b200/csrc/hopb_machine.h"

struct Ev { int32_t frame, iatom, s0, s, tau, tbar; };

extern "C" {

// labels [F][N] int32 (global frames frame0 .. frame0+F-1); slabs of `slab_len` frames are cut into
// chunks of `chunk_len` frames, exactly like ranks x chunks on the GPUs.  Events are written
// unsorted to ev_out (cap entries); returns the number of events, final state per atom in
// state_out[N][4] and orbit column in orbit_out[F][N].
int64_t emul_scan(const int32_t* labels, int64_t F, int64_t N, int64_t slab_len, int64_t chunk_len,
                  Ev* ev_out, int64_t cap, int32_t* state_out, int32_t* orbit_out) {
  std::vector<Ev> ev;
  const int64_t nslabs = (F + slab_len - 1) / slab_len;
  for (int64_t a = 0; a < N; ++a) {
    auto emit = [&](int s0, int s, int frame, int tau, int tbar) {
      ev.push_back(Ev{frame, (int32_t)a, s0, s, tau, tbar});
    };
    // pass 1: per chunk summaries (events of FULL phases emitted directly), per slab combine
    std::vector<std::vector<HopbSummary>> chunks(nslabs);
    std::vector<HopbSummary> slabs(nslabs);
    for (int64_t r = 0; r < nslabs; ++r) {
      const int64_t s_begin = r * slab_len, s_end = std::min(F, s_begin + slab_len);
      HopbSummary slab = hopb_summary_empty();
      int64_t done = 0;
      for (int64_t c0 = s_begin; c0 < s_end; c0 += chunk_len) {
        const int64_t c1 = std::min(s_end, c0 + chunk_len);
        HopbSummary T = hopb_summary_empty();
        int ypre = 0, y = 0; bool yknown = false;
        for (int64_t f = c0; f < c1; ++f) {
          const int lab = labels[f * N + a];
          hopb_push_frame(T, lab, (int)f, emit);
          if (lab > 0) { y = lab; yknown = true; }
          if (!yknown) ++ypre;
          orbit_out[f * N + a] = yknown ? y : 0;
        }
        T.ypre = ypre;
        // round trip through the packed layout
        int32_t packed[HOPB_SUMMARY_FIELDS];
        hopb_summary_store(packed, 1, T);
        T = hopb_summary_load(packed, 1);
        chunks[r].push_back(T);
        slab = hopb_combine(slab, T, (int)done);
        done += c1 - c0;
      }
      slabs[r] = slab;
    }
    // pass 2: per slab, incoming state from the earlier slabs' summaries, then fix-up per chunk
    for (int64_t r = 0; r < nslabs; ++r) {
      HopbState S = hopb_state_virgin();
      for (int64_t q = 0; q < r; ++q) hopb_apply(S, slabs[q], [](int, int, int, int, int) {});
      const int64_t s_begin = r * slab_len;
      for (size_t c = 0; c < chunks[r].size(); ++c) {
        const HopbSummary& T = chunks[r][c];
        const int yin = S.s0 > 0 ? S.s0 : 0;
        const int64_t c0 = s_begin + (int64_t)c * chunk_len;
        if (yin != 0) for (int k = 0; k < T.ypre; ++k) orbit_out[(c0 + k) * N + a] = yin;
        hopb_apply(S, T, emit);
      }
      if (r == nslabs - 1) { state_out[a * 4 + 0] = S.s0; state_out[a * 4 + 1] = S.t0; state_out[a * 4 + 2] = S.t1; state_out[a * 4 + 3] = S.on; }
    }
    // cross-check: slab-level apply must reach the same final state as chunk-level apply
    HopbState S2 = hopb_state_virgin();
    for (int64_t q = 0; q < nslabs; ++q) hopb_apply(S2, slabs[q], [](int, int, int, int, int) {});
    if (S2.s0 != state_out[a * 4 + 0] || S2.t0 != state_out[a * 4 + 1] || S2.on != state_out[a * 4 + 3] ||
        (!S2.on && S2.t1 != state_out[a * 4 + 2]))
      return -2;
  }
  if ((int64_t)ev.size() > cap) return -1;
  std::memcpy(ev_out, ev.data(), ev.size() * sizeof(Ev));
  return (int64_t)ev.size();
}

// Binning: mode 0 = histogram bin (1..n counted), mode 1 = hoptraj bin ([-1, n+1]); modes 2 / 3 are
// the same through the branch-free pair-table search the kernels use when fast_ok is set (returned);
// mode 4 = histogram bin through the table-free "sure" bin with the pair-table search as fallback (what
// hist_kernel does), mode 5 = 1 where the sure bin applies, else 0.
int emul_bin(const double* ex, int nx, const double* ey, int ny, const double* ez, int nz,
             const int* decimals, const float* xyz, int64_t npts, int mode, int32_t* out /*[npts][3]*/,
             float* top_round_out /*[3]*/) {
  const double* e[3] = {ex, ey, ez};
  const int n[3] = {nx, ny, nz};
  HopbGridTab tab;
  std::vector<float> table;
  std::vector<HopbPair> table2;
  if (hopb_build_gridtab(e, n, decimals, &tab, &table, &table2) != 0) return -1;
  for (int d = 0; d < 3; ++d) top_round_out[d] = tab.ax[d].top_round;
  for (int64_t i = 0; i < npts; ++i)
    for (int d = 0; d < 3; ++d) {
      const float x = xyz[i * 3 + d];
      const float* eup = table.data() + tab.ax[d].off;
      const HopbPair* pr = table2.data() + tab.ax[d].off2;
      int b;
      if (mode == 0) b = hopb_bin_hist(x, tab.ax[d], eup);
      else if (mode == 1) b = hopb_bin_hop(x, tab.ax[d], eup);
      else if (mode == 2) b = hopb_bin_hist_fast(x, tab.ax[d], pr);
      else if (mode == 3) b = hopb_bin_hop_fast(x, tab.ax[d], pr);
      else {
        const HopbAxis& a = tab.ax[d];
        const int s = tab.sure_eps > 0.0f ? hopb_bin_sure(x, a.lo, a.inv, (float)a.n, tab.sure_eps) : 0;
        if (mode == 4) b = s ? s : hopb_bin_hist_fast(x, a, pr);
        else b = s ? 1 : 0;
      }
      out[i * 3 + d] = b;
    }
  return tab.fast_ok;
}

}  // extern "C"

// ---- kernel-by-kernel CPU mirrors of csrc/hoptraj.cu, for the multi-process (gloo) tests --------------
extern "C" {

// mirrors hoptraj_kernel<1>: labels [F][N] of one slab -> summaries [n_chunks][11][N] + events
int64_t emul_slab_scan(const int32_t* labels, int64_t F, int64_t N, int64_t frame0, int n_chunks, int64_t chunk_len,
                       int32_t* summaries, Ev* ev_out, int64_t cap) {
  int64_t n_ev = 0;
  for (int c = 0; c < n_chunks; ++c)
    for (int64_t a = 0; a < N; ++a) {
      auto emit = [&](int s0, int s, int frame, int tau, int tbar) {
        if (n_ev < cap) ev_out[n_ev] = Ev{frame, (int32_t)a, s0, s, tau, tbar};
        ++n_ev;
      };
      HopbSummary T = hopb_summary_empty();
      int ypre = 0; bool yknown = false;
      const int64_t f0 = (int64_t)c * chunk_len, f1 = std::min(F, f0 + chunk_len);
      for (int64_t f = f0; f < f1; ++f) {
        const int lab = labels[f * N + a];
        hopb_push_frame(T, lab, (int)(frame0 + f), emit);
        if (lab > 0) yknown = true;
        if (!yknown) ++ypre;
      }
      T.ypre = ypre;
      hopb_summary_store(summaries + ((int64_t)c * HOPB_SUMMARY_FIELDS) * N + a, N, T);
    }
  return n_ev;
}

// mirrors hop_combine_kernel
void emul_combine(const int32_t* summaries, int64_t N, int n_chunks, int64_t chunk_len, int64_t F, int32_t* slab) {
  for (int64_t a = 0; a < N; ++a) {
    HopbSummary R = hopb_summary_empty();
    int64_t done = 0;
    for (int c = 0; c < n_chunks; ++c) {
      const HopbSummary T = hopb_summary_load(summaries + ((int64_t)c * HOPB_SUMMARY_FIELDS) * N + a, N);
      R = hopb_combine(R, T, (int)done);
      done += std::min(chunk_len, F - (int64_t)c * chunk_len);
    }
    hopb_summary_store(slab + a, N, R);
  }
}

// mirrors hop_advance_kernel: state [4][N]
void emul_advance(const int32_t* summaries, int n, int64_t N, int init, int32_t* state) {
  for (int64_t a = 0; a < N; ++a) {
    HopbState S = hopb_state_virgin();
    if (!init) { S.s0 = state[a]; S.t0 = state[N + a]; S.t1 = state[2 * N + a]; S.on = state[3 * N + a]; }
    for (int c = 0; c < n; ++c)
      hopb_apply(S, hopb_summary_load(summaries + ((int64_t)c * HOPB_SUMMARY_FIELDS) * N + a, N), [](int, int, int, int, int) {});
    state[a] = S.s0; state[N + a] = S.t0; state[2 * N + a] = S.t1; state[3 * N + a] = S.on;
  }
}

// mirrors hop_fixup_kernel (events only)
int64_t emul_fixup(const int32_t* summaries, int n_chunks, int64_t N, int32_t* state, Ev* ev_out, int64_t cap) {
  int64_t n_ev = 0;
  for (int64_t a = 0; a < N; ++a) {
    auto emit = [&](int s0, int s, int frame, int tau, int tbar) {
      if (n_ev < cap) ev_out[n_ev] = Ev{frame, (int32_t)a, s0, s, tau, tbar};
      ++n_ev;
    };
    HopbState S; S.s0 = state[a]; S.t0 = state[N + a]; S.t1 = state[2 * N + a]; S.on = state[3 * N + a];
    for (int c = 0; c < n_chunks; ++c)
      hopb_apply(S, hopb_summary_load(summaries + ((int64_t)c * HOPB_SUMMARY_FIELDS) * N + a, N), emit);
    state[a] = S.s0; state[N + a] = S.t0; state[2 * N + a] = S.t1; state[3 * N + a] = S.on;
  }
  return n_ev;
}

}  // extern "C"
