// Per-atom transition state machine of TransportNetwork._analyze_hops (hop/graph.py:190-285,
// SURVEY.md Appendix A.4), written as an algebra of frame-range *summaries* so that a trajectory
// can be cut into frame chunks (within a GPU) and frame slabs (across GPUs) and stitched back
// together exactly.  Host + device: the same code is compiled into the CUDA kernels and into the
// CPU-side unit tests (tests/emul), which replay it against the oracle.
//
// Reference machine, per atom (s = site label of this frame, outliers replaced by slast):
//     if s != s0:
//         if slast == s0: t1 = t                     # left s0 at t
//         if s != 0:
//             if s0 != 0: emit (s0 -> s): tau = t1 - t0, tbarrier = t - t1, frame = t
//             s0 = s; t0 = t
//     slast = s
// Invariant: slast is either s0 ("on") or 0 ("off"), so the state is (s0, t0, t1, on).  Outlier
// frames are no-ops (s = slast).  The first frame of the trajectory sets s0 = x[0] *raw*; that is
// the same as running the machine from the virgin state (s0 = 0) except that a raw -1 must be
// treated as a visit to the pseudo-site -1 (hop/graph.py:190-208, the "(-1, N) edges" quirk).
//
// What a frame range does to an *unknown* incoming state only depends on that state until the
// second distinct site of the range is reached:
//   EMPTY  only outliers                      -> identity
//   PRE    only interstitial frames; z = first of them (sets t1 = z if the atom was "on")
//   SPEC   first site A1 at f1 (z before it, if any); no second site yet.  The incoming s0 decides
//          whether f1 is a transition (s0_in != A1: event E1, t0 = f1) or a return (t0 = t0_in).
//   FULL   second site A2 != A1 at f2: event E2 = (A1 -> A2) whose tau = t1_2 - t0 is still open
//          in t0; everything after f2 is independent of the incoming state and is emitted directly.
// Summaries are closed under concatenation (hopb_combine), and applying one to a known state
// (hopb_apply) yields at most the two open events E1, E2.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HOPB_HD __host__ __device__ __forceinline__
#else
#define HOPB_HD inline
#endif

enum { HOPB_EMPTY = 0, HOPB_PRE = 1, HOPB_SPEC = 2, HOPB_FULL = 3 };

#define HOPB_SUMMARY_FIELDS 11  // ints per (chunk, atom) in the packed [chunk][field][atom] layout
#define HOPB_STATE_FIELDS 4     // s0, t0, t1, on

struct HopbSummary {
  int32_t kind;  // HOPB_*
  int32_t on;    // slast == s0 at the end of the range (SPEC / FULL)
  int32_t z;     // first interstitial frame before f1, or -1
  int32_t A1, f1;
  int32_t A2, f2, t1_2;
  int32_t s0, t0, t1;  // end state (FULL: absolute; SPEC: t1 only, valid when !on)
  int32_t ypre;        // leading frames of the range whose orbit site is inherited
};

struct HopbState {
  int32_t s0, t0, t1, on;
};

HOPB_HD HopbSummary hopb_summary_empty() {
  HopbSummary T;
  T.kind = HOPB_EMPTY; T.on = 1; T.z = -1;
  T.A1 = 0; T.f1 = 0; T.A2 = 0; T.f2 = 0; T.t1_2 = 0;
  T.s0 = 0; T.t0 = 0; T.t1 = 0; T.ypre = 0;
  return T;
}

HOPB_HD HopbState hopb_state_virgin() {
  HopbState S; S.s0 = 0; S.t0 = 0; S.t1 = 0; S.on = 1;
  return S;
}

// Feed one frame (global frame number t, raw label) into the running summary of a chunk.
// `emit(s0, s, frame, tau, tbarrier)` is called for every fully determined transition.
template <class Emit>
HOPB_HD void hopb_push_frame(HopbSummary& T, int label, int t, Emit&& emit) {
  if (label == -1 && t != 0) return;  // outlier: s = slast, nothing changes
  switch (T.kind) {
    case HOPB_EMPTY:
    case HOPB_PRE:
      if (label == 0) {
        if (T.kind == HOPB_EMPTY) { T.kind = HOPB_PRE; T.z = t; }
      } else {
        T.kind = HOPB_SPEC; T.A1 = label; T.f1 = t; T.on = 1;
      }
      break;
    case HOPB_SPEC:
      if (label == T.A1) {
        T.on = 1;
      } else if (label == 0) {
        if (T.on) { T.t1 = t; T.on = 0; }
      } else {
        T.t1_2 = T.on ? t : T.t1;
        T.A2 = label; T.f2 = t;
        T.kind = HOPB_FULL; T.s0 = label; T.t0 = t; T.on = 1;
      }
      break;
    default:  // HOPB_FULL: the plain machine on (s0, t0, t1, on)
      if (label == T.s0) {
        T.on = 1;
      } else if (label == 0) {
        if (T.on) { T.t1 = t; T.on = 0; }
      } else {
        const int t1 = T.on ? t : T.t1;
        emit(T.s0, label, t, t1 - T.t0, t - t1);
        T.s0 = label; T.t0 = t; T.on = 1;
      }
      break;
  }
}

// Apply a summary to a known state; emits the (at most two) events the summary left open.
template <class Emit>
HOPB_HD void hopb_apply(HopbState& S, const HopbSummary& T, Emit&& emit) {
  if (T.kind == HOPB_EMPTY) return;
  if (T.z >= 0 && S.on && S.s0 != 0) { S.t1 = T.z; S.on = 0; }
  if (T.kind == HOPB_PRE) return;
  if (T.A1 != S.s0) {
    if (S.on) S.t1 = T.f1;
    if (S.s0 != 0) emit(S.s0, T.A1, T.f1, S.t1 - S.t0, T.f1 - S.t1);
    S.s0 = T.A1; S.t0 = T.f1;
  }
  S.on = 1;
  if (T.kind == HOPB_SPEC) {
    S.on = T.on;
    if (!T.on) S.t1 = T.t1;
    return;
  }
  emit(T.A1, T.A2, T.f2, T.t1_2 - S.t0, T.f2 - T.t1_2);
  S.s0 = T.s0; S.t0 = T.t0; S.t1 = T.t1; S.on = T.on;
}

// Concatenate two summaries (T earlier, U later) into the summary of the joint range.  Events that
// become determined by the concatenation are NOT emitted here: hopb_apply re-derives them chunk by
// chunk once the incoming state is known (see the fix-up kernel), so this is only used to build
// the one-summary-per-slab message that ranks exchange.
HOPB_HD HopbSummary hopb_combine(const HopbSummary& T, const HopbSummary& U, int T_len) {
  // orbit prefix: T.ypre == T_len means T never saw a real site (label > 0)
  const int ypre = (T.ypre >= T_len) ? T_len + U.ypre : T.ypre;
  HopbSummary R;
  if (U.kind == HOPB_EMPTY) {
    R = T;
  } else if (T.kind == HOPB_EMPTY || T.kind == HOPB_PRE) {
    R = U;
    if (T.kind == HOPB_PRE) R.z = T.z;
  } else if (T.kind == HOPB_SPEC) {
    R = T;  // keeps z, A1, f1 of the earlier range
    int on = T.on, t1 = T.t1;
    if (U.z >= 0 && on) { t1 = U.z; on = 0; }
    if (U.kind == HOPB_PRE) {
      R.on = on; R.t1 = t1;
    } else if (U.A1 == T.A1) {
      if (U.kind == HOPB_SPEC) {
        R.on = U.on; R.t1 = U.on ? t1 : U.t1;
      } else {
        R.kind = HOPB_FULL; R.A2 = U.A2; R.f2 = U.f2; R.t1_2 = U.t1_2;
        R.s0 = U.s0; R.t0 = U.t0; R.t1 = U.t1; R.on = U.on;
      }
    } else {
      R.kind = HOPB_FULL; R.A2 = U.A1; R.f2 = U.f1; R.t1_2 = on ? U.f1 : t1;
      if (U.kind == HOPB_SPEC) { R.s0 = U.A1; R.t0 = U.f1; R.t1 = U.t1; R.on = U.on; }
      else { R.s0 = U.s0; R.t0 = U.t0; R.t1 = U.t1; R.on = U.on; }
    }
  } else {
    // T is FULL: its end state is absolute
    R = T;
    HopbState S; S.s0 = T.s0; S.t0 = T.t0; S.t1 = T.t1; S.on = T.on;
    hopb_apply(S, U, [](int, int, int, int, int) {});
    R.s0 = S.s0; R.t0 = S.t0; R.t1 = S.t1; R.on = S.on;
  }
  R.ypre = ypre;
  return R;
}

// packed <-> struct ([field][atom] with a stride between fields)
HOPB_HD void hopb_summary_store(int32_t* base, int64_t stride, const HopbSummary& T) {
  base[0 * stride] = T.kind | (T.on << 2);
  base[1 * stride] = T.z;
  base[2 * stride] = T.A1;
  base[3 * stride] = T.f1;
  base[4 * stride] = T.A2;
  base[5 * stride] = T.f2;
  base[6 * stride] = T.t1_2;
  base[7 * stride] = T.s0;
  base[8 * stride] = T.t0;
  base[9 * stride] = T.t1;
  base[10 * stride] = T.ypre;
}

HOPB_HD HopbSummary hopb_summary_load(const int32_t* base, int64_t stride) {
  HopbSummary T;
  const int k = base[0 * stride];
  T.kind = k & 3; T.on = (k >> 2) & 1;
  T.z = base[1 * stride];
  T.A1 = base[2 * stride];
  T.f1 = base[3 * stride];
  T.A2 = base[4 * stride];
  T.f2 = base[5 * stride];
  T.t1_2 = base[6 * stride];
  T.s0 = base[7 * stride];
  T.t0 = base[8 * stride];
  T.t1 = base[9 * stride];
  T.ypre = base[10 * stride];
  return T;
}
