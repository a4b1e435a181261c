// Rate constants from waiting times on the GPU: HoppingGraph._rate (hop/graph.py:1677-1748)
// for every edge of a hopping graph at once.
//
// Per edge the reference (a) evaluates the survival function S(t) = <Theta(tau_i - t)>, Theta(0) = 1/2, on the mesh
// t = 0, 1, 2, ... < max(tau) + 2 as a sum of len(tau) step functions per mesh point (hop/graph.py:2433-2518), and
// (b) fits a * exp(-k1 t) + (1 - a) * exp(-k2 t) -- or exp(-k t) when there are at most 10 waiting times or the
// double exponential comes out degenerate -- with scipy.optimize.leastsq, i.e. MINPACK's lmdif (fit_func,
// hop/graph.py:2354-2419).  "Hours" for a big graph (hop/graph.py:587-589).
//
// Here: (a) one histogram + suffix sum per edge (S on an integer mesh only depends on how many tau lie above /
// exactly on each mesh point: sums of 0, 1/2 and 1, exact); (b) one warp per edge runs lmdif itself -- the
// algorithm of MINPACK restated: forward-difference Jacobian (fdjac2), Householder QR with column pivoting
// (qrfac), the Levenberg-Marquardt parameter from lmpar / qrsolv, the same scaling, step-bound updates,
// tolerances (ftol = xtol = 1.49012e-8, gtol = 0, factor = 100, maxfev = 200 (n + 1), epsfcn = machine epsilon:
// scipy's defaults) and start values ([0.5, 0.1, 1e-4] and [1.0]).  The m-vectors live in global memory and are
// processed by the 32 lanes; the n <= 3 sized algebra is done redundantly by every lane.  Differences to MINPACK
// are of rounding order only (plain instead of range-scaled sums of squares, exp of the CUDA math library).
#include "hopb_common.cuh"

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr double kEpsMch = 2.220446049250313e-16;
constexpr double kDwarf = 2.2250738585072014e-308;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);   // same value, bit for bit, on every lane
  return v;
}

// ---- survival function -------------------------------------------------------------------------------------
// hist[k] = #tau with ceil(tau) == k, eq[k] = #tau == k; then #tau > j = sum_{k > j} hist[k].
__global__ void __launch_bounds__(256)
survival_kernel(const double* __restrict__ tau, const int64_t* __restrict__ off_tau, const int64_t* __restrict__ off_x, int64_t n_edges,
                double* __restrict__ S, uint32_t* __restrict__ work /* 2 x total mesh points, zeroed */, int64_t total_x) {
  for (int64_t e = blockIdx.x; e < n_edges; e += gridDim.x) {
    const int64_t t0 = off_tau[e], t1 = off_tau[e + 1], x0 = off_x[e], m = off_x[e + 1] - off_x[e];
    uint32_t* hist = work + x0;
    uint32_t* eq = work + total_x + x0;
    for (int64_t i = t0 + threadIdx.x; i < t1; i += blockDim.x) {
      const double t = tau[i];
      const double c = ceil(t);
      if (c >= 0.0 && c < (double)m) atomicAdd(hist + (int64_t)c, 1u);
      if (c == t && c >= 0.0 && c < (double)m) atomicAdd(eq + (int64_t)c, 1u);
    }
    __syncthreads();
    // suffix sums by one warp (meshes are short compared with the events)
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const double n = (double)(t1 - t0);
      uint32_t carry = 0;     // #tau with ceil > the chunk being processed
      for (int64_t hi = m; hi > 0; hi -= 32) {
        const int64_t j = hi - 1 - lane;                  // lane 0 takes the top mesh point of the chunk
        uint32_t above = j >= 0 && j + 1 < m ? hist[j + 1] : 0u;
        // inclusive prefix over lanes = number of tau above mesh point j inside this chunk
        uint32_t incl = above;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += v; }
        if (j >= 0) S[x0 + j] = ((double)(carry + incl) + 0.5 * (double)eq[j]) / n;
        carry += __shfl_sync(kFullMask, incl, 31);
      }
    }
    __syncthreads();
  }
}

// ---- lmdif, one warp per problem --------------------------------------------------------------------------------
template <int NP>
struct Fit {
  const double* S;     // data y_j at x_j = j
  int64_t m;
  double* fvec;        // [m]
  double* wa4;         // [m]
  double* fjac;        // [NP][m], column-major
  int lane;
  unsigned noise;      // != 0: every exp() is nudged by +-1 ulp, pseudo-randomly (the "second opinion" of hopb_rate_fit)
  mutable unsigned calls;

  __device__ __forceinline__ double ex(double arg, unsigned salt) const {
    double v = exp(arg);
    if (noise) {
      unsigned hsh = (salt * 2654435761u) ^ (calls * 0x9E3779B9u) ^ noise;
      hsh ^= hsh >> 15; hsh *= 0x2c1b3c6du; hsh ^= hsh >> 12;
      v = __longlong_as_double(__double_as_longlong(v) + ((hsh & 1u) ? 1ll : -1ll));
    }
    return v;
  }

  __device__ __forceinline__ double model(const double* p, double x, unsigned j) const {
    if (NP == 1) return ex(-p[0] * x, j);
    return p[0] * ex(-p[1] * x, 2u * j) + (1.0 - p[0]) * ex(-p[2] * x, 2u * j + 1u);
  }
  __device__ __forceinline__ void eval(const double* p, double* out) const {
    ++calls;
    for (int64_t j = lane; j < m; j += 32) out[j] = model(p, (double)j, (unsigned)j) - S[j];
    __syncwarp();
  }
  __device__ __forceinline__ double enorm_m(const double* v, int64_t from = 0) const {
    double s = 0.0;
    for (int64_t j = from + lane; j < m; j += 32) s += v[j] * v[j];
    return sqrt(warp_sum(s));
  }
};

template <int NP>
__device__ __forceinline__ double enorm_n(const double* v) {
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < NP; ++j) s += v[j] * v[j];
  return sqrt(s);
}

// qrsolv (MINPACK): r is the NP x NP upper triangle (full square storage, the strict lower part is scratch = S^T)
template <int NP>
__device__ void qrsolv(double r[NP][NP], const int* ipvt, const double* diag, const double* qtb, double* x, double* sdiag) {
  double wa[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    for (int i = j; i < NP; ++i) r[i][j] = r[j][i];
    x[j] = r[j][j];
    wa[j] = qtb[j];
  }
  for (int j = 0; j < NP; ++j) {
    const int l = ipvt[j];
    if (diag[l] != 0.0) {
      for (int k = j; k < NP; ++k) sdiag[k] = 0.0;
      sdiag[j] = diag[l];
      double qtbpj = 0.0;
      for (int k = j; k < NP; ++k) {
        if (sdiag[k] == 0.0) continue;
        double c, s;
        if (fabs(r[k][k]) >= fabs(sdiag[k])) {
          const double t = sdiag[k] / r[k][k];
          c = 0.5 / sqrt(0.25 + 0.25 * t * t);
          s = c * t;
        } else {
          const double ct = r[k][k] / sdiag[k];
          s = 0.5 / sqrt(0.25 + 0.25 * ct * ct);
          c = s * ct;
        }
        r[k][k] = c * r[k][k] + s * sdiag[k];
        const double temp = c * wa[k] + s * qtbpj;
        qtbpj = -s * wa[k] + c * qtbpj;
        wa[k] = temp;
        for (int i = k + 1; i < NP; ++i) {
          const double t2 = c * r[i][k] + s * sdiag[i];
          sdiag[i] = -s * r[i][k] + c * sdiag[i];
          r[i][k] = t2;
        }
      }
    }
    sdiag[j] = r[j][j];
    r[j][j] = x[j];
  }
  int nsing = NP;
  for (int j = 0; j < NP; ++j) {
    if (sdiag[j] == 0.0 && nsing == NP) nsing = j;
    if (nsing < NP) wa[j] = 0.0;
  }
  for (int k = 0; k < nsing; ++k) {
    const int j = nsing - 1 - k;
    double sum = 0.0;
    for (int i = j + 1; i < nsing; ++i) sum += r[i][j] * wa[i];
    wa[j] = (wa[j] - sum) / sdiag[j];
  }
  for (int j = 0; j < NP; ++j) x[ipvt[j]] = wa[j];
}

// lmpar (MINPACK)
template <int NP>
__device__ void lmpar(double r[NP][NP], const int* ipvt, const double* diag, const double* qtb, double delta, double& par, double* x,
                      double* sdiag) {
  double wa1[NP], wa2[NP];
  int nsing = NP;
  for (int j = 0; j < NP; ++j) {
    wa1[j] = qtb[j];
    if (r[j][j] == 0.0 && nsing == NP) nsing = j;
    if (nsing < NP) wa1[j] = 0.0;
  }
  for (int k = 0; k < nsing; ++k) {
    const int j = nsing - 1 - k;
    wa1[j] = wa1[j] / r[j][j];
    const double temp = wa1[j];
    for (int i = 0; i < j; ++i) wa1[i] -= r[i][j] * temp;
  }
  for (int j = 0; j < NP; ++j) x[ipvt[j]] = wa1[j];
  int iter = 0;
  for (int j = 0; j < NP; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = enorm_n<NP>(wa2);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) { par = 0.0; return; }
  double parl = 0.0;
  if (nsing >= NP) {
    for (int j = 0; j < NP; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
    for (int j = 0; j < NP; ++j) {
      double sum = 0.0;
      for (int i = 0; i < j; ++i) sum += r[i][j] * wa1[i];
      wa1[j] = (wa1[j] - sum) / r[j][j];
    }
    const double temp = enorm_n<NP>(wa1);
    parl = ((fp / delta) / temp) / temp;
  }
  for (int j = 0; j < NP; ++j) {
    double sum = 0.0;
    for (int i = 0; i <= j; ++i) sum += r[i][j] * qtb[i];
    wa1[j] = sum / diag[ipvt[j]];
  }
  const double gnorm = enorm_n<NP>(wa1);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = kDwarf / fmin(delta, 0.1);
  par = fmax(par, parl);
  par = fmin(par, paru);
  if (par == 0.0) par = gnorm / dxnorm;
  while (true) {
    ++iter;
    if (par == 0.0) par = fmax(kDwarf, 0.001 * paru);
    double temp = sqrt(par);
    for (int j = 0; j < NP; ++j) wa1[j] = temp * diag[j];
    qrsolv<NP>(r, ipvt, wa1, qtb, x, sdiag);
    for (int j = 0; j < NP; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = enorm_n<NP>(wa2);
    temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    for (int j = 0; j < NP; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
    for (int j = 0; j < NP; ++j) {
      wa1[j] = wa1[j] / sdiag[j];
      const double t = wa1[j];
      for (int i = j + 1; i < NP; ++i) wa1[i] -= r[i][j] * t;
    }
    temp = enorm_n<NP>(wa1);
    const double parc = ((fp / delta) / temp) / temp;
    if (fp > 0.0) parl = fmax(parl, par);
    if (fp < 0.0) paru = fmin(paru, par);
    par = fmax(parl, par + parc);
  }
  if (iter == 0) par = 0.0;
}

// lmdif with scipy.optimize.leastsq's defaults.  x: start values in, solution out.  Returns MINPACK's info.
template <int NP>
__device__ int lmdif(const Fit<NP>& F, double* x, int* nfev_out) {
  const double ftol = 1.49012e-8, xtol = 1.49012e-8, gtol = 0.0, factor = 100.0;
  const int maxfev = 200 * (NP + 1);
  const double eps = sqrt(kEpsMch);     // epsfcn = machine epsilon
  const int64_t m = F.m;
  const int lane = F.lane;
  double diag[NP], qtf[NP], wa1[NP], wa2[NP], wa3[NP], acnorm[NP], rdiag[NP];
  int ipvt[NP];
  double r[NP][NP];
  int info = 0, nfev = 0, iter = 1;
  double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
  F.eval(x, F.fvec);
  nfev = 1;
  double fnorm = F.enorm_m(F.fvec);
  while (true) {
    // fdjac2
    for (int j = 0; j < NP; ++j) {
      const double temp = x[j];
      double h = eps * fabs(temp);
      if (h == 0.0) h = eps;
      x[j] = temp + h;
      double* col = F.fjac + (int64_t)j * m;
      F.eval(x, col);
      x[j] = temp;
      for (int64_t i = lane; i < m; i += 32) col[i] = (col[i] - F.fvec[i]) / h;
      __syncwarp();
    }
    nfev += NP;
    // qrfac with column pivoting; cols[] maps the logical to the stored column (swaps move the mapping, not the data)
    int cols[NP];
    double wa[NP];
    for (int j = 0; j < NP; ++j) {
      cols[j] = j;
      ipvt[j] = j;
      acnorm[j] = F.enorm_m(F.fjac + (int64_t)j * m);
      rdiag[j] = acnorm[j];
      wa[j] = rdiag[j];
    }
    const int minmn = m < NP ? (int)m : NP;
    for (int j = 0; j < minmn; ++j) {
      int kmax = j;
      for (int k = j; k < NP; ++k) if (rdiag[k] > rdiag[kmax]) kmax = k;
      if (kmax != j) {
        int t = cols[j]; cols[j] = cols[kmax]; cols[kmax] = t;
        rdiag[kmax] = rdiag[j];
        wa[kmax] = wa[j];
        t = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = t;
      }
      double* aj = F.fjac + (int64_t)cols[j] * m;
      double ajnorm = F.enorm_m(aj, j);
      if (ajnorm != 0.0) {
        const double ajj = aj[j];
        if (ajj < 0.0) ajnorm = -ajnorm;
        __syncwarp();
        for (int64_t i = j + lane; i < m; i += 32) aj[i] = aj[i] / ajnorm;
        __syncwarp();
        if (lane == 0) aj[j] += 1.0;
        __syncwarp();
        const double ajj1 = aj[j];
        for (int k = j + 1; k < NP; ++k) {
          double* ak = F.fjac + (int64_t)cols[k] * m;
          double s = 0.0;
          for (int64_t i = j + lane; i < m; i += 32) s += aj[i] * ak[i];
          const double temp = warp_sum(s) / ajj1;
          for (int64_t i = j + lane; i < m; i += 32) ak[i] -= temp * aj[i];
          __syncwarp();
          if (rdiag[k] != 0.0) {
            double t2 = ak[j] / rdiag[k];
            rdiag[k] *= sqrt(fmax(0.0, 1.0 - t2 * t2));
            t2 = rdiag[k] / wa[k];
            if (0.05 * t2 * t2 <= kEpsMch) {
              rdiag[k] = F.enorm_m(ak, j + 1);
              wa[k] = rdiag[k];
            }
          }
        }
      }
      rdiag[j] = -ajnorm;
    }
    // acnorm is indexed by original column (ipvt maps logical -> original)
    if (iter == 1) {
      for (int j = 0; j < NP; ++j) { diag[j] = acnorm[j]; if (acnorm[j] == 0.0) diag[j] = 1.0; }
      for (int j = 0; j < NP; ++j) wa3[j] = diag[j] * x[j];
      xnorm = enorm_n<NP>(wa3);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    // (Q^T) fvec -> qtf; R into registers
    for (int64_t i = lane; i < m; i += 32) F.wa4[i] = F.fvec[i];
    __syncwarp();
    for (int j = 0; j < NP; ++j) {
      double* aj = F.fjac + (int64_t)cols[j] * m;
      const double ajj = j < m ? aj[j] : 0.0;
      if (ajj != 0.0) {
        double s = 0.0;
        for (int64_t i = j + lane; i < m; i += 32) s += aj[i] * F.wa4[i];
        const double temp = -warp_sum(s) / ajj;
        for (int64_t i = j + lane; i < m; i += 32) F.wa4[i] += aj[i] * temp;
        __syncwarp();
      }
      qtf[j] = j < m ? F.wa4[j] : 0.0;
    }
    for (int j = 0; j < NP; ++j)
      for (int i = 0; i < NP; ++i) r[i][j] = 0.0;
    for (int j = 0; j < NP; ++j) {
      const double* aj = F.fjac + (int64_t)cols[j] * m;
      for (int i = 0; i < j; ++i) r[i][j] = i < m ? aj[i] : 0.0;
      r[j][j] = rdiag[j];
    }
    // norm of the scaled gradient
    gnorm = 0.0;
    if (fnorm != 0.0) {
      for (int j = 0; j < NP; ++j) {
        const int l = ipvt[j];
        if (acnorm[l] != 0.0) {
          double sum = 0.0;
          for (int i = 0; i <= j; ++i) sum += r[i][j] * (qtf[i] / fnorm);
          gnorm = fmax(gnorm, fabs(sum / acnorm[l]));
        }
      }
    }
    if (gnorm <= gtol) info = 4;
    if (info != 0) break;
    for (int j = 0; j < NP; ++j) diag[j] = fmax(diag[j], acnorm[j]);
    // inner loop
    double ratio = 0.0;
    do {
      double sdiag[NP];
      lmpar<NP>(r, ipvt, diag, qtf, delta, par, wa1, sdiag);
      for (int j = 0; j < NP; ++j) { wa1[j] = -wa1[j]; wa2[j] = x[j] + wa1[j]; wa3[j] = diag[j] * wa1[j]; }
      const double pnorm = enorm_n<NP>(wa3);
      if (iter == 1) delta = fmin(delta, pnorm);
      F.eval(wa2, F.wa4);
      ++nfev;
      const double fnorm1 = F.enorm_m(F.wa4);
      double actred = -1.0;
      if (0.1 * fnorm1 < fnorm) actred = 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
      for (int j = 0; j < NP; ++j) wa3[j] = 0.0;
      for (int j = 0; j < NP; ++j) {
        const double temp = wa1[ipvt[j]];
        for (int i = 0; i <= j; ++i) wa3[i] += r[i][j] * temp;
      }
      const double temp1 = enorm_n<NP>(wa3) / fnorm, temp2 = (sqrt(par) * pnorm) / fnorm;
      const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
      const double dirder = -(temp1 * temp1 + temp2 * temp2);
      ratio = prered != 0.0 ? actred / prered : 0.0;
      if (ratio <= 0.25) {
        double temp = 0.5;
        if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
        delta = temp * fmin(delta, pnorm / 0.1);
        par = par / temp;
      } else if (par == 0.0 || ratio >= 0.75) {
        delta = pnorm / 0.5;
        par = 0.5 * par;
      }
      if (ratio >= 1.0e-4) {
        for (int j = 0; j < NP; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
        for (int64_t i = lane; i < m; i += 32) F.fvec[i] = F.wa4[i];
        __syncwarp();
        xnorm = enorm_n<NP>(wa2);
        fnorm = fnorm1;
        ++iter;
      }
      const bool small = fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0;
      if (small) info = 1;
      if (delta <= xtol * xnorm) info = 2;
      if (small && info == 2) info = 3;
      if (info != 0) break;
      if (nfev >= maxfev) info = 5;
      if (fabs(actred) <= kEpsMch && prered <= kEpsMch && 0.5 * ratio <= 1.0) info = 6;
      if (delta <= kEpsMch * xnorm) info = 7;
      if (gnorm <= kEpsMch) info = 8;
      if (info != 0) break;
    } while (ratio < 1.0e-4);
    if (info != 0) break;
  }
  *nfev_out = nfev;
  return info;
}

struct RateArgs {
  const double* S;
  const int64_t* off_x;
  const int64_t* off_tau;
  int64_t n_edges;
  double* work;            // 5 doubles per mesh point (fvec, wa4, fjac[3]), laid out per edge at 5 * off_x[e]
  double* out;             // [n_edges][8]: kind (0 exp, 1 exp2), p0, p1, p2, k, info, nfev, N
  double perturb;          // start values are multiplied by 1 + perturb (0: the reference's; see hopb_rate_fit)
  unsigned noise;
};

__global__ void __launch_bounds__(128) ratefit_kernel(RateArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= a.n_edges) return;
  const int64_t x0 = a.off_x[e], m = a.off_x[e + 1] - x0;
  const int64_t n_tau = a.off_tau[e + 1] - a.off_tau[e];
  double* w = a.work + 5 * x0;
  double kind = 0.0, k = 0.0, p[3] = {0.0, 0.0, 0.0};
  int info = 0, nfev = 0;
  bool single = n_tau <= 10;            // hop/graph.py:1727: the double exponential needs more than 10 waiting times
  if (!single) {
    Fit<3> F{a.S + x0, m, w, w + m, w + 2 * m, lane, a.noise, 0u};
    p[0] = 0.5 * (1.0 + a.perturb); p[1] = 0.1 * (1.0 + a.perturb); p[2] = 1.0e-4 * (1.0 + a.perturb);   // fitExp2.initial_values
    info = lmdif<3>(F, p, &nfev);
    kind = 1.0;
    single = fabs(0.5 - p[0]) > 0.49;                            // very asymmetric fit
    if (p[1] < 0.0 || p[2] < 0.0) single = true;                 // at least one k < 0: no sum of exponentials
    else k = fmax(p[1] > 0.0 ? p[1] : -1.0e308, p[2] > 0.0 ? p[2] : -1.0e308);   // the dominant (fast) rate
  }
  if (single) {
    Fit<1> F{a.S + x0, m, w, w + m, w + 2 * m, lane, a.noise, 0u};
    double q[1] = {1.0 + a.perturb};                              // fitExp.initial_values
    info = lmdif<1>(F, q, &nfev);
    kind = 0.0;
    p[0] = q[0]; p[1] = 0.0; p[2] = 0.0;
    k = q[0];
  }
  if (lane == 0) {
    double* o = a.out + e * 8;
    o[0] = kind; o[1] = p[0]; o[2] = p[1]; o[3] = p[2]; o[4] = k; o[5] = (double)info; o[6] = (double)nfev; o[7] = (double)n_tau;
  }
}

}  // namespace

extern "C" {

int hopb_rate_fit(hopb_handle* h, void* stream, const double* tau, const int64_t* off_tau, const int64_t* off_x, int64_t n_edges,
                  int64_t total_x, double perturb, unsigned noise, double* S, double* out) {
  HOPB_REQUIRE(h, tau && off_tau && off_x && S && out && n_edges >= 0 && total_x >= 0, "bad argument");
  if (n_edges == 0) return HOPB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* work = nullptr;
  // histogram scratch (2 x u32 per mesh point) and the lmdif work vectors (5 doubles per mesh point)
  const size_t bytes_hist = (size_t)total_x * 2 * sizeof(uint32_t);
  const size_t bytes_fit = (size_t)total_x * 5 * sizeof(double);
  int rc = hopb_scratch(h, 16, bytes_fit + bytes_hist + 64, (void**)&work);
  if (rc) return rc;
  uint32_t* hist = reinterpret_cast<uint32_t*>(work + bytes_fit);
  HOPB_CUDA(h, cudaMemsetAsync(hist, 0, bytes_hist, s));
  const int64_t sblocks = n_edges < (int64_t)h->num_sms * 8 ? n_edges : (int64_t)h->num_sms * 8;
  HOPB_K(survival_kernel)<<<(unsigned)sblocks, 256, 0, s>>>(tau, off_tau, off_x, n_edges, S, hist, total_x);
  RateArgs a{S, off_x, off_tau, n_edges, reinterpret_cast<double*>(work), out, perturb, noise};
  HOPB_K(ratefit_kernel)<<<(unsigned)((n_edges * 32 + 127) / 128), 128, 0, s>>>(a);
  HOPB_LAUNCH_CHECK(h);
  return HOPB_OK;
}

}  // extern "C"
