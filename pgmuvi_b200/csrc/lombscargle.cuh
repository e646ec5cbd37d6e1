// pgmuvi_b200 - N2: batched generalised Lomb-Scargle periodogram + peak picking on device.
//
// The step immediately BEFORE the hot path: Lightcurve.fit seeds the spectral-mixture means
// from the highest periodogram peaks (pgmuvi/lightcurve.py:4214-4611 fit_LS, 5475-5653).  The
// reference calls astropy.timeseries.LombScargle(t, y, dy).power(autofrequency()) (floating
// mean, centred data, 'standard' normalisation) and scipy.signal.find_peaks(distance=...);
// here one launch evaluates the exact O(n * nf) floating-mean periodogram (Zechmeister &
// Kuerster 2009, the formulation of astropy's "slow" implementation) for B light curves.
//
// Work split: block = (1024 consecutive frequencies, light curve); each thread owns R = 4
// consecutive frequencies and walks all points, which are staged in shared memory together
// with the per-point rotation (cos, sin)(2 pi df t_i): one direct sincospi per (thread, point),
// then three rotations.  Six running sums per frequency (S, C, S2, C2, YS, YC); the time-shift
// tau and the shifted sums follow algebraically at the end.  FP64-pipe bound: about 19 FP64
// operations per (frequency, point) pair; no HBM traffic beyond 3n doubles in, nf doubles out.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace pgm {

constexpr int LS_THREADS = 256;
constexpr int LS_R = 4;                       // consecutive frequencies per thread
constexpr int LS_FPB = LS_THREADS * LS_R;     // frequencies per block
constexpr int LS_CHUNK = 512;                 // points staged per pass

#define PGM_LS_FIT_MEAN 1
#define PGM_LS_CENTER_DATA 2

struct LsArgs {
  const double* t;        // [B, n_max]
  const int32_t* n_valid; // [B] or null
  const double* y;        // [B, n_max]
  const double* dy;       // [B, n_max] or null (unit errors)
  const double* f0;       // [B] first frequency
  const double* df;       // [B] frequency step
  const int32_t* nf;      // [B] frequencies per light curve (<= nf_max)
  int B, n_max, nf_max, flags;
  double* power;          // [B, nf_max]
};

__device__ __forceinline__ double ls_block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w8 = 0; w8 < LS_THREADS / 32; ++w8) s += red[w8];
  return s;
}

__global__ void __launch_bounds__(LS_THREADS) ls_power_kernel(LsArgs A) {
  __shared__ double s_x[LS_CHUNK];            // 2 (t_i - t_0): argument of sincospi per unit f
  __shared__ double s_w[LS_CHUNK];            // normalised weight
  __shared__ double s_wy[LS_CHUNK];           // w * (y - mean)
  __shared__ double2 s_rot[LS_CHUNK];         // (cos, sin)(2 pi df (t_i - t_0))
  __shared__ double red[LS_THREADS / 32];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int n = A.n_valid ? A.n_valid[b] : A.n_max;
  const int nf = A.nf[b];
  const int k0 = blockIdx.x * LS_FPB + tid * LS_R;   // first frequency of this thread
  if (blockIdx.x * LS_FPB >= nf || n < 1) return;
  const double* tb = A.t + (size_t)b * A.n_max;
  const double* yb = A.y + (size_t)b * A.n_max;
  const double* eb = A.dy ? A.dy + (size_t)b * A.n_max : nullptr;
  const double f0 = A.f0[b], df = A.df[b];
  const bool fit_mean = (A.flags & PGM_LS_FIT_MEAN) != 0;
  const bool center = fit_mean || (A.flags & PGM_LS_CENTER_DATA) != 0;

  // weights w = dy^-2 / sum, weighted mean and YY = sum w y^2 (of the centred data)
  double sw = 0.0, swy = 0.0;
  for (int i = tid; i < n; i += LS_THREADS) {
    const double w = eb ? 1.0 / (eb[i] * eb[i]) : 1.0;
    sw += w;
    swy += w * yb[i];
  }
  const double wsum = ls_block_sum(sw, red);
  const double ymean = center ? ls_block_sum(swy, red) / wsum : 0.0;
  double syy = 0.0, sy = 0.0;
  for (int i = tid; i < n; i += LS_THREADS) {
    const double w = (eb ? 1.0 / (eb[i] * eb[i]) : 1.0) / wsum;
    const double yc = yb[i] - ymean;
    syy += w * yc * yc;
    sy += w * yc;
  }
  const double YY = ls_block_sum(syy, red);
  const double Y = ls_block_sum(sy, red);
  const double t0 = tb[0];

  double S[LS_R], C[LS_R], S2[LS_R], C2[LS_R], YS[LS_R], YC[LS_R];
#pragma unroll
  for (int r = 0; r < LS_R; ++r) S[r] = C[r] = S2[r] = C2[r] = YS[r] = YC[r] = 0.0;
  const double fk = f0 + (double)k0 * df;

  for (int base = 0; base < n; base += LS_CHUNK) {
    const int cnt = min(LS_CHUNK, n - base);
    __syncthreads();
    for (int i = tid; i < cnt; i += LS_THREADS) {
      const double w = (eb ? 1.0 / (eb[base + i] * eb[base + i]) : 1.0) / wsum;
      const double x = 2.0 * (tb[base + i] - t0);
      s_x[i] = x;
      s_w[i] = w;
      s_wy[i] = w * (yb[base + i] - ymean);
      double sn, cs;
      sincospi(df * x, &sn, &cs);
      s_rot[i] = make_double2(cs, sn);
    }
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < cnt; ++i) {
      const double w = s_w[i], wy = s_wy[i];
      const double2 rot = s_rot[i];
      double sn, cs;
      sincospi(fk * s_x[i], &sn, &cs);
#pragma unroll
      for (int r = 0; r < LS_R; ++r) {
        S[r] = fma(w, sn, S[r]);
        C[r] = fma(w, cs, C[r]);
        YS[r] = fma(wy, sn, YS[r]);
        YC[r] = fma(wy, cs, YC[r]);
        S2[r] = fma(w, sn * cs, S2[r]);                  // x2 at the end
        C2[r] = fma(w, fma(cs, cs, -sn * sn), C2[r]);
        if (r + 1 < LS_R) {                              // rotate to the next frequency
          const double c2 = fma(cs, rot.x, -sn * rot.y);
          sn = fma(sn, rot.x, cs * rot.y);
          cs = c2;
        }
      }
    }
  }

#pragma unroll
  for (int r = 0; r < LS_R; ++r) {
    const int k = k0 + r;
    if (k >= nf) break;
    double s2 = 2.0 * S2[r], c2 = C2[r];
    if (fit_mean) {
      s2 -= 2.0 * S[r] * C[r];
      c2 -= C[r] * C[r] - S[r] * S[r];
    }
    // omega tau = atan2(s2, c2) / 2; shifted sums by the angle-difference identities
    const double two_wt = atan2(s2, c2);
    double st, ct, s2t, c2t;
    sincos(0.5 * two_wt, &st, &ct);
    sincos(two_wt, &s2t, &c2t);
    double yct = YC[r] * ct + YS[r] * st;
    double yst = YS[r] * ct - YC[r] * st;
    const double C2raw = C2[r], S2raw = 2.0 * S2[r];
    double cct = 0.5 * (1.0 + C2raw * c2t + S2raw * s2t);
    double sst = 0.5 * (1.0 - C2raw * c2t - S2raw * s2t);
    if (fit_mean) {
      const double ctau = C[r] * ct + S[r] * st;
      const double stau = S[r] * ct - C[r] * st;
      yct -= Y * ctau;
      yst -= Y * stau;
      cct -= ctau * ctau;
      sst -= stau * stau;
    }
    A.power[(size_t)b * A.nf_max + k] = (yct * yct / cct + yst * yst / sst) / YY;
  }
}

// ---- peak picking: scipy.signal.find_peaks(power, distance=d), highest first ---------------
// Local maxima (strict neighbours; flat tops report their midpoint, scipy's rule), then the
// greedy height-ordered distance filter: the highest remaining peak is kept and every peak
// closer than `distance` samples to it is dropped.  One block per light curve, num_peaks rounds
// of a block-wide arg-max; the candidate mask lives in global scratch (`mask`, [B, nf_max]).
struct LsPeakArgs {
  const double* power;   // [B, nf_max]
  const int32_t* nf;     // [B]
  int B, nf_max, distance, num_peaks;
  uint8_t* mask;         // [B, nf_max] scratch
  int32_t* peak_idx;     // [B, num_peaks], -1 padded
  double* peak_power;    // [B, num_peaks], NaN padded
};

__global__ void __launch_bounds__(LS_THREADS) ls_peaks_kernel(LsPeakArgs A) {
  __shared__ double s_val[LS_THREADS / 32];
  __shared__ int s_idx[LS_THREADS / 32];
  __shared__ int s_best;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nf = A.nf[b];
  const double* p = A.power + (size_t)b * A.nf_max;
  uint8_t* m = A.mask + (size_t)b * A.nf_max;
  for (int i = tid; i < nf; i += LS_THREADS) {
    uint8_t is_peak = 0;
    if (i > 0 && i < nf - 1 && p[i - 1] < p[i]) {
      int ahead = i + 1;
      while (ahead < nf - 1 && p[ahead] == p[i]) ++ahead;
      if (p[ahead] < p[i]) is_peak = 2;               // rising edge of a (flat) maximum
      if (is_peak && ahead - 1 > i) is_peak = 3;      // plateau: mark, midpoint resolved below
    }
    m[i] = is_peak;
  }
  __syncthreads();
  // plateaus: move the mark from the left edge to the midpoint
  for (int i = tid; i < nf; i += LS_THREADS) {
    if (m[i] == 3) {
      int ahead = i + 1;
      while (ahead < nf - 1 && p[ahead] == p[i]) ++ahead;
      m[i] = 0;
      m[(i + ahead - 1) / 2] = 1;
    } else if (m[i] == 2) {
      m[i] = 1;
    }
  }
  __syncthreads();
  for (int round = 0; round < A.num_peaks; ++round) {
    double best = -INFINITY;
    int bi = -1;
    for (int i = tid; i < nf; i += LS_THREADS)
      if (m[i] == 1 && (p[i] > best || (p[i] == best && i > bi))) { best = p[i]; bi = i; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi > bi))) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double bv = -INFINITY;
      int bj = -1;
      for (int w8 = 0; w8 < LS_THREADS / 32; ++w8)
        if (s_idx[w8] >= 0 && (bj < 0 || s_val[w8] > bv || (s_val[w8] == bv && s_idx[w8] > bj))) {
          bv = s_val[w8];
          bj = s_idx[w8];
        }
      s_best = bj;
      A.peak_idx[(size_t)b * A.num_peaks + round] = bj;
      A.peak_power[(size_t)b * A.num_peaks + round] = bj >= 0 ? bv : nan("");
    }
    __syncthreads();
    const int bj = s_best;
    if (bj < 0) {   // no peak left: pad the tail and stop (uniform over the block)
      for (int r2 = round + 1 + tid; r2 < A.num_peaks; r2 += LS_THREADS) {
        A.peak_idx[(size_t)b * A.num_peaks + r2] = -1;
        A.peak_power[(size_t)b * A.num_peaks + r2] = nan("");
      }
      break;
    }
    const int lo = max(0, bj - A.distance + 1), hi = min(nf - 1, bj + A.distance - 1);
    for (int i = lo + tid; i <= hi; i += LS_THREADS) m[i] = 0;
    __syncthreads();
  }
}

}  // namespace pgm

// ------------------------------------------------------------------------------------
// N4 (first stage): spectral-mixture PSD of every fitted light curve on the reference's
// log-spaced frequency grid and its dominant peak
//   PSD(f) = sum_q w_q exp(-0.5 ((f - mu_q) / sigma_q)^2)          (lightcurve.py:6537-6578)
//   grid   = logspace(log10 fmin, log10 fmax, n_grid)              (:7474-7482)
//   dominant = highest local maximum (scipy.signal.find_peaks), or the arg-max when the PSD
//   has no interior peak                                           (:7931-7940)
// One block per light curve; the PSD row is optional output.
// ------------------------------------------------------------------------------------
namespace pgm {
struct PsdArgs {
  const double* freq;    // [B, Q] component frequencies (raw units)
  const double* fscale;  // [B, Q] component frequency scales
  const double* weight;  // [B, Q]
  const double* fmin;    // [B]
  const double* fmax;    // [B]
  int B, Q, n_grid;
  double* grid;          // [B, n_grid] or null
  double* psd;           // [B, n_grid] (required: scratch + output)
  int32_t* dom_idx;      // [B]
  double* dom_freq;      // [B]
  double* dom_height;    // [B]
  int32_t* n_peaks;      // [B] number of local maxima
};

__global__ void __launch_bounds__(LS_THREADS) sm_psd_peak_kernel(PsdArgs A) {
  __shared__ double s_mu[8], s_is[8], s_w[8];
  __shared__ double s_val[LS_THREADS / 32];
  __shared__ int s_idx[LS_THREADS / 32], s_cnt[LS_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < A.Q) {
    s_mu[tid] = A.freq[(size_t)b * A.Q + tid];
    s_is[tid] = 1.0 / A.fscale[(size_t)b * A.Q + tid];
    s_w[tid] = A.weight[(size_t)b * A.Q + tid];
  }
  __syncthreads();
  const double l0 = log10(A.fmin[b]), l1 = log10(A.fmax[b]);
  const double step = (A.n_grid > 1) ? (l1 - l0) / (double)(A.n_grid - 1) : 0.0;
  double* p = A.psd + (size_t)b * A.n_grid;
  for (int k = tid; k < A.n_grid; k += LS_THREADS) {
    // numpy.logspace: 10 ** (start + k * step), last sample pinned to the end point
    const double f = (k == A.n_grid - 1 && A.n_grid > 1) ? pow(10.0, l1) : pow(10.0, l0 + k * step);
    double s = 0.0;
    for (int q = 0; q < A.Q; ++q) {
      const double z = (f - s_mu[q]) * s_is[q];
      s += s_w[q] * exp(-0.5 * z * z);
    }
    p[k] = s;
    if (A.grid) A.grid[(size_t)b * A.n_grid + k] = f;
  }
  __syncthreads();
  // highest strict local maximum (plateaus: left-edge rule is enough for a smooth PSD), the
  // global arg-max as fall-back, and the number of local maxima
  double best = -INFINITY, gbest = -INFINITY;
  int bi = -1, gi = -1, cnt = 0;
  for (int k = tid; k < A.n_grid; k += LS_THREADS) {
    const double v = p[k];
    if (v > gbest || (v == gbest && k < gi)) { gbest = v; gi = k; }
    if (k > 0 && k < A.n_grid - 1 && p[k - 1] < v && v > p[k + 1]) {
      ++cnt;
      if (v > best || (v == best && k < bi)) { best = v; bi = k; }
    }
  }
  auto merge = [](double& v, int& i, double ov, int oi) {
    if (oi >= 0 && (i < 0 || ov > v || (ov == v && oi < i))) { v = ov; i = oi; }
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    merge(best, bi, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
    merge(gbest, gi, __shfl_xor_sync(0xffffffffu, gbest, o), __shfl_xor_sync(0xffffffffu, gi, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  // two rounds through shared memory: peaks first, then the global maximum
  if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; s_cnt[warp] = cnt; }
  __syncthreads();
  double fb = -INFINITY;
  int fi = -1, fc = 0;
  if (tid == 0)
    for (int w8 = 0; w8 < LS_THREADS / 32; ++w8) { merge(fb, fi, s_val[w8], s_idx[w8]); fc += s_cnt[w8]; }
  __syncthreads();
  if (lane == 0) { s_val[warp] = gbest; s_idx[warp] = gi; }
  __syncthreads();
  if (tid == 0) {
    double gb = -INFINITY;
    int gj = -1;
    for (int w8 = 0; w8 < LS_THREADS / 32; ++w8) merge(gb, gj, s_val[w8], s_idx[w8]);
    const int d = fi >= 0 ? fi : gj;
    A.dom_idx[b] = d;
    A.dom_height[b] = p[d];
    const double f = (d == A.n_grid - 1 && A.n_grid > 1) ? pow(10.0, l1) : pow(10.0, l0 + d * step);
    A.dom_freq[b] = f;
    A.n_peaks[b] = fc;
  }
}
}  // namespace pgm
