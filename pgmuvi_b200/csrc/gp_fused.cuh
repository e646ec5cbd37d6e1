// pgmuvi_b200 - fused exact-GP evaluation for sm_100a (B200).
//
// One thread block owns one light curve at a time and runs the whole path for it:
//
//   (1) spectral-mixture K(x,x') tiles generated on the fly (never written to HBM)
//   (2) left-looking blocked Cholesky (64x64 tiles, trailing updates on FP64 tensor cores
//       through mma.sync.m8n8k4.f64 -> SASS DMMA), forward solve, log-det
//   (3) tiled triangular inverse + K^-1 = X^T X regenerated tile by tile and contracted at
//       once with dK/dtheta (regenerated per tile) -> gradient; K^-1 never materialised
//   (4) constraint chain rule, optional optimiser step (fit kernel)
//
// Reference semantics: gpytorch SpectralMixtureKernel / ExactMarginalLogLikelihood as driven
// by pgmuvi/trainers.py:177-182 and pgmuvi/gps.py:205-220, 302-318.
//
// Per-block scratch lives in a global workspace indexed by blockIdx.x (reused for every
// light curve the block processes, so it stays L2-resident); only L / L^-1 tiles go there.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace pgm {

constexpr int TS = 64;          // tile edge
constexpr int TT = TS * TS;     // elements per tile
constexpr int KC = 16;          // k-chunk per pipeline stage
constexpr int NTHREADS = 256;   // 8 warps: 2 (M) x 4 (N), warp tile 32 x 16
constexpr int NSTAGES = 3;
constexpr int LD_NT = 20;       // smem ld of a [64][KC] chunk   (ld % 16 == 4 -> conflict free)
constexpr int LD_KM = 68;       // smem ld of a [KC][64] chunk
constexpr int OPBUF = 1280;     // elements per operand per stage  (>= 64*20, 16*68)
constexpr int LD_S = 68;        // smem ld of the 64x64 work tile
constexpr int STAGE_ELEMS = NSTAGES * 2 * OPBUF;   // 7680
constexpr int S_ELEMS = TS * LD_S;                 // 4352

#define PGM_KIND_SM1D 0
#define PGM_KIND_SM_ARD_PRODSUM 1
#define PGM_KIND_SM_ARD_SUMPROD 2
#define PGM_FLAG_GRAD 1
#define PGM_FLAG_LEARN_NOISE 2
#define PGM_FLAG_BOUNDS_PER_LC 4

// ------------------------------------------------------------------------------------
// small PTX wrappers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_f64(double (&d)[2], double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d[0]), "+d"(d[1])
      : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ double shfl_d(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}

__host__ __device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// ------------------------------------------------------------------------------------
// static configuration per (kernel kind, padded mixture count, input dims)
// ------------------------------------------------------------------------------------
template <int KIND, int QT, int D>
struct Cfg {
  static constexpr int NFB = D + 2 * D * QT;  // per-point fields: x[D], (cos,sin)[D][QT]
  static constexpr int NF = NFB + 1;          // + alpha
  static constexpr int NG = QT + 2 * QT * D;  // kernel-gradient accumulators
  static constexpr int NV = NG + 1;           // + tr W
  static constexpr int PMAX = 2 + QT + 2 * QT * D;
  // shared memory (doubles)
  static constexpr int SM_STAGES = 0;
  static constexpr int SM_S = STAGE_ELEMS;
  static constexpr int SM_ROW = SM_S + S_ELEMS;
  static constexpr int SM_COL = SM_ROW + NF * TS;
  static constexpr int SM_PAR = SM_COL + NF * TS;
  // par block: theta[PMAX] jac[PMAX] w[QT] a[QT*D] red[8*NV] fin[NV+4] tmp[64]
  static constexpr int PAR_THETA = 0;
  static constexpr int PAR_JAC = PAR_THETA + PMAX;
  static constexpr int PAR_W = PAR_JAC + PMAX;
  static constexpr int PAR_A = PAR_W + QT;
  static constexpr int PAR_RED = PAR_A + QT * D;
  static constexpr int PAR_FIN = PAR_RED + 8 * (NV + 2);
  static constexpr int PAR_TMP = PAR_FIN + NV + 4;
  static constexpr int PAR_RAW = PAR_TMP + 4 * TS;   // raw / adam state for the fit kernel
  static constexpr int PAR_END = PAR_RAW + 3 * PMAX;
  static constexpr int SM_TOTAL = SM_PAR + PAR_END + 8;
  static constexpr size_t SMEM_BYTES = (size_t)SM_TOTAL * sizeof(double);
};

// per-block global scratch layout (doubles)
struct Scratch {
  double* tiles;  // ntri * TT : L_ij (i>j) then X_ij; diagonal slots hold X_jj = L_jj^-1
  double* ctmp;   // TT
  double* fld;    // NF * npad : x[D], trig, alpha
  double* rhs;    // npad  (y - mean)
  double* z;      // npad  (L^-1 rhs)
  double* dn;     // npad  (diagonal noise)
};
__host__ __device__ inline size_t scratch_elems(int n_max, int NF) {
  int N = (n_max + TS - 1) / TS;
  size_t npad = (size_t)N * TS;
  return (size_t)tri(N, 0) * TT + TT + (size_t)NF * npad + 3 * npad + 64;
}

// ------------------------------------------------------------------------------------
// the tensor-core tile engine:  acc += sum_kt  opA(tileA(kt)) * opB(tileB(kt))^T
//   opX element (row r of the product operand, k):   NT: tile[r*64 + k]   KM: tile[k*64 + r]
// A,B tiles are global (L2-resident scratch); 3-stage cp.async pipeline into padded smem.
// ------------------------------------------------------------------------------------
template <bool A_KM, bool B_KM, typename FA, typename FB>
__device__ __forceinline__ void gemm_tiles(double (&acc)[4][2][2], int nk, FA tileA, FB tileB,
                                           double* __restrict__ stages) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int nchunks = nk * (TS / KC);

  auto issue = [&](int c) {
    const int kt = c >> 2, kc = c & 3;
    double* sA = stages + (c % NSTAGES) * (2 * OPBUF);
    double* sB = sA + OPBUF;
    const double* gA = tileA(kt);
    const double* gB = tileB(kt);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int p = tid + h * NTHREADS;
      if (!A_KM) {
        const int row = p >> 3, seg = p & 7;
        cp_async16(sA + row * LD_NT + seg * 2, gA + row * TS + kc * KC + seg * 2);
      } else {
        const int kr = p >> 5, seg = p & 31;
        cp_async16(sA + kr * LD_KM + seg * 2, gA + (kc * KC + kr) * TS + seg * 2);
      }
      if (!B_KM) {
        const int row = p >> 3, seg = p & 7;
        cp_async16(sB + row * LD_NT + seg * 2, gB + row * TS + kc * KC + seg * 2);
      } else {
        const int kr = p >> 5, seg = p & 31;
        cp_async16(sB + kr * LD_KM + seg * 2, gB + (kc * KC + kr) * TS + seg * 2);
      }
    }
  };

  __syncthreads();  // previous users of the stage buffers / producers of the tiles are done
#pragma unroll
  for (int s = 0; s < NSTAGES - 1; ++s) {
    if (s < nchunks) issue(s);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<NSTAGES - 2>();
    __syncthreads();
    if (c + NSTAGES - 1 < nchunks) issue(c + NSTAGES - 1);
    cp_async_commit();
    const double* sA = stages + (c % NSTAGES) * (2 * OPBUF);
    const double* sB = sA + OPBUF;
#pragma unroll
    for (int k4 = 0; k4 < KC / 4; ++k4) {
      double a[4], b[2];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int r = wm * 32 + mi * 8 + g;
        a[mi] = A_KM ? sA[(k4 * 4 + tq) * LD_KM + r] : sA[r * LD_NT + k4 * 4 + tq];
      }
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        const int r = wn * 16 + ni * 8 + g;
        b[ni] = B_KM ? sB[(k4 * 4 + tq) * LD_KM + r] : sB[r * LD_NT + k4 * 4 + tq];
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) mma_f64(acc[mi][ni], a[mi], b[ni]);
    }
  }
  cp_async_wait<0>();
}

__device__ __forceinline__ void zero_acc(double (&acc)[4][2][2]) {
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
}

// fragment entry -> (row, col) inside the 64x64 tile
#define PGM_FRAG_LOOP(mi, ni, e, r, c)                                      \
  _Pragma("unroll") for (int mi = 0; mi < 4; ++mi)                          \
  _Pragma("unroll") for (int ni = 0; ni < 2; ++ni)                          \
  _Pragma("unroll") for (int e = 0; e < 2; ++e)                             \
    if (const int r = (threadIdx.x >> 7) * 32 + mi * 8 + ((threadIdx.x & 31) >> 2); true) \
      if (const int c = ((threadIdx.x >> 5) & 3) * 16 + ni * 8 + (threadIdx.x & 3) * 2 + e; true)

__device__ __forceinline__ void store_acc_tile(const double (&acc)[4][2][2], double* tile,
                                               double sign) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3, wm = warp >> 2, wn = warp & 3;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) {
      const int r = wm * 32 + mi * 8 + g, c = wn * 16 + ni * 8 + tq * 2;
      double2 v = make_double2(sign * acc[mi][ni][0], sign * acc[mi][ni][1]);
      *reinterpret_cast<double2*>(tile + r * TS + c) = v;
    }
}

// ------------------------------------------------------------------------------------
// kernel entry K(x_i, x_j) and its hyper-parameter derivatives.
// rowv / colv: per-point fields of the 64 rows / cols of the tile, field-major [f][64]:
//   f < D: centred x;  f = D + (dd*QT + q)*2 + {0,1}: cos / sin of 2 pi mu_qd x_d.
// cos(2 pi mu tau) = c_i c_j + s_i s_j,  sin(2 pi mu tau) = s_i c_j - c_i s_j.
// ------------------------------------------------------------------------------------
template <int KIND, int QT, int D>
__device__ __forceinline__ double k_entry(const double* __restrict__ rowv,
                                          const double* __restrict__ colv, int r, int c,
                                          const double* __restrict__ w,
                                          const double* __restrict__ a) {
  double tau2[D];
#pragma unroll
  for (int dd = 0; dd < D; ++dd) {
    const double tau = rowv[dd * TS + r] - colv[dd * TS + c];
    tau2[dd] = tau * tau;
  }
  if (KIND == PGM_KIND_SM_ARD_SUMPROD) {
    double k = 0.0;
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      double pr = w[q];
#pragma unroll
      for (int dd = 0; dd < D; ++dd) {
        const int f = D + (dd * QT + q) * 2;
        const double C = rowv[f * TS + r] * colv[f * TS + c] +
                         rowv[(f + 1) * TS + r] * colv[(f + 1) * TS + c];
        pr *= exp(-a[q * D + dd] * tau2[dd]) * C;
      }
      k += pr;
    }
    return k;
  } else {
    double k = 1.0;
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        const int f = D + (dd * QT + q) * 2;
        const double C = rowv[f * TS + r] * colv[f * TS + c] +
                         rowv[(f + 1) * TS + r] * colv[(f + 1) * TS + c];
        s += w[q] * exp(-a[q * D + dd] * tau2[dd]) * C;
      }
      k *= s;
    }
    return k;
  }
}

// accumulate  wgt * dK/dtheta  into ga[NG] = { gw[q], gmu[q*D+dd], gsg[q*D+dd] } (raw sums;
// the constant factors -2 pi w_q and -4 pi^2 sigma w_q are applied once at the end).
template <int KIND, int QT, int D>
__device__ __forceinline__ void k_grad_entry(const double* __restrict__ rowv,
                                             const double* __restrict__ colv, int r, int c,
                                             const double* __restrict__ w,
                                             const double* __restrict__ a, double wgt,
                                             double (&ga)[QT + 2 * QT * D]) {
  double tau[D], EC[D][QT], ES[D][QT], Ssum[D];
#pragma unroll
  for (int dd = 0; dd < D; ++dd) {
    tau[dd] = rowv[dd * TS + r] - colv[dd * TS + c];
    const double t2 = tau[dd] * tau[dd];
    Ssum[dd] = 0.0;
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      const int f = D + (dd * QT + q) * 2;
      const double ci = rowv[f * TS + r], si = rowv[(f + 1) * TS + r];
      const double cj = colv[f * TS + c], sj = colv[(f + 1) * TS + c];
      const double E = exp(-a[q * D + dd] * t2);
      EC[dd][q] = E * (ci * cj + si * sj);
      ES[dd][q] = E * (si * cj - ci * sj);
      Ssum[dd] += w[q] * EC[dd][q];
    }
  }
#pragma unroll
  for (int dd = 0; dd < D; ++dd) {
    const double wt = wgt * tau[dd];
    const double wt2 = wt * tau[dd];
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      double R = 1.0;  // product of the other dimensions' factor
      if (D == 2) R = (KIND == PGM_KIND_SM_ARD_SUMPROD) ? EC[1 - dd][q] : Ssum[1 - dd];
      const double ecr = EC[dd][q] * R;
      if (KIND == PGM_KIND_SM_ARD_SUMPROD) {
        if (dd == 0) ga[q] += wgt * ecr;
      } else {
        ga[q] += wgt * ecr;
      }
      ga[QT + q * D + dd] += wt * (ES[dd][q] * R);
      ga[QT + QT * D + q * D + dd] += wt2 * ecr;
    }
  }
}

// ------------------------------------------------------------------------------------
// 64x64 diagonal block:  S (lower) -> L in place,  X = L^-1 into S2.   256 threads.
// Panel width 8: warp 0 factors the panel in registers (shuffles), everybody applies the
// trailing update and forms the next 8 rows of L^-1 by block forward substitution.
// Returns sum log(pivot) (= log det of the block) in *logdet_out (thread 0), sets *fail.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void potrf_inv_64(double* __restrict__ S, double* __restrict__ S2,
                                             int* fail, double* logdet_acc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int p = 0; p < 8; ++p) {
    const int c0 = p * 8;
    if (warp == 0) {
      const int r0 = c0 + lane, r1 = c0 + lane + 32;
      double a0[8], a1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a0[k] = (r0 < TS) ? S[r0 * LD_S + c0 + k] : 0.0;
        a1[k] = (r1 < TS) ? S[r1 * LD_S + c0 + k] : 0.0;
      }
      double ld = 0.0;
      bool bad = false, isnan_ = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double dpiv = shfl_d(a0[k], k);
        if (!(dpiv > 0.0)) { bad = true; if (dpiv != dpiv) isnan_ = true; }
        const double rs = bad ? 1.0 : rsqrt(dpiv);
        ld += bad ? 0.0 : log(dpiv);
        a0[k] *= rs;
        a1[k] *= rs;
#pragma unroll
        for (int c = k + 1; c < 8; ++c) {
          const double lc = shfl_d(a0[k], c);
          a0[c] -= a0[k] * lc;
          a1[c] -= a1[k] * lc;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (r0 < TS) S[r0 * LD_S + c0 + k] = a0[k];
        if (r1 < TS) S[r1 * LD_S + c0 + k] = a1[k];
      }
      if (lane == 0) {
        *logdet_acc += ld;
        if (bad) atomicOr(fail, isnan_ ? 2 : 1);
      }
    }
    __syncthreads();
    // (i) trailing update of the lower triangle below/right of the panel
    const int m = TS - c0 - 8;
    for (int idx = tid; idx < m * m; idx += NTHREADS) {
      const int rr = idx / m, cc = idx - rr * m;
      if (cc <= rr) {
        const int r = c0 + 8 + rr, c = c0 + 8 + cc;
        double s = S[r * LD_S + c];
#pragma unroll
        for (int k = 0; k < 8; ++k) s -= S[r * LD_S + c0 + k] * S[c * LD_S + c0 + k];
        S[r * LD_S + c] = s;
      }
    }
    // (ii) rows c0..c0+7 of T = I - L[c0.., :c0] X[:c0, :]   (columns 0..c0+7)
    for (int idx = tid; idx < 8 * (c0 + 8); idx += NTHREADS) {
      const int rr = idx / (c0 + 8), c = idx - rr * (c0 + 8);
      const int r = c0 + rr;
      double s = (r == c) ? 1.0 : 0.0;
      for (int k = c; k < c0; ++k) s -= S[r * LD_S + k] * S2[k * LD_S + c];
      S2[r * LD_S + c] = s;
    }
    __syncthreads();
    // (iii) 8x8 triangular solve per column (threads 32.., overlaps the next panel on warp 0)
    if (tid >= 32 && tid < 32 + c0 + 8) {
      const int c = tid - 32;
      double v[8];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) v[rr] = S2[(c0 + rr) * LD_S + c];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        double s = v[rr];
#pragma unroll
        for (int kk = 0; kk < rr; ++kk) s -= S[(c0 + rr) * LD_S + c0 + kk] * v[kk];
        v[rr] = s / S[(c0 + rr) * LD_S + c0 + rr];
      }
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) S2[(c0 + rr) * LD_S + c] = v[rr];
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------
// arguments
// ------------------------------------------------------------------------------------
struct EvalArgs {
  const double* x;
  const int32_t* n_valid;
  const double* y;
  const double* fixed_noise;
  const double* raw;
  const int32_t* con_kind;
  const double* con_lb;
  const double* con_ub;
  int B, n_max, Q, flags;
  double* mll;
  double* grad;
  int32_t* info;
  double* ws;
  size_t ws_per_block;  // elements
};

struct FitArgs {
  EvalArgs e;
  double* raw_io;
  int optim_kind;
  double lr, beta1, beta2, eps, weight_decay, stop;
  int maxiter, miniter, stopavg;
  double* loss_hist;
  double* raw_hist;
  int32_t* n_iter;
};

__device__ __forceinline__ double softplus_d(double x) {
  return x > 30.0 ? x : log1p(exp(x));
}
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

template <int NVAL>
__device__ __forceinline__ void block_reduce(double (&v)[NVAL], double* red, double* fin) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < NVAL; ++i) {
    double s = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += shfl_xor_d(s, o);
    if (lane == 0) red[warp * NVAL + i] = s;
  }
  __syncthreads();
  if (tid < NVAL) {
    double s = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < NTHREADS / 32; ++w8) s += red[w8 * NVAL + tid];
    fin[tid] = s;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------
// one full evaluation of light curve b with raw parameters `raw` (global or shared).
// Results: smem par[PAR_FIN..]: mll in fin[NV+0]; gradient written to grad_out[P] (may be
// shared or global); returns info.
// ------------------------------------------------------------------------------------
template <int KIND, int QT, int D>
__device__ int eval_lightcurve(const EvalArgs& A, int b, const double* raw, double* sm,
                               const Scratch& sc, double* mll_out, double* grad_out) {
  using C = Cfg<KIND, QT, D>;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Q = A.Q;
  const bool learn_noise = (A.flags & PGM_FLAG_LEARN_NOISE) != 0;
  const bool want_grad = (A.flags & PGM_FLAG_GRAD) != 0;
  const int P = 1 + Q + 2 * Q * D + (learn_noise ? 1 : 0);
  const int n = A.n_valid ? A.n_valid[b] : A.n_max;
  const int N = (n + TS - 1) / TS;
  const int npad = N * TS;

  double* stages = sm + C::SM_STAGES;
  double* S2 = stages;  // aliases the pipeline buffers (idle during diagonal blocks)
  double* S = sm + C::SM_S;
  double* rowv = sm + C::SM_ROW;
  double* colv = sm + C::SM_COL;
  double* par = sm + C::SM_PAR;
  double* theta = par + C::PAR_THETA;
  double* jac = par + C::PAR_JAC;
  double* wq = par + C::PAR_W;
  double* aq = par + C::PAR_A;
  double* red = par + C::PAR_RED;
  double* fin = par + C::PAR_FIN;
  double* tmpv = par + C::PAR_TMP;
  int* s_fail = reinterpret_cast<int*>(sm + C::SM_PAR + C::PAR_END);
  double* s_logdet = sm + C::SM_PAR + C::PAR_END + 1;

  __syncthreads();
  // ---- constraints: raw -> theta, d theta / d raw  (A.2) -----------------------------
  if (tid < P) {
    const double rv = raw[tid];
    const int kd = A.con_kind[tid];
    const size_t bo = (A.flags & PGM_FLAG_BOUNDS_PER_LC) ? (size_t)b * P : 0;
    const double lb = A.con_lb[bo + tid], ub = A.con_ub[bo + tid];
    double th = rv, jc = 1.0;
    if (kd == 1) {
      th = softplus_d(rv) + lb;
      jc = sigmoid_d(rv);
    } else if (kd == 2) {
      const double s = sigmoid_d(rv);
      th = lb + (ub - lb) * s;
      jc = (ub - lb) * s * (1.0 - s);
    }
    theta[tid] = th;
    jac[tid] = jc;
  }
  __syncthreads();
  if (tid < QT) wq[tid] = (tid < Q) ? theta[1 + tid] : 0.0;
  if (tid < QT * D) {
    const int q = tid / D, dd = tid - q * D;
    const double sg = (q < Q) ? theta[1 + Q + Q * D + q * D + dd] : 0.0;
    aq[q * D + dd] = 2.0 * M_PI * M_PI * sg * sg;  // indexed [q*D + dd]
  }
  const double mean = theta[0];
  const double lnoise = learn_noise ? theta[P - 1] : 0.0;
  // ---- per-point fields into the block's scratch ------------------------------------
  const double* xb = A.x + (size_t)b * A.n_max * D;
  const double* yb = A.y + (size_t)b * A.n_max;
  const double* fnb = A.fixed_noise ? A.fixed_noise + (size_t)b * A.n_max : nullptr;
  for (int i = tid; i < npad; i += NTHREADS) {
    const bool valid = i < n;
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      const double xc = valid ? (xb[(size_t)i * D + dd] - xb[dd]) : 0.0;
      sc.fld[(size_t)dd * npad + i] = xc;
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        double sn = 0.0, cs = 1.0;
        if (valid && q < Q) sincospi(2.0 * theta[1 + Q + q * D + dd] * xc, &sn, &cs);
        sc.fld[(size_t)(D + (dd * QT + q) * 2) * npad + i] = cs;
        sc.fld[(size_t)(D + (dd * QT + q) * 2 + 1) * npad + i] = sn;
      }
    }
    sc.rhs[i] = valid ? (yb[i] - mean) : 0.0;
    sc.dn[i] = valid ? ((fnb ? fnb[i] : 0.0) + lnoise) : 0.0;
  }
  __syncthreads();

  auto load_side = [&](double* vec, int I, int nf) {
    for (int idx = tid; idx < nf * TS; idx += NTHREADS) {
      const int f = idx >> 6, r = idx & 63;
      vec[idx] = sc.fld[(size_t)f * npad + I * TS + r];
    }
  };

  const int g = lane >> 2, tq = lane & 3, wm = warp >> 2, wn = warp & 3;
  double acc[4][2][2];
  double iq_part = 0.0;  // partial of z^T z
  int info = 0;

  // ================= phase P: Cholesky + forward solve, with the jitter ladder ========
  for (int attempt = 0; attempt <= 3; ++attempt) {
    double jitter = 0.0;
    if (attempt > 0) {
      jitter = 1e-8;
      for (int t = 1; t < attempt; ++t) jitter *= 10.0;
    }
    if (tid == 0) { *s_fail = 0; *s_logdet = 0.0; }
    iq_part = 0.0;
    bool failed = false;
    for (int j = 0; j < N && !failed; ++j) {
      for (int i = j; i < N; ++i) {
        zero_acc(acc);
        gemm_tiles<false, false>(
            acc, j, [&](int k) { return sc.tiles + (size_t)tri(i, k) * TT; },
            [&](int k) { return sc.tiles + (size_t)tri(j, k) * TT; }, stages);
        // rowv/colv are free here: every reader passed the barrier inside gemm_tiles
        load_side(rowv, i, C::NFB);
        load_side(colv, j, C::NFB);
        __syncthreads();
        // epilogue: C = Ktilde_ij - acc
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int r = wm * 32 + mi * 8 + g, c = wn * 16 + ni * 8 + tq * 2 + e;
              const int gi = i * TS + r, gj = j * TS + c;
              double kv = 0.0;
              if (gi < n && gj < n) kv = k_entry<KIND, QT, D>(rowv, colv, r, c, wq, aq);
              if (gi == gj) kv = (gi < n) ? (kv + sc.dn[gi] + jitter) : 1.0;
              acc[mi][ni][e] = kv - acc[mi][ni][e];
            }
        if (i == j) {
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int r = wm * 32 + mi * 8 + g, c = wn * 16 + ni * 8 + tq * 2 + e;
                S[r * LD_S + c] = acc[mi][ni][e];
              }
          __syncthreads();
          potrf_inv_64(S, S2, s_fail, s_logdet);
          if (*s_fail) { failed = true; break; }
          // X_jj -> tile(j,j) (explicit zeros above the diagonal)
          double* dt = sc.tiles + (size_t)tri(j, j) * TT;
          for (int idx = tid; idx < TT / 2; idx += NTHREADS) {
            const int r = idx >> 5, c2 = (idx & 31) * 2;
            double2 v;
            v.x = (c2 <= r) ? S2[r * LD_S + c2] : 0.0;
            v.y = (c2 + 1 <= r) ? S2[r * LD_S + c2 + 1] : 0.0;
            *reinterpret_cast<double2*>(dt + r * TS + c2) = v;
          }
          // forward solve: z_j = X_jj (rhs_j - sum_{k<j} L_jk z_k)   (4 lanes per row)
          {
            const int r = tid >> 2, l4 = tid & 3;
            double u = 0.0;
            for (int k = 0; k < j; ++k) {
              const double* Lt = sc.tiles + (size_t)tri(j, k) * TT + r * TS;
              const double* zk = sc.z + k * TS;
#pragma unroll 4
              for (int c = l4; c < TS; c += 4) u += Lt[c] * zk[c];
            }
            u += shfl_xor_d(u, 1);
            u += shfl_xor_d(u, 2);
            if (l4 == 0) tmpv[r] = sc.rhs[j * TS + r] - u;
            __syncthreads();
            double zz = 0.0;
            for (int c = l4; c <= r; c += 4) zz += S2[r * LD_S + c] * tmpv[c];
            zz += shfl_xor_d(zz, 1);
            zz += shfl_xor_d(zz, 2);
            if (l4 == 0) {
              sc.z[j * TS + r] = zz;
              iq_part += zz * zz;
            }
          }
          __syncthreads();
        } else {
          // L_ij = C * X_jj^T  via the tile engine (C staged through the block's scratch)
          store_acc_tile(acc, sc.ctmp, 1.0);
          zero_acc(acc);
          gemm_tiles<false, false>(
              acc, 1, [&](int) { return sc.ctmp; },
              [&](int) { return sc.tiles + (size_t)tri(j, j) * TT; }, stages);
          store_acc_tile(acc, sc.tiles + (size_t)tri(i, j) * TT, 1.0);
        }
      }
    }
    __syncthreads();
    const int fl = *s_fail;
    if (!fl) { info = attempt; break; }
    if (fl & 2) { info = -1; break; }
    info = -2;
    __syncthreads();
  }

  // ---- MLL ---------------------------------------------------------------------------
  {
    double v1[1] = {iq_part};
    block_reduce<1>(v1, red, fin);
  }
  const double inv_quad = fin[0];
  const double logdet = *s_logdet;
  const double mll = (info >= 0)
                         ? -0.5 * (inv_quad + logdet + (double)n * 1.8378770664093454836) / n
                         : nan("");
  __syncthreads();
  if (tid == 0) *mll_out = mll;
  if (!want_grad) return info;
  if (info < 0) {
    if (tid < P) grad_out[tid] = nan("");
    return info;
  }

  // ================= phase T: X = L^-1 (in place, column by column) ===================
  for (int j = 0; j < N - 1; ++j) {
    for (int i = j + 1; i < N; ++i) {
      zero_acc(acc);
      gemm_tiles<false, true>(
          acc, i - j, [&](int kk) { return sc.tiles + (size_t)tri(i, j + kk) * TT; },
          [&](int kk) { return sc.tiles + (size_t)tri(j + kk, j) * TT; }, stages);
      store_acc_tile(acc, sc.ctmp, 1.0);
      zero_acc(acc);
      gemm_tiles<false, true>(
          acc, 1, [&](int) { return sc.tiles + (size_t)tri(i, i) * TT; },
          [&](int) { return sc.ctmp; }, stages);
      store_acc_tile(acc, sc.tiles + (size_t)tri(i, j) * TT, -1.0);
    }
  }
  __syncthreads();
  // ---- alpha = X^T z ------------------------------------------------------------------
  double* alpha = sc.fld + (size_t)C::NFB * npad;
  {
    const int c = tid & 63, rg = tid >> 6;
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int i = j; i < N; ++i) {
        const double* Xt = sc.tiles + (size_t)tri(i, j) * TT;
        const double* zi = sc.z + i * TS;
#pragma unroll 4
        for (int r = rg; r < TS; r += 4) s += Xt[r * TS + c] * zi[r];
      }
      tmpv[rg * TS + c] = s;
      __syncthreads();
      if (tid < TS)
        alpha[j * TS + tid] = tmpv[tid] + tmpv[TS + tid] + tmpv[2 * TS + tid] + tmpv[3 * TS + tid];
      __syncthreads();
    }
  }
  // ================= phase G: K^-1 tiles -> gradient contraction =======================
  double ga[C::NG];
#pragma unroll
  for (int t = 0; t < C::NG; ++t) ga[t] = 0.0;
  double trW = 0.0;
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j <= i; ++j) {
      zero_acc(acc);
      gemm_tiles<true, true>(
          acc, N - i, [&](int kk) { return sc.tiles + (size_t)tri(i + kk, i) * TT; },
          [&](int kk) { return sc.tiles + (size_t)tri(i + kk, j) * TT; }, stages);
      load_side(rowv, i, C::NF);
      load_side(colv, j, C::NF);
      __syncthreads();
      const double* al_r = rowv + C::NFB * TS;
      const double* al_c = colv + C::NFB * TS;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = wm * 32 + mi * 8 + g, c = wn * 16 + ni * 8 + tq * 2 + e;
            const int gi = i * TS + r, gj = j * TS + c;
            if (gi < n && gj <= gi) {
              const double W = al_r[r] * al_c[c] - acc[mi][ni][e];
              if (gi == gj) trW += W;
              const double wgt = (gi == gj) ? W : 2.0 * W;
              k_grad_entry<KIND, QT, D>(rowv, colv, r, c, wq, aq, wgt, ga);
            }
          }
    }
  }
  // ---- reduce, apply constants and the constraint Jacobian ----------------------------
  {
    double v[C::NV + 1];
#pragma unroll
    for (int t = 0; t < C::NG; ++t) v[t] = ga[t];
    v[C::NG] = trW;
    double sa = 0.0;
    for (int i2 = tid; i2 < n; i2 += NTHREADS) sa += alpha[i2];
    v[C::NV] = sa;
    __syncthreads();
    block_reduce<C::NV + 1>(v, red, fin);
  }
  if (tid < P) {
    const double half = 0.5 / (double)n;
    double gv;
    if (tid == 0) {
      gv = fin[C::NV] / (double)n;
    } else if (tid < 1 + Q) {
      gv = half * fin[tid - 1];
    } else if (tid < 1 + Q + Q * D) {
      const int t = tid - 1 - Q, q = t / D, dd = t - q * D;
      gv = half * (-2.0 * M_PI * wq[q]) * fin[QT + q * D + dd];
    } else if (tid < 1 + Q + 2 * Q * D) {
      const int t = tid - 1 - Q - Q * D, q = t / D, dd = t - q * D;
      gv = half * (-4.0 * M_PI * M_PI * theta[tid] * wq[q]) * fin[QT + QT * D + q * D + dd];
    } else {
      gv = half * fin[C::NG];
    }
    grad_out[tid] = gv * jac[tid];
  }
  __syncthreads();
  return info;
}

template <int KIND, int QT, int D>
__device__ __forceinline__ Scratch make_scratch(double* base, int n_max) {
  using C = Cfg<KIND, QT, D>;
  const int N = (n_max + TS - 1) / TS;
  const size_t npad = (size_t)N * TS;
  Scratch sc;
  sc.tiles = base;
  sc.ctmp = sc.tiles + (size_t)tri(N, 0) * TT;
  sc.fld = sc.ctmp + TT;
  sc.rhs = sc.fld + (size_t)C::NF * npad;
  sc.z = sc.rhs + npad;
  sc.dn = sc.z + npad;
  return sc;
}

// ------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------
template <int KIND, int QT, int D>
__global__ void __launch_bounds__(NTHREADS, (Cfg<KIND, QT, D>::SMEM_BYTES <= 113 * 1024) ? 2 : 1)
    sm_mll_grad_kernel(EvalArgs A) {
  extern __shared__ __align__(16) double sm[];
  // NOTE: field layout of a light curve in scratch depends on ITS npad, so make_scratch is
  // sized for n_max and eval_lightcurve only uses the leading part.
  Scratch sc = make_scratch<KIND, QT, D>(A.ws + (size_t)blockIdx.x * A.ws_per_block, A.n_max);
  const int P = 1 + A.Q + 2 * A.Q * D + ((A.flags & PGM_FLAG_LEARN_NOISE) ? 1 : 0);
  for (int b = blockIdx.x; b < A.B; b += gridDim.x) {
    double* gout = A.grad ? A.grad + (size_t)b * P : nullptr;
    const int info =
        eval_lightcurve<KIND, QT, D>(A, b, A.raw + (size_t)b * P, sm, sc, A.mll + b, gout);
    if (threadIdx.x == 0) A.info[b] = info;
  }
}

// dense K + D for parity tests / large-n path: one block per (light curve, 64x64 tile)
template <int KIND, int QT, int D>
__global__ void __launch_bounds__(NTHREADS)
    sm_kernel_dense_kernel(EvalArgs A, double* __restrict__ Kout) {
  using C = Cfg<KIND, QT, D>;
  __shared__ double rowv[C::NFB * TS];
  __shared__ double colv[C::NFB * TS];
  __shared__ double theta[C::PMAX], wq[QT], aq[QT * D];
  const int tid = threadIdx.x;
  const int b = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
  const int Q = A.Q;
  const bool learn_noise = (A.flags & PGM_FLAG_LEARN_NOISE) != 0;
  const int P = 1 + Q + 2 * Q * D + (learn_noise ? 1 : 0);
  const int n = A.n_valid ? A.n_valid[b] : A.n_max;
  if (ti * TS >= n || tj * TS >= n) return;
  if (tid < P) {
    const double rv = A.raw[(size_t)b * P + tid];
    const int kd = A.con_kind[tid];
    const size_t bo = (A.flags & PGM_FLAG_BOUNDS_PER_LC) ? (size_t)b * P : 0;
    const double lb = A.con_lb[bo + tid], ub = A.con_ub[bo + tid];
    double th = rv;
    if (kd == 1) th = softplus_d(rv) + lb;
    else if (kd == 2) th = lb + (ub - lb) * sigmoid_d(rv);
    theta[tid] = th;
  }
  __syncthreads();
  if (tid < QT) wq[tid] = (tid < Q) ? theta[1 + tid] : 0.0;
  if (tid < QT * D) {
    const int q = tid / D, dd = tid - q * D;
    const double sg = (q < Q) ? theta[1 + Q + Q * D + q * D + dd] : 0.0;
    aq[q * D + dd] = 2.0 * M_PI * M_PI * sg * sg;
  }
  const double* xb = A.x + (size_t)b * A.n_max * D;
  for (int idx = tid; idx < 2 * TS; idx += NTHREADS) {
    const int side = idx >> 6, r = idx & 63;
    const int gi = (side ? tj : ti) * TS + r;
    double* vec = side ? colv : rowv;
    const bool valid = gi < n;
    for (int dd = 0; dd < D; ++dd) {
      const double xc = valid ? (xb[(size_t)gi * D + dd] - xb[dd]) : 0.0;
      vec[dd * TS + r] = xc;
      for (int q = 0; q < QT; ++q) {
        double sn = 0.0, cs = 1.0;
        if (valid && q < Q) sincospi(2.0 * theta[1 + Q + q * D + dd] * xc, &sn, &cs);
        vec[(D + (dd * QT + q) * 2) * TS + r] = cs;
        vec[(D + (dd * QT + q) * 2 + 1) * TS + r] = sn;
      }
    }
  }
  __syncthreads();
  const double lnoise = learn_noise ? theta[P - 1] : 0.0;
  const double* fnb = A.fixed_noise ? A.fixed_noise + (size_t)b * A.n_max : nullptr;
  double* Kb = Kout + (size_t)b * A.n_max * A.n_max;
  for (int idx = tid; idx < TT; idx += NTHREADS) {
    const int r = idx >> 6, c = idx & 63;
    const int gi = ti * TS + r, gj = tj * TS + c;
    if (gi < n && gj < n) {
      double kv = k_entry<KIND, QT, D>(rowv, colv, r, c, wq, aq);
      if (gi == gj) kv += (fnb ? fnb[gi] : 0.0) + lnoise;
      Kb[(size_t)gi * A.n_max + gj] = kv;
    }
  }
}

// torch.optim.{SGD,Adam,AdamW} step on one parameter (trainers.py:141-147 defaults; A.7).
// g is d loss / d raw.
__device__ __forceinline__ double optim_update(double p, double g, double& m, double& v,
                                               int kind, double lr, double b1, double b2,
                                               double eps, double wd, int step) {
  if (kind == 0) return p - lr * g;
  if (kind == 2) p *= (1.0 - lr * wd);
  else if (wd != 0.0) g += wd * p;
  m = b1 * m + (1.0 - b1) * g;
  v = b2 * v + (1.0 - b2) * g * g;
  const double bc1 = 1.0 - pow(b1, (double)step);
  const double bc2 = 1.0 - pow(b2, (double)step);
  const double denom = sqrt(v) / sqrt(bc2) + eps;
  return p - (lr / bc1) * (m / denom);
}

// whole training loop of trainers.py:177-207 per light curve, on device
template <int KIND, int QT, int D>
__global__ void __launch_bounds__(NTHREADS, (Cfg<KIND, QT, D>::SMEM_BYTES <= 113 * 1024) ? 2 : 1)
    sm_fit_kernel(FitArgs F) {
  using C = Cfg<KIND, QT, D>;
  extern __shared__ __align__(16) double sm[];
  const EvalArgs& A = F.e;
  Scratch sc = make_scratch<KIND, QT, D>(A.ws + (size_t)blockIdx.x * A.ws_per_block, A.n_max);
  const int P = 1 + A.Q + 2 * A.Q * D + ((A.flags & PGM_FLAG_LEARN_NOISE) ? 1 : 0);
  double* par = sm + C::SM_PAR;
  double* s_raw = par + C::PAR_RAW;
  double* s_m = s_raw + C::PMAX;
  double* s_v = s_m + C::PMAX;
  double* s_grad = par + C::PAR_TMP + 2 * TS;  // tmpv upper half is free between evals
  double* s_mll = par + C::PAR_FIN + C::NV + 2;
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < A.B; b += gridDim.x) {
    __syncthreads();
    if (tid < P) {
      s_raw[tid] = F.raw_io[(size_t)b * P + tid];
      s_m[tid] = 0.0;
      s_v[tid] = 0.0;
      if (F.raw_hist) F.raw_hist[(size_t)b * P + tid] = s_raw[tid];
    }
    int it = 0, info = 0;
    for (; it < F.maxiter; ++it) {
      __syncthreads();
      info = eval_lightcurve<KIND, QT, D>(A, b, s_raw, sm, sc, s_mll, s_grad);
      __syncthreads();
      const double loss = -(*s_mll);
      if (tid == 0) F.loss_hist[(size_t)it * A.B + b] = loss;
      if (info < 0) { ++it; break; }
      if (tid < P) {
        s_raw[tid] = optim_update(s_raw[tid], -s_grad[tid], s_m[tid], s_v[tid], F.optim_kind,
                                  F.lr, F.beta1, F.beta2, F.eps, F.weight_decay, it + 1);
        if (F.raw_hist) F.raw_hist[((size_t)(it + 1) * A.B + b) * P + tid] = s_raw[tid];
      }
      // early stop: stop and i > miniter and std(loss[-stopavg:]) < stop  (np.std, ddof 0)
      if (F.stop > 0.0 && it > F.miniter) {
        const int cnt = min(F.stopavg, it + 1);
        double s1 = 0.0, s2 = 0.0;
        for (int t = 0; t < cnt; ++t) s1 += (t == 0) ? loss : F.loss_hist[(size_t)(it - t) * A.B + b];
        const double mu = s1 / cnt;
        for (int t = 0; t < cnt; ++t) {
          const double lv = (t == 0) ? loss : F.loss_hist[(size_t)(it - t) * A.B + b];
          s2 += (lv - mu) * (lv - mu);
        }
        if (sqrt(s2 / cnt) < F.stop) { ++it; break; }
      }
    }
    __syncthreads();
    if (tid < P) F.raw_io[(size_t)b * P + tid] = s_raw[tid];
    if (tid == 0) { F.n_iter[b] = it; A.info[b] = info; }
    // mark the unused tail of the history
    for (int t = it + tid; t < F.maxiter; t += NTHREADS) F.loss_hist[(size_t)t * A.B + b] = nan("");
  }
}

}  // namespace pgm
