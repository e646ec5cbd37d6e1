// C ABI of pgmuvi_b200 (see include/pgmuvi_b200.h).  Host-side dispatch only; all device
// code is in gp_fused.cuh.
#include "gp_fused.cuh"
#include "gp_large.cuh"
#include "gp_large_tc.cuh"
#include "lombscargle.cuh"
#include "../../include/pgmuvi_b200.h"

#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <string>

namespace pgm {
// launchers are explicitly instantiated in inst.cu (one translation unit per config)
template <int KIND, int QT, int D> int launch_eval(const EvalArgs& A0, cudaStream_t st);
template <int KIND, int QT, int D> int launch_fit(const FitArgs& F, cudaStream_t st);
template <int KIND, int QT, int D> int launch_dense(const EvalArgs& A, double* K, cudaStream_t st);
template <int KIND, int QT, int D>
int launch_large(const LargeArgs& A, int want_grad, cudaStream_t st, int predict_only);
template <int KIND, int QT, int D> int launch_predict(const PredictArgs& PA, cudaStream_t st);

thread_local std::string g_err;
int fail(const std::string& m) {
  g_err = m;
  return -1;
}
int cuda_fail(const char* what, cudaError_t e) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return -2;
}
int device_sms() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}
}  // namespace pgm

namespace pgm {
__global__ void optim_step_kernel(double* raw, const double* grad_mll, double* m, double* v,
                                  const int32_t* active, int B, int P, int kind, double lr,
                                  double b1, double b2, double eps, double wd, double bc1,
                                  double bc2s) {
  // HBM-bound: 4 reads + 3 writes of 8 bytes per parameter; the bias corrections (two pow per
  // step) come from the host so that no transcendental is left in the kernel
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * P) return;
  if (active && !active[idx / P]) return;
  double mm = m ? m[idx] : 0.0, vv = v ? v[idx] : 0.0;
  raw[idx] = optim_update_bc(raw[idx], -grad_mll[idx], mm, vv, kind, lr, b1, b2, eps, wd, bc1, bc2s);
  if (m) m[idx] = mm;
  if (v) v[idx] = vv;
}


// fp32 <-> fp64 staging of the _f32 entry points (fp32 storage, fp64 arithmetic)
__global__ void widen_kernel(const float* __restrict__ src, double* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (double)src[i];
}
__global__ void narrow_kernel(const double* __restrict__ src, float* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}
}  // namespace pgm

namespace {
using pgm::cuda_fail;
using pgm::fail;
using pgm::launch_dense;
using pgm::launch_eval;
using pgm::launch_fit;
using pgm::launch_large;
using pgm::launch_predict;

int pad_q(int Q) { return Q <= 1 ? 1 : Q <= 2 ? 2 : Q <= 4 ? 4 : 8; }

// NF of the instantiated config serving (d, Q): sizes the workspace (upper bound over kinds)
int nf_for(int d, int Q) { return d + 2 * d * pad_q(Q <= 4 ? 4 : Q) + 1; }  // Q = 0 -> QT = 4

#define PGM_DISPATCH_Q(KIND, D, FN, ...)                     \
  switch (pad_q(Q)) {                                        \
    case 1: return FN<KIND, 1, D>(__VA_ARGS__);              \
    case 2: return FN<KIND, 2, D>(__VA_ARGS__);              \
    case 4: return FN<KIND, 4, D>(__VA_ARGS__);              \
    default: return FN<KIND, 8, D>(__VA_ARGS__);             \
  }

#define PGM_DISPATCH_Q48(KIND, D, FN, ...)                   \
  if (Q <= 4) return FN<KIND, 4, D>(__VA_ARGS__);            \
  return FN<KIND, 8, D>(__VA_ARGS__);

#define PGM_DISPATCH(FN, ...)                                                        \
  do {                                                                               \
    if (kernel_kind == PGM_KIND_SM1D) {                                              \
      PGM_DISPATCH_Q(PGM_KIND_SM1D, 1, FN, __VA_ARGS__)                              \
    } else if (kernel_kind == PGM_KIND_SM_ARD_PRODSUM) {                             \
      PGM_DISPATCH_Q(PGM_KIND_SM_ARD_PRODSUM, 2, FN, __VA_ARGS__)                    \
    } else if (kernel_kind == PGM_KIND_SM_ARD_SUMPROD) {                             \
      PGM_DISPATCH_Q(PGM_KIND_SM_ARD_SUMPROD, 2, FN, __VA_ARGS__)                    \
    } else if (kernel_kind == PGM_KIND_SEP_RBF) {                                    \
      PGM_DISPATCH_Q48(PGM_KIND_SEP_RBF, 2, FN, __VA_ARGS__)                         \
    } else if (kernel_kind == PGM_KIND_SEP_MATERN15) {                               \
      PGM_DISPATCH_Q48(PGM_KIND_SEP_MATERN15, 2, FN, __VA_ARGS__)                    \
    } else if (kernel_kind == PGM_KIND_SEP_RQ) {                                     \
      PGM_DISPATCH_Q48(PGM_KIND_SEP_RQ, 2, FN, __VA_ARGS__)                          \
    } else if (kernel_kind == PGM_KIND_SEP_CONST) {                                  \
      PGM_DISPATCH_Q48(PGM_KIND_SEP_CONST, 2, FN, __VA_ARGS__)                       \
    } else {                                                                         \
      switch (kernel_kind) {                                                         \
        case PGM_KIND_STAT(0, 0): return FN<PGM_KIND_STAT(0, 0), 4, 1>(__VA_ARGS__); \
        case PGM_KIND_STAT(0, 1): return FN<PGM_KIND_STAT(0, 1), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(0, 2): return FN<PGM_KIND_STAT(0, 2), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(0, 3): return FN<PGM_KIND_STAT(0, 3), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(0, 4): return FN<PGM_KIND_STAT(0, 4), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(1, 0): return FN<PGM_KIND_STAT(1, 0), 4, 1>(__VA_ARGS__); \
        case PGM_KIND_STAT(1, 1): return FN<PGM_KIND_STAT(1, 1), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(1, 2): return FN<PGM_KIND_STAT(1, 2), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(1, 3): return FN<PGM_KIND_STAT(1, 3), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(1, 4): return FN<PGM_KIND_STAT(1, 4), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(2, 0): return FN<PGM_KIND_STAT(2, 0), 4, 1>(__VA_ARGS__); \
        case PGM_KIND_STAT(2, 1): return FN<PGM_KIND_STAT(2, 1), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(2, 2): return FN<PGM_KIND_STAT(2, 2), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(2, 3): return FN<PGM_KIND_STAT(2, 3), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(2, 4): return FN<PGM_KIND_STAT(2, 4), 4, 2>(__VA_ARGS__); \
        case PGM_KIND_STAT(3, 0): return FN<PGM_KIND_STAT(3, 0), 4, 1>(__VA_ARGS__); \
        case PGM_KIND_STAT(4, 0): return FN<PGM_KIND_STAT(4, 0), 4, 1>(__VA_ARGS__); \
        default: return FN<PGM_KIND_STAT(5, 0), 4, 1>(__VA_ARGS__);                  \
      }                                                                              \
    }                                                                                \
  } while (0)

int check_common(int B, int n_max, int d, int Q, int kernel_kind) {
  if (B < 0 || n_max < 1) return fail("B must be >= 0 and n_max >= 1");
  if (kernel_kind >= PGM_KIND_STAT_BASE && kernel_kind <= PGM_KIND_STAT(5, 0)) {
    if (Q != 0) return fail("stationary kinds have no mixtures: pass Q == 0");
    const int wk = (kernel_kind - PGM_KIND_STAT_BASE) % 5;
    if ((kernel_kind - PGM_KIND_STAT_BASE) / 5 >= 3 && wk != 0)
      return fail("time kernels 3..5 exist for 1-D models only (wk = 0)");
    if (d != (wk == 0 ? 1 : 2)) return fail("stationary kind / d mismatch (d = 1 without, 2 with a wavelength kernel)");
    return 0;
  }
  if (Q < 1 || Q > 8) return fail("Q (num_mixtures) must be in 1..8");
  if (kernel_kind == PGM_KIND_SM1D) {
    if (d != 1) return fail("PGM_KIND_SM1D needs d == 1");
  } else if (kernel_kind == PGM_KIND_SM_ARD_PRODSUM || kernel_kind == PGM_KIND_SM_ARD_SUMPROD) {
    if (d != 2) return fail("ARD spectral-mixture kinds need d == 2");
  } else if (kernel_kind >= PGM_KIND_SEP_RBF && kernel_kind <= PGM_KIND_SEP_CONST) {
    if (d != 2) return fail("separable kinds need d == 2 (time, wavelength)");
  } else {
    return fail("unknown kernel_kind");
  }
  return 0;
}


// number of packed raw parameters of a model (layout in include/pgmuvi_b200.h)
int host_param_count(int d, int Q, int kernel_kind, int flags) {
  const bool sep = kernel_kind >= PGM_KIND_SEP_RBF;
  const int ds = sep ? 1 : d;
  const int nl = (kernel_kind >= PGM_KIND_STAT_BASE)
                     ? pgm::stat_num_time(kernel_kind) + pgm::sep_num_lam(pgm::stat_wave_atom(kernel_kind))
                     : pgm::sep_num_lam(kernel_kind);
  return 1 + Q + 2 * Q * ds + ((flags & PGM_FLAG_LEARN_NOISE) ? 1 : 0) + nl;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// fp64 staging area of the _f32 entry points, carved from the tail of the workspace
struct Stage32 {
  double *x, *y, *fn, *raw, *lb, *ub, *mll, *grad, *loss, *hist;
  size_t bytes;
};
Stage32 carve32(char* base, int B, int n_max, int d, int P, bool bounds_per_lc, int maxiter,
                bool want_hist) {
  Stage32 s;
  size_t off = 0;
  auto take = [&](size_t elems) {
    double* p = base ? reinterpret_cast<double*>(base + off) : nullptr;
    off += align256(elems * sizeof(double));
    return p;
  };
  const size_t Bn = (size_t)B * n_max, BP = (size_t)B * P;
  s.x = take(Bn * d); s.y = take(Bn); s.fn = take(Bn); s.raw = take(BP);
  s.lb = take(bounds_per_lc ? BP : (size_t)P); s.ub = take(bounds_per_lc ? BP : (size_t)P);
  s.mll = take((size_t)B); s.grad = take(BP);
  s.loss = take((size_t)maxiter * B);
  s.hist = take(want_hist ? (size_t)(maxiter + 1) * BP : 0);
  s.bytes = off;
  return s;
}
int widen(const float* src, double* dst, size_t n, cudaStream_t st) {
  if (!n) return 0;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  pgm::widen_kernel<<<blocks, 256, 0, st>>>(src, dst, n);
  return 0;
}
int narrow(const double* src, float* dst, size_t n, cudaStream_t st) {
  if (!n) return 0;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  pgm::narrow_kernel<<<blocks, 256, 0, st>>>(src, dst, n);
  return 0;
}

template <int KIND, int QT, int D>
int fused_blocks_per_sm() { return pgm::Cfg<KIND, QT, D>::F_BYTES <= 113 * 1024 ? 2 : 1; }
int fused_occ_dispatch(int d, int Q, int kernel_kind) {
  PGM_DISPATCH(fused_blocks_per_sm);
  return 1;
}

}  // namespace

extern "C" {

int pgm_version(void) { return 101; }
const char* pgm_last_error(void) { return pgm::g_err.c_str(); }

size_t pgm_workspace_bytes(int elem_size, int n_max, int d, int Q, int device) {
  (void)elem_size;
  int sms = 148;
  if (device >= 0) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  else { int dv = 0; if (cudaGetDevice(&dv) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dv); }
  if (sms <= 0) sms = 148;
  const size_t per_block = pgm::scratch_elems(n_max, nf_for(d, Q));
  // up to 2 resident blocks per SM, + the light-curve ticket counters of the fused kernels
  return per_block * sizeof(double) * (size_t)sms * 2 + pgm::PGM_SCHED_INTS * sizeof(int);
}

int pgm_sm_mll_grad_f64(const double* x, const int32_t* n_valid, const double* y,
                        const double* fixed_noise, const double* raw, const int32_t* con_kind,
                        const double* con_lb, const double* con_ub, int B, int n_max, int d,
                        int Q, int kernel_kind, int flags, double* mll, double* grad_raw,
                        int32_t* info, void* workspace, size_t workspace_bytes, void* stream) {
  return pgm_sm_mll_grad_alpha_f64(x, n_valid, y, fixed_noise, raw, con_kind, con_lb, con_ub, B,
                                   n_max, d, Q, kernel_kind, flags, mll, grad_raw, nullptr, info,
                                   workspace, workspace_bytes, stream);
}

int pgm_sm_mll_grad_alpha_f64(const double* x, const int32_t* n_valid, const double* y,
                              const double* fixed_noise, const double* raw,
                              const int32_t* con_kind, const double* con_lb, const double* con_ub,
                              int B, int n_max, int d, int Q, int kernel_kind, int flags,
                              double* mll, double* grad_raw, double* alpha_out, int32_t* info,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (int r = check_common(B, n_max, d, Q, kernel_kind)) return r;
  if (alpha_out && !(flags & PGM_FLAG_GRAD)) return fail("alpha_out needs PGM_FLAG_GRAD");
  if (B == 0) return 0;
  if (!x || !y || !raw || !con_kind || !con_lb || !con_ub || !mll || !info || !workspace)
    return fail("null pointer argument");
  if ((flags & PGM_FLAG_GRAD) && !grad_raw) return fail("PGM_FLAG_GRAD needs grad_raw");
  const size_t per_block = pgm::scratch_elems(n_max, nf_for(d, Q));
  if (workspace_bytes < pgm_workspace_bytes(8, n_max, d, Q, -1))
    return fail("workspace too small (see pgm_workspace_bytes)");
  pgm::EvalArgs A;
  A.x = x; A.n_valid = n_valid; A.y = y; A.fixed_noise = fixed_noise; A.raw = raw;
  A.con_kind = con_kind; A.con_lb = con_lb; A.con_ub = con_ub;
  A.B = B; A.n_max = n_max; A.Q = Q; A.flags = flags;
  A.mll = mll; A.grad = grad_raw; A.info = info;
  A.ws = static_cast<double*>(workspace); A.ws_per_block = per_block;
  A.alpha_out = alpha_out;
  A.sched = reinterpret_cast<int*>(static_cast<char*>(workspace) + pgm_workspace_bytes(8, n_max, d, Q, -1)
                                   - pgm::PGM_SCHED_INTS * sizeof(int));
  A.sms = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PGM_DISPATCH(launch_eval, A, st);
  return 0;
}

int pgm_fused_grid(int d, int Q, int kernel_kind) {
  if (check_common(1, 64, d, Q, kernel_kind)) return -1;
  return pgm::device_sms() * fused_occ_dispatch(d, Q, kernel_kind);
}

size_t pgm_staged_workspace_bytes(int n_max, int B) {
  return (n_max < 1 || B < 1) ? 0 : pgm::large_ws_bytes(n_max, B);
}

int pgm_sm_mll_grad_staged_f64(const double* x, const int32_t* n_valid, const double* y,
                               const double* fixed_noise, const double* raw,
                               const int32_t* con_kind, const double* con_lb,
                               const double* con_ub, int B, int n_max, int d, int Q,
                               int kernel_kind, int flags, double* mll, double* grad_raw,
                               int32_t* info, void* workspace, size_t workspace_bytes,
                               void* stream) {
  return pgm_sm_mll_grad_staged_alpha_f64(x, n_valid, y, fixed_noise, raw, con_kind, con_lb,
                                          con_ub, B, n_max, d, Q, kernel_kind, flags, mll,
                                          grad_raw, nullptr, info, workspace, workspace_bytes,
                                          stream);
}

int pgm_sm_mll_grad_staged_alpha_f64(const double* x, const int32_t* n_valid, const double* y,
                                     const double* fixed_noise, const double* raw,
                                     const int32_t* con_kind, const double* con_lb,
                                     const double* con_ub, int B, int n_max, int d, int Q,
                                     int kernel_kind, int flags, double* mll, double* grad_raw,
                                     double* alpha_out, int32_t* info, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (int r = check_common(B, n_max, d, Q, kernel_kind)) return r;
  if (alpha_out && !(flags & PGM_FLAG_GRAD)) return fail("alpha_out needs PGM_FLAG_GRAD");
  if (B == 0) return 0;
  if (!x || !y || !raw || !con_kind || !con_lb || !con_ub || !mll || !info || !workspace)
    return fail("null pointer argument");
  if ((flags & PGM_FLAG_GRAD) && !grad_raw) return fail("PGM_FLAG_GRAD needs grad_raw");
  if (workspace_bytes < ((flags & PGM_FLAG_TF32X3) ? pgm_staged_tf32x3_workspace_bytes(n_max, B)
                                                   : pgm_staged_workspace_bytes(n_max, B)))
    return fail("workspace too small (see pgm_staged_workspace_bytes / pgm_staged_tf32x3_workspace_bytes)");
  pgm::LargeArgs A;
  A.x = x; A.n_valid = n_valid; A.y = y; A.fixed_noise = fixed_noise; A.raw = raw;
  A.con_kind = con_kind; A.con_lb = con_lb; A.con_ub = con_ub;
  A.B = B; A.n_max = n_max; A.Q = Q; A.flags = flags;
  A.mll = mll; A.grad = grad_raw; A.info = info;
  A.ws = static_cast<double*>(workspace);
  A.alpha_out = alpha_out;
  const int want_grad = (flags & PGM_FLAG_GRAD) ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PGM_DISPATCH(launch_large, A, want_grad, st, 0);
  return 0;
}

size_t pgm_staged_tf32x3_workspace_bytes(int n_max, int B) {
  return (n_max < 1 || B < 1) ? 0 : pgm::large_ws_bytes_tc(n_max, B);
}

int pgm_sm_mll_grad_staged_tf32x3_f64(const double* x, const int32_t* n_valid, const double* y,
                                      const double* fixed_noise, const double* raw,
                                      const int32_t* con_kind, const double* con_lb,
                                      const double* con_ub, int B, int n_max, int d, int Q,
                                      int kernel_kind, int flags, double* mll, double* grad_raw,
                                      int32_t* info, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  return pgm_sm_mll_grad_staged_alpha_f64(x, n_valid, y, fixed_noise, raw, con_kind, con_lb,
                                          con_ub, B, n_max, d, Q, kernel_kind,
                                          flags | PGM_FLAG_TF32X3, mll, grad_raw, nullptr, info,
                                          workspace, workspace_bytes, stream);
}

int pgm_sm_mll_grad_tf32x3_f32(const float* x, const int32_t* n_valid, const float* y,
                               const float* fixed_noise, const float* raw,
                               const int32_t* con_kind, const float* con_lb, const float* con_ub,
                               int B, int n_max, int d, int Q, int kernel_kind, int flags,
                               float* mll, float* grad_raw, int32_t* info, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (int r = check_common(B, n_max, d, Q, kernel_kind)) return r;
  if (B == 0) return 0;
  if (!x || !y || !raw || !con_kind || !con_lb || !con_ub || !mll || !info || !workspace)
    return fail("null pointer argument");
  if ((flags & PGM_FLAG_GRAD) && !grad_raw) return fail("PGM_FLAG_GRAD needs grad_raw");
  const int P = host_param_count(d, Q, kernel_kind, flags);
  const bool per_lc = (flags & PGM_FLAG_BOUNDS_PER_LC) != 0;
  const size_t base = align256(pgm_staged_tf32x3_workspace_bytes(n_max, B));
  const size_t need = base + pgm_f32_staging_bytes(B, n_max, d, Q, kernel_kind, flags, 0, 0);
  if (workspace_bytes < need)
    return fail("workspace too small (pgm_staged_tf32x3_workspace_bytes rounded up to 256 + "
                "pgm_f32_staging_bytes)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Stage32 s = carve32(static_cast<char*>(workspace) + base, B, n_max, d, P, per_lc, 0, false);
  const size_t Bn = (size_t)B * n_max, BP = (size_t)B * P;
  widen(x, s.x, Bn * d, st); widen(y, s.y, Bn, st);
  if (fixed_noise) widen(fixed_noise, s.fn, Bn, st);
  widen(raw, s.raw, BP, st);
  widen(con_lb, s.lb, per_lc ? BP : (size_t)P, st);
  widen(con_ub, s.ub, per_lc ? BP : (size_t)P, st);
  if (int r = pgm_sm_mll_grad_staged_alpha_f64(
          s.x, n_valid, s.y, fixed_noise ? s.fn : nullptr, s.raw, con_kind, s.lb, s.ub, B, n_max, d,
          Q, kernel_kind, flags | PGM_FLAG_JITTER_F32 | PGM_FLAG_TF32X3 | PGM_FLAG_TF32X3_CHOL, s.mll,
          (flags & PGM_FLAG_GRAD) ? s.grad : nullptr, nullptr, info, workspace, base, stream))
    return r;
  narrow(s.mll, mll, (size_t)B, st);
  if (flags & PGM_FLAG_GRAD) narrow(s.grad, grad_raw, BP, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("f32 staging", e);
  return 0;
}

size_t pgm_predict_workspace_bytes(int n_max, int B, int device) {
  if (n_max < 1 || B < 1) return 0;
  int sms = 148;
  if (device >= 0) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  else sms = pgm::device_sms();
  if (sms <= 0) sms = 148;
  const size_t N = (n_max + pgm::TS - 1) / pgm::TS;
  size_t staged = (pgm::large_ws_bytes(n_max, B) + 255) & ~(size_t)255;
  return staged + (size_t)2 * sms * N * pgm::TT * sizeof(double);
}

int pgm_sm_predict_f64(const double* x, const int32_t* n_valid, const double* y,
                       const double* fixed_noise, const double* raw, const int32_t* con_kind,
                       const double* con_lb, const double* con_ub, int B, int n_max, int d, int Q,
                       int kernel_kind, int flags, const double* xstar, int m, double* mean,
                       double* var, int32_t* info, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (int r = check_common(B, n_max, d, Q, kernel_kind)) return r;
  if (B == 0 || m == 0) return 0;
  if (m < 0) return fail("m must be >= 0");
  if (!x || !y || !raw || !con_kind || !con_lb || !con_ub || !xstar || !mean || !var || !info ||
      !workspace)
    return fail("null pointer argument");
  if (workspace_bytes < pgm_predict_workspace_bytes(n_max, B, -1))
    return fail("workspace too small (see pgm_predict_workspace_bytes)");
  pgm::PredictArgs PA;
  pgm::LargeArgs& A = PA.a;
  A.x = x; A.n_valid = n_valid; A.y = y; A.fixed_noise = fixed_noise; A.raw = raw;
  A.con_kind = con_kind; A.con_lb = con_lb; A.con_ub = con_ub;
  A.B = B; A.n_max = n_max; A.Q = Q; A.flags = flags & ~PGM_FLAG_GRAD;
  A.mll = nullptr; A.grad = nullptr; A.info = info;
  A.ws = static_cast<double*>(workspace);
  A.alpha_out = nullptr;
  PA.xstar = xstar; PA.m = m; PA.mean = mean; PA.var = var;
  const size_t staged = (pgm::large_ws_bytes(n_max, B) + 255) & ~(size_t)255;
  PA.kscratch = reinterpret_cast<double*>(static_cast<char*>(workspace) + staged);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PGM_DISPATCH(launch_predict, PA, st);
  return 0;
}

int pgm_sm_kernel_dense_f64(const double* x, const int32_t* n_valid, const double* fixed_noise,
                            const double* raw, const int32_t* con_kind, const double* con_lb,
                            const double* con_ub, int B, int n_max, int d, int Q,
                            int kernel_kind, int flags, double* K_out, void* stream) {
  if (int r = check_common(B, n_max, d, Q, kernel_kind)) return r;
  if (B == 0) return 0;
  if (!x || !raw || !con_kind || !con_lb || !con_ub || !K_out) return fail("null pointer argument");
  if (B > 65535) return fail("B > 65535 not supported by the dense builder");
  pgm::EvalArgs A;
  memset(&A, 0, sizeof(A));
  A.x = x; A.n_valid = n_valid; A.fixed_noise = fixed_noise; A.raw = raw;
  A.con_kind = con_kind; A.con_lb = con_lb; A.con_ub = con_ub;
  A.B = B; A.n_max = n_max; A.Q = Q; A.flags = flags;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PGM_DISPATCH(launch_dense, A, K_out, st);
  return 0;
}

int pgm_optim_step_f64(double* raw, const double* grad_mll, double* exp_avg, double* exp_avg_sq,
                       const int32_t* active, int B, int P, int optim_kind, double lr,
                       double beta1, double beta2, double eps, double weight_decay, int step,
                       void* stream) {
  if (B < 0 || P < 1) return fail("bad B / P");
  if (B == 0) return 0;
  if (!raw || !grad_mll) return fail("null pointer argument");
  if (optim_kind != PGM_OPT_SGD && (!exp_avg || !exp_avg_sq))
    return fail("Adam / AdamW need exp_avg and exp_avg_sq");
  if (optim_kind < 0 || optim_kind > 2) return fail("unknown optim_kind");
  if (step < 1) return fail("step counts from 1");
  const size_t total = (size_t)B * P;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  // torch.optim computes the bias corrections in host double precision as well (1 - beta ** step)
  const double bc1 = 1.0 - std::pow(beta1, (double)step);
  const double bc2s = std::sqrt(1.0 - std::pow(beta2, (double)step));
  pgm::optim_step_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      raw, grad_mll, exp_avg, exp_avg_sq, active, B, P, optim_kind, lr, beta1, beta2, eps,
      weight_decay, bc1, bc2s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("optim_step_kernel launch", e);
  return 0;
}

int pgm_sm_fit_f64(const double* x, const int32_t* n_valid, const double* y,
                   const double* fixed_noise, double* raw, const int32_t* con_kind,
                   const double* con_lb, const double* con_ub, int B, int n_max, int d, int Q,
                   int kernel_kind, int flags, int optim_kind, double lr, double beta1,
                   double beta2, double eps, double weight_decay, int maxiter, int miniter,
                   double stop, int stopavg, double* loss_hist, double* raw_hist,
                   int32_t* n_iter, int32_t* info, double* opt_state, void* workspace,
                   size_t workspace_bytes, void* stream) {
  (void)opt_state;  // optimiser state lives in shared memory for the whole loop
  if (int r = check_common(B, n_max, d, Q, kernel_kind)) return r;
  if (B == 0) return 0;
  if (!x || !y || !raw || !con_kind || !con_lb || !con_ub || !loss_hist || !n_iter || !info ||
      !workspace)
    return fail("null pointer argument");
  if (maxiter < 1) return fail("maxiter must be >= 1");
  if (optim_kind < 0 || optim_kind > 2) return fail("unknown optim_kind");
  if (stopavg < 1) stopavg = 1;
  const size_t per_block = pgm::scratch_elems(n_max, nf_for(d, Q));
  if (workspace_bytes < pgm_workspace_bytes(8, n_max, d, Q, -1))
    return fail("workspace too small (see pgm_workspace_bytes)");
  pgm::FitArgs F;
  pgm::EvalArgs& A = F.e;
  A.x = x; A.n_valid = n_valid; A.y = y; A.fixed_noise = fixed_noise; A.raw = raw;
  A.con_kind = con_kind; A.con_lb = con_lb; A.con_ub = con_ub;
  A.B = B; A.n_max = n_max; A.Q = Q; A.flags = flags | PGM_FLAG_GRAD;
  A.mll = nullptr; A.grad = nullptr; A.info = info;
  A.ws = static_cast<double*>(workspace); A.ws_per_block = per_block;
  A.alpha_out = nullptr;
  A.sched = reinterpret_cast<int*>(static_cast<char*>(workspace) + pgm_workspace_bytes(8, n_max, d, Q, -1)
                                   - pgm::PGM_SCHED_INTS * sizeof(int));
  A.sms = 0;
  F.raw_io = raw; F.optim_kind = optim_kind; F.lr = lr; F.beta1 = beta1; F.beta2 = beta2;
  F.eps = eps; F.weight_decay = weight_decay; F.stop = stop; F.maxiter = maxiter;
  F.miniter = miniter; F.stopavg = stopavg; F.loss_hist = loss_hist; F.raw_hist = raw_hist;
  F.n_iter = n_iter;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PGM_DISPATCH(launch_fit, F, st);
  return 0;
}

size_t pgm_f32_staging_bytes(int B, int n_max, int d, int Q, int kernel_kind, int flags,
                             int maxiter, int want_raw_hist) {
  if (B < 1 || n_max < 1) return 0;
  const int P = host_param_count(d, Q, kernel_kind, flags);
  return carve32(nullptr, B, n_max, d, P, (flags & PGM_FLAG_BOUNDS_PER_LC) != 0,
                 maxiter < 0 ? 0 : maxiter, want_raw_hist != 0).bytes;
}

int pgm_sm_mll_grad_f32(const float* x, const int32_t* n_valid, const float* y,
                        const float* fixed_noise, const float* raw, const int32_t* con_kind,
                        const float* con_lb, const float* con_ub, int B, int n_max, int d, int Q,
                        int kernel_kind, int flags, float* mll, float* grad_raw, int32_t* info,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (int r = check_common(B, n_max, d, Q, kernel_kind)) return r;
  if (B == 0) return 0;
  if (!x || !y || !raw || !con_kind || !con_lb || !con_ub || !mll || !info || !workspace)
    return fail("null pointer argument");
  if ((flags & PGM_FLAG_GRAD) && !grad_raw) return fail("PGM_FLAG_GRAD needs grad_raw");
  const int P = host_param_count(d, Q, kernel_kind, flags);
  const bool per_lc = (flags & PGM_FLAG_BOUNDS_PER_LC) != 0;
  const size_t base = align256(pgm_workspace_bytes(8, n_max, d, Q, -1));
  const size_t need = base + pgm_f32_staging_bytes(B, n_max, d, Q, kernel_kind, flags, 0, 0);
  if (workspace_bytes < need)
    return fail("workspace too small (pgm_workspace_bytes rounded up to 256 + pgm_f32_staging_bytes)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Stage32 s = carve32(static_cast<char*>(workspace) + base, B, n_max, d, P, per_lc, 0, false);
  const size_t Bn = (size_t)B * n_max, BP = (size_t)B * P;
  widen(x, s.x, Bn * d, st); widen(y, s.y, Bn, st);
  if (fixed_noise) widen(fixed_noise, s.fn, Bn, st);
  widen(raw, s.raw, BP, st);
  widen(con_lb, s.lb, per_lc ? BP : (size_t)P, st);
  widen(con_ub, s.ub, per_lc ? BP : (size_t)P, st);
  if (int r = pgm_sm_mll_grad_f64(s.x, n_valid, s.y, fixed_noise ? s.fn : nullptr, s.raw, con_kind,
                                  s.lb, s.ub, B, n_max, d, Q, kernel_kind,
                                  flags | PGM_FLAG_JITTER_F32, s.mll,
                                  (flags & PGM_FLAG_GRAD) ? s.grad : nullptr, info, workspace,
                                  base, stream))
    return r;
  narrow(s.mll, mll, (size_t)B, st);
  if (flags & PGM_FLAG_GRAD) narrow(s.grad, grad_raw, BP, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("f32 staging", e);
  return 0;
}

int pgm_sm_fit_f32(const float* x, const int32_t* n_valid, const float* y,
                   const float* fixed_noise, float* raw, const int32_t* con_kind,
                   const float* con_lb, const float* con_ub, int B, int n_max, int d, int Q,
                   int kernel_kind, int flags, int optim_kind, double lr, double beta1,
                   double beta2, double eps, double weight_decay, int maxiter, int miniter,
                   double stop, int stopavg, float* loss_hist, float* raw_hist, int32_t* n_iter,
                   int32_t* info, void* workspace, size_t workspace_bytes, void* stream) {
  if (int r = check_common(B, n_max, d, Q, kernel_kind)) return r;
  if (B == 0) return 0;
  if (!x || !y || !raw || !con_kind || !con_lb || !con_ub || !loss_hist || !n_iter || !info ||
      !workspace)
    return fail("null pointer argument");
  if (maxiter < 1) return fail("maxiter must be >= 1");
  const int P = host_param_count(d, Q, kernel_kind, flags);
  const bool per_lc = (flags & PGM_FLAG_BOUNDS_PER_LC) != 0;
  const size_t base = align256(pgm_workspace_bytes(8, n_max, d, Q, -1));
  const size_t need = base + pgm_f32_staging_bytes(B, n_max, d, Q, kernel_kind, flags, maxiter,
                                                   raw_hist != nullptr);
  if (workspace_bytes < need)
    return fail("workspace too small (pgm_workspace_bytes rounded up to 256 + pgm_f32_staging_bytes)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Stage32 s = carve32(static_cast<char*>(workspace) + base, B, n_max, d, P, per_lc, maxiter,
                      raw_hist != nullptr);
  const size_t Bn = (size_t)B * n_max, BP = (size_t)B * P;
  widen(x, s.x, Bn * d, st); widen(y, s.y, Bn, st);
  if (fixed_noise) widen(fixed_noise, s.fn, Bn, st);
  widen(raw, s.raw, BP, st);
  widen(con_lb, s.lb, per_lc ? BP : (size_t)P, st);
  widen(con_ub, s.ub, per_lc ? BP : (size_t)P, st);
  if (int r = pgm_sm_fit_f64(s.x, n_valid, s.y, fixed_noise ? s.fn : nullptr, s.raw, con_kind, s.lb,
                             s.ub, B, n_max, d, Q, kernel_kind, flags | PGM_FLAG_JITTER_F32,
                             optim_kind, lr, beta1, beta2,
                             eps, weight_decay, maxiter, miniter, stop, stopavg, s.loss,
                             raw_hist ? s.hist : nullptr, n_iter, info, nullptr, workspace, base,
                             stream))
    return r;
  narrow(s.raw, raw, BP, st);
  narrow(s.loss, loss_hist, (size_t)maxiter * B, st);
  if (raw_hist) narrow(s.hist, raw_hist, (size_t)(maxiter + 1) * BP, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("f32 staging", e);
  return 0;
}

int pgm_lombscargle_f64(const double* t, const int32_t* n_valid, const double* y, const double* dy,
                        int B, int n_max, const double* f0, const double* df, const int32_t* nf,
                        int nf_max, int flags, double* power, void* stream) {
  if (B < 0 || n_max < 1 || nf_max < 1) return fail("B >= 0, n_max >= 1, nf_max >= 1 required");
  if (B == 0) return 0;
  if (B > 65535) return fail("B > 65535 (split the batch)");
  if (!t || !y || !f0 || !df || !nf || !power) return fail("null pointer argument");
  pgm::LsArgs A;
  A.t = t; A.n_valid = n_valid; A.y = y; A.dy = dy; A.f0 = f0; A.df = df; A.nf = nf;
  A.B = B; A.n_max = n_max; A.nf_max = nf_max; A.flags = flags; A.power = power;
  dim3 grid((nf_max + pgm::LS_FPB - 1) / pgm::LS_FPB, B);
  pgm::ls_power_kernel<<<grid, pgm::LS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(A);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("ls_power_kernel launch", e);
  return 0;
}

int pgm_ls_peaks_f64(const double* power, const int32_t* nf, int B, int nf_max, int distance,
                     int num_peaks, int32_t* peak_idx, double* peak_power, void* scratch,
                     size_t scratch_bytes, void* stream) {
  if (B < 0 || nf_max < 1 || num_peaks < 1) return fail("bad B / nf_max / num_peaks");
  if (B == 0) return 0;
  if (distance < 1) return fail("distance must be >= 1 (scipy.signal.find_peaks)");
  if (!power || !nf || !peak_idx || !peak_power || !scratch) return fail("null pointer argument");
  if (scratch_bytes < (size_t)B * nf_max) return fail("scratch too small (B * nf_max bytes)");
  pgm::LsPeakArgs A;
  A.power = power; A.nf = nf; A.B = B; A.nf_max = nf_max; A.distance = distance;
  A.num_peaks = num_peaks; A.mask = static_cast<uint8_t*>(scratch); A.peak_idx = peak_idx;
  A.peak_power = peak_power;
  pgm::ls_peaks_kernel<<<B, pgm::LS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(A);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("ls_peaks_kernel launch", e);
  return 0;
}

int pgm_sm_psd_peak_f64(const double* freq, const double* fscale, const double* weight,
                        const double* fmin, const double* fmax, int B, int Q, int n_grid,
                        double* grid, double* psd, int32_t* dom_idx, double* dom_freq,
                        double* dom_height, int32_t* n_peaks, void* stream) {
  if (B < 0 || Q < 1 || Q > 8 || n_grid < 3) return fail("bad B / Q (1..8) / n_grid (>= 3)");
  if (B == 0) return 0;
  if (!freq || !fscale || !weight || !fmin || !fmax || !psd || !dom_idx || !dom_freq ||
      !dom_height || !n_peaks)
    return fail("null pointer argument");
  pgm::PsdArgs A;
  A.freq = freq; A.fscale = fscale; A.weight = weight; A.fmin = fmin; A.fmax = fmax;
  A.B = B; A.Q = Q; A.n_grid = n_grid; A.grid = grid; A.psd = psd; A.dom_idx = dom_idx;
  A.dom_freq = dom_freq; A.dom_height = dom_height; A.n_peaks = n_peaks;
  pgm::sm_psd_peak_kernel<<<B, pgm::LS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(A);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("sm_psd_peak_kernel launch", e);
  return 0;
}

}  // extern "C"

// ---- FP64 yardsticks ---------------------------------------------------------------------
namespace pgm {
__global__ void __launch_bounds__(256) probe_dmma(int iters, double* out) {
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) mma_f64(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) probe_dfma(int iters, double* out) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}
// DMMA and DFMA interleaved (equal flops each): tells whether they share one FP64 pipe
__global__ void __launch_bounds__(256) probe_mixed(int iters, double* out) {
  double acc[4][2], f[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-9 + i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  const double fa = 1.0000001, fb = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mma_f64(acc[i], a, b);                        // 512 flops / warp
#pragma unroll
      for (int j = 0; j < 8; ++j) f[(i * 8 + j) & 15] = fma(f[(i * 8 + j) & 15], fa, fb);  // 8 x 64
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += acc[i][0] + acc[i][1];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) probe_ffma(int iters, double* out) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9f + i;
  const float a = 1.0000001f, b = 1e-9f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}
// dense TF32 tcgen05.mma rate (M = N = 128, K = 8 per instruction): one CTA per SM, one thread issues
// `iters` rounds of 16 MMAs (4 k-steps x 2 accumulators x A/B image pairs) from a resident stage
__global__ void __launch_bounds__(64, 1) probe_tcgen05_tf32(int iters, double* out) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  const unsigned base = (smem_u32(smraw) + 1023u) & ~1023u;
  const unsigned bar = base + 2 * tc::IMG_BYTES, tslot = bar + 16;
  float* img = reinterpret_cast<float*>(smraw + (base - smem_u32(smraw)));
  for (int i = threadIdx.x; i < 2 * tc::IMG_FLOATS; i += 64) img[i] = 1e-3f * (float)((i * 37) % 101);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
  }
  if ((threadIdx.x >> 5) == 1) tc::tmem_alloc(tslot, 256);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  unsigned tmem;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem) : "r"(tslot));
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = tc::idesc_tf32(128, 128);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t a = tc::smem_desc_sw128(base + ks * 32);
        const uint64_t b = tc::smem_desc_sw128(base + tc::IMG_BYTES + ks * 32);
        tc::mma_tf32(tmem, a, b, idesc, (it | ks) ? 1u : 0u);
        tc::mma_tf32(tmem + 128, b, a, idesc, (it | ks) ? 1u : 0u);
        tc::mma_tf32(tmem, b, a, idesc, 1u);
        tc::mma_tf32(tmem + 128, a, b, idesc, 1u);
      }
    tc::mma_commit(bar);
    mbar_wait(bar, 0);
    tc::fence_after_sync();
  }
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    float v[16];
    tc::tmem_ld16(tmem + ((unsigned)(32 * 1) << 16), v);
    if (v[0] == 123.456f) out[0] = v[0];
    tc::fence_before_sync();
    tc::tmem_dealloc(tmem, 256);
  }
}
}  // namespace pgm

extern "C" int pgm_peak_probe(int kind, int iters, double* tflops_host, void* stream) {
  if (!tflops_host || iters < 1) return fail("bad arguments");
  const int sms = pgm::device_sms();
  double* dout = nullptr;
  cudaError_t e = cudaMalloc(&dout, 8);
  if (e != cudaSuccess) return cuda_fail("cudaMalloc", e);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int blocks = sms * 8, threads = 256;
  double flops_per_thread_iter = 0.0;
  float best = 1e30f;
  const size_t tc_smem = 2 * pgm::tc::IMG_BYTES + 1024 + 64;
  if (kind == 4) {
    blocks = sms; threads = 64;
    cudaFuncSetAttribute(pgm::probe_tcgen05_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem);
  }
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, st);
    if (kind == 4) {
      pgm::probe_tcgen05_tf32<<<blocks, threads, tc_smem, st>>>(iters, dout);
      flops_per_thread_iter = 16.0 * (128.0 * 128 * 8 * 2) / 64.0;
    } else if (kind == 0) {
      pgm::probe_dmma<<<blocks, threads, 0, st>>>(iters, dout);
      flops_per_thread_iter = 8.0 * (8 * 8 * 4 * 2) / 32.0;
    } else if (kind == 1) {
      pgm::probe_dfma<<<blocks, threads, 0, st>>>(iters, dout);
      flops_per_thread_iter = 16.0 * 2;
    } else if (kind == 3) {
      pgm::probe_mixed<<<blocks, threads, 0, st>>>(iters, dout);
      flops_per_thread_iter = 4.0 * (8 * 8 * 4 * 2) / 32.0 + 32.0 * 2;
    } else {
      pgm::probe_ffma<<<blocks, threads, 0, st>>>(iters, dout);
      flops_per_thread_iter = 16.0 * 2;
    }
    cudaEventRecord(e1, st);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaFree(dout); return cuda_fail("probe", e); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dout);
  const double total = flops_per_thread_iter * (double)iters * (double)blocks * threads;
  *tflops_host = total / (best * 1e-3) / 1e12;
  return 0;
}
