// Host launchers of the fused kernels; explicitly instantiated per (KIND, QT, D) in inst.cu so
// the heavy kernels compile in parallel translation units.
#pragma once
#include "gp_fused.cuh"
#include "gp_large.cuh"
#include "gp_large_tc.cuh"
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>
#include <cstdio>

namespace pgm {
int fail(const std::string& m);
int cuda_fail(const char* what, cudaError_t e);
int device_sms();
inline int predict_blocks() { return 2 * device_sms(); }   // persistent blocks of lg_predict

template <int KIND, int QT, int D>
int launch_eval(const pgm::EvalArgs& A0, cudaStream_t st) {
  using C = pgm::Cfg<KIND, QT, D>;
  pgm::EvalArgs A = A0;
  auto kern = pgm::sm_mll_grad_kernel<KIND, QT, D>;
  size_t smem = C::F_BYTES;
  if (const char* f = getenv("PGM_DEBUG_SMEM_KB")) smem = std::max(smem, (size_t)atoi(f) * 1024);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return cuda_fail("cudaFuncSetAttribute", e);
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, pgm::NTHREADS, smem);
  if (e != cudaSuccess) return cuda_fail("occupancy", e);
  if (occ < 1) return fail("kernel does not fit on an SM");
  int grid = device_sms() * occ;
  if (grid > A.B) grid = A.B;
  A.sms = device_sms();
  if (A.sched && (occ != 2 || grid != 2 * A.sms || getenv("PGM_STATIC_STRIDE"))) A.sched = nullptr;
  if (A.sched) cudaMemsetAsync(A.sched, 0, PGM_SCHED_INTS * sizeof(int), st);
#ifdef PGM_DEBUG_HOOKS
  {
    int dbg = 0;
    if (const char* f = getenv("PGM_DEBUG_MODE")) dbg = (int)strtol(f, nullptr, 0);
    cudaMemcpyToSymbolAsync(pgm::c_dbg, &dbg, sizeof(int), 0, cudaMemcpyHostToDevice, st);
  }
  long long* dprof = nullptr;
  if (getenv("PGM_DEBUG_PROF")) {
    cudaMalloc(&dprof, (size_t)grid * 16 * sizeof(long long));
    cudaMemsetAsync(dprof, 0, (size_t)grid * 16 * sizeof(long long), st);
  }
  cudaMemcpyToSymbolAsync(pgm::c_prof, &dprof, sizeof(dprof), 0, cudaMemcpyHostToDevice, st);
#endif
  kern<<<grid, pgm::NTHREADS, smem, st>>>(A);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("sm_mll_grad_kernel launch", e);
#ifdef PGM_DEBUG_HOOKS
  if (dprof) {
    cudaStreamSynchronize(st);
    std::vector<long long> h((size_t)grid * 16);
    cudaMemcpy(h.data(), dprof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dprof);
    static const char* nm[13] = {"setup", "P.gemm", "P.kepi", "P.potrf", "P.diagpost", "P.trsm", "mll",
                                 "T.gemm1", "T.gemm2", "alpha", "G.gemm", "G.epi", "final"};
    double tot = 0, v[13];
    for (int s2 = 0; s2 < 13; ++s2) { v[s2] = 0; for (int b = 0; b < grid; ++b) v[s2] += (double)h[(size_t)b * 16 + s2]; v[s2] /= grid; tot += v[s2]; }
    fprintf(stderr, "[pgm prof] per-CTA cycles (avg over %d CTAs), total %.0f:", grid, tot);
    for (int s2 = 0; s2 < 13; ++s2) fprintf(stderr, " %s %.0f (%.1f%%)", nm[s2], v[s2], 100 * v[s2] / tot);
    fprintf(stderr, "\n");
  }
#endif
  return 0;
}

template <int KIND, int QT, int D>
int launch_fit(const pgm::FitArgs& F, cudaStream_t st) {
  using C = pgm::Cfg<KIND, QT, D>;
  auto kern = pgm::sm_fit_kernel<KIND, QT, D>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C::F_BYTES);
  if (e != cudaSuccess) return cuda_fail("cudaFuncSetAttribute", e);
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, pgm::NTHREADS, C::F_BYTES);
  if (e != cudaSuccess) return cuda_fail("occupancy", e);
  if (occ < 1) return fail("kernel does not fit on an SM");
  int grid = device_sms() * occ;
  if (grid > F.e.B) grid = F.e.B;
  pgm::FitArgs F2 = F;
  F2.e.sms = device_sms();
  if (F2.e.sched && (occ != 2 || grid != 2 * F2.e.sms || getenv("PGM_STATIC_STRIDE"))) F2.e.sched = nullptr;
  if (F2.e.sched) cudaMemsetAsync(F2.e.sched, 0, PGM_SCHED_INTS * sizeof(int), st);
  kern<<<grid, pgm::NTHREADS, C::F_BYTES, st>>>(F2);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("sm_fit_kernel launch", e);
  return 0;
}

template <int KIND, int QT, int D>
int launch_dense(const pgm::EvalArgs& A, double* K, cudaStream_t st) {
  const int N = (A.n_max + pgm::TS - 1) / pgm::TS;
  if (A.B > 65535) return fail("sm_kernel_dense: at most 65535 light curves per call");
  auto kern = pgm::sm_kernel_dense_kernel<KIND, QT, D>;
  const size_t smem = pgm::DenseSmem<KIND, QT, D>::BYTES;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(sm_kernel_dense_kernel)", e);
  dim3 grid(N * (N + 1) / 2, A.B);
  kern<<<grid, pgm::NTHREADS, smem, st>>>(A, K);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("sm_kernel_dense_kernel launch", e);
  return 0;
}


// Staged engine (gp_large.cuh): B light curves advance stage by stage, one launch per
// dependency stage with B x (tiles of the stage) blocks; B = 1 is the single large GP.
// Host-orchestrated on `st`; synchronises once per Cholesky pass to read how many light curves
// must repeat it with more jitter (psd_safe_cholesky's ladder) - the call is blocking.
// pieces per level-1 fold of the tcgen05 drain (gp_large_tc.cuh::tc_drain)
inline int tc_fold() {
  int f = 4;
  if (const char* e = getenv("PGM_TC_FOLD")) f = std::max(1, atoi(e));
  return f;
}

template <int KIND, int QT, int D>
int launch_large(const pgm::LargeArgs& A, int want_grad, cudaStream_t st, int predict_only = 0) {
  using C = pgm::Cfg<KIND, QT, D>;
  using namespace pgm;
  const int n = A.n_max, N = (n + TS - 1) / TS, npad = N * TS, B = A.B;
  if (B > 65535) return fail("staged engine: B > 65535 (split the batch)");
  // tile columns per panel (one right-looking trailing update every NB columns): wider panels
  // amortise the C-tile round trip of the trailing update once there are enough tile rows to
  // keep the in-panel launches busy (B200, C4: NB 8 -> 32 takes the P phase from 497 to 416 ms)
  int NB = (N <= 64) ? 8 : (N <= 256) ? 16 : 32;
  if (const char* f = getenv("PGM_STAGED_NB")) NB = std::max(1, atoi(f));
  auto k_upd = lg_update<KIND, QT, D>;
  auto k_grad = lg_grad<KIND, QT, D>;
  auto k_chol = lg_chol_all<KIND, QT, D>;
  cudaError_t e;
  constexpr size_t UPD_SMEM = UpdSmem<KIND, QT, D>::BYTES;
  e = cudaFuncSetAttribute(k_upd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM);
  if (e != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(lg_update)", e);
  constexpr size_t CHOL_SMEM = CholAllSmem<KIND, QT, D>::BYTES;
  cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_SMEM);
  constexpr size_t GRAD_SMEM = GradSmem<KIND, QT, D>::BYTES;
  e = cudaFuncSetAttribute(k_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAD_SMEM);
  if (e != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(lg_grad)", e);
  cudaFuncSetAttribute(lg_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LG_DIAG_SMEM);
  cudaFuncSetAttribute(lg_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LG_TRSM_SMEM);
  cudaFuncSetAttribute(lg_inv_row, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LG_INV_SMEM);
  cudaFuncSetAttribute(lg_inv_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LG_INV_SMEM);
  BatchState bs = make_batch_state(A.ws, n, B);
  cudaMemsetAsync(bs.state, 0, (size_t)(3 * B + 4) * sizeof(int), st);
  const dim3 blk(NTHREADS);
  for (int pass = 0; pass <= 3; ++pass) {
    cudaMemsetAsync(bs.count, 0, 4 * sizeof(int), st);   // [0] repeat count, [1] / [2] job tickets
    lg_setup<KIND, QT, D><<<dim3((npad + NTHREADS - 1) / NTHREADS, B), blk, 0, st>>>(A);
    const long long ncb = (long long)(N * (N + 1) / 2) * B;
    // up to PGM_STAGED_CHOL_ALL_N tile rows (default 200) the left-looking one-launch schedule
    // wins (no launch latencies, potrf overlapped: n=2048 1.62 -> 1.16 ms, C3 11.8 -> 8.6 ms); the
    // two meet at N = 256 (63.1 vs 64.4 ms) and beyond the panel scheme's operand reuse in L2
    // wins (C4: 425 vs 500 ms)
    int all_n = 200;
    if (const char* f = getenv("PGM_STAGED_CHOL_ALL_N")) all_n = atoi(f);
    // ... and only while a dependency stage cannot fill the device by itself: big batches have
    // all the parallelism they need in one launch per stage (C5: 52.8 vs 62.9 ms)
    const bool few = (long long)B * N < 1024;
    const bool one_launch = !getenv("PGM_STAGED_ROWWISE") && few && N <= all_n && ncb <= 2147483647LL;
    if (one_launch) {   // dataflow Cholesky: the whole pass in one flag-ordered launch
      cudaMemsetAsync(bs.tflag, 0, (size_t)B * large_ntri(n) * sizeof(int), st);
      k_chol<<<dim3((unsigned)ncb), blk, CHOL_SMEM, st>>>(A);
    }
    for (int J0 = 0; J0 < N && !one_launch; J0 += NB) {
      const int J1 = std::min(J0 + NB, N);
      const int build = (J0 == 0) ? 1 : 0;
      for (int j = J0; j < J1; ++j) {
        if (build || j > J0)
          k_upd<<<dim3(N - j, B), blk, UPD_SMEM, st>>>(A, 0, j, J0, j, build);
        lg_diag<<<dim3(1, B), blk, LG_DIAG_SMEM, st>>>(A, j);
        if (j + 1 < N) lg_trsm<<<dim3(N - j - 1, B), blk, LG_TRSM_SMEM, st>>>(A, j);
      }
      if (J1 < N && (A.flags & PGM_FLAG_TF32X3_CHOL) && !(J1 & 1)) {
        // trailing update on tcgen05 (3xTF32): pack the panel's L tiles, one CTA per 128x128 tile
        using TS_ = TcGradSmem<KIND, QT, D>;
        auto k_utc = lg_update_tc<KIND, QT, D>;
        int nst = 3;
        while (nst > 1 && TS_::bytes(nst) > 227 * 1024) --nst;
        if (TS_::bytes(nst) > 227 * 1024) return fail("lg_update_tc does not fit in shared memory");
        e = cudaFuncSetAttribute(k_utc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_::bytes(nst));
        if (e != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(lg_update_tc)", e);
        int halfp = 1;
        if (const char* f = getenv("PGM_TC_HALF_PIECES")) halfp = atoi(f) ? 1 : 0;
        const int NT = (N + 1) / 2, Mt = NT - J1 / 2;
        lg_pack_panel_tf32<<<dim3(J1 - J0, 2 * NT - J1, B), blk, 0, st>>>(A, J0, J1);
        k_utc<<<dim3(tc_grid(Mt), B), TC_THREADS, TS_::bytes(nst), st>>>(A, J0, J1, build, nst, halfp, tc_fold());
      } else if (J1 < N) {
        const int M = N - J1;
        k_upd<<<dim3(M * (M + 1) / 2, B), blk, UPD_SMEM, st>>>(A, 1, J1, J0, J1, build);
      }
    }
    lg_ladder<<<(B + 255) / 256, 256, 0, st>>>(A);
    if (A.flags & PGM_FLAG_NOSYNC) {
      // no host round trip: all four ladder passes are enqueued; for a light curve that is already
      // factored (or failed) every kernel of a later pass returns at once
      if (!one_launch) return fail("PGM_FLAG_NOSYNC needs the one-launch schedule (B * N < 1024, N <= 200)");
      continue;
    }
    int again = 0;
    e = cudaMemcpyAsync(&again, bs.count, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail("staged engine, Cholesky phase", e);
    if (!again) break;
  }
  if (want_grad || predict_only) {
    if (getenv("PGM_STAGED_ROWWISE") || (long long)B * N >= 1024) {   // one launch per tile row
      for (int i = 1; i < N; ++i) lg_inv_row<<<dim3(i, B), blk, LG_INV_SMEM, st>>>(A, i);
    } else if (N > 1) {                        // whole T phase in one launch, flag-ordered
      cudaMemsetAsync(bs.tflag, 0, (size_t)B * large_ntri(n) * sizeof(int), st);
      cudaMemsetAsync(bs.count + 2, 0, sizeof(int), st);
      const long long nblk = (long long)(N * (N - 1) / 2) * B;
      if (nblk > 2147483647LL) return fail("staged engine: too many tiles x light curves");
      lg_inv_all<<<dim3((unsigned)nblk), blk, LG_INV_SMEM, st>>>(A);
    }
    lg_alpha<<<dim3(N, B), blk, 0, st>>>(A);
    if (!predict_only && (A.flags & PGM_FLAG_TF32X3)) {
      // G phase on tcgen05: X^T -> packed TF32 hi / lo operand images, then one CTA per 128x128
      // tile of K~^-1 (3xTF32 products in tensor memory, FP64 contraction epilogue)
      const int NT = (N + 1) / 2;
      using TS_ = TcGradSmem<KIND, QT, D>;
      auto k_tc = lg_grad_tc<KIND, QT, D>;
      int nst = 3;
      if (const char* f = getenv("PGM_TC_STAGES")) nst = std::min(3, std::max(1, atoi(f)));
      while (nst > 1 && TS_::bytes(nst) > 227 * 1024) --nst;
      if (TS_::bytes(nst) > 227 * 1024) return fail("lg_grad_tc does not fit in shared memory");
      e = cudaFuncSetAttribute(k_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_::bytes(nst));
      if (e != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(lg_grad_tc)", e);
      lg_pack_tf32<<<dim3(N, 2 * NT, B), blk, 0, st>>>(A);
      // pieces of 16 k (two truncating hi hi additions each) by default: accuracy first.  32-k pieces
      // (PGM_TC_HALF_PIECES=0) halve the drain work - C4's G phase 170 -> 112 ms - at 1.8x the error.
      int halfp = 1;
      if (const char* f = getenv("PGM_TC_HALF_PIECES")) halfp = atoi(f) ? 1 : 0;
      k_tc<<<dim3(tc_grid(NT), B), TC_THREADS, TS_::bytes(nst), st>>>(A, nst, halfp, tc_fold());
    } else if (!predict_only) {
      k_grad<<<dim3(N * (N + 1) / 2, B), blk, GRAD_SMEM, st>>>(A);
    }
  }
  if (predict_only) {
    lg_transpose<<<dim3(N * (N + 1) / 2, B), blk, 0, st>>>(A);
  } else {
    lg_finish<KIND, QT, D><<<dim3(1, B), blk, 0, st>>>(A, want_grad);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("staged engine launch", e);
  return 0;
}

// N1: posterior prediction = P + T phases of the staged engine, then lg_predict (persistent
// blocks over (light curve, test tile) jobs, each with its own K* panel scratch).
template <int KIND, int QT, int D>
int launch_predict(const pgm::PredictArgs& PA0, cudaStream_t st) {
  using C = pgm::Cfg<KIND, QT, D>;
  using namespace pgm;
  PredictArgs PA = PA0;
  if (int r = launch_large<KIND, QT, D>(PA.a, 0, st, 1)) return r;
  auto kern = lg_predict<KIND, QT, D>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C::SMEM_BYTES);
  if (e != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(lg_predict)", e);
  const int Mt = (PA.m + TS - 1) / TS;
  int grid = predict_blocks();
  const long long jobs = (long long)PA.a.B * Mt;
  if (jobs < grid) grid = (int)jobs;
  if (grid < 1) return 0;
  kern<<<grid, NTHREADS, C::SMEM_BYTES, st>>>(PA);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("lg_predict launch", e);
  return 0;
}

}  // namespace pgm
