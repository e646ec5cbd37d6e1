// 64x64 diagonal-block micro-kernel (potrf_inv_64: S -> L in place, X = L^-1 into S2): cycles per
// call with one / two blocks per SM, and the residuals |L L^T - A|, |X L - I|.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pgmuvi_b200/csrc -o scratch/potrf_bench scratch/potrf_bench.cu
#include "gp_fused.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
using namespace pgm;

// timing variants of potrf_inv_64 (results are wrong unless VAR == 0):
//   1: panel factorisations only   2: no inverse rows (step 2b)   3: no trailing step 2a   4: panels + step 1 only
template <int VAR>
__device__ __forceinline__ void potrf_var(double* __restrict__ S, double* __restrict__ S2,
                                          double* __restrict__ dinv, int* fail) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  if (warp == 0) potrf_panel8(S, dinv, 0, lane, fail);
  for (int p = 0; p < 8; ++p) {
    const int c0 = p * 8;
    __syncthreads();
    if (VAR != 1 && p < 7 && warp < 7 - p) potrf_syrk_tile(S, c0, p + 1 + warp, p + 1, g, tq);
    __syncthreads();
    if (p < 7 && warp == 0) {
      potrf_panel8(S, dinv, c0 + 8, lane, fail);
    } else if (VAR != 1 && VAR != 4) {
      const int nw = (p < 7) ? 7 : 8, w = (p < 7) ? warp - 1 : warp;
      const int m8 = 6 - p;
      const int cnt = (m8 > 0) ? m8 * (m8 + 1) / 2 : 0;
      if (VAR != 3)
      for (int t = w; t < cnt; t += nw) {
        int a_ = 0;
        while ((a_ + 1) * (a_ + 2) / 2 <= t) ++a_;
        const int b_ = t - a_ * (a_ + 1) / 2;
        potrf_syrk_tile(S, c0, p + 2 + a_, p + 2 + b_, g, tq);
      }
      if (VAR != 2)
      for (int nt = w; nt <= p; nt += nw) {
        double c2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) c2[e] = (c0 + g == 8 * nt + 2 * tq + e) ? 1.0 : 0.0;
        for (int kt = nt; kt < p; ++kt) {
#pragma unroll
          for (int s = 0; s < 2; ++s)
            mma_f64(c2, -S[(c0 + g) * LD_S + 8 * kt + 4 * s + tq],
                    S2[(8 * kt + 4 * s + tq) * LD_S + 8 * nt + g]);
        }
        *reinterpret_cast<double2*>(S2 + (c0 + g) * LD_S + 8 * nt + 2 * tq) = make_double2(c2[0], c2[1]);
        __syncwarp();
        if (lane < 8) {
          const int c = 8 * nt + lane;
          double v[8];
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) v[rr] = S2[(c0 + rr) * LD_S + c];
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            double sacc = v[rr];
#pragma unroll
            for (int kk = 0; kk < rr; ++kk) sacc -= S[(c0 + rr) * LD_S + c0 + kk] * v[kk];
            v[rr] = sacc * dinv[c0 + rr];
          }
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) S2[(c0 + rr) * LD_S + c] = v[rr];
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
}
#ifndef PVAR
#define PVAR 0
#endif
#define POTRF potrf_var<PVAR>

__global__ void __launch_bounds__(256, 2) bench(const double* A, int iters, long long* cyc, double* Lout,
                                                double* Xout, int* failout) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;
  double* S2 = sm + S_ELEMS;
  double* dinv = S2 + S_ELEMS;
  __shared__ int fail;
  const int tid = threadIdx.x;
  if (tid == 0) fail = 0;
  long long total = 0;
  for (int it = 0; it < iters; ++it) {
    for (int idx = tid; idx < TT; idx += 256) {
      const int r = idx >> 6, c = idx & 63;
      S[r * LD_S + c] = A[idx];
      S2[r * LD_S + c] = 0.0;
    }
    __syncthreads();
    const long long t0 = clock64();
    POTRF(S, S2, dinv, &fail);
    const long long t1 = clock64();
    total += t1 - t0;
    __syncthreads();
  }
  if (tid == 0) { cyc[blockIdx.x] = total / iters; failout[blockIdx.x] = fail; }
  if (blockIdx.x == 0)
    for (int idx = tid; idx < TT; idx += 256) {
      const int r = idx >> 6, c = idx & 63;
      Lout[idx] = (c <= r) ? S[r * LD_S + c] : 0.0;
      Xout[idx] = (c <= r) ? S2[r * LD_S + c] : 0.0;
    }
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  std::vector<double> M(TT), A(TT);
  srand(1);
  for (auto& v : M) v = (rand() / (double)RAND_MAX) - 0.5;
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < 64; ++k) s += M[i * 64 + k] * M[j * 64 + k];
      A[i * 64 + j] = s;
    }
  double *dA, *dL, *dX; long long* dc; int* df;
  cudaMalloc(&dA, TT * 8); cudaMalloc(&dL, TT * 8); cudaMalloc(&dX, TT * 8);
  cudaMalloc(&dc, sms * 2 * 8); cudaMalloc(&df, sms * 2 * 4);
  cudaMemcpy(dA, A.data(), TT * 8, cudaMemcpyHostToDevice);
  const size_t smem = (2 * S_ELEMS + 64) * 8;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int bps = 1; bps <= 2; ++bps) {
    bench<<<sms * bps, 256, smem>>>(dA, 50, dc, dL, dX, df);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> c(sms * bps); std::vector<int> f(sms * bps);
    cudaMemcpy(c.data(), dc, sms * bps * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(f.data(), df, sms * bps * 4, cudaMemcpyDeviceToHost);
    double avg = 0; int nf = 0; for (int i = 0; i < sms * bps; ++i) { avg += c[i]; nf += f[i] != 0; }
    printf("blocks/SM %d: %.0f cycles per potrf_inv_64 (avg over %d blocks), failures %d\n", bps, avg / (sms * bps),
           sms * bps, nf);
  }
  std::vector<double> L(TT), X(TT);
  cudaMemcpy(L.data(), dL, TT * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(X.data(), dX, TT * 8, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0;
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) {
      double s = 0, t = 0;
      for (int k = 0; k < 64; ++k) { s += L[i * 64 + k] * L[j * 64 + k]; t += X[i * 64 + k] * L[k * 64 + j]; }
      e1 = fmax(e1, fabs(s - A[i * 64 + j]));
      e2 = fmax(e2, fabs(t - (i == j ? 1.0 : 0.0)));
    }
  printf("max |L L^T - A| = %.3e   max |X L - I| = %.3e\n", e1, e2);
  return 0;
}
