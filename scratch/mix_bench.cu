// Do a DMMA stream and a DFMA stream issued by DIFFERENT warps of an SM share the FP64 pipe fairly?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pgmuvi_b200/csrc -o scratch/mix_bench scratch/mix_bench.cu
#include "gp_fused.cuh"
#include <cstdio>
using namespace pgm;

// 512 threads: warps 0-7 = class A (DMMA, 8 accumulators), warps 8-15 = class B (DFMA, ILP chains)
// itA / itB = 0 switches a class off.  out[2*block + cls] = cycles of warp 0 / 8 of the class.
template <int ILP>
__global__ void __launch_bounds__(512, 1) mix(int itA, int itB, long long* out, double* sink) {
  const int tid = threadIdx.x, warp = tid >> 5;
  double a = 1.0 + 1e-9 * tid, b = 1.0 - 1e-9 * tid;
  __syncthreads();
  const long long t0 = clock64();
  double s = 0;
  if (warp < 8) {
    double acc[8][2];
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0;
    for (int it = 0; it < itA; ++it)
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) mma_f64(acc[i], a, b);
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  } else {
    double v[ILP];
    for (int i = 0; i < ILP; ++i) v[i] = 1e-3 * i;
    for (int it = 0; it < itB; ++it)
#pragma unroll
      for (int u = 0; u < 64 / ILP; ++u)
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    for (int i = 0; i < ILP; ++i) s += v[i];
  }
  const long long t1 = clock64();
  if ((tid & 255) == 0) out[2 * blockIdx.x + (warp >= 8)] = t1 - t0;
  if (s == 123.456) sink[0] = s;
}

template <int ILP>
void run(int itA, int itB) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d; double* sk; cudaMalloc(&d, sms * 2 * 8); cudaMalloc(&sk, 8);
  cudaMemset(d, 0, sms * 16);
  mix<ILP><<<sms, 512>>>(itA, itB, d, sk);
  mix<ILP><<<sms, 512>>>(itA, itB, d, sk);
  cudaDeviceSynchronize();
  long long h[2 * 256]; cudaMemcpy(h, d, sms * 16, cudaMemcpyDeviceToHost);
  double ca = 0, cb = 0; for (int i = 0; i < sms; ++i) { ca += h[2 * i]; cb += h[2 * i + 1]; }
  ca /= sms; cb /= sms;
  // pipe cycles needed per SM: A: itA*64 DMMA per warp * 8 warps * 4 cyc; B: itB*64 DFMA per warp * 8 warps * 0.5 cyc
  const double needA = itA * 64.0 * 8 * 4, needB = itB * 64.0 * 8 * 0.5;
  printf("ILP %d itA %6d itB %7d : A %10.0f cyc (pipe need %9.0f, util %5.1f%%)   B %10.0f cyc (need %9.0f, util %5.1f%%)\n",
         ILP, itA, itB, ca, needA, itA ? 100 * needA / ca : 0.0, cb, needB, itB ? 100 * needB / cb : 0.0);
  cudaFree(d); cudaFree(sk);
}

int main() {
  run<8>(2000, 0);
  run<8>(0, 16000);
  run<8>(2000, 16000);     // equal pipe need
  run<8>(2000, 4000);      // B needs 1/4 of A
  run<8>(4000, 2000);
  run<4>(0, 16000);
  run<4>(2000, 4000);
  run<2>(0, 16000);
  run<2>(2000, 2000);
  run<1>(0, 8000);
  run<1>(2000, 1000);
  return 0;
}
