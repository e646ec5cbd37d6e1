// FP64 pipe arbitration probes (follow-up of mix_bench.cu):
//  (1) does the warp slot matter?  DFMA class in the LOW warps, DMMA class in the high warps
//  (2) 4 DMMA warps (one per SM sub-partition) next to DFMA warps
//  (3) a dependent DMMA chain (one accumulator) next to DMMA streams: does a DMMA-typed chain get the pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pgmuvi_b200/csrc -o scratch/mix2_bench scratch/mix2_bench.cu
#include "gp_fused.cuh"
#include <cstdio>
using namespace pgm;

// MODE 0: warps 0-7 DMMA stream, 8-15 DFMA chain (ILP 1)      [reference]
// MODE 1: warps 0-7 DFMA chain, 8-15 DMMA stream
// MODE 2: warps 0-3 DMMA stream (one per sub-partition), 8-15 DFMA chain, 4-7 idle
// MODE 3: warps 0-7 DMMA stream, 8-15 dependent DMMA chain (one accumulator)
// MODE 4: warps 0-7 DMMA stream with one accumulator each (dependent), 8-15 DFMA chain
template <int MODE>
__global__ void __launch_bounds__(512, 1) mix(int itA, int itB, long long* out, double* sink) {
  const int tid = threadIdx.x, warp = tid >> 5;
  double a = 1.0 + 1e-9 * tid, b = 1.0 - 1e-9 * tid;
  const bool clsA = (MODE == 1) ? (warp >= 8) : (MODE == 2) ? (warp < 4) : (warp < 8);
  const bool clsB = (MODE == 1) ? (warp < 8) : (warp >= 8);
  __syncthreads();
  const long long t0 = clock64();
  double s = 0;
  if (clsA) {
    double acc[8][2];
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0;
    for (int it = 0; it < itA; ++it)
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) mma_f64(acc[MODE == 4 ? 0 : i], a, b);
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  } else if (clsB) {
    if (MODE == 3) {
      double acc[2] = {0, 0};
      for (int it = 0; it < itB; ++it)
#pragma unroll
        for (int u = 0; u < 64; ++u) mma_f64(acc, a, b);
      s = acc[0] + acc[1];
    } else {
      double v = 1e-3;
      for (int it = 0; it < itB; ++it)
#pragma unroll
        for (int u = 0; u < 64; ++u) v = fma(v, a, b);
      s = v;
    }
  }
  const long long t1 = clock64();
  if ((tid & 31) == 0 && (warp == 0 || warp == 8)) out[2 * blockIdx.x + (clsB ? 1 : 0)] = t1 - t0;
  if (s == 123.456) sink[0] = s;
}

template <int MODE>
void run(const char* name, int itA, int itB) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d; double* sk; cudaMalloc(&d, sms * 2 * 8); cudaMalloc(&sk, 8);
  cudaMemset(d, 0, sms * 16);
  mix<MODE><<<sms, 512>>>(itA, itB, d, sk);
  mix<MODE><<<sms, 512>>>(itA, itB, d, sk);
  cudaDeviceSynchronize();
  long long h[2 * 256]; cudaMemcpy(h, d, sms * 16, cudaMemcpyDeviceToHost);
  double ca = 0, cb = 0; for (int i = 0; i < sms; ++i) { ca += h[2 * i]; cb += h[2 * i + 1]; }
  printf("%-58s itA %5d itB %5d : A %9.0f cyc   B %9.0f cyc\n", name, itA, itB, ca / sms, cb / sms);
  cudaFree(d); cudaFree(sk);
}

int main() {
  run<0>("0: DMMA stream w0-7 | DFMA chain w8-15, B alone", 0, 1000);
  run<0>("0: DMMA stream w0-7 | DFMA chain w8-15", 2000, 1000);
  run<1>("1: DFMA chain w0-7 | DMMA stream w8-15", 2000, 1000);
  run<2>("2: DMMA stream w0-3 (1 per SMSP) | DFMA chain w8-15, A alone", 2000, 0);
  run<2>("2: DMMA stream w0-3 (1 per SMSP) | DFMA chain w8-15", 2000, 1000);
  run<3>("3: DMMA stream w0-7 | dependent DMMA chain w8-15, B alone", 0, 200);
  run<3>("3: DMMA stream w0-7 | dependent DMMA chain w8-15", 2000, 200);
  run<4>("4: dependent DMMA w0-7 | DFMA chain w8-15, A alone", 2000, 0);
  run<4>("4: dependent DMMA w0-7 | DFMA chain w8-15", 2000, 1000);
  return 0;
}
