// k-loop microbenchmark: how much of the DMMA peak do 8 / 16 warps per SM reach with the tile
// engine's compute_chunk (operands from shared memory), and with register-only DMMA streams?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pgmuvi_b200/csrc -o scratch/kloop_bench scratch/kloop_bench.cu
#include "gp_fused.cuh"
#include <cstdio>
#include <cstdlib>
using namespace pgm;

template <int VAR>
__global__ void __launch_bounds__(256, 2) bench(int iters, double* out) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3, wm = warp >> 2, wn = warp & 3;
  for (int i = tid; i < 2 * 2 * OPBUF; i += 256) sm[i] = 1e-3 * ((i * 37) % 101);
  __syncthreads();
  double acc[4][2][2];
  zero_acc(acc);
  if (VAR == 0) {
    for (int it = 0; it < iters; ++it) {
      const double* sA = sm + (it & 1) * 2 * OPBUF;
      compute_chunk<M_FULL, false>(acc, sA, sA + OPBUF, 0, wm, wn, g, tq);
    }
  } else if (VAR == 1) {   // registers only, 8 independent accumulators, h outer
    double a = sm[tid], b = sm[tid + 256];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 2; ++ni) mma_f64(acc[mi][ni], a, b);
    }
  } else if (VAR == 2) {   // registers only, 4 independent accumulators
    double a = sm[tid], b = sm[tid + 256];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u)
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) mma_f64(acc[mi][0], a, b);
    }
  } else if (VAR == 3) {   // registers only, 2 independent accumulators
    double a = sm[tid], b = sm[tid + 256];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 32; ++u)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) mma_f64(acc[mi][0], a, b);
    }
  } else if (VAR == 4) {   // dependent chain
    double a = sm[tid], b = sm[tid + 256];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 64; ++u) mma_f64(acc[0][0], a, b);
    }
  }
  double s = 0;
  for (int mi = 0; mi < 4; ++mi) for (int ni = 0; ni < 2; ++ni) s += acc[mi][ni][0] + acc[mi][ni][1];
  if (s == 123.456) out[0] = s;
}

template <int VAR>
void run(const char* name, int bps, int threads_note) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  size_t smem = bps == 1 ? 120 * 1024 : bps == 2 ? 100 * 1024 : bps == 3 ? 70 * 1024 : 50 * 1024;
  cudaFuncSetAttribute(bench<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bench<VAR>, 256, smem);
  double* d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); bench<VAR><<<sms * occ, 256, smem>>>(iters, d); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double dmma_per_warp = 64.0 * iters;   // every variant issues 64 DMMA per warp per iteration
  const double fl = dmma_per_warp * 512.0 * 8 * sms * occ;
  printf("%-34s blocks/SM %d (occ %d): %8.3f ms  %6.2f TFLOP/s\n", name, bps, occ, best, fl / best / 1e9);
  cudaFree(d);
}

int main() {
  for (int bps = 1; bps <= 4; ++bps) {
    run<0>("compute_chunk (smem operands)", bps, 0);
    run<1>("regs, 8 accumulators", bps, 0);
    run<2>("regs, 4 accumulators", bps, 0);
    run<3>("regs, 2 accumulators", bps, 0);
    run<4>("regs, 1 accumulator (chain)", bps, 0);
  }
  return 0;
}
