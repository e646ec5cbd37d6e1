// stand-alone check of the tcgen05 3xTF32 product: D = A B^T, A, B [128 x K] doubles split into packed
// SWIZZLE_128B images on the host; compares with the fp64 product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pgmuvi_b200/csrc -o scratch/tc_test scratch/tc_test.cu
#include "gp_fused.cuh"
#include "tc_tf32.cuh"
#include <cstdio>
#include <cstring>
#include <vector>
#include <cmath>
#include <cstdlib>
using namespace pgm;
using namespace pgm::tc;

constexpr int NST = 3;
__global__ void __launch_bounds__(192, 1)
tc_prod(const float* Ahi, const float* Alo, const float* Bhi, const float* Blo, int nchunks, float* D, int mode) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  const unsigned base = (smem_u32(smraw) + 1023u) & ~1023u;      // operand images need 1024-B alignment
  const unsigned bars = base + NST * 4 * IMG_BYTES;               // full[NST], empty[NST], tmem_full
  const unsigned tslot = bars + 8 * (2 * NST + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (NST + s), 1); }
    mbar_init(bars + 8 * 2 * NST, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tslot, 128);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  unsigned tmem;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem) : "r"(tslot));
  if (warp == 0 && lane == 0) {
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % NST, use = c / NST;
      if (use > 0) mbar_wait(bars + 8 * (NST + s), (use - 1) & 1);
      const unsigned bar = bars + 8 * s, dst = base + s * 4 * IMG_BYTES;
      mbar_expect_tx(bar, 4 * IMG_BYTES);
      bulk_g2s(dst, Ahi + (size_t)c * IMG_FLOATS, IMG_BYTES, bar);
      bulk_g2s(dst + IMG_BYTES, Alo + (size_t)c * IMG_FLOATS, IMG_BYTES, bar);
      bulk_g2s(dst + 2 * IMG_BYTES, Bhi + (size_t)c * IMG_FLOATS, IMG_BYTES, bar);
      bulk_g2s(dst + 3 * IMG_BYTES, Blo + (size_t)c * IMG_FLOATS, IMG_BYTES, bar);
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = idesc_tf32(128, 128);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % NST, use = c / NST;
      mbar_wait(bars + 8 * s, use & 1);
      fence_after_sync();
      const unsigned st = base + s * 4 * IMG_BYTES;
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ahi = smem_desc_sw128(st + ks * 32), alo = smem_desc_sw128(st + IMG_BYTES + ks * 32);
        const uint64_t bhi = smem_desc_sw128(st + 2 * IMG_BYTES + ks * 32);
        const uint64_t blo = smem_desc_sw128(st + 3 * IMG_BYTES + ks * 32);
        if (mode == 0) {
          mma_tf32(tmem, alo, bhi, idesc, (c | ks) ? 1u : 0u);
          mma_tf32(tmem, ahi, blo, idesc, 1u);
          mma_tf32(tmem, ahi, bhi, idesc, 1u);
        } else if (mode == 1) {   // hi hi only
          mma_tf32(tmem, ahi, bhi, idesc, (c | ks) ? 1u : 0u);
        } else {                  // cross terms only
          mma_tf32(tmem, alo, bhi, idesc, (c | ks) ? 1u : 0u);
          mma_tf32(tmem, ahi, blo, idesc, 1u);
        }
      }
      mma_commit(bars + 8 * (NST + s));
    }
    mma_commit(bars + 8 * 2 * NST);
  } else if (warp >= 2) {
    mbar_wait(bars + 8 * 2 * NST, 0);
    fence_after_sync();
    const int q = warp & 3, row = 32 * q + lane;
    for (int c0 = 0; c0 < 128; c0 += 16) {
      float v[16];
      tmem_ld16(tmem + ((unsigned)(32 * q) << 16) + c0, v);
      for (int i = 0; i < 16; ++i) D[row * 128 + c0 + i] = v[i];
    }
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

int run(int K) {
  const int nch = K / 32;
  std::vector<double> A(128 * K), B(128 * K);
  srand(1);
  for (auto& v : A) v = (rand() / (double)RAND_MAX - 0.5) * 2;
  for (auto& v : B) v = (rand() / (double)RAND_MAX - 0.5) * 2;
  std::vector<float> ahi(nch * IMG_FLOATS), alo(ahi.size()), bhi(ahi.size()), blo(ahi.size());
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < K; ++k) {
      const size_t o = (size_t)(k / 32) * IMG_FLOATS + sw128_idx(r, k % 32);
      split_tf32(A[r * K + k], ahi[o], alo[o]);
      split_tf32(B[r * K + k], bhi[o], blo[o]);
    }
  float *dah, *dal, *dbh, *dbl, *dD;
  const size_t by = ahi.size() * 4;
  cudaMalloc(&dah, by); cudaMalloc(&dal, by); cudaMalloc(&dbh, by); cudaMalloc(&dbl, by); cudaMalloc(&dD, 128 * 128 * 4);
  cudaMemcpy(dah, ahi.data(), by, cudaMemcpyHostToDevice); cudaMemcpy(dal, alo.data(), by, cudaMemcpyHostToDevice);
  cudaMemcpy(dbh, bhi.data(), by, cudaMemcpyHostToDevice); cudaMemcpy(dbl, blo.data(), by, cudaMemcpyHostToDevice);
  const size_t smem = NST * 4 * IMG_BYTES + 1024 + 256;
  cudaFuncSetAttribute(tc_prod, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> D(128 * 128);
  for (int mode = 0; mode < 3; ++mode) {
    tc_prod<<<1, 192, smem>>>(dah, dal, dbh, dbl, nch, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, sumabs = 0, errsplit = 0, bias = 0, biasabs = 0; int nb = 0;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 128; ++c) {
        double ref = 0, sa = 0, full = 0;
        for (int k = 0; k < K; ++k) {
          const double a = A[r * K + k], b = B[c * K + k];
          float ah, al, bh, bl; split_tf32(a, ah, al); split_tf32(b, bh, bl);
          double t = mode == 0 ? (double)ah * bh + (double)ah * bl + (double)al * bh
                   : mode == 1 ? (double)ah * bh : (double)ah * bl + (double)al * bh;
          ref += t; sa += fabs(a * b); full += a * b;
        }
        maxerr = fmax(maxerr, fabs(D[r * 128 + c] - ref));
        if (fabs(ref) > 1e-3) { const double rel = (D[r * 128 + c] - ref) / fabs(ref) * (ref > 0 ? 1 : -1); bias += rel; biasabs += fabs(rel); ++nb; }
        if (mode == 0) errsplit = fmax(errsplit, fabs(full - ref));
        sumabs = fmax(sumabs, sa);
      }
    printf("K=%5d mode %d (%s): max |D - exact sum of the SAME split products| = %.3e  (= %.2e of sum|terms| %.1f)%s\n", K, mode,
           mode == 0 ? "hh+hl+lh" : mode == 1 ? "hh only " : "hl+lh   ", maxerr, maxerr / sumabs, sumabs, "");
    printf("         signed relative error toward larger magnitude: mean %.3e, mean |.| %.3e  (fp32 ulp = 6e-8..1.2e-7)\n", bias / nb, biasabs / nb);
    if (mode == 0) printf("         split error (dropped lo lo + representation) = %.3e (%.2e of sum|terms|)\n", errsplit, errsplit / sumabs);
  }
  cudaFree(dah); cudaFree(dal); cudaFree(dbh); cudaFree(dbl); cudaFree(dD);
  return 0;
}
int main() { for (int K : {32, 64, 384, 4096, 32768}) if (run(K)) return 1; return 0; }
