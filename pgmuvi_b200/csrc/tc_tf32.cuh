// pgmuvi_b200 - Blackwell tensor-core (tcgen05 / TMEM) primitives for the 3xTF32 products of the
// fp32 model path (north star: "trailing SYRK/GEMM updates on tensor cores, FP64 or TF32-refined,
// matching the reference dtype"; the reference's default dtype is float32,
// pgmuvi/lightcurve.py:2434-2446).
//
// Operand images.  A tcgen05.mma operand is a K-major [128 rows x 32 k] float tile in the canonical
// SWIZZLE_128B layout: row pitch 128 B, 8-row groups 1024 B apart, 16-byte chunk index XOR-ed with
// (row & 7).  HBM holds byte-for-byte the same 16 KB image ("packed panel"), so one cp.async.bulk
// per operand fills a pipeline stage - the convention of the FP64 tile engine (gp_fused.cuh).
//
// 3xTF32: x = hi + lo with hi = tf32(x), lo = tf32(x - hi) (both round-to-nearest: x is represented
// to 2^-24); a b ~= hi_a hi_b + hi_a lo_b + lo_a hi_b (the lo lo term, 2^-24 relative, is dropped).
//
// Accumulation (measured on B200, scratch/tc_test.cu, gpurun_out/r02f_tc_test.log): every
// tcgen05.mma that adds into a TMEM accumulator TRUNCATES - about one fp32 ulp of the accumulator is
// lost per instruction, with a bias, so the error grows linearly in the number of MMAs (K = 32768:
// 2.7e-6 of sum |terms| for the hi hi products alone, 3x that with the cross terms on top), while
// the split itself is good to 2e-9.  Hence the users of these primitives (gp_large_tc.cuh)
//   * keep the small cross terms hi lo + lo hi in an accumulator of their own (2^-11 of the
//     magnitude, so its truncation is harmless), and
//   * accumulate the hi hi products in short PIECES (8 MMAs) that the epilogue warps drain and add
//     with round-to-nearest on the CUDA cores into a master accumulator (also in TMEM).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pgm {
namespace tc {

constexpr int TM = 128;                 // rows of an operand image / of the accumulator tile
constexpr int TKC = 32;                 // floats per row of an operand image (one 128-byte swizzle atom)
constexpr int IMG_FLOATS = TM * TKC;    // 4096 floats
constexpr unsigned IMG_BYTES = IMG_FLOATS * 4;   // 16 KB

// float index of element (r, k) inside a [128 x 32] SWIZZLE_128B K-major image
__host__ __device__ __forceinline__ int sw128_idx(int r, int k) {
  return (r >> 3) * 256 + (r & 7) * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3));
}

// fp64 -> (hi, lo) TF32 pair.  hi keeps the 10 explicit mantissa bits of TF32 (round to nearest on
// the dropped 13 bits of the float), lo is the float nearest to the remainder, cut to TF32 as well
// (the tensor core ignores the low 13 bits of its fp32 inputs).
__host__ __device__ __forceinline__ float tf32_rn(float v) {
#ifdef __CUDA_ARCH__
  unsigned u = __float_as_uint(v);
#else
  unsigned u;
  memcpy(&u, &v, 4);
#endif
  u += 0x00000fffu + ((u >> 13) & 1u);
  u &= 0xffffe000u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float o;
  memcpy(&o, &u, 4);
  return o;
#endif
}
__host__ __device__ __forceinline__ void split_tf32(double x, float& hi, float& lo) {
  hi = tf32_rn((float)x);
  lo = tf32_rn((float)(x - (double)hi));
}

#ifdef __CUDACC__
// shared-memory matrix descriptor of a K-major SWIZZLE_128B operand whose 8-row groups are 1024 B
// apart (cute::UMMA::SmemDescriptor: start >> 4 | LBO << 16 | SBO << 32 | version 1 << 46 | layout << 61)
__device__ __forceinline__ uint64_t smem_desc_sw128(unsigned smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(unsigned smem_dst, unsigned ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {    // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, one thread issues for the CTA
__device__ __forceinline__ void mma_tf32(unsigned taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         unsigned accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(taddr),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier gets one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar)
               : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id)
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
  // the registers are written asynchronously: tie them to the wait so no use is scheduled above it
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                 "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]),
                 "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// Batched form: issue several loads, ONE tmem_wait_ld(), then tmem_tie() every destination array (an
// empty volatile asm that redefines the registers after the wait, so that no use of the asynchronously
// written registers can be scheduled above it).
__device__ __forceinline__ void tmem_ld16_async(unsigned taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15}, [%16];\n"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
        "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]),
        "=f"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_tie(float (&v)[16]) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]),
                    "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]),
                    "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
               :
               : "memory");
}
// ... and back: 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(unsigned taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
#endif  // __CUDACC__

}  // namespace tc
}  // namespace pgm
