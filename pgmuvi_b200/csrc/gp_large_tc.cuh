// pgmuvi_b200 - the G phase of the staged engine on Blackwell tensor cores (tcgen05 / TMEM):
//
//   K~^-1 = X^T X   (X = L^-1, the "lauum" third of the n^3 flops)   as 3xTF32 products with FP32
//   accumulators in tensor memory, each 128x128 tile contracted at once - in the epilogue, straight
//   out of TMEM - with W = alpha alpha^T - K~^-1 and the regenerated dK/dtheta in FP64: the north
//   star's kernel (3) "0.5 tr((alpha alpha^T - K^-1) dK/dtheta) without ever materialising it",
//   for float32 models (the reference's default dtype, pgmuvi/lightcurve.py:2434-2446; loss.backward()
//   of pgmuvi/trainers.py:181).  The MLL itself (Cholesky, solves, log-det) stays on the FP64 path.
//
//   lg_pack_tf32   X^T tiles (fp64 64x64 images of the T phase) -> packed TF32 hi / lo operand images
//                  U[c, k] = X[k, c], one 16 KB SWIZZLE_128B image per (128 rows, 32 k) - byte for byte
//                  what tcgen05.mma reads from shared memory, so a pipeline stage is four bulk copies
//   lg_grad_tc     one CTA per 128x128 tile (I >= J) of K~^-1:  warp 0 = bulk-copy producer,
//                  warp 1 = MMA issuer (one thread; owns the TMEM allocation), warps 2-9 = epilogue
//                  (tcgen05.ld -> W -> dK/dtheta contraction in FP64 -> per-tile partial sums)
//
// Tensor memory (all 512 columns, one CTA per SM).  tcgen05.mma truncates on every accumulating
// instruction (tc_tf32.cuh), so the sum over k is split into PIECES of 16 k (half a 32-k chunk):
//   [256,512) two piece buffers, ping-pong between the MMA warp and the epilogue warps (mbarriers
//             pfull[2] / pempty[2]).  Per piece: the 4 cross-term MMAs (hi lo + lo hi) first - the
//             accumulator is still 2^-11 small, their truncation is harmless - then the 2 hi hi MMAs,
//             i.e. two truncating additions per piece instead of K / 8 * 3 per tile;
//   [  0,256) master accumulator as an unevaluated float pair (hi [0,128), lo [128,256)): the epilogue
//             warps drain every piece and add it with Fast2Sum on the CUDA cores (round to nearest,
//             the rounding error of each addition kept in lo).
#pragma once
#include "gp_large.cuh"
#include "tc_tf32.cuh"

namespace pgm {

#define PGM_FLAG_TF32X3 16   // G phase on tcgen05 (3xTF32); set by pgm_sm_mll_grad_tf32x3_f32

constexpr int TC_THREADS = 320;        // producer warp + MMA warp + 8 epilogue warps
constexpr int TC_EPI_THREADS = 256;

// packed operand images of one light curve: piece (0 = hi, 1 = lo), 128-row block I, k-chunk kc
__host__ __device__ inline size_t tc_pack_floats(int n_max) {   // per light curve, both pieces
  const size_t N = (n_max + TS - 1) / TS, NT = (N + 1) / 2, KCH = 2 * N;
  return 2 * NT * KCH * (size_t)tc::IMG_FLOATS;
}
__host__ __device__ inline size_t large_ws_bytes_tc(int n_max, int B) {
  return ((large_ws_bytes(n_max, B) + 1023) & ~(size_t)1023) + tc_pack_floats(n_max) * 4 * (size_t)B;
}
__host__ __device__ inline float* tc_pack_base(double* ws, int n_max, int B, int b) {
  char* p = reinterpret_cast<char*>(ws) + ((large_ws_bytes(n_max, B) + 1023) & ~(size_t)1023);
  return reinterpret_cast<float*>(p) + tc_pack_floats(n_max) * (size_t)b;
}

// grid (Nmax, 2 * NTmax, B): block (a, b64) converts the 64 rows of tile column b64 of X (rows of U)
// against the 64 k of tile row a.  a > b64: the stored X_ab^T image; a == b64: X_bb^T (tilesT);
// a < b64 inside the same 128-row block or b64 >= N: zeros.
static __global__ void __launch_bounds__(NTHREADS) lg_pack_tf32(LargeArgs A) {
  const LcView v = lc_view(A, (int)blockIdx.z);
  if (v.st.state[v.b] != LG_FACTORED) return;
  const int a = blockIdx.x, b64 = blockIdx.y, I = b64 >> 1;
  if (a >= v.N || a < 2 * I || I >= (v.N + 1) / 2) return;
  const int Nmax = (A.n_max + TS - 1) / TS, NTmax = (Nmax + 1) / 2, KCH = 2 * Nmax;
  float* hi = tc_pack_base(A.ws, A.n_max, A.B, v.b);
  float* lo = hi + (size_t)NTmax * KCH * tc::IMG_FLOATS;
  const double* src = nullptr;
  if (b64 < v.N) {
    if (a > b64) src = lg_tile(v.w.tilesX, a, b64);
    else if (a == b64) src = v.w.tilesT + (size_t)b64 * TT;
  }
  const int row0 = (b64 & 1) * TS;
  for (int idx = threadIdx.x; idx < TT / 2; idx += NTHREADS) {
    const int c = idx >> 5, k = (idx & 31) * 2;     // row c of the image (column of X), k pair
    float2 h = make_float2(0.f, 0.f), l = make_float2(0.f, 0.f);
    if (src) {
      const double2 x = *reinterpret_cast<const double2*>(src + img(c, k));
      tc::split_tf32(x.x, h.x, l.x);
      tc::split_tf32(x.y, h.y, l.y);
    }
    const size_t o = ((size_t)I * KCH + 2 * a + (k >> 5)) * tc::IMG_FLOATS + tc::sw128_idx(row0 + c, k & 31);
    *reinterpret_cast<float2*>(hi + o) = h;
    *reinterpret_cast<float2*>(lo + o) = l;
  }
}

template <int KIND, int QT, int D>
struct TcGradSmem {
  using C = Cfg<KIND, QT, D>;
  static constexpr int SIDE = 2 * C::NF * TS;                 // doubles per side (two 64-point sub-tiles)
  static constexpr int PAR_TAB = 0;                           // exp tables [96]
  static constexpr int PAR_RED = PAR_TAB + EXP_TAB;           // [8][NV] warp partials
  static constexpr int PAR_BAR = PAR_RED + 8 * C::NV;         // mbarriers (10) + TMEM slot
  static constexpr int PAR_END = PAR_BAR + 14;
  static __host__ __device__ constexpr size_t bytes(int nst) {
    return 1024 + (size_t)nst * 4 * tc::IMG_BYTES + (size_t)(2 * SIDE + PAR_END) * sizeof(double);
  }
};

template <int KIND, int QT, int D>
__global__ void __launch_bounds__(TC_THREADS, 1) lg_grad_tc(LargeArgs A, int nst, int half_pieces) {
  using C = Cfg<KIND, QT, D>;
  using SM = TcGradSmem<KIND, QT, D>;
  constexpr int DS = C::DS;
  static_assert(C::NV <= LG_GP, "gradient partial slot too small");
  extern __shared__ __align__(1024) unsigned char smraw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const LcView v = lc_view(A);
  if (v.st.state[v.b] != LG_FACTORED) return;
  const LargeWs& w = v.w;
  const int n = v.n, N = v.N, npad = v.npad;
  int I, J;
  tri_unrank(blockIdx.x, I, J);
  if (I >= (N + 1) / 2) return;
  const int Nmax = (A.n_max + TS - 1) / TS, NTmax = (Nmax + 1) / 2, KCH = 2 * Nmax;
  const float* phi = tc_pack_base(A.ws, A.n_max, A.B, v.b);
  const float* plo = phi + (size_t)NTmax * KCH * tc::IMG_FLOATS;
  const int kc0 = 4 * I, kc1 = 2 * N;      // U[128 I .., k] = 0 for k < 128 I

  const unsigned base = (smem_u32(smraw) + 1023u) & ~1023u;
  unsigned char* gen = smraw + (base - smem_u32(smraw));
  double* sd = reinterpret_cast<double*>(gen + (size_t)nst * 4 * tc::IMG_BYTES);
  double* rowv = sd;
  double* colv = sd + SM::SIDE;
  double* par = sd + 2 * SM::SIDE;
  double* tab = par + SM::PAR_TAB;
  double* red = par + SM::PAR_RED;
  // full[3] | empty[3] at +24 | pfull[2] at +48 | pempty[2] at +64 | TMEM slot at +96
  const unsigned bars = smem_u32(par + SM::PAR_BAR);
  const unsigned bar_full = bars, bar_empty = bars + 24, bar_pfull = bars + 48, bar_pempty = bars + 64;
  const unsigned tslot = bars + 96;
  constexpr unsigned T_MASTER = 0, T_MLO = 128, T_PIECE = 256;
  // pieces per chunk: 2 (16 k, two truncating hi hi additions; default) or 1 (32 k, four)
  const int ppc = half_pieces ? 2 : 1, kpp = 4 / ppc;
  const int npieces = (kc1 - kc0) * ppc;

  if (tid == 0) {
    for (int s = 0; s < nst; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_pfull + 8 * s, 1); mbar_init(bar_pempty + 8 * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 1) tc::tmem_alloc(tslot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  unsigned tmem;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem) : "r"(tslot));

  if (warp == 0) {
    if (lane == 0) {
      for (int kc = kc0; kc < kc1; ++kc) {
        const int c = kc - kc0, s = c % nst, use = c / nst;
        if (use > 0) mbar_wait(bar_empty + 8 * s, (use - 1) & 1);
        const unsigned bar = bar_full + 8 * s, dst = base + (unsigned)s * 4 * tc::IMG_BYTES;
        const size_t oa = ((size_t)I * KCH + kc) * tc::IMG_FLOATS, ob = ((size_t)J * KCH + kc) * tc::IMG_FLOATS;
        mbar_expect_tx(bar, 4 * tc::IMG_BYTES);
        bulk_g2s(dst, phi + oa, tc::IMG_BYTES, bar);
        bulk_g2s(dst + tc::IMG_BYTES, plo + oa, tc::IMG_BYTES, bar);
        bulk_g2s(dst + 2 * tc::IMG_BYTES, phi + ob, tc::IMG_BYTES, bar);
        bulk_g2s(dst + 3 * tc::IMG_BYTES, plo + ob, tc::IMG_BYTES, bar);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(128, 128);
      for (int kc = kc0; kc < kc1; ++kc) {
        const int c = kc - kc0, s = c % nst, use = c / nst;
        mbar_wait(bar_full + 8 * s, use & 1);
        tc::fence_after_sync();
        const unsigned st = base + (unsigned)s * 4 * tc::IMG_BYTES;
        for (int hp = 0; hp < ppc; ++hp) {
          const int piece = c * ppc + hp, buf = piece & 1;
          if (piece >= 2) {                  // the epilogue has drained this piece buffer
            mbar_wait(bar_pempty + 8 * buf, ((piece >> 1) - 1) & 1);
            tc::fence_after_sync();
          }
          const unsigned tp = tmem + T_PIECE + 128u * (unsigned)buf;
          const int ks0 = hp * kpp;
          for (int ks = ks0; ks < ks0 + kpp; ++ks) {     // cross terms first: hi lo + lo hi
            const uint64_t ahi = tc::smem_desc_sw128(st + ks * 32);
            const uint64_t alo = tc::smem_desc_sw128(st + tc::IMG_BYTES + ks * 32);
            const uint64_t bhi = tc::smem_desc_sw128(st + 2 * tc::IMG_BYTES + ks * 32);
            const uint64_t blo = tc::smem_desc_sw128(st + 3 * tc::IMG_BYTES + ks * 32);
            tc::mma_tf32(tp, alo, bhi, idesc, ks > ks0 ? 1u : 0u);
            tc::mma_tf32(tp, ahi, blo, idesc, 1u);
          }
          for (int ks = ks0; ks < ks0 + kpp; ++ks) {     // then hi hi
            const uint64_t ahi = tc::smem_desc_sw128(st + ks * 32);
            const uint64_t bhi = tc::smem_desc_sw128(st + 2 * tc::IMG_BYTES + ks * 32);
            tc::mma_tf32(tp, ahi, bhi, idesc, 1u);
          }
          tc::mma_commit(bar_pfull + 8 * buf);  // piece complete
        }
        tc::mma_commit(bar_empty + 8 * s);      // the stage is free once these MMAs have read it
      }
    }
  } else {
    // ---------------- epilogue warps: per-point fields while the products run ----------------
    const int et = tid - 64;                              // 0..255
    const int q4 = warp & 3, half = (warp - 2) >> 2;      // TMEM lane quarter, column half
    const int r = 32 * q4 + lane;                         // accumulator row of this thread
    if (et < EXP_TAB) tab[et] = c_exp2_tab[et];
    for (int side = 0; side < 2; ++side) {
      double* vec = side ? colv : rowv;
      const int T0 = 2 * (side ? J : I);                  // first 64-point tile of the side
      for (int idx = et; idx < 2 * C::NF * TS; idx += TC_EPI_THREADS) {
        const int sub = idx / (C::NF * TS), o = idx - sub * C::NF * TS;
        const int f = o >> 6, p = o & 63;                 // only for the x / alpha fields below
        const int gp = (T0 + sub) * TS;                   // first point of the sub-tile
        double val = 0.0;
        if (T0 + sub < N) {
          if (o < D * TS) {
            val = w.fx[(size_t)f * npad + gp + p];
          } else if (o < C::NFB * TS) {
            const int o2 = o - D * TS, fc = o2 >> 7, pp = o2 & 127;   // (cos, sin) pairs: 128 doubles per field
            val = w.fcs[((size_t)fc * npad + gp) * 2 + pp];
          } else {
            val = w.alpha[gp + p];
          }
        } else if (o >= D * TS && o < C::NFB * TS) {
          val = ((o - D * TS) & 1) ? 0.0 : 1.0;           // padded points: cos = 1, sin = 0
        }
        vec[sub * C::NF * TS + o] = val;
      }
    }
    double wreg[QT], areg[QT * DS], lam[4];
#pragma unroll
    for (int q = 0; q < QT; ++q) wreg[q] = w.par[LG_PAR_WQ + q];
#pragma unroll
    for (int q = 0; q < QT * DS; ++q) areg[q] = w.par[LG_PAR_AQ + q];
#pragma unroll
    for (int q = 0; q < 4; ++q) lam[q] = w.par[LG_PAR_LM + q];
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    double ga[C::NG];
#pragma unroll
    for (int t = 0; t < C::NG; ++t) ga[t] = 0.0;
    double trW = 0.0;
    const double* rv = rowv + (r >> 6) * C::NF * TS;
    const double* cv = colv + half * C::NF * TS;
    const int rr = r & 63;
    const int gi = I * 128 + r;
    const double al_i = rv[C::NFB * TS + rr];
    const unsigned tlane = tmem + ((unsigned)(32 * q4) << 16) + (unsigned)(half * 64);
    // drain the pieces into the master accumulator: (hi, lo) += piece by Fast2Sum (|hi| >= |piece|
    // but for the first few pieces, where the missed rounding error is of no consequence)
    for (int piece = 0; piece < npieces; ++piece) {
      const int buf = piece & 1;
      mbar_wait(bar_pfull + 8 * buf, (piece >> 1) & 1);
      tc::fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        float pv[16], lo[16];
        tc::tmem_ld16(tlane + T_PIECE + 128u * (unsigned)buf + (unsigned)c0, pv);
        if (piece > 0) {
          float hi[16];
          tc::tmem_ld16(tlane + T_MASTER + (unsigned)c0, hi);
          tc::tmem_ld16(tlane + T_MLO + (unsigned)c0, lo);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float s2 = __fadd_rn(hi[e], pv[e]);
            const float er = __fsub_rn(pv[e], __fsub_rn(s2, hi[e]));
            lo[e] = __fadd_rn(lo[e], er);
            pv[e] = s2;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) lo[e] = 0.f;
        }
        tc::tmem_st16(tlane + T_MASTER + (unsigned)c0, pv);
        tc::tmem_st16(tlane + T_MLO + (unsigned)c0, lo);
      }
      tc::tmem_wait_st();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pempty + 8 * buf);
    }
    // tcgen05.ld is warp-collective: the chunk loop is uniform over the warp (rows 32 q4 .. + 31),
    // the lower-triangle / n masks act per entry
    const int gi_hi = I * 128 + 32 * q4 + 31;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 16) {
      const int gj0 = J * 128 + half * 64 + c0;
      if (gj0 > gi_hi || I * 128 + 32 * q4 >= n) break;
      float kv[16], xv[16];
      tc::tmem_ld16(tlane + T_MASTER + (unsigned)c0, kv);
      tc::tmem_ld16(tlane + T_MLO + (unsigned)c0, xv);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int gj = gj0 + e;
        if (gi < n && gj <= gi) {
          const double W = al_i * cv[C::NFB * TS + c0 + e] - ((double)kv[e] + (double)xv[e]);
          if (gi == gj) trW += W;
          const double wgt = (gi == gj) ? W : 2.0 * W;
          k_grad_entry<KIND, QT, D>(rv, cv, rr, c0 + e, wreg, areg, lam, tab, wgt, ga);
        }
      }
    }
    tc::fence_before_sync();
    // reduction over the 256 epilogue threads (fixed order)
    double vv[C::NV];
#pragma unroll
    for (int t = 0; t < C::NG; ++t) vv[t] = ga[t];
    vv[C::NG] = trW;
#pragma unroll
    for (int t = 0; t < C::NV; ++t) {
      double s = vv[t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += shfl_xor_d(s, o);
      if (lane == 0) red[(warp - 2) * C::NV + t] = s;
    }
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    if (et < C::NV) {
      double s = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * C::NV + et];
      w.gpart[(size_t)blockIdx.x * LG_GP + et] = s;
    }
  }
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace pgm
