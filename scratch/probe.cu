// dev probe: pipe sharing DMMA/DFMA and inner-loop ceiling (not shipped)
#include "../pgmuvi_b200/csrc/gp_fused.cuh"
#include <cstdio>
using namespace pgm;
__global__ void __launch_bounds__(256) p_dmma(int iters, double* out) {
  double acc[8][2]; for (int i=0;i<8;++i) acc[i][0]=acc[i][1]=0.0;
  double a=1.0+threadIdx.x*1e-9,b=1.0-threadIdx.x*1e-9;
  for (int it=0;it<iters;++it){
#pragma unroll
    for(int i=0;i<8;++i) mma_f64(acc[i],a,b);}
  double s=0; for(int i=0;i<8;++i) s+=acc[i][0]+acc[i][1]; if(s==123.456) out[0]=s;
}
__global__ void __launch_bounds__(256) p_dfma(int iters, double* out) {
  double f[16]; for(int i=0;i<16;++i) f[i]=threadIdx.x*1e-9+i;
  const double fa=1.0000001,fb=1e-9;
  for(int it=0;it<iters;++it){
#pragma unroll
    for(int i=0;i<16;++i) f[i]=fma(f[i],fa,fb);}
  double s=0; for(int i=0;i<16;++i) s+=f[i]; if(s==123.456) out[0]=s;
}
// R dfma per dmma (per thread): dmma = 512 flop/warp = 16 flop/thread; dfma = 2 flop/thread
template<int R>
__global__ void __launch_bounds__(256) p_mixed(int iters, double* out) {
  double acc[8][2], f[16];
  for (int i=0;i<8;++i) acc[i][0]=acc[i][1]=0.0;
  for(int i=0;i<16;++i) f[i]=threadIdx.x*1e-9+i;
  double a=1.0+threadIdx.x*1e-9,b=1.0-threadIdx.x*1e-9; const double fa=1.0000001,fb=1e-9;
  for(int it=0;it<iters;++it){
#pragma unroll
    for(int i=0;i<8;++i){ mma_f64(acc[i],a,b);
#pragma unroll
      for(int j=0;j<R;++j) f[(i*R+j)&15]=fma(f[(i*R+j)&15],fa,fb);}
  }
  double s=0; for(int i=0;i<8;++i) s+=acc[i][0]+acc[i][1]; for(int i=0;i<16;++i) s+=f[i]; if(s==123.456) out[0]=s;
}
// FFMA alongside DMMA
template<int R>
__global__ void __launch_bounds__(256) p_mixed_f32(int iters, double* out) {
  double acc[8][2]; float f[16];
  for (int i=0;i<8;++i) acc[i][0]=acc[i][1]=0.0;
  for(int i=0;i<16;++i) f[i]=threadIdx.x*1e-9f+i;
  double a=1.0+threadIdx.x*1e-9,b=1.0-threadIdx.x*1e-9; const float fa=1.0000001f,fb=1e-9f;
  for(int it=0;it<iters;++it){
#pragma unroll
    for(int i=0;i<8;++i){ mma_f64(acc[i],a,b);
#pragma unroll
      for(int j=0;j<R;++j) f[(i*R+j)&15]=fmaf(f[(i*R+j)&15],fa,fb);}
  }
  double s=0; for(int i=0;i<8;++i) s+=acc[i][0]+acc[i][1]; for(int i=0;i<16;++i) s+=f[i]; if(s==123.456) out[0]=s;
}
// inner loop of the tile engine on resident smem chunks
__global__ void __launch_bounds__(256,2) p_inner(int iters, double* out) {
  extern __shared__ __align__(16) double sm[];
  for (int i=threadIdx.x;i<2*OPBUF;i+=256) sm[i]=1.0+1e-9*i;
  __syncthreads();
  const int tid=threadIdx.x, lane=tid&31, warp=tid>>5, g=lane>>2,tq=lane&3,wm=warp>>2,wn=warp&3;
  double acc[4][2][2]; zero_acc(acc);
  for(int it=0;it<iters;++it) compute_chunk<M_FULL,false>(acc, sm, sm+OPBUF, 0, wm, wn, g, tq);
  double s=0; for(int a=0;a<4;++a)for(int b=0;b<2;++b) s+=acc[a][b][0]+acc[a][b][1];
  if(s==123.456) out[0]=s;
}
template<typename F> float timeit(F f){ cudaEvent_t e0,e1; cudaEventCreate(&e0);cudaEventCreate(&e1); float best=1e30f;
  for(int r=0;r<4;++r){cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(r&&ms<best)best=ms;} return best;}
int main(){
  double* d; cudaMalloc(&d,8); int sms=148; const int it=8192; const int blocks=sms*8;
  double thr=(double)blocks*256*it;
  float t;
  t=timeit([&]{p_dmma<<<blocks,256>>>(it,d);}); printf("dmma   %.2f TF/s\n", thr*8*16/t/1e9);
  t=timeit([&]{p_dfma<<<blocks,256>>>(it,d);}); printf("dfma   %.2f TF/s\n", thr*16*2/t/1e9);
  t=timeit([&]{p_mixed<1><<<blocks,256>>>(it,d);}); printf("mixed R=1: dmma %.2f + dfma %.2f TF/s (%.3f ms)\n", thr*8*16/t/1e9, thr*8*2/t/1e9,t);
  t=timeit([&]{p_mixed<2><<<blocks,256>>>(it,d);}); printf("mixed R=2: dmma %.2f + dfma %.2f TF/s\n", thr*8*16/t/1e9, thr*16*2/t/1e9);
  t=timeit([&]{p_mixed<4><<<blocks,256>>>(it,d);}); printf("mixed R=4: dmma %.2f + dfma %.2f TF/s\n", thr*8*16/t/1e9, thr*32*2/t/1e9);
  t=timeit([&]{p_mixed<8><<<blocks,256>>>(it,d);}); printf("mixed R=8: dmma %.2f + dfma %.2f TF/s\n", thr*8*16/t/1e9, thr*64*2/t/1e9);
  t=timeit([&]{p_mixed_f32<8><<<blocks,256>>>(it,d);}); printf("mixed32 R=8: dmma %.2f + ffma %.2f TF/s\n", thr*8*16/t/1e9, thr*64*2/t/1e9);
  t=timeit([&]{p_mixed_f32<16><<<blocks,256>>>(it,d);}); printf("mixed32 R=16: dmma %.2f + ffma %.2f TF/s\n", thr*8*16/t/1e9, thr*128*2/t/1e9);
  cudaFuncSetAttribute(p_inner, cudaFuncAttributeMaxDynamicSharedMemorySize, 2*OPBUF*8);
  const int it2=2048;
  for (int occ=1; occ<=2; ++occ){
    t=timeit([&]{p_inner<<<sms*occ,256,2*OPBUF*8>>>(it2,d);});
    printf("inner loop occ=%d: %.2f TF/s\n", occ, (double)sms*occ*it2*(64.0*64*32*2)/t/1e9);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
