// Latency microbenchmarks on B200 (development aid): cycles per dependent op, single warp / single CTA.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__device__ __forceinline__ double rsq(double p){double y;asm("rsqrt.approx.ftz.f64 %0, %1;":"=d"(y):"d"(p));
 const double py=p*y;const double e=fma(-py,y,1.0);const double ye=y*e;const double q=fma(e,0.375,0.5);return fma(ye,q,y);}
__global__ void k(int mode, double* out, long long* cyc, double seed){
  __shared__ double sm[64];
  double x = seed + threadIdx.x*1e-9, y = 1.0000001, z=0.5;
  sm[threadIdx.x & 63] = x;
  __syncthreads();
  long long t0 = clock64();
  if (mode==0) { for(int i=0;i<N;i++) x = fma(x,y,z); }                       // dependent DFMA
  else if (mode==1) { float f=(float)x; for(int i=0;i<N;i++) f = fmaf(f,1.0000001f,0.5f); x=f; }  // dependent FFMA
  else if (mode==2) { for(int i=0;i<N;i++){ x = rsq(x)+1.5; } }               // rsqrt chain (+1 dadd)
  else if (mode==3) { for(int i=0;i<N;i++){ __syncthreads(); } }              // barrier
  else if (mode==4) { for(int i=0;i<N;i++){ x = __shfl_sync(0xffffffffu, x, (threadIdx.x+1)&31); } } // 64-bit shfl chain
  else if (mode==5) { int idx=threadIdx.x&63; for(int i=0;i<N;i++){ idx = (int)sm[idx & 63] & 63; } x=idx; } // LDS chain (+cvt)
  else if (mode==6) { double d0=0,d1=0; for(int i=0;i<N;i++){ asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};":"+d"(d0),"+d"(d1):"d"(x),"d"(y)); } x=d0+d1; } // dependent DMMA
  else if (mode==7) { for(int i=0;i<N;i++){ sm[threadIdx.x&63]=x; __syncthreads(); x = sm[(threadIdx.x+1)&63]*1.0000001; } } // STS->BAR->LDS->DMUL
  else if (mode==8) { double a0=x,a1=x+1,a2=x+2,a3=x+3; for(int i=0;i<N;i++){ a0=fma(a0,y,z);a1=fma(a1,y,z);a2=fma(a2,y,z);a3=fma(a3,y,z);} x=a0+a1+a2+a3; } // 4 indep DFMA chains
  long long t1 = clock64();
  if (threadIdx.x==0) cyc[blockIdx.x] = t1-t0;
  out[blockIdx.x*blockDim.x+threadIdx.x] = x;
}
int main(){
  double* out; long long* cyc; cudaMalloc(&out, 8*1024*1024); cudaMalloc(&cyc, 8*4096);
  const char* names[]={"dep DFMA","dep FFMA","rsqrt(4op)+dadd","bar.sync(256thr)","shfl64 chain","LDS chain","dep DMMA","STS-BAR-LDS-DMUL (256thr)","4 indep DFMA chains (per iter)"};
  for (int threads : {32, 256}) for (int grid : {1, 296}) for (int mode=0; mode<9; ++mode){
    k<<<grid,threads>>>(mode,out,cyc,1.25); cudaDeviceSynchronize();
    k<<<grid,threads>>>(mode,out,cyc,1.25); cudaDeviceSynchronize();
    long long h[4096]; cudaMemcpy(h,cyc,8*grid,cudaMemcpyDeviceToHost);
    double s=0; for(int i=0;i<grid;i++) s+=h[i];
    printf("threads=%3d grid=%3d %-32s %.1f cyc/iter\n",threads,grid,names[mode],s/grid/N);
  }
  return 0;
}
