// Pipe-throughput microbenchmarks for B200 (sm_100a): FFMA2 alone, MUFU.RCP alone, and mixed at the ratios the
// intersection kernel uses.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipes.cu -o pipes && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pack2(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void unpack2(unsigned long long v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){ unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ float rcpa(float x){ float r; asm volatile("rcp.approx.ftz.f32 %0, %1;":"=f"(r):"f"(x)); return r;}

// MODE: 0 = 10 FFMA2 / iter; 1 = 10 FFMA2 + 2 MUFU / iter (kernel ratio); 2 = 10 FFMA2 + 1 MUFU; 3 = 8 MUFU / iter only
// 4 = 10 FFMA2 + 2 MUFU + 1 FMNMX3-like (fmin) ; 5 = scalar FFMA x20 + 2 MUFU
template<int MODE>
__global__ void __launch_bounds__(256) k(int iters, float* out){
  float seed = 1.0f + threadIdx.x*1e-6f;
  unsigned long long x[10]; float u[8];
  for(int j=0;j<10;j++) x[j]=pack2(seed+j, seed-j);
  for(int j=0;j<8;j++) u[j]=seed+0.1f*j;
  const unsigned long long b=pack2(0.9999f,1.0001f), c=pack2(1e-4f,2e-4f);
  float m = 1e30f;
  for(int i=0;i<iters;i++){
    if(MODE==0||MODE==1||MODE==2||MODE==4){
      #pragma unroll
      for(int j=0;j<10;j++) x[j]=fma2(x[j],b,c);
    }
    if(MODE==1||MODE==4){ u[0]=rcpa(u[0]); u[1]=rcpa(u[1]); }
    if(MODE==2){ u[0]=rcpa(u[0]); }
    if(MODE==3){
      #pragma unroll
      for(int j=0;j<8;j++) u[j]=rcpa(u[j]);
    }
    if(MODE==4){ float a0,a1; unpack2(x[i&7?0:1],a0,a1); m=fminf(m,fminf(a0,a1)); }
    if(MODE==6){   // FFMA2 with a scalar .F32 broadcast multiplicand (as in the intersection kernel)
      #pragma unroll
      for(int j=0;j<10;j++) x[j]=fma2(pack2(u[j&7],u[j&7]),x[j],c);
    }
    if(MODE==7){   // FFMA2 with a scalar .F32 broadcast addend
      #pragma unroll
      for(int j=0;j<10;j++) x[j]=fma2(x[j],b,pack2(u[j&7],u[j&7]));
    }
    if(MODE==8){   // square-accumulate form: fma2(x, x, y)
      #pragma unroll
      for(int j=0;j<10;j++) x[j]=fma2(x[(j+1)%10],x[(j+1)%10],x[j]);
    }
    if(MODE==9){   // three distinct register pairs per FFMA2 (register-file read-port pressure)
      unsigned long long yy[10], zz[10];
      #pragma unroll
      for(int j=0;j<10;j++){ yy[j]=pack2(u[j&7]+j, u[(j+1)&7]); zz[j]=pack2(u[(j+2)&7], u[(j+3)&7]-j); }
      #pragma unroll
      for(int r=0;r<4;r++){
        #pragma unroll
        for(int j=0;j<10;j++) x[j]=fma2(yy[j],zz[(j+r)%10],x[j]);
      }
    }
    if(MODE==5){
      float a0[20];
      #pragma unroll
      for(int j=0;j<10;j++){ float lo,hi; unpack2(x[j],lo,hi); lo=fmaf(lo,0.9999f,1e-4f); hi=fmaf(hi,1.0001f,2e-4f); x[j]=pack2(lo,hi);}  
      u[0]=rcpa(u[0]); u[1]=rcpa(u[1]);
    }
  }
  float s=m; for(int j=0;j<10;j++){float lo,hi; unpack2(x[j],lo,hi); s+=lo+hi;} for(int j=0;j<8;j++) s+=u[j];
  if(s==12345.6789f) out[0]=s;
}
template<int MODE> void run(const char* name, double fma2_per_iter, double mufu_per_iter){
  float* out; cudaMalloc(&out,4); int iters=20000; int grid=148*8;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid,256>>>(iters,out); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<grid,256>>>(iters,out); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double warps = (double)grid*8; double cyc = ms*1e-3*1.965e9; // assume max clock
  double warp_iters_per_smsp = warps*iters/(148.0*4.0);
  double cyc_per_iter = cyc/warp_iters_per_smsp;
  printf("%-40s %8.3f ms  cycles/warp-iter/SMSP %.2f  (FFMA2 %.0f -> ideal %.1f cyc; MUFU %.0f)\n", name, ms, cyc_per_iter, fma2_per_iter, fma2_per_iter*2, mufu_per_iter);
  cudaFree(out);
}
int main(){
  run<0>("10 FFMA2", 10, 0);
  run<1>("10 FFMA2 + 2 MUFU.RCP", 10, 2);
  run<2>("10 FFMA2 + 1 MUFU.RCP", 10, 1);
  run<3>("8 MUFU.RCP", 0, 8);
  run<4>("10 FFMA2 + 2 MUFU + FMNMX3", 10, 2);
  run<5>("20 FFMA scalar + 2 MUFU", 10, 2);
  run<6>("10 FFMA2 bcast multiplicand", 10, 0);
  run<7>("10 FFMA2 bcast addend", 10, 0);
  run<8>("10 FFMA2 x*x+y", 10, 0);
  run<9>("40 FFMA2, 3 distinct reg pairs", 40, 0);
  return 0;
}
