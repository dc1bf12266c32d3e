#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){ unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__global__ void k2(float* out, float w0, float w1, int n){
  unsigned long long acc[8]; 
  for(int i=0;i<8;++i) acc[i]=pk(threadIdx.x+i, i);
  unsigned long long w = pk(w0,w1), a = pk(w1,w0);
  for(int it=0;it<n;++it){
#pragma unroll
    for(int i=0;i<8;++i) acc[i]=fma2(w,a,acc[i]);
  }
  float s=0; for(int i=0;i<8;++i){float x,y; upk(acc[i],x,y); s+=x+y;}
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k1(float* out, float w0, float w1, int n){
  float acc[16];
  for(int i=0;i<16;++i) acc[i]=threadIdx.x+i;
  for(int it=0;it<n;++it){
#pragma unroll
    for(int i=0;i<16;++i) acc[i]=fmaf(w0,w1,acc[i]);
  }
  float s=0; for(int i=0;i<16;++i) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){ float* d; cudaMalloc(&d, 148*1024*4*4); cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
 for(int rep=0;rep<2;++rep){
  cudaEventRecord(e0); k1<<<148*2,512>>>(d,1.0001f,0.5f,20000); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1);
  printf("FFMA : %.3f ms  %.1f TFLOP/s\n", ms, 148.0*2*512*20000*16*2/ms/1e9);
  cudaEventRecord(e0); k2<<<148*2,512>>>(d,1.0001f,0.5f,20000); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
  printf("FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, 148.0*2*512*20000*16*2/ms/1e9);
 }
 return 0; }
