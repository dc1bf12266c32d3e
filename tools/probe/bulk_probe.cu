#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void k(const float* __restrict__ src, float* out, int n) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long bar;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(n * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"((uint32_t)__cvta_generic_to_shared(sm)), "l"(src), "r"(n * 4), "r"(b) : "memory");
    }
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(0) : "memory");
    float s = 0; for (int i = threadIdx.x; i < n; i += blockDim.x) s += sm[i];
    out[threadIdx.x] = s;
}
int main(){ float *d,*o; int n=8192; cudaMalloc(&d,n*4); cudaMalloc(&o,128*4); float* h=(float*)malloc(n*4); for(int i=0;i<n;++i)h[i]=1.f; cudaMemcpy(d,h,n*4,cudaMemcpyHostToDevice);
 cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, n*4); k<<<1,128,n*4>>>(d,o,n); float r[128]; cudaMemcpy(r,o,512,cudaMemcpyDeviceToHost); printf("%f %s\n", r[0], cudaGetErrorString(cudaGetLastError())); }
