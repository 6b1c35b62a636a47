#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill(float4* p, size_t n, float v){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st) p[i]=make_float4(v,v,v,v);}
__global__ void rd(const float4* p, size_t n, float* out){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; float a=0; for(;i<n;i+=st){float4 v=p[i]; a+=v.x+v.y+v.z+v.w;} if(a==12345.f) *out=a;}
__global__ void rmw(float4* p, size_t n){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st){float4 v=p[i]; v.x+=1; p[i]=v;}}
int main(){ size_t bytes=2ull<<30; float4* p; cudaMalloc(&p,bytes); float* o; cudaMalloc(&o,4); size_t n=bytes/16; cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
 for(int k=0;k<3;k++){ fill<<<148*16,512>>>(p,n,1.f);} cudaEventRecord(a); for(int k=0;k<5;k++) fill<<<148*16,512>>>(p,n,1.f); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b); printf("write %.0f GB/s\n", 5*bytes/ms*1e-6);
 cudaEventRecord(a); for(int k=0;k<5;k++) rd<<<148*16,512>>>(p,n,o); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b); printf("read %.0f GB/s\n", 5*bytes/ms*1e-6);
 cudaEventRecord(a); for(int k=0;k<5;k++) rmw<<<148*16,512>>>(p,n); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b); printf("rmw in place %.0f GB/s (read+write bytes)\n", 2*5*bytes/ms*1e-6);
 cudaEventRecord(a); for(int k=0;k<5;k++) cudaMemsetAsync(p,0,bytes); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b); printf("memset %.0f GB/s\n", 5*bytes/ms*1e-6);
 return 0;}
