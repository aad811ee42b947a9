// micro.cu -- throughput probes that decide the demean data-path design (scratch; not part of the library)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// each thread: rows i (grid-stride); row -> key = hash(i) % G ; adds C doubles to table[key*C + c]
template<int C> __global__ void k_red(double* table, uint32_t G, int64_t n, int replicas){
  int64_t stride=(int64_t)gridDim.x*blockDim.x;
  double* tb = table + (size_t)(blockIdx.x % replicas) * G * C;
  for(int64_t i=(int64_t)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=stride){
    uint32_t k = hash32((uint32_t)i) % G;
    double v = (double)(i&7);
    #pragma unroll
    for(int c=0;c<C;++c) atomicAdd(tb + (size_t)k*C + c, v);
  }
}
template<int C> __global__ void k_smem_red(double* out, uint32_t G, int64_t n){
  extern __shared__ double bins[];
  for(int i=threadIdx.x;i<G*C;i+=blockDim.x) bins[i]=0;
  __syncthreads();
  int64_t stride=(int64_t)gridDim.x*blockDim.x;
  for(int64_t i=(int64_t)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=stride){
    uint32_t k = hash32((uint32_t)i) % G;
    double v = (double)(i&7);
    #pragma unroll
    for(int c=0;c<C;++c) atomicAdd(bins + (size_t)k*C + c, v);
  }
  __syncthreads();
  for(int i=threadIdx.x;i<G*C;i+=blockDim.x) atomicAdd(out+i, bins[i]);
}
// gather C doubles (row) from table by random key, sum -> out (prevent DCE)
template<int C> __global__ void k_gather_rows(const double* __restrict__ table, uint32_t G, int64_t n, double* out){
  int64_t stride=(int64_t)gridDim.x*blockDim.x; double s=0;
  for(int64_t i=(int64_t)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=stride){
    uint32_t k = hash32((uint32_t)i) % G;
    const double2* r = reinterpret_cast<const double2*>(table + (size_t)k*C);
    #pragma unroll
    for(int c=0;c<C/2;++c){ double2 t=__ldg(r+c); s+=t.x+t.y; }
  }
  if(s==123.456) out[0]=s;
}
__global__ void k_gather8(const double* __restrict__ src, int64_t n, double* __restrict__ dst){
  int64_t stride=(int64_t)gridDim.x*blockDim.x;
  for(int64_t i=(int64_t)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=stride){
    uint32_t k = hash32((uint32_t)i) % (uint32_t)n;
    dst[i]=src[k];
  }
}
__global__ void k_copy(const double* __restrict__ src, int64_t n, double* __restrict__ dst){
  int64_t stride=(int64_t)gridDim.x*blockDim.x;
  for(int64_t i=(int64_t)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=stride) dst[i]=src[i];
}
template<typename F> float timeit(F f,int reps=3){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); float best=1e9;
  for(int r=0;r<reps;++r){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms;} return best; }
int main(){
  const int64_t n=50000000; const int C=6;
  double *table,*out,*src,*dst; CK(cudaMalloc(&table,(size_t)512<<20)); CK(cudaMalloc(&out,1<<20)); CK(cudaMemset(table,0,(size_t)512<<20));
  CK(cudaMalloc(&src,n*8)); CK(cudaMalloc(&dst,n*8)); CK(cudaMemset(src,0,n*8));
  int grid=148*8;
  printf("copy 400MB: %.3f ms\n", timeit([&]{k_copy<<<grid,256>>>(src,n,dst);}));
  printf("random 8B gather (400MB src) 50M: %.3f ms\n", timeit([&]{k_gather8<<<grid*2,256>>>(src,n,dst);}));
  uint32_t Gs[]={500,16000,1000000,10000000};
  for(uint32_t G:Gs){
    printf("G=%u (table %.1f MB): RED x6/row 50M rows: %.3f ms", G, G*48.0/1e6, timeit([&]{k_red<6><<<grid,256>>>(table,G,n,1);}));
    printf(" | RED x1/row: %.3f ms", timeit([&]{k_red<1><<<grid,256>>>(table,G,n,1);}));
    printf(" | gather 48B rows: %.3f ms\n", timeit([&]{k_gather_rows<6><<<grid,256>>>(table,G,n,out);}));
  }
  for(int rep: {8,64,256,1184}) printf("G=500 replicas=%d RED x6/row: %.3f ms\n", rep, timeit([&]{k_red<6><<<grid,256>>>(table,500,n,rep);}));
  CK(cudaFuncSetAttribute(k_smem_red<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 500*6*8));
  printf("smem fp64 atomicAdd G=500 x6/row (148x4 CTAs): %.3f ms\n", timeit([&]{k_smem_red<6><<<148*4,256,500*6*8>>>(out,500,n);}));
  printf("smem fp64 atomicAdd G=500 x6/row (148x8 CTAs x512): %.3f ms\n", timeit([&]{k_smem_red<6><<<148*4,512,500*6*8>>>(out,500,n);}));
  return 0;
}
