// micro_s0_replica.cu -- round-2 probe for the S0 pass: per row, 6 values go to a big by-key table (1M rows) AND to a
// small one (500 rows).  A: bulk reduce (big) + shared-memory fp64 atomics (small)  [what k_s0_fused does today]
//                        B: bulk reduce (big) + bulk reduce into a per-CTA replica of the small table in global/L2
// (scratch probe; not part of the library)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)
__device__ __forceinline__ uint32_t hash32(uint32_t x){ x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__device__ __forceinline__ void bulk6(double* gdst, double* slot, double v){
  double2* d = reinterpret_cast<double2*>(slot);
  d[0] = make_double2(v, v+1); d[1] = make_double2(v+2, v+3); d[2] = make_double2(v+4, v+5);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  uint32_t sa = (uint32_t)__cvta_generic_to_shared(d);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 48;" :: "l"(gdst), "r"(sa) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// MODE 0: big only (bulk).  1: big bulk + small smem CAS.  2: big bulk + small bulk into replica.  3: small smem only.
template<int MODE> __global__ void __launch_bounds__(256) k(double* big, uint32_t G, double* small_rep, uint32_t Gs, int64_t n){
  extern __shared__ __align__(16) double sm[];
  double* bins = sm;                                   // [Gs*6] (modes 1,3)
  double* slots = sm + ((MODE==1||MODE==3) ? Gs*6 : 0) + (size_t)threadIdx.x * 4 * 6;   // 4 slots per thread
  if (MODE==1||MODE==3){ for(int i=threadIdx.x;i<(int)Gs*6;i+=blockDim.x) bins[i]=0; __syncthreads(); }
  double* rep = small_rep + (size_t)blockIdx.x * Gs * 6;
  int64_t stride=(int64_t)gridDim.x*blockDim.x;
  int s = 0;
  for(int64_t i=(int64_t)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=stride){
    uint32_t h = hash32((uint32_t)i);
    uint32_t k = h % G, t = (h >> 7) % Gs;
    double v = (double)(i&7);
    if (MODE != 3) {
      asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(MODE==2 ? 2 : 1) : "memory");
      bulk6(big + (size_t)k*6, slots + s*6, v); s = (s+1)&3;
    }
    if (MODE==1||MODE==3){
      #pragma unroll
      for(int c=0;c<6;++c) atomicAdd(bins + (size_t)t*6 + c, v + c);
    } else if (MODE==2){
      bulk6(rep + (size_t)t*6, slots + s*6, v); s = (s+1)&3;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (MODE==1||MODE==3){ __syncthreads(); for(int i=threadIdx.x;i<(int)Gs*6;i+=blockDim.x) if(bins[i]!=0) atomicAdd(small_rep+i, bins[i]); }
}
template<typename F> float timeit(F f,int reps=3){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); float best=1e9;
  for(int r=0;r<reps;++r){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms;} return best; }
int main(){
  const int64_t n=50000000; const uint32_t G=1000000, Gs=500;
  double *big,*rep; CK(cudaMalloc(&big,(size_t)G*48)); CK(cudaMemset(big,0,(size_t)G*48));
  for (int grid : {296, 592}) {
    CK(cudaMalloc(&rep,(size_t)grid*Gs*48)); CK(cudaMemset(rep,0,(size_t)grid*Gs*48));
    size_t sm_bins = (size_t)Gs*48, sm_slots = 256*4*48;
    CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sm_bins+sm_slots)));
    CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sm_slots)));
    CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sm_slots)));
    CK(cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sm_bins+sm_slots)));
    printf("grid=%d  big bulk only: %.3f ms", grid, timeit([&]{k<0><<<grid,256,sm_slots>>>(big,G,rep,Gs,n);}));
    printf(" | small smem CAS only: %.3f ms", timeit([&]{k<3><<<grid,256,sm_bins+sm_slots>>>(big,G,rep,Gs,n);}));
    printf(" | A big bulk + small smem CAS: %.3f ms", timeit([&]{k<1><<<grid,256,sm_bins+sm_slots>>>(big,G,rep,Gs,n);}));
    printf(" | B big bulk + small bulk->replica: %.3f ms\n", timeit([&]{k<2><<<grid,256,sm_slots>>>(big,G,rep,Gs,n);}));
    cudaFree(rep);
  }
  return 0;
}
