// micro_tma_red.cu -- can one TMA bulk reduce (cp.reduce.async.bulk ... .add.f64, 48 bytes) per row replace six
// RED.ADD.F64 lane-ops per row?  (scratch probe; not part of the library)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)
__device__ __forceinline__ uint32_t hash32(uint32_t x){ x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void k_red6(double* table, uint32_t G, int64_t n){
  int64_t stride=(int64_t)gridDim.x*blockDim.x;
  for(int64_t i=(int64_t)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=stride){
    uint32_t k = hash32((uint32_t)i) % G;
    double v = (double)(i&7);
    #pragma unroll
    for(int c=0;c<6;++c) atomicAdd(table + (size_t)k*6 + c, v + c);
  }
}
template<int SLOTS> __global__ void __launch_bounds__(256) k_bulk6(double* table, uint32_t G, int64_t n){
  extern __shared__ __align__(16) double slots[];   // [256][SLOTS][6]
  double* mine = slots + (size_t)threadIdx.x * SLOTS * 6;
  int64_t stride=(int64_t)gridDim.x*blockDim.x;
  int s = 0;
  for(int64_t i=(int64_t)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=stride){
    uint32_t k = hash32((uint32_t)i) % G;
    double v = (double)(i&7);
    // the slot written SLOTS rounds ago must have been read by the TMA engine
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(SLOTS-1) : "memory");
    double2* d = reinterpret_cast<double2*>(mine + s*6);
    d[0] = make_double2(v, v+1); d[1] = make_double2(v+2, v+3); d[2] = make_double2(v+4, v+5);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(d);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 48;" :: "l"(table + (size_t)k*6), "r"(sa) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    s = (s + 1 == SLOTS) ? 0 : s + 1;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// one elected lane issues the bulk reduce for the whole warp's 32 rows?  not possible (32 different rows);
// variant: warp stages 32 rows contiguously and each LANE issues its own (same as above) -- kept single variant.
__global__ void k_sum(const double* t, size_t n, double* out){
  double s=0; for(size_t i=(size_t)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=(size_t)gridDim.x*blockDim.x) s+=t[i];
  atomicAdd(out,s);
}
template<typename F> float timeit(F f,int reps=3){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); float best=1e9;
  for(int r=0;r<reps;++r){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms;} return best; }
int main(){
  const int64_t n=50000000;
  for (uint32_t G : {1000000u, 4000000u}) {
    double *table,*out; CK(cudaMalloc(&table,(size_t)G*48)); CK(cudaMalloc(&out,8));
    auto total=[&]{ double h=0; CK(cudaMemset(out,0,8)); k_sum<<<592,256>>>(table,(size_t)G*6,out); CK(cudaMemcpy(&h,out,8,cudaMemcpyDeviceToHost)); return h; };
    CK(cudaMemset(table,0,(size_t)G*48)); k_red6<<<148*8,256>>>(table,G,n); double ref=total();
    printf("G=%u RED x6/row: %.3f ms (sum %.6e)\n", G, timeit([&]{k_red6<<<148*8,256>>>(table,G,n);}), ref);
    for (int blocks : {148*2, 148*4, 148*8}) {
      CK(cudaMemset(table,0,(size_t)G*48)); k_bulk6<2><<<blocks,256,256*2*48>>>(table,G,n); double got=total();
      printf("  bulk reduce 48B/row slots=2 grid=%d: %.3f ms (sum %.6e %s)\n", blocks, timeit([&]{k_bulk6<2><<<blocks,256,256*2*48>>>(table,G,n);}), got, got==ref?"OK":"MISMATCH");
      CK(cudaFuncSetAttribute(k_bulk6<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256*8*48));
      CK(cudaMemset(table,0,(size_t)G*48)); k_bulk6<8><<<blocks,256,256*8*48>>>(table,G,n); got=total();
      printf("  bulk reduce 48B/row slots=8 grid=%d: %.3f ms (sum %.6e %s)\n", blocks, timeit([&]{k_bulk6<8><<<blocks,256,256*8*48>>>(table,G,n);}), got, got==ref?"OK":"MISMATCH");
    }
    cudaFree(table); cudaFree(out);
  }
  return 0;
}
