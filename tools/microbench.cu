// Pipe and memory ceilings of the box the kernels are tuned against:
// FFMA / FFMA2 / DFMA issue rates and a float4 copy.  Build: see tools/run_microbench.sh
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

template<int ILP> __global__ void k_ffma(float* out, float a, float b, int iters){
  float v[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) v[i]=threadIdx.x+i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<ILP;i++) v[i]=fmaf(v[i],a,b);
  }
  float s=0; 
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_ffma2(float2* out, float a, float b, int iters){
  unsigned long long v[ILP];
  float2 av=make_float2(a,a), bv=make_float2(b,b);
  unsigned long long aa=*reinterpret_cast<unsigned long long*>(&av), bb=*reinterpret_cast<unsigned long long*>(&bv);
  #pragma unroll
  for(int i=0;i<ILP;i++){ float2 t=make_float2(threadIdx.x+i, i); v[i]=*reinterpret_cast<unsigned long long*>(&t);} 
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<ILP;i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(aa), "l"(bb));
  }
  float2 s=make_float2(0,0);
  #pragma unroll
  for(int i=0;i<ILP;i++){ float2 t=*reinterpret_cast<float2*>(&v[i]); s.x+=t.x; s.y+=t.y; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_dfma(double* out, double a, double b, int iters){
  double v[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) v[i]=threadIdx.x+i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<ILP;i++) v[i]=fma(v[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n){
  for(size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; i<n; i+=(size_t)gridDim.x*blockDim.x) out[i]=in[i];
}
// shared-memory read bandwidth: every lane reads 16 B per iteration, conflict free
__global__ void k_lds(float* out, int iters){
  __shared__ float4 s[1024];
  for(int i=threadIdx.x;i<1024;i+=blockDim.x) s[i]=make_float4(i,i,i,i);
  __syncthreads();
  float4 acc=make_float4(0,0,0,0);
  int idx=threadIdx.x;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int u=0;u<8;u++){ float4 t=s[(idx+u*32)&1023]; acc.x+=t.x; acc.y+=t.y; acc.z+=t.z; acc.w+=t.w; }
    idx=(idx+1)&1023;
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc.x+acc.y+acc.z+acc.w;
}
template<typename F> float timeit(F f, int rep=5){
  cudaEvent_t a,b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<rep;r++){ CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms,a,b)); if(ms<best)best=ms; }
  return best;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount; printf("device %s sms %d clock %d kHz smem_optin %zu\n", p.name, sms, p.clockRate, p.sharedMemPerBlockOptin);
  void* buf; CK(cudaMalloc(&buf, (size_t)sms*8*1024*16));
  const int iters=4096; const int threads=1024; int blocks=sms*2;
  { float ms=timeit([&]{k_ffma<16><<<blocks,threads>>>((float*)buf,1.0001f,0.5f,iters);});
    double fma=(double)blocks*threads*16.0*iters; printf("FFMA  : %.2f TFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", fma/ms/1e9, fma/ms/1e6/sms/1965.0*1e3/1e3); }
  { float ms=timeit([&]{k_ffma2<16><<<blocks,threads>>>((float2*)buf,1.0001f,0.5f,iters);});
    double fma=(double)blocks*threads*16.0*2*iters; printf("FFMA2 : %.2f TFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", fma/ms/1e9, fma/ms/1e6/sms/1965.0*1e3/1e3); }
  { float ms=timeit([&]{k_dfma<16><<<blocks,threads>>>((double*)buf,1.0001,0.5,iters/4);});
    double fma=(double)blocks*threads*16.0*(iters/4); printf("DFMA  : %.2f TFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", fma/ms/1e9, fma/ms/1e6/sms/1965.0*1e3/1e3); }
  { float ms=timeit([&]{k_lds<<<blocks,threads>>>((float*)buf,2048);});
    double bytes=(double)blocks*threads*2048.0*8*16; printf("LDS.128: %.2f TB/s (%.1f B/clk/SM @1.965GHz)\n", bytes/ms/1e9, bytes/ms/1e6/sms/1965.0); }
  { size_t n=(size_t)1<<30; float4 *a,*b; CK(cudaMalloc(&a,n)); CK(cudaMalloc(&b,n)); CK(cudaMemset(a,1,n));
    float ms=timeit([&]{k_copy<<<sms*16,512>>>(a,b,n/16);}); printf("copy float4 1 GiB: %.1f GB/s (read+write)\n", 2.0*n/ms/1e6);
    CK(cudaFree(a)); CK(cudaFree(b)); }
  return 0;
}
