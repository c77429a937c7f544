// Debug aid: which TMA box shapes / coordinates does the hardware accept?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int bytes, float* out, int nout, int use_elect){
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  if(threadIdx.x==0){
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  bool issue = use_elect ? (threadIdx.x < 32) : (threadIdx.x == 0);
  if(issue){
    bool leader = true;
    if(use_elect){ uint32_t pred; asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0,1,0,P;\n}":"=r"(pred)); leader = pred; }
    if(leader){
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(&bar)),"r"(bytes):"memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(s32(smem)),"l"((uint64_t)&tm),"r"(c0),"r"(c1),"r"(s32(&bar)):"memory");
    }
  }
  uint32_t ok=0; 
  while(!ok){ asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}":"=r"(ok):"r"(s32(&bar)):"memory"); }
  const float* s=(const float*)smem;
  for(int i=threadIdx.x;i<nout;i+=blockDim.x) out[i]=s[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(){
  struct Cfg{int box0, box1, c0, c1, elect; const char* name;};
  std::vector<Cfg> cfgs={{64,8,0,0,1,"64x8 pos elect"},{64,8,0,0,0,"64x8 pos tid0"},{64,8,-20,0,0,"64x8 neg"},{164,8,0,0,0,"164x8 pos"},{164,128,0,0,0,"164x128 pos"},{164,128,-20,0,0,"164x128 neg"},{164,128,-20,0,1,"164x128 neg elect"},{256,64,-3,0,0,"256x64 neg"},{132,128,4000,0,0,"132x128 tail"}};
  for(auto& c: cfgs){
    void* fp=nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&fp,cudaEnableDefault,&q);
    EncodeFn enc=(EncodeFn)fp;
    const int N=4412, C=5; float* x; cudaMalloc(&x,(size_t)N*C*4);
    std::vector<float> h((size_t)N*C); for(size_t i=0;i<h.size();i++) h[i]=(float)(i%N)+1000.f*(i/N);
    cudaMemcpy(x,h.data(),h.size()*4,cudaMemcpyHostToDevice);
    CUtensorMap tm; memset(&tm,0,sizeof(tm));
    cuuint64_t dims[2]={(cuuint64_t)N,(cuuint64_t)C}; cuuint64_t strides[1]={(cuuint64_t)N*4}; cuuint32_t box[2]={(cuuint32_t)c.box0,(cuuint32_t)c.box1}; cuuint32_t es[2]={1,1};
    CUresult r=enc(&tm,CU_TENSOR_MAP_DATA_TYPE_FLOAT32,2,x,dims,strides,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_L2_256B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int bytes=c.box0*c.box1*4; float* out; cudaMalloc(&out,bytes);
    cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,bytes);
    k<<<1,128,bytes>>>(tm,c.c0,c.c1,bytes,out,c.box0*c.box1,c.elect);
    cudaError_t e=cudaDeviceSynchronize();
    std::vector<float> o((size_t)c.box0*c.box1,-1.f);
    if(e==cudaSuccess) cudaMemcpy(o.data(),out,bytes,cudaMemcpyDeviceToHost);
    printf("%-22s encode=%d run=%s  s[0]=%g s[20]=%g s[box0]=%g s[box0+21]=%g\n",c.name,(int)r,cudaGetErrorString(e),o[0],o[20],o[c.box0],o[c.box0+21]);
    cudaDeviceReset();
  }
  return 0;
}
