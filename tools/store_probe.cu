// How fast can a persistent kernel write y[1024][480000] fp32 in [128 outputs x 256 channels] tiles?
// Variants of the store shape, to separate LSU issue limits from DRAM write locality.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/store_probe.cu -o tools/bin/store_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int kCh = 1024, kN = 480000, kTM = 128, kTN = 256;

// mode 0: 4 warps, warp w lane l writes y[c][m0 + 32w + l] for c = 0..255 (4-byte lanes, 128 B per warp store)
// mode 1: 4 warps, warp w handles channels c = w, w+4, ...; lane l writes float4 y[c][m0 + 4l .. +3] (512 B per warp store)
// mode 2: as mode 0 but 16 warps (each warp a 64-channel slice)
// mode 3: as mode 1 but 16 warps
__global__ void __launch_bounds__(512) probe(float* y, int mode, int n_tt, long long n_tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long ct = tile / n_tt, tt = tile - ct * n_tt;
    const long long m0 = tt * kTM;
    if (mode == 0 || mode == 2) {
      const int w4 = warp & 3, slice = warp >> 2, nslice = nw >> 2;
      const long long m = m0 + w4 * 32 + lane;
      if (m < kN)
        for (int c = slice; c < kTN; c += nslice) y[(ct * kTN + c) * (long long)kN + m] = 1.0f;
    } else {
      const long long m = m0 + 4 * lane;
      if (m < kN)
        for (int c = warp; c < kTN; c += nw)
          *reinterpret_cast<float4*>(&y[(ct * kTN + c) * (long long)kN + m]) = make_float4(1.f, 2.f, 3.f, 4.f);
    }
  }
}

int main() {
  float* y;
  cudaMalloc(&y, (size_t)kCh * kN * 4);
  const int n_tt = (kN + kTM - 1) / kTM;
  const long long n_tiles = (long long)n_tt * (kCh / kTN);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct { int mode, threads; const char* name; } cfg[] = {
      {0, 128, "4 warps, 128 B per warp-store (4 B lanes)"}, {1, 128, "4 warps, 512 B per warp-store (float4 lanes)"},
      {2, 512, "16 warps, 128 B per warp-store"}, {3, 512, "16 warps, 512 B per warp-store"}};
  for (auto& c : cfg) {
    for (int rep = 0; rep < 2; ++rep) probe<<<148, c.threads>>>(y, c.mode, n_tt, n_tiles);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 5; ++rep) probe<<<148, c.threads>>>(y, c.mode, n_tt, n_tiles);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("%-48s %.3f ms  %.0f GB/s\n", c.name, ms, (double)kCh * kN * 4 / ms / 1e6);
  }
  return 0;
}
