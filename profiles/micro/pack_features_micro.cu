// Stand-alone timing of the fp32 NCHW -> bf16 pixel-major feature pack (8 x 2048 x 64 x 128, 537 MB in, 268 MB out) in several forms:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/pf profiles/micro/pack_features_micro.cu && /tmp/pf
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

// V0: the library kernel (64 ch x 64 px tile, scalar loads, bf16x2 stores, bounds checks everywhere)
__global__ void __launch_bounds__(256) v0(const float* __restrict__ x, int Cin, int hw, __nv_bfloat16* __restrict__ Xp) {
  __shared__ float tile[64][65];
  const int s0 = blockIdx.x * 64, c0 = blockIdx.y * 64, n = blockIdx.z;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const float* src = x + ((long long)n * Cin + c0) * hw + s0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int ci = ty + i * 4;
    tile[ci][tx] = (c0 + ci < Cin && s0 + tx < hw) ? __ldcs(src + (long long)ci * hw + tx) : 0.f;
  }
  __syncthreads();
  const int cpair = threadIdx.x & 31, prow = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = prow + i * 8;
    if (s0 + s < hw && c0 + 2 * cpair < Cin) {
      const __nv_bfloat162 v = __floats2bfloat162_rn(tile[2 * cpair][s], tile[2 * cpair + 1][s]);
      *reinterpret_cast<__nv_bfloat162*>(Xp + ((long long)n * hw + s0 + s) * Cin + c0 + 2 * cpair) = v;
    }
  }
}

// V1: same tile, interior fast path without bounds checks, 32-bit offsets from per-block base pointers
__global__ void __launch_bounds__(256) v1(const float* __restrict__ x, int Cin, int hw, __nv_bfloat16* __restrict__ Xp) {
  __shared__ float tile[64][65];
  const int s0 = blockIdx.x * 64, c0 = blockIdx.y * 64, n = blockIdx.z;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const float* src = x + ((long long)n * Cin + c0 + ty) * hw + s0 + tx;
  const int step = 4 * hw;
#pragma unroll
  for (int i = 0; i < 16; ++i) tile[ty + i * 4][tx] = __ldcs(src + i * step);
  __syncthreads();
  const int cpair = threadIdx.x & 31, prow = threadIdx.x >> 5;
  __nv_bfloat16* dst = Xp + ((long long)n * hw + s0 + prow) * Cin + c0 + 2 * cpair;
  const int dstep = 8 * Cin;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = prow + i * 8;
    *reinterpret_cast<__nv_bfloat162*>(dst + i * dstep) = __floats2bfloat162_rn(tile[2 * cpair][s], tile[2 * cpair + 1][s]);
  }
}

// V2: 128 ch x 64 px tile (256-byte output rows), float4 loads along pixels, 8-byte stores (4 channels), interior only
__global__ void __launch_bounds__(256) v2(const float* __restrict__ x, int Cin, int hw, __nv_bfloat16* __restrict__ Xp) {
  __shared__ float tile[128][65];
  const int s0 = blockIdx.x * 64, c0 = blockIdx.y * 128, n = blockIdx.z;
  const int q = threadIdx.x & 15, r = threadIdx.x >> 4;                  // 16 float4 per channel row, 16 rows per pass
  const float4* src = reinterpret_cast<const float4*>(x + ((long long)n * Cin + c0 + r) * hw + s0) + q;
  const int step = 16 * hw / 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 v = __ldcs(src + i * step);
    float* t = &tile[r + i * 16][4 * q];
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
  }
  __syncthreads();
  const int cq = threadIdx.x & 31, prow = threadIdx.x >> 5;              // 32 channel quads x 8 pixel rows
  __nv_bfloat16* dst = Xp + ((long long)n * hw + s0 + prow) * Cin + c0 + 4 * cq;
  const int dstep = 8 * Cin;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = prow + i * 8;
    const __nv_bfloat162 a = __floats2bfloat162_rn(tile[4 * cq][s], tile[4 * cq + 1][s]);
    const __nv_bfloat162 b = __floats2bfloat162_rn(tile[4 * cq + 2][s], tile[4 * cq + 3][s]);
    uint2 w;
    w.x = *reinterpret_cast<const uint32_t*>(&a);
    w.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i * dstep) = w;
  }
}

// V3: V1 with two 64 x 64 tiles per block along the pixel axis (loads of tile 2 in flight while tile 1 is stored)
__global__ void __launch_bounds__(256) v3(const float* __restrict__ x, int Cin, int hw, __nv_bfloat16* __restrict__ Xp) {
  __shared__ float tile[2][64][65];
  const int c0 = blockIdx.y * 64, n = blockIdx.z;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int cpair = threadIdx.x & 31, prow = threadIdx.x >> 5;
  const int step = 4 * hw, dstep = 8 * Cin;
  float v[16];
  const float* src = x + ((long long)n * Cin + c0 + ty) * hw + blockIdx.x * 128 + tx;
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __ldcs(src + i * step);
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) tile[h2][ty + i * 4][tx] = v[i];
    if (h2 == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __ldcs(src + 64 + i * step);
    }
    __syncthreads();
    __nv_bfloat16* dst = Xp + ((long long)n * hw + blockIdx.x * 128 + h2 * 64 + prow) * Cin + c0 + 2 * cpair;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int s = prow + i * 8;
      *reinterpret_cast<__nv_bfloat162*>(dst + i * dstep) = __floats2bfloat162_rn(tile[h2][2 * cpair][s], tile[h2][2 * cpair + 1][s]);
    }
  }
}

int main() {
  const int N = 8, Cin = 2048, hw = 8192;
  float* x; cudaMalloc(&x, (size_t)N * Cin * hw * 4); cudaMemset(x, 0, (size_t)N * Cin * hw * 4);
  __nv_bfloat16* Xp; cudaMalloc(&Xp, (size_t)N * Cin * hw * 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int variant = 0; variant < 4; ++variant)
    for (int rep = 0; rep < 2; ++rep) {
      float tot = 0;
      for (int it = 0; it < 8; ++it) {
        cudaEventRecord(e0);
        if (variant == 0) v0<<<dim3(hw / 64, Cin / 64, N), 256>>>(x, Cin, hw, Xp);
        if (variant == 1) v1<<<dim3(hw / 64, Cin / 64, N), 256>>>(x, Cin, hw, Xp);
        if (variant == 2) v2<<<dim3(hw / 64, Cin / 128, N), 256>>>(x, Cin, hw, Xp);
        if (variant == 3) v3<<<dim3(hw / 128, Cin / 64, N), 256>>>(x, Cin, hw, Xp);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) tot += ms;
      }
      printf("variant V%d rep %d: %.1f us  (%.2f TB/s)  (%s)\n", variant, rep, tot / 6 * 1e3, 805.3e6 / (tot / 6 * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
