// Stand-alone timing of the split-K reduction of the head's weight gradient (variants of wgrad_reduce_kernel):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/wr profiles/micro/wgrad_reduce_micro.cu && /tmp/wr
#include <cstdio>
#include <cuda_runtime.h>
struct MutPtrList { float* p[8]; };

// A: as in the library (one ci per thread, 9 taps, transposed through smem)
__global__ void __launch_bounds__(256) reduce_a(const float* __restrict__ part, int S, long long slab, int R, int C, int Cin, MutPtrList gw) {
  __shared__ __align__(16) float sm[256 * 9];
  const int ci0 = blockIdx.x * 256, ci = ci0 + threadIdx.x, c = blockIdx.y, r = blockIdx.z;
  const float* src[9];
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int t = (k == 4) ? 8 * R : r * 8 + (k < 4 ? k : k - 1);
    src[k] = part + (long long)(t * C + c) * Cin + ci;
    acc[k] = 0.f;
  }
#pragma unroll 2
  for (int s = 0; s < S; ++s) {
    float v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = __ldcs(src[k] + s * slab);
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] += v[k];
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) sm[threadIdx.x * 9 + k] = acc[k];
  __syncthreads();
  float* dst = gw.p[r] + ((long long)c * Cin + ci0) * 9;
  for (int i = threadIdx.x; i < 256 * 9 / 4; i += 256) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(sm)[i];
}

// B: 64 ci per block, 4 ci per thread as float4 loads, S splits unrolled by template
template <int S>
__global__ void __launch_bounds__(144) reduce_b(const float* __restrict__ part, long long slab, int R, int C, int Cin, MutPtrList gw) {
  // block = (64 ci) x (9 taps): thread = (tap k, group of 4 ci): 9 * 16 = 144 threads
  __shared__ __align__(16) float sm[64 * 9];
  const int ci0 = blockIdx.x * 64, c = blockIdx.y, r = blockIdx.z;
  const int k = threadIdx.x / 16, q = threadIdx.x % 16;
  const int t = (k == 4) ? 8 * R : r * 8 + (k < 4 ? k : k - 1);
  const float4* src = reinterpret_cast<const float4*>(part + (long long)(t * C + c) * Cin + ci0) + q;
  float4 v[S];
#pragma unroll
  for (int s = 0; s < S; ++s) v[s] = __ldcs(src + s * (slab / 4));
  float4 a = v[0];
#pragma unroll
  for (int s = 1; s < S; ++s) { a.x += v[s].x; a.y += v[s].y; a.z += v[s].z; a.w += v[s].w; }
  sm[(4 * q + 0) * 9 + k] = a.x; sm[(4 * q + 1) * 9 + k] = a.y; sm[(4 * q + 2) * 9 + k] = a.z; sm[(4 * q + 3) * 9 + k] = a.w;
  __syncthreads();
  float* dst = gw.p[r] + ((long long)c * Cin + ci0) * 9;
  if (threadIdx.x < 64 * 9 / 4) reinterpret_cast<float4*>(dst)[threadIdx.x] = reinterpret_cast<const float4*>(sm)[threadIdx.x];
}

int main() {
  const int S = 6, R = 4, C = 19, Cin = 2048, NJ = 640;
  const long long slab = (long long)NJ * Cin;
  float* part; cudaMalloc(&part, S * slab * 4); cudaMemset(part, 0, S * slab * 4);
  MutPtrList gw = {};
  for (int r = 0; r < R; ++r) cudaMalloc(&gw.p[r], (size_t)C * Cin * 9 * 4);
  float* flush; cudaMalloc(&flush, 512 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int variant = 0; variant < 2; ++variant)
    for (int cold = 0; cold < 2; ++cold) {
      float tot = 0;
      for (int it = 0; it < 12; ++it) {
        if (cold) cudaMemsetAsync(flush, it, 512 << 20); else cudaMemsetAsync(part, 0, S * slab * 4);
        cudaEventRecord(e0);
        if (variant == 0) reduce_a<<<dim3(Cin / 256, C, R), 256>>>(part, S, slab, R, C, Cin, gw);
        else reduce_b<S><<<dim3(Cin / 64, C, R), 144>>>(part, slab, R, C, Cin, gw);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) tot += ms;
      }
      printf("variant %c, %s: %.2f us   (%s)\n", 'A' + variant, cold ? "slabs cold (L2 flushed)" : "slabs just written (L2)", tot / 10 * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
