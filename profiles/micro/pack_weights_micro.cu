// Stand-alone timing of the head's weight pack (4 x [19,2048,3,3] fp32 -> Wp [640,2048] bf16, WpT [2048,640] bf16) in two forms:
//   A: the library kernel (8 input channels per block: 16-byte segments of every Wp row)
//   B: 16 input channels per block (32-byte = whole-sector segments of Wp rows), table and centre sums precomputed in shared memory,
//      no integer divisions in the load loop
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/pw profiles/micro/pack_weights_micro.cu && /tmp/pw
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
struct PtrList { const float* p[8]; };

template <int PW_CI>
__device__ __forceinline__ float packed_weight(const float* __restrict__ w_s, const int* __restrict__ tab, int R, int C, int j, int ci_l) {
  const int e = tab[j];
  if (e >= 0) return w_s[e + ci_l * 9];
  float v = 0.f;
  if (e != INT_MIN) {
    const int c = -1 - e;
    for (int r = 0; r < R; ++r) v += w_s[((r * C + c) * PW_CI + ci_l) * 9 + 4];
  }
  return v;
}

template <int PW_CI, bool NODIV>
__global__ void __launch_bounds__(256) pack_kernel(PtrList w, int R, int C, int Cin, int NJ, __nv_bfloat16* __restrict__ Wp,
                                                   __nv_bfloat16* __restrict__ WpT) {
  extern __shared__ float pw_s[];
  int* tab = reinterpret_cast<int*>(pw_s + R * C * PW_CI * 9);
  const int ci0 = blockIdx.x * PW_CI;
  for (int j = threadIdx.x; j < NJ; j += 256) {
    const int t = j / C, c = j - t * C;
    int e = INT_MIN;
    if (t < 8 * R) { const int r = t >> 3, q = t & 7; e = ((r * C + c) * PW_CI) * 9 + (q < 4 ? q : q + 1); }
    else if (t == 8 * R) e = -1 - c;
    tab[j] = e;
  }
  constexpr int SEG = PW_CI * 9;
  if (NODIV) {
    // one (branch, class) segment of SEG contiguous floats per warp trip: no divisions
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rc = warp; rc < R * C; rc += 8) {
      const int r = rc / C, c = rc - r * C;
      const float* src = w.p[r] + ((long long)c * Cin + ci0) * 9;
      for (int off = lane; off < SEG; off += 32) pw_s[rc * SEG + off] = __ldg(src + off);
    }
  } else {
    for (int i = threadIdx.x; i < R * C * SEG; i += 256) {
      const int rc = i / SEG, off = i - rc * SEG;
      const int r = rc / C, c = rc - r * C;
      pw_s[i] = __ldg(w.p[r] + ((long long)c * Cin + ci0) * 9 + off);
    }
  }
  __syncthreads();
  const int half = NJ >> 1;
  for (int ci_l = 0; ci_l < PW_CI; ++ci_l)
    for (int jp = threadIdx.x; jp < half; jp += 256) {
      const int j = 2 * jp;
      const __nv_bfloat162 v = __floats2bfloat162_rn(packed_weight<PW_CI>(pw_s, tab, R, C, j, ci_l), packed_weight<PW_CI>(pw_s, tab, R, C, j + 1, ci_l));
      *reinterpret_cast<__nv_bfloat162*>(WpT + (long long)(ci0 + ci_l) * NJ + j) = v;
    }
  for (int idx = threadIdx.x; idx < NJ * (PW_CI / 8); idx += 256) {
    const int j = idx / (PW_CI / 8), seg = idx - j * (PW_CI / 8);
    uint32_t w4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 v = __floats2bfloat162_rn(packed_weight<PW_CI>(pw_s, tab, R, C, j, seg * 8 + 2 * e), packed_weight<PW_CI>(pw_s, tab, R, C, j, seg * 8 + 2 * e + 1));
      w4[e] = *reinterpret_cast<const uint32_t*>(&v);
    }
    *reinterpret_cast<uint4*>(Wp + (long long)j * Cin + ci0 + seg * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
}

template <int PW_CI, bool NODIV>
static void run(const char* name, PtrList w, int R, int C, int Cin, int NJ, __nv_bfloat16* Wp, __nv_bfloat16* WpT, float* flush) {
  const size_t smem = (size_t)R * C * PW_CI * 9 * 4 + NJ * 4;
  cudaFuncSetAttribute(pack_kernel<PW_CI, NODIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int cold = 0; cold < 2; ++cold) {
    float tot = 0;
    for (int it = 0; it < 12; ++it) {
      if (cold) cudaMemsetAsync(flush, it, 512 << 20);
      cudaEventRecord(e0);
      pack_kernel<PW_CI, NODIV><<<Cin / PW_CI, 256, smem>>>(w, R, C, Cin, NJ, Wp, WpT);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 2) tot += ms;
    }
    printf("%s, %s: %.2f us (%s)\n", name, cold ? "cold (L2 flushed)" : "warm (weights in L2)", tot / 10 * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  const int R = 4, C = 19, Cin = 2048, NJ = 640;
  PtrList w = {};
  for (int r = 0; r < R; ++r) { float* p; cudaMalloc(&p, (size_t)C * Cin * 9 * 4); cudaMemset(p, 0, (size_t)C * Cin * 9 * 4); w.p[r] = p; }
  __nv_bfloat16 *Wp, *WpT; cudaMalloc(&Wp, (size_t)NJ * Cin * 2); cudaMalloc(&WpT, (size_t)NJ * Cin * 2);
  float* flush; cudaMalloc(&flush, 512 << 20);
  run<8, false>("A  8 ci / block (library)", w, R, C, Cin, NJ, Wp, WpT, flush);
  run<8, true>("A' 8 ci / block, no divisions in the load loop", w, R, C, Cin, NJ, Wp, WpT, flush);
  run<16, true>("B 16 ci / block, no divisions", w, R, C, Cin, NJ, Wp, WpT, flush);
  run<32, true>("C 32 ci / block, no divisions", w, R, C, Cin, NJ, Wp, WpT, flush);
  return 0;
}
