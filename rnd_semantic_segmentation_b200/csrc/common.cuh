// Shared helpers for the b200seg kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define B200SEG_OK 0
#define B200SEG_ERR_ARG 1
#define B200SEG_ERR_CUDA 2
#define B200SEG_ERR_UNSUPPORTED 3

namespace b200seg {

void set_error(const char* fmt, ...);           // api.cu
int num_sms();                                   // api.cu (cached per process)
void count_launch();                             // api.cu: kernels launched by this library (bench evidence)
// api.cu: when profiling is on, bracket a launch with CUDA events on its stream under `tag`
void profile_begin(int tag, cudaStream_t stream);
void profile_end(int tag, cudaStream_t stream);

#define B200SEG_CHECK_ARG(cond, ...)                                  \
  do {                                                                \
    if (!(cond)) {                                                    \
      ::b200seg::set_error(__VA_ARGS__);                              \
      return B200SEG_ERR_ARG;                                         \
    }                                                                 \
  } while (0)

#define B200SEG_CUDA(call)                                                              \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::b200seg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                           __FILE__, __LINE__);                                         \
      return B200SEG_ERR_CUDA;                                                          \
    }                                                                                   \
  } while (0)

#define B200SEG_LAUNCH_CHECK()             \
  do {                                     \
    ::b200seg::count_launch();             \
    B200SEG_CUDA(cudaGetLastError());      \
  } while (0)

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- align_corners=True bilinear taps, the ATen way -----------------------------------------
// (ATen/native/UpSample.h area_pixel_compute_scale / area_pixel_compute_source_index with
//  align_corners=true: scale = (float)(in-1)/(out-1) [0 when out<=1]; src = scale*dst;
//  i0 = (int)src; i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1.)
__host__ __device__ __forceinline__ float ac_scale(int in_size, int out_size) {
  return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.0f;
}
struct Tap {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Tap ac_tap(float scale, int dst, int in_size) {
  Tap t;
  const float src = scale * (float)dst;
  t.i0 = (int)src;
  t.i1 = t.i0 + ((t.i0 < in_size - 1) ? 1 : 0);
  t.l1 = src - (float)t.i0;
  t.l0 = 1.0f - t.l1;
  return t;
}
// First output index whose i0 is >= cell (monotone in dst); returns out_size if none.
__device__ __forceinline__ int ac_first_dst(float scale, int cell, int in_size, int out_size) {
  if (cell <= 0) return 0;
  if (scale <= 0.0f) return out_size;
  int d = (int)((float)cell / scale);
  if (d > out_size) d = out_size;
  if (d < 0) d = 0;
  while (d > 0 && (int)(scale * (float)(d - 1)) >= cell) --d;
  while (d < out_size && (int)(scale * (float)d) < cell) ++d;
  return d;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- warp / block reductions ----------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- streaming 128-bit loads/stores (read-once data: bypass L1 allocation) --------------------
__device__ __forceinline__ int4 ld_stream_v4(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ long long ld_stream_s64(const void* p) {
  long long r;
  asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(void* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace b200seg
