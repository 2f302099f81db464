// K1: the fused four-dilation ASPP classifier head (DeepLabV2), forward and backward.
//
// Replaces (reference file:line): ASPP_Classifier_V2.forward, core/models/classifiers/aspp/classifier.py:26-29
//   out = sum_r Conv2d(Cin, C, 3, padding=r, dilation=r)(x),  r in dilation_series (6/12/18/24)
// and its autograd backward (dgrad, wgrad, bias grad).
//
// Formulation ("tap-packed GEMM").  With T = 8R+1 distinct taps (the R centre taps hit the same
// pixel, so their weights are pre-summed) and offset d_t of tap t:
//     out[p, c] = bias[c] + sum_t sum_ci X[p + d_t, ci] * W[t, c, ci]
// Padding the class dimension to an MMA-friendly N (19 -> 32) would waste 41 % of the tensor
// pipe and leave it shared-memory-bound (a 128 x 32 x 16 MMA re-reads 4 KB of A for 16 cycles of
// math).  Instead (t, c) is packed into ONE GEMM dimension j = t*C + c  (33*19 = 627 -> 640):
//     fwd    Yt[j, p]   = sum_ci Wp[j, ci] * Xp[p, ci]          M=640  N=P     K=Cin   (K-major x K-major)
//            out[p, c]  = bias[c] + sum_t Yt[t*C+c, p + d_t]     shift-and-add gather, coalesced along p
//     G'[p, j]          = gO[p - d_t, c]                         bf16 "im2col of the output gradient"
//     dgrad  dX[ci, p]  = sum_j WpT[ci, j] * G'[p, j]            M=Cin  N=P     K=640   (writes fp32 NCHW directly)
//     wgrad  dWp[j, ci] = sum_p G'[p, j] * Xp[p, ci]             M=640  N=Cin   K=P     (MN-major x MN-major, split-K)
// All three GEMMs run on the tcgen05 core in gemm_sm100.cuh with 98 % useful MMA work at C=19.
// Operands are bf16 (features, packed weights, output gradient), accumulation is fp32 in TMEM.
#include "common.cuh"
#include "gemm_sm100.cuh"

namespace b200seg {

constexpr int MAX_RATES = 8;
constexpr int MAX_TAPS = 8 * MAX_RATES + 1;

struct TapTable {
  int n_taps;              // 8R + 1 (last = merged centre)
  int dy[MAX_TAPS], dx[MAX_TAPS];
};

static void make_taps(TapTable& tt, const int* rates, int R) {
  int t = 0;
  for (int r = 0; r < R; ++r)
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        if (ky == 1 && kx == 1) continue;
        tt.dy[t] = (ky - 1) * rates[r];
        tt.dx[t] = (kx - 1) * rates[r];
        ++t;
      }
  tt.dy[t] = 0; tt.dx[t] = 0;
  tt.n_taps = t + 1;
}

int aspp_nj(int C, int R) { return ceil_div((8 * R + 1) * C, 128) * 128; }

struct PtrList {
  const float* p[MAX_RATES];
};
struct MutPtrList {
  float* p[MAX_RATES];
};

// ------------------------------------------------------------------------------------------
// pack weights: R x [C, Cin, 3, 3] fp32 -> Wp [NJ, Cin] bf16, WpT [Cin, NJ] bf16, bias_sum [C]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weights_kernel(PtrList w, PtrList b, int R, int C, int Cin, int NJ,
                                                           __nv_bfloat16* __restrict__ Wp, __nv_bfloat16* __restrict__ WpT,
                                                           float* __restrict__ bias_sum) {
  const long long idx = blockIdx.x * 256LL + threadIdx.x;
  if (idx < C) {
    float s = b.p[0] ? b.p[0][idx] : 0.f;
    for (int r = 1; r < R; ++r) s += b.p[r] ? b.p[r][idx] : 0.f;        // ((b0+b1)+b2)+b3, the reference's order
    bias_sum[idx] = s;
  }
  if (idx >= (long long)NJ * Cin) return;
  const int j = (int)(idx / Cin), ci = (int)(idx - (long long)j * Cin);
  const int t = j / C, c = j - t * C;
  float v = 0.f;
  if (t < 8 * R) {
    const int r = t >> 3, q = t & 7;
    const int k = q < 4 ? q : q + 1;                                       // skip the centre (k == 4)
    v = w.p[r][((long long)c * Cin + ci) * 9 + k];
  } else if (t == 8 * R) {
    for (int r = 0; r < R; ++r) v += w.p[r][((long long)c * Cin + ci) * 9 + 4];
  }
  const __nv_bfloat16 h = __float2bfloat16(v);
  Wp[idx] = h;
  WpT[(long long)ci * NJ + j] = h;
}

// ------------------------------------------------------------------------------------------
// pack features: x fp32 NCHW [N, Cin, hw] -> Xp bf16 [N*hw, Cin]  (64 ci x 64 px tiles through smem)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_features_kernel(const float* __restrict__ x, int Cin, int hw,
                                                            __nv_bfloat16* __restrict__ Xp) {
  __shared__ float tile[64][65];
  const int s0 = blockIdx.x * 64, c0 = blockIdx.y * 64, n = blockIdx.z;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;          // 64 x 4
  const float* src = x + ((long long)n * Cin + c0) * hw + s0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int ci = ty + i * 4;
    tile[ci][tx] = (c0 + ci < Cin && s0 + tx < hw) ? __ldcs(src + (long long)ci * hw + tx) : 0.f;
  }
  __syncthreads();
  const int cpair = threadIdx.x & 31, prow = threadIdx.x >> 5;     // 32 channel pairs x 8 pixel rows
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = prow + i * 8;
    if (s0 + s < hw && c0 + 2 * cpair < Cin) {
      const __nv_bfloat162 v = __floats2bfloat162_rn(tile[2 * cpair][s], tile[2 * cpair + 1][s]);
      *reinterpret_cast<__nv_bfloat162*>(Xp + ((long long)n * hw + s0 + s) * Cin + c0 + 2 * cpair) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward gather: logits[n,c,y,x] = bias[c] + sum_t Yt[t*C+c][p(n, y+dy_t, x+dx_t)]
// ------------------------------------------------------------------------------------------
// One thread per low-res pixel, all classes in registers: per tap one bounds test and C loads at a fixed stride
// (coalesced along x across the warp), so the kernel issues ~3 instructions per Yt element and runs at the HBM rate
// of the Yt read instead of being issue-bound on per-element address arithmetic.
constexpr int GATHER_THREADS = 64;
template <int CT>
__global__ void __launch_bounds__(GATHER_THREADS) head_gather_kernel(const float* __restrict__ Yt, long long ypitch, TapTable tt,
                                                                     const float* __restrict__ bias_sum, int C, int nchunk,
                                                                     int h, int w, float* __restrict__ logits) {
  // blockIdx.z = image * nchunk + class chunk; this thread owns classes [c0, c0 + nc) of one pixel (nc <= CT)
  const int n = blockIdx.z / nchunk;
  const int c0 = (blockIdx.z - n * nchunk) * CT;
  const int nc = min(CT, C - c0);
  const int xq = blockIdx.x * GATHER_THREADS + threadIdx.x;
  const int y = blockIdx.y;
  if (xq >= w) return;
  const long long pbase = (long long)n * h * w;
  const long long cstep = ypitch * 4;                         // bytes between classes of one tap
  float acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = (c < nc) ? bias_sum[c0 + c] : 0.f;
  const char* tap_base = reinterpret_cast<const char*>(Yt + pbase + (long long)y * w + xq) + cstep * c0;
  const long long tstep = cstep * C;                          // bytes between taps
  constexpr int TG = 3;                                        // taps per batch: TG * CT loads in flight per thread
#pragma unroll 1
  for (int t0 = 0; t0 < tt.n_taps; t0 += TG, tap_base += TG * tstep) {
    float v[TG][CT];
#pragma unroll
    for (int k = 0; k < TG; ++k) {
      const int t = min(t0 + k, tt.n_taps - 1);
      const int dy = tt.dy[t], dx = tt.dx[t];
      const int yy = y + dy, xx = xq + dx;
      const bool inb = (t0 + k < tt.n_taps) && yy >= 0 && yy < h && xx >= 0 && xx < w;
      const char* src = tap_base + k * tstep + ((long long)dy * w + dx) * 4;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        v[k][c] = 0.f;
        if (inb && c < nc) v[k][c] = __ldcs(reinterpret_cast<const float*>(src));
        src += cstep;
      }
    }
#pragma unroll
    for (int k = 0; k < TG; ++k)
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[c] += v[k][c];
  }
  const long long hw = (long long)h * w;
  float* dst = logits + ((long long)n * C + c0) * hw + (long long)y * w + xq;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (c < nc) dst[c * hw] = acc[c];
}

// ------------------------------------------------------------------------------------------
// backward prep: gO fp32 NCHW [N,C,hw] -> gOt bf16 [P][32]; bias grad; G' bf16 [P][NJ]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_to_pixel_major_kernel(const float* __restrict__ g, int C, int hw, long long P,
                                                                  __nv_bfloat16* __restrict__ gOt) {
  const long long p = blockIdx.x * 256LL + threadIdx.x;
  if (p >= P) return;
  const long long n = p / hw, s = p - n * hw;
  uint32_t wds[16];
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    const float a = c < C ? g[(n * C + c) * hw + s] : 0.f;
    const float b = c + 1 < C ? g[(n * C + c + 1) * hw + s] : 0.f;
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    wds[c >> 1] = *reinterpret_cast<const uint32_t*>(&v);
  }
  int4* dst = reinterpret_cast<int4*>(gOt + p * 32);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    dst[i] = make_int4((int)wds[4 * i], (int)wds[4 * i + 1], (int)wds[4 * i + 2], (int)wds[4 * i + 3]);
}

// deterministic per-class sum over all pixels (bias gradient), one block per class
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ g, int N, int C, int hw, MutPtrList gb, int R) {
  __shared__ double red[8];
  const int c = blockIdx.x;
  double acc = 0.0;
  for (int n = 0; n < N; ++n) {
    const float* src = g + ((long long)n * C + c) * hw;
    for (int s = threadIdx.x; s < hw; s += 256) acc += (double)src[s];
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    for (int r = 0; r < R; ++r)
      if (gb.p[r]) gb.p[r][c] = (float)t;
  }
}

// G'[p][j] = gO[p - d_t][c]  (j = t*C + c; zero outside the image and for j >= T*C).
// One warp assembles one G' row (NJ bf16) in shared memory: lane t copies the C-element run of tap t from the 64-byte
// pixel-major gOt row (three 16-byte loads, 2-byte shared stores at the run's unaligned offset t*C), then the warp
// streams the finished row out with 16-byte coalesced stores.  ~70 warp instructions per pixel instead of ~80 per
// 16-byte vector: the kernel is bound by the G' write, not by issue.
constexpr int GP_WARPS = 8;          // warps (= rows in flight) per block
constexpr int GP_PX_PER_WARP = 8;    // consecutive pixels per warp
__device__ __forceinline__ void sts_u16(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }

template <int CT>                     // CT == C for the instantiated class counts, 0 = runtime C (<= 32)
__global__ void __launch_bounds__(GP_WARPS * 32) build_gprime_kernel(const __nv_bfloat16* __restrict__ gOt, TapTable tt, int C_rt,
                                                                     int NJ, int h, int w, long long P,
                                                                     __nv_bfloat16* __restrict__ Gp) {
  extern __shared__ __align__(16) uint8_t gp_smem[];
  const int C = CT ? CT : C_rt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ntap = tt.n_taps;
  const int row_bytes = NJ * 2;
  const unsigned row_s = smem_u32(gp_smem) + warp * row_bytes;
  const int hw = h * w;
  long long p = ((long long)blockIdx.x * GP_WARPS + warp) * GP_PX_PER_WARP;
  if (p >= P) return;
  const long long n0 = p / hw;
  int s = (int)(p - n0 * hw);
  int y = s / w, x = s - y * w;
  const uint4* img = reinterpret_cast<const uint4*>(gOt) + n0 * hw * 4;       // 4 x 16 B per pixel row

  // this lane's taps t = lane, lane + 32, lane + 64 held in registers (a lane-indexed read of the kernel-parameter
  // table would serialise in the constant cache); absent taps get an offset that is always outside the image
  constexpr int TS = (MAX_TAPS + 31) / 32;
  int my_dy[TS], my_dx[TS];
#pragma unroll
  for (int i = 0; i < TS; ++i) {
    const int t = lane + 32 * i;
    my_dy[i] = 1 << 20; my_dx[i] = 1 << 20;
#pragma unroll 1
    for (int q = 0; q < ntap; ++q)                             // uniform index -> plain constant loads
      if (q == t) { my_dy[i] = tt.dy[q]; my_dx[i] = tt.dx[q]; }
  }
  const int nvec = NJ >> 3;
  const int pad0 = ntap * C;

  for (int k = 0; k < GP_PX_PER_WARP && p < P; ++k, ++p) {
#pragma unroll
    for (int i = 0; i < TS; ++i) {
      const int t = lane + 32 * i;
      if (t < ntap) {
        const int yy = y - my_dy[i], xx = x - my_dx[i];
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = make_uint4(0, 0, 0, 0);
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
          const uint4* src = img + ((long long)yy * w + xx) * 4;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (u * 8 < C) q[u] = __ldg(src + u);
        }
        const unsigned wd[16] = {q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, q[1].z, q[1].w,
                                 q[2].x, q[2].y, q[2].z, q[2].w, q[3].x, q[3].y, q[3].z, q[3].w};
        const unsigned dst = row_s + (unsigned)(t * C) * 2;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < C) sts_u16(dst + c * 2, (c & 1) ? (wd[c >> 1] >> 16) : wd[c >> 1]);
      }
    }
    for (int j = pad0 + lane; j < NJ; j += 32) sts_u16(row_s + j * 2, 0u);              // padding columns
    __syncwarp();
    uint4* out = reinterpret_cast<uint4*>(Gp + p * NJ);
    for (int v = lane; v < nvec; v += 32) {
      uint4 r;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(row_s + v * 16));
      out[v] = r;
    }
    __syncwarp();
    if (++x == w) { x = 0; if (++y == h) { y = 0; img += (long long)hw * 4; } }
  }
}

// ------------------------------------------------------------------------------------------
// wgrad epilogue: sum split-K partials [S][NJ][Cin] and scatter to the R conv weight grads [C,Cin,3,3]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int S, long long slab, int R, int C,
                                                           int Cin, MutPtrList gw) {
  const int ci = blockIdx.x * 256 + threadIdx.x;
  const int c = blockIdx.y, r = blockIdx.z;
  if (ci >= Cin || !gw.p[r]) return;
  float out[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int t = (k == 4) ? 8 * R : r * 8 + (k < 4 ? k : k - 1);
    const long long off = (long long)(t * C + c) * Cin + ci;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += part[s * slab + off];
    out[k] = acc;
  }
  float* dst = gw.p[r] + ((long long)c * Cin + ci) * 9;
#pragma unroll
  for (int k = 0; k < 9; ++k) dst[k] = out[k];
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
int aspp_pack_weights(const float* const* w, const float* const* b, int R, int C, int Cin, void* Wp, void* WpT,
                      float* bias_sum, cudaStream_t stream) {
  B200SEG_CHECK_ARG(R >= 1 && R <= MAX_RATES, "aspp: %d dilation branches unsupported (1..%d)", R, MAX_RATES);
  B200SEG_CHECK_ARG(w && Wp && WpT && bias_sum, "aspp_pack_weights: null pointer");
  B200SEG_CHECK_ARG(Cin % 8 == 0, "aspp: in_channels=%d must be a multiple of 8", Cin);
  PtrList pw, pb;
  for (int r = 0; r < MAX_RATES; ++r) { pw.p[r] = r < R ? w[r] : nullptr; pb.p[r] = (r < R && b) ? b[r] : nullptr; }
  const int NJ = aspp_nj(C, R);
  const long long total = (long long)NJ * Cin;
  pack_weights_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(pw, pb, R, C, Cin, NJ, (__nv_bfloat16*)Wp,
                                                                               (__nv_bfloat16*)WpT, bias_sum);
  B200SEG_LAUNCH_CHECK();
  return B200SEG_OK;
}

int aspp_pack_features(const float* x, int N, int Cin, int h, int w, void* Xp, cudaStream_t stream) {
  B200SEG_CHECK_ARG(x && Xp && N > 0 && h > 0 && w > 0, "aspp_pack_features: bad arguments");
  B200SEG_CHECK_ARG(Cin % 8 == 0, "aspp: in_channels=%d must be a multiple of 8", Cin);
  dim3 grid(ceil_div(h * w, 64), ceil_div(Cin, 64), N);
  profile_begin(3, stream);
  pack_features_kernel<<<grid, 256, 0, stream>>>(x, Cin, h * w, (__nv_bfloat16*)Xp);
  profile_end(3, stream);
  B200SEG_LAUNCH_CHECK();
  return B200SEG_OK;
}

long long aspp_yt_bytes(int N, int C, int h, int w, int R) {
  const long long P = (long long)N * h * w;
  return (long long)aspp_nj(C, R) * (ceil_div_ll(P, 4) * 4) * 4;
}

int aspp_forward(const void* Xp, const void* Wp, const float* bias_sum, const int* rates, int R, int N, int Cin, int C, int h,
                 int w, float* Yt, float* logits, cudaStream_t stream) {
  B200SEG_CHECK_ARG(Xp && Wp && bias_sum && rates && Yt && logits, "aspp_forward: null pointer");
  B200SEG_CHECK_ARG(R >= 1 && R <= MAX_RATES, "aspp: %d dilation branches unsupported", R);
  const long long P = (long long)N * h * w;
  B200SEG_CHECK_ARG(P < (1LL << 31), "aspp_forward: too many pixels");
  const int NJ = aspp_nj(C, R);
  const long long ypitch = ceil_div_ll(P, 4) * 4;
  gemm::Operand a{(const __nv_bfloat16*)Wp, false, Cin};
  gemm::Operand b{(const __nv_bfloat16*)Xp, false, Cin};
  int rc = gemm::launch(a, b, NJ, (int)P, Cin, 1, Yt, ypitch, 0, 0, 0, stream, nullptr, 0, gemm::SHARE_A);
  if (rc) return rc;
  TapTable tt;
  make_taps(tt, rates, R);
  // class chunks of CT per thread: enough threads in flight to cover the HBM latency of the strided Yt reads
  const int CT = (C == 19) ? 5 : (C <= 4 ? 4 : 8);
  const int nchunk = ceil_div(C, CT);
  dim3 grid(ceil_div(w, GATHER_THREADS), h, N * nchunk);
  profile_begin(4, stream);
  if (CT == 5) head_gather_kernel<5><<<grid, GATHER_THREADS, 0, stream>>>(Yt, ypitch, tt, bias_sum, C, nchunk, h, w, logits);
  else if (CT == 4) head_gather_kernel<4><<<grid, GATHER_THREADS, 0, stream>>>(Yt, ypitch, tt, bias_sum, C, nchunk, h, w, logits);
  else head_gather_kernel<8><<<grid, GATHER_THREADS, 0, stream>>>(Yt, ypitch, tt, bias_sum, C, nchunk, h, w, logits);
  profile_end(4, stream);
  B200SEG_LAUNCH_CHECK();
  return B200SEG_OK;
}

// scratch: gOt [P][32] bf16 | Gp [P][NJ] bf16 | wpart [S][NJ][Cin] fp32
long long aspp_bwd_scratch_bytes(int N, int Cin, int C, int h, int w, int R, int splits) {
  const long long P = (long long)N * h * w;
  const int NJ = aspp_nj(C, R);
  return P * 32 * 2 + 256 + P * NJ * 2 + 256 + (long long)splits * NJ * Cin * 4 + 256;
}

// scratch layout: gOt [P][32] bf16 | Gp [P][NJ] bf16 | wpart [S][NJ][Cin] fp32
static void aspp_bwd_carve(void* scratch, long long P, int NJ, __nv_bfloat16** gOt, __nv_bfloat16** Gp, float** wpart) {
  uint8_t* sp = reinterpret_cast<uint8_t*>(scratch);
  *gOt = reinterpret_cast<__nv_bfloat16*>(sp);
  sp += (P * 32 * 2 + 255) / 256 * 256;
  *Gp = reinterpret_cast<__nv_bfloat16*>(sp);
  sp += (P * NJ * 2 + 255) / 256 * 256;
  *wpart = reinterpret_cast<float*>(sp);
}

void* aspp_bwd_gOt_ptr(void* scratch) { return scratch; }

// backward GEMMs from the packed bf16 pixel-major output gradient gOt [P][32]
int aspp_backward_packed(const void* gOt_in, const void* Xp, const void* WpT, const int* rates, int R, int N, int Cin, int C,
                         int h, int w, void* scratch, long long scratch_bytes, int splits, float* grad_x, float* const* grad_w,
                         cudaStream_t stream) {
  B200SEG_CHECK_ARG(gOt_in && Xp && WpT && rates && scratch, "aspp_backward: null pointer");
  B200SEG_CHECK_ARG(R >= 1 && R <= MAX_RATES, "aspp: %d dilation branches unsupported", R);
  B200SEG_CHECK_ARG(C <= 32, "aspp_backward: num_classes=%d > 32 is not supported", C);
  if (splits < 1) splits = 1;
  B200SEG_CHECK_ARG(scratch_bytes >= aspp_bwd_scratch_bytes(N, Cin, C, h, w, R, splits), "aspp_backward: scratch too small");
  const long long P = (long long)N * h * w;
  const int NJ = aspp_nj(C, R);
  const int hw = h * w;
  __nv_bfloat16 *gOt_own, *Gp;
  float* wpart;
  aspp_bwd_carve(scratch, P, NJ, &gOt_own, &Gp, &wpart);
  const __nv_bfloat16* gOt = reinterpret_cast<const __nv_bfloat16*>(gOt_in);
  TapTable tt;
  make_taps(tt, rates, R);
  {
    const unsigned blocks = (unsigned)ceil_div_ll(P, (long long)GP_WARPS * GP_PX_PER_WARP);
    const size_t smem = (size_t)GP_WARPS * NJ * 2;
    profile_begin(5, stream);
    if (C == 19) build_gprime_kernel<19><<<blocks, GP_WARPS * 32, smem, stream>>>(gOt, tt, C, NJ, h, w, P, Gp);
    else build_gprime_kernel<0><<<blocks, GP_WARPS * 32, smem, stream>>>(gOt, tt, C, NJ, h, w, P, Gp);
    profile_end(5, stream);
    B200SEG_LAUNCH_CHECK();
  }
  if (grad_x) {
    // dX[ci, p] = WpT[ci, :] . G'[p, :]   -> fp32 NCHW: column p = (image, pixel), row = channel
    gemm::Operand a{(const __nv_bfloat16*)WpT, false, NJ};
    gemm::Operand b{Gp, false, NJ};
    int rc = gemm::launch(a, b, Cin, (int)P, NJ, 1, grad_x, hw, hw, (long long)Cin * hw, 0, stream, nullptr, 1, gemm::SHARE_B);
    if (rc) return rc;
  }
  if (grad_w) {
    MutPtrList gw;
    bool any = false;
    for (int r = 0; r < MAX_RATES; ++r) { gw.p[r] = r < R ? grad_w[r] : nullptr; any |= gw.p[r] != nullptr; }
    if (any) {
      // dWp[j, ci] = sum_p G'[p, j] * Xp[p, ci]   both operands MN-major, split-K over pixels
      gemm::Operand a{Gp, true, NJ};
      gemm::Operand b{(const __nv_bfloat16*)Xp, true, Cin};
      int used = 1;
      const long long slab = (long long)NJ * Cin;
      int rc = gemm::launch(a, b, NJ, Cin, (int)P, splits, wpart, Cin, 0, 0, slab, stream, &used, 2, gemm::SHARE_A);
      if (rc) return rc;
      dim3 grid(ceil_div(Cin, 256), C, R);
      profile_begin(10, stream);
      wgrad_reduce_kernel<<<grid, 256, 0, stream>>>(wpart, used, slab, R, C, Cin, gw);
      profile_end(10, stream);
      B200SEG_LAUNCH_CHECK();
    }
  }
  return B200SEG_OK;
}

// backward from the fp32 NCHW output gradient (generic autograd contract)
int aspp_backward(const float* grad_logits, const void* Xp, const void* WpT, const int* rates, int R, int N, int Cin, int C,
                  int h, int w, void* scratch, long long scratch_bytes, int splits, float* grad_x, float* const* grad_w,
                  float* const* grad_b, cudaStream_t stream) {
  B200SEG_CHECK_ARG(grad_logits && scratch, "aspp_backward: null pointer");
  B200SEG_CHECK_ARG(R >= 1 && R <= MAX_RATES, "aspp: %d dilation branches unsupported", R);
  B200SEG_CHECK_ARG(C <= 32, "aspp_backward: num_classes=%d > 32 is not supported", C);
  if (splits < 1) splits = 1;
  B200SEG_CHECK_ARG(scratch_bytes >= aspp_bwd_scratch_bytes(N, Cin, C, h, w, R, splits), "aspp_backward: scratch too small");
  const long long P = (long long)N * h * w;
  const int hw = h * w;
  __nv_bfloat16* gOt = reinterpret_cast<__nv_bfloat16*>(scratch);
  grad_to_pixel_major_kernel<<<(unsigned)ceil_div_ll(P, 256), 256, 0, stream>>>(grad_logits, C, hw, P, gOt);
  B200SEG_LAUNCH_CHECK();
  if (grad_b) {
    MutPtrList gb;
    bool any = false;
    for (int r = 0; r < MAX_RATES; ++r) { gb.p[r] = r < R ? grad_b[r] : nullptr; any |= gb.p[r] != nullptr; }
    if (any) {
      bias_grad_kernel<<<C, 256, 0, stream>>>(grad_logits, N, C, hw, gb, R);
      B200SEG_LAUNCH_CHECK();
    }
  }
  return aspp_backward_packed(gOt, Xp, WpT, rates, R, N, Cin, C, h, w, scratch, scratch_bytes, splits, grad_x, grad_w, stream);
}

}  // namespace b200seg
