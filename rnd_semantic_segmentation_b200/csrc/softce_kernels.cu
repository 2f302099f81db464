// K3: soft-label cross-entropy on materialised operands (the reference's API-level contract).
//
// Replaces (reference file:line):
//   soft_label_cross_entropy(pred, soft_label, pixel_weights=None)   core/utils/utility.py:172-177
//     loss = mean_{N,H,W}( [w *] sum_c( -q_c * log_softmax(p)_c ) )
//     d loss / d p_c = ( softmax(p)_c * sum_k q_k - q_c ) * w / (N*H*W)
//   callers: core/combos/aspp_fada.py:111,120,124 on [N, 2C, H, W] fp32 operands.
//
// Pure HBM streaming: forward reads p and q once (8*K B/px), backward reads p and q and writes the
// gradient (12*K B/px).  One thread per pixel; the K class values of a pixel sit HW*4 bytes apart
// (NCHW), so every load/store instruction of a warp is one fully coalesced 128-byte line, and all K
// loads of a thread are issued before the first use (K independent requests in flight per thread).
#include "common.cuh"

namespace b200seg {

constexpr int K3_THREADS = 256;

template <int KT>
__global__ void __launch_bounds__(K3_THREADS) k3_softce_fwd(const float* __restrict__ pred, const float* __restrict__ soft,
                                                            const float* __restrict__ wts, int K, long long HW,
                                                            long long total_px, float* __restrict__ partial) {
  __shared__ float red[K3_THREADS / 32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)K3_THREADS + threadIdx.x; i < total_px; i += (long long)gridDim.x * K3_THREADS) {
    const long long n = i / HW;
    const long long base = n * K * HW + (i - n * HW);
    float p[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) p[k] = __ldcs(pred + base + k * HW);
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) m = fmaxf(m, p[k]);
    float s = 0.f, sq = 0.f, sqp = 0.f;
    const float mneg = -m * 1.4426950408889634f;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) {
        const float q = __ldcs(soft + base + k * HW);
        s += fast_exp2(fmaf(p[k], 1.4426950408889634f, mneg));
        sq += q;
        sqp = fmaf(q, p[k], sqp);
      }
    float l = (m + __logf(s)) * sq - sqp;
    if (wts) l *= __ldcs(wts + i);
    acc += l;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < K3_THREADS / 32; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) k3_finalize(const float* partial, int n, double inv_count, float* out) {
  __shared__ double sl[8];
  double l = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) l += (double)partial[i];
  l = warp_sum_d(l);
  if ((threadIdx.x & 31) == 0) sl[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    double L = 0.0;
    for (int i = 0; i < 8; ++i) L += sl[i];
    out[0] = (float)(L * inv_count);
  }
}

template <int KT>
__global__ void __launch_bounds__(K3_THREADS) k3_softce_bwd(const float* __restrict__ pred, const float* __restrict__ soft,
                                                            const float* __restrict__ wts, const float* __restrict__ grad_out,
                                                            int K, long long HW, long long total_px, float inv_count,
                                                            float* __restrict__ grad) {
  const float go = (grad_out ? grad_out[0] : 1.f) * inv_count;
  for (long long i = blockIdx.x * (long long)K3_THREADS + threadIdx.x; i < total_px; i += (long long)gridDim.x * K3_THREADS) {
    const long long n = i / HW;
    const long long base = n * K * HW + (i - n * HW);
    float p[KT], q[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) p[k] = __ldcs(pred + base + k * HW);
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) q[k] = __ldcs(soft + base + k * HW);
    float m = -INFINITY, sq = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) { m = fmaxf(m, p[k]); sq += q[k]; }
    float s = 0.f;
    const float mneg = -m * 1.4426950408889634f;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) {
        p[k] = fast_exp2(fmaf(p[k], 1.4426950408889634f, mneg));
        s += p[k];
      }
    const float wgt = (wts ? __ldcs(wts + i) : 1.f) * go;
    const float a = wgt * sq * __fdividef(1.f, s);
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) __stcs(grad + base + k * HW, fmaf(a, p[k], -wgt * q[k]));
  }
}

long long k3_workspace_bytes() { return 4096 * 4 + 64; }

static int k3_grid(long long total_px) {
  long long blocks = ceil_div_ll(total_px, K3_THREADS);
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks > 4096) blocks = 4096;
  return (int)blocks;
}

int k3_forward(const float* pred, const float* soft, const float* wts, int N, int K, int H, int W, void* workspace,
               long long workspace_bytes, float* loss_out, cudaStream_t stream) {
  B200SEG_CHECK_ARG(pred && soft && workspace && loss_out, "soft_ce_forward: null pointer");
  B200SEG_CHECK_ARG(N > 0 && K > 0 && H > 0 && W > 0, "soft_ce_forward: bad shape");
  B200SEG_CHECK_ARG(K <= 64, "soft_ce_forward: %d channels > 64 is not supported", K);
  B200SEG_CHECK_ARG(workspace_bytes >= k3_workspace_bytes(), "soft_ce_forward: workspace too small");
  const long long HW = (long long)H * W, total = HW * N;
  const int grid = k3_grid(total);
  float* partial = reinterpret_cast<float*>(workspace);
  profile_begin(8, stream);
  if (K <= 4) k3_softce_fwd<4><<<grid, K3_THREADS, 0, stream>>>(pred, soft, wts, K, HW, total, partial);
  else if (K <= 38) k3_softce_fwd<38><<<grid, K3_THREADS, 0, stream>>>(pred, soft, wts, K, HW, total, partial);
  else k3_softce_fwd<64><<<grid, K3_THREADS, 0, stream>>>(pred, soft, wts, K, HW, total, partial);
  profile_end(8, stream);
  B200SEG_LAUNCH_CHECK();
  k3_finalize<<<1, 256, 0, stream>>>(partial, grid, 1.0 / (double)total, loss_out);
  B200SEG_LAUNCH_CHECK();
  return B200SEG_OK;
}

int k3_backward(const float* pred, const float* soft, const float* wts, const float* grad_out, int N, int K, int H, int W,
                float* grad_pred, cudaStream_t stream) {
  B200SEG_CHECK_ARG(pred && soft && grad_pred, "soft_ce_backward: null pointer");
  B200SEG_CHECK_ARG(N > 0 && K > 0 && H > 0 && W > 0, "soft_ce_backward: bad shape");
  B200SEG_CHECK_ARG(K <= 64, "soft_ce_backward: %d channels > 64 is not supported", K);
  const long long HW = (long long)H * W, total = HW * N;
  const int grid = k3_grid(total);
  const float inv = (float)(1.0 / (double)total);
  profile_begin(9, stream);
  if (K <= 4) k3_softce_bwd<4><<<grid, K3_THREADS, 0, stream>>>(pred, soft, wts, grad_out, K, HW, total, inv, grad_pred);
  else if (K <= 38) k3_softce_bwd<38><<<grid, K3_THREADS, 0, stream>>>(pred, soft, wts, grad_out, K, HW, total, inv, grad_pred);
  else k3_softce_bwd<64><<<grid, K3_THREADS, 0, stream>>>(pred, soft, wts, grad_out, K, HW, total, inv, grad_pred);
  profile_end(9, stream);
  B200SEG_LAUNCH_CHECK();
  return B200SEG_OK;
}

}  // namespace b200seg
