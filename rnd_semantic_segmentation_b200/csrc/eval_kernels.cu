// K4: align-corners bilinear upsample -> argmax of the softmax probabilities -> C x C confusion
// matrix, in ONE streaming pass that never materialises the full-resolution logits.
//
// Replaces, for one frame (reference file:line):
//   F.interpolate(..., bilinear, align_corners=True)      core/utils/utility.py:185
//   F.softmax(dim=1)                                      core/utils/utility.py:186
//   output.max(1)[1]                                      core/testers/aspp_tester.py:63
//   confusion_matrix(cfg, pd, gt)   (per-pixel Python loop) core/utils/utility.py:347-359
//   intersectionAndUnionGPU(...)    (3x CPU histc)         core/utils/utility.py:148-161  [= diag/row/col sums]
//
// Bit-exactness: the upsampled value uses ATen's expression form
//   h0*(w0*a + w1*b) + h1*(w0*c + w1*d)   (ATen/native/cuda/UpSampleBilinear2d.cu)
// and argmax(softmax(v)) == first-max(v) unless a lower-indexed class is within ~1 ulp of the
// maximum after exp/divide; pixels whose top-2 gap is <= 1e-6 re-run ATen's exact spatial-softmax
// sequence (max, sum of expf(v-max) in class order, expf(v-max)/sum, first maximum).
//
// Layout / mapping: one thread per output column, warps walk rows (row taps are warp-uniform);
// the horizontal lerps t_c (upper source row) and u_c (lower source row) of all classes live in
// registers and are reused for every output row of the same source-row pair.  Labels (int64, the
// only large stream) are read with streaming 64-bit loads, eight rows in flight per thread.
// Histogram: each lane owns private uint8 counters hist[bin][lane] in shared memory (no atomics,
// no contention even for i.i.d. labels), folded into an int32 CTA histogram once per tile and into
// the int64 global matrix once per frame per CTA.
#include "common.cuh"

namespace b200seg {

constexpr int K4_THREADS = 128;
constexpr int K4_TILE_W = 128;
constexpr int K4_TILE_H = 32;
constexpr int K4_STRIP = 8;
constexpr float K4_NEAR_TIE = 1e-6f;

struct K4Params {
  const float* logits;      // [N, C, h, w]
  const long long* labels;  // [N, H, W] or null
  long long* cm;            // [C, C] (+ frame * cm_frame_stride) or null
  long long* pred;          // [N, H, W] or null
  int N, C, h, w, H, W;
  int ignore_index;
  long long cm_frame_stride;
  float scale_h, scale_w;
  int tiles_x, tiles_y, total_tiles, tiles_per_cta;
  int use_u8;               // per-lane uint8 counters (C*C*32*4 bytes of smem) or aggregated atomics
};

template <int FMA>
__device__ __forceinline__ float lerp2(float wa, float a, float wb, float b) {
  if (FMA == 0) return wa * a + wb * b;                       // compiler's contraction (same source form as ATen)
  if (FMA == 1) return fmaf(wa, a, wb * b);
  if (FMA == 2) return fmaf(wb, b, wa * a);
  return __fadd_rn(__fmul_rn(wa, a), __fmul_rn(wb, b));       // no contraction
}

// Exact ATen spatial-softmax sequence for one pixel (rare path): sequential float sum of expf(v - max) in class
// order, IEEE divide, first maximum.  TOP/BOT select which register array holds the upper source row.
template <int CT, int FMA, bool SWAP, bool EXACT>
__device__ __forceinline__ int exact_softmax_argmax(const float (&a0)[CT], const float (&a1)[CT], int C, float h0, float h1,
                                                    float m) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (EXACT || c < C) s += expf((SWAP ? lerp2<FMA>(h0, a1[c], h1, a0[c]) : lerp2<FMA>(h0, a0[c], h1, a1[c])) - m);
  float pbest = -1.f;
  int idx = 0;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (EXACT || c < C) {
      const float pc = expf((SWAP ? lerp2<FMA>(h0, a1[c], h1, a0[c]) : lerp2<FMA>(h0, a0[c], h1, a1[c])) - m) / s;
      if (pc > pbest) { pbest = pc; idx = c; }
    }
  return idx;
}

// argmax over classes of the interpolated logits of one pixel (first index on ties), with the near-tie rescue.
template <int CT, int FMA, bool SWAP, bool EXACT>
__device__ __forceinline__ int k4_pixel_argmax(const float (&a0)[CT], const float (&a1)[CT], int C, float h0, float h1) {
  float best = -INFINITY, second = -INFINITY;
  int idx = 0;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (EXACT || c < C) {
      const float v = SWAP ? lerp2<FMA>(h0, a1[c], h1, a0[c]) : lerp2<FMA>(h0, a0[c], h1, a1[c]);
      const bool gt = v > best;
      second = fmaxf(second, fminf(best, v));
      idx = gt ? c : idx;
      best = fmaxf(best, v);
    }
  if (best - second <= K4_NEAR_TIE) idx = exact_softmax_argmax<CT, FMA, SWAP, EXACT>(a0, a1, C, h0, h1, best);
  return idx;
}

// horizontal lerp of one source row for all classes into a register array
template <int CT, int FMA, bool EXACT>
__device__ __forceinline__ void k4_load_row(float (&dst)[CT], const float* __restrict__ lg, int C, long long hw, int row, int w,
                                            const Tap& tapx) {
  const float* p0 = lg + (long long)row * w + tapx.i0;
  const float* p1 = lg + (long long)row * w + tapx.i1;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (EXACT || c < C) {
      dst[c] = lerp2<FMA>(tapx.l0, __ldg(p0), tapx.l1, __ldg(p1));
      p0 += hw; p1 += hw;
    }
}

template <int CT, int FMA, bool EXACT>
__global__ void __launch_bounds__(K4_THREADS, 4) k4_upsample_argmax_confusion(const K4Params p) {
  extern __shared__ __align__(16) uint8_t k4_smem[];
  const int C = EXACT ? CT : p.C, CC = C * C;
  int* cta_hist = reinterpret_cast<int*>(k4_smem);                       // [CC]
  const int hist_off = ((CC * 4 + 15) / 16) * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_hist_bytes = ((CC * 32 + 15) / 16) * 16;
  uint8_t* whist = k4_smem + hist_off + warp * warp_hist_bytes;           // [CC][32] uint8, this warp's
  const bool do_cm = (p.cm != nullptr) && (p.labels != nullptr);

  if (do_cm) {
    for (int i = threadIdx.x; i < CC; i += K4_THREADS) cta_hist[i] = 0;
    if (p.use_u8) {
      int4* z = reinterpret_cast<int4*>(whist);
      for (int i = lane; i < warp_hist_bytes / 16; i += 32) z[i] = make_int4(0, 0, 0, 0);
    }
  }
  __syncthreads();

  const int tile_begin = blockIdx.x * p.tiles_per_cta;
  const int tile_end = min(p.total_tiles, tile_begin + p.tiles_per_cta);
  const int tiles_per_frame = p.tiles_x * p.tiles_y;
  const long long hw = (long long)p.h * p.w;
  int cur_frame = -1;

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int n = tile / tiles_per_frame;
    const int trem = tile - n * tiles_per_frame;
    const int ty = trem / p.tiles_x;
    const int tx = trem - ty * p.tiles_x;

    if (do_cm && n != cur_frame) {           // frame boundary: publish this CTA's counts for the finished frame
      if (cur_frame >= 0) {
        __syncthreads();
        long long* dst = p.cm + (long long)cur_frame * p.cm_frame_stride;
        for (int i = threadIdx.x; i < CC; i += K4_THREADS) {
          const int v = cta_hist[i];
          if (v) atomicAdd(reinterpret_cast<unsigned long long*>(dst + i), (unsigned long long)v);
          cta_hist[i] = 0;
        }
        __syncthreads();
      }
      cur_frame = n;
    }

    const int x = tx * K4_TILE_W + threadIdx.x;
    const bool xvalid = x < p.W;
    const Tap tapx = ac_tap(p.scale_w, xvalid ? x : p.W - 1, p.w);
    const float* lg = p.logits + (long long)n * C * hw;
    const long long* lab_ptr = p.labels ? p.labels + (long long)n * p.H * p.W + x : nullptr;
    long long* pred_ptr = p.pred ? p.pred + (long long)n * p.H * p.W + x : nullptr;
    const bool load_labels = do_cm && xvalid;

    // a0 / a1 hold the horizontally interpolated logits of two source rows; `swap` says which one is the upper row.
    float a0[CT], a1[CT];
    int row0 = -1, row1 = -1;
    bool swap = false;
    const int y_begin = ty * K4_TILE_H;
    const int y_end = min(p.H, y_begin + K4_TILE_H);

    // label prefetch pipeline, 4 rows deep
    long long q0 = -1, q1 = -1, q2 = -1, q3 = -1;
    if (load_labels) {
      if (y_begin + 0 < y_end) q0 = ld_stream_s64(lab_ptr + (long long)(y_begin + 0) * p.W);
      if (y_begin + 1 < y_end) q1 = ld_stream_s64(lab_ptr + (long long)(y_begin + 1) * p.W);
      if (y_begin + 2 < y_end) q2 = ld_stream_s64(lab_ptr + (long long)(y_begin + 2) * p.W);
      if (y_begin + 3 < y_end) q3 = ld_stream_s64(lab_ptr + (long long)(y_begin + 3) * p.W);
    }

#pragma unroll 1
    for (int y = y_begin; y < y_end; ++y) {
      const long long g = q0;
      q0 = q1; q1 = q2; q2 = q3;
      q3 = (load_labels && y + 4 < y_end) ? ld_stream_s64(lab_ptr + (long long)(y + 4) * p.W) : -1;

      const Tap tapy = ac_tap(p.scale_h, y, p.h);                 // warp-uniform
      const int top = swap ? row1 : row0, bot = swap ? row0 : row1;
      if (tapy.i0 != top || tapy.i1 != bot) {                      // new source-row pair (uniform branch)
        if (row0 == tapy.i0) {
          swap = false;
          if (row1 != tapy.i1) { k4_load_row<CT, FMA, EXACT>(a1, lg, C, hw, tapy.i1, p.w, tapx); row1 = tapy.i1; }
        } else if (row1 == tapy.i0) {
          swap = true;
          if (row0 != tapy.i1) { k4_load_row<CT, FMA, EXACT>(a0, lg, C, hw, tapy.i1, p.w, tapx); row0 = tapy.i1; }
        } else {
          swap = false;
          k4_load_row<CT, FMA, EXACT>(a0, lg, C, hw, tapy.i0, p.w, tapx); row0 = tapy.i0;
          if (row1 != tapy.i1) { k4_load_row<CT, FMA, EXACT>(a1, lg, C, hw, tapy.i1, p.w, tapx); row1 = tapy.i1; }
        }
      }
      const int idx = swap ? k4_pixel_argmax<CT, FMA, true, EXACT>(a0, a1, C, tapy.l0, tapy.l1)
                           : k4_pixel_argmax<CT, FMA, false, EXACT>(a0, a1, C, tapy.l0, tapy.l1);
      if (xvalid) {
        if (pred_ptr) pred_ptr[(long long)y * p.W] = (long long)idx;
        if (do_cm) {
          const bool count = (g >= 0) && (g < C) && (g != p.ignore_index);
          if (p.use_u8) {
            if (count) {
              uint8_t* hp = whist + ((int)g * C + idx) * 32 + lane;
              *hp = (uint8_t)(*hp + 1);
            }
          } else {
            const int bin = count ? (int)g * C + idx : -1;
            const unsigned peers = __match_any_sync(__activemask(), bin);
            if (count && lane == (__ffs(peers) - 1)) atomicAdd(&cta_hist[bin], __popc(peers));
          }
        }
      }
    }

    if (do_cm && p.use_u8) {                  // fold this warp's uint8 counters (<= 32 per lane per tile)
      __syncwarp();
      const uint4* wh = reinterpret_cast<const uint4*>(whist);
      for (int b = lane; b < CC; b += 32) {
        const uint4 a = wh[b * 2], c4 = wh[b * 2 + 1];
        unsigned sum = 0;
        sum = __dp4a(a.x, 0x01010101u, sum); sum = __dp4a(a.y, 0x01010101u, sum);
        sum = __dp4a(a.z, 0x01010101u, sum); sum = __dp4a(a.w, 0x01010101u, sum);
        sum = __dp4a(c4.x, 0x01010101u, sum); sum = __dp4a(c4.y, 0x01010101u, sum);
        sum = __dp4a(c4.z, 0x01010101u, sum); sum = __dp4a(c4.w, 0x01010101u, sum);
        if (sum) atomicAdd(&cta_hist[b], (int)sum);
      }
      __syncwarp();
      int4* z = reinterpret_cast<int4*>(whist);
      for (int i = lane; i < warp_hist_bytes / 16; i += 32) z[i] = make_int4(0, 0, 0, 0);
      __syncwarp();
    }
  }

  if (do_cm && cur_frame >= 0) {
    __syncthreads();
    long long* dst = p.cm + (long long)cur_frame * p.cm_frame_stride;
    for (int i = threadIdx.x; i < CC; i += K4_THREADS) {
      const int v = cta_hist[i];
      if (v) atomicAdd(reinterpret_cast<unsigned long long*>(dst + i), (unsigned long long)v);
    }
  }
}

template <int CT, int FMA, bool EXACT>
static int k4_launch_t(const K4Params& p, int grid, size_t smem, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    B200SEG_CUDA(cudaFuncSetAttribute(k4_upsample_argmax_confusion<CT, FMA, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      100 * 1024));
    configured = true;
  }
  profile_begin(7, stream);
  k4_upsample_argmax_confusion<CT, FMA, EXACT><<<grid, K4_THREADS, smem, stream>>>(p);
  profile_end(7, stream);
  B200SEG_LAUNCH_CHECK();
  return B200SEG_OK;
}

template <int CT, bool EXACT>
static int k4_launch_c(const K4Params& p, int fma_mode, int grid, size_t smem, cudaStream_t stream) {
  switch (fma_mode) {
    case 0: return k4_launch_t<CT, 0, EXACT>(p, grid, smem, stream);
    case 1: return k4_launch_t<CT, 1, EXACT>(p, grid, smem, stream);
    case 2: return k4_launch_t<CT, 2, EXACT>(p, grid, smem, stream);
    default: return k4_launch_t<CT, 3, EXACT>(p, grid, smem, stream);
  }
}

int k4_launch(const float* logits, int N, int C, int h, int w, const long long* labels, int H, int W, int ignore_index,
              long long* cm, long long cm_frame_stride, long long* pred, int fma_mode, cudaStream_t stream) {
  B200SEG_CHECK_ARG(logits != nullptr, "upsample_argmax_confusion: logits is null");
  B200SEG_CHECK_ARG(N > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "upsample_argmax_confusion: bad shape N=%d C=%d h=%d w=%d H=%d W=%d", N, C, h, w, H, W);
  B200SEG_CHECK_ARG(C <= 32, "upsample_argmax_confusion: num_classes=%d > 32 is not supported", C);
  B200SEG_CHECK_ARG(cm == nullptr || labels != nullptr, "upsample_argmax_confusion: confusion matrix requested without labels");
  B200SEG_CHECK_ARG(cm != nullptr || pred != nullptr, "upsample_argmax_confusion: nothing to compute (cm and pred both null)");
  K4Params p;
  p.logits = logits; p.labels = labels; p.cm = cm; p.pred = pred;
  p.N = N; p.C = C; p.h = h; p.w = w; p.H = H; p.W = W;
  p.ignore_index = ignore_index;
  p.cm_frame_stride = cm_frame_stride;
  p.scale_h = ac_scale(h, H);
  p.scale_w = ac_scale(w, W);
  p.tiles_x = ceil_div(W, K4_TILE_W);
  p.tiles_y = ceil_div(H, K4_TILE_H);
  p.total_tiles = N * p.tiles_x * p.tiles_y;
  const int CC = C * C;
  p.use_u8 = (CC * 32 * (K4_THREADS / 32) <= 64 * 1024) ? 1 : 0;
  const size_t smem = ((CC * 4 + 15) / 16) * 16 + (p.use_u8 ? (size_t)(K4_THREADS / 32) * (((CC * 32 + 15) / 16) * 16) : 0);
  int max_ctas = num_sms() * (smem > 48 * 1024 ? 3 : 4);
  int grid = p.total_tiles < max_ctas ? p.total_tiles : max_ctas;
  p.tiles_per_cta = ceil_div(p.total_tiles, grid);
  grid = ceil_div(p.total_tiles, p.tiles_per_cta);
  if (C == 2) return k4_launch_c<2, true>(p, fma_mode, grid, smem, stream);
  if (C == 19) return k4_launch_c<19, true>(p, fma_mode, grid, smem, stream);
  if (C <= 8) return k4_launch_c<8, false>(p, fma_mode, grid, smem, stream);
  return k4_launch_c<32, false>(p, fma_mode, grid, smem, stream);
}

// ---------------------------------------------------------------------------------------------
// API-compat: confusion matrix / IoU areas from an already-materialised prediction map
// (confusion_matrix(cfg, pd, gt) utility.py:347-359 and intersectionAndUnionGPU utility.py:148-161,
//  which also overwrites output[target == ignore] = ignore in place).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) confusion_pairs_kernel(long long* pd, const long long* gt, long long n, int C,
                                                              int ignore_index, int mutate_pd, long long* cm) {
  extern __shared__ int cp_hist[];
  const int CC = C * C;
  for (int i = threadIdx.x; i < CC; i += blockDim.x) cp_hist[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_round = ceil_div_ll(n, 32) * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_round; i += stride) {
    int bin = -1;
    if (i < n) {
      const long long g = gt[i];
      const long long q = pd[i];
      if (g == ignore_index) {
        if (mutate_pd) pd[i] = ignore_index;
      } else if (g >= 0 && g < C && q >= 0 && q < C) {
        bin = (int)g * C + (int)q;
      }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(&cp_hist[bin], __popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CC; i += blockDim.x) {
    const int v = cp_hist[i];
    if (v) atomicAdd(reinterpret_cast<unsigned long long*>(cm + i), (unsigned long long)v);
  }
}

int confusion_pairs_launch(long long* pd, const long long* gt, long long n, int C, int ignore_index, int mutate_pd,
                           long long* cm, cudaStream_t stream) {
  B200SEG_CHECK_ARG(pd && gt && cm, "confusion_from_pred: null pointer");
  B200SEG_CHECK_ARG(C > 0 && C <= 104, "confusion_from_pred: num_classes=%d unsupported (1..104)", C);
  if (n <= 0) return B200SEG_OK;
  long long blocks = ceil_div_ll(n, 256 * 8);
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  confusion_pairs_kernel<<<(unsigned)blocks, 256, C * C * sizeof(int), stream>>>(pd, gt, n, C, ignore_index, mutate_pd, cm);
  B200SEG_LAUNCH_CHECK();
  return B200SEG_OK;
}

}  // namespace b200seg
