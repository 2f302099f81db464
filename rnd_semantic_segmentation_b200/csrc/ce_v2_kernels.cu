// K2, warp-tile kernel: align-corners bilinear upsample fused with softmax cross-entropy (ignore_index) forward AND the
// gradient with respect to the LOW-RESOLUTION logits -- the same contract, workspace layout and finalize kernels as the
// CTA-tile kernel in ce_kernels.cu (reference lines: classifier.py:30-31, aspp_fada.py:93-94, aspp_trainer.py:61,91-92), for
// the shapes the reference actually runs (upsampling by >= ~5.3x horizontally, not downsampling vertically; C = 19 or 2).
//
// What is different, and why (ncu of the CTA-tile kernel, profiles/r1i: 286 warp instructions per 32-pixel row at 35 % issue
// utilisation; 17 % of the stall samples at CTA barriers, 15 % on un-prefetched source-row loads, the rest fixed-latency waits):
//   * every WARP owns a 32-column x 32-row tile and never synchronises with another warp: the per-column gradient sums are
//     reduced onto their source columns inside the warp (shared memory + __syncwarp) and the finished source row is written
//     straight to the per-tile partial block in global memory -- no __syncthreads in the kernel, no shared gradient block;
//   * the gradient toward a source row is carried in registers across the two row segments that touch it (as lower row, then as
//     upper row), so each source row is reduced and written once per tile instead of twice;
//   * source rows are fetched one segment ahead (registers) and labels one 8-row strip ahead (cp.async ring for int64 labels,
//     plain byte loads for uint8 labels), also across tiles;
//   * an output row is one fma per class pair (upper row + l1y * (lower - upper)), and the labelled-class logit never enters the
//     row loop: sum over pixels of logit[label] = sum over source rows and classes of row[c] * (one-hot weight the pixels put on
//     it), and those weights are exactly what the gradient's one-hot part accumulates anyway -- one dot product per row flush;
//   * soft-max without a per-pixel maximum: the two horizontally interpolated source rows are kept in the log2 domain, pre-scaled
//     by log2(e) / T and shifted by M = max over both rows and all classes (one max tree per SEGMENT, not per pixel).  Every
//     interpolated logit is a convex combination of the two rows, hence <= M: exp2 cannot overflow, and no per-pixel max tree /
//     subtraction is needed.  A pixel whose sum of exponentials underflows (logit range across two adjacent source pixels above
//     ~60 in natural units) takes a rare exact path with its own maximum;
//   * tiles are handed out by an atomic counter (persistent warps): no tail wave, and results do not depend on the schedule
//     (every tile owns its slots of the workspace), so they stay run-to-run bit-identical.
#include "ce_geom.cuh"

namespace b200seg {

constexpr int V2_THREADS = 128;
constexpr int V2_WARPS = V2_THREADS / 32;
constexpr int V2_TW = 32;
constexpr int V2_TH = 32;
constexpr int V2_STRIP = 8;
constexpr int V2_SPAN = K2V2_SPAN;
constexpr float V2_PAD = -1e30f;
constexpr int V2_OH_PITCH = 34 * 8;             // bytes per class-pair row of a one-hot / staging slot: 34 float2 (16-byte aligned rows)

__host__ __device__ constexpr int v2_np(int CT) { return (CT + 1) / 2; }
__host__ __device__ constexpr int v2_ns(int CT) { return (CT * V2_SPAN + 31) / 32; }

// per-warp shared memory (bytes)
struct V2Smem {
  int raw, oh, wt, runs, tab, ring, total;
};
__host__ __device__ constexpr V2Smem v2_smem_layout(int CT, int lm) {
  V2Smem L{};
  int o = 0;
  L.raw = o; o += 2 * CT * V2_SPAN * 4;                    // two staging windows [CT][SPAN] for source rows in flight (cp.async)
  o = (o + 15) / 16 * 16;
  L.oh = o; o += 3 * v2_np(CT) * V2_OH_PITCH;              // three row slots [pair][lane] float2: -sum of one-hot weights
  o = (o + 15) / 16 * 16;
  L.wt = o; o += V2_SPAN * 32 * 4;                         // wt[s][lane]: weight of the lane's column toward source column s
  L.runs = o; o += V2_SPAN * 8;                            // {first lane, lane count} contributing to source column s
  o = (o + 15) / 16 * 16;
  L.tab = o; o += (V2_TH + 1) * 16;                        // per tile row {l0y, l1y, bits(i0y), bits(segment end)} (+1: look-ahead read)
  L.ring = o; o += lm == 2 ? (V2_STRIP + 1) * 32 : (2 * V2_STRIP + 1) * 32 * (lm == 0 ? 8 : 1);      // (+1 row: look-ahead read)
  o = (o + 15) / 16 * 16;
  L.total = o;
  return L;
}

__device__ __forceinline__ void sts_v2f32(unsigned a, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory"); }
__device__ __forceinline__ float2 lds_v2f32(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
// s is a sum of exponentials in [1e-30, 32]: the bare approximate instructions, without the range fix-ups of __log2f / __fdividef
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts_v2u32(unsigned a, unsigned x, unsigned y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }

// LM: how the labels arrive.  0: int64, per-lane 8-byte cp.async ring.  1: uint8 rows that are 4-byte aligned (W % 4 == 0, aligned
// base): every fourth lane copies 4 label bytes with one 4-byte cp.async -- the same two-stage ring at an eighth of the bytes.
// 2: uint8 at any alignment: plain byte loads into registers a strip ahead, parked in shared memory at the start of their strip.
template <int CT, bool GRAD, int LM>
__global__ void __launch_bounds__(V2_THREADS, 3) k2v2_upsample_ce_main(const __grid_constant__ K2Params p) {
  constexpr bool LU8 = LM != 0, LREG = LM == 2;
  constexpr int NP = v2_np(CT);
  constexpr int NS = v2_ns(CT);
  constexpr V2Smem L = v2_smem_layout(CT, LM);
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  constexpr int LAB_PITCH = LU8 ? 32 : 256;
  extern __shared__ __align__(16) uint8_t v2_smem[];
  const K2Geom& g = p.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned base_s = smem_u32(v2_smem) + warp * L.total;
  const unsigned raw_s = base_s + L.raw;
  const unsigned oh_s = base_s + L.oh + lane * 8;                    // this lane's float2 in pair row 0 of slot 0
  const unsigned wt_s = base_s + L.wt;
  const unsigned runs_s = base_s + L.runs;
  const unsigned tab_s = base_s + L.tab;
  const unsigned ring_s = LU8 ? base_s + L.ring + lane : base_s + L.ring + lane * 8;
  constexpr int SLOT_BYTES = NP * V2_OH_PITCH;

  const long long hw = (long long)g.h * g.w;
  const long long HW = (long long)g.H * g.W;
  const int tiles_per_frame = g.tiles_x * g.tiles_y;
  const int total_tiles = g.N * tiles_per_frame;
  const int blk_floats = g.ispan_max * g.jspan_max * g.C;
  const float K = p.inv_T * LOG2E;
  const unsigned ign32 = (p.ignore_index >= 0) ? (unsigned)p.ignore_index : 0xffffffffu;
  const long long lab_row_bytes = (long long)g.W * (LU8 ? 1 : 8);

  // zero the three one-hot slots once (every flush leaves its slot zeroed again)
#pragma unroll
  for (int q = 0; q < 3 * NP; ++q) sts_v2f32(oh_s + q * V2_OH_PITCH, make_float2(0.f, 0.f));

  // ---- dynamic tile hand-out: one atomic per warp and tile, issued a tile ahead by lane 0 and broadcast only where the id is
  //      first needed (two label strips later), so the atomic's latency is never waited for ------------------------------------
  auto grab = [&]() -> int { return (lane == 0) ? (int)atomicAdd(p.counter, 1u) : 0; };
  auto bcast = [&](int t) -> int { return __shfl_sync(0xffffffffu, t, 0); };

  // ---- label prefetch cursor: strip pf_st of tile pf_tile goes in flight next ----------------------------------------------
  int pf_tile, pf_st = 0, pf_nst = 1, pf_stage = 0, pf_rows = 0;
  bool pf_copy = false;                             // LM == 1: this lane copies 4 label bytes (every fourth lane, inside the image)
  const char* pf_src = nullptr;
  unsigned nb[LREG ? V2_STRIP : 1];
  auto pf_setup = [&]() {
    pf_rows = 0;
    pf_st = 0;
    pf_nst = 1;
    if (pf_tile < total_tiles) {
      const int n = pf_tile / tiles_per_frame;
      const int trem = pf_tile - n * tiles_per_frame;
      const int ty = trem / g.tiles_x;
      const int tx = trem - ty * g.tiles_x;
      const int x = tx * V2_TW + lane;
      const int y0 = ty * V2_TH;
      pf_src = reinterpret_cast<const char*>(p.labels) + (n * HW + (long long)y0 * g.W + min(x, g.W - 1)) * (LU8 ? 1 : 8);
      pf_rows = min(V2_TH, g.H - y0);
      pf_copy = ((lane & 3) == 0) && (x < g.W);
      pf_nst = (pf_rows + V2_STRIP - 1) / V2_STRIP;            // strips this tile really has (the row loop runs as many)
    }
  };
  auto issue_strip = [&]() {
    const unsigned dst = ring_s + pf_stage * (V2_STRIP * LAB_PITCH);
    const char* src = pf_src;
#pragma unroll
    for (int r = 0; r < V2_STRIP; ++r) {
      if constexpr (LREG) {
        nb[r] = (r < pf_rows) ? (unsigned)__ldg(reinterpret_cast<const unsigned char*>(src)) : 0xffu;
      } else if constexpr (LU8) {
        if (r < pf_rows && pf_copy) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + r * 32), "l"(src) : "memory");
      } else {
        if (r < pf_rows) cp_async_8s(dst + r * 256, src);
      }
      src += lab_row_bytes;
    }
    if constexpr (!LREG) cp_async_commit();
    pf_src = src;
    pf_rows -= V2_STRIP;
    pf_stage ^= 1;
    ++pf_st;
  };

  int tile = bcast(grab());
  int tile_next_raw = grab();                       // lane 0 holds the id; everybody gets it at the first strip that needs it
  pf_tile = tile;
  pf_setup();
  int cur_stage = 0;
  if (tile < total_tiles) issue_strip();

#pragma unroll 1
  while (tile < total_tiles) {
    int tile_next = -1;                             // not broadcast yet
    const int n = tile / tiles_per_frame;
    const int trem = tile - n * tiles_per_frame;
    const int ty = trem / g.tiles_x;
    const int tx = trem - ty * g.tiles_x;
    const int x = tx * V2_TW + lane;
    const bool xvalid = x < g.W;
    const Tap tapx = ac_tap(g.scale_w, xvalid ? x : g.W - 1, g.w);
    const int j_lo = __shfl_sync(0xffffffffu, tapx.i0, 0);
    const int rel0 = tapx.i0 - j_lo, rel1 = tapx.i1 - j_lo;
    const int y0 = ty * V2_TH;
    const int y_end = min(g.H, y0 + V2_TH);
    const int i_lo = (int)(g.scale_h * (float)y0);
    const float* lg = p.logits + (long long)n * g.C * hw;
    const unsigned c_lim = xvalid ? (unsigned)g.C : 0u;
    float* blk = p.blocks + (long long)tile * blk_floats;

    __syncwarp();
    // column -> source-column weights and the lane runs feeding each source column
    if constexpr (GRAD) {
#pragma unroll
      for (int s = 0; s < V2_SPAN; ++s) {
        const float wv = xvalid ? ((rel0 == s ? tapx.l0 : 0.f) + (rel1 == s ? tapx.l1 : 0.f)) : 0.f;
        sts_f32(wt_s + (s * 32 + lane) * 4, wv);
        const unsigned m = __ballot_sync(0xffffffffu, xvalid && (rel0 == s || rel1 == s));
        if (lane == 0) sts_v2u32(runs_s + s * 8, m ? (unsigned)(__ffs(m) - 1) : 0u, (unsigned)__popc(m));
      }
    }
    // per-row table: vertical taps and the end of each run of rows sharing one source-row pair.  A row whose lower tap is
    // clamped onto its upper tap (last source row) carries its whole weight on the upper tap.
    {
      const int yy = min(y0 + lane, g.H - 1);
      const Tap t = ac_tap(g.scale_h, yy, g.h);
      const bool clamped = t.i1 == t.i0;
      const int nxt = __shfl_down_sync(0xffffffffu, t.i0, 1);
      const bool last = (lane == 31) || (y0 + lane >= y_end - 1) || (nxt != t.i0);
      const unsigned ends = __ballot_sync(0xffffffffu, last);
      const int seg_end = y0 + lane + __ffs(ends >> lane);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tab_s + lane * 16), "f"(clamped ? t.l0 + t.l1 : t.l0),
                   "f"(clamped ? 0.f : t.l1), "f"(__int_as_float(t.i0)), "f"(__int_as_float(seg_end))
                   : "memory");
    }
    // raw-window element offsets of this lane (element e = q*32 + lane: class e / SPAN, column j_lo + e % SPAN)
    // staging-window element e = q*32 + lane is class q*4 + lane/8, source column j_lo + lane%8: one offset, stride 4*hw per q
    const int soff0 = (lane >> 3) * (int)hw + min(j_lo + (lane & 7), g.w - 1);
    constexpr unsigned RAW_BYTES = CT * V2_SPAN * 4;
    const unsigned t0_o = (unsigned)rel0 * 4, t1_o = (unsigned)rel1 * 4;
    const float kx0 = K * tapx.l0, kx1 = K * tapx.l1;
    __syncwarp();

    // a_top: upper source row of the current segment (log2 domain, horizontally interpolated, shifted by M);
    // dlt  : lower row minus upper row, so an output row is ONE fma per class pair: v = a_top + l1y * dlt
    float2 a_top[NP], dlt[NP];
    float2 acc_top[GRAD ? NP : 1], acc_bot[GRAD ? NP : 1];
    if constexpr (GRAD) {
#pragma unroll
      for (int i = 0; i < NP; ++i) { acc_top[i] = make_float2(0.f, 0.f); acc_bot[i] = make_float2(0.f, 0.f); }
    }
    int pre_row = -1;
    int row_top = -1, row_bot = -1;
    float M = 0.f, m_bot = 0.f;
    float loss_acc = 0.f, cnt_acc = 0.f;

    // source row -> staging window (row & 1) by 4-byte cp.async: no registers are held while it is in flight
    auto row_fetch = [&](int row) {
      const float* rb = lg + (long long)row * g.w + soff0;
      const unsigned dst = raw_s + (row & 1) * RAW_BYTES + lane * 4;
      __syncwarp();                                 // every lane is done reading the previous contents of this window
#pragma unroll
      for (int q = 0; q < NS; ++q)
        if (q * 32 + lane < CT * V2_SPAN)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + q * 128), "l"(rb + (long long)q * 4 * hw) : "memory");
      cp_async_commit();
      pre_row = row;
    };
    // interpolate the staged row horizontally in the log2 domain (un-shifted); returns the row maximum
    auto row_commit = [&](float2 (&dst)[NP], int row) -> float {
      cp_async_wait<0>();
      __syncwarp();
      const unsigned t0_s = raw_s + (row & 1) * RAW_BYTES + t0_o, t1_s = raw_s + (row & 1) * RAW_BYTES + t1_o;
      const float2 L0 = make_float2(kx0, kx0), L1 = make_float2(kx1, kx1);
      float f[2 * NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
        a.x = lds_f32(t0_s + (2 * i) * V2_SPAN * 4); b.x = lds_f32(t1_s + (2 * i) * V2_SPAN * 4);
        if (2 * i + 1 < CT) { a.y = lds_f32(t0_s + (2 * i + 1) * V2_SPAN * 4); b.y = lds_f32(t1_s + (2 * i + 1) * V2_SPAN * 4); }
        dst[i] = fma2(L0, a, mul2(L1, b));
        if (2 * i + 1 >= CT) dst[i].y = V2_PAD;
        f[2 * i] = dst[i].x; f[2 * i + 1] = dst[i].y;
      }
      return tree_max3<0, 2 * NP, 2 * NP>(f);
    };

    // A source row is complete: (1) the labelled-class logits of all its pixels leave the loss as ONE dot product per lane,
    // sum_c row[c] * W[c] with W[c] = the one-hot weights the pixels of this column put on class c of this row (the slot holds
    // -W); (2) GRAD: gradient sums (softmax part in `acc`, one-hot part in the slot) are reduced onto the source columns and
    // written to block row `li`.  `row` = the row's values shifted by M.  Leaves the slot zeroed.
    auto flush_row = [&](const float2 (&row)[NP], float2 (&acc)[GRAD ? NP : 1], int slot, int li) {
      const unsigned so = oh_s + slot * SLOT_BYTES;
      float2 dot = make_float2(0.f, 0.f), sw = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 o = lds_v2f32(so + i * V2_OH_PITCH);
        float2 r = row[i];
        if (2 * i + 1 >= CT) r.y = 0.f;                           // padding class: weight 0, value -1e30
        dot = fma2(r, o, dot);
        sw = add2(sw, o);
        if constexpr (GRAD) sts_v2f32(so + i * V2_OH_PITCH, add2(o, acc[i]));
      }
      loss_acc += (dot.x + dot.y) + M * (sw.x + sw.y);            // -= sum_c (row[c] + M) * W[c]
      if constexpr (GRAD) {
        __syncwarp();
        const int s = lane >> 2, pg = lane & 3;
        unsigned lo, cnt;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(cnt) : "r"(runs_s + s * 8));
        float2 t[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        const unsigned sb = base_s + L.oh + slot * SLOT_BYTES + pg * V2_OH_PITCH;
        const unsigned wb = wt_s + s * 128;
        // two lanes per step (the weight table is zero outside a column's run, so starting at an even lane is harmless)
#pragma unroll 1
        for (unsigned k = lo & ~1u; k < lo + cnt; k += 2) {
          const float2 wk = lds_v2f32(wb + k * 4);
          const float2 W0 = make_float2(wk.x, wk.x), W1 = make_float2(wk.y, wk.y);
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            if (pg + 4 * u < NP) {
              const float4 v = lds_v4f32(sb + u * (4 * V2_OH_PITCH) + k * 8);
              t[u] = fma2(W0, make_float2(v.x, v.y), t[u]);
              t[u] = fma2(W1, make_float2(v.z, v.w), t[u]);
            }
          }
        }
        float* dst = blk + (li * V2_SPAN + s) * g.C;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int i = pg + 4 * u;
          if (i < NP) {
            dst[2 * i] = t[u].x;
            if (2 * i + 1 < CT) dst[2 * i + 1] = t[u].y;
          }
        }
        __syncwarp();
      }
#pragma unroll
      for (int i = 0; i < NP; ++i) sts_v2f32(so + i * V2_OH_PITCH, make_float2(0.f, 0.f));
    };

#pragma unroll 1
    for (int ys = y0; ys < y_end; ys += V2_STRIP) {
      // this strip's labels: park / wait, then put the next strip (possibly of the next tile) in flight
      unsigned lab_s;
      if constexpr (LREG) {
#pragma unroll
        for (int r = 0; r < V2_STRIP; ++r) sts_u8(ring_s + r * 32, nb[r]);
      }
      if (pf_st >= pf_nst) {                        // the cursor moves on to the next tile: its id is needed now
        if (tile_next < 0) tile_next = bcast(tile_next_raw);
        pf_tile = tile_next;
        pf_setup();
      }
      issue_strip();
      if constexpr (LREG) {
        lab_s = ring_s;
      } else {
        cp_async_wait<1>();                         // everything but the strip just issued has landed (in-order groups)
        if constexpr (LU8) __syncwarp();            // (a lane reads label bytes that another lane's copy brought in)
        lab_s = ring_s + cur_stage * (V2_STRIP * LAB_PITCH);
        cur_stage ^= 1;
      }
      const int ye = min(y_end, ys + V2_STRIP);
      int y = ys;
#pragma unroll 1
      while (y < ye) {
        float4 tr = lds_v4f32(tab_s + (y - y0) * 16);
        const int i0 = __float_as_int(tr.z);
        const int i1 = i0 + ((i0 < g.h - 1) ? 1 : 0);
        const int yend = min(__float_as_int(tr.w), ye);
        if (i0 != row_top) {                                       // new source-row pair (warp-uniform)
          float M_old = M;
          if (row_top < 0) {                                       // first segment of the tile: the upper row too
            row_fetch(i0);
            if (i1 != i0) row_fetch(i1);                           // both rows of the first pair in flight together
            m_bot = row_commit(a_top, i0);                         // (un-shifted for the moment: M_old = 0)
            M_old = 0.f;
            pre_row = i1;
          } else {                                                 // rows advance by one: the lower row becomes the upper row
            flush_row(a_top, acc_top, row_top % 3, row_top - i_lo);
#pragma unroll
            for (int i = 0; i < NP; ++i) a_top[i] = add2(a_top[i], dlt[i]);       // old lower row, still shifted by M_old
            if constexpr (GRAD) {
#pragma unroll
              for (int i = 0; i < NP; ++i) { acc_top[i] = acc_bot[i]; acc_bot[i] = make_float2(0.f, 0.f); }
            }
          }
          float m_new = m_bot;
          if (i1 != i0) {
            if (pre_row != i1) row_fetch(i1);
            m_new = row_commit(dlt, i1);                           // un-shifted new lower row
          }
          const float M_new = fmaxf(m_bot, m_new);
          const float2 D = make_float2(M_old - M_new, M_old - M_new), NM = make_float2(-M_new, -M_new);
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            a_top[i] = add2(a_top[i], D);                          // re-shifted to the new M
            dlt[i] = (i1 != i0) ? add2(add2(dlt[i], NM), make_float2(-a_top[i].x, -a_top[i].y)) : make_float2(0.f, 0.f);
          }
          M = M_new;
          m_bot = m_new;
          row_top = i0; row_bot = i1;
          if (i1 + 1 < g.h) row_fetch(i1 + 1);                     // the next segment's lower row goes in flight now
        }
        // this lane's one-hot slots of the two rows; a clamped pair (row_bot == row_top) sends its zero lower weight to the free slot
        const unsigned ohT_s = oh_s + (row_top % 3) * SLOT_BYTES, ohB_s = oh_s + ((row_top + 1) % 3) * SLOT_BYTES;
        unsigned la = lab_s + (y - ys) * LAB_PITCH;
        unsigned ts = tab_s + (y - y0) * 16;
        // software pipeline: the table entry and the label of row y + 1 are loaded while row y is computed
        unsigned gl, gh = 0u;
        if constexpr (LU8) gl = lds_u8(la);
        else { const uint2 g2 = lds_v2u32(la); gl = g2.x; gh = g2.y; }
#pragma unroll 1
        for (; y < yend; ++y) {
          ts += 16; la += LAB_PITCH;
          const float4 tr_n = lds_v4f32(ts);
          unsigned gl_n, gh_n = 0u;
          if constexpr (LU8) gl_n = lds_u8(la);
          else { const uint2 g2 = lds_v2u32(la); gl_n = g2.x; gh_n = g2.y; }
          const float w0 = tr.x, w1 = tr.y;
          const bool valid = (gh == 0u) && (gl < c_lim) && (gl != ign32);
          const unsigned gi = valid ? gl : 0u;
          const unsigned oo = (gi >> 1) * V2_OH_PITCH + (gi & 1) * 4;
          const float ot = lds_f32(ohT_s + oo), ob = lds_f32(ohB_s + oo);       // early: consumed after the exponentials
          const float2 W1 = make_float2(w1, w1);
          float e[2 * NP];
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const float2 v = fma2(W1, dlt[i], a_top[i]);
            e[2 * i] = fast_exp2(v.x);
            e[2 * i + 1] = (2 * i + 1 < CT) ? fast_exp2(v.y) : 0.f;
          }
          float s;
          {
            float2 t2[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) t2[i] = make_float2(e[2 * i], e[2 * i + 1]);
#pragma unroll
            for (int nn = NP; nn > 1; nn = (nn + 1) / 2) {
#pragma unroll
              for (int i = 0; i < nn / 2; ++i) t2[i] = add2(t2[i], t2[nn - 1 - i]);
            }
            s = t2[0].x + t2[0].y;
          }
          float madd = 0.f;                                       // exact path: the pixel's own maximum (log2 domain, relative to M)
          if (!(s > 1e-30f)) {                                    // sum of exponentials underflowed (or NaN logits): rare
            float f[2 * NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              const float2 v = fma2(W1, dlt[i], a_top[i]);
              f[2 * i] = v.x; f[2 * i + 1] = (2 * i + 1 < CT) ? v.y : V2_PAD;
            }
            madd = tree_max3<0, 2 * NP, 2 * NP>(f);
            s = 0.f;
#pragma unroll
            for (int c = 0; c < 2 * NP; ++c) { e[c] = (c < CT) ? fast_exp2(f[c] - madd) : 0.f; s += e[c]; }
          }
          if (valid) {
            loss_acc += fast_log2(s) + (madd + M);                  // the labelled-class logit is subtracted at the row flushes
            cnt_acc += 1.f;
          }
          const float d0 = valid ? w0 : 0.f, d1 = valid ? w1 : 0.f;
          if constexpr (GRAD) {
            const float inv_s = fast_rcp(s);
            const float c0 = d0 * inv_s, c1 = d1 * inv_s;
            const float2 C0 = make_float2(c0, c0), C1 = make_float2(c1, c1);
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              const float2 pr = make_float2(e[2 * i], e[2 * i + 1]);
              acc_top[i] = fma2(C0, pr, acc_top[i]);
              acc_bot[i] = fma2(C1, pr, acc_bot[i]);
            }
          }
          sts_f32(ohT_s + oo, ot - d0);
          sts_f32(ohB_s + oo, ob - d1);
          tr = tr_n; gl = gl_n; gh = gh_n;
        }
      }
    }
    // end of tile: the last upper and lower rows
    if (row_top >= 0) {
      flush_row(a_top, acc_top, row_top % 3, row_top - i_lo);
      if (row_bot != row_top) {
#pragma unroll
        for (int i = 0; i < NP; ++i) a_top[i] = add2(a_top[i], dlt[i]);
        flush_row(a_top, acc_bot, row_bot % 3, row_bot - i_lo);
      }
    }
    loss_acc = warp_sum(loss_acc);
    cnt_acc = warp_sum(cnt_acc);
    if (lane == 0) {
      p.loss_part[tile] = loss_acc * LN2;
      p.cnt_part[tile] = cnt_acc;
    }
    if (tile_next < 0) tile_next = bcast(tile_next_raw);
    tile = tile_next;
    tile_next_raw = (tile < total_tiles) ? grab() : 0;
  }
  cp_async_wait<0>();
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
// shapes the warp-tile kernel can run (the launcher decides whether it SHOULD: measured faster at 19 classes -- 70.7 vs 111 us
// at the bench shape -- and a tie or slightly slower at 2 classes, where the per-pixel class work it streamlines is tiny)
bool k2v2_eligible(int N, int C, int h, int w, int H, int W) {
  if (!(C == 19 || C == 2)) return false;
  if (N <= 0 || h <= 0 || w <= 0 || H < h || W < w) return false;                 // rows must advance by at most one source row
  if ((long long)C * h * w >= (1LL << 31)) return false;
  if (H < 2 || W < 2) return false;
  const float sw = ac_scale(w, W), sh = ac_scale(h, H);
  if (sh > 1.0f || sw > 1.0f) return false;
  const int tiles_x = ceil_div(W, V2_TW);
  for (int tx = 0; tx < tiles_x; ++tx) {
    const int xa = tx * V2_TW, xb = (xa + V2_TW < W ? xa + V2_TW : W) - 1;
    const int lo = host_i0(sw, xa);
    int hi = host_i0(sw, xb); hi += (hi < w - 1) ? 1 : 0;
    if (hi - lo + 1 > V2_SPAN) return false;
  }
  return true;
}

void k2v2_geometry(K2Geom& g, int N, int C, int h, int w, int H, int W) {
  k2_geometry_tiled(g, N, C, h, w, H, W, V2_TW, V2_TH);
  g.jspan_max = V2_SPAN;                                                            // block rows are [SPAN][C], fully written
}

template <int CT, bool GRAD, int LM>
static int k2v2_launch_t(const K2Params& p, cudaStream_t stream) {
  constexpr V2Smem L = v2_smem_layout(CT, LM);
  const size_t smem = (size_t)L.total * V2_WARPS;
  static bool configured = false;
  if (!configured) {
    B200SEG_CUDA(cudaFuncSetAttribute(k2v2_upsample_ce_main<CT, GRAD, LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const long long tiles = k2_tiles(p.g);
  long long grid = ceil_div_ll(tiles, V2_WARPS);
  const long long cap = (long long)num_sms() * 3;
  if (grid > cap) grid = cap;
  B200SEG_CUDA(cudaMemsetAsync(p.counter, 0, 4, stream));
  profile_begin(6, stream);
  k2v2_upsample_ce_main<CT, GRAD, LM><<<(unsigned)grid, V2_THREADS, smem, stream>>>(p);
  profile_end(6, stream);
  B200SEG_LAUNCH_CHECK();
  return B200SEG_OK;
}

template <int CT>
static int k2v2_launch_c(const K2Params& p, bool grad, cudaStream_t stream) {
  if (p.label_u8) {
    const bool aligned = (p.g.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.labels) & 3) == 0);
    if (aligned) return grad ? k2v2_launch_t<CT, true, 1>(p, stream) : k2v2_launch_t<CT, false, 1>(p, stream);
    return grad ? k2v2_launch_t<CT, true, 2>(p, stream) : k2v2_launch_t<CT, false, 2>(p, stream);
  }
  return grad ? k2v2_launch_t<CT, true, 0>(p, stream) : k2v2_launch_t<CT, false, 0>(p, stream);
}

int k2v2_main_launch(const K2Params& p, bool grad, cudaStream_t stream) {
  if (p.g.C == 19) return k2v2_launch_c<19>(p, grad, stream);
  if (p.g.C == 2) return k2v2_launch_c<2>(p, grad, stream);
  set_error("upsample_ce: the warp-tile kernel has no instantiation for %d classes", p.g.C);
  return B200SEG_ERR_UNSUPPORTED;
}

}  // namespace b200seg
