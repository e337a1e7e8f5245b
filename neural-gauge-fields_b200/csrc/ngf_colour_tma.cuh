// ngf_colour_tma.cuh — the colour kernel of the TriPlane render path with a TMA-staged plane gather
// (compute_rgb, TriPlane/models/Field.py:93-105: 3 x grid_sample over the 48 appearance channels -> rgb_decoder).
//
// What bounded the first colour kernel was the gather: 12 taps x 96 B per sample as 16-byte ld.global.nc requests
// (9216 per 128-sample tile), bound by L1/L2 latency with the loads of one tile exposed in front of its MMAs
// (profiles/r01_ncu_*).  Here the plane texels a tile needs come in through the TMA unit instead:
//
//   * The 32 samples of one march warp burst (one 8x4-pixel block at neighbouring depths) touch a small patch of each
//     plane.  For every (group of 32 samples, plane) the patch origin is the minimum tap coordinate of the group and ONE
//     cp.async.bulk.tensor.3d box of 48 channels x 5 x 5 texels (2400 B) is fetched into shared memory — 12 boxes per
//     tile — from a 3-D tensor map over the channels-last fp16 plane [H][W][48].  Out-of-plane texels arrive as zeros,
//     which is grid_sample's zero padding.  A group whose taps do not fit a 5x5 patch (wide footprints: planes much finer
//     than the pixel grid, samples of very different depth in one burst) keeps the direct gather for that plane.
//   * The boxes of tile t+1 are requested while tile t is in its MMAs / epilogues (its queue items are read and its tap
//     sets computed right after the layer-1 MMAs of tile t have been issued), so by the time the CTA reaches the blend of
//     tile t+1 the texels are in shared memory: the blend reads 4 x 16 B per (sample, 8-channel chunk) with ld.shared and
//     no exposed global latency.
//   * Blend (packed half2 FMAs, same rounding as before), layer-1 operand, MMAs and epilogues are those of ngf_mlp.cuh.
#pragma once
#include "ngf_mlp.cuh"

namespace ngf {

constexpr int kBoxW = 5, kBoxH = 5;                        // texels per staged patch
constexpr uint32_t kBoxBytes = kBoxW * kBoxH * 96;         // 2400: what one TMA box delivers (48 fp16 channels per texel)
constexpr uint32_t kBoxSlot = 2432;                        // 128-byte aligned slot per (group, plane)

struct __align__(16) TapC {                                // tap set of one (sample, plane), 16 bytes
  uint16_t off[4];                                         // texel index inside the patch, times 6 (units of 16 bytes)
  __half w[4];
};

struct TmaSmem {                                           // shared-memory carve-up (TriPlane only)
  static constexpr int K1 = Cfg<0>::K1;
  static constexpr int NKC = K1 / 8;
  static constexpr uint32_t kW1Bytes = NKC * kMid * 16;
  static constexpr uint32_t kW2Bytes = (kMid / 8) * kMid * 16;
  static constexpr uint32_t kLboA = kTileM * 16 + 16;
  static constexpr uint32_t kABytes = NKC * kLboA;
  static constexpr bool kPermuteRows = false;
  static constexpr bool kAliasH = true;                    // the hidden tile reuses the (dead by then) layer-1 operand
  static constexpr uint32_t offW1 = 0;
  static constexpr uint32_t offW2 = offW1 + kW1Bytes;
  static constexpr uint32_t offA = offW2 + kW2Bytes;
  static constexpr uint32_t offH = offA;
  static constexpr uint32_t offQueue = offA + kABytes;
  static constexpr uint32_t offPart = offQueue + kQueueCap * sizeof(QEntry);
  static constexpr uint32_t offTap = offPart + kTileM * 16;                       // TapC [3][128]
  static constexpr uint32_t offBox = ((offTap + 3 * kTileM * 16 + 127) / 128) * 128;   // 12 patch slots
  static constexpr uint32_t offCtl = offBox + 12 * kBoxSlot;
  static constexpr uint32_t kCtlBytes = 64;
  static constexpr uint32_t offEnd = offCtl + kCtlBytes;
};

struct TmaCtl {                 // at offCtl; the first two members are MlpCtl's
  uint64_t bar;                 // tcgen05.commit
  uint32_t tmem_base;
  uint32_t staged;              // bit (4*pl + g): patch of (group g, plane pl) was requested through the TMA
  uint64_t tma_bar;             // 4 arrivals (one per group warp) + the bytes of the requested boxes
  uint32_t pad[10];
};
static_assert(sizeof(TmaCtl) == TmaSmem::kCtlBytes, "ctl size");

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one 3-D box: coordinates (channel, x, y) of its first element; out-of-range elements are zero-filled
__device__ __forceinline__ void tma_box_3d(void* dst_smem, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// Warps 0..3 (thread = sample of the tile): read the tile's work items, build the tap sets, request the patches.
// Returns the thread's work item (held in registers until the tile becomes current).
__device__ __forceinline__ void tma_prefetch(const FieldDev& f, uint8_t* smem, const float4* __restrict__ queue,
                                             uint32_t first, uint32_t count, float4& e0, float4& e1, uint32_t& n_direct) {
  using L = TmaSmem;
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;          // g < 4
  constexpr unsigned FULL = 0xffffffffu;
  TmaCtl* ctl = reinterpret_cast<TmaCtl*>(smem + L::offCtl);
  const uint32_t item = first + (uint32_t)tid;
  e0 = make_float4(0.f, 0.f, 0.f, 0.f);
  e1 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
  if (item < count) { e0 = __ldg(queue + (size_t)item * 2); e1 = __ldg(queue + (size_t)item * 2 + 1); }
  const bool valid = __float_as_int(e1.w) >= 0;
  const float c[6] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y};
  TapC* tapbuf = reinterpret_cast<TapC*>(smem + L::offTap);
  uint32_t bytes = 0, staged = 0;
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
    const PlaneDev& P = f.plane[pl];
    // same arithmetic as make_taps (ngf_common.cuh): unnormalise, clamp (also maps NaN to a bound), floor, weights
    float ix = ((c[2 * pl] + 1.f) * 0.5f) * P.wm1, iy = ((c[2 * pl + 1] + 1.f) * 0.5f) * P.hm1;
    ix = fminf(fmaxf(ix, -2.f), (float)P.W + 1.f);
    iy = fminf(fmaxf(iy, -2.f), (float)P.H + 1.f);
    const float x0f = floorf(ix), y0f = floorf(iy), fx = ix - x0f, fy = iy - y0f;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const bool vx0 = (unsigned)x0 < (unsigned)P.W, vx1 = (unsigned)(x0 + 1) < (unsigned)P.W;
    const bool vy0 = (unsigned)y0 < (unsigned)P.H, vy1 = (unsigned)(y0 + 1) < (unsigned)P.H;
    // patch origin = minimum tap coordinate of the group's valid samples
    int xmin = valid ? x0 : 0x7fffffff, xmax = valid ? x0 : (int)0x80000000;
    int ymin = valid ? y0 : 0x7fffffff, ymax = valid ? y0 : (int)0x80000000;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      xmin = min(xmin, __shfl_xor_sync(FULL, xmin, s)); xmax = max(xmax, __shfl_xor_sync(FULL, xmax, s));
      ymin = min(ymin, __shfl_xor_sync(FULL, ymin, s)); ymax = max(ymax, __shfl_xor_sync(FULL, ymax, s));
    }
    const bool any = xmax >= xmin;                         // at least one real sample in the group
    const int span_x = any ? xmax - xmin : 1 << 20, span_y = any ? ymax - ymin : 1 << 20;
    const bool fits = span_x + 2 <= kBoxW && span_y + 2 <= kBoxH;
    TapC t;
    const int lx = valid && fits ? x0 - xmin : 0, ly = valid && fits ? y0 - ymin : 0;
    const int o = (ly * kBoxW + lx) * 6;
    t.off[0] = (uint16_t)o; t.off[1] = (uint16_t)(o + 6); t.off[2] = (uint16_t)(o + kBoxW * 6); t.off[3] = (uint16_t)(o + kBoxW * 6 + 6);
    t.w[0] = __float2half_rn((valid && vx0 && vy0) ? (1.f - fx) * (1.f - fy) : 0.f);
    t.w[1] = __float2half_rn((valid && vx1 && vy0) ? fx * (1.f - fy) : 0.f);
    t.w[2] = __float2half_rn((valid && vx0 && vy1) ? (1.f - fx) * fy : 0.f);
    t.w[3] = __float2half_rn((valid && vx1 && vy1) ? fx * fy : 0.f);
    tapbuf[pl * kTileM + tid] = t;
    if (any && !fits) ++n_direct;
    if (fits) {
      staged |= 1u << (4 * pl + g);
      if (lane == 0) {
        tma_box_3d(smem + L::offBox + (uint32_t)(pl * 4 + g) * kBoxSlot, reinterpret_cast<const uint8_t*>(f.tmap) + pl * 128, 0,
                   xmin, ymin, &ctl->tma_bar);
        bytes += kBoxBytes;
      }
    }
  }
  if (lane == 0) {
    if (staged) atomicOr(&ctl->staged, staged);
    mbar_arrive_expect_tx(&ctl->tma_bar, bytes);
  }
}

// Blend the four taps of every (sample, plane, 8-channel chunk) into the layer-1 operand.
__device__ __forceinline__ void tma_blend(const FieldDev& f, uint8_t* smem, uint32_t staged) {
  using L = TmaSmem;
  constexpr int AC = Cfg<0>::AC;
  const int tid = threadIdx.x;
  const TapC* tapbuf = reinterpret_cast<const TapC*>(smem + L::offTap);
  const QEntry* q = reinterpret_cast<const QEntry*>(smem + L::offQueue);
  uint8_t* A = smem + L::offA;
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {                          // 128 rows x 6 chunks = 768 items over 256 threads
      const int it = tid + kThreads * j, chunk = it % 6, m = it / 6, g = m >> 5;
      __half2 acc[4];
      if ((staged >> (4 * pl + g)) & 1u) {
        const TapC t = tapbuf[pl * kTileM + m];
        const uint4* box = reinterpret_cast<const uint4*>(smem + L::offBox + (uint32_t)(pl * 4 + g) * kBoxSlot) + chunk;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint4 raw = box[t.off[k]];
          const __half2* h = reinterpret_cast<const __half2*>(&raw);
          const __half2 wk = __half2half2(t.w[k]);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[e] = k == 0 ? __hmul2(wk, h[e]) : __hfma2(wk, h[e], acc[e]);
        }
      } else {                                             // patch did not fit: direct gather of this (group, plane)
        const QEntry& e = q[m];
        const PlaneDev& P = f.plane[pl];
        const TapsH t = to_half_taps(make_taps(e.c[2 * pl], e.c[2 * pl + 1], P.W, P.H, P.wm1, P.hm1));
        uint4 raw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) raw[k] = __ldg(reinterpret_cast<const uint4*>(P.app + (size_t)t.off[k] * AC + chunk * 8));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __half2* h = reinterpret_cast<const __half2*>(&raw[k]);
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) acc[e2] = k == 0 ? __hmul2(t.w[0], h[e2]) : __hfma2(t.w[k], h[e2], acc[e2]);
        }
      }
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&acc[0]); o.y = *reinterpret_cast<uint32_t*>(&acc[1]);
      o.z = *reinterpret_cast<uint32_t*>(&acc[2]); o.w = *reinterpret_cast<uint32_t*>(&acc[3]);
      *reinterpret_cast<uint4*>(A + (size_t)(pl * 6 + chunk) * L::kLboA + m * 16) = o;
    }
  }
}

// Persistent CTAs, one 128-item tile of the colour queue per iteration, patches of the next tile in flight.
__global__ void __launch_bounds__(kThreads, 2) ngf_colour_tma_kernel(const __grid_constant__ FieldDev f,
                                                                     const __grid_constant__ RenderArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  using L = TmaSmem;
  uint32_t count = *reinterpret_cast<volatile const uint32_t*>(a.queue_count);
  if (count > a.queue_cap) count = a.queue_cap;
  const uint32_t n_tiles = (count + kTileM - 1) / kTileM;
  if (blockIdx.x >= n_tiles) return;
  const int tid = threadIdx.x;
  TmaCtl* ctl = reinterpret_cast<TmaCtl*>(smem + L::offCtl);
  // ---- set-up: weights -> shared memory, TMEM, barriers (as mlp_setup, with this kernel's layout)
  {
    const uint4* s1 = reinterpret_cast<const uint4*>(f.w1p);
    uint4* d1 = reinterpret_cast<uint4*>(smem + L::offW1);
    for (int i = tid; i < (int)(L::kW1Bytes / 16); i += kThreads) d1[i] = __ldg(s1 + i);
    const uint4* s2 = reinterpret_cast<const uint4*>(f.w2p);
    uint4* d2 = reinterpret_cast<uint4*>(smem + L::offW2);
    for (int i = tid; i < (int)(L::kW2Bytes / 16); i += kThreads) d2[i] = __ldg(s2 + i);
    if (tid == 0) {
      ctl->tmem_base = 0;
      ctl->staged = 0;
      mbar_init(&ctl->bar, 1);
      mbar_init(&ctl->tma_bar, 4);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid < 32) tmem_alloc(&ctl->tmem_base, kTmemCols);
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
    tc_fence_after();
  }
  const float4* src = reinterpret_cast<const float4*>(a.queue);
  float4* q = reinterpret_cast<float4*>(smem + L::offQueue);
  uint32_t phase = 0, tma_phase = 0, done = 0;
  float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
  uint32_t n_direct = 0;                                   // (group, plane) patches that did not fit (lane 0 of warps 0-3)
  if (tid < kTileM) tma_prefetch(f, smem, src, blockIdx.x * kTileM, count, e0, e1, n_direct);
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++done) {
    if (tid < kTileM) { q[2 * tid] = e0; q[2 * tid + 1] = e1; }
    __syncthreads();                                       // work items, tap sets and the staged mask of this tile
    const uint32_t staged = ctl->staged;
    mbar_wait(&ctl->tma_bar, tma_phase);                   // the tile's patches have landed
    tma_phase ^= 1u;
    tma_blend(f, smem, staged);
    mlp_view_columns<L, Cfg<0>::F>(reinterpret_cast<const QEntry*>(q), smem + L::offA, tid, a.rays + 3, a.ray_stride,
                                   a.cam_on ? &a.cam : nullptr);
    const uint32_t next = tile + gridDim.x;
    // mlp_layers starts with a CTA barrier (every thread is past the blend: patches, tap sets and mask are free), issues
    // the layer-1 MMAs and then calls this: the next tile's items, taps and TMA requests hide under the tensor core
    auto between = [&]() {
      if (tid >= kTileM) return;                           // warps 0-3 (thread = sample of the next tile)
      if (tid == 0) ctl->staged = 0;
      asm volatile("bar.sync 1, 128;" ::: "memory");       // the mask is cleared before anyone ORs into it
      if (next < n_tiles) tma_prefetch(f, smem, src, next * kTileM, count, e0, e1, n_direct);
    };
    mlp_layers<L, 0, true>(f, smem, 0u, phase, a.rgb, between);
  }
  mlp_teardown<0>(smem, L::offCtl);
  if (tid == 0) atomicAdd(a.stats + 3, (unsigned long long)done);
  if (tid < kTileM && (tid & 31) == 0 && n_direct) atomicAdd(a.stats + 4, (unsigned long long)n_direct);
}

}  // namespace ngf
