// ngf_infoinv_tc.cuh — the InfoInv march with its density MLP on the tensor cores.
//
// InfoInv's density is a 72 -> 32 -> 32 -> 1 MLP per kept sample (InfoInv/models/Field.py:52-70, networks.py:34-54).  Inside
// the one-ray-per-lane march it runs per lane on CUDA cores at ~19 % lane efficiency (0.78 of the 0.97 ms frame).  Here the
// march of InfoInv/models/FieldBase.py:228-282 is three kernels:
//
//   ngf_ii_find_kernel       one ray per lane, 8x4-pixel warp tiles: the skip loop of ngf_march_kernel (sample positions,
//                            bbox, occupancy — the same decision chain, bit for bit) records EVERY kept sample of the ray
//                            (normalised position, t, delta) in a device list, linked per ray in sample order.
//   ngf_ii_density_kernel    persistent CTAs, 128 samples per tile: bilinear gather of the 3 x 24 density channels (fp32
//                            texels, one (sample, float4) per thread so a texel's 96 bytes are read by six neighbouring
//                            lanes), phase code (InfoInv's sin/cos(2^k x), k < 4, same recurrence as sigma_infoinv),
//                            both MLP layers as tcgen05.mma kind::f16 with SPLIT fp16 operands — x = hi + lo for
//                            activations and weights, three MMAs per K step (hi.hi + lo.hi + hi.lo), fp32 accumulation in
//                            TMEM, biases as a constant-one K column — the 32 -> 1 head, softplus(. - 10) on CUDA cores.
//                            Plain fp16 operands are not an option: a 1 % error in sigma moves pixels by 4e-3.
//   ngf_ii_composite_kernel  one ray per lane: walk the ray's samples in order, alpha / transmittance / weights
//                            (raw2alpha, FieldBase.py:12-19), acc and depth sums, colour work items for weights above the
//                            threshold, early-out at T <= tstop.
//
// Samples behind an opaque surface get a density they never use (the march no longer knows T when it records them); the
// tensor-core evaluation is cheap enough to pay for that.
#pragma once
#include "ngf_internal.h"
#include "ngf_mlp.cuh"

namespace ngf {

struct __align__(16) IiEntry {              // 32 bytes
  float n[3];                               // normalised position; plane coordinates are (x,y), (y,z), (x,z)
  float t, delta, sigma;
  int next;                                 // next sample of the same ray (-1: last)
  int ray;
};

struct IiTcWs {
  unsigned int* counts;                     // [0] entries recorded, [1] tile counter of the find kernel, [2] overflow flag
  int* head;                                // [R] first entry of the ray, -1 if none
  IiEntry* entry;
  unsigned int cap;
};

// ---------------------------------------------------------------------------------------------------------------------
// find
// ---------------------------------------------------------------------------------------------------------------------
template <bool JIT>
__global__ void __launch_bounds__(256, 3) ngf_ii_find_kernel(const __grid_constant__ FieldDev f,
                                                             const __grid_constant__ RenderArgs a,
                                                             const __grid_constant__ IiTcWs ws) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int S = a.S;
  uint32_t st_box = 0;
  for (;;) {
    int tile = 0;
    if (lane == 0) tile = (int)atomicAdd(ws.counts + 1, 1u);
    tile = __shfl_sync(FULL, tile, 0);
    if (tile >= a.n_tiles) break;
    long long ray;
    if (a.img_w > 0) {
      const int tiles_x = (a.img_w + 7) >> 3;
      const int px = (tile % tiles_x) * 8 + (lane & 7), py = (tile / tiles_x) * 4 + (lane >> 3);
      ray = (px < a.img_w && py < a.img_h) ? (long long)py * a.img_w + px : -1;
    } else {
      ray = (long long)tile * 32 + lane;
      if (ray >= a.n_rays) ray = -1;
    }
    float o[3] = {0, 0, 0}, d[3] = {0, 0, 1}, t0 = 0.f, jit = 0.f;
    int i = 0, i_end = 0, prev = -1;
    bool live = false;
    if (ray >= 0) {
      if (a.cam_on) {
        camera_ray(a.cam, ray, o, d);
      } else {
        const float* rp = a.rays + ray * a.ray_stride;
#pragma unroll
        for (int k = 0; k < 3; ++k) { o[k] = __ldg(rp + k); d[k] = __ldg(rp + 3 + k); }
      }
      t0 = ray_t0(f, o, d);
      if (JIT) jit = __ldg(a.jitter + ray);
      int lo_i, hi_i;
      ray_index_range(f, o, d, t0, S, lo_i, hi_i, JIT ? 1.f : 0.f);
      i = lo_i; i_end = hi_i + 1;
      live = i < i_end;
      float* rgb = a.rgb + ray * 3;                        // the colour kernel accumulates into it
      rgb[0] = 0.f; rgb[1] = 0.f; rgb[2] = 0.f;
      ws.head[ray] = -1;
    }
    while (__any_sync(FULL, live)) {
      bool found = false;
      float t = 0.f, p[3] = {0, 0, 0};
      while (live && !found) {
        t = JIT ? sample_t(f, t0, i, jit) : sample_t(f, t0, i);
        bool in = sample_pos(f, o, d, t, p);
        if (in) ++st_box;
        if (in && f.has_occ) in = occ_keep(f, p);
        if (in) found = true;
        else if (++i >= i_end) live = false;
      }
      const unsigned fm = __ballot_sync(FULL, found);
      if (fm) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(ws.counts, (unsigned)__popc(fm));
        base = __shfl_sync(FULL, base, 0);
        if (found) {
          const unsigned slot = base + __popc(fm & ((1u << lane) - 1u));
          if (slot < ws.cap) {
            IiEntry e;
            unit_coords(f, p, e.n);
            const float tn = JIT ? sample_t(f, t0, i + 1, jit) : sample_t(f, t0, i + 1);
            e.t = t;
            e.delta = (i == S - 1) ? 0.f : __fmul_rn(__fsub_rn(tn, t), f.dscale);
            e.sigma = 0.f; e.next = -1; e.ray = (int)ray;
            ws.entry[slot] = e;
            if (prev >= 0) ws.entry[prev].next = (int)slot;
            else ws.head[ray] = (int)slot;
            prev = (int)slot;
          } else {
            atomicExch(ws.counts + 2, 1u);
            live = false;
          }
          if (++i >= i_end) live = false;
        }
      }
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) st_box += __shfl_xor_sync(FULL, st_box, s);
  if (lane == 0 && st_box) atomicAdd(a.stats + 0, (unsigned long long)st_box);
}

// ---------------------------------------------------------------------------------------------------------------------
// density on the tensor cores
// ---------------------------------------------------------------------------------------------------------------------
struct IiSmem {
  static constexpr uint32_t kLbo = kTileM * 16 + 16;                  // K-group stride of the A-side operands (padded)
  static constexpr uint32_t kW1 = 10 * 32 * 16, kW2 = 6 * 32 * 16;    // per hi / lo part
  static constexpr uint32_t offW1h = 0, offW1l = offW1h + kW1, offW2h = offW1l + kW1, offW2l = offW2h + kW2;
  static constexpr uint32_t offAh = offW2l + kW2, offAl = offAh + 10 * kLbo;
  static constexpr uint32_t offHh = offAl + 10 * kLbo, offHl = offHh + 6 * kLbo;   // hidden operand; tap sets alias offHh
  static constexpr uint32_t offPe = offHl + 6 * kLbo;                 // [128][24] phase code
  static constexpr uint32_t offTail = offPe + kTileM * 24 * 4;        // w3 [32], b3
  static constexpr uint32_t offPart = offTail + 36 * 4;               // [128] partial head sums
  static constexpr uint32_t offCtl = offPart + kTileM * 4;
  static constexpr uint32_t offEnd = offCtl + 64;
};
static_assert(3 * kTileM * sizeof(Taps) <= 6 * IiSmem::kLbo, "tap sets must fit the hidden operand they alias");

__device__ __forceinline__ void split_store4(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, const float v[4]) {
  __half h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h[e] = __float2half_rn(v[e]);
    l[e] = __float2half_rn(v[e] - __half2float(h[e]));
  }
  *reinterpret_cast<uint2*>(hi_base + off) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(lo_base + off) = *reinterpret_cast<const uint2*>(l);
}

__global__ void __launch_bounds__(kThreads, 2) ngf_ii_density_kernel(const __grid_constant__ FieldDev f,
                                                                     const __grid_constant__ IiTcWs ws,
                                                                     unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(128) uint8_t smem[];
  using L = IiSmem;
  uint32_t count = *reinterpret_cast<volatile const uint32_t*>(ws.counts);
  if (count > ws.cap) count = ws.cap;
  const uint32_t n_tiles = (count + kTileM - 1) / kTileM;
  if (blockIdx.x >= n_tiles) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  MlpCtl* ctl = reinterpret_cast<MlpCtl*>(smem + L::offCtl);
  // ---- set-up: split weights -> shared memory, constant columns of the operands, TMEM, barrier
  {
    const uint4* src = reinterpret_cast<const uint4*>(f.ii_w);
    uint4* dst = reinterpret_cast<uint4*>(smem + L::offW1h);
    for (int i = tid; i < (int)((2 * L::kW1 + 2 * L::kW2) / 16); i += kThreads) dst[i] = __ldg(src + i);
    float* tl = reinterpret_cast<float*>(smem + L::offTail);
    if (tid < 33) tl[tid] = __ldg(f.ii_tail + tid);
    // zero both A parts and both hidden parts once; then the constant-one columns (A col 72, H col 32) of the hi parts
    uint4* z = reinterpret_cast<uint4*>(smem + L::offAh);
    for (int i = tid; i < (int)((L::offPe - L::offAh) / 16); i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
      ctl->tmem_base = 0;
      mbar_init(&ctl->bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid < kTileM) {
      const __half one = __float2half_rn(1.f);
      *reinterpret_cast<__half*>(smem + L::offAh + 9 * L::kLbo + tid * 16) = one;       // column 72 = group 9, element 0
    }
    if (tid < 32) tmem_alloc(&ctl->tmem_base, 64);
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem = ctl->tmem_base;
  const int row = (warp & 3) * 32 + lane, chalf = warp >> 2;
  float* pe = reinterpret_cast<float*>(smem + L::offPe);
  Taps* taps = reinterpret_cast<Taps*>(smem + L::offHh);
  const float* tl = reinterpret_cast<const float*>(smem + L::offTail);
  float* part = reinterpret_cast<float*>(smem + L::offPart);
  uint32_t phase = 0, done = 0;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++done) {
    const uint32_t first = tile * kTileM;
    // ---- phase code and tap sets of the tile's 128 samples
    if (tid < kTileM) {
      const uint32_t item = first + tid;
      float n[3] = {0.f, 0.f, 0.f};
      if (item < count) { const IiEntry& e = ws.entry[item]; n[0] = e.n[0]; n[1] = e.n[1]; n[2] = e.n[2]; }
      float* p = pe + tid * 24;
      if (f.infoinv) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {                      // same sincosf + double-angle recurrence as sigma_infoinv
          float s_, c_;
          sincosf(n[c], &s_, &c_);
          p[4 * c] = s_; p[12 + 4 * c] = c_;
#pragma unroll
          for (int k = 1; k < 4; ++k) {
            const float s2 = 2.f * s_ * c_, c2 = 1.f - 2.f * s_ * s_;
            s_ = s2; c_ = c2;
            p[4 * c + k] = s_; p[12 + 4 * c + k] = c_;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 24; ++k) p[k] = 1.f;
      }
      const float uv[3][2] = {{n[0], n[1]}, {n[1], n[2]}, {n[0], n[2]}};
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        const PlaneDev& P = f.plane[pl];
        taps[pl * kTileM + tid] = make_taps(uv[pl][0], uv[pl][1], P.W, P.H, P.wm1, P.hm1);
      }
    }
    __syncthreads();
    // ---- gather: one (sample, plane, float4 of channels) per thread and step; blend, phase, split fp16, store
#pragma unroll 3
    for (int j = 0; j < 9; ++j) {
      const int it = tid + kThreads * j, m = it / 18, r = it - m * 18, pl = r / 6, q = r - pl * 6;
      const Taps t = taps[pl * kTileM + m];
      const float4* src = reinterpret_cast<const float4*>(f.plane[pl].dens) + q;
      float4 x[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) x[k] = __ldg(src + (size_t)t.off[k] * 6);
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {                        // same order of operations as sigma_infoinv
        v[0] += t.w[k] * x[k].x; v[1] += t.w[k] * x[k].y; v[2] += t.w[k] * x[k].z; v[3] += t.w[k] * x[k].w;
      }
      const float* p = pe + m * 24 + 4 * q;
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] *= p[e];
      const int col = pl * 24 + 4 * q;                     // 4 consecutive columns inside one K group
      split_store4(smem + L::offAh, smem + L::offAl, (uint32_t)(col >> 3) * L::kLbo + m * 16 + (col & 7) * 2, v);
    }
    fence_async_smem();
    __syncthreads();
    // ---- layer 1: [128 x 80] x [80 x 32], split: hi.hi + lo.hi + hi.lo -> TMEM columns [0, 32)
    constexpr uint32_t idesc = umma_idesc(kTileM, 32);
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ah = smem_u32(smem + L::offAh), al = smem_u32(smem + L::offAl);
      const uint32_t bh = smem_u32(smem + L::offW1h), bl = smem_u32(smem + L::offW1l);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const uint64_t dah = umma_desc(ah + j * 2 * L::kLbo, L::kLbo, 128), dal = umma_desc(al + j * 2 * L::kLbo, L::kLbo, 128);
        const uint64_t dbh = umma_desc(bh + j * 2 * 512, 512, 128), dbl = umma_desc(bl + j * 2 * 512, 512, 128);
        umma_f16(tmem, dah, dbh, idesc, j > 0);
        umma_f16(tmem, dal, dbh, idesc, 1u);
        umma_f16(tmem, dah, dbl, idesc, 1u);
      }
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase);
    phase ^= 1u;
    tc_fence_after();
    {
      float acc[16];
      tmem_ld16_issue(tmem + ((uint32_t)((warp & 3) * 32) << 16) + chalf * 16, acc);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 2; ++g) {                        // ReLU, split, hidden operand K groups chalf*2 + g
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaxf(acc[8 * g + e], 0.f);
        const uint32_t off = (uint32_t)(chalf * 2 + g) * L::kLbo + row * 16;
        split_store4(smem + L::offHh, smem + L::offHl, off, v);
        split_store4(smem + L::offHh, smem + L::offHl, off + 8, v + 4);
      }
      if (chalf == 0) {                                    // constant-one column 32 (group 4, element 0) for the bias of layer 2
        uint4 one = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<__half*>(&one) = __float2half_rn(1.f);
        *reinterpret_cast<uint4*>(smem + L::offHh + 4 * L::kLbo + row * 16) = one;
        *reinterpret_cast<uint4*>(smem + L::offHl + 4 * L::kLbo + row * 16) = make_uint4(0, 0, 0, 0);
      } else {
        *reinterpret_cast<uint4*>(smem + L::offHh + 5 * L::kLbo + row * 16) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(smem + L::offHl + 5 * L::kLbo + row * 16) = make_uint4(0, 0, 0, 0);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- layer 2: [128 x 48] x [48 x 32] -> TMEM columns [32, 64)
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ah = smem_u32(smem + L::offHh), al = smem_u32(smem + L::offHl);
      const uint32_t bh = smem_u32(smem + L::offW2h), bl = smem_u32(smem + L::offW2l);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint64_t dah = umma_desc(ah + j * 2 * L::kLbo, L::kLbo, 128), dal = umma_desc(al + j * 2 * L::kLbo, L::kLbo, 128);
        const uint64_t dbh = umma_desc(bh + j * 2 * 512, 512, 128), dbl = umma_desc(bl + j * 2 * 512, 512, 128);
        umma_f16(tmem + 32, dah, dbh, idesc, j > 0);
        umma_f16(tmem + 32, dal, dbh, idesc, 1u);
        umma_f16(tmem + 32, dah, dbl, idesc, 1u);
      }
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase);
    phase ^= 1u;
    tc_fence_after();
    {
      float acc[16];
      tmem_ld16_issue(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 32 + chalf * 16, acc);
      tmem_ld_wait();
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) s += fmaxf(acc[e], 0.f) * tl[chalf * 16 + e];
      if (chalf == 1) part[row] = s;
      tc_fence_before();
      __syncthreads();
      if (chalf == 0) {
        const uint32_t item = first + row;
        if (item < count) ws.entry[item].sigma = softplus_torch(s + part[row] + tl[32] + f.dshift);
      }
    }
    __syncthreads();                                       // taps alias the hidden operand; part / pe are rewritten
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
  if (tid == 0 && blockIdx.x == 0) atomicAdd(stats + 1, (unsigned long long)count);
}

// ---------------------------------------------------------------------------------------------------------------------
// composite
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ngf_ii_composite_kernel(const __grid_constant__ FieldDev f,
                                                               const __grid_constant__ RenderArgs a,
                                                               const __grid_constant__ IiTcWs ws) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_gwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  uint32_t st_col = 0;
  for (long long base = gwarp * 32; base < a.n_rays; base += n_gwarps * 32) {
    const long long ray = base + lane;
    const bool has = ray < a.n_rays;
    int e = has ? ws.head[ray] : -1;
    float T = 1.f, acc = 0.f, dep = 0.f;
    while (__any_sync(FULL, e >= 0)) {
      bool push = false;
      QEntry q;
      if (e >= 0) {
        const IiEntry s = ws.entry[e];
        const float alpha = __fsub_rn(1.f, expf(-__fmul_rn(s.sigma, s.delta)));
        const float w = __fmul_rn(alpha, T);
        T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f));
        acc += w;
        dep += w * s.t;
        push = w > f.wthres;
        if (push) {
          q.c[0] = s.n[0]; q.c[1] = s.n[1]; q.c[2] = s.n[1]; q.c[3] = s.n[2]; q.c[4] = s.n[0]; q.c[5] = s.n[2];
          q.w = w; q.id = (int)ray;
          ++st_col;
        }
        e = T <= f.tstop ? -1 : s.next;
      }
      const unsigned pm = __ballot_sync(FULL, push);
      if (pm) {
        unsigned qb = 0;
        if (lane == 0) qb = atomicAdd(a.queue_count, (unsigned)__popc(pm));
        qb = __shfl_sync(FULL, qb, 0);
        if (push) {
          const unsigned slot = qb + __popc(pm & ((1u << lane) - 1u));
          if (slot < a.queue_cap) a.queue[slot] = q;
        }
      }
    }
    if (has) {
      float last_col;
      if (a.cam_on) {
        float o[3], d[3];
        camera_ray(a.cam, ray, o, d);
        last_col = d[2];
      } else {
        last_col = __ldg(a.rays + ray * a.ray_stride + a.ray_stride - 1);
      }
      a.acc[ray] = acc;
      a.depth[ray] = dep + (1.f - acc) * last_col;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) st_col += __shfl_xor_sync(FULL, st_col, s);
  if (lane == 0 && st_col) atomicAdd(a.stats + 2, (unsigned long long)st_col);
}

}  // namespace ngf
