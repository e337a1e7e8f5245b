// ngf_kernels.cu — sm_100a kernels of the render path and their launchers.
//
// The per-frame path that replaces Base.forward (TriPlane/models/FieldBase.py:251-312, InfoInv/models/
// FieldBase.py:228-282) and the chunk loop around it (TriPlane/main.py:60-71) is a family of three kernels:
//
//   ngf_march_kernel     one ray per lane, warps own 8x4-pixel tiles (dynamic tile counter).  Sample position + bbox
//                        test (sample_ray), occupancy test on the packed bit grids (AlphaGridMask), gauge lookup
//                        (compute_gauge), density (compute_density), alpha / transmittance / weight (raw2alpha) and
//                        the acc / depth sums — all in registers, nothing materialised.  Samples whose weight
//                        exceeds rayMarch_weight_thres are compacted (warp ballot -> per-warp shared-memory stage ->
//                        32-entry coalesced bursts) into a device queue of 32-byte colour work items.
//   ngf_colour_kernel    persistent CTAs take 128-item tiles of that queue: bilinear gather of the appearance
//                        texels straight into the tcgen05 A operand, the colour MLP on the tensor cores with
//                        accumulators in TMEM (ngf_mlp.cuh), sigmoid, weight * rgb accumulated per ray.
//   ngf_finalize_kernel  white background + clamp (FieldBase.py:299-302).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>

#include "ngf_handle.h"
#include "ngf_colour_tma.cuh"
#include "ngf_infoinv_march.cuh"
#include "ngf_infoinv_tc.cuh"

namespace ngf {

static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1); }
#define NGF_COUNT_LAUNCH() g_launches.fetch_add(1)

constexpr int kMarchThreads = 256;
constexpr int kMarchWarps = kMarchThreads / 32;
constexpr int kStage = 64;        // staged colour items per warp (flushed 32 at a time)

template <int V>
struct MarchSmem {
  static constexpr uint32_t offStage = 0;
  static constexpr uint32_t offDmlp = offStage + kMarchWarps * kStage * sizeof(QEntry);
  static constexpr uint32_t kBytes = offDmlp + (V == 1 ? ((kDmlpFloats * 4 + 15) / 16) * 16 : 0);
};

// JIT: training-time sampling, every sample of ray r shifted by u_r steps (a.jitter); the JIT = false instantiation is
// the evaluation kernel, unchanged.
template <int V, bool JIT = false>
__global__ void __launch_bounds__(kMarchThreads, V == 0 ? 3 : 2) ngf_march_kernel(const __grid_constant__ FieldDev f,
                                                                const __grid_constant__ RenderArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr unsigned FULL = 0xffffffffu;
  float4* stage = reinterpret_cast<float4*>(smem + MarchSmem<V>::offStage) + warp * kStage * 2;   // 2 float4 per item
  const float* dmlp_s = nullptr;
  if (V == 1) {
    float* dm = reinterpret_cast<float*>(smem + MarchSmem<V>::offDmlp);
    for (int i = tid; i < kDmlpFloats; i += kMarchThreads) dm[i] = __ldg(f.dmlp + i);
    dmlp_s = dm;
    __syncthreads();
  }

  // per-lane ray state
  float o[3] = {0, 0, 0}, d[3] = {0, 0, 1}, t0 = 0.f, T = 1.f, acc = 0.f, dep = 0.f, last_col = 0.f;
  float jit = 0.f;
  int i = 0, i_end = 0;
  long long ray = -1;
  bool live = false;
  int n_staged = 0;                                      // warp-uniform
  uint32_t st_box = 0, st_den = 0, st_col = 0;
  const int S = a.S;
  const unsigned lt_mask = (1u << lane) - 1u;

  auto flush = [&](int n) {                              // write the first n (<= 32) staged items to the queue
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(a.queue_count, (uint32_t)n);
    base = __shfl_sync(FULL, base, 0);
    float4* dst = reinterpret_cast<float4*>(a.queue + base);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = lane + 32 * h;
      if (idx < 2 * n && base + (uint32_t)(idx >> 1) < a.queue_cap) dst[idx] = stage[idx];
    }
  };

  for (;;) {
    if (!__any_sync(FULL, live)) {
      int tile = 0;
      if (lane == 0) tile = (int)atomicAdd(a.tile_counter, 1u);
      tile = __shfl_sync(FULL, tile, 0);
      if (tile >= a.n_tiles) break;
      if (a.img_w > 0) {
        const int tiles_x = (a.img_w + 7) >> 3;
        const int px = (tile % tiles_x) * 8 + (lane & 7), py = (tile / tiles_x) * 4 + (lane >> 3);
        ray = (px < a.img_w && py < a.img_h) ? (long long)py * a.img_w + px : -1;
      } else {
        ray = (long long)tile * 32 + lane;
        if (ray >= a.n_rays) ray = -1;
      }
      if (ray >= 0) {
        if (a.cam_on) {
          camera_ray(a.cam, ray, o, d);
          last_col = d[2];                               // a 6-column ray's last column (FieldBase.py:306)
        } else {
          const float* rp = a.rays + ray * a.ray_stride;
#pragma unroll
          for (int k = 0; k < 3; ++k) { o[k] = __ldg(rp + k); d[k] = __ldg(rp + 3 + k); }
          last_col = __ldg(rp + a.ray_stride - 1);
        }
        t0 = ray_t0(f, o, d);
        if (JIT) jit = __ldg(a.jitter + ray);
        int lo_i, hi_i;
        ray_index_range(f, o, d, t0, S, lo_i, hi_i, JIT ? 1.f : 0.f);
        i = lo_i; i_end = hi_i + 1;
        T = 1.f; acc = 0.f; dep = 0.f;
        live = i < i_end;
        float* rgb = a.rgb + ray * 3;                    // the colour kernel accumulates into it
        rgb[0] = 0.f; rgb[1] = 0.f; rgb[2] = 0.f;
        if (!live) {                                     // no sample can be kept: background only
          a.acc[ray] = 0.f;
          a.depth[ray] = last_col;
          ray = -1;
        }
      }
      continue;
    }
    // ---- skip empty space until this lane finds a sample that needs the field (or runs out)
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
    // ---- density + alpha compositing weights for the found samples (lanes converge here)
    bool push = false;
    float c[6], w = 0.f;
    if (found) {
      float n[3];
      unit_coords(f, p, n);
      gauge_coords(f, n, V == 0 && f.gauge_on, c);
      const float sigma = (V == 0) ? sigma_triplane(f, c) : sigma_infoinv(f, c, dmlp_s);
      ++st_den;
      // raw2alpha (FieldBase.py:12-19) with dists from the rounded t values (FieldBase.py:258, 288)
      const float tn = JIT ? sample_t(f, t0, i + 1, jit) : sample_t(f, t0, i + 1);
      const float delta = (i == S - 1) ? 0.f : __fmul_rn(__fsub_rn(tn, t), f.dscale);
      const float alpha = __fsub_rn(1.f, expf(-__fmul_rn(sigma, delta)));
      w = __fmul_rn(alpha, T);
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f));
      acc += w;
      dep += w * t;
      push = w > f.wthres;
      if (++i >= i_end || T <= f.tstop) live = false;
    }
    const unsigned pm = __ballot_sync(FULL, push);
    if (pm) {
      if (push) {
        ++st_col;
        const int slot = n_staged + __popc(pm & lt_mask);
        stage[2 * slot] = make_float4(c[0], c[1], c[2], c[3]);
        stage[2 * slot + 1] = make_float4(c[4], c[5], w, __int_as_float((int)ray));
      }
      n_staged += __popc(pm);
      __syncwarp();
      if (n_staged >= 32) {
        flush(32);
        const int rem = n_staged - 32;                   // <= 31 items move to the front of the stage
        float4 m0 = make_float4(0, 0, 0, 0), m1 = m0;
        if (lane < rem) { m0 = stage[64 + 2 * lane]; m1 = stage[64 + 2 * lane + 1]; }
        __syncwarp();
        if (lane < rem) { stage[2 * lane] = m0; stage[2 * lane + 1] = m1; }
        __syncwarp();
        n_staged = rem;
      }
    }
    if (!live && ray >= 0) {                             // ray finished: acc_map / depth_map (FieldBase.py:296,305-306)
      a.acc[ray] = acc;
      a.depth[ray] = dep + (1.f - acc) * last_col;
      ray = -1;
    }
  }
  if (n_staged > 0) flush(n_staged);

  // statistics
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    st_box += __shfl_xor_sync(FULL, st_box, s);
    st_den += __shfl_xor_sync(FULL, st_den, s);
    st_col += __shfl_xor_sync(FULL, st_col, s);
  }
  if (lane == 0) {
    atomicAdd(a.stats + 0, (unsigned long long)st_box);
    atomicAdd(a.stats + 1, (unsigned long long)st_den);
    atomicAdd(a.stats + 2, (unsigned long long)st_col);
  }
}

// Colour MLP over the compacted queue: persistent CTAs, one 128-item tile per iteration.
template <int V, int IMPL>
__global__ void __launch_bounds__(kThreads, 2) ngf_colour_kernel(const __grid_constant__ FieldDev f,
                                                                 const __grid_constant__ RenderArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  using L = MlpSmem<V>;
  uint32_t count = *reinterpret_cast<volatile const uint32_t*>(a.queue_count);
  if (count > a.queue_cap) count = a.queue_cap;
  const uint32_t n_tiles = (count + kTileM - 1) / kTileM;
  if (blockIdx.x >= n_tiles) return;
  mlp_setup<V, IMPL>(f, smem);
  __syncthreads();
  float4* q = reinterpret_cast<float4*>(smem + L::offQueue);
  const float4* src = reinterpret_cast<const float4*>(a.queue);
  uint32_t phase = 0;
  uint32_t done = 0;
  // half a work item (16 bytes) per thread; the items of the next tile are requested before this tile is processed, so
  // their global-memory latency hides under the tile instead of standing at the head of every iteration
  auto fetch = [&](uint32_t tile) {
    const uint32_t first = tile * kTileM, item = first + (threadIdx.x >> 1);
    float4 v = (threadIdx.x & 1) ? make_float4(0.f, 0.f, 0.f, __int_as_float(-1)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (item < count) v = __ldg(src + (size_t)first * 2 + threadIdx.x);
    return v;
  };
  float4 v = fetch(blockIdx.x);
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++done) {
    q[threadIdx.x] = v;
    if (tile + gridDim.x < n_tiles) v = fetch(tile + gridDim.x);
    __syncthreads();
    mlp_tile<V, IMPL, true>(f, smem, 0u, phase, a.rays + 3, a.ray_stride, a.rgb, a.cam_on ? &a.cam : nullptr);
  }
  mlp_teardown<IMPL>(smem, L::offCtl);
  if (threadIdx.x == 0) atomicAdd(a.stats + 3, (unsigned long long)done);
}

// rgb_map = clamp(sum w*rgb + [white_bg](1 - acc), 0, 1)   (FieldBase.py:297-302)
__global__ void ngf_finalize_kernel(float* __restrict__ rgb, const float* __restrict__ acc, long long n_rays,
                                    int white_bg) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * 3) return;
  float v = rgb[i];
  if (white_bg) v += 1.f - acc[i / 3];
  rgb[i] = fminf(fmaxf(v, 0.f), 1.f);
}

// Resident CTAs per SM from the kernel's own resource use (registers, shared memory, threads).
// cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for the tcgen05 kernels here although two CTAs fit
// (ncu: register and shared-memory limits both 2), so the limits are evaluated directly.
static int blocks_per_sm(const void* kern, int threads, size_t dyn_smem) {
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) { cudaGetLastError(); return 1; }
  const int regs_per_warp = ((fa.numRegs * 32 + 511) / 512) * 512;        // allocation granularity: 512 regs / warp
  const int warps = (threads + 31) / 32;
  int by_regs = 65536 / (regs_per_warp * warps);
  int by_smem = (int)((227 * 1024 + 1024) / (dyn_smem + fa.sharedSizeBytes + 1024));
  int by_threads = 2048 / threads;
  int occ = by_regs < by_smem ? by_regs : by_smem;
  if (by_threads < occ) occ = by_threads;
  if (occ > 32) occ = 32;
  return occ < 1 ? 1 : occ;
}

template <int V, bool JIT>
static cudaError_t launch_march_t(const FieldDev& f, const RenderArgs& a, int num_sms, cudaStream_t st) {
  auto kern = ngf_march_kernel<V, JIT>;
  const size_t smem = MarchSmem<V>::kBytes;
  static PerDevice<int> occ_of;
  bool fresh = false;
  int& occ = *occ_of.get(&fresh);
  if (fresh || occ == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { occ_of.retry(); return e; }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    occ = blocks_per_sm(reinterpret_cast<const void*>(kern), kMarchThreads, smem);
  }
  long long want = ((long long)a.n_tiles + kMarchWarps - 1) / kMarchWarps;
  long long grid = (long long)num_sms * occ;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kMarchThreads, smem, st>>>(f, a);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

template <int V, int IMPL>
static cudaError_t launch_colour_t(const FieldDev& f, const RenderArgs& a, int num_sms, cudaStream_t st) {
  auto kern = ngf_colour_kernel<V, IMPL>;
  const size_t smem = MlpSmem<V>::offEnd;
  static PerDevice<int> occ_of;
  bool fresh = false;
  int& occ = *occ_of.get(&fresh);
  if (fresh || occ == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { occ_of.retry(); return e; }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    occ = blocks_per_sm(reinterpret_cast<const void*>(kern), kThreads, smem);
    if (occ > 2) occ = 2;
  }
  long long grid = (long long)num_sms * occ;
  long long worst = ((long long)a.queue_cap + kTileM - 1) / kTileM;
  if (grid > worst) grid = worst;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kThreads, smem, st>>>(f, a);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

// InfoInv: three-phase cooperative march (ngf_infoinv_march.cuh).  a.ii_ws: 192 bytes per ray + 64.
template <bool JIT>
static cudaError_t launch_infoinv_march_t(const FieldDev& f, const RenderArgs& a, int num_sms, cudaStream_t st) {
  auto kern = ngf_infoinv_march_kernel<JIT>;
  const size_t smem = ((kDmlpFloats * 4 + 127) / 128) * 128;
  static PerDevice<int> occ_of;
  bool fresh = false;
  int& occ = *occ_of.get(&fresh);
  if (fresh || occ == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { occ_of.retry(); return e; }
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kIiThreads, smem);
    if (e != cudaSuccess || n < 1) { occ_of.retry(); return e != cudaSuccess ? e : cudaErrorLaunchOutOfResources; }
    occ = n;
  }
  uint8_t* base = static_cast<uint8_t*>(a.ii_ws);
  const size_t R = (size_t)a.n_rays;
  IiWs ws;
  ws.counts = reinterpret_cast<unsigned int*>(base);
  ws.ray = reinterpret_cast<IiRay*>(base + 64);
  ws.sample = reinterpret_cast<IiSample*>(base + 64 + R * 32);
  ws.list[0] = reinterpret_cast<int*>(base + 64 + R * 160);
  ws.list[1] = ws.list[0] + R;
  ws.hit = ws.list[1] + R;
  dim3 grid((unsigned)(num_sms * occ)), block(kIiThreads);
  void* args[] = {const_cast<FieldDev*>(&f), const_cast<RenderArgs*>(&a), &ws};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), grid, block, args, smem, st);
  NGF_COUNT_LAUNCH();
  return e != cudaSuccess ? e : cudaGetLastError();
}

// InfoInv: find / tensor-core density / composite (ngf_infoinv_tc.cuh)
template <bool JIT>
static cudaError_t launch_infoinv_tc_t(const FieldDev& f, const RenderArgs& a, int num_sms, cudaStream_t st) {
  uint8_t* base = static_cast<uint8_t*>(a.ii_ws);
  IiTcWs ws;
  ws.counts = reinterpret_cast<unsigned int*>(base);
  ws.head = reinterpret_cast<int*>(base + 64);
  const size_t head_bytes = (((size_t)a.n_rays * 4 + 63) / 64) * 64;
  ws.entry = reinterpret_cast<IiEntry*>(base + 64 + head_bytes);
  ws.cap = (unsigned int)((long long)a.n_rays * a.S < 0xfffffff0ll ? (long long)a.n_rays * a.S : 0xfffffff0ll);
  cudaError_t e = cudaMemsetAsync(ws.counts, 0, 64, st);
  if (e != cudaSuccess) return e;
  {
    auto kern = ngf_ii_find_kernel<JIT>;
    long long want = ((long long)a.n_tiles + 7) / 8, grid = (long long)num_sms * 3;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, 256, 0, st>>>(f, a, ws);
    NGF_COUNT_LAUNCH();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  {
    auto kern = ngf_ii_density_kernel;
    const size_t smem = IiSmem::offEnd;
    static PerDevice<int> configured;
    bool fresh = false;
    configured.get(&fresh);
    if (fresh) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { configured.retry(); return e; }
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    long long grid = (long long)num_sms * 2, worst = ((long long)ws.cap + kTileM - 1) / kTileM;
    if (grid > worst) grid = worst;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kThreads, smem, st>>>(f, ws, a.stats);
    NGF_COUNT_LAUNCH();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  {
    long long grid = (a.n_rays + 255) / 256;
    if (grid > (long long)num_sms * 8) grid = (long long)num_sms * 8;
    ngf_ii_composite_kernel<<<(unsigned)grid, 256, 0, st>>>(f, a, ws);
    NGF_COUNT_LAUNCH();
    e = cudaGetLastError();
  }
  return e;
}

cudaError_t launch_march(const FieldDev& f, const RenderArgs& a, int num_sms, cudaStream_t st) {
  if (f.variant == 1 && a.ii_ws && a.ii_tc)
    return a.jitter ? launch_infoinv_tc_t<true>(f, a, num_sms, st) : launch_infoinv_tc_t<false>(f, a, num_sms, st);
  if (f.variant == 1 && a.ii_ws) return a.jitter ? launch_infoinv_march_t<true>(f, a, num_sms, st) : launch_infoinv_march_t<false>(f, a, num_sms, st);
  if (a.jitter) return f.variant == 0 ? launch_march_t<0, true>(f, a, num_sms, st) : launch_march_t<1, true>(f, a, num_sms, st);
  return f.variant == 0 ? launch_march_t<0, false>(f, a, num_sms, st) : launch_march_t<1, false>(f, a, num_sms, st);
}

// TriPlane colour kernel with the TMA-staged gather (ngf_colour_tma.cuh)
static cudaError_t launch_colour_tma(const FieldDev& f, const RenderArgs& a, int num_sms, cudaStream_t st) {
  auto kern = ngf_colour_tma_kernel;
  const size_t smem = TmaSmem::offEnd;
  static PerDevice<int> occ_of;
  bool fresh = false;
  int& occ = *occ_of.get(&fresh);
  if (fresh || occ == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { occ_of.retry(); return e; }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    occ = blocks_per_sm(reinterpret_cast<const void*>(kern), kThreads, smem);
    if (occ > 2) occ = 2;
  }
  long long grid = (long long)num_sms * occ;
  long long worst = ((long long)a.queue_cap + kTileM - 1) / kTileM;
  if (grid > worst) grid = worst;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kThreads, smem, st>>>(f, a);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

cudaError_t launch_colour(const FieldDev& f, const RenderArgs& a, int mlp_impl, int num_sms, cudaStream_t st) {
  if (f.variant == 0 && mlp_impl == 0 && f.tmap) return launch_colour_tma(f, a, num_sms, st);
  if (f.variant == 0) return mlp_impl == 0 ? launch_colour_t<0, 0>(f, a, num_sms, st) : launch_colour_t<0, 1>(f, a, num_sms, st);
  return mlp_impl == 0 ? launch_colour_t<1, 0>(f, a, num_sms, st) : launch_colour_t<1, 1>(f, a, num_sms, st);
}

// Ray-sharded variant: local ray l of this rank is global ray ((l / block) * world + rank) * block + l % block;
// its (r, g, b, depth) row goes, as one 16-byte store, to the same position of every destination frame buffer — this
// rank's own and, in the fused mode, the peer-mapped buffers of the other ranks (st.global over NVLink: the all-gather
// is the epilogue of the render).
__global__ void ngf_finalize_shard_kernel(const float* __restrict__ rgb, const float* __restrict__ acc,
                                          const float* __restrict__ depth, long long n_local, int white_bg,
                                          const __grid_constant__ ShardOut so) {
  for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < n_local; l += (long long)gridDim.x * blockDim.x) {
    const float bg = white_bg ? 1.f - acc[l] : 0.f;
    float4 v;
    v.x = fminf(fmaxf(rgb[l * 3 + 0] + bg, 0.f), 1.f);
    v.y = fminf(fmaxf(rgb[l * 3 + 1] + bg, 0.f), 1.f);
    v.z = fminf(fmaxf(rgb[l * 3 + 2] + bg, 0.f), 1.f);
    v.w = depth[l];
    const long long g = ((l / so.block) * so.world + so.rank) * so.block + l % so.block;
    for (int d = 0; d < so.n_dst; ++d) so.dst[d][g] = v;
  }
}

cudaError_t launch_finalize_shard(const float* rgb, const float* acc, const float* depth, long long n_local,
                                  int white_bg, const ShardOut& so, cudaStream_t st) {
  if (n_local <= 0) return cudaSuccess;
  long long blocks = (n_local + 255) / 256;
  if (so.n_dst > 1 && blocks > 592) blocks = 592;        // remote stores only need enough warps to keep the links busy
  ngf_finalize_shard_kernel<<<(unsigned)blocks, 256, 0, st>>>(rgb, acc, depth, n_local, white_bg, so);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

cudaError_t launch_finalize(float* rgb, const float* acc, long long n_rays, int white_bg, cudaStream_t st) {
  if (n_rays <= 0) return cudaSuccess;
  long long n = n_rays * 3;
  ngf_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rgb, acc, n_rays, white_bg);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

// ==========================================================================================================
// Point-wise kernels (API parity with the reference's public methods)
// ==========================================================================================================

// Base.sample_ray, eval branch (FieldBase.py:118-137)
__global__ void ngf_sample_ray_kernel(const __grid_constant__ FieldDev f, const float* __restrict__ rays,
                                      long long n_rays, int stride, int S, const float* __restrict__ jitter,
                                      float* __restrict__ pts, float* __restrict__ tout, uint8_t* __restrict__ inside) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * S) return;
  long long r = idx / S;
  int i = (int)(idx - r * S);
  const float* rp = rays + r * stride;
  float o[3] = {rp[0], rp[1], rp[2]}, d[3] = {rp[3], rp[4], rp[5]};
  float t0 = ray_t0(f, o, d);
  float t = jitter ? sample_t(f, t0, i, jitter[r]) : sample_t(f, t0, i), p[3];
  bool in = sample_pos(f, o, d, t, p);
  pts[idx * 3 + 0] = p[0]; pts[idx * 3 + 1] = p[1]; pts[idx * 3 + 2] = p[2];
  tout[idx] = t;
  inside[idx] = in ? 1 : 0;
}

// AlphaGridMask.sample_alpha(pts) > 0 (FieldBase.py:33-37)
__global__ void ngf_alpha_keep_kernel(const __grid_constant__ FieldDev f, const float* __restrict__ pts, long long n,
                                      uint8_t* __restrict__ keep) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[3] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]};
  keep[i] = (!f.has_occ || occ_keep(f, p)) ? 1 : 0;
}

// AlphaGridMask.sample_alpha(pts) itself (FieldBase.py:33-37): the trilinear value of the {0,1} volume
// (grid_sample, align_corners=True, zeros padding) — sum of the corner weights of the set corners, in ATen's order of
// operations for the weights ((x1 - ix) * (y1 - iy) * (z1 - iz) ...).
__global__ void ngf_alpha_value_kernel(const __grid_constant__ FieldDev f, const float* __restrict__ pts, long long n,
                                       float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 1.f;
  if (f.has_occ) {
    const int dims[3] = {f.occ_w, f.occ_h, f.occ_d};
    float fi[3], fl[3];
    int i0[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float q = __fsub_rn(__fmul_rn(__fsub_rn(pts[i * 3 + k], f.occ_lo[k]), f.occ_inv[k]), 1.f);
      fi[k] = __fmul_rn(__fmul_rn(__fadd_rn(q, 1.f), 0.5f), (float)(dims[k] - 1));
      fl[k] = floorf(fi[k]);
      i0[k] = (int)fminf(fmaxf(fl[k], -2.f), (float)dims[k] + 1.f);
    }
    // ATen grid_sampler_3d: tnw = (ix_bse - ix) * (iy_bse - iy) * (iz_bse - iz) etc. with ix_bse = floor(ix) + 1
    const float wx[2] = {__fsub_rn(__fadd_rn(fl[0], 1.f), fi[0]), __fsub_rn(fi[0], fl[0])};
    const float wy[2] = {__fsub_rn(__fadd_rn(fl[1], 1.f), fi[1]), __fsub_rn(fi[1], fl[1])};
    const float wz[2] = {__fsub_rn(__fadd_rn(fl[2], 1.f), fi[2]), __fsub_rn(fi[2], fl[2])};
    v = 0.f;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
          if (occ_bit(f, i0[0] + dx, i0[1] + dy, i0[2] + dz)) v = __fadd_rn(v, __fmul_rn(__fmul_rn(wx[dx], wy[dy]), wz[dz]));
  }
  out[i] = v;
}

// compute_gauge (Field.py:53-75) / transform (InfoInv Field.py:43-50) on normalised coordinates
__global__ void ngf_gauge_kernel(const __grid_constant__ FieldDev f, const float* __restrict__ xyz, long long n,
                                 int gauge_on, float* __restrict__ xy, float* __restrict__ yz,
                                 float* __restrict__ xz) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float nn[3] = {xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2]}, c[6];
  gauge_coords(f, nn, f.variant == 0 && gauge_on && f.gauge[0].g != nullptr, c);
  xy[i * 2] = c[0]; xy[i * 2 + 1] = c[1];
  yz[i * 2] = c[2]; yz[i * 2 + 1] = c[3];
  xz[i * 2] = c[4]; xz[i * 2 + 1] = c[5];
}

// compute_density (Field.py:77-91 / InfoInv Field.py:52-70)
__global__ void ngf_density_kernel(const __grid_constant__ FieldDev f, const float* __restrict__ xy,
                                   const float* __restrict__ yz, const float* __restrict__ xz, long long n,
                                   float* __restrict__ sigma) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float c[6] = {xy[i * 2], xy[i * 2 + 1], yz[i * 2], yz[i * 2 + 1], xz[i * 2], xz[i * 2 + 1]};
  sigma[i] = f.variant == 0 ? sigma_triplane(f, c) : sigma_infoinv(f, c, f.dmlp);
}

// compute_alpha's field evaluation (FieldBase.py:140-156): world points -> sigma, 0 where the mask rejects
__global__ void ngf_sigma_world_kernel(const __grid_constant__ FieldDev f, const float* __restrict__ pts, long long n,
                                       int use_gauge, float* __restrict__ sigma) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[3] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]};
  float s = 0.f;
  if (!f.has_occ || occ_keep(f, p)) {
    float nn[3], c[6];
    unit_coords(f, p, nn);
    gauge_coords(f, nn, f.variant == 0 && use_gauge && f.gauge[0].g != nullptr, c);
    s = f.variant == 0 ? sigma_triplane(f, c) : sigma_infoinv(f, c, f.dmlp);
  }
  sigma[i] = s;
}

// compute_rgb (Field.py:93-105 / InfoInv Field.py:72-89): persistent CTAs, 128 rows per MLP tile
template <int V, int IMPL>
__global__ void __launch_bounds__(kThreads, 2) ngf_rgb_kernel(const __grid_constant__ FieldDev f,
                                                              const float* __restrict__ xy,
                                                              const float* __restrict__ yz,
                                                              const float* __restrict__ xz,
                                                              const float* __restrict__ dirs, long long n,
                                                              float* __restrict__ rgb) {
  extern __shared__ __align__(128) uint8_t smem[];
  using L = MlpSmem<V>;
  mlp_setup<V, IMPL>(f, smem);
  __syncthreads();
  QEntry* queue = reinterpret_cast<QEntry*>(smem + L::offQueue);
  uint32_t phase = 0;
  const long long n_tiles = (n + kTileM - 1) / kTileM;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (threadIdx.x < kTileM) {
      long long r = tile * kTileM + threadIdx.x;
      float4* dst = reinterpret_cast<float4*>(&queue[threadIdx.x]);
      if (r < n) {
        dst[0] = make_float4(xy[r * 2], xy[r * 2 + 1], yz[r * 2], yz[r * 2 + 1]);
        dst[1] = make_float4(xz[r * 2], xz[r * 2 + 1], 1.f, __int_as_float((int)r));
      } else {
        dst[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        dst[1] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
      }
    }
    __syncthreads();
    mlp_tile<V, IMPL, false>(f, smem, 0u, phase, dirs, 3, rgb);
  }
  mlp_teardown<IMPL>(smem, L::offCtl);
}

template <int V, int IMPL>
static cudaError_t launch_rgb_t(const FieldDev& f, const float* xy, const float* yz, const float* xz,
                                const float* dirs, long long n, float* rgb, int num_sms, cudaStream_t st) {
  auto kern = ngf_rgb_kernel<V, IMPL>;
  const size_t smem = MlpSmem<V>::offEnd;
  static PerDevice<int> configured;
  bool fresh = false;
  configured.get(&fresh);
  if (fresh) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { configured.retry(); return e; }
  }
  long long n_tiles = (n + kTileM - 1) / kTileM;
  long long grid = (long long)num_sms * 2;
  if (grid > n_tiles) grid = n_tiles;
  kern<<<(unsigned)grid, kThreads, smem, st>>>(f, xy, yz, xz, dirs, n, rgb);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

cudaError_t launch_rgb(const FieldDev& f, const float* xy, const float* yz, const float* xz, const float* dirs,
                       long long n, float* rgb, int mlp_impl, int num_sms, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (f.variant == 0)
    return mlp_impl == 0 ? launch_rgb_t<0, 0>(f, xy, yz, xz, dirs, n, rgb, num_sms, st)
                         : launch_rgb_t<0, 1>(f, xy, yz, xz, dirs, n, rgb, num_sms, st);
  return mlp_impl == 0 ? launch_rgb_t<1, 0>(f, xy, yz, xz, dirs, n, rgb, num_sms, st)
                       : launch_rgb_t<1, 1>(f, xy, yz, xz, dirs, n, rgb, num_sms, st);
}

static inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

cudaError_t launch_sample_ray(const FieldDev& f, const float* rays, long long n_rays, int stride, int S,
                              const float* jitter, float* pts, float* t, uint8_t* inside, cudaStream_t st) {
  long long n = n_rays * S;
  if (n <= 0) return cudaSuccess;
  ngf_sample_ray_kernel<<<blocks_for(n, 256), 256, 0, st>>>(f, rays, n_rays, stride, S, jitter, pts, t, inside);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_alpha_keep(const FieldDev& f, const float* pts, long long n, uint8_t* keep, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  ngf_alpha_keep_kernel<<<blocks_for(n, 256), 256, 0, st>>>(f, pts, n, keep);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_alpha_value(const FieldDev& f, const float* pts, long long n, float* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  ngf_alpha_value_kernel<<<blocks_for(n, 256), 256, 0, st>>>(f, pts, n, out);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_gauge(const FieldDev& f, const float* xyz, long long n, int gauge_on, float* xy, float* yz,
                         float* xz, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  ngf_gauge_kernel<<<blocks_for(n, 256), 256, 0, st>>>(f, xyz, n, gauge_on, xy, yz, xz);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_density(const FieldDev& f, const float* xy, const float* yz, const float* xz, long long n,
                           float* sigma, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  ngf_density_kernel<<<blocks_for(n, 128), 128, 0, st>>>(f, xy, yz, xz, n, sigma);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_sigma_world(const FieldDev& f, const float* pts, long long n, int use_gauge, float* sigma,
                               cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  ngf_sigma_world_kernel<<<blocks_for(n, 128), 128, 0, st>>>(f, pts, n, use_gauge, sigma);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

// ==========================================================================================================
// Packing kernels: reference NCHW fp32 parameters -> gather-friendly shadows
// ==========================================================================================================
// plane [C][H][W] -> dens [H][W][DC] fp32 + app [H][W][C-DC] fp16
__global__ void ngf_pack_plane_kernel(const float* __restrict__ src, int C, int H, int W, int DC,
                                      float* __restrict__ dens, __half* __restrict__ app) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over C*H*W, x fastest (coalesced reads)
  long long hw = (long long)H * W;
  if (idx >= hw * C) return;
  int c = (int)(idx / hw);
  long long t = idx - (long long)c * hw;
  float v = src[idx];
  if (c < DC) dens[t * DC + c] = v;
  else app[t * (C - DC) + (c - DC)] = __float2half_rn(v);
}

// gauge [2][H][W] -> float2 [H][W]
__global__ void ngf_pack_gauge_kernel(const float* __restrict__ src, long long hw, float2* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  out[i] = make_float2(src[i], src[hw + i]);
}

// occupancy volume fp32 (>0 == occupied) -> bits
__global__ void ngf_pack_occ_kernel(const float* __restrict__ vol, long long n_vox, uint32_t* __restrict__ bits) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w * 32 >= n_vox) return;
  uint32_t m = 0;
  for (int b = 0; b < 32; ++b) {
    long long i = w * 32 + b;
    if (i < n_vox && vol[i] > 0.f) m |= 1u << b;
  }
  bits[w] = m;
}

// raw bits -> occ2: cell (X,Y,Z), X = x0+1 in [0,W], holds OR of the raw bits at (x0..x0+1, y0..y0+1, z0..z0+1)
// (out-of-range corners read as 0); 4x4x2 cells per word.  One thread per occ2 cell.
__global__ void ngf_pack_occ2_kernel(const uint32_t* __restrict__ bits, int W, int H, int D,
                                     uint32_t* __restrict__ occ2, int nxb, int nyb, uint32_t* __restrict__ coarse,
                                     int cx, int cy, int* __restrict__ bbox) {
  const long long n = (long long)(W + 1) * (H + 1) * (D + 1);
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int X = (int)(idx % (W + 1)), Y = (int)((idx / (W + 1)) % (H + 1)), Z = (int)(idx / ((long long)(W + 1) * (H + 1)));
  bool any = false;
  for (int dz = 0; dz < 2; ++dz)
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const int x = X - 1 + dx, y = Y - 1 + dy, z = Z - 1 + dz;
        if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H && (unsigned)z < (unsigned)D) {
          const uint32_t b = ((uint32_t)z * (uint32_t)H + (uint32_t)y) * (uint32_t)W + (uint32_t)x;
          any = any || ((bits[b >> 5] >> (b & 31)) & 1u);
        }
      }
  if (!any) return;
  const uint32_t word = ((uint32_t)(Z >> 1) * (uint32_t)nyb + (uint32_t)(Y >> 2)) * (uint32_t)nxb + (uint32_t)(X >> 2);
  atomicOr(occ2 + word, 1u << ((X & 3) | ((Y & 3) << 2) | ((Z & 1) << 4)));
  const uint32_t cb = ((uint32_t)(Z >> 3) * (uint32_t)cy + (uint32_t)(Y >> 3)) * (uint32_t)cx + (uint32_t)(X >> 3);
  atomicOr(coarse + (cb >> 5), 1u << (cb & 31));
  atomicMin(bbox + 0, X); atomicMin(bbox + 1, Y); atomicMin(bbox + 2, Z);
  atomicMax(bbox + 3, X); atomicMax(bbox + 4, Y); atomicMax(bbox + 5, Z);
}

// dsum[t] = sum_c dens[t][c] * w[c]
__global__ void ngf_pack_dsum_kernel(const float* __restrict__ dens, long long hw, int DC, const float* __restrict__ w,
                                     float* __restrict__ dsum) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= hw) return;
  float s = 0.f;
  for (int c = 0; c < DC; ++c) s += dens[t * DC + c] * w[c];
  dsum[t] = s;
}

cudaError_t launch_pack_plane(const float* nchw, int C, int H, int W, int DC, float* dens, __half* app,
                              cudaStream_t st) {
  long long n = (long long)C * H * W;
  ngf_pack_plane_kernel<<<blocks_for(n, 256), 256, 0, st>>>(nchw, C, H, W, DC, dens, app);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_pack_gauge(const float* nchw, int H, int W, float2* out, cudaStream_t st) {
  long long hw = (long long)H * W;
  ngf_pack_gauge_kernel<<<blocks_for(hw, 256), 256, 0, st>>>(nchw, hw, out);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_pack_occ(const float* vol, long long n_vox, uint32_t* bits, cudaStream_t st) {
  long long words = (n_vox + 31) / 32;
  ngf_pack_occ_kernel<<<blocks_for(words, 256), 256, 0, st>>>(vol, n_vox, bits);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_pack_occ2(const uint32_t* bits, int W, int H, int D, uint32_t* occ2, int nxb, int nyb, int nzb,
                             uint32_t* coarse, int cx, int cy, int cz, int* bbox, cudaStream_t st) {
  (void)nzb; (void)cz;
  const long long n = (long long)(W + 1) * (H + 1) * (D + 1);
  ngf_pack_occ2_kernel<<<blocks_for(n, 256), 256, 0, st>>>(bits, W, H, D, occ2, nxb, nyb, coarse, cx, cy, bbox);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_pack_dsum(const float* dens, long long hw, int DC, const float* w_dev, float* dsum, cudaStream_t st) {
  ngf_pack_dsum_kernel<<<blocks_for(hw, 256), 256, 0, st>>>(dens, hw, DC, w_dev, dsum);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

// ==========================================================================================================
// Frame post-processing: what evaluation() does on the CPU after .cpu() (TriPlane/main.py:99-116)
//   u8  = (rgb * 255).astype('uint8')      (fp32 multiply, truncation)
//   sse = sum((rgb - gt)^2)                 (PSNR = -10 ln(sse / n) / ln 10 on the host)
// ==========================================================================================================
__global__ void ngf_frame_post_kernel(const float* __restrict__ rgb, const float* __restrict__ gt, long long n,
                                      uint8_t* __restrict__ u8, double* __restrict__ sse) {
  double local = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = rgb[i];
    if (u8) u8[i] = (uint8_t)__fmul_rn(v, 255.f);
    if (gt) {
      const float d = __fsub_rn(v, gt[i]);
      local += (double)__fmul_rn(d, d);
    }
  }
  if (!gt) return;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    atomicAdd(sse, t);
  }
}

cudaError_t launch_frame_post(const float* rgb, const float* gt, long long n, uint8_t* u8, double* sse, int num_sms,
                              cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
  ngf_frame_post_kernel<<<(unsigned)blocks, 256, 0, st>>>(rgb, gt, n, u8, sse);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

// Depth visualisation of evaluation() (TriPlane/main.py:102 -> utils.py:32-47 visualize_depth_numpy with minmax = near_far):
//   x = nan_to_num(depth); x = (x - mi) / (ma - mi + 1e-8); u8 = (255 * x).astype(uint8); cv2.applyColorMap(u8, COLORMAP_JET)
// in numpy's fp32 arithmetic (the Python scalars are cast to fp32), the float -> uint8 cast truncating like the host's.
#include "ngf_jet_lut.h"
__constant__ unsigned char c_jet_lut[256 * 3];
__global__ void ngf_depth_colormap_kernel(const float* __restrict__ depth, long long n, float mi, float den,
                                          uint8_t* __restrict__ bgr) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float x = depth[i];
    if (isnan(x)) x = 0.f;
    else if (isinf(x)) x = x > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
    const float v = __fmul_rn(255.f, __fdiv_rn(__fsub_rn(x, mi), den));
    const int idx = (int)v & 255;                          // cvttss2si + low byte, as numpy's astype(uint8) on the host
    bgr[i * 3 + 0] = c_jet_lut[idx * 3 + 0];
    bgr[i * 3 + 1] = c_jet_lut[idx * 3 + 1];
    bgr[i * 3 + 2] = c_jet_lut[idx * 3 + 2];
  }
}

cudaError_t launch_depth_colormap(const float* depth, long long n, float mi, float den, uint8_t* bgr, int num_sms,
                                  cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  static PerDevice<int> lut_done;
  bool fresh = false;
  lut_done.get(&fresh);
  if (fresh) {
    cudaError_t e = cudaMemcpyToSymbol(c_jet_lut, kJetLutBgr, sizeof(kJetLutBgr));
    if (e != cudaSuccess) { lut_done.retry(); return e; }
  }
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
  ngf_depth_colormap_kernel<<<(unsigned)blocks, 256, 0, st>>>(depth, n, mi, den, bgr);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

// ==========================================================================================================
// Ray sharding: ray g belongs to rank (g / block) % world, local index
// (g / (block*world)) * block + g % block.
// ==========================================================================================================
__global__ void ngf_shard_gather_kernel(const float* __restrict__ src, long long n_rays, int width, int block,
                                        int rank, int world, long long n_local, float* __restrict__ dst) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_local * width) return;
  long long l = idx / width;
  int c = (int)(idx - l * width);
  long long g = ((l / block) * world + rank) * block + (l % block);
  dst[idx] = src[g * width + c];
}

__global__ void ngf_shard_scatter_kernel(const float* __restrict__ src, long long n_rays, int width, int block,
                                         int world, long long max_shard, float* __restrict__ dst) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * width) return;
  long long g = idx / width;
  int c = (int)(idx - g * width);
  long long blk = g / block;
  int rank = (int)(blk % world);
  long long l = (blk / world) * block + (g % block);
  dst[idx] = src[((long long)rank * max_shard + l) * width + c];
}

static long long shard_count(long long n, int block, int rank, int world) {
  long long per_cycle = (long long)block * world;
  long long full = n / per_cycle, rem = n - full * per_cycle;
  long long extra = rem - (long long)rank * block;
  if (extra < 0) extra = 0;
  if (extra > block) extra = block;
  return full * block + extra;
}

cudaError_t launch_shard_gather(const float* src, long long n_rays, int width, int block, int rank, int world,
                                float* dst, cudaStream_t st) {
  long long nl = shard_count(n_rays, block, rank, world);
  if (nl <= 0) return cudaSuccess;
  ngf_shard_gather_kernel<<<blocks_for(nl * width, 256), 256, 0, st>>>(src, n_rays, width, block, rank, world, nl, dst);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_shard_scatter(const float* src, long long n_rays, int width, int block, int world,
                                 long long max_shard, float* dst, cudaStream_t st) {
  if (n_rays <= 0) return cudaSuccess;
  ngf_shard_scatter_kernel<<<blocks_for(n_rays * width, 256), 256, 0, st>>>(src, n_rays, width, block, world,
                                                                            max_shard, dst);
  NGF_COUNT_LAUNCH();
  return cudaGetLastError();
}

}  // namespace ngf
