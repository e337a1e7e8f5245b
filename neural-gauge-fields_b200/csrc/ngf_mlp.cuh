// ngf_mlp.cuh — the colour MLP of rgb_decoder (TriPlane/models/networks.py:12-32) for one tile of 128 samples,
// run by a whole CTA of 256 threads.
//
//   features (3 x bilinear over the appearance channels, Field.py:97-103; x phase code for InfoInv,
//   InfoInv/models/Field.py:74-86)  ->  basis  ->  cat[., d, sin/cos(d*2^j)]  ->  64  -> ReLU -> 64 -> ReLU -> 3
//   -> sigmoid
//
// The bias-free `basis` layer is folded into mlp.0 on the host (W1' = W1[:, :F] . B), the view-direction terms
// and b1 ride in extra K columns, so layer 1 is one [128 x K1] x [K1 x 64] product (K1 = 160 / 240).
// Layers 1 and 2 run as tcgen05.mma kind::f16 (fp16 operands, fp32 accumulators in TMEM); layer 3 (64 -> 3) and
// the sigmoid run on CUDA cores in the TMEM epilogue of layer 2.  An all-CUDA-core variant with the same operand
// rounding exists for cross-checking (NGF_MLP_SIMT).
//
// Shared-memory operand layout (both A tiles and weights): tcgen05 canonical K-major, no swizzle — 8x8 fp16
// "core matrices" of 128 contiguous bytes (8 rows x 16 B); core matrices of one 8-wide K chunk are contiguous
// over rows (stride-dimension byte offset SBO = 128 B), K chunks follow each other at the leading-dimension byte
// offset LBO = rows * 16 B.  Element (row r, col k) lives at (k/8)*rows*16 + r*16 + (k%8)*2.
#pragma once
#include "ngf_common.cuh"
#include "ngf_queue.h"

namespace ngf {

constexpr int kTileM = 128;                 // samples per MLP tile == UMMA M
constexpr int kThreads = 256;               // CTA size of every kernel that calls mlp_tile
constexpr int kQueueCap = 128;              // work items of the tile in flight (power of two)
constexpr uint32_t kTmemCols = 128;         // 64 fp32 columns per layer accumulator


// ----------------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    // an MMA that never commits would otherwise spin forever and wedge the device: fail the launch instead
    if (!ok && ++spins > (1u << 24)) __trap();
  } while (!ok);
}
// Same wait for warps that are not on the critical path: back off between polls so that many waiting warps do not take
// issue slots from the single MMA-issuing thread that shares their scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(32);
    if (++spins > (1u << 22)) __trap();
  }
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {     // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE (layout_type 0), descriptor version 1 (Blackwell).
// bits [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::f16: D=f32 (bits 4-5 = 1), A=B=f16 (0), both K-major, N>>3 at bit 17, M>>4 at 24.
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t of the warp = lane base + t).
// tmem_ld32_issue + tmem_ld_wait let several loads be in flight before the wait.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float v[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float v[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  tmem_ld32_issue(taddr, v);
  tmem_ld_wait();
}

// ----------------------------------------------------------------------------------------------------------
// Shared-memory carve-up used by every kernel that runs MLP tiles.
// ----------------------------------------------------------------------------------------------------------
template <int V>
struct MlpSmem {
  static constexpr int K1 = Cfg<V>::K1;
  static constexpr int NKC = K1 / 8;                       // 8-wide K chunks in layer 1
  static constexpr uint32_t kW1Bytes = NKC * kMid * 16;    // 20480 | 30720
  static constexpr uint32_t kW2Bytes = (kMid / 8) * kMid * 16;     // 8192
  // K-group stride (LBO) of the layer-1 A operand: one 16-byte slot of padding per group rotates the shared-memory
  // banks so that the six lanes that write the six K groups of one row do not collide
  static constexpr uint32_t kLboA = kTileM * 16 + 16;
  static constexpr uint32_t kABytes = NKC * kLboA;         // 41280 | 61920
  static constexpr uint32_t kHBytes = (kMid / 8) * kTileM * 16;    // 16384
  // InfoInv aliases the hidden tile onto the (dead by then) A tile to stay within two CTAs per SM.
  static constexpr bool kAliasH = (V == 1);
  // TriPlane: sample g of the tile lives in operand row tile_row(g) (see tile_row below)
  static constexpr bool kPermuteRows = (V == 0);
  static constexpr uint32_t offW1 = 0;
  static constexpr uint32_t offW2 = offW1 + kW1Bytes;
  static constexpr uint32_t offA = offW2 + kW2Bytes;
  static constexpr uint32_t offH = kAliasH ? offA : offA + kABytes;
  // (the layer-3 constants are read from the kernel parameters, FieldDev::tail_c: no shared-memory copy)
  static constexpr uint32_t offQueue = (kAliasH ? offA + kABytes : offH + kHBytes);
  static constexpr uint32_t offPart = offQueue + kQueueCap * sizeof(QEntry);   // layer-3 partial sums [128][4]
  static constexpr uint32_t offCtl = offPart + kTileM * 16;
  static constexpr uint32_t kCtlBytes = 64;
  static constexpr uint32_t offEnd = offCtl + kCtlBytes;
};

// Which operand row (= TMEM lane) holds sample g of a tile.  The TriPlane gather gives six consecutive lanes the six
// 16-byte chunks of one sample, so a quarter warp stores chunks 0..5 of sample g and 0..1 of sample g+1 (then 2..5 | 0..3,
// 4..5 | 0..5); with rows in sample order the two partial rows collide in the shared-memory banks (bank = chunk + row
// mod 8 with the padded K-group stride): every operand store took two wavefronts per quarter warp.  Rows taken in the order
// r0, r0+6, r0+4, r0+2 (mod 8) within each group of four samples make the eight banks of every quarter warp distinct; two
// groups (r0 = 0 and r0 = 1) fill an aligned block of eight rows, so the map is a permutation of each 8-row core matrix.
template <bool PERM>
__device__ __forceinline__ int tile_row(int g) {
  return PERM ? (g & ~7) | ((((g >> 2) & 1) + 6 * (g & 3)) & 7) : g;
}
template <bool PERM>
__device__ __forceinline__ int tile_sample(int row) {
  if (!PERM) return row;
  const int res = row & 7, p = res & 1;
  return (row & ~7) | (p << 2) | (((8 - (res - p)) >> 1) & 3);
}

struct MlpCtl {               // lives at offCtl
  uint64_t bar;               // mbarrier for tcgen05.commit
  uint32_t tmem_base;
  uint32_t pad[5];
};

// Copy the packed weights into shared memory, allocate TMEM, init the mbarrier.  All threads call it.
template <int V, int IMPL>
__device__ __forceinline__ void mlp_setup(const FieldDev& f, uint8_t* smem) {
  using L = MlpSmem<V>;
  const int tid = threadIdx.x;
  const uint4* s1 = reinterpret_cast<const uint4*>(f.w1p);
  uint4* d1 = reinterpret_cast<uint4*>(smem + L::offW1);
  for (int i = tid; i < (int)(L::kW1Bytes / 16); i += kThreads) d1[i] = __ldg(s1 + i);
  const uint4* s2 = reinterpret_cast<const uint4*>(f.w2p);
  uint4* d2 = reinterpret_cast<uint4*>(smem + L::offW2);
  for (int i = tid; i < (int)(L::kW2Bytes / 16); i += kThreads) d2[i] = __ldg(s2 + i);
  // zero the A tile once: K padding chunks are never written again
  uint4* da = reinterpret_cast<uint4*>(smem + L::offA);
  for (int i = tid; i < (int)(L::kABytes / 16); i += kThreads) da[i] = make_uint4(0, 0, 0, 0);
  MlpCtl* ctl = reinterpret_cast<MlpCtl*>(smem + L::offCtl);
  if (tid == 0) {
    ctl->tmem_base = 0;
    if (IMPL == 0) {
      mbar_init(&ctl->bar, 1);
      fence_mbar_init();
    }
  }
  __syncthreads();
  if (IMPL == 0) {
    if (tid < 32) tmem_alloc(&ctl->tmem_base, kTmemCols);
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
    tc_fence_after();
  }
}

template <int IMPL>
__device__ __forceinline__ void mlp_teardown(uint8_t* smem, uint32_t ctl_off) {
  if (IMPL == 0) {
    MlpCtl* ctl = reinterpret_cast<MlpCtl*>(smem + ctl_off);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(ctl->tmem_base, kTmemCols);
  }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Blend 8 fp16 channels from 4 taps with packed half2 FMAs (texels are stored in fp16 and the result is an fp16 MMA
// operand anyway; the fp16 accumulation adds about one more rounding of 2^-11 relative to the fp32 blend).
struct __align__(16) TapsH {
  int off[4];
  __half2 w[4];
};
__device__ __forceinline__ TapsH to_half_taps(const Taps& t) {
  TapsH h;
#pragma unroll
  for (int k = 0; k < 4; ++k) { h.off[k] = t.off[k]; h.w[k] = __float2half2_rn(t.w[k]); }
  return h;
}
__device__ __forceinline__ uint4 blend8h(const __half* __restrict__ base, int chan_off, int AC, const TapsH& t) {
  __half2 acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(base + (size_t)t.off[k] * AC + chan_off));
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] = k == 0 ? __hmul2(t.w[0], h[e]) : __hfma2(t.w[k], h[e], acc[e]);
  }
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&acc[0]); o.y = *reinterpret_cast<uint32_t*>(&acc[1]);
  o.z = *reinterpret_cast<uint32_t*>(&acc[2]); o.w = *reinterpret_cast<uint32_t*>(&acc[3]);
  return o;
}

// ----------------------------------------------------------------------------------------------------------
// Fill the A tile rows [0,128) from queue slots head .. head+127.
//   dir / dir_stride : view direction of entry id is dir[id*dir_stride + 0..2]; with a camera (cam != nullptr) it is
//                      regenerated from the pixel index instead
// ----------------------------------------------------------------------------------------------------------
// q: the tile's 128 work items; A: the layer-1 operand; tapbuf: 12 KiB of scratch (TriPlane only); tid: 0..NT-1 within the
// NT threads (256 or 384) that share the work; sync(): a barrier over exactly those threads.
template <class L, int F>
__device__ __forceinline__ void mlp_view_columns(const QEntry* __restrict__ q, uint8_t* A, int tid,
                                                 const float* __restrict__ dir, int dir_stride, const CamDev* cam);

template <int V, int NT, class Sync>
__device__ __forceinline__ void mlp_gather_at(const FieldDev& f, const QEntry* __restrict__ q, uint8_t* A, TapsH* tapbuf,
                                              int tid, const float* __restrict__ dir, int dir_stride, const CamDev* cam,
                                              Sync sync) {
  using L = MlpSmem<V>;
  constexpr int AC = Cfg<V>::AC, F = Cfg<V>::F;
  constexpr uint32_t head = 0;
  if (V == 0) {
    // Phase 1: bilinear tap sets of the 128 x 3 (row, plane) pairs -> shared memory (aliases the layer-2 operand, which is
    // only written after this tile's layer-1 MMA).
    for (int it = tid; it < kTileM * 3; it += NT) {
      const int m = it & (kTileM - 1), pl = it >> 7;
      const QEntry& e = q[(head + m) & (kQueueCap - 1)];
      const PlaneDev& P = f.plane[pl];
      // offsets and weights in two arrays of 16-byte slots: a 32-byte struct per lane would cost two shared-memory
      // wavefronts per quarter warp
      const TapsH th = to_half_taps(make_taps(e.c[2 * pl], e.c[2 * pl + 1], P.W, P.H, P.wm1, P.hm1));
      reinterpret_cast<int4*>(tapbuf)[it] = *reinterpret_cast<const int4*>(th.off);
      reinterpret_cast<uint4*>(tapbuf)[kTileM * 3 + it] = *reinterpret_cast<const uint4*>(th.w);
    }
    sync();
    // Phase 2: consecutive lanes take consecutive 16-byte chunks of the SAME texel (6 chunks = 96 contiguous bytes per
    // tap), so a warp-wide load touches ~8 cache lines instead of 32.  The gather is latency-bound, so all 4 x J texel
    // loads of a plane are issued before the first one is used.
    constexpr int J = 768 / NT;                            // items per thread and plane (128 rows x 6 chunks = 768)
    // The gather is bound by load latency, so the loads run one plane ahead of the blend: item j of
    // plane pl + 1 is requested as soon as item j of plane pl has been blended and its registers are free.
    TapsH t[J];
    uint4 raw[J][4];
    auto request = [&](int pl, int j) {
      const int it = tid + NT * j, chunk = it % 6;
      *reinterpret_cast<int4*>(t[j].off) = reinterpret_cast<const int4*>(tapbuf)[pl * kTileM + it / 6];
      *reinterpret_cast<uint4*>(t[j].w) = reinterpret_cast<const uint4*>(tapbuf)[kTileM * 3 + pl * kTileM + it / 6];
      const __half* app = f.plane[pl].app;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        raw[j][k] = __ldg(reinterpret_cast<const uint4*>(app + (size_t)t[j].off[k] * AC + chunk * 8));
    };
#pragma unroll
    for (int j = 0; j < J; ++j) request(0, j);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int it = tid + NT * j, chunk = it % 6, m = it / 6;
        __half2 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __half2* h = reinterpret_cast<const __half2*>(&raw[j][k]);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[e] = k == 0 ? __hmul2(t[j].w[0], h[e]) : __hfma2(t[j].w[k], h[e], acc[e]);
        }
        if (pl < 2) request(pl + 1, j);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&acc[0]); o.y = *reinterpret_cast<uint32_t*>(&acc[1]);
        o.z = *reinterpret_cast<uint32_t*>(&acc[2]); o.w = *reinterpret_cast<uint32_t*>(&acc[3]);
        *reinterpret_cast<uint4*>(A + (size_t)(pl * 6 + chunk) * L::kLboA + tile_row<L::kPermuteRows>(m) * 16) = o;
      }
    }
  } else {
    // Phase 1 (lanes = consecutive rows, one 8-channel chunk per warp iteration, so the phase code below is warp-uniform):
    // the eight fp32 phase factors of item (row m, chunk) are parked in the operand slots that the same item fills in phase
    // 2 for planes 0 and 1 — 16 bytes each at (plane * 9 + chunk, m) — so the table needs no shared memory of its own
    for (int it = tid; it < kTileM * 9; it += NT) {
      const int m = it & (kTileM - 1), chunk = it >> 7;
      const QEntry& e = q[(head + m) & (kQueueCap - 1)];
      const float xyz[3] = {e.c[0], e.c[1], e.c[3]};
      float pe[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) pe[k] = f.infoinv ? phase_value<12>(xyz, chunk * 8 + k) : 1.f;
      *reinterpret_cast<float4*>(A + (size_t)chunk * L::kLboA + m * 16) = make_float4(pe[0], pe[1], pe[2], pe[3]);
      *reinterpret_cast<float4*>(A + (size_t)(9 + chunk) * L::kLboA + m * 16) = make_float4(pe[4], pe[5], pe[6], pe[7]);
    }
    sync();
    // Phase 2: nine consecutive lanes take the nine 16-byte chunks of the SAME texel (144 contiguous bytes per tap), as in
    // the TriPlane gather.  An item reads its own phase factors and then overwrites exactly those slots, and with the padded
    // K-group stride the stores of a quarter warp fall into banks (chunk + m) mod 8 = it mod 8: conflict-free.
    for (int it = tid; it < kTileM * 9; it += NT) {
      const int m = it / 9, chunk = it - 9 * m;
      const QEntry& e = q[(head + m) & (kQueueCap - 1)];
      // all twelve texel loads of the item are requested before the first blend (the gather is latency bound)
      Taps t[3];
      uint4 raw[3][4];
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        const PlaneDev& P = f.plane[pl];
        t[pl] = make_taps(e.c[2 * pl], e.c[2 * pl + 1], P.W, P.H, P.wm1, P.hm1);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          raw[pl][k] = __ldg(reinterpret_cast<const uint4*>(P.app + (size_t)t[pl].off[k] * AC + chunk * 8));
      }
      const float4 pe_lo = *reinterpret_cast<const float4*>(A + (size_t)chunk * L::kLboA + m * 16);
      const float4 pe_hi = *reinterpret_cast<const float4*>(A + (size_t)(9 + chunk) * L::kLboA + m * 16);
      const float pe[8] = {pe_lo.x, pe_lo.y, pe_lo.z, pe_lo.w, pe_hi.x, pe_hi.y, pe_hi.z, pe_hi.w};
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        float v[8];
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) v[k2] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                    // same order of operations as blend8
          const __half2* hh = reinterpret_cast<const __half2*>(&raw[pl][k]);
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const float2 x = __half22float2(hh[e2]);
            v[2 * e2] += t[pl].w[k] * x.x;
            v[2 * e2 + 1] += t[pl].w[k] * x.y;
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] *= pe[k];
        uint4 o = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                             pack_half2(v[6], v[7]));
        *reinterpret_cast<uint4*>(A + (size_t)(pl * 9 + chunk) * L::kLboA + m * 16) = o;
      }
    }
  }
  mlp_view_columns<L, F>(q, A, tid, dir, dir_stride, cam);
}

// view-direction columns F..F+15 of the layer-1 operand: [d(3), sin(d_k*2^j) (6), cos (6), 1]  (networks.py:27-29, 205-216)
template <class L, int F>
__device__ __forceinline__ void mlp_view_columns(const QEntry* __restrict__ q, uint8_t* A, int tid,
                                                 const float* __restrict__ dir, int dir_stride, const CamDev* cam) {
  constexpr uint32_t head = 0;
  if (tid < kTileM) {
    const int m = tid;
    const QEntry& e = q[(head + m) & (kQueueCap - 1)];
    float v[16];
    if (e.id >= 0) {
      float d[3];
      if (cam) {
        float o_unused[3];
        camera_ray(*cam, e.id, o_unused, d);
      } else {
        const float* dp = dir + (size_t)e.id * dir_stride;
        d[0] = __ldg(dp); d[1] = __ldg(dp + 1); d[2] = __ldg(dp + 2);
      }
      v[0] = d[0]; v[1] = d[1]; v[2] = d[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float a1 = d[k], a2 = d[k] * 2.f;
        if (fabsf(a1) <= 4.f) {   // unit view directions: the fast intrinsics err by ~4e-7 there (far below fp16 rounding)
          v[3 + 2 * k] = __sinf(a1); v[4 + 2 * k] = __sinf(a2);
          v[9 + 2 * k] = __cosf(a1); v[10 + 2 * k] = __cosf(a2);
        } else {
          v[3 + 2 * k] = sinf(a1); v[4 + 2 * k] = sinf(a2);
          v[9 + 2 * k] = cosf(a1); v[10 + 2 * k] = cosf(a2);
        }
      }
      v[15] = 1.f;
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = 0.f;
    }
    uint4 o0 = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                          pack_half2(v[6], v[7]));
    uint4 o1 = make_uint4(pack_half2(v[8], v[9]), pack_half2(v[10], v[11]), pack_half2(v[12], v[13]),
                          pack_half2(v[14], v[15]));
    const int r = tile_row<L::kPermuteRows>(m);
    *reinterpret_cast<uint4*>(A + (size_t)(F / 8) * L::kLboA + r * 16) = o0;
    *reinterpret_cast<uint4*>(A + (size_t)(F / 8 + 1) * L::kLboA + r * 16) = o1;
  }
}

template <int V>
__device__ __forceinline__ void mlp_gather(const FieldDev& f, uint8_t* smem, uint32_t head,
                                           const float* __restrict__ dir, int dir_stride, const CamDev* cam) {
  using L = MlpSmem<V>;
  (void)head;                                   // every caller stages the tile at slot 0
  mlp_gather_at<V, kThreads>(f, reinterpret_cast<const QEntry*>(smem + L::offQueue), smem + L::offA,
                   reinterpret_cast<TapsH*>(smem + L::offH), (int)threadIdx.x, dir, dir_stride, cam,
                   [] { __syncthreads(); });
}

struct NoBetween {
  __device__ __forceinline__ void operator()() const {}
};

// The layers of one MLP tile, after the layer-1 operand has been written to shared memory (layout L: MlpSmem<V> or the
// TMA-staged colour kernel's).  `between` is called by every thread after the layer-1 MMAs have been issued and before
// their completion is awaited (work that should hide under the tensor core, e.g. prefetching the next tile).
// Layer 3 over 32 hidden units: h = relu(acc + b2), p += h * w3.  The constants are read from the kernel parameters
// (FieldDev::tail_c) with compile-time offsets, so they are constant-bank operands of the FADD/FFMAs: no shared-memory
// wavefronts (a quarter of the kernel's shared-memory traffic when they were LDS.128 broadcasts).
template <int CH>
__device__ __forceinline__ void mlp_head_partial(const FieldDev& f, const float (&acc)[32], float& p0, float& p1, float& p2) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    constexpr int base = CH * 128;
    const float h = fmaxf(acc[j] + f.tail_c[base + 4 * j], 0.f);
    p0 += h * f.tail_c[base + 4 * j + 1];
    p1 += h * f.tail_c[base + 4 * j + 2];
    p2 += h * f.tail_c[base + 4 * j + 3];
  }
}

template <class L, int IMPL, bool ATOMIC, class Between>
__device__ __forceinline__ void mlp_layers(const FieldDev& f, uint8_t* smem, uint32_t head, uint32_t& phase,
                                           float* __restrict__ out, Between between) {
  constexpr int NKC = L::NKC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane;        // TMEM lane == tile row handled by this thread
  const int chalf = warp >> 2;                   // which 32 of the 64 hidden columns
  MlpCtl* ctl = reinterpret_cast<MlpCtl*>(smem + L::offCtl);
  uint8_t* H = smem + L::offH;
  float acc[32];

  if (IMPL == 0) {
    fence_async_smem();
    __syncthreads();
    const uint32_t tmem = ctl->tmem_base;
    constexpr uint32_t lboA = L::kLboA, sboA = 128u, lboB = kMid * 16u, sboB = 128u;
    constexpr uint32_t idesc = umma_idesc(kTileM, kMid);
    // ---- layer 1: [128 x K1] x [K1 x 64] -> TMEM columns [0,64)
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(smem + L::offA), b0 = smem_u32(smem + L::offW1);
#pragma unroll
      for (int j = 0; j < NKC / 2; ++j)
        umma_f16(tmem, umma_desc(a0 + j * 2 * lboA, lboA, sboA), umma_desc(b0 + j * 2 * kMid * 16, lboB, sboB),
                 idesc, j > 0);
      umma_commit(&ctl->bar);
    }
    between();
    mbar_wait(&ctl->bar, phase);
    phase ^= 1u;
    tc_fence_after();
    tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + chalf * 32, acc);
  } else {
    __syncthreads();
    between();
    // CUDA-core layer 1 with the same fp16 operands
    const uint8_t* A = smem + L::offA;
    const uint8_t* W1 = smem + L::offW1;
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int kc = 0; kc < NKC; ++kc) {
      uint4 araw = *reinterpret_cast<const uint4*>(A + (size_t)kc * L::kLboA + row * 16);
      const __half2* ah = reinterpret_cast<const __half2*>(&araw);
      float a[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { float2 v = __half22float2(ah[e]); a[2 * e] = v.x; a[2 * e + 1] = v.y; }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        uint4 wraw = *reinterpret_cast<const uint4*>(W1 + (size_t)kc * (kMid * 16) + (chalf * 32 + j) * 16);
        const __half2* wh = reinterpret_cast<const __half2*>(&wraw);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 w = __half22float2(wh[e]);
          acc[j] += a[2 * e] * w.x + a[2 * e + 1] * w.y;
        }
      }
    }
    if (L::kAliasH) __syncthreads();             // everyone done reading A before H (aliased) is written
  }

  // ---- epilogue 1: ReLU, round to fp16, store as the layer-2 A operand (K chunks chalf*4 .. +3)
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 o = make_uint4(pack_half2(fmaxf(acc[8 * c], 0.f), fmaxf(acc[8 * c + 1], 0.f)),
                         pack_half2(fmaxf(acc[8 * c + 2], 0.f), fmaxf(acc[8 * c + 3], 0.f)),
                         pack_half2(fmaxf(acc[8 * c + 4], 0.f), fmaxf(acc[8 * c + 5], 0.f)),
                         pack_half2(fmaxf(acc[8 * c + 6], 0.f), fmaxf(acc[8 * c + 7], 0.f)));
    *reinterpret_cast<uint4*>(H + (size_t)(chalf * 4 + c) * (kTileM * 16) + row * 16) = o;
  }

  if (IMPL == 0) {
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    const uint32_t tmem = ctl->tmem_base;
    constexpr uint32_t lboA = kTileM * 16u, sboA = 128u, lboB = kMid * 16u, sboB = 128u;
    constexpr uint32_t idesc = umma_idesc(kTileM, kMid);
    // ---- layer 2: [128 x 64] x [64 x 64] -> TMEM columns [64,128)
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(H), b0 = smem_u32(smem + L::offW2);
#pragma unroll
      for (int j = 0; j < kMid / 16; ++j)
        umma_f16(tmem + 64, umma_desc(a0 + j * 2 * kTileM * 16, lboA, sboA),
                 umma_desc(b0 + j * 2 * kMid * 16, lboB, sboB), idesc, j > 0);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase);
    phase ^= 1u;
    tc_fence_after();
    tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 64 + chalf * 32, acc);
  } else {
    __syncthreads();
    const uint8_t* W2 = smem + L::offW2;
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int kc = 0; kc < kMid / 8; ++kc) {
      uint4 araw = *reinterpret_cast<const uint4*>(H + (size_t)kc * (kTileM * 16) + row * 16);
      const __half2* ah = reinterpret_cast<const __half2*>(&araw);
      float a[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { float2 v = __half22float2(ah[e]); a[2 * e] = v.x; a[2 * e + 1] = v.y; }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        uint4 wraw = *reinterpret_cast<const uint4*>(W2 + (size_t)kc * (kMid * 16) + (chalf * 32 + j) * 16);
        const __half2* wh = reinterpret_cast<const __half2*>(&wraw);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 w = __half22float2(wh[e]);
          acc[j] += a[2 * e] * w.x + a[2 * e + 1] * w.y;
        }
      }
    }
  }

  // ---- epilogue 2: + b2, ReLU, layer 3 (64 -> 3) partial sums over this thread's 32 hidden units
  float p0 = 0.f, p1 = 0.f, p2 = 0.f;
  if (chalf == 0) mlp_head_partial<0>(f, acc, p0, p1, p2);
  else mlp_head_partial<1>(f, acc, p0, p1, p2);
  float4* part = reinterpret_cast<float4*>(smem + L::offPart);
  if (chalf == 1) part[row] = make_float4(p0, p1, p2, 0.f);
  if (IMPL == 0) tc_fence_before();
  __syncthreads();
  if (chalf == 0) {
    const QEntry* q = reinterpret_cast<const QEntry*>(smem + L::offQueue);
    // (w, id) of the sample in this row, one 8-byte load
    const float2 wi = *reinterpret_cast<const float2*>(&q[(head + tile_sample<L::kPermuteRows>(row)) & (kQueueCap - 1)].w);
    struct { float w; int id; } e = {wi.x, __float_as_int(wi.y)};
    if (e.id >= 0) {
      float4 o = part[row];
      float r = 1.f / (1.f + expf(-(p0 + o.x + f.tail_c[256])));
      float g = 1.f / (1.f + expf(-(p1 + o.y + f.tail_c[257])));
      float b = 1.f / (1.f + expf(-(p2 + o.z + f.tail_c[258])));
      float* dst = out + (size_t)e.id * 3;
      if (ATOMIC) {
        atomicAdd(dst, e.w * r);
        atomicAdd(dst + 1, e.w * g);
        atomicAdd(dst + 2, e.w * b);
      } else {
        dst[0] = r; dst[1] = g; dst[2] = b;
      }
    }
  }
  __syncthreads();
}

// One MLP tile.  All 256 threads call it with the same arguments; `phase` is the running mbarrier parity
// (register, uniform).  Output: ATOMIC -> atomicAdd(out[id*3+c], w*rgb_c);  else out[id*3+c] = rgb_c.
// The caller guarantees a __syncthreads() between the last write to the queue slots and this call; the function
// ends with a __syncthreads().
template <int V, int IMPL, bool ATOMIC>
__device__ __forceinline__ void mlp_tile(const FieldDev& f, uint8_t* smem, uint32_t head, uint32_t& phase,
                                         const float* __restrict__ dir, int dir_stride, float* __restrict__ out,
                                         const CamDev* cam = nullptr) {
  mlp_gather<V>(f, smem, head, dir, dir_stride, cam);
  mlp_layers<MlpSmem<V>, IMPL, ATOMIC>(f, smem, head, phase, out, NoBetween{});
}

}  // namespace ngf
