// ngf_neutex.cu — sm_100a kernels of the UV-Mapping (NeuTex) render path (see ngf_neutex.cuh for the reference map).
//
//   ntx_raygen_kernel   cube_ray_generation (model/renderer.py:79-141), one ray per thread: slab test, jittered
//                       segment lengths, end points by a double-precision running sum (torch's CPU cumsum accumulates
//                       in double), mid points, p = campos + raydir * mid, strict in-cube test.  In-cube samples are
//                       compacted (warp prefix sum, one atomic per warp) into a work list.
//   ntx_mlp_kernel      persistent CTAs (one per SM), 256-sample tiles.  The 25-layer weight stream is copied ring stage
//                       after ring stage (cp.async.bulk -> mbarrier ring) by a producer warp; every K=16 slice feeds two
//                       M=128 tcgen05.mma tiles whose accumulators fill the SM's TMEM (2 x 256 fp32 columns); 512
//                       worker threads (two per sample row, half of a layer's columns each) build the sinusoidal
//                       encodings, run the epilogues TMEM -> activation -> fp16 -> shared-memory A operand of the next
//                       layer, and evaluate the narrow heads (density, uv, colour) in fp32 from weights staged in
//                       shared memory.  Biases ride on the MMA (a constant-one column times a bias slice).  The gauge
//                       network runs split-fp16 (hi + lo, 3 MMAs per slice) because its output is multiplied by 2^9
//                       inside PE(uv, 10).  The MMA-issue loop is written for the tensor pipe's short queue: whole ring
//                       stages as straight-line code, the wait for the next stage under the last slice's MMAs
//                       (DESIGN.md 4.3).  NGF_NTX_CG=2 (when packing) runs CTA pairs with cta_group::2 MMAs instead.
//   ntx_march_kernel    ray_march + alpha_blend + background + simple_tone_map (model/renderer.py:4-11,176-247;
//                       model/model.py:46-50), one ray per thread, transmittance as a double running product.
#include "ngf_neutex.cuh"

#include <atomic>

#include "ngf_internal.h"
#include "ngf_mlp.cuh"

namespace ngf {
uint64_t launch_count();
void count_launch();
namespace ntx {

// ---------------------------------------------------------------------------------------------------------
// PTX helpers beyond ngf_mlp.cuh
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ bool elect_one() {      // one lane of a converged warp
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- CTA-pair (cta_group::2) helpers: the two CTAs of a cluster run one M=256 MMA together; only rank 0 issues ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {        // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope
  uint32_t ok = 0, spins = 0;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* dst_smem, uint32_t cols) {   // the same warp of both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D (128 lanes x N columns in each CTA) += A (each CTA's own 128 rows) . B (N rows: rank 0 holds [0, N/2), rank 1 the rest)
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in both CTAs of the pair once every MMA issued so far has completed
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// shared-memory map of ntx_mlp_kernel
constexpr uint32_t kKGroupBytes = kRows * 16;            // one 8-wide K group of the 256-row A operand
constexpr uint32_t offA = 0;                             // [32 K groups][256 rows][8 halves]  (split: lo half at +64 KiB)
constexpr uint32_t kABytes = 32 * kKGroupBytes;          // 131072
constexpr uint32_t offA2 = offA + kABytes;               // view-direction operand, K = 48: 39 encoding columns, two
constexpr uint32_t kA2Bytes = 6 * kKGroupBytes;          //   constant-one columns (39, 40) that carry the biases, 7 zeros
constexpr uint32_t offRing = offA2 + kA2Bytes;
constexpr uint32_t offHx = offRing + kStages * kStageBytes;   // head partial sums exchanged between a row's two threads
constexpr uint32_t kHxBytes = 2 * kRows * 12;                 //   [2 column halves][256 rows][3 floats]
constexpr uint32_t offHw = offHx + kHxBytes;                  // fp32 weights of the head layer in flight (<= 3 x 256)
constexpr uint32_t kHwBytes = 3 * 256 * 4 + 16;                // + the head's bias (<= 3 floats) at float index kHwBias
constexpr int kHwBias = 3 * 256;
constexpr uint32_t offBar = offHw + kHwBytes;
static_assert(offBar + 256 <= 227 * 1024, "shared memory budget");
constexpr uint32_t kSmemBytes = offBar + 256;
constexpr uint32_t kALoOff = 16 * kKGroupBytes;          // lo operand of split layers (K <= 128)
constexpr int kOnesCol = 39;                             // A2 columns 39 and 40 are 1.0

struct Bars {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t acc_ready;
  uint64_t a_ready;
  uint64_t peer_full[kStages];   // CTA pairs, rank 0 only: rank 1's ring stage has landed (relayed by rank 1)
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkerThreads) : "memory"); }

// ---------------------------------------------------------------------------------------------------------
// A-operand writers.  Element (row r, column k) lives at (k / 8) * kKGroupBytes + r * 16 + (k % 8) * 2.
// ---------------------------------------------------------------------------------------------------------
template <bool SPLIT>
__device__ __forceinline__ void store_group(uint8_t* A, int group, int row, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = v[2 * e], b = v[2 * e + 1];
    __half2 h = __floats2half2_rn(a, b);
    hi[e] = *reinterpret_cast<uint32_t*>(&h);
    if (SPLIT) {
      float2 back = __half22float2(h);
      __half2 l = __floats2half2_rn(a - back.x, b - back.y);
      lo[e] = *reinterpret_cast<uint32_t*>(&l);
    }
  }
  *reinterpret_cast<uint4*>(A + group * kKGroupBytes + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (SPLIT) *reinterpret_cast<uint4*>(A + kALoOff + group * kKGroupBytes + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// sin and cos of |x| <~ 2000 to about 1 ulp: Cody-Waite reduction by pi/2 in three FMA steps, then the single-precision
// minimax polynomials on [-pi/4, pi/4] (Cephes sinf / cosf coefficients).  ~35 instructions for both values; sincosf's
// generic path (with its Payne-Hanek fallback in local memory) costs several times that.
__device__ __forceinline__ void sincos_cw(float x, float& s, float& c) {
  const float kf = rintf(x * 0.63661977236758134f);
  float r = fmaf(kf, -1.57079637050628662109375f, x);
  r = fmaf(kf, 4.37113900018624283e-8f, r);
  r = fmaf(kf, 1.71512451805164e-15f, r);
  const float r2 = r * r;
  const float sp = fmaf(r * r2, fmaf(r2, fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), r);
  const float cp = fmaf(r2 * r2, fmaf(r2, fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f),
                        fmaf(r2, -0.5f, 1.f));
  const int q = (int)kf & 3;
  const float s0 = (q & 1) ? cp : sp, c0 = (q & 1) ? sp : cp;
  s = (q & 2) ? -s0 : s0;
  c = ((q + 1) & 2) ? -c0 : c0;
}

// One fp16 hi (+ lo) element of the A operand
__device__ __forceinline__ void store_split_elem(uint8_t* A, int col, int row, float v) {
  const __half h = __float2half_rn(v);
  const uint32_t off = (uint32_t)(col >> 3) * kKGroupBytes + (uint32_t)row * 16u + (uint32_t)(col & 7) * 2u;
  *reinterpret_cast<__half*>(A + off) = h;
  *reinterpret_cast<__half*>(A + kALoOff + off) = __float2half_rn(v - __half2float(h));
}

// Gauge-network input [p, sin(p_d 2^f), cos(p_d 2^f)] (63 columns + one zero), split fp16, every value accurate to ~1e-7:
// the row's two threads take 15 of the 30 (coordinate, octave) pairs each and write sin and cos of a pair from one
// sincos_cw call.
__device__ __forceinline__ void write_gauge_encoding(uint8_t* A, int row, int ch, const float* x) {
#pragma unroll
  for (int j = 0; j < 15; ++j) {
    const int i = ch * 15 + j;                      // pair index = d * 10 + f
    float s, c;
    const int d0 = i / 10;
    const float xd = d0 == 0 ? x[0] : (d0 == 1 ? x[1] : x[2]);
    sincos_cw(xd * (float)(1 << (i % 10)), s, c);
    store_split_elem(A, 3 + i, row, s);
    store_split_elem(A, 33 + i, row, c);
  }
  if (ch == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) store_split_elem(A, d, row, x[d]);
  } else {
    store_split_elem(A, 63, row, 0.f);
  }
}

// Encoding [x, sin(x_d * 2^f), cos(x_d * 2^f)] (column order of util.py:427-438: d-major, then f), zero padded to
// 8*NG columns; with ONES, columns kOnesCol and kOnesCol+1 are 1.0.  The row's two threads each write half of the K
// groups (CH = 0 / 1).  ACCURATE: every value from sinf / cosf (gauge network, 1e-6 budget); otherwise one sincosf per
// coordinate and the double-angle recurrence (error grows ~2x per octave to <= 3e-5 at 2^9, far below the fp16
// rounding of the operand it feeds).
template <int D, int F, int NG, bool SPLIT, bool ACCURATE, bool ONES, int CH>
__device__ __forceinline__ void write_encoding_half(uint8_t* A, int row, const float* x) {
  float sn[D * F], cs[D * F];
  if (!ACCURATE) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float s, c;
      sincosf(x[d], &s, &c);
      sn[d * F] = s; cs[d * F] = c;
#pragma unroll
      for (int f = 1; f < F; ++f) {
        const float s2 = 2.f * s * c, c2 = 1.f - 2.f * s * s;
        s = s2; c = c2;
        sn[d * F + f] = s; cs[d * F + f] = c;
      }
    }
  }
  constexpr int G0 = CH * (NG / 2), G1 = G0 + NG / 2;
#pragma unroll
  for (int g = G0; g < G1; ++g) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = 8 * g + e;
      float val = 0.f;
      if (col < D) val = x[col];
      else if (col < D + D * F) {
        const int i = col - D;
        val = ACCURATE ? sinf(x[i / F] * (float)(1 << (i % F))) : sn[i];
      } else if (col < D + 2 * D * F) {
        const int i = col - D - D * F;
        val = ACCURATE ? cosf(x[i / F] * (float)(1 << (i % F))) : cs[i];
      } else if (ONES && (col == kOnesCol || col == kOnesCol + 1)) val = 1.f;
      v[e] = val;
    }
    store_group<SPLIT>(A, g, row, v);
  }
}
template <int D, int F, int NG, bool SPLIT, bool ACCURATE, bool ONES>
__device__ __forceinline__ void write_encoding(uint8_t* A, int row, int ch, const float* x) {
  if (ch == 0) write_encoding_half<D, F, NG, SPLIT, ACCURATE, ONES, 0>(A, row, x);
  else write_encoding_half<D, F, NG, SPLIT, ACCURATE, ONES, 1>(A, row, x);
}

// ---------------------------------------------------------------------------------------------------------
// Epilogue of one layer for one (row, column half): TMEM (bias already accumulated by the bias slice) -> activation ->
// fp16 A operand of the next layer and/or fp32 partial dot products with up to 3 head weight rows.
//   MODE 0: activation on packed halves, write   MODE 1: fp32 activation, split hi/lo write   MODE 2: no write
// ---------------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ uint32_t act_pack(float a, float b) {
  if (ACT == 0) {      // ReLU folded into the conversion: max(rn(x), 0) == rn(max(x, 0))
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
  __half2 h = __floats2half2_rn(a, b);
  h = __hmax2(h, __hmul2(h, __float2half2_rn(0.2f)));
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int N, int ACT, int MODE, int NH>
__device__ __forceinline__ void epilogue(uint32_t taddr, uint8_t* A, int row, int ch, const float* __restrict__ headw,
                                         float* hacc) {
  constexpr int NC = N / 2;                     // columns of this thread
  const int col0 = ch * NC;
  if (MODE == 0 && NH == 0) {
    static_assert(NC % 64 == 0 || NC == 32, "column count");
    constexpr int STEP = NC >= 64 ? 64 : 32;
#pragma unroll 1
    for (int c0 = 0; c0 < NC; c0 += STEP) {
      float v[STEP];
      tmem_ld32_issue(taddr + col0 + c0, v);
      if (STEP == 64) tmem_ld32_issue(taddr + col0 + c0 + 32, v + 32);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < STEP / 8; ++g) {
        const uint4 o = make_uint4(act_pack<ACT>(v[8 * g], v[8 * g + 1]), act_pack<ACT>(v[8 * g + 2], v[8 * g + 3]),
                                   act_pack<ACT>(v[8 * g + 4], v[8 * g + 5]), act_pack<ACT>(v[8 * g + 6], v[8 * g + 7]));
        *reinterpret_cast<uint4*>(A + ((col0 + c0) / 8 + g) * kKGroupBytes + row * 16) = o;
      }
    }
  } else {
#pragma unroll 1
    for (int c0 = 0; c0 < NC; c0 += 32) {
      float v[32];
      tmem_ld32(taddr + col0 + c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = ACT == 0 ? fmaxf(v[j], 0.f) : fmaxf(v[j], 0.2f * v[j]);
      if (NH > 0) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          float2 s2 = make_float2(hacc[h], 0.f);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(headw + h * N + col0 + c0) + q);
            s2 = __ffma2_rn(make_float2(v[4 * q], v[4 * q + 1]), make_float2(w.x, w.y), s2);
            s2 = __ffma2_rn(make_float2(v[4 * q + 2], v[4 * q + 3]), make_float2(w.z, w.w), s2);
          }
          hacc[h] = s2.x + s2.y;
        }
      }
      if (MODE != 2) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (MODE == 1) store_group<true>(A, (col0 + c0) / 8 + g, row, v + 8 * g);
          else store_group<false>(A, (col0 + c0) / 8 + g, row, v + 8 * g);
        }
      }
    }
  }
}

// Epilogue of a layer that feeds a narrow fp32 head (density, uv, color1, block2 output): fp32 activation, NH partial dot
// products over this thread's NC columns, optionally the fp16 A operand of the next layer (WRITE_A).  The head's weight
// rows [NH][N] were staged in shared memory by stage_head() while the layer's MMAs ran: every lane reads the same
// address (one broadcast LDS.128 per four weights).  Read from global memory they cost ~10 000 cycles per head layer:
// with 227 KB of shared memory carved out nothing stays in L1, so each of the 96 loads was an L2 round trip; as
// constant-bank operands (kernel parameters) the 3 KB per head thrashed the constant cache just the same.
template <int N, int ACT, bool WRITE_A, int NH, int CH>
__device__ __forceinline__ void epilogue_head_ch(const float* __restrict__ sw, uint32_t taddr, uint8_t* A, int row,
                                                 float* hacc) {
  constexpr int NC = N / 2, col0 = CH * NC;
  float2 part[NH][2];        // packed fp32 FMAs (two exact fp32 FMAs per instruction on sm_100)
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int q = 0; q < 2; ++q) part[h][q] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int c0 = 0; c0 < NC; c0 += 32) {
    float v[32];
    tmem_ld32(taddr + col0 + c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ACT == 0 ? fmaxf(v[j], 0.f) : fmaxf(v[j], 0.2f * v[j]);
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(sw + h * N + col0 + c0 + 4 * q);
        part[h][0] = __ffma2_rn(make_float2(v[4 * q], v[4 * q + 1]), make_float2(w.x, w.y), part[h][0]);
        part[h][1] = __ffma2_rn(make_float2(v[4 * q + 2], v[4 * q + 3]), make_float2(w.z, w.w), part[h][1]);
      }
    if (WRITE_A) {
#pragma unroll
      for (int g = 0; g < 4; ++g) store_group<false>(A, (col0 + c0) / 8 + g, row, v + 8 * g);
    }
  }
#pragma unroll
  for (int h = 0; h < NH; ++h) hacc[h] += (part[h][0].x + part[h][0].y) + (part[h][1].x + part[h][1].y);
}
template <int N, int ACT, bool WRITE_A, int NH>
__device__ __forceinline__ void epilogue_head(const float* sw, uint32_t taddr, uint8_t* A, int row, int ch, float* hacc) {
  if (ch == 0) epilogue_head_ch<N, ACT, WRITE_A, NH, 0>(sw, taddr, A, row, hacc);
  else epilogue_head_ch<N, ACT, WRITE_A, NH, 1>(sw, taddr, A, row, hacc);
}

// sample_square (util.py:277-282): bilinear, align_corners=False, border padding, over tex [h][w][c]
__device__ __forceinline__ void sample_texture(const NetDev& net, float u, float v, float out[3]) {
  const int W = net.tex_w, H = net.tex_h, C = net.tex_c;
  float ix = ((u + 1.f) * (float)W - 1.f) * 0.5f, iy = ((v + 1.f) * (float)H - 1.f) * 0.5f;
  ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
  const float fx = ix - x0f, fy = iy - y0f;
  const float w00 = (1.f - fx) * (1.f - fy), w10 = fx * (1.f - fy), w01 = (1.f - fx) * fy, w11 = fx * fy;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int cc = c < C ? c : C - 1;
    float s = w00 * __ldg(net.texture + ((size_t)y0 * W + x0) * C + cc);
    if (x1 < W) s += w10 * __ldg(net.texture + ((size_t)y0 * W + x1) * C + cc);
    if (y1 < H) s += w01 * __ldg(net.texture + ((size_t)y1 * W + x0) * C + cc);
    if (x1 < W && y1 < H) s += w11 * __ldg(net.texture + ((size_t)y1 * W + x1) * C + cc);
    out[c] = s;
  }
}

// CG = 1: one CTA per 256-sample tile, every CTA streams the whole weight stream (cta_group::1 MMAs, M = 128).
// CG = 2: the two CTAs of a cluster take 256 samples each and run every MMA together (cta_group::2, M = 256): each
//         CTA streams only the weights of half of every layer's output rows — the pair's tensor cores read both halves —
//         so the per-SM weight stream and its shared-memory operand traffic are halved and the 64 KiB ring reaches a
//         whole 256x256 layer ahead.  Rank 0 issues the MMAs; its commits are multicast to the barriers of both CTAs;
//         rank 1 relays "my ring stage has landed" and "my A operand is written" to rank 0's barriers.
template <int CG>
__global__ void __launch_bounds__(kThreads, 1) ntx_mlp_kernel(const __grid_constant__ NetDev net,
                                                              const __grid_constant__ RenderArgsN a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + offBar);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t count = *reinterpret_cast<volatile const unsigned int*>(a.counters);
  // a unit = the CG tiles one CTA (pair) works on together; both CTAs of a pair see the same unit sequence
  const uint32_t n_tiles = (count + kRows * CG - 1) / (kRows * CG);
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
  const uint32_t unit0 = CG == 2 ? blockIdx.x >> 1 : blockIdx.x, unit_step = CG == 2 ? gridDim.x >> 1 : gridDim.x;
  if (unit0 >= n_tiles) return;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
      mbar_init(&bars->peer_full[s], 1);
    }
    mbar_init(&bars->acc_ready, 1);
    mbar_init(&bars->a_ready, kWorkerWarps * CG);
    fence_mbar_init();
  }
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == kWorkerWarps) {
    if (CG == 2) tmem_alloc_cg2(&bars->tmem_base, 512); else tmem_alloc(&bars->tmem_base, 512);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  const uint32_t nst = net.stream_bytes / kStageBytes;       // ring stages per unit
  if (warp == kWorkerWarps + 1) {
    // ------------------------------------------------------------------ weight producer
    // this rank's stream is a run of whole ring stages; it is copied stage after stage, unit after unit
    if (lane == 0) {
      const uint8_t* wsrc = net.wstream + (size_t)crank * net.stream_bytes;
      uint32_t it = 0;
      const uint32_t q0 = net.layer[kTraceLayer].off / kStageBytes;
      for (uint32_t unit = unit0; unit < n_tiles; unit += unit_step) {
        for (uint32_t q = 0; q < nst; ++q, ++it) {
          const uint32_t s = it % kStages;
          mbar_wait(&bars->empty[s], ((it / kStages) & 1u) ^ 1u);
          if ((net.dbg & 4) && blockIdx.x == 0 && unit == unit0 && q - q0 < 12u) net.trace[100 + 32 + (q - q0)] = clock64();
          if (net.dbg & 8) { mbar_arrive(&bars->full[s]); continue; }
          mbar_expect_tx(&bars->full[s], kStageBytes);
          bulk_g2s(smem + offRing + s * kStageBytes, wsrc + (size_t)q * kStageBytes, kStageBytes, &bars->full[s]);
        }
      }
    }
  } else if (warp == kWorkerWarps && CG == 2 && crank != 0) {
    // ------------------------------------------------------------------ rank 1: relay "stage landed" to rank 0
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t unit = unit0; unit < n_tiles; unit += unit_step) {
        for (uint32_t q = 0; q < nst; ++q, ++it) {
          const uint32_t s = it % kStages;
          mbar_wait(&bars->full[s], (it / kStages) & 1u);
          mbar_arrive_cluster(map_to_cta(smem_u32(&bars->peer_full[s]), 0));
        }
      }
    }
  } else if (warp == kWorkerWarps) {
    // ------------------------------------------------------------------ MMA issuer (rank 0 of a pair)
    // The warp walks the layers together and one lane, chosen with elect.sync, issues a whole layer.  (With the loop
    // under `if (lane == 0)` ptxas wrapped every tcgen05.mma in a replay loop over possibly divergent operands and the
    // per-slice address arithmetic sat between the MMAs: ~226 cycles per M128 N256 K16 MMA against the pipe's 128,
    // scripts/micro/umma_issue.cu.  Under elect.sync it knows a single lane is active.)
    const uint32_t n_units = __shfl_sync(0xffffffffu, n_tiles, 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t a0 = smem_u32(smem + offA), a2 = smem_u32(smem + offA2), r0 = smem_u32(smem + offRing);
    constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);              // SBO = 128 B, descriptor version 1
    constexpr uint32_t kALo = (kKGroupBytes >> 4) << 16;                // LBO of the A operand
    const bool skip = (net.dbg & 2) != 0;
    // The tensor pipe queues only about two MMAs behind the one it is executing, so everything the issuing lane does
    // between the last MMA of one ring stage and the first of the next (commit, barrier wait, descriptor set-up) has
    // to fit in ~256 cycles or the pipe drains and pays its ~300-cycle start-up again: no divisions, no warp-wide
    // hand-offs inside a layer; where a layer's first / last stage is shared with its neighbours is precomputed on the
    // host (LayerDesc::first_have, tail_release).
    uint32_t par_a = 0, ui = 0;
    for (uint32_t unit = unit0; unit < n_units; unit += unit_step, ++ui) {
#pragma unroll 1
      for (int l = 0; l < kNumLayers; ++l) {
        const LayerDesc& L = net.layer[l];
        const uint32_t idesc = umma_idesc(128 * CG, L.N);
        const uint32_t rowsB = (uint32_t)L.N / CG;
        const uint32_t bLo = ((rowsB * 16u) >> 4) << 16;                // LBO of the B operand
        const uint32_t bLoOff = (rowsB * 32u) >> 4;                     // lo part of a split slice, in descriptor units
        const uint32_t slice = L.slice;
        const int nk_main = L.K >> 4, nk_ext = L.Kext >> 4, nk = nk_main + nk_ext + L.bias_slice;
        const bool split = L.split != 0;
        if (CG == 2) mbar_wait_cluster(&bars->a_ready, par_a); else mbar_wait(&bars->a_ready, par_a);
        par_a ^= 1u;
        tc_fence_after();
        const bool tr = (net.dbg & 4) && blockIdx.x == 0 && unit == unit0;
        if (elect_one()) {
          if (tr) net.trace[l * 4 + 0] = clock64();
          // ring state of the layer's first slice
          const uint32_t g0 = ui * nst + L.off / kStageBytes;
          uint32_t s = g0 % kStages, par = (g0 / kStages) & 1u;
          uint32_t rem = kStageBytes - (L.off & (kStageBytes - 1u));      // bytes of the stage still unread
          auto wait_stage = [&](uint32_t st, uint32_t ph) {
            mbar_wait(&bars->full[st], ph);
            if (CG == 2) mbar_wait_cluster(&bars->peer_full[st], ph);
            tc_fence_after();
          };
          if (!L.first_have) wait_stage(s, par);                 // else the previous layer left this stage acquired
          uint32_t b_lo = (((r0 + s * kStageBytes + (L.off & (kStageBytes - 1u))) & 0x3FFFFu) >> 4) | bLo;
          const uint32_t b_step = slice >> 4;
          const bool tail_release = L.tail_release != 0;
          uint32_t acc = 0;                                      // 0 only for the layer's first K step
          int left = nk;                                         // slices of the layer still to issue
          // `n` consecutive K=16 slices whose A operand starts at abase and advances by two K groups per slice
          auto run = [&](uint32_t abase, int n, bool lo_terms) {
            uint32_t a_lo = ((abase & 0x3FFFFu) >> 4) | kALo;
            // fast path: whole ring stages of a 256-wide layer as straight-line code (2 * CG slices, 4 * CG MMAs),
            // so that the descriptors stay in uniform registers from one MMA to the next
            constexpr int SPS = (int)(kStageBytes / (kSliceBytes / CG));
            const bool fast = !lo_terms && slice == kSliceBytes / CG && !skip;
#pragma unroll 1
            for (; n > 0; --n, --left) {
              while (fast && n >= SPS && rem == kStageBytes) {
                const uint32_t s_next = (s + 1u) % kStages, par_next = par ^ (s == kStages - 1u ? 1u : 0u);
#pragma unroll
                for (int j = 0; j < SPS; ++j) {
                  if (j == SPS - 1 && left > SPS) wait_stage(s_next, par_next);   // one slice early, under queued MMAs
                  const uint64_t db = ((uint64_t)kDescHi << 32) | (b_lo + (uint32_t)j * ((kSliceBytes / CG) >> 4));
                  const uint64_t da0 = ((uint64_t)kDescHi << 32) | (a_lo + (uint32_t)j * ((2u * kKGroupBytes) >> 4));
                  const uint64_t da1 = da0 + ((128u * 16u) >> 4);
                  if (CG == 2) {
                    umma_f16_cg2(tm, da0, db, idesc, j == 0 ? acc : 1u);
                    umma_f16_cg2(tm + 256u, da1, db, idesc, j == 0 ? acc : 1u);
                  } else {
                    umma_f16(tm, da0, db, idesc, j == 0 ? acc : 1u);
                    umma_f16(tm + 256u, da1, db, idesc, j == 0 ? acc : 1u);
                  }
                }
                if (CG == 2) umma_commit_cg2(&bars->empty[s]); else umma_commit(&bars->empty[s]);
                acc = 1u;
                a_lo += (uint32_t)SPS * ((2u * kKGroupBytes) >> 4);
                n -= SPS;
                left -= SPS;
                s = s_next;
                par = par_next;
                b_lo = (((r0 + s * kStageBytes) & 0x3FFFFu) >> 4) | bLo;
              }
              if (n == 0) break;
              // generic path, one slice: stages shared with other layers, split layers, view-direction and bias slices.
              // The wait for the next stage goes one slice early, under the MMAs that are still queued
              if (rem == slice && left > 1) wait_stage((s + 1u) % kStages, par ^ (s == kStages - 1u ? 1u : 0u));
              if (!skip) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                  const uint32_t d = tm + (uint32_t)half * 256u;
                  const uint64_t da_hi = ((uint64_t)kDescHi << 32) | (a_lo + (uint32_t)half * ((128u * 16u) >> 4));
                  const uint64_t db_hi = ((uint64_t)kDescHi << 32) | b_lo;
                  if (CG == 2) umma_f16_cg2(d, da_hi, db_hi, idesc, acc); else umma_f16(d, da_hi, db_hi, idesc, acc);
                  if (lo_terms) {
                    const uint64_t da_lo = da_hi + (kALoOff >> 4);
                    const uint64_t db_lo = db_hi + bLoOff;
                    if (CG == 2) {
                      umma_f16_cg2(d, da_lo, db_hi, idesc, 1u);
                      umma_f16_cg2(d, da_hi, db_lo, idesc, 1u);
                    } else {
                      umma_f16(d, da_lo, db_hi, idesc, 1u);
                      umma_f16(d, da_hi, db_lo, idesc, 1u);
                    }
                  }
                }
              }
              acc = 1u;
              a_lo += (2u * kKGroupBytes) >> 4;
              b_lo += b_step;
              rem -= slice;
              // hand the stage back once its MMAs have completed: it is used up, or the rest of it is padding
              if (rem == 0u || (left == 1 && tail_release)) {
                if (CG == 2) umma_commit_cg2(&bars->empty[s]); else umma_commit(&bars->empty[s]);
              }
              if (rem == 0u) {
                s = (s + 1u) % kStages;
                par ^= (s == 0u) ? 1u : 0u;
                rem = kStageBytes;
                b_lo = (((r0 + s * kStageBytes) & 0x3FFFFu) >> 4) | bLo;
              }
            }
          };
          run(a0, nk_main, split);                               // the layer's own inputs
          run(a2, nk_ext, split);                                // view-direction inputs (block2 layer 0)
          run(a2 + 4u * kKGroupBytes, L.bias_slice, false);      // constant-one columns x (bias_hi, bias_lo)
          if (CG == 2) umma_commit_cg2(&bars->acc_ready); else umma_commit(&bars->acc_ready);
          if (tr) net.trace[l * 4 + 1] = clock64();
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ row workers (warps 0..15)
    const int half = (warp >> 2) & 1, ch = warp >> 3;
    const int row = half * 128 + (warp & 3) * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)half * 256u;
    uint8_t* A = smem + offA;
    float* hx = reinterpret_cast<float*>(smem + offHx);        // [2 column halves][256 rows][3]
    float* hw = reinterpret_cast<float*>(smem + offHw);
    const float* heads = net.heads;
    // copy the next head layer's weight rows (n floats, a multiple of 4) into shared memory; called before wait_acc of
    // that layer, whose barrier orders it against the readers; the previous head's readers are several barriers back
    auto stage_head = [&](int off, int n, int bias_off, int n_bias) {
      const int t = warp * 32 + lane, i = t * 4;
      if (i < n) *reinterpret_cast<float4*>(hw + i) = __ldg(reinterpret_cast<const float4*>(heads + off + i));
      else if (t - 256 >= 0 && t - 256 < n_bias) hw[kHwBias + t - 256] = __ldg(heads + bias_off + t - 256);
    };
    uint32_t par_acc = 0;
    int trace_l = 0;
    bool trace_on = false;
    const uint32_t a_ready_rank0 = CG == 2 ? map_to_cta(smem_u32(&bars->a_ready), 0) : 0u;
    auto signal_a = [&]() {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(a_ready_rank0);     // both CTAs' 16 worker warps arrive on rank 0's barrier
        else mbar_arrive(&bars->a_ready);
      }
    };
    auto mark_done = [&]() { if (trace_on && trace_l < kNumLayers) { net.trace[trace_l * 4 + 3] = clock64(); ++trace_l; } };
    // One warp watches the mbarrier, the other fifteen sleep on a hardware barrier (which also orders stage_head's
    // shared-memory writes against the head epilogue's reads)
    auto wait_acc = [&]() {
      if (warp == 0) mbar_wait(&bars->acc_ready, par_acc);
      asm volatile("bar.sync 2, %0;" ::"n"(kWorkerThreads) : "memory");
      par_acc ^= 1u;
      tc_fence_after();
      if (trace_on && trace_l < kNumLayers) net.trace[trace_l * 4 + 2] = clock64();
    };
    // add the partial head sums of the row's other column half
    auto combine = [&](float* v, int n) {
      float* mine = hx + (ch * kRows + row) * 3;
      mine[0] = v[0];
      if (n > 1) mine[1] = v[1];
      if (n > 2) mine[2] = v[2];
      worker_bar();
      const float* other = hx + ((ch ^ 1) * kRows + row) * 3;
      v[0] += other[0];
      if (n > 1) v[1] += other[1];
      if (n > 2) v[2] += other[2];
      worker_bar();              // the buffer may be rewritten by the next head layer only after every read
    };
    {   // the constant-one columns must exist before the first bias slice is multiplied
      const float zero[3] = {0.f, 0.f, 0.f};
      write_encoding<3, 6, 6, false, false, true>(smem + offA2, row, ch, zero);
    }
    for (uint32_t tile = unit0; tile < n_tiles; tile += unit_step) {
      trace_on = (net.dbg & 4) && blockIdx.x == 0 && tile == unit0 && tid == 0;
      trace_l = 0;
      const uint32_t item = (tile * CG + crank) * kRows + row;
      int id = -1;
      float p[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
      if (item < count) {
        const float4 e = __ldg(a.work + item);
        id = __float_as_int(e.x);
        p[0] = e.y; p[1] = e.z; p[2] = e.w;
        const float* rd = a.raydir + (size_t)(id >> 6) * 3;
        dir[0] = __ldg(rd); dir[1] = __ldg(rd + 1); dir[2] = __ldg(rd + 2);
      }
      // ---- geometry: [p, PE(p,10)] -> 256 -> 10 x 256 (ReLU) -> 1 -> softplus          (decoder.py:201-237)
      write_encoding<3, 10, 8, false, false, false>(A, row, ch, p);
      signal_a();
      for (int l = 0; l < 10; ++l) {
        wait_acc();
        epilogue<256, 0, 0, 0>(taddr, A, row, ch, nullptr, nullptr);
        mark_done();
        signal_a();
      }
      float raw = 0.f;
      stage_head(kHeadGeo, 256, kHeadGeoB, 1);
      wait_acc();
      epilogue_head<256, 0, false, 1>(hw, taddr, A, row, ch, &raw);
      mark_done();
      combine(&raw, 1);
      const float sigma = softplus_t(raw + hw[kHwBias]);
      // ---- gauge: [p, PE(p,10)] -> 64 -> 128 -> 128 -> 128 (ReLU) -> 2 -> tanh, split fp16   (gauge_fields.py:8-74)
      write_gauge_encoding(A, row, ch, p);
      signal_a();
      wait_acc();
      epilogue<64, 0, 1, 0>(taddr, A, row, ch, nullptr, nullptr);
      mark_done();
      signal_a();
      for (int r = 0; r < 2; ++r) {
        wait_acc();
        epilogue<128, 0, 1, 0>(taddr, A, row, ch, nullptr, nullptr);
        mark_done();
        signal_a();
      }
      float uvr[3] = {0.f, 0.f, 0.f}, uv[3];
      if (net.sphere) {
        // sphere primitive: 3 outputs, uv = F.normalize(out) = out / max(|out|, 1e-12)   (gauge_fields.py:71-74)
        stage_head(kHeadGauge, 384, kHeadGaugeB, 3);
        wait_acc();
        epilogue_head<128, 0, false, 3>(hw, taddr, A, row, ch, uvr);
        mark_done();
        combine(uvr, 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) uvr[k] += hw[kHwBias + k];
        const float nrm = fmaxf(sqrtf(uvr[0] * uvr[0] + uvr[1] * uvr[1] + uvr[2] * uvr[2]), 1e-12f);
#pragma unroll
        for (int k = 0; k < 3; ++k) uv[k] = uvr[k] / nrm;
        // ---- texture block1: [uv3, PE(uv3,10)] (63) -> 256 -> ...   (model.py:22: uv_dim 3)
        write_encoding<3, 10, 8, false, false, false>(A, row, ch, uv);
      } else {
        stage_head(kHeadGauge, 256, kHeadGaugeB, 2);
        wait_acc();
        epilogue_head<128, 0, false, 2>(hw, taddr, A, row, ch, uvr);
        mark_done();
        combine(uvr, 2);
        uv[0] = tanhf(uvr[0] + hw[kHwBias]); uv[1] = tanhf(uvr[1] + hw[kHwBias + 1]); uv[2] = 0.f;
        // ---- texture block1: [uv, PE(uv,10)] -> 256 -> 5 x 256 (LeakyReLU 0.2); color1 256 -> 3 softplus   (decoder.py:56-78)
        write_encoding<2, 10, 6, false, false, false>(A, row, ch, uv);
      }
      write_encoding<3, 6, 6, false, false, true>(smem + offA2, row, ch, dir);
      signal_a();
      for (int r = 0; r < 5; ++r) {
        wait_acc();
        epilogue<256, 1, 0, 0>(taddr, A, row, ch, nullptr, nullptr);
        mark_done();
        signal_a();
      }
      float c1[3] = {0.f, 0.f, 0.f};
      stage_head(kHeadC1, 768, kHeadC1B, 3);
      wait_acc();
      epilogue_head<256, 1, true, 3>(hw, taddr, A, row, ch, c1);
      mark_done();
      signal_a();
      combine(c1, 3);
#pragma unroll
      for (int k = 0; k < 3; ++k) c1[k] = softplus_t(c1[k] + hw[kHwBias + k]);
      // ---- texture block2: [h, d, PE(d,6)] -> 256 -> 3 x 256 (LeakyReLU) -> 3
      for (int r = 0; r < 3; ++r) {
        wait_acc();
        epilogue<256, 1, 0, 0>(taddr, A, row, ch, nullptr, nullptr);
        mark_done();
        signal_a();
      }
      float c2[3] = {0.f, 0.f, 0.f};
      stage_head(kHeadB2, 768, kHeadB2B, 3);
      wait_acc();
      epilogue_head<256, 1, false, 3>(hw, taddr, A, row, ch, c2);
      mark_done();
      combine(c2, 3);
      if (ch == 0 && id >= 0) {
        float rgb[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) rgb[k] = c1[k] + c2[k] + hw[kHwBias + k];
        if (net.texture == nullptr) {
#pragma unroll
          for (int k = 0; k < 3; ++k) rgb[k] = fmaxf(rgb[k], 0.f);                     // (c1 + c2).clamp(min=0)
        } else {                                                                       // decoder.py:91-103, mode 0
          float m = 0.f;
#pragma unroll
          for (int k = 0; k < 3; ++k) m += fminf(fmaxf(rgb[k] * 8.f, 0.f), 1.f);
          m = m / 3.f;
          float tx[3];
          sample_texture(net, uv[0], uv[1], tx);
#pragma unroll
          for (int k = 0; k < 3; ++k) rgb[k] = tx[k] * m;
        }
        a.sample_out[id] = make_float4(sigma, rgb[0], rgb[1], rgb[2]);
      }
    }
  }
  __syncwarp();            // lane 0 of the producer / issuer / relay warps re-joins its warp
  tc_fence_before();
  if (CG == 2) {
    // every worker of both CTAs has seen the last acc_ready, i.e. every MMA and every multicast commit of rank 0 is
    // done; neither CTA may release its shared memory or TMEM before the other has got this far
    cluster_sync_all();
    if (warp == kWorkerWarps) tmem_dealloc_cg2(tmem, 512);
  } else {
    __syncthreads();
    if (warp == kWorkerWarps) tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------
// cube_ray_generation (model/renderer.py:79-141)
// ---------------------------------------------------------------------------------------------------------
struct RaySetup {
  float c[3], d[3], t;
};

__device__ __forceinline__ RaySetup ray_setup(const RenderArgsN& a, long long ray) {
  RaySetup r;
  float tmin = -INFINITY, tmax = INFINITY;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r.c[k] = __ldg(a.campos + k);
    r.d[k] = __ldg(a.raydir + ray * 3 + k);
    const float t1 = __fdiv_rn(__fsub_rn(-1.f, r.c[k]), r.d[k]);
    const float t2 = __fdiv_rn(__fsub_rn(1.f, r.c[k]), r.d[k]);
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
  }
  r.t = fmaxf(tmin < tmax ? tmin : 0.f, 0.f);
  return r;
}

// Counter-based jitter numbers (Philox 4x32, 10 rounds; Salmon et al. 2011): word (idx & 3) of block idx >> 2 under the
// 64-bit key `seed`, top 24 bits -> U[0,1) like torch.rand's fp32 grid.  Stateless, so the ray generator and the march draw
// the same number twice instead of storing it (256 B per ray).
__device__ __forceinline__ float jitter_uniform(unsigned long long seed, unsigned long long idx) {
  unsigned int c0 = (unsigned int)(idx >> 2), c1 = (unsigned int)(idx >> 34), c2 = 0u, c3 = 0u;
  unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const unsigned int h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const unsigned int sel = (unsigned int)idx & 3u;
  const unsigned int x = sel == 0 ? c0 : sel == 1 ? c1 : sel == 2 ? c2 : c3;
  return (float)(x >> 8) * (1.f / 16777216.f);
}
__device__ __forceinline__ float jitter_at(const RenderArgsN& a, long long ray, int S, int i) {
  if (a.noise) return __ldg(a.noise + ray * S + i);
  return a.seeded ? jitter_uniform(a.seed, (unsigned long long)(a.ray0 + ray) * 64ull + (unsigned)i) : 0.5f;
}

__global__ void ntx_noise_kernel(unsigned long long seed, long long ray0, long long n, int S, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n * S) return;
  const long long ray = i / S;
  out[i] = jitter_uniform(seed, (unsigned long long)(ray0 + ray) * 64ull + (unsigned)(i - ray * S));
}

cudaError_t launch_neutex_noise(unsigned long long seed, long long ray0, long long n_rays, int S, float* noise_out, cudaStream_t st) {
  if (n_rays <= 0) return cudaSuccess;
  const long long n = n_rays * S;
  ntx_noise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seed, ray0, n_rays, S, noise_out);
  return cudaGetLastError();
}

// seg_i = dt + dt*jitter*(U_i - 0.5) (renderer.py:111-118); dt = 2/S and dt*jitter are Python doubles cast to fp32
__device__ __forceinline__ float seg_len(float dt, float dj, float u) { return __fadd_rn(dt, __fmul_rn(dj, __fsub_rn(u, 0.5f))); }

template <bool EMIT>
__device__ __forceinline__ unsigned long long walk_ray(const RenderArgsN& a, const RaySetup& r, long long ray, int S, float dt,
                                                       float dj, float4* dst) {
  unsigned long long mask = 0ull;
  double run = 0.0;
  float e_prev = __fadd_rn(r.t, 0.f);
  int n = 0;
#pragma unroll 4
  for (int i = 0; i < S; ++i) {
    const float u = jitter_at(a, ray, S, i);
    run += (double)seg_len(dt, dj, u);
    const float e = __fadd_rn(r.t, (float)run);
    const float mid = __fmul_rn(__fadd_rn(e_prev, e), 0.5f);
    e_prev = e;
    float p[3];
    bool in = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      p[k] = __fadd_rn(r.c[k], __fmul_rn(r.d[k], mid));
      in = in && (p[k] > -1.f) && (p[k] < 1.f);
    }
    if (in) {
      mask |= 1ull << i;
      if (EMIT) dst[n++] = make_float4(__int_as_float((int)(ray * kS + i)), p[0], p[1], p[2]);
    }
  }
  return mask;
}

__global__ void __launch_bounds__(128) ntx_raygen_kernel(const __grid_constant__ NetDev net,
                                                         const __grid_constant__ RenderArgsN a) {
  const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = ray < a.n_rays;
  const float dj = net.dj, dt = net.dt;
  const int S = net.S;
  RaySetup r{};
  unsigned long long mask = 0ull;
  if (live) {
    r = ray_setup(a, ray);
    mask = walk_ray<false>(a, r, ray, S, dt, dj, nullptr);
    a.valid_mask[ray] = mask;
  }
  // warp prefix sum of the per-ray counts, one atomic per warp
  const int cnt = __popcll(mask);
  int incl = cnt;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, s);
    if (lane >= s) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned int base = 0;
  if (lane == 31 && total > 0) base = atomicAdd(a.counters, (unsigned int)total);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (cnt > 0) walk_ray<true>(a, r, ray, S, dt, dj, a.work + base + (incl - cnt));
}

// ---------------------------------------------------------------------------------------------------------
// ray_march (renderer.py:176-247) + background (model.py:48-49) + simple_tone_map (renderer.py:7-8)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ntx_march_kernel(const __grid_constant__ NetDev net,
                                                        const __grid_constant__ RenderArgsN a) {
  const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= a.n_rays) return;
  const float dj = net.dj, dt = net.dt;
  const int S = net.S;
  const unsigned long long mask = a.valid_mask[ray];
  double T = 1.0;                           // torch.cumprod on the CPU accumulates in double
  float col[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < S; ++i) {
    if (!((mask >> i) & 1ull)) continue;    // sigma * 0 -> opacity 0 -> weight 0, transmittance factor 1 + 1e-10 == 1.f
    const float u = jitter_at(a, ray, S, i);
    const float seg = seg_len(dt, dj, u);
    const float4 s = a.sample_out[ray * kS + i];
    const float op = __fsub_rn(1.f, expf(-__fmul_rn(s.x, seg)));
    const float w = __fmul_rn(op, (float)T);
    col[0] += s.y * w; col[1] += s.z * w; col[2] += s.w * w;
    T *= (double)__fadd_rn(__fsub_rn(1.f, op), 1e-10f);
  }
  const float bg_t = (float)T;
  const float gamma = (float)(1.0 / 2.2);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float c = col[k];
    if (a.background) c += __ldg(a.background + k) * bg_t;
    a.color[ray * 3 + k] = fminf(fmaxf(powf(c + 1e-5f, gamma), 0.f), 1.f);
  }
  a.transmittance[ray] = bg_t;
}

// ---------------------------------------------------------------------------------------------------------
cudaError_t launch_neutex_raygen(const NetDev& net, const RenderArgsN& a, cudaStream_t st) {
  if (a.n_rays <= 0) return cudaSuccess;
  ntx_raygen_kernel<<<(unsigned)((a.n_rays + 127) / 128), 128, 0, st>>>(net, a);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_neutex_mlp(const NetDev& net, const RenderArgsN& a, int num_sms, cudaStream_t st) {
  static PerDevice<int> configured;
  bool fresh = false;
  configured.get(&fresh);
  if (fresh) {
    cudaError_t e = cudaFuncSetAttribute(ntx_mlp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(ntx_mlp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) { configured.retry(); return e; }
  }
  if (net.cg == 2) {
    // one CTA pair (a cluster of two, i.e. the two SMs of a TPC) per 512 work items
    long long worst = (a.n_rays * kS + 2 * kRows - 1) / (2 * kRows);
    long long pairs = num_sms / 2 < worst ? num_sms / 2 : worst;
    if (pairs < 1) pairs = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, ntx_mlp_kernel<2>, net, a);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
  }
  long long worst = (a.n_rays * kS + kRows - 1) / kRows;
  long long grid = num_sms < worst ? num_sms : worst;
  if (grid < 1) grid = 1;
  ntx_mlp_kernel<1><<<(unsigned)grid, kThreads, kSmemBytes, st>>>(net, a);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// fp32 reference evaluation (one warp per in-cube sample): the same networks as ntx_mlp_kernel, from the unpacked fp32
// parameters, every layer a plain dot product per output.  ~30x slower than the tensor-core kernel; used by
// ngf_neutex_self_check and as the NGF_NTX_FP32 / ngf_neutex_set_precision fallback for networks whose fp16 evaluation
// leaves the 1e-3 image budget (ill-conditioned trained weights).
// ---------------------------------------------------------------------------------------------------------
constexpr int kRefWarps = 8, kRefWidth = 320;
__device__ __forceinline__ void ref_dense(const RawNet& raw, int l, const float* __restrict__ x, float* __restrict__ y, int act,
                                          int lane) {
  const int in = raw.in[l], out = raw.out[l];
  for (int n = lane; n < out; n += 32) {
    const float* w = raw.w[l] + (size_t)n * in;
    float acc = __ldg(raw.b[l] + n);
    for (int k = 0; k < in; ++k) acc = fmaf(__ldg(w + k), x[k], acc);
    y[n] = act == 1 ? fmaxf(acc, 0.f) : act == 2 ? (acc > 0.f ? acc : 0.2f * acc) : acc;
  }
  __syncwarp();
}
__device__ __forceinline__ void ref_encode(const float* v, int dims, int freqs, float* out, int lane) {
  // [x, sin(x_d 2^f) (d-major), cos(same)]   (util.py:427-438)
  for (int i = lane; i < dims; i += 32) out[i] = v[i];
  for (int i = lane; i < dims * freqs; i += 32) {
    const float a = v[i / freqs] * (float)(1 << (i % freqs));
    out[dims + i] = sinf(a);
    out[dims + dims * freqs + i] = cosf(a);
  }
  __syncwarp();
}
__global__ void __launch_bounds__(kRefWarps * 32) ntx_ref_kernel(const __grid_constant__ NetDev net,
                                                                 const __grid_constant__ RawNet raw,
                                                                 const __grid_constant__ RenderArgsN a) {
  __shared__ float buf[kRefWarps][2][kRefWidth];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned count = a.counters[0];
  float* x = buf[warp][0];
  float* y = buf[warp][1];
  for (unsigned item = blockIdx.x * kRefWarps + warp; item < count; item += gridDim.x * kRefWarps) {
    const float4 e = a.work[item];
    const int id = __float_as_int(e.x);
    const float p[3] = {e.y, e.z, e.w};
    const float* rd = a.raydir + (size_t)(id >> 6) * 3;
    const float dir[3] = {rd[0], rd[1], rd[2]};
    // geometry
    ref_encode(p, 3, 10, x, lane);
    for (int l = 0; l < 11; ++l) { ref_dense(raw, l, x, y, 1, lane); float* t = x; x = y; y = t; }
    ref_dense(raw, 11, x, y, 0, lane);
    const float sigma = softplus_t(y[0]);
    __syncwarp();
    // gauge
    ref_encode(p, 3, 10, x, lane);
    for (int l = 12; l < 16; ++l) { ref_dense(raw, l, x, y, 1, lane); float* t = x; x = y; y = t; }
    ref_dense(raw, 16, x, y, 0, lane);
    float uv[3];
    if (net.sphere) {
      const float nrm = fmaxf(sqrtf(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]), 1e-12f);
      uv[0] = y[0] / nrm; uv[1] = y[1] / nrm; uv[2] = y[2] / nrm;
    } else {
      uv[0] = tanhf(y[0]); uv[1] = tanhf(y[1]); uv[2] = 0.f;
    }
    __syncwarp();
    // texture block1, color1
    ref_encode(uv, net.sphere ? 3 : 2, 10, x, lane);
    for (int l = 17; l < 23; ++l) { ref_dense(raw, l, x, y, 2, lane); float* t = x; x = y; y = t; }
    ref_dense(raw, 23, x, y, 0, lane);                    // color1: x = h (256), y = 3
    const float c1[3] = {softplus_t(y[0]), softplus_t(y[1]), softplus_t(y[2])};
    __syncwarp();
    // block2: [h, d, PE(d, 6)] (295)
    ref_encode(dir, 3, 6, x + 256, lane);
    for (int l = 24; l < 28; ++l) { ref_dense(raw, l, x, y, 2, lane); float* t = x; x = y; y = t; }
    ref_dense(raw, 28, x, y, 0, lane);
    if (lane == 0) {
      float rgb[3] = {c1[0] + y[0], c1[1] + y[1], c1[2] + y[2]};
      if (net.texture == nullptr) {
        for (int k = 0; k < 3; ++k) rgb[k] = fmaxf(rgb[k], 0.f);
      } else {
        float m = 0.f;
        for (int k = 0; k < 3; ++k) m += fminf(fmaxf(rgb[k] * 8.f, 0.f), 1.f);
        m = m / 3.f;
        float tx[3];
        sample_texture(net, uv[0], uv[1], tx);
        for (int k = 0; k < 3; ++k) rgb[k] = tx[k] * m;
      }
      a.sample_out[id] = make_float4(sigma, rgb[0], rgb[1], rgb[2]);
    }
    __syncwarp();
  }
}

cudaError_t launch_neutex_ref(const NetDev& net, const RawNet& raw, const RenderArgsN& a, cudaStream_t st) {
  long long worst = (a.n_rays * kS + kRefWarps - 1) / kRefWarps;
  long long grid = worst < 148 * 8 ? worst : 148 * 8;
  if (grid < 1) grid = 1;
  ntx_ref_kernel<<<(unsigned)grid, kRefWarps * 32, 0, st>>>(net, raw, a);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_neutex_march(const NetDev& net, const RenderArgsN& a, cudaStream_t st) {
  if (a.n_rays <= 0) return cudaSuccess;
  ntx_march_kernel<<<(unsigned)((a.n_rays + 127) / 128), 128, 0, st>>>(net, a);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ntx
}  // namespace ngf
