// ngf_train.cu — backward pass of the render path: the gradients of one training step
// (TriPlane/main.py:272-302: field(rays, is_train=True) -> loss(rgb_map) -> backward) with respect to every parameter
// Base.forward reads — feature planes, gauge planes, rgb_decoder (basis + 3 layers), density head — given dL/d(rgb_map).
//
// Nothing is saved by the forward: the backward re-marches the rays with the same jitter and the same decision chain
// (sample positions, bbox, occupancy, early-out) as ngf_march_kernel, so it differentiates exactly the samples the forward
// composited.  fp32 throughout (CUDA cores): a training batch is 4096 rays (main.py:272), three orders of magnitude less
// work than an evaluation frame, and the gradients are compared with torch autograd at 1e-3.
//
//   ngf_bwd_march_kernel      one ray per lane: re-march, append one record per valid sample (coords, t, delta, sigma, T, w)
//                             to a device list, linked per ray (prev index) for the reverse walk; samples with
//                             w > rayMarch_weight_thres also go to the active list.
//   ngf_bwd_feat_kernel       active samples: appearance features X [A][F] (bilinear from the fp16 planes, x phase code
//                             for InfoInv) and the view-direction columns of the MLP input.
//   ngf_gemm_kernel           strided fp32 GEMM (64x64 tiles) used for the colour MLP forward, its backward and the
//                             weight gradients (split over samples, atomic accumulation).
//   ngf_bwd_colour_out_kernel sigmoid, dL/dz3 = (dL/drgb_map . w) c (1 - c)                  (FieldBase.py:297)
//   ngf_bwd_colour_scatter_kernel  dL/dX -> appearance channels of the plane gradients (grid_sample backward) and the
//                             gradient of the plane coordinates (for the gauge planes).
//   ngf_bwd_composite_kernel  one ray per thread, reverse walk over its records: dL/dw, dL/dalpha, dL/dsigma
//                             (raw2alpha backward, FieldBase.py:12-19; white background term, :299-300).
//   ngf_bwd_density_kernel    per record: softplus', density head (Linear(48,1)) gradient, density channels of the plane
//                             gradients, coordinate gradient, scatter of the total coordinate gradient into the three
//                             gauge planes (compute_gauge backward, Field.py:53-75).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ngf_handle.h"

using namespace ngf;

#define CU(expr)                                                                                                   \
  do {                                                                                                             \
    cudaError_t _e = (expr);                                                                                       \
    if (_e != cudaSuccess)                                                                                         \
      return ngf_set_error(NGF_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);        \
  } while (0)

namespace {

struct __align__(16) BRec {
  float c[6];        // plane coordinates after the gauge: u_xy v_xy u_yz v_yz u_xz v_xz
  float n[3];        // normalised coordinates before the gauge
  float t, delta, sigma, T, w;
  int ray, prev;     // previous record of the same ray (-1: first)
  float rgb[3];      // colour of the sample (active records), 0 otherwise
  float dsigma;      // dL/dsigma
  int active;        // index in the active list, -1 if w <= rayMarch_weight_thres
  float dc[6];       // dL/d(plane coordinates) from the colour branch
  float pad[5];
};
static_assert(sizeof(BRec) == 128, "BRec size");

struct BwdArgs {
  const float* rays;
  const float* jitter;      // nullptr: evaluation-time sampling
  long long n_rays;
  int ray_stride, S;
  BRec* rec;
  unsigned int rec_cap;
  int* tail;                // [R] last record of the ray, -1 if none
  int* active_list;         // [rec_cap] record index of active sample a
  unsigned int* counters;   // [0] records, [1] active, [2] overflow
};

// ---------------------------------------------------------------------------------------------------------------------
// re-march (same chain as ngf_march_kernel in ngf_kernels.cu; one ray per lane, no image tiling)
// ---------------------------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) ngf_bwd_march_kernel(const __grid_constant__ FieldDev f,
                                                            const __grid_constant__ BwdArgs a) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool has = ray < a.n_rays;
  float o[3] = {0, 0, 0}, d[3] = {0, 0, 1}, t0 = 0.f, jit = 0.f, T = 1.f;
  int i = 0, i_end = 0, last = -1;
  const int S = a.S;
  bool live = false;
  if (has) {
    const float* rp = a.rays + ray * a.ray_stride;
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = rp[k]; d[k] = rp[3 + k]; }
    t0 = ray_t0(f, o, d);
    if (a.jitter) jit = a.jitter[ray];
    int lo_i, hi_i;
    ray_index_range(f, o, d, t0, S, lo_i, hi_i, a.jitter ? 1.f : 0.f);
    i = lo_i; i_end = hi_i + 1;
    live = i < i_end;
  }
  const float* dmlp = f.dmlp;
  while (__any_sync(FULL, live)) {
    bool found = false;
    float t = 0.f, p[3] = {0, 0, 0};
    while (live && !found) {
      t = a.jitter ? sample_t(f, t0, i, jit) : sample_t(f, t0, i);
      bool in = sample_pos(f, o, d, t, p);
      if (in && f.has_occ) in = occ_keep(f, p);
      if (in) found = true;
      else if (++i >= i_end) live = false;
    }
    BRec r;
    bool act = false;
    if (found) {
      unit_coords(f, p, r.n);
      gauge_coords(f, r.n, V == 0 && f.gauge_on, r.c);
      r.sigma = (V == 0) ? sigma_triplane(f, r.c) : sigma_infoinv(f, r.c, dmlp);
      const float tn = a.jitter ? sample_t(f, t0, i + 1, jit) : sample_t(f, t0, i + 1);
      r.delta = (i == S - 1) ? 0.f : __fmul_rn(__fsub_rn(tn, t), f.dscale);
      const float alpha = __fsub_rn(1.f, expf(-__fmul_rn(r.sigma, r.delta)));
      r.t = t; r.T = T;
      r.w = __fmul_rn(alpha, T);
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f));
      act = r.w > f.wthres;
      if (++i >= i_end || T <= f.tstop) live = false;
    }
    const unsigned fm = __ballot_sync(FULL, found), am = __ballot_sync(FULL, act);
    if (fm) {
      unsigned base = 0, abase = 0;
      if (lane == 0) {
        base = atomicAdd(a.counters, (unsigned)__popc(fm));
        if (am) abase = atomicAdd(a.counters + 1, (unsigned)__popc(am));
      }
      base = __shfl_sync(FULL, base, 0);
      abase = __shfl_sync(FULL, abase, 0);
      if (found) {
        const unsigned slot = base + __popc(fm & ((1u << lane) - 1u));
        if (slot < a.rec_cap) {
          r.ray = (int)ray; r.prev = last;
          r.rgb[0] = r.rgb[1] = r.rgb[2] = 0.f;
          r.dsigma = 0.f;
          r.active = -1;
#pragma unroll
          for (int k = 0; k < 6; ++k) r.dc[k] = 0.f;
          if (act) {
            const unsigned ai = abase + __popc(am & ((1u << lane) - 1u));
            r.active = (int)ai;
            a.active_list[ai] = (int)slot;
          }
          a.rec[slot] = r;
          last = (int)slot;
        } else {
          atomicExch(a.counters + 2, 1u);
          live = false;
        }
      }
    }
  }
  if (has) a.tail[ray] = last;
}

// ---------------------------------------------------------------------------------------------------------------------
// strided fp32 GEMM: C[m][n] = epi( sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n]) ), 64x64 tiles, 256 threads,
// 4x4 outputs per thread.  split > 1: the k range is divided over blockIdx.z and partial sums are atomically added.
//   epi 0: none; 1: relu; 2: multiply by (mask[m*ldc + n] > 0)
// ---------------------------------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A; long long sam, sak;
  const float* B; long long sbk, sbn;
  const float* bias;
  const float* mask;
  float* C; long long ldc;
  long long M; int N; long long K;
  int epi, atomic;
};

__global__ void __launch_bounds__(256) ngf_gemm_kernel(const __grid_constant__ GemmArgs g) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const long long kper = (g.K + gridDim.z - 1) / gridDim.z;
  const long long kb = (long long)blockIdx.z * kper;
  long long ke = kb + kper;
  if (ke > g.K) ke = g.K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long k0 = kb; k0 < ke; k0 += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + 256 * e;                 // 1024 elements per tile
      {
        // A tile [64 m][16 k]: consecutive threads walk the dimension with the smaller stride
        int mm, kk;
        if (g.sak <= g.sam) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
        const long long m = m0 + mm, k = k0 + kk;
        As[kk][mm] = (m < g.M && k < ke) ? g.A[m * g.sam + k * g.sak] : 0.f;
      }
      {
        int nn, kk;
        if (g.sbk <= g.sbn) { kk = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kk = idx >> 6; }
        const int n = n0 + nn;
        const long long k = k0 + kk;
        Bs[kk][nn] = (n < g.N && k < ke) ? g.B[k * g.sbk + (long long)n * g.sbn] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias && blockIdx.z == 0) v += g.bias[n];
      float* dst = g.C + m * g.ldc + n;
      if (g.atomic) { atomicAdd(dst, v); continue; }
      if (g.epi == 1) v = fmaxf(v, 0.f);
      else if (g.epi == 2) v = g.mask[m * g.ldc + n] > 0.f ? v : 0.f;
      *dst = v;
    }
  }
}

int gemm(cudaStream_t st, const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
         const float* bias, const float* mask, float* C, long long ldc, long long M, int N, long long K, int epi,
         int split = 1) {
  if (M <= 0 || N <= 0 || K <= 0) return NGF_OK;
  GemmArgs g{A, sam, sak, B, sbk, sbn, bias, mask, C, ldc, M, N, K, epi, split > 1 ? 1 : 0};
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64), (unsigned)split);
  ngf_gemm_kernel<<<grid, 256, 0, st>>>(g);
  count_launch();
  CU(cudaGetLastError());
  return NGF_OK;
}

// column sums of D [A][N] (row stride ld) added to out[N]
__global__ void ngf_colsum_kernel(const float* __restrict__ D, long long A, int N, long long ld, float* __restrict__ out) {
  const int n = blockIdx.y * 32 + (threadIdx.x & 31);
  float s = 0.f;
  if (n < N)
    for (long long a = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); a < A; a += (long long)gridDim.x * 8) s += D[a * ld + n];
  if (n < N && s != 0.f) atomicAdd(out + n, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// colour branch
// ---------------------------------------------------------------------------------------------------------------------
struct PlaneParams {
  const float* plane[3];   // the fp32 feature-plane parameters [C][H][W], or NULL: use the handle's fp16 appearance texels
};

// X [A][F]: appearance features of the active samples; IN[:, F..F+15): view-direction columns [d, sin(d_k 2^j), cos(..)]
template <int V>
__global__ void ngf_bwd_feat_kernel(const __grid_constant__ FieldDev f, const BRec* __restrict__ rec,
                                    const int* __restrict__ active_list, long long n_active,
                                    const float* __restrict__ rays, int ray_stride, float* __restrict__ X,
                                    float* __restrict__ IN, const __grid_constant__ PlaneParams pp) {
  constexpr int AC = Cfg<V>::AC, DC = Cfg<V>::DC, F = Cfg<V>::F, CH = AC / 8, LD = F + 15;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long a = idx / (3 * CH);
  if (a >= n_active) return;
  const int rem = (int)(idx - a * (3 * CH)), pl = rem / CH, chunk = rem - pl * CH;
  const BRec& r = rec[active_list[a]];
  const PlaneDev& P = f.plane[pl];
  const Taps t = make_taps(r.c[2 * pl], r.c[2 * pl + 1], P.W, P.H, P.wm1, P.hm1);
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (pp.plane[pl]) {                                    // the fp32 parameter itself, [C][H][W]: the reference's features
    const size_t hw = (size_t)P.H * P.W;
    const float* src = pp.plane[pl] + (size_t)(DC + chunk * 8) * hw;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += t.w[k] * __ldg(src + (size_t)e * hw + t.off[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(P.app + (size_t)t.off[k] * AC + chunk * 8));
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = __half22float2(h[e]);
        v[2 * e] += t.w[k] * x.x;
        v[2 * e + 1] += t.w[k] * x.y;
      }
    }
  }
  if (V == 1 && f.infoinv) {
    const float xyz[3] = {r.c[0], r.c[1], r.c[3]};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= phase_value<12>(xyz, chunk * 8 + e);
  }
  float* dst = X + a * F + pl * AC + chunk * 8;
#pragma unroll
  for (int e = 0; e < 8; ++e) dst[e] = v[e];
  if (rem == 0) {
    const float* dp = rays + (size_t)r.ray * ray_stride + 3;
    float* o = IN + a * LD + F;
    const float d[3] = {dp[0], dp[1], dp[2]};
    o[0] = d[0]; o[1] = d[1]; o[2] = d[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {                        // positional_encoding(d, 2): [sin(d_k*1), sin(d_k*2)] k-major, then cos
      o[3 + 2 * k] = sinf(d[k]); o[4 + 2 * k] = sinf(d[k] * 2.f);
      o[9 + 2 * k] = cosf(d[k]); o[10 + 2 * k] = cosf(d[k] * 2.f);
    }
  }
}

// Z [A][3] -> rgb = sigmoid(Z) into the record; dZ = (g_ray * w) * rgb * (1 - rgb) in place
__global__ void ngf_bwd_colour_out_kernel(BRec* __restrict__ rec, const int* __restrict__ active_list, long long n_active,
                                          const float* __restrict__ grad_rgb, float* __restrict__ Z) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_active) return;
  BRec& r = rec[active_list[a]];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s = 1.f / (1.f + expf(-Z[a * 3 + c]));
    r.rgb[c] = s;
    Z[a * 3 + c] = grad_rgb[(size_t)r.ray * 3 + c] * r.w * s * (1.f - s);
  }
}

// Tap set of one bilinear lookup with what grid_sampler_2d's backward needs (ATen GridSampler, align_corners=True, zeros):
// in-bounds flags and the derivatives of the four weights with respect to the unnormalised coordinates (ix, iy).  A tap
// that is in bounds contributes to the coordinate gradient even when its weight is exactly 0.
struct TapsG {
  int off[4];
  float w[4], dwx[4], dwy[4];
  bool ok[4];
};
__device__ __forceinline__ TapsG make_taps_grad(float u, float v, int W, int H, float wm1, float hm1) {
  float ix = ((u + 1.f) * 0.5f) * wm1, iy = ((v + 1.f) * 0.5f) * hm1;
  ix = fminf(fmaxf(ix, -2.f), (float)W + 1.f);
  iy = fminf(fmaxf(iy, -2.f), (float)H + 1.f);
  const float x0f = floorf(ix), y0f = floorf(iy), fx = ix - x0f, fy = iy - y0f;
  const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
  const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)x1 < (unsigned)W;
  const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)y1 < (unsigned)H;
  const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1), cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
  TapsG t;
  t.off[0] = cy0 * W + cx0; t.ok[0] = vx0 && vy0; t.w[0] = (1.f - fx) * (1.f - fy); t.dwx[0] = -(1.f - fy); t.dwy[0] = -(1.f - fx);
  t.off[1] = cy0 * W + cx1; t.ok[1] = vx1 && vy0; t.w[1] = fx * (1.f - fy);         t.dwx[1] = (1.f - fy);  t.dwy[1] = -fx;
  t.off[2] = cy1 * W + cx0; t.ok[2] = vx0 && vy1; t.w[2] = (1.f - fx) * fy;         t.dwx[2] = -fy;         t.dwy[2] = (1.f - fx);
  t.off[3] = cy1 * W + cx1; t.ok[3] = vx1 && vy1; t.w[3] = fx * fy;                 t.dwx[3] = fy;          t.dwy[3] = fx;
  return t;
}

struct PlaneGrads {
  float* plane[3];     // [C][H][W] fp32 (the reference's NCHW parameter layout), accumulated into
  float* gauge[3];     // [2][Hg][Wg]
};

// dX [A][F] -> appearance channels of the plane gradients; coordinate gradient of the colour branch -> rec.dc
template <int V>
__global__ void ngf_bwd_colour_scatter_kernel(const __grid_constant__ FieldDev f, BRec* __restrict__ rec,
                                              const int* __restrict__ active_list, long long n_active,
                                              const float* __restrict__ dX, const __grid_constant__ PlaneGrads pg,
                                              int want_coord, const __grid_constant__ PlaneParams pp) {
  constexpr int AC = Cfg<V>::AC, DC = Cfg<V>::DC, F = Cfg<V>::F;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long a = idx / 3;
  if (a >= n_active) return;
  const int pl = (int)(idx - a * 3);
  BRec& r = rec[active_list[a]];
  const PlaneDev& P = f.plane[pl];
  const float u = r.c[2 * pl], v = r.c[2 * pl + 1];
  const TapsG t = make_taps_grad(u, v, P.W, P.H, P.wm1, P.hm1);
  const size_t hw = (size_t)P.H * P.W;
  float gu = 0.f, gv = 0.f;
  const float xyz[3] = {r.c[0], r.c[1], r.c[3]};
  for (int ch = 0; ch < AC; ++ch) {
    float gfeat = dX[a * F + pl * AC + ch];
    if (V == 1 && f.infoinv) gfeat *= phase_value<12>(xyz, ch);
    float* gp = pg.plane[pl] + (size_t)(DC + ch) * hw;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!t.ok[k]) continue;                          // a tap outside the plane contributes nothing
      if (t.w[k] != 0.f) atomicAdd(gp + t.off[k], t.w[k] * gfeat);
      if (want_coord) {
        const float tex = pp.plane[pl] ? __ldg(pp.plane[pl] + (size_t)(DC + ch) * hw + t.off[k])
                                       : __half2float(P.app[(size_t)t.off[k] * AC + ch]);
        gu += gfeat * t.dwx[k] * tex;
        gv += gfeat * t.dwy[k] * tex;
      }
    }
  }
  if (want_coord) {
    r.dc[2 * pl] = gu * 0.5f * P.wm1;
    r.dc[2 * pl + 1] = gv * 0.5f * P.hm1;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// compositing backward (raw2alpha, FieldBase.py:12-19; rgb_map, :296-302): one ray per thread, reverse walk
// ---------------------------------------------------------------------------------------------------------------------
__global__ void ngf_bwd_composite_kernel(BRec* __restrict__ rec, const int* __restrict__ tail, long long n_rays,
                                         const float* __restrict__ grad_rgb, int white_bg) {
  const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= n_rays) return;
  const float g0 = grad_rgb[ray * 3], g1 = grad_rgb[ray * 3 + 1], g2 = grad_rgb[ray * 3 + 2];
  const float gbg = white_bg ? (g0 + g1 + g2) : 0.f;       // rgb_map += 1 - acc_map
  float suffix = 0.f;                                      // sum_{j > i} dL/dw_j * w_j
  for (int s = tail[ray]; s >= 0;) {
    BRec& r = rec[s];
    const float dLdw = g0 * r.rgb[0] + g1 * r.rgb[1] + g2 * r.rgb[2] - gbg;
    const float alpha = __fsub_rn(1.f, expf(-__fmul_rn(r.sigma, r.delta)));
    const float om = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
    const float dalpha = dLdw * r.T - suffix / om;
    suffix += dLdw * r.w;
    r.dsigma = dalpha * r.delta * (1.f - alpha);            // d(1 - exp(-sigma delta))/dsigma = delta exp(-sigma delta)
    s = r.prev;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// density branch + gauge scatter, TriPlane (Field.py:48-50, 53-91)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ngf_bwd_density_kernel(const __grid_constant__ FieldDev f,
                                                              const BRec* __restrict__ rec, long long n_rec,
                                                              const __grid_constant__ PlaneGrads pg,
                                                              float* __restrict__ g_dw, float* __restrict__ g_db) {
  __shared__ float s_dw[49];
  if (threadIdx.x < 49) s_dw[threadIdx.x] = 0.f;
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rec) {
    const BRec& r = rec[i];
    // sigma = softplus(raw - 10): d sigma / d raw = sigmoid(raw - 10) = 1 - exp(-sigma)
    const float draw = r.dsigma * (-expm1f(-r.sigma));
    float dc[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) dc[k] = r.dc[k];
    if (draw != 0.f) {
      atomicAdd(&s_dw[48], draw);
      for (int pl = 0; pl < 3; ++pl) {
        const PlaneDev& P = f.plane[pl];
        const float u = r.c[2 * pl], v = r.c[2 * pl + 1];
        const TapsG t = make_taps_grad(u, v, P.W, P.H, P.wm1, P.hm1);
        const size_t hw = (size_t)P.H * P.W;
        float gu = 0.f, gv = 0.f;
        for (int q = 0; q < 4; ++q) {                       // 4 x float4 = 16 density channels
          float4 tex[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) tex[k] = __ldg(reinterpret_cast<const float4*>(P.dens) + (size_t)t.off[k] * 4 + q);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ch = 4 * q + e;
            const float gfeat = draw * f.dw[16 * pl + ch];
            float feat = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float x = e == 0 ? tex[k].x : e == 1 ? tex[k].y : e == 2 ? tex[k].z : tex[k].w;
              if (!t.ok[k]) continue;
              feat += t.w[k] * x;
              if (t.w[k] != 0.f) atomicAdd(pg.plane[pl] + (size_t)ch * hw + t.off[k], t.w[k] * gfeat);
              gu += gfeat * t.dwx[k] * x;
              gv += gfeat * t.dwy[k] * x;
            }
            atomicAdd(&s_dw[16 * pl + ch], draw * feat);
          }
        }
        dc[2 * pl] += gu * 0.5f * P.wm1;
        dc[2 * pl + 1] += gv * 0.5f * P.hm1;
      }
    }
    if (f.gauge_on) {
      // compute_gauge backward: c0 = (x+gxy.x)+gxz.x, c1 = (y+gxy.y)+gyz.x, c2 = (y+gyz.x)+gxy.y, c3 = (z+gyz.y)+gxz.y,
      // c4 = (x+gxz.x)+gxy.x, c5 = (z+gxz.y)+gyz.y
      const float dg[3][2] = {{dc[0] + dc[4], dc[1] + dc[2]}, {dc[1] + dc[2], dc[3] + dc[5]}, {dc[0] + dc[4], dc[3] + dc[5]}};
      const float uv[3][2] = {{r.n[0], r.n[1]}, {r.n[1], r.n[2]}, {r.n[0], r.n[2]}};
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        if (dg[pl][0] == 0.f && dg[pl][1] == 0.f) continue;
        const GaugeDev& G = f.gauge[pl];
        const Taps t = make_taps(uv[pl][0], uv[pl][1], G.W, G.H, G.wm1, G.hm1);
        const size_t hw = (size_t)G.H * G.W;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (t.w[k] != 0.f) {
            atomicAdd(pg.gauge[pl] + t.off[k], t.w[k] * dg[pl][0]);
            atomicAdd(pg.gauge[pl] + hw + t.off[k], t.w[k] * dg[pl][1]);
          }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 48 && s_dw[threadIdx.x] != 0.f) atomicAdd(g_dw + threadIdx.x, s_dw[threadIdx.x]);
  if (threadIdx.x == 48 && s_dw[48] != 0.f) atomicAdd(g_db, s_dw[48]);
}

// ---------------------------------------------------------------------------------------------------------------------
// density branch, InfoInv (InfoInv/models/Field.py:52-70, networks.py:34-54): features of the records (x phase code)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void ngf_bwd_dfeat_kernel(const __grid_constant__ FieldDev f, const BRec* __restrict__ rec, long long first,
                                     long long n, float* __restrict__ Xd) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long j = idx / 18;                           // 3 planes x 6 float4
  if (j >= n) return;
  const int rem = (int)(idx - j * 18), pl = rem / 6, q = rem - pl * 6;
  const BRec& r = rec[first + j];
  const PlaneDev& P = f.plane[pl];
  const Taps t = make_taps(r.c[2 * pl], r.c[2 * pl + 1], P.W, P.H, P.wm1, P.hm1);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(P.dens) + (size_t)t.off[k] * 6 + q);
    v.x += t.w[k] * x.x; v.y += t.w[k] * x.y; v.z += t.w[k] * x.z; v.w += t.w[k] * x.w;
  }
  if (f.infoinv) {
    const float xyz[3] = {r.c[0], r.c[1], r.c[3]};
    v.x *= phase_value<4>(xyz, 4 * q); v.y *= phase_value<4>(xyz, 4 * q + 1);
    v.z *= phase_value<4>(xyz, 4 * q + 2); v.w *= phase_value<4>(xyz, 4 * q + 3);
  }
  *reinterpret_cast<float4*>(Xd + j * 72 + pl * 24 + 4 * q) = v;
}

// dZ[j] = dsigma * softplus'(raw - 10) for records first .. first+n
__global__ void ngf_bwd_dsig_kernel(const BRec* __restrict__ rec, long long first, long long n, float* __restrict__ dZ) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const BRec& r = rec[first + j];
  dZ[j] = r.dsigma * (-expm1f(-r.sigma));
}

// dXd [n][72] -> density channels of the plane gradients (no coordinate gradient: InfoInv has no gauge planes)
__global__ void ngf_bwd_dscatter_kernel(const __grid_constant__ FieldDev f, const BRec* __restrict__ rec, long long first,
                                        long long n, const float* __restrict__ dXd, const __grid_constant__ PlaneGrads pg) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long j = idx / 3;
  if (j >= n) return;
  const int pl = (int)(idx - j * 3);
  const BRec& r = rec[first + j];
  const PlaneDev& P = f.plane[pl];
  const Taps t = make_taps(r.c[2 * pl], r.c[2 * pl + 1], P.W, P.H, P.wm1, P.hm1);
  const size_t hw = (size_t)P.H * P.W;
  const float xyz[3] = {r.c[0], r.c[1], r.c[3]};
  for (int ch = 0; ch < 24; ++ch) {
    float gfeat = dXd[j * 72 + pl * 24 + ch];
    if (f.infoinv) gfeat *= phase_value<4>(xyz, ch);
    if (gfeat == 0.f) continue;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (t.w[k] != 0.f) atomicAdd(pg.plane[pl] + (size_t)ch * hw + t.off[k], t.w[k] * gfeat);
  }
}

unsigned grid_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// One Adam step over a flat fp32 parameter (torch.optim.Adam, no weight decay, no amsgrad; TriPlane/main.py:237,300-302):
//   m = b1 m + (1 - b1) g ; v = b2 v + (1 - b2) g^2 ; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// in torch's order of operations (lerp for m, addcmul for v, addcdiv for p), one pass over the four arrays.
__global__ void ngf_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long long n, float b1, float b2, float one_minus_b2, float step_size,
                                float bc2_sqrt, float eps) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);               // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * b2 + one_minus_b2 * gi * gi;            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);                          // param.addcdiv_(exp_avg, denom, value=-step_size)
  }
}

}  // namespace

extern "C" int ngf_adam_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n,
                             double lr, double beta1, double beta2, double eps, int64_t step, void* stream) {
  if (!param_dev || !grad_dev || !exp_avg_dev || !exp_avg_sq_dev) return ngf_set_error(NGF_EINVAL, "NULL pointer");
  if (n < 0 || step < 1) return ngf_set_error(NGF_EINVAL, "n=%lld step=%lld", (long long)n, (long long)step);
  if (n == 0) return NGF_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ngf_adam_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      param_dev, grad_dev, exp_avg_dev, exp_avg_sq_dev, n, (float)beta1, (float)beta2, (float)(1.0 - beta2), (float)(lr / bc1),
      (float)sqrt(bc2), (float)eps);
  count_launch();
  CU(cudaGetLastError());
  return NGF_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
struct TrainWs {               // owned by the field handle (NgfField_::train)
  BRec* rec = nullptr;
  int* active_list = nullptr;
  long long rec_cap = 0;
  int* tail = nullptr;
  long long tail_cap = 0;
  unsigned int* counters = nullptr;
  float* buf = nullptr;        // activation workspace
  long long buf_floats = 0;
};

void ngf_train_free(void* p) {
  TrainWs* w = static_cast<TrainWs*>(p);
  if (!w) return;
  cudaFree(w->rec); cudaFree(w->active_list); cudaFree(w->tail); cudaFree(w->counters); cudaFree(w->buf);
  delete w;
}

static long long train_budget_bytes() {
  const char* e = getenv("NGF_TRAIN_MIB");
  long long mib = e && atoll(e) > 0 ? atoll(e) : 2048;
  return mib << 20;
}

extern "C" int ngf_field_backward(NgfField h, const float* rays_dev, int64_t n_rays, int32_t ray_stride,
                                  int32_t n_samples, int32_t white_bg, const float* jitter_dev,
                                  const float* grad_rgb_dev, const NgfFieldGrads* grads, void* stream) {
  if (!h || !grads) return ngf_set_error(NGF_EINVAL, "NULL argument");
  if (n_rays < 0 || n_rays > 0x7fffffffll) return ngf_set_error(NGF_EINVAL, "n_rays=%lld", (long long)n_rays);
  if (n_rays == 0) return NGF_OK;
  if (!rays_dev || !grad_rgb_dev) return ngf_set_error(NGF_EINVAL, "NULL ray / gradient pointer");
  if (ray_stride < 6) return ngf_set_error(NGF_EINVAL, "ray_stride=%d (< 6)", ray_stride);
  const FieldDev& f = h->dev;
  const int V = f.variant;
  for (int i = 0; i < 3; ++i)
    if (!grads->plane[i]) return ngf_set_error(NGF_EINVAL, "grads.plane[%d] is NULL", i);
  if (!grads->rgb_basis || !grads->rgb_l1_w || !grads->rgb_l1_b || !grads->rgb_l2_w || !grads->rgb_l2_b || !grads->rgb_l3_w ||
      !grads->rgb_l3_b || !grads->dens_l1_w || !grads->dens_l1_b)
    return ngf_set_error(NGF_EINVAL, "a network gradient pointer is NULL");
  if (V == 1 && (!grads->dens_l2_w || !grads->dens_l2_b || !grads->dens_l3_w || !grads->dens_l3_b))
    return ngf_set_error(NGF_EINVAL, "InfoInv needs the gradients of all three density layers");
  const bool gauge = V == 0 && f.gauge_on;
  if (gauge)
    for (int i = 0; i < 3; ++i)
      if (!grads->gauge[i]) return ngf_set_error(NGF_EINVAL, "gauge is on but grads.gauge[%d] is NULL", i);
  if (!h->raw_w) return ngf_set_error(NGF_EINVAL, "the handle holds no fp32 network weights (packed by an older library?)");
  int dev_prev = -1;
  cudaGetDevice(&dev_prev);
  if (dev_prev != h->device) CU(cudaSetDevice(h->device));
  struct Restore { int d; ~Restore() { int c = -1; if (d >= 0 && cudaGetDevice(&c) == cudaSuccess && c != d) cudaSetDevice(d); } } restore{dev_prev};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int S = n_samples > 0 ? n_samples : h->n_samples_default;
  if (S < 1) return ngf_set_error(NGF_EINVAL, "n_samples resolves to %d", S);
  if (!h->train) h->train = new TrainWs();
  TrainWs* w = static_cast<TrainWs*>(h->train);

  // ---- record list: capacity from the budget (half of it), at most n_rays * S
  long long cap = n_rays * (long long)S;
  const long long cap_budget = train_budget_bytes() / 2 / (long long)(sizeof(BRec) + sizeof(int));
  if (cap > cap_budget) cap = cap_budget;
  if (cap > 0x7ffffff0ll) cap = 0x7ffffff0ll;
  if (w->rec_cap < cap) {
    CU(cudaStreamSynchronize(st));
    cudaFree(w->rec); cudaFree(w->active_list);
    w->rec = nullptr; w->active_list = nullptr; w->rec_cap = 0;
    CU(cudaMalloc(reinterpret_cast<void**>(&w->rec), (size_t)cap * sizeof(BRec)));
    CU(cudaMalloc(reinterpret_cast<void**>(&w->active_list), (size_t)cap * sizeof(int)));
    w->rec_cap = cap;
  }
  if (w->tail_cap < n_rays) {
    CU(cudaStreamSynchronize(st));
    cudaFree(w->tail);
    w->tail = nullptr; w->tail_cap = 0;
    CU(cudaMalloc(reinterpret_cast<void**>(&w->tail), (size_t)n_rays * sizeof(int)));
    w->tail_cap = n_rays;
  }
  if (!w->counters) CU(cudaMalloc(reinterpret_cast<void**>(&w->counters), 16));
  CU(cudaMemsetAsync(w->counters, 0, 16, st));
  BwdArgs a{rays_dev, jitter_dev, (long long)n_rays, ray_stride, S, w->rec, (unsigned)cap, w->tail, w->active_list, w->counters};
  if (V == 0) ngf_bwd_march_kernel<0><<<grid_for(n_rays, 256), 256, 0, st>>>(f, a);
  else ngf_bwd_march_kernel<1><<<grid_for(n_rays, 256), 256, 0, st>>>(f, a);
  count_launch();
  CU(cudaGetLastError());
  unsigned int cnt[4] = {0, 0, 0, 0};
  CU(cudaMemcpyAsync(cnt, w->counters, 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (cnt[2] || cnt[0] > (unsigned)cap)
    return ngf_set_error(NGF_ENOMEM, "the batch has more than %lld valid samples: raise NGF_TRAIN_MIB (now %lld MiB) or render "
                         "fewer rays per step", cap, train_budget_bytes() >> 20);
  const long long n_rec = cnt[0], n_act = cnt[1];
  PlaneGrads pg{};
  PlaneParams pp{};
  for (int i = 0; i < 3; ++i) pp.plane[i] = grads->plane_param[i];
  for (int i = 0; i < 3; ++i) { pg.plane[i] = grads->plane[i]; pg.gauge[i] = gauge ? grads->gauge[i] : nullptr; }

  // ---- colour branch on the active samples, in chunks that fit the activation workspace
  const int F = V == 0 ? Cfg<0>::F : Cfg<1>::F, LD = F + 15;
  const long long per_act = (long long)F + LD + 64 + 64 + 3 + 64 + 64 + F + F;      // X IN H1 H2 Z dH2 dH1 dIN dX
  const long long per_den = V == 1 ? (72 + 32 + 32 + 1 + 32 + 32 + 72) : 0;          // Xd G1 G2 dZ dG2 dG1 dXd
  long long want = train_budget_bytes() / 2 / 4;
  if (w->buf_floats < want) {
    cudaFree(w->buf);
    w->buf = nullptr; w->buf_floats = 0;
    CU(cudaMalloc(reinterpret_cast<void**>(&w->buf), (size_t)want * 4));
    w->buf_floats = want;
  }
  const float* RW = h->raw_w;                       // [B | W1 | b1 | W2 | b2 | W3 | b3]
  const float* B_ = RW; const float* W1 = B_ + (size_t)F * F; const float* b1 = W1 + (size_t)64 * LD;
  const float* W2 = b1 + 64; const float* b2 = W2 + 64 * 64; const float* W3 = b2 + 64; const float* b3 = W3 + 3 * 64;
  const long long chunk_max = w->buf_floats / per_act;
  if (chunk_max < 64) return ngf_set_error(NGF_ENOMEM, "NGF_TRAIN_MIB too small");
  int rc = NGF_OK;
  for (long long a0 = 0; a0 < n_act && rc == NGF_OK; a0 += chunk_max) {
    const long long A = (n_act - a0) < chunk_max ? (n_act - a0) : chunk_max;
    const int* al = w->active_list + a0;
    float* X = w->buf; float* IN = X + A * F; float* H1 = IN + A * LD; float* H2 = H1 + A * 64; float* Z = H2 + A * 64;
    float* dH2 = Z + A * 3; float* dH1 = dH2 + A * 64; float* dIN = dH1 + A * 64; float* dX = dIN + A * F;
    const int split = (int)((A + 2047) / 2048 > 256 ? 256 : (A + 2047) / 2048);
    if (V == 0) ngf_bwd_feat_kernel<0><<<grid_for(A * 18, 256), 256, 0, st>>>(f, w->rec, al, A, rays_dev, ray_stride, X, IN, pp);
    else ngf_bwd_feat_kernel<1><<<grid_for(A * 27, 256), 256, 0, st>>>(f, w->rec, al, A, rays_dev, ray_stride, X, IN, pp);
    count_launch();
    CU(cudaGetLastError());
    // forward: IN[:, :F] = X B^T ; H1 = relu(IN W1^T + b1) ; H2 = relu(H1 W2^T + b2) ; Z = H2 W3^T + b3
    if ((rc = gemm(st, X, F, 1, B_, 1, F, nullptr, nullptr, IN, LD, A, F, F, 0))) break;
    if ((rc = gemm(st, IN, LD, 1, W1, 1, LD, b1, nullptr, H1, 64, A, 64, LD, 1))) break;
    if ((rc = gemm(st, H1, 64, 1, W2, 1, 64, b2, nullptr, H2, 64, A, 64, 64, 1))) break;
    if ((rc = gemm(st, H2, 64, 1, W3, 1, 64, b3, nullptr, Z, 3, A, 3, 64, 0))) break;
    ngf_bwd_colour_out_kernel<<<grid_for(A, 256), 256, 0, st>>>(w->rec, al, A, grad_rgb_dev, Z);
    count_launch();
    CU(cudaGetLastError());
    // backward through the layers (Z now holds dL/dz3)
    if ((rc = gemm(st, Z, 3, 1, W3, 64, 1, nullptr, H2, dH2, 64, A, 64, 3, 2))) break;
    if ((rc = gemm(st, dH2, 64, 1, W2, 64, 1, nullptr, H1, dH1, 64, A, 64, 64, 2))) break;
    if ((rc = gemm(st, dH1, 64, 1, W1, LD, 1, nullptr, nullptr, dIN, F, A, F, 64, 0))) break;
    if ((rc = gemm(st, dIN, F, 1, B_, F, 1, nullptr, nullptr, dX, F, A, F, F, 0))) break;
    // weight gradients: dW[n][k] += sum_a D[a][n] * Act[a][k]
    if ((rc = gemm(st, Z, 1, 3, H2, 64, 1, nullptr, nullptr, grads->rgb_l3_w, 64, 3, 64, A, 0, split > 1 ? split : 2))) break;
    if ((rc = gemm(st, dH2, 1, 64, H1, 64, 1, nullptr, nullptr, grads->rgb_l2_w, 64, 64, 64, A, 0, split > 1 ? split : 2))) break;
    if ((rc = gemm(st, dH1, 1, 64, IN, LD, 1, nullptr, nullptr, grads->rgb_l1_w, LD, 64, LD, A, 0, split > 1 ? split : 2))) break;
    if ((rc = gemm(st, dIN, 1, F, X, F, 1, nullptr, nullptr, grads->rgb_basis, F, F, F, A, 0, split > 1 ? split : 2))) break;
    ngf_colsum_kernel<<<dim3(64, 1), 256, 0, st>>>(Z, A, 3, 3, grads->rgb_l3_b);
    ngf_colsum_kernel<<<dim3(64, 2), 256, 0, st>>>(dH2, A, 64, 64, grads->rgb_l2_b);
    ngf_colsum_kernel<<<dim3(64, 2), 256, 0, st>>>(dH1, A, 64, 64, grads->rgb_l1_b);
    count_launch(); count_launch(); count_launch();
    CU(cudaGetLastError());
    if (V == 0) ngf_bwd_colour_scatter_kernel<0><<<grid_for(A * 3, 128), 128, 0, st>>>(f, w->rec, al, A, dX, pg, gauge ? 1 : 0, pp);
    else ngf_bwd_colour_scatter_kernel<1><<<grid_for(A * 3, 128), 128, 0, st>>>(f, w->rec, al, A, dX, pg, 0, pp);
    count_launch();
    CU(cudaGetLastError());
  }
  if (rc) return rc;

  // ---- compositing backward
  ngf_bwd_composite_kernel<<<grid_for(n_rays, 128), 128, 0, st>>>(w->rec, w->tail, n_rays, grad_rgb_dev, white_bg ? 1 : 0);
  count_launch();
  CU(cudaGetLastError());
  if (n_rec == 0) return NGF_OK;

  // ---- density branch
  if (V == 0) {
    ngf_bwd_density_kernel<<<grid_for(n_rec, 128), 128, 0, st>>>(f, w->rec, n_rec, pg, grads->dens_l1_w, grads->dens_l1_b);
    count_launch();
    CU(cudaGetLastError());
    return NGF_OK;
  }
  const float* DW = h->raw_dw;                      // [W1 32x72 | b1 | W2 32x32 | b2 | W3 32 | b3]
  const float* D1 = DW; const float* db1 = D1 + 32 * 72; const float* D2 = db1 + 32; const float* db2 = D2 + 32 * 32;
  const float* D3 = db2 + 32; const float* db3 = D3 + 32;
  const long long dchunk_max = w->buf_floats / per_den;
  for (long long r0 = 0; r0 < n_rec; r0 += dchunk_max) {
    const long long n = (n_rec - r0) < dchunk_max ? (n_rec - r0) : dchunk_max;
    float* Xd = w->buf; float* G1 = Xd + n * 72; float* G2 = G1 + n * 32; float* dZ = G2 + n * 32;
    float* dG2 = dZ + n; float* dG1 = dG2 + n * 32; float* dXd = dG1 + n * 32;
    const int split = (int)((n + 2047) / 2048 > 256 ? 256 : ((n + 2047) / 2048 < 2 ? 2 : (n + 2047) / 2048));
    ngf_bwd_dfeat_kernel<<<grid_for(n * 18, 256), 256, 0, st>>>(f, w->rec, r0, n, Xd);
    count_launch();
    if ((rc = gemm(st, Xd, 72, 1, D1, 1, 72, db1, nullptr, G1, 32, n, 32, 72, 1))) return rc;
    if ((rc = gemm(st, G1, 32, 1, D2, 1, 32, db2, nullptr, G2, 32, n, 32, 32, 1))) return rc;
    ngf_bwd_dsig_kernel<<<grid_for(n, 256), 256, 0, st>>>(w->rec, r0, n, dZ);
    count_launch();
    CU(cudaGetLastError());
    if ((rc = gemm(st, dZ, 1, 1, D3, 32, 1, nullptr, G2, dG2, 32, n, 32, 1, 2))) return rc;
    if ((rc = gemm(st, dG2, 32, 1, D2, 32, 1, nullptr, G1, dG1, 32, n, 32, 32, 2))) return rc;
    if ((rc = gemm(st, dG1, 32, 1, D1, 72, 1, nullptr, nullptr, dXd, 72, n, 72, 32, 0))) return rc;
    if ((rc = gemm(st, dZ, 1, 1, G2, 32, 1, nullptr, nullptr, grads->dens_l3_w, 32, 1, 32, n, 0, split))) return rc;
    if ((rc = gemm(st, dG2, 1, 32, G1, 32, 1, nullptr, nullptr, grads->dens_l2_w, 32, 32, 32, n, 0, split))) return rc;
    if ((rc = gemm(st, dG1, 1, 32, Xd, 72, 1, nullptr, nullptr, grads->dens_l1_w, 72, 32, 72, n, 0, split))) return rc;
    ngf_colsum_kernel<<<dim3(64, 1), 256, 0, st>>>(dZ, n, 1, 1, grads->dens_l3_b);
    ngf_colsum_kernel<<<dim3(64, 1), 256, 0, st>>>(dG2, n, 32, 32, grads->dens_l2_b);
    ngf_colsum_kernel<<<dim3(64, 1), 256, 0, st>>>(dG1, n, 32, 32, grads->dens_l1_b);
    ngf_bwd_dscatter_kernel<<<grid_for(n * 3, 128), 128, 0, st>>>(f, w->rec, r0, n, dXd, pg);
    count_launch(); count_launch(); count_launch(); count_launch();
    CU(cudaGetLastError());
  }
  return NGF_OK;
}
