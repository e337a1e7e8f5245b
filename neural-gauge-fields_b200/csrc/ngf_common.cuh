// ngf_common.cuh — device-side field description and the point-wise building blocks of the render path.
//
// Every function names the reference code it implements (paths relative to /root/reference).  The
// decision-critical chain (sample position, bbox test, occupancy test) uses __f*_rn intrinsics so nvcc cannot
// contract mul+add into FMA: eager PyTorch rounds after every op, and a 1-ulp position change flips mask
// decisions (SURVEY.md §7 "bit-level decision parity").  Never build this with --use_fast_math.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ngf {

constexpr int kMid = 64;          // rgb_decoder middle_dim as instantiated (Field.py:28)
constexpr int kDensMid = 32;      // InfoInv density_decoder middle_dim (InfoInv/models/Field.py:23)

template <int V> struct Cfg;
template <> struct Cfg<0> {       // TriPlane
  static constexpr int DC = 16, AC = 48, F = 144, K1 = 160;
};
template <> struct Cfg<1> {       // InfoInv
  static constexpr int DC = 24, AC = 72, F = 216, K1 = 240;
};

struct PlaneDev {
  const float* dens;              // [H][W][DC] fp32, channels-last (InfoInv: feeds the density MLP)
  const float* dsum;              // [H][W] fp32: <density channels, density_decoder.weight slice> per texel (TriPlane)
  const __half* app;              // [H][W][AC] fp16, channels-last
  int H, W;
  float wm1, hm1;                 // float(W-1), float(H-1)
};

struct GaugeDev {
  const float2* g;                // [H][W] (du, dv)
  int H, W;
  float wm1, hm1;
};

constexpr int kTailFloats = 3 * 64 + 64 + 4;                        // 260

struct FieldDev {
  int variant;
  PlaneDev plane[3];              // xy (u=x,v=y), yz (u=y,v=z), xz (u=x,v=z)
  GaugeDev gauge[3];
  int gauge_on;
  int infoinv;
  float lo[3], hi[3], inv[3];
  float step, near_t, far_t, dscale, wthres, dshift;
  float tstop;                    // the march stops once T <= tstop: min(1e-6, wthres), 0 (never) when wthres <= 0
  // occupancy ("alpha mask"): raw bit grid + derived acceleration grids (built by ngf_field_pack)
  const uint32_t* occ;            // raw volume bits, bit (z*H + y)*W + x
  const uint32_t* occ2;           // "any of the 8 cell corners set" grid over cells x0 in [-1, W-1] (index x0+1),
                                  // 4x4x2-cell bricks per 32-bit word
  const uint32_t* occ_coarse;     // any occ2 bit set in an 8x8x8 block of occ2 cells, linear bits
  int occ_w, occ_h, occ_d;
  int occ2_nxb, occ2_nyb;         // bricks per row / per slab of occ2
  int occ_cx, occ_cy;             // coarse grid row / slab sizes
  int has_occ;
  float occ_lo[3], occ_inv[3];
  float clip_lo[3], clip_hi[3];   // world box containing every sample the mask can keep (conservative)
  // TriPlane density head Linear(48,1)
  float dw[48];
  float db;
  // InfoInv density MLP 72->32->32->1, fp32: [w1t 72x32 (input-major)][b1 32][w2 32x32][b2 32][w3 32][b3 1]
  const float* dmlp;
  // colour MLP, packed: w1p/w2p fp16 in tcgen05 K-major core-matrix order, tail fp32 [64 x (b2, w3_r, w3_g, w3_b)][b3 3][pad]
  const __half* w1p;
  const __half* w2p;
  const float* tail;
  // the same tail by value: kernels read it from their parameter bank with compile-time offsets (mlp_head_partial)
  float tail_c[kTailFloats];
  // three 128-byte TMA tensor maps (device memory) over the appearance planes [H][W][48] fp16 with 48 x 5 x 5 boxes, or
  // nullptr (InfoInv; driver without cuTensorMapEncodeTiled): ngf_colour_tma.cuh
  const void* tmap;
  // InfoInv density MLP for the tensor-core march (ngf_infoinv_tc.cuh): split fp16 weights in tcgen05 K-major order
  // [W1 hi | W1 lo] ([32 x 80], column 72 = b1) [W2 hi | W2 lo] ([32 x 48], column 32 = b2), and fp32 [w3 32][b3]
  const __half* ii_w;
  const float* ii_tail;
};

// Pinhole camera for on-device ray generation (TriPlane/dataLoader/ray_utils.py:24-42,66-87 + blender.py:46-52)
struct CamDev {
  float c2w[12];                  // row-major [3][4] camera-to-world
  float fx, fy, cx, cy;
  int W, H;
  long long base;                 // pixel index of local ray 0 (frames rendered in row blocks)
  // ray-sharded batches of several frames (ngf_comm.cu): local ray l is ray ((l / block) * world + rank) * block + l % block
  // of the batch, which is pixel g % (W*H) of frame g / (W*H); every frame has its own pose, poses[frame][12].
  const float* poses;             // nullptr: one frame, pose c2w
  int shard_block, shard_rank, shard_world;
};

// Ray of pixel `ray` (row-major): camera-space direction ((i+0.5-cx)/fx, (j+0.5-cy)/fy, 1), normalised
// (blender.py:52), rotated by c2w[:3,:3] (get_rays: directions @ c2w[:3,:3].T); origin = c2w[:3,3].
__device__ __forceinline__ void camera_ray(const CamDev& c, long long ray, float o[3], float d[3]) {
  ray += c.base;
  const float* c2w = c.c2w;
  if (c.poses) {
    const long long g = ((ray / c.shard_block) * c.shard_world + c.shard_rank) * c.shard_block + ray % c.shard_block;
    const long long per = (long long)c.W * c.H, frame = g / per;
    c2w = c.poses + frame * 12;
    ray = g - frame * per;
  }
  const int j = (int)(ray / c.W), i = (int)(ray - (long long)j * c.W);
  const float x = __fdiv_rn(__fsub_rn(__fadd_rn((float)i, 0.5f), c.cx), c.fx);
  const float y = __fdiv_rn(__fsub_rn(__fadd_rn((float)j, 0.5f), c.cy), c.fy);
  const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), 1.f));
  const float dx = __fdiv_rn(x, nrm), dy = __fdiv_rn(y, nrm), dz = __fdiv_rn(1.f, nrm);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    d[k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, c2w[4 * k]), __fmul_rn(dy, c2w[4 * k + 1])), __fmul_rn(dz, c2w[4 * k + 2]));
    o[k] = c2w[4 * k + 3];
  }
}

constexpr int kDmlpFloats = 32 * 72 + 32 + 32 * 32 + 32 + 32 + 1;   // 3457

// ---------------------------------------------------------------------------------------------------------
// Base.sample_ray prologue (FieldBase.py:121-125): first sample distance t0.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ray_t0(const FieldDev& f, const float o[3], const float d[3]) {
  float best = -INFINITY;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float vec = (d[k] == 0.f) ? 1e-6f : d[k];
    float ra = __fdiv_rn(__fsub_rn(f.hi[k], o[k]), vec);
    float rb = __fdiv_rn(__fsub_rn(f.lo[k], o[k]), vec);
    best = fmaxf(best, fminf(ra, rb));
  }
  return fminf(fmaxf(best, f.near_t), f.far_t);
}

// Conservative superset [i_lo, i_hi] of the sample indices whose position can be inside the box; the exact
// per-sample test still runs inside it, so results do not depend on this clip.  The slab distances are widened
// by the rounding error of p = o + d*t mapped back to t (huge for near-zero direction components, which then
// leave the axis unconstrained) plus two whole steps.
__device__ __forceinline__ void ray_index_range(const FieldDev& f, const float o[3], const float d[3], float t0,
                                                int S, int& i_lo, int& i_hi, float extra = 0.f) {
  float t_in = -INFINITY, t_out = INFINITY;
  bool empty = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // with an alpha mask only samples inside clip_lo..clip_hi (a superset of the occupied cells) can be kept
    const float lo = f.has_occ ? fmaxf(f.lo[k], f.clip_lo[k]) : f.lo[k];
    const float hi = f.has_occ ? fminf(f.hi[k], f.clip_hi[k]) : f.hi[k];
    if (lo > hi) empty = true;
    if (d[k] == 0.f) {                       // p_k == o_k exactly for every sample
      if (lo > o[k] || o[k] > hi) empty = true;
      continue;
    }
    float inv_d = 1.f / d[k];
    float a = (hi - o[k]) * inv_d, b = (lo - o[k]) * inv_d;
    float m = 4e-6f * (fabsf(o[k]) + fabsf(d[k]) * f.far_t + fmaxf(fabsf(f.lo[k]), fabsf(f.hi[k]))) * fabsf(inv_d);
    t_in = fmaxf(t_in, fminf(a, b) - m);     // NaN operands are ignored by fminf/fmaxf: axis left unconstrained
    t_out = fminf(t_out, fmaxf(a, b) + m);
  }
  if (empty || t_out < t_in) { i_lo = 0; i_hi = -1; return; }
  float a = floorf((t_in - t0) / f.step) - 2.f - extra;     // extra: a jittered sample i sits at t0 + step * (i + u)
  float b = ceilf((t_out - t0) / f.step) + 2.f + extra;
  a = fminf(fmaxf(a, 0.f), (float)S);
  b = fminf(fmaxf(b, -1.f), (float)(S - 1));
  i_lo = (int)a;
  i_hi = (int)b;
}

// t_i = t0 + stepSize * i  (FieldBase.py:131-132: mul, then add)
__device__ __forceinline__ float sample_t(const FieldDev& f, float t0, int i) {
  return __fadd_rn(t0, __fmul_rn(f.step, (float)i));
}
// training-time sampling (FieldBase.py:128-132): rng = i + u with one u ~ U[0,1) per ray, step = stepSize * rng
__device__ __forceinline__ float sample_t(const FieldDev& f, float t0, int i, float u) {
  return __fadd_rn(t0, __fmul_rn(f.step, __fadd_rn((float)i, u)));
}

// p = o + d * t (FieldBase.py:134) and the bbox test (FieldBase.py:135)
__device__ __forceinline__ bool sample_pos(const FieldDev& f, const float o[3], const float d[3], float t,
                                           float p[3]) {
  bool in = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    p[k] = __fadd_rn(o[k], __fmul_rn(d[k], t));
    in = in && !(f.lo[k] > p[k] || p[k] > f.hi[k]);
  }
  return in;
}

// ---------------------------------------------------------------------------------------------------------
// AlphaGridMask.sample_alpha(p) > 0 (FieldBase.py:33-40, used at :261-267) on the bit-packed volume.
// ATen grid_sampler_3d, align_corners=True, zeros padding: i = ((q+1)/2)*(size-1); corner weights are products of
// (floor(i)+1-i) (always > 0) and (i-floor(i)) (> 0 iff i is not an integer); the {0,1} volume sample is > 0 iff
// an in-range corner with non-zero weight is set.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool occ_bit(const FieldDev& f, int x, int y, int z) {
  if ((unsigned)x >= (unsigned)f.occ_w || (unsigned)y >= (unsigned)f.occ_h || (unsigned)z >= (unsigned)f.occ_d)
    return false;
  uint32_t idx = ((uint32_t)z * (uint32_t)f.occ_h + (uint32_t)y) * (uint32_t)f.occ_w + (uint32_t)x;
  return (__ldg(f.occ + (idx >> 5)) >> (idx & 31)) & 1u;
}

// Fast path: when no coordinate is an exact lattice value all 8 corners carry non-zero weight, so the answer is
// the precomputed "any corner set" bit of the cell (one load, after an L1-resident coarse test).  Exact lattice
// coordinates (measure zero, but the reference's own dense alpha grid produces them) take the 8-corner path.
__device__ __forceinline__ bool occ_keep(const FieldDev& f, const float p[3]) {
  float fi[3];
  int i0[3];
  bool frac[3];
  const int dims[3] = {f.occ_w, f.occ_h, f.occ_d};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float q = __fsub_rn(__fmul_rn(__fsub_rn(p[k], f.occ_lo[k]), f.occ_inv[k]), 1.f);
    fi[k] = __fmul_rn(__fmul_rn(__fadd_rn(q, 1.f), 0.5f), (float)(dims[k] - 1));
    float fl = floorf(fi[k]);
    frac[k] = fi[k] != fl;
    fl = fminf(fmaxf(fl, -2.f), (float)dims[k] + 1.f);
    i0[k] = (int)fl;
  }
  if (frac[0] && frac[1] && frac[2]) {
    const int X = i0[0] + 1, Y = i0[1] + 1, Z = i0[2] + 1;
    if ((unsigned)X > (unsigned)f.occ_w || (unsigned)Y > (unsigned)f.occ_h || (unsigned)Z > (unsigned)f.occ_d)
      return false;
    const uint32_t cb = ((uint32_t)(Z >> 3) * (uint32_t)f.occ_cy + (uint32_t)(Y >> 3)) * (uint32_t)f.occ_cx + (uint32_t)(X >> 3);
    if (!((__ldg(f.occ_coarse + (cb >> 5)) >> (cb & 31)) & 1u)) return false;
    const uint32_t word = ((uint32_t)(Z >> 1) * (uint32_t)f.occ2_nyb + (uint32_t)(Y >> 2)) * (uint32_t)f.occ2_nxb + (uint32_t)(X >> 2);
    const uint32_t bit = (uint32_t)(X & 3) | ((uint32_t)(Y & 3) << 2) | ((uint32_t)(Z & 1) << 4);
    return (__ldg(f.occ2 + word) >> bit) & 1u;
  }
  bool keep = false;
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        bool w = (!dx || frac[0]) && (!dy || frac[1]) && (!dz || frac[2]);
        if (w) keep = keep || occ_bit(f, i0[0] + dx, i0[1] + dy, i0[2] + dz);
      }
  return keep;
}

// Base.normalize_coord (FieldBase.py:88-89): (p - aabb0) * invaabbSize - 1
__device__ __forceinline__ void unit_coords(const FieldDev& f, const float p[3], float n[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) n[k] = __fsub_rn(__fmul_rn(__fsub_rn(p[k], f.lo[k]), f.inv[k]), 1.f);
}

// ---------------------------------------------------------------------------------------------------------
// grid_sample(bilinear, align_corners=True, zeros) tap set for one lookup (ATen GridSampler: ix=((u+1)/2)*(W-1)).
// Taps outside the plane get weight 0 and a clamped (in-bounds) address.
// ---------------------------------------------------------------------------------------------------------
struct Taps {
  int off[4];     // texel index y*W+x
  float w[4];
};

__device__ __forceinline__ Taps make_taps(float u, float v, int W, int H, float wm1, float hm1) {
  float ix = ((u + 1.f) * 0.5f) * wm1;
  float iy = ((v + 1.f) * 0.5f) * hm1;
  ix = fminf(fmaxf(ix, -2.f), (float)W + 1.f);      // also maps NaN to a bound: never an OOB address
  iy = fminf(fmaxf(iy, -2.f), (float)H + 1.f);
  float x0f = floorf(ix), y0f = floorf(iy);
  float fx = ix - x0f, fy = iy - y0f;
  int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
  bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)x1 < (unsigned)W;
  bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)y1 < (unsigned)H;
  int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
  int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
  Taps t;
  t.off[0] = cy0 * W + cx0; t.w[0] = (vx0 && vy0) ? (1.f - fx) * (1.f - fy) : 0.f;
  t.off[1] = cy0 * W + cx1; t.w[1] = (vx1 && vy0) ? fx * (1.f - fy) : 0.f;
  t.off[2] = cy1 * W + cx0; t.w[2] = (vx0 && vy1) ? (1.f - fx) * fy : 0.f;
  t.off[3] = cy1 * W + cx1; t.w[3] = (vx1 && vy1) ? fx * fy : 0.f;
  return t;
}

__device__ __forceinline__ float2 gauge_lookup(const GaugeDev& g, float u, float v) {
  Taps t = make_taps(u, v, g.W, g.H, g.wm1, g.hm1);
  float2 r = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 s = __ldg(g.g + t.off[k]);
    r.x += t.w[k] * s.x;
    r.y += t.w[k] * s.y;
  }
  return r;
}

// TriPlane.compute_gauge (TriPlane/models/Field.py:53-75) / InfoInv transform (InfoInv/models/Field.py:43-50).
// c = {u_xy, v_xy, u_yz, v_yz, u_xz, v_xz}.  Association order of the adds follows the reference.
__device__ __forceinline__ void gauge_coords(const FieldDev& f, const float n[3], bool gauge_on, float c[6]) {
  const float x = n[0], y = n[1], z = n[2];
  if (!gauge_on) {
    c[0] = x; c[1] = y; c[2] = y; c[3] = z; c[4] = x; c[5] = z;
    return;
  }
  float2 gxy = gauge_lookup(f.gauge[0], x, y);
  float2 gyz = gauge_lookup(f.gauge[1], y, z);
  float2 gxz = gauge_lookup(f.gauge[2], x, z);
  c[0] = __fadd_rn(__fadd_rn(x, gxy.x), gxz.x);
  c[1] = __fadd_rn(__fadd_rn(y, gxy.y), gyz.x);
  c[2] = __fadd_rn(__fadd_rn(y, gyz.x), gxy.y);
  c[3] = __fadd_rn(__fadd_rn(z, gyz.y), gxz.y);
  c[4] = __fadd_rn(__fadd_rn(x, gxz.x), gxy.x);
  c[5] = __fadd_rn(__fadd_rn(z, gxz.y), gyz.y);
}

// F.softplus(x) with torch defaults (beta=1, threshold=20)
__device__ __forceinline__ float softplus_torch(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// ---------------------------------------------------------------------------------------------------------
// compute_density, TriPlane (Field.py:77-91): 3 x bilinear over channels [0,16), Linear(48,1), softplus(.-10).
// Both steps are linear, so the Linear is applied per texel when the field is packed (PlaneDev::dsum) and the
// lookup blends one scalar per tap: sum_planes sum_taps w_tap * <texel, weight slice>.
__device__ __forceinline__ float sigma_triplane(const FieldDev& f, const float c[6]) {
  float acc = f.db;
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
    const PlaneDev& P = f.plane[pl];
    Taps t = make_taps(c[2 * pl], c[2 * pl + 1], P.W, P.H, P.wm1, P.hm1);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += t.w[k] * __ldg(P.dsum + t.off[k]);
  }
  return softplus_torch(acc + f.dshift);
}

// positional_encoding value for channel ch of PE(xyz, NF) (InfoInv/models/networks.py:227-237): layout
// [sin(x*2^0..2^(NF-1)), sin(y*..), sin(z*..), cos(same)].  x*2^j is exact in fp32, as in torch.
template <int NF>
__device__ __forceinline__ float phase_value(const float xyz[3], int ch) {
  bool is_cos = ch >= 3 * NF;
  int r = is_cos ? ch - 3 * NF : ch;
  float a = xyz[r / NF] * (float)(1 << (r % NF));
  return is_cos ? cosf(a) : sinf(a);
}

// compute_density, InfoInv (InfoInv/models/Field.py:52-70 + networks.py:34-54): 3 x bilinear over channels
// [0,24) (x phase code of xyz with 4 bands when infoinv), MLP 72->32->32->1 (ReLU), softplus(.-10).
// `w` points at the fp32 density MLP (shared memory in the render kernel, global in the point-wise kernel).
// fp32 throughout; the two hidden layers use the packed fma.rn.f32x2 of sm_100 (two exact fp32 FMAs per instruction).
__device__ __forceinline__ float sigma_infoinv(const FieldDev& f, const float c[6], const float* __restrict__ w) {
  const float xyz[3] = {c[0], c[1], c[3]};
  // phase code PE(xyz, 4): sin / cos of x * 2^k, k < 4 — one sincosf per coordinate, then the exact double-angle identities
  // (arguments <= 8 rad; the recurrence error stays below 1e-6)
  float sn[12], cs[12];
  if (f.infoinv) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s_, c_;
      sincosf(xyz[a], &s_, &c_);
      sn[4 * a] = s_; cs[4 * a] = c_;
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        const float s2 = 2.f * s_ * c_, c2 = 1.f - 2.f * s_ * s_;
        s_ = s2; c_ = c2;
        sn[4 * a + k] = s_; cs[4 * a + k] = c_;
      }
    }
  }
  float2 h1[kDensMid / 2];
  const float2* b1 = reinterpret_cast<const float2*>(w + 32 * 72);
#pragma unroll
  for (int j = 0; j < kDensMid / 2; ++j) h1[j] = b1[j];
  for (int pl = 0; pl < 3; ++pl) {
    const PlaneDev& P = f.plane[pl];
    Taps t = make_taps(c[2 * pl], c[2 * pl + 1], P.W, P.H, P.wm1, P.hm1);
#pragma unroll
    for (int q = 0; q < 6; ++q) {                    // 6 x float4 = 24 channels
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 a = __ldg(reinterpret_cast<const float4*>(P.dens) + (size_t)t.off[k] * 6 + q);
        v.x += t.w[k] * a.x; v.y += t.w[k] * a.y; v.z += t.w[k] * a.z; v.w += t.w[k] * a.w;
      }
      float fe[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ch = 4 * q + e;                    // channel of PE(xyz,4): [sin x(4) sin y(4) sin z(4) cos x(4) cos y(4) cos z(4)]
        float x = fe[e];
        if (f.infoinv) x *= ch < 12 ? sn[ch] : cs[ch - 12];
        const float2 xx = make_float2(x, x);
        const float2* wrow = reinterpret_cast<const float2*>(w + (pl * 24 + ch) * 32);   // [72][32] input-major
#pragma unroll
        for (int j = 0; j < kDensMid / 2; ++j) h1[j] = __ffma2_rn(wrow[j], xx, h1[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kDensMid / 2; ++j) { h1[j].x = fmaxf(h1[j].x, 0.f); h1[j].y = fmaxf(h1[j].y, 0.f); }
  const float* b2 = w + 32 * 72 + 32 + 32 * 32;
  const float2* w2 = reinterpret_cast<const float2*>(w + 32 * 72 + 32);
  const float* w3 = b2 + 32;
  float out = w3[32];
#pragma unroll 4
  for (int j = 0; j < kDensMid; ++j) {
    float2 s2 = make_float2(b2[j], 0.f);
#pragma unroll
    for (int k = 0; k < kDensMid / 2; ++k) s2 = __ffma2_rn(w2[j * 16 + k], h1[k], s2);
    out += w3[j] * fmaxf(s2.x + s2.y, 0.f);
  }
  return softplus_torch(out + f.dshift);
}

}  // namespace ngf
