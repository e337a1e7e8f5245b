// ngf_infoinv_march.cuh — the march of the InfoInv field (InfoInv/models/FieldBase.py:228-282) as a three-phase
// cooperative kernel.
//
// InfoInv's density is a 72 -> 32 -> 32 -> 1 MLP per kept sample (InfoInv/models/Field.py:52-70, networks.py:34-54).  Inside
// the one-ray-per-lane march it ran at ~19 % lane efficiency: a warp evaluates the MLP whenever ANY of its 32 rays has found
// a sample (0.78 of the 0.97 ms frame, profiles/r01).  Here the rays are marched in rounds of up to kLook kept samples:
//
//   phase 1  one ray per lane (8x4-pixel warp tiles in round 0, the compacted list of still-live rays afterwards): skip empty
//            space exactly as ngf_march_kernel does and record the next <= kLook kept samples (normalised position, t, delta);
//            rays that found any go to the round's hit list.
//   phase 2  one kept sample per LANE over the hit list: the density MLP in fp32 (sigma_infoinv, bit-identical to the old
//            path) at full lane occupancy.
//   phase 3  one hit ray per lane: alpha / transmittance / weights in sample order (raw2alpha, FieldBase.py:12-19), acc and
//            depth sums, colour work items for weights above the threshold, early-out at T <= tstop; rays with samples
//            left go to the next round's live list.
//
// The phases are separated by grid-wide barriers (cooperative launch, one resident grid), so the whole march is still ONE
// kernel launch with no host round trip; a ray behind an opaque surface costs at most kLook - 1 wasted density evaluations.
#pragma once
#include <cooperative_groups.h>
#include <cstdio>

#include "ngf_internal.h"
#include "ngf_mlp.cuh"

namespace ngf {

constexpr int kLook = 4;                    // kept samples per ray and round
constexpr int kIiThreads = 256;

struct __align__(16) IiSample {             // 32 bytes
  float n[3];                               // normalised position (InfoInv has no gauge: plane coords are (x,y), (y,z), (x,z))
  float t, delta, sigma;
  float pad[2];
};

struct __align__(16) IiRay {                // 32 bytes of per-ray state between rounds
  float T, acc, dep, last_col;
  int i, i_end;                             // next sample index / end of the conservative index range
  int n_found;                              // kept samples recorded this round
  int pad;
};

struct IiWs {                               // carved out of the render workspace by the host (ngf_abi.cu)
  IiRay* ray;                               // [R]
  IiSample* sample;                         // [R][kLook]
  int* list[2];                             // live lists (ping-pong), [R] each
  int* hit;                                 // [R] rays that recorded samples this round
  unsigned int* counts;                     // [4]: live[0], live[1], hit, unused   (zeroed by the kernel)
};

namespace cg = cooperative_groups;

template <bool JIT>
__global__ void __launch_bounds__(kIiThreads, 2) ngf_infoinv_march_kernel(const __grid_constant__ FieldDev f,
                                                                          const __grid_constant__ RenderArgs a,
                                                                          const __grid_constant__ IiWs ws) {
  extern __shared__ __align__(128) uint8_t smem[];
  cg::grid_group grid = cg::this_grid();
  constexpr unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warps_per_cta = kIiThreads / 32;
  const long long gwarp = (long long)blockIdx.x * warps_per_cta + warp, n_gwarps = (long long)gridDim.x * warps_per_cta;
  const long long gtid = (long long)blockIdx.x * kIiThreads + tid, n_gthreads = (long long)gridDim.x * kIiThreads;
  // density MLP weights in shared memory (phase 2) and the per-warp colour staging area (phase 3)
  float* dm = reinterpret_cast<float*>(smem);
  for (int i = tid; i < kDmlpFloats; i += kIiThreads) dm[i] = __ldg(f.dmlp + i);
  if (gtid < 4) ws.counts[gtid] = 0;
  uint32_t st_box = 0, st_den = 0, st_col = 0;
  const int S = a.S;
  __syncthreads();
  grid.sync();

  for (int round = 0;; ++round) {
    const int cur = round & 1;
    // ------------------------------------------------------------------ phase 1: find the next <= kLook kept samples
    const long long n_units = round == 0 ? (long long)a.n_tiles : ((long long)ws.counts[cur] + 31) / 32;
    for (long long unit = gwarp; unit < n_units; unit += n_gwarps) {
      long long ray = -1;
      if (round == 0) {
        if (a.img_w > 0) {
          const int tiles_x = (a.img_w + 7) >> 3;
          const int px = (int)(unit % tiles_x) * 8 + (lane & 7), py = (int)(unit / tiles_x) * 4 + (lane >> 3);
          ray = (px < a.img_w && py < a.img_h) ? (long long)py * a.img_w + px : -1;
        } else {
          ray = unit * 32 + lane;
          if (ray >= a.n_rays) ray = -1;
        }
      } else {
        const long long k = unit * 32 + lane;
        ray = k < (long long)ws.counts[cur] ? ws.list[cur][k] : -1;
      }
      bool hit = false;
      if (ray >= 0) {
        float o[3], d[3], last_col;
        if (a.cam_on) {
          camera_ray(a.cam, ray, o, d);
          last_col = d[2];
        } else {
          const float* rp = a.rays + ray * a.ray_stride;
#pragma unroll
          for (int k = 0; k < 3; ++k) { o[k] = __ldg(rp + k); d[k] = __ldg(rp + 3 + k); }
          last_col = __ldg(rp + a.ray_stride - 1);
        }
        const float t0 = ray_t0(f, o, d);
        const float jit = JIT ? __ldg(a.jitter + ray) : 0.f;
        IiRay rs;
        if (round == 0) {
          int lo_i, hi_i;
          ray_index_range(f, o, d, t0, S, lo_i, hi_i, JIT ? 1.f : 0.f);
          rs.T = 1.f; rs.acc = 0.f; rs.dep = 0.f; rs.last_col = last_col;
          rs.i = lo_i; rs.i_end = hi_i + 1;
          float* rgb = a.rgb + ray * 3;                    // the colour kernel accumulates into it
          rgb[0] = 0.f; rgb[1] = 0.f; rgb[2] = 0.f;
        } else {
          rs = ws.ray[ray];
        }
        int n_found = 0, i = rs.i;
        IiSample* out = ws.sample + ray * kLook;
        while (i < rs.i_end && n_found < kLook) {
          const float t = JIT ? sample_t(f, t0, i, jit) : sample_t(f, t0, i);
          float p[3];
          bool in = sample_pos(f, o, d, t, p);
          if (in) ++st_box;
          if (in && f.has_occ) in = occ_keep(f, p);
          if (in) {
            IiSample s;
            unit_coords(f, p, s.n);
            const float tn = JIT ? sample_t(f, t0, i + 1, jit) : sample_t(f, t0, i + 1);
            s.t = t;
            s.delta = (i == S - 1) ? 0.f : __fmul_rn(__fsub_rn(tn, t), f.dscale);
            s.sigma = 0.f; s.pad[0] = 0.f; s.pad[1] = 0.f;
            out[n_found++] = s;
          }
          ++i;
        }
        rs.i = i;
        rs.n_found = n_found;
        hit = n_found > 0;
        if (hit || round == 0) ws.ray[ray] = rs;
        if (!hit) {                                        // the ray has run out of samples: acc_map / depth_map
          a.acc[ray] = rs.acc;
          a.depth[ray] = rs.dep + (1.f - rs.acc) * rs.last_col;
        }
      }
      const unsigned hm = __ballot_sync(FULL, hit);
      if (hm) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(ws.counts + 2, (unsigned)__popc(hm));
        base = __shfl_sync(FULL, base, 0);
        if (hit) ws.hit[base + __popc(hm & ((1u << lane) - 1u))] = (int)ray;
      }
    }
    grid.sync();
    const unsigned n_hit = ws.counts[2];
#ifdef NGF_II_TRACE
    unsigned long long tr0 = 0, tr1 = 0, tr2 = 0;
    if (gtid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr0));
#endif
    if (n_hit == 0) break;
    // ------------------------------------------------------------------ phase 2: density of every recorded sample
    if (gtid == 0) ws.counts[cur ^ 1] = 0;                 // next round's live list: filled in phase 3, last read a round ago
    for (long long idx = gtid; idx < (long long)n_hit * kLook; idx += n_gthreads) {
      const int ray = ws.hit[idx / kLook], k = (int)(idx % kLook);
      if (k < ws.ray[ray].n_found) {
        IiSample* s = ws.sample + (long long)ray * kLook + k;
        const float c[6] = {s->n[0], s->n[1], s->n[1], s->n[2], s->n[0], s->n[2]};
        s->sigma = sigma_infoinv(f, c, dm);
        ++st_den;
      }
    }
    grid.sync();
#ifdef NGF_II_TRACE
    if (gtid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr1));
#endif
    // ------------------------------------------------------------------ phase 3: composite the round's samples in order
    for (long long base_k = gwarp * 32; base_k < (long long)n_hit; base_k += n_gwarps * 32) {
      const long long k = base_k + lane;
      const int ray = k < (long long)n_hit ? ws.hit[k] : -1;
      IiRay rs{};
      bool live = false;
      if (ray >= 0) rs = ws.ray[ray];
      const int n_found = ray >= 0 ? rs.n_found : 0;
      bool stopped = false;
      for (int j = 0; j < kLook; ++j) {                    // lock-step over the sample slot so pushes can be aggregated
        bool push = false;
        QEntry e;
        if (j < n_found && !stopped) {
          const IiSample s = ws.sample[(long long)ray * kLook + j];
          const float alpha = __fsub_rn(1.f, expf(-__fmul_rn(s.sigma, s.delta)));
          const float w = __fmul_rn(alpha, rs.T);
          rs.T = __fmul_rn(rs.T, __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f));
          rs.acc += w;
          rs.dep += w * s.t;
          push = w > f.wthres;
          if (push) {
            e.c[0] = s.n[0]; e.c[1] = s.n[1]; e.c[2] = s.n[1]; e.c[3] = s.n[2]; e.c[4] = s.n[0]; e.c[5] = s.n[2];
            e.w = w; e.id = ray;
            ++st_col;
          }
          if (rs.T <= f.tstop) stopped = true;
        }
        const unsigned pm = __ballot_sync(FULL, push);
        if (pm) {
          unsigned qb = 0;
          if (lane == 0) qb = atomicAdd(a.queue_count, (unsigned)__popc(pm));
          qb = __shfl_sync(FULL, qb, 0);
          if (push) {
            const unsigned slot = qb + __popc(pm & ((1u << lane) - 1u));
            if (slot < a.queue_cap) a.queue[slot] = e;
          }
        }
      }
      if (ray >= 0) {
        live = !stopped && rs.i < rs.i_end;
        if (live) {
          ws.ray[ray] = rs;
        } else {
          a.acc[ray] = rs.acc;
          a.depth[ray] = rs.dep + (1.f - rs.acc) * rs.last_col;
        }
      }
      const unsigned lm = __ballot_sync(FULL, live);
      if (lm) {
        unsigned lb = 0;
        if (lane == 0) lb = atomicAdd(ws.counts + (cur ^ 1), (unsigned)__popc(lm));
        lb = __shfl_sync(FULL, lb, 0);
        if (live) ws.list[cur ^ 1][lb + __popc(lm & ((1u << lane) - 1u))] = ray;
      }
    }
    grid.sync();
    if (gtid == 0) ws.counts[2] = 0;                       // hit count of the next round (phase 1 comes after this barrier...)
    const unsigned n_live = ws.counts[cur ^ 1];
#ifdef NGF_II_TRACE
    if (gtid == 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr2));
      printf("round %d: hit %u live %u | t(end of phase 1) %llu ns, phase 2 %llu ns, phase 3 %llu ns\n", round, n_hit, n_live, tr0, tr1 - tr0, tr2 - tr1);
    }
#endif
    grid.sync();                                           // ... and nobody may still be reading the old value
    if (n_live == 0) break;
  }

  // statistics
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    st_box += __shfl_xor_sync(FULL, st_box, s);
    st_den += __shfl_xor_sync(FULL, st_den, s);
    st_col += __shfl_xor_sync(FULL, st_col, s);
  }
  if (lane == 0) {
    if (st_box) atomicAdd(a.stats + 0, (unsigned long long)st_box);
    if (st_den) atomicAdd(a.stats + 1, (unsigned long long)st_den);
    if (st_col) atomicAdd(a.stats + 2, (unsigned long long)st_col);
  }
}

}  // namespace ngf
