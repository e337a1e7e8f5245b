// Micro-benchmark: how should 32 lanes fetch the 4 x 3 bilinear taps (48 fp16 channels = 96 B per texel) of 128 samples so
// that the L1 data pipe sees the fewest wavefronts?  Same access statistics as ngf_colour_kernel's gather (4 bursts of 32
// lock-step rays per tile, 0.42 texel per pixel, +-1.5 texels of per-ray offset), no MMA, result folded into a checksum.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o gather_probe gather_probe.cu && ./gather_probe
// Variants:
//   0  96-B texels, 6 lanes x LDG.128 per tap (the kernel's mapping)
//   1  128-B texels (padded), 6 lanes x LDG.128 per tap
//   2  128-B texels, 8 lanes x LDG.128 per tap (quarter-warp == one aligned line; 2 lanes fetch padding)
//   3  96-B texels, 3 lanes x LDG.256 per tap
//   4  96-B texels, 12 lanes x LDG.64 per tap
//   5  96-B texels, x-pair merged: 12 lanes x LDG.128 fetch taps (x0,y),(x0+1,y) = 192 contiguous bytes
//   6  128-B texels, x-pair merged: 16 lanes x LDG.128 (half-warp == 2 aligned lines)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kRes = 256, kTile = 128, kThreads = 256;

struct Tap { int x0, y0; float fx, fy; };                // per (tile, plane, sample)

template <int BYTES>
__device__ __forceinline__ void ld(const char* p, uint32_t* r);
template <> __device__ __forceinline__ void ld<8>(const char* p, uint32_t* r) {
  asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "l"(p));
}
template <> __device__ __forceinline__ void ld<16>(const char* p, uint32_t* r) {
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
}
template <> __device__ __forceinline__ void ld<32>(const char* p, uint32_t* r) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}

// LPT lanes per tap, BYTES per lane load, STRIDE texel stride in bytes, PAIR: one load covers the x-pair
template <int LPT, int BYTES, int STRIDE, bool PAIR, int USED = (PAIR ? 2 : 1) * STRIDE>
__global__ void __launch_bounds__(kThreads, 2) gather_kernel(const char* __restrict__ planes, size_t plane_bytes,
                                                             const Tap* __restrict__ taps, int n_tiles, uint32_t* out) {
  constexpr int W = BYTES / 4;
  constexpr int ITEMS = kTile * LPT;                     // (sample, lane-slot) items per plane and tile
  constexpr int J = (ITEMS + kThreads - 1) / kThreads;
  constexpr int NT = PAIR ? 2 : 4;                       // loads per item
  uint32_t cs = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    uint32_t raw[J][NT][W];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      const char* base = planes + pl * plane_bytes;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int it = threadIdx.x + kThreads * j;
        if (ITEMS % kThreads != 0 && it >= ITEMS) continue;
        const int m = it / LPT, slot = it % LPT;
        const Tap t = taps[((size_t)tile * 3 + pl) * kTile + m];
        if (PAIR) {
          // slot covers bytes [slot*BYTES, +BYTES) of the 2-texel run starting at (x0, y)
          if (slot * BYTES < USED) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              ld<BYTES>(base + ((size_t)(t.y0 + k) * kRes + t.x0) * STRIDE + slot * BYTES, raw[j][k]);
          }
        } else {
          if (slot * BYTES < USED) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ld<BYTES>(base + ((size_t)(t.y0 + (k >> 1)) * kRes + t.x0 + (k & 1)) * STRIDE + slot * BYTES, raw[j][k]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int it = threadIdx.x + kThreads * j;
        if (ITEMS % kThreads != 0 && it >= ITEMS) continue;
        const int slot = it % LPT;
        if (slot * BYTES < USED) {
#pragma unroll
          for (int k = 0; k < NT; ++k)
#pragma unroll
            for (int w = 0; w < W; ++w) cs += raw[j][k][w] * (k + 1);
        }
      }
    }
  }
  if (cs == 0x12345678u) out[0] = cs;
}

int main(int argc, char** argv) {
  const int n_tiles = argc > 1 ? atoi(argv[1]) : 148 * 2 * 64;
  const int only = argc > 2 ? atoi(argv[2]) : -1;
  std::vector<Tap> taps((size_t)n_tiles * 3 * kTile);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((s >> 11) & 0xFFFFFF) / 16777216.f; };
  for (int t = 0; t < n_tiles; ++t)
    for (int pl = 0; pl < 3; ++pl)
      for (int b = 0; b < 4; ++b) {
        const float bx = 8.f + rnd() * (kRes - 24.f), by = 8.f + rnd() * (kRes - 24.f);
        for (int i = 0; i < 32; ++i) {
          float x = bx + (i % 8) * 0.42f + (rnd() - 0.5f) * 3.f, y = by + (i / 8) * 0.42f + (rnd() - 0.5f) * 3.f;
          x = fminf(fmaxf(x, 0.f), kRes - 2.001f); y = fminf(fmaxf(y, 0.f), kRes - 2.001f);
          Tap tp; tp.x0 = (int)x; tp.y0 = (int)y; tp.fx = x - tp.x0; tp.fy = y - tp.y0;
          taps[((size_t)t * 3 + pl) * kTile + b * 32 + i] = tp;
        }
      }
  Tap* d_taps; CK(cudaMalloc(&d_taps, taps.size() * sizeof(Tap)));
  CK(cudaMemcpy(d_taps, taps.data(), taps.size() * sizeof(Tap), cudaMemcpyHostToDevice));
  const size_t pb = (size_t)kRes * kRes * 128;           // room for either stride
  char* d_planes; CK(cudaMalloc(&d_planes, 3 * pb)); CK(cudaMemset(d_planes, 1, 3 * pb));
  uint32_t* d_out; CK(cudaMalloc(&d_out, 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 2;
  auto run = [&](int v, const char* name, auto kern) {
    if (only >= 0 && only != v) return;
    for (int i = 0; i < 2; ++i) kern<<<grid, kThreads>>>(d_planes, pb, d_taps, n_tiles, d_out);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) kern<<<grid, kThreads>>>(d_planes, pb, d_taps, n_tiles, d_out);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double samples = (double)n_tiles * kTile;
    printf("variant %d %-44s %8.3f ms  %6.2f G samples/s  %7.1f GB/s useful\n", v, name, ms, samples / ms * 1e-6,
           samples * 1152 / ms * 1e-6);
  };
  run(0, "96B texel, 6 x LDG.128", gather_kernel<6, 16, 96, false>);
  run(1, "128B texel, 6 x LDG.128", gather_kernel<6, 16, 128, false>);
  run(2, "128B texel, 8 x LDG.128 (whole line)", gather_kernel<8, 16, 128, false>);
  run(7, "128B texel, 8 lanes, 2 predicated off", gather_kernel<8, 16, 128, false, 96>);
  run(3, "96B texel, 3 x LDG.256", gather_kernel<3, 32, 96, false>);
  run(4, "96B texel, 12 x LDG.64", gather_kernel<12, 8, 96, false>);
  run(5, "96B texel, x-pair 12 x LDG.128", gather_kernel<12, 16, 96, true>);
  run(6, "128B texel, x-pair 16 x LDG.128", gather_kernel<16, 16, 128, true>);
  return 0;
}
