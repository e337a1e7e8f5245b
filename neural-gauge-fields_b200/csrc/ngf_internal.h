// ngf_internal.h — host-side launch interface between the C ABI (ngf_abi.cu) and the kernels (ngf_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ngf_common.cuh"
#include "ngf_queue.h"

namespace ngf {

// Function attributes (dynamic shared-memory limit, carve-out) belong to the function ON ONE DEVICE: a process that
// renders on several devices has to set them on each.  PerDevice<T> keeps one lazily filled slot per device ordinal.
template <class T>
struct PerDevice {
  T slot[64] = {};
  bool done[64] = {};
  // returns the slot of the current device and whether it still has to be initialised
  T* get(bool* fresh) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    *fresh = !done[dev];
    done[dev] = true;
    return &slot[dev];
  }
  void retry() {               // initialisation failed: try again on the next call
    int dev = 0;
    cudaGetDevice(&dev);
    done[dev & 63] = false;
  }
};

struct RenderArgs {
  CamDev cam;                  // cam_on: rays are generated from this camera (pixel = ray index), `rays` is unused
  int cam_on;
  const float* rays;
  const float* jitter;         // [R] or nullptr: per-ray u of the training-time sampling (FieldBase.py:128-130)
  long long n_rays;
  int ray_stride;
  int S;
  int white_bg;
  int img_w, img_h;            // > 0: rays are row-major pixels of an img_w x img_h image -> 8x4 pixel warp tiles
  float* rgb;                  // [R][3]: zeroed by the march kernel, accumulated by the colour kernel, finalised in place
  float* depth;                // [R]
  float* acc;                  // [R]
  unsigned int* tile_counter;  // zero-initialised
  unsigned int* queue_count;   // zero-initialised: colour work items appended so far
  QEntry* queue;               // [queue_cap] colour work items (march kernel -> colour kernel)
  unsigned int queue_cap;
  void* ii_ws;                 // InfoInv: workspace of the tensor-core march (ngf_infoinv_tc.cuh: 64 B + 4 B per ray + 32 B per
                               // sample slot) or of the opt-in three-phase march (ngf_infoinv_march.cuh: 192 B per ray + 64 B)
  int ii_tc;                   // 1: ii_ws is laid out for the tensor-core march
  unsigned long long* stats;   // [5]: samples_in_box, samples_density, samples_colour, mlp_tiles, direct_patches (accumulated)
  int n_tiles;
};

// All launchers return cudaGetLastError() after enqueueing on `st`.
cudaError_t launch_march(const FieldDev& f, const RenderArgs& a, int num_sms, cudaStream_t st);
cudaError_t launch_colour(const FieldDev& f, const RenderArgs& a, int mlp_impl, int num_sms, cudaStream_t st);
cudaError_t launch_finalize(float* rgb, const float* acc, long long n_rays, int white_bg, cudaStream_t st);
cudaError_t launch_sample_ray(const FieldDev& f, const float* rays, long long n_rays, int stride, int S,
                              const float* jitter, float* pts, float* t, uint8_t* inside, cudaStream_t st);
cudaError_t launch_alpha_keep(const FieldDev& f, const float* pts, long long n, uint8_t* keep, cudaStream_t st);
cudaError_t launch_alpha_value(const FieldDev& f, const float* pts, long long n, float* out, cudaStream_t st);
cudaError_t launch_gauge(const FieldDev& f, const float* xyz, long long n, int gauge_on, float* xy, float* yz,
                         float* xz, cudaStream_t st);
cudaError_t launch_density(const FieldDev& f, const float* xy, const float* yz, const float* xz, long long n,
                           float* sigma, cudaStream_t st);
cudaError_t launch_sigma_world(const FieldDev& f, const float* pts, long long n, int use_gauge, float* sigma,
                               cudaStream_t st);
cudaError_t launch_rgb(const FieldDev& f, const float* xy, const float* yz, const float* xz, const float* dirs,
                       long long n, float* rgb, int mlp_impl, int num_sms, cudaStream_t st);

// packing kernels
cudaError_t launch_pack_plane(const float* nchw, int C, int H, int W, int DC, float* dens, __half* app,
                              cudaStream_t st);
cudaError_t launch_pack_gauge(const float* nchw, int H, int W, float2* out, cudaStream_t st);
cudaError_t launch_pack_occ(const float* vol, long long n_vox, uint32_t* bits, cudaStream_t st);
// raw bits -> "any corner set" brick grid + coarse grid + occupied index box (bbox[6] = min xyz, max xyz of occ2 cells)
cudaError_t launch_pack_occ2(const uint32_t* bits, int W, int H, int D, uint32_t* occ2, int nxb, int nyb, int nzb,
                             uint32_t* coarse, int cx, int cy, int cz, int* bbox, cudaStream_t st);
// TriPlane: dsum[t] = <dens[t][0..DC), w[0..DC)>
cudaError_t launch_pack_dsum(const float* dens, long long hw, int DC, const float* w_dev, float* dsum, cudaStream_t st);

cudaError_t launch_frame_post(const float* rgb, const float* gt, long long n, uint8_t* u8, double* sse, int num_sms,
                              cudaStream_t st);

cudaError_t launch_depth_colormap(const float* depth, long long n, float mi, float den, uint8_t* bgr, int num_sms,
                                  cudaStream_t st);

// ray sharding
cudaError_t launch_shard_gather(const float* src, long long n_rays, int width, int block, int rank, int world,
                                float* dst, cudaStream_t st);
cudaError_t launch_shard_scatter(const float* src, long long n_rays, int width, int block, int world,
                                 long long max_shard, float* dst, cudaStream_t st);

uint64_t launch_count();
void count_launch();

}  // namespace ngf
