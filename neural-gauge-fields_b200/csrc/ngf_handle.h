// ngf_handle.h — the field handle behind NgfField (shared by ngf_abi.cu and ngf_comm.cu).  Host code only.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <vector>

#include "../../include/ngf_b200.h"
#include "ngf_internal.h"

// ---------------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------------
struct HostChunk {            // one in-flight chunk of the host-buffer render path
  cudaStream_t stream = nullptr;      // compute stream (slots 0..kHostComp-1 own one; the others borrow slot % kHostComp)
  cudaEvent_t ev_in = nullptr, ev_comp = nullptr, ev_out = nullptr;   // upload done | kernels done | download done
  float* rays = nullptr;
  float* rgb = nullptr;
  float* depth = nullptr;
  float* acc = nullptr;
  uint8_t* u8 = nullptr;              // uint8 image of the chunk (ngf_field_render_camera_u8_host_async)
  unsigned int* counters = nullptr;
  ngf::QEntry* queue = nullptr;
  long long queue_cap = 0;
};

struct NgfField_ {
  int device = 0;
  int num_sms = 0;
  int has_gauge = 0;
  ngf::FieldDev dev{};
  int n_samples_default = 0;
  int plane_c = 0;
  // owned device memory
  float* dens[3] = {nullptr, nullptr, nullptr};
  __half* app[3] = {nullptr, nullptr, nullptr};
  float2* gauge[3] = {nullptr, nullptr, nullptr};
  uint32_t* occ = nullptr;
  uint32_t* occ2 = nullptr;
  uint32_t* occ_coarse = nullptr;
  float* dsum[3] = {nullptr, nullptr, nullptr};
  float* dmlp = nullptr;
  __half* w1p = nullptr;
  __half* w2p = nullptr;
  float* tail = nullptr;
  // fp32 network weights in nn.Linear layout for the backward pass (ngf_train.cu)
  __half* ii_w = nullptr;             // InfoInv: split fp16 density MLP (ngf_infoinv_tc.cuh)
  float* ii_tail = nullptr;
  void* tmaps = nullptr;              // 3 CUtensorMap objects in device memory (TriPlane appearance planes)
  float* raw_w = nullptr;             // [basis F*F | mlp.0 W 64*(F+15) | b 64 | mlp.2 W 64*64 | b 64 | mlp.4 W 3*64 | b 3]
  float* raw_dw = nullptr;            // InfoInv density MLP [W1 32*72 | b1 32 | W2 32*32 | b2 32 | W3 32 | b3 1]
  void* train = nullptr;              // TrainWs of ngf_train.cu (record list + activation workspace)
  // render workspace
  float* acc_ws = nullptr;
  long long acc_cap = 0;
  unsigned int* counters = nullptr;   // [0] tile counter, [1] queue count, [2..9] = 4 x u64 stats
  // ngf_field_render / _jitter / _camera keep one workspace (colour queue, counters, acc scratch) per caller stream, so
  // renders issued on different streams of one handle are independent and may overlap (the march of one frame beside the
  // colour pass of another).  Slot 0 is the handle's own workspace above; a slot whose stream has not been seen for the
  // longest is recycled (after synchronising that stream) when more than kStreamSlots streams are used.
  struct StreamWs {
    cudaStream_t stream = nullptr;
    bool used = false;
    unsigned long long last_use = 0;
    unsigned int* counters = nullptr;
    ngf::QEntry* queue = nullptr;
    long long queue_cap = 0;
    float* acc_ws = nullptr;
    long long acc_cap = 0;
  } sws[4];
  unsigned long long sws_clock = 0;
  const unsigned int* stats_counters = nullptr;   // counters of the most recent device render (ngf_field_stats)
  ngf::QEntry* queue = nullptr;            // colour work items of the device-resident path
  long long queue_cap = 0;
  // kernel timing (ngf_field_timing_*)
  std::vector<cudaEvent_t> ev;        // 3 per timed march+colour pair
  int ev_used = 0;
  // host path: kHostSlots chunks in flight, each on its own stream; whole-frame pipelines are replayed as CUDA graphs
  HostChunk chunk[6];
  cudaStream_t s_in = nullptr, s_out = nullptr;       // dedicated upload / download streams
  int next_slot = 0;                                  // round robin over the chunk slots, across frames
  long long chunk_cap = 0;
  int chunk_stride = 0;
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  struct HostGraph {
    const void *rays, *rgb, *depth;
    long long n_rays;
    int stride, n_samples, white_bg, tile_w, impl;
    unsigned long long epoch;
    cudaGraphExec_t exec;            // nullptr: capture failed once, stay eager for this key
  };
  std::vector<HostGraph> graphs;
  unsigned long long epoch = 0;       // bumped whenever anything a captured kernel argument depends on changes
  cudaEvent_t frame_done[8] = {};     // completion of the last 8 asynchronous host frames (ticket % 8)
  unsigned long long next_ticket = 1;
};
constexpr int kHostSlots = 6;    // chunk buffers in flight
constexpr int kHostComp = 3;     // compute streams

constexpr int kCounterBytes = 64;

namespace ngf {
// Where the finished rows of a ray-sharded render go (ngf_comm.cu): local ray l is global ray
// ((l / block) * world + rank) * block + l % block of the batch; (r, g, b, depth) is written as one float4 row into
// every dst[] (dst[0] = this rank's frame buffer; further entries are peer mappings of the other ranks' buffers).
struct ShardOut {
  float4* dst[16];
  int n_dst;
  int block, rank, world;
};
cudaError_t launch_finalize_shard(const float* rgb, const float* acc, const float* depth, long long n_local,
                                  int white_bg, const ShardOut& so, cudaStream_t st);
}  // namespace ngf

// ngf_train.cu
void ngf_train_free(void* train_ws);

// ngf_abi.cu
int ngf_set_error(int code, const char* fmt, ...);
int ngf_render_dev(NgfField h, const float* rays, long long n_rays, int ray_stride, int n_samples, int white_bg,
                   int tile_w, float* rgb, float* depth, float* acc, unsigned int* counters, ngf::QEntry** queue,
                   long long* queue_cap, int mlp_impl, cudaStream_t st, const ngf::CamDev* cam = nullptr,
                   const float* jitter = nullptr, const ngf::ShardOut* shard = nullptr);
