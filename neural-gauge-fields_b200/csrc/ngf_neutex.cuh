// ngf_neutex.cuh — shared declarations of the UV-Mapping (NeuTex) render path (ngf_neutex.cu / ngf_neutex_abi.cu).
//
// Reference (paths relative to /root/reference/UV-Mapping): model/model.py:27-59 NeuTex.forward wires
//   cube_ray_generation (model/renderer.py:79-141) -> GeometryMlpDecoder (model/decoder.py:201-237)
//   -> GaugeTransform (model/gauge_fields.py:8-74) -> TextureMlpDecoder (model/decoder.py:11-121)
//   -> ray_march / alpha_blend / simple_tone_map (model/renderer.py:4-11,176-247).
// The reference pushes all R*64 samples through all three MLP stacks; samples outside the unit cube have their
// density multiplied by 0 (renderer.py:222), so they cannot reach the image.  Here the ray-generation kernel compacts
// the in-cube samples into a work list and only those are evaluated.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ngf {
namespace ntx {

constexpr int kS = 64;                 // samples per ray (dtu_test.sh: --sample_num 64); the kernels are built for it
constexpr int kRows = 256;             // work items per MLP tile: two M=128 tcgen05 tiles sharing every weight chunk
constexpr int kWorkerThreads = 512;    // two threads per tile row (TMEM lane): each owns half of the layer's columns
constexpr int kWorkerWarps = kWorkerThreads / 32;
constexpr int kThreads = kWorkerThreads + 64;   // + 1 MMA-issue warp + 1 weight-producer warp
constexpr int kStages = 4;             // weight ring
constexpr int kSlicesPerStage = 2;     // K=16 weight slices per ring stage: one wait / copy / commit per stage
constexpr uint32_t kSliceBytes = 8192; // one K=16 slice of a 256-wide layer (or hi+lo slices of a <=128-wide one)
constexpr uint32_t kStageBytes = kSlicesPerStage * kSliceBytes;
constexpr int kNumLayers = 25;         // MMA layers per tile: geometry 11, gauge 4, texture block1 6, block2 4

struct LayerDesc {
  int K;                // K of the main A operand (multiple of 16)
  int Kext;             // extra K taken from the view-direction operand (block2 layer 0: 48), else 0
  int N;                // output width (64 / 128 / 256)
  int split;            // 1: hi/lo split-fp16 operands, 3 MMAs per K step (gauge network)
  int bias_slice;       // 1: one more K=16 slice whose A operand is the constant-one columns of the view operand and
                        //    whose weights are (bias_hi, bias_lo); 0: the bias rides in the Kext slices (block2 layer 0)
  uint32_t w_off;       // byte offset of the layer's first weight chunk in the packed weight stream
  uint32_t chunk_bytes; // bytes per K=16 chunk (N*32, doubled when split)
};

struct NetDev {
  LayerDesc layer[kNumLayers];
  const uint8_t* wpack;   // all weight chunks, tcgen05 K-major core-matrix order, in execution order
  uint32_t wpack_stride;  // optional replicas of the stream (NGF_NTX_COPIES, default 1): CTA b reads copy b % w_copies
  int w_copies;
  const float* heads;     // fp32 head weights (16-byte aligned rows first, then the biases), offsets kHead* below
  const float* texture;   // [h][w][c] edited texture or nullptr
  int tex_h, tex_w, tex_c;
  float jitter;
  int dbg;                // NGF_NTX_DBG (profiling experiments only): 2 = skip MMA issue, 4 = record a timeline
  long long* trace;       // dbg & 4: [25 layers][4] clock64 stamps of CTA 0's first tile (a_ready seen, MMAs issued,
                          //          acc_ready seen by worker 0, epilogue done by worker 0)
};

// geometry head [256] | gauge head [2][128] | color1 [3][256] | block2 head [3][256] | biases 1 + 2 + 3 + 3
constexpr int kHeadGeo = 0, kHeadGauge = 256, kHeadC1 = 512, kHeadB2 = 1280, kHeadGeoB = 2048, kHeadGaugeB = 2049,
              kHeadC1B = 2051, kHeadB2B = 2054, kHeadFloats = 2060;

struct RenderArgsN {
  const float* campos;     // [3]
  const float* raydir;     // [R][3]
  const float* background; // [3] or nullptr
  const float* noise;      // [R][64] or nullptr (no jitter)
  long long n_rays;
  float4* work;            // [R*64] compacted in-cube samples: (bit-cast sample id, x, y, z)
  unsigned int* counters;  // [0] work count, [1] tile counter
  unsigned long long* valid_mask;   // [R] bit i = sample i is inside the cube
  float4* sample_out;      // [R*64] (sigma, r, g, b) of evaluated samples
  float* color;            // [R][3]
  float* transmittance;    // [R]
};

cudaError_t launch_neutex_raygen(const NetDev& net, const RenderArgsN& a, cudaStream_t st);
cudaError_t launch_neutex_mlp(const NetDev& net, const RenderArgsN& a, int num_sms, cudaStream_t st);
cudaError_t launch_neutex_march(const NetDev& net, const RenderArgsN& a, cudaStream_t st);

}  // namespace ntx
}  // namespace ngf
