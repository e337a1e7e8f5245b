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

constexpr int kS = 64;                 // sample slots per ray in the workspaces (the in-cube mask is 64 bits): sample_num <= 64
constexpr int kRows = 256;             // work items per MLP tile: two M=128 tcgen05 tiles sharing every weight chunk
constexpr int kWorkerThreads = 512;    // two threads per tile row (TMEM lane): each owns half of the layer's columns
constexpr int kWorkerWarps = kWorkerThreads / 32;
constexpr int kThreads = kWorkerThreads + 64;   // + 1 MMA-issue warp + 1 weight-producer warp
#ifndef NTX_STAGES           // experiment knobs (scripts/gpu_stage_variants.sh); the defaults are what ships
#define NTX_STAGES 4
#endif
#ifndef NTX_STAGE_BYTES
#define NTX_STAGE_BYTES 16384
#endif
constexpr int kStages = NTX_STAGES;    // weight ring
constexpr uint32_t kSliceBytes = 8192; // largest K=16 slice: a 256-wide layer (or hi+lo of a 128-wide one) at cg = 1
constexpr uint32_t kStageBytes = NTX_STAGE_BYTES; // one cp.async.bulk per stage: two 8 KiB slices (cg = 1), four 4 KiB (cg = 2)
static_assert((kStageBytes & (kStageBytes - 1)) == 0 && kStageBytes >= kSliceBytes, "ring stage size");
constexpr int kTraceWords = 25 * 4 + 64;   // NGF_NTX_DBG=4: per-layer stamps + ring stamps of layer kTraceLayer
constexpr int kTraceLayer = 5;
constexpr int kNumLayers = 25;         // MMA layers per tile: geometry 11, gauge 4, texture block1 6, block2 4

// geometry head [256] | gauge head [2 or 3][128] | color1 [3][256] | block2 head [3][256] | biases 1 + 3 + 3 + 3
constexpr int kHeadGeo = 0, kHeadGauge = 256, kHeadC1 = 768, kHeadB2 = 1536, kHeadGeoB = 2304, kHeadGaugeB = 2305,
              kHeadC1B = 2308, kHeadB2B = 2311, kHeadFloats = 2316;

struct LayerDesc {
  int K;                // K of the main A operand (multiple of 16)
  int Kext;             // extra K taken from the view-direction operand (block2 layer 0: 48), else 0
  int N;                // output width (64 / 128 / 256)
  int split;            // 1: hi/lo split-fp16 operands, 3 MMAs per K step (gauge network)
  int bias_slice;       // 1: one more K=16 slice whose A operand is the constant-one columns of the view operand and
                        //    whose weights are (bias_hi, bias_lo); 0: the bias rides in the Kext slices (block2 layer 0)
  uint32_t off;         // byte offset of the layer's first K=16 slice inside one rank's weight stream (8 KiB aligned)
  uint32_t slice;       // bytes per K=16 slice of one rank: (N / cg) * 32, doubled when split
  int first_have;       // 1: the layer's first slice lies in the ring stage the previous layer ended in (already waited for)
  int tail_release;     // 1: the layer's last (partly used) stage holds nothing of the next layer: hand it back
};

struct NetDev {
  LayerDesc layer[kNumLayers];
  // Weight stream(s), tcgen05 K-major core-matrix order, layers in execution order, padded to whole ring stages.
  // cg = 1: one stream with all N output rows of every layer.  cg = 2: [rank 0 stream][rank 1 stream]; rank r holds the
  // output rows [r*N/2, (r+1)*N/2) of every layer (its half of the B operand of the pair's M=256 MMAs).
  const uint8_t* wstream;
  uint32_t stream_bytes;  // bytes of one rank's stream
  int cg;                 // 1: one CTA per 256-sample tile (cta_group::1); 2: CTA pairs (cta_group::2) (NGF_NTX_CG)
  const float* heads;     // fp32 head weights (16-byte aligned rows first, then the biases), offsets kHead* above
  const float* texture;   // [h][w][c] edited texture or nullptr
  int tex_h, tex_w, tex_c;
  float jitter;
  int S;                  // opt.sample_num (1..64; dtu_test.sh: 64)
  float dt, dj;           // fp32(2 / S) and fp32((2 / S) * jitter): the Python doubles of renderer.py:111-118 as torch casts them
  int sphere;             // 1: primitive_type 'sphere' (gauge_fields.py:55-56,71-74): 3 gauge outputs, uv = normalize(.)
  int dbg;                // NGF_NTX_DBG (profiling experiments only): 2 = skip MMA issue, 4 = record a timeline,
                          //          8 = skip the weight copies (ring stages are released without data)
  long long* trace;       // dbg & 4: [25 layers][4] clock64 stamps of CTA 0's first tile (a_ready seen, MMAs issued,
                          //          acc_ready seen by worker 0, epilogue done by worker 0)
};

struct RenderArgsN {
  const float* campos;     // [3]
  const float* raydir;     // [R][3]
  const float* background; // [3] or nullptr
  const float* noise;      // [R][S] caller-supplied U[0,1) numbers, or nullptr
  // noise == nullptr && seeded: the numbers are drawn in the kernels, jitter_uniform(seed, (ray0 + ray) * 64 + i);
  // noise == nullptr && !seeded: no jitter (every U = 0.5)
  int seeded;
  unsigned long long seed;
  long long ray0;          // index of this launch's first ray within the frame (batches and host chunks draw the same numbers)
  long long n_rays;
  float4* work;            // [R*64] compacted in-cube samples: (bit-cast sample id, x, y, z)
  unsigned int* counters;  // [0] work count, [1] tile counter
  unsigned long long* valid_mask;   // [R] bit i = sample i is inside the cube
  float4* sample_out;      // [R*64] (sigma, r, g, b) of evaluated samples
  float* color;            // [R][3]
  float* transmittance;    // [R]
};

// fp32 CUDA-core evaluation of the three networks from the unpacked parameters (fallback / self-check): raw = every layer's
// W [out][in] then b [out], in NgfNeutexDesc order (geometry 12, gauge 5, tex_block1 6, tex_color1, tex_block2 5)
struct RawNet {
  const float* w[29];
  const float* b[29];
  int in[29], out[29];
};
// Fill noise_out[n_rays][S] with the numbers a seeded render draws (tests feed them to the oracle).
cudaError_t launch_neutex_noise(unsigned long long seed, long long ray0, long long n_rays, int S, float* noise_out, cudaStream_t st);
cudaError_t launch_neutex_ref(const NetDev& net, const RawNet& raw, const RenderArgsN& a, cudaStream_t st);

cudaError_t launch_neutex_raygen(const NetDev& net, const RenderArgsN& a, cudaStream_t st);
cudaError_t launch_neutex_mlp(const NetDev& net, const RenderArgsN& a, int num_sms, cudaStream_t st);
cudaError_t launch_neutex_march(const NetDev& net, const RenderArgsN& a, cudaStream_t st);

}  // namespace ntx
}  // namespace ngf
