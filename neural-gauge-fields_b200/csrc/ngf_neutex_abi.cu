// ngf_neutex_abi.cu — C ABI of the UV-Mapping (NeuTex) render path: weight packing, workspaces, render entry points.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/ngf_b200.h"
#include "ngf_neutex.cuh"

using namespace ngf::ntx;

int ngf_set_error(int code, const char* fmt, ...);   // ngf_abi.cu

#define CUN(expr)                                                                                              \
  do {                                                                                                         \
    cudaError_t _e = (expr);                                                                                   \
    if (_e != cudaSuccess) return ngf_set_error(NGF_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                                                __FILE__, __LINE__);                                           \
  } while (0)

struct NtxChunk {
  cudaStream_t stream = nullptr;
  float* raydir = nullptr;
  float* noise = nullptr;
  float* color = nullptr;
  float* trans = nullptr;
};

struct NgfNeutex_ {
  int device = 0;
  int num_sms = 0;
  NetDev net{};
  uint8_t* wstream = nullptr;                // weight stream(s) of the MLP kernel (one per CTA rank)
  float* heads = nullptr;
  float* texture = nullptr;
  // workspace for up to cap_rays rays
  long long cap_rays = 0;
  float4* work = nullptr;
  float4* sample_out = nullptr;
  unsigned long long* valid_mask = nullptr;
  unsigned int* counters = nullptr;          // [0] work count, [1] spare, [2..3] u64 valid-sample total
  float* cam_bg = nullptr;                   // [6]: campos, background (host path)
  float* rawbuf = nullptr;                   // unpacked fp32 parameters (fp32 fallback / self-check)
  RawNet raw{};
  int precision = 0;                         // 0: tcgen05 fp16 / split fp16 (ntx_mlp_kernel), 1: fp32 CUDA cores (ntx_ref_kernel)
  NtxChunk chunk[2];
  long long chunk_cap = 0;
  std::vector<cudaEvent_t> ev;               // 4 per timed render: raygen | mlp | march
  int ev_used = 0;
  unsigned long long last_valid = 0;
};

static void ntx_free_ws(NgfNeutex_* h) {
  cudaFree(h->work); cudaFree(h->sample_out); cudaFree(h->valid_mask);
  h->work = h->sample_out = nullptr; h->valid_mask = nullptr; h->cap_rays = 0;
}
static void ntx_free_chunks(NgfNeutex_* h) {
  for (auto& c : h->chunk) {
    if (c.stream) cudaStreamDestroy(c.stream);
    cudaFree(c.raydir); cudaFree(c.noise); cudaFree(c.color); cudaFree(c.trans);
    c = NtxChunk{};
  }
  h->chunk_cap = 0;
}
static void ntx_free_all(NgfNeutex_* h) {
  ntx_free_ws(h);
  ntx_free_chunks(h);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  cudaFree(h->wstream); cudaFree(h->heads); cudaFree(h->texture); cudaFree(h->counters); cudaFree(h->cam_bg);
  cudaFree(h->rawbuf);
}

struct Guard {
  int prev = -1;
  explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
  ~Guard() { int cur = -1; if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev); }
};

static int check_lin(const NgfLinear& l, int in_dim, int out_dim, const char* name, int idx) {
  if (!l.w || !l.b) return ngf_set_error(NGF_EINVAL, "%s[%d]: NULL weight or bias", name, idx);
  if (l.in_dim != in_dim || l.out_dim != out_dim)
    return ngf_set_error(NGF_EUNSUPPORTED, "%s[%d] is %dx%d, kernels are built for %dx%d", name, idx, l.out_dim, l.in_dim,
                         out_dim, in_dim);
  return NGF_OK;
}

static cudaError_t fetch(std::vector<float>& dst, const float* src, size_t n) {
  dst.resize(n);
  return cudaMemcpy(dst.data(), src, n * sizeof(float), cudaMemcpyDefault);
}

// Append one layer to the weight stream(s).  ranks = 1: one stream with all N output rows.  ranks = 2 (CTA pairs): rank r
// gets the rows [r*N/2, (r+1)*N/2).  Per K=16 step a rank's slice is [hi: 2 K-groups x rows x 8 halves][lo: same, split
// layers only] in tcgen05 K-major core-matrix order; W is [N][K_total] row-major.
static void pack_layer(std::vector<uint8_t>* out, int ranks, const std::vector<float>& W, int N, int K_total, bool split) {
  const int nk = K_total / 16, rows = N / ranks;
  const size_t slice = (size_t)rows * 32 * (split ? 2 : 1);
  for (int rank = 0; rank < ranks; ++rank) {
    const size_t base = out[rank].size();
    out[rank].resize(base + slice * nk, 0);
    for (int kk = 0; kk < nk; ++kk) {
      __half* hi = reinterpret_cast<__half*>(out[rank].data() + base + slice * kk);
      __half* lo = hi + (size_t)rows * 16;
      for (int kg = 0; kg < 2; ++kg)
        for (int r = 0; r < rows; ++r)
          for (int e = 0; e < 8; ++e) {
            const float w = W[(size_t)(rank * rows + r) * K_total + kk * 16 + kg * 8 + e];
            const __half h = __float2half_rn(w);
            hi[((size_t)kg * rows + r) * 8 + e] = h;
            if (split) lo[((size_t)kg * rows + r) * 8 + e] = __float2half_rn(w - __half2float(h));
          }
    }
  }
}

extern "C" {

int ngf_neutex_pack(const NgfNeutexDesc* d, int device, NgfNeutex* out) {
  if (!out) return ngf_set_error(NGF_EINVAL, "out is NULL");
  *out = nullptr;
  if (!d) return ngf_set_error(NGF_EINVAL, "desc is NULL");
  if (d->sample_num < 1 || d->sample_num > kS)
    return ngf_set_error(NGF_EUNSUPPORTED, "sample_num=%d (1..%d: the in-cube mask of a ray is a 64-bit word)", d->sample_num, kS);
  int rc;
  for (int i = 0; i < 12; ++i) {
    const int in_dim = i == 0 ? 63 : 256, out_dim = i == 11 ? 1 : 256;
    if ((rc = check_lin(d->geometry[i], in_dim, out_dim, "geometry", i))) return rc;
  }
  if (d->primitive != 0 && d->primitive != 1) return ngf_set_error(NGF_EINVAL, "primitive=%d (0 square, 1 sphere)", d->primitive);
  const int sphere = d->primitive;
  if (sphere && d->texture)
    return ngf_set_error(NGF_EUNSUPPORTED, "the edited-texture branch of the sphere primitive samples a cube map "
                         "(decoder.py:96-98); only the square primitive's texture swap is built");
  const int gi[5] = {63, 64, 128, 128, 128}, go[5] = {64, 128, 128, 128, sphere ? 3 : 2};
  for (int i = 0; i < 5; ++i)
    if ((rc = check_lin(d->gauge[i], gi[i], go[i], "gauge", i))) return rc;
  for (int i = 0; i < 6; ++i)
    if ((rc = check_lin(d->tex_block1[i], i == 0 ? (sphere ? 63 : 42) : 256, 256, "tex_block1", i))) return rc;
  if ((rc = check_lin(d->tex_color1, 256, 3, "tex_color1", 0))) return rc;
  for (int i = 0; i < 5; ++i)
    if ((rc = check_lin(d->tex_block2[i], i == 0 ? 295 : 256, i == 4 ? 3 : 256, "tex_block2", i))) return rc;
  if (d->texture && (d->tex_h < 1 || d->tex_w < 1 || d->tex_c < 1)) return ngf_set_error(NGF_EINVAL, "bad texture shape");
  int ndev = 0;
  CUN(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return ngf_set_error(NGF_EINVAL, "device %d out of range (%d visible)", device, ndev);
  Guard g(device);
  int major = 0;
  CUN(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return ngf_set_error(NGF_EUNSUPPORTED, "device %d has compute capability %d.x; sm_100a only", device, major);

  NgfNeutex_* h = new NgfNeutex_();
  h->device = device;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);

  int cg = 1;
  { const char* ce = getenv("NGF_NTX_CG"); if (ce && atoi(ce) == 2 && h->num_sms >= 2) cg = 2; }
  std::vector<uint8_t> wp[2];
  std::vector<float> heads(kHeadFloats, 0.f), W, B, Wb, rawv;
  std::vector<size_t> raw_off;
  std::vector<int> raw_in, raw_out;
  auto keep_raw = [&](const NgfLinear& l) {
    raw_off.push_back(rawv.size());
    raw_in.push_back(l.in_dim); raw_out.push_back(l.out_dim);
    rawv.insert(rawv.end(), W.begin(), W.begin() + (size_t)l.out_dim * l.in_dim);
    rawv.insert(rawv.end(), B.begin(), B.begin() + l.out_dim);
  };
  int li = 0;
  // The bias is accumulated by the tensor core: the view-direction operand carries two constant-one columns
  // (kOnesCol, kOnesCol + 1 = columns 7 and 8 of its third K=16 slice) and the weights hold (bias_hi, bias_lo) there.
  // Layers without view inputs get one extra K=16 slice for it; block2 layer 0 has it inside its 48 view columns.
  auto add = [&](const NgfLinear& l, int K, int Kext, bool split) -> int {
    if (fetch(W, l.w, (size_t)l.out_dim * l.in_dim) != cudaSuccess || fetch(B, l.b, l.out_dim) != cudaSuccess)
      return ngf_set_error(NGF_ECUDA, "cannot read layer %d parameters: %s", li, cudaGetErrorString(cudaGetLastError()));
    keep_raw(l);
    LayerDesc& L = h->net.layer[li++];
    const int N = l.out_dim;
    const bool merged = Kext > 0;
    const int K_total = merged ? K + Kext : K + 16;
    const int bias_col = merged ? K + 39 : K + 7;
    L.K = K; L.Kext = Kext; L.N = N; L.split = split ? 1 : 0; L.bias_slice = merged ? 0 : 1;
    const size_t al = kSliceBytes / cg;      // slices are powers of two <= al: none straddles a 16 KiB ring stage
    for (int r = 0; r < cg; ++r) wp[r].resize((wp[r].size() + al - 1) / al * al, 0);
    L.off = (uint32_t)wp[0].size();
    L.slice = (uint32_t)(N / cg) * 32u * (split ? 2u : 1u);
    Wb.assign((size_t)N * K_total, 0.f);
    for (int r = 0; r < N; ++r) {
      for (int k = 0; k < l.in_dim; ++k) Wb[(size_t)r * K_total + k] = W[(size_t)r * l.in_dim + k];
      const float bh = __half2float(__float2half_rn(B[r]));
      Wb[(size_t)r * K_total + bias_col] = bh;
      Wb[(size_t)r * K_total + bias_col + 1] = B[r] - bh;
    }
    pack_layer(wp, cg, Wb, N, K_total, split);
    return NGF_OK;
  };
  auto head = [&](const NgfLinear& l, int w_off, int b_off) -> int {
    if (fetch(W, l.w, (size_t)l.out_dim * l.in_dim) != cudaSuccess || fetch(B, l.b, l.out_dim) != cudaSuccess)
      return ngf_set_error(NGF_ECUDA, "cannot read head parameters: %s", cudaGetErrorString(cudaGetLastError()));
    keep_raw(l);
    memcpy(heads.data() + w_off, W.data(), W.size() * sizeof(float));
    memcpy(heads.data() + b_off, B.data(), B.size() * sizeof(float));
    return NGF_OK;
  };
  rc = NGF_OK;
  for (int i = 0; i < 11 && !rc; ++i) rc = add(d->geometry[i], i == 0 ? 64 : 256, 0, false);
  if (!rc) rc = head(d->geometry[11], kHeadGeo, kHeadGeoB);
  for (int i = 0; i < 4 && !rc; ++i) rc = add(d->gauge[i], i < 2 ? 64 : 128, 0, true);
  if (!rc) rc = head(d->gauge[4], kHeadGauge, kHeadGaugeB);
  for (int i = 0; i < 6 && !rc; ++i) rc = add(d->tex_block1[i], i == 0 ? (sphere ? 64 : 48) : 256, 0, false);
  if (!rc) rc = head(d->tex_color1, kHeadC1, kHeadC1B);
  for (int i = 0; i < 4 && !rc; ++i) rc = add(d->tex_block2[i], 256, i == 0 ? 48 : 0, false);
  if (!rc) rc = head(d->tex_block2[4], kHeadB2, kHeadB2B);
  if (rc) { ntx_free_all(h); delete h; return rc; }

  // which layers share a ring stage with their neighbours (the MMA issuer waits for / hands back a stage exactly once)
  for (int l = 0; l < kNumLayers; ++l) {
    LayerDesc& L = h->net.layer[l];
    const uint32_t nk = (uint32_t)((L.K + L.Kext) / 16 + L.bias_slice), end = L.off + nk * L.slice;
    L.first_have = 0;
    L.tail_release = 0;
    if (l > 0) {
      const LayerDesc& P = h->net.layer[l - 1];
      const uint32_t pend = P.off + (uint32_t)((P.K + P.Kext) / 16 + P.bias_slice) * P.slice;
      L.first_have = (pend % kStageBytes != 0 && pend / kStageBytes == L.off / kStageBytes) ? 1 : 0;
    }
    if (end % kStageBytes != 0) {
      const bool last = l == kNumLayers - 1;
      L.tail_release = (last || h->net.layer[l + 1].off / kStageBytes != end / kStageBytes) ? 1 : 0;
    }
  }
  // every rank's stream is padded to whole ring stages: the producer copies stage after stage, unit after unit
  for (int r = 0; r < cg; ++r) wp[r].resize((wp[r].size() + kStageBytes - 1) / kStageBytes * kStageBytes, 0);
  h->net.stream_bytes = (uint32_t)wp[0].size();
  h->net.cg = cg;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&h->wstream), (size_t)cg * wp[0].size());
  for (int r = 0; r < cg && e == cudaSuccess; ++r)
    e = cudaMemcpy(h->wstream + (size_t)r * wp[0].size(), wp[r].data(), wp[r].size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->heads), heads.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(h->heads, heads.data(), heads.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->counters), 64);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->cam_bg), 6 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->rawbuf), rawv.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(h->rawbuf, rawv.data(), rawv.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && raw_off.size() == 29) {
    for (int l = 0; l < 29; ++l) {
      h->raw.w[l] = h->rawbuf + raw_off[l];
      h->raw.b[l] = h->raw.w[l] + (size_t)raw_in[l] * raw_out[l];
      h->raw.in[l] = raw_in[l]; h->raw.out[l] = raw_out[l];
    }
  }
  { const char* pe = getenv("NGF_NTX_FP32"); h->precision = pe && pe[0] == '1' ? 1 : 0; }
  if (e == cudaSuccess && d->texture) {
    const size_t n = (size_t)d->tex_h * d->tex_w * d->tex_c;
    e = cudaMalloc(reinterpret_cast<void**>(&h->texture), n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(h->texture, d->texture, n * sizeof(float), cudaMemcpyDefault);
  }
  if (e != cudaSuccess) {
    ntx_free_all(h); delete h;
    return ngf_set_error(NGF_ECUDA, "ngf_neutex_pack: %s", cudaGetErrorString(e));
  }
  h->net.wstream = h->wstream; h->net.heads = h->heads;
  h->net.texture = h->texture; h->net.tex_h = d->tex_h; h->net.tex_w = d->tex_w; h->net.tex_c = d->tex_c;
  h->net.jitter = d->jitter;
  h->net.S = d->sample_num;
  h->net.dt = (float)(2.0 / d->sample_num);
  h->net.dj = (float)((2.0 / d->sample_num) * (double)d->jitter);
  h->net.sphere = sphere;
  { const char* e = getenv("NGF_NTX_DBG"); h->net.dbg = e ? atoi(e) : 0; }
  h->net.trace = nullptr;
  if (h->net.dbg & 4) {
    cudaMalloc(reinterpret_cast<void**>(&h->net.trace), kTraceWords * sizeof(long long));
    cudaMemset(h->net.trace, 0, kTraceWords * sizeof(long long));
  }
  *out = h;
  return NGF_OK;
}

void ngf_neutex_free(NgfNeutex h) {
  if (!h) return;
  Guard g(h->device);
  cudaDeviceSynchronize();
  ntx_free_all(h);
  delete h;
}

static int ntx_ensure_ws(NgfNeutex_* h, long long n_rays, cudaStream_t st) {
  if (h->cap_rays >= n_rays) return NGF_OK;
  CUN(cudaStreamSynchronize(st));
  ntx_free_ws(h);
  const size_t n = (size_t)n_rays * kS;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&h->work), n * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->sample_out), n * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->valid_mask), (size_t)n_rays * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    cudaGetLastError();
    ntx_free_ws(h);
    return ngf_set_error(NGF_ENOMEM, "NeuTex workspace for %lld rays: %s", n_rays, cudaGetErrorString(e));
  }
  h->cap_rays = n_rays;
  return NGF_OK;
}

// rays per internal batch: 2 x 16 B x 64 per ray of workspace -> 1 Mi rays = 2 GiB
static const long long kNtxBatch = 1ll << 20;

// where the jitter numbers of a render come from: caller's array | drawn on the device from (seed, frame ray index) | none
struct NoiseSrc {
  const float* ptr = nullptr;
  bool seeded = false;
  unsigned long long seed = 0;
  long long ray0 = 0;
};

static int ntx_render_dev(NgfNeutex_* h, const float* campos, const float* raydir, const float* background,
                          const NoiseSrc& nz, long long n_rays, float* color, float* trans, cudaStream_t st) {
  const long long per = n_rays < kNtxBatch ? n_rays : kNtxBatch;
  int rc = ntx_ensure_ws(h, per, st);
  if (rc) return rc;
  for (long long s0 = 0; s0 < n_rays; s0 += per) {
    const long long n = (n_rays - s0) < per ? (n_rays - s0) : per;
    RenderArgsN a{};
    a.campos = campos; a.raydir = raydir + s0 * 3; a.background = background;
    a.noise = nz.ptr ? nz.ptr + s0 * h->net.S : nullptr;
    a.seeded = nz.seeded ? 1 : 0; a.seed = nz.seed; a.ray0 = nz.ray0 + s0;
    a.n_rays = n;
    a.work = h->work; a.counters = h->counters; a.valid_mask = h->valid_mask; a.sample_out = h->sample_out;
    a.color = color + s0 * 3; a.transmittance = trans + s0;
    CUN(cudaMemsetAsync(h->counters, 0, 8, st));
    const bool timed = h->ev_used + 4 <= (int)h->ev.size();
    if (timed) CUN(cudaEventRecord(h->ev[h->ev_used], st));
    CUN(launch_neutex_raygen(h->net, a, st));
    if (timed) CUN(cudaEventRecord(h->ev[h->ev_used + 1], st));
    if (h->precision == 1) CUN(launch_neutex_ref(h->net, h->raw, a, st));
    else CUN(launch_neutex_mlp(h->net, a, h->num_sms, st));
    if (timed) CUN(cudaEventRecord(h->ev[h->ev_used + 2], st));
    CUN(launch_neutex_march(h->net, a, st));
    if (timed) {
      CUN(cudaEventRecord(h->ev[h->ev_used + 3], st));
      h->ev_used += 4;
    }
  }
  return NGF_OK;
}

static int ntx_render_checked(NgfNeutex h, const float* campos_dev, const float* raydir_dev, const float* background_dev,
                              const NoiseSrc& nz, int64_t n_rays, float* color_dev, float* transmittance_dev, void* stream) {
  if (!h) return ngf_set_error(NGF_EINVAL, "handle is NULL");
  if (n_rays < 0 || n_rays > (1ll << 31) / kS * 16) return ngf_set_error(NGF_EINVAL, "n_rays=%lld", (long long)n_rays);
  if (n_rays == 0) return NGF_OK;
  if (!campos_dev || !raydir_dev || !color_dev || !transmittance_dev) return ngf_set_error(NGF_EINVAL, "NULL pointer");
  Guard g(h->device);
  return ntx_render_dev(h, campos_dev, raydir_dev, background_dev, nz, n_rays, color_dev, transmittance_dev,
                        reinterpret_cast<cudaStream_t>(stream));
}

int ngf_neutex_render(NgfNeutex h, const float* campos_dev, const float* raydir_dev, const float* background_dev,
                      const float* noise_dev, int64_t n_rays, float* color_dev, float* transmittance_dev, void* stream) {
  NoiseSrc nz;
  nz.ptr = noise_dev;
  return ntx_render_checked(h, campos_dev, raydir_dev, background_dev, nz, n_rays, color_dev, transmittance_dev, stream);
}

int ngf_neutex_render_seeded(NgfNeutex h, const float* campos_dev, const float* raydir_dev, const float* background_dev,
                             uint64_t seed, int64_t first_ray, int64_t n_rays, float* color_dev, float* transmittance_dev,
                             void* stream) {
  if (first_ray < 0) return ngf_set_error(NGF_EINVAL, "first_ray=%lld", (long long)first_ray);
  NoiseSrc nz;
  nz.seeded = true; nz.seed = seed; nz.ray0 = first_ray;
  return ntx_render_checked(h, campos_dev, raydir_dev, background_dev, nz, n_rays, color_dev, transmittance_dev, stream);
}

int ngf_neutex_noise(NgfNeutex h, uint64_t seed, int64_t first_ray, int64_t n_rays, float* noise_dev, void* stream) {
  if (!h || !noise_dev) return ngf_set_error(NGF_EINVAL, "NULL argument");
  if (first_ray < 0 || n_rays < 0) return ngf_set_error(NGF_EINVAL, "first_ray=%lld n_rays=%lld", (long long)first_ray, (long long)n_rays);
  Guard g(h->device);
  CUN(launch_neutex_noise(seed, first_ray, n_rays, h->net.S, noise_dev, reinterpret_cast<cudaStream_t>(stream)));
  return NGF_OK;
}

static int ntx_render_host_impl(NgfNeutex h, const float* campos_host, const float* raydir_host, const float* background_host,
                                const float* noise_host, bool seeded, uint64_t seed, int64_t n_rays, float* color_host,
                                float* transmittance_host) {
  if (!h) return ngf_set_error(NGF_EINVAL, "handle is NULL");
  if (n_rays < 0) return ngf_set_error(NGF_EINVAL, "n_rays=%lld", (long long)n_rays);
  if (n_rays == 0) return NGF_OK;
  if (!campos_host || !raydir_host || !color_host || !transmittance_host) return ngf_set_error(NGF_EINVAL, "NULL pointer");
  Guard g(h->device);
  const long long chunk = n_rays < 65536 ? n_rays : 65536;
  if (h->chunk_cap < chunk) {
    ntx_free_chunks(h);
    for (auto& c : h->chunk) {
      CUN(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
      CUN(cudaMalloc(reinterpret_cast<void**>(&c.raydir), (size_t)chunk * 3 * sizeof(float)));
      CUN(cudaMalloc(reinterpret_cast<void**>(&c.noise), (size_t)chunk * kS * sizeof(float)));
      CUN(cudaMalloc(reinterpret_cast<void**>(&c.color), (size_t)chunk * 3 * sizeof(float)));
      CUN(cudaMalloc(reinterpret_cast<void**>(&c.trans), (size_t)chunk * sizeof(float)));
    }
    h->chunk_cap = chunk;
  }
  float cb[6] = {campos_host[0], campos_host[1], campos_host[2], 0.f, 0.f, 0.f};
  if (background_host) { cb[3] = background_host[0]; cb[4] = background_host[1]; cb[5] = background_host[2]; }
  CUN(cudaMemcpy(h->cam_bg, cb, sizeof(cb), cudaMemcpyHostToDevice));
  // the MLP workspace is shared, so chunks run back to back on one stream; copies of chunk i+1 overlap on the other
  cudaStream_t run = h->chunk[0].stream, copy = h->chunk[1].stream;
  cudaEvent_t up[2], done[2];
  for (int i = 0; i < 2; ++i) { CUN(cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming)); CUN(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming)); }
  int rc = NGF_OK, ci = 0;
  for (long long s = 0; s < n_rays && rc == NGF_OK; s += chunk, ci ^= 1) {
    const long long n = (n_rays - s) < chunk ? (n_rays - s) : chunk;
    NtxChunk& c = h->chunk[ci];
    cudaStreamWaitEvent(copy, done[ci], 0);               // buffers of this slot free again
    cudaMemcpyAsync(c.raydir, raydir_host + s * 3, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, copy);
    if (noise_host) cudaMemcpyAsync(c.noise, noise_host + s * h->net.S, (size_t)n * h->net.S * sizeof(float), cudaMemcpyHostToDevice, copy);
    cudaEventRecord(up[ci], copy);
    cudaStreamWaitEvent(run, up[ci], 0);
    NoiseSrc nz;
    nz.ptr = noise_host ? c.noise : nullptr;
    nz.seeded = seeded; nz.seed = seed; nz.ray0 = s;
    rc = ntx_render_dev(h, h->cam_bg, c.raydir, background_host ? h->cam_bg + 3 : nullptr, nz, n, c.color, c.trans, run);
    cudaMemcpyAsync(color_host + s * 3, c.color, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, run);
    cudaMemcpyAsync(transmittance_host + s, c.trans, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, run);
    cudaEventRecord(done[ci], run);
  }
  cudaError_t e1 = cudaStreamSynchronize(run), e2 = cudaStreamSynchronize(copy);
  for (int i = 0; i < 2; ++i) { cudaEventDestroy(up[i]); cudaEventDestroy(done[i]); }
  if (rc) return rc;
  CUN(e1);
  CUN(e2);
  return NGF_OK;
}

int ngf_neutex_render_host(NgfNeutex h, const float* campos_host, const float* raydir_host, const float* background_host,
                           const float* noise_host, int64_t n_rays, float* color_host, float* transmittance_host) {
  return ntx_render_host_impl(h, campos_host, raydir_host, background_host, noise_host, false, 0, n_rays, color_host,
                              transmittance_host);
}

int ngf_neutex_render_host_seeded(NgfNeutex h, const float* campos_host, const float* raydir_host, const float* background_host,
                                  uint64_t seed, int64_t n_rays, float* color_host, float* transmittance_host) {
  return ntx_render_host_impl(h, campos_host, raydir_host, background_host, nullptr, true, seed, n_rays, color_host,
                              transmittance_host);
}

int ngf_neutex_set_precision(NgfNeutex h, int32_t mode) {
  if (!h) return ngf_set_error(NGF_EINVAL, "handle is NULL");
  if (mode != 0 && mode != 1) return ngf_set_error(NGF_EINVAL, "mode=%d (0 tensor cores, 1 fp32)", mode);
  h->precision = mode;
  return NGF_OK;
}

// Evaluate n_points random in-cube points (and unit view directions) through the tensor-core kernel and through the fp32
// reference kernel and report how far they are apart.
int ngf_neutex_self_check(NgfNeutex h, int32_t n_points, uint64_t seed, float* report, void* stream) {
  if (!h || !report) return ngf_set_error(NGF_EINVAL, "NULL argument");
  if (n_points < 1 || n_points > (1 << 20)) return ngf_set_error(NGF_EINVAL, "n_points=%d", n_points);
  Guard g(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = ntx_ensure_ws(h, n_points, st);
  if (rc) return rc;
  std::vector<float4> work((size_t)n_points);
  std::vector<float> dirs((size_t)n_points * 3);
  uint64_t x = seed * 6364136223846793005ull + 1442695040888963407ull;
  auto uni = [&]() { x = x * 6364136223846793005ull + 1442695040888963407ull; return (float)((x >> 40) & 0xFFFFFF) / 16777216.f; };
  for (int i = 0; i < n_points; ++i) {
    const int id = i * kS;
    float4 w;
    memcpy(&w.x, &id, 4);
    w.y = uni() * 1.98f - 0.99f; w.z = uni() * 1.98f - 0.99f; w.w = uni() * 1.98f - 0.99f;
    work[i] = w;
    float d[3] = {uni() * 2.f - 1.f, uni() * 2.f - 1.f, uni() * 2.f - 1.f};
    const float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) + 1e-5f;
    for (int k = 0; k < 3; ++k) dirs[(size_t)i * 3 + k] = d[k] / n;
  }
  float* dirs_dev = nullptr;
  CUN(cudaMalloc(reinterpret_cast<void**>(&dirs_dev), dirs.size() * sizeof(float)));
  cudaError_t e = cudaMemcpyAsync(dirs_dev, dirs.data(), dirs.size() * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->work, work.data(), work.size() * sizeof(float4), cudaMemcpyHostToDevice, st);
  const unsigned int cnt[2] = {(unsigned)n_points, 0u};
  std::vector<float4> out_tc((size_t)n_points), out_ref((size_t)n_points);
  RenderArgsN a{};
  a.raydir = dirs_dev; a.n_rays = n_points;
  a.work = h->work; a.counters = h->counters; a.valid_mask = h->valid_mask; a.sample_out = h->sample_out;
  for (int pass = 0; pass < 2 && e == cudaSuccess; ++pass) {
    e = cudaMemcpyAsync(h->counters, cnt, sizeof(cnt), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = pass == 0 ? launch_neutex_mlp(h->net, a, h->num_sms, st) : launch_neutex_ref(h->net, h->raw, a, st);
    // sample_out is indexed by sample id = point * 64
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(pass == 0 ? out_tc.data() : out_ref.data(), sizeof(float4), h->sample_out, sizeof(float4) * kS,
                            sizeof(float4), (size_t)n_points, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  cudaFree(dirs_dev);
  if (e != cudaSuccess) { cudaGetLastError(); return ngf_set_error(NGF_ECUDA, "ngf_neutex_self_check: %s", cudaGetErrorString(e)); }
  double max_sig = 0.0, max_rgb = 0.0, mean_rgb = 0.0, max_ref_rgb = 0.0;
  for (int i = 0; i < n_points; ++i) {
    const float4 t = out_tc[i], r = out_ref[i];
    max_sig = fmax(max_sig, fabs((double)t.x - r.x) / (1.0 + fabs((double)r.x)));
    const double d = fmax(fabs((double)t.y - r.y), fmax(fabs((double)t.z - r.z), fabs((double)t.w - r.w)));
    // fmax drops NaNs: a non-finite tensor-core result must read as a failure, not as agreement
    const bool finite = std::isfinite(t.x) && std::isfinite(t.y) && std::isfinite(t.z) && std::isfinite(t.w);
    max_rgb = finite ? fmax(max_rgb, d) : INFINITY;
    if (!finite) max_sig = INFINITY;
    mean_rgb += finite ? d : 0.0;
    max_ref_rgb = fmax(max_ref_rgb, fmax(fabs((double)r.y), fmax(fabs((double)r.z), fabs((double)r.w))));
  }
  report[0] = (float)max_sig; report[1] = (float)max_rgb; report[2] = (float)(mean_rgb / n_points); report[3] = (float)max_ref_rgb;
  return NGF_OK;
}

int ngf_neutex_last_valid_samples(NgfNeutex h, uint64_t* n_valid, void* stream) {
  if (!h || !n_valid) return ngf_set_error(NGF_EINVAL, "NULL argument");
  Guard g(h->device);
  unsigned int c = 0;
  CUN(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  CUN(cudaMemcpy(&c, h->counters, sizeof(c), cudaMemcpyDeviceToHost));
  *n_valid = c;
  return NGF_OK;
}

int ngf_neutex_copy_samples(NgfNeutex h, int64_t first_sample, int64_t n, float* sigma_rgb_host, uint64_t* valid_mask_host,
                            int64_t first_ray, int64_t n_mask_rays) {
  if (!h) return ngf_set_error(NGF_EINVAL, "handle is NULL");
  Guard g(h->device);
  CUN(cudaDeviceSynchronize());
  if (first_sample < 0 || n < 0 || first_sample + n > h->cap_rays * kS || first_ray < 0 || n_mask_rays < 0 ||
      first_ray + n_mask_rays > h->cap_rays)
    return ngf_set_error(NGF_EINVAL, "range outside the last render's workspace");
  if (sigma_rgb_host && n) CUN(cudaMemcpy(sigma_rgb_host, h->sample_out + first_sample, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost));
  if (valid_mask_host && n_mask_rays)
    CUN(cudaMemcpy(valid_mask_host, h->valid_mask + first_ray, (size_t)n_mask_rays * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return NGF_OK;
}

int ngf_neutex_debug_trace(NgfNeutex h, long long* out_host) {     /* NGF_NTX_DBG=4 only: 25 x 4 + 64 clock64 stamps */
  if (!h || !out_host) return ngf_set_error(NGF_EINVAL, "NULL argument");
  if (!h->net.trace) return ngf_set_error(NGF_EINVAL, "tracing is off (NGF_NTX_DBG=4)");
  Guard g(h->device);
  CUN(cudaDeviceSynchronize());
  CUN(cudaMemcpy(out_host, h->net.trace, kTraceWords * sizeof(long long), cudaMemcpyDeviceToHost));
  return NGF_OK;
}

int ngf_neutex_timing_begin(NgfNeutex h, int32_t capacity) {
  if (!h) return ngf_set_error(NGF_EINVAL, "handle is NULL");
  if (capacity < 0 || capacity > 65536) return ngf_set_error(NGF_EINVAL, "capacity=%d", capacity);
  Guard g(h->device);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  h->ev.assign((size_t)capacity * 4, nullptr);
  h->ev_used = 0;
  for (auto& e : h->ev) CUN(cudaEventCreate(&e));
  return NGF_OK;
}

int ngf_neutex_timing_read(NgfNeutex h, int32_t* n_renders, double* raygen_ms, double* mlp_ms, double* march_ms) {
  if (!h || !n_renders || !raygen_ms || !mlp_ms || !march_ms) return ngf_set_error(NGF_EINVAL, "NULL argument");
  Guard g(h->device);
  double s[3] = {0, 0, 0};
  for (int i = 0; i + 3 < h->ev_used; i += 4) {
    CUN(cudaEventSynchronize(h->ev[i + 3]));
    for (int k = 0; k < 3; ++k) {
      float ms = 0.f;
      CUN(cudaEventElapsedTime(&ms, h->ev[i + k], h->ev[i + k + 1]));
      s[k] += ms;
    }
  }
  *n_renders = h->ev_used / 4;
  *raygen_ms = s[0]; *mlp_ms = s[1]; *march_ms = s[2];
  h->ev_used = 0;
  return NGF_OK;
}

}  // extern "C"
