// ngf_abi.cu — the C ABI of libngf_b200.so (include/ngf_b200.h): handle management, parameter packing, render
// entry points.  Host code only; kernels live in ngf_kernels.cu.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ngf_handle.h"

using namespace ngf;

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// shared with ngf_neutex_abi.cu
int ngf_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(expr)                                                                                          \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) return fail(NGF_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                                       __LINE__);                                                         \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};


static void drop_graphs(NgfField_* h) {
  for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
}

static void free_chunks(NgfField_* h) {
  drop_graphs(h);
  if (h->ev_fork) { cudaEventDestroy(h->ev_fork); h->ev_fork = nullptr; }
  for (auto& e : h->ev_join) if (e) { cudaEventDestroy(e); e = nullptr; }
  for (auto& e : h->frame_done) if (e) { cudaEventDestroy(e); e = nullptr; }
  for (int i = 0; i < kHostSlots; ++i) {
    HostChunk& c = h->chunk[i];
    if (c.stream && i < kHostComp) cudaStreamDestroy(c.stream);
    if (c.ev_in) cudaEventDestroy(c.ev_in);
    if (c.ev_comp) cudaEventDestroy(c.ev_comp);
    if (c.ev_out) cudaEventDestroy(c.ev_out);
    cudaFree(c.rays); cudaFree(c.rgb); cudaFree(c.depth); cudaFree(c.acc); cudaFree(c.u8); cudaFree(c.counters); cudaFree(c.queue);
    c = HostChunk{};
  }
  if (h->s_in) { cudaStreamDestroy(h->s_in); h->s_in = nullptr; }
  if (h->s_out) { cudaStreamDestroy(h->s_out); h->s_out = nullptr; }
  h->chunk_cap = 0;
}

static void free_events(NgfField_* h) {
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  h->ev.clear();
  h->ev_used = 0;
}

static void free_all(NgfField_* h) {
  free_events(h);
  for (int i = 0; i < 3; ++i) { cudaFree(h->dens[i]); cudaFree(h->app[i]); cudaFree(h->gauge[i]); }
  for (int i = 0; i < 3; ++i) cudaFree(h->dsum[i]);
  cudaFree(h->occ); cudaFree(h->occ2); cudaFree(h->occ_coarse);
  cudaFree(h->dmlp); cudaFree(h->w1p); cudaFree(h->w2p); cudaFree(h->tail);
  cudaFree(h->raw_w); cudaFree(h->raw_dw); cudaFree(h->tmaps); cudaFree(h->ii_w); cudaFree(h->ii_tail);
  ngf_train_free(h->train);
  h->train = nullptr;
  cudaFree(h->acc_ws); cudaFree(h->counters); cudaFree(h->queue);
  for (auto& w : h->sws) { cudaFree(w.counters); cudaFree(w.queue); cudaFree(w.acc_ws); }
  free_chunks(h);
}

// ---------------------------------------------------------------------------------------------------------
// validation
// ---------------------------------------------------------------------------------------------------------
static int check_linear(const NgfLinear& l, int in_dim, int out_dim, bool need_bias, const char* name) {
  if (!l.w) return fail(NGF_EINVAL, "%s.w is NULL", name);
  if (need_bias && !l.b) return fail(NGF_EINVAL, "%s.b is NULL", name);
  if (l.in_dim != in_dim || l.out_dim != out_dim)
    return fail(NGF_EUNSUPPORTED, "%s is %dx%d, kernels are built for %dx%d", name, l.out_dim, l.in_dim, out_dim, in_dim);
  return NGF_OK;
}

static int validate(const NgfFieldDesc* d) {
  if (!d) return fail(NGF_EINVAL, "desc is NULL");
  if (d->variant != NGF_TRIPLANE && d->variant != NGF_INFOINV) return fail(NGF_EINVAL, "unknown variant %d", d->variant);
  const int C = d->variant == NGF_TRIPLANE ? 64 : 96, DC = d->variant == NGF_TRIPLANE ? 16 : 24;
  if (d->plane_c != C || d->density_c != DC)
    return fail(NGF_EUNSUPPORTED, "plane_c=%d density_c=%d; variant %d is built for %d/%d", d->plane_c, d->density_c,
                d->variant, C, DC);
  for (int i = 0; i < 3; ++i) {
    if (!d->plane[i]) return fail(NGF_EINVAL, "plane[%d] is NULL", i);
    if (d->plane_h[i] < 1 || d->plane_w[i] < 1 || d->plane_h[i] > 32768 || d->plane_w[i] > 32768)
      return fail(NGF_EINVAL, "plane[%d] has shape %dx%d", i, d->plane_h[i], d->plane_w[i]);
    if (d->variant == NGF_TRIPLANE && d->gauge_on && !d->gauge[i])
      return fail(NGF_EINVAL, "gauge_on but gauge[%d] is NULL", i);
    if (d->variant == NGF_TRIPLANE && d->gauge[i] && (d->gauge_h[i] < 1 || d->gauge_w[i] < 1))
      return fail(NGF_EINVAL, "gauge[%d] has shape %dx%d", i, d->gauge_h[i], d->gauge_w[i]);
  }
  const int F = 3 * (C - DC);
  if (d->view_pe != 2) return fail(NGF_EUNSUPPORTED, "view_pe=%d (built for 2)", d->view_pe);
  int rc;
  if ((rc = check_linear(d->rgb_basis, F, F, false, "rgb_basis"))) return rc;
  if ((rc = check_linear(d->rgb_l1, F + 15, 64, true, "rgb_l1"))) return rc;
  if ((rc = check_linear(d->rgb_l2, 64, 64, true, "rgb_l2"))) return rc;
  if ((rc = check_linear(d->rgb_l3, 64, 3, true, "rgb_l3"))) return rc;
  if (d->variant == NGF_TRIPLANE) {
    if ((rc = check_linear(d->dens_l1, 48, 1, true, "dens_l1"))) return rc;
  } else {
    if ((rc = check_linear(d->dens_l1, 72, 32, true, "dens_l1"))) return rc;
    if ((rc = check_linear(d->dens_l2, 32, 32, true, "dens_l2"))) return rc;
    if ((rc = check_linear(d->dens_l3, 32, 1, true, "dens_l3"))) return rc;
  }
  if (!(d->step_size > 0.f)) return fail(NGF_EINVAL, "step_size must be > 0");
  for (int k = 0; k < 3; ++k)
    if (!(d->aabb[3 + k] > d->aabb[k])) return fail(NGF_EINVAL, "aabb is empty along axis %d", k);
  if (d->alpha_volume) {
    long long n = 1;
    for (int k = 0; k < 3; ++k) {
      if (d->alpha_dims[k] < 1 || d->alpha_dims[k] > 2048) return fail(NGF_EINVAL, "alpha_dims[%d]=%d", k, d->alpha_dims[k]);
      n *= d->alpha_dims[k];
    }
    if (n > (1ll << 31)) return fail(NGF_EUNSUPPORTED, "alpha volume larger than 2^31 voxels");
  }
  return NGF_OK;
}

// ---------------------------------------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t dev_alloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)); }

static cudaError_t fetch(std::vector<float>& dst, const float* src, size_t n) {
  dst.resize(n);
  return cudaMemcpy(dst.data(), src, n * sizeof(float), cudaMemcpyDefault);
}

// K-major core-matrix order used by the tcgen05 descriptors: element (row r, col k) of an [rows][K] operand at
// (k/8)*rows*8 + r*8 + (k%8)   (in halves)
static void put_kmajor(std::vector<__half>& dst, int rows, int r, int k, float v) {
  dst[(size_t)(k / 8) * rows * 8 + (size_t)r * 8 + (k % 8)] = __float2half_rn(v);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
  }
  return fn;
}

// Tensor maps over the channels-last fp16 appearance planes [H][W][48] with boxes of 48 channels x kBox x kBox texels
// (ngf_colour_tma.cuh).  Failure is not an error: the colour kernel then keeps the direct gather.
static void build_tensor_maps(NgfField_* h) {
  h->dev.tmap = nullptr;
  if (h->dev.variant != 0) return;
  // opt-in: on the 800x800 / 256^2-plane workload only ~36 % of the (group, plane) taps fit a 5x5 patch (86 % would fit
  // 8x8, which does not fit shared memory twice per SM), and the kernel is slower than the direct gather
  const char* e = getenv("NGF_COLOUR_TMA");
  if (!(e && e[0] == '1')) return;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return;
  alignas(64) CUtensorMap maps[3];
  for (int i = 0; i < 3; ++i) {
    const PlaneDev& P = h->dev.plane[i];
    const cuuint64_t dims[3] = {48, (cuuint64_t)P.W, (cuuint64_t)P.H};
    const cuuint64_t strides[2] = {96, (cuuint64_t)P.W * 96};
    const cuuint32_t box[3] = {48, 5, 5};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(P.app), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return;
  }
  if (!h->tmaps && cudaMalloc(&h->tmaps, sizeof(maps)) != cudaSuccess) { cudaGetLastError(); h->tmaps = nullptr; return; }
  if (cudaMemcpy(h->tmaps, maps, sizeof(maps), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); return; }
  h->dev.tmap = h->tmaps;
}

static int pack_params(NgfField_* h, const NgfFieldDesc* d, bool allocate) {
  const int V = d->variant;
  const int C = d->plane_c, DC = d->density_c, AC = C - DC, F = 3 * AC;
  const int K1 = V == 0 ? Cfg<0>::K1 : Cfg<1>::K1;
  FieldDev& f = h->dev;
  f.variant = V;
  f.infoinv = d->infoinv ? 1 : 0;
  const bool has_gauge = V == 0 && d->gauge[0] && d->gauge[1] && d->gauge[2];
  h->has_gauge = has_gauge ? 1 : 0;
  f.gauge_on = (has_gauge && d->gauge_on) ? 1 : 0;
  for (int k = 0; k < 3; ++k) {
    f.lo[k] = d->aabb[k];
    f.hi[k] = d->aabb[3 + k];
    f.inv[k] = d->inv_aabb_size[k];
  }
  f.step = d->step_size; f.near_t = d->near_t; f.far_t = d->far_t;
  f.dscale = d->distance_scale; f.wthres = d->weight_thres; f.dshift = d->density_shift;
  // early-out of the march: once T <= tstop every later weight is <= tstop <= weight_thres, so no later sample can be
  // colour-active and acc / depth lose at most tstop; a threshold <= 0 keeps every sample, as the reference does
  f.tstop = d->weight_thres >= 1e-6f ? 1e-6f : (d->weight_thres > 0.f ? d->weight_thres : 0.f);
  h->n_samples_default = d->n_samples;
  h->plane_c = C;

  // ---- planes
  for (int i = 0; i < 3; ++i) {
    const int H = d->plane_h[i], W = d->plane_w[i];
    const size_t hw = (size_t)H * W;
    if (allocate) {
      CU(dev_alloc(&h->dens[i], hw * DC));
      CU(dev_alloc(&h->app[i], hw * AC));
      if (V == 0) CU(dev_alloc(&h->dsum[i], hw));
    } else if (H != f.plane[i].H || W != f.plane[i].W) {
      return fail(NGF_EINVAL, "repack: plane[%d] changed shape (%dx%d -> %dx%d); pack a new handle", i, f.plane[i].H,
                  f.plane[i].W, H, W);
    }
    float* tmp = nullptr;
    CU(dev_alloc(&tmp, hw * C));
    cudaError_t e = cudaMemcpy(tmp, d->plane[i], hw * C * sizeof(float), cudaMemcpyDefault);
    if (e == cudaSuccess) e = launch_pack_plane(tmp, C, H, W, DC, h->dens[i], h->app[i], 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(tmp);
    CU(e);
    f.plane[i] = PlaneDev{h->dens[i], h->dsum[i], h->app[i], H, W, (float)(W - 1), (float)(H - 1)};
  }
  // ---- gauge planes
  for (int i = 0; i < 3; ++i) {
    if (!has_gauge) { if (allocate) f.gauge[i] = GaugeDev{nullptr, 1, 1, 0.f, 0.f}; continue; }
    const int H = d->gauge_h[i], W = d->gauge_w[i];
    const size_t hw = (size_t)H * W;
    if (allocate || !h->gauge[i]) {
      if (h->gauge[i]) { cudaFree(h->gauge[i]); h->gauge[i] = nullptr; }
      CU(dev_alloc(&h->gauge[i], hw));
    } else if (H != f.gauge[i].H || W != f.gauge[i].W) {
      return fail(NGF_EINVAL, "repack: gauge[%d] changed shape", i);
    }
    float* tmp = nullptr;
    CU(dev_alloc(&tmp, hw * 2));
    cudaError_t e = cudaMemcpy(tmp, d->gauge[i], hw * 2 * sizeof(float), cudaMemcpyDefault);
    if (e == cudaSuccess) e = launch_pack_gauge(tmp, H, W, h->gauge[i], 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(tmp);
    CU(e);
    f.gauge[i] = GaugeDev{h->gauge[i], H, W, (float)(W - 1), (float)(H - 1)};
  }
  // ---- occupancy grid: raw bits, "any corner" brick grid, coarse grid, occupied box
  cudaFree(h->occ); cudaFree(h->occ2); cudaFree(h->occ_coarse);
  h->occ = h->occ2 = h->occ_coarse = nullptr;
  if (d->alpha_volume) {
    const int W = d->alpha_dims[0], H = d->alpha_dims[1], D = d->alpha_dims[2];
    const long long n = (long long)W * H * D;
    CU(dev_alloc(&h->occ, (size_t)((n + 31) / 32)));
    float* tmp = nullptr;
    CU(dev_alloc(&tmp, (size_t)n));
    cudaError_t e = cudaMemcpy(tmp, d->alpha_volume, (size_t)n * sizeof(float), cudaMemcpyDefault);
    if (e == cudaSuccess) e = launch_pack_occ(tmp, n, h->occ, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(tmp);
    CU(e);
    const int nxb = (W + 1 + 3) / 4, nyb = (H + 1 + 3) / 4, nzb = (D + 1 + 1) / 2;
    const int cx = (W + 1 + 7) / 8, cy = (H + 1 + 7) / 8, cz = (D + 1 + 7) / 8;
    const size_t n2 = (size_t)nxb * nyb * nzb, nc = ((size_t)cx * cy * cz + 31) / 32;
    CU(dev_alloc(&h->occ2, n2));
    CU(dev_alloc(&h->occ_coarse, nc));
    CU(cudaMemset(h->occ2, 0, n2 * sizeof(uint32_t)));
    CU(cudaMemset(h->occ_coarse, 0, nc * sizeof(uint32_t)));
    int* bbox_dev = nullptr;
    CU(dev_alloc(&bbox_dev, 6));
    int bbox[6] = {1 << 30, 1 << 30, 1 << 30, -1, -1, -1};
    e = cudaMemcpy(bbox_dev, bbox, sizeof(bbox), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_pack_occ2(h->occ, W, H, D, h->occ2, nxb, nyb, nzb, h->occ_coarse, cx, cy, cz, bbox_dev, 0);
    if (e == cudaSuccess) e = cudaMemcpy(bbox, bbox_dev, sizeof(bbox), cudaMemcpyDeviceToHost);
    cudaFree(bbox_dev);
    CU(e);
    f.occ = h->occ; f.occ2 = h->occ2; f.occ_coarse = h->occ_coarse;
    f.occ_w = W; f.occ_h = H; f.occ_d = D;
    f.occ2_nxb = nxb; f.occ2_nyb = nyb; f.occ_cx = cx; f.occ_cy = cy;
    f.has_occ = 1;
    const int dims[3] = {W, H, D};
    for (int k = 0; k < 3; ++k) {
      f.occ_lo[k] = d->alpha_aabb[k];
      f.occ_inv[k] = d->alpha_inv[k];
      // A kept sample has floor(fi) + 1 in [bbox_min, bbox_max], i.e. fi in [bbox_min - 1, bbox_max); half a cell of
      // slack covers the rounding of the fi chain.  World x = lo + fi / (dim - 1) * size.
      const double size = (double)d->alpha_aabb[3 + k] - (double)d->alpha_aabb[k];
      const double cell = dims[k] > 1 ? size / (dims[k] - 1) : size;
      if (bbox[3 + k] < 0) {                 // empty mask: nothing can be kept
        f.clip_lo[k] = 1.f; f.clip_hi[k] = -1.f;
      } else if (dims[k] > 1) {
        f.clip_lo[k] = (float)(d->alpha_aabb[k] + (bbox[k] - 1.5) * cell);
        f.clip_hi[k] = (float)(d->alpha_aabb[k] + (bbox[3 + k] + 0.5) * cell);
      } else {
        f.clip_lo[k] = -INFINITY; f.clip_hi[k] = INFINITY;
      }
    }
  } else {
    f.occ = f.occ2 = f.occ_coarse = nullptr; f.has_occ = 0; f.occ_w = f.occ_h = f.occ_d = 1;
    f.occ2_nxb = f.occ2_nyb = f.occ_cx = f.occ_cy = 1;
    for (int k = 0; k < 3; ++k) { f.occ_lo[k] = 0.f; f.occ_inv[k] = 0.f; f.clip_lo[k] = -INFINITY; f.clip_hi[k] = INFINITY; }
  }
  // ---- density head
  std::vector<float> w, b;
  if (V == 0) {
    CU(fetch(w, d->dens_l1.w, 48));
    CU(fetch(b, d->dens_l1.b, 1));
    memcpy(f.dw, w.data(), 48 * sizeof(float));
    f.db = b[0];
    f.dmlp = nullptr;
    f.ii_w = nullptr; f.ii_tail = nullptr;
    float* w_dev = nullptr;
    CU(dev_alloc(&w_dev, (size_t)48));
    cudaError_t e = cudaMemcpy(w_dev, w.data(), 48 * sizeof(float), cudaMemcpyHostToDevice);
    for (int i = 0; i < 3 && e == cudaSuccess; ++i)
      e = launch_pack_dsum(h->dens[i], (long long)f.plane[i].H * f.plane[i].W, 16, w_dev + 16 * i, h->dsum[i], 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(w_dev);
    CU(e);
  } else {
    std::vector<float> m(kDmlpFloats, 0.f), w2, b2, w3, b3;
    CU(fetch(w, d->dens_l1.w, 32 * 72)); CU(fetch(b, d->dens_l1.b, 32));
    CU(fetch(w2, d->dens_l2.w, 32 * 32)); CU(fetch(b2, d->dens_l2.b, 32));
    CU(fetch(w3, d->dens_l3.w, 32)); CU(fetch(b3, d->dens_l3.b, 1));
    for (int j = 0; j < 32; ++j)
      for (int k = 0; k < 72; ++k) m[(size_t)k * 32 + j] = w[(size_t)j * 72 + k];        // input-major
    float* p = m.data() + 32 * 72;
    memcpy(p, b.data(), 32 * 4); p += 32;
    memcpy(p, w2.data(), 32 * 32 * 4); p += 32 * 32;
    memcpy(p, b2.data(), 32 * 4); p += 32;
    memcpy(p, w3.data(), 32 * 4); p += 32;
    p[0] = b3[0];
    if (allocate) CU(dev_alloc(&h->dmlp, (size_t)kDmlpFloats));
    CU(cudaMemcpy(h->dmlp, m.data(), kDmlpFloats * sizeof(float), cudaMemcpyHostToDevice));
    {
      std::vector<float> raw;                       // nn.Linear layout, for the backward pass
      raw.insert(raw.end(), w.begin(), w.end()); raw.insert(raw.end(), b.begin(), b.end());
      raw.insert(raw.end(), w2.begin(), w2.end()); raw.insert(raw.end(), b2.begin(), b2.end());
      raw.insert(raw.end(), w3.begin(), w3.end()); raw.insert(raw.end(), b3.begin(), b3.end());
      if (allocate) CU(dev_alloc(&h->raw_dw, raw.size()));
      CU(cudaMemcpy(h->raw_dw, raw.data(), raw.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    {
      // split fp16 operands of the tensor-core density MLP: x = hi + lo, hi = rn16(x), lo = rn16(x - hi); the biases ride
      // in a constant-one K column (layer 1: column 72 of 80, layer 2: column 32 of 48)
      std::vector<__half> iw((size_t)2 * 32 * 80 + (size_t)2 * 32 * 48, __float2half_rn(0.f));
      auto put = [&](size_t base_hi, size_t base_lo, int r, int k, float v) {
        const __half hi = __float2half_rn(v);
        const size_t o = (size_t)(k / 8) * 32 * 8 + (size_t)r * 8 + (k % 8);
        iw[base_hi + o] = hi;
        iw[base_lo + o] = __float2half_rn(v - __half2float(hi));
      };
      const size_t o1h = 0, o1l = 32 * 80, o2h = 2 * 32 * 80, o2l = 2 * 32 * 80 + 32 * 48;
      for (int r = 0; r < 32; ++r) {
        for (int k = 0; k < 72; ++k) put(o1h, o1l, r, k, w[(size_t)r * 72 + k]);
        put(o1h, o1l, r, 72, b[r]);
        for (int k = 0; k < 32; ++k) put(o2h, o2l, r, k, w2[(size_t)r * 32 + k]);
        put(o2h, o2l, r, 32, b2[r]);
      }
      std::vector<float> tl(36, 0.f);
      memcpy(tl.data(), w3.data(), 32 * 4);
      tl[32] = b3[0];
      if (allocate) { CU(dev_alloc(&h->ii_w, iw.size())); CU(dev_alloc(&h->ii_tail, tl.size())); }
      CU(cudaMemcpy(h->ii_w, iw.data(), iw.size() * sizeof(__half), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(h->ii_tail, tl.data(), tl.size() * sizeof(float), cudaMemcpyHostToDevice));
      f.ii_w = h->ii_w; f.ii_tail = h->ii_tail;
    }
    f.dmlp = h->dmlp;
    memset(f.dw, 0, sizeof(f.dw));
    f.db = 0.f;
  }
  // ---- colour MLP: fold the bias-free basis into layer 1 (W1' = W1[:, :F] . B), append view columns and b1
  {
    std::vector<float> B, W1, b1, W2, b2, W3, b3;
    CU(fetch(B, d->rgb_basis.w, (size_t)F * F));
    CU(fetch(W1, d->rgb_l1.w, (size_t)64 * (F + 15))); CU(fetch(b1, d->rgb_l1.b, 64));
    CU(fetch(W2, d->rgb_l2.w, 64 * 64)); CU(fetch(b2, d->rgb_l2.b, 64));
    CU(fetch(W3, d->rgb_l3.w, 3 * 64)); CU(fetch(b3, d->rgb_l3.b, 3));
    std::vector<__half> w1p((size_t)K1 * 64, __float2half_rn(0.f)), w2p((size_t)64 * 64);
    std::vector<double> row(F);
    for (int j = 0; j < 64; ++j) {
      std::fill(row.begin(), row.end(), 0.0);
      for (int m = 0; m < F; ++m) {
        const double wjm = W1[(size_t)j * (F + 15) + m];
        const float* brow = &B[(size_t)m * F];
        for (int k = 0; k < F; ++k) row[k] += wjm * (double)brow[k];
      }
      for (int k = 0; k < F; ++k) put_kmajor(w1p, 64, j, k, (float)row[k]);
      for (int k = 0; k < 15; ++k) put_kmajor(w1p, 64, j, F + k, W1[(size_t)j * (F + 15) + F + k]);
      put_kmajor(w1p, 64, j, F + 15, b1[j]);
      for (int k = 0; k < 64; ++k) put_kmajor(w2p, 64, j, k, W2[(size_t)j * 64 + k]);
    }
    std::vector<float> tail(kTailFloats, 0.f);
    for (int c = 0; c < 64; ++c) {          // one float4 per hidden unit: (b2, w3[0], w3[1], w3[2])
      tail[4 * c] = b2[c];
      for (int o = 0; o < 3; ++o) tail[4 * c + 1 + o] = W3[(size_t)o * 64 + c];
    }
    memcpy(tail.data() + 256, b3.data(), 3 * 4);
    if (allocate) {
      CU(dev_alloc(&h->w1p, w1p.size()));
      CU(dev_alloc(&h->w2p, w2p.size()));
      CU(dev_alloc(&h->tail, tail.size()));
    }
    CU(cudaMemcpy(h->w1p, w1p.data(), w1p.size() * sizeof(__half), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->w2p, w2p.data(), w2p.size() * sizeof(__half), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->tail, tail.data(), tail.size() * sizeof(float), cudaMemcpyHostToDevice));
    f.w1p = h->w1p; f.w2p = h->w2p; f.tail = h->tail;
    memcpy(f.tail_c, tail.data(), sizeof(f.tail_c));
    {
      std::vector<float> raw;                       // unfolded fp32 weights, for the backward pass
      raw.insert(raw.end(), B.begin(), B.end()); raw.insert(raw.end(), W1.begin(), W1.end());
      raw.insert(raw.end(), b1.begin(), b1.end()); raw.insert(raw.end(), W2.begin(), W2.end());
      raw.insert(raw.end(), b2.begin(), b2.end()); raw.insert(raw.end(), W3.begin(), W3.end());
      raw.insert(raw.end(), b3.begin(), b3.end());
      if (allocate) CU(dev_alloc(&h->raw_w, raw.size()));
      CU(cudaMemcpy(h->raw_w, raw.data(), raw.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
  }
  if (allocate) {
    CU(cudaMalloc(reinterpret_cast<void**>(&h->counters), kCounterBytes));
    CU(cudaMemset(h->counters, 0, kCounterBytes));
  }
  build_tensor_maps(h);
  CU(cudaDeviceSynchronize());
  return NGF_OK;
}

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int ngf_abi_version(void) { return NGF_ABI_VERSION; }
const char* ngf_last_error(void) { return g_err; }
uint64_t ngf_launch_count(void) { return launch_count(); }

int ngf_field_pack(const NgfFieldDesc* desc, int device, NgfField* out) {
  if (!out) return fail(NGF_EINVAL, "out is NULL");
  *out = nullptr;
  int rc = validate(desc);
  if (rc) return rc;
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(NGF_EINVAL, "device %d out of range (%d visible)", device, ndev);
  DeviceGuard g(device);
  if (!g.ok) return fail(NGF_ECUDA, "cannot select device %d", device);
  int major = 0;
  CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(NGF_EUNSUPPORTED, "device %d has compute capability %d.x; this library is sm_100a only", device, major);
  NgfField_* h = new NgfField_();
  h->device = device;
  CU(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  rc = pack_params(h, desc, true);
  if (rc) { free_all(h); delete h; return rc; }
  *out = h;
  return NGF_OK;
}

int ngf_field_repack(NgfField h, const NgfFieldDesc* desc) {
  if (!h) return fail(NGF_EINVAL, "field is NULL");
  int rc = validate(desc);
  if (rc) return rc;
  if (desc->variant != h->dev.variant) return fail(NGF_EINVAL, "repack: variant changed");
  DeviceGuard g(h->device);
  if (!g.ok) return fail(NGF_ECUDA, "cannot select device %d", h->device);
  ++h->epoch;
  drop_graphs(h);
  return pack_params(h, desc, false);
}

void ngf_field_free(NgfField h) {
  if (!h) return;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  free_all(h);
  delete h;
}

// Colour-queue budget: the march kernel may emit up to n_rays * S work items of 32 bytes; a render is split into
// ray batches so that the worst case fits the workspace (4 GiB by default: one 800x800x192 frame is one batch).
static long long queue_budget_items() {
  static long long v = 0;
  if (v == 0) {
    const char* e = getenv("NGF_QUEUE_MIB");
    long long mib = e ? atoll(e) : 4096;
    if (mib < 1) mib = 1;
    v = mib * (1ll << 20) / (long long)sizeof(QEntry);
  }
  return v;
}

static int ensure_queue(QEntry** q, long long* cap, long long want, cudaStream_t st) {
  if (*cap >= want) return NGF_OK;
  CU(cudaStreamSynchronize(st));
  cudaFree(*q);
  *q = nullptr; *cap = 0;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(q), (size_t)want * sizeof(QEntry));
  if (e != cudaSuccess) { cudaGetLastError(); return fail(NGF_ENOMEM, "colour queue of %lld MiB: %s", want * (long long)sizeof(QEntry) >> 20, cudaGetErrorString(e)); }
  *cap = want;
  return NGF_OK;
}

}  // extern "C"

int ngf_render_dev(NgfField h, const float* rays, long long n_rays, int ray_stride, int n_samples, int white_bg,
                   int tile_w, float* rgb, float* depth, float* acc, unsigned int* counters, QEntry** queue,
                   long long* queue_cap, int mlp_impl, cudaStream_t st, const CamDev* cam, const float* jitter,
                   const ShardOut* shard) {
  if (n_rays == 0) return NGF_OK;
  const int S = n_samples > 0 ? n_samples : h->n_samples_default;
  if (S < 1) return fail(NGF_EINVAL, "n_samples resolves to %d", S);
  const bool img = tile_w > 0 && n_rays % tile_w == 0;
  // rays per batch: worst-case queue use within budget; whole 4-row groups when image-tiled, whole warps otherwise
  long long per = queue_budget_items() / S;
  const long long unit = img ? 4ll * tile_w : 32;
  per = per / unit * unit;
  if (per < unit) per = unit;
  if (per > n_rays) per = n_rays;
  long long want = per * S;
  if (want > 0xfffffff0ll) { per = 0xfffffff0ll / S / unit * unit; want = per * S; }
  // InfoInv: the three-phase march keeps 192 bytes of state per ray (+ 64) behind the queue, in the same allocation
  // InfoInv, opt-in (NGF_INFOINV_TC=1): the march as find / tensor-core density / composite (ngf_infoinv_tc.cuh), with one
  // 32-byte record per kept sample (worst case n * S, like the queue) + 4 bytes per ray behind the queue, in the same
  // allocation: queue and records share the budget.  Parity-green, but it evaluates every kept sample of a ray (13.2 per ray
  // on the hull field against 3.3 with the in-march early-out) and the fp32 density gather costs as much per sample as the
  // colour gather: 1.92 ms per frame against 0.77.  Default: the in-march per-lane MLP.
  static int ii_tc = -1;
  if (ii_tc < 0) { const char* e = getenv("NGF_INFOINV_TC"); ii_tc = e && e[0] == '1' ? 1 : 0; }
  const bool use_tc = h->dev.variant == 1 && ii_tc && !getenv("NGF_INFOINV_PHASED");
  if (use_tc) {
    per = per / 2 / unit * unit;
    if (per < unit) per = unit;
    if (per > n_rays) per = n_rays;
    want = per * S;
  }
  static int ii_phased = -1;
  // opt-in (NGF_INFOINV_PHASED=1): parity-green but slower on B200 than the in-march MLP — 10 rounds x 4 grid barriers of
  // ~12 us each and a per-lane MLP that is latency-bound even at full lane occupancy (profiles/r02_infoinv_phased_trace.txt)
  if (ii_phased < 0) { const char* e = getenv("NGF_INFOINV_PHASED"); ii_phased = e && e[0] == '1' ? 1 : 0; }
  const long long ii_items = use_tc ? per * S + per / 8 + 4                               // in 32-byte queue items
                                    : (h->dev.variant == 1 && ii_phased) ? per * 6 + 2 : 0;
  int rc = ensure_queue(queue, queue_cap, want + ii_items, st);
  if (rc) return rc;
  for (long long s0 = 0; s0 < n_rays; s0 += per) {
    const long long n = (n_rays - s0) < per ? (n_rays - s0) : per;
    RenderArgs a{};
    a.rays = rays ? rays + s0 * ray_stride : nullptr; a.n_rays = n; a.ray_stride = ray_stride;
    a.jitter = jitter ? jitter + s0 : nullptr;
    a.cam_on = cam ? 1 : 0;
    if (cam) { a.cam = *cam; a.cam.base = cam->base + s0; }
    a.S = S;
    a.white_bg = white_bg ? 1 : 0;
    if (img) {
      a.img_w = tile_w; a.img_h = (int)(n / tile_w);
      a.n_tiles = ((a.img_w + 7) / 8) * ((a.img_h + 3) / 4);
    } else {
      a.img_w = 0; a.img_h = 0;
      a.n_tiles = (int)((n + 31) / 32);
    }
    a.rgb = rgb + s0 * 3; a.depth = depth + s0; a.acc = acc + s0;
    a.tile_counter = counters;
    a.queue_count = counters + 1;
    a.queue = *queue;
    a.queue_cap = (unsigned int)(n * S < want ? n * S : want);
    a.ii_ws = ii_items ? static_cast<void*>(*queue + want) : nullptr;
    a.ii_tc = use_tc ? 1 : 0;
    a.stats = reinterpret_cast<unsigned long long*>(counters + 2);
    CU(cudaMemsetAsync(counters, 0, s0 == 0 ? kCounterBytes : 8, st));   // first batch also clears the statistics
    const bool timed = h->ev_used + 3 <= (int)h->ev.size();
    if (timed) CU(cudaEventRecord(h->ev[h->ev_used], st));
    CU(launch_march(h->dev, a, h->num_sms, st));
    if (timed) CU(cudaEventRecord(h->ev[h->ev_used + 1], st));
    CU(launch_colour(h->dev, a, mlp_impl, h->num_sms, st));
    if (timed) {
      CU(cudaEventRecord(h->ev[h->ev_used + 2], st));
      h->ev_used += 3;
    }
  }
  if (shard) CU(launch_finalize_shard(rgb, acc, depth, n_rays, white_bg ? 1 : 0, *shard, st));
  else CU(launch_finalize(rgb, acc, n_rays, white_bg ? 1 : 0, st));
  return NGF_OK;
}

extern "C" {

// The workspace of a device-resident render issued on stream `st` (see NgfField_::StreamWs).
struct WsRef {
  unsigned int* counters;
  QEntry** queue;
  long long* queue_cap;
  float** acc_ws;
  long long* acc_cap;
};

static int ws_for_stream(NgfField h, cudaStream_t st, WsRef* out) {
  constexpr int kSlots = (int)(sizeof(h->sws) / sizeof(h->sws[0]));
  int slot = -1, free_slot = -1, oldest = 0;
  for (int i = 0; i < kSlots; ++i) {
    if (h->sws[i].used && h->sws[i].stream == st) { slot = i; break; }
    if (!h->sws[i].used && free_slot < 0) free_slot = i;
    if (h->sws[i].last_use < h->sws[oldest].last_use) oldest = i;
  }
  if (slot < 0) {
    slot = free_slot >= 0 ? free_slot : oldest;
    NgfField_::StreamWs& w = h->sws[slot];
    if (w.used && cudaStreamSynchronize(w.stream) != cudaSuccess) {
      // recycled: its last render must be done with the queue.  The caller may have destroyed that stream since (CUDA then
      // finishes its work on its own): fall back to a device-wide wait
      cudaGetLastError();
      CU(cudaDeviceSynchronize());
    }
    w.stream = st; w.used = true;
    if (slot > 0 && !w.counters) {
      CU(cudaMalloc(reinterpret_cast<void**>(&w.counters), kCounterBytes));
      CU(cudaMemset(w.counters, 0, kCounterBytes));
    }
  }
  NgfField_::StreamWs& w = h->sws[slot];
  w.last_use = ++h->sws_clock;
  if (slot == 0) *out = WsRef{h->counters, &h->queue, &h->queue_cap, &h->acc_ws, &h->acc_cap};
  else *out = WsRef{w.counters, &w.queue, &w.queue_cap, &w.acc_ws, &w.acc_cap};
  h->stats_counters = out->counters;
  return NGF_OK;
}

static int ws_acc(const WsRef& ws, long long n_rays, cudaStream_t st, float** acc) {
  if (*ws.acc_cap < n_rays) {
    CU(cudaStreamSynchronize(st));
    cudaFree(*ws.acc_ws);
    *ws.acc_ws = nullptr; *ws.acc_cap = 0;
    CU(dev_alloc(ws.acc_ws, (size_t)n_rays));
    *ws.acc_cap = n_rays;
  }
  *acc = *ws.acc_ws;
  return NGF_OK;
}

static int field_render(NgfField h, const float* rays_dev, int64_t n_rays, int32_t ray_stride, int32_t n_samples,
                        int32_t white_bg, int32_t tile_w, const float* jitter_dev, float* rgb_dev, float* depth_dev,
                        float* acc_dev, int32_t mlp_impl, void* stream) {
  if (!h) return fail(NGF_EINVAL, "field is NULL");
  if (n_rays < 0 || n_rays > 0x7fffffffll) return fail(NGF_EINVAL, "n_rays=%lld", (long long)n_rays);
  if (n_rays == 0) return NGF_OK;
  if (!rays_dev || !rgb_dev || !depth_dev) return fail(NGF_EINVAL, "NULL ray/output pointer");
  if (ray_stride < 6) return fail(NGF_EINVAL, "ray_stride=%d (< 6)", ray_stride);
  if (mlp_impl != NGF_MLP_TCGEN05 && mlp_impl != NGF_MLP_SIMT) return fail(NGF_EINVAL, "mlp_impl=%d", mlp_impl);
  DeviceGuard g(h->device);
  if (!g.ok) return fail(NGF_ECUDA, "cannot select device %d", h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  WsRef ws;
  int rc = ws_for_stream(h, st, &ws);
  if (rc) return rc;
  float* acc = acc_dev;
  if (!acc && (rc = ws_acc(ws, n_rays, st, &acc))) return rc;
  return ngf_render_dev(h, rays_dev, n_rays, ray_stride, n_samples, white_bg, tile_w, rgb_dev, depth_dev, acc,
                        ws.counters, ws.queue, ws.queue_cap, mlp_impl, st, nullptr, jitter_dev);
}

int ngf_field_render(NgfField h, const float* rays_dev, int64_t n_rays, int32_t ray_stride, int32_t n_samples,
                     int32_t white_bg, int32_t tile_w, float* rgb_dev, float* depth_dev, float* acc_dev,
                     int32_t mlp_impl, void* stream) {
  return field_render(h, rays_dev, n_rays, ray_stride, n_samples, white_bg, tile_w, nullptr, rgb_dev, depth_dev, acc_dev,
                      mlp_impl, stream);
}

int ngf_field_render_jitter(NgfField h, const float* rays_dev, int64_t n_rays, int32_t ray_stride, int32_t n_samples,
                            int32_t white_bg, int32_t tile_w, const float* jitter_dev, float* rgb_dev, float* depth_dev,
                            float* acc_dev, int32_t mlp_impl, void* stream) {
  if (!jitter_dev && n_rays > 0) return fail(NGF_EINVAL, "jitter is NULL");
  return field_render(h, rays_dev, n_rays, ray_stride, n_samples, white_bg, tile_w, jitter_dev, rgb_dev, depth_dev,
                      acc_dev, mlp_impl, stream);
}

// Enqueue one whole frame as a three-stage pipeline over chunks: uploads back to back on s_in, kernels on kHostComp
// compute streams, downloads back to back on s_out, chained by events; a chunk buffer (slot) is reused every kHostSlots
// chunks.  The frame is complete when s_out has drained (origin of the fork/join is s_in, so the same code runs under
// stream capture).
static int host_enqueue(NgfField h, const float* rays_host, long long n_rays, int ray_stride, int n_samples, int white_bg,
                        int tile_w, float* rgb_host, float* depth_host, int mlp_impl, long long chunk, bool img,
                        bool join, const CamDev* cam = nullptr, uint8_t* u8_host = nullptr) {
  CU(cudaEventRecord(h->ev_fork, h->s_in));
  CU(cudaStreamWaitEvent(h->s_out, h->ev_fork, 0));
  for (int i = 0; i < kHostComp; ++i) CU(cudaStreamWaitEvent(h->chunk[i].stream, h->ev_fork, 0));
  int ci = h->next_slot;
  for (long long s = 0; s < n_rays; s += chunk, ci = (ci + 1) % kHostSlots) {
    const long long n = (n_rays - s) < chunk ? (n_rays - s) : chunk;
    HostChunk& c = h->chunk[ci];
    // upload: the slot's ray buffer is free once the kernels of its previous chunk are done
    if (!cam) {
      CU(cudaStreamWaitEvent(h->s_in, c.ev_comp, 0));
      CU(cudaMemcpyAsync(c.rays, rays_host + s * ray_stride, (size_t)n * ray_stride * sizeof(float),
                         cudaMemcpyHostToDevice, h->s_in));
      CU(cudaEventRecord(c.ev_in, h->s_in));
      CU(cudaStreamWaitEvent(c.stream, c.ev_in, 0));     // kernels need the rays ...
    }
    CU(cudaStreamWaitEvent(c.stream, c.ev_out, 0));      // ... and the slot's result buffers must have been downloaded
    CamDev cam_chunk{};
    if (cam) { cam_chunk = *cam; cam_chunk.base = s; }
    // slots i and i + kHostComp run on the same compute stream, one after the other: they share one colour queue and one
    // set of counters (the queue is sized for the worst case of a chunk, n * S items)
    HostChunk& qs = h->chunk[ci % kHostComp];
    int rc = ngf_render_dev(h, cam ? nullptr : c.rays, n, ray_stride, n_samples, white_bg, img ? tile_w : 0, c.rgb, c.depth,
                        c.acc, qs.counters, &qs.queue, &qs.queue_cap, mlp_impl, c.stream, cam ? &cam_chunk : nullptr);
    if (rc) return rc;
    if (u8_host) CU(launch_frame_post(c.rgb, nullptr, n * 3, c.u8, nullptr, h->num_sms, c.stream));   // main.py:116
    CU(cudaEventRecord(c.ev_comp, c.stream));
    // download
    CU(cudaStreamWaitEvent(h->s_out, c.ev_comp, 0));
    if (u8_host) CU(cudaMemcpyAsync(u8_host + s * 3, c.u8, (size_t)n * 3, cudaMemcpyDeviceToHost, h->s_out));
    if (rgb_host) CU(cudaMemcpyAsync(rgb_host + s * 3, c.rgb, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->s_out));
    if (depth_host) CU(cudaMemcpyAsync(depth_host + s, c.depth, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->s_out));
    CU(cudaEventRecord(c.ev_out, h->s_out));
  }
  h->next_slot = ci;
  if (!join) return NGF_OK;               // eager callers wait on s_out; joining would stall the next frame's uploads
  // stream capture needs every forked stream to flow back into the origin (s_in)
  for (int i = 0; i < kHostComp; ++i) {
    CU(cudaEventRecord(h->ev_join[i], h->chunk[i].stream));
    CU(cudaStreamWaitEvent(h->s_in, h->ev_join[i], 0));
  }
  CU(cudaEventRecord(h->ev_join[kHostComp], h->s_out));
  CU(cudaStreamWaitEvent(h->s_in, h->ev_join[kHostComp], 0));
  return NGF_OK;
}

// Validate, pick the chunk size and make sure the slot buffers exist.
static int host_prepare(NgfField h, const float* rays_host, int64_t n_rays, int32_t ray_stride, int32_t tile_w,
                        float* rgb_host, float* depth_host, int32_t mlp_impl, bool pipelined, long long* chunk_out,
                        bool* img_out, bool camera = false, bool u8_only = false) {
  if (n_rays < 0 || n_rays > 0x7fffffffll) return fail(NGF_EINVAL, "n_rays=%lld", (long long)n_rays);
  if ((!rays_host && !camera) || ((!rgb_host || !depth_host) && !u8_only)) return fail(NGF_EINVAL, "NULL ray/output pointer");
  if (ray_stride < 6) return fail(NGF_EINVAL, "ray_stride=%d (< 6)", ray_stride);
  if (mlp_impl != NGF_MLP_TCGEN05 && mlp_impl != NGF_MLP_SIMT) return fail(NGF_EINVAL, "mlp_impl=%d", mlp_impl);
  // Chunking (whole groups of 4 image rows when the image width is known).  A single synchronous frame overlaps its own
  // copies and kernels best with ~128 Ki-ray chunks; when frames are pipelined (async API) the overlap comes from the
  // neighbouring frames and larger chunks keep the kernels efficient.  [B200: 0.72 ms sync, 0.38 ms pipelined per
  // 640 000-ray frame]
  const bool img = tile_w > 0 && n_rays % tile_w == 0;
  static long long chunk_sync = 0, chunk_async = 0;
  if (chunk_sync == 0) {
    const char* e = getenv("NGF_HOST_CHUNK");
    chunk_sync = e && atoll(e) > 0 ? atoll(e) : 128 * 1024;
    const char* ea = getenv("NGF_HOST_CHUNK_ASYNC");
    chunk_async = ea && atoll(ea) > 0 ? atoll(ea) : (e && atoll(e) > 0 ? atoll(e) : 320 * 1000);
  }
  long long chunk = pipelined ? chunk_async : chunk_sync;
  if (img) {
    long long rows = chunk / tile_w;
    rows = rows < 4 ? 4 : rows - rows % 4;
    chunk = rows * tile_w;
  }
  if (chunk > n_rays) chunk = n_rays;
  if (h->chunk_cap < chunk || h->chunk_stride != ray_stride) {
    if (h->s_in) { CU(cudaStreamSynchronize(h->s_in)); CU(cudaStreamSynchronize(h->s_out)); }
    free_chunks(h);
    CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    for (auto& e : h->ev_join) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : h->frame_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < kHostSlots; ++i) {
      HostChunk& c = h->chunk[i];
      if (i < kHostComp) CU(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
      else c.stream = h->chunk[i % kHostComp].stream;
      CU(cudaEventCreateWithFlags(&c.ev_in, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c.ev_comp, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c.ev_out, cudaEventDisableTiming));
      CU(dev_alloc(&c.rays, (size_t)chunk * ray_stride));
      CU(dev_alloc(&c.rgb, (size_t)chunk * 3));
      CU(dev_alloc(&c.depth, (size_t)chunk));
      CU(dev_alloc(&c.acc, (size_t)chunk));
      CU(dev_alloc(&c.u8, (size_t)chunk * 3));
      CU(cudaMalloc(reinterpret_cast<void**>(&c.counters), kCounterBytes));
    }
    h->chunk_cap = chunk;
    h->chunk_stride = ray_stride;
  }
  *chunk_out = chunk;
  *img_out = img;
  return NGF_OK;
}

int ngf_field_render_host(NgfField h, const float* rays_host, int64_t n_rays, int32_t ray_stride,
                          int32_t n_samples, int32_t white_bg, int32_t tile_w, float* rgb_host,
                          float* depth_host, int32_t mlp_impl) {
  if (!h) return fail(NGF_EINVAL, "field is NULL");
  if (n_rays == 0) return NGF_OK;
  DeviceGuard g(h->device);
  if (!g.ok) return fail(NGF_ECUDA, "cannot select device %d", h->device);
  long long chunk = 0;
  bool img = false;
  int rc = host_prepare(h, rays_host, n_rays, ray_stride, tile_w, rgb_host, depth_host, mlp_impl, false, &chunk, &img);
  if (rc) return rc;
  static int use_graphs = -1;
  if (use_graphs < 0) {
    const char* ge = getenv("NGF_HOST_GRAPH");
    use_graphs = ge && ge[0] == '0' ? 0 : 1;
  }
  cudaStream_t s0 = h->s_in;

  // A frame that was already rendered once from the same host buffers is replayed as one CUDA-graph launch instead of
  // ~50 stream calls.
  const bool graph_ok = use_graphs == 1 && h->ev.empty();
  NgfField_::HostGraph* hit = nullptr;
  if (graph_ok)
    for (auto& e : h->graphs)
      if (e.rays == rays_host && e.rgb == rgb_host && e.depth == depth_host && e.n_rays == n_rays &&
          e.stride == ray_stride && e.n_samples == n_samples && e.white_bg == white_bg && e.tile_w == tile_w &&
          e.impl == mlp_impl && e.epoch == h->epoch) { hit = &e; break; }
  if (hit && hit->exec) {
    CU(cudaGraphLaunch(hit->exec, s0));
    CU(cudaStreamSynchronize(s0));
    return NGF_OK;
  }
  rc = host_enqueue(h, rays_host, n_rays, ray_stride, n_samples, white_bg, tile_w, rgb_host, depth_host, mlp_impl,
                    chunk, img, false);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->s_out));
  if (graph_ok && !hit) {                 // every buffer now has its final size: record the same pipeline for next time
    if (h->graphs.size() >= 64) drop_graphs(h);
    NgfField_::HostGraph e{rays_host, rgb_host, depth_host, (long long)n_rays, ray_stride, n_samples, white_bg, tile_w,
                           mlp_impl, h->epoch, nullptr};
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      const int crc = host_enqueue(h, rays_host, n_rays, ray_stride, n_samples, white_bg, tile_w, rgb_host, depth_host,
                                   mlp_impl, chunk, img, true);
      const cudaError_t ce = cudaStreamEndCapture(s0, &graph);
      if (crc == NGF_OK && ce == cudaSuccess && graph) {
        if (cudaGraphInstantiate(&e.exec, graph, 0) != cudaSuccess) e.exec = nullptr;
      }
      if (graph) cudaGraphDestroy(graph);
    }
    cudaGetLastError();                   // a failed capture only means this key stays on the eager path
    g_err[0] = 0;
    h->graphs.push_back(e);
  }
  return NGF_OK;
}

int ngf_field_render_host_async(NgfField h, const float* rays_host, int64_t n_rays, int32_t ray_stride,
                                int32_t n_samples, int32_t white_bg, int32_t tile_w, float* rgb_host,
                                float* depth_host, int32_t mlp_impl, uint64_t* ticket) {
  if (!h || !ticket) return fail(NGF_EINVAL, "NULL argument");
  DeviceGuard g(h->device);
  if (!g.ok) return fail(NGF_ECUDA, "cannot select device %d", h->device);
  long long chunk = 0;
  bool img = false;
  int rc = host_prepare(h, rays_host, n_rays, ray_stride, tile_w, rgb_host, depth_host, mlp_impl, true, &chunk, &img);
  if (rc) return rc;
  const unsigned long long t = h->next_ticket;
  cudaEvent_t done = h->frame_done[t % 8];
  CU(cudaEventSynchronize(done));         // at most 8 frames in flight: the slot's previous frame must be finished
  if (n_rays > 0) {
    rc = host_enqueue(h, rays_host, n_rays, ray_stride, n_samples, white_bg, tile_w, rgb_host, depth_host, mlp_impl,
                      chunk, img, false);
    if (rc) return rc;
  }
  CU(cudaEventRecord(done, h->s_out));
  *ticket = t;
  ++h->next_ticket;
  return NGF_OK;
}

static int check_camera(const NgfCamera* c, CamDev* out) {
  if (!c) return fail(NGF_EINVAL, "camera is NULL");
  if (c->width < 1 || c->height < 1 || (long long)c->width * c->height > 0x7fffffffll) return fail(NGF_EINVAL, "camera image %dx%d", c->width, c->height);
  if (!(c->fx != 0.f) || !(c->fy != 0.f)) return fail(NGF_EINVAL, "camera focal length is zero");
  memcpy(out->c2w, c->c2w, sizeof(out->c2w));
  out->fx = c->fx; out->fy = c->fy; out->cx = c->cx; out->cy = c->cy;
  out->W = c->width; out->H = c->height;
  out->base = 0;
  out->poses = nullptr; out->shard_block = 1; out->shard_rank = 0; out->shard_world = 1;
  return NGF_OK;
}

int ngf_field_render_camera(NgfField h, const NgfCamera* camera, int32_t n_samples, int32_t white_bg, float* rgb_dev,
                            float* depth_dev, float* acc_dev, int32_t mlp_impl, void* stream) {
  if (!h) return fail(NGF_EINVAL, "field is NULL");
  CamDev cam{};
  int rc = check_camera(camera, &cam);
  if (rc) return rc;
  if (!rgb_dev || !depth_dev) return fail(NGF_EINVAL, "NULL output pointer");
  if (mlp_impl != NGF_MLP_TCGEN05 && mlp_impl != NGF_MLP_SIMT) return fail(NGF_EINVAL, "mlp_impl=%d", mlp_impl);
  DeviceGuard g(h->device);
  if (!g.ok) return fail(NGF_ECUDA, "cannot select device %d", h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n_rays = (long long)cam.W * cam.H;
  WsRef ws;
  rc = ws_for_stream(h, st, &ws);
  if (rc) return rc;
  float* acc = acc_dev;
  if (!acc && (rc = ws_acc(ws, n_rays, st, &acc))) return rc;
  return ngf_render_dev(h, nullptr, n_rays, 6, n_samples, white_bg, cam.W, rgb_dev, depth_dev, acc, ws.counters, ws.queue,
                        ws.queue_cap, mlp_impl, st, &cam);
}

int ngf_field_render_camera_host_async(NgfField h, const NgfCamera* camera, int32_t n_samples, int32_t white_bg,
                                       float* rgb_host, float* depth_host, int32_t mlp_impl, uint64_t* ticket) {
  if (!h || !ticket) return fail(NGF_EINVAL, "NULL argument");
  CamDev cam{};
  int rc = check_camera(camera, &cam);
  if (rc) return rc;
  DeviceGuard g(h->device);
  if (!g.ok) return fail(NGF_ECUDA, "cannot select device %d", h->device);
  const long long n_rays = (long long)cam.W * cam.H;
  long long chunk = 0;
  bool img = false;
  rc = host_prepare(h, nullptr, n_rays, 6, cam.W, rgb_host, depth_host, mlp_impl, true, &chunk, &img, true);
  if (rc) return rc;
  const unsigned long long t = h->next_ticket;
  cudaEvent_t done = h->frame_done[t % 8];
  CU(cudaEventSynchronize(done));
  rc = host_enqueue(h, nullptr, n_rays, 6, n_samples, white_bg, cam.W, rgb_host, depth_host, mlp_impl, chunk, img, false, &cam);
  if (rc) return rc;
  CU(cudaEventRecord(done, h->s_out));
  *ticket = t;
  ++h->next_ticket;
  return NGF_OK;
}

int ngf_field_render_camera_u8_host_async(NgfField h, const NgfCamera* camera, int32_t n_samples, int32_t white_bg,
                                          uint8_t* u8_host, float* depth_host, int32_t mlp_impl, uint64_t* ticket) {
  if (!h || !ticket || !u8_host) return fail(NGF_EINVAL, "NULL argument");
  CamDev cam{};
  int rc = check_camera(camera, &cam);
  if (rc) return rc;
  DeviceGuard g(h->device);
  if (!g.ok) return fail(NGF_ECUDA, "cannot select device %d", h->device);
  const long long n_rays = (long long)cam.W * cam.H;
  long long chunk = 0;
  bool img = false;
  rc = host_prepare(h, nullptr, n_rays, 6, cam.W, nullptr, depth_host, mlp_impl, true, &chunk, &img, true, true);
  if (rc) return rc;
  const unsigned long long t = h->next_ticket;
  cudaEvent_t done = h->frame_done[t % 8];
  CU(cudaEventSynchronize(done));
  rc = host_enqueue(h, nullptr, n_rays, 6, n_samples, white_bg, cam.W, nullptr, depth_host, mlp_impl, chunk, img, false, &cam,
                    u8_host);
  if (rc) return rc;
  CU(cudaEventRecord(done, h->s_out));
  *ticket = t;
  ++h->next_ticket;
  return NGF_OK;
}

int ngf_field_host_wait(NgfField h, uint64_t ticket) {
  if (!h) return fail(NGF_EINVAL, "field is NULL");
  if (ticket == 0 || ticket >= h->next_ticket) return fail(NGF_EINVAL, "unknown ticket %llu", (unsigned long long)ticket);
  if (ticket + 8 < h->next_ticket) return NGF_OK;     // its event slot was recycled by ticket + 8, which waited for it
  DeviceGuard g(h->device);
  CU(cudaEventSynchronize(h->frame_done[ticket % 8]));
  return NGF_OK;
}

int ngf_field_set_gauge(NgfField h, int32_t on) {
  if (!h) return fail(NGF_EINVAL, "field is NULL");
  if (on && !h->has_gauge) return fail(NGF_EINVAL, "field was packed without gauge planes");
  if (h->dev.gauge_on != (on ? 1 : 0)) ++h->epoch;
  h->dev.gauge_on = on ? 1 : 0;
  return NGF_OK;
}

int ngf_field_set_infoinv(NgfField h, int32_t on) {
  if (!h) return fail(NGF_EINVAL, "field is NULL");
  if (h->dev.infoinv != (on ? 1 : 0)) ++h->epoch;
  h->dev.infoinv = on ? 1 : 0;
  return NGF_OK;
}

int ngf_field_stats(NgfField h, NgfStats* out, void* stream) {
  if (!h || !out) return fail(NGF_EINVAL, "NULL argument");
  DeviceGuard g(h->device);
  unsigned long long s[5];
  CU(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  CU(cudaMemcpy(s, (h->stats_counters ? h->stats_counters : h->counters) + 2, sizeof(s), cudaMemcpyDeviceToHost));
  out->rays = 0;
  out->samples_in_box = s[0]; out->samples_density = s[1]; out->samples_colour = s[2]; out->mlp_tiles = s[3];
  out->direct_patches = s[4];
  return NGF_OK;
}

int ngf_field_timing_begin(NgfField h, int32_t capacity) {
  if (!h) return fail(NGF_EINVAL, "field is NULL");
  if (capacity < 0 || capacity > 65536) return fail(NGF_EINVAL, "capacity=%d", capacity);
  DeviceGuard g(h->device);
  free_events(h);
  h->ev.resize((size_t)capacity * 3);
  for (auto& e : h->ev) e = nullptr;
  for (auto& e : h->ev) CU(cudaEventCreate(&e));
  return NGF_OK;
}

int ngf_field_timing_read(NgfField h, int32_t* n_launches, double* march_ms, double* colour_ms) {
  if (!h || !n_launches || !march_ms || !colour_ms) return fail(NGF_EINVAL, "NULL argument");
  DeviceGuard g(h->device);
  double sm = 0.0, sc = 0.0;
  for (int i = 0; i + 2 < h->ev_used; i += 3) {
    float a = 0.f, b = 0.f;
    CU(cudaEventSynchronize(h->ev[i + 2]));
    CU(cudaEventElapsedTime(&a, h->ev[i], h->ev[i + 1]));
    CU(cudaEventElapsedTime(&b, h->ev[i + 1], h->ev[i + 2]));
    sm += a; sc += b;
  }
  *n_launches = h->ev_used / 3;
  *march_ms = sm;
  *colour_ms = sc;
  h->ev_used = 0;
  return NGF_OK;
}

#define NGF_POINTWISE_PROLOGUE()                                              \
  if (!h) return fail(NGF_EINVAL, "field is NULL");                           \
  if (n < 0) return fail(NGF_EINVAL, "negative count");                       \
  if (n == 0) return NGF_OK;                                                  \
  DeviceGuard g(h->device);                                                   \
  if (!g.ok) return fail(NGF_ECUDA, "cannot select device %d", h->device);    \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

int ngf_field_sample_ray(NgfField h, const float* rays_dev, int64_t n, int32_t ray_stride, int32_t n_samples,
                         float* pts_dev, float* t_dev, uint8_t* inside_dev, void* stream) {
  NGF_POINTWISE_PROLOGUE();
  if (!rays_dev || !pts_dev || !t_dev || !inside_dev) return fail(NGF_EINVAL, "NULL pointer");
  if (ray_stride < 6) return fail(NGF_EINVAL, "ray_stride=%d (< 6)", ray_stride);
  const int S = n_samples > 0 ? n_samples : h->n_samples_default;
  CU(launch_sample_ray(h->dev, rays_dev, n, ray_stride, S, nullptr, pts_dev, t_dev, inside_dev, st));
  return NGF_OK;
}

int ngf_field_sample_ray_jitter(NgfField h, const float* rays_dev, int64_t n, int32_t ray_stride, int32_t n_samples,
                                const float* jitter_dev, float* pts_dev, float* t_dev, uint8_t* inside_dev, void* stream) {
  NGF_POINTWISE_PROLOGUE();
  if (!rays_dev || !jitter_dev || !pts_dev || !t_dev || !inside_dev) return fail(NGF_EINVAL, "NULL pointer");
  if (ray_stride < 6) return fail(NGF_EINVAL, "ray_stride=%d (< 6)", ray_stride);
  const int S = n_samples > 0 ? n_samples : h->n_samples_default;
  CU(launch_sample_ray(h->dev, rays_dev, n, ray_stride, S, jitter_dev, pts_dev, t_dev, inside_dev, st));
  return NGF_OK;
}

int ngf_field_alpha_keep(NgfField h, const float* pts_dev, int64_t n, uint8_t* keep_dev, void* stream) {
  NGF_POINTWISE_PROLOGUE();
  if (!pts_dev || !keep_dev) return fail(NGF_EINVAL, "NULL pointer");
  CU(launch_alpha_keep(h->dev, pts_dev, n, keep_dev, st));
  return NGF_OK;
}

int ngf_field_alpha_value(NgfField h, const float* pts_dev, int64_t n, float* value_dev, void* stream) {
  NGF_POINTWISE_PROLOGUE();
  if (!pts_dev || !value_dev) return fail(NGF_EINVAL, "NULL pointer");
  CU(launch_alpha_value(h->dev, pts_dev, n, value_dev, st));
  return NGF_OK;
}

int ngf_field_gauge(NgfField h, const float* xyz_norm_dev, int64_t n, int32_t gauge_on, float* xy_dev,
                    float* yz_dev, float* xz_dev, void* stream) {
  NGF_POINTWISE_PROLOGUE();
  if (!xyz_norm_dev || !xy_dev || !yz_dev || !xz_dev) return fail(NGF_EINVAL, "NULL pointer");
  if (gauge_on && h->dev.variant == 0 && !h->has_gauge)
    return fail(NGF_EINVAL, "gauge requested but the field was packed without gauge planes");
  CU(launch_gauge(h->dev, xyz_norm_dev, n, gauge_on, xy_dev, yz_dev, xz_dev, st));
  return NGF_OK;
}

int ngf_field_density(NgfField h, const float* xy_dev, const float* yz_dev, const float* xz_dev, int64_t n,
                      float* sigma_dev, void* stream) {
  NGF_POINTWISE_PROLOGUE();
  if (!xy_dev || !yz_dev || !xz_dev || !sigma_dev) return fail(NGF_EINVAL, "NULL pointer");
  CU(launch_density(h->dev, xy_dev, yz_dev, xz_dev, n, sigma_dev, st));
  return NGF_OK;
}

int ngf_field_rgb(NgfField h, const float* xy_dev, const float* yz_dev, const float* xz_dev,
                  const float* viewdirs_dev, int64_t n, float* rgb_dev, int32_t mlp_impl, void* stream) {
  NGF_POINTWISE_PROLOGUE();
  if (!xy_dev || !yz_dev || !xz_dev || !viewdirs_dev || !rgb_dev) return fail(NGF_EINVAL, "NULL pointer");
  if (n > 0x7fffffffll) return fail(NGF_EINVAL, "n too large");
  if (mlp_impl != NGF_MLP_TCGEN05 && mlp_impl != NGF_MLP_SIMT) return fail(NGF_EINVAL, "mlp_impl=%d", mlp_impl);
  CU(launch_rgb(h->dev, xy_dev, yz_dev, xz_dev, viewdirs_dev, n, rgb_dev, mlp_impl, h->num_sms, st));
  return NGF_OK;
}

int ngf_field_sigma_world(NgfField h, const float* pts_dev, int64_t n, int32_t use_gauge, float* sigma_dev,
                          void* stream) {
  NGF_POINTWISE_PROLOGUE();
  if (!pts_dev || !sigma_dev) return fail(NGF_EINVAL, "NULL pointer");
  CU(launch_sigma_world(h->dev, pts_dev, n, use_gauge, sigma_dev, st));
  return NGF_OK;
}

int ngf_frame_post(const float* rgb_dev, const float* gt_dev, int64_t n_values, uint8_t* u8_dev, double* sse_dev,
                   void* stream) {
  if (!rgb_dev) return fail(NGF_EINVAL, "rgb is NULL");
  if (n_values < 0) return fail(NGF_EINVAL, "negative count");
  if (gt_dev && !sse_dev) return fail(NGF_EINVAL, "gt given but sse is NULL");
  if (!gt_dev && !u8_dev) return fail(NGF_EINVAL, "nothing to do: neither u8 nor gt given");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  CU(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (gt_dev) CU(cudaMemsetAsync(sse_dev, 0, sizeof(double), st));
  CU(launch_frame_post(rgb_dev, gt_dev, n_values, u8_dev, sse_dev, sms, st));
  return NGF_OK;
}

int ngf_depth_colormap(const float* depth_dev, int64_t n, double min_depth, double max_depth, uint8_t* bgr_dev,
                       void* stream) {
  if (!depth_dev || !bgr_dev) return fail(NGF_EINVAL, "NULL pointer");
  if (n < 0) return fail(NGF_EINVAL, "negative count");
  int dev = 0, sms = 148;
  CU(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // numpy casts the Python scalars mi and (ma - mi + 1e-8) to the array's fp32
  CU(launch_depth_colormap(depth_dev, n, (float)min_depth, (float)(max_depth - min_depth + 1e-8), bgr_dev, sms,
                           reinterpret_cast<cudaStream_t>(stream)));
  return NGF_OK;
}

int64_t ngf_shard_count(int64_t n_rays, int32_t block, int32_t rank, int32_t world) {
  if (n_rays < 0 || block < 1 || world < 1 || rank < 0 || rank >= world) return -1;
  const long long per_cycle = (long long)block * world;
  const long long full = n_rays / per_cycle, rem = n_rays - full * per_cycle;
  long long extra = rem - (long long)rank * block;
  if (extra < 0) extra = 0;
  if (extra > block) extra = block;
  return full * block + extra;
}

int ngf_shard_gather(const float* src_dev, int64_t n_rays, int32_t width, int32_t block, int32_t rank,
                     int32_t world, float* dst_dev, void* stream) {
  if (!src_dev || !dst_dev) return fail(NGF_EINVAL, "NULL pointer");
  if (ngf_shard_count(n_rays, block, rank, world) < 0 || width < 1) return fail(NGF_EINVAL, "bad shard arguments");
  CU(launch_shard_gather(src_dev, n_rays, width, block, rank, world, dst_dev, reinterpret_cast<cudaStream_t>(stream)));
  return NGF_OK;
}

int ngf_shard_scatter(const float* src_dev, int64_t n_rays, int32_t width, int32_t block, int32_t world,
                      int64_t max_shard, float* dst_dev, void* stream) {
  if (!src_dev || !dst_dev) return fail(NGF_EINVAL, "NULL pointer");
  if (n_rays < 0 || block < 1 || world < 1 || width < 1 || max_shard < ngf_shard_count(n_rays, block, 0, world))
    return fail(NGF_EINVAL, "bad shard arguments");
  CU(launch_shard_scatter(src_dev, n_rays, width, block, world, max_shard, dst_dev,
                          reinterpret_cast<cudaStream_t>(stream)));
  return NGF_OK;
}

}  // extern "C"
