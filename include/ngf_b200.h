/*
 * ngf_b200.h — C ABI of libngf_b200.so: the B200 (sm_100a) implementation of the volumetric-rendering hot
 * path of fnzhan/Neural-Gauge-Fields (TriPlane / InfoInv sub-projects).
 *
 * The reference is pure Python/PyTorch and has no FFI; its "plugin" seam is the model class that
 * `eval(args.model_name)(**kwargs)` instantiates (TriPlane/main.py:37,226,230) and whose forward() the
 * `renderer` chunk loop calls (TriPlane/main.py:60-71).  Each entry point below states the reference
 * interface it replaces (paths relative to /root/reference).  INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every function returns 0 (NGF_OK) or a negative NgfStatus; ngf_last_error() gives the thread-local text.
 *   - "dev" pointers are CUDA device pointers on the handle's device; "host" pointers are CPU memory
 *     (pinned memory makes the copies asynchronous and fast; pageable memory also works).
 *   - the caller owns every buffer it passes; the library copies what it needs during ngf_field_pack and never
 *     retains caller pointers after a call returns.  The handle owns its packed shadows and workspaces.
 *   - all device work is enqueued on the `stream` argument (a cudaStream_t passed as void*; NULL = default
 *     stream) and is asynchronous with respect to the host unless stated otherwise.  No host synchronisation
 *     happens inside ngf_field_render (the reference forward() has ~15 per 4096-ray chunk).
 *   - one handle per device; a handle may be used from one stream at a time.
 */
#ifndef NGF_B200_H_
#define NGF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGF_ABI_VERSION 1

typedef enum NgfStatus {
  NGF_OK = 0,
  NGF_EINVAL = -1,       /* bad shape / stride / null pointer / alignment */
  NGF_ECUDA = -2,        /* CUDA runtime error; text in ngf_last_error()   */
  NGF_EUNSUPPORTED = -3, /* configuration outside what the kernels are built for */
  NGF_ENOMEM = -4,
  NGF_ECOMM = -5         /* a peer rank did not answer in time (multi-GPU frame exchange) */
} NgfStatus;

typedef enum NgfVariant {
  NGF_TRIPLANE = 0, /* TriPlane/models/Field.py: 64-ch planes (16 density + 48 appearance), learned gauge planes */
  NGF_INFOINV = 1   /* InfoInv/models/Field.py: 96-ch planes (24 + 72), sinusoidal phase product, density MLP   */
} NgfVariant;

typedef enum NgfMlpImpl {
  NGF_MLP_TCGEN05 = 0, /* colour MLP on 5th-gen tensor cores (fp16 operands, fp32 accumulate in TMEM) */
  NGF_MLP_SIMT = 1     /* same packed weights on CUDA cores; debugging / cross-check only            */
} NgfMlpImpl;

/* A dense linear layer y = W x + b in PyTorch nn.Linear layout: w is [out][in] row-major fp32, b is [out]
 * (b may be NULL for a bias-free layer). */
typedef struct NgfLinear {
  const float* w;
  const float* b;
  int32_t in_dim;
  int32_t out_dim;
} NgfLinear;

/*
 * Field description: the parameters and attributes Base.forward() reads
 * (TriPlane/models/FieldBase.py:251-312, InfoInv/models/FieldBase.py:228-282).  Parameter pointers may be host
 * or device memory (they are copied with cudaMemcpyDefault while packing).  All arrays are fp32 in the
 * reference's own NCHW layouts, so a state_dict tensor's data_ptr can be passed as is.
 */
typedef struct NgfFieldDesc {
  int32_t variant; /* NgfVariant */

  /* plane_xy [1,C,Hy,Wx], plane_yz [1,C,Hz,Wy], plane_xz [1,C,Hz,Wx] (Field.py:19-21; non-square after
   * up_sampling/shrink, Field.py:108-132) */
  const float* plane[3];
  int32_t plane_h[3];
  int32_t plane_w[3];
  int32_t plane_c;   /* 64 (TriPlane) | 96 (InfoInv) */
  int32_t density_c; /* 16 | 24: channels [0,density_c) feed density, the rest feed colour */

  /* gauge_xy / gauge_yz / gauge_xz [1,2,Hg,Wg] (TriPlane/models/Field.py:23-26); uploaded when all three are
   * non-NULL.  gauge_on = (iteration >= gauge_start), TriPlane/models/Field.py:58: initial value of the switch
   * ngf_field_set_gauge() flips. */
  const float* gauge[3];
  int32_t gauge_h[3];
  int32_t gauge_w[3];
  int32_t gauge_on;

  /* rgb_decoder (networks.py:12-32): basis (bias-free F x F), mlp.0 (F+3+12 -> 64), mlp.2 (64 -> 64),
   * mlp.4 (64 -> 3); view_pe = 2. */
  NgfLinear rgb_basis;
  NgfLinear rgb_l1, rgb_l2, rgb_l3;
  int32_t view_pe;

  /* density head: TriPlane: density_decoder Linear(48,1) in dens_l1 (dens_l2/l3 unused, Field.py:29);
   * InfoInv: density_decoder.mlp.{0,2,4} = 72->32->32->1 (InfoInv/models/networks.py:34-54). */
  NgfLinear dens_l1, dens_l2, dens_l3;
  float density_shift; /* feature2density default -10 (Field.py:48) */
  int32_t infoinv;     /* forward(..., infoinv=True): multiply plane features by the phase code */

  /* Base attributes (FieldBase.py:45-74).  inv_aabb_size / step_size must be the fp32 values torch computed
   * (Base.invaabbSize, Base.stepSize) so every mask decision rounds exactly like the reference. */
  float aabb[6];          /* aabb[0] = min xyz, aabb[1] = max xyz */
  float inv_aabb_size[3]; /* 2 / (aabb[1]-aabb[0]) */
  float step_size;
  int32_t n_samples; /* Base.nSamples: used when N_samples <= 0 */
  float near_t, far_t;
  float distance_scale;
  float weight_thres; /* rayMarch_weight_thres */

  /* AlphaGridMask (FieldBase.py:22-40) or NULL: alpha_volume is the [D][H][W] fp32 {0,1} volume,
   * alpha_dims = {W, H, D}, alpha_aabb its own box, alpha_inv = AlphaGridMask.invgridSize (fp32 from torch). */
  const float* alpha_volume;
  int32_t alpha_dims[3];
  float alpha_aabb[6];
  float alpha_inv[3];
} NgfFieldDesc;

/* Counters of the last render (device-side sums copied on request; forces a stream sync). */
typedef struct NgfStats {
  uint64_t rays;
  uint64_t samples_in_box;  /* samples inside the box after the conservative range clip   */
  uint64_t samples_density; /* density evaluations: valid samples (bbox and alpha mask)    */
  uint64_t samples_colour;  /* colour-MLP evaluations: weight > weight_thres               */
  uint64_t mlp_tiles;       /* 128-sample colour-MLP tiles issued                          */
  uint64_t direct_patches;  /* (32-sample group, plane) pairs whose taps did not fit a TMA patch and were gathered directly
                               (of 12 per tile; TMA-staged colour kernel only) */
} NgfStats;

typedef struct NgfField_* NgfField;

int ngf_abi_version(void);
const char* ngf_last_error(void);
/* Number of kernels this library has launched in the calling process so far (bench.py's gpu_launches). */
uint64_t ngf_launch_count(void);

/*
 * Build the device-side shadows of a field on `device`: density texels projected through the density head
 * (TriPlane) or channels-last fp32 (InfoInv), channels-last fp16 appearance texels, float2 gauge texels,
 * bit-packed occupancy grids, folded (basis . mlp.0) fp16 weights in tcgen05 shared-memory layout.
 * Replaces: TriPlane.__init__/init_model + Base.load (Field.py:14-32, FieldBase.py:111-116) as far as the
 * render path is concerned.  Synchronous.
 */
int ngf_field_pack(const NgfFieldDesc* desc, int device, NgfField* out);
/* Re-read parameters after an optimizer step / load_state_dict: same shapes required. Synchronous. */
int ngf_field_repack(NgfField f, const NgfFieldDesc* desc);
void ngf_field_free(NgfField f);

/*
 * Render rays.  Replaces Base.forward(rays_chunk, white_bg, is_train=False, N_samples, ...) for a whole
 * frame at once (FieldBase.py:251-312) and therefore also the chunk loop `renderer` (main.py:60-71).
 *   rays_dev   [R][ray_stride] fp32, columns 0-2 origin, 3-5 direction (ray_stride >= 6; the LAST column feeds
 *              the depth background term exactly as FieldBase.py:306 does)
 *   n_samples  N_samples argument (<= 0: use desc.n_samples)
 *   rgb_dev    [R][3] fp32 out (rgb_map), depth_dev [R] fp32 out (depth_map); acc_dev [R] fp32 out or NULL
 *   tile_w     0, or the image width when rays are the row-major pixels of an image: lets the kernel map
 *              warps to 8x4 pixel blocks for texel locality (results do not depend on it)
 *   stream     the cudaStream_t the kernels are enqueued on (asynchronous).  The handle keeps one workspace (colour queue,
 *              counters) per caller stream, so frames issued on different streams are independent and overlap on the
 *              device — on B200 three streams render independent 640 000-ray frames in 0.29 ms each against 0.38 ms back
 *              to back on one.  Up to four streams are remembered; a fifth recycles the least recently used workspace
 *              after synchronising its stream.  The colour queue of a workspace is sized for the worst case of a call,
 *              min(n_rays * n_samples * 32 B, NGF_QUEUE_MIB (default 4096) MiB), allocated on first use.
 */
int ngf_field_render(NgfField f, const float* rays_dev, int64_t n_rays, int32_t ray_stride, int32_t n_samples,
                     int32_t white_bg, int32_t tile_w, float* rgb_dev, float* depth_dev, float* acc_dev,
                     int32_t mlp_impl, void* stream);

/*
 * Same with the training-time sampling of Base.sample_ray(is_train=True) (FieldBase.py:128-132): sample i of ray r sits
 * at t_min + stepSize * (i + jitter[r]), jitter_dev [R] fp32 in [0, 1) (the reference draws it with torch.rand_like on
 * the host; the caller does the same and passes it in, so both implementations see the same numbers).  Forward only:
 * there is no backward pass in this library.  jitter 0 everywhere gives ngf_field_render bit for bit.
 */
int ngf_field_render_jitter(NgfField f, const float* rays_dev, int64_t n_rays, int32_t ray_stride, int32_t n_samples,
                            int32_t white_bg, int32_t tile_w, const float* jitter_dev, float* rgb_dev, float* depth_dev,
                            float* acc_dev, int32_t mlp_impl, void* stream);

/*
 * Backward pass of the training step (TriPlane/main.py:272-302: rgb_map = field(rays,
 * is_train=True)['rgb_map'] -> loss -> loss.backward()).  Given dL/d(rgb_map) [R][3] it ADDS dL/d(parameter) into the
 * caller's gradient buffers, which have the reference's own parameter layouts (feature planes [C][H][W], gauge planes
 * [2][Hg][Wg], nn.Linear weights [out][in] and biases), i.e. the `.grad` tensors of the module's parameters.  The forward is
 * not saved: the rays are re-marched with the same jitter (NULL = evaluation-time sampling) and the same sample decisions
 * as ngf_field_render_jitter on the handle's current (packed) parameters, so call it before the parameters change.
 * depth_map carries no gradient (it is computed under torch.no_grad(), FieldBase.py:304-306).  gauge[] may be NULL when
 * the gauge is off; dens_l2 / dens_l3 are InfoInv's (density_decoder.mlp.{2,4}).  fp32 on CUDA cores; synchronises `stream`
 * once (to size its workspace, NGF_TRAIN_MIB MiB at most, default 2048).
 */
typedef struct NgfFieldGrads {
  float* plane[3];
  float* gauge[3];
  /* optional (all three or NULL): the feature-plane parameters themselves, fp32 [C][H][W] device pointers.  The forward
   * of the colour MLP is then recomputed from the reference's fp32 features, so the ReLU masks of the backward are the
   * reference's; with NULL the handle's fp16 appearance texels are used (hidden units within ~1e-3 of zero may flip). */
  const float* plane_param[3];
  float *rgb_basis, *rgb_l1_w, *rgb_l1_b, *rgb_l2_w, *rgb_l2_b, *rgb_l3_w, *rgb_l3_b;
  float *dens_l1_w, *dens_l1_b, *dens_l2_w, *dens_l2_b, *dens_l3_w, *dens_l3_b;
} NgfFieldGrads;
int ngf_field_backward(NgfField f, const float* rays_dev, int64_t n_rays, int32_t ray_stride, int32_t n_samples,
                       int32_t white_bg, const float* jitter_dev, const float* grad_rgb_dev, const NgfFieldGrads* grads,
                       void* stream);

/* One Adam step (torch.optim.Adam without weight decay / amsgrad, the optimiser of TriPlane/main.py:237,300-302) over a flat
 * fp32 parameter in one pass: exp_avg / exp_avg_sq are the optimiser state, `step` the 1-based step count.  The feature
 * planes are 50 MB of parameters; torch's unfused step makes ~10 passes over them. */
int ngf_adam_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, double lr,
                  double beta1, double beta2, double eps, int64_t step, void* stream);

/*
 * Same through HOST buffers: H2D of the rays, render, D2H of rgb/depth, chunked and overlapped on internal
 * streams; returns after the results are in rgb_host/depth_host.  This is the call the reference-facing
 * `renderer(rays_cpu, field, ...)` maps to when rays live on the CPU (main.py:64-65 does the H2D per chunk).
 */
int ngf_field_render_host(NgfField f, const float* rays_host, int64_t n_rays, int32_t ray_stride,
                          int32_t n_samples, int32_t white_bg, int32_t tile_w, float* rgb_host,
                          float* depth_host, int32_t mlp_impl);

/*
 * Asynchronous variant for loops over many frames (the reference's evaluation() renders 200 test frames one after the
 * other, main.py:89-100): enqueues the frame and returns a ticket at once; consecutive frames pipeline on the device
 * (frame k+1's H2D overlaps frame k's compute and D2H).  The host buffers of a frame must stay untouched until
 * ngf_field_host_wait(ticket) returns.  At most 8 frames may be in flight (older tickets are waited for implicitly).
 */
int ngf_field_render_host_async(NgfField f, const float* rays_host, int64_t n_rays, int32_t ray_stride,
                                int32_t n_samples, int32_t white_bg, int32_t tile_w, float* rgb_host,
                                float* depth_host, int32_t mlp_impl, uint64_t* ticket);
int ngf_field_host_wait(NgfField f, uint64_t ticket);

/*
 * On-device ray generation: the rays of a pinhole camera are produced inside the march kernel,
 * so a frame needs 18 numbers instead of a 15 MB ray tensor.  Replaces get_ray_directions + the unit normalisation of
 * blender.py:52 + get_rays (TriPlane/dataLoader/ray_utils.py:24-42,66-87; blender.py:46-52) as evaluation_path uses
 * them per frame (TriPlane/main.py:155-161): pixel (i, j) -> ((i+0.5-cx)/fx, (j+0.5-cy)/fy, 1) / |.| rotated by
 * c2w[:3,:3], origin c2w[:3,3]; rays are the row-major pixels, results as in ngf_field_render with tile_w = width.
 */
typedef struct NgfCamera {
  float c2w[12];          /* row-major [3][4] camera-to-world */
  float fx, fy, cx, cy;   /* focal lengths and principal point in pixels (blender.py: fx = fy = focal, cx = W/2, cy = H/2) */
  int32_t width, height;
} NgfCamera;
int ngf_field_render_camera(NgfField f, const NgfCamera* camera, int32_t n_samples, int32_t white_bg, float* rgb_dev,
                            float* depth_dev, float* acc_dev, int32_t mlp_impl, void* stream);
/* Same into HOST buffers, pipelined like ngf_field_render_host_async (wait with ngf_field_host_wait). */
int ngf_field_render_camera_host_async(NgfField f, const NgfCamera* camera, int32_t n_samples, int32_t white_bg,
                                       float* rgb_host, float* depth_host, int32_t mlp_impl, uint64_t* ticket);

/* evaluation_path's whole per-frame job (TriPlane/main.py:155-161 + the uint8 conversion of main.py:116): camera in,
 * rgb_map as uint8 [H*W][3] out (value = (uint8)(rgb * 255), as numpy's astype truncates), depth_map fp32 optional
 * (NULL: not downloaded).  1.92 MB leave the device per 800x800 frame instead of 10.24 MB. */
int ngf_field_render_camera_u8_host_async(NgfField f, const NgfCamera* camera, int32_t n_samples, int32_t white_bg,
                                          uint8_t* u8_host, float* depth_host, int32_t mlp_impl, uint64_t* ticket);

/* Per-call switches of forward(): TriPlane `iteration >= gauge_start` (TriPlane/models/Field.py:58) and InfoInv
 * `infoinv=` (InfoInv/models/FieldBase.py:228).  They only flip a flag in the handle; no repack. */
int ngf_field_set_gauge(NgfField f, int32_t on);
int ngf_field_set_infoinv(NgfField f, int32_t on);

/* Copy the counters of the last render on `stream` (synchronises that stream). */
int ngf_field_stats(NgfField f, NgfStats* out, void* stream);

/*
 * Device-side timing of the two render kernels (bench.py's roofline): after ngf_field_timing_begin every
 * ngf_field_render brackets its ngf_march_kernel and ngf_colour_kernel launches with CUDA events recorded on the
 * launching stream (up to `capacity` march+colour pairs; 0 switches it off).  ngf_field_timing_read synchronises
 * the recorded events and returns the number of pairs timed and the summed durations in milliseconds, then rearms.
 */
int ngf_field_timing_begin(NgfField f, int32_t capacity);
int ngf_field_timing_read(NgfField f, int32_t* n_launches, double* march_ms, double* colour_ms);

/*
 * Point-wise queries (API parity with the reference's public methods).
 *  ngf_field_sample_ray : Base.sample_ray, eval branch (FieldBase.py:118-137): pts [R][S][3], t [R][S],
 *                         inside [R][S] (uint8 0/1).
 *  ngf_field_alpha_keep : AlphaGridMask.sample_alpha(...) > 0 (FieldBase.py:33-37): world pts [N][3] -> uint8.
 *  ngf_field_alpha_value: AlphaGridMask.sample_alpha(...) itself: the trilinear value of the {0,1} volume -> fp32 [N]
 *                         (1 everywhere when the field has no mask).
 *  ngf_field_gauge      : Base.normalize_coord is NOT applied: in = normalised xyz [N][3];
 *                         out xy/yz/xz [N][2] each = TriPlane.compute_gauge (Field.py:53-75) or
 *                         InfoInv transform (InfoInv/models/Field.py:43-50).
 *  ngf_field_density    : compute_density(xy, yz, xz) (Field.py:77-91 / InfoInv Field.py:52-70) -> sigma [N].
 *  ngf_field_rgb        : compute_rgb(xy, yz, xz, viewdirs) (Field.py:93-105 / InfoInv Field.py:72-89) -> [N][3].
 *  ngf_field_sigma_world: compute_alpha's inner part (FieldBase.py:140-156): world pts -> sigma with the alpha
 *                         mask applied and gauge optional (the reference passes iteration=-1, i.e. off).
 */
int ngf_field_sample_ray(NgfField f, const float* rays_dev, int64_t n_rays, int32_t ray_stride, int32_t n_samples,
                         float* pts_dev, float* t_dev, uint8_t* inside_dev, void* stream);
/* Base.sample_ray, is_train=True branch: as above with t_i = t_min + stepSize * (i + jitter[r]), jitter_dev [R]. */
int ngf_field_sample_ray_jitter(NgfField f, const float* rays_dev, int64_t n_rays, int32_t ray_stride, int32_t n_samples,
                                const float* jitter_dev, float* pts_dev, float* t_dev, uint8_t* inside_dev, void* stream);
int ngf_field_alpha_keep(NgfField f, const float* pts_dev, int64_t n, uint8_t* keep_dev, void* stream);
int ngf_field_alpha_value(NgfField f, const float* pts_dev, int64_t n, float* value_dev, void* stream);
int ngf_field_gauge(NgfField f, const float* xyz_norm_dev, int64_t n, int32_t gauge_on, float* xy_dev,
                    float* yz_dev, float* xz_dev, void* stream);
int ngf_field_density(NgfField f, const float* xy_dev, const float* yz_dev, const float* xz_dev, int64_t n,
                      float* sigma_dev, void* stream);
int ngf_field_rgb(NgfField f, const float* xy_dev, const float* yz_dev, const float* xz_dev,
                  const float* viewdirs_dev, int64_t n, float* rgb_dev, int32_t mlp_impl, void* stream);
int ngf_field_sigma_world(NgfField f, const float* pts_dev, int64_t n, int32_t use_gauge, float* sigma_dev,
                          void* stream);

/*
 * Frame post-processing on the device, replacing what evaluation() does on the CPU after
 * `.cpu()` (TriPlane/main.py:99-116): u8 = (rgb * 255).astype('uint8') for the PNG / video writers and
 * sse = sum((rgb - gt)^2), from which PSNR = -10 ln(sse / n_values) / ln 10 (main.py:105-106).  Either output may be
 * NULL.  Arrays are device pointers on the current device; sse_dev is one double, zeroed by the call.
 */
int ngf_frame_post(const float* rgb_dev, const float* gt_dev, int64_t n_values, uint8_t* u8_dev, double* sse_dev,
                   void* stream);

/* The depth image evaluation() writes next to every frame (TriPlane/main.py:102,168 -> visualize_depth_numpy,
 * utils.py:32-47, with minmax = near_far): x = nan_to_num(depth); x = (x - min) / (max - min + 1e-8);
 * cv2.applyColorMap((255 * x).astype(uint8), cv2.COLORMAP_JET) -> bgr_dev [n][3] uint8 (B, G, R as OpenCV returns). */
int ngf_depth_colormap(const float* depth_dev, int64_t n, double min_depth, double max_depth, uint8_t* bgr_dev,
                       void* stream);

/*
 * Multi-GPU helpers (the reference has no distributed code).  Rays of a frame are dealt to
 * ranks in interleaved blocks of `block` rays: global ray g belongs to rank (g / block) % world.
 *  ngf_shard_count   : number of rays rank owns.
 *  ngf_shard_gather  : dst[i] = src[global index of the i-th ray of rank] for a [n][width] fp32 device array
 *  ngf_shard_scatter : inverse of the all-gather: src is [world][max_shard][width] (rank-major, as
 *                      ncclAllGather leaves it), dst [n][width] in frame order.
 */
int64_t ngf_shard_count(int64_t n_rays, int32_t block, int32_t rank, int32_t world);
int ngf_shard_gather(const float* src_dev, int64_t n_rays, int32_t width, int32_t block, int32_t rank,
                     int32_t world, float* dst_dev, void* stream);
int ngf_shard_scatter(const float* src_dev, int64_t n_rays, int32_t width, int32_t block, int32_t world,
                      int64_t max_shard, float* dst_dev, void* stream);

/*
 * Ray-sharded multi-GPU frames with the frame all-gather inside the library ("ngf_comm_init /
 * ngf_frame_allgather").  The reference has nothing here — its only multi-GPU code is the vestigial
 * nn.DataParallel(NeuTex) at UV-Mapping/model/model.py:285 — so the contract is the survey's: one process per GPU, rays
 * of a batch dealt to ranks in interleaved `block`-ray blocks (ngf_shard_count), every rank renders its blocks, ONE
 * all-gather per batch leaves the whole [n_rays][4] fp32 (r, g, b, depth) batch in frame order on every rank.
 *
 * The exchange runs over NVLink peer mappings of the ranks' frame buffers (CUDA IPC), not through a NCCL kernel:
 *   NGF_COMM_COPY   the render's last kernel writes the rank's rows in frame order; one strided copy per peer on the
 *                   copy engines lands them in every peer's buffer (no SM is used by the collective);
 *   NGF_COMM_STORE  the render's last kernel stores each finished row into every rank's buffer itself (the all-gather
 *                   is the epilogue of the render).
 * Completion / back-pressure are device-side step counters polled by one-CTA kernels; the host never blocks.
 *
 *   ngf_comm_init     allocate this rank's `n_slots` (2..4) frame buffers + flags on `device`.
 *   ngf_comm_export   write ngf_comm_handle_bytes() opaque bytes to hand to every other rank (any transport:
 *                     torch.distributed all_gather, MPI, a file).
 *   ngf_comm_connect  blobs = the world's exported bytes, rank-major; maps the peers' buffers.  Needs peer access
 *                     between the devices (NGF_EUNSUPPORTED otherwise).
 *   ngf_field_render_sharded
 *                     render this rank's rays (local order = its blocks, in order; n_local == ngf_comm_local_rays) of
 *                     the next batch on `stream` and start the exchange; returns a ticket.  At most n_slots batches may
 *                     be unreleased.
 *   ngf_frame_allgather
 *                     make `stream` wait until every rank's rows of `ticket` have landed; *frame_dev is then the whole
 *                     batch [n_rays][4] in frame order on this rank (valid until ngf_frame_release + n_slots - 1 more
 *                     batches).
 *   ngf_frame_release the consumer on `stream` is done with the buffer: peers may overwrite it.
 *   ngf_field_render_sharded_host_async
 *                     the same through HOST buffers, pipelined over internal streams like ngf_field_render_host_async:
 *                     H2D of the rank's rays, render, exchange, D2H of rows [first_row, first_row + n_rows) of the
 *                     gathered batch into frame_host [n_rows][4]; the buffer is released internally.  Wait with
 *                     ngf_comm_wait(ticket); NGF_ECOMM if a peer never answered (NGF_COMM_TIMEOUT_S, default 10 s).
 */
typedef enum NgfCommMode { NGF_COMM_COPY = 0, NGF_COMM_STORE = 1 } NgfCommMode;
typedef struct NgfComm_* NgfComm;
int ngf_comm_init(int32_t rank, int32_t world, int32_t device, int64_t n_rays, int32_t block, int32_t n_slots,
                  int32_t mode, NgfComm* out);
int64_t ngf_comm_handle_bytes(void);
int ngf_comm_export(NgfComm c, void* blob);
int ngf_comm_connect(NgfComm c, const void* blobs);
void ngf_comm_free(NgfComm c);
int64_t ngf_comm_local_rays(NgfComm c);
int ngf_field_render_sharded(NgfField f, NgfComm c, const float* rays_local_dev, int64_t n_local, int32_t ray_stride,
                             int32_t n_samples, int32_t white_bg, int32_t tile_w, int32_t mlp_impl, void* stream,
                             uint64_t* ticket);
int ngf_frame_allgather(NgfComm c, uint64_t ticket, void* stream, const float** frame_dev);
int ngf_frame_release(NgfComm c, uint64_t ticket, void* stream);
int ngf_field_render_sharded_host_async(NgfField f, NgfComm c, const float* rays_local_host, int64_t n_local,
                                        int32_t ray_stride, int32_t n_samples, int32_t white_bg, int32_t tile_w,
                                        int32_t mlp_impl, float* frame_host, int64_t first_row, int64_t n_rows,
                                        uint64_t* ticket);
int ngf_comm_wait(NgfComm c, uint64_t ticket);
/* Camera batches (evaluation_path's per-frame job, TriPlane/main.py:155-161, for a ray-sharded batch of frames): the batch is
 * n_frames (<= 64) frames of one pinhole model (`camera`: intrinsics and image size; its c2w is ignored) with one pose each,
 * poses [n_frames][12] = row-major [3][4] camera-to-world; n_frames * width * height must be the comm's n_rays.  Rays are
 * generated in the march kernel from the pixel index, so a step's input is 48 bytes per frame.
 *   ngf_field_render_sharded_camera                 device-resident: as ngf_field_render_sharded (ticket for
 *                                                   ngf_frame_allgather / ngf_frame_release), poses in device memory
 *   ngf_field_render_sharded_camera_u8_host_async   poses from host memory in; rows [first_row, first_row + n_rows) of the
 *                                                   gathered batch out as uint8 rgb [n_rows][3] ((rgb * 255) truncated,
 *                                                   main.py:116); wait with ngf_comm_wait */
int ngf_field_render_sharded_camera(NgfField f, NgfComm c, const NgfCamera* camera, const float* poses_dev, int32_t n_frames,
                                    int32_t n_samples, int32_t white_bg, int32_t mlp_impl, void* stream, uint64_t* ticket);
int ngf_field_render_sharded_camera_u8_host_async(NgfField f, NgfComm c, const NgfCamera* camera, const float* poses_host,
                                                  int32_t n_frames, int32_t n_samples, int32_t white_bg, int32_t mlp_impl,
                                                  uint8_t* u8_host, int64_t first_row, int64_t n_rows, uint64_t* ticket);

/* =====================================================================================================
 * UV-Mapping (NeuTex) render path.  Reference (paths relative to /root/reference/UV-Mapping):
 * NeuTex.forward (model/model.py:27-59) = cube_ray_generation (model/renderer.py:79-141) -> GeometryMlpDecoder
 * (model/decoder.py:201-237) -> GaugeTransform (model/gauge_fields.py:8-74) -> TextureMlpDecoder
 * (model/decoder.py:11-121) -> ray_march / simple_tone_map (model/renderer.py:4-11,176-247), as called per 1024-ray
 * chunk by test.py:108-114 through Model.test (model/model.py:362-373).  primitive_type 'square' or 'sphere'.
 * ===================================================================================================== */
typedef struct NgfNeutexDesc {
  NgfLinear geometry[12];   /* net_geometry_decoder.block.{0,2,...,22}: 63->256, 10 x 256->256, 256->1          */
  NgfLinear gauge[5];       /* gauge_transform.encoder.{linear1, linear2, linear_list.0, linear_list.1, last_linear} */
  NgfLinear tex_block1[6];  /* net_texture.block1.{0,2,...,10}: 42->256, 5 x 256->256                            */
  NgfLinear tex_color1;     /* net_texture.color1: 256->3                                                        */
  NgfLinear tex_block2[5];  /* net_texture.block2.{0,2,4,6,8}: 295->256, 3 x 256->256, 256->3                    */
  int32_t sample_num;       /* opt.sample_num (64)                                                               */
  float jitter;             /* 0.05, hard-coded at model/model.py:30                                             */
  const float* texture;     /* TextureMlpDecoder.cubemap_ after load_square: [h][w][c] fp32 in [0,1], or NULL    */
  int32_t tex_h, tex_w, tex_c;
  int32_t primitive;        /* opt.primitive_type: 0 = 'square' (gauge 128->2, uv = tanh, tex_block1[0] 42->256),
                               1 = 'sphere' (gauge 128->3, uv = normalize, tex_block1[0] 63->256; gauge_fields.py:53-56,71-74,
                               model.py:22); the sphere's edited-texture (cube map) branch is not built */
} NgfNeutexDesc;

typedef struct NgfNeutex_* NgfNeutex;

/* Pack the three MLP stacks into tcgen05 operand order (fp16; the gauge network as hi+lo split fp16).  Parameter
 * pointers may be host or device memory.  Synchronous. */
int ngf_neutex_pack(const NgfNeutexDesc* desc, int device, NgfNeutex* out);
void ngf_neutex_free(NgfNeutex h);

/*
 * Render rays of ONE camera: replaces NeuTex.forward(camera_position[1,3], ray_direction[1,R,3], background_color[1,3])
 * -> output["color"] [1,R,3], output["transmittance"] [1,R].
 *   noise_dev  [R][64] U[0,1) numbers the reference draws with torch.rand inside cube_ray_generation
 *              (renderer.py:113-118), or NULL for no jitter
 *   background_dev  [3] or NULL (model.py:48-49)
 */
int ngf_neutex_render(NgfNeutex h, const float* campos_dev, const float* raydir_dev, const float* background_dev,
                      const float* noise_dev, int64_t n_rays, float* color_dev, float* transmittance_dev, void* stream);
/* Same through HOST buffers (what test.py's chunk loop does with model.set_input / .cpu()); returns when the results
 * are in color_host / transmittance_host. */
int ngf_neutex_render_host(NgfNeutex h, const float* campos_host, const float* raydir_host, const float* background_host,
                           const float* noise_host, int64_t n_rays, float* color_host, float* transmittance_host);
/* The same two calls with the jitter numbers drawn on the device instead of uploaded: sample i of frame ray r uses
 * U = philox4x32-10(key = seed, index = r * 64 + i) >> 8 scaled to [0,1) — the role of the torch.rand call inside
 * cube_ray_generation (renderer.py:113-118), without the 256 B per ray of host-drawn numbers.  `first_ray` is the frame
 * index of raydir_dev[0] (a frame rendered in several calls draws the same numbers as in one).  ngf_neutex_noise writes
 * the numbers a seeded render of rays [first_ray, first_ray + n_rays) uses into noise_dev[n_rays][sample_num], so that a
 * checker can hand the reference the identical jitter. */
int ngf_neutex_render_seeded(NgfNeutex h, const float* campos_dev, const float* raydir_dev, const float* background_dev,
                             uint64_t seed, int64_t first_ray, int64_t n_rays, float* color_dev, float* transmittance_dev,
                             void* stream);
int ngf_neutex_render_host_seeded(NgfNeutex h, const float* campos_host, const float* raydir_host,
                                  const float* background_host, uint64_t seed, int64_t n_rays, float* color_host,
                                  float* transmittance_host);
int ngf_neutex_noise(NgfNeutex h, uint64_t seed, int64_t first_ray, int64_t n_rays, float* noise_dev, void* stream);
/* Arithmetic of the three MLP stacks: 0 (default) = tcgen05 tensor cores, fp16 operands / fp32 accumulation with the
 * gauge network and the first geometry layer in split fp16; 1 = plain fp32 on the CUDA cores from the unpacked
 * parameters (several times slower; the fall-back for checkpoints whose activations leave fp16's range or precision).
 * The environment variable NGF_NTX_FP32=1 selects 1 at pack time. */
int ngf_neutex_set_precision(NgfNeutex h, int32_t mode);
/* Evaluate n_points seeded random in-cube points with random unit view directions through BOTH arithmetic paths and
 * report, in report[4] (host): max |sigma_tc - sigma_fp32| / (1 + |sigma_fp32|), max |rgb_tc - rgb_fp32|, the mean of
 * that rgb deviation, and max |rgb_fp32|.  What a caller runs after loading a trained checkpoint to decide whether the
 * fp16 path is good enough for it.  Synchronous; overwrites the per-sample workspace of the last render. */
int ngf_neutex_self_check(NgfNeutex h, int32_t n_points, uint64_t seed, float* report, void* stream);
/* In-cube samples the last render evaluated (synchronises `stream`). */
int ngf_neutex_last_valid_samples(NgfNeutex h, uint64_t* n_valid, void* stream);
/* Per-sample view of the last render (tests): (sigma, r, g, b) of samples [first_sample, first_sample+n) — defined only
 * where the corresponding valid_mask bit is set — and the per-ray 64-bit in-cube masks (bit i = sample i), i.e. the
 * density / radiance / valid tensors NeuTex.forward hands to ray_march (model.py:40-47).  Synchronises the device. */
int ngf_neutex_copy_samples(NgfNeutex h, int64_t first_sample, int64_t n, float* sigma_rgb_host, uint64_t* valid_mask_host,
                            int64_t first_ray, int64_t n_mask_rays);
/* CUDA-event timing of the three kernels of a render (ray generation | MLP | march), as ngf_field_timing_*. */
int ngf_neutex_timing_begin(NgfNeutex h, int32_t capacity);
int ngf_neutex_timing_read(NgfNeutex h, int32_t* n_renders, double* raygen_ms, double* mlp_ms, double* march_ms);

/* Environment read by ngf_neutex_pack: NGF_NTX_CG=2 packs per-CTA-rank weight streams and renders with CTA pairs
 * (cta_group::2 MMAs, M = 256) instead of one CTA per 256-sample tile (default 1; same results within fp32 rounding of the
 * accumulation order); NGF_NTX_DBG is a profiling switch (2: skip the MMAs, 4: record a timeline, 8: skip the weight copies
 * — results are then meaningless). */
/* Profiling aid (library built as is, NGF_NTX_DBG=4 in the environment when packing): per-layer clock64 stamps of CTA 0's
 * first tile — [25][4] = MMA warp saw a_ready | MMA warp issued the layer | worker 0 saw acc_ready | worker 0 finished the
 * epilogue — followed by 64 stamps of the weight ring during layer 5 ([2i], [2i+1] = MMA warp saw stage i full | issued its
 * MMAs; [32+i] = producer saw the stage's slot empty).  out_host holds 164 values. */
int ngf_neutex_debug_trace(NgfNeutex h, long long* out_host);

#ifdef __cplusplus
}
#endif
#endif /* NGF_B200_H_ */
