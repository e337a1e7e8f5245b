// ngf_comm.cu — ray-sharded multi-GPU frames inside the C ABI (SURVEY.md §8b "ngf_comm_init / ngf_frame_allgather",
// §8e).  One process per GPU; the reference has no distributed code (nn.DataParallel at UV-Mapping/model/model.py:285
// is vestigial), so the contract is the survey's: rays are dealt to ranks in interleaved blocks, every rank renders its
// blocks and ONE all-gather per batch leaves the whole batch, in frame order, on every rank.
//
// The all-gather uses no NCCL kernel and (by default) no SM at all.  Every rank owns `n_slots` frame buffers of
// [n_rays][4] fp32 (r, g, b, depth) in one cudaMalloc allocation that is exported with cudaIpcGetMemHandle and mapped
// by every peer (NVLink 5 / NVSwitch peer access).  ngf_finalize_shard_kernel writes the rank's finished rows straight
// into frame order; then
//   mode NGF_COMM_COPY   one cudaMemcpy2DAsync per peer (row = block * 16 B, pitch = world * block * 16 B) on the copy
//                        engines pushes the rows into the same positions of every peer's buffer;
//   mode NGF_COMM_STORE  the finalize kernel itself stores every row into all peers' buffers (st.global over NVLink):
//                        the all-gather is the epilogue of the render.
// Completion and back-pressure are step counters in the same allocation, written remotely and polled locally:
//   arrived[slot][src]   src's rows of step k are in my buffer  (src stores k+1 after its copies / stores)
//   freed[slot][peer]    peer has finished reading its own buffer of step k (I may overwrite it with step k+n_slots)
// written by a one-CTA signal kernel (after __threadfence_system) and awaited by a one-CTA wait kernel on the consuming
// stream, so the host never blocks and no rank needs a host-side barrier.  A wait that sees nothing for ~10 s sets
// an error word (pinned host memory) and gives up instead of wedging the device.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

#include "ngf_handle.h"

using namespace ngf;

#define CU(expr)                                                                                                   \
  do {                                                                                                             \
    cudaError_t _e = (expr);                                                                                       \
    if (_e != cudaSuccess)                                                                                         \
      return ngf_set_error(NGF_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);        \
  } while (0)

namespace {

constexpr int kMaxWorld = 16;
constexpr int kMaxSlots = 4;
constexpr int kCommStreams = 2;
constexpr int kCompStreams = 3;          // compute streams of the host pipelines: batch k renders on stream (k % n_slots) % 3
constexpr int kMaxFrames = 64;                // frames per camera batch
constexpr size_t kFlagBytes = 4096;           // arrived [kMaxSlots][kMaxWorld] u32 @0, freed [kMaxSlots][kMaxWorld] u32 @1024

struct CommBlob {                             // what ranks exchange (ngf_comm_export / ngf_comm_connect)
  cudaIpcMemHandle_t mem;
  int32_t rank, world, device, block, n_slots, pid;
  int64_t n_rays;
  uint64_t base;                              // the allocation's address: used instead of the IPC handle by ranks that
                                              // live in the exporting process (single-process tests)
  uint32_t magic;
  uint32_t pad;
};
static_assert(sizeof(CommBlob) <= 128, "blob size");
constexpr uint32_t kMagic = 0x4e474643u;      // "NGFC"

struct PeerList {
  uint32_t* p[kMaxWorld];
  int n;
};

// store `value` into every listed flag (the flags live in the peers' allocations)
__global__ void ngf_comm_signal_kernel(const __grid_constant__ PeerList pl, uint32_t value) {
  const int i = threadIdx.x;
  if (i < pl.n) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(pl.p[i]) = value;
  }
}

// wait until every listed (local) flag has reached `want` (wrap-safe); give up after `timeout` cycles
__global__ void ngf_comm_wait_kernel(const __grid_constant__ PeerList pl, uint32_t want, long long timeout,
                                     uint32_t* __restrict__ err) {
  const int i = threadIdx.x;
  if (i < pl.n) {
    const volatile uint32_t* f = reinterpret_cast<const volatile uint32_t*>(pl.p[i]);
    const long long t0 = clock64();
    while ((int32_t)(*f - want) < 0) {
      __nanosleep(200);
      if (clock64() - t0 > timeout) { *reinterpret_cast<volatile uint32_t*>(err) = 1u + (uint32_t)i; break; }
    }
    __threadfence_system();
  }
}

// rows [first, first + n) of a gathered [.][4] frame -> uint8 rgb, (rgb * 255) truncated (TriPlane/main.py:116)
__global__ void ngf_frame4_u8_kernel(const float4* __restrict__ frame, long long first, long long n, uint8_t* __restrict__ u8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = frame[first + i];
    u8[i * 3 + 0] = (uint8_t)__fmul_rn(v.x, 255.f);
    u8[i * 3 + 1] = (uint8_t)__fmul_rn(v.y, 255.f);
    u8[i * 3 + 2] = (uint8_t)__fmul_rn(v.z, 255.f);
  }
}

}  // namespace

struct NgfComm_ {
  int rank = 0, world = 1, device = 0, block = 1, n_slots = 3, mode = NGF_COMM_COPY;
  long long n_rays = 0;                       // rays of a whole batch (all ranks)
  long long n_local = 0;                      // rays this rank renders
  size_t frame_bytes = 0, bytes = 0;
  uint8_t* base = nullptr;                    // [flags | frame 0 | frame 1 | ...]
  uint8_t* peer[kMaxWorld] = {};              // peer[rank] == base
  bool peer_ipc[kMaxWorld] = {};              // mapped with cudaIpcOpenMemHandle (to be closed)
  bool connected = false;
  uint32_t* err_host = nullptr;               // pinned, mapped
  uint32_t* err_dev = nullptr;
  long long timeout_cycles = 0;
  cudaStream_t s_in = nullptr, s_comp[kCompStreams] = {}, s_out = nullptr, s_comm[kCommStreams] = {};
  struct Slot {
    cudaEvent_t ev_in = nullptr;              // rays uploaded (host path)
    cudaEvent_t ev_rendered = nullptr;        // my rows are in my frame buffer
    cudaEvent_t ev_sent[kCommStreams] = {};   // my rows have been pushed to the peers (copy mode)
    cudaEvent_t ev_consumed = nullptr;        // the consumer has finished reading the frame buffer
    cudaEvent_t ev_done = nullptr;            // host path: results are in the caller's host buffer
    float* rays = nullptr;                    // host path staging
    float* poses = nullptr;                   // camera path: [kMaxFrames][12] c2w of the batch's frames
    uint8_t* u8 = nullptr;                    // camera path: uint8 image rows on their way to the host
    long long u8_cap = 0;
    unsigned long long ticket = 0;            // ticket occupying the slot (0 = none)
    bool released = true;
    // render workspace of the slot: batches in different slots may be rendered on different streams and overlap (the
    // march of one beside the colour pass of another)
    float *rgb = nullptr, *depth = nullptr, *acc = nullptr;
    unsigned int* counters = nullptr;
    QEntry* queue = nullptr;
    long long queue_cap = 0;
  } slot[kMaxSlots];
  long long rays_cap = 0;
  int rays_stride = 0;
  bool ws_ready = false;
  unsigned long long next_step = 0;

  float4* frame(int q, int s) const { return reinterpret_cast<float4*>(peer[q] + kFlagBytes + (size_t)s * frame_bytes); }
  uint32_t* arrived(int q, int s, int src) const { return reinterpret_cast<uint32_t*>(peer[q]) + s * kMaxWorld + src; }
  uint32_t* freed(int q, int s, int who) const { return reinterpret_cast<uint32_t*>(peer[q] + 1024) + s * kMaxWorld + who; }
};

namespace {

struct Guard {
  int prev = -1;
  explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
  ~Guard() { int cur = -1; if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev); }
};

void comm_destroy(NgfComm_* c) {
  Guard g(c->device);
  cudaDeviceSynchronize();
  for (int q = 0; q < c->world; ++q)
    if (q != c->rank && c->peer[q] && c->peer_ipc[q]) cudaIpcCloseMemHandle(c->peer[q]);
  for (auto& s : c->slot) {
    if (s.ev_in) cudaEventDestroy(s.ev_in);
    if (s.ev_rendered) cudaEventDestroy(s.ev_rendered);
    for (auto& e : s.ev_sent) if (e) cudaEventDestroy(e);
    if (s.ev_consumed) cudaEventDestroy(s.ev_consumed);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    cudaFree(s.rays); cudaFree(s.poses); cudaFree(s.u8);
    cudaFree(s.rgb); cudaFree(s.depth); cudaFree(s.acc); cudaFree(s.counters); cudaFree(s.queue);
  }
  if (c->s_in) cudaStreamDestroy(c->s_in);
  for (auto& st : c->s_comp) if (st) cudaStreamDestroy(st);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  for (auto& s : c->s_comm) if (s) cudaStreamDestroy(s);
  cudaFree(c->base);
  if (c->err_host) cudaFreeHost(c->err_host);
  delete c;
}

// peers served by comm stream j (copy mode): every kCommStreams-th rank after mine, so ranks do not all hit one target
PeerList stream_peers(const NgfComm_* c, int j, int s, bool arrived_flags, int own_index_rank) {
  PeerList pl{};
  for (int d = 1; d < c->world; ++d) {
    if ((d - 1) % kCommStreams != j) continue;
    const int q = (c->rank + d) % c->world;
    pl.p[pl.n++] = arrived_flags ? c->arrived(q, s, own_index_rank) : c->freed(q, s, own_index_rank);
  }
  return pl;
}

int launch_signal(const PeerList& pl, uint32_t value, cudaStream_t st) {
  if (pl.n == 0) return NGF_OK;
  ngf_comm_signal_kernel<<<1, 32, 0, st>>>(pl, value);
  count_launch();
  CU(cudaGetLastError());
  return NGF_OK;
}

int launch_wait(const NgfComm_* c, const PeerList& pl, uint32_t want, cudaStream_t st) {
  if (pl.n == 0) return NGF_OK;
  ngf_comm_wait_kernel<<<1, 32, 0, st>>>(pl, want, c->timeout_cycles, c->err_dev);
  count_launch();
  CU(cudaGetLastError());
  return NGF_OK;
}

// flags of MY allocation that the other ranks write: arrived[s][q] (q != rank) or freed[s][q]
PeerList local_flags(const NgfComm_* c, int s, bool arrived_flags, int j = -1) {
  PeerList pl{};
  for (int d = 1; d < c->world; ++d) {
    if (j >= 0 && (d - 1) % kCommStreams != j) continue;
    const int q = (c->rank + d) % c->world;
    pl.p[pl.n++] = arrived_flags ? c->arrived(c->rank, s, q) : c->freed(c->rank, s, q);
  }
  return pl;
}

int check_ready(const NgfComm_* c) {
  if (!c) return ngf_set_error(NGF_EINVAL, "comm is NULL");
  if (c->world > 1 && !c->connected) return ngf_set_error(NGF_EINVAL, "comm is not connected: call ngf_comm_connect first");
  return NGF_OK;
}

int check_err(const NgfComm_* c) {
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(c->err_host);
  if (e) return ngf_set_error(NGF_ECOMM, "rank %d: timed out waiting for a peer flag (wait list entry %u): a peer is "
                              "stalled or has exited", c->rank, e - 1);
  return NGF_OK;
}

// Steps 2-3 of a sharded frame on stream `st`: render my rays into frame order of slot s, then hand the rows to the peers.
int render_and_push(NgfField f, NgfComm_* c, int s, unsigned long long k, const float* rays_dev, long long n_local,
                    int ray_stride, int n_samples, int white_bg, int tile_w, int mlp_impl, cudaStream_t st,
                    const CamDev* cam = nullptr) {
  NgfComm_::Slot& sl = c->slot[s];
  // my rows of slot s may be overwritten once (a) the local consumer of step k - n_slots has released the buffer and
  // (b) the copies that pushed them to the peers have finished reading them
  CU(cudaStreamWaitEvent(st, sl.ev_consumed, 0));
  for (int j = 0; j < kCommStreams; ++j) CU(cudaStreamWaitEvent(st, sl.ev_sent[j], 0));
  // (c) the slot's render workspace: its previous batch may have been rendered on another stream
  CU(cudaStreamWaitEvent(st, sl.ev_rendered, 0));
  ShardOut so{};
  so.block = c->block; so.rank = c->rank; so.world = c->world;
  so.dst[so.n_dst++] = c->frame(c->rank, s);
  const uint32_t want_freed = k >= (unsigned long long)c->n_slots ? (uint32_t)(k - c->n_slots + 1) : 0u;
  if (c->mode == NGF_COMM_STORE && c->world > 1) {
    // the finalize kernel stores into the peers' buffers: they must have been released by their owners first
    if (want_freed) { int rc = launch_wait(c, local_flags(c, s, false), want_freed, st); if (rc) return rc; }
    for (int d = 1; d < c->world; ++d) so.dst[so.n_dst++] = c->frame((c->rank + d) % c->world, s);
  }
  int rc = ngf_render_dev(f, rays_dev, n_local, ray_stride, n_samples, white_bg, tile_w, sl.rgb, sl.depth, sl.acc,
                          sl.counters, &sl.queue, &sl.queue_cap, mlp_impl, st, cam, nullptr, &so);
  if (rc) return rc;
  if (c->mode == NGF_COMM_STORE && c->world > 1) {
    PeerList pl{};
    for (int d = 1; d < c->world; ++d) pl.p[pl.n++] = c->arrived((c->rank + d) % c->world, s, c->rank);
    rc = launch_signal(pl, (uint32_t)(k + 1), st);
    if (rc) return rc;
  }
  CU(cudaEventRecord(sl.ev_rendered, st));
  if (c->mode == NGF_COMM_COPY && c->world > 1) {
    // rows of mine: blocks rank, rank + world, ...; all full except possibly the very last block of the batch
    const long long nb_total = (c->n_rays + c->block - 1) / c->block;
    const long long nb_mine = nb_total > c->rank ? (nb_total - c->rank + c->world - 1) / c->world : 0;
    const bool last_partial = c->n_rays % c->block != 0;
    const bool last_mine = nb_total > 0 && (nb_total - 1) % c->world == c->rank;
    const long long full = nb_mine - ((last_partial && last_mine) ? 1 : 0);
    const size_t row = (size_t)c->block * 16, pitch = row * c->world, off = (size_t)c->rank * row;
    for (int j = 0; j < kCommStreams; ++j) {
      cudaStream_t cs = c->s_comm[j];
      const PeerList targets = stream_peers(c, j, s, true, c->rank);
      if (targets.n == 0) continue;
      CU(cudaStreamWaitEvent(cs, sl.ev_rendered, 0));
      if (want_freed) { rc = launch_wait(c, local_flags(c, s, false, j), want_freed, cs); if (rc) return rc; }
      for (int d = 1; d < c->world; ++d) {
        if ((d - 1) % kCommStreams != j) continue;
        const int q = (c->rank + d) % c->world;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(c->frame(c->rank, s)) + off;
        uint8_t* dst = reinterpret_cast<uint8_t*>(c->frame(q, s)) + off;
        if (full > 0) CU(cudaMemcpy2DAsync(dst, pitch, src, pitch, row, (size_t)full, cudaMemcpyDeviceToDevice, cs));
        if (last_partial && last_mine) {
          const size_t o2 = (size_t)(nb_total - 1) * row - off;      // offset of the last block relative to my first
          CU(cudaMemcpyAsync(dst + o2, src + o2, (size_t)(c->n_rays % c->block) * 16, cudaMemcpyDeviceToDevice, cs));
        }
      }
      rc = launch_signal(targets, (uint32_t)(k + 1), cs);
      if (rc) return rc;
      CU(cudaEventRecord(sl.ev_sent[j], cs));
    }
  }
  return NGF_OK;
}

int ensure_workspace(NgfComm_* c, long long n_local) {
  if (c->ws_ready) return NGF_OK;
  const size_t n = (size_t)(n_local > 0 ? n_local : 1);
  for (int i = 0; i < c->n_slots; ++i) {
    NgfComm_::Slot& sl = c->slot[i];
    CU(cudaMalloc(reinterpret_cast<void**>(&sl.rgb), n * 3 * sizeof(float)));
    CU(cudaMalloc(reinterpret_cast<void**>(&sl.depth), n * sizeof(float)));
    CU(cudaMalloc(reinterpret_cast<void**>(&sl.acc), n * sizeof(float)));
    CU(cudaMalloc(reinterpret_cast<void**>(&sl.counters), kCounterBytes));
    CU(cudaMemset(sl.counters, 0, kCounterBytes));
  }
  c->ws_ready = true;
  return NGF_OK;
}

int begin_step(NgfField f, NgfComm_* c, long long n_local, int mlp_impl, int* slot_out, unsigned long long* k_out) {
  if (!f) return ngf_set_error(NGF_EINVAL, "field is NULL");
  int rc = check_ready(c);
  if (rc) return rc;
  if (f->device != c->device) return ngf_set_error(NGF_EINVAL, "field is on device %d, comm on device %d", f->device, c->device);
  if (n_local != c->n_local)
    return ngf_set_error(NGF_EINVAL, "rank %d of %d renders %lld rays of a %lld-ray batch in %d-ray blocks, got %lld", c->rank,
                         c->world, c->n_local, c->n_rays, c->block, n_local);
  if (mlp_impl != NGF_MLP_TCGEN05 && mlp_impl != NGF_MLP_SIMT) return ngf_set_error(NGF_EINVAL, "mlp_impl=%d", mlp_impl);
  if ((rc = check_err(c))) return rc;
  const unsigned long long k = c->next_step;
  const int s = (int)(k % c->n_slots);
  if (c->slot[s].ticket && !c->slot[s].released)
    return ngf_set_error(NGF_EINVAL, "frame buffer %d still holds ticket %llu: release it (ngf_frame_release) before submitting "
                         "%d more batches", s, c->slot[s].ticket, c->n_slots);
  if ((rc = ensure_workspace(c, c->n_local))) return rc;
  *slot_out = s;
  *k_out = k;
  return NGF_OK;
}

}  // namespace

extern "C" {

int ngf_comm_init(int32_t rank, int32_t world, int32_t device, int64_t n_rays, int32_t block, int32_t n_slots,
                  int32_t mode, NgfComm* out) {
  if (!out) return ngf_set_error(NGF_EINVAL, "out is NULL");
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return ngf_set_error(NGF_EINVAL, "rank %d of world %d (max %d)", rank, world, kMaxWorld);
  if (n_rays < 1 || n_rays > 0x7fffffffll || block < 1) return ngf_set_error(NGF_EINVAL, "n_rays=%lld block=%d", (long long)n_rays, block);
  if (n_slots < 2 || n_slots > kMaxSlots) return ngf_set_error(NGF_EINVAL, "n_slots=%d (2..%d)", n_slots, kMaxSlots);
  if (mode != NGF_COMM_COPY && mode != NGF_COMM_STORE) return ngf_set_error(NGF_EINVAL, "mode=%d", mode);
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return ngf_set_error(NGF_EINVAL, "device %d out of range (%d visible)", device, ndev);
  Guard g(device);
  NgfComm_* c = new NgfComm_();
  c->rank = rank; c->world = world; c->device = device; c->block = block; c->n_slots = n_slots; c->mode = mode;
  c->n_rays = n_rays;
  c->n_local = ngf_shard_count(n_rays, block, rank, world);
  c->frame_bytes = (((size_t)n_rays * 16) + 255) / 256 * 256;
  c->bytes = kFlagBytes + c->frame_bytes * n_slots;
  int khz = 1965000;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  const char* te = getenv("NGF_COMM_TIMEOUT_S");
  const double secs = te && atof(te) > 0 ? atof(te) : 10.0;
  c->timeout_cycles = (long long)(secs * 1e3 * khz);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->base), c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->base, 0, c->bytes);
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&c->err_host), 64, cudaHostAllocMapped);
  if (e == cudaSuccess) { memset(c->err_host, 0, 64); e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->err_dev), c->err_host, 0); }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
  for (auto& st : c->s_comp) if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
  for (int j = 0; j < kCommStreams && e == cudaSuccess; ++j) e = cudaStreamCreateWithFlags(&c->s_comm[j], cudaStreamNonBlocking);
  for (int s = 0; s < n_slots && e == cudaSuccess; ++s) {
    NgfComm_::Slot& sl = c->slot[s];
    e = cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.ev_rendered, cudaEventDisableTiming);
    for (int j = 0; j < kCommStreams && e == cudaSuccess; ++j) e = cudaEventCreateWithFlags(&sl.ev_sent[j], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.ev_consumed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    comm_destroy(c);
    cudaGetLastError();
    return ngf_set_error(NGF_ECUDA, "ngf_comm_init: %s", cudaGetErrorString(e));
  }
  c->peer[rank] = c->base;
  c->connected = world == 1;
  *out = c;
  return NGF_OK;
}

int64_t ngf_comm_handle_bytes(void) { return 128; }

int ngf_comm_export(NgfComm c, void* blob) {
  if (!c || !blob) return ngf_set_error(NGF_EINVAL, "NULL argument");
  Guard g(c->device);
  CommBlob b{};
  CU(cudaIpcGetMemHandle(&b.mem, c->base));
  b.rank = c->rank; b.world = c->world; b.device = c->device; b.block = c->block; b.n_slots = c->n_slots;
  b.pid = (int32_t)getpid(); b.n_rays = c->n_rays; b.magic = kMagic;
  b.base = (uint64_t)(uintptr_t)c->base;
  memset(blob, 0, 128);
  memcpy(blob, &b, sizeof(b));
  return NGF_OK;
}

int ngf_comm_connect(NgfComm c, const void* blobs) {
  if (!c || !blobs) return ngf_set_error(NGF_EINVAL, "NULL argument");
  if (c->connected) return NGF_OK;
  Guard g(c->device);
  for (int q = 0; q < c->world; ++q) {
    CommBlob b;
    memcpy(&b, static_cast<const uint8_t*>(blobs) + (size_t)q * 128, sizeof(b));
    if (b.magic != kMagic || b.rank != q || b.world != c->world || b.n_rays != c->n_rays || b.block != c->block ||
        b.n_slots != c->n_slots)
      return ngf_set_error(NGF_EINVAL, "handle %d does not describe rank %d of the same batch geometry", q, q);
    if (q == c->rank) continue;
    if (b.pid == (int32_t)getpid()) {
      // the peer rank lives in this process (tests drive several ranks from one process): its allocation is directly
      // addressable, on the same device or after enabling peer access
      if (b.device != c->device) {
        int can_p = 0;
        CU(cudaDeviceCanAccessPeer(&can_p, c->device, b.device));
        if (!can_p) return ngf_set_error(NGF_EUNSUPPORTED, "device %d cannot access device %d", c->device, b.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
        cudaGetLastError();
      }
      c->peer[q] = reinterpret_cast<uint8_t*>((uintptr_t)b.base);
      continue;
    }
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, c->device, b.device));
    if (!can) return ngf_set_error(NGF_EUNSUPPORTED, "device %d cannot access device %d (no NVLink / PCIe peer path)", c->device, b.device);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, b.mem, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return ngf_set_error(NGF_ECUDA, "cudaIpcOpenMemHandle(rank %d): %s", q, cudaGetErrorString(e));
    }
    c->peer[q] = static_cast<uint8_t*>(p);
    c->peer_ipc[q] = true;
  }
  c->connected = true;
  return NGF_OK;
}

void ngf_comm_free(NgfComm c) {
  if (c) comm_destroy(c);
}

int64_t ngf_comm_local_rays(NgfComm c) { return c ? c->n_local : -1; }

int ngf_field_render_sharded(NgfField f, NgfComm c, const float* rays_local_dev, int64_t n_local, int32_t ray_stride,
                             int32_t n_samples, int32_t white_bg, int32_t tile_w, int32_t mlp_impl, void* stream,
                             uint64_t* ticket) {
  if (!ticket) return ngf_set_error(NGF_EINVAL, "ticket is NULL");
  int s = 0;
  unsigned long long k = 0;
  int rc = begin_step(f, c, n_local, mlp_impl, &s, &k);
  if (rc) return rc;
  if (n_local > 0 && !rays_local_dev) return ngf_set_error(NGF_EINVAL, "rays is NULL");
  if (ray_stride < 6) return ngf_set_error(NGF_EINVAL, "ray_stride=%d (< 6)", ray_stride);
  Guard g(c->device);
  rc = render_and_push(f, c, s, k, rays_local_dev, n_local, ray_stride, n_samples, white_bg, tile_w, mlp_impl,
                       reinterpret_cast<cudaStream_t>(stream));
  if (rc) return rc;
  c->slot[s].ticket = k + 1;
  c->slot[s].released = false;
  c->next_step = k + 1;
  *ticket = k + 1;
  return NGF_OK;
}

int ngf_frame_allgather(NgfComm c, uint64_t ticket, void* stream, const float** frame_dev) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (!frame_dev) return ngf_set_error(NGF_EINVAL, "frame_dev is NULL");
  const int s = (int)((ticket - 1) % c->n_slots);
  if (ticket == 0 || c->slot[s].ticket != ticket) return ngf_set_error(NGF_EINVAL, "ticket %llu is not buffered any more", (unsigned long long)ticket);
  Guard g(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CU(cudaStreamWaitEvent(st, c->slot[s].ev_rendered, 0));
  if ((rc = launch_wait(c, local_flags(c, s, true), (uint32_t)ticket, st))) return rc;
  *frame_dev = reinterpret_cast<const float*>(c->frame(c->rank, s));
  return NGF_OK;
}

int ngf_frame_release(NgfComm c, uint64_t ticket, void* stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  const int s = (int)((ticket - 1) % c->n_slots);
  if (ticket == 0 || c->slot[s].ticket != ticket) return ngf_set_error(NGF_EINVAL, "ticket %llu is not buffered any more", (unsigned long long)ticket);
  if (c->slot[s].released) return NGF_OK;
  Guard g(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CU(cudaEventRecord(c->slot[s].ev_consumed, st));
  PeerList pl{};
  for (int d = 1; d < c->world; ++d) pl.p[pl.n++] = c->freed((c->rank + d) % c->world, s, c->rank);
  if ((rc = launch_signal(pl, (uint32_t)ticket, st))) return rc;
  c->slot[s].released = true;
  return NGF_OK;
}

int ngf_field_render_sharded_host_async(NgfField f, NgfComm c, const float* rays_local_host, int64_t n_local,
                                        int32_t ray_stride, int32_t n_samples, int32_t white_bg, int32_t tile_w,
                                        int32_t mlp_impl, float* frame_host, int64_t first_row, int64_t n_rows,
                                        uint64_t* ticket) {
  if (!ticket) return ngf_set_error(NGF_EINVAL, "ticket is NULL");
  int s = 0;
  unsigned long long k = 0;
  int rc = begin_step(f, c, n_local, mlp_impl, &s, &k);
  if (rc) return rc;
  if (n_local > 0 && !rays_local_host) return ngf_set_error(NGF_EINVAL, "rays is NULL");
  if (ray_stride < 6) return ngf_set_error(NGF_EINVAL, "ray_stride=%d (< 6)", ray_stride);
  if (first_row < 0 || n_rows < 0 || first_row + n_rows > c->n_rays || (n_rows > 0 && !frame_host))
    return ngf_set_error(NGF_EINVAL, "rows [%lld, %lld) of a %lld-ray batch", (long long)first_row, (long long)(first_row + n_rows), c->n_rays);
  Guard g(c->device);
  NgfComm_::Slot& sl = c->slot[s];
  if (c->rays_stride != ray_stride || c->rays_cap < n_local) {
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < c->n_slots; ++i) {
      cudaFree(c->slot[i].rays);
      c->slot[i].rays = nullptr;
      CU(cudaMalloc(reinterpret_cast<void**>(&c->slot[i].rays), (size_t)(n_local > 0 ? n_local : 1) * ray_stride * sizeof(float)));
    }
    c->rays_stride = ray_stride;
    c->rays_cap = n_local;
  }
  // upload: the slot's ray buffer is free once the render that read it (step k - n_slots) is done
  CU(cudaStreamWaitEvent(c->s_in, sl.ev_rendered, 0));
  if (n_local > 0)
    CU(cudaMemcpyAsync(sl.rays, rays_local_host, (size_t)n_local * ray_stride * sizeof(float), cudaMemcpyHostToDevice, c->s_in));
  CU(cudaEventRecord(sl.ev_in, c->s_in));
  cudaStream_t comp = c->s_comp[s % kCompStreams];
  CU(cudaStreamWaitEvent(comp, sl.ev_in, 0));
  rc = render_and_push(f, c, s, k, sl.rays, n_local, ray_stride, n_samples, white_bg, tile_w, mlp_impl, comp);
  if (rc) return rc;
  // download: my rows (local event) + everybody else's (flags), then give the buffer back to the peers
  CU(cudaStreamWaitEvent(c->s_out, sl.ev_rendered, 0));
  if ((rc = launch_wait(c, local_flags(c, s, true), (uint32_t)(k + 1), c->s_out))) return rc;
  if (n_rows > 0)
    CU(cudaMemcpyAsync(frame_host, reinterpret_cast<const float*>(c->frame(c->rank, s)) + first_row * 4, (size_t)n_rows * 16,
                       cudaMemcpyDeviceToHost, c->s_out));
  CU(cudaEventRecord(sl.ev_consumed, c->s_out));
  PeerList pl{};
  for (int d = 1; d < c->world; ++d) pl.p[pl.n++] = c->freed((c->rank + d) % c->world, s, c->rank);
  if ((rc = launch_signal(pl, (uint32_t)(k + 1), c->s_out))) return rc;
  CU(cudaEventRecord(sl.ev_done, c->s_out));
  sl.ticket = k + 1;
  sl.released = true;
  c->next_step = k + 1;
  *ticket = k + 1;
  return NGF_OK;
}

// Camera batches: the batch is n_frames frames of one pinhole camera model (intrinsics of `camera`, one pose each), rays are
// generated inside the march kernel from the pixel index.  poses: [n_frames][12] row-major [3][4] camera-to-world.
static int camera_batch(NgfComm_* c, const NgfCamera* camera, int n_frames, int rank_check, CamDev* out) {
  if (!camera) return ngf_set_error(NGF_EINVAL, "camera is NULL");
  if (n_frames < 1 || n_frames > kMaxFrames) return ngf_set_error(NGF_EINVAL, "n_frames=%d (1..%d)", n_frames, kMaxFrames);
  if (camera->width < 1 || camera->height < 1 || !(camera->fx != 0.f) || !(camera->fy != 0.f))
    return ngf_set_error(NGF_EINVAL, "bad camera");
  if ((long long)n_frames * camera->width * camera->height != c->n_rays)
    return ngf_set_error(NGF_EINVAL, "%d frames of %dx%d pixels are not the comm's %lld-ray batch", n_frames, camera->width,
                         camera->height, c->n_rays);
  (void)rank_check;
  CamDev cam{};
  memcpy(cam.c2w, camera->c2w, sizeof(cam.c2w));
  cam.fx = camera->fx; cam.fy = camera->fy; cam.cx = camera->cx; cam.cy = camera->cy;
  cam.W = camera->width; cam.H = camera->height;
  cam.base = 0;
  cam.shard_block = c->block; cam.shard_rank = c->rank; cam.shard_world = c->world;
  *out = cam;
  return NGF_OK;
}

int ngf_field_render_sharded_camera(NgfField f, NgfComm c, const NgfCamera* camera, const float* poses_dev, int32_t n_frames,
                                    int32_t n_samples, int32_t white_bg, int32_t mlp_impl, void* stream, uint64_t* ticket) {
  if (!ticket || !poses_dev) return ngf_set_error(NGF_EINVAL, "NULL argument");
  int s = 0;
  unsigned long long k = 0;
  int rc = begin_step(f, c, c ? c->n_local : 0, mlp_impl, &s, &k);
  if (rc) return rc;
  CamDev cam{};
  if ((rc = camera_batch(c, camera, n_frames, c->rank, &cam))) return rc;
  cam.poses = poses_dev;
  Guard g(c->device);
  const int tile_w = (c->block % (4 * cam.W) == 0) ? cam.W : 0;       // whole 4-row groups per block: 8x4-pixel warp tiles
  rc = render_and_push(f, c, s, k, nullptr, c->n_local, 6, n_samples, white_bg, tile_w, mlp_impl,
                       reinterpret_cast<cudaStream_t>(stream), &cam);
  if (rc) return rc;
  c->slot[s].ticket = k + 1;
  c->slot[s].released = false;
  c->next_step = k + 1;
  *ticket = k + 1;
  return NGF_OK;
}

int ngf_field_render_sharded_camera_u8_host_async(NgfField f, NgfComm c, const NgfCamera* camera, const float* poses_host,
                                                  int32_t n_frames, int32_t n_samples, int32_t white_bg, int32_t mlp_impl,
                                                  uint8_t* u8_host, int64_t first_row, int64_t n_rows, uint64_t* ticket) {
  if (!ticket || !poses_host) return ngf_set_error(NGF_EINVAL, "NULL argument");
  int s = 0;
  unsigned long long k = 0;
  int rc = begin_step(f, c, c ? c->n_local : 0, mlp_impl, &s, &k);
  if (rc) return rc;
  CamDev cam{};
  if ((rc = camera_batch(c, camera, n_frames, c->rank, &cam))) return rc;
  if (first_row < 0 || n_rows < 0 || first_row + n_rows > c->n_rays || (n_rows > 0 && !u8_host))
    return ngf_set_error(NGF_EINVAL, "rows [%lld, %lld) of a %lld-ray batch", (long long)first_row, (long long)(first_row + n_rows), c->n_rays);
  Guard g(c->device);
  NgfComm_::Slot& sl = c->slot[s];
  if (!sl.poses) CU(cudaMalloc(reinterpret_cast<void**>(&sl.poses), kMaxFrames * 12 * sizeof(float)));
  if (sl.u8_cap < n_rows) {
    CU(cudaStreamSynchronize(c->s_out));
    cudaFree(sl.u8);
    sl.u8 = nullptr; sl.u8_cap = 0;
    CU(cudaMalloc(reinterpret_cast<void**>(&sl.u8), (size_t)(n_rows > 0 ? n_rows : 1) * 3));
    sl.u8_cap = n_rows;
  }
  // upload the poses: the slot's pose buffer is free once the render that read it (step k - n_slots) is done
  CU(cudaStreamWaitEvent(c->s_in, sl.ev_rendered, 0));
  CU(cudaMemcpyAsync(sl.poses, poses_host, (size_t)n_frames * 12 * sizeof(float), cudaMemcpyHostToDevice, c->s_in));
  CU(cudaEventRecord(sl.ev_in, c->s_in));
  cudaStream_t comp = c->s_comp[s % kCompStreams];
  CU(cudaStreamWaitEvent(comp, sl.ev_in, 0));
  cam.poses = sl.poses;
  const int tile_w = (c->block % (4 * cam.W) == 0) ? cam.W : 0;
  rc = render_and_push(f, c, s, k, nullptr, c->n_local, 6, n_samples, white_bg, tile_w, mlp_impl, comp, &cam);
  if (rc) return rc;
  CU(cudaStreamWaitEvent(c->s_out, sl.ev_rendered, 0));
  if ((rc = launch_wait(c, local_flags(c, s, true), (uint32_t)(k + 1), c->s_out))) return rc;
  if (n_rows > 0) {
    long long blocks = (n_rows + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    ngf_frame4_u8_kernel<<<(unsigned)blocks, 256, 0, c->s_out>>>(c->frame(c->rank, s), first_row, n_rows, sl.u8);
    count_launch();
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(sl.ev_consumed, c->s_out));          // the frame buffer has been read: peers may overwrite it
  PeerList pl{};
  for (int d = 1; d < c->world; ++d) pl.p[pl.n++] = c->freed((c->rank + d) % c->world, s, c->rank);
  if ((rc = launch_signal(pl, (uint32_t)(k + 1), c->s_out))) return rc;
  if (n_rows > 0) CU(cudaMemcpyAsync(u8_host, sl.u8, (size_t)n_rows * 3, cudaMemcpyDeviceToHost, c->s_out));
  CU(cudaEventRecord(sl.ev_done, c->s_out));
  sl.ticket = k + 1;
  sl.released = true;
  c->next_step = k + 1;
  *ticket = k + 1;
  return NGF_OK;
}

int ngf_comm_wait(NgfComm c, uint64_t ticket) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (ticket == 0 || ticket > c->next_step) return ngf_set_error(NGF_EINVAL, "unknown ticket %llu", (unsigned long long)ticket);
  Guard g(c->device);
  // a recycled slot's event now stands for a later ticket, whose completion implies this one's
  CU(cudaEventSynchronize(c->slot[(ticket - 1) % c->n_slots].ev_done));
  return check_err(c);
}

}  // extern "C"
