// Microbenchmark 3: tcgen05.mma issue rate inside a full/empty mbarrier ring against a producer thread (the structure of
// the NeuTex MLP kernel's weight ring), cta_group::1, M=128 N=256 K=16, kind::f16.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_ring umma_ring.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define DEVI __device__ __forceinline__
DEVI uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVI uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
DEVI void mma1(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
DEVI void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory"); }
DEVI void wait_try(uint64_t* bar, uint32_t par) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(bar)), "r"(par) : "memory");
    if (!ok && ++spins > (1u << 22)) __trap();
  } while (!ok);
}
DEVI void wait_test(uint64_t* bar, uint32_t par) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(bar)), "r"(par) : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();
  } while (!ok);
}
struct P { int per, nst, pad, test, fence, workers, iters; };
// pad: barriers 128 B apart instead of 8;  test: test_wait spin instead of try_wait;  fence: tcgen05.fence::after_thread_sync
// after the full wait;  workers: that many extra warps spin (try_wait) on a barrier that only completes at the end
__global__ void __launch_bounds__(640, 1) k(P p, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(128) uint64_t bars[16 * 16 + 32];
  __shared__ uint32_t tmem_base;
  const int stride = p.pad ? 16 : 1;
  uint64_t* full = bars;
  uint64_t* empty = bars + 8 * stride;
  uint64_t* done = bars + 16 * stride;
  for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u, w[4];
    for (int j = 0; j < 4; ++j) { h = h * 1664525u + 1013904223u; w[j] = (h & 0x83FF83FFu) | 0x34003400u; }
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i * stride])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[i * stride])));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(done)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t id = idesc(128, 256);
  const uint32_t aBase = s32(smem), bBase = s32(smem) + 128 * 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 1 && lane == 0) {               // producer
    for (int it = 0, g = 0; it < p.iters; it += p.per, ++g) {
      const int st = g % p.nst;
      if (p.test) wait_test(&empty[st * stride], ((g / p.nst) & 1) ^ 1); else wait_try(&empty[st * stride], ((g / p.nst) & 1) ^ 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&full[st * stride])) : "memory");
    }
  } else if (warp == 0 && lane == 0) {        // issuer
    const uint64_t a0 = desc(aBase, 4096, 128), a1 = desc(aBase + 2048, 4096, 128);
    t0 = clock64();
    for (int it = 0, g = 0; it < p.iters; it += p.per, ++g) {
      const int st = g % p.nst;
      if (p.test) wait_test(&full[st * stride], (g / p.nst) & 1); else wait_try(&full[st * stride], (g / p.nst) & 1);
      if (p.fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t b0 = desc(bBase + (st & 3) * 16384, 4096, 128);
      for (int j = 0; j < p.per; ++j) mma1(tmem + (j & 1) * 256, (j & 1) ? a1 : a0, b0 + ((j >> 1) & 1) * (8192 >> 4), id, (uint32_t)(it + j > 1));
      commit(&empty[st * stride]);
    }
    t1 = clock64(); commit(done); wait_try(done, 0); t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0;
  } else if (warp >= 2 && warp < 2 + p.workers) {
    wait_try(done, 0);
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 192 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out; cudaMallocManaged(&out, sms * 2 * sizeof(long long));
  const int iters = 4800;
  for (int workers : {0, 16})
    for (int pad : {0, 1})
      for (int test : {0, 1})
        for (int nst : {2, 4})
          for (int per : {2, 4, 6, 8, 16}) {
            P p{per, nst, pad, test, 1, workers, iters};
            k<<<sms, 640, smem>>>(p, out);
            cudaError_t e = cudaDeviceSynchronize();
            printf("workers %2d pad %d %s nst %d per %2d: %.1f cyc/mma  [%s]\n", workers, pad, test ? "test_wait" : "try_wait ", nst, per, (double)out[1] / iters, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
          }
  return 0;
}
