// Microbenchmark: issue rate of tcgen05.mma kind::f16 (cycles per instruction) for the operand layouts / modes the
// NeuTex MLP kernel could use.  Data are zeros; only the timing matters.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define DEVI __device__ __forceinline__
DEVI uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVI uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
DEVI void mma1(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
DEVI void mma1_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
DEVI void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
DEVI void mma2_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
DEVI void wait_bar(uint64_t* bar, uint32_t par) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(bar)), "r"(par) : "memory");
    if (!ok && ++spins > (1u << 22)) __trap();
  } while (!ok);
}
struct Cfg { int cg, N, layout, ts, alt, kslices, iters; };   // layout 0 none, 2 swizzle-128B; ts: A from TMEM; alt: alternate 2 accumulators

template <int CG>
__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
  if (threadIdx.x < 32) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0 && rank == 0) {
    const int M = CG == 2 ? 256 : 128;
    const uint32_t id = idesc(M, c.N);
    const uint32_t rowsB = CG == 2 ? c.N / 2 : c.N;
    const uint32_t aBase = s32(smem), bBase = s32(smem) + 128 * 1024;
    // layout 0: [K-groups][rows][16 B]: A has 256 rows per K group (two 128-row halves), B rowsB rows
    // layout 2: 128B swizzle, K-major: 8-row groups of 1024 B, 64 K elements per row; K=16 step = +32 B
    for (int rep = 0; rep < 2; ++rep) {
      long long t0 = clock64();
      for (int it = 0; it < c.iters; ++it) {
        const int ks = it % c.kslices;
        const int half = c.alt ? (it & 1) : 0;
        uint64_t da, db;
        if (c.layout == 0) {
          da = desc(aBase + ks * 2 * 4096 + half * 2048, 4096, 128, 0);
          db = desc(bBase + (ks & 7) * rowsB * 32, rowsB * 16, 128, 0);
        } else {
          da = desc(aBase + half * 16384 + (ks & 3) * 32 + (ks >> 2) * 32768, 16, 1024, 2);
          db = desc(bBase + (ks & 3) * 32 + ((ks >> 2) & 1) * 32768, 16, 1024, 2);
        }
        const uint32_t d = tmem + (c.N <= 128 || !c.ts ? half * 256 : 0);
        if (CG == 2) { if (c.ts) mma2_ts(d, tmem + 384 + ks * 8 % 64, db, id, it > 0); else mma2(d, da, db, id, it > 0); }
        else { if (c.ts) mma1_ts(d, tmem + 384 + ks * 8 % 64, db, id, it > 0); else mma1(d, da, db, id, it > 0); }
      }
      long long t1 = clock64();
      if (CG == 2) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      wait_bar(&bar, rep & 1);
      long long t2 = clock64();
      if (rep == 1) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
  if (threadIdx.x < 32) {
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 192 * 1024;
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out; cudaMallocManaged(&out, sms * 2 * sizeof(long long));
  const Cfg cfgs[] = {
      {1, 256, 0, 0, 1, 16, 2048}, {1, 256, 0, 0, 0, 16, 2048}, {1, 256, 2, 0, 1, 16, 2048}, {1, 256, 2, 0, 0, 16, 2048},
      {1, 128, 0, 0, 1, 16, 2048}, {1, 128, 2, 0, 1, 16, 2048}, {1, 64, 0, 0, 1, 16, 2048},  {1, 64, 2, 0, 1, 16, 2048},
      {1, 256, 0, 1, 0, 16, 2048}, {1, 128, 0, 1, 1, 16, 2048}, {1, 128, 2, 1, 1, 16, 2048}, {1, 256, 0, 0, 1, 1, 2048},
      {2, 256, 0, 0, 1, 16, 2048}, {2, 256, 2, 0, 1, 16, 2048}, {2, 256, 0, 0, 0, 16, 2048}, {2, 128, 0, 0, 1, 16, 2048},
      {2, 128, 2, 0, 1, 16, 2048}, {2, 256, 0, 1, 0, 16, 2048}, {2, 128, 2, 1, 1, 16, 2048}, {2, 256, 0, 0, 1, 1, 2048},
  };
  for (int grid : {2, (sms / 2) * 2})
    for (const Cfg& c : cfgs) {
      for (int i = 0; i < sms * 2; ++i) out[i] = 0;
      cudaError_t e;
      if (c.cg == 2) {
        cudaLaunchConfig_t lc{}; lc.gridDim = dim3(grid); lc.blockDim = dim3(128); lc.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        e = cudaLaunchKernelEx(&lc, k<2>, c, out);
      } else {
        k<1><<<grid, 128, smem>>>(c, out);
        e = cudaGetLastError();
      }
      cudaError_t e2 = cudaDeviceSynchronize();
      long long mx = 0; for (int i = 0; i < grid; ++i) if (out[i * 2 + 1] > mx) mx = out[i * 2 + 1];
      printf("grid %3d cg%d N=%3d layout=%d A=%s alt=%d kslices=%2d: issue %.1f cyc/mma, complete %.1f cyc/mma (cta0), worst cta %.1f  [%s %s]\n", grid, c.cg, c.N,
             c.layout, c.ts ? "tmem" : "smem", c.alt, c.kslices, (double)out[0] / c.iters, (double)out[1] / c.iters, (double)mx / c.iters,
             cudaGetErrorString(e), cudaGetErrorString(e2));
      if (e2 != cudaSuccess) return 1;
    }
  return 0;
}
