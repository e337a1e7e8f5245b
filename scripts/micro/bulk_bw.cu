// Microbenchmark: per-SM throughput of cp.async.bulk (global/L2 -> shared) for different copy sizes and ring depths.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bw bulk_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, size_t src_bytes, int copy_bytes, int stages, int iters, int lanes) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 196608);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + s)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  // `lanes` producer threads (one per warp), each owning stages s % lanes
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < lanes) {
    size_t off = ((size_t)blockIdx.x * 7919 * 4096) % (src_bytes - (size_t)copy_bytes);
    for (int it = w; it < iters; it += lanes) {
      const int s = it % stages;
      if (it >= stages) {   // wait for the previous use of this stage
        uint32_t ok = 0, par = ((it / stages) - 1) & 1;
        do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(bar + s)), "r"(par)); } while (!ok);
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + s)), "r"(copy_bytes));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + (size_t)s * copy_bytes)), "l"(src + off), "r"(copy_bytes), "r"(s32(bar + s)) : "memory");
      off += copy_bytes; if (off + copy_bytes > src_bytes) off = 0;
    }
    // drain
    for (int s = 0; s < stages && s < iters; ++s) {
      int last = ((iters - 1 - s) / stages) * stages + s; if (last % lanes != w) continue;
      uint32_t ok = 0, par = (last / stages) & 1;
      do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(bar + s)), "r"(par)); } while (!ok);
    }
  }
}
int main() {
  const size_t src_bytes = 3u << 20;   // 3 MB: L2 resident, like the NeuTex weight stream
  uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608 + 256);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int lanes : {1, 2})
    for (int cb : {4096, 8192, 16384, 32768})
      for (int st : {2, 4, 6}) {
        if ((size_t)cb * st > 196608) continue;
        const int iters = (64 << 20) / cb;     // 64 MB per CTA
        k<<<sms, 128, 196608 + 256>>>(src, src_bytes, cb, st, iters, lanes);
        cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<sms, 128, 196608 + 256>>>(src, src_bytes, cb, st, iters, lanes); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("producers %d copy %6d B stages %d: %.1f GB/s per SM, %.2f TB/s total  (%s)\n", lanes, cb, st, (double)cb * iters / ms / 1e6, (double)cb * iters * sms / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
