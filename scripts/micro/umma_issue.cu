// Microbenchmark 2: what limits the tcgen05.mma issue cadence of one thread?  cta_group::1, M=128, kind::f16, zeros.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_issue umma_issue.cu
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
DEVI void wait_bar(uint64_t* bar, uint32_t par) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(bar)), "r"(par) : "memory");
    if (!ok && ++spins > (1u << 22)) __trap();
  } while (!ok);
}
DEVI bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
  return pred != 0;
}
// mode 0: one thread, descriptors computed per iteration     mode 1: one thread, two fixed descriptor pairs, unrolled
// mode 2: whole warp runs the loop, elect.sync lane issues    mode 3: two warps issue concurrently (iters/2 each)
// mode 4: like 1 but the accumulate predicate is constant 1   mode 5: like 1, one commit after every 4 MMAs
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int iters, long long* out, int rnd) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint64_t junk;
  __shared__ uint64_t full[8], empty[8];
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (rnd) {      // fp16 pairs with random mantissas and signs, magnitudes in [0.25, 0.5): finite sums
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      uint32_t w[4];
      for (int j = 0; j < 4; ++j) { h = h * 1664525u + 1013904223u; w[j] = (h & 0x83FF83FFu) | 0x34003400u; }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    reinterpret_cast<uint4*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[1])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(s32(&junk)));
    for (int i = 0; i < 8; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[i])));
    }
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
  const uint32_t id = idesc(128, N);
  const uint32_t aBase = s32(smem), bBase = s32(smem) + 128 * 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (mode == 0 && threadIdx.x == 0) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int ks = it & 15, half = it & 1;
      mma1(tmem + half * 256, desc(aBase + ks * 8192 + half * 2048, 4096, 128), desc(bBase + (ks & 7) * N * 32, N * 16, 128), id, it > 0);
    }
    t1 = clock64(); commit(&bar[0]); wait_bar(&bar[0], 0); t2 = clock64();
  } else if ((mode == 1 || mode == 4 || mode == 5) && threadIdx.x == 0) {
    const uint64_t a0 = desc(aBase, 4096, 128), a1 = desc(aBase + 2048, 4096, 128), b0 = desc(bBase, N * 16, 128);
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mma1(tmem + (j & 1) * 256, (j & 1) ? a1 : a0, b0, id, mode == 4 ? 1u : (uint32_t)(it + j > 1));
        if (mode == 5 && (j & 3) == 3) commit(&junk);
      }
    }
    t1 = clock64(); commit(&bar[0]); wait_bar(&bar[0], 0); t2 = clock64();
  } else if (mode == 2 && warp == 0) {
    const uint64_t a0 = desc(aBase, 4096, 128), a1 = desc(aBase + 2048, 4096, 128), b0 = desc(bBase, N * 16, 128);
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it += 8) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mma1(tmem + (j & 1) * 256, (j & 1) ? a1 : a0, b0, id, (uint32_t)(it + j > 1));
      }
      __syncwarp();
    }
    t1 = clock64();
    if (elect_one()) commit(&bar[0]);
    __syncwarp();
    wait_bar(&bar[0], 0); t2 = clock64();
  } else if (mode == 3 && warp < 2 && lane == 0) {
    const uint64_t a0 = desc(aBase + warp * 2048, 4096, 128), b0 = desc(bBase, N * 16, 128);
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters / 2; it += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) mma1(tmem + warp * 256, a0, b0, id, (uint32_t)(it + j > 0));
    }
    t1 = clock64(); commit(&bar[warp]); wait_bar(&bar[warp], 0); t2 = clock64();
  }
  // mode 6: stages of 4 MMAs, commit to a junk barrier + try_wait on an already completed barrier per stage
  // mode 7: a real 4-stage full/empty ring against a producer thread (no copies), 4 MMAs per stage
  // mode 8: same, 8 MMAs per stage      mode 9: 8 stages of 4 MMAs      mode 10: mode 7, producer also copies 16 KiB
  if (mode == 6 && threadIdx.x == 0) {
    const uint64_t a0 = desc(aBase, 4096, 128), a1 = desc(aBase + 2048, 4096, 128), b0 = desc(bBase, N * 16, 128);
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it += 4) {
      wait_bar(&full[(it >> 2) & 3], 1);          // fresh barrier: the phase with parity 1 counts as complete
#pragma unroll
      for (int j = 0; j < 4; ++j) mma1(tmem + (j & 1) * 256, (j & 1) ? a1 : a0, b0, id, (uint32_t)(it + j > 1));
      commit(&junk);
    }
    t1 = clock64(); commit(&bar[0]); wait_bar(&bar[0], 0); t2 = clock64();
  } else if (mode >= 7 && mode <= 10) {
    const int per = mode == 8 ? 8 : 4, nst = mode == 9 ? 8 : 4;
    if (warp == 1 && lane == 0) {               // producer
      for (int it = 0, g = 0; it < iters; it += per, ++g) {
        const int st = g % nst;
        wait_bar(&empty[st], ((g / nst) & 1) ^ 1);
        if (mode == 10) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(16384) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + 128 * 1024 + st * 16384)), "l"(reinterpret_cast<const uint8_t*>(out) + 65536 + (size_t)(g & 63) * 16384), "r"(16384), "r"(s32(&full[st])) : "memory");
        } else {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&full[st])) : "memory");
        }
      }
    } else if (warp == 0 && lane == 0) {        // issuer
      const uint64_t a0 = desc(aBase, 4096, 128), a1 = desc(aBase + 2048, 4096, 128);
      t0 = clock64();
      for (int it = 0, g = 0; it < iters; it += per, ++g) {
        const int st = g % nst;
        wait_bar(&full[st], (g / nst) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t b0 = desc(bBase + (st & 3) * 16384, N * 16, 128);
        for (int j = 0; j < per; ++j) mma1(tmem + (j & 1) * 256, (j & 1) ? a1 : a0, b0 + ((j >> 1) & 1) * (8192 >> 4), id, (uint32_t)(it + j > 1));
        commit(&empty[st]);
      }
      t1 = clock64(); commit(&bar[0]); wait_bar(&bar[0], 0); t2 = clock64();
    }
  }
  if (lane == 0 && warp < 2) { out[(blockIdx.x * 2 + warp) * 2] = t1 - t0; out[(blockIdx.x * 2 + warp) * 2 + 1] = t2 - t0; }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 192 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out; cudaMallocManaged(&out, 65536 + 64 * 16384 + 16384);
  for (int rnd : {1})
  for (int iters : {4096})
  for (int grid : {1, sms})
    for (int N : {256})
      for (int mode : {0, 5, 6, 7, 8, 9, 10}) {
        for (int i = 0; i < sms * 4; ++i) out[i] = 0;
        k<<<grid, 128, smem>>>(mode, N, iters, out, rnd);
        cudaError_t e = cudaDeviceSynchronize();
        long long w1 = out[0], w2 = out[1]; if (mode == 3) { if (out[2] > w1) w1 = out[2]; if (out[3] > w2) w2 = out[3]; }
        printf("data %s iters %6d grid %3d N=%3d mode %d: issue %.1f cyc/mma, complete %.1f cyc/mma  [%s]\n", rnd ? "random" : "zeros ", iters, grid, N, mode, (double)w1 / iters, (double)w2 / iters, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
      }
  return 0;
}
