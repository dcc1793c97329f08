// Micro-benchmark: mbarrier hand-off latency between two warps of a CTA (sm_100a): plain arrive vs tcgen05.commit,
// spin on try_wait.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_latency.bin mbar_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void commit(uint64_t *b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
}
// MODE 0: lane-0-only ping-pong with arrive; 1: whole warps wait (lane 0 arrives after __syncwarp); 2: A signals with tcgen05.commit
template <int MODE>
__global__ void k(int iters, long long *out) {
  __shared__ uint64_t b1, b2;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&b1, 1); mbar_init(&b2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32u) : "memory");
                   asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  __syncthreads();
  long long t0 = clock64();
  if (MODE == 0) {
    if (lane == 0) for (int i = 0; i < iters; ++i) {
      if (warp == 0) { mbar_arrive(&b1); mbar_wait(&b2, i & 1); } else { mbar_wait(&b1, i & 1); mbar_arrive(&b2); }
    }
  } else {
    for (int i = 0; i < iters; ++i) {
      if (warp == 0) {
        if (lane == 0) { if (MODE == 2) commit(&b1); else mbar_arrive(&b1); }
        __syncwarp();
        mbar_wait(&b2, i & 1);
      } else {
        mbar_wait(&b1, i & 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(&b2);
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(32u) : "memory");
}
int main() {
  long long *out, h; cudaMalloc(&out, 8);
  const int iters = 20000;
  k<0><<<1, 64>>>(iters, out); cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); printf("lane-0 ping-pong (arrive):      %.1f cycles per hand-off\n", (double)h / iters / 2);
  k<1><<<1, 64>>>(iters, out); cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); printf("warp-wide wait (arrive):        %.1f cycles per hand-off\n", (double)h / iters / 2);
  k<2><<<1, 64>>>(iters, out); cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); printf("warp-wide, A signals by commit: %.1f cycles per round trip (%s)\n", (double)h / iters, cudaGetErrorString(e));
  return 0;
}
