// Rate probe for CTA-pair MMAs: tcgen05.mma.cta_group::2.kind::f16, M = 256 (128 rows per CTA), K = 16.
// Each CTA of the pair holds its 128 A rows and N/2 B rows at the same shared-memory offsets; the leader
// CTA issues.  Prints cycles per instruction for N = 32 .. 256 next to the single-CTA numbers of
// umma_f16_probe.cu (85 / 89 / 97 / 113 / 171 cycles for N = 16 / 32 / 64 / 128 / 256).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_2cta_probe umma_2cta_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../tf_image_compression_b200/csrc/tic_ptx.cuh"
using namespace tic::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) rate2_kernel(int N, int count, int R, long long* cycles, uint32_t layout, uint32_t sbo, uint32_t a_start) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_rank();
  for (int i = tid; i < (32768 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(256, N);
    // A: `layout` rows (2 = SWIZZLE_128B, 4 = 64B, 6 = 32B), 8-row groups `sbo` bytes apart, start offset a_start
    // (the "one box, nine taps" operands of tic_umma16.cuh start on arbitrary rows: sbo = 9 or 10 rows)
    const uint64_t ad0 = make_desc(smem_u32(smem) + a_start, sbo, layout);
    const uint64_t bd0 = make_desc(smem_u32(smem + 32768), layout == 2 ? 1024 : layout == 4 ? 512 : 256, layout);
    long long t0 = clock64();
    uint32_t r = 0;
    for (int i = 0; i < count; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tmem_base + r * (uint32_t)N;
          r = (r + 1 == (uint32_t)R) ? 0 : r + 1;
          mma2_f16_ss(d, ad0 + 2 * k, bd0 + 2 * k, idesc, 1);
        }
      }
      r = __shfl_sync(0xffffffffu, r, 0);
      __syncwarp();
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar_mma)), "h"((uint16_t)1) : "memory");
    __syncwarp();
    long long t1 = clock64();
    while (!mbar_try_wait(&bar_mma, 0)) { if (clock64() - t1 > 2000000000LL) break; }
    long long t2 = clock64();
    if ((tid & 31) == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

int main() {
  const size_t smem_bytes = 32768 + 32768;
  CK(cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, 16));
  for (int N : {32, 64, 128, 256})
    for (int R : {1, 2}) {
      if (R * N > 512) continue;
      const int count = 4096;
      rate2_kernel<<<2, 128, smem_bytes>>>(N, count, R, d_cyc, 2u, 1024u, 0u);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel error %s (N=%d)\n", cudaGetErrorString(e), N); return 3; }
      long long c[2];
      CK(cudaMemcpy(c, d_cyc, 16, cudaMemcpyDeviceToHost));
      printf("2CTA f16 K=16 M=256 N=%3d R=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  -> %.0f MAC/cycle/SM (ideal %d cycles)\n", N, R,
             (double)c[0] / count, (double)c[1] / count, 128.0 * N * 16 / ((double)c[1] / count), N / 2);
    }
  // operand layouts of the 32-channel layers (64-byte rows) against the 64-channel ones (128-byte rows)
  struct Cfg { const char* name; uint32_t layout, sbo, start; };
  const Cfg cfgs[] = {{"SW128 sbo=1024", 2, 1024, 0},     {"SW128 sbo=1280 (10-col box) start=128", 2, 1280, 128},
                      {"SW64  sbo=512", 4, 512, 0},       {"SW64  sbo=576 (9-col box) start=64", 4, 576, 64},
                      {"SW64  sbo=640 (10-col box) start=64", 4, 640, 64}, {"SW32  sbo=256", 6, 256, 0}};
  for (const Cfg& c : cfgs)
    for (int N : {16, 32, 64, 128}) {
      const int count = 4096;
      rate2_kernel<<<2, 128, smem_bytes>>>(N, count, 1, d_cyc, c.layout, c.sbo, c.start);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel error %s (%s N=%d)\n", cudaGetErrorString(e), c.name, N); return 3; }
      long long cy[2];
      CK(cudaMemcpy(cy, d_cyc, 16, cudaMemcpyDeviceToHost));
      printf("2CTA f16 K=16 M=256 N=%3d %-40s: %.1f cyc/MMA\n", N, c.name, (double)cy[1] / count);
    }
  return 0;
}
