// Probe: rounding behaviour of the tcgen05 fp32 accumulator (kind::tf32).
// Every row of A and of B holds the same 8-vector per step, so D[m][n] = sum_steps sum_k a[s][k]*b[s][k].
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../tf_image_compression_b200/csrc/tic_ptx.cuh"
using namespace tic::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

__global__ void __launch_bounds__(128) accum_kernel(const float* av, const float* bv, int steps, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;          // 128 x 128 B
  uint8_t* sB = smem + 16384;  // 16 x 128 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base, 32); tmem_relinquish(); }
  for (int i = tid; i < (16384 + 2048) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  tc_fence_before(); __syncthreads(); tc_fence_after();
  for (int s = 0; s < steps; ++s) {
    for (int i = tid; i < 128 * 8; i += 128) *reinterpret_cast<float*>(sA + sw128_offset(i / 8, i % 8)) = av[s * 8 + i % 8];
    for (int i = tid; i < 16 * 8; i += 128) *reinterpret_cast<float*>(sB + sw128_offset(i / 8, i % 8)) = bv[s * 8 + i % 8];
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mma_tf32_ss(tmem_base, make_smem_desc_sw128(smem_u32(sA), 1024), make_smem_desc_sw128(smem_u32(sB), 1024),
                  make_idesc_tf32(128, 16), s > 0);
      tc_commit(&bar);
    }
    mbar_wait(&bar, s & 1);
    tc_fence_after();
    __syncthreads();
  }
  float v[16];
  tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), v);
  if (tid == 0) out[0] = v[0];
  if (tid == 127) out[1] = v[15];
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 32);
}

static float run(const std::vector<float>& a, const std::vector<float>& b) {
  int steps = (int)a.size() / 8;
  float *da, *db, *dout;
  CK(cudaMalloc(&da, a.size() * 4)); CK(cudaMalloc(&db, b.size() * 4)); CK(cudaMalloc(&dout, 8));
  CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  accum_kernel<<<1, 128, 16384 + 2048>>>(da, db, steps, dout);
  CK(cudaDeviceSynchronize());
  float h[2]; CK(cudaMemcpy(h, dout, 8, cudaMemcpyDeviceToHost));
  if (h[0] != h[1]) printf("  (row/col mismatch %a vs %a)\n", h[0], h[1]);
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return h[0];
}

int main() {
  const float ulp = ldexpf(1.f, -23);
  auto test = [&](const char* name, float first, float addend_ulps, int nadd, int per_step) {
    // step 0: D = first; then nadd steps each adding per_step products of (addend_ulps/per_step) ulp
    std::vector<float> a, b;
    for (int k = 0; k < 8; ++k) { a.push_back(k == 0 ? 1.f : 0.f); b.push_back(k == 0 ? first : 0.f); }
    for (int s = 0; s < nadd; ++s)
      for (int k = 0; k < 8; ++k) { a.push_back(k < per_step ? 1.f : 0.f); b.push_back(k < per_step ? addend_ulps * ulp / per_step : 0.f); }
    float got = run(a, b);
    double exact = (double)first + (double)nadd * addend_ulps * ulp;
    printf("%-44s got 1+%8.3f ulp   exact 1+%8.3f ulp\n", name, (got - first) / ulp, (exact - first) / ulp);
  };
  test("64 x (+0.75 ulp)  [RN: +64, RZ: 0, wide: 48]", 1.f, 0.75f, 64, 1);
  test("64 x (+0.25 ulp)  [RN: 0, RZ: 0, wide: 16]", 1.f, 0.25f, 64, 1);
  test("64 x (+0.50 ulp)  [RN-even: 0, RZ: 0, wide: 32]", 1.f, 0.5f, 64, 1);
  test("64 x (+1.50 ulp)  [RN: 128/64?, RZ: 64, wide: 96]", 1.f, 1.5f, 64, 1);
  test("8 x (8 products of 0.25 ulp in one MMA) [wide sum: +16]", 1.f, 2.0f, 8, 8);
  test("8 x (8 products of 0.125 ulp in one MMA) [sum 1 ulp each: +8]", 1.f, 1.0f, 8, 8);
  test("64 x (-0.25 ulp) from 1.0 [RZ toward zero: -64*?]", 1.f, -0.25f, 64, 1);
  test("64 x (-0.75 ulp) from 1.5", 1.5f, -0.75f, 64, 1);
  return 0;
}
