// Hardware probe for the fp16-pair tensor path (run on a B200 via gpurun; prints PASS/FAIL + timings):
//   F1  tcgen05.mma kind::f16 (fp16 x fp16 -> fp32), M=128, K-major operands in SWIZZLE_128B / 64B / 32B
//       rows (64 / 32 / 16 channels), A staged by a 4-D TMA box with re-ordered dims (C, W, N, H) and
//       the filter-row (kh) shift as an aligned start offset
//   F2  "one box, nine taps": a 10-column box (x-1 .. x+8); the filter-column (kw) shift is a start
//       offset of ONE row (not a multiple of the swizzle atom) and the 8-row groups are 10 rows apart
//       (SBO = 10 * row bytes).  Tried with descriptor base-offset 0 and (addr >> 7) & 7.
//   F3  issue-rate: cycles per back-to-back MMA for N = 16 .. 256, kind::f16 (K=16) and kind::tf32 (K=8)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_f16_probe umma_f16_probe.cu -lcuda
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../tf_image_compression_b200/csrc/tic_ptx.cuh"

using namespace tic::ptx;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

__device__ int g_timeout_flag = 0;
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity, int where) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) {
      atomicExch(&g_timeout_flag, where);
      return false;
    }
  }
  return true;
}

// linear byte offset inside a dense K-major tile (rows `rowbytes` apart) -> swizzled offset
__host__ __device__ inline uint32_t swz(uint32_t off, int rowbytes) {
  const uint32_t mask = rowbytes == 128 ? 7u : rowbytes == 64 ? 3u : 1u;
  return off ^ (((off >> 7) & mask) << 4);
}
__host__ __device__ inline uint32_t layout_code(int rowbytes) { return rowbytes == 128 ? 2u : rowbytes == 64 ? 4u : 6u; }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}
// kind::f16: D f32 (1 << 4), A/B format 0 = f16, K-major both
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

struct ProbeArgs {
  int coord[4];
  uint32_t a_bytes, a_off, sbo, base_off;
  int rowbytes;   // 128 | 64 | 32
  int N;
  const __half* w;  // [N][rowbytes/2] K-major
  float* d_out;     // [128][N]
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmap, ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;          // up to 48 KB
  uint8_t* sB = smem + 49152;  // up to 256 x 128 B
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 256);
    tmem_relinquish();
  }
  const int kel = p.rowbytes / 2;
  for (int i = tid; i < p.N * kel; i += 128) {
    int r = i / kel, k = i % kel;
    *reinterpret_cast<__half*>(sB + swz((uint32_t)r * p.rowbytes + k * 2, p.rowbytes)) = p.w[r * kel + k];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_expect_tx(&bar_tma, p.a_bytes);
    tma_load_4d(sA, &tmap, &bar_tma, p.coord[0], p.coord[1], p.coord[2], p.coord[3]);
  }
  if (!wait_bounded(&bar_tma, 0, 1)) return;
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, p.N);
    const uint32_t lc = layout_code(p.rowbytes);
    for (int k = 0; k < p.rowbytes / 32; ++k) {
      uint64_t ad = make_desc(smem_u32(sA) + p.a_off + k * 32, p.sbo, lc, p.base_off);
      uint64_t bd = make_desc(smem_u32(sB) + k * 32, 8 * p.rowbytes, lc, 0);
      mma_f16_ss(tmem_base, ad, bd, idesc, k > 0);
    }
    tc_commit(&bar_mma);
  }
  if (!wait_bounded(&bar_mma, 0, 2)) return;
  tc_fence_after();
  for (int c = 0; c < p.N; c += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 16; ++i) p.d_out[(warp * 32 + lane) * p.N + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// F3: one thread issues `count` MMAs (4 K-steps per operand pair, round-robin over 2 accumulators).
__global__ void __launch_bounds__(128) rate_kernel(int N, int kind_f16, int count, int R, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    // the whole warp runs the loop; one elected lane issues (the pattern ptxas turns into straight-line
    // predicated UTCHMMA — an `if (tid == 0)` region makes it wrap every MMA in an ELECT/BRA.U.ANY loop)
    const uint32_t idesc = kind_f16 ? make_idesc_f16(128, N) : make_idesc_tf32(128, N);
    const uint64_t ad0 = make_desc(smem_u32(smem), 1024, 2, 0);
    const uint64_t bd0 = make_desc(smem_u32(smem + 16384), 1024, 2, 0);
    long long t0 = clock64();
    uint32_t r = 0;
    for (int i = 0; i < count; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tmem_base + r * (uint32_t)N;
          r = (r + 1 == (uint32_t)R) ? 0 : r + 1;
          if (kind_f16)
            mma_f16_ss(d, ad0 + 2 * k, bd0 + 2 * k, idesc, 1);
          else
            mma_tf32_ss(d, ad0 + 2 * k, bd0 + 2 * k, idesc, 1);
        }
      }
      r = __shfl_sync(0xffffffffu, r, 0);
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar_mma);
    __syncwarp();
    long long t1 = clock64();
    while (!mbar_try_wait(&bar_mma, 0)) {
      if (clock64() - t1 > 2000000000LL) break;
    }
    long long t2 = clock64();
    if ((tid & 31) == 0) {
      cycles[0] = t1 - t0;
      cycles[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) {
    printf("cuTensorMapEncodeTiled not available\n");
    exit(2);
  }
  return (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

static float frand() { return (float)((double)rand() / RAND_MAX * 2.0 - 1.0); }

static int check_timeout() {
  int flag = 0;
  CK(cudaMemcpyFromSymbol(&flag, g_timeout_flag, sizeof(int)));
  if (flag) printf("TIMEOUT waiting on barrier %d (1 = TMA, 2 = MMA)\n", flag);
  return flag;
}

int main() {
  auto encode = get_encode();
  srand(4321);
  int fails = 0;
  float* d_out;
  CK(cudaMalloc(&d_out, 128 * 256 * 4));
  const size_t smem_bytes = 49152 + 32768;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));

  const int NB = 4, H = 8, W = 16, C = 64;
  std::vector<__half> hx((size_t)NB * H * W * C);
  std::vector<float> fx(hx.size());
  for (size_t i = 0; i < hx.size(); ++i) {
    hx[i] = __float2half(frand());
    fx[i] = __half2float(hx[i]);
  }
  __half* dx;
  CK(cudaMalloc(&dx, hx.size() * 2));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  auto X = [&](int n, int h, int x, int c) -> float {
    if (h < 0 || h >= H || x < 0 || x >= W) return 0.f;
    return fx[(((size_t)n * H + h) * W + x) * C + c];
  };

  for (int rowbytes : {128, 64, 32}) {
    const int kel = rowbytes / 2;
    for (int N : {64, 32, 16}) {
      std::vector<__half> hw((size_t)N * kel);
      std::vector<float> fw(hw.size());
      for (size_t i = 0; i < hw.size(); ++i) {
        hw[i] = __float2half(frand());
        fw[i] = __half2float(hw[i]);
      }
      __half* dw;
      CK(cudaMalloc(&dw, hw.size() * 2));
      CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
      const CUtensorMapSwizzle sw = rowbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                    : rowbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
      // variant 0: 8-column box per kw (F1); variants 1,2: one 10-column box, kw as a row offset (F2),
      // base offset 0 / (addr >> 7) & 7
      for (int variant = 0; variant < 3; ++variant) {
        const int bw = variant == 0 ? 8 : 10;
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)NB, (cuuint64_t)H};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)kel, (cuuint32_t)bw, 2, 10};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
          printf("encode failed rowbytes=%d variant=%d: %d\n", rowbytes, variant, (int)r);
          ++fails;
          continue;
        }
        int bad_cases = 0;
        double worst = 0;
        for (int x0 : {0, 8})
          for (int kw = 0; kw < 3; ++kw)
            for (int kh = 0; kh < 3; ++kh) {
              const int c0 = (C - kel), n0 = 2;
              ProbeArgs p{};
              p.coord[0] = c0;
              p.coord[2] = n0;
              p.coord[3] = -1;
              p.rowbytes = rowbytes;
              p.N = N;
              p.w = dw;
              p.d_out = d_out;
              if (variant == 0) {
                p.coord[1] = x0 + kw - 1;
                p.a_bytes = 10 * 2 * 8 * rowbytes;
                p.a_off = kh * 2 * 8 * rowbytes;
                p.sbo = 8 * rowbytes;
                p.base_off = 0;
              } else {
                p.coord[1] = x0 - 1;
                p.a_bytes = 10 * 2 * 10 * rowbytes;
                p.a_off = (kh * 2 * 10 + kw) * rowbytes;
                p.sbo = 10 * rowbytes;
                p.base_off = variant == 1 ? 0 : ((p.a_off >> 7) & 7);
              }
              CK(cudaMemset(d_out, 0xff, 128 * 256 * 4));
              probe_kernel<<<1, 128, smem_bytes>>>(tm, p);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) {
                printf("kernel error %s (rowbytes=%d N=%d variant=%d kw=%d kh=%d)\n", cudaGetErrorString(e), rowbytes, N, variant, kw, kh);
                return 3;
              }
              if (check_timeout()) return 3;
              std::vector<float> out(128 * N);
              CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
              double err = 0;
              for (int m = 0; m < 128; ++m)
                for (int oc = 0; oc < N; ++oc) {
                  int x = m % 8, nb = (m / 8) % 2, h = m / 16;
                  double s = 0;
                  for (int k = 0; k < kel; ++k) s += (double)X(n0 + nb, h + kh - 1, x0 + x + kw - 1, c0 + k) * (double)fw[oc * kel + k];
                  err = fmax(err, fabs(out[m * N + oc] - s));
                }
              worst = fmax(worst, err);
              if (!(err < 1e-3)) ++bad_cases;
            }
        const char* vn = variant == 0 ? "F1 8-col box per kw" : variant == 1 ? "F2 10-col box, base_off 0" : "F2 10-col box, base_off (a>>7)&7";
        printf("rowbytes=%3d N=%2d %-34s: %2d/18 bad, worst |err| %.3e %s\n", rowbytes, N, vn, bad_cases, worst,
               bad_cases == 0 ? "PASS" : "FAIL");
        if (bad_cases && variant == 0) ++fails;
      }
      cudaFree(dw);
    }
  }

  // ---------------- F3: issue rate ----------------------------------------------------------------
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, 16));
  for (int kind = 1; kind >= 0; --kind)
    for (int N : {16, 32, 64, 128, 256})
      for (int R : {1, 2, 3, 4, 6, 8}) {
        if (R * N > 512) continue;
        const int count = 4096;
        rate_kernel<<<1, 128, smem_bytes>>>(N, kind, count, R, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("rate kernel error %s\n", cudaGetErrorString(e));
          return 3;
        }
        long long c[2];
        CK(cudaMemcpy(c, d_cyc, 16, cudaMemcpyDeviceToHost));
        printf("F3 %s M=128 N=%3d R=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal compute %.1f, smem 128B/clk model %.1f)\n",
               kind ? "f16 K=16" : "tf32 K=8", N, R, (double)c[0] / count, (double)c[1] / count, N / 2.0, (128 + N) * 32 / 128.0);
      }
  printf("%s (%d failures)\n", fails ? "PROBE FAILED" : "PROBE OK", fails);
  return fails ? 1 : 0;
}
