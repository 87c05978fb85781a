// Hardware probe for the tensor path's assumptions (run on a B200 via gpurun; prints PASS/FAIL):
//   P1  TMA 4-D box with re-ordered dims (C, W, N, H), OOB zero fill, SWIZZLE_128B -> shared image
//   P2  tcgen05.mma kind::tf32 SS, M=128 N=64, K-major SW128 descriptors with a start offset (kh shift)
//       and non-default SBO; operand truncation vs rounding of the low 13 mantissa bits
//   P3  TMA 5-D map for stride-2 (inner dim = (w parity, C)), SBO = 2048
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../tf_image_compression_b200/csrc/tic_ptx.cuh"

using namespace tic::ptx;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e = (x);                                                                       \
    if (e != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);           \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

__device__ int g_timeout_flag = 0;
// bounded wait: a wrong byte count / descriptor must not hang the GPU box
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

struct ProbeArgs {
  int rank;
  int coord[5];
  uint32_t a_bytes;      // TMA box bytes
  uint32_t a_off;        // descriptor start offset inside the A region (tap shift)
  uint32_t sbo;          // stride between 8-row groups
  const float* w;        // [64][32] K-major weights (global)
  float* a_dump;         // raw shared image of the A region (a_bytes)
  float* d_out;          // [128][64]
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmap, ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // up to 40 KB
  uint8_t* sB = smem + 40960;               // 64 x 128 B
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 64);
    tmem_relinquish();
  }
  // B tile: K-major SW128 image written by hand (validates sw128_offset)
  for (int i = tid; i < 64 * 32; i += 128) {
    int r = i / 32, k = i % 32;
    *reinterpret_cast<float*>(sB + sw128_offset(r, k)) = p.w[r * 32 + k];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_expect_tx(&bar_tma, p.a_bytes);
    if (p.rank == 4)
      tma_load_4d(sA, &tmap, &bar_tma, p.coord[0], p.coord[1], p.coord[2], p.coord[3]);
    else
      tma_load_5d(sA, &tmap, &bar_tma, p.coord[0], p.coord[1], p.coord[2], p.coord[3], p.coord[4]);
  }
  if (!wait_bounded(&bar_tma, 0, 1)) return;
  for (uint32_t i = tid; i < p.a_bytes / 4; i += 128) p.a_dump[i] = reinterpret_cast<float*>(sA)[i];
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, 64);
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = make_smem_desc_sw128(smem_u32(sA) + p.a_off + k * 32, p.sbo);
      uint64_t bd = make_smem_desc_sw128(smem_u32(sB) + k * 32, 1024);
      mma_tf32_ss(tmem_base, ad, bd, idesc, k > 0);
    }
    tc_commit(&bar_mma);
  }
  if (!wait_bounded(&bar_mma, 0, 2)) return;
  tc_fence_after();
  for (int c = 0; c < 64; c += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 16; ++i) p.d_out[(warp * 32 + lane) * 64 + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
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

static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static float tf32_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x00000FFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

static float frand() { return (float)((double)rand() / RAND_MAX * 2.0 - 1.0); }

int main() {
  auto encode = get_encode();
  srand(1234);
  int fails = 0;
  std::vector<float> hw(64 * 32);
  for (auto& v : hw) v = frand();
  float* dw;
  CK(cudaMalloc(&dw, hw.size() * 4));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
  float *d_dump, *d_out;
  CK(cudaMalloc(&d_dump, 40960));
  CK(cudaMalloc(&d_out, 128 * 64 * 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960 + 8192));

  // ---------------- P1 / P2: stride-1, X[N=4][H=8][W=8][C=64], dims (C, W, N, H) -----------------
  {
    const int N = 4, H = 8, W = 8, C = 64;
    std::vector<float> hx((size_t)N * H * W * C);
    for (auto& v : hx) v = frand();
    float* dx;
    CK(cudaMalloc(&dx, hx.size() * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)N, (cuuint64_t)H};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)H * W * C * 4, (cuuint64_t)W * C * 4};
    cuuint32_t box[4] = {32, 8, 2, 10};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("P1 encode 4D: %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 2;
    for (int kw = 0; kw < 3; ++kw)
      for (int kh = 0; kh < 3; ++kh) {
        const int c0 = 32, n0 = 2;
        ProbeArgs p{};
        p.rank = 4;
        p.coord[0] = c0;
        p.coord[1] = kw - 1;
        p.coord[2] = n0;
        p.coord[3] = -1;
        p.a_bytes = 160 * 128;
        p.a_off = kh * 2 * 1024;
        p.sbo = 1024;
        p.w = dw;
        p.a_dump = d_dump;
        p.d_out = d_out;
        CK(cudaMemset(d_out, 0xff, 128 * 64 * 4));
        probe_kernel<<<1, 128, 40960 + 8192>>>(tm, p);
        CK(cudaDeviceSynchronize());
        {
          int flag = 0;
          CK(cudaMemcpyFromSymbol(&flag, g_timeout_flag, sizeof(int)));
          if (flag) {
            printf("TIMEOUT waiting on barrier %d (1 = TMA, 2 = MMA)\n", flag);
            return 3;
          }
        }
        std::vector<float> dump(160 * 32), out(128 * 64);
        CK(cudaMemcpy(dump.data(), d_dump, dump.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
        // expected shared image: row = (h_box * 2 + n_box) * 8 + x ; value X[n0+n_box][h_box-1][x+kw-1][c0+k] or 0
        auto X = [&](int n, int h, int x, int c) -> float {
          if (h < 0 || h >= H || x < 0 || x >= W) return 0.f;
          return hx[(((size_t)n * H + h) * W + x) * C + c];
        };
        int bad_layout = 0;
        for (int row = 0; row < 160; ++row)
          for (int k = 0; k < 32; ++k) {
            int x = row % 8, nb = (row / 8) % 2, hb = row / 16;
            float want = X(n0 + nb, hb - 1, x + kw - 1, c0 + k);
            float got = dump[sw128_offset(row, k) / 4];
            if (want != got) ++bad_layout;
          }
        // expected D: row m -> x = m%8, n = (m/8)%2, h = m/16 ; tap (kh, kw) only
        double err_t = 0, err_r = 0, ref_max = 0;
        for (int m = 0; m < 128; ++m)
          for (int oc = 0; oc < 64; ++oc) {
            int x = m % 8, nb = (m / 8) % 2, h = m / 16;
            double st = 0, sr = 0;
            for (int k = 0; k < 32; ++k) {
              float xv = X(n0 + nb, h + kh - 1, x + kw - 1, c0 + k);
              st += (double)tf32_trunc(xv) * (double)tf32_trunc(hw[oc * 32 + k]);
              sr += (double)tf32_rn(xv) * (double)tf32_rn(hw[oc * 32 + k]);
            }
            err_t = fmax(err_t, fabs(out[m * 64 + oc] - st));
            err_r = fmax(err_r, fabs(out[m * 64 + oc] - sr));
            ref_max = fmax(ref_max, fabs(st));
          }
        bool ok = bad_layout == 0 && (err_t < 1e-4 || err_r < 1e-4);
        printf("P1/P2 kw=%d kh=%d: layout mismatches %d; |D - trunc model| %.3e, |D - rn model| %.3e (max |D| %.2f) %s\n", kw, kh,
               bad_layout, err_t, err_r, ref_max, ok ? "PASS" : "FAIL");
        if (!ok) ++fails;
      }
    cudaFree(dx);
  }

  // ---------------- P3: stride-2, X[N=2][H=16][W=16][C=32], dims ((wpar,C), W2, hpar, N, H2) ------
  {
    const int N = 2, H = 16, W = 16, C = 32;
    std::vector<float> hx((size_t)N * H * W * C);
    for (auto& v : hx) v = frand();
    float* dx;
    CK(cudaMalloc(&dx, hx.size() * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)W / 2, 2, (cuuint64_t)N, (cuuint64_t)H / 2};
    cuuint64_t strides[4] = {(cuuint64_t)2 * C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, (cuuint64_t)2 * W * C * 4};
    cuuint32_t box[5] = {32, 8, 2, 2, 9};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("P3 encode 5D: %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 2;
    for (int kw = 0; kw < 3; ++kw)
      for (int kh = 0; kh < 3; ++kh) {
        ProbeArgs p{};
        p.rank = 5;
        p.coord[0] = (kw & 1) * C;  // w parity selects the inner-dim half
        p.coord[1] = kw >> 1;       // w2 start (output x0 = 0)
        p.coord[2] = 0;
        p.coord[3] = 0;
        p.coord[4] = 0;             // h2 start (output y0 = 0)
        p.a_bytes = 288 * 128;
        p.a_off = (kh == 0 ? 0 : kh == 1 ? 1024 : 2 * 2 * 1024);
        p.sbo = 2048;
        p.w = dw;
        p.a_dump = d_dump;
        p.d_out = d_out;
        probe_kernel<<<1, 128, 40960 + 8192>>>(tm, p);
        CK(cudaDeviceSynchronize());
        std::vector<float> out(128 * 64);
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
        auto X = [&](int n, int h, int x, int c) -> float {
          if (h < 0 || h >= H || x < 0 || x >= W) return 0.f;
          return hx[(((size_t)n * H + h) * W + x) * C + c];
        };
        double err_t = 0, err_r = 0;
        for (int m = 0; m < 128; ++m)
          for (int oc = 0; oc < 64; ++oc) {
            int ox = m % 8, nb = (m / 8) % 2, oy = m / 16;
            double st = 0, sr = 0;
            for (int k = 0; k < 32; ++k) {
              float xv = X(nb, 2 * oy + kh, 2 * ox + kw, k);
              st += (double)tf32_trunc(xv) * (double)tf32_trunc(hw[oc * 32 + k]);
              sr += (double)tf32_rn(xv) * (double)tf32_rn(hw[oc * 32 + k]);
            }
            err_t = fmax(err_t, fabs(out[m * 64 + oc] - st));
            err_r = fmax(err_r, fabs(out[m * 64 + oc] - sr));
          }
        bool ok = (err_t < 1e-4 || err_r < 1e-4);
        printf("P3 stride2 kw=%d kh=%d: |D - trunc| %.3e, |D - rn| %.3e %s\n", kw, kh, err_t, err_r, ok ? "PASS" : "FAIL");
        if (!ok) ++fails;
      }
    cudaFree(dx);
  }
  printf("%s (%d failures)\n", fails ? "PROBE FAILED" : "PROBE OK", fails);
  return fails ? 1 : 0;
}
