// HBM write bandwidth by store pattern: what the pair-plane epilogues write per 128-pixel tile (16 map rows x 8
// pixels x 32 channels: 16 runs of 512 bytes, 4 KB apart, in each of two planes) against longer contiguous runs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o write_pattern_probe write_pattern_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

// tensor [n][H][W] pixels of `px_bytes`; a tile is th rows x tw pixels; tiles are walked in (n, ty, tx) order by
// persistent CTAs (tile = blockIdx.x + k * gridDim.x), each warp writes whole runs with 16-byte stores.
__global__ void write_tiles(uint4* __restrict__ plane0, uint4* __restrict__ plane1, int H, int W, int px16, int th, int tw,
                            long long num_tiles, int blocked) {
  const int tiles_x = W / tw, tiles_y = H / th;
  const int run16 = tw * px16;                 // 16-byte units per run
  const int per_tile = th * run16;
  // blocked: CTA c walks its own contiguous range of tiles; else tiles are dealt round-robin (t = c + k * grid)
  const long long per_cta = (num_tiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = blocked ? blockIdx.x * per_cta : blockIdx.x;
  const long long t_end = blocked ? (t_begin + per_cta < num_tiles ? t_begin + per_cta : num_tiles) : num_tiles;
  const long long t_step = blocked ? 1 : gridDim.x;
  // a warp owns whole tiles (like an epilogue warp group): warp w of the CTA takes every (blockDim/32)-th tile of the CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (long long t = t_begin + warp * t_step; t < t_end; t += t_step * nw) {
    const unsigned tu = (unsigned)t;
    const unsigned tx = tu % (unsigned)tiles_x, q = tu / (unsigned)tiles_x;
    const unsigned ty = q % (unsigned)tiles_y, n = q / (unsigned)tiles_y;
    const long long base = (((long long)n * H + (long long)ty * th) * W + (long long)tx * tw) * px16;
    const uint4 v = make_uint4(lane, tx, ty, n);
#pragma unroll 8
    for (int i = lane; i < per_tile; i += 32) {
      const int r = i / run16, c = i - r * run16;
      const long long off = base + (long long)r * W * px16 + c;
      plane0[off] = v;
      plane1[off] = v;
    }
  }
}

int main() {
  const int n = 12288, H = 64, W = 64, px16 = 4;  // 64 bytes per pixel and plane (32 channels fp16)
  const size_t plane_bytes = (size_t)n * H * W * px16 * 16;
  uint4 *p0, *p1;
  CK(cudaMalloc(&p0, plane_bytes));
  CK(cudaMalloc(&p1, plane_bytes));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  struct Cfg { int th, tw; };
  const Cfg cfgs[] = {{16, 8}, {8, 16}, {4, 32}, {2, 64}, {64, 64}};
  for (const Cfg& c : cfgs)
    for (int blocked : {0, 1}) {
      const int threads = 1024;
      const long long num_tiles = (long long)n * (H / c.th) * (W / c.tw);
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        write_tiles<<<148, threads>>>(p0, p1, H, W, px16, c.th, c.tw, num_tiles, blocked);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
      }
      printf("tile %2d rows x %2d px (runs of %4d B, %d per plane): %s  %.3f ms  %.0f GB/s\n", c.th, c.tw, c.tw * 64, c.th,
             blocked ? "blocked per CTA" : "round-robin    ", best, 2.0 * plane_bytes / best / 1e6);
    }
  return 0;
}
