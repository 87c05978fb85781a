// C-ABI implementation (include/tic.h): handle, graph executor, chunked H2D -> compute -> D2H
// pipeline, and kernel dispatch.  No torch types, no exceptions across the boundary.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tic.h"
#include "tic_simt.cuh"
#include "tic_umma.cuh"
#include "tic_umma16.cuh"
#include "tic_first16.cuh"
#include "tic_fused16.cuh"
#include "tic_fused_enc16.cuh"
#include "tic_entropy.cuh"

using namespace tic;

namespace {

std::string g_create_error;

struct Layer {
  tic_layer_desc d{};
  float* w = nullptr;     // device [9][cin][cout]
  float* b = nullptr;     // device [cout]
  std::vector<float> hb;  // host copy of the bias (the fused kernels take it as launch constants)
  UmmaWeights uw;         // tensor-path operand images (built lazily from w)
  U16Weights uw16;        // fp16-pair operand images
  F16Weights fw16;        // fp16-pair first-layer (cin = 3) operand image
  FusedDecWeights fdw;    // operand images of this layer and the next one for the fused transposed-conv pair
  FusedEncWeights few;    // ... for the fused first two layers of an encoder
  bool loaded = false;
};

struct Graph {
  std::vector<Layer> layers;
  float mean[3] = {0, 0, 0}, stdv[3] = {1, 1, 1};
  bool has_norm = false;
  float* d_normlut = nullptr;  // [3][256]
};

struct IoSpec {
  int mode = IO_ACT;
  const void* in = nullptr;
  void* out = nullptr;
  Geo geo{};
};

}  // namespace

struct tic_codec {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_in[2]{}, ev_comp[2]{}, ev_out[2]{}, ev_enc[2]{}, ev_t0 = nullptr, ev_t1 = nullptr, ev_rt = nullptr;
  Graph g[3];
  int q = 0;
  float* d_symlut = nullptr;  // [256]
  unsigned long long* d_hist = nullptr;  // [256]
  int mode = TIC_COMPUTE_FP32;
  int chunk128 = 4096;
  // workspaces
  float* act[3] = {nullptr, nullptr, nullptr};  // rotating (sub-)chunk activations
  size_t act_bytes = 0;
  float* bnd[2] = {nullptr, nullptr};           // chunk-level tensors at layer-group boundaries
  size_t bnd_bytes = 0;
  void* stage_in[2] = {nullptr, nullptr};
  void* stage_out[2] = {nullptr, nullptr};
  void* stage_sym[2] = {nullptr, nullptr};      // round trips: the symbols between the two graphs
  size_t stage_in_bytes = 0, stage_out_bytes = 0, stage_sym_bytes = 0;
  int64_t launches = 0;
  float last_ms = 0.f;
  bool profile = false;
  struct ProfRec {
    int graph, layer;
    cudaEvent_t e0, e1;
  };
  std::vector<ProfRec> prof_pending;
  std::vector<float> prof_ms[3];
  std::vector<int64_t> prof_n[3];
  int num_sms = 148;
  unsigned int* h_oflow = nullptr;  // host-mapped sticky flags: [0] a pair-plane split saw |x| >= 65504 (fp16 range),
                                    // [1] entropy stage (1: symbol outside the table / of zero width, 2: slot too small)
  unsigned int* d_oflow = nullptr;  // their device address
  uint32_t* d_cum = nullptr;        // entropy stage: the call's cumulative-frequency table [257]
  uint8_t* d_seg = nullptr;         // entropy stage: worst-case slots of a segmented call's pieces (grow-only)
  long long* d_seg_bytes = nullptr; //   and their stored lengths
  int64_t seg_cap = 0;              //   pieces both hold
  cudaEvent_t ev_switch = nullptr;  // orders the old stream before the new one in tic_set_stream
  std::string err;
};

namespace {

int fail(tic_codec* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

#define TIC_CUDA(h, expr)                                                                         \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail((h), TIC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// Error paths of the staged (host-buffer) drivers: nothing may still be copying into / out of the caller's buffers
// or the staging sets when the call returns.
void quiesce(tic_codec* h) {
  cudaStreamSynchronize(h->s_h2d);
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->s_d2h);
}

const char* kOverflowMsg =
    "fp16-pair tensor mode: an activation reached the fp16 range limit (|x| >= 65504); results are invalid. "
    "Use compute mode fp32 or 3xtf32 for these weights (tic_set_compute_mode clears the flag)";

// Sticky fp16-range status: checked at the start of every hot-path call and after every call that synchronises.
int overflow_status(tic_codec* h) {
  if (h->h_oflow && *reinterpret_cast<volatile unsigned int*>(h->h_oflow)) return fail(h, TIC_ERR_UNSUPPORTED, "%s", kOverflowMsg);
  if (h->h_oflow) {
    const unsigned int es = *reinterpret_cast<volatile unsigned int*>(h->h_oflow + 1);
    if (es) {
      *reinterpret_cast<volatile unsigned int*>(h->h_oflow + 1) = 0u;  // reported once
      return fail(h, TIC_ERR_INVALID, es == 1 ? "entropy stage: symbol outside the table or of zero probability"
                                                : "entropy stage: output slot too small (use tic_entropy_bound)");
    }
  }
  return TIC_OK;
}

int pow2ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// TF SAME: out = ceil(in / s); pad_total = max((out-1)*s + 3 - in, 0); before = total / 2
void same_pad(int in, int s, int* out, int* before) {
  *out = (in + s - 1) / s;
  int total = std::max((*out - 1) * s + 3 - in, 0);
  *before = total / 2;
}

// ---- SIMT launch planning -------------------------------------------------------------------
template <int OCB, int S>
int launch_conv_t(tic_codec* h, const LayerArgs& a, const SmemPlan& sp, dim3 grid, size_t smem) {
  auto k = conv3x3_simt_kernel<OCB, S>;
  static SmemAttrCache cache;
  TIC_CUDA(h, cache.ensure(reinterpret_cast<const void*>(k), smem, 48 * 1024));
  k<<<grid, kThreads, smem, h->stream>>>(a, sp);
  return TIC_OK;
}
template <int OCB>
int launch_deconv_t(tic_codec* h, const LayerArgs& a, const SmemPlan& sp, dim3 grid, size_t smem) {
  auto k = deconv3x3_simt_kernel<OCB>;
  static SmemAttrCache cache;
  TIC_CUDA(h, cache.ensure(reinterpret_cast<const void*>(k), smem, 48 * 1024));
  k<<<grid, kThreads, smem, h->stream>>>(a, sp);
  return TIC_OK;
}

// Pick the row pitch so that the (up to four) row groups a warp spans land in distinct banks.
int pick_pitch(int iw, int lanes_x, int row_words /* words between a warp's row groups, per pitch unit */, int s) {
  if (lanes_x * s >= 32) return iw | 1;
  for (int c = iw; c < iw + 32; ++c)
    if ((row_words * c) % 32 == (lanes_x * s) % 32) return c;
  return iw;
}

int launch_simt(tic_codec* h, LayerArgs a, int kind, int stride) {
  SmemPlan sp{};
  int ocb, grid_y;
  if (kind == TIC_CONV) {
    ocb = a.cout >= 64 ? 64 : a.cout > 16 ? 32 : a.cout > 8 ? 16 : 8;
    const int ppc = 8192 / ocb;  // output pixels per CTA
    a.TW = std::min(32, pow2ceil(a.wout));
    a.TH = std::min(std::max(4, pow2ceil(a.hout)), ppc / a.TW);
    a.TP = ppc / (a.TW * a.TH);
    a.tiles_x = (a.wout + a.TW - 1) / a.TW;
    a.tiles_y = (a.hout + a.TH - 1) / a.TH;
    sp.ih = (a.TH - 1) * stride + 3;
    sp.iw = (a.TW - 1) * stride + 3;
    sp.iwp = pick_pitch(sp.iw, a.TW, 4 * stride, stride);
  } else {
    ocb = a.cout >= 32 ? 32 : a.cout > 8 ? 16 : a.cout > 4 ? 8 : 4;
    const int ipc = 4096 / ocb;  // input pixels per CTA
    a.TW = std::min(32, pow2ceil(a.win));
    a.TH = std::min(std::max(4, pow2ceil(a.hin)), ipc / a.TW);
    a.TP = ipc / (a.TW * a.TH);
    a.tiles_x = (a.win + a.TW - 1) / a.TW;
    a.tiles_y = (a.hin + a.TH - 1) / a.TH;
    sp.ih = a.TH + 1;
    sp.iw = a.TW + 1;
    sp.iwp = pick_pitch(sp.iw, a.TW, 4, 1);
  }
  sp.cstr = sp.ih * sp.iwp;
  sp.pstr = kIcc * sp.cstr;
  {
    // lanes of one warp may straddle patches: offset consecutive patches to fresh banks
    const int lanes_per_patch = a.TW * a.TH / 4;
    if (lanes_per_patch < 32 && a.TP > 1) {
      const int want = (lanes_per_patch * (kind == TIC_CONV ? stride : 1)) % 32;
      while (sp.pstr % 32 != want) ++sp.pstr;
    }
  }
  grid_y = (a.cout + ocb - 1) / ocb;
  const long long groups = (a.n + a.TP - 1) / a.TP;
  const long long gx = (long long)a.tiles_x * a.tiles_y * groups;
  if (gx <= 0 || gx > 0x7fffffffLL) return fail(h, TIC_ERR_INVALID, "grid too large (%lld)", gx);
  dim3 grid((unsigned)gx, (unsigned)grid_y);
  const size_t smem = ((size_t)a.TP * sp.pstr + 9 * kIcc * ocb) * sizeof(float);
  if (smem > 200 * 1024) return fail(h, TIC_ERR_UNSUPPORTED, "tile needs %zu B shared memory", smem);
  int rc = TIC_OK;
  if (kind == TIC_CONV) {
    if (stride == 1) {
      switch (ocb) {
        case 64: rc = launch_conv_t<64, 1>(h, a, sp, grid, smem); break;
        case 32: rc = launch_conv_t<32, 1>(h, a, sp, grid, smem); break;
        case 16: rc = launch_conv_t<16, 1>(h, a, sp, grid, smem); break;
        default: rc = launch_conv_t<8, 1>(h, a, sp, grid, smem); break;
      }
    } else {
      switch (ocb) {
        case 64: rc = launch_conv_t<64, 2>(h, a, sp, grid, smem); break;
        case 32: rc = launch_conv_t<32, 2>(h, a, sp, grid, smem); break;
        case 16: rc = launch_conv_t<16, 2>(h, a, sp, grid, smem); break;
        default: rc = launch_conv_t<8, 2>(h, a, sp, grid, smem); break;
      }
    }
  } else {
    switch (ocb) {
      case 32: rc = launch_deconv_t<32>(h, a, sp, grid, smem); break;
      case 16: rc = launch_deconv_t<16>(h, a, sp, grid, smem); break;
      case 8: rc = launch_deconv_t<8>(h, a, sp, grid, smem); break;
      default: rc = launch_deconv_t<4>(h, a, sp, grid, smem); break;
    }
  }
  if (rc != TIC_OK) return rc;
  h->launches++;
  TIC_CUDA(h, cudaGetLastError());
  return TIC_OK;
}

// ---- graph executor -------------------------------------------------------------------------
struct Shape {
  int h, w, c;
};

int graph_out_shape(const Graph& g, Shape in, Shape* out, size_t* max_act_elems, int limit = -1) {
  Shape s = in;
  size_t mx = 0;
  const size_t L = limit > 0 ? std::min<size_t>((size_t)limit, g.layers.size()) : g.layers.size();
  for (size_t i = 0; i < L; ++i) {
    const tic_layer_desc& d = g.layers[i].d;
    if (d.cin != s.c) return -1;
    if (d.kind == TIC_CONV) {
      int pb;
      same_pad(s.h, d.stride, &s.h, &pb);
      same_pad(s.w, d.stride, &s.w, &pb);
    } else {
      s.h *= 2;
      s.w *= 2;
    }
    s.c = d.cout;
    if (i + 1 < L) mx = std::max(mx, (size_t)s.h * s.w * s.c);
  }
  *out = s;
  if (max_act_elems) *max_act_elems = mx;
  return 0;
}

int check_graph_ready(tic_codec* h, int gi, bool need_norm) {
  Graph& g = h->g[gi];
  if (g.layers.empty()) return fail(h, TIC_ERR_STATE, "graph %d not configured (tic_set_graph)", gi);
  for (size_t i = 0; i < g.layers.size(); ++i)
    if (!g.layers[i].loaded) return fail(h, TIC_ERR_STATE, "graph %d layer %zu has no weights (tic_load_weights)", gi, i);
  if (need_norm && !g.has_norm) return fail(h, TIC_ERR_STATE, "graph %d has no normalisation (tic_set_norm)", gi);
  return TIC_OK;
}

// Layer groups of one graph execution (fp16-pair mode).  With every layer at (or near) the HBM roofline the
// remaining lever is to keep inter-layer tensors in the 126 MB L2: consecutive layers run over a SUB-chunk of
// m patches so that the tensors between them (m x bytes per patch) fit the L2 budget, and only the tensor at a
// group boundary is written out for the whole chunk.  A group is closed when the next layer would push m so
// low that some layer of the group could not fill the 148 SMs twice (model_0: {encode_0, encode_1} at m = 111
// and {encode_2 .. encode_4} at m = 888; decoder {decode_4 .. decode_2} and {decode_1, decode_0}).
struct LayerShape {
  int hin, win, cin, hout, wout, cout, pad_t, pad_l;
};
struct LayerGroup {
  int first, last, m;
};

long long l2_budget_bytes_per_group() {
  // 0 (default, and the only value outside -DTIC_ABLATE builds) disables grouping.  Measured on B200 (model_0, 12288
  // patches): 56 MB -> 20.3 ms per encode+decode step against 14.0 ms ungrouped: each extra (20 us) launch costs ~8 us
  // of prologue (weight tiles, TMEM, cluster sync) and tail.  The schedule pays off only once a group runs as one
  // persistent kernel.
  const long long v = tic_env_int("TIC_L2_BUDGET_MB", 0);
  return (v < 0 ? 0 : std::min<long long>(v, 4096)) << 20;
}

std::vector<LayerGroup> plan_groups(const Graph& g, const std::vector<LayerShape>& sh, int L, int n, bool enable) {
  std::vector<LayerGroup> groups;
  const long long budget = l2_budget_bytes_per_group();
  if (!enable || budget <= 0 || n < 64) {
    groups.push_back({0, L - 1, n});
    return groups;
  }
  auto tiles_pp = [&](int i) {
    const bool dc = g.layers[i].d.kind == TIC_DECONV;
    return (double)(dc ? sh[i].hin * sh[i].win : sh[i].hout * sh[i].wout) / 128.0;
  };
  auto out_bytes = [&](int i) { return (long long)sh[i].hout * sh[i].wout * sh[i].cout * 4; };
  const double min_tiles = 296.0;  // two tiles per SM (one CTA pair per TPC, twice)
  int first = 0;
  while (first < L) {
    int last = first;
    long long tmax = 0;
    bool in_res = g.layers[first].d.res_begin && !g.layers[first].d.res_end;
    while (last + 1 < L) {
      const long long tnew = std::max(tmax, out_bytes(last));
      const long long mnew = std::min<long long>(n, std::max<long long>(1, budget / std::max<long long>(tnew, 1)));
      bool ok = true;
      if (mnew < n)
        for (int j = first; j <= last + 1; ++j) ok = ok && tiles_pp(j) * (double)mnew >= min_tiles;
      if (!ok && !in_res) break;
      ++last;
      tmax = tnew;
      if (g.layers[last].d.res_begin) in_res = true;
      if (g.layers[last].d.res_end) in_res = false;
    }
    long long m = tmax > 0 ? std::min<long long>(n, std::max<long long>(1, budget / tmax)) : n;
    if (m < n) {
      // a whole number of waves for the layer with the fewest tiles: m a multiple of q patches
      double tp_min = 1e30;
      for (int j = first; j <= last; ++j) tp_min = std::min(tp_min, tiles_pp(j));
      const long long q = std::max<long long>(1, (long long)(296.0 / tp_min + 0.5));
      long long mr = ((m + q / 2) / q) * q;
      if (mr < q) mr = q;
      if (mr * tmax > budget + budget / 4) mr = std::max<long long>(q, (m / q) * q);
      m = std::min<long long>(n, mr);
    }
    groups.push_back({first, last, (int)m});
    first = last + 1;
  }
  return groups;
}

int ensure_buffers(tic_codec* h, float** bufs, int count, size_t* have, size_t need) {
  if (need <= *have) return TIC_OK;
  TIC_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < count; ++i) {
    if (bufs[i]) cudaFree(bufs[i]);
    bufs[i] = nullptr;
  }
  *have = 0;
  for (int i = 0; i < count; ++i) {
    cudaError_t e = cudaMalloc((void**)&bufs[i], need);
    if (e != cudaSuccess) return fail(h, TIC_ERR_NOMEM, "cudaMalloc(%zu) for activations failed: %s", need, cudaGetErrorString(e));
  }
  *have = need;
  return TIC_OK;
}

// Run graph gi over n patches whose first-layer input map is h0 x w0 x c0.
int run_graph(tic_codec* h, int gi, const IoSpec& io_in, const IoSpec& io_out, int n, int h0, int w0, int limit = -1) {
  Graph& g = h->g[gi];
  const int L = limit > 0 ? std::min(limit, (int)g.layers.size()) : (int)g.layers.size();
  std::vector<LayerShape> sh(L);
  {
    Shape s{h0, w0, g.layers[0].d.cin};
    for (int i = 0; i < L; ++i) {
      const tic_layer_desc& d = g.layers[i].d;
      if (d.cin != s.c) return fail(h, TIC_ERR_INVALID, "layer channel chain is inconsistent");
      LayerShape& t = sh[i];
      t.hin = s.h;
      t.win = s.w;
      t.cin = s.c;
      t.pad_t = t.pad_l = 0;
      if (d.kind == TIC_CONV) {
        same_pad(s.h, d.stride, &t.hout, &t.pad_t);
        same_pad(s.w, d.stride, &t.wout, &t.pad_l);
      } else {
        t.hout = 2 * s.h;
        t.wout = 2 * s.w;
      }
      t.cout = d.cout;
      s = Shape{t.hout, t.wout, t.cout};
    }
  }
  const bool pair16 = h->mode == TIC_COMPUTE_TENSOR_F16X3;  // intermediate activations are fp16 pair planes
  // pair mode: f32 / symbol inputs of a tensor-capable first layer are split into a workspace buffer first
  bool split_input = false;
  if (pair16 && (io_in.mode == IO_ACT || io_in.mode == IO_U8_SYMLUT)) {
    LayerArgs t{};
    const tic_layer_desc& d0 = g.layers[0].d;
    t.n = n;
    t.hin = sh[0].hin;
    t.win = sh[0].win;
    t.cin = sh[0].cin;
    t.hout = sh[0].hout;
    t.wout = sh[0].wout;
    t.cout = sh[0].cout;
    t.in_mode = IO_ACT16;
    t.out_mode = L == 1 ? io_out.mode : IO_ACT16;
    split_input = u16_supported(t, d0.kind, d0.stride);
  }
  const std::vector<LayerGroup> groups = plan_groups(g, sh, L, n, pair16);
  // workspaces: three rotating sub-chunk buffers, two chunk-level group-boundary buffers
  size_t intra = 0, bnd = 0;
  for (size_t gx = 0; gx < groups.size(); ++gx) {
    const LayerGroup& G = groups[gx];
    for (int i = G.first; i < G.last; ++i) intra = std::max(intra, (size_t)G.m * sh[i].hout * sh[i].wout * sh[i].cout);
    if (gx == 0 && split_input) intra = std::max(intra, (size_t)G.m * sh[0].hin * sh[0].win * sh[0].cin);
    if (gx + 1 < groups.size()) bnd = std::max(bnd, (size_t)n * sh[G.last].hout * sh[G.last].wout * sh[G.last].cout);
  }
  int rc0 = ensure_buffers(h, h->act, 3, &h->act_bytes, intra * sizeof(float));
  if (rc0 != TIC_OK) return rc0;
  rc0 = ensure_buffers(h, h->bnd, 2, &h->bnd_bytes, bnd * sizeof(float));
  if (rc0 != TIC_OK) return rc0;

  for (size_t gx = 0; gx < groups.size(); ++gx) {
    const LayerGroup& G = groups[gx];
    const bool first_group = gx == 0, last_group = gx + 1 == groups.size();
    for (int s0 = 0; s0 < n; s0 += G.m) {
      const int ns = std::min(G.m, n - s0);
      int cur = -1;        // act[] index holding the current activation (-1: the group's input)
      int res_buf = -1;    // act[] index of the live residual source (-1: none, or not an act[] buffer)
      const void* res_ptr = nullptr;
      long long res_lo = 0;
      if (first_group && split_input) {
        const long long per = (long long)sh[0].hin * sh[0].win * sh[0].cin;
        const long long count = (long long)ns * per;
        __half* hi = reinterpret_cast<__half*>(h->act[0]);
        const int blocks = (int)std::min<long long>((count + 255) / 256, (long long)h->num_sms * 8);
        if (io_in.mode == IO_ACT)
          u16_split_f32_kernel<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const float*>(io_in.in) + (long long)s0 * per, hi,
                                                             hi + count, count, h->d_oflow);
        else
          u16_split_symlut_kernel<<<blocks, 256, 0, h->stream>>>(
              reinterpret_cast<const uint8_t*>(io_in.in) + ((long long)io_in.geo.n0 + s0) * per, h->d_symlut, h->q, hi, hi + count, count,
              h->d_oflow);
        h->launches++;
        TIC_CUDA(h, cudaGetLastError());
        cur = 0;
      }
      for (int i = G.first; i <= G.last; ++i) {
        Layer& ly = g.layers[i];
        const tic_layer_desc& d = ly.d;
        const LayerShape& t = sh[i];
        LayerArgs a{};
        a.n = ns;
        a.hin = t.hin;
        a.win = t.win;
        a.cin = t.cin;
        a.hout = t.hout;
        a.wout = t.wout;
        a.cout = t.cout;
        a.pad_t = t.pad_t;
        a.pad_l = t.pad_l;
        a.act = d.act;
        a.wgt = ly.w;
        a.bias = ly.b;
        a.q = h->q;
        a.hist = h->d_hist;
        a.oflow = h->d_oflow;
        a.dbg = tic_env_int("TIC_DBG", 0);  // -DTIC_ABLATE builds only
        for (int c = 0; c < 3; ++c) {
          a.mean[c] = g.mean[c];
          a.stdv[c] = g.stdv[c];
        }
        const long long ein = (long long)t.hin * t.win * t.cin, eout = (long long)t.hout * t.wout * t.cout;
        // ---- input
        if (i == G.first && cur < 0) {
          if (first_group) {
            a.in_mode = io_in.mode;
            a.geo = io_in.geo;
            a.geo.n0 += s0;
            a.lut = (io_in.mode == IO_U8_NORM) ? g.d_normlut : (io_in.mode == IO_U8_SYMLUT ? h->d_symlut : nullptr);
            a.in = io_in.mode == IO_ACT ? (const void*)(reinterpret_cast<const float*>(io_in.in) + (long long)s0 * ein) : io_in.in;
          } else {
            a.in = reinterpret_cast<const __half*>(h->bnd[(gx - 1) & 1]) + (long long)s0 * ein;
            a.in_mode = IO_ACT16;
            a.in_lo_off = (long long)n * ein;
          }
        } else {
          a.in = h->act[cur];
          a.in_mode = pair16 ? IO_ACT16 : IO_ACT;
          a.in_lo_off = (long long)ns * ein;
        }
        if (d.res_begin) {
          res_buf = cur;
          res_ptr = a.in;
          res_lo = a.in_lo_off;
        }
        // ---- output: a workspace buffer that is neither the input nor the live residual
        int ob = -1;
        if (i == G.last) {
          if (last_group) {
            a.out_mode = io_out.mode;
            a.out = io_out.mode == IO_ACT ? (void*)(reinterpret_cast<float*>(io_out.out) + (long long)s0 * eout) : io_out.out;
            // a single-layer graph shares one geo between prologue and epilogue only if both need it
            if (!(first_group && i == G.first && io_out.mode == IO_ACT)) {
              a.geo = io_out.geo;
              a.geo.n0 += s0;
            }
          } else {
            a.out = reinterpret_cast<__half*>(h->bnd[gx & 1]) + (long long)s0 * eout;
            a.out_mode = IO_ACT16;
            a.out_lo_off = (long long)n * eout;
          }
        } else {
          for (int b = 0; b < 3; ++b)
            if (b != cur && b != res_buf) {
              ob = b;
              break;
            }
          a.out = h->act[ob];
          a.out_mode = pair16 ? IO_ACT16 : IO_ACT;
          a.out_lo_off = (long long)ns * eout;
        }
        if (d.res_end) {
          if (!res_ptr) return fail(h, TIC_ERR_INVALID, "res_end without res_begin at layer %d", i);
          a.res = reinterpret_cast<const float*>(res_ptr);
          a.res16 = pair16 ? 1 : 0;
          a.res_lo_off = res_lo;
        }
        int rc;
        tic_codec::ProfRec pr{gi, i, nullptr, nullptr};
        if (h->profile) {
          cudaEventCreate(&pr.e0);
          cudaEventCreate(&pr.e1);
          cudaEventRecord(pr.e0, h->stream);
        }
        // the last two transposed convs of a decoder (32 -> 32 -> 3 into the image) run as one back-to-back kernel: the
        // tensor between them never leaves shared memory (tic_fused16.cuh)
        bool fused = false;
        if (pair16 && last_group && i + 1 == G.last && i + 1 == L - 1 && !d.res_begin && !d.res_end && !g.layers[i + 1].d.res_begin &&
            !g.layers[i + 1].d.res_end && tic_env_int("TIC_FUSE_DEC", 1) != 0) {
          Layer& ly2 = g.layers[i + 1];
          const LayerShape& t2 = sh[i + 1];
          LayerArgs a2 = a;
          a2.hin = t2.hin;
          a2.win = t2.win;
          a2.cin = t2.cin;
          a2.hout = t2.hout;
          a2.wout = t2.wout;
          a2.cout = t2.cout;
          a2.pad_t = t2.pad_t;
          a2.pad_l = t2.pad_l;
          a2.act = ly2.d.act;
          a2.wgt = ly2.w;
          a2.bias = ly2.b;
          a2.res = nullptr;
          a2.in = nullptr;
          a2.in_mode = IO_ACT16;
          a2.out_mode = io_out.mode;
          a2.out = io_out.mode == IO_ACT ? nullptr : io_out.out;
          a2.geo = io_out.geo;
          a2.geo.n0 += s0;
          if (a2.out && fused_dec_supported(a, d.kind, a2, ly2.d.kind)) {
            int nl = 0;
            rc = launch_fused_dec(h->stream, a, a2, ly.w, ly2.w, ly.hb.data(), ly2.hb.data(), &ly.fdw, h->num_sms, &h->err, &nl);
            h->launches += nl;
            fused = true;
          }
        }
        // ... and the first two layers of an encoder (u8 image -> 3 -> 32 -> 32, both stride 2): tic_fused_enc16.cuh
        if (!fused && pair16 && first_group && i == 0 && cur < 0 && i + 1 < G.last && !g.layers[1].d.res_begin && !g.layers[1].d.res_end &&
            tic_env_int("TIC_FUSE_ENC", 1) != 0) {
          Layer& ly2 = g.layers[1];
          const LayerShape& t2 = sh[1];
          LayerArgs a2 = a;
          a2.hin = t2.hin;
          a2.win = t2.win;
          a2.cin = t2.cin;
          a2.hout = t2.hout;
          a2.wout = t2.wout;
          a2.cout = t2.cout;
          a2.pad_t = t2.pad_t;
          a2.pad_l = t2.pad_l;
          a2.act = ly2.d.act;
          a2.wgt = ly2.w;
          a2.bias = ly2.b;
          a2.res = nullptr;
          a2.in = nullptr;
          a2.in_mode = IO_ACT16;
          a2.out_mode = IO_ACT16;
          a2.out = h->act[0];
          a2.out_lo_off = (long long)ns * t2.hout * t2.wout * t2.cout;
          if (fused_enc_supported(a, d.kind, d.stride, a2, ly2.d.kind, ly2.d.stride)) {
            int nl = 0;
            rc = launch_fused_enc(h->stream, a, a2, ly.w, ly2.w, ly.hb.data(), &ly.few, h->num_sms, &h->err, &nl);
            h->launches += nl;
            fused = true;
            ob = 0;  // the pair's output lives in act[0]
          }
        }
        if (fused) {
          ++i;  // the next layer ran inside this launch (its time is reported with this layer)
        } else if (pair16 && f16_first_supported(a, d.kind, d.stride)) {
          int nl = 0;
          rc = launch_first16(h->stream, a, d.stride, ly.w, &ly.fw16, h->num_sms, &h->err, &nl);
          h->launches += nl;
        } else if (pair16 && u16_supported(a, d.kind, d.stride)) {
          int nl = 0;
          rc = launch_u16(h->stream, a, d.kind, d.stride, ly.w, &ly.uw16, h->num_sms, &h->err, &nl);
          h->launches += nl;
        } else if (h->mode != TIC_COMPUTE_FP32 && !pair16 && umma_supported(a, d.kind, d.stride)) {
          int nl = 0;
          rc = launch_umma(h->stream, a, d.kind, d.stride, ly.w, &ly.uw, h->mode == TIC_COMPUTE_TENSOR_3XTF32, h->num_sms,
                           &h->err, &nl);
          h->launches += nl;
        } else {
          rc = launch_simt(h, a, d.kind, d.stride);
        }
        if (h->profile) {
          cudaEventRecord(pr.e1, h->stream);
          h->prof_pending.push_back(pr);
        }
        if (rc != TIC_OK) return rc < 0 && rc >= TIC_ERR_UNSUPPORTED ? rc : TIC_ERR_CUDA;
        if (d.res_end) {
          res_buf = -1;
          res_ptr = nullptr;
        }
        cur = ob;
      }
    }
  }
  return TIC_OK;
}

int patches_per_chunk(const tic_codec* h, int P) {
  double scale = (128.0 / P) * (128.0 / P);
  int c = (int)(h->chunk128 * scale);
  return std::max(1, c);
}

// Host-staged calls overlap H2D, kernels and D2H chunk by chunk; what does not overlap is the first chunk's H2D and
// the last chunk's kernels + D2H, so chunks are equal-sized and small: 16 of them as long as a chunk keeps >= 768
// patches of 128x128 (measured on the 64-image round trip: 8 / 12 / 16 / 24 chunks -> 12.8 / 12.9 / 13.2 / 13.1
// Gpixel/s).  A ramped schedule (chunks doubling from 384 patches to the workspace chunk and halving again:
// 2, 4, 8, 18, 18, 8, 4, 2 images) was measured SLOWER (11.3 Gpixel/s): the round trip is bound by the D2H direction
// running next to the H2D (654 MB at 50 GB/s under duplex load, tools/pcie_probe.py), and large chunks leave it idle
// at both ends.
// TIC_HOST_RAMP=1 selects the ramp, TIC_HOST_CHUNKS the equal-chunk count.
std::vector<int64_t> host_chunk_schedule(int64_t units, int64_t cap, double patches128_per_unit) {
  const bool ramp = tic_env_set("TIC_HOST_RAMP");                      // -DTIC_ABLATE builds only
  const int want = std::max(1, tic_env_int("TIC_HOST_CHUNKS", 16));   // -DTIC_ABLATE builds only
  cap = std::max<int64_t>(1, std::min(cap, units));
  std::vector<int64_t> out;
  if (!ramp) {
    int64_t chunks = (units + cap - 1) / cap;
    const int64_t by_size = (int64_t)((double)units * patches128_per_unit / 768.0);
    chunks = std::max<int64_t>(chunks, std::min<int64_t>(want, by_size));
    chunks = std::max<int64_t>(1, std::min<int64_t>(chunks, units));
    const int64_t upc = (units + chunks - 1) / chunks;
    for (int64_t u0 = 0; u0 < units; u0 += upc) out.push_back(std::min<int64_t>(upc, units - u0));
    return out;
  }
  std::vector<int64_t> up;
  int64_t ramp_units = 0;
  int64_t v = std::max<int64_t>(1, (int64_t)(384.0 / std::max(patches128_per_unit, 1e-9) + 0.999));
  while (v < cap && 2 * (ramp_units + v) <= units / 2) {
    up.push_back(v);
    ramp_units += v;
    v *= 2;
  }
  out = up;
  const int64_t rem = units - 2 * ramp_units;
  const int64_t nmid = (rem + cap - 1) / cap;
  for (int64_t i = 0, left = rem; i < nmid; ++i) {
    const int64_t c = (left + (nmid - i) - 1) / (nmid - i);
    out.push_back(c);
    left -= c;
  }
  for (auto it = up.rbegin(); it != up.rend(); ++it) out.push_back(*it);
  return out;
}

// Generic chunked driver.  `units` are processed `upc` (units per chunk) at a time; a unit is a
// patch or a whole image.  in_unit_bytes / out_unit_bytes describe the caller's buffers; run()
// executes one chunk given device pointers (chunk-local when staged from host) and the unit offset.
template <typename RunFn>
int drive(tic_codec* h, int mem, const void* in, void* out, int64_t units, int64_t upc, size_t in_unit_bytes,
          size_t out_unit_bytes, bool inout_same, RunFn run, double patches128_per_unit = 1.0) {
  if (units <= 0) return TIC_OK;
  TIC_CUDA(h, cudaSetDevice(h->device));
  if (int st = overflow_status(h)) return st;
  TIC_CUDA(h, cudaEventRecord(h->ev_t0, h->stream));
  if (mem == TIC_MEM_DEVICE) {
    upc *= 4;  // no staging to overlap with: fewer, longer launch sequences (up to 16384 patches of 128x128 each)
    upc = (units + (units + upc - 1) / upc - 1) / ((units + upc - 1) / upc);  // equal-sized sequences
    for (int64_t u0 = 0; u0 < units; u0 += upc) {
      int64_t nu = std::min<int64_t>(upc, units - u0);
      int rc = run(in, out, u0, nu, /*local=*/false);
      if (rc != TIC_OK) return rc;
    }
    TIC_CUDA(h, cudaEventRecord(h->ev_t1, h->stream));
    return TIC_OK;
  }
  // host buffers: double-buffered staging, copies on side streams
  const std::vector<int64_t> sched = host_chunk_schedule(units, upc, patches128_per_unit);
  upc = *std::max_element(sched.begin(), sched.end());
  const size_t in_need = (size_t)upc * in_unit_bytes, out_need = (size_t)upc * out_unit_bytes;
  if (h->stage_in_bytes < in_need) {
    for (int b = 0; b < 2; ++b) {
      if (h->stage_in[b]) cudaFree(h->stage_in[b]);
      h->stage_in[b] = nullptr;
    }
    h->stage_in_bytes = 0;
    for (int b = 0; b < 2; ++b) TIC_CUDA(h, cudaMalloc(&h->stage_in[b], in_need));
    h->stage_in_bytes = in_need;
  }
  if (!inout_same && h->stage_out_bytes < out_need) {
    for (int b = 0; b < 2; ++b) {
      if (h->stage_out[b]) cudaFree(h->stage_out[b]);
      h->stage_out[b] = nullptr;
    }
    h->stage_out_bytes = 0;
    for (int b = 0; b < 2; ++b) TIC_CUDA(h, cudaMalloc(&h->stage_out[b], out_need));
    h->stage_out_bytes = out_need;
  }
  // the pipeline proper runs inside a lambda: ANY failure (a CUDA call, a layer launch) drains the three streams before
  // the error is returned, so no copy is still in flight into or out of the caller's buffers
  const int rc_pipe = [&]() -> int {
  int64_t u0 = 0;
  for (int64_t idx = 0; idx < (int64_t)sched.size(); u0 += sched[idx], ++idx) {
    const int b = (int)(idx & 1);
    const int64_t nu = sched[idx];
    // staging buffer b is free once the D2H (or compute) of chunk idx-2 finished
    if (idx >= 2) {
      TIC_CUDA(h, cudaStreamWaitEvent(h->s_h2d, inout_same ? h->ev_out[b] : h->ev_comp[b], 0));
    }
    TIC_CUDA(h, cudaMemcpyAsync(h->stage_in[b], (const char*)in + (size_t)u0 * in_unit_bytes, (size_t)nu * in_unit_bytes,
                                cudaMemcpyHostToDevice, h->s_h2d));
    TIC_CUDA(h, cudaEventRecord(h->ev_in[b], h->s_h2d));
    TIC_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_in[b], 0));
    if (idx >= 2 && !inout_same) TIC_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_out[b], 0));
    void* dout = inout_same ? h->stage_in[b] : h->stage_out[b];
    int rc = run(h->stage_in[b], dout, 0, nu, /*local=*/true);
    if (rc != TIC_OK) return rc;
    TIC_CUDA(h, cudaEventRecord(h->ev_comp[b], h->stream));
    TIC_CUDA(h, cudaStreamWaitEvent(h->s_d2h, h->ev_comp[b], 0));
    TIC_CUDA(h, cudaMemcpyAsync((char*)out + (size_t)u0 * out_unit_bytes, dout, (size_t)nu * out_unit_bytes,
                                cudaMemcpyDeviceToHost, h->s_d2h));
    TIC_CUDA(h, cudaEventRecord(h->ev_out[b], h->s_d2h));
  }
  TIC_CUDA(h, cudaEventRecord(h->ev_t1, h->stream));
  TIC_CUDA(h, cudaStreamSynchronize(h->s_d2h));
  TIC_CUDA(h, cudaStreamSynchronize(h->stream));
  return TIC_OK;
  }();
  if (rc_pipe != TIC_OK) {
    quiesce(h);
    return rc_pipe;
  }
  return overflow_status(h);
}

Geo patch_geo(int P, long long n0) {
  Geo g{};
  g.H = P;
  g.W = P;
  g.gh = 1;
  g.gw = 1;
  g.oy = 0;
  g.ox = 0;
  g.P = P;
  g.n0 = n0;
  geo_finish(g);
  return g;
}

Geo image_geo(int H, int W, int P, long long n0) {
  Geo g{};
  g.H = H;
  g.W = W;
  g.gh = (H + P - 1) / P;
  g.gw = (W + P - 1) / P;
  g.oy = 0;
  g.ox = 0;
  g.P = P;
  g.n0 = n0;
  geo_finish(g);
  return g;
}

}  // namespace

// =============================================================================================
extern "C" {

const char* tic_version(void) { return "tic-b200 0.1 (sm_100a)"; }
const char* tic_create_error(void) { return g_create_error.c_str(); }

int tic_create(tic_codec** out, int device) {
  if (!out) return TIC_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                     " (this library has no CPU fallback)";
    return TIC_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_create_error = "device index out of range";
    return TIC_ERR_INVALID;
  }
  tic_codec* h = new tic_codec();
  h->device = device;
  auto bail = [&](const char* what, cudaError_t err) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    delete h;
    return TIC_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  h->num_sms = prop.multiProcessorCount;
  if (prop.major != 10) {
    g_create_error = "this build targets sm_100a (B200); device is sm_" + std::to_string(prop.major * 10 + prop.minor);
    delete h;
    return TIC_ERR_UNSUPPORTED;
  }
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  h->own_stream = true;
  if ((e = cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  for (int b = 0; b < 2; ++b) {
    cudaEventCreateWithFlags(&h->ev_in[b], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_comp[b], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_out[b], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_enc[b], cudaEventDisableTiming);
  }
  cudaEventCreate(&h->ev_rt);
  cudaEventCreate(&h->ev_t0);
  cudaEventCreate(&h->ev_t1);
  if ((e = cudaMalloc(&h->d_hist, 256 * sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc", e);
  cudaMemset(h->d_hist, 0, 256 * sizeof(unsigned long long));
  if ((e = cudaMalloc(&h->d_symlut, 256 * sizeof(float))) != cudaSuccess) return bail("cudaMalloc", e);
  cudaMemset(h->d_symlut, 0, 256 * sizeof(float));
  if ((e = cudaHostAlloc((void**)&h->h_oflow, 2 * sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable)) != cudaSuccess)
    return bail("cudaHostAlloc", e);
  h->h_oflow[0] = h->h_oflow[1] = 0u;
  if ((e = cudaMalloc(&h->d_cum, (kEntropyMaxSymbols + 1) * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaHostGetDevicePointer((void**)&h->d_oflow, h->h_oflow, 0)) != cudaSuccess) return bail("cudaHostGetDevicePointer", e);
  cudaEventCreateWithFlags(&h->ev_switch, cudaEventDisableTiming);
  *out = h;
  return TIC_OK;
}

void tic_destroy(tic_codec* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (int gi = 0; gi < 3; ++gi) {
    for (auto& l : h->g[gi].layers) {
      if (l.w) cudaFree(l.w);
      if (l.b) cudaFree(l.b);
      l.uw.release();
      l.uw16.release();
      l.fw16.release();
      l.fdw.release();
      l.few.release();
    }
    if (h->g[gi].d_normlut) cudaFree(h->g[gi].d_normlut);
  }
  for (int i = 0; i < 3; ++i)
    if (h->act[i]) cudaFree(h->act[i]);
  for (int i = 0; i < 2; ++i)
    if (h->bnd[i]) cudaFree(h->bnd[i]);
  for (int b = 0; b < 2; ++b) {
    if (h->stage_in[b]) cudaFree(h->stage_in[b]);
    if (h->stage_out[b]) cudaFree(h->stage_out[b]);
    if (h->stage_sym[b]) cudaFree(h->stage_sym[b]);
    cudaEventDestroy(h->ev_enc[b]);
    cudaEventDestroy(h->ev_in[b]);
    cudaEventDestroy(h->ev_comp[b]);
    cudaEventDestroy(h->ev_out[b]);
  }
  cudaEventDestroy(h->ev_rt);
  cudaEventDestroy(h->ev_t0);
  cudaEventDestroy(h->ev_t1);
  if (h->d_hist) cudaFree(h->d_hist);
  if (h->d_symlut) cudaFree(h->d_symlut);
  if (h->h_oflow) cudaFreeHost(h->h_oflow);
  if (h->d_cum) cudaFree(h->d_cum);
  if (h->d_seg) cudaFree(h->d_seg);
  if (h->d_seg_bytes) cudaFree(h->d_seg_bytes);
  if (h->ev_switch) cudaEventDestroy(h->ev_switch);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  delete h;
}

const char* tic_last_error(const tic_codec* h) { return h ? h->err.c_str() : "null handle"; }

int tic_set_stream(tic_codec* h, void* cuda_stream) {
  if (!h) return TIC_ERR_INVALID;
  cudaStream_t next = (cudaStream_t)cuda_stream;
  if (next == h->stream) return TIC_OK;
  TIC_CUDA(h, cudaSetDevice(h->device));
  // The handle's workspaces (rotating activation buffers, weight images under construction, the histogram, the
  // staging sets) are shared by everything it launches: work queued on the old stream must finish before work on the
  // new stream may touch them.  Device-side ordering (event), no host stall — except when leaving the private
  // stream, which is destroyed.
  if (h->own_stream && h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  } else {
    TIC_CUDA(h, cudaEventRecord(h->ev_switch, h->stream));
    TIC_CUDA(h, cudaStreamWaitEvent(next, h->ev_switch, 0));
  }
  h->stream = next;
  h->own_stream = false;
  return TIC_OK;
}

int tic_set_compute_mode(tic_codec* h, int mode) {
  if (!h) return TIC_ERR_INVALID;
  if (mode < TIC_COMPUTE_FP32 || mode > TIC_COMPUTE_TENSOR_F16X3) return fail(h, TIC_ERR_INVALID, "unknown compute mode %d", mode);
  h->mode = mode;
  if (h->h_oflow) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    *h->h_oflow = 0u;  // the fp16-range flag belongs to the previous mode's runs
  }
  return TIC_OK;
}

int tic_set_chunk_patches(tic_codec* h, int chunk) {
  if (!h || chunk <= 0) return h ? fail(h, TIC_ERR_INVALID, "chunk must be positive") : TIC_ERR_INVALID;
  h->chunk128 = chunk;
  return TIC_OK;
}

int tic_set_graph(tic_codec* h, int graph, const tic_layer_desc* layers, int n_layers) {
  if (!h) return TIC_ERR_INVALID;
  if (graph < 0 || graph > 2 || !layers || n_layers <= 0) return fail(h, TIC_ERR_INVALID, "bad graph arguments");
  TIC_CUDA(h, cudaSetDevice(h->device));
  int prev = -1;
  bool open_res = false;
  for (int i = 0; i < n_layers; ++i) {
    const tic_layer_desc& d = layers[i];
    if (d.kind != TIC_CONV && d.kind != TIC_DECONV) return fail(h, TIC_ERR_INVALID, "layer %d: bad kind", i);
    if (d.cin <= 0 || d.cout <= 0 || d.cin > 1024 || d.cout > 1024) return fail(h, TIC_ERR_INVALID, "layer %d: bad channels", i);
    if (d.kind == TIC_CONV && d.stride != 1 && d.stride != 2) return fail(h, TIC_ERR_INVALID, "layer %d: conv stride must be 1 or 2", i);
    if (d.kind == TIC_DECONV && d.stride != 2) return fail(h, TIC_ERR_INVALID, "layer %d: deconv stride must be 2", i);
    if (prev >= 0 && d.cin != prev) return fail(h, TIC_ERR_INVALID, "layer %d: cin %d != previous cout %d", i, d.cin, prev);
    if (d.res_begin) {
      if (open_res || i == 0) return fail(h, TIC_ERR_INVALID, "layer %d: bad res_begin", i);
      open_res = true;
    }
    if (d.res_end) {
      if (!open_res || d.kind != TIC_CONV || d.stride != 1) return fail(h, TIC_ERR_INVALID, "layer %d: bad res_end", i);
      open_res = false;
    }
    if (open_res && (d.kind != TIC_CONV || d.stride != 1 || d.cin != d.cout))
      return fail(h, TIC_ERR_INVALID, "layer %d: residual blocks need stride-1 convs with cin == cout", i);
    if (i == n_layers - 1 && open_res) return fail(h, TIC_ERR_INVALID, "unterminated residual block");
    prev = d.cout;
  }
  Graph& g = h->g[graph];
  for (auto& l : g.layers) {
    if (l.w) cudaFree(l.w);
    if (l.b) cudaFree(l.b);
    l.uw.release();
    l.uw16.release();
    l.fw16.release();
    l.fdw.release();
    l.few.release();
  }
  g.layers.assign(n_layers, Layer());
  for (int i = 0; i < n_layers; ++i) {
    Layer& l = g.layers[i];
    l.d = layers[i];
    TIC_CUDA(h, cudaMalloc(&l.w, (size_t)9 * l.d.cin * l.d.cout * sizeof(float)));
    TIC_CUDA(h, cudaMalloc(&l.b, (size_t)l.d.cout * sizeof(float)));
  }
  return TIC_OK;
}

int tic_load_weights(tic_codec* h, int graph, int layer, const float* kernel, const float* bias) {
  if (!h) return TIC_ERR_INVALID;
  if (graph < 0 || graph > 2 || !kernel || !bias) return fail(h, TIC_ERR_INVALID, "bad arguments");
  Graph& g = h->g[graph];
  if (layer < 0 || layer >= (int)g.layers.size()) return fail(h, TIC_ERR_INVALID, "layer index %d out of range", layer);
  TIC_CUDA(h, cudaSetDevice(h->device));
  Layer& l = g.layers[layer];
  const int cin = l.d.cin, cout = l.d.cout;
  std::vector<float> w((size_t)9 * cin * cout);
  if (l.d.kind == TIC_CONV) {
    // TF HWIO [3][3][cin][cout] is already [tap][cin][cout]
    std::memcpy(w.data(), kernel, w.size() * sizeof(float));
  } else {
    // TF conv2d_transpose filter [3][3][cout][cin] (basic_block.py:53) -> [tap][cin][cout]
    for (int t = 0; t < 9; ++t)
      for (int oc = 0; oc < cout; ++oc)
        for (int ic = 0; ic < cin; ++ic) w[((size_t)t * cin + ic) * cout + oc] = kernel[((size_t)t * cout + oc) * cin + ic];
  }
  TIC_CUDA(h, cudaStreamSynchronize(h->stream));
  TIC_CUDA(h, cudaMemcpy(l.w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  TIC_CUDA(h, cudaMemcpy(l.b, bias, (size_t)cout * sizeof(float), cudaMemcpyHostToDevice));
  l.hb.assign(bias, bias + cout);
  l.uw.release();
  l.uw16.release();
  l.fw16.release();
  l.fdw.release();
  l.few.release();
  if (layer > 0) {  // the fused pair kernels keep this layer's operand image with the previous layer
    g.layers[layer - 1].fdw.release();
    g.layers[layer - 1].few.release();
  }
  l.loaded = true;
  return TIC_OK;
}

int tic_set_norm(tic_codec* h, int graph, const float* mean3, const float* std3) {
  if (!h) return TIC_ERR_INVALID;
  if (graph < 0 || graph > 2 || !mean3 || !std3) return fail(h, TIC_ERR_INVALID, "bad arguments");
  TIC_CUDA(h, cudaSetDevice(h->device));
  Graph& g = h->g[graph];
  std::vector<float> lut(3 * 256);
  for (int c = 0; c < 3; ++c) {
    if (!(std3[c] != 0.0f)) return fail(h, TIC_ERR_INVALID, "std[%d] is zero", c);
    g.mean[c] = mean3[c];
    g.stdv[c] = std3[c];
    for (int v = 0; v < 256; ++v) lut[c * 256 + v] = tic_normalize((float)v, mean3[c], std3[c]);
  }
  if (!g.d_normlut) TIC_CUDA(h, cudaMalloc(&g.d_normlut, lut.size() * sizeof(float)));
  TIC_CUDA(h, cudaStreamSynchronize(h->stream));
  TIC_CUDA(h, cudaMemcpy(g.d_normlut, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice));
  g.has_norm = true;
  return TIC_OK;
}

int tic_set_quantizer(tic_codec* h, int quan_scale, const float* inv_sigmoid_lut) {
  if (!h) return TIC_ERR_INVALID;
  if (quan_scale < 2 || quan_scale > 256 || !inv_sigmoid_lut)
    return fail(h, TIC_ERR_INVALID, "quan_scale must be in [2, 256] (symbols are stored as uint8)");
  TIC_CUDA(h, cudaSetDevice(h->device));
  TIC_CUDA(h, cudaStreamSynchronize(h->stream));
  TIC_CUDA(h, cudaMemcpy(h->d_symlut, inv_sigmoid_lut, (size_t)quan_scale * sizeof(float), cudaMemcpyHostToDevice));
  h->q = quan_scale;
  return TIC_OK;
}

int tic_bottleneck_shape(tic_codec* h, int P, int* hb, int* wb, int* cb) {
  if (!h) return TIC_ERR_INVALID;
  Graph& g = h->g[TIC_GRAPH_ENCODER];
  if (g.layers.empty()) return fail(h, TIC_ERR_STATE, "encoder graph not configured");
  Shape so;
  if (graph_out_shape(g, Shape{P, P, g.layers[0].d.cin}, &so, nullptr) != 0) return fail(h, TIC_ERR_INVALID, "inconsistent graph");
  if (hb) *hb = so.h;
  if (wb) *wb = so.w;
  if (cb) *cb = so.c;
  return TIC_OK;
}

// ---- hot path ---------------------------------------------------------------------------------
int tic_encode_patches(tic_codec* h, const void* patches, int in_dtype, int64_t n, int P, void* symbols,
                       int out_dtype, int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n < 0 || P <= 0 || (n > 0 && (!patches || !symbols))) return fail(h, TIC_ERR_INVALID, "bad arguments");
  if (in_dtype != TIC_U8 && in_dtype != TIC_F32) return fail(h, TIC_ERR_INVALID, "bad in_dtype");
  if (out_dtype != TIC_U8 && out_dtype != TIC_F32) return fail(h, TIC_ERR_INVALID, "bad out_dtype");
  int rc = check_graph_ready(h, TIC_GRAPH_ENCODER, true);
  if (rc != TIC_OK) return rc;
  if (h->q < 2) return fail(h, TIC_ERR_STATE, "quantiser not configured (tic_set_quantizer)");
  if (h->g[TIC_GRAPH_ENCODER].layers[0].d.cin != 3) return fail(h, TIC_ERR_INVALID, "encoder must take 3 channels");
  int hb, wb, cb;
  rc = tic_bottleneck_shape(h, P, &hb, &wb, &cb);
  if (rc != TIC_OK) return rc;
  const size_t in_unit = (size_t)P * P * 3 * (in_dtype == TIC_U8 ? 1 : 4);
  const size_t out_unit = (size_t)hb * wb * cb * (out_dtype == TIC_U8 ? 1 : 4);
  return drive(h, mem, patches, symbols, n, patches_per_chunk(h, P), in_unit, out_unit, false,
               [&](const void* din, void* dout, int64_t u0, int64_t nu, bool) {
                 IoSpec i, o;
                 i.mode = in_dtype == TIC_U8 ? IO_U8_NORM : IO_F32_NORM;
                 i.in = din;
                 i.geo = patch_geo(P, u0);
                 o.mode = out_dtype == TIC_U8 ? IO_QUANT_U8 : IO_QUANT_F32;
                 o.out = dout;
                 o.geo = patch_geo(P, u0);
                 return run_graph(h, TIC_GRAPH_ENCODER, i, o, (int)nu, P, P);
               }, (double)P * P / 16384.0);
}

int tic_encode_images(tic_codec* h, const uint8_t* images, int64_t n_images, int H, int W, int P, uint8_t* symbols,
                      int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n_images < 0 || H <= 0 || W <= 0 || P <= 0 || (n_images > 0 && (!images || !symbols)))
    return fail(h, TIC_ERR_INVALID, "bad arguments");
  if ((H % P != 0 && H < 2) || (W % P != 0 && W < 2)) return fail(h, TIC_ERR_INVALID, "reflect padding needs at least 2 pixels");
  int rc = check_graph_ready(h, TIC_GRAPH_ENCODER, true);
  if (rc != TIC_OK) return rc;
  if (h->q < 2) return fail(h, TIC_ERR_STATE, "quantiser not configured (tic_set_quantizer)");
  if (h->g[TIC_GRAPH_ENCODER].layers[0].d.cin != 3) return fail(h, TIC_ERR_INVALID, "encoder must take 3 channels");
  int hb, wb, cb;
  rc = tic_bottleneck_shape(h, P, &hb, &wb, &cb);
  if (rc != TIC_OK) return rc;
  const Geo g0 = image_geo(H, W, P, 0);
  const int ppi = g0.gh * g0.gw;
  const int64_t ipc = std::max<int64_t>(1, patches_per_chunk(h, P) / ppi);
  const size_t in_unit = (size_t)H * W * 3;
  const size_t out_unit = (size_t)ppi * hb * wb * cb;
  return drive(h, mem, images, symbols, n_images, ipc, in_unit, out_unit, false,
               [&](const void* din, void* dout, int64_t u0, int64_t nu, bool) {
                 IoSpec i, o;
                 i.mode = IO_U8_NORM;
                 i.in = din;
                 i.geo = image_geo(H, W, P, u0 * ppi);
                 o.mode = IO_QUANT_U8;
                 o.out = dout;
                 o.geo = i.geo;
                 return run_graph(h, TIC_GRAPH_ENCODER, i, o, (int)(nu * ppi), P, P);
               }, (double)ppi * P * P / 16384.0);
}

int tic_decode_patches(tic_codec* h, const uint8_t* symbols, int64_t n, int hb, int wb, float* recon, int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n < 0 || hb <= 0 || wb <= 0 || (n > 0 && (!symbols || !recon))) return fail(h, TIC_ERR_INVALID, "bad arguments");
  int rc = check_graph_ready(h, TIC_GRAPH_DECODER, true);
  if (rc != TIC_OK) return rc;
  if (h->q < 2) return fail(h, TIC_ERR_STATE, "quantiser not configured (tic_set_quantizer)");
  Graph& g = h->g[TIC_GRAPH_DECODER];
  const int cb = g.layers[0].d.cin;
  Shape so;
  if (graph_out_shape(g, Shape{hb, wb, cb}, &so, nullptr) != 0) return fail(h, TIC_ERR_INVALID, "inconsistent graph");
  if (so.c != 3 || so.h != so.w) return fail(h, TIC_ERR_INVALID, "decoder must produce square RGB patches (got %dx%dx%d)", so.h, so.w, so.c);
  const int P = so.h;
  return drive(h, mem, symbols, recon, n, patches_per_chunk(h, P), (size_t)hb * wb * cb, (size_t)P * P * 3 * 4, false,
               [&](const void* din, void* dout, int64_t u0, int64_t nu, bool) {
                 IoSpec i, o;
                 i.mode = IO_U8_SYMLUT;
                 i.in = din;
                 i.geo = patch_geo(P, u0);
                 o.mode = IO_DENORM_F32;
                 o.out = dout;
                 o.geo = patch_geo(P, u0);
                 return run_graph(h, TIC_GRAPH_DECODER, i, o, (int)nu, hb, wb);
               }, (double)P * P / 16384.0);
}

int tic_decode_images(tic_codec* h, const uint8_t* symbols, int64_t n_images, int H, int W, int P, void* images,
                      int out_dtype, int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n_images < 0 || H <= 0 || W <= 0 || P <= 0 || (n_images > 0 && (!symbols || !images)))
    return fail(h, TIC_ERR_INVALID, "bad arguments");
  if (out_dtype != TIC_U8 && out_dtype != TIC_F32) return fail(h, TIC_ERR_INVALID, "bad out_dtype");
  int rc = check_graph_ready(h, TIC_GRAPH_DECODER, true);
  if (rc != TIC_OK) return rc;
  if (h->q < 2) return fail(h, TIC_ERR_STATE, "quantiser not configured (tic_set_quantizer)");
  Graph& g = h->g[TIC_GRAPH_DECODER];
  const int cb = g.layers[0].d.cin;
  // bottleneck map size for patch size P: invert the decoder's upsampling
  int up = 1;
  for (auto& l : g.layers)
    if (l.d.kind == TIC_DECONV) up *= 2;
  if (P % up != 0) return fail(h, TIC_ERR_INVALID, "patch size %d is not a multiple of the decoder upsampling %d", P, up);
  const int hb = P / up, wb = P / up;
  Shape so;
  if (graph_out_shape(g, Shape{hb, wb, cb}, &so, nullptr) != 0 || so.c != 3 || so.h != P)
    return fail(h, TIC_ERR_INVALID, "decoder graph does not map %dx%d symbols to %dx%d RGB", hb, wb, P, P);
  const Geo g0 = image_geo(H, W, P, 0);
  const int ppi = g0.gh * g0.gw;
  const int64_t ipc = std::max<int64_t>(1, patches_per_chunk(h, P) / ppi);
  const size_t in_unit = (size_t)ppi * hb * wb * cb;
  const size_t out_unit = (size_t)H * W * 3 * (out_dtype == TIC_U8 ? 1 : 4);
  return drive(h, mem, symbols, images, n_images, ipc, in_unit, out_unit, false,
               [&](const void* din, void* dout, int64_t u0, int64_t nu, bool) {
                 IoSpec i, o;
                 i.mode = IO_U8_SYMLUT;
                 i.in = din;
                 i.geo = image_geo(H, W, P, u0 * ppi);
                 o.mode = out_dtype == TIC_U8 ? IO_DENORM_U8 : IO_DENORM_F32;
                 o.out = dout;
                 o.geo = i.geo;
                 return run_graph(h, TIC_GRAPH_DECODER, i, o, (int)(nu * ppi), hb, wb);
               }, (double)ppi * P * P / 16384.0);
}

int tic_roundtrip_images(tic_codec* h, const uint8_t* images, int64_t n_images, int H, int W, int P, uint8_t* symbols,
                         void* recon, int out_dtype, int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n_images < 0 || H <= 0 || W <= 0 || P <= 0 || (n_images > 0 && (!images || !recon)))
    return fail(h, TIC_ERR_INVALID, "bad arguments");
  if (out_dtype != TIC_U8 && out_dtype != TIC_F32) return fail(h, TIC_ERR_INVALID, "bad out_dtype");
  if (mem == TIC_MEM_DEVICE) {
    if (!symbols && n_images > 0) return fail(h, TIC_ERR_INVALID, "device-resident round trips need a symbol buffer");
    int rc = tic_encode_images(h, images, n_images, H, W, P, symbols, mem);
    if (rc != TIC_OK) return rc;
    cudaEvent_t t0 = h->ev_t0;  // keep the encoder's start event: last_kernel_ms covers both graphs
    h->ev_t0 = h->ev_rt;
    rc = tic_decode_images(h, symbols, n_images, H, W, P, recon, out_dtype, mem);
    h->ev_rt = h->ev_t0;
    h->ev_t0 = t0;
    return rc;
  }
  if ((H % P != 0 && H < 2) || (W % P != 0 && W < 2)) return fail(h, TIC_ERR_INVALID, "reflect padding needs at least 2 pixels");
  int rc = check_graph_ready(h, TIC_GRAPH_ENCODER, true);
  if (rc != TIC_OK) return rc;
  rc = check_graph_ready(h, TIC_GRAPH_DECODER, true);
  if (rc != TIC_OK) return rc;
  if (h->q < 2) return fail(h, TIC_ERR_STATE, "quantiser not configured (tic_set_quantizer)");
  if (h->g[TIC_GRAPH_ENCODER].layers[0].d.cin != 3) return fail(h, TIC_ERR_INVALID, "encoder must take 3 channels");
  int hb, wb, cb;
  rc = tic_bottleneck_shape(h, P, &hb, &wb, &cb);
  if (rc != TIC_OK) return rc;
  Graph& gd = h->g[TIC_GRAPH_DECODER];
  Shape so;
  if (gd.layers[0].d.cin != cb || graph_out_shape(gd, Shape{hb, wb, cb}, &so, nullptr) != 0 || so.c != 3 || so.h != P || so.w != P)
    return fail(h, TIC_ERR_INVALID, "decoder graph does not map the encoder's %dx%dx%d symbols back to %dx%d RGB", hb, wb, cb, P, P);
  if (n_images == 0) return TIC_OK;
  TIC_CUDA(h, cudaSetDevice(h->device));
  const Geo g0 = image_geo(H, W, P, 0);
  const int ppi = g0.gh * g0.gw;
  const std::vector<int64_t> sched =
      host_chunk_schedule(n_images, std::max<int64_t>(1, patches_per_chunk(h, P) / ppi), (double)ppi * P * P / 16384.0);
  const int64_t ipc = *std::max_element(sched.begin(), sched.end());
  const size_t img_unit = (size_t)H * W * 3;
  const size_t sym_unit = (size_t)ppi * hb * wb * cb;
  const size_t rec_unit = img_unit * (out_dtype == TIC_U8 ? 1 : 4);
  auto grow = [&](void** bufs, size_t* have, size_t need) -> int {
    if (*have >= need) return TIC_OK;
    TIC_CUDA(h, cudaDeviceSynchronize());
    for (int b = 0; b < 2; ++b) {
      if (bufs[b]) cudaFree(bufs[b]);
      bufs[b] = nullptr;
    }
    *have = 0;
    for (int b = 0; b < 2; ++b) TIC_CUDA(h, cudaMalloc(&bufs[b], need));
    *have = need;
    return TIC_OK;
  };
  if ((rc = grow(h->stage_in, &h->stage_in_bytes, (size_t)ipc * img_unit)) != TIC_OK) return rc;
  if ((rc = grow(h->stage_sym, &h->stage_sym_bytes, (size_t)ipc * sym_unit)) != TIC_OK) return rc;
  if ((rc = grow(h->stage_out, &h->stage_out_bytes, (size_t)ipc * rec_unit)) != TIC_OK) return rc;
  // Three streams, two staging sets: H2D of chunk i+1 (one PCIe direction) runs under the kernels of chunk i and
  // the D2H of chunk i-1 (the other direction).
  if (int st = overflow_status(h)) return st;
  const int rc_pipe = [&]() -> int {  // any failure drains the three streams before returning (see drive())
  TIC_CUDA(h, cudaEventRecord(h->ev_t0, h->stream));
  int64_t u0 = 0;
  for (int64_t idx = 0; idx < (int64_t)sched.size(); u0 += sched[idx], ++idx) {
    const int b = (int)(idx & 1);
    const int64_t nu = sched[idx];
    if (idx >= 2) TIC_CUDA(h, cudaStreamWaitEvent(h->s_h2d, h->ev_enc[b], 0));  // encoder of chunk idx-2 consumed stage_in[b]
    TIC_CUDA(h, cudaMemcpyAsync(h->stage_in[b], images + (size_t)u0 * img_unit, (size_t)nu * img_unit, cudaMemcpyHostToDevice,
                                h->s_h2d));
    TIC_CUDA(h, cudaEventRecord(h->ev_in[b], h->s_h2d));
    TIC_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_in[b], 0));
    if (idx >= 2) TIC_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_out[b], 0));  // stage_sym / stage_out[b] drained
    IoSpec i, o;
    i.mode = IO_U8_NORM;
    i.in = h->stage_in[b];
    i.geo = image_geo(H, W, P, 0);
    o.mode = IO_QUANT_U8;
    o.out = h->stage_sym[b];
    o.geo = i.geo;
    rc = run_graph(h, TIC_GRAPH_ENCODER, i, o, (int)(nu * ppi), P, P);
    if (rc != TIC_OK) return rc;
    TIC_CUDA(h, cudaEventRecord(h->ev_enc[b], h->stream));
    IoSpec di, dout;
    di.mode = IO_U8_SYMLUT;
    di.in = h->stage_sym[b];
    di.geo = i.geo;
    dout.mode = out_dtype == TIC_U8 ? IO_DENORM_U8 : IO_DENORM_F32;
    dout.out = h->stage_out[b];
    dout.geo = i.geo;
    rc = run_graph(h, TIC_GRAPH_DECODER, di, dout, (int)(nu * ppi), hb, wb);
    if (rc != TIC_OK) return rc;
    TIC_CUDA(h, cudaEventRecord(h->ev_comp[b], h->stream));
    if (symbols) {
      TIC_CUDA(h, cudaStreamWaitEvent(h->s_d2h, h->ev_enc[b], 0));
      TIC_CUDA(h, cudaMemcpyAsync(symbols + (size_t)u0 * sym_unit, h->stage_sym[b], (size_t)nu * sym_unit, cudaMemcpyDeviceToHost,
                                  h->s_d2h));
    }
    TIC_CUDA(h, cudaStreamWaitEvent(h->s_d2h, h->ev_comp[b], 0));
    TIC_CUDA(h, cudaMemcpyAsync((char*)recon + (size_t)u0 * rec_unit, h->stage_out[b], (size_t)nu * rec_unit, cudaMemcpyDeviceToHost,
                                h->s_d2h));
    TIC_CUDA(h, cudaEventRecord(h->ev_out[b], h->s_d2h));
  }
  TIC_CUDA(h, cudaEventRecord(h->ev_t1, h->stream));
  TIC_CUDA(h, cudaStreamSynchronize(h->s_d2h));
  TIC_CUDA(h, cudaStreamSynchronize(h->stream));
  return TIC_OK;
  }();
  if (rc_pipe != TIC_OK) {
    quiesce(h);
    return rc_pipe;
  }
  return overflow_status(h);
}

int tic_postfilter_patches(tic_codec* h, const float* tiles, int64_t n, int P, float* out, int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n < 0 || P <= 0 || (n > 0 && (!tiles || !out))) return fail(h, TIC_ERR_INVALID, "bad arguments");
  int rc = check_graph_ready(h, TIC_GRAPH_POSTFILTER, true);
  if (rc != TIC_OK) return rc;
  Graph& g = h->g[TIC_GRAPH_POSTFILTER];
  Shape so;
  if (g.layers[0].d.cin != 3 || graph_out_shape(g, Shape{P, P, 3}, &so, nullptr) != 0 || so.c != 3 || so.h != P || so.w != P)
    return fail(h, TIC_ERR_INVALID, "post-filter graph must map PxPx3 to PxPx3");
  const size_t unit = (size_t)P * P * 3 * 4;
  return drive(h, mem, tiles, out, n, patches_per_chunk(h, P), unit, unit, false,
               [&](const void* din, void* dout, int64_t u0, int64_t nu, bool) {
                 IoSpec i, o;
                 i.mode = IO_F32_NORM;
                 i.in = din;
                 i.geo = patch_geo(P, u0);
                 o.mode = IO_DENORM_F32;
                 o.out = dout;
                 o.geo = patch_geo(P, u0);
                 return run_graph(h, TIC_GRAPH_POSTFILTER, i, o, (int)nu, P, P);
               });
}

int tic_postfilter_images(tic_codec* h, float* images, int64_t n_images, int H, int W, int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n_images < 0 || H <= 0 || W <= 0 || (n_images > 0 && !images)) return fail(h, TIC_ERR_INVALID, "bad arguments");
  int rc = check_graph_ready(h, TIC_GRAPH_POSTFILTER, true);
  if (rc != TIC_OK) return rc;
  Graph& g = h->g[TIC_GRAPH_POSTFILTER];
  const int P = 128, off = 64;  // submit/2/rmbe/rmbe.py:12,16
  Shape so;
  if (g.layers[0].d.cin != 3 || graph_out_shape(g, Shape{P, P, 3}, &so, nullptr) != 0 || so.c != 3 || so.h != P || so.w != P)
    return fail(h, TIC_ERR_INVALID, "post-filter graph must map 128x128x3 to 128x128x3");
  const size_t unit = (size_t)H * W * 3 * 4;
  // tiles per image of the two passes; whole images per chunk so that pass 2 reads pass-1 output
  const int t1 = (H / P) * ((W - off) / P), t2 = ((H - off) / P) * (W / P);
  const int64_t ipc = std::max<int64_t>(1, patches_per_chunk(h, P) / std::max(1, std::max(t1, t2)));
  return drive(h, mem, images, images, n_images, ipc, unit, unit, true,
               [&](const void* din, void* dout, int64_t u0, int64_t nu, bool) {
                 (void)din;
                 for (int pass = 0; pass < 2; ++pass) {
                   Geo gg{};
                   gg.H = H;
                   gg.W = W;
                   gg.P = P;
                   if (pass == 0) {  // rmbe_height: rows i*128, cols 64 + j*128 (rmbe.py:70-89)
                     gg.gh = H / P;
                     gg.gw = W >= off ? (W - off) / P : 0;
                     gg.oy = 0;
                     gg.ox = off;
                   } else {          // rmbe_width: rows 64 + i*128, cols j*128 (rmbe.py:92-111)
                     gg.gh = H >= off ? (H - off) / P : 0;
                     gg.gw = W / P;
                     gg.oy = off;
                     gg.ox = 0;
                   }
                   const int tiles = gg.gh * gg.gw;
                   if (tiles <= 0) continue;
                   gg.n0 = u0 * tiles;
                   geo_finish(gg);
                   IoSpec i, o;
                   i.mode = IO_F32_NORM;
                   i.in = dout;  // in place: pass 2 reads what pass 1 wrote
                   i.geo = gg;
                   o.mode = IO_DENORM_F32;
                   o.out = dout;
                   o.geo = gg;
                   int r2 = run_graph(h, TIC_GRAPH_POSTFILTER, i, o, (int)(nu * tiles), P, P);
                   if (r2 != TIC_OK) return r2;
                 }
                 return (int)TIC_OK;
               });
}

int tic_run_layers(tic_codec* h, int graph, const float* in, int64_t n, int h0, int w0, int n_layers, float* out,
                   int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (graph < 0 || graph > 2 || n < 0 || h0 <= 0 || w0 <= 0 || n_layers <= 0 || (n > 0 && (!in || !out)))
    return fail(h, TIC_ERR_INVALID, "bad arguments");
  int rc = check_graph_ready(h, graph, false);
  if (rc != TIC_OK) return rc;
  Graph& g = h->g[graph];
  if (n_layers > (int)g.layers.size()) return fail(h, TIC_ERR_INVALID, "graph has only %zu layers", g.layers.size());
  if (g.layers[n_layers - 1].d.res_begin) return fail(h, TIC_ERR_INVALID, "cannot stop inside a residual block");
  const int cin = g.layers[0].d.cin;
  Shape so;
  if (graph_out_shape(g, Shape{h0, w0, cin}, &so, nullptr, n_layers) != 0) return fail(h, TIC_ERR_INVALID, "inconsistent graph");
  const size_t in_unit = (size_t)h0 * w0 * cin * 4, out_unit = (size_t)so.h * so.w * so.c * 4;
  const int64_t upc = std::max<int64_t>(1, (int64_t)h->chunk128 * 128 * 128 / ((int64_t)std::max(h0 * w0, so.h * so.w)));
  return drive(h, mem, in, out, n, upc, in_unit, out_unit, false,
               [&](const void* din, void* dout, int64_t u0, int64_t nu, bool) {
                 IoSpec i, o;
                 i.mode = IO_ACT;
                 i.in = (const char*)din + (size_t)u0 * in_unit;
                 o.mode = IO_ACT;
                 o.out = (char*)dout + (size_t)u0 * out_unit;
                 return run_graph(h, graph, i, o, (int)nu, h0, w0, n_layers);
               });
}

int tic_round_u8(tic_codec* h, const float* src, uint8_t* dst, int64_t count, int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (count < 0 || (count > 0 && (!src || !dst))) return fail(h, TIC_ERR_INVALID, "bad arguments");
  const int64_t upc = 1 << 26;
  return drive(h, mem, src, dst, count, upc, 4, 1, false, [&](const void* din, void* dout, int64_t u0, int64_t nu, bool) {
    const float* s = (const float*)din + u0;
    uint8_t* d = (uint8_t*)dout + u0;
    int blocks = (int)std::min<int64_t>((nu + 255) / 256, (int64_t)h->num_sms * 16);
    round_u8_kernel<<<blocks, 256, 0, h->stream>>>(s, d, nu);
    h->launches++;
    TIC_CUDA(h, cudaGetLastError());
    return (int)TIC_OK;
  });
}

// ---- statistics -------------------------------------------------------------------------------
int tic_hist_reset(tic_codec* h) {
  if (!h) return TIC_ERR_INVALID;
  TIC_CUDA(h, cudaSetDevice(h->device));
  TIC_CUDA(h, cudaMemsetAsync(h->d_hist, 0, 256 * sizeof(unsigned long long), h->stream));
  return TIC_OK;
}

int tic_hist_read(tic_codec* h, uint64_t* counts, int q) {
  if (!h) return TIC_ERR_INVALID;
  if (!counts || q <= 0 || q > 256) return fail(h, TIC_ERR_INVALID, "bad arguments");
  TIC_CUDA(h, cudaSetDevice(h->device));
  TIC_CUDA(h, cudaMemcpyAsync(counts, h->d_hist, (size_t)q * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
  TIC_CUDA(h, cudaStreamSynchronize(h->stream));
  return TIC_OK;
}

int tic_hist_device_ptr(tic_codec* h, void** dev_ptr) {
  if (!h || !dev_ptr) return TIC_ERR_INVALID;
  *dev_ptr = h->d_hist;
  return TIC_OK;
}

int tic_position_sums(tic_codec* h, const uint8_t* symbols, int64_t n, int64_t npos, uint64_t* sums, int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n < 0 || npos <= 0 || !sums || (n > 0 && !symbols)) return fail(h, TIC_ERR_INVALID, "bad arguments");
  TIC_CUDA(h, cudaSetDevice(h->device));
  const uint8_t* dsym = symbols;
  unsigned long long* dsum = (unsigned long long*)sums;
  void *tmp_sym = nullptr, *tmp_sum = nullptr;
  if (mem == TIC_MEM_HOST) {
    TIC_CUDA(h, cudaMalloc(&tmp_sym, (size_t)std::max<int64_t>(1, n * npos)));
    TIC_CUDA(h, cudaMalloc(&tmp_sum, (size_t)npos * 8));
    TIC_CUDA(h, cudaMemcpyAsync(tmp_sym, symbols, (size_t)(n * npos), cudaMemcpyHostToDevice, h->stream));
    TIC_CUDA(h, cudaMemcpyAsync(tmp_sum, sums, (size_t)npos * 8, cudaMemcpyHostToDevice, h->stream));
    dsym = (const uint8_t*)tmp_sym;
    dsum = (unsigned long long*)tmp_sum;
  }
  position_sums_kernel<<<(unsigned)((npos + 255) / 256), 256, 0, h->stream>>>(dsym, n, npos, dsum);
  h->launches++;
  TIC_CUDA(h, cudaGetLastError());
  if (mem == TIC_MEM_HOST) {
    TIC_CUDA(h, cudaMemcpyAsync(sums, tmp_sum, (size_t)npos * 8, cudaMemcpyDeviceToHost, h->stream));
    TIC_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(tmp_sym);
    cudaFree(tmp_sum);
  }
  return TIC_OK;
}

int tic_position_sums_batched(tic_codec* h, const uint8_t* symbols, int64_t n, int64_t npos, int64_t batch, uint64_t* sums,
                              int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n < 0 || npos <= 0 || batch <= 0 || !sums || (n > 0 && !symbols)) return fail(h, TIC_ERR_INVALID, "bad arguments");
  const int64_t nb = (n + batch - 1) / batch;
  if (nb == 0) return TIC_OK;
  if (nb > 65535) return fail(h, TIC_ERR_INVALID, "too many batches (%lld)", (long long)nb);
  TIC_CUDA(h, cudaSetDevice(h->device));
  const uint8_t* dsym = symbols;
  unsigned long long* dsum = (unsigned long long*)sums;
  void *tmp_sym = nullptr, *tmp_sum = nullptr;
  if (mem == TIC_MEM_HOST) {
    TIC_CUDA(h, cudaMalloc(&tmp_sym, (size_t)(n * npos)));
    TIC_CUDA(h, cudaMalloc(&tmp_sum, (size_t)(nb * npos) * 8));
    TIC_CUDA(h, cudaMemcpyAsync(tmp_sym, symbols, (size_t)(n * npos), cudaMemcpyHostToDevice, h->stream));
    dsym = (const uint8_t*)tmp_sym;
    dsum = (unsigned long long*)tmp_sum;
  }
  position_sums_batched_kernel<<<dim3((unsigned)((npos + 255) / 256), (unsigned)nb), 256, 0, h->stream>>>(dsym, n, npos, batch, dsum);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && mem == TIC_MEM_HOST)
    e = cudaMemcpyAsync(sums, tmp_sum, (size_t)(nb * npos) * 8, cudaMemcpyDeviceToHost, h->stream);
  if (mem == TIC_MEM_HOST) {
    cudaStreamSynchronize(h->stream);
    cudaFree(tmp_sym);
    cudaFree(tmp_sum);
  }
  if (e != cudaSuccess) return fail(h, TIC_ERR_CUDA, "position sums failed: %s", cudaGetErrorString(e));
  return TIC_OK;
}

static int prof_collect(tic_codec* h) {
  for (auto& r : h->prof_pending) {
    float ms = 0.f;
    cudaEventSynchronize(r.e1);
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    auto& v = h->prof_ms[r.graph];
    auto& c = h->prof_n[r.graph];
    if ((int)v.size() <= r.layer) {
      v.resize(r.layer + 1, 0.f);
      c.resize(r.layer + 1, 0);
    }
    v[r.layer] += ms;
    c[r.layer] += 1;
  }
  h->prof_pending.clear();
  return TIC_OK;
}

int tic_profile_enable(tic_codec* h, int on) {
  if (!h) return TIC_ERR_INVALID;
  h->profile = on != 0;
  return TIC_OK;
}

int tic_profile_reset(tic_codec* h) {
  if (!h) return TIC_ERR_INVALID;
  prof_collect(h);
  for (int g = 0; g < 3; ++g) {
    h->prof_ms[g].clear();
    h->prof_n[g].clear();
  }
  return TIC_OK;
}

int tic_profile_read(tic_codec* h, int graph, float* ms, int64_t* launches, int n_layers) {
  if (!h) return TIC_ERR_INVALID;
  if (graph < 0 || graph > 2 || !ms || !launches || n_layers <= 0) return fail(h, TIC_ERR_INVALID, "bad arguments");
  prof_collect(h);
  for (int i = 0; i < n_layers; ++i) {
    ms[i] = i < (int)h->prof_ms[graph].size() ? h->prof_ms[graph][i] : 0.f;
    launches[i] = i < (int)h->prof_n[graph].size() ? h->prof_n[graph][i] : 0;
  }
  return TIC_OK;
}

// ---- GPU entropy stage ------------------------------------------------------------------------
int64_t tic_entropy_bound(int64_t stream_len) { return (tic_rc_bound(stream_len < 0 ? 0 : stream_len) + 15) & ~(int64_t)15; }

static int entropy_table(tic_codec* h, const uint32_t* cum, int n_cum) {
  if (!cum || n_cum < 2 || n_cum - 1 > kEntropyMaxSymbols) return fail(h, TIC_ERR_INVALID, "entropy stage: tables of 1..256 symbols");
  if (cum[0] != 0) return fail(h, TIC_ERR_INVALID, "entropy stage: cumulative frequencies must start at 0");
  for (int i = 1; i < n_cum; ++i)
    if (cum[i] < cum[i - 1]) return fail(h, TIC_ERR_INVALID, "entropy stage: cumulative frequencies must not decrease");
  if (cum[n_cum - 1] == 0 || cum[n_cum - 1] > TIC_RC_MAX_TOTAL) return fail(h, TIC_ERR_INVALID, "entropy stage: frequency total must be in [1, 65536]");
  TIC_CUDA(h, cudaMemcpyAsync(h->d_cum, cum, (size_t)n_cum * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  return TIC_OK;
}

int tic_entropy_encode(tic_codec* h, const uint8_t* symbols, int64_t n_streams, int64_t stream_len, const uint32_t* cum_freq,
                       int n_cum, uint8_t* out, int64_t out_stride, int64_t* out_bytes, int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n_streams < 0 || stream_len < 0 || out_stride < 0 || (n_streams > 0 && (!symbols || !out || !out_bytes)))
    return fail(h, TIC_ERR_INVALID, "bad arguments");
  if (n_streams == 0) return TIC_OK;
  if (n_streams > 0x7fffffffLL) return fail(h, TIC_ERR_INVALID, "too many streams");
  TIC_CUDA(h, cudaSetDevice(h->device));
  if (int st = overflow_status(h)) return st;
  int rc = entropy_table(h, cum_freq, n_cum);
  if (rc != TIC_OK) return rc;
  const uint8_t* dsym = symbols;
  uint8_t* dout = out;
  long long* dbytes = reinterpret_cast<long long*>(out_bytes);
  void *t_sym = nullptr, *t_out = nullptr, *t_bytes = nullptr;
  if (mem == TIC_MEM_HOST) {
    TIC_CUDA(h, cudaMalloc(&t_sym, (size_t)std::max<int64_t>(16, n_streams * stream_len)));
    TIC_CUDA(h, cudaMalloc(&t_out, (size_t)std::max<int64_t>(16, n_streams * out_stride)));
    TIC_CUDA(h, cudaMalloc(&t_bytes, (size_t)n_streams * 8));
    TIC_CUDA(h, cudaMemcpyAsync(t_sym, symbols, (size_t)(n_streams * stream_len), cudaMemcpyHostToDevice, h->stream));
    dsym = (const uint8_t*)t_sym;
    dout = (uint8_t*)t_out;
    dbytes = (long long*)t_bytes;
  }
  const int64_t nseg = tic_rc_segments(stream_len);
  if (nseg > 0x7fffffffLL / std::max<int64_t>(1, n_streams)) return fail(h, TIC_ERR_INVALID, "too many segments");
  if (nseg > 0 && n_streams * nseg > h->seg_cap) {
    // grow-only scratch: the stream is drained first, earlier calls may still be packing out of the old buffer
    TIC_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->d_seg) cudaFree(h->d_seg);
    if (h->d_seg_bytes) cudaFree(h->d_seg_bytes);
    h->d_seg = nullptr;
    h->d_seg_bytes = nullptr;
    h->seg_cap = 0;
    TIC_CUDA(h, cudaMalloc(&h->d_seg, (size_t)(n_streams * nseg * kEntropySegSlot)));
    TIC_CUDA(h, cudaMalloc(&h->d_seg_bytes, (size_t)(n_streams * nseg) * sizeof(long long)));
    h->seg_cap = n_streams * nseg;
  }
  TIC_CUDA(h, cudaEventRecord(h->ev_t0, h->stream));
  if (nseg == 0) {
    rc_encode_kernel<<<(unsigned)n_streams, 32, 0, h->stream>>>(dsym, stream_len, stream_len, 1, h->d_cum, n_cum, dout, out_stride, dbytes,
                                                                  h->d_oflow + 1);
    h->launches++;
  } else {
    rc_encode_kernel<<<(unsigned)(n_streams * nseg), 32, 0, h->stream>>>(dsym, stream_len, TIC_RC_SEGMENT_SYMBOLS, (int)nseg, h->d_cum, n_cum,
                                                                           h->d_seg, kEntropySegSlot, h->d_seg_bytes, h->d_oflow + 1);
    rc_pack_kernel<<<(unsigned)(n_streams * nseg), 256, 0, h->stream>>>(h->d_seg, h->d_seg_bytes, (int)nseg, dout, out_stride, dbytes,
                                                                                      h->d_oflow + 1);
    h->launches += 2;
  }
  cudaError_t e = cudaGetLastError();
  cudaEventRecord(h->ev_t1, h->stream);
  if (mem == TIC_MEM_HOST) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_bytes, t_bytes, (size_t)n_streams * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    // only the stored bytes of every stream come back
    for (int64_t i = 0; e == cudaSuccess && i < n_streams; ++i)
      if (out_bytes[i] > 0)
        e = cudaMemcpyAsync(out + i * out_stride, (uint8_t*)t_out + i * out_stride, (size_t)out_bytes[i], cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(t_sym);
    cudaFree(t_out);
    cudaFree(t_bytes);
    if (e != cudaSuccess) return fail(h, TIC_ERR_CUDA, "entropy encode failed: %s", cudaGetErrorString(e));
    return overflow_status(h);
  }
  if (e != cudaSuccess) return fail(h, TIC_ERR_CUDA, "entropy encode launch failed: %s", cudaGetErrorString(e));
  return TIC_OK;
}

int tic_entropy_decode(tic_codec* h, const uint8_t* in, int64_t n_streams, int64_t in_stride, const int64_t* in_bytes,
                       const uint32_t* cum_freq, int n_cum, uint8_t* symbols, int64_t stream_len, int mem) {
  if (!h) return TIC_ERR_INVALID;
  if (n_streams < 0 || stream_len < 0 || in_stride < 0 || (n_streams > 0 && (!symbols || !in || !in_bytes)))
    return fail(h, TIC_ERR_INVALID, "bad arguments");
  if (n_streams == 0) return TIC_OK;
  if (n_streams > 0x7fffffffLL) return fail(h, TIC_ERR_INVALID, "too many streams");
  if (in_stride % 16 != 0) return fail(h, TIC_ERR_INVALID, "entropy decode: in_stride must be a multiple of 16 bytes");
  TIC_CUDA(h, cudaSetDevice(h->device));
  if (int st = overflow_status(h)) return st;
  int rc = entropy_table(h, cum_freq, n_cum);
  if (rc != TIC_OK) return rc;
  const uint8_t* din = in;
  const long long* dbytes = reinterpret_cast<const long long*>(in_bytes);
  uint8_t* dsym = symbols;
  void *t_in = nullptr, *t_bytes = nullptr, *t_sym = nullptr;
  if (mem == TIC_MEM_HOST) {
    for (int64_t i = 0; i < n_streams; ++i)
      if (in_bytes[i] < 0 || in_bytes[i] > in_stride)
        return fail(h, TIC_ERR_INVALID, "entropy decode: stream %lld has %lld bytes in a %lld-byte slot", (long long)i,
                    (long long)in_bytes[i], (long long)in_stride);
    TIC_CUDA(h, cudaMalloc(&t_in, (size_t)std::max<int64_t>(16, n_streams * in_stride)));
    TIC_CUDA(h, cudaMalloc(&t_bytes, (size_t)n_streams * 8));
    TIC_CUDA(h, cudaMalloc(&t_sym, (size_t)std::max<int64_t>(16, n_streams * stream_len)));
    for (int64_t i = 0; i < n_streams; ++i) {
      if (in_bytes[i] > 0)
        TIC_CUDA(h, cudaMemcpyAsync((uint8_t*)t_in + i * in_stride, in + i * in_stride, (size_t)in_bytes[i], cudaMemcpyHostToDevice, h->stream));
    }
    TIC_CUDA(h, cudaMemcpyAsync(t_bytes, in_bytes, (size_t)n_streams * 8, cudaMemcpyHostToDevice, h->stream));
    din = (const uint8_t*)t_in;
    dbytes = (const long long*)t_bytes;
    dsym = (uint8_t*)t_sym;
  } else if ((reinterpret_cast<uintptr_t>(in) & 15) != 0) {
    return fail(h, TIC_ERR_INVALID, "entropy decode: the stream buffer must be 16-byte aligned");
  }
  TIC_CUDA(h, cudaEventRecord(h->ev_t0, h->stream));
  const int64_t nseg = tic_rc_segments(stream_len);
  if (nseg > 0x7fffffffLL / n_streams) return fail(h, TIC_ERR_INVALID, "too many segments");
  rc_decode_kernel<<<(unsigned)(n_streams * std::max<int64_t>(1, nseg)), 32, 0, h->stream>>>(din, in_stride, dbytes, h->d_cum, n_cum, dsym,
                                                                                                stream_len, (int)nseg);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  cudaEventRecord(h->ev_t1, h->stream);
  if (mem == TIC_MEM_HOST) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(symbols, t_sym, (size_t)(n_streams * stream_len), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(t_in);
    cudaFree(t_bytes);
    cudaFree(t_sym);
  }
  if (e != cudaSuccess) return fail(h, TIC_ERR_CUDA, "entropy decode failed: %s", cudaGetErrorString(e));
  return TIC_OK;
}

#ifdef TIC_ABLATE
// ablation builds only (tools/fused_waits.py): the fused kernels' wait-time counters of cluster 0
extern "C" int tic_debug_prof_read(unsigned long long* out64, int reset) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (out64 && cudaMemcpyFromSymbol(out64, g_tic_prof, sizeof(unsigned long long) * 64) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[64] = {};
    if (cudaMemcpyToSymbol(g_tic_prof, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}
#endif

int tic_check_status(tic_codec* h) {
  if (!h) return TIC_ERR_INVALID;
  TIC_CUDA(h, cudaSetDevice(h->device));
  TIC_CUDA(h, cudaStreamSynchronize(h->stream));
  return overflow_status(h);
}

uint32_t tic_crc32c(const void* data, uint64_t n) {
  static uint32_t table[8][256];
  static const bool ready = [] {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
      table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xffu];
    return true;
  }();
  (void)ready;
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = 0xffffffffu;
  while (n >= 8) {  // slicing-by-8
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = table[7][lo & 0xffu] ^ table[6][(lo >> 8) & 0xffu] ^ table[5][(lo >> 16) & 0xffu] ^ table[4][lo >> 24] ^
        table[3][hi & 0xffu] ^ table[2][(hi >> 8) & 0xffu] ^ table[1][(hi >> 16) & 0xffu] ^ table[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ table[0][(c ^ *p++) & 0xffu];
  return c ^ 0xffffffffu;
}

int64_t tic_launch_count(const tic_codec* h) { return h ? h->launches : 0; }

float tic_last_kernel_ms(const tic_codec* h) {
  if (!h) return 0.f;
  float ms = 0.f;
  if (cudaEventSynchronize(h->ev_t1) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1) != cudaSuccess) return -1.f;
  return ms;
}

}  // extern "C"
