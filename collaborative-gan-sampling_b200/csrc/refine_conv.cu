// Image-path refinement: layer -> gathered-GEMM parameter mapping, the fused head kernel (logit GEMV +
// BCE gradient + best-of-K select + image keep), and the K-step loop of sampling/collaborator.py:41-88.
#include "common.h"
#include "conv_gemm.cuh"
#include "edge_conv.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace cgs {

namespace {

inline int cstride(int c) { return (c + 3) & ~3; }   // NHWC channel stride (floats)

inline int same_pad_before(int in_size, int k) {      // TF SAME, stride 2
  const int out = (in_size + 1) / 2;
  int total = (out - 1) * 2 + k - in_size;
  if (total < 0) total = 0;
  return total / 2;
}

struct LayerShape {
  int hout, wout, cs_in, cs_out;
};

LayerShape layer_shape(const cgs_layer_desc& L) {
  LayerShape s;
  s.cs_in = cstride(L.cin);
  s.cs_out = cstride(L.cout);
  if (L.type == CGS_LAYER_CONV) {
    s.hout = (L.hin + 1) / 2;
    s.wout = (L.win + 1) / 2;
  } else if (L.type == CGS_LAYER_DECONV) {
    s.hout = L.hin * 2;
    s.wout = L.win * 2;
  } else {
    s.hout = s.wout = 1;
  }
  return s;
}

int check_layer(const cgs_layer_desc& L) {
  if (L.type == CGS_LAYER_FC) {
    if (L.cin % 32) return set_error(CGS_ERR_UNSUPPORTED, "fc input size %d must be a multiple of 32", L.cin);
    return CGS_OK;
  }
  if (L.k < 2 || L.k > 5) return set_error(CGS_ERR_UNSUPPORTED, "kernel size %d unsupported (2..5)", L.k);
  if ((L.hin & 1) && L.type == CGS_LAYER_CONV)
    return set_error(CGS_ERR_UNSUPPORTED, "odd conv input size %d unsupported", L.hin);
  const bool in_ok = (L.cin % 32 == 0) || L.cin <= 4;
  const bool out_ok = (L.cout % 32 == 0) || L.cout <= 4;
  if (!in_ok || !out_ok)
    return set_error(CGS_ERR_UNSUPPORTED, "channels (%d -> %d) must be <= 4 or multiples of 32", L.cin, L.cout);
  // the transposed-type pass gathers 32-channel blocks of its input
  if (L.type == CGS_LAYER_DECONV && L.cin % 32) return set_error(CGS_ERR_UNSUPPORTED, "deconv cin %d", L.cin);
  if (L.type == CGS_LAYER_CONV && L.cout % 32) return set_error(CGS_ERR_UNSUPPORTED, "conv cout %d", L.cout);
  return CGS_OK;
}

// Strided-type pass ("F"): out pixel (j,i) reads in pixels (2j+ky-p, 2i+kx-p); K order (ky,kx,c).
void fill_strided(ConvGemmParams& p, int k, int pad_y, int pad_x, int cin_k) {
  p.nclasses = 1;
  GemmClass& g = p.cls[0];
  g.k0 = 0;
  g.oy0 = g.ox0 = 0;
  g.ntaps = k * k;
  g.nkx = k;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      g.dy[ky * k + kx] = (signed char)(ky - pad_y);
      g.dx[ky * k + kx] = (signed char)(kx - pad_x);
    }
  if (cin_k % 32 == 0) {
    p.cblocks = cin_k / 32;
    g.nkb = g.ntaps * p.cblocks;
  } else {
    p.cblocks = 0;                          // pixel mode: 8 taps x 4 channels per K block
    g.nkb = (g.ntaps + 7) / 8;
  }
  p.S = 2;
  p.os = 1;
}

// Transposed-type pass ("T"): output pixel (2j+py, 2i+px) reads in pixels (j+dy, i+dx),
// dy = (py + p - ky)/2 over the ky with matching parity; K order (class, tap, c).  Classes are emitted
// heaviest first so the persistent tile loop tails off on the short ones.
void fill_transposed(ConvGemmParams& p, int k, int pad_y, int pad_x, int cin_k) {
  p.nclasses = 4;
  p.cblocks = cin_k / 32;
  p.S = 1;
  p.os = 2;
  int k0 = 0;
  int ci = 0;
  // order parities by tap count (descending)
  int py_order[2] = {0, 1}, px_order[2] = {0, 1};
  auto ntap = [&](int par, int pad) { int n = 0; for (int kk = 0; kk < k; ++kk) if (((par + pad - kk) & 1) == 0) ++n; return n; };
  if (ntap(1, pad_y) > ntap(0, pad_y)) { py_order[0] = 1; py_order[1] = 0; }
  if (ntap(1, pad_x) > ntap(0, pad_x)) { px_order[0] = 1; px_order[1] = 0; }
  // CANONICAL SHIFT ORDER shared by all classes: a class's taps are enumerated in the order of their input shifts
  // (dy, dx), shifts used by more classes first (ties: dy descending, dx descending).  The class-fused tcgen05 tiles
  // (conv_gemm.cu) walk the shifts in this same order and may merge the classes of a shift into one wide MMA, so every
  // class accumulates its taps in the same order fused or not -- results never depend on the tile choice.
  auto has_tap = [&](int par, int pad, int d) { const int kk = par + pad - 2 * d; return kk >= 0 && kk < k; };
  struct Shift { int dy, dx, users; };
  Shift sh[64];
  int nsh = 0;
  for (int dy = 2; dy >= -2; --dy)
    for (int dx = 2; dx >= -2; --dx) {
      int users = 0;
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
          if (has_tap(py, pad_y, dy) && has_tap(px, pad_x, dx)) ++users;
      if (users) sh[nsh++] = Shift{dy, dx, users};
    }
  for (int a = 1; a < nsh; ++a)                       // stable insertion sort by users, descending
    for (int b = a; b > 0 && sh[b].users > sh[b - 1].users; --b) { const Shift t = sh[b]; sh[b] = sh[b - 1]; sh[b - 1] = t; }
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const int py = py_order[a], px = px_order[b];
      GemmClass& g = p.cls[ci++];
      g.k0 = k0;
      g.oy0 = py;
      g.ox0 = px;
      int t = 0;
      for (int si = 0; si < nsh; ++si) {
        if (!has_tap(py, pad_y, sh[si].dy) || !has_tap(px, pad_x, sh[si].dx)) continue;
        g.dy[t] = (signed char)sh[si].dy;
        g.dx[t] = (signed char)sh[si].dx;
        ++t;
      }
      g.ntaps = t;
      g.nkx = ntap(px, pad_x);
      g.nkb = t * p.cblocks;
      k0 += t * cin_k;
    }
}

}  // namespace

// Parameters of the forward pass of one layer.
int make_forward_params(const cgs_layer_desc& L, int64_t B, const float* x, float* y, ConvGemmParams& p) {
  if (int rc = check_layer(L)) return rc;
  std::memset(&p, 0, sizeof(p));
  const LayerShape s = layer_shape(L);
  p.in = x;
  p.out = y;
  p.bias = L.bias;
  p.epi = EPI_FWD;
  p.act = L.act;
  p.N = L.cout;
  p.ON = s.cs_out;
  p.OH = s.hout;
  p.OW = s.wout;
  p.IH = L.hin;
  p.IW = L.win;
  p.Cs = s.cs_in;
  if (L.type == CGS_LAYER_CONV) {
    fill_strided(p, L.k, same_pad_before(L.hin, L.k), same_pad_before(L.win, L.k), s.cs_in);
    p.MH = s.hout;
    p.MW = s.wout;
  } else if (L.type == CGS_LAYER_DECONV) {
    fill_transposed(p, L.k, same_pad_before(s.hout, L.k), same_pad_before(s.wout, L.k), L.cin);
    p.MH = L.hin;
    p.MW = L.win;
  } else {
    p.nclasses = 1;
    p.cls[0].ntaps = 1;
    p.cls[0].nkx = 1;
    p.cblocks = L.cin / 32;
    p.cls[0].nkb = p.cblocks;
    p.S = 1;
    p.os = 1;
    p.MH = p.MW = 1;
    p.IH = p.IW = 1;
    p.Cs = L.cin;
  }
  p.M = (int)(B * p.MH * p.MW);
  return CGS_OK;
}

// Parameters of the data-gradient pass of one layer: dy (grad w.r.t. the layer's pre-activation output) -> dx.
int make_backward_params(const cgs_layer_desc& L, int64_t B, const float* dy, float* dx, ConvGemmParams& p) {
  if (int rc = check_layer(L)) return rc;
  std::memset(&p, 0, sizeof(p));
  const LayerShape s = layer_shape(L);
  p.in = dy;
  p.out = dx;
  p.epi = EPI_RAW;
  p.N = L.cin;
  p.ON = s.cs_in;
  p.OH = L.hin;
  p.OW = L.win;
  p.IH = s.hout;
  p.IW = s.wout;
  p.Cs = s.cs_out;
  if (L.type == CGS_LAYER_CONV) {
    // conv data-gradient is a transposed-type pass over dy
    fill_transposed(p, L.k, same_pad_before(L.hin, L.k), same_pad_before(L.win, L.k), L.cout);
    p.MH = s.hout;
    p.MW = s.wout;
  } else if (L.type == CGS_LAYER_DECONV) {
    // deconv data-gradient is a strided-type pass over dy (== conv fprop with the same filter)
    fill_strided(p, L.k, same_pad_before(s.hout, L.k), same_pad_before(s.wout, L.k), s.cs_out);
    p.MH = L.hin;
    p.MW = L.win;
  } else {
    if (L.cout % 32) return set_error(CGS_ERR_UNSUPPORTED, "fc output size %d must be a multiple of 32", L.cout);
    p.nclasses = 1;
    p.cls[0].ntaps = 1;
    p.cls[0].nkx = 1;
    p.cblocks = L.cout / 32;
    p.cls[0].nkb = p.cblocks;
    p.S = 1;
    p.os = 1;
    p.MH = p.MW = 1;
    p.IH = p.IW = 1;
    p.OH = p.OW = 1;
    p.Cs = L.cout;
    p.ON = L.cin;
  }
  p.M = (int)(B * p.MH * p.MW);
  return CGS_OK;
}


// ---------------------------------------------------------------------------------------------
// Image-like tensors (<= 4 channels: the generated image and its gradient) live in a PITCHED layout inside the
// chain: [B][H][W + 8][4] with pixel x stored at column x + 2 and zero columns around it.  That lets the strided
// passes over them ("window" lowering below) fetch a whole K block with one TMA box and makes horizontal padding
// free; vertical padding is TMA out-of-bounds zero fill.  The public API stays dense: [B][H][W][4].
// ---------------------------------------------------------------------------------------------
constexpr int IMG_XOFF = 2;
inline int img_pitch(int w) { return w + 8; }
// layout of an image-like tensor handed to / produced by a pass
enum : int { IMG_PITCHED = 0,   // [B][H][W + 8][4] (inside the chain, mma.sync edge kernels / general lowerings)
             IMG_DENSE = 1,     // [B][H][W][4]     (single-layer entry points)
             IMG_S2D = 2 };     // [B][H/2][W/2][16] space-to-depth (inside the chain, tcgen05 edge kernels: edge_tc.cu)
inline size_t tensor_elems(int h, int w, int c) {        // per-sample elements of an activation inside the chain
  return c <= 4 ? (size_t)h * img_pitch(w) * 4 : (size_t)h * w * cstride(c);
}

// Window lowering: strided-type passes whose INPUT is image-like (first conv forward, last deconv data-gradient).
bool use_window(const cgs_layer_desc& L, bool backward) {
  return (L.type == CGS_LAYER_CONV && !backward && L.cin <= 4) || (L.type == CGS_LAYER_DECONV && backward && L.cout <= 4);
}
inline int window_kcols(const cgs_layer_desc& L) { return L.k * 32; }

// in = pitched image-like tensor; K block ky = 8 stored pixels x 4 channels starting at (2j + ky - pad_y, 2i + x0)
int make_window_params(const cgs_layer_desc& L, bool backward, int64_t B, const float* in, float* out,
                       ConvGemmParams& p) {
  if (int rc = check_layer(L)) return rc;
  if (L.k > 8) return set_error(CGS_ERR_UNSUPPORTED, "window lowering needs k <= 8");
  std::memset(&p, 0, sizeof(p));
  const LayerShape s = layer_shape(L);
  const int ih = backward ? s.hout : L.hin, iw = backward ? s.wout : L.win;      // image-like input
  const int oh = backward ? L.hin : s.hout, ow = backward ? L.win : s.wout;      // output grid = M-space
  const int pad_y = same_pad_before(ih, L.k), pad_x = same_pad_before(iw, L.k);
  if (pad_x > IMG_XOFF) return set_error(CGS_ERR_UNSUPPORTED, "horizontal padding %d exceeds the image margin", pad_x);
  p.in = in;
  p.out = out;
  p.bias = backward ? nullptr : L.bias;
  p.epi = EPI_RAW;
  p.N = backward ? L.cin : L.cout;
  p.ON = cstride(p.N);
  p.OH = oh; p.OW = ow;
  p.IH = ih; p.IW = iw;
  p.Cs = 4;
  p.cblocks = 1;
  p.MH = oh; p.MW = ow;
  p.S = 2; p.os = 1;
  p.nclasses = 1;
  GemmClass& g = p.cls[0];
  g.k0 = 0; g.oy0 = g.ox0 = 0;
  g.ntaps = L.k; g.nkx = 1; g.nkb = L.k;
  for (int ky = 0; ky < L.k; ++ky) { g.dy[ky] = (signed char)(ky - pad_y); g.dx[ky] = 0; }
  p.window = 1;
  p.win_k = L.k;
  p.win_x0 = IMG_XOFF - pad_x;
  p.in_pitch_px = img_pitch(iw);
  p.M = (int)(B * p.MH * p.MW);
  return CGS_OK;
}

// dense [B][H][W][4]  <->  pitched [B][H][W+8][4] (test entry points and API copies only)
__global__ void __launch_bounds__(256) pad_image_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int H,
                                                        int W, long long rows) {
  const int P = W + 8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < rows * P;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % P);
    const long long r = idx / P;
    const int x = c - IMG_XOFF;
    dst[idx] = (x >= 0 && x < W) ? src[r * W + x] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
int pad_image(const float* dense, float* pitched, int64_t B, int H, int W, cudaStream_t st) {
  const long long rows = (long long)B * H;
  long long blocks = (rows * (W + 8) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pad_image_kernel<<<(int)blocks, 256, 0, st>>>((const float4*)dense, (float4*)pitched, H, W, rows); count_launch();
  return check_launch("pad_image_kernel");
}
int unpad_image(const float* pitched, float* dense, int64_t B, int H, int W, cudaStream_t st) {
  cudaError_t e = cudaMemcpy2DAsync(dense, (size_t)W * 16, pitched + IMG_XOFF * 4, (size_t)img_pitch(W) * 16, (size_t)W * 16,
                                    (size_t)B * H, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return set_error(CGS_ERR_CUDA, "cudaMemcpy2DAsync: %s", cudaGetErrorString(e));
  return CGS_OK;
}

// ---------------------------------------------------------------------------------------------
// Scatter formulation for transposed-type passes with <= 4 output channels (deconv -> image, conv1 data-gradient).
// The gather form would re-read every input pixel k^2/4 * 4 times for a 1..3-wide output; instead
//   stage 1  col[b,iy,ix,(ky,kx,c)] = sum_ci in[b,iy,ix,ci] * W[(ky,kx,c)][ci]     one GEMM, input read once
//   stage 2  out[b,y,x,c] = epi( sum_{(ky,kx): y = 2 iy + ky - p, x = 2 ix + kx - p} col[b,iy,ix,(ky,kx,c)] )
// Same MACs as the true-tap count; stage 2 is a small memory-bound kernel with the fused epilogue.
// ---------------------------------------------------------------------------------------------
bool use_scatter(const cgs_layer_desc& L, bool backward) {
  return (L.type == CGS_LAYER_DECONV && !backward && L.cout <= 4) || (L.type == CGS_LAYER_CONV && backward && L.cin <= 4);
}
inline int scatter_cols(const cgs_layer_desc& L) { return L.k * L.k * 4; }      // col row pitch (floats)

// Split-K for the forward pass of a long-K fc layer (d_fc3: 6272 -> 1024): at batch 1024 there are only 8 x 4 tiles of
// 128 x 256, and a CTA pays one pipeline hand-shake per K block whatever the tile width, so narrow tiles over the full
// K (128 tiles of 128 x 64, 98 K blocks each) run at a third of the wide-tile rate.  Instead: 128 x 256 tiles over a
// quarter of K each (the split is a "class" of the gathered GEMM: own K range, own output row), FP32 partial sums in
// the pass scratch buffer, and a small kernel that adds the four partials in a fixed order and applies the epilogue.
constexpr int kFcSplit = 4;
inline int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
inline int fc_split(const cgs_layer_desc& L, bool backward) {
  static const int S = env_int("CGS_FC_SPLIT", kFcSplit);      // developer knob
  return (L.type == CGS_LAYER_FC && !backward && L.cin >= 4096 && L.cin % (32 * S) == 0) ? S : 1;
}

inline bool use_fc_split(const cgs_layer_desc& L, bool backward, int64_t B, int math, const float* scratch) {
  return fc_split(L, backward) > 1 && math == CGS_MATH_TF32_TENSOR && scratch && !(debug_flags() & 16384) &&
         ((B + 127) / 128) * ((L.cout + 255) / 256) * fc_split(L, backward) <= 2 * 148;
}

size_t scatter_col_elems(const cgs_layer_desc& L, bool backward) {     // per-sample elements of the pass scratch buffer
  if (fc_split(L, backward) > 1) return (size_t)fc_split(L, backward) * cstride(L.cout);
  if (!use_scatter(L, backward)) return 0;
  const LayerShape s = layer_shape(L);
  const size_t pixels = backward ? (size_t)s.hout * s.wout : (size_t)L.hin * L.win;   // stage-1 rows per sample
  return pixels * scatter_cols(L);
}

// stage 1 as a 1-tap "fc over pixels": rows = B * pixels, K = big channel count, N = k*k*4
int make_scatter_gemm_params(const cgs_layer_desc& L, bool backward, int64_t B, const float* in, float* col,
                             ConvGemmParams& p) {
  if (int rc = check_layer(L)) return rc;
  std::memset(&p, 0, sizeof(p));
  const LayerShape s = layer_shape(L);
  const int kch = backward ? L.cout : L.cin;                   // reduced (big) channel count
  const int64_t pixels = backward ? (int64_t)s.hout * s.wout : (int64_t)L.hin * L.win;
  p.in = in;
  p.out = col;
  p.epi = EPI_RAW;
  p.N = scatter_cols(L);
  p.ON = scatter_cols(L);
  p.OH = p.OW = 1;
  p.IH = p.IW = 1;
  p.Cs = kch;
  p.nclasses = 1;
  p.cls[0].ntaps = 1;
  p.cls[0].nkx = 1;
  p.cblocks = kch / 32;
  p.cls[0].nkb = p.cblocks;
  p.S = 1;
  p.os = 1;
  p.MH = p.MW = 1;
  if (B * pixels >= (1ll << 31) / p.ON) return set_error(CGS_ERR_UNSUPPORTED, "batch too large for 32-bit indexing; split the batch");
  p.M = (int)(B * pixels);
  return CGS_OK;
}

struct Col2imParams {
  const float* col;
  float* out;
  const float* bias;
  const float* aux;
  int IH, IW, OH, OW, k, pad_y, pad_x, pitch;
  int epi, act, round_out;
  int out_pitch, out_xoff;   // row pitch (pixels) and first data column of out / aux (dense: OW, 0)
  long long pixels;          // B * OH * out_pitch (pad columns are written as zeros)
  const int* live;           // early exit: images still in the batch (null = all)
};

__global__ void __launch_bounds__(256) col2im_kernel(const Col2imParams p) {
  long long pixels = p.pixels;
  if (p.live) {
    const long long per_img = (long long)p.OH * p.out_pitch;
    pixels = (long long)live_images(p.live, (int)(p.pixels / per_img)) * per_img;
  }
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < pixels;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xc = (int)(idx % p.out_pitch);
    const long long t = idx / p.out_pitch;
    const int y = (int)(t % p.OH);
    const long long b = t / p.OH;
    const int x = xc - p.out_xoff;
    if (x < 0 || x >= p.OW) {                       // margin of the pitched layout
      *reinterpret_cast<float4*>(p.out + idx * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ky = 0; ky < p.k; ++ky) {
      const int ty = y + p.pad_y - ky;
      if (ty < 0 || (ty & 1)) continue;
      const int iy = ty >> 1;
      if (iy >= p.IH) continue;
      for (int kx = 0; kx < p.k; ++kx) {
        const int tx = x + p.pad_x - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int ix = tx >> 1;
        if (ix >= p.IW) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(
            p.col + ((b * p.IH + iy) * p.IW + ix) * p.pitch + (ky * p.k + kx) * 4));
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    float4 o = a;
    if (p.epi == EPI_FWD) {
      const float4 bb = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias)) : make_float4(0.f, 0.f, 0.f, 0.f);
      o.x = act_apply(a.x + bb.x, p.act); o.y = act_apply(a.y + bb.y, p.act);
      o.z = act_apply(a.z + bb.z, p.act); o.w = act_apply(a.w + bb.w, p.act);
    } else if (p.epi == EPI_BWD) {
      const float4 yv = __ldg(reinterpret_cast<const float4*>(p.aux + idx * 4));
      o.x = a.x * act_grad_from_output(yv.x, p.act); o.y = a.y * act_grad_from_output(yv.y, p.act);
      o.z = a.z * act_grad_from_output(yv.z, p.act); o.w = a.w * act_grad_from_output(yv.w, p.act);
    }
    if (p.round_out) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
    *reinterpret_cast<float4*>(p.out + idx * 4) = o;
  }
}

struct PassEpi {
  int epi = EPI_RAW;
  int act = ACT_NONE;
  const float* aux = nullptr;
  int round_out = 0;
  const ConvGemmParams* upd = nullptr;   // EPI_UPDATE fields
  const int* live = nullptr;             // early exit: device-resident count of the images still in the batch
};

// Epilogue description of an image-edge kernel (edge_conv.cuh) from the pass epilogue.
EdgeEpi make_edge_epi(const PassEpi& e, const float* bias) {
  EdgeEpi x;
  std::memset(&x, 0, sizeof(x));
  x.epi = e.epi;
  x.act_tanh = (e.act == ACT_TANH);
  x.slope = e.act == ACT_RELU ? 0.f : (e.act == ACT_LRELU ? 0.2f : 1.f);
  x.round_out = e.round_out;
  x.bias = (e.epi == EPI_FWD) ? bias : nullptr;
  x.aux = e.aux;
  x.live = e.live;
  if (e.upd) {
    x.epi = EPI_UPDATE;
    x.mom = e.upd->mom; x.first = e.upd->first; x.sgd = e.upd->sgd; x.rate = e.upd->rate; x.alpha = e.upd->alpha;
    x.clip = e.upd->clip; x.vmin = e.upd->vmin; x.vmax = e.upd->vmax;
    x.round_out = 0;
  }
  return x;
}

// CGS_DEBUG bit 4096 (or cgs_debug_set_flags): keep the image-edge passes on the general tcgen05 lowerings
bool edge_kernels_enabled() { return !(debug_flags() & 4096); }

int launch_gemm(const ConvGemmParams& p, const float* w, int rows, int cols, int math, cudaStream_t stream) {
  if (math == CGS_MATH_FP32_SIMT) return launch_conv_gemm_simt(p, w, rows, cols, stream);
  if (math == CGS_MATH_TF32_TENSOR) return launch_conv_gemm_tc(p, w, rows, cols, stream);
  return set_error(CGS_ERR_INVALID, "unknown math mode %d", math);
}

// out[b][n] = epi( part[b][0][n] + part[b][1][n] + ... ) in a fixed order; same epilogue code as the GEMM kernels
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const __grid_constant__ ConvGemmParams p,
                                                            const float* __restrict__ part, int S, long long total4) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = p.ON / 4;
  if (p.live) total4 = (long long)live_images(p.live, (int)(total4 / groups)) * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / groups;
    const int ng = (int)(idx - b * groups);
    const int off = (int)(b * p.ON) + ng * 4;
    float4 a = __ldg(reinterpret_cast<const float4*>(part + (size_t)b * S * p.ON + ng * 4));
    for (int s = 1; s < S; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(part + ((size_t)b * S + s) * p.ON + ng * 4));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
    if (p.epi == EPI_FWD) { if (p.bias) x0 = __ldg(reinterpret_cast<const float4*>(p.bias + ng * 4)); }
    else if (p.epi == EPI_BWD) x0 = __ldg(reinterpret_cast<const float4*>(p.aux + off));
    else if (p.epi == EPI_UPDATE) {
      x0 = *reinterpret_cast<const float4*>(p.out + off);
      if (!p.sgd && !p.first) x1 = *reinterpret_cast<const float4*>(p.mom + off);
    }
    *reinterpret_cast<float4*>(p.out + off) = epilogue4(p, off, a, x0, x1);
  }
}

int launch_splitk_reduce(ConvGemmParams pe, const float* part, int S, int64_t B, cudaStream_t st) {
  pe.act_tanh = (pe.act == ACT_TANH);
  pe.slope = pe.act == ACT_RELU ? 0.f : (pe.act == ACT_LRELU ? 0.2f : 1.f);
  const long long total4 = (long long)B * (pe.ON / 4);
  long long blocks = (total4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaError_t le = launch_pdl(splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, st, pe, part, S, total4);
  count_launch();
  if (le != cudaSuccess) return set_error(CGS_ERR_CUDA, "splitk_reduce_kernel: %s", cudaGetErrorString(le));
  return check_launch("splitk_reduce_kernel");
}

// Parameters of the fused image-edge kernels (edge_conv.cuh) for a scatter- / window-lowered pass; false when the
// shape is not covered (the general tcgen05 lowering runs instead) or the kernels are switched off.
bool make_edge_narrow(const cgs_layer_desc& L, bool backward, int64_t B, const float* in, float* out, const PassEpi& e,
                      int img_mode, int w_cols, EdgeNarrowParams& q) {
  const LayerShape s = layer_shape(L);
  const int kch = backward ? L.cout : L.cin, cimg = backward ? L.cin : L.cout;
  const int iw = backward ? s.wout : L.win;
  if (!edge_kernels_enabled() || !use_scatter(L, backward) || !edge_narrow_supported(kch, L.k, cimg, iw) || w_cols != kch)
    return false;
  std::memset(&q, 0, sizeof(q));
  q.in = in; q.out = out; q.w = backward ? L.w_bwd : L.w_fwd;
  q.B = (int)B; q.K = kch; q.k = L.k; q.cimg = cimg;
  if (!backward) {
    q.IH = L.hin; q.IW = L.win; q.OH = s.hout; q.OW = s.wout;
    q.pad_y = same_pad_before(s.hout, L.k); q.pad_x = same_pad_before(s.wout, L.k);
  } else {
    q.IH = s.hout; q.IW = s.wout; q.OH = L.hin; q.OW = L.win;
    q.pad_y = same_pad_before(L.hin, L.k); q.pad_x = same_pad_before(L.win, L.k);
  }
  q.out_pitch = img_mode == IMG_PITCHED ? img_pitch(q.OW) : q.OW;
  q.out_xoff = img_mode == IMG_PITCHED ? IMG_XOFF : 0;
  q.s2d = img_mode == IMG_S2D;
  q.e = make_edge_epi(e, L.bias);
  return true;
}

bool make_edge_wide(const cgs_layer_desc& L, bool backward, int64_t B, const float* in, float* out, const PassEpi& e,
                    const ConvGemmParams& wp, EdgeWideParams& q) {     // wp = make_window_params of the same pass
  const int cimg = backward ? L.cout : L.cin;
  if (!edge_kernels_enabled() || !edge_wide_supported(wp.N, L.k, cimg) || wp.ON != wp.N) return false;
  std::memset(&q, 0, sizeof(q));
  q.in = in; q.out = out; q.w = backward ? L.w_bwd : L.w_fwd;
  q.IH = wp.IH; q.pitch = wp.in_pitch_px; q.xoff = IMG_XOFF; q.OH = wp.OH; q.OW = wp.OW; q.ON = wp.ON; q.N = wp.N;
  q.k = L.k; q.cimg = cimg;
  q.pad_y = same_pad_before(wp.IH, L.k); q.pad_x = same_pad_before(wp.IW, L.k);
  q.M = (long long)B * wp.OH * wp.OW;
  q.e = make_edge_epi(e, backward ? nullptr : L.bias);
  return true;
}

// The narrow pass of layer `ln` followed by the wide pass of layer `lw` on its output (forward: G's last deconv + D's
// first conv; backward: their data-gradients in the opposite order) as ONE kernel when both fit a one-image tile.
// Returns 1 when the pair was launched, 0 when the caller should run the two passes separately, < 0 on error.
int try_edge_pair(const cgs_layer_desc& ln, const cgs_layer_desc& lw, bool backward, int64_t B, const float* in,
                  float* mid, float* out, const PassEpi& en, const PassEpi& ew, int store_mid, int math, cudaStream_t st) {
  if (math != CGS_MATH_TF32_TENSOR || (debug_flags() & 65536) || en.upd) return 0;
  if (!use_scatter(ln, backward) || !use_window(lw, backward)) return 0;
  EdgeNarrowParams qn;
  if (!make_edge_narrow(ln, backward, B, in, mid, en, false, backward ? ln.kcols_bwd : ln.kcols_fwd, qn)) return 0;
  ConvGemmParams wp;
  if (make_window_params(lw, backward, B, mid, out, wp) != CGS_OK) return 0;
  if ((backward ? lw.rows_bwd : lw.rows_fwd) != (backward ? lw.cin : lw.cout) ||
      (backward ? lw.kcols_bwd : lw.kcols_fwd) != window_kcols(lw)) return 0;
  EdgeWideParams qw;
  if (!make_edge_wide(lw, backward, B, mid, out, ew, wp, qw)) return 0;
  if (!edge_pair_supported(qn, qw)) return 0;
  if (int rc = launch_edge_pair(qn, qw, store_mid, st)) return rc;
  return 1;
}

// One layer pass (forward or data-gradient) with its fused epilogue; picks the gather or the scatter lowering.
int run_pass(const cgs_layer_desc& L, bool backward, int64_t B, const float* in, float* out, const PassEpi& e,
             float* col, int math, cudaStream_t st, int img_mode = IMG_PITCHED, bool defer_reduce = false) {
  const float* w = backward ? L.w_bwd : L.w_fwd;
  const int rows = backward ? L.rows_bwd : L.rows_fwd;
  const int cols = backward ? L.kcols_bwd : L.kcols_fwd;
  if (!w) return set_error(CGS_ERR_INVALID, "layer has no packed %s weights", backward ? "backward" : "forward");
  if (use_scatter(L, backward)) {
    if (!col) return set_error(CGS_ERR_WORKSPACE, "scatter pass needs a column workspace");
    if (rows != scatter_cols(L)) return set_error(CGS_ERR_INVALID, "weights of this pass must be in scatter layout (%d rows)", scatter_cols(L));
    {
      EdgeNarrowParams q;
      if (math == CGS_MATH_TF32_TENSOR && !e.upd && make_edge_narrow(L, backward, B, in, out, e, img_mode, cols, q))
        return launch_edge_narrow(q, st);
      if (img_mode == IMG_S2D) return set_error(CGS_ERR_UNSUPPORTED, "s2d image layout needs the tcgen05 edge kernels");
    }
    ConvGemmParams p;
    if (int rc = make_scatter_gemm_params(L, backward, B, in, col, p)) return rc;
    // (early exit: the 1-tap GEMM over pixels runs over the full batch -- its rows are pixels, not images; the
    //  col2im below stops at the live images)
    if (int rc = launch_gemm(p, w, rows, cols, math, st)) return rc;
    const LayerShape s = layer_shape(L);
    Col2imParams c;
    c.col = col;
    c.out = out;
    c.bias = (e.epi == EPI_FWD) ? L.bias : nullptr;
    c.aux = e.aux;
    c.k = L.k;
    c.pitch = scatter_cols(L);
    c.epi = e.epi;
    c.act = e.act;
    c.round_out = e.round_out;
    if (!backward) {           // deconv forward: col over the input grid, out = 2x grid
      c.IH = L.hin; c.IW = L.win; c.OH = s.hout; c.OW = s.wout;
      c.pad_y = same_pad_before(s.hout, L.k); c.pad_x = same_pad_before(s.wout, L.k);
    } else {                   // conv data-gradient: col over the conv-output grid, out = conv-input grid
      c.IH = s.hout; c.IW = s.wout; c.OH = L.hin; c.OW = L.win;
      c.pad_y = same_pad_before(L.hin, L.k); c.pad_x = same_pad_before(L.win, L.k);
    }
    c.out_pitch = img_mode == IMG_PITCHED ? img_pitch(c.OW) : c.OW;
    c.out_xoff = img_mode == IMG_PITCHED ? IMG_XOFF : 0;
    c.pixels = (long long)B * c.OH * c.out_pitch;
    c.live = e.live;
    long long blocks = (c.pixels + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    col2im_kernel<<<(int)blocks, 256, 0, st>>>(c); count_launch();
    return check_launch("col2im_kernel");
  }
  ConvGemmParams p;
  if (const int S = fc_split(L, backward); use_fc_split(L, backward, B, math, col)) {
    if (int rc = make_forward_params(L, B, in, col, p)) return rc;
    ConvGemmParams pe = p;                    // epilogue description of the real output
    pe.out = out;
    pe.bias = L.bias;
    pe.epi = e.epi; pe.act = e.act; pe.aux = e.aux; pe.round_out = e.round_out;
    if (e.upd) {
      pe.epi = EPI_UPDATE;
      pe.mom = e.upd->mom; pe.first = e.upd->first; pe.sgd = e.upd->sgd; pe.rate = e.upd->rate; pe.alpha = e.upd->alpha;
      pe.clip = e.upd->clip; pe.vmin = e.upd->vmin; pe.vmax = e.upd->vmax;
      pe.round_out = 0;
    }
    p.bias = nullptr;
    p.epi = EPI_RAW;
    p.live = e.live;
    pe.live = e.live;
    p.nclasses = S;
    p.cblocks = L.cin / 32 / S;
    for (int c = 0; c < S; ++c) {
      GemmClass& g = p.cls[c];
      std::memset(&g, 0, sizeof(g));
      g.k0 = c * p.cblocks * 32;
      g.nkb = p.cblocks;
      g.ntaps = 1;
      g.nkx = 1;
      g.oy0 = c;                              // partial sums of split c land in row (b * S + c) of the scratch buffer
      g.cb0 = c * p.cblocks;
    }
    p.OH = S;
    { static const int bn = env_int("CGS_FC_BN", 256); p.force_bn = bn; }
    if (int rc = launch_gemm(p, w, rows, cols, math, st)) return rc;
    if (defer_reduce) return CGS_OK;          // the head kernel adds the partial sums itself (head_takes_partials)
    return launch_splitk_reduce(pe, col, S, B, st);
  }
  if (use_window(L, backward)) {
    if (rows != (backward ? L.cin : L.cout) || cols != window_kcols(L))
      return set_error(CGS_ERR_INVALID, "weights of this pass must be in window layout (%d columns)", window_kcols(L));
    if (int rc = make_window_params(L, backward, B, in, out, p)) return rc;
    EdgeWideParams q;
    if (img_mode == IMG_S2D) {               // image-like input in s2d layout: tcgen05 resident-patch kernel
      if (math == CGS_MATH_TF32_TENSOR && make_edge_wide(L, backward, B, in, out, e, p, q) && edge_wide_tc_supported(q))
        return launch_edge_wide_tc(q, (int)B, st);
      return set_error(CGS_ERR_UNSUPPORTED, "s2d image layout needs the tcgen05 edge kernels");
    }
    if (math == CGS_MATH_TF32_TENSOR && make_edge_wide(L, backward, B, in, out, e, p, q)) return launch_edge_wide(q, st);
  } else if (int rc = backward ? make_backward_params(L, B, in, out, p) : make_forward_params(L, B, in, out, p)) {
    return rc;
  }
  if (!backward) p.bias = L.bias;
  p.epi = e.epi;
  p.act = e.act;
  p.aux = e.aux;
  p.round_out = e.round_out;
  p.live = e.live;
  if (e.upd) {
    p.epi = EPI_UPDATE;
    p.mom = e.upd->mom; p.first = e.upd->first; p.sgd = e.upd->sgd; p.rate = e.upd->rate; p.alpha = e.upd->alpha;
    p.clip = e.upd->clip; p.vmin = e.upd->vmin; p.vmax = e.upd->vmax;
    p.round_out = 0;
  }
  return launch_gemm(p, w, rows, cols, math, st);
}

// ---------------------------------------------------------------------------------------------
// Head kernel: final linear (-> 1 logit), BCE-with-ones gradient, selection, image keep.
//   logit_b = <feat_b, w> + bias                                   (nsgan/ops.py:81, GAN.py:68)
//   d pre_k = (sigmoid(logit_b) - 1) * w_k * act'(feat_bk)         (GAN.py:176-177, collaborator.py:30-31)
//   deterministic: update iff logit > best (strict)                (collaborator.py:79-83)
//   probabilistic: update iff prob_indices[b] == step_index        (collaborator.py:77)
// One CTA per sample; the dot product is reduced in a fixed order that depends only on K (never on B).
// ---------------------------------------------------------------------------------------------
struct HeadParams {
  const float* feat;    // [B, K]  output of the last hidden layer (post activation)
  const float* w;       // [K]
  const float* bias;    // [1] device scalar (folded)
  int K;
  int act;              // activation of the producer of feat
  float* dpre;          // [B, K] or nullptr (no backward wanted)
  const float* img;     // [B, img_elems] current image
  float* best_img;      // [B, img_elems]
  const float* feature; // [B, feat_elems] current feature (only if best_feature wanted)
  float* best_feature;
  int img_elems, feat_elems;   // dense per-sample sizes (what the caller sees)
  int img_h, img_w4, img_pitch4, img_xoff4;   // chain-side image layout in float4 units (pitched if <= 4 channels)
  int img_s2d;          // chain-side image is in the space-to-depth layout [H/2][W/2][16] (tcgen05 edge kernels)
  float* cur_logit;     // [B]
  float* best_logit;    // [B]
  float* best_step;     // [B]
  float* default_logit; // [B] or nullptr
  const int* prob_indices;
  int step;             // -1 = initial evaluation (collaborator.py:49-60), else loop index i
  int mode;
  unsigned char* done;  // early-exit flags [B] or nullptr
  const int* orig;      // compact row -> original sample index (early-exit compaction) or nullptr (identity)
  float* final_feature; // [B_orig, feat_elems]: state of a sample at the step it exits (early-exit) or nullptr
  float exit_logit;
  int round_out;        // round dpre to TF32 (RN): it feeds a kind::tf32 MMA
  // split-K producer (fc_split): feat is not materialised; feat[b][k] = act(sum_s part[b][s][k] + fc_bias[k])
  const float* part;    // [B][S][K] partial sums or nullptr
  int S;
  const float* fc_bias;
  float fc_slope;       // relu 0, lrelu 0.2, none 1 (same expression as the GEMM epilogue: bit-identical)
  int fc_tanh;
  const int* live;      // early exit: device-resident count of the rows still in the batch (null = gridDim.x)
};

// Input of the head for elements k..k+3 of sample b: the stored activation, or the split-K partial sums added in
// a fixed order with the fc layer's bias and activation applied (what splitk_reduce_kernel would have stored).
__device__ __forceinline__ float4 head_feat4(const HeadParams& p, const float* f, int b, int k) {
  if (!p.part) return *reinterpret_cast<const float4*>(f + k);
  const float* pr = p.part + (size_t)b * p.S * p.K + k;
  float4 a = __ldg(reinterpret_cast<const float4*>(pr));
  for (int s = 1; s < p.S; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(pr + (size_t)s * p.K));
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  const float4 bb = p.fc_bias ? __ldg(reinterpret_cast<const float4*>(p.fc_bias + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float vx = a.x + bb.x, vy = a.y + bb.y, vz = a.z + bb.z, vw = a.w + bb.w;
  if (p.fc_tanh) return make_float4(tanhf(vx), tanhf(vy), tanhf(vz), tanhf(vw));
  return make_float4(fmaxf(vx, vx * p.fc_slope), fmaxf(vy, vy * p.fc_slope), fmaxf(vz, vz * p.fc_slope),
                     fmaxf(vw, vw * p.fc_slope));
}

__global__ void __launch_bounds__(256) head_kernel(const HeadParams p) {
  __shared__ float red[8];
  __shared__ float s_logit;
  __shared__ int s_update;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  if (p.live && b >= *reinterpret_cast<const volatile int*>(p.live)) return;   // row left the batch (early exit)
  const int ob = p.orig ? p.orig[b] : b;           // where this sample's results live
  const float* f = p.feat + (size_t)b * p.K;
  float acc = 0.f;
  float4 a_first = make_float4(0.f, 0.f, 0.f, 0.f);   // this thread's first four inputs, reused by the gradient pass
  for (int k = threadIdx.x * 4; k < p.K; k += 256 * 4) {
    const float4 a = head_feat4(p, f, b, k);
    if (k < 256 * 4) a_first = a;
    const float4 ww = __ldg(reinterpret_cast<const float4*>(p.w + k));
    acc = fmaf(a.x, ww.x, acc);
    acc = fmaf(a.y, ww.y, acc);
    acc = fmaf(a.z, ww.z, acc);
    acc = fmaf(a.w, ww.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    const float logit = t + (p.bias ? __ldg(p.bias) : 0.f);
    s_logit = logit;
    p.cur_logit[b] = logit;
    int upd;
    if (p.step < 0) {
      upd = 1;
      p.best_logit[ob] = logit;
      p.best_step[ob] = 1.0f;                                  // collaborator.py:60 (sic)
      if (p.default_logit) p.default_logit[ob] = logit;        // collaborator.py:52
    } else {
      const bool frozen = p.done && p.done[b];
      if (p.mode == CGS_MODE_PROBABILISTIC) upd = (p.prob_indices[ob] == p.step);
      else upd = logit > p.best_logit[ob];
      if (frozen) upd = 0;
      if (upd) {
        p.best_logit[ob] = logit;
        p.best_step[ob] = (float)(p.step + 1);
      }
    }
    int exit_now = 0;
    if (p.done && !p.done[b] && logit >= p.exit_logit) {       // opt-in early exit (README.md:13)
      p.done[b] = 1;
      exit_now = 1;
    }
    s_update = upd | (exit_now << 1);
  }
  __syncthreads();
  const float logit = s_logit;
  if (p.dpre) {
    // d softplus(-l)/dl = sigmoid(l) - 1
    const float dl = 1.f / (1.f + expf(-logit)) - 1.f;
    float* d = p.dpre + (size_t)b * p.K;
    for (int k = threadIdx.x * 4; k < p.K; k += 256 * 4) {
      const float4 a = (k < 256 * 4) ? a_first : head_feat4(p, f, b, k);
      const float4 ww = __ldg(reinterpret_cast<const float4*>(p.w + k));
      float4 o;
      o.x = dl * ww.x * act_grad_from_output(a.x, p.act);
      o.y = dl * ww.y * act_grad_from_output(a.y, p.act);
      o.z = dl * ww.z * act_grad_from_output(a.z, p.act);
      o.w = dl * ww.w * act_grad_from_output(a.w, p.act);
      if (p.round_out) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
      *reinterpret_cast<float4*>(d + k) = o;
    }
  }
  if (s_update & 1) {
    // chain-side image rows (possibly pitched) -> dense best_img
    float4* dst = reinterpret_cast<float4*>(p.best_img + (size_t)ob * p.img_elems);
    if (p.img_s2d) {
      // s2d pixel (y/2, x/2), sub-pixel (y & 1, x & 1) -> dense pixel (y, x); img_w4 = W here (one float4 per pixel)
      const float4* src = reinterpret_cast<const float4*>(p.img) + (size_t)b * p.img_h * p.img_w4;
      const int W = p.img_w4;
      for (int i = threadIdx.x; i < p.img_h * W; i += 256) {
        const int y = i / W, x = i - y * W;
        dst[i] = src[((size_t)(y >> 1) * (W >> 1) + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))];
      }
    } else {
      const float4* src = reinterpret_cast<const float4*>(p.img) + (size_t)b * p.img_h * p.img_pitch4 + p.img_xoff4;
      for (int i = threadIdx.x; i < p.img_h * p.img_w4; i += 256) {
        const int y = i / p.img_w4;
        dst[i] = src[(size_t)y * p.img_pitch4 + (i - y * p.img_w4)];
      }
    }
    if (p.best_feature) {
      const float4* fs = reinterpret_cast<const float4*>(p.feature + (size_t)b * p.feat_elems);
      float4* fd = reinterpret_cast<float4*>(p.best_feature + (size_t)ob * p.feat_elems);
      for (int i = threadIdx.x; i < p.feat_elems / 4; i += 256) fd[i] = fs[i];
    }
  }
  if ((s_update & 2) && p.final_feature) {                    // the sample leaves the batch: keep its final state
    const float4* fs = reinterpret_cast<const float4*>(p.feature + (size_t)b * p.feat_elems);
    float4* fd = reinterpret_cast<float4*>(p.final_feature + (size_t)ob * p.feat_elems);
    for (int i = threadIdx.x; i < p.feat_elems / 4; i += 256) fd[i] = fs[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Early exit (README.md:13; opt-in, the reference itself always runs K steps: collaborator.py:63-83), device side.
// After every fused policy step the rows D already classifies as real leave the batch: ONE single-CTA kernel turns the
// per-row exit flags into the ordered list of the rows that stay and their count (ballot + warp-sum scan), ONE kernel
// gathers feature / momentum / row-map rows through that list.  The count lives in device memory and every later
// launch of the step reads it (ConvGemmParams::live and friends), so there is no host synchronisation, the launch
// sequence is static and the whole K loop can be captured in a CUDA graph.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) live_compact_kernel(unsigned char* __restrict__ done, const int* __restrict__ live_in,
                                                            int B, int* __restrict__ idx, int* __restrict__ live_out) {
  __shared__ int warp_sums[32];
  __shared__ int s_base;
  const int live = min(max(*live_in, 0), B);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < B; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const bool keep = i < live && !done[i];
    if (i < B) done[i] = 0;                               // rows are renumbered: the flags start afresh
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_sums[warp] = __popc(m);
    __syncthreads();
    int before = s_base;
    for (int w2 = 0; w2 < warp; ++w2) before += warp_sums[w2];
    if (keep) idx[before + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = s_base;
      for (int w2 = 0; w2 < 32; ++w2) t += warp_sums[w2];
      s_base = t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *live_out = s_base;
}

// rows idx[0 .. *count) of (feature, momentum, row map) -> rows 0 .. *count of the partner buffers (ordered)
__global__ void __launch_bounds__(256) live_gather_kernel(const float4* __restrict__ feat_src, float4* __restrict__ feat_dst,
                                                          const float4* __restrict__ mom_src, float4* __restrict__ mom_dst,
                                                          const int* __restrict__ orig_src, int* __restrict__ orig_dst,
                                                          const int* __restrict__ idx, const int* __restrict__ count,
                                                          int elems4) {
  const int rows = *count;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int s = idx[r];
    const float4* fs = feat_src + (size_t)s * elems4;
    float4* fd = feat_dst + (size_t)r * elems4;
    for (int i = threadIdx.x; i < elems4; i += blockDim.x) fd[i] = fs[i];
    if (mom_src) {
      const float4* ms = mom_src + (size_t)s * elems4;
      float4* md = mom_dst + (size_t)r * elems4;
      for (int i = threadIdx.x; i < elems4; i += blockDim.x) md[i] = ms[i];
    }
    if (threadIdx.x == 0) orig_dst[r] = orig_src[s];
  }
}
// feature rows of the samples that never exited -> their slot in the caller's buffer
__global__ void scatter_active_kernel(const float4* __restrict__ feat, const int* __restrict__ orig,
                                      const unsigned char* __restrict__ done, const int* __restrict__ live,
                                      float4* __restrict__ out, int elems4) {
  const int b = blockIdx.x;
  if (b >= *live || done[b]) return;
  const float4* s = feat + (size_t)b * elems4;
  float4* d = out + (size_t)orig[b] * elems4;
  for (int i = threadIdx.x; i < elems4; i += blockDim.x) d[i] = s[i];
}
__global__ void iota_kernel(int* p, int n, int* live) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i;
  if (live && blockIdx.x == 0 && threadIdx.x == 0) *live = n;
}

// ---------------------------------------------------------------------------------------------
// Chain executor
// ---------------------------------------------------------------------------------------------
struct Chain {
  int n;                         // GEMM layers (G-tail + D without the 1-logit head)
  cgs_layer_desc layers[2 * CGS_MAX_LAYERS];
  size_t act_elems[2 * CGS_MAX_LAYERS + 1];   // per-sample elements of act[i] (act[0] = feature)
  int n_gtail;
  cgs_layer_desc head;
  size_t max_elems;
  size_t col_elems;              // per-sample elements of the scatter column buffer (0 if unused)
};

static int build_chain(const cgs_net_desc* gtail, const cgs_net_desc* d, Chain& c) {
  if (!gtail || !d) return set_error(CGS_ERR_INVALID, "null network descriptor");
  if (gtail->n_layers < 1 || gtail->n_layers > CGS_MAX_LAYERS || d->n_layers < 2 || d->n_layers > CGS_MAX_LAYERS)
    return set_error(CGS_ERR_INVALID, "bad layer counts (%d, %d)", gtail->n_layers, d->n_layers);
  c.n = 0;
  for (int i = 0; i < gtail->n_layers; ++i) c.layers[c.n++] = gtail->layers[i];
  c.n_gtail = gtail->n_layers;
  for (int i = 0; i < d->n_layers - 1; ++i) c.layers[c.n++] = d->layers[i];
  c.head = d->layers[d->n_layers - 1];
  if (c.head.type != CGS_LAYER_FC || c.head.cout != 1)
    return set_error(CGS_ERR_UNSUPPORTED, "discriminator must end in a linear layer with one logit");
  const cgs_layer_desc& L0 = c.layers[0];
  if (L0.cin <= 4) return set_error(CGS_ERR_UNSUPPORTED, "the refined feature map must have >= 32 channels");
  c.act_elems[0] = (size_t)L0.hin * L0.win * cstride(L0.cin);
  c.max_elems = c.act_elems[0];
  c.col_elems = 0;
  for (int i = 0; i < c.n; ++i) {
    if (int rc = check_layer(c.layers[i])) return rc;
    for (int bw = 0; bw < 2; ++bw) {
      const size_t ce = scatter_col_elems(c.layers[i], bw != 0);
      if (ce > c.col_elems) c.col_elems = ce;
    }
    const LayerShape s = layer_shape(c.layers[i]);
    c.act_elems[i + 1] = tensor_elems(s.hout, s.wout, c.layers[i].cout);
    if (c.act_elems[i + 1] > c.max_elems) c.max_elems = c.act_elems[i + 1];
    if (i + 1 < c.n) {
      const cgs_layer_desc& Ln = c.layers[i + 1];
      const size_t expect = (Ln.type == CGS_LAYER_FC) ? (size_t)Ln.cin : tensor_elems(Ln.hin, Ln.win, Ln.cin);
      if (expect != c.act_elems[i + 1])
        return set_error(CGS_ERR_INVALID, "layer %d output (%zu) does not feed layer %d input (%zu)", i,
                         c.act_elems[i + 1], i + 1, expect);
    }
  }
  if ((size_t)c.head.cin != c.act_elems[c.n]) return set_error(CGS_ERR_INVALID, "head input size mismatch");
  if (c.head.cin % 4) return set_error(CGS_ERR_UNSUPPORTED, "head input size must be a multiple of 4");
  return CGS_OK;
}

// Image-like tensors of the chain live in the s2d layout when both image-edge layers run on the tcgen05 edge kernels
// (edge_tc.cu): G's last deconv (64 -> <= 3 channels) and D's first conv (<= 3 -> 64 channels), rows of >= 16 pixels.
// CGS_DEBUG bits 4096 (general lowerings) and 2097152 (mma.sync edge kernels) keep the pitched layout.
static bool chain_s2d(const Chain& c, int math) {
  if (math != CGS_MATH_TF32_TENSOR || (debug_flags() & (4096 | 2097152))) return false;
  if (c.n_gtail < 1 || c.n_gtail >= c.n) return false;
  const cgs_layer_desc& Ln = c.layers[c.n_gtail - 1];
  const cgs_layer_desc& Lw = c.layers[c.n_gtail];
  if (Ln.type != CGS_LAYER_DECONV || Lw.type != CGS_LAYER_CONV) return false;
  if (Ln.cin != 64 || Ln.cout > 3 || Lw.cout != 64 || Lw.cin != Ln.cout) return false;
  if ((Ln.k != 4 && Ln.k != 5) || (Lw.k != 4 && Lw.k != 5)) return false;
  if (Ln.win < 16 || Ln.win > 64 || Ln.hin != Ln.win || Lw.hin != 2 * Ln.hin || Lw.win != 2 * Ln.win) return false;
  if (same_pad_before(Lw.hin, Lw.k) != 1 || same_pad_before(2 * Ln.hin, Ln.k) != 1) return false;
  return true;
}

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct Workspace {
  float* act[2 * CGS_MAX_LAYERS + 1];
  float* g[2];
  float* mom;
  float* col;
  float* cur_logit;
  unsigned char* done;
  // early-exit compaction: ping-pong feature / momentum buffers, row maps, index list, the two live-row counters
  float* feat2;
  float* feat3;
  float* mom2;
  int* orig[2];
  int* idx;
  int* count;            // count[0], count[1]: live rows, ping-pong per step
  size_t total;
};

// act[0] is the caller's feature buffer; everything else is carved out of the workspace.
static void carve(const Chain& c, int64_t B, void* base, Workspace& w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = base ? (char*)base + off : nullptr; off += align256(bytes); return p; };
  w.act[0] = nullptr;
  for (int i = 1; i <= c.n; ++i) w.act[i] = (float*)take((size_t)B * c.act_elems[i] * 4);
  w.g[0] = (float*)take((size_t)B * c.max_elems * 4);
  w.g[1] = (float*)take((size_t)B * c.max_elems * 4);
  w.mom = (float*)take((size_t)B * c.act_elems[0] * 4);
  w.col = c.col_elems ? (float*)take((size_t)B * c.col_elems * 4) : nullptr;
  w.cur_logit = (float*)take((size_t)B * 4);
  w.done = (unsigned char*)take((size_t)B);
  w.feat2 = (float*)take((size_t)B * c.act_elems[0] * 4);
  w.feat3 = (float*)take((size_t)B * c.act_elems[0] * 4);
  w.mom2 = (float*)take((size_t)B * c.act_elems[0] * 4);
  w.orig[0] = (int*)take((size_t)B * 4);
  w.orig[1] = (int*)take((size_t)B * 4);
  w.idx = (int*)take((size_t)B * 4);
  w.count = (int*)take(16);
  w.total = off;
}

// The last hidden layer runs split-K and nothing but the head reads its output: the head kernel then adds the partial
// sums itself and the activation is never written (CGS_DEBUG bit 131072 keeps the separate reduce kernel).
static bool head_takes_partials(const Chain& c, const Workspace& w, int64_t B, int math) {
  const cgs_layer_desc& L = c.layers[c.n - 1];
  return use_fc_split(L, false, B, math, w.col) && !(debug_flags() & 131072) && cstride(L.cout) == c.head.cin;
}

static int run_forward(const Chain& c, const Workspace& w, int64_t B, int math, cudaStream_t st,
                       const int* live = nullptr) {
  const int img_mode = chain_s2d(c, math) ? IMG_S2D : IMG_PITCHED;
  for (int i = 0; i < c.n; ++i) {
    PassEpi e;
    e.live = live;
    e.epi = EPI_FWD;
    e.act = c.layers[i].act;
    // TF32 path: activations that feed another MMA are rounded to TF32 (RN) where they are produced, so the
    // tensor core's operand truncation is exact; the image (returned to the caller) and the head input stay FP32
    e.round_out = (math == CGS_MATH_TF32_TENSOR) && (i != c.n_gtail - 1) && (i != c.n - 1);
    if (i == c.n_gtail - 1 && i + 1 < c.n && img_mode == IMG_PITCHED) {
      // generator's last deconv + discriminator's first conv on the same image: one kernel when the shapes allow
      PassEpi e2;
      e2.live = live;
      e2.epi = EPI_FWD;
      e2.act = c.layers[i + 1].act;
      e2.round_out = (math == CGS_MATH_TF32_TENSOR) && (i + 1 != c.n - 1);
      const int rc = try_edge_pair(c.layers[i], c.layers[i + 1], false, B, w.act[i], w.act[i + 1], w.act[i + 2], e, e2, 1, math, st);
      if (rc < 0) return rc;
      if (rc == 1) { ++i; continue; }
    }
    const bool defer = (i == c.n - 1) && head_takes_partials(c, w, B, math);
    if (int rc = run_pass(c.layers[i], false, B, w.act[i], w.act[i + 1], e, w.col, math, st, img_mode, defer)) return rc;
  }
  return CGS_OK;
}

// Backward from dpre[c.n-1] (already in w.g[0]) down to the feature.  `upd` != null fuses the policy step into
// the last GEMM's epilogue; otherwise the raw gradient is written to grad_out.
static int run_backward(const Chain& c, const Workspace& w, int64_t B, int math, const ConvGemmParams* upd,
                        float* grad_out, cudaStream_t st, const int* live = nullptr) {
  const int img_mode = chain_s2d(c, math) ? IMG_S2D : IMG_PITCHED;
  int cur = 0;
  for (int i = c.n - 1; i >= 0; --i) {
    float* dst = (i == 0) ? (upd ? w.act[0] : grad_out) : w.g[cur ^ 1];
    PassEpi e;
    e.live = live;
    if (i > 0) {
      e.epi = EPI_BWD;
      e.aux = w.act[i];
      e.act = c.layers[i - 1].act;
      e.round_out = (math == CGS_MATH_TF32_TENSOR);
    } else {
      e.upd = upd;
    }
    if (i == c.n_gtail && i >= 1 && img_mode == IMG_PITCHED) {
      // data-gradients of D's first conv and G's last deconv: one kernel, the image gradient never leaves the SM
      PassEpi e2;
      e2.live = live;
      float* dst2 = (i - 1 == 0) ? (upd ? w.act[0] : grad_out) : w.g[cur ^ 1];
      if (i - 1 > 0) {
        e2.epi = EPI_BWD;
        e2.aux = w.act[i - 1];
        e2.act = c.layers[i - 2].act;
        e2.round_out = (math == CGS_MATH_TF32_TENSOR);
      } else {
        e2.upd = upd;
      }
      // the intermediate (image gradient) buffer is only described, never written: any valid pointer of that size
      const int rc = try_edge_pair(c.layers[i], c.layers[i - 1], true, B, w.g[cur], w.g[cur ^ 1], dst2, e, e2, 0, math, st);
      if (rc < 0) return rc;
      if (rc == 1) { --i; cur ^= 1; continue; }
    }
    if (int rc = run_pass(c.layers[i], true, B, w.g[cur], dst, e, w.col, math, st, img_mode)) return rc;
    cur ^= 1;
  }
  return CGS_OK;
}

}  // namespace cgs

using namespace cgs;

// Largest batch one call can take (rows are indexed with 32 bits); larger batches are refined in chunks by the caller.
extern "C" int64_t cgs_refine_max_batch(const cgs_net_desc* gtail, const cgs_net_desc* d) {
  Chain c;
  if (int rc = build_chain(gtail, d, c)) return rc;
  const size_t per = c.max_elems > c.col_elems ? c.max_elems : c.col_elems;
  return (int64_t)(((1ull << 31) - 1) / per) - 1;
}

extern "C" size_t cgs_refine_workspace_bytes(const cgs_net_desc* gtail, const cgs_net_desc* d, int64_t B) {
  Chain c;
  if (build_chain(gtail, d, c) != CGS_OK || B < 0) return 0;
  Workspace w;
  carve(c, B, nullptr, w);
  return w.total + 256;
}

static int head_launch(const Chain& c, const Workspace& w, int64_t B, HeadParams hp, int math, cudaStream_t st) {
  hp.feat = w.act[c.n];
  if (head_takes_partials(c, w, B, math)) {          // run_forward left the split-K partial sums in the scratch buffer
    const cgs_layer_desc& L = c.layers[c.n - 1];
    hp.part = w.col;
    hp.S = fc_split(L, false);
    hp.fc_bias = L.bias;
    hp.fc_slope = L.act == ACT_RELU ? 0.f : (L.act == ACT_LRELU ? 0.2f : 1.f);
    hp.fc_tanh = (L.act == ACT_TANH);
  }
  hp.w = c.head.w_fwd;
  hp.bias = c.head.bias;
  hp.K = c.head.cin;
  hp.act = c.layers[c.n - 1].act;
  hp.img = w.act[c.n_gtail];
  {
    const cgs_layer_desc& Li = c.layers[c.n_gtail - 1];          // producer of the image
    const LayerShape si = layer_shape(Li);
    const int c4 = si.cs_out / 4;
    const bool pitched = Li.cout <= 4;
    hp.img_h = si.hout;
    hp.img_w4 = si.wout * c4;
    hp.img_pitch4 = (pitched ? img_pitch(si.wout) : si.wout) * c4;
    hp.img_xoff4 = pitched ? IMG_XOFF * c4 : 0;
    hp.img_elems = si.hout * si.wout * si.cs_out;
    if (chain_s2d(c, math)) {
      hp.img_s2d = 1;
      hp.img_w4 = si.wout;                      // one float4 per pixel
    }
  }
  hp.feature = w.act[0];
  hp.feat_elems = (int)c.act_elems[0];
  hp.cur_logit = w.cur_logit;
  cudaError_t le = launch_pdl(head_kernel, dim3((unsigned)B), dim3(256), (size_t)0, st, hp);
  count_launch();
  if (le != cudaSuccess) return set_error(CGS_ERR_CUDA, "head_kernel: %s", cudaGetErrorString(le));
  return check_launch("head_kernel");
}

extern "C" int cgs_refine_conv(const cgs_net_desc* gtail, const cgs_net_desc* d, const cgs_refine_cfg* cfg,
                               int64_t B, float* feature, float* best_img, float* best_logit, float* best_step,
                               float* default_logit, const int32_t* prob_indices, float* best_feature,
                               void* workspace, size_t workspace_bytes, cgs_stream_t stream) {
  if (int rc = require_sm100()) return rc;
  if (!cfg || !feature || !best_img || !best_logit || !best_step) return set_error(CGS_ERR_INVALID, "null argument");
  if (B <= 0) return B == 0 ? CGS_OK : set_error(CGS_ERR_INVALID, "negative batch");
  if (cfg->method != CGS_POLICY_SGD && cfg->method != CGS_POLICY_MOMENTUM)
    return set_error(CGS_ERR_UNSUPPORTED, "graph refiner supports sgd / momentum only (sampling/policy.py:51)");
  if (cfg->mode == CGS_MODE_PROBABILISTIC && !prob_indices)
    return set_error(CGS_ERR_INVALID, "probabilistic mode needs prob_indices");
  if (cfg->mode != CGS_MODE_PROBABILISTIC && cfg->mode != CGS_MODE_DETERMINISTIC)
    return set_error(CGS_ERR_UNSUPPORTED, "unknown mode %d", cfg->mode);
  Chain c;
  if (int rc = build_chain(gtail, d, c)) return rc;
  // programmatic dependent launch for this call's kernels when the average pass is large (common.h launch_pdl):
  // forward + data-gradient FLOPs of one iteration over its 2 n + 1 launches
  double macs = 0.0;
  for (int i = 0; i < c.n; ++i) {
    const cgs_layer_desc& L = c.layers[i];
    const double taps = L.type == CGS_LAYER_FC ? 1.0 : (double)L.k * L.k;
    const double px = L.type == CGS_LAYER_FC ? 1.0 : (L.type == CGS_LAYER_CONV ? (double)((L.hin + 1) / 2) * ((L.win + 1) / 2)
                                                                              : (double)L.hin * L.win);
    macs += px * taps * L.cin * L.cout;
  }
  const PdlScope pdl_scope(4.0 * macs * (double)B / (2.0 * c.n + 1.0) >= 12e9);     // MNIST B=1024: 7.3 GFLOP (off), DCGAN-32: 18 (on), DCGAN-64: 78 (on)
  Workspace w;
  carve(c, B, (void*)(((uintptr_t)workspace + 255) & ~uintptr_t(255)), w);
  if (!workspace || w.total + 256 > workspace_bytes) return set_error(CGS_ERR_WORKSPACE, "workspace too small");
  if (B > cgs_refine_max_batch(gtail, d))
    return set_error(CGS_ERR_UNSUPPORTED, "batch too large for 32-bit row indexing; split it (cgs_refine_max_batch)");
  w.act[0] = feature;
  cudaStream_t st = (cudaStream_t)stream;
  HeadParams hp;
  std::memset(&hp, 0, sizeof(hp));
  hp.best_img = best_img;
  hp.best_logit = best_logit;
  hp.best_step = best_step;
  hp.default_logit = default_logit;
  hp.best_feature = best_feature;
  hp.prob_indices = prob_indices;
  hp.mode = cfg->mode;
  hp.exit_logit = cfg->exit_logit;
  hp.round_out = (cfg->math == CGS_MATH_TF32_TENSOR);
  if (cfg->early_exit) {
    hp.done = w.done;
    cudaMemsetAsync(w.done, 0, (size_t)B, st);
  }
  const int K = cfg->steps;
  const bool compacting = cfg->early_exit != 0;
  if (compacting && cfg->mode != CGS_MODE_DETERMINISTIC)
    return set_error(CGS_ERR_UNSUPPORTED, "early exit is defined for the deterministic mode only");
  float* feature_out = feature;             // the caller's buffer: final state of every sample ends up here
  float* mom_cur = w.mom;
  float* mom_alt = w.mom2;
  const int* live = nullptr;                // device-resident count of the rows still in the batch (early exit)
  int cur = 0;                              // which orig[] map / live counter is current
  if (compacting) {
    // work on a private copy so that exited samples can be dropped and the rest compacted (ping-pong feat2 / feat3)
    cudaMemcpyAsync(w.feat2, feature, (size_t)B * c.act_elems[0] * 4, cudaMemcpyDeviceToDevice, st);
    w.act[0] = w.feat2;
    iota_kernel<<<148, 256, 0, st>>>(w.orig[0], (int)B, w.count); count_launch();
    hp.orig = w.orig[0];
    hp.final_feature = feature_out;
    live = w.count;
    hp.live = live;
  }
  // initial evaluation: collaborator.py:48-60
  if (int rc = run_forward(c, w, B, cfg->math, st, live)) return rc;
  hp.step = -1;
  hp.dpre = K > 0 ? w.g[0] : nullptr;
  if (int rc = head_launch(c, w, B, hp, cfg->math, st)) return rc;
  ConvGemmParams upd;
  std::memset(&upd, 0, sizeof(upd));
  upd.sgd = cfg->method == CGS_POLICY_SGD;
  upd.rate = cfg->rate;
  upd.alpha = cfg->alpha;
  upd.clip = cfg->clip;
  upd.vmin = cfg->vmin;
  upd.vmax = cfg->vmax;
  const int elems4 = (int)(c.act_elems[0] / 4);
  for (int i = 0; i < K; ++i) {                       // collaborator.py:63-83
    upd.first = (i == 0);
    upd.mom = mom_cur;
    if (int rc = run_backward(c, w, B, cfg->math, &upd, nullptr, st, live)) return rc;   // grad + policy step (:66-70)
    if (compacting) {
      // drop the samples D already classifies as real (README.md:13): ordered compaction of the rows that stay.
      // Only the feature map, its momentum and the row map are live here (activations / gradients are dead).
      // Everything is sized by the device-side counter: no host read-back, static launch sequence.
      int* live_next = w.count + (cur ^ 1);
      live_compact_kernel<<<1, 1024, 0, st>>>(w.done, live, (int)B, w.idx, live_next); count_launch();
      float* feat_src = w.act[0];
      float* feat_dst = (feat_src == w.feat2) ? w.feat3 : w.feat2;
      const int grid = (int)(B < 148 * 8 ? B : 148 * 8);
      live_gather_kernel<<<grid, 256, 0, st>>>((const float4*)feat_src, (float4*)feat_dst,
                                               upd.sgd ? nullptr : (const float4*)mom_cur, (float4*)mom_alt,
                                               w.orig[cur], w.orig[cur ^ 1], w.idx, live_next, elems4); count_launch();
      if (int rc = check_launch("early-exit compaction")) return rc;
      w.act[0] = feat_dst;
      float* t = mom_cur; mom_cur = mom_alt; mom_alt = t;
      cur ^= 1;
      hp.orig = w.orig[cur];
      live = live_next;
      hp.live = live;
    }
    if (int rc = run_forward(c, w, B, cfg->math, st, live)) return rc;                   // :73
    hp.step = i;
    hp.dpre = (i + 1 < K) ? w.g[0] : nullptr;         // the gradient after the last step is never consumed
    if (int rc = head_launch(c, w, B, hp, cfg->math, st)) return rc;                          // :76-83
  }
  if (compacting) {
    scatter_active_kernel<<<(unsigned)B, 128, 0, st>>>((const float4*)w.act[0], hp.orig, w.done, live, (float4*)feature_out,
                                                       elems4); count_launch();
    if (int rc = check_launch("scatter_active_kernel")) return rc;
  }
  return CGS_OK;
}

extern "C" int cgs_forward_logits_and_grad(const cgs_net_desc* gtail, const cgs_net_desc* d, int math, int64_t B,
                                           const float* feature, float* logit_out, float* grad_out, float* img_out,
                                           void* workspace, size_t workspace_bytes, cgs_stream_t stream) {
  if (int rc = require_sm100()) return rc;
  if (!feature || !logit_out) return set_error(CGS_ERR_INVALID, "null argument");
  if (B <= 0) return B == 0 ? CGS_OK : set_error(CGS_ERR_INVALID, "negative batch");
  Chain c;
  if (int rc = build_chain(gtail, d, c)) return rc;
  Workspace w;
  carve(c, B, (void*)(((uintptr_t)workspace + 255) & ~uintptr_t(255)), w);
  if (!workspace || w.total + 256 > workspace_bytes) return set_error(CGS_ERR_WORKSPACE, "workspace too small");
  if (B > cgs_refine_max_batch(gtail, d))
    return set_error(CGS_ERR_UNSUPPORTED, "batch too large for 32-bit row indexing; split it (cgs_refine_max_batch)");
  w.act[0] = const_cast<float*>(feature);
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = run_forward(c, w, B, math, st)) return rc;
  HeadParams hp;
  std::memset(&hp, 0, sizeof(hp));
  // selection outputs are not wanted here: point them at scratch so the kernel stays branch-free
  hp.best_img = w.g[1];
  hp.best_logit = w.cur_logit;
  hp.best_step = w.mom;          // scratch (>= B floats)
  hp.step = -1;
  hp.round_out = (math == CGS_MATH_TF32_TENSOR);
  hp.dpre = grad_out ? w.g[0] : nullptr;
  if (int rc = head_launch(c, w, B, hp, math, st)) return rc;
  cudaMemcpyAsync(logit_out, w.cur_logit, (size_t)B * 4, cudaMemcpyDeviceToDevice, st);
  if (img_out) {
    const cgs_layer_desc& Li = c.layers[c.n_gtail - 1];
    const LayerShape si = layer_shape(Li);
    if (Li.cout <= 4 && chain_s2d(c, math)) {
      if (int rc = s2d_to_image(w.act[c.n_gtail], img_out, B, si.hout, si.wout, st)) return rc;
    } else if (Li.cout <= 4) {
      if (int rc = unpad_image(w.act[c.n_gtail], img_out, B, si.hout, si.wout, st)) return rc;
    } else {
      cudaMemcpyAsync(img_out, w.act[c.n_gtail], (size_t)B * c.act_elems[c.n_gtail] * 4, cudaMemcpyDeviceToDevice, st);
    }
  }
  if (grad_out) {
    if (int rc = run_backward(c, w, B, math, nullptr, grad_out, st)) return rc;
  }
  return check_launch("cgs_forward_logits_and_grad");
}

static size_t layer_stage_elems(const cgs_layer_desc& L, bool backward) {   // pitched copy of an image-like input
  if (!use_window(L, backward)) return 0;
  const LayerShape s = layer_shape(L);
  return backward ? tensor_elems(s.hout, s.wout, L.cout) : tensor_elems(L.hin, L.win, L.cin);
}

extern "C" size_t cgs_layer_workspace_bytes(const cgs_layer_desc* L, int64_t B) {
  if (!L || B < 0) return 0;
  size_t ce = scatter_col_elems(*L, false);
  if (scatter_col_elems(*L, true) > ce) ce = scatter_col_elems(*L, true);
  size_t se = layer_stage_elems(*L, false);
  if (layer_stage_elems(*L, true) > se) se = layer_stage_elems(*L, true);
  return (ce + se) * (size_t)B * 4 + 512;
}

// The single-layer entry points take and return DENSE tensors; image-like inputs of window-lowered passes are
// re-laid out into the workspace first (inside the refinement chain they are produced pitched, no copy).
static int layer_pass_dense(const cgs_layer_desc& L, bool backward, int math, int64_t B, const float* in, float* out,
                            const PassEpi& e, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const bool needs_ws = use_scatter(L, backward) || use_window(L, backward);
  if (needs_ws && (!workspace || workspace_bytes < cgs_layer_workspace_bytes(&L, B)))
    return set_error(CGS_ERR_WORKSPACE, "workspace too small");
  float* base = (float*)(((uintptr_t)workspace + 255) & ~uintptr_t(255));
  float* col = (workspace && workspace_bytes >= cgs_layer_workspace_bytes(&L, B)) ? base : nullptr;
  size_t ce = scatter_col_elems(L, false);
  if (scatter_col_elems(L, true) > ce) ce = scatter_col_elems(L, true);
  float* stage = base + ((ce * (size_t)B + 63) & ~size_t(63));
  int img_mode = IMG_DENSE;
  if (use_window(L, backward)) {
    const LayerShape s = layer_shape(L);
    const int h = backward ? s.hout : L.hin, w = backward ? s.wout : L.win;
    // the image-like input is re-laid out first (inside the chain it is produced in that layout: no copy there)
    ConvGemmParams wp;
    EdgeWideParams q;
    const bool tc = math == CGS_MATH_TF32_TENSOR && !(debug_flags() & (4096 | 2097152)) && w >= 32 && w <= 128 &&
                    make_window_params(L, backward, B, in, out, wp) == CGS_OK && make_edge_wide(L, backward, B, in, out, e, wp, q) &&
                    edge_wide_tc_supported(q);
    if (tc) {
      if (int rc = image_to_s2d(in, stage, B, h, w, st)) return rc;
      img_mode = IMG_S2D;
    } else {
      if (int rc = pad_image(in, stage, B, h, w, st)) return rc;
      img_mode = IMG_PITCHED;                 // (the window passes read the pitched layout; their output is not image-like)
    }
    in = stage;
  }
  return run_pass(L, backward, B, in, out, e, col, math, st, img_mode);
}

extern "C" int cgs_layer_forward(const cgs_layer_desc* L, int math, int64_t B, const float* x, float* y,
                                 void* workspace, size_t workspace_bytes, cgs_stream_t stream) {
  if (int rc = require_sm100()) return rc;
  if (!L || !x || !y) return set_error(CGS_ERR_INVALID, "null argument");
  PassEpi e;
  e.epi = EPI_FWD;
  e.act = L->act;
  return layer_pass_dense(*L, false, math, B, x, y, e, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int cgs_layer_backward(const cgs_layer_desc* L, int math, int64_t B, const float* dy, float* dx,
                                  const float* x_fwd, int prev_act, void* workspace, size_t workspace_bytes,
                                  cgs_stream_t stream) {
  if (int rc = require_sm100()) return rc;
  if (!L || !dy || !dx) return set_error(CGS_ERR_INVALID, "null argument");
  PassEpi e;
  if (x_fwd && prev_act != CGS_ACT_NONE) {
    e.epi = EPI_BWD;
    e.aux = x_fwd;
    e.act = prev_act;
  }
  return layer_pass_dense(*L, true, math, B, dy, dx, e, workspace, workspace_bytes, (cudaStream_t)stream);
}

// Host-only: 0 = the pass uses the gather layout described by cgs_pack_map, 1 = scatter layout
// (rows = (ky*k + kx)*4 + small channel, K = the large channel count; see "Scatter formulation" above),
// 2 = window layout (rows = large channel, K index = ky*32 + kx*4 + small channel, zero elsewhere).
extern "C" int cgs_pass_layout(const cgs_layer_desc* L, int backward) {
  if (!L) return set_error(CGS_ERR_INVALID, "null layer");
  if (use_window(*L, backward != 0)) return 2;
  return use_scatter(*L, backward != 0) ? 1 : 0;
}

// Host-only: the K ordering of a layer's packed weight matrix, so the packer (cgs/pack.py) never has to
// re-derive the class / tap enumeration.  For K index q: ky[q], kx[q] = filter tap (or -1 for zero padding),
// ch[q] = reduced channel.  Returns the K length (multiple of 32), or a negative status.
extern "C" int64_t cgs_pack_map(const cgs_layer_desc* L, int backward, int32_t* ky, int32_t* kx, int32_t* ch,
                                int64_t capacity) {
  if (!L) return set_error(CGS_ERR_INVALID, "null layer");
  ConvGemmParams p;
  int rc = backward ? make_backward_params(*L, 1, nullptr, nullptr, p) : make_forward_params(*L, 1, nullptr, nullptr, p);
  if (rc) return rc;
  int64_t total = 0;
  for (int c = 0; c < p.nclasses; ++c) total += (int64_t)p.cls[c].nkb * 32;
  if (!ky || !kx || !ch) return total;
  if (capacity < total) return set_error(CGS_ERR_INVALID, "pack map capacity too small");
  if (L->type == CGS_LAYER_FC) {
    for (int64_t q = 0; q < total; ++q) { ky[q] = 0; kx[q] = 0; ch[q] = (int32_t)q; }
    return total;
  }
  // recover (ky,kx) from (dy,dx): strided pass dy = ky - pad ; transposed pass dy = (py + pad - ky)/2
  const bool strided = (p.S == 2);
  const int size_for_pad = (L->type == CGS_LAYER_CONV) ? L->hin : L->hin * 2;
  const int size_for_pad_x = (L->type == CGS_LAYER_CONV) ? L->win : L->win * 2;
  const int pad_y = same_pad_before(size_for_pad, L->k), pad_x = same_pad_before(size_for_pad_x, L->k);
  for (int c = 0; c < p.nclasses; ++c) {
    const GemmClass& g = p.cls[c];
    const int cin_k = p.cblocks ? p.cblocks * 32 : 4;
    const int64_t klen = (int64_t)g.nkb * 32;
    for (int64_t q = 0; q < klen; ++q) {
      const int t = (int)(q / cin_k);
      const int cc = (int)(q % cin_k);
      const int64_t o = g.k0 + q;
      if (t >= g.ntaps) { ky[o] = -1; kx[o] = -1; ch[o] = cc; continue; }
      if (strided) {
        ky[o] = g.dy[t] + pad_y;
        kx[o] = g.dx[t] + pad_x;
      } else {
        ky[o] = g.oy0 + pad_y - 2 * g.dy[t];
        kx[o] = g.ox0 + pad_x - 2 * g.dx[t];
      }
      ch[o] = cc;
    }
  }
  return total;
}

// Host-only introspection (no GPU needed): the gathered-GEMM parameters a layer pass is lowered to, flattened to
// int32 so tests can replay the exact gather on the CPU.  Layout: [IH, IW, Cs, cblocks, MH, MW, S, M, OH, OW, ON,
// os, N, nclasses, window, win_k, win_x0, in_pitch_px] then per class [k0, nkb, ntaps, oy0, ox0, dy[32], dx[32]].  Returns the number of ints.
// Host-only introspection of the class-fused lowering (include/cgs.h).
extern "C" int64_t cgs_debug_fusion_plan(const cgs_layer_desc* L, int backward, int64_t B, int32_t* out, int64_t capacity) {
  if (!L) return set_error(CGS_ERR_INVALID, "null layer");
  if (use_window(*L, backward != 0) || use_scatter(*L, backward != 0)) return 0;
  ConvGemmParams p;
  int rc = backward ? make_backward_params(*L, B, nullptr, nullptr, p) : make_forward_params(*L, B, nullptr, nullptr, p);
  if (rc) return rc;
  const int ns = plan_fusion(p);
  if (!ns) return 0;
  int nshf = 0;
  for (int g = 0; g < p.ngroups; ++g) nshf += p.grp[g].nshifts;
  const int64_t need = 2 + (int64_t)p.ngroups * 7 + (int64_t)nshf * 48;
  if (!out) return need;
  if (capacity < need) return set_error(CGS_ERR_INVALID, "capacity too small");
  int32_t* o = out;
  *o++ = ns; *o++ = p.ngroups;
  for (int g = 0; g < p.ngroups; ++g) {
    const FuseGroup& G = p.grp[g];
    *o++ = G.nshifts; *o++ = G.shift0; *o++ = G.ncls;
    for (int q = 0; q < 4; ++q) *o++ = G.cls[q];
  }
  for (int i = 0; i < nshf; ++i) {
    const FuseShift& s = p.shf[i];
    *o++ = s.dy; *o++ = s.dx; *o++ = s.ncls; *o++ = s.nrun;
    for (int q = 0; q < 4; ++q) *o++ = s.slot[q];
    for (int q = 0; q < 4; ++q) *o++ = s.katom0[q];
    for (int q = 0; q < 4; ++q) *o++ = s.run_slot[q];
    for (int q = 0; q < 4; ++q) *o++ = s.run_len[q];
    for (int q = 0; q < 4; ++q) *o++ = s.run_acc[q];
    for (int r = 0; r < 2; ++r) for (int q = 0; q < 4; ++q) *o++ = s.pc_slot[r][q];
    for (int r = 0; r < 2; ++r) for (int q = 0; q < 4; ++q) *o++ = s.pc_half[r][q];
    for (int r = 0; r < 2; ++r) for (int q = 0; q < 4; ++q) *o++ = s.pc_katom[r][q];
  }
  return need;
}

extern "C" int64_t cgs_debug_gemm_params(const cgs_layer_desc* L, int backward, int64_t B, int32_t* out,
                                         int64_t capacity) {
  if (!L) return set_error(CGS_ERR_INVALID, "null layer");
  ConvGemmParams p;
  int rc = use_window(*L, backward != 0) ? make_window_params(*L, backward != 0, B, nullptr, nullptr, p)
           : use_scatter(*L, backward != 0) ? make_scatter_gemm_params(*L, backward != 0, B, nullptr, nullptr, p)
           : backward ? make_backward_params(*L, B, nullptr, nullptr, p) : make_forward_params(*L, B, nullptr, nullptr, p);
  if (rc) return rc;
  const int64_t need = 18 + (int64_t)p.nclasses * (5 + 2 * kMaxTaps);
  if (!out) return need;
  if (capacity < need) return set_error(CGS_ERR_INVALID, "capacity too small");
  int32_t* o = out;
  const int head[18] = {p.IH, p.IW, p.Cs, p.cblocks, p.MH, p.MW, p.S, p.M, p.OH, p.OW, p.ON, p.os, p.N, p.nclasses,
                        p.window, p.win_k, p.win_x0, p.in_pitch_px};
  for (int i = 0; i < 18; ++i) *o++ = head[i];
  for (int c = 0; c < p.nclasses; ++c) {
    const GemmClass& g = p.cls[c];
    *o++ = g.k0; *o++ = g.nkb; *o++ = g.ntaps; *o++ = g.oy0; *o++ = g.ox0;
    for (int t = 0; t < kMaxTaps; ++t) *o++ = g.dy[t];
    for (int t = 0; t < kMaxTaps; ++t) *o++ = g.dx[t];
  }
  return need;
}
