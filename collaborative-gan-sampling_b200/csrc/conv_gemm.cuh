// Gathered-A implicit GEMM: the one contraction kernel behind every conv / deconv / fc pass of the
// refinement loop (forward AND data-gradient; weights are frozen so there is no wgrad).
//
//   out[b, j*os+oy0, i*os+ox0, n] = epi( sum_{t<ntaps} sum_{c<Cin} in[b, j*S+dy[t], i*S+dx[t], c] * W[n][k0 + t*Cin + c] )
//
// * stride-2 conv fprop (nsgan/ops.py:41) and deconv-backward: S=2, os=1, one class, taps = all (ky,kx)
// * stride-2 deconv fprop (nsgan/ops.py:55) and conv data-gradient: S=1, os=2, four output-parity classes,
//   each with its own tap subset (k4: 2x2 taps; k5: 3x3/3x2/2x3/2x2) -- no zero-insertion MACs
// * fc (nsgan/ops.py:81): IH=IW=MH=MW=1, one tap
// Rows m = (b, j, i) are gathered (zero-filled outside the image) into 128B-swizzled shared memory;
// W is a plain K-major matrix fetched by TMA.  Pixel mode (Cs == 4) packs 8 taps x 4 channels per K block
// for the 1-/3-channel image layers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cgs {

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3 };
enum : int {
  EPI_FWD = 0,     // out = act(acc + bias)
  EPI_BWD = 1,     // out = acc * act'(aux)      (aux = forward OUTPUT of the layer being differentiated)
  EPI_UPDATE = 2,  // acc is d loss / d feature: momentum/sgd step fused (sampling/policy.py:27-37)
  EPI_RAW = 3      // out = acc
};

constexpr int kMaxTaps = 32;
constexpr int kMaxClasses = 4;

struct GemmClass {
  int k0;      // first K element of this class inside the packed weight matrix
  int nkb;     // K blocks (32 floats) of this class
  int ntaps;   // taps of this class
  int oy0, ox0;
  signed char dy[kMaxTaps];
  signed char dx[kMaxTaps];
};

struct ConvGemmParams {
  const float* in;
  float* out;
  const float* bias;
  const float* aux;
  float* mom;
  int IH, IW, Cs;     // input image, channel stride (floats per pixel)
  int cblocks;        // Cin / 32 in block mode; 0 in pixel mode
  int MH, MW, S;      // rows per image = MH*MW; input pixel = (j*S+dy, i*S+dx)
  int M;              // B*MH*MW
  int OH, OW, ON, os; // output image, channel stride, output pixel stride
  int N;              // valid output channels
  int n_tiles;        // ceil(N / BLOCK_N)
  int m_tiles;        // ceil(M / 128)
  int nclasses;
  int epi, act;
  int first, clip, sgd;
  float rate, alpha, vmin, vmax;
  GemmClass cls[kMaxClasses];
};

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_LRELU: return fmaxf(v, 0.2f * v);   // nsgan/ops.py:69-70
    case ACT_TANH: return tanhf(v);
    default: return v;
  }
}
// derivative expressed through the layer OUTPUT y = act(pre)
__device__ __forceinline__ float act_grad_from_output(float y, int act) {
  switch (act) {
    case ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case ACT_LRELU: return y > 0.f ? 1.f : 0.2f;
    case ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}

// One output element.  `off` = element offset of (row, n) in out / aux / mom.
__device__ __forceinline__ float epilogue_value(const ConvGemmParams& p, size_t off, int n, float acc) {
  if (p.epi == EPI_FWD) {
    return act_apply(acc + (p.bias ? __ldg(p.bias + n) : 0.f), p.act);
  } else if (p.epi == EPI_BWD) {
    return acc * act_grad_from_output(__ldg(p.aux + off), p.act);
  } else if (p.epi == EPI_UPDATE) {
    float m;
    if (p.sgd) {
      m = p.rate * acc;                                            // policy.py:28
    } else {
      m = p.first ? p.rate * acc : p.alpha * p.mom[off] + p.rate * acc;   // policy.py:32-35
      p.mom[off] = m;
    }
    float h = p.out[off] - m;                                      // policy.py:28,36
    if (p.clip) h = fminf(fmaxf(h, p.vmin), p.vmax);               // collaborator.py:69-70
    return h;
  }
  return acc;
}

// host launchers (conv_gemm.cu)
int launch_conv_gemm_tc(const ConvGemmParams& p, const float* w, int w_rows, int w_cols, cudaStream_t stream);
int launch_conv_gemm_simt(const ConvGemmParams& p, const float* w, int w_rows, int w_cols, cudaStream_t stream);

}  // namespace cgs
