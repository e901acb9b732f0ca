// Image-edge passes (see edge_conv.cuh): fused im2col / col2im + mma.sync TF32 + epilogue streaming kernels.
//
// Reference semantics: nsgan/ops.py:41 (conv2d, k x k, stride 2, SAME), nsgan/ops.py:55 (deconv2d) and their
// tf.gradients data-gradients (sampling/collaborator.py:31); epilogues as in conv_gemm.cuh.
#include "edge_conv.cuh"

#include <cstdint>
#include <type_traits>

#include "common.h"
#include "conv_gemm.cuh"
#include "ptx.cuh"

namespace cgs {

namespace {

// D(16x8, fp32) += A(16x8, tf32, row) * B(8x8, tf32, col); g = lane / 4, t = lane % 4:
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);  b0 (k = t, n = g)  b1 (k = t+4, n = g)
//   c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Scalar twin of epilogue4 (conv_gemm.cuh): same formulas, same separately rounded policy arithmetic.
__device__ __forceinline__ float epilogue1(const EdgeEpi& e, float a, float x0, float x1, float* m_out) {
  float o = a;
  if (e.epi == EPI_FWD) {
    const float v = a + x0;
    o = e.act_tanh ? tanhf(v) : fmaxf(v, v * e.slope);
  } else if (e.epi == EPI_BWD) {
    o = e.act_tanh ? a * (1.f - x0 * x0) : a * (x0 > 0.f ? 1.f : e.slope);
  } else if (e.epi == EPI_UPDATE) {
    float m = __fmul_rn(e.rate, a);                                   // sampling/policy.py:27-37
    if (!e.sgd) {
      if (!e.first) m = __fadd_rn(__fmul_rn(e.alpha, x1), m);
      *m_out = m;
    }
    o = __fsub_rn(x0, m);
    if (e.clip) o = fminf(fmaxf(o, e.vmin), e.vmax);                  // collaborator.py:69-70
    return o;
  }
  return e.round_out ? tf32_rn(o) : o;
}

// ---------------------------------------------------------------------------------------------
// edge_wide: per tile (image b, band of RO output rows <= 256 pixels) the input rows the band needs are staged in
// shared memory by cp.async (rows outside the image zero-filled), double-buffered so the next tile's rows arrive
// while this tile computes; the im2col gather of the A fragments is then plain LDS with no bounds tests.
// N = 64, K = k*k*cimg.  CTA = 8 warps; a warp takes 16 pixels x 64 channels at a time (8 mma tiles per k-step,
// 32 accumulator registers) so that four CTAs fit an SM: the kernel is latency bound, warps in flight are what
// buys bandwidth.  Weights live in shared memory in fragment order (one LDS.64 per B fragment).
// ---------------------------------------------------------------------------------------------
constexpr int EW_THREADS = 256;
constexpr int EW_NT = 8;            // N = 64
constexpr int EW_KMAX = 104;        // k <= 5, cimg <= 4
constexpr int EW_TILE_PX = EW_THREADS;   // 32 pixels per warp

// Epilogue modes compiled into the kernel (the per-element code must stay a handful of instructions: it runs 64
// times per thread and tile): slope-type forward / derivative, the policy step, and a generic fallback (tanh, raw).
enum : int { EM_FWD_SLOPE = 0, EM_BWD_SLOPE = 1, EM_UPDATE = 2, EM_GENERIC = 3 };

template <int MODE, bool ROUND>
__device__ __forceinline__ float epi_fast(const EdgeEpi& e, float a, float x0, float x1, float* m_out) {
  float o;
  if (MODE == EM_FWD_SLOPE) {
    const float v = a + x0;
    o = fmaxf(v, v * e.slope);
  } else if (MODE == EM_BWD_SLOPE) {
    o = a * (x0 > 0.f ? 1.f : e.slope);
  } else {
    return epilogue1(e, a, x0, x1, m_out);
  }
  return ROUND ? tf32_rn(o) : o;
}

// One-time set-up of an edge_wide CTA: k -> patch offset table, bias, weights in mma fragment order.
__device__ __forceinline__ void wide_setup(const EdgeWideParams& p, int ksteps, float2* bfrag, int* koff_s, float* bias_s) {
  const int tid = threadIdx.x;
  const int CI = p.cimg;
  const int kreal = p.k * p.k * CI;
  const int prow = p.pitch * 4;              // floats per staged row
  for (int kk = tid; kk < ksteps * 8; kk += EW_THREADS) {
    int off = 0;                             // padding k slots read a valid word and multiply a zero weight
    if (kk < kreal) {
      const int tap = kk / CI, c = kk - tap * CI;
      const int ky = tap / p.k, kx = tap - ky * p.k;
      off = ky * prow + (kx - p.pad_x + p.xoff) * 4 + c;
    }
    koff_s[kk] = off;
  }
  if (tid < EW_NT * 8) bias_s[tid] = (p.e.epi == EPI_FWD && p.e.bias && tid < p.N) ? __ldg(p.e.bias + tid) : 0.f;
  for (int idx = tid; idx < ksteps * EW_NT * 32; idx += EW_THREADS) {
    const int l = idx & 31, nt = (idx >> 5) % EW_NT, ks = idx / (EW_NT * 32);
    const int n = nt * 8 + (l >> 2);
    float b[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = ks * 8 + (l & 3) + 4 * h;
      b[h] = 0.f;
      if (kk < kreal && n < p.N) {
        const int tap = kk / CI, c = kk - tap * CI;
        const int ky = tap / p.k, kx = tap - ky * p.k;
        b[h] = __ldg(p.w + (size_t)n * (p.k * 32) + ky * 32 + kx * 4 + c);
      }
    }
    bfrag[idx] = make_float2(b[0], b[1]);
  }
}

// The band's GEMM + epilogue: `patch` holds the staged input rows ([rows][pitch][4] floats, row 0 = image row
// 2*oy0 - pad_y), npx output pixels starting at global pixel m0.  Each warp takes 16 pixels x 64 channels at a time.
template <int MODE, bool ROUND>
__device__ __forceinline__ void wide_band(const EdgeWideParams& p, int ksteps, const float* patch, const float2* bfrag,
                                          const int* koff_s, const float* bias_s, int npx, int m0,
                                          unsigned long long fd_ow) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int prow = p.pitch * 4;
  for (int mt = warp; mt * 16 < npx; mt += EW_THREADS / 32) {
    // the two pixels (rows g, g+8 of the m16 tile) this thread gathers for and stores
    int base[2], mloc[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ml = mt * 16 + h * 8 + g;
      const int oyr = fast_div(ml, fd_ow), ox = ml - oyr * p.OW;
      const bool ok = ml < npx;
      base[h] = ok ? (2 * oyr * prow + 2 * ox * 4) : 0;
      mloc[h] = ok ? ml : -1;
    }
    float acc[EW_NT][4];
#pragma unroll
    for (int nt = 0; nt < EW_NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll 2
    for (int ks = 0; ks < ksteps; ++ks) {
      const int o0 = koff_s[ks * 8 + t];
      const int o1 = koff_s[ks * 8 + t + 4];
      const uint32_t a0 = __float_as_uint(patch[base[0] + o0]);
      const uint32_t a1 = __float_as_uint(patch[base[1] + o0]);
      const uint32_t a2 = __float_as_uint(patch[base[0] + o1]);
      const uint32_t a3 = __float_as_uint(patch[base[1] + o1]);
#pragma unroll
      for (int nt = 0; nt < EW_NT; ++nt) {
        const float2 bq = bfrag[(ks * EW_NT + nt) * 32 + lane];
        mma_tf32(acc[nt], a0, a1, a2, a3, __float_as_uint(bq.x), __float_as_uint(bq.y));
      }
    }
    // epilogue: thread owns columns nt*8 + 2t, +1 of its two rows; a quad writes one 32-byte sector
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (mloc[h] < 0) continue;
      const size_t ro = (size_t)(m0 + mloc[h]) * p.ON + 2 * t;
      float* orow = p.out + ro;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float2 x0[EW_NT / 2], x1[EW_NT / 2];
#pragma unroll
        for (int q = 0; q < EW_NT / 2; ++q) {
          const int nt = half * (EW_NT / 2) + q;
          x1[q] = make_float2(0.f, 0.f);
          if (MODE == EM_FWD_SLOPE) {
            x0[q] = *reinterpret_cast<const float2*>(bias_s + nt * 8 + 2 * t);
          } else if (MODE == EM_BWD_SLOPE) {
            x0[q] = __ldg(reinterpret_cast<const float2*>(p.e.aux + ro + nt * 8));
          } else {
            x0[q] = *reinterpret_cast<const float2*>(bias_s + nt * 8 + 2 * t);
            if (p.e.epi == EPI_BWD) {
              x0[q] = __ldg(reinterpret_cast<const float2*>(p.e.aux + ro + nt * 8));
            } else if (p.e.epi == EPI_UPDATE) {
              x0[q] = *reinterpret_cast<const float2*>(orow + nt * 8);
              if (!p.e.sgd && !p.e.first) x1[q] = *reinterpret_cast<const float2*>(p.e.mom + ro + nt * 8);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < EW_NT / 2; ++q) {
          const int nt = half * (EW_NT / 2) + q;
          float2 m2 = make_float2(0.f, 0.f), o;
          o.x = epi_fast<MODE, ROUND>(p.e, acc[nt][2 * h], x0[q].x, x1[q].x, &m2.x);
          o.y = epi_fast<MODE, ROUND>(p.e, acc[nt][2 * h + 1], x0[q].y, x1[q].y, &m2.y);
          if (MODE >= EM_UPDATE && p.e.epi == EPI_UPDATE && !p.e.sgd) *reinterpret_cast<float2*>(p.e.mom + ro + nt * 8) = m2;
          *reinterpret_cast<float2*>(orow + nt * 8) = o;
        }
      }
    }
  }
}

template <int MODE, bool ROUND>
__global__ void __launch_bounds__(EW_THREADS, 4) edge_wide_kernel(const EdgeWideParams p, int ksteps, int ntiles, int RO,
                                                                  int bands, unsigned long long fd_bands,
                                                                  unsigned long long fd_ow) {
  extern __shared__ float4 ew_smem4[];
  const int PR = 2 * RO + p.k - 2;           // staged input rows per band
  const int patch_px = PR * p.pitch;         // pixels (float4) per patch buffer
  float2* bfrag = reinterpret_cast<float2*>(ew_smem4);                          // [ksteps][8][32]
  float4* patch4 = ew_smem4 + ksteps * EW_NT * 32 / 2;                          // [2][PR][pitch]
  __shared__ int koff_s[EW_KMAX];
  __shared__ float bias_s[EW_NT * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  wide_setup(p, ksteps, bfrag, koff_s, bias_s);

  // async copy of the input rows of one tile into patch buffer `buf`
  auto stage_patch = [&](int tile, int buf) {
    const int b = fast_div(tile, fd_bands);
    const int iy0 = 2 * (tile - b * bands) * RO - p.pad_y;       // image row of staged row 0
    const uint32_t dst0 = smem_u32(patch4 + buf * patch_px);
    for (int pr = warp; pr < PR; pr += EW_THREADS / 32) {
      const int iy = iy0 + pr;
      const bool ok = (unsigned)iy < (unsigned)p.IH;
      const float4* src = reinterpret_cast<const float4*>(p.in) + (size_t)(b * p.IH + (ok ? iy : 0)) * p.pitch;
      for (int x = lane; x < p.pitch; x += 32)
        cp_async_16(dst0 + (pr * p.pitch + x) * 16, src + x, ok ? 16u : 0u);
    }
  };

  pdl_launch_dependents();
  pdl_wait();                                // tables and weights above are constants; activations from here on
  if (p.e.live) ntiles = min(ntiles, live_images(p.e.live, ntiles / bands) * bands);
  int tile = blockIdx.x;
  if (tile < ntiles) stage_patch(tile, 0);
  cp_async_commit();
  for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    const int cur = it & 1;
    const int b = fast_div(tile, fd_bands);
    const int oy0 = (tile - b * bands) * RO;
    const int npx = min(RO, p.OH - oy0) * p.OW;
    const int m0 = (b * p.OH + oy0) * p.OW;                      // first output pixel of the band
    if (tile + (int)gridDim.x < ntiles) stage_patch(tile + gridDim.x, cur ^ 1);
    cp_async_commit();
    cp_async_wait<1>();                      // this tile's input rows have landed
    __syncthreads();
    wide_band<MODE, ROUND>(p, ksteps, reinterpret_cast<const float*>(patch4 + cur * patch_px), bfrag, koff_s, bias_s, npx,
                           m0, fd_ow);
    __syncthreads();                         // patch[cur] may be overwritten from here on
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// edge_narrow: per tile (image b, band of R input rows + halo) compute col[pixel][(ky,kx,c)] = in[pixel][:] . W with
// mma.sync (A fragments straight from global memory as 16-byte loads, K permuted so that a float4 feeds two k-steps),
// park col in shared memory, then every output pixel of the band sums its taps in a fixed order (ky, kx ascending)
// and applies the epilogue.  K = 64.
// ---------------------------------------------------------------------------------------------
constexpr int EN_THREADS = 256;
constexpr int EN_ROWS = 256;        // col rows (input pixels incl. halo) per tile
constexpr int EN_KSTEPS = 8;        // K = 64

// Weights of an edge_narrow pass in mma fragment order: [8 k-steps][NT][32 lanes] float2.
template <int NT, int CI>
__device__ __forceinline__ void narrow_setup(const EdgeNarrowParams& p, float2* bfrag) {
  const int nreal = p.k * p.k * CI;
  for (int idx = threadIdx.x; idx < EN_KSTEPS * NT * 32; idx += EN_THREADS) {
    const int l = idx & 31, nt = (idx >> 5) % NT, ks = idx / (NT * 32);
    const int n = nt * 8 + (l >> 2);
    // logical k slots (t, t+4) of k-step ks are input channels 16*(ks/2) + 4t + 2*(ks%2) + {0, 1}
    const int ch = 16 * (ks >> 1) + 4 * (l & 3) + 2 * (ks & 1);
    float2 b = make_float2(0.f, 0.f);
    if (n < nreal) {
      const int tap = n / CI, c = n - tap * CI;
      b = __ldg(reinterpret_cast<const float2*>(p.w + (size_t)(tap * 4 + c) * p.K + ch));
    }
    bfrag[idx] = b;
  }
}

// col[pixel][(ky,kx,c)] = in[pixel][:] . W for the mt_rows input pixels starting at src; result in shared memory.
template <int NT>
__device__ __forceinline__ void narrow_mma(const float* src, int K, int mt_rows, const float2* bfrag, float* col_s) {
  constexpr int CP = NT * 8 + 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nmt = (mt_rows + 15) >> 4;
  for (int mt = warp; mt < nmt; mt += EN_THREADS / 32) {
    const int ra = mt * 16 + g, rb = ra + 8;
    float4 va[4], vb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      va[j] = ra < mt_rows ? __ldg(reinterpret_cast<const float4*>(src + (size_t)ra * K + 16 * j + 4 * t))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      vb[j] = rb < mt_rows ? __ldg(reinterpret_cast<const float4*>(src + (size_t)rb * K + 16 * j + 4 * t))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float2 be = bfrag[((2 * j) * NT + nt) * 32 + lane];
        mma_tf32(acc[nt], __float_as_uint(va[j].x), __float_as_uint(vb[j].x), __float_as_uint(va[j].y),
                 __float_as_uint(vb[j].y), __float_as_uint(be.x), __float_as_uint(be.y));
        const float2 bo = bfrag[((2 * j + 1) * NT + nt) * 32 + lane];
        mma_tf32(acc[nt], __float_as_uint(va[j].z), __float_as_uint(vb[j].z), __float_as_uint(va[j].w),
                 __float_as_uint(vb[j].w), __float_as_uint(bo.x), __float_as_uint(bo.y));
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      *reinterpret_cast<float2*>(col_s + ra * CP + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<float2*>(col_s + rb * CP + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
    }
  }
}

// col2im + epilogue over output rows [y_lo, y_hi) of image b (margins of the pitched layout are written as zeros): one
// warp per output row, so the ky taps are warp-uniform; only taps of the right parity are visited, in a fixed order.
// The result goes to global memory (unless `out` is null) and, if `patch` is given, into a shared-memory copy of the
// pitched image whose row 0 is image row -patch_row0 (the consumer's staged input, fused pair kernels).
template <int NT, int CI>
__device__ __forceinline__ void narrow_col2im(const EdgeNarrowParams& p, const float* col_s, int b, int iy_lo, int rows_in,
                                              int y_lo, int y_hi, float* out, float4* patch, int patch_row0) {
  constexpr int CP = NT * 8 + 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 bias4 = (p.e.epi == EPI_FWD && p.e.bias) ? __ldg(reinterpret_cast<const float4*>(p.e.bias))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int y = y_lo + warp; y < y_hi; y += EN_THREADS / 32) {
    const int ky0 = (y + p.pad_y) & 1;
    const size_t grow = ((size_t)b * p.OH + y) * p.out_pitch * 4;
    for (int xc = lane; xc < p.out_pitch; xc += 32) {
      const int x = xc - p.out_xoff;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      if (x >= 0 && x < p.OW) {
        float a[CI];
#pragma unroll
        for (int c = 0; c < CI; ++c) a[c] = 0.f;
        const int kx0 = (x + p.pad_x) & 1;
        for (int ky = ky0; ky < p.k; ky += 2) {
          const int lr = ((y + p.pad_y - ky) >> 1) - iy_lo;      // negative (above the image / tile) -> skipped
          if ((unsigned)lr >= (unsigned)rows_in) continue;
          for (int kx = kx0; kx < p.k; kx += 2) {
            const int ix = (x + p.pad_x - kx) >> 1;
            if ((unsigned)ix >= (unsigned)p.IW) continue;
            const float* cp = col_s + (lr * p.IW + ix) * CP + (ky * p.k + kx) * CI;
#pragma unroll
            for (int c = 0; c < CI; ++c) a[c] += cp[c];
          }
        }
        float4 x0 = bias4;
        if (p.e.epi == EPI_BWD) x0 = __ldg(reinterpret_cast<const float4*>(p.e.aux + grow + xc * 4));
        const float xs[3] = {x0.x, x0.y, x0.z};
        float unused;
#pragma unroll
        for (int c = 0; c < CI; ++c) o[c] = epilogue1(p.e, a[c], xs[c], 0.f, &unused);
      }
      const float4 v = make_float4(o[0], o[1], o[2], o[3]);
      if (out) *reinterpret_cast<float4*>(out + grow + xc * 4) = v;
      if (patch) patch[(y + patch_row0) * p.out_pitch + xc] = v;
    }
  }
}

template <int NT, int CI>
__global__ void __launch_bounds__(EN_THREADS) edge_narrow_kernel(const EdgeNarrowParams p, int ntiles) {
  extern __shared__ float2 en_smem[];
  float2* bfrag = en_smem;                                              // [8][NT][32]
  float* col_s = reinterpret_cast<float*>(en_smem + EN_KSTEPS * NT * 32);   // [EN_ROWS][CP]
  narrow_setup<NT, CI>(p, bfrag);
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();                                // weights above are constants; activations from here on
  if (p.e.live) ntiles = min(ntiles, live_images(p.e.live, p.B) * p.bands);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / p.bands;
    const int r0 = (tile - b * p.bands) * p.R;
    const int iy_lo = max(0, r0 - p.halo_lo);
    const int iy_hi = min(p.IH, r0 + p.R + p.halo_hi);
    narrow_mma<NT>(p.in + ((size_t)b * p.IH + iy_lo) * p.IW * p.K, p.K, (iy_hi - iy_lo) * p.IW, bfrag, col_s);
    __syncthreads();
    narrow_col2im<NT, CI>(p, col_s, b, iy_lo, iy_hi - iy_lo, 2 * r0, min(p.OH, 2 * (r0 + p.R)), p.out, nullptr, 0);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// edge_narrow2: the same pass (64 channels -> image-like output, stride-2 transposed type) in GATHER form with the
// accumulators in registers -- no column buffer, no shared-memory col2im, no halo recompute.
//
// Output pixel (2j + py, 2i + px) sums the taps (ky, kx) of matching parity; tap ky reads input row j + dy with
// dy = (py + pad - ky) / 2 in {-1, 0, 1} (pad = 1 for k = 4 and k = 5 under TF SAME), likewise dx.  So for the input
// grid point (j, i) the 4 parity classes x 4 padded channels are 16 GEMM columns and the reduction runs over
// (dy, dx, 64 channels): acc[(j, i)][(py, px, c)] += in[j + dy][i + dx][:] . Wshift[dy][dx][:][(py, px, c)], with
// Wshift zero where a class has no tap at that shift.  n-tile 0 holds the py = 0 classes, n-tile 1 the py = 1 ones,
// so (dy, n-tile) pairs without any tap are skipped at compile time: 15 of 18 (k = 5) / 12 of 18 (k = 4).
//
// A WARP owns a job = (image, band of R input rows) and sweeps the input rows once ("input stationary"): row r feeds
// the accumulator sets of output-row pairs j = r + 1, r, r - 1 (three register sets, rotated by a compile-time phase),
// so every A fragment is loaded once per dx instead of once per (dy, dx).  Rows are staged by the warp itself with
// cp.async into a private double buffer (swizzled 16-byte chunks: conflict-free LDS.128 fragment loads); warps never
// synchronise with each other.  When row r is done the pair j = r - 1 is complete: bias / tanh (or x tanh') and the
// pitched store straight from the mma accumulator layout (a quad writes two adjacent pixels = 32 bytes, a warp
// instruction 256 contiguous bytes).
// ---------------------------------------------------------------------------------------------
constexpr int N2_PX = 34;                      // staged pixels per row: zero | <= 32 data | zero
constexpr int N2_ROW_FLOATS = N2_PX * 64;      // 8704 bytes
constexpr int N2_BFRAG = 9 * EN_KSTEPS * 2 * 32;   // float2 per CTA: [shift][k-step][n-tile][lane]

template <int KS>
__host__ __device__ constexpr bool n2_need(int nt, int dyi) {       // does class row py = nt have a tap at dy = dyi - 1 ?
  return (nt + 3 - 2 * dyi) >= 0 && (nt + 3 - 2 * dyi) < KS;         // ky = py + pad - 2 dy, pad = 1
}

// Four 8x4 TF32 sub-matrices (8 rows of 16 bytes each) -> the A fragment of mma.m16n8k8 in register order: lanes 0-7
// address the rows of (pixels g, k 0-3), lanes 8-15 (pixels g+8, k 0-3), lanes 16-23 (g, k 4-7), lanes 24-31 (g+8, k 4-7).
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void n2_setup(const EdgeNarrowParams& p, float2* bfrag) {
  for (int idx = threadIdx.x; idx < N2_BFRAG; idx += blockDim.x) {
    const int l = idx & 31, nt = (idx >> 5) & 1, ks = (idx >> 6) & 7, sh = idx >> 9;
    const int dy = sh / 3 - 1, dx = sh % 3 - 1;
    const int g = l >> 2, t = l & 3;
    const int py = nt, px = g >> 2, c = g & 3;
    const int ky = py + p.pad_y - 2 * dy, kx = px + p.pad_x - 2 * dx;
    float2 b = make_float2(0.f, 0.f);                                // b0 = W[k = t][n = g], b1 = W[k = t + 4][n = g]
    if (c < p.cimg && ky >= 0 && ky < p.k && kx >= 0 && kx < p.k) {
      const float* wr = p.w + (size_t)((ky * p.k + kx) * 4 + c) * p.K + 8 * ks + t;
      b = make_float2(__ldg(wr), __ldg(wr + 4));
    }
    bfrag[idx] = b;
  }
}

template <int KS, int MT>
__global__ void __launch_bounds__(256, 1) edge_narrow2_kernel(const EdgeNarrowParams p, int njobs) {
  extern __shared__ float4 n2_smem4[];
  float2* bfrag = reinterpret_cast<float2*>(n2_smem4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float* rowbuf = reinterpret_cast<float*>(bfrag + N2_BFRAG) + (size_t)warp * 2 * N2_ROW_FLOATS;   // [2][34][64]
  n2_setup(p, bfrag);
  for (int i = lane; i < 2 * N2_ROW_FLOATS / 4; i += 32) reinterpret_cast<float4*>(rowbuf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  if (p.e.live) njobs = min(njobs, live_images(p.e.live, p.B) * p.bands);
  const uint32_t rowbuf_u32 = smem_u32(rowbuf);
  const float4 bias4 = (p.e.epi == EPI_FWD && p.e.bias) ? __ldg(reinterpret_cast<const float4*>(p.e.bias))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
  const float bias0 = (t & 1) ? bias4.z : bias4.x, bias1 = (t & 1) ? bias4.w : bias4.y;
  const int c0 = (t & 1) * 2;                       // image channels this lane holds: c0, c0 + 1
  const bool bwd = p.e.epi == EPI_BWD;
  // ldmatrix row address of this lane: stored pixel (lane & 7) + 8 * bit 3 (+ 16 m + dx), 16-byte chunk 2 ks + bit 4
  const int lm_px = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int lm_half = lane >> 4;

  for (int job = blockIdx.x * nwarps + warp; job < njobs; job += gridDim.x * nwarps) {
    const int b = job / p.bands;
    const int j0 = (job - b * p.bands) * p.R;
    const int j1 = min(p.IH, j0 + p.R);              // output-row pairs [j0, j1)
    const float* src_img = p.in + (size_t)b * p.IH * p.IW * 64;
    // three accumulator sets: output-row pairs j = r + 1, r, r - 1 of the input row r being swept
    float acc[3][MT][2][4];
#pragma unroll
    for (int s3 = 0; s3 < 3; ++s3)
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[s3][m][nt][q] = 0.f;

    // stage input row r into buffer `buf`: 16-byte chunk c of stored pixel x lands at chunk (c & 8) | ((c ^ x) & 7), so
    // that the eight 16-byte rows of an ldmatrix sub-matrix (8 consecutive pixels, same chunk) hit distinct banks
    auto stage_row = [&](int r, int buf) {
      if ((unsigned)r < (unsigned)p.IH) {
        const float* src = src_img + (size_t)r * p.IW * 64;
        const uint32_t dst = rowbuf_u32 + buf * (N2_ROW_FLOATS * 4);
        for (int q = lane; q < p.IW * 16; q += 32) {
          const int px = (q >> 4) + 1, ck = q & 15;
          cp_async_16(dst + px * 256 + (((ck & 8) | ((ck ^ px) & 7)) << 4), src + q * 4, 16u);
        }
      }
      cp_async_commit();
    };

    // epilogue of the output-row pair j held in accumulator set `a` (stale sets are zeroed, never stored)
    auto finalize = [&](int j, bool store, float (&a)[MT][2][4], const float2 (&aux)[MT][2][2]) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float* orow = p.out + ((size_t)b * p.OH + 2 * j + nt) * p.out_pitch * 4;
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = 16 * m + 8 * h + g;
            if (store && i < p.IW) {
              const float x0a = bwd ? aux[m][nt][h].x : bias0, x0b = bwd ? aux[m][nt][h].y : bias1;
              float unused;
              float2 o;
              o.x = c0 < p.cimg ? epilogue1(p.e, a[m][nt][2 * h], x0a, 0.f, &unused) : 0.f;
              o.y = c0 + 1 < p.cimg ? epilogue1(p.e, a[m][nt][2 * h + 1], x0b, 0.f, &unused) : 0.f;
              *reinterpret_cast<float2*>(orow + (size_t)(2 * i + (t >> 1) + p.out_xoff) * 4 + c0) = o;
            }
            a[m][nt][2 * h] = 0.f;
            a[m][nt][2 * h + 1] = 0.f;
          }
        // margins of the pitched layout are zeros
        const int nmargin = p.out_pitch - p.OW;
        if (store && lane < nmargin) {
          const int col = lane < p.out_xoff ? lane : p.OW + lane;
          *reinterpret_cast<float4*>(orow + (size_t)col * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };

    // one input row: PH = (r - (j0 - 1)) % 3 selects which register set belongs to which output-row pair.  All three
    // targets are always accumulated (no branches around the MMAs: 10 independent accumulators per k-step hide the
    // mma.sync latency with only two warps per scheduler); the set of a pair outside the band collects garbage and is
    // zeroed, never stored, when its turn to be finalised comes.  The k-step loop is deliberately NOT fully unrolled:
    // the fully unrolled kernel was 140 KB of code and spent 20 % of its issue slots on instruction-cache misses.
    auto row_step = [&](int r, int buf, auto ph_tag) {
      constexpr int PH = decltype(ph_tag)::value;
      cp_async_wait<0>();
      __syncwarp();                                   // row r has landed; every lane is done with the other buffer
      if (r + 1 <= j1) stage_row(r + 1, buf ^ 1);
      // derivative operand of the pair that completes after this row (prefetched: its latency hides behind the MMAs)
      float2 aux[MT][2][2];
      const int jf = r - 1;
      const bool fin = jf >= j0 && jf < j1;
      if (bwd && fin) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const float* arow = p.e.aux + ((size_t)b * p.OH + 2 * jf + nt) * p.out_pitch * 4;
#pragma unroll
          for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int i = 16 * m + 8 * h + g;
              aux[m][nt][h] = i < p.IW ? __ldg(reinterpret_cast<const float2*>(arow + (size_t)(2 * i + (t >> 1) + p.out_xoff) * 4 + c0))
                                       : make_float2(0.f, 0.f);
            }
        }
      }
      if ((unsigned)r < (unsigned)p.IH) {
        const uint32_t rb = rowbuf_u32 + buf * (N2_ROW_FLOATS * 4);
        constexpr int kSetOf[3] = {(PH + 1) % 3, PH, (PH + 2) % 3};    // dy = -1, 0, +1  ->  pair j = r + 1, r, r - 1
#pragma unroll
        for (int dxi = 0; dxi < 3; ++dxi) {
          const int swz = (lm_px + dxi) & 7;                            // low bits of the stored pixel index
          const uint32_t abase = rb + (uint32_t)(lm_px + dxi) * 256u;
          const float2* bsh = bfrag + (size_t)dxi * (EN_KSTEPS * 2 * 32) + lane;   // shift (dy, dx) at + dyi * 3 * 512
#pragma unroll 2
          for (int ks = 0; ks < EN_KSTEPS; ++ks) {
            const int ck = 2 * ks + lm_half;
            const uint32_t coff = (uint32_t)(((ck & 8) | ((ck ^ swz) & 7)) << 4);
            uint32_t af[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m) ldmatrix_x4(af[m], abase + (uint32_t)m * (16u * 256u) + coff);
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) {
                if (!n2_need<KS>(nt, dyi)) continue;                    // compile time
                const float2 bq = bsh[(size_t)dyi * (3 * EN_KSTEPS * 2 * 32) + (ks * 2 + nt) * 32];
#pragma unroll
                for (int m = 0; m < MT; ++m)
                  mma_tf32(acc[kSetOf[dyi]][m][nt], af[m][0], af[m][1], af[m][2], af[m][3], __float_as_uint(bq.x),
                           __float_as_uint(bq.y));
              }
          }
        }
      }
      finalize(jf, fin, acc[(PH + 2) % 3], aux);      // store if the pair is in the band; zero the set either way
    };

    __syncwarp();                                     // previous job's reads of the row buffers are done
    stage_row(j0 - 1, 0);
    int buf = 0;
    for (int r = j0 - 1; r <= j1; r += 3) {
      row_step(r, buf, std::integral_constant<int, 0>());
      buf ^= 1;
      if (r + 1 > j1) break;
      row_step(r + 1, buf, std::integral_constant<int, 1>());
      buf ^= 1;
      if (r + 2 > j1) break;
      row_step(r + 2, buf, std::integral_constant<int, 2>());
      buf ^= 1;
    }
    cp_async_wait<0>();
  }
}

template <int KS, int MT>
int launch_narrow2_inst(const EdgeNarrowParams& p, int warps, cudaStream_t st) {
  const size_t smem = (size_t)N2_BFRAG * sizeof(float2) + (size_t)warps * 2 * N2_ROW_FLOATS * sizeof(float);
  static DynSmemCache smem_cache;
  cudaError_t e = ensure_dyn_smem(edge_narrow2_kernel<KS, MT>, smem, smem_cache);
  if (e != cudaSuccess) return set_error(CGS_ERR_CUDA, "cudaFuncSetAttribute(edge_narrow2): %s", cudaGetErrorString(e));
  const long long jobs = (long long)p.B * p.bands;
  if (jobs >= (1ll << 31)) return set_error(CGS_ERR_UNSUPPORTED, "batch too large; split the batch");
  long long grid = (jobs + warps - 1) / warps;
  if (grid > device_num_sms()) grid = device_num_sms();
  cudaError_t le = launch_pdl(edge_narrow2_kernel<KS, MT>, dim3((unsigned)grid), dim3(warps * 32), smem, st, p, (int)jobs);
  count_launch();
  if (le != cudaSuccess) return set_error(CGS_ERR_CUDA, "edge_narrow2_kernel: %s", cudaGetErrorString(le));
  return check_launch("edge_narrow2_kernel");
}

// Rows per job and warps per CTA: as few bands as possible (a band re-reads two halo rows) while the jobs still fill
// the machine in whole rounds.  Pure function of (batch, image height, SM count): never of the data.
void narrow2_geometry(EdgeNarrowParams& p, int& warps) {
  const int sms = device_num_sms();
  double best = 1e30;
  int best_bands = 1, best_w = 8;
  for (int bands = 1; bands <= p.IH / 4 || bands == 1; bands *= 2) {
    const int R = (p.IH + bands - 1) / bands;
    const int nb = (p.IH + R - 1) / R;
    const double per_job = 5.0 * R + 3.0 + 0.5 * (R + 2);           // MMA units of a band + staging of its rows
    for (int w = 8; w >= 5; --w) {
      const long long jobs = (long long)p.B * nb;
      const long long slots = (long long)sms * w;
      const long long rounds = (jobs + slots - 1) / slots;
      const double cost = (double)rounds * per_job * (1.0 + 0.02 * (8 - w));   // a round takes a job's time
      if (cost < best) { best = cost; best_bands = nb; best_w = w; p.R = R; }
    }
  }
  p.bands = best_bands;
  p.R = (p.IH + best_bands - 1) / best_bands;
  warps = best_w;
}

// ---------------------------------------------------------------------------------------------
// edge_pair: the narrow pass of one layer followed by the wide pass of the next on the same image, in one kernel:
// forward   G's last deconv (-> image, tanh) + D's first conv (-> 64 channels, LeakyReLU)
// backward  D's first conv data-gradient (x tanh') + G's last deconv data-gradient (x ReLU' / policy step)
// The image (or its gradient) goes from the col2im of the first pass straight into the staged-input buffer of the
// second; the forward image is also written to global memory (best-of-K keep, derivative operand), its gradient is
// not.  One image per tile: used when a whole image fits one tile of both passes (MNIST-sized nets).
// ---------------------------------------------------------------------------------------------
template <int NT, int CI, int MODE, bool ROUND>
__global__ void __launch_bounds__(EN_THREADS, 4) edge_pair_kernel(const EdgeNarrowParams pn, const EdgeWideParams pw,
                                                                  int ksteps_w, int store_image,
                                                                  unsigned long long fd_ow) {
  static_assert(EN_THREADS == EW_THREADS, "both passes use the same CTA shape");
  extern __shared__ float4 ep_smem4[];
  const int PR = 2 * pw.OH + pw.k - 2;                                   // staged rows of the wide pass (whole image)
  float2* bfrag_n = reinterpret_cast<float2*>(ep_smem4);                 // [8][NT][32]
  float2* bfrag_w = bfrag_n + EN_KSTEPS * NT * 32;                       // [ksteps_w][8][32]
  float4* patch4 = reinterpret_cast<float4*>(bfrag_w + ksteps_w * EW_NT * 32);   // [PR][pitch]
  float* col_s = reinterpret_cast<float*>(patch4 + PR * pw.pitch);       // [IH*IW rounded to 16][CP]
  __shared__ int koff_s[EW_KMAX];
  __shared__ float bias_s[EW_NT * 8];
  narrow_setup<NT, CI>(pn, bfrag_n);
  wide_setup(pw, ksteps_w, bfrag_w, koff_s, bias_s);
  // rows of the staged image outside the image stay zero for the whole kernel
  for (int idx = threadIdx.x; idx < PR * pw.pitch; idx += EN_THREADS) {
    const int iy = idx / pw.pitch - pw.pad_y;
    if ((unsigned)iy >= (unsigned)pw.IH) patch4[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  const int npx = pw.OH * pw.OW;
  const int B_live = live_images(pw.e.live, pn.B);
  for (int b = blockIdx.x; b < B_live; b += gridDim.x) {
    narrow_mma<NT>(pn.in + (size_t)b * pn.IH * pn.IW * pn.K, pn.K, pn.IH * pn.IW, bfrag_n, col_s);
    __syncthreads();
    narrow_col2im<NT, CI>(pn, col_s, b, 0, pn.IH, 0, pn.OH, store_image ? pn.out : nullptr, patch4, pw.pad_y);
    __syncthreads();
    wide_band<MODE, ROUND>(pw, ksteps_w, reinterpret_cast<const float*>(patch4), bfrag_w, koff_s, bias_s, npx, b * npx, fd_ow);
    __syncthreads();
  }
}

int num_sms() { return device_num_sms(); }

template <int NT, int CI>
int launch_narrow_nt(const EdgeNarrowParams& p, cudaStream_t st) {
  constexpr int CP = NT * 8 + 4;
  constexpr size_t smem = (size_t)EN_KSTEPS * NT * 32 * sizeof(float2) + (size_t)EN_ROWS * CP * sizeof(float);
  int ctas_per_sm = 0;
  {
    static DynSmemCache smem_cache;
    cudaError_t e = ensure_dyn_smem(edge_narrow_kernel<NT, CI>, smem, smem_cache);
    if (e != cudaSuccess) return set_error(CGS_ERR_CUDA, "cudaFuncSetAttribute(edge_narrow): %s", cudaGetErrorString(e));
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, edge_narrow_kernel<NT, CI>, EN_THREADS, smem);
    if (e != cudaSuccess || ctas_per_sm < 1) return set_error(CGS_ERR_CUDA, "edge_narrow occupancy query failed");
  }
  const long long tiles = (long long)p.B * p.bands;
  if (tiles >= (1ll << 31)) return set_error(CGS_ERR_UNSUPPORTED, "batch too large; split the batch");
  long long grid = (long long)num_sms() * ctas_per_sm;
  if (grid > tiles) grid = tiles;
  cudaError_t le = launch_pdl(edge_narrow_kernel<NT, CI>, dim3((unsigned)grid), dim3(EN_THREADS), smem, st, p, (int)tiles);
  count_launch();
  if (le != cudaSuccess) return set_error(CGS_ERR_CUDA, "edge_narrow_kernel: %s", cudaGetErrorString(le));
  return check_launch("edge_narrow_kernel");
}

}  // namespace

bool edge_wide_supported(int N, int k, int cimg) {
  return N == 64 && k >= 1 && k <= 5 && cimg >= 1 && cimg <= 4;
}

bool edge_narrow_supported(int K, int k, int cimg, int IW) {
  return K == 64 && (k == 4 || k == 5) && (cimg == 1 || cimg == 3) && IW <= EN_ROWS / 4;
}

template <int MODE, bool ROUND>
int launch_wide_mode(const EdgeWideParams& p, int B, cudaStream_t st) {
  const int ksteps = (p.k * p.k * p.cimg + 7) / 8;
  const int RO = (p.OH * p.OW <= EW_TILE_PX) ? p.OH : EW_TILE_PX / p.OW;   // whole image per tile when it fits
  const int bands = (p.OH + RO - 1) / RO;
  const long long tiles_ll = (long long)B * bands;
  if (tiles_ll >= (1ll << 31)) return set_error(CGS_ERR_UNSUPPORTED, "batch too large; split the batch");
  const int tiles = (int)tiles_ll;
  const int PR = 2 * RO + p.k - 2;
  size_t smem = (size_t)ksteps * EW_NT * 32 * sizeof(float2) + (size_t)2 * PR * p.pitch * sizeof(float4);
  if (smem > 200 * 1024) return set_error(CGS_ERR_UNSUPPORTED, "edge_wide: image band does not fit shared memory");
  {
    static DynSmemCache smem_cache;
    cudaError_t e = ensure_dyn_smem(edge_wide_kernel<MODE, ROUND>, smem, smem_cache);
    if (e != cudaSuccess) return set_error(CGS_ERR_CUDA, "cudaFuncSetAttribute(edge_wide): %s", cudaGetErrorString(e));
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, edge_wide_kernel<MODE, ROUND>, EW_THREADS, smem);
  if (e != cudaSuccess || per_sm < 1) return set_error(CGS_ERR_CUDA, "edge_wide occupancy query failed");
  long long grid = (long long)num_sms() * per_sm;
  if (grid > tiles) grid = tiles;
  cudaError_t le = launch_pdl(edge_wide_kernel<MODE, ROUND>, dim3((unsigned)grid), dim3(EW_THREADS), smem, st, p, ksteps, tiles, RO,
                              bands, fast_div_magic((unsigned)bands), fast_div_magic((unsigned)p.OW));
  count_launch();
  if (le != cudaSuccess) return set_error(CGS_ERR_CUDA, "edge_wide_kernel: %s", cudaGetErrorString(le));
  return check_launch("edge_wide_kernel");
}

int launch_edge_wide(const EdgeWideParams& p, cudaStream_t st) {
  if (p.M <= 0) return CGS_OK;
  if (p.M * p.ON >= (1ll << 31)) return set_error(CGS_ERR_UNSUPPORTED, "batch too large for 32-bit indexing; split the batch");
  if (p.OW > EW_TILE_PX) return set_error(CGS_ERR_UNSUPPORTED, "edge_wide: output row wider than a tile");
  const int B = (int)(p.M / ((long long)p.OH * p.OW));
  const bool rnd = p.e.round_out != 0;
  if (p.e.epi == EPI_FWD && !p.e.act_tanh)
    return rnd ? launch_wide_mode<EM_FWD_SLOPE, true>(p, B, st) : launch_wide_mode<EM_FWD_SLOPE, false>(p, B, st);
  if (p.e.epi == EPI_BWD && !p.e.act_tanh)
    return rnd ? launch_wide_mode<EM_BWD_SLOPE, true>(p, B, st) : launch_wide_mode<EM_BWD_SLOPE, false>(p, B, st);
  if (p.e.epi == EPI_UPDATE) return launch_wide_mode<EM_UPDATE, false>(p, B, st);
  return launch_wide_mode<EM_GENERIC, false>(p, B, st);
}

namespace {
int narrow_geometry(EdgeNarrowParams& p) {
  // rows of the input needed above / below a band of R input rows (its 2R output rows gather taps
  // iy = (y + pad_y - ky) / 2): brute force over a sample band, the answer does not depend on R or r0
  int lo = 0, hi = 0;
  {
    const int Rs = 8;
    for (int y = 0; y < 2 * Rs; ++y)
      for (int ky = 0; ky < p.k; ++ky) {
        const int ty = y + p.pad_y - ky;
        if (ty & 1) continue;
        const int iy = ty >= 0 ? ty / 2 : -((-ty) / 2);
        if (-iy > lo) lo = -iy;
        if (iy - (Rs - 1) > hi) hi = iy - (Rs - 1);
      }
  }
  p.halo_lo = lo;
  p.halo_hi = hi;
  const int rows_max = EN_ROWS / p.IW;
  if (p.IH <= rows_max) {
    p.R = p.IH;
  } else {
    p.R = rows_max - lo - hi;
    if (p.R < 1) return set_error(CGS_ERR_UNSUPPORTED, "edge_narrow: image row too wide for the tile");
  }
  p.bands = (p.IH + p.R - 1) / p.R;
  return CGS_OK;
}

template <int NT, int CI, int MODE, bool ROUND>
int launch_pair_inst(const EdgeNarrowParams& pn, const EdgeWideParams& pw, int store_image, cudaStream_t st) {
  constexpr int CP = NT * 8 + 4;
  const int ksteps_w = (pw.k * pw.k * pw.cimg + 7) / 8;
  const int PR = 2 * pw.OH + pw.k - 2;
  const int col_rows = ((pn.IH * pn.IW + 15) / 16) * 16;
  const size_t smem = (size_t)EN_KSTEPS * NT * 32 * sizeof(float2) + (size_t)ksteps_w * EW_NT * 32 * sizeof(float2) +
                      (size_t)PR * pw.pitch * sizeof(float4) + (size_t)col_rows * CP * sizeof(float);
  {
    static DynSmemCache smem_cache;
    cudaError_t e = ensure_dyn_smem(edge_pair_kernel<NT, CI, MODE, ROUND>, smem, smem_cache);
    if (e != cudaSuccess) return set_error(CGS_ERR_CUDA, "cudaFuncSetAttribute(edge_pair): %s", cudaGetErrorString(e));
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, edge_pair_kernel<NT, CI, MODE, ROUND>, EN_THREADS, smem);
  if (e != cudaSuccess || per_sm < 1) return set_error(CGS_ERR_CUDA, "edge_pair occupancy query failed");
  long long grid = (long long)num_sms() * per_sm;
  if (grid > pn.B) grid = pn.B;
  e = launch_pdl(edge_pair_kernel<NT, CI, MODE, ROUND>, dim3((unsigned)grid), dim3(EN_THREADS), smem, st, pn, pw, ksteps_w,
                 store_image, fast_div_magic((unsigned)pw.OW));
  count_launch();
  if (e != cudaSuccess) return set_error(CGS_ERR_CUDA, "edge_pair_kernel: %s", cudaGetErrorString(e));
  return check_launch("edge_pair_kernel");
}

template <int NT, int CI>
int launch_pair_mode(const EdgeNarrowParams& pn, const EdgeWideParams& pw, int store_image, cudaStream_t st) {
  const bool rnd = pw.e.round_out != 0;
  if (pw.e.epi == EPI_FWD && !pw.e.act_tanh)
    return rnd ? launch_pair_inst<NT, CI, EM_FWD_SLOPE, true>(pn, pw, store_image, st)
               : launch_pair_inst<NT, CI, EM_FWD_SLOPE, false>(pn, pw, store_image, st);
  if (pw.e.epi == EPI_BWD && !pw.e.act_tanh)
    return rnd ? launch_pair_inst<NT, CI, EM_BWD_SLOPE, true>(pn, pw, store_image, st)
               : launch_pair_inst<NT, CI, EM_BWD_SLOPE, false>(pn, pw, store_image, st);
  if (pw.e.epi == EPI_UPDATE) return launch_pair_inst<NT, CI, EM_UPDATE, false>(pn, pw, store_image, st);
  return launch_pair_inst<NT, CI, EM_GENERIC, false>(pn, pw, store_image, st);
}
}  // namespace

// One image per tile in both passes, single-channel image, same pitched layout on both sides; the shared-memory
// footprint (column buffer + staged image + two weight tiles) must leave room for three CTAs per SM.
bool edge_pair_supported(const EdgeNarrowParams& pn, const EdgeWideParams& pw) {
  if (!edge_narrow_supported(pn.K, pn.k, pn.cimg, pn.IW) || !edge_wide_supported(pw.N, pw.k, pw.cimg)) return false;
  if (pn.cimg != 1 || pw.cimg != 1) return false;
  if (pn.IH * pn.IW > EN_ROWS || pw.OH * pw.OW > EW_TILE_PX) return false;
  if (pn.OH != pw.IH || pn.out_pitch != pw.pitch || pn.out_xoff != pw.xoff || pw.ON != pw.N) return false;
  return true;
}

int launch_edge_pair(EdgeNarrowParams pn, const EdgeWideParams& pw, int store_image, cudaStream_t st) {
  if (pn.B <= 0) return CGS_OK;
  if ((long long)pn.B * pw.OH * pw.OW * pw.ON >= (1ll << 31)) return set_error(CGS_ERR_UNSUPPORTED, "batch too large for 32-bit indexing; split the batch");
  if (int rc = narrow_geometry(pn)) return rc;
  if (pn.k == 4) return launch_pair_mode<2, 1>(pn, pw, store_image, st);
  if (pn.k == 5) return launch_pair_mode<4, 1>(pn, pw, store_image, st);
  return set_error(CGS_ERR_UNSUPPORTED, "edge_pair: unsupported kernel size");
}

int launch_edge_narrow(EdgeNarrowParams p, cudaStream_t st) {
  if (p.B <= 0) return CGS_OK;
  // tcgen05 form with the input patch resident in shared memory (edge_tc.cu); CGS_DEBUG bit 2097152 keeps mma.sync
  if (!(debug_flags() & 2097152) && p.IW >= 16 && edge_narrow_tc_supported(p)) return launch_edge_narrow_tc(p, st);
  // gather form with register accumulators (edge_narrow2); CGS_DEBUG bit 262144 keeps the column-buffer kernel
  // (rows narrower than 16 pixels leave half of every 16-row MMA tile empty: the MNIST-sized nets keep the other form)
  if (!(debug_flags() & 262144) && p.K == 64 && p.pad_y == 1 && p.pad_x == 1 && p.IW >= 16 && p.IW <= 32 && (p.k == 4 || p.k == 5) &&
      p.OW == 2 * p.IW && p.OH == 2 * p.IH && p.e.epi != EPI_UPDATE) {
    int warps = 8;
    narrow2_geometry(p, warps);
    if (p.k == 5) return p.IW > 16 ? launch_narrow2_inst<5, 2>(p, warps, st) : launch_narrow2_inst<5, 1>(p, warps, st);
    return p.IW > 16 ? launch_narrow2_inst<4, 2>(p, warps, st) : launch_narrow2_inst<4, 1>(p, warps, st);
  }
  if (int rc = narrow_geometry(p)) return rc;
  if (p.k == 4 && p.cimg == 1) return launch_narrow_nt<2, 1>(p, st);
  if (p.k == 5 && p.cimg == 1) return launch_narrow_nt<4, 1>(p, st);
  if (p.k == 4 && p.cimg == 3) return launch_narrow_nt<6, 3>(p, st);
  if (p.k == 5 && p.cimg == 3) return launch_narrow_nt<10, 3>(p, st);
  return set_error(CGS_ERR_UNSUPPORTED, "edge_narrow: unsupported k / channel combination");
}

}  // namespace cgs
