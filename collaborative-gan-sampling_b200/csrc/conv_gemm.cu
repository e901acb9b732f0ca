// tcgen05 / TMEM implicit-GEMM kernel (TF32 inputs, FP32 accumulate) + an FP32 SIMT twin used by the
// parity tests and as the exact-FP32 mode.  See conv_gemm.cuh for the contraction it computes.
//
// CTA = 10 warps, persistent over output tiles (128 rows x BN channels):
//   warps 0-3  epilogue: tcgen05.ld the accumulator (TMEM lane = row), fused bias/activation/derivative/
//              momentum update, vectorised global stores
//   warps 4-7  A producers: cp.async gather of 128 rows x 128 B per K block into SWIZZLE_128B smem
//   warp  8    MMA issuer (one elected lane): 4 x tcgen05.mma.kind::tf32 (K=8 each) per K block
//   warp  9    TMA producer for the weight tile (BN rows x 128 B, SWIZZLE_128B)
// Pipelines: full/empty mbarriers per smem stage, tmem_full/tmem_empty per accumulator buffer
// (2 x BN TMEM columns, so the epilogue of tile i overlaps the main loop of tile i+1).
#include "conv_gemm.cuh"
#include "ptx.cuh"
#include "common.h"

#include <cuda.h>

namespace cgs {

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                       // floats per K block = one 128-byte swizzle row
constexpr int kProducerThreads = 128;
constexpr int kThreads = 320;
constexpr int A_STAGE_BYTES = BM * BK * 4;   // 16 KB

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int LAG = STAGES / 2;     // cp.async groups kept in flight per producer thread
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct RowInfo {
  int base;   // element offset of pixel (j*S, i*S) of image b; -1 if the row is past M
  int y0, x0;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ ConvGemmParams p, const __grid_constant__ CUtensorMap tmap_w) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned stage bases
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full_bar = bars + 2 * C::STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], kProducerThreads + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
  if (warp == 9 && lane == 0) tma_prefetch_desc(&tmap_w);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int tiles_per_class = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_per_class * p.nclasses;

  if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ A producers
    const int pw = warp - 4;
    const int chunk = lane & 7;           // 16-byte chunk inside the 128-byte row
    const int rsub = lane >> 3;           // 0..3
    const uint32_t smem_a_u32 = smem_u32(smem_a);
    const bool pixel_mode = (p.cblocks == 0);
    uint32_t it_global = 0;               // K blocks issued by this thread (== commit groups)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ci = tile / tiles_per_class;
      const int rem = tile - ci * tiles_per_class;
      const int m_tile = rem / p.n_tiles;
      const GemmClass& gc = p.cls[ci];
      RowInfo ri[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = pw * 32 + it * 4 + rsub;
        const int m = m_tile * BM + r;
        if (m < p.M) {
          const int per_img = p.MH * p.MW;
          const int b = m / per_img;
          const int q = m - b * per_img;
          const int j = q / p.MW;
          const int i = q - j * p.MW;
          ri[it].y0 = j * p.S;
          ri[it].x0 = i * p.S;
          ri[it].base = ((b * p.IH + ri[it].y0) * p.IW + ri[it].x0) * p.Cs;
        } else {
          ri[it].base = -1;
          ri[it].y0 = 0;
          ri[it].x0 = 0;
        }
      }
      for (int kb = 0; kb < gc.nkb; ++kb, ++it_global) {
        const int s = it_global % C::STAGES;
        const uint32_t ph = (it_global / C::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        int dy, dx, coff;
        bool tap_ok = true;
        if (pixel_mode) {
          const int t = kb * 8 + chunk;   // one tap (4 channels = 16 B) per chunk
          tap_ok = t < gc.ntaps;
          dy = tap_ok ? gc.dy[t] : 0;
          dx = tap_ok ? gc.dx[t] : 0;
          coff = 0;
        } else {
          const int t = kb / p.cblocks;
          const int cb = kb - t * p.cblocks;
          dy = gc.dy[t];
          dx = gc.dx[t];
          coff = cb * BK + chunk * 4;
        }
        const int tap_off = (dy * p.IW + dx) * p.Cs + coff;
        const uint32_t stage_base = smem_a_u32 + s * A_STAGE_BYTES;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = pw * 32 + it * 4 + rsub;
          const bool ok = tap_ok && ri[it].base >= 0 && (unsigned)(ri[it].y0 + dy) < (unsigned)p.IH &&
                          (unsigned)(ri[it].x0 + dx) < (unsigned)p.IW;
          const float* src = ok ? p.in + (ri[it].base + tap_off) : p.in;
          const uint32_t dst = stage_base + r * 128 + ((chunk ^ (r & 7)) << 4);
          cp_async_16(dst, src, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (it_global >= (uint32_t)C::LAG) {
          cp_async_wait<C::LAG>();
          fence_proxy_async_smem();
          mbar_arrive(&full_bar[(it_global - C::LAG) % C::STAGES]);
        }
      }
    }
    // drain: the last min(LAG, it_global) groups have not been published yet
    cp_async_wait<0>();
    fence_proxy_async_smem();
    const uint32_t pending = it_global < (uint32_t)C::LAG ? it_global : (uint32_t)C::LAG;
    for (uint32_t g = it_global - pending; g < it_global; ++g) mbar_arrive(&full_bar[g % C::STAGES]);
  } else if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer (weights)
    if (lane == 0) {
      const uint32_t smem_b_u32 = smem_u32(smem_b);
      uint32_t it_global = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ci = tile / tiles_per_class;
        const int rem = tile - ci * tiles_per_class;
        const int n_tile = rem % p.n_tiles;
        const GemmClass& gc = p.cls[ci];
        for (int kb = 0; kb < gc.nkb; ++kb, ++it_global) {
          const int s = it_global % C::STAGES;
          const uint32_t ph = (it_global / C::STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], C::B_STAGE_BYTES);
          tma_load_2d(smem_b_u32 + s * C::B_STAGE_BYTES, &tmap_w, &full_bar[s], gc.k0 + kb * BK, n_tile * BN);
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
      const uint32_t smem_a_u32 = smem_u32(smem_a);
      const uint32_t smem_b_u32 = smem_u32(smem_b);
      uint32_t it_global = 0;
      uint32_t tile_count = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_count) {
        const int ci = tile / tiles_per_class;
        const int nkb = p.cls[ci].nkb;
        const uint32_t acc = tile_count & 1;
        const uint32_t acc_ph = (tile_count >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it_global) {
          const int s = it_global % C::STAGES;
          const uint32_t ph = (it_global / C::STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_a_u32 + s * A_STAGE_BYTES);
          const uint64_t db = make_smem_desc_sw128(smem_b_u32 + s * C::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            // +32 bytes per K=8 step inside the 128-byte swizzle row (address field is in 16-byte units)
            umma_tf32_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);          // smem stage reusable once these MMAs have read it
        }
        umma_commit(&tmem_full_bar[acc]);      // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3)
    uint32_t tile_count = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_count) {
      const int ci = tile / tiles_per_class;
      const int rem = tile - ci * tiles_per_class;
      const int m_tile = rem / p.n_tiles;
      const int n_tile = rem - m_tile * p.n_tiles;
      const GemmClass& gc = p.cls[ci];
      const uint32_t acc = tile_count & 1;
      const uint32_t acc_ph = (tile_count >> 1) & 1;
      const int r = warp * 32 + lane;
      const int m = m_tile * BM + r;
      size_t row_off = 0;
      const bool row_ok = m < p.M;
      if (row_ok) {
        const int per_img = p.MH * p.MW;
        const int b = m / per_img;
        const int q = m - b * per_img;
        const int j = q / p.MW;
        const int i = q - j * p.MW;
        row_off = ((size_t)(b * p.OH + j * p.os + gc.oy0) * p.OW + (i * p.os + gc.ox0)) * p.ON;
      }
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(warp * 32) << 16);
      constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t v[CH];
        if constexpr (CH == 32) tmem_ld_32x32b_x32(taddr + c0, v); else tmem_ld_32x32b_x16(taddr + c0, v);
        tmem_ld_wait();
        const int nbase = n_tile * BN + c0;
        if (row_ok && nbase < p.ON) {
#pragma unroll
          for (int q4 = 0; q4 < CH; q4 += 4) {
            const int n = nbase + q4;
            if (n < p.ON) {          // ON is a multiple of 4
              float4 o;
              o.x = epilogue_value(p, row_off + n + 0, n + 0, __uint_as_float(v[q4 + 0]));
              o.y = epilogue_value(p, row_off + n + 1, n + 1, __uint_as_float(v[q4 + 1]));
              o.z = epilogue_value(p, row_off + n + 2, n + 2, __uint_as_float(v[q4 + 2]));
              o.w = epilogue_value(p, row_off + n + 3, n + 3, __uint_as_float(v[q4 + 3]));
              *reinterpret_cast<float4*>(p.out + row_off + n) = o;
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// FP32 SIMT twin: one thread per (row, 4 output channels); same parameters, same epilogue.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_gemm_simt_kernel(const __grid_constant__ ConvGemmParams p, const float* __restrict__ w, int w_cols) {
  const int ngroups = p.ON / 4;
  const long long total = (long long)p.nclasses * p.M * ngroups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ng = (int)(idx % ngroups);
    const long long t = idx / ngroups;
    const int m = (int)(t % p.M);
    const int ci = (int)(t / p.M);
    const GemmClass& gc = p.cls[ci];
    const int per_img = p.MH * p.MW;
    const int b = m / per_img;
    const int q = m - b * per_img;
    const int j = q / p.MW;
    const int i = q - j * p.MW;
    const int cin = p.cblocks ? p.cblocks * BK : 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tp = 0; tp < gc.ntaps; ++tp) {
      const int y = j * p.S + gc.dy[tp];
      const int x = i * p.S + gc.dx[tp];
      if ((unsigned)y >= (unsigned)p.IH || (unsigned)x >= (unsigned)p.IW) continue;
      const float* src = p.in + ((size_t)(b * p.IH + y) * p.IW + x) * p.Cs;
      for (int c = 0; c < cin; ++c) {
        const float a = __ldg(src + c);
        const int k = gc.k0 + tp * cin + c;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = ng * 4 + u;
          if (n < p.N) acc[u] = fmaf(a, __ldg(w + (size_t)n * w_cols + k), acc[u]);
        }
      }
    }
    const size_t row_off = ((size_t)(b * p.OH + j * p.os + gc.oy0) * p.OW + (i * p.os + gc.ox0)) * p.ON;
    float4 o;
    o.x = epilogue_value(p, row_off + ng * 4 + 0, ng * 4 + 0, acc[0]);
    o.y = epilogue_value(p, row_off + ng * 4 + 1, ng * 4 + 1, acc[1]);
    o.z = epilogue_value(p, row_off + ng * 4 + 2, ng * 4 + 2, acc[2]);
    o.w = epilogue_value(p, row_off + ng * 4 + 3, ng * 4 + 3, acc[3]);
    *reinterpret_cast<float4*>(p.out + row_off + ng * 4) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int pick_bn(int N) {
  if (N <= 16) return 16;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return 256;
}

template <int BN>
int launch_tc(ConvGemmParams p, const float* w, int w_rows, int w_cols, cudaStream_t stream) {
  using C = Cfg<BN>;
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(CGS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {(cuuint64_t)w_cols, (cuuint64_t)w_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)w_cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(CGS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  p.n_tiles = (p.N + BN - 1) / BN;
  p.m_tiles = (p.M + BM - 1) / BM;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(CGS_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.nclasses;
  const int grid = total < num_sms ? total : num_sms;
  conv_gemm_tc_kernel<BN><<<grid, kThreads, C::SMEM_BYTES, stream>>>(p, tmap); count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(CGS_ERR_CUDA, "conv_gemm_tc launch: %s", cudaGetErrorString(e));
  return CGS_OK;
}

int validate(const ConvGemmParams& p, int w_cols) {
  if (p.ON % 4 != 0) return set_error(CGS_ERR_INVALID, "output channel stride %d must be a multiple of 4", p.ON);
  if (p.cblocks == 0 && p.Cs != 4) return set_error(CGS_ERR_INVALID, "pixel mode needs channel stride 4");
  if (p.nclasses < 1 || p.nclasses > kMaxClasses) return set_error(CGS_ERR_INVALID, "bad class count");
  if (w_cols % 4 != 0) return set_error(CGS_ERR_INVALID, "weight row length must be a multiple of 4 floats");
  for (int c = 0; c < p.nclasses; ++c) {
    if (p.cls[c].ntaps > kMaxTaps) return set_error(CGS_ERR_INVALID, "too many taps");
    if (p.cls[c].k0 + p.cls[c].nkb * BK > w_cols) return set_error(CGS_ERR_INVALID, "class K range exceeds weight matrix");
  }
  return CGS_OK;
}

}  // namespace

int launch_conv_gemm_tc(const ConvGemmParams& p, const float* w, int w_rows, int w_cols, cudaStream_t stream) {
  if (int rc = validate(p, w_cols)) return rc;
  if (p.M <= 0) return CGS_OK;
  switch (pick_bn(p.N)) {
    case 16: return launch_tc<16>(p, w, w_rows, w_cols, stream);
    case 32: return launch_tc<32>(p, w, w_rows, w_cols, stream);
    case 64: return launch_tc<64>(p, w, w_rows, w_cols, stream);
    case 128: return launch_tc<128>(p, w, w_rows, w_cols, stream);
    default: return launch_tc<256>(p, w, w_rows, w_cols, stream);
  }
}

int launch_conv_gemm_simt(const ConvGemmParams& p, const float* w, int w_rows, int w_cols, cudaStream_t stream) {
  (void)w_rows;
  if (int rc = validate(p, w_cols)) return rc;
  if (p.M <= 0) return CGS_OK;
  const long long total = (long long)p.nclasses * p.M * (p.ON / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  conv_gemm_simt_kernel<<<(int)blocks, 256, 0, stream>>>(p, w, w_cols); count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(CGS_ERR_CUDA, "conv_gemm_simt launch: %s", cudaGetErrorString(e));
  return CGS_OK;
}

}  // namespace cgs
