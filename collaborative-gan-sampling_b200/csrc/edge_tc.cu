// edge_narrow_tc: the narrow image-edge pass (64 channels -> image-like output, stride-2 transposed type: the
// generator's last deconv forward, nsgan/ops.py:55, and the data-gradient of the discriminator's first conv,
// sampling/collaborator.py:31) on the 5th-generation tensor cores, with the input PATCH RESIDENT in shared memory.
//
// Gather form (see edge_narrow2 in edge_conv.cu): for the input-grid point (j, i) the 4 output-parity classes x 4
// padded channels are 16 GEMM columns and the reduction runs over (dy, dx, 64 channels), dy, dx in {-1, 0, 1}:
//     acc[(j, i)][(py, px, c)] += in[j + dy][i + dx][:] . Wshift[dy][dx][:][(py, px, c)].
// The legacy mma.sync path is throughput-bound on this part (one HMMA.1688.TF32 per ~23 cycles per SM sub-partition,
// ~100 TFLOP/s chip-wide: profiles/round2_ncu_summary.md), so the 9x MMA inflation of the gather form made the
// mma.sync kernels 4x slower than the HBM floor.  tcgen05 has the rate, and its shared-memory descriptors make the
// nine shifted A operands FREE: a band of input rows (+ one halo row / column each side, zero-filled by TMA) is loaded
// once as [pixel][32 channels] 128-byte rows (SWIZZLE_128B, two channel halves); the A tile of shift (dy, dx) for the
// 128 "virtual pixels" v .. v+127 is the same memory starting (1 + dy) * pitch + (1 + dx) rows later -- a descriptor
// whose start address is simply advanced by a multiple of 128 bytes (the hardware swizzles on absolute address bits:
// tools/probes/umma_shift_probe.cu).  Virtual pixels run over the halo columns too; those rows are never stored.
//
// CTA = 6 warps, persistent over (image, band) tiles: warp 0 TMA producer (one box per channel half into a ring of
// half-patch stages), warp 1 MMA issuer (tcgen05.mma kind::tf32, M = 128, N = 16, accumulators in TMEM, double-buffered
// per band), warps 2-5 epilogue (tcgen05.ld, bias + tanh or x tanh', pitched float4 stores).  HBM-bound by design:
// every input byte is read once (+ 2 halo rows per band, L2 hits), every output byte written once.
#include "edge_conv.cuh"

#include <cuda.h>
#include <cstdint>
#include <cstring>

#include "common.h"
#include "conv_gemm.cuh"
#include "ptx.cuh"

namespace cgs {

namespace {

constexpr int TC_THREADS = 192;
constexpr int TC_MAX_MT = 4;                   // M tiles (128 virtual pixels) per band
constexpr int TC_B_BYTES = 9 * 2 * 16 * 128;   // weights: [shift][channel half][16 columns][32 k] = 36 KB
constexpr int TC_SLACK_BYTES = 16 * 1024;      // rows a (never stored) virtual pixel past the band may read

struct NarrowTcParams {
  EdgeNarrowParams p;
  int R, bands, MT, PW, PR;                    // rows per band, bands per image, M tiles per band, patch pitch / rows
  int stage_bytes, nstages, ntiles;
};

// scalar epilogue of an image-edge pass (same formulas as epilogue4 in conv_gemm.cuh; no policy step here)
__device__ __forceinline__ float epilogue1_tc(const EdgeEpi& e, float a, float x0, float*) {
  float o = a;
  if (e.epi == EPI_FWD) {
    const float v = a + x0;
    o = e.act_tanh ? tanhf(v) : fmaxf(v, v * e.slope);
  } else if (e.epi == EPI_BWD) {
    o = e.act_tanh ? a * (1.f - x0 * x0) : a * (x0 > 0.f ? 1.f : e.slope);
  }
  return e.round_out ? tf32_rn(o) : o;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
edge_narrow_tc_kernel(const __grid_constant__ NarrowTcParams q, const __grid_constant__ CUtensorMap tmap_in) {
  const EdgeNarrowParams& p = q.p;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;                                    // [9][2][16 rows][128 B], swizzled
  uint8_t* smem_a = smem + TC_B_BYTES;                       // ring of half-patch stages (+ slack behind the last)
  __shared__ uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_ptr_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // weights -> SWIZZLE_128B K-major tiles: tile (shift, half), row n = (py * 2 + px) * 4 + c, k = channel in the half
  for (int idx = threadIdx.x; idx < 9 * 2 * 16 * 8; idx += TC_THREADS) {
    const int ck = idx & 7, n = (idx >> 3) & 15, h = (idx >> 7) & 1, sh = idx >> 8;
    const int dy = sh / 3 - 1, dx = sh % 3 - 1;
    const int py = n >> 3, px = (n >> 2) & 1, c = n & 3;
    const int ky = py + p.pad_y - 2 * dy, kx = px + p.pad_x - 2 * dx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < p.cimg && ky >= 0 && ky < p.k && kx >= 0 && kx < p.k)
      v = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)((ky * p.k + kx) * 4 + c) * p.K + h * 32 + ck * 4));
    *reinterpret_cast<float4*>(smem_b + ((sh * 2 + h) * 16 + n) * 128 + ((ck ^ (n & 7)) << 4)) = v;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < q.nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
    fence_barrier_init();
  }
  fence_proxy_async_smem();                                  // generic-proxy weight stores -> tensor-core reads
  if (warp == 1) tmem_alloc(&tmem_ptr_smem, 2 * TC_MAX_MT * 16);
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_in);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();
  int ntiles = q.ntiles;
  if (p.e.live) ntiles = min(ntiles, live_images(p.e.live, p.B) * q.bands);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      const uint32_t bytes = (uint32_t)q.PR * q.PW * 128u;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / q.bands;
        const int j0 = (tile - b * q.bands) * q.R;
        for (int h = 0; h < 2; ++h) {
          const uint32_t s = stage, ph = phase;
          if (++stage == (uint32_t)q.nstages) { stage = 0; phase ^= 1u; }
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          // box {32 channels, PW pixels from x = -1, PR rows from y = j0 - 1, 1 image}; outside the image reads zeros
          tma_load_4d(smem_u32(smem_a) + s * q.stage_bytes, &tmap_in, &full_bar[s], h * 32, -1, j0 - 1, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_tf32(128, 16);
      const uint64_t da0 = make_smem_desc_sw128(smem_u32(smem_a));
      const uint64_t db0 = make_smem_desc_sw128(smem_u32(smem_b));
      uint32_t stage = 0, phase = 0, tile_count = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_count) {
        const uint32_t acc = tile_count & 1, acc_ph = (tile_count >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
        for (int h = 0; h < 2; ++h) {
          const uint32_t s = stage;
          mbar_wait(&full_bar[s], phase);
          if (++stage == (uint32_t)q.nstages) { stage = 0; phase ^= 1u; }
          tcgen05_fence_after();
          for (int mt = 0; mt < q.MT; ++mt) {
            const uint32_t tmem_d = tmem_base + (acc * TC_MAX_MT + mt) * 16;
#pragma unroll
            for (int sh = 0; sh < 9; ++sh) {
              // rows of the A tile = stored pixels 128 mt + (dyi * PW + dxi) ..: descriptor start in 16-byte units
              const uint32_t row0 = 128u * mt + (uint32_t)(sh / 3) * q.PW + (uint32_t)(sh % 3);
              const uint64_t da = da0 + (uint64_t)((s * (uint32_t)q.stage_bytes + row0 * 128u) >> 4);
              const uint64_t db = db0 + (uint64_t)(((sh * 2 + h) * 16 * 128) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (h > 0 || sh > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);                          // half-patch stage reusable once these MMAs have read it
        }
        umma_commit(&tmem_full_bar[acc]);                      // all accumulators of the band complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2-5: TMEM lane quarter = warp % 4)
    const int quarter = warp & 3;
    const bool bwd = p.e.epi == EPI_BWD;
    const float4 bias4 = (p.e.epi == EPI_FWD && p.e.bias) ? __ldg(reinterpret_cast<const float4*>(p.e.bias))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t tile_count = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_count) {
      const int b = tile / q.bands;
      const int j0 = (tile - b * q.bands) * q.R;
      const int j1 = min(p.IH, j0 + q.R);
      const uint32_t acc = tile_count & 1, acc_ph = (tile_count >> 1) & 1;
      bool waited = false;
      for (int mt = 0; mt < q.MT; ++mt) {
        const int v = 128 * mt + quarter * 32 + lane;          // virtual pixel of this lane
        const int jr = v / q.PW, i = v - jr * q.PW, j = j0 + jr;
        const bool valid = i < p.IW && j < j1;
        float* obase = p.out + ((size_t)b * p.OH + 2 * j) * p.out_pitch * 4 + (size_t)(2 * i + p.out_xoff) * 4;
        // derivative operand (the forward image at the four output pixels): in flight while the MMAs still run
        float4 aux[2][2];
        if (bwd && valid) {
          const float* abase = p.e.aux + (obase - p.out);
#pragma unroll
          for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) aux[py][px] = __ldg(reinterpret_cast<const float4*>(abase + (size_t)py * p.out_pitch * 4 + px * 4));
        }
        if (!waited) {
          mbar_wait(&tmem_full_bar[acc], acc_ph);
          tcgen05_fence_after();
          waited = true;
        }
        uint32_t a[16];
        tmem_ld_32x32b_x16(tmem_base + (acc * TC_MAX_MT + mt) * 16 + (static_cast<uint32_t>(quarter * 32) << 16), a);
        tmem_ld_wait();
        if (mt == q.MT - 1) {                                  // accumulators read: hand the TMEM buffer back
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (valid) {
#pragma unroll
          for (int py = 0; py < 2; ++py) {
            float* orow = obase + (size_t)py * p.out_pitch * 4;
#pragma unroll
            for (int px = 0; px < 2; ++px) {
              const int n0 = (py * 2 + px) * 4;
              const float4 x0 = bwd ? aux[py][px] : bias4;
              float unused;
              float4 o;
              o.x = epilogue1_tc(p.e, __uint_as_float(a[n0]), x0.x, &unused);
              o.y = p.cimg > 1 ? epilogue1_tc(p.e, __uint_as_float(a[n0 + 1]), x0.y, &unused) : 0.f;
              o.z = p.cimg > 2 ? epilogue1_tc(p.e, __uint_as_float(a[n0 + 2]), x0.z, &unused) : 0.f;
              o.w = 0.f;
              *reinterpret_cast<float4*>(orow + px * 4) = o;
            }
            // margins of the pitched layout are zeros: written by the lanes at the two ends of the row
            if (i == 0)
              for (int c = 0; c < p.out_xoff; ++c) *reinterpret_cast<float4*>(orow - (size_t)(p.out_xoff - c) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i == p.IW - 1)
              for (int c = p.out_xoff + p.OW; c < p.out_pitch; ++c)
                *reinterpret_cast<float4*>(orow + (size_t)(c - p.out_xoff - 2 * i) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * TC_MAX_MT * 16);
}

typedef CUresult (*PFN_encodeTiledTc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiledTc encode_fn() {
  static PFN_encodeTiledTc fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiledTc>(ptr);
  }
  return fn;
}

}  // namespace

bool edge_narrow_tc_supported(const EdgeNarrowParams& p) {
  return p.K == 64 && p.pad_y == 1 && p.pad_x == 1 && p.IW >= 8 && p.IW <= 64 && (p.k == 4 || p.k == 5) &&
         p.OW == 2 * p.IW && p.OH == 2 * p.IH && p.e.epi != EPI_UPDATE && p.cimg >= 1 && p.cimg <= 3;
}

int launch_edge_narrow_tc(const EdgeNarrowParams& p, cudaStream_t st) {
  if (p.B <= 0) return CGS_OK;
  NarrowTcParams q;
  std::memset(&q, 0, sizeof(q));
  q.p = p;
  q.PW = p.IW + 2;
  // rows per band: the largest R <= 16 whose virtual pixels ((R - 1) * PW + IW) fill their 128-row M tiles best
  int bestR = 1;
  double best = -1.0;
  for (int R = 1; R <= 16 && R <= p.IH; ++R) {
    const int valid = (R - 1) * q.PW + p.IW;
    const int mt = (valid + 127) / 128;
    if (mt > TC_MAX_MT) break;
    const int bands = (p.IH + R - 1) / R;
    // cost model: M tiles per image (epilogue work) + a small charge per band (two halo rows re-read)
    const int last = p.IH - (bands - 1) * R;
    const int mt_last = ((last - 1) * q.PW + p.IW + 127) / 128;
    const double cost = (bands - 1) * mt + mt_last + 0.15 * bands;
    const double score = 1.0 / cost;
    if (score > best) { best = score; bestR = R; }
  }
  q.R = bestR;
  q.bands = (p.IH + q.R - 1) / q.R;
  q.PR = q.R + 2;
  q.MT = ((q.R - 1) * q.PW + p.IW + 127) / 128;
  q.stage_bytes = ((q.PR * q.PW * 128 + 1023) / 1024) * 1024;
  const int budget = 227 * 1024 - TC_B_BYTES - TC_SLACK_BYTES - 2048;
  q.nstages = budget / q.stage_bytes;
  if (q.nstages > 8) q.nstages = 8;
  if (q.nstages < 2) return set_error(CGS_ERR_UNSUPPORTED, "edge_narrow_tc: band does not fit shared memory");
  // a virtual pixel past the band may read up to 128 * MT + 2 * PW + 2 rows from the stage start
  if ((128 * q.MT + 2 * q.PW + 2) * 128 > q.stage_bytes + TC_SLACK_BYTES)
    return set_error(CGS_ERR_UNSUPPORTED, "edge_narrow_tc: slack too small");
  const long long tiles = (long long)p.B * q.bands;
  if (tiles >= (1ll << 31)) return set_error(CGS_ERR_UNSUPPORTED, "batch too large; split the batch");
  q.ntiles = (int)tiles;
  PFN_encodeTiledTc enc = encode_fn();
  if (!enc) return set_error(CGS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap tmap;
  cuuint64_t gdim[4] = {64, (cuuint64_t)p.IW, (cuuint64_t)p.IH, (cuuint64_t)p.B};
  cuuint64_t gstr[3] = {64 * 4, (cuuint64_t)p.IW * 64 * 4, (cuuint64_t)p.IH * p.IW * 64 * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)q.PW, (cuuint32_t)q.PR, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.in), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(CGS_ERR_CUDA, "cuTensorMapEncodeTiled (edge_narrow_tc) failed (%d)", (int)r);
  const size_t smem = (size_t)TC_B_BYTES + (size_t)q.nstages * q.stage_bytes + TC_SLACK_BYTES + 1024;
  static DynSmemCache smem_cache;
  cudaError_t e = ensure_dyn_smem(edge_narrow_tc_kernel, smem, smem_cache);
  if (e != cudaSuccess) return set_error(CGS_ERR_CUDA, "cudaFuncSetAttribute(edge_narrow_tc): %s", cudaGetErrorString(e));
  long long grid = tiles < device_num_sms() ? tiles : device_num_sms();
  cudaError_t le = launch_pdl(edge_narrow_tc_kernel, dim3((unsigned)grid), dim3(TC_THREADS), smem, st, q, tmap);
  count_launch();
  if (le != cudaSuccess) return set_error(CGS_ERR_CUDA, "edge_narrow_tc_kernel: %s", cudaGetErrorString(le));
  return check_launch("edge_narrow_tc_kernel");
}

}  // namespace cgs
