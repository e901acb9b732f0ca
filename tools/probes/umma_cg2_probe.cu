// Probe: tcgen05.mma.cta_group::2 (kind::tf32, M = 256 over a cluster of two CTAs) with the operand split this
// repo's conv_gemm kernel wants to use:
//   * each CTA holds ITS 128 rows of A and HALF the N rows of B (CTA r: rows [r N/2, (r+1) N/2)),
//   * both CTAs load with cp.async.bulk.tensor ... .cta_group::2 and signal the LEADER's mbarrier,
//   * the peer releases a gate on the leader with a remote mbarrier.arrive (the accumulator-free hand-shake),
//   * the leader issues the MMAs and commits with .multicast::cluster to the barrier at the same offset in both CTAs,
//   * each CTA reads its own 128 TMEM lanes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I collaborative-gan-sampling_b200/csrc -I include \
//             -o /tmp/cg2_probe tools/probes/umma_cg2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda.h>
#include "ptx.cuh"
using namespace cgs;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// 2-SM TMA load: data into THIS CTA's smem, completion bytes on the barrier at the same offset in the pair's even CTA
__device__ __forceinline__ void tma2_load_2d(uint32_t dst_smem, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
      "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* d, int N) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                    // [128][32] floats, SWIZZLE_128B
  uint8_t* sb = smem + 128 * 128;        // [N/2][32] floats
  __shared__ uint64_t full_bar, done_bar, gate_bar;
  __shared__ uint32_t tmem_ptr;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    mbar_init(&gate_bar, 2);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc2(&tmem_ptr, 256);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t bytes_cta = 128 * 128 + (N / 2) * 128;
    if (rank == 0) mbar_arrive_expect_tx(&full_bar, 2 * bytes_cta);
    tma2_load_2d(smem_u32(sa), &tm_a, &full_bar, 0, (int)rank * 128);
    tma2_load_2d(smem_u32(sb), &tm_b, &full_bar, 0, (int)rank * (N / 2));
    if (rank == 0) mbar_arrive(&gate_bar); else mbar_arrive_remote(&gate_bar, 0);
    if (rank == 0) {
      mbar_wait(&gate_bar, 0);
      mbar_wait(&full_bar, 0);
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_tf32(256, (uint32_t)N);
      const uint64_t da = make_smem_desc_sw128(smem_u32(sa));
      const uint64_t db = make_smem_desc_sw128(smem_u32(sb));
      for (int k = 0; k < 4; ++k) umma2_tf32_ss(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
      umma2_commit_mc(&done_bar, 3);
    }
  }
  mbar_wait(&done_bar, 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem + ((threadIdx.x & ~31u) << 16) + c0, v);
    tmem_ld_wait();
    for (int n = 0; n < 16; ++n) d[(rank * 128 + threadIdx.x) * N + c0 + n] = __uint_as_float(v[n]);
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc2(tmem, 256);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(PFN_encodeTiled enc, float* p, int rows, int box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {32, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
  PFN_encodeTiled enc = (PFN_encodeTiled)fp;
  const int Ns[] = {64, 128, 256, 32, 192};
  for (int N : Ns) {
    std::vector<float> a(256 * 32), b(N * 32), d(256 * N);
    for (auto& x : a) x = (float)((rand() % 17) - 8);      // small integers: exact in TF32
    for (auto& x : b) x = (float)((rand() % 9) - 4);
    float *da, *db, *dd;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, d.size() * 4);
    CUtensorMap ta = make_map(enc, da, 256, 128), tb = make_map(enc, db, N, N / 2);
    const size_t smem = 128 * 128 + (N / 2) * 128 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<2, 128, smem>>>(ta, tb, dd, N);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bad_top = 0;
    for (int r = 0; r < 256; ++r)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < 32; ++k) ref += a[r * 32 + k] * b[n * 32 + k];
        if (fabsf(ref - d[r * N + n]) > 1e-3f) { ++bad; if (r < 128) ++bad_top; }
      }
    printf("N %3d : %s (%d of %d wrong, %d in the leader's rows)\n", N, bad ? "MISMATCH" : "ok", bad, 256 * N, bad_top);
    cudaFree(da); cudaFree(db); cudaFree(dd);
  }
  return 0;
}
