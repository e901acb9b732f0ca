// Probe: tcgen05 SWIZZLE_64B K-major A operand (rows of 64 bytes = 16 floats, 8-row groups of 512 bytes) with a start
// address shifted by s x 64 bytes, second k-step at +32 bytes; B operand SWIZZLE_128B.  D[128][64] = A[128][16] . B[64][16]^T
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "ptx.cuh"
using namespace cgs;

constexpr int ROWS = 320;
constexpr int N = 64;

__device__ __forceinline__ uint64_t desc_sw(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
  const uint32_t lo = (smem_addr >> 4) & 0x3FFFu;
  const uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (layout_type << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

__global__ void __launch_bounds__(128, 1) probe(const float* a, const float* b, float* d, int shift, int layout_a) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  float* sa = reinterpret_cast<float*>(smem);                       // [ROWS][16] floats, 64-byte rows, SW64: chunk ^= (row >> 1) & 3
  float* sb = reinterpret_cast<float*>(smem + 32768);               // [N][32] floats SW128 (only k 0..15 used)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < ROWS * 16; i += 128) {
    const int r = i / 16, c = i % 16;
    sa[r * 16 + (((c / 4) ^ ((r >> 1) & 3)) * 4) + (c % 4)] = a[i];
  }
  for (int i = threadIdx.x; i < N * 32; i += 128) {
    const int r = i / 32, c = i % 32;
    sb[r * 32 + (((c / 4) ^ (r & 7)) * 4) + (c % 4)] = c < 16 ? b[r * 16 + c] : 0.f;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc(&tmem_ptr, 64);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint64_t da = desc_sw(smem_u32(sa) + shift * 64, 512, (uint32_t)layout_a);
    const uint64_t db = desc_sw(smem_u32(sb), 1024, 2);
    for (int k = 0; k < 2; ++k) umma_tf32_ss(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem + c0 + ((threadIdx.x & ~31u) << 16), v);
    tmem_ld_wait();
    for (int n = 0; n < 16; ++n) d[threadIdx.x * N + c0 + n] = __uint_as_float(v[n]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<float> a(ROWS * 16), b(N * 16), d(128 * N);
  for (auto& x : a) x = (float)((rand() % 17) - 8);
  for (auto& x : b) x = (float)((rand() % 9) - 4);
  float *da, *db, *dd;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 32768 + N * 128 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int shifts[] = {0, 8, 1, 2, 3, 5, 35, 70, 150};
  for (int layout : {4, 6, 1})                 // candidate encodings of SWIZZLE_64B
    for (int s : shifts) {
      cudaMemset(dd, 0, d.size() * 4);
      probe<<<1, 128, smem>>>(da, db, dd, s, layout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("layout %d shift %d: CUDA error %s\n", layout, s, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          float ref = 0;
          for (int k = 0; k < 16; ++k) ref += a[(r + s) * 16 + k] * b[n * 16 + k];
          if (fabsf(ref - d[r * N + n]) > 1e-3f) ++bad;
        }
      printf("layout_type %d shift %3d : %s (%d of %d wrong)\n", layout, s, bad ? "MISMATCH" : "ok", bad, 128 * N);
    }
  return 0;
}
