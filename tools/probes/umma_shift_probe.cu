// Probe: does a tcgen05 SWIZZLE_128B K-major A descriptor whose start address is shifted by s x 128 bytes (not 1024-byte
// aligned) read rows s .. s+127 of a larger resident matrix?  Tries base_offset = 0 and base_offset = (addr >> 7) & 7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I collaborative-gan-sampling_b200/csrc -I include -o /tmp/probe tools/probes/umma_shift_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "ptx.cuh"
using namespace cgs;

constexpr int ROWS = 256;   // resident "patch" rows of 128 bytes (32 floats)
constexpr int N = 16;

__device__ __forceinline__ uint64_t desc_sw128_off(uint32_t smem_addr, int use_base_offset) {
  const uint32_t lo = (smem_addr >> 4) & 0x3FFFu;
  uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  if (use_base_offset) hi |= ((smem_addr >> 7) & 7u) << 17;      // bits 49-51 of the 64-bit descriptor
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

__global__ void __launch_bounds__(128, 1) probe(const float* a, const float* b, float* d, int shift, int use_bo) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  float* sa = reinterpret_cast<float*>(smem);                       // [ROWS][32], swizzled
  float* sb = reinterpret_cast<float*>(smem + ROWS * 128);          // [N][32], swizzled (1024-aligned: ROWS*128 = 32 KB)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < ROWS * 32; i += 128) {
    const int r = i / 32, c = i % 32;
    sa[r * 32 + (((c / 4) ^ (r & 7)) * 4) + (c % 4)] = a[i];
  }
  for (int i = threadIdx.x; i < N * 32; i += 128) {
    const int r = i / 32, c = i % 32;
    sb[r * 32 + (((c / 4) ^ (r & 7)) * 4) + (c % 4)] = b[i];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc(&tmem_ptr, 32);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint64_t da = desc_sw128_off(smem_u32(sa) + shift * 128, use_bo);
    const uint64_t db = desc_sw128_off(smem_u32(sb), 0);
    for (int k = 0; k < 4; ++k) umma_tf32_ss(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  uint32_t v[16];
  tmem_ld_32x32b_x16(tmem + ((threadIdx.x & ~31u) << 16), v);
  tmem_ld_wait();
  for (int n = 0; n < N; ++n) d[threadIdx.x * N + n] = __uint_as_float(v[n]);
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

int main() {
  std::vector<float> a(ROWS * 32), b(N * 32), d(128 * N);
  for (auto& x : a) x = (float)((rand() % 17) - 8);      // small integers: exact in TF32
  for (auto& x : b) x = (float)((rand() % 9) - 4);
  float *da, *db, *dd;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = ROWS * 128 + N * 128 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int shifts[] = {0, 8, 1, 3, 5, 35, 70, 127};
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int s : shifts) {
      cudaMemset(dd, 0, d.size() * 4);
      probe<<<1, 128, smem>>>(da, db, dd, s, use_bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d bo %d: CUDA error %s\n", s, use_bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          float ref = 0;
          for (int k = 0; k < 32; ++k) ref += a[(r + s) * 32 + k] * b[n * 32 + k];
          if (fabsf(ref - d[r * N + n]) > 1e-3f) ++bad;
        }
      printf("shift %3d  base_offset %s : %s (%d of %d wrong)\n", s, use_bo ? "set " : "zero", bad ? "MISMATCH" : "ok", bad, 128 * N);
    }
  return 0;
}
