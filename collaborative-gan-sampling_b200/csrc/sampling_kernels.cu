// Memory-bound kernels of the accept-reject stage and the standalone update policy:
//   policy step     sampling/policy.py:26-64
//   DRS             sampling/rejector.py:11-38   (FP64, decisions bit-exact given the same uniforms)
//   MH independence sampling/idpsampler.py:17-53 (exact parallel restatement of the sequential chain)
//   ordered compaction + row gather
// All of them are HBM/latency bound: coalesced grid-stride loops, warp-shuffle reductions, no tensor cores.
#include "common.h"
#include "policy.cuh"
#include "philox.cuh"

#include <cfloat>
#include <cmath>

namespace cgs {
namespace {

constexpr int kBlock = 256;
constexpr double kClipLo = 1e-14;            // rejector.py:12,18
constexpr double kClipHi = 1.0 - 1e-14;

inline int grid_for(int64_t n, int per_block = kBlock, int cap = 148 * 8) {
  int64_t g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------ policy
__global__ void policy_loss_avg_kernel(PolicyConsts c, const float* __restrict__ loss, float* __restrict__ loss_avg,
                                       int first, int64_t rows) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
    loss_avg[r] = ladam_loss_avg(c, loss_avg[r], loss[r], first);
}

__global__ void policy_step_kernel(PolicyConsts c, float* __restrict__ theta, const float* __restrict__ grad,
                                   float* __restrict__ mom, float* __restrict__ msq,
                                   const float* __restrict__ loss_avg, int first, int64_t rows, int64_t cols,
                                   int clip_hi) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = grad[i];
    float th = theta[i];
    if (c.method == CGS_POLICY_SGD) {
      th = sgd_update(c, th, g);
    } else if (c.method == CGS_POLICY_MOMENTUM) {
      float m = first ? 0.f : mom[i];
      th = momentum_update(c, th, g, m, first);
      mom[i] = m;
    } else {
      float m = first ? 0.f : mom[i];
      float v = first ? 0.f : msq[i];
      th = ladam_update(c, th, g, m, v, loss_avg[i / cols], first, clip_hi);
      mom[i] = m;
      msq[i] = v;
    }
    theta[i] = th;
  }
}

// ------------------------------------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double block_max(double v) {
  __shared__ double sm[kBlock / 32];
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < kBlock / 32 ? sm[threadIdx.x] : -DBL_MAX;
    t = warp_max(t);
    if (threadIdx.x == 0) sm[0] = t;
  }
  __syncthreads();
  const double r = sm[0];
  __syncthreads();
  return r;
}

__device__ __forceinline__ double load_score(const void* s, int dtype, int64_t i) {
  return dtype == CGS_F64 ? static_cast<const double*>(s)[i] : (double)static_cast<const float*>(s)[i];
}
__device__ __forceinline__ double clip_score(double s) { return fmin(fmax(s, kClipLo), kClipHi); }
__device__ __forceinline__ double logit_f64(double s) { return log(s / (1.0 - s)); }     // scipy.special.logit

// ------------------------------------------------------------------------------------------------ DRS
// pass 1: D_tilde = logit(clip(sigmoid)), per-block maxima                      rejector.py:18-21
__global__ void drs_logit_kernel(const void* sig, int dtype, int64_t n, double* __restrict__ dt,
                                 double* __restrict__ block_maxima) {
  double mx = -DBL_MAX;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = logit_f64(clip_score(load_score(sig, dtype, i)));
    dt[i] = v;
    mx = fmax(mx, v);
  }
  mx = block_max(mx);
  if (threadIdx.x == 0) block_maxima[blockIdx.x] = mx;
}
// M <- max(M, max D_tilde)                                                       rejector.py:22
__global__ void drs_update_max_kernel(const double* __restrict__ block_maxima, int nblocks, double* m_inout) {
  double mx = -DBL_MAX;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) mx = fmax(mx, block_maxima[i]);
  mx = block_max(mx);
  if (threadIdx.x == 0) *m_inout = fmax(*m_inout, mx);
}
// pass 2: F = D_delta - log(1 - exp(D_delta - eps)), per-block maxima             rejector.py:25-26
__global__ void drs_f_kernel(double* __restrict__ dt_to_f, int64_t n, const double* __restrict__ m, double eps,
                             double* __restrict__ block_maxima) {
  const double M = *m;
  double mx = -DBL_MAX;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double delta = dt_to_f[i] - M;
    const double f = delta - log(1.0 - exp(delta - eps));
    dt_to_f[i] = f;
    mx = fmax(mx, f);
  }
  mx = block_max(mx);
  if (threadIdx.x == 0) block_maxima[blockIdx.x] = mx;
}
__global__ void reduce_max_kernel(const double* __restrict__ block_maxima, int nblocks, double* out) {
  double mx = -DBL_MAX;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) mx = fmax(mx, block_maxima[i]);
  mx = block_max(mx);
  if (threadIdx.x == 0) *out = mx;
}

// Order statistics for a general percentile (rejector.py:28, numpy 'linear' method): MSB-first radix select on
// order-preserving 64-bit keys.  sel[0] = prefix, sel[1] = remaining rank, hist = 256 bins.
__device__ __forceinline__ unsigned long long f64_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
__global__ void select_hist_kernel(const double* __restrict__ f, int64_t n, const unsigned long long* sel, int pass,
                                   unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const unsigned long long prefix = sel[0];
  const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = f64_key(f[i]);
    if ((k & mask) == prefix) atomicAdd(&sh[(k >> shift) & 0xff], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
__global__ void select_pick_kernel(unsigned long long* sel, int pass, unsigned int* hist) {
  if (threadIdx.x == 0) {
    const int shift = 56 - 8 * pass;
    unsigned long long rank = sel[1];
    int b = 0;
    for (; b < 256; ++b) {
      const unsigned int c = hist[b];
      if (rank < c) break;
      rank -= c;
    }
    if (b > 255) b = 255;
    sel[0] |= (unsigned long long)b << shift;
    sel[1] = rank;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
}
__global__ void select_init_kernel(unsigned long long* sel, unsigned long long rank, unsigned int* hist) {
  if (threadIdx.x == 0) {
    sel[0] = 0;
    sel[1] = rank;
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
}
// gamma = lerp(F_(lo), F_(hi), t) exactly as numpy's _lerp (function_base.py): a + (b-a)*t, and b - (b-a)*(1-t) for t >= .5
__global__ void percentile_lerp_kernel(const unsigned long long* sel_lo, const unsigned long long* sel_hi, double t,
                                       double* gamma) {
  const double a = key_f64(sel_lo[0]);
  const double b = key_f64(sel_hi[0]);
  const double d = b - a;
  double r = __dadd_rn(a, __dmul_rn(d, t));
  if (t >= 0.5) r = __dsub_rn(b, __dmul_rn(d, 1.0 - t));
  if (t == 0.0) r = a;
  *gamma = r;
}
__global__ void set_double_kernel(double* p, double v) { *p = v; }

// pass 3: P = expit(F - gamma), accept = u < P                                    rejector.py:29-33
__global__ void drs_accept_kernel(const double* __restrict__ f, int64_t n, const double* __restrict__ gamma,
                                  const double* __restrict__ uniforms, uint64_t seed, uint64_t offset,
                                  uint8_t* __restrict__ accept, double* __restrict__ prob_out) {
  const double g = *gamma;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double p = 1.0 / (1.0 + exp(-(f[i] - g)));            // scipy.special.expit
    const double u = uniforms ? uniforms[i] : philox_uniform_f64(seed, offset + (uint64_t)i);
    accept[i] = (u < p) ? 1 : 0;
    if (prob_out) prob_out[i] = p;
  }
}

__global__ void drs_score_max_kernel(const void* s, int dtype, double* m) {
  *m = logit_f64(clip_score(load_score(s, dtype, 0)));          // rejector.py:12-14
}

// ------------------------------------------------------------------------------------------------ ordered compaction
// flags [n] -> ascending indices of set flags + count.  Three small kernels; chunk = 1024 flags per block.
constexpr int kChunk = 1024;
__global__ void compact_count_kernel(const uint8_t* __restrict__ flags, int64_t n, int* __restrict__ block_counts) {
  __shared__ int sm[kBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kChunk;
  int c = 0;
  for (int t = threadIdx.x; t < kChunk; t += kBlock) {
    const int64_t i = base + t;
    c += (i < n && flags[i]) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < kBlock / 32; ++i) t += sm[i];
    block_counts[blockIdx.x] = t;
  }
}
__global__ void compact_scan_kernel(int* __restrict__ block_counts, int nblocks, int* __restrict__ total) {
  // single block, serial over tiles of blockDim: exclusive scan in place
  __shared__ int sm[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? block_counts[i] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
      const int t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0;
      __syncthreads();
      sm[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) block_counts[i] = carry + sm[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sm[blockDim.x - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void compact_scatter_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                       const int* __restrict__ block_offsets, int* __restrict__ idx_out) {
  // warp w of the block owns flags [base + 32*k*8 ...] in order: do an ordered in-block scan with ballots
  __shared__ int warp_tot[kBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kChunk;
  int running = block_offsets[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int sub = 0; sub < kChunk; sub += kBlock) {
    const int64_t i = base + sub + threadIdx.x;
    const int f = (i < n && flags[i]) ? 1 : 0;
    const unsigned b = __ballot_sync(0xffffffffu, f);
    const int before = __popc(b & ((1u << lane) - 1));
    if (lane == 0) warp_tot[warp] = __popc(b);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) {
      const int t = warp_tot[w];
      if (w < warp) woff += t;
      tot += t;
    }
    if (f) idx_out[running + woff + before] = (int)i;
    running += tot;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ MH
// acceptance test of idpsampler.py:48-51 in the dtype numpy would use
__device__ __forceinline__ bool mh_move(double d_state, bool f64_math, const void* sig, int dtype, int64_t j, double u) {
  if (f64_math) {
    const double dn = load_score(sig, dtype, j);
    const double num = __dmul_rn(dn, __dsub_rn(1.0, d_state));
    const double den = __dmul_rn(d_state, __dsub_rn(1.0, dn));
    const double ratio = __ddiv_rn(num, den);
    const double alpha = ratio < 1.0 ? ratio : 1.0;             // python min(1.0, x): NaN -> 1.0
    return !(u > alpha);
  } else {
    const float dn = static_cast<const float*>(sig)[j];
    const float dc = (float)d_state;
    const float num = __fmul_rn(dn, __fsub_rn(1.0f, dc));
    const float den = __fmul_rn(dc, __fsub_rn(1.0f, dn));
    const float ratio = __fdiv_rn(num, den);
    const float alpha = ratio < 1.0f ? ratio : 1.0f;
    return !(u > (double)alpha);
  }
}

// next[i+1] = first j > i that the chain would accept when its state is the score of row i (i = -1: carried state).
// Each thread first looks at the kShortScan rows after its own (the usual gap is one or two rows); rows that are still
// undecided after that - a sticky state whose next move is far away - are finished one at a time by the whole warp,
// 32 candidates per step with a ballot, so a long gap costs gap/32 steps and no lane waits on a serial tail.
constexpr int kShortScan = 8;
__global__ void mh_next_kernel(const void* sig, int dtype, int64_t n, const double* __restrict__ uniforms,
                               uint64_t seed, uint64_t offset, const double* d_curr, const int* d_kind,
                               int* __restrict__ next) {
  const int kind = *d_kind;
  const int ushift = (kind == 0) ? 1 : 0;     // with d_curr None the first row draws no uniform (idpsampler.py:47)
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = (n + 1 + stride - 1) / stride;       // the same trip count for every lane of a warp
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t r = 0; r < rounds; ++r, t += stride) {
    const bool active = t <= n;
    const int64_t i = t - 1;
    int64_t j = n;
    bool found = !active;
    double d_state = 0.5;
    bool f64_math = false;
    if (active) {
      if (i < 0 && kind == 0) {
        j = 0;                                 // unconditional first move
        found = true;
      } else {
        if (i < 0) {
          d_state = *d_curr;
          f64_math = (dtype == CGS_F64) || (kind == 2);
        } else {
          d_state = load_score(sig, dtype, i);
          f64_math = (dtype == CGS_F64);
        }
        const int64_t lim = min(n, i + 1 + kShortScan);
        for (j = i + 1; j < lim; ++j) {
          const int64_t ui = j - ushift;
          const double u = uniforms ? uniforms[ui] : philox_uniform_f64(seed, offset + (uint64_t)ui);
          if (mh_move(d_state, f64_math, sig, dtype, j, u)) { found = true; break; }
        }
        if (j >= n) { j = n; found = true; }
      }
    }
    unsigned pending = __ballot_sync(0xffffffffu, !found);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const double ds = __shfl_sync(0xffffffffu, d_state, src);
      const int fm = __shfl_sync(0xffffffffu, (int)f64_math, src);
      const int64_t j0 = __shfl_sync(0xffffffffu, j, src);   // first row not yet examined
      int64_t hit = n;
      for (int64_t base = j0; base < n; base += 32) {
        const int64_t jj = base + lane;
        bool ok = false;
        if (jj < n) {
          const int64_t ui = jj - ushift;
          const double u = uniforms ? uniforms[ui] : philox_uniform_f64(seed, offset + (uint64_t)ui);
          ok = mh_move(ds, fm != 0, sig, dtype, jj, u);
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) { hit = base + (__ffs(m) - 1); break; }
      }
      if (lane == src) j = hit;
    }
    if (active) next[t] = (int)j;
  }
}

// phase A: for every node of a 1024-node segment, the first chain node at or beyond the segment end
__global__ void mh_exit_kernel(const int* __restrict__ next, int64_t n, int* __restrict__ exit_node) {
  __shared__ int nx[kChunk];
  const int64_t base = (int64_t)blockIdx.x * kChunk;
  const int end = (int)min((int64_t)n, base + kChunk);
  for (int t = threadIdx.x; t < kChunk; t += blockDim.x) {
    const int64_t i = base + t;
    nx[t] = i < n ? next[i + 1] : (int)n;
  }
  __syncthreads();
  for (int round = 0; round < 10; ++round) {
    int v[kChunk / kBlock];
#pragma unroll
    for (int q = 0; q < kChunk / kBlock; ++q) {
      const int t = threadIdx.x + q * kBlock;
      const int a = nx[t];
      v[q] = a < end ? nx[a - (int)base] : a;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kChunk / kBlock; ++q) nx[threadIdx.x + q * kBlock] = v[q];
    __syncthreads();
  }
  for (int t = threadIdx.x; t < kChunk; t += blockDim.x) {
    const int64_t i = base + t;
    if (i < n) exit_node[i] = nx[t];
  }
}
// phase B: hop segment to segment from the chain start, recording each segment's entry node
__global__ void mh_entries_kernel(const int* __restrict__ next, const int* __restrict__ exit_node, int64_t n,
                                  int* __restrict__ entry, int nseg) {
  for (int i = threadIdx.x; i < nseg; i += blockDim.x) entry[i] = -1;
  __syncthreads();
  if (threadIdx.x == 0) {
    int cur = next[0];
    while (cur < n) {
      entry[cur / kChunk] = cur;
      cur = exit_node[cur];
    }
  }
}
// phase C: walk the chain inside each segment, flagging accepted rows
__global__ void mh_mark_kernel(const int* __restrict__ next, int64_t n, const int* __restrict__ entry,
                               uint8_t* __restrict__ accepted) {
  __shared__ int nx[kChunk];
  __shared__ uint8_t fl[kChunk];
  const int64_t base = (int64_t)blockIdx.x * kChunk;
  const int end = (int)min((int64_t)n, base + kChunk);
  for (int t = threadIdx.x; t < kChunk; t += blockDim.x) {
    const int64_t i = base + t;
    nx[t] = i < n ? next[i + 1] : (int)n;
    fl[t] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int cur = entry[blockIdx.x];
    while (cur >= 0 && cur < end) {
      fl[cur - (int)base] = 1;
      cur = nx[cur - (int)base];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < kChunk; t += blockDim.x) {
    const int64_t i = base + t;
    if (i < n) accepted[i] = fl[t];
  }
}

// thinning + state update (idpsampler.py:29-39).  acc_idx = ascending accepted rows, n_acc their count.
// plan[0] = first emission position, plan[1] = number of emissions.
__global__ void mh_plan_kernel(const void* sig, int dtype, int64_t n, const int* __restrict__ acc_idx,
                               const int* __restrict__ n_acc, int thin, int burn_in, double* d_curr, int* d_kind,
                               int* cnt_chain, int* plan, int* count_out) {
  const int na = *n_acc;
  const int c0 = *cnt_chain;
  int first_emit = -1, n_emit = 0;
  if (na > burn_in) {
    const int f = acc_idx[burn_in];                       // first row at which curr_sample exists
    const int e = c0 > thin ? 0 : thin + 1 - c0;
    first_emit = f + e;
    int c_end;
    if ((int64_t)first_emit <= n - 1) {
      n_emit = (int)((n - 1 - first_emit) / (thin + 1)) + 1;
      const int64_t last = (int64_t)first_emit + (int64_t)(n_emit - 1) * (thin + 1);
      c_end = 1 + (int)(n - 1 - last);
    } else {
      first_emit = -1;
      c_end = c0 + (int)(n - f);
    }
    *cnt_chain = c_end;
  }
  if (na > 0) {
    *d_curr = load_score(sig, dtype, acc_idx[na - 1]);    // idpsampler.py:52
    *d_kind = dtype == CGS_F64 ? 2 : 1;
  }
  plan[0] = first_emit;
  plan[1] = n_emit;
  *count_out = n_emit;
}
__global__ void mh_emit_kernel(const int* __restrict__ acc_idx, const int* __restrict__ n_acc,
                               const int* __restrict__ plan, int thin, int* __restrict__ emit_src) {
  const int n_emit = plan[1];
  const int first = plan[0];
  const int na = *n_acc;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_emit; e += gridDim.x * blockDim.x) {
    const int pos = first + e * (thin + 1);
    // last accepted row <= pos  (upper_bound - 1)
    int lo = 0, hi = na;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (acc_idx[mid] <= pos) lo = mid + 1; else hi = mid;
    }
    emit_src[e] = acc_idx[lo - 1];
  }
}

// ------------------------------------------------------------------------------------------------ gather
__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, int64_t row_bytes, const int* __restrict__ idx,
                                   const int* __restrict__ count, int64_t max_rows, uint8_t* __restrict__ dst) {
  const int64_t rows = min((int64_t)*count, max_rows);
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const uint8_t* s = src + (int64_t)idx[r] * row_bytes;
    uint8_t* d = dst + r * row_bytes;
    if ((row_bytes & 15) == 0 && ((uintptr_t)s & 15) == 0 && ((uintptr_t)d & 15) == 0) {
      const int4* s4 = reinterpret_cast<const int4*>(s);
      int4* d4 = reinterpret_cast<int4*>(d);
      for (int64_t i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) d4[i] = __ldg(s4 + i);
    } else if ((row_bytes & 3) == 0 && ((uintptr_t)s & 3) == 0 && ((uintptr_t)d & 3) == 0) {
      const int* s1 = reinterpret_cast<const int*>(s);
      int* d1 = reinterpret_cast<int*>(d);
      for (int64_t i = threadIdx.x; i < row_bytes / 4; i += blockDim.x) d1[i] = __ldg(s1 + i);
    } else {
      for (int64_t i = threadIdx.x; i < row_bytes; i += blockDim.x) d[i] = s[i];
    }
  }
}

inline size_t al(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

// flags [n] -> ascending indices of the set flags + their count (device); block_counts needs ceil(n/1024)+1 ints
int compact_flags(const uint8_t* flags, int64_t n, int* block_counts, int* idx_out, int* count_out, cudaStream_t st) {
  const int nb = (int)((n + kChunk - 1) / kChunk);
  compact_count_kernel<<<nb, kBlock, 0, st>>>(flags, n, block_counts); count_launch();
  compact_scan_kernel<<<1, 1024, 0, st>>>(block_counts, nb, count_out); count_launch();
  compact_scatter_kernel<<<nb, kBlock, 0, st>>>(flags, n, block_counts, idx_out); count_launch();
  return check_launch("compact_flags");
}


int gather_rows(const void* src, int64_t row_bytes, const int* idx, const int* count, int64_t max_rows, void* dst,
                cudaStream_t st) {
  if (max_rows <= 0 || row_bytes <= 0) return CGS_OK;
  int grid = (int)(max_rows < 148 * 16 ? max_rows : 148 * 16);
  gather_rows_kernel<<<grid, row_bytes >= 4096 ? 256 : 64, 0, st>>>((const uint8_t*)src, row_bytes, idx, count, max_rows,
                                                                      (uint8_t*)dst); count_launch();
  return check_launch("gather_rows");
}

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_policy_step(const cgs_policy_cfg* cfg, float* theta, const float* grad, const float* loss,
                               float* momentum, float* mean_square, float* loss_avg, int first, int64_t rows,
                               int64_t cols, cgs_stream_t stream) {
  if (int rc = require_sm100()) return rc;
  if (!cfg || !theta || !grad) return set_error(CGS_ERR_INVALID, "null argument");
  if (rows < 0 || cols <= 0) return set_error(CGS_ERR_INVALID, "bad shape");
  if (cfg->method < CGS_POLICY_SGD || cfg->method > CGS_POLICY_LADAM)
    return set_error(CGS_ERR_UNSUPPORTED, "unknown policy method %d (sampling/policy.py:64)", cfg->method);
  if (cfg->method == CGS_POLICY_MOMENTUM && !momentum) return set_error(CGS_ERR_INVALID, "momentum buffer required");
  if (cfg->method == CGS_POLICY_LADAM && (!momentum || !mean_square || !loss_avg || !loss))
    return set_error(CGS_ERR_INVALID, "ladam needs momentum, mean_square, loss_avg and loss (sampling/policy.py:51)");
  if (cfg->method == CGS_POLICY_LADAM && cfg->degree != 2)
    return set_error(CGS_ERR_UNSUPPORTED, "ladam degree %d (only degree_ = 2, sampling/policy.py:16)", cfg->degree);
  if (rows == 0) return CGS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const PolicyConsts c = make_policy_consts(*cfg);
  if (cfg->method == CGS_POLICY_LADAM) {
    policy_loss_avg_kernel<<<grid_for(rows), kBlock, 0, st>>>(c, loss, loss_avg, first, rows); count_launch();
  }
  policy_step_kernel<<<grid_for(rows * cols), kBlock, 0, st>>>(c, theta, grad, momentum, mean_square, loss_avg, first,
                                                               rows, cols, cols > 2 ? 1 : 0); count_launch();
  return check_launch("cgs_policy_step");
}

// workspace layout (DRS): F/D_tilde [n] f64 | block maxima [grid] f64 | gamma f64 | sel_lo[2], sel_hi[2] u64 |
// hist[256] u32 | block counts [ceil(n/1024)] i32
extern "C" size_t cgs_drs_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return al((size_t)n * 8) + al(148 * 8 * 8) + al(8) + al(64) + al(1024) + al((size_t)((n + 1023) / 1024 + 1) * 4) + 256;
}

extern "C" int cgs_drs_set_score_max(const void* score_max, int dtype, double* d_tilde_m, cgs_stream_t stream) {
  if (int rc = require_sm100()) return rc;
  if (!score_max || !d_tilde_m) return set_error(CGS_ERR_INVALID, "null argument");
  drs_score_max_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(score_max, dtype, d_tilde_m); count_launch();
  return check_launch("cgs_drs_set_score_max");
}

extern "C" int cgs_drs_accept(const void* sigmoids, int sig_dtype, int64_t n, const double* uniforms,
                              uint64_t philox_seed, uint64_t philox_offset, double* d_tilde_m, double epsilon,
                              double shift_percent, uint8_t* accept_out, int32_t* idx_out, int32_t* count_out,
                              double* prob_out, void* workspace, size_t workspace_bytes, cgs_stream_t stream) {
  if (int rc = require_sm100()) return rc;
  if (n < 0) return set_error(CGS_ERR_INVALID, "negative n");
  if (!d_tilde_m || !count_out) return set_error(CGS_ERR_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    cudaMemsetAsync(count_out, 0, 4, st);
    return CGS_OK;
  }
  if (!sigmoids || !accept_out || !idx_out) return set_error(CGS_ERR_INVALID, "null argument");
  if (sig_dtype != CGS_F32 && sig_dtype != CGS_F64) return set_error(CGS_ERR_INVALID, "bad score dtype");
  if (shift_percent > 100.0) return set_error(CGS_ERR_INVALID, "Percentiles must be in the range [0, 100]");
  if (!workspace || workspace_bytes < cgs_drs_workspace_bytes(n)) return set_error(CGS_ERR_WORKSPACE, "workspace too small");
  char* base = (char*)(((uintptr_t)workspace + 255) & ~uintptr_t(255));
  double* f = (double*)base;                 base += al((size_t)n * 8);
  double* bmax = (double*)base;              base += al(148 * 8 * 8);
  double* gamma = (double*)base;             base += al(8);
  unsigned long long* sel = (unsigned long long*)base;  base += al(64);
  unsigned int* hist = (unsigned int*)base;  base += al(1024);
  int* bcounts = (int*)base;
  const int grid = grid_for(n);
  drs_logit_kernel<<<grid, kBlock, 0, st>>>(sigmoids, sig_dtype, n, f, bmax); count_launch();
  drs_update_max_kernel<<<1, kBlock, 0, st>>>(bmax, grid, d_tilde_m); count_launch();
  drs_f_kernel<<<grid, kBlock, 0, st>>>(f, n, d_tilde_m, epsilon, bmax); count_launch();
  if (shift_percent < 0.0) {
    set_double_kernel<<<1, 1, 0, st>>>(gamma, 0.0); count_launch();                       // shift_percent=None: no shift
  } else if (shift_percent == 100.0) {
    reduce_max_kernel<<<1, kBlock, 0, st>>>(bmax, grid, gamma); count_launch();           // numpy percentile(.,100) == max
  } else {
    // numpy 'linear': virtual index (n-1)*q/100 -> floor / ceil order statistics, lerp
    const double q = shift_percent / 100.0;
    const double vidx = (double)(n - 1) * q;
    double lo_f = floor(vidx);
    const double t = vidx - lo_f;
    int64_t lo = (int64_t)lo_f;
    int64_t hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
    for (int which = 0; which < 2; ++which) {
      unsigned long long* s = sel + 2 * which;
      select_init_kernel<<<1, 256, 0, st>>>(s, (unsigned long long)(which ? hi : lo), hist); count_launch();
      for (int pass = 0; pass < 8; ++pass) {
        select_hist_kernel<<<grid, kBlock, 0, st>>>(f, n, s, pass, hist); count_launch();
        select_pick_kernel<<<1, 256, 0, st>>>(s, pass, hist); count_launch();
      }
    }
    percentile_lerp_kernel<<<1, 1, 0, st>>>(sel, sel + 2, t, gamma); count_launch();
  }
  drs_accept_kernel<<<grid, kBlock, 0, st>>>(f, n, gamma, uniforms, philox_seed, philox_offset, accept_out, prob_out); count_launch();
  if (int rc = check_launch("cgs_drs_accept")) return rc;
  return compact_flags(accept_out, n, bcounts, idx_out, count_out, st);
}

// workspace layout (MH): next [n+1] i32 | exit [n] i32 | entry [nseg] i32 | acc_idx [n] i32 | n_acc i32 | plan[2] i32 |
// block counts [nseg+1] i32
extern "C" size_t cgs_mh_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  const size_t nseg = (size_t)((n + 1023) / 1024 + 1);
  return al((size_t)(n + 1) * 4) + al((size_t)n * 4) + al(nseg * 4) + al((size_t)n * 4) + al(4) + al(8) + al(nseg * 4) + 256;
}

extern "C" int cgs_mh_accept(const void* sigmoids, int sig_dtype, int64_t n, const double* uniforms,
                             uint64_t philox_seed, uint64_t philox_offset, double* d_curr, int32_t* d_kind,
                             int32_t* cnt_chain, int thin_period, int burn_in, uint8_t* accepted_out,
                             int32_t* emit_src_out, int32_t* count_out, void* workspace, size_t workspace_bytes,
                             cgs_stream_t stream) {
  if (int rc = require_sm100()) return rc;
  if (n < 0) return set_error(CGS_ERR_INVALID, "negative n");
  if (!d_curr || !d_kind || !cnt_chain || !count_out) return set_error(CGS_ERR_INVALID, "null argument");
  if (thin_period < 0 || burn_in < 0) return set_error(CGS_ERR_INVALID, "negative thin_period / burn_in");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    cudaMemsetAsync(count_out, 0, 4, st);
    return CGS_OK;
  }
  if (n >= (1ll << 31) - 2) return set_error(CGS_ERR_UNSUPPORTED, "n too large for 32-bit row indices");
  if (!sigmoids || !accepted_out || !emit_src_out) return set_error(CGS_ERR_INVALID, "null argument");
  if (sig_dtype != CGS_F32 && sig_dtype != CGS_F64) return set_error(CGS_ERR_INVALID, "bad score dtype");
  if (!workspace || workspace_bytes < cgs_mh_workspace_bytes(n)) return set_error(CGS_ERR_WORKSPACE, "workspace too small");
  const int nseg = (int)((n + kChunk - 1) / kChunk);
  char* base = (char*)(((uintptr_t)workspace + 255) & ~uintptr_t(255));
  int* next = (int*)base;      base += al((size_t)(n + 1) * 4);
  int* exitn = (int*)base;     base += al((size_t)n * 4);
  int* entry = (int*)base;     base += al((size_t)(nseg + 1) * 4);
  int* acc_idx = (int*)base;   base += al((size_t)n * 4);
  int* n_acc = (int*)base;     base += al(4);
  int* plan = (int*)base;      base += al(8);
  int* bcounts = (int*)base;
  mh_next_kernel<<<grid_for(n + 1, 128, 148 * 16), 128, 0, st>>>(sigmoids, sig_dtype, n, uniforms, philox_seed,
                                                                 philox_offset, d_curr, d_kind, next); count_launch();
  mh_exit_kernel<<<nseg, kBlock, 0, st>>>(next, n, exitn); count_launch();
  mh_entries_kernel<<<1, 256, 0, st>>>(next, exitn, n, entry, nseg); count_launch();
  mh_mark_kernel<<<nseg, kBlock, 0, st>>>(next, n, entry, accepted_out); count_launch();
  if (int rc = check_launch("cgs_mh_accept")) return rc;
  if (int rc = compact_flags(accepted_out, n, bcounts, acc_idx, n_acc, st)) return rc;
  mh_plan_kernel<<<1, 1, 0, st>>>(sigmoids, sig_dtype, n, acc_idx, n_acc, thin_period, burn_in, d_curr, d_kind,
                                  cnt_chain, plan, count_out); count_launch();
  const int64_t max_emit = n / (thin_period + 1) + 1;
  mh_emit_kernel<<<grid_for(max_emit), kBlock, 0, st>>>(acc_idx, n_acc, plan, thin_period, emit_src_out); count_launch();
  return check_launch("cgs_mh_accept");
}

extern "C" int cgs_gather_rows(const void* src, int64_t row_bytes, const int32_t* idx, const int32_t* count,
                               int64_t max_rows, void* dst, cgs_stream_t stream) {
  if (int rc = require_sm100()) return rc;
  if (row_bytes < 0 || max_rows < 0) return set_error(CGS_ERR_INVALID, "negative size");
  if (max_rows == 0 || row_bytes == 0) return CGS_OK;
  if (!src || !idx || !count || !dst) return set_error(CGS_ERR_INVALID, "null argument");
  return gather_rows(src, row_bytes, idx, count, max_rows, dst, (cudaStream_t)stream);
}
