// K2  RunningMeanStd: Welford column moments (+ Chan merges, warp-shuffle / smem tree), the
// running-state merge, the pre-computed normaliser schedule of one PPO update, and the fused
// normalise(+clamp) / unnormalise.  Replaces mineral/common/normalizers.py:44-78.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 256;           // columns per CTA (gridDim.z chunks when cols > 256)
constexpr int kElemsPerThread = 64;   // slab sizing (4 rounds of 16 loads in flight)

struct MomentsPlan {
  int n_groups, slabs, chunks;
  int64_t slab_rows;
};

MomentsPlan plan_moments(int64_t rows, int cols, int64_t group_rows) {
  MomentsPlan p;
  p.n_groups = (int)((rows + group_rows - 1) / group_rows);
  p.chunks = (cols + kChunk - 1) / kChunk;
  const int cw = cols < kChunk ? cols : kChunk;
  int64_t slab_rows = ((int64_t)kThreads * kElemsPerThread) / cw;
  if (slab_rows < 1) slab_rows = 1;
  if (slab_rows > group_rows) slab_rows = group_rows;
  p.slabs = (int)((group_rows + slab_rows - 1) / slab_rows);
  // no more CTAs than ~4 per SM over all (group, chunk) pairs: every extra slab is one more partial for the final merge
  const int64_t max_slabs = ((int64_t)mnr_sm_count() * 4 + (int64_t)p.n_groups * p.chunks - 1) / ((int64_t)p.n_groups * p.chunks);
  if (p.slabs > max_slabs) {
    p.slabs = (int)(max_slabs < 1 ? 1 : max_slabs);
    slab_rows = (group_rows + p.slabs - 1) / p.slabs;
    p.slabs = (int)((group_rows + slab_rows - 1) / slab_rows);
  }
  p.slab_rows = slab_rows;
  return p;
}

// x (rows, cols); CTA (slab, group, chunk).  Thread (r_local, c) strides over the slab's rows; for
// cols <= 256 consecutive threads read consecutive addresses (r_local*cols + c == tid).
__global__ void __launch_bounds__(kThreads) batch_moments_kernel(const float* __restrict__ x, int64_t rows, int cols,
                                                                 int64_t group_rows, int64_t slab_rows, int slabs,
                                                                 int correction, float* __restrict__ mean_out,
                                                                 float* __restrict__ var_out,
                                                                 mnr::Moments* __restrict__ partials,
                                                                 unsigned int* __restrict__ tickets) {
  __shared__ mnr::Moments sm[kThreads];
  const int slab = blockIdx.x, group = blockIdx.y, chunk = blockIdx.z;
  const int col0 = chunk * kChunk;
  const int cw = min(kChunk, cols - col0);
  const int rpar = kThreads / cw;  // rows processed side by side
  const int tid = threadIdx.x;
  const int r_local = tid / cw, c = tid - r_local * cw;
  const bool active = r_local < rpar;

  const int64_t g_begin = (int64_t)group * group_rows;
  const int64_t g_end = min(rows, g_begin + group_rows);
  const int64_t r_begin = g_begin + (int64_t)slab * slab_rows;
  const int64_t r_end = min(g_end, r_begin + slab_rows);

  mnr::Welford w;
  w.init();
  if (active) {
    const float* p = x + col0 + c;
    int64_t r = r_begin + r_local;
    // 16 independent loads in flight per thread: with 4 the kernel sat at 5 % of the HBM peak, bound by
    // bytes-in-flight (Little's law), not by the Welford arithmetic
    for (; r + 15 * (int64_t)rpar < r_end; r += 16 * (int64_t)rpar) {
      float a[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) a[u] = mnr::ld_stream(p + (r + (int64_t)u * rpar) * cols);
#pragma unroll
      for (int u = 0; u < 16; ++u) w.push(a[u]);
    }
    for (; r + 3 * (int64_t)rpar < r_end; r += 4 * (int64_t)rpar) {
      const float a0 = mnr::ld_stream(p + r * cols), a1 = mnr::ld_stream(p + (r + rpar) * cols);
      const float a2 = mnr::ld_stream(p + (r + 2 * rpar) * cols), a3 = mnr::ld_stream(p + (r + 3 * rpar) * cols);
      w.push(a0), w.push(a1), w.push(a2), w.push(a3);
    }
    for (; r < r_end; r += rpar) w.push(mnr::ld_stream(p + r * cols));
  }
  sm[tid] = active ? w.to_moments() : mnr::Moments{0, 0, 0};
  __syncthreads();
  // tree over r_local (fixed order): entry (r, c) absorbs (r + s, c)
  int top = 1;
  while (top < rpar) top <<= 1;
  for (int s = top >> 1; s > 0; s >>= 1) {
    if (active && r_local < s && r_local + s < rpar) sm[tid] = mnr::moments_merge(sm[tid], sm[tid + s * cw]);
    __syncthreads();
  }
  const size_t pbase = ((size_t)group * slabs + slab) * cols + col0;
  if (tid < cw) partials[pbase + tid] = sm[tid];

  unsigned int* ticket = tickets + (size_t)group * gridDim.z + chunk;
  if (!mnr::last_block_ticket(ticket, (unsigned int)slabs)) return;
  // last CTA of the (group, chunk): merge the slabs' partials.  Thread (r_local, c) takes slabs r_local, r_local + rpar, ..
  // (independent loads, short dependent chain of f64 Chan merges), then the same fixed-order tree over r_local as above:
  // a single thread per column walking all slabs was a serial chain of `slabs` L2 round trips + f64 divisions and took
  // most of the kernel's time.
  mnr::Moments m{0, 0, 0};
  if (active) {
    const mnr::Moments* pp = partials + (size_t)group * slabs * cols + col0 + c;
    int s = r_local;
    for (; s + 3 * rpar < slabs; s += 4 * rpar) {  // 4 partials in flight, merged in slab order
      const mnr::Moments p0 = pp[(size_t)s * cols], p1 = pp[(size_t)(s + rpar) * cols];
      const mnr::Moments p2 = pp[(size_t)(s + 2 * rpar) * cols], p3 = pp[(size_t)(s + 3 * rpar) * cols];
      m = mnr::moments_merge(mnr::moments_merge(mnr::moments_merge(mnr::moments_merge(m, p0), p1), p2), p3);
    }
    for (; s < slabs; s += rpar) m = mnr::moments_merge(m, pp[(size_t)s * cols]);
  }
  __syncthreads();
  sm[tid] = m;
  __syncthreads();
  for (int s = top >> 1; s > 0; s >>= 1) {
    if (active && r_local < s && r_local + s < rpar) sm[tid] = mnr::moments_merge(sm[tid], sm[tid + s * cw]);
    __syncthreads();
  }
  if (tid < cw) {
    const mnr::Moments t = sm[tid];
    mean_out[(size_t)group * cols + col0 + tid] = (float)t.mean;
    var_out[(size_t)group * cols + col0 + tid] = (float)(t.m2 / (t.n - (double)correction));
  }
}

template <typename S>
__global__ void __launch_bounds__(kThreads) rms_merge_kernel(S* __restrict__ mean, S* __restrict__ var,
                                                             S* __restrict__ count, const float* __restrict__ bmean,
                                                             const float* __restrict__ bvar, double bcount, int cols) {
  const S c = *count;
  __syncthreads();
  for (int i = threadIdx.x; i < cols; i += kThreads) {
    S m = mean[i], v = var[i];
    mnr::chan_update<S>(m, v, c, bmean[i], bvar[i], bcount);
    mean[i] = m, var[i] = v;
  }
  if (threadIdx.x == 0) *count = mnr::add_rn(c, (S)bcount);
}

// The whole obs-RMS trajectory of one PPO update (ppo.py:258-262 executed n_steps times) as one
// sequential-in-steps, parallel-in-columns kernel.
template <typename S>
__global__ void __launch_bounds__(kThreads) rms_schedule_kernel(S* __restrict__ mean, S* __restrict__ var,
                                                                S* __restrict__ count,
                                                                const float* __restrict__ group_mean,
                                                                const float* __restrict__ group_var, int64_t group_rows,
                                                                int64_t rows, int n_groups, int n_steps, int cols,
                                                                float eps, float* __restrict__ mean_tab,
                                                                float* __restrict__ denom_tab) {
  const S c0 = *count;
  __syncthreads();
  S c = c0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < cols; i += gridDim.x * kThreads) {
    S m = mean[i], v = var[i];
    c = c0;
    for (int s = 0; s < n_steps; ++s) {
      const int g = s % n_groups;
      const double bc = (double)min(group_rows, rows - (int64_t)g * group_rows);
      mnr::chan_update<S>(m, v, c, __ldg(group_mean + (size_t)g * cols + i), __ldg(group_var + (size_t)g * cols + i), bc);
      c = mnr::add_rn(c, (S)bc);
      mean_tab[(size_t)s * cols + i] = (float)m;
      denom_tab[(size_t)s * cols + i] = __fsqrt_rn(__fadd_rn((float)v, eps));
    }
    mean[i] = m, var[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    c = c0;
    for (int s = 0; s < n_steps; ++s) {
      const int g = s % n_groups;
      c = mnr::add_rn(c, (S)(double)min(group_rows, rows - (int64_t)g * group_rows));
    }
    *count = c;
  }
}

template <typename S>
__global__ void rms_prepare_kernel(const S* __restrict__ mean, const S* __restrict__ var, int cols, float eps,
                                   float* __restrict__ mean_f32, float* __restrict__ denom_f32) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cols) return;
  mean_f32[i] = (float)mean[i];
  denom_f32[i] = __fsqrt_rn(__fadd_rn((float)var[i], eps));
}

__device__ __forceinline__ float apply1(float x, float m, float d, float lo, float hi, int inverse) {
  if (inverse) {  // normalizers.py:70-73: clamp the input, then sqrt(var+eps) * x + mean
    if (lo <= hi) x = fminf(fmaxf(x, lo), hi);
    return __fadd_rn(__fmul_rn(d, x), m);
  }
  float y = __fdiv_rn(__fsub_rn(x, m), d);  // normalizers.py:65-68
  if (lo <= hi) y = fminf(fmaxf(y, lo), hi);
  return y;
}

// CTA = (row band) x all columns; a thread keeps its column fixed when cols divides the stride.
__global__ void __launch_bounds__(kThreads) rms_apply_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                             int64_t total, int cols, const float* __restrict__ mean,
                                                             const float* __restrict__ denom, float lo, float hi,
                                                             int inverse) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % cols);
  const int step_c = (int)(stride % cols);
  for (; i < total; i += stride) {
    y[i] = apply1(mnr::ld_stream(x + i), __ldg(mean + c), __ldg(denom + c), lo, hi, inverse);
    c += step_c;
    if (c >= cols) c -= cols;
  }
}

// cols % 4 == 0 and 16 B aligned: float4 path
__global__ void __launch_bounds__(kThreads) rms_apply_vec4_kernel(const float4* __restrict__ x, float4* __restrict__ y,
                                                                  int64_t total4, int cols4,
                                                                  const float4* __restrict__ mean,
                                                                  const float4* __restrict__ denom, float lo, float hi,
                                                                  int inverse) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total4) return;
  int c = (int)(i % cols4);
  const int step_c = (int)(stride % cols4);
  for (; i < total4; i += stride) {
    const float4 v = mnr::ld_stream(x + i), m = __ldg(mean + c), d = __ldg(denom + c);
    float4 o;
    o.x = apply1(v.x, m.x, d.x, lo, hi, inverse), o.y = apply1(v.y, m.y, d.y, lo, hi, inverse);
    o.z = apply1(v.z, m.z, d.z, lo, hi, inverse), o.w = apply1(v.w, m.w, d.w, lo, hi, inverse);
    y[i] = o;
    c += step_c;
    if (c >= cols4) c -= cols4;
  }
}

// padded-output variant: y row stride y_ld (feeds TMA, which needs 16 B-multiple row strides)
__global__ void __launch_bounds__(kThreads) rms_apply_ld_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                int64_t total, int cols, int y_ld,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ denom, float lo, float hi) {
  mnr::pdl_trigger();   // PDL: see common.cuh
  mnr::pdl_wait();
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    y[r * y_ld + c] = apply1(mnr::ld_stream(x + i), __ldg(mean + c), __ldg(denom + c), lo, hi, 0);
  }
}

int apply_grid(int64_t work) {
  int64_t g = (work + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)mnr_sm_count() * 8;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" size_t mnr_rms_moments_workspace_bytes(int64_t rows, int cols, int64_t group_rows) {
  if (rows <= 0 || cols <= 0 || group_rows <= 0) return 0;
  const MomentsPlan p = plan_moments(rows, cols, group_rows);
  const size_t tickets = ((size_t)p.n_groups * p.chunks * sizeof(unsigned int) + 255) / 256 * 256;
  return tickets + (size_t)p.n_groups * p.slabs * cols * sizeof(mnr::Moments);
}

extern "C" int mnr_rms_batch_moments_f32(const float* x, int64_t rows, int cols, int64_t group_rows, int correction,
                                         float* mean_out, float* var_out, void* workspace, size_t workspace_bytes,
                                         mnr_stream_t stream) {
  MNR_CHECK_PTR(x); MNR_CHECK_PTR(mean_out); MNR_CHECK_PTR(var_out); MNR_CHECK_PTR(workspace);
  if (rows <= 0 || cols <= 0 || group_rows <= 0) return MNR_ESHAPE;
  if (workspace_bytes < mnr_rms_moments_workspace_bytes(rows, cols, group_rows)) return MNR_EWORKSPACE;
  if (!mnr_aligned(workspace, 16)) return MNR_EALIGN;
  const MomentsPlan p = plan_moments(rows, cols, group_rows);
  if (p.n_groups > 65535 || p.chunks > 65535) return MNR_ESHAPE;
  const size_t tickets = ((size_t)p.n_groups * p.chunks * sizeof(unsigned int) + 255) / 256 * 256;
  // tickets: zero on entry (caller zero-fills a fresh workspace once), reset to zero by the winner CTA
  dim3 grid(p.slabs, p.n_groups, p.chunks);
  batch_moments_kernel<<<grid, kThreads, 0, mnr_cs(stream)>>>(
      x, rows, cols, group_rows, p.slab_rows, p.slabs, correction, mean_out, var_out,
      reinterpret_cast<mnr::Moments*>((unsigned char*)workspace + tickets), reinterpret_cast<unsigned int*>(workspace));
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_rms_merge(void* mean, void* var, void* count, int state_f64, const float* batch_mean,
                             const float* batch_var, double batch_count, int cols, mnr_stream_t stream) {
  MNR_CHECK_PTR(mean); MNR_CHECK_PTR(var); MNR_CHECK_PTR(count); MNR_CHECK_PTR(batch_mean); MNR_CHECK_PTR(batch_var);
  if (cols <= 0) return MNR_ESHAPE;
  if (state_f64)
    rms_merge_kernel<double><<<1, kThreads, 0, mnr_cs(stream)>>>((double*)mean, (double*)var, (double*)count, batch_mean,
                                                                 batch_var, batch_count, cols);
  else
    rms_merge_kernel<float><<<1, kThreads, 0, mnr_cs(stream)>>>((float*)mean, (float*)var, (float*)count, batch_mean,
                                                                batch_var, batch_count, cols);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_rms_schedule(void* mean, void* var, void* count, int state_f64, const float* group_mean,
                                const float* group_var, int64_t group_rows, int64_t rows, int n_steps, int cols,
                                float eps, float* mean_tab, float* denom_tab, mnr_stream_t stream) {
  MNR_CHECK_PTR(mean); MNR_CHECK_PTR(var); MNR_CHECK_PTR(count); MNR_CHECK_PTR(group_mean); MNR_CHECK_PTR(group_var);
  MNR_CHECK_PTR(mean_tab); MNR_CHECK_PTR(denom_tab);
  if (cols <= 0 || n_steps <= 0 || rows <= 0 || group_rows <= 0) return MNR_ESHAPE;
  const int n_groups = (int)((rows + group_rows - 1) / group_rows);
  const int grid = 1;  // single CTA: every thread must read *count before thread 0 rewrites it
  if (state_f64)
    rms_schedule_kernel<double><<<grid, kThreads, 0, mnr_cs(stream)>>>((double*)mean, (double*)var, (double*)count,
                                                                      group_mean, group_var, group_rows, rows, n_groups,
                                                                      n_steps, cols, eps, mean_tab, denom_tab);
  else
    rms_schedule_kernel<float><<<grid, kThreads, 0, mnr_cs(stream)>>>((float*)mean, (float*)var, (float*)count,
                                                                     group_mean, group_var, group_rows, rows, n_groups,
                                                                     n_steps, cols, eps, mean_tab, denom_tab);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_rms_prepare(const void* mean, const void* var, int state_f64, int cols, float eps, float* mean_f32,
                               float* denom_f32, mnr_stream_t stream) {
  MNR_CHECK_PTR(mean); MNR_CHECK_PTR(var); MNR_CHECK_PTR(mean_f32); MNR_CHECK_PTR(denom_f32);
  if (cols <= 0) return MNR_ESHAPE;
  const int grid = (cols + 127) / 128;
  if (state_f64)
    rms_prepare_kernel<double><<<grid, 128, 0, mnr_cs(stream)>>>((const double*)mean, (const double*)var, cols, eps,
                                                                 mean_f32, denom_f32);
  else
    rms_prepare_kernel<float><<<grid, 128, 0, mnr_cs(stream)>>>((const float*)mean, (const float*)var, cols, eps,
                                                                mean_f32, denom_f32);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_rms_apply_f32(const float* x, float* y, int64_t rows, int cols, const float* mean_f32,
                                 const float* denom_f32, float clamp_lo, float clamp_hi, int inverse,
                                 mnr_stream_t stream) {
  MNR_CHECK_PTR(x); MNR_CHECK_PTR(y); MNR_CHECK_PTR(mean_f32); MNR_CHECK_PTR(denom_f32);
  if (rows < 0 || cols <= 0) return MNR_ESHAPE;
  if (rows == 0) return MNR_OK;
  const int64_t total = rows * cols;
  if (cols % 4 == 0 && mnr_aligned(x, 16) && mnr_aligned(y, 16) && mnr_aligned(mean_f32, 16) && mnr_aligned(denom_f32, 16)) {
    rms_apply_vec4_kernel<<<apply_grid(total / 4), kThreads, 0, mnr_cs(stream)>>>(
        (const float4*)x, (float4*)y, total / 4, cols / 4, (const float4*)mean_f32, (const float4*)denom_f32, clamp_lo,
        clamp_hi, inverse);
  } else {
    rms_apply_kernel<<<apply_grid(total), kThreads, 0, mnr_cs(stream)>>>(x, y, total, cols, mean_f32, denom_f32,
                                                                        clamp_lo, clamp_hi, inverse);
  }
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_rms_apply_ld_f32(const float* x, float* y, int64_t rows, int cols, int y_ld, const float* mean_f32,
                                    const float* denom_f32, float clamp_lo, float clamp_hi, mnr_stream_t stream) {
  MNR_CHECK_PTR(x); MNR_CHECK_PTR(y); MNR_CHECK_PTR(mean_f32); MNR_CHECK_PTR(denom_f32);
  if (rows < 0 || cols <= 0 || y_ld < cols) return MNR_ESHAPE;
  if (rows == 0) return MNR_OK;
  const int64_t total = rows * cols;
  const cudaError_t le = mnr_launch_pdl(rms_apply_ld_kernel, dim3(apply_grid(total)), dim3(kThreads), 0, mnr_cs(stream), x, y, total, cols,
                                        y_ld, mean_f32, denom_f32, clamp_lo, clamp_hi);
  if (le != cudaSuccess) return (int)le;
  MNR_LAUNCH_RESULT();
}
