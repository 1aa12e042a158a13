// Shared device helpers for the mineral_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mineral_b200.h"

#define MNR_CHECK_PTR(p) \
  do {                   \
    if ((p) == nullptr) return MNR_ENULL; \
  } while (0)

#define MNR_LAUNCH_RESULT()                  \
  do {                                       \
    cudaError_t e__ = cudaPeekAtLastError(); \
    return e__ == cudaSuccess ? MNR_OK : (int)e__; \
  } while (0)

// deferred folds (include/mineral_b200.h, mnr_fold_defer): the list registered by this host thread, or NULL (api.cu)
mnr_fold_list_t* mnr_fold_target();
// the next free fold slot of the registered list, or NULL when none is registered / the list is full
mnr_fold_job_t* mnr_fold_next();

static inline cudaStream_t mnr_cs(mnr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool mnr_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

__device__ __forceinline__ bool mnr_aligned_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// Kernels of the update chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel of the
// stream may be scheduled - and run its prologue (barrier init, TMEM allocation, descriptor prefetch) - while this one is
// still draining, SM by SM.  Contract for a kernel launched this way: call mnr::pdl_wait() BEFORE its first global-memory
// access (it returns once every preceding grid has completed and its writes are visible; a no-op under a plain launch), and
// mnr::pdl_trigger() as early as it likes (dependents still wait for this grid's completion inside their own pdl_wait).
// Opt-in with MNR_PDL=1 in the environment: on the PPO update graph it measured 1-5 % slower than plain launches (the
// persistent tcgen05 kernels fill an SM's shared memory, so a dependent CTA cannot become resident early anyway, and the
// programmatic graph edges cost more than they hide); the default is plain launches.
bool mnr_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t mnr_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = mnr_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Opt-in ceiling of dynamic shared memory per CTA on sm_100a.  Kernels launched with a shape-dependent amount set the function
// attribute to this CEILING, not to the amount of the launch at hand: the attribute belongs to the function, not to the launch,
// so a later launch with a smaller request would lower it under the feet of an earlier one recorded in a CUDA graph (plain
// graph replays keep working, but a profiler that re-launches graph kernel nodes one by one - ncu - gets LaunchFailed).
constexpr int kMnrMaxDynSmem = 227 * 1024;

// set `kernel`'s dynamic shared-memory ceiling to the most this device allows beside the kernel's static shared memory
// (looked up once per function and host thread)
template <typename K>
static inline cudaError_t mnr_allow_max_dyn_smem(K kernel) {
  static thread_local const void* done[64];
  static thread_local int n_done = 0;
  const void* key = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < n_done; ++i)
    if (done[i] == key) return cudaSuccess;
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMnrMaxDynSmem - (int)fa.sharedSizeBytes);
  if (e == cudaSuccess && n_done < 64) done[n_done++] = key;
  return e;
}

namespace mnr {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int kWarp = 32;

// separately-rounded arithmetic (no FMA contraction) in f32 / f64
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// RunningMeanStd._update_from_moments (mineral/common/normalizers.py:44-55), op for op in the
// state dtype S: batch moments arrive as f32 and are promoted, the count is a Python int.
template <typename S>
__device__ __forceinline__ void chan_update(S& mean, S& var, S count, float bmean, float bvar, double bcount) {
  const S bm = (S)bmean, bv = (S)bvar, bc = (S)bcount;
  const S delta = sub_rn(bm, mean);
  const S tot = add_rn(count, bc);
  const S new_mean = add_rn(mean, div_rn(mul_rn(delta, bc), tot));
  const S m_a = mul_rn(var, count);
  const S m_b = mul_rn(bv, bc);
  const S m2 = add_rn(add_rn(m_a, m_b), div_rn(mul_rn(mul_rn(mul_rn(delta, delta), count), bc), tot));
  mean = new_mean;
  var = div_rn(m2, tot);
}

// ---- (count, mean, M2) moment triple with Chan's parallel merge -----------------------------
struct Moments {
  double n, mean, m2;
};

__device__ __forceinline__ Moments moments_merge(const Moments& a, const Moments& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Moments r;
  r.n = a.n + b.n;
  const double d = b.mean - a.mean;
  const double f = b.n / r.n;
  r.mean = a.mean + d * f;
  r.m2 = a.m2 + b.m2 + d * d * a.n * f;
  return r;
}

__device__ __forceinline__ Moments moments_shfl_down(const Moments& m, int delta) {
  Moments o;
  o.n = __shfl_down_sync(0xffffffffu, m.n, delta);
  o.mean = __shfl_down_sync(0xffffffffu, m.mean, delta);
  o.m2 = __shfl_down_sync(0xffffffffu, m.m2, delta);
  return o;
}

// warp-shuffle Chan merge: lane 0 ends with the merge of all 32 lanes, in a fixed tree order.
__device__ __forceinline__ Moments moments_warp_reduce(Moments m) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = moments_merge(m, moments_shfl_down(m, d));
  return m;
}

// f32 Welford accumulator (one thread's private stream of samples)
struct Welford {
  float n, mean, m2;
  __device__ __forceinline__ void init() { n = 0.f, mean = 0.f, m2 = 0.f; }
  __device__ __forceinline__ void push(float x) {
    n += 1.f;
    const float d = x - mean;
    mean += __fdividef(d, n);
    m2 = fmaf(d, x - mean, m2);
  }
  __device__ __forceinline__ Moments to_moments() const { return Moments{(double)n, (double)mean, (double)m2}; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;
}

// "last block done" ticket: returns true in exactly one CTA (the last to arrive); that CTA may
// then read every other CTA's global partials.  Counter must be zero before the launch and is
// reset to zero by the winner so graphs can replay.
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, unsigned int n_blocks) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == n_blocks - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// Grid-wide barrier among the CTAs of one launch, ALL of which must be resident (grid <= SMs x occupancy).
// `state` = two zero-initialised 64-bit words {arrivals, generation}; the last CTA to arrive resets the arrivals and
// bumps the generation, so the words need no reset between launches / graph replays and any grid size may use them.
// `n_ctas`: how many CTAs take part (default: the whole grid).
__device__ __forceinline__ void grid_barrier(unsigned long long* state, unsigned int n_ctas = 0) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long gen, v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(gen) : "l"(state + 1) : "memory");
    __threadfence();
    if (atomicAdd(state, 1ull) == (unsigned long long)(n_ctas ? n_ctas : gridDim.x) - 1ull) {
      state[0] = 0ull;
      __threadfence();
      asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(state + 1), "l"(gen + 1ull) : "memory");
    } else {
      do {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(state + 1) : "memory");
      } while (v == gen);
    }
  }
  __syncthreads();
}

// CTA-wide sum of one double per thread; the total is returned to thread 0 (others get garbage).
// blockDim.x <= 1024, multiple of 32.  Fixed tree order => deterministic.
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double s_bs[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect s_bs across back-to-back calls
  if (lane == 0) s_bs[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < nw; ++w) t += s_bs[w];
  return t;
}

// same for a Moments triple (Chan merges)
__device__ __forceinline__ Moments block_merge(Moments m) {
  __shared__ Moments s_bm[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  m = moments_warp_reduce(m);
  __syncthreads();
  if (lane == 0) s_bm[warp] = m;
  __syncthreads();
  Moments t{0, 0, 0};
  if (threadIdx.x == 0)
    for (int w = 0; w < nw; ++w) t = moments_merge(t, s_bm[w]);
  return t;
}

// streaming global accessors (read-once / write-once data: keep L1 clean)
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

}  // namespace mnr
