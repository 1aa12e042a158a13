// K1a GAE and K1b SHAC/SAPO critic-target reverse scans over the horizon, parallel across envs.
// One lane per environment: at every step t a warp reads 32 consecutive envs of the time-major
// (T,N) arrays (one 128 B line per array), so every sector fetched is fully used.  The scan's
// loop-carried dependency is arithmetic only; loads of later steps are independent and are
// batched by the unrolled loop.  All arithmetic uses explicit round-to-nearest intrinsics so
// nvcc cannot contract a*b+c into an FMA: the reference rounds after every op and the result
// must be bit-exact.
#include "common.cuh"

namespace {

constexpr int kScanThreads = 128;
constexpr int kScanChunk = 16;   // horizon steps whose inputs are loaded together

// mineral/agents/ppo/experience.py:96-108
__global__ void __launch_bounds__(kScanThreads) gae_kernel(const float* __restrict__ rewards,
                                                            const float* __restrict__ values,
                                                            const uint8_t* __restrict__ dones,
                                                            const float* __restrict__ last_values,
                                                            float* __restrict__ returns,
                                                            float* __restrict__ advantages, int T, int N,
                                                            float gamma, float gamma_tau) {
  const int n = blockIdx.x * kScanThreads + threadIdx.x;
  if (n >= N) return;
  float next_v = last_values[n];
  float carry = 0.f;
#pragma unroll 8
  for (int t = T - 1; t >= 0; --t) {
    const size_t i = (size_t)t * N + n;
    const float r = __ldg(rewards + i);
    const float v = __ldg(values + i);
    const float alive = __fsub_rn(1.0f, (float)__ldg(dones + i));
    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, next_v), alive)), v);
    carry = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_tau, alive), carry));
    if (advantages) advantages[i] = carry;
    returns[i] = __fadd_rn(carry, v);
    next_v = v;
  }
}

__device__ __forceinline__ float entropy_term(float h, int flags, float abs_t, float inv_abs_t) {
  if (flags & 1) h = -h;                                      // -1.0 * logprobs
  if (flags & 2) h = __fmul_rn(__fadd_rn(h, abs_t), 0.5f);     // (H + |Ht|) * 0.5
  if (flags & 4) h = __fmul_rn(h, inv_abs_t);                  // H * (1/|Ht|)
  return h;
}

// mineral/agents/diffrl/shac.py:877-916
__global__ void __launch_bounds__(kScanThreads) critic_targets_kernel(
    const float* __restrict__ rew, const float* __restrict__ done_mask, const float* __restrict__ next_values,
    const float* __restrict__ entropy, float* __restrict__ target, int T, int N, int method, float gamma,
    float lam, float one_minus_lam, float lam_gamma, float alpha, int ent_flags, float abs_t, float inv_abs_t) {
  const int n = blockIdx.x * kScanThreads + threadIdx.x;
  if (n >= N) return;
  float Ai = 0.f, Bi = 0.f, w = 1.f;
  // The scan is sequential in t but its INPUTS are not: every chunk's loads are issued together (one DRAM round trip per
  // kScanChunk steps instead of one per step: 36.9 -> 12.2 us at cfg3), then the recurrence runs out of registers.  (The GAE
  // kernel above gets the same effect from `#pragma unroll 8`: its body has no branch that stops the hoisting.)
  for (int t_hi = T - 1; t_hi >= 0; t_hi -= kScanChunk) {
    float rr[kScanChunk], nvv[kScanChunk], dd[kScanChunk], hh[kScanChunk];
#pragma unroll
    for (int k = 0; k < kScanChunk; ++k) {
      const int t = t_hi - k;
      if (t >= 0) {
        const size_t i = (size_t)t * N + n;
        rr[k] = __ldg(rew + i), nvv[k] = __ldg(next_values + i);
        dd[k] = method == 0 ? 0.f : __ldg(done_mask + i);
        hh[k] = entropy ? __ldg(entropy + i) : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < kScanChunk; ++k) {
      const int t = t_hi - k;
      if (t < 0) continue;
      const size_t i = (size_t)t * N + n;
      float r = rr[k];
      const float nv = nvv[k];
      if (entropy) r = __fadd_rn(r, __fmul_rn(alpha, entropy_term(hh[k], ent_flags, abs_t, inv_abs_t)));
      if (method == 0) {  // one-step
        target[i] = __fadd_rn(r, __fmul_rn(gamma, nv));
        continue;
      }
      const float d = dd[k];
      const float live = __fsub_rn(1.0f, d);
      w = __fadd_rn(__fmul_rn(__fmul_rn(w, lam), live), d);
      const float adj = __fmul_rn(__fdiv_rn(__fsub_rn(1.0f, w), one_minus_lam), r);
      Ai = __fmul_rn(live, __fadd_rn(__fadd_rn(__fmul_rn(lam_gamma, Ai), __fmul_rn(gamma, nv)), adj));
      Bi = __fadd_rn(__fmul_rn(gamma, __fadd_rn(__fmul_rn(nv, d), __fmul_rn(Bi, live))), r);
      target[i] = __fadd_rn(__fmul_rn(one_minus_lam, Ai), __fmul_rn(w, Bi));
    }
  }
}

}  // namespace

extern "C" int mnr_gae_f32(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                           float* returns, float* advantages, int T, int N, float gamma, float gamma_tau,
                           mnr_stream_t stream) {
  MNR_CHECK_PTR(rewards); MNR_CHECK_PTR(values); MNR_CHECK_PTR(dones); MNR_CHECK_PTR(last_values); MNR_CHECK_PTR(returns);
  if (T <= 0 || N <= 0) return MNR_ESHAPE;
  const int grid = (N + kScanThreads - 1) / kScanThreads;
  gae_kernel<<<grid, kScanThreads, 0, mnr_cs(stream)>>>(rewards, values, dones, last_values, returns, advantages, T, N,
                                                        gamma, gamma_tau);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_critic_targets_f32(const float* rew, const float* done_mask, const float* next_values,
                                      const float* entropy, float* target, int T, int N, int method, float gamma,
                                      float lam, float one_minus_lam, float lam_gamma, float alpha, int ent_flags,
                                      float abs_target_entropy, float inv_abs_target_entropy, mnr_stream_t stream) {
  MNR_CHECK_PTR(rew); MNR_CHECK_PTR(next_values); MNR_CHECK_PTR(target);
  if (T <= 0 || N <= 0) return MNR_ESHAPE;
  if (method != 0 && method != 1) return MNR_EUNSUP;
  if (method == 1) MNR_CHECK_PTR(done_mask);
  const int grid = (N + kScanThreads - 1) / kScanThreads;
  critic_targets_kernel<<<grid, kScanThreads, 0, mnr_cs(stream)>>>(rew, done_mask, next_values, entropy, target, T, N,
                                                                   method, gamma, lam, one_minus_lam, lam_gamma, alpha,
                                                                   ent_flags, abs_target_entropy, inv_abs_target_entropy);
  MNR_LAUNCH_RESULT();
}

// ---- a16  SHAC / SAPO critic regression loss + gradient (mineral/agents/diffrl/shac.py:869-875) ----------------
namespace {
constexpr int kMseThreads = 256;
__global__ void __launch_bounds__(kMseThreads) critic_mse_kernel(const mnr_critic_set_t set, int n_critics, const float* __restrict__ target,
                                                                 int64_t rows, float* __restrict__ loss_out, unsigned char* __restrict__ workspace,
                                                                 int sum_over_critics) {
  // mean over critics (SHAC, shac.py:869-875) or sum over critics (SAC / DDPG, sac.py:303, ddpg.py:296)
  const float n_div = sum_over_critics ? 1.0f : (float)n_critics;
  const float gscale = 2.0f / ((float)rows * n_div);
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kMseThreads + threadIdx.x; i < rows; i += (int64_t)gridDim.x * kMseThreads) {
    const float t = __ldg(target + i);
    float s = 0.f;
    for (int k = 0; k < n_critics; ++k) {
      const float d = __ldg(set.pred[k] + i) - t;
      set.grad[k][i] = gscale * d;
      s = fmaf(d, d, s);
    }
    acc += (double)s;
  }
  double* parts = reinterpret_cast<double*>(workspace + 256);
  const double tot = mnr::block_sum(acc);
  if (threadIdx.x == 0) parts[blockIdx.x] = tot;
  if (!mnr::last_block_ticket(reinterpret_cast<unsigned int*>(workspace), gridDim.x)) return;
  double mine = 0.0;
  for (unsigned int x = threadIdx.x; x < gridDim.x; x += kMseThreads) mine += parts[x];
  const double all = mnr::block_sum(mine);
  if (threadIdx.x == 0) *loss_out = (float)(all / ((double)rows * (double)n_div));
}
int mse_grid(int64_t rows) {
  int64_t g = (rows + kMseThreads - 1) / kMseThreads;
  const int64_t cap = (int64_t)mnr_sm_count() * 4;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}
}  // namespace

extern "C" size_t mnr_critic_mse_workspace_bytes(int64_t rows) {
  (void)rows;
  return 256 + (size_t)mnr_sm_count() * 4 * sizeof(double);
}

static int critic_mse_launch(const mnr_critic_set_t* set, int n_critics, const float* target, int64_t rows, float* loss_out,
                             void* workspace, size_t workspace_bytes, mnr_stream_t stream, int sum_over_critics) {
  MNR_CHECK_PTR(set); MNR_CHECK_PTR(target); MNR_CHECK_PTR(loss_out); MNR_CHECK_PTR(workspace);
  if (n_critics <= 0 || n_critics > MNR_MAX_CRITICS || rows <= 0) return MNR_ESHAPE;
  if (workspace_bytes < mnr_critic_mse_workspace_bytes(rows)) return MNR_EWORKSPACE;
  if (!mnr_aligned(workspace, 8)) return MNR_EALIGN;
  for (int k = 0; k < n_critics; ++k)
    if (!set->pred[k] || !set->grad[k]) return MNR_ENULL;
  critic_mse_kernel<<<mse_grid(rows), kMseThreads, 0, mnr_cs(stream)>>>(*set, n_critics, target, rows, loss_out, (unsigned char*)workspace,
                                                                        sum_over_critics);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_critic_mse_fwd_bwd(const mnr_critic_set_t* set, int n_critics, const float* target, int64_t rows, float* loss_out,
                                      void* workspace, size_t workspace_bytes, mnr_stream_t stream) {
  return critic_mse_launch(set, n_critics, target, rows, loss_out, workspace, workspace_bytes, stream, 0);
}

extern "C" int mnr_critic_mse_sum_fwd_bwd(const mnr_critic_set_t* set, int n_critics, const float* target, int64_t rows, float* loss_out,
                                          void* workspace, size_t workspace_bytes, mnr_stream_t stream) {
  return critic_mse_launch(set, n_critics, target, rows, loss_out, workspace, workspace_bytes, stream, 1);
}
