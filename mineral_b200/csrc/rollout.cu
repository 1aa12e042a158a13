// 8(f)-1  rollout-side policy step.  Between two simulator steps the reference runs ~50 small torch ops: the
// actor / critic heads, Dist, sample, log-prob, value un-normalisation (ActorCritic.act, mineral/agents/ppo/models.py:117-133;
// PPO.model_act, mineral/agents/ppo/ppo.py:125-131), five `update_data` copies into the rollout storage (ppo.py:146-150)
// and, after env.step, the reward shaping + time-out bootstrap + two more copies (ppo.py:152-163).  Here the towers run
// on the tcgen05 / SIMT linear kernels (nets/fused_towers.py) and everything behind them is TWO launches that write
// straight into the time-major storage rows of step t.
#include "common.cuh"
#include "loss_math.cuh"

namespace {
using namespace mnr_loss;

constexpr int kPolicyThreads = 256;  // 8 rows per CTA: one warp per env
constexpr int kMaxA = 32, kMaxH = 512;

struct PolicyArgs {
  const float *y3a, *y3v, *w_mu, *b_mu, *w_v, *b_v, *w_sig, *b_sig, *eps;
  const float *value_mean, *value_denom;  // f32 scalars prepared by mnr_rms_prepare, or NULL: values stay normalised
  float clamp_lo, clamp_hi;
  float *actions, *neglogp, *values, *mu, *sigma, *actions_clamped;
  int N, A, H;
};

// warp per env: lanes stride the hidden vector (coalesced), every head output is a shuffle-reduced dot product; lane j then
// owns action column j.
__global__ void __launch_bounds__(kPolicyThreads) policy_step_kernel(const __grid_constant__ PolicyArgs a,
                                                                     const __grid_constant__ mnr_ppo_loss_cfg_t cfg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * (kPolicyThreads / 32) + warp;
  if (n >= a.N) return;
  const int A = a.A, H = a.H;
  const bool learned = a.w_sig != nullptr;
  float ya[kMaxH / 32], yv[kMaxH / 32];
#pragma unroll
  for (int k = 0; k < kMaxH / 32; ++k) {
    const int u = lane + 32 * k;
    ya[k] = u < H ? __ldg(a.y3a + (size_t)n * H + u) : 0.f;
    yv[k] = u < H ? __ldg(a.y3v + (size_t)n * H + u) : 0.f;
  }
  float value = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxH / 32; ++k) {
    const int u = lane + 32 * k;
    if (u < H) value = fmaf(yv[k], __ldg(a.w_v + u), value);
  }
  value = mnr::warp_sum(value);
  value = __shfl_sync(0xffffffffu, value, 0) + __ldg(a.b_v);
  float my_mu_raw = 0.f, my_ls = 0.f;  // lane j keeps output j
  for (int j = 0; j < A; ++j) {
    float pm = 0.f, pl = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxH / 32; ++k) {
      const int u = lane + 32 * k;
      if (u < H) {
        pm = fmaf(ya[k], __ldg(a.w_mu + (size_t)j * H + u), pm);
        if (learned) pl = fmaf(ya[k], __ldg(a.w_sig + (size_t)j * H + u), pl);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      pm += __shfl_xor_sync(0xffffffffu, pm, d);
      if (learned) pl += __shfl_xor_sync(0xffffffffu, pl, d);
    }
    if (lane == j) my_mu_raw = pm, my_ls = pl;
  }
  float nlp = 0.f;
  if (lane < A) {
    const float m = my_mu_raw + __ldg(a.b_mu + lane);
    const float l = learned ? my_ls + __ldg(a.b_sig + lane) : __ldg(a.b_sig + lane);
    float mu, sg, dmu, dsg;
    dist_fwd(cfg, m, l, mu, sg, dmu, dsg);
    // Normal(mu, sigma).sample() == torch.normal(mu, sigma) == mu + sigma * eps, two separately rounded ops
    const float act = __fadd_rn(mu, __fmul_rn(sg, __ldg(a.eps + (size_t)n * A + lane)));
    const float diff = act - mu;
    nlp = diff * diff / (2.f * sg * sg) + logf(sg) + kLogSqrt2Pi;  // -log_prob
    const size_t o = (size_t)n * A + lane;
    a.actions[o] = act, a.mu[o] = mu, a.sigma[o] = sg;
    if (a.actions_clamped) a.actions_clamped[o] = fminf(fmaxf(act, -1.0f), 1.0f);  // ppo.py:151
  }
  nlp = mnr::warp_sum(nlp);
  if (lane == 0) {
    a.neglogp[n] = nlp;
    float v = value;
    if (a.value_mean) {  // RunningMeanStd.unnormalize (normalizers.py:70-73): clamp the INPUT, then scale back
      if (a.clamp_lo <= a.clamp_hi) v = fminf(fmaxf(v, a.clamp_lo), a.clamp_hi);
      v = __fadd_rn(__fmul_rn(v, __ldg(a.value_denom)), __ldg(a.value_mean));
    }
    a.values[n] = v;
  }
}

__global__ void __launch_bounds__(256) post_step_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones,
                                                        const uint8_t* __restrict__ time_outs, const float* __restrict__ values,
                                                        float scale, float gamma, float* __restrict__ rewards_t,
                                                        uint8_t* __restrict__ dones_t, int N) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  float r = __fmul_rn(__ldg(rewards + n), scale);  // RewardShaper 'scale'
  // value bootstrap on time-outs (ppo.py:158-160): += gamma * V(s_t) * time_out, each op rounded separately as torch does
  if (time_outs) r = __fadd_rn(r, __fmul_rn(__fmul_rn(gamma, __ldg(values + n)), time_outs[n] ? 1.0f : 0.0f));
  rewards_t[n] = r;
  dones_t[n] = dones[n] ? 1 : 0;
}

}  // namespace

extern "C" int mnr_policy_step(const float* y3a, const float* y3v, const float* w_mu, const float* b_mu, const float* w_v,
                               const float* b_v, const float* w_sig, const float* b_sig, const float* eps,
                               const float* value_mean_f32, const float* value_denom_f32, float clamp_lo, float clamp_hi,
                               float* actions_t, float* neglogp_t, float* values_t, float* mu_t, float* sigma_t,
                               float* actions_clamped, int N, int A, int H, const mnr_ppo_loss_cfg_t* cfg, mnr_stream_t stream) {
  MNR_CHECK_PTR(y3a); MNR_CHECK_PTR(y3v); MNR_CHECK_PTR(w_mu); MNR_CHECK_PTR(b_mu); MNR_CHECK_PTR(w_v); MNR_CHECK_PTR(b_v);
  MNR_CHECK_PTR(b_sig); MNR_CHECK_PTR(eps); MNR_CHECK_PTR(actions_t); MNR_CHECK_PTR(neglogp_t); MNR_CHECK_PTR(values_t);
  MNR_CHECK_PTR(mu_t); MNR_CHECK_PTR(sigma_t); MNR_CHECK_PTR(cfg);
  if ((value_mean_f32 == nullptr) != (value_denom_f32 == nullptr)) return MNR_ENULL;
  if (N <= 0 || A <= 0 || A > kMaxA || H <= 0 || H > kMaxH) return MNR_ESHAPE;
  if (cfg->dist_type != MNR_DIST_NORMAL && cfg->dist_type != MNR_DIST_DREAMERV3_NORMAL) return MNR_EUNSUP;
  PolicyArgs a{y3a, y3v, w_mu, b_mu, w_v, b_v, w_sig, b_sig, eps, value_mean_f32, value_denom_f32, clamp_lo, clamp_hi,
               actions_t, neglogp_t, values_t, mu_t, sigma_t, actions_clamped, N, A, H};
  const int rows_per_cta = kPolicyThreads / 32;
  policy_step_kernel<<<(N + rows_per_cta - 1) / rows_per_cta, kPolicyThreads, 0, mnr_cs(stream)>>>(a, *cfg);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_post_step(const float* rewards, const void* dones, const void* time_outs, const float* values_t,
                             float reward_scale, float gamma, float* rewards_t, uint8_t* dones_t, int N, mnr_stream_t stream) {
  MNR_CHECK_PTR(rewards); MNR_CHECK_PTR(dones); MNR_CHECK_PTR(rewards_t); MNR_CHECK_PTR(dones_t);
  if (time_outs) MNR_CHECK_PTR(values_t);
  if (N <= 0) return MNR_ESHAPE;
  post_step_kernel<<<(N + 255) / 256, 256, 0, mnr_cs(stream)>>>(rewards, (const uint8_t*)dones, (const uint8_t*)time_outs, values_t,
                                                               reward_scale, gamma, rewards_t, dones_t, N);
  MNR_LAUNCH_RESULT();
}

// ---- 8(f)-4  episode accounting on the device (mineral/common/metrics.py:59-67, tracker.py:9-22) ----------------------
// current_rewards += r; current_lengths += 1; the finished envs (in env order, as the reference's boolean indexing
// yields them) push (reward, length) into the two windows of the last L finished episodes (a deque(maxlen=L): ring
// slot = running episode count mod L) and reset.  ONE CTA: block-wide exclusive scan of the done flags gives every
// finished env its deterministic slot, so the per-step `torch.where(dones)[0].tolist()` host sync of the reference
// (ppo.py:165) disappears from the rollout.
namespace {
constexpr int kEpiThreads = 1024;
__global__ void __launch_bounds__(kEpiThreads) episode_scores_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones,
                                                                     float* __restrict__ cur_rewards, int64_t* __restrict__ cur_lengths,
                                                                     float* __restrict__ win_rewards, float* __restrict__ win_lengths,
                                                                     int64_t* __restrict__ counters /* [0] episodes so far */, int N, int L) {
  __shared__ int s_warp[kEpiThreads / 32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  const int64_t total0 = counters[0];
  // pass 1: count the finished envs (needed to drop all but the last L of them when more than L finish at once)
  int mine = 0;
  for (int n = tid; n < N; n += kEpiThreads) mine += dones[n] ? 1 : 0;
  mine = (int)mnr::warp_sum((float)mine);
  if (lane == 0) s_warp[warp] = mine;
  __syncthreads();
  int n_done = 0;
  for (int w = 0; w < kEpiThreads / 32; ++w) n_done += s_warp[w];
  __syncthreads();
  // pass 2: chunk by chunk in env order, exclusive scan inside the chunk
  for (int c0 = 0; c0 < N; c0 += kEpiThreads) {
    const int n = c0 + tid;
    const bool in = n < N;
    float r = 0.f;
    int64_t len = 0;
    bool d = false;
    if (in) {
      r = cur_rewards[n] + __ldg(rewards + n);
      len = cur_lengths[n] + 1;
      d = dones[n] != 0;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, d);
    const int in_warp = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int rank = before + in_warp;   // how many finished envs precede this one in this step
    if (in) {
      if (d) {
        if (rank >= n_done - L) {   // a deque(maxlen=L) keeps only the last L pushed
          const int slot = (int)((total0 + rank) % L);
          win_rewards[slot] = r, win_lengths[slot] = (float)len;
        }
        r = 0.f, len = 0;
      }
      cur_rewards[n] = r, cur_lengths[n] = len;
    }
    __syncthreads();
    if (tid == 0) {
      int t = s_base;
      for (int w = 0; w < kEpiThreads / 32; ++w) t += s_warp[w];
      s_base = t;
    }
    __syncthreads();
  }
  if (tid == 0) counters[0] = total0 + n_done;
}
}  // namespace

extern "C" int mnr_episode_scores(const float* rewards, const void* dones, float* cur_rewards, int64_t* cur_lengths,
                                  float* win_rewards, float* win_lengths, int64_t* counters, int N, int window,
                                  mnr_stream_t stream) {
  MNR_CHECK_PTR(rewards); MNR_CHECK_PTR(dones); MNR_CHECK_PTR(cur_rewards); MNR_CHECK_PTR(cur_lengths);
  MNR_CHECK_PTR(win_rewards); MNR_CHECK_PTR(win_lengths); MNR_CHECK_PTR(counters);
  if (N <= 0 || window <= 0) return MNR_ESHAPE;
  episode_scores_kernel<<<1, kEpiThreads, 0, mnr_cs(stream)>>>(rewards, (const uint8_t*)dones, cur_rewards, cur_lengths, win_rewards,
                                                               win_lengths, counters, N, window);
  MNR_LAUNCH_RESULT();
}
