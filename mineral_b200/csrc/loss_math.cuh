// Per-row PPO loss math shared by the loss kernels (csrc/ppo_loss.cu, csrc/heads_fused.cu):
// Dist.forward (mineral/nets/dist.py:31-47), actor_loss / critic_loss / bounds_loss / policy_kl
// (mineral/agents/ppo/ppo.py:449-498) and their derivatives.
#pragma once

#include "common.cuh"

namespace mnr_loss {

constexpr float kLogSqrt2Pi = 0.91893853320467274178f;  // log(sqrt(2*pi))

struct LossPartial {
  double a, c, b, h, clip, kl;
  mnr::Moments d;    // returns - values
  mnr::Moments ret;  // returns
};

__device__ __forceinline__ void dist_fwd(const mnr_ppo_loss_cfg_t& cfg, float mu_raw, float ls, float& mu, float& sigma,
                                         float& dmu, float& dsigma) {
  if (cfg.dist_type == MNR_DIST_NORMAL) {  // dist.py:32-34
    mu = mu_raw, sigma = expf(ls), dmu = 1.f, dsigma = sigma;
  } else {  // dreamerv3_normal, dist.py:42-47
    const float s = 1.f / (1.f + expf(-(ls + 2.0f)));
    const float span = cfg.maxstd - cfg.minstd;
    mu = tanhf(mu_raw);
    sigma = span * s + cfg.minstd;
    dmu = 1.f - mu * mu;
    dsigma = span * s * (1.f - s);
  }
}

__device__ __forceinline__ float smooth_clamp(float x, float lo, float hi, float* dx) {
  // ppo.py:449-450
  const float s = 1.f / (1.f + expf((-(x - lo) / (hi - lo) + 0.5f) * 4.f));
  *dx = 4.f * s * (1.f - s);
  return s * (hi - lo) + lo;
}

// d max(a,b): weight of a (1, 0, or 0.5 on ties as torch.max/maximum backward does)
__device__ __forceinline__ float max_weight(float a, float b) { return a > b ? 1.f : (a < b ? 0.f : 0.5f); }

__device__ __forceinline__ float actor_row(float old_nlp, float nlp, float adv, float e_clip, int smooth, float* dloss_dnlp,
                                           float* clipped) {
  const float ratio = expf(old_nlp - nlp);
  const float lo = 1.0f - e_clip, hi = 1.0f + e_clip;
  float dclamp, cr;
  if (smooth) cr = smooth_clamp(ratio, lo, hi, &dclamp);
  else cr = fminf(fmaxf(ratio, lo), hi), dclamp = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
  const float s1 = -(adv * ratio), s2 = -(adv * cr);
  const float w1 = max_weight(s1, s2);
  const float dl_dratio = -adv * (w1 + (1.f - w1) * dclamp);
  *dloss_dnlp = -dl_dratio * ratio;
  *clipped = fabsf(ratio - 1.0f) > e_clip ? 1.f : 0.f;
  return fmaxf(s1, s2);
}

__device__ __forceinline__ float critic_row(float v_old, float v, float ret, float e_clip, int clip, float* dloss_dv) {
  const float d1 = v - ret;
  const float c1 = d1 * d1;
  if (!clip) {
    *dloss_dv = 2.f * d1;
    return c1;
  }
  const float dv = v - v_old;
  const float d2 = (v_old + fminf(fmaxf(dv, -e_clip), e_clip)) - ret;
  const float c2 = d2 * d2;
  const float w1 = max_weight(c1, c2);
  const float inside = (dv >= -e_clip && dv <= e_clip) ? 1.f : 0.f;
  *dloss_dv = w1 * 2.f * d1 + (1.f - w1) * 2.f * d2 * inside;
  return fmaxf(c1, c2);
}

__device__ __forceinline__ float bounds_elem(float mu, int type, float soft, float* db) {
  if (type == MNR_BOUNDS_REG) {
    *db = 2.f * mu;
    return mu * mu;
  }
  const float hi = fmaxf(mu - soft, 0.f), lo = fminf(mu + soft, 0.f);
  *db = 2.f * hi + 2.f * lo;
  return lo * lo + hi * hi;
}

__device__ __forceinline__ float kl_elem(float mu, float sigma, float mu_old, float sigma_old) {
  // policy_kl(p0 = new, p1 = old), ppo.py:492-497
  const float dm = mu_old - mu;
  return logf(sigma_old / sigma + 1e-5f) + (sigma * sigma + dm * dm) / (2.0f * (sigma_old * sigma_old + 1e-5f)) - 0.5f;
}

}  // namespace mnr_loss
