// 8(f)-2  SAC / DDPG update math that sits on top of the fused replay sample (csrc/rows.cu):
//   * Polyak target-network update          soft_update, mineral/agents/ddpg/utils.py:8-11 (sac.py:271-273, ddpg.py:261-263)
//   * TD target of the critic regression    sac.py:293-297 (entropy-backed twin-Q minimum), ddpg.py:292-293
//   * C51 categorical projection            distl_projection, mineral/agents/ddpg/utils.py:14-31 (ddpg.py:276-290)
// Element-wise / row-wise fp32 with the reference's operation order (separately rounded ops), so the results are
// bit-identical to the reference's CPU torch evaluation.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) soft_update_kernel(float* __restrict__ tar, const float* __restrict__ cur, int64_t n,
                                                               float one_minus_tau, float tau) {
  // tar.mul_(1.0 - tau).add_(cur * tau)
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
    tar[i] = __fadd_rn(__fmul_rn(tar[i], one_minus_tau), __fmul_rn(__ldg(cur + i), tau));
}

struct QSet {
  const float* q[MNR_MAX_CRITICS];
};
struct GSet {
  float* g[MNR_MAX_CRITICS];
};

__global__ void __launch_bounds__(kThreads) td_target_kernel(const QSet qs, int n_q, const float* __restrict__ log_prob,
                                                             const float* __restrict__ reward, const float* __restrict__ done,
                                                             float alpha, float gamma_n, float* __restrict__ out, int64_t rows) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < rows; i += stride) {
    float t = __ldg(qs.q[0] + i);
    for (int k = 1; k < n_q; ++k) t = fminf(t, __ldg(qs.q[k] + i));                    // get_q_min: torch.min over the ensemble
    if (log_prob) t = __fsub_rn(t, __fmul_rn(alpha, __ldg(log_prob + i)));             // target_Q -= alpha * log_prob
    const float nd = __fsub_rn(1.0f, __ldg(done + i));                                  // (1 - done)
    out[i] = __fadd_rn(__ldg(reward + i), __fmul_rn(__fmul_rn(nd, gamma_n), t));        // reward + (1 - done) * gamma**nstep * target_Q
  }
}

constexpr int kMaxAtoms = 128;

// thread per batch row; the two index_add_ passes of the reference (all lower-atom contributions, then all upper-atom
// ones, atoms in ascending order) are replayed in that order into a private row -> bit-identical accumulation.
__global__ void __launch_bounds__(128) distl_projection_kernel(const float* __restrict__ next_dist, const float* __restrict__ reward,
                                                               const float* __restrict__ done, const float* __restrict__ support,
                                                               float gamma, float v_min, float v_max, float delta_z, int n_atoms,
                                                               float* __restrict__ out, int64_t rows) {
  const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (row >= rows) return;
  float proj[kMaxAtoms];
  for (int j = 0; j < n_atoms; ++j) proj[j] = 0.f;
  const float r = __ldg(reward + row);
  const float g = __fmul_rn(__fsub_rn(1.0f, __ldg(done + row)), gamma);  // (1 - done) * gamma
  const float* p = next_dist + row * n_atoms;
  for (int pass = 0; pass < 2; ++pass) {
    for (int j = 0; j < n_atoms; ++j) {
      float tz = __fadd_rn(r, __fmul_rn(g, __ldg(support + j)));
      tz = fminf(fmaxf(tz, v_min), v_max);
      const float b = __fdiv_rn(__fsub_rn(tz, v_min), delta_z);
      int l = (int)floorf(b), u = (int)ceilf(b);
      if (u > 0 && l == u) l -= 1;
      if (l < n_atoms - 1 && l == u) u += 1;
      const float pj = __ldg(p + j);
      if (pass == 0) proj[l] = __fadd_rn(proj[l], __fmul_rn(pj, __fsub_rn((float)u, b)));
      else proj[u] = __fadd_rn(proj[u], __fmul_rn(pj, __fsub_rn(b, (float)l)));
    }
  }
  for (int j = 0; j < n_atoms; ++j) out[row * n_atoms + j] = proj[j];
}

// ---- SAC / DDPG agents (mineral/agents/sac/sac.py, mineral/agents/ddpg/ddpg.py): the row-wise math around the tower GEMMs ----
constexpr int kMaxAct = 32;
constexpr float kHalfLog2Pi = 0.91893853320467274178f;  // log(sqrt(2 pi)), Normal.log_prob
constexpr float kLog2 = 0.69314718055994530942f;

__global__ void __launch_bounds__(kThreads) copy_cols_kernel(const float* __restrict__ src, int src_ld, float* __restrict__ dst, int dst_ld,
                                                             int64_t rows, int cols) {
  const int64_t n = rows * cols, stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    dst[r * dst_ld + c] = __ldg(src + r * src_ld + c);
  }
}

// Dist.forward (mineral/nets/dist.py:31-58) for the squashed distributions: logstd -> sigma and d sigma / d logstd
struct SquashCfg {
  int dist_type;  // 0 squashed_normal, 1 dreamerv3_squashed_normal
  int has_min, has_max;
  float minlogstd, maxlogstd, minstd, maxstd;
};
__device__ __forceinline__ float squash_sigma(const SquashCfg& c, float ls_raw, float* dsig_dls) {
  if (c.dist_type == 0) {
    float ls = ls_raw;
    bool pass = true;  // torch.clamp: gradient 1 inside the closed range, 0 outside
    if (c.has_min && ls < c.minlogstd) ls = c.minlogstd, pass = false;
    if (c.has_max && ls > c.maxlogstd) ls = c.maxlogstd, pass = false;
    const float sig = expf(ls);
    *dsig_dls = pass ? sig : 0.f;
    return sig;
  }
  const float s = 1.0f / (1.0f + expf(-(ls_raw + 2.0f)));
  const float w = c.maxstd - c.minstd;
  *dsig_dls = w * s * (1.0f - s);
  return w * s + c.minstd;
}
__device__ __forceinline__ float softplus_torch(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // F.softplus, threshold 20

// SAC.get_actions (sac.py:119-136): u = mu + sigma * eps (Normal.rsample), a = tanh(u) (TanhTransform),
// log_prob = sum_j [Normal.log_prob(u_j) - 2 (log 2 - u_j - softplus(-2 u_j))] (mineral/nets/distributions.py:10-37 with the
// cached pre-tanh sample).  eps == NULL: the deterministic branch returns the UN-squashed mean (sac.py:129-130).
__global__ void __launch_bounds__(kThreads) squashed_sample_kernel(const float* __restrict__ mu, const float* __restrict__ logstd, int ls_bcast,
                                                                   const float* __restrict__ eps, int64_t rows, int A, SquashCfg cfg,
                                                                   float* __restrict__ act, int act_ld, float* __restrict__ logp) {
  const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (r >= rows) return;
  float lp = 0.f;
  for (int j = 0; j < A; ++j) {
    const float m = __ldg(mu + r * A + j);
    if (!eps) {
      act[r * act_ld + j] = m;
      continue;
    }
    float dsig;
    const float sig = squash_sigma(cfg, __ldg(logstd + (ls_bcast ? j : r * A + j)), &dsig);
    const float u = __fadd_rn(m, __fmul_rn(__ldg(eps + r * A + j), sig));
    act[r * act_ld + j] = tanhf(u);
    const float d = __fsub_rn(u, m);
    const float base = -(d * d) / (2.0f * sig * sig) - logf(sig) - kHalfLog2Pi;
    const float ladj = 2.0f * (kLog2 - u - softplus_torch(-2.0f * u));
    lp += base - ladj;
  }
  if (logp && eps) logp[r] = lp;
}

// gradient of  c_logp * sum_rows log_prob + <g_act, a>  w.r.t. (mu, logstd) through the reparameterised sample:
//   d/du = c_logp * 2 tanh(u) + g_a (1 - tanh(u)^2);  d/dmu = d/du;  d/dsigma = d/du * eps - c_logp / sigma
// (the Normal.log_prob terms through u and through (mu, sigma) cancel to -1/sigma; checked against autograd in the oracle).
// c_logp = alpha / rows with alpha = exp(*log_alpha) (sac.py:323: actor_loss = (alpha * log_prob - Q).mean()) or alpha_host.
__global__ void __launch_bounds__(kThreads) squashed_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ logstd, int ls_bcast,
                                                                const float* __restrict__ eps, const float* __restrict__ g_act,
                                                                const float* __restrict__ g_act2, int g_ld,
                                                                const float* __restrict__ log_alpha, float alpha_host, int64_t rows, int A,
                                                                SquashCfg cfg, float* __restrict__ g_mu, float* __restrict__ g_ls) {
  const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (r >= rows) return;
  const float alpha = log_alpha ? expf(__ldg(log_alpha)) : alpha_host;
  const float c = alpha / (float)rows;
  for (int j = 0; j < A; ++j) {
    const float m = __ldg(mu + r * A + j), e = __ldg(eps + r * A + j);
    float dsig;
    const float sig = squash_sigma(cfg, __ldg(logstd + (ls_bcast ? j : r * A + j)), &dsig);
    const float u = __fadd_rn(m, __fmul_rn(e, sig));
    const float a = tanhf(u);
    float ga = g_act ? __ldg(g_act + r * g_ld + j) : 0.f;
    if (g_act2) ga += __ldg(g_act2 + r * g_ld + j);   // the second critic tower's input gradient (same layout)
    const float gu = c * 2.0f * a + ga * (1.0f - a * a);
    g_mu[r * A + j] = gu;
    g_ls[r * A + j] = (gu * e - c / sig) * dsig;
  }
}

// DDPG actor output (ddpg/models.py:55-58, ddpg.py:98-117,132-147, noise.py:22-55): a = tanh(mu) [+ clamp(noise, +-noise_bound)]
// [clamped to out bounds]; noise (rows, A) is the already-scaled Normal(0, std) draw.
__global__ void __launch_bounds__(kThreads) tanh_action_kernel(const float* __restrict__ mu, const float* __restrict__ noise, int64_t rows, int A,
                                                               float noise_lo, float noise_hi, float out_lo, float out_hi,
                                                               float* __restrict__ act, int act_ld) {
  const int64_t n = rows * A, stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / A;
    const int j = (int)(i - r * A);
    float a = tanhf(__ldg(mu + i));
    if (noise) {
      float z = __ldg(noise + i);
      if (noise_lo <= noise_hi) z = fminf(fmaxf(z, noise_lo), noise_hi);
      a = __fadd_rn(a, z);
      if (out_lo <= out_hi) a = fminf(fmaxf(a, out_lo), out_hi);
    }
    act[r * act_ld + j] = a;
  }
}
__global__ void __launch_bounds__(kThreads) tanh_action_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ g_act,
                                                                   const float* __restrict__ g_act2, int g_ld, int64_t rows, int A,
                                                                   float* __restrict__ g_mu) {
  const int64_t n = rows * A, stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / A;
    const int j = (int)(i - r * A);
    const float a = tanhf(__ldg(mu + i));
    float ga = __ldg(g_act + r * g_ld + j);
    if (g_act2) ga += __ldg(g_act2 + r * g_ld + j);
    g_mu[i] = ga * (1.0f - a * a);
  }
}

constexpr int kOneCta = 1024;
constexpr int kRowChunk = 4;   // rows per thread whose inputs are loaded together in the one-CTA kernels
__device__ __forceinline__ float block_minmax(float v, bool want_max) {
  __shared__ float s_mm[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int d = 16; d > 0; d >>= 1) {
    const float o = __shfl_down_sync(0xffffffffu, v, d);
    v = want_max ? fmaxf(v, o) : fminf(v, o);
  }
  __syncthreads();
  if (lane == 0) s_mm[warp] = v;
  __syncthreads();
  float t = s_mm[0];
  if (threadIdx.x == 0)
    for (int w = 1; w < nw; ++w) t = want_max ? fmaxf(t, s_mm[w]) : fminf(t, s_mm[w]);
  return t;
}

// td target with the temperature read on the DEVICE (alpha = exp(*log_alpha), sac.py:275-285) + the target statistics the
// agent logs (sac.py:306-311: mean, unbiased std, max, min).  ONE CTA: a batch is a few thousand rows; fixed order.
__global__ void __launch_bounds__(kOneCta) td_target_stats_kernel(const QSet qs, int n_q, const float* __restrict__ log_prob,
                                                                  const float* __restrict__ reward, const float* __restrict__ done,
                                                                  const float* __restrict__ log_alpha, float alpha_host, float gamma_n,
                                                                  float* __restrict__ out, int64_t rows, float* __restrict__ stats) {
  const float alpha = log_alpha ? expf(__ldg(log_alpha)) : alpha_host;
  double sum = 0.0;
  float mx = -INFINITY, mn = INFINITY;
  // one CTA, so a thread owns rows/1024 rows: their inputs are loaded kRowChunk rows at a time BEFORE any of them is used
  // (otherwise every row is its own DRAM round trip: 15.6 us for 8192 rows)
  for (int64_t i0 = threadIdx.x; i0 < rows; i0 += (int64_t)kOneCta * kRowChunk) {
    float q[kRowChunk][MNR_MAX_CRITICS], lp[kRowChunk], rw[kRowChunk], dn[kRowChunk];
#pragma unroll
    for (int u = 0; u < kRowChunk; ++u) {
      const int64_t i = i0 + (int64_t)u * kOneCta;
      if (i < rows) {
#pragma unroll
        for (int k = 0; k < MNR_MAX_CRITICS; ++k)
          if (k < n_q) q[u][k] = __ldg(qs.q[k] + i);
        lp[u] = log_prob ? __ldg(log_prob + i) : 0.f;
        rw[u] = __ldg(reward + i), dn[u] = __ldg(done + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kRowChunk; ++u) {
      const int64_t i = i0 + (int64_t)u * kOneCta;
      if (i < rows) {
        float t = q[u][0];
#pragma unroll
        for (int k = 1; k < MNR_MAX_CRITICS; ++k)
          if (k < n_q) t = fminf(t, q[u][k]);
        if (log_prob) t = __fsub_rn(t, __fmul_rn(alpha, lp[u]));
        const float nd = __fsub_rn(1.0f, dn[u]);
        const float y = __fadd_rn(rw[u], __fmul_rn(__fmul_rn(nd, gamma_n), t));
        out[i] = y;
        sum += (double)y;
        mx = fmaxf(mx, y), mn = fminf(mn, y);
      }
    }
  }
  if (!stats) return;
  __shared__ double s_mean;
  const double tot = mnr::block_sum(sum);
  if (threadIdx.x == 0) s_mean = tot / (double)rows;
  __syncthreads();
  const double mean = s_mean;
  double ss = 0.0;
  for (int64_t i = threadIdx.x; i < rows; i += kOneCta) {  // each thread re-reads the rows it wrote itself
    const double d = (double)out[i] - mean;
    ss += d * d;
  }
  const double ss_tot = mnr::block_sum(ss);
  const float mx_all = block_minmax(mx, true), mn_all = block_minmax(mn, false);
  if (threadIdx.x == 0) {
    stats[0] = (float)mean;
    stats[1] = (float)sqrt(ss_tot / (double)(rows - 1));
    stats[2] = mx_all;
    stats[3] = mn_all;
  }
}

// Actor objective on the critics' outputs (sac.py:313-332, ddpg.py:299-310): q_min = min_k Q_k (first minimum on ties, as torch.min),
// actor_loss = mean(alpha * log_prob - q_min) [log_prob NULL: mean(-q_min)], seeds d actor_loss / d Q_k = -1/rows at the argmin,
// entropy = mean(-log_prob), alpha_loss = mean(alpha * (-log_prob - target_entropy)) and its gradient w.r.t. log_alpha
// (= alpha_loss itself, since d exp(x)/dx = exp(x)).  out = {actor_loss, entropy, alpha_loss, alpha}.  ONE CTA, fixed order.
__global__ void __launch_bounds__(kOneCta) actor_seed_kernel(const QSet qs, int n_q, const float* __restrict__ log_prob,
                                                             const float* __restrict__ log_alpha, float alpha_host, float target_entropy,
                                                             int64_t rows, GSet gs, float* __restrict__ out, float* __restrict__ log_alpha_grad) {
  const float alpha = log_alpha ? expf(__ldg(log_alpha)) : alpha_host;
  const float seed = -1.0f / (float)rows;
  double s_loss = 0.0, s_lp = 0.0;
  for (int64_t i0 = threadIdx.x; i0 < rows; i0 += (int64_t)kOneCta * kRowChunk) {   // chunked loads, see td_target_stats_kernel
    float qv[kRowChunk][MNR_MAX_CRITICS], lpv[kRowChunk];
#pragma unroll
    for (int u = 0; u < kRowChunk; ++u) {
      const int64_t i = i0 + (int64_t)u * kOneCta;
      if (i < rows) {
#pragma unroll
        for (int k = 0; k < MNR_MAX_CRITICS; ++k)
          if (k < n_q) qv[u][k] = __ldg(qs.q[k] + i);
        lpv[u] = log_prob ? __ldg(log_prob + i) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kRowChunk; ++u) {
      const int64_t i = i0 + (int64_t)u * kOneCta;
      if (i < rows) {
        float q = qv[u][0];
        int arg = 0;
#pragma unroll
        for (int k = 1; k < MNR_MAX_CRITICS; ++k)
          if (k < n_q && qv[u][k] < q) q = qv[u][k], arg = k;
#pragma unroll
        for (int k = 0; k < MNR_MAX_CRITICS; ++k)
          if (k < n_q) gs.g[k][i] = k == arg ? seed : 0.f;
        s_loss += (double)(alpha * lpv[u] - q);
        s_lp += (double)lpv[u];
      }
    }
  }
  const double t_loss = mnr::block_sum(s_loss);
  const double t_lp = mnr::block_sum(s_lp);
  if (threadIdx.x == 0) {
    const double n = (double)rows;
    const double ent = -t_lp / n;
    const float a_loss = (float)((double)alpha * (ent - (double)target_entropy));
    out[0] = (float)(t_loss / n);
    out[1] = (float)ent;
    out[2] = a_loss;
    out[3] = alpha;
    if (log_alpha_grad) *log_alpha_grad = a_loss;
  }
}

// One env step of SAC/DDPG.explore_env (sac.py:163-177, ddpg.py:176-190) into the (N, T, .) trajectory tensors: obs, actions,
// next_obs, rewards * reward_scale (RewardShaper 'scale', applied by the reference to the whole trajectory afterwards: the same
// single multiply) and dones masked by the time-outs (Agent._handle_timeout, agent.py:158-167: dones * ~time_outs), stored as
// float like the reference's `traj_dones`.  Replaces five strided slice copies + the mask + the shaper.
__global__ void __launch_bounds__(kThreads) traj_store_kernel(const float* __restrict__ obs, const float* __restrict__ next_obs,
                                                              const float* __restrict__ actions, const float* __restrict__ rewards,
                                                              const uint8_t* __restrict__ dones, const uint8_t* __restrict__ time_outs,
                                                              int N, int O, int A, int T, int t, float reward_scale,
                                                              float* __restrict__ traj_obs, float* __restrict__ traj_next_obs,
                                                              float* __restrict__ traj_actions, float* __restrict__ traj_rewards,
                                                              float* __restrict__ traj_dones) {
  const int W = 2 * O + A + 2;
  const int64_t total = (int64_t)N * W, stride = (int64_t)gridDim.x * kThreads;
  for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < total; q += stride) {
    const int64_t n = q / W;
    const int e = (int)(q - n * W);
    const int64_t slot = n * T + t;
    if (e < O) traj_obs[slot * O + e] = __ldg(obs + n * O + e);
    else if (e < 2 * O) traj_next_obs[slot * O + (e - O)] = __ldg(next_obs + n * O + (e - O));
    else if (e < 2 * O + A) traj_actions[slot * A + (e - 2 * O)] = __ldg(actions + n * A + (e - 2 * O));
    else if (e == 2 * O + A) traj_rewards[slot] = __fmul_rn(__ldg(rewards + n), reward_scale);
    else traj_dones[slot] = (__ldg(dones + n) && !(time_outs && __ldg(time_outs + n))) ? 1.0f : 0.0f;
  }
}

int grid_for(int64_t n) {
  int64_t g = (n + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)mnr_sm_count() * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

extern "C" int mnr_soft_update(float* target, const float* current, int64_t n, float one_minus_tau, float tau, mnr_stream_t stream) {
  MNR_CHECK_PTR(target); MNR_CHECK_PTR(current);
  if (n <= 0) return MNR_ESHAPE;
  soft_update_kernel<<<grid_for(n), kThreads, 0, mnr_cs(stream)>>>(target, current, n, one_minus_tau, tau);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_td_target(const mnr_critic_set_t* qs, int n_q, const float* log_prob, const float* reward, const float* done,
                             float alpha, float gamma_n, float* out, int64_t rows, mnr_stream_t stream) {
  MNR_CHECK_PTR(qs); MNR_CHECK_PTR(reward); MNR_CHECK_PTR(done); MNR_CHECK_PTR(out);
  if (n_q <= 0 || n_q > MNR_MAX_CRITICS || rows <= 0) return MNR_ESHAPE;
  QSet s;
  for (int k = 0; k < MNR_MAX_CRITICS; ++k) s.q[k] = k < n_q ? qs->pred[k] : nullptr;
  for (int k = 0; k < n_q; ++k)
    if (!s.q[k]) return MNR_ENULL;
  td_target_kernel<<<grid_for(rows), kThreads, 0, mnr_cs(stream)>>>(s, n_q, log_prob, reward, done, alpha, gamma_n, out, rows);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_distl_projection(const float* next_dist, const float* reward, const float* done, const float* support, float gamma,
                                    float v_min, float v_max, float delta_z, int n_atoms, float* out, int64_t rows,
                                    mnr_stream_t stream) {
  MNR_CHECK_PTR(next_dist); MNR_CHECK_PTR(reward); MNR_CHECK_PTR(done); MNR_CHECK_PTR(support); MNR_CHECK_PTR(out);
  if (rows <= 0 || n_atoms < 2 || n_atoms > kMaxAtoms) return MNR_ESHAPE;
  distl_projection_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, mnr_cs(stream)>>>(next_dist, reward, done, support, gamma, v_min, v_max,
                                                                                      delta_z, n_atoms, out, rows);
  MNR_LAUNCH_RESULT();
}

static int pack_qset(const mnr_critic_set_t* qs, int n_q, QSet* s, GSet* g) {
  if (!qs) return MNR_ENULL;
  if (n_q <= 0 || n_q > MNR_MAX_CRITICS) return MNR_ESHAPE;
  for (int k = 0; k < MNR_MAX_CRITICS; ++k) {
    s->q[k] = k < n_q ? qs->pred[k] : nullptr;
    if (g) g->g[k] = k < n_q ? qs->grad[k] : nullptr;
  }
  for (int k = 0; k < n_q; ++k)
    if (!s->q[k] || (g && !g->g[k])) return MNR_ENULL;
  return MNR_OK;
}

static SquashCfg pack_squash(const mnr_squash_cfg_t* c) {
  SquashCfg s;
  s.dist_type = c->dist_type, s.has_min = c->has_minlogstd, s.has_max = c->has_maxlogstd;
  s.minlogstd = c->minlogstd, s.maxlogstd = c->maxlogstd, s.minstd = c->minstd, s.maxstd = c->maxstd;
  return s;
}

extern "C" int mnr_copy_cols_f32(const float* src, int src_ld, float* dst, int dst_ld, int64_t rows, int cols, mnr_stream_t stream) {
  MNR_CHECK_PTR(src); MNR_CHECK_PTR(dst);
  if (rows <= 0 || cols <= 0 || src_ld < cols || dst_ld < cols) return MNR_ESHAPE;
  copy_cols_kernel<<<grid_for(rows * cols), kThreads, 0, mnr_cs(stream)>>>(src, src_ld, dst, dst_ld, rows, cols);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_squashed_sample(const float* mu, const float* logstd, int logstd_broadcast, const float* eps, int64_t rows, int A,
                                   const mnr_squash_cfg_t* cfg, float* actions, int act_ld, float* log_prob, mnr_stream_t stream) {
  MNR_CHECK_PTR(mu); MNR_CHECK_PTR(cfg); MNR_CHECK_PTR(actions);
  if (eps) MNR_CHECK_PTR(logstd);
  if (rows <= 0 || A <= 0 || A > kMaxAct || act_ld < A) return MNR_ESHAPE;
  if (cfg->dist_type < 0 || cfg->dist_type > 1) return MNR_EUNSUP;
  squashed_sample_kernel<<<(unsigned)((rows + kThreads - 1) / kThreads), kThreads, 0, mnr_cs(stream)>>>(
      mu, logstd, logstd_broadcast, eps, rows, A, pack_squash(cfg), actions, act_ld, log_prob);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_squashed_bwd(const float* mu, const float* logstd, int logstd_broadcast, const float* eps, const float* g_actions,
                                const float* g_actions2, int g_ld, const float* log_alpha, float alpha_host, int64_t rows, int A, const mnr_squash_cfg_t* cfg, float* g_mu,
                                float* g_logstd, mnr_stream_t stream) {
  MNR_CHECK_PTR(mu); MNR_CHECK_PTR(logstd); MNR_CHECK_PTR(eps); MNR_CHECK_PTR(cfg); MNR_CHECK_PTR(g_mu); MNR_CHECK_PTR(g_logstd);
  if (rows <= 0 || A <= 0 || A > kMaxAct || (g_actions && g_ld < A)) return MNR_ESHAPE;
  if (cfg->dist_type < 0 || cfg->dist_type > 1) return MNR_EUNSUP;
  squashed_bwd_kernel<<<(unsigned)((rows + kThreads - 1) / kThreads), kThreads, 0, mnr_cs(stream)>>>(
      mu, logstd, logstd_broadcast, eps, g_actions, g_actions ? g_actions2 : nullptr, g_ld, log_alpha, alpha_host, rows, A, pack_squash(cfg), g_mu,
      g_logstd);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_tanh_action(const float* mu, const float* noise, int64_t rows, int A, float noise_lo, float noise_hi, float out_lo,
                               float out_hi, float* actions, int act_ld, mnr_stream_t stream) {
  MNR_CHECK_PTR(mu); MNR_CHECK_PTR(actions);
  if (rows <= 0 || A <= 0 || act_ld < A) return MNR_ESHAPE;
  tanh_action_kernel<<<grid_for(rows * A), kThreads, 0, mnr_cs(stream)>>>(mu, noise, rows, A, noise_lo, noise_hi, out_lo, out_hi, actions,
                                                                         act_ld);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_tanh_action_bwd(const float* mu, const float* g_actions, const float* g_actions2, int g_ld, int64_t rows, int A,
                                   float* g_mu, mnr_stream_t stream) {
  MNR_CHECK_PTR(mu); MNR_CHECK_PTR(g_actions); MNR_CHECK_PTR(g_mu);
  if (rows <= 0 || A <= 0 || g_ld < A) return MNR_ESHAPE;
  tanh_action_bwd_kernel<<<grid_for(rows * A), kThreads, 0, mnr_cs(stream)>>>(mu, g_actions, g_actions2, g_ld, rows, A, g_mu);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_td_target_stats(const mnr_critic_set_t* qs, int n_q, const float* log_prob, const float* reward, const float* done,
                                   const float* log_alpha, float alpha_host, float gamma_n, float* out, int64_t rows, float* stats,
                                   mnr_stream_t stream) {
  MNR_CHECK_PTR(reward); MNR_CHECK_PTR(done); MNR_CHECK_PTR(out);
  if (rows <= 0 || (stats && rows < 2)) return MNR_ESHAPE;
  QSet s;
  const int rc = pack_qset(qs, n_q, &s, nullptr);
  if (rc != MNR_OK) return rc;
  td_target_stats_kernel<<<1, kOneCta, 0, mnr_cs(stream)>>>(s, n_q, log_prob, reward, done, log_alpha, alpha_host, gamma_n, out, rows, stats);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_actor_seed(const mnr_critic_set_t* qs, int n_q, const float* log_prob, const float* log_alpha, float alpha_host,
                              float target_entropy, int64_t rows, float* out4, float* log_alpha_grad, mnr_stream_t stream) {
  MNR_CHECK_PTR(out4);
  if (rows <= 0) return MNR_ESHAPE;
  QSet s;
  GSet g;
  const int rc = pack_qset(qs, n_q, &s, &g);
  if (rc != MNR_OK) return rc;
  actor_seed_kernel<<<1, kOneCta, 0, mnr_cs(stream)>>>(s, n_q, log_prob, log_alpha, alpha_host, target_entropy, rows, g, out4, log_alpha_grad);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_traj_store_step(const float* obs, const float* next_obs, const float* actions, const float* rewards, const void* dones,
                                   const void* time_outs, int N, int O, int A, int T, int t, float reward_scale, float* traj_obs,
                                   float* traj_next_obs, float* traj_actions, float* traj_rewards, float* traj_dones, mnr_stream_t stream) {
  MNR_CHECK_PTR(obs); MNR_CHECK_PTR(next_obs); MNR_CHECK_PTR(actions); MNR_CHECK_PTR(rewards); MNR_CHECK_PTR(dones);
  MNR_CHECK_PTR(traj_obs); MNR_CHECK_PTR(traj_next_obs); MNR_CHECK_PTR(traj_actions); MNR_CHECK_PTR(traj_rewards); MNR_CHECK_PTR(traj_dones);
  if (N <= 0 || O <= 0 || A <= 0 || T <= 0 || t < 0 || t >= T) return MNR_ESHAPE;
  traj_store_kernel<<<grid_for((int64_t)N * (2 * O + A + 2)), kThreads, 0, mnr_cs(stream)>>>(
      obs, next_obs, actions, rewards, (const uint8_t*)dones, (const uint8_t*)time_outs, N, O, A, T, t, reward_scale, traj_obs,
      traj_next_obs, traj_actions, traj_rewards, traj_dones);
  MNR_LAUNCH_RESULT();
}
