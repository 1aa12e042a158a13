// K4  fused PPO loss forward + backward (one pass over a minibatch) and the un-reduced per-row
// twins of the reference's module-level loss functions.
// Replaces Dist.forward + entropy + neglogp (mineral/nets/dist.py:31-47,
// mineral/agents/ppo/models.py:93-108), actor_loss / critic_loss / bounds_loss / policy_kl
// (mineral/agents/ppo/ppo.py:449-498), the means + total (ppo.py:283-285), autograd's backward of
// all of that, and update_mu_sigma (mineral/agents/ppo/experience.py:83-87).
//
// One thread per minibatch row; the A action columns of a row are contiguous, read as float4 when
// A % 4 == 0 (a warp then covers 32 consecutive rows = one contiguous run per array).  Pass 1 sums
// the row's neg-log-prob / entropy / bounds / KL, pass 2 (inputs now L1-resident) emits the
// gradients, which need the row's probability ratio.  Scalars reduce warp-shuffle -> CTA -> f64
// partials in the workspace -> last CTA (ticket) in a fixed order, so results are deterministic.
#include "common.cuh"
#include "loss_math.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kHeader = 256;                             // ticket + padding
constexpr int kMaxA = 64;
using namespace mnr_loss;

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void load_rw(const float* p) { v[0] = *p; }  // memory this kernel also writes
  __device__ __forceinline__ void store(float* p) const { p[0] = v[0]; }
};
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  __device__ __forceinline__ void load_rw(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

template <int VEC>
__global__ void __launch_bounds__(kThreads) ppo_loss_kernel(
    const float* __restrict__ mu_raw, const float* __restrict__ logstd, int ls_stride, const float* __restrict__ value,
    const float* __restrict__ actions, const float* __restrict__ old_nlp, const float* __restrict__ adv,
    const float* __restrict__ old_values, const float* __restrict__ returns, float* __restrict__ old_mu,
    float* __restrict__ old_sigma, int write_back, float* __restrict__ g_mu_raw, float* __restrict__ g_logstd,
    float* __restrict__ g_value, float* __restrict__ metrics_out, float* __restrict__ mu_mean_out,
    float* __restrict__ sigma_mean_out, int64_t mb, int A, mnr_ppo_loss_cfg_t cfg, unsigned char* __restrict__ workspace) {
  __shared__ float s_col[2 * kMaxA];
  __shared__ double s_red[6][kThreads / 32];
  __shared__ mnr::Moments s_mom[2][kThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int j = tid; j < 2 * A; j += kThreads) s_col[j] = 0.f;
  __syncthreads();

  const float w = cfg.grad_scale / (float)mb;
  float t_a = 0.f, t_c = 0.f, t_b = 0.f, t_h = 0.f, t_clip = 0.f, t_kl = 0.f;
  mnr::Welford w_d, w_ret;
  w_d.init(), w_ret.init();

  for (int64_t row0 = (int64_t)blockIdx.x * kThreads; row0 < mb; row0 += (int64_t)gridDim.x * kThreads) {
    // warp-uniform trip count (shuffles below need all 32 lanes); out-of-range lanes redo the last row
    const bool valid = row0 + tid < mb;
    const int64_t row = valid ? row0 + tid : mb - 1;
    const size_t base = (size_t)row * A;
    const float* ls_row = logstd + (size_t)row * ls_stride;
    // ---- pass 1: row sums ------------------------------------------------------------------
    float nlp = 0.f, ent = 0.f, bsum = 0.f, kl = 0.f;
    for (int j = 0; j < A; j += VEC) {
      Vec<VEC> m, l, a, om, os;
      m.load(mu_raw + base + j), l.load(ls_row + j), a.load(actions + base + j);
      om.load_rw(old_mu + base + j), os.load_rw(old_sigma + base + j);
#pragma unroll
      for (int u = 0; u < VEC; ++u) {
        float mu, sg, dmu, dsg, db;
        dist_fwd(cfg, m.v[u], l.v[u], mu, sg, dmu, dsg);
        const float diff = a.v[u] - mu;
        const float lsg = logf(sg);
        nlp += diff * diff / (2.f * sg * sg) + lsg + kLogSqrt2Pi;  // -Normal.log_prob
        ent += 0.5f + kLogSqrt2Pi + lsg;                           // Normal.entropy
        bsum += bounds_elem(mu, cfg.bounds_type, cfg.soft_bound, &db);
        kl += kl_elem(mu, sg, om.v[u], os.v[u]);
      }
    }
    float dl_dnlp, clipped, dl_dv;
    const float advr = __ldg(adv + row), ret = __ldg(returns + row), v = __ldg(value + row);
    const float a_loss = actor_row(__ldg(old_nlp + row), nlp, advr, cfg.e_clip, cfg.use_smooth_clamp, &dl_dnlp, &clipped);
    const float c_loss = critic_row(__ldg(old_values + row), v, ret, cfg.e_clip, cfg.clip_value_loss, &dl_dv);
    if (valid) {
      t_a += a_loss, t_c += c_loss, t_b += bsum, t_h += ent, t_clip += clipped, t_kl += kl;
      w_d.push(ret - v), w_ret.push(ret);
      g_value[row] = w * 0.5f * cfg.critic_coef * dl_dv;
    }
    const float g_nlp = w * dl_dnlp;
    __syncwarp();  // pass 1 of every lane has read old_mu/old_sigma before any lane overwrites them
    // ---- pass 2: gradients, write-back, column sums -------------------------------------------
    for (int j = 0; j < A; j += VEC) {
      Vec<VEC> m, l, a, gm, gl, nm, ns;
      m.load(mu_raw + base + j), l.load(ls_row + j), a.load(actions + base + j);
#pragma unroll
      for (int u = 0; u < VEC; ++u) {
        float mu, sg, dmu, dsg, db;
        dist_fwd(cfg, m.v[u], l.v[u], mu, sg, dmu, dsg);
        const float diff = a.v[u] - mu;
        const float inv_var = 1.f / (sg * sg);
        bounds_elem(mu, cfg.bounds_type, cfg.soft_bound, &db);
        const float g_mu = g_nlp * (-diff * inv_var) + w * cfg.bounds_loss_coef * db;
        const float g_sg = g_nlp * (1.f - diff * diff * inv_var) / sg - w * cfg.entropy_coef / sg;
        gm.v[u] = g_mu * dmu, gl.v[u] = g_sg * dsg;
        nm.v[u] = mu, ns.v[u] = sg;
        const float cm = mnr::warp_sum(valid ? mu : 0.f), cs = mnr::warp_sum(valid ? sg : 0.f);
        if (lane == 0) atomicAdd(&s_col[j + u], cm), atomicAdd(&s_col[A + j + u], cs);
      }
      if (valid) {
        gm.store(g_mu_raw + base + j), gl.store(g_logstd + base + j);
        if (write_back) nm.store(old_mu + base + j), ns.store(old_sigma + base + j);
      }
    }
  }

  // ---- CTA reduction ---------------------------------------------------------------------------
  float vals[6] = {t_a, t_c, t_b, t_h, t_clip, t_kl};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float s = mnr::warp_sum(vals[k]);
    if (lane == 0) s_red[k][warp] = (double)s;
  }
  const mnr::Moments md = mnr::moments_warp_reduce(w_d.to_moments());
  const mnr::Moments mr = mnr::moments_warp_reduce(w_ret.to_moments());
  if (lane == 0) s_mom[0][warp] = md, s_mom[1][warp] = mr;
  __syncthreads();
  LossPartial* parts = reinterpret_cast<LossPartial*>(workspace + kHeader);
  float* col_parts = reinterpret_cast<float*>(workspace + kHeader + sizeof(LossPartial) * gridDim.x);
  if (tid == 0) {
    LossPartial p;
    double acc[6];
    for (int k = 0; k < 6; ++k) {
      acc[k] = 0;
      for (int x = 0; x < kThreads / 32; ++x) acc[k] += s_red[k][x];
    }
    p.a = acc[0], p.c = acc[1], p.b = acc[2], p.h = acc[3], p.clip = acc[4], p.kl = acc[5];
    p.d = s_mom[0][0], p.ret = s_mom[1][0];
    for (int x = 1; x < kThreads / 32; ++x) p.d = mnr::moments_merge(p.d, s_mom[0][x]), p.ret = mnr::moments_merge(p.ret, s_mom[1][x]);
    parts[blockIdx.x] = p;
  }
  for (int j = tid; j < 2 * A; j += kThreads) col_parts[(size_t)blockIdx.x * 2 * A + j] = s_col[j];

  if (!mnr::last_block_ticket(reinterpret_cast<unsigned int*>(workspace), gridDim.x)) return;
  // last CTA: every thread folds a strided share of the per-CTA partials, then one CTA-wide tree
  double acc[6] = {0, 0, 0, 0, 0, 0};
  mnr::Moments md2{0, 0, 0}, mr2{0, 0, 0};
  for (unsigned int x = tid; x < gridDim.x; x += kThreads) {
    const LossPartial& p = parts[x];
    acc[0] += p.a, acc[1] += p.c, acc[2] += p.b, acc[3] += p.h, acc[4] += p.clip, acc[5] += p.kl;
    md2 = mnr::moments_merge(md2, p.d), mr2 = mnr::moments_merge(mr2, p.ret);
  }
  double tot[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) tot[k] = mnr::block_sum(acc[k]);
  const mnr::Moments td = mnr::block_merge(md2), tr = mnr::block_merge(mr2);
  if (tid == 0) {
    const double n = (double)mb;
    const float a = (float)(tot[0] / n), c = (float)(tot[1] / n), b = (float)(tot[2] / n), h = (float)(tot[3] / n);
    // ppo.py:285
    metrics_out[0] = a + 0.5f * c * cfg.critic_coef - h * cfg.entropy_coef + b * cfg.bounds_loss_coef;
    metrics_out[1] = a, metrics_out[2] = c, metrics_out[3] = b, metrics_out[4] = h;
    metrics_out[5] = (float)(tot[4] / n);
    const float var_d = (float)(td.m2 / (td.n - 1.0)), var_r = (float)(tr.m2 / (tr.n - 1.0));
    metrics_out[6] = fminf(fmaxf(1.f - var_d / (var_r + 1e-8f), 0.f), 1.f);  // ppo.py:475
    metrics_out[7] = (float)(tot[5] / n);
    metrics_out[8] = 0.f, metrics_out[9] = 0.f;
    for (int k = 0; k < MNR_PPO_METRICS; ++k) metrics_out[k] *= cfg.metric_scale;
  }
  if (mu_mean_out && sigma_mean_out) {
    // column sums: thread (slice, j) folds CTAs slice, slice+S, ...; then S slices per column
    __shared__ double s_cols[kThreads];
    const int cols = 2 * A;
    const int S = kThreads / cols;  // >= 1 (A <= kMaxA = 64)
    const int sl = tid / cols, j = tid - sl * cols;
    double sacc = 0;
    if (sl < S)
      for (unsigned int x = sl; x < gridDim.x; x += S) sacc += (double)col_parts[(size_t)x * cols + j];
    s_cols[tid] = sacc;
    __syncthreads();
    if (tid < cols) {
      double t = 0;
      for (int q = 0; q < S; ++q) t += s_cols[q * cols + tid];
      const float mean = (float)(t / (double)mb);
      if (tid < A) mu_mean_out[tid] = mean;
      else sigma_mean_out[tid - A] = mean;
    }
  }
}

// =====================================================================================================
// Fused policy/value HEADS + loss:  mu = Y3a Wmu^T + bmu, logstd = sigma | Y3a Wsig^T + bsig, value = Y3v Wv^T + bv
// (mineral/agents/ppo/models.py:136-151) -> the whole PPO objective and its backward (as ppo_loss_kernel)
// -> gradients w.r.t. the towers' last activations dY3a = g_mu Wmu + g_ls Wsig, dY3v = g_v Wv.
// 8 lanes cooperate on one row: lane s owns a slice of the hidden vector (for the dot products and for
// dY3) and action column(s) j = s, s+8, .. (for the per-element loss math); head weights live in smem.
// =====================================================================================================
constexpr int kHeadThreads = 256, kTPR = 8, kMaxHS = 16;  // hidden width H <= 128, H % 8 == 0

struct HeadsArgs {
  const float *y3a, *y3v;                    // (M,H)
  const float *w_mu, *b_mu, *w_v, *b_v;      // (A,H) (A) (H) (1)
  const float *w_sig, *b_sig;                // learned sigma: (A,H) (A); fixed: w_sig = NULL, b_sig = sigma vector (A)
  const float *actions, *old_nlp, *adv, *old_values, *returns;
  float *old_mu, *old_sigma;
  float *dy3a, *dy3v;                        // (M,H)
  float *g_mu, *g_ls, *g_v;                  // (M,A) (M,A) (M)
  float *metrics, *mu_mean, *sigma_mean;
  unsigned char* workspace;
  int64_t M;
  int A, H, write_back;
};

template <int AMAX, int HS>
__global__ void __launch_bounds__(kHeadThreads) ppo_heads_loss_kernel(const __grid_constant__ HeadsArgs a,
                                                                      const __grid_constant__ mnr_ppo_loss_cfg_t cfg) {
  extern __shared__ float sh[];
  __shared__ float s_col[2 * kMaxA];
  __shared__ double s_red[6][kHeadThreads / 32];
  __shared__ mnr::Moments s_mom[2][kHeadThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int A = a.A, H = a.H;
  constexpr int hs = HS;  // hidden elements per lane (H == 8 * HS)
  const bool learned = a.w_sig != nullptr;
  float* s_wmu = sh;                    // (A,H)
  float* s_wsig = s_wmu + A * H;        // (A,H) (learned only)
  float* s_wv = s_wsig + (learned ? A * H : 0);
  for (int i = tid; i < A * H; i += kHeadThreads) {
    s_wmu[i] = __ldg(a.w_mu + i);
    if (learned) s_wsig[i] = __ldg(a.w_sig + i);
  }
  for (int i = tid; i < H; i += kHeadThreads) s_wv[i] = __ldg(a.w_v + i);
  for (int j = tid; j < 2 * A; j += kHeadThreads) s_col[j] = 0.f;
  __syncthreads();

  const int s = tid & (kTPR - 1);
  const float w = cfg.grad_scale / (float)a.M;
  const float b_v = __ldg(a.b_v);
  float t_a = 0.f, t_c = 0.f, t_b = 0.f, t_h = 0.f, t_clip = 0.f, t_kl = 0.f;
  mnr::Welford w_d, w_ret;
  w_d.init(), w_ret.init();
  constexpr int kRowsPerIter = kHeadThreads / kTPR;

  for (int64_t row0 = (int64_t)blockIdx.x * kRowsPerIter; row0 < a.M; row0 += (int64_t)gridDim.x * kRowsPerIter) {
    const int64_t rraw = row0 + (tid >> 3);
    const bool valid = rraw < a.M;
    const int64_t row = valid ? rraw : a.M - 1;
    // ---- head forward: partial dots over this lane's hidden slice, all-reduced over the 8 lanes ----
    float ya[HS], yv[HS];
    const float* pa = a.y3a + (size_t)row * H + s * hs;
    const float* pv = a.y3v + (size_t)row * H + s * hs;
    if (HS % 4 == 0) {
#pragma unroll
      for (int u = 0; u < HS; u += 4) {
        const float4 fa = __ldg(reinterpret_cast<const float4*>(pa + u)), fv = __ldg(reinterpret_cast<const float4*>(pv + u));
        ya[u] = fa.x, ya[u + 1] = fa.y, ya[u + 2] = fa.z, ya[u + 3] = fa.w;
        yv[u] = fv.x, yv[u + 1] = fv.y, yv[u + 2] = fv.z, yv[u + 3] = fv.w;
      }
    } else {
#pragma unroll
      for (int u = 0; u < HS; ++u) ya[u] = __ldg(pa + u), yv[u] = __ldg(pv + u);
    }
    float mu_raw[AMAX], ls[AMAX];
    float value = 0.f;
#pragma unroll
    for (int u = 0; u < HS; ++u) value = fmaf(yv[u], s_wv[s * hs + u], value);
#pragma unroll
    for (int j = 0; j < AMAX; ++j) {
      float pm = 0.f, pl = 0.f;
      if (j < A) {
#pragma unroll
        for (int u = 0; u < HS; ++u) {
          pm = fmaf(ya[u], s_wmu[j * H + s * hs + u], pm);
          if (learned) pl = fmaf(ya[u], s_wsig[j * H + s * hs + u], pl);
        }
      }
      mu_raw[j] = pm, ls[j] = pl;
    }
#pragma unroll
    for (int d = 1; d < kTPR; d <<= 1) {
      value += __shfl_xor_sync(0xffffffffu, value, d);
#pragma unroll
      for (int j = 0; j < AMAX; ++j) {
        mu_raw[j] += __shfl_xor_sync(0xffffffffu, mu_raw[j], d);
        if (learned) ls[j] += __shfl_xor_sync(0xffffffffu, ls[j], d);
      }
    }
    value += b_v;
    // ---- per-element loss terms: lane s owns columns j = s, s + 8, ... ----------------------------------
    float nlp = 0.f, ent = 0.f, bsum = 0.f, kl = 0.f;
    float e_mu[(AMAX + kTPR - 1) / kTPR], e_sg[(AMAX + kTPR - 1) / kTPR], e_dmu[(AMAX + kTPR - 1) / kTPR],
        e_dsg[(AMAX + kTPR - 1) / kTPR], e_diff[(AMAX + kTPR - 1) / kTPR];
#pragma unroll
    for (int jj = 0; jj < (AMAX + kTPR - 1) / kTPR; ++jj) {
      const int j = s + jj * kTPR;
      e_mu[jj] = 0.f, e_sg[jj] = 1.f, e_dmu[jj] = 0.f, e_dsg[jj] = 0.f, e_diff[jj] = 0.f;
      if (j < A) {
        float m = 0.f, l = 0.f;
#pragma unroll
        for (int q = 0; q < AMAX; ++q)
          if (q == j) m = mu_raw[q], l = ls[q];
        m += __ldg(a.b_mu + j);
        l = learned ? l + __ldg(a.b_sig + j) : __ldg(a.b_sig + j);
        float mu, sg, dmu, dsg, db;
        dist_fwd(cfg, m, l, mu, sg, dmu, dsg);
        const float diff = __ldg(a.actions + (size_t)row * A + j) - mu;
        const float lsg = logf(sg);
        nlp += diff * diff / (2.f * sg * sg) + lsg + kLogSqrt2Pi;
        ent += 0.5f + kLogSqrt2Pi + lsg;
        bsum += bounds_elem(mu, cfg.bounds_type, cfg.soft_bound, &db);
        kl += kl_elem(mu, sg, a.old_mu[(size_t)row * A + j], a.old_sigma[(size_t)row * A + j]);
        e_mu[jj] = mu, e_sg[jj] = sg, e_dmu[jj] = dmu, e_dsg[jj] = dsg, e_diff[jj] = diff;
      }
    }
#pragma unroll
    for (int d = 1; d < kTPR; d <<= 1) {
      nlp += __shfl_xor_sync(0xffffffffu, nlp, d), ent += __shfl_xor_sync(0xffffffffu, ent, d);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, d), kl += __shfl_xor_sync(0xffffffffu, kl, d);
    }
    float dl_dnlp, clipped, dl_dv;
    const float ret = __ldg(a.returns + row);
    const float a_loss = actor_row(__ldg(a.old_nlp + row), nlp, __ldg(a.adv + row), cfg.e_clip, cfg.use_smooth_clamp, &dl_dnlp, &clipped);
    const float c_loss = critic_row(__ldg(a.old_values + row), value, ret, cfg.e_clip, cfg.clip_value_loss, &dl_dv);
    const float g_v = w * 0.5f * cfg.critic_coef * dl_dv;
    if (valid && s == 0) {
      t_a += a_loss, t_c += c_loss, t_b += bsum, t_h += ent, t_clip += clipped, t_kl += kl;
      w_d.push(ret - value), w_ret.push(ret);
      a.g_v[row] = g_v;
    }
    const float g_nlp = w * dl_dnlp;
    // ---- gradients of the owned columns, write-back, column sums ----------------------------------------------
    float gm_own[(AMAX + kTPR - 1) / kTPR], gl_own[(AMAX + kTPR - 1) / kTPR];
#pragma unroll
    for (int jj = 0; jj < (AMAX + kTPR - 1) / kTPR; ++jj) {
      const int j = s + jj * kTPR;
      gm_own[jj] = 0.f, gl_own[jj] = 0.f;
      float cm = 0.f, cs = 0.f;
      if (j < A) {
        const float mu = e_mu[jj], sg = e_sg[jj], diff = e_diff[jj];
        const float inv_var = 1.f / (sg * sg);
        float db;
        bounds_elem(mu, cfg.bounds_type, cfg.soft_bound, &db);
        const float g_mu = g_nlp * (-diff * inv_var) + w * cfg.bounds_loss_coef * db;
        const float g_sg = g_nlp * (1.f - diff * diff * inv_var) / sg - w * cfg.entropy_coef / sg;
        gm_own[jj] = g_mu * e_dmu[jj], gl_own[jj] = g_sg * e_dsg[jj];
        if (valid) {
          a.g_mu[(size_t)row * A + j] = gm_own[jj], a.g_ls[(size_t)row * A + j] = gl_own[jj];
          if (a.write_back) a.old_mu[(size_t)row * A + j] = mu, a.old_sigma[(size_t)row * A + j] = sg;
          cm = mu, cs = sg;
        }
      }
      // the 4 rows of this warp that share column j: lanes s, s+8, s+16, s+24
      cm += __shfl_xor_sync(0xffffffffu, cm, 8), cs += __shfl_xor_sync(0xffffffffu, cs, 8);
      cm += __shfl_xor_sync(0xffffffffu, cm, 16), cs += __shfl_xor_sync(0xffffffffu, cs, 16);
      if (lane < kTPR && j < A) atomicAdd(&s_col[j], cm), atomicAdd(&s_col[A + j], cs);
    }
    // ---- dY3 = g W over this lane's hidden slice (all-gather the row's g over its 8 lanes) ---------------------
    float da[HS];
#pragma unroll
    for (int u = 0; u < HS; ++u) da[u] = 0.f;
#pragma unroll
    for (int j = 0; j < AMAX; ++j) {
      if (j < A) {  // warp-uniform
        const int src = (lane & ~(kTPR - 1)) | (j & (kTPR - 1));
        float gmj = 0.f, glj = 0.f;
#pragma unroll
        for (int jj = 0; jj < (AMAX + kTPR - 1) / kTPR; ++jj) {
          const float tm = __shfl_sync(0xffffffffu, gm_own[jj], src), tl = __shfl_sync(0xffffffffu, gl_own[jj], src);
          if (jj == j / kTPR) gmj = tm, glj = tl;
        }
#pragma unroll
        for (int u = 0; u < HS; ++u) {
          da[u] = fmaf(gmj, s_wmu[j * H + s * hs + u], da[u]);
          if (learned) da[u] = fmaf(glj, s_wsig[j * H + s * hs + u], da[u]);
        }
      }
    }
    if (valid) {
      float* oa = a.dy3a + (size_t)row * H + s * hs;
      float* ov = a.dy3v + (size_t)row * H + s * hs;
      if (HS % 4 == 0) {
#pragma unroll
        for (int u = 0; u < HS; u += 4) {
          *reinterpret_cast<float4*>(oa + u) = make_float4(da[u], da[u + 1], da[u + 2], da[u + 3]);
          *reinterpret_cast<float4*>(ov + u) = make_float4(g_v * s_wv[s * hs + u], g_v * s_wv[s * hs + u + 1],
                                                           g_v * s_wv[s * hs + u + 2], g_v * s_wv[s * hs + u + 3]);
        }
      } else {
#pragma unroll
        for (int u = 0; u < HS; ++u) oa[u] = da[u], ov[u] = g_v * s_wv[s * hs + u];
      }
    }
  }

  // ---- CTA reduction + last-CTA finalize (same scheme as ppo_loss_kernel) ----------------------------------------
  float vals[6] = {t_a, t_c, t_b, t_h, t_clip, t_kl};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float sum = mnr::warp_sum(vals[k]);
    if (lane == 0) s_red[k][warp] = (double)sum;
  }
  const mnr::Moments md = mnr::moments_warp_reduce(w_d.to_moments());
  const mnr::Moments mr = mnr::moments_warp_reduce(w_ret.to_moments());
  if (lane == 0) s_mom[0][warp] = md, s_mom[1][warp] = mr;
  __syncthreads();
  LossPartial* parts = reinterpret_cast<LossPartial*>(a.workspace + kHeader);
  float* col_parts = reinterpret_cast<float*>(a.workspace + kHeader + sizeof(LossPartial) * gridDim.x);
  if (tid == 0) {
    LossPartial p;
    double acc[6];
    for (int k = 0; k < 6; ++k) {
      acc[k] = 0;
      for (int x = 0; x < kHeadThreads / 32; ++x) acc[k] += s_red[k][x];
    }
    p.a = acc[0], p.c = acc[1], p.b = acc[2], p.h = acc[3], p.clip = acc[4], p.kl = acc[5];
    p.d = s_mom[0][0], p.ret = s_mom[1][0];
    for (int x = 1; x < kHeadThreads / 32; ++x) p.d = mnr::moments_merge(p.d, s_mom[0][x]), p.ret = mnr::moments_merge(p.ret, s_mom[1][x]);
    parts[blockIdx.x] = p;
  }
  for (int j = tid; j < 2 * A; j += kHeadThreads) col_parts[(size_t)blockIdx.x * 2 * A + j] = s_col[j];
  if (!mnr::last_block_ticket(reinterpret_cast<unsigned int*>(a.workspace), gridDim.x)) return;
  double acc2[6] = {0, 0, 0, 0, 0, 0};
  mnr::Moments md2{0, 0, 0}, mr2{0, 0, 0};
  for (unsigned int x = tid; x < gridDim.x; x += kHeadThreads) {
    const LossPartial& p = parts[x];
    acc2[0] += p.a, acc2[1] += p.c, acc2[2] += p.b, acc2[3] += p.h, acc2[4] += p.clip, acc2[5] += p.kl;
    md2 = mnr::moments_merge(md2, p.d), mr2 = mnr::moments_merge(mr2, p.ret);
  }
  double tot[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) tot[k] = mnr::block_sum(acc2[k]);
  const mnr::Moments td = mnr::block_merge(md2), tr = mnr::block_merge(mr2);
  if (tid == 0) {
    const double n = (double)a.M;
    const float av = (float)(tot[0] / n), c = (float)(tot[1] / n), b = (float)(tot[2] / n), h = (float)(tot[3] / n);
    a.metrics[0] = av + 0.5f * c * cfg.critic_coef - h * cfg.entropy_coef + b * cfg.bounds_loss_coef;
    a.metrics[1] = av, a.metrics[2] = c, a.metrics[3] = b, a.metrics[4] = h;
    a.metrics[5] = (float)(tot[4] / n);
    const float var_d = (float)(td.m2 / (td.n - 1.0)), var_r = (float)(tr.m2 / (tr.n - 1.0));
    a.metrics[6] = fminf(fmaxf(1.f - var_d / (var_r + 1e-8f), 0.f), 1.f);
    a.metrics[7] = (float)(tot[5] / n);
    a.metrics[8] = 0.f, a.metrics[9] = 0.f;
    for (int k = 0; k < MNR_PPO_METRICS; ++k) a.metrics[k] *= cfg.metric_scale;
  }
  if (a.mu_mean && a.sigma_mean) {
    __shared__ double s_cols2[kHeadThreads];
    const int cols = 2 * A;
    const int S = kHeadThreads / cols;
    const int sl = tid / cols, j = tid - sl * cols;
    double sacc = 0;
    if (sl < S)
      for (unsigned int x = sl; x < gridDim.x; x += S) sacc += (double)col_parts[(size_t)x * cols + j];
    s_cols2[tid] = sacc;
    __syncthreads();
    if (tid < cols) {
      double t = 0;
      for (int q = 0; q < S; ++q) t += s_cols2[q * cols + tid];
      const float mean = (float)(t / (double)a.M);
      if (tid < A) a.mu_mean[tid] = mean;
      else a.sigma_mean[tid - A] = mean;
    }
  }
}

// Head parameter gradients in one pass over Y3a / Y3v:  dW_mu = g_mu^T Y3a, db_mu = colsum g_mu, same for the
// sigma head (learned) or d sigma = colsum g_ls (fixed), dW_v = g_v^T Y3v, db_v = sum g_v.
// Thread (rp, h): column h of the hidden vector, row phase rp; the chunk's g rows are staged in shared memory.
struct HeadGradArgs {
  const float *y3a, *y3v, *g_mu, *g_ls, *g_v;
  float *dw_mu, *db_mu, *dw_sig, *db_sig, *dw_v, *db_v;  // dw_sig NULL: fixed sigma, db_sig = d sigma vector
  float* partials;
  unsigned int* ticket;
  int64_t M;
  int A, H, rows_per_cta;
};

template <int AMAX>
__global__ void __launch_bounds__(256) head_grads_kernel(const __grid_constant__ HeadGradArgs a) {
  extern __shared__ float sg[];  // [rows_per_cta][2A+1] : g_mu | g_ls | g_v
  const int tid = threadIdx.x, A = a.A, H = a.H;
  const int GW = 2 * A + 1;
  const int64_t r0 = (int64_t)blockIdx.x * a.rows_per_cta;
  const int nrows = (int)min((int64_t)a.rows_per_cta, a.M - r0);
  for (int i = tid; i < nrows * A; i += 256) {
    const int r = i / A, j = i - r * A;
    sg[r * GW + j] = __ldg(a.g_mu + (r0 + r) * A + j);
    sg[r * GW + A + j] = __ldg(a.g_ls + (r0 + r) * A + j);
  }
  for (int r = tid; r < nrows; r += 256) sg[r * GW + 2 * A] = __ldg(a.g_v + r0 + r);
  __syncthreads();
  const int rpar = 256 / H;  // H <= 256
  const int rp = tid / H, h = tid - rp * H;
  float am[AMAX], as[AMAX], av = 0.f;
#pragma unroll
  for (int j = 0; j < AMAX; ++j) am[j] = 0.f, as[j] = 0.f;
  if (rp < rpar) {
    int r = rp;
    for (; r + 3 * rpar < nrows; r += 4 * rpar) {  // 8 independent loads in flight
      float ya[4], yv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) ya[q] = __ldg(a.y3a + (r0 + r + q * rpar) * H + h), yv[q] = __ldg(a.y3v + (r0 + r + q * rpar) * H + h);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float* g = sg + (r + q * rpar) * GW;
        av = fmaf(g[2 * A], yv[q], av);
#pragma unroll
        for (int j = 0; j < AMAX; ++j)
          if (j < A) am[j] = fmaf(g[j], ya[q], am[j]), as[j] = fmaf(g[A + j], ya[q], as[j]);
      }
    }
    for (; r < nrows; r += rpar) {
      const float ya = __ldg(a.y3a + (r0 + r) * H + h), yv = __ldg(a.y3v + (r0 + r) * H + h);
      const float* g = sg + r * GW;
      av = fmaf(g[2 * A], yv, av);
#pragma unroll
      for (int j = 0; j < AMAX; ++j)
        if (j < A) am[j] = fmaf(g[j], ya, am[j]), as[j] = fmaf(g[A + j], ya, as[j]);
    }
  }
  // per-CTA partial layout: [GW][H] weights (mu rows, sigma rows, value row) then [GW] bias sums
  const int PW = GW * H + GW;
  float* part = a.partials + (size_t)blockIdx.x * PW;
  __syncthreads();
  float* red = sg;  // reuse: [rpar][GW][H] needs rpar*GW*H <= 256*GW floats (host sizes the buffer for it)
  // bias sums first (they read sg), by the first GW threads
  float bias_sum = 0.f;
  if (tid < GW)
    for (int r = 0; r < nrows; ++r) bias_sum += sg[r * GW + tid];
  __syncthreads();
  if (rp < rpar) {
#pragma unroll
    for (int j = 0; j < AMAX; ++j)
      if (j < A) red[(rp * GW + j) * H + h] = am[j], red[(rp * GW + A + j) * H + h] = as[j];
    red[(rp * GW + 2 * A) * H + h] = av;
  }
  __syncthreads();
  for (int i = tid; i < GW * H; i += 256) {
    float t = 0.f;
    for (int q = 0; q < rpar; ++q) t += red[q * GW * H + i];
    part[i] = t;
  }
  if (tid < GW) part[GW * H + tid] = bias_sum;
  if (!mnr::last_block_ticket(a.ticket, gridDim.x)) return;
  for (int i = tid; i < PW; i += 256) {
    float t = 0.f;
    unsigned int x = 0;
    for (; x + 4 <= gridDim.x; x += 4) {  // 4 independent loads in flight, fixed summation order
      const float p0 = a.partials[(size_t)x * PW + i], p1 = a.partials[(size_t)(x + 1) * PW + i];
      const float p2 = a.partials[(size_t)(x + 2) * PW + i], p3 = a.partials[(size_t)(x + 3) * PW + i];
      t += p0, t += p1, t += p2, t += p3;
    }
    for (; x < gridDim.x; ++x) t += a.partials[(size_t)x * PW + i];
    if (i < GW * H) {
      const int row = i / H, hh = i - row * H;
      if (row < A) a.dw_mu[row * H + hh] = t;
      else if (row < 2 * A) { if (a.dw_sig) a.dw_sig[(row - A) * H + hh] = t; }
      else a.dw_v[hh] = t;
    } else {
      const int j = i - GW * H;
      if (j < A) a.db_mu[j] = t;
      else if (j < 2 * A) a.db_sig[j - A] = t;
      else a.db_v[0] = t;
    }
  }
}

int heads_grid(int64_t M) {
  int64_t g = (M + (kHeadThreads / kTPR) - 1) / (kHeadThreads / kTPR);
  const int64_t cap = (int64_t)mnr_sm_count() * 4;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

int loss_grid(int64_t mb) {
  int64_t g = (mb + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)mnr_sm_count() * 4;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

// ---- generic "rows -> scalar" helpers for the module-level twins -------------------------------------
constexpr int kRowThreads = 256;

struct RowsPartial {
  double s;
  mnr::Moments m0, m1;
};

// true (in every thread of the block) only in the last CTA to finish; there thread 0 gets the totals
__device__ bool block_finish(double my_sum, const mnr::Welford& w0, const mnr::Welford& w1, unsigned char* workspace,
                             RowsPartial* total) {
  __shared__ double s_s[kRowThreads / 32];
  __shared__ mnr::Moments s_m[2][kRowThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double s = mnr::warp_sum(my_sum);
  const mnr::Moments a = mnr::moments_warp_reduce(w0.to_moments()), b = mnr::moments_warp_reduce(w1.to_moments());
  if (lane == 0) s_s[warp] = s, s_m[0][warp] = a, s_m[1][warp] = b;
  __syncthreads();
  RowsPartial* parts = reinterpret_cast<RowsPartial*>(workspace + kHeader);
  if (threadIdx.x == 0) {
    RowsPartial p{s_s[0], s_m[0][0], s_m[1][0]};
    for (int x = 1; x < kRowThreads / 32; ++x) {
      p.s += s_s[x];
      p.m0 = mnr::moments_merge(p.m0, s_m[0][x]), p.m1 = mnr::moments_merge(p.m1, s_m[1][x]);
    }
    parts[blockIdx.x] = p;
  }
  if (!mnr::last_block_ticket(reinterpret_cast<unsigned int*>(workspace), gridDim.x)) return false;
  double acc = 0;
  mnr::Moments m0{0, 0, 0}, m1{0, 0, 0};
  for (unsigned int x = threadIdx.x; x < gridDim.x; x += kRowThreads) {
    acc += parts[x].s;
    m0 = mnr::moments_merge(m0, parts[x].m0), m1 = mnr::moments_merge(m1, parts[x].m1);
  }
  const double ts = mnr::block_sum(acc);
  const mnr::Moments t0 = mnr::block_merge(m0), t1 = mnr::block_merge(m1);
  if (threadIdx.x == 0) *total = RowsPartial{ts, t0, t1};
  return true;
}

__global__ void __launch_bounds__(kRowThreads) actor_rows_kernel(const float* __restrict__ old_nlp,
                                                                 const float* __restrict__ nlp,
                                                                 const float* __restrict__ adv, int64_t n, float e_clip,
                                                                 int smooth, float* __restrict__ loss,
                                                                 float* __restrict__ dloss, float* __restrict__ clip_frac,
                                                                 unsigned char* __restrict__ workspace) {
  double cnt = 0;
  mnr::Welford z0, z1;
  z0.init(), z1.init();
  for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kRowThreads) {
    float d, c;
    loss[i] = actor_row(old_nlp[i], nlp[i], adv[i], e_clip, smooth, &d, &c);
    if (dloss) dloss[i] = d;
    cnt += c;
  }
  RowsPartial t;
  const bool won = block_finish(cnt, z0, z1, workspace, &t);
  if (won && threadIdx.x == 0 && clip_frac) *clip_frac = (float)(t.s / (double)n);
}

__global__ void __launch_bounds__(kRowThreads) critic_rows_kernel(const float* __restrict__ v_old,
                                                                  const float* __restrict__ v,
                                                                  const float* __restrict__ ret, int64_t n, float e_clip,
                                                                  int clip, float* __restrict__ loss,
                                                                  float* __restrict__ dloss, float* __restrict__ ev,
                                                                  unsigned char* __restrict__ workspace) {
  mnr::Welford wd, wr;
  wd.init(), wr.init();
  for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kRowThreads) {
    float d;
    loss[i] = critic_row(v_old[i], v[i], ret[i], e_clip, clip, &d);
    if (dloss) dloss[i] = d;
    wd.push(ret[i] - v[i]), wr.push(ret[i]);
  }
  RowsPartial t;
  const bool won = block_finish(0.0, wd, wr, workspace, &t);
  if (won && threadIdx.x == 0 && ev) {
    const float var_d = (float)(t.m0.m2 / (t.m0.n - 1.0)), var_r = (float)(t.m1.m2 / (t.m1.n - 1.0));
    *ev = fminf(fmaxf(1.f - var_d / (var_r + 1e-8f), 0.f), 1.f);
  }
}

__global__ void __launch_bounds__(kRowThreads) bounds_rows_kernel(const float* __restrict__ mu, int64_t n, int A, int type,
                                                                  float soft, float* __restrict__ loss,
                                                                  float* __restrict__ dmu) {
  for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kRowThreads) {
    float s = 0.f;
    for (int j = 0; j < A; ++j) {
      float d;
      s += bounds_elem(mu[i * A + j], type, soft, &d);
      if (dmu) dmu[i * A + j] = d;
    }
    loss[i] = s;
  }
}

__global__ void __launch_bounds__(kRowThreads) kl_rows_kernel(const float* __restrict__ p0_mu,
                                                              const float* __restrict__ p0_sigma,
                                                              const float* __restrict__ p1_mu,
                                                              const float* __restrict__ p1_sigma, int64_t n, int A,
                                                              float* __restrict__ kl_rows, float* __restrict__ kl_mean,
                                                              unsigned char* __restrict__ workspace) {
  double acc = 0;
  mnr::Welford z0, z1;
  z0.init(), z1.init();
  for (int64_t i = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kRowThreads) {
    float s = 0.f;
    for (int j = 0; j < A; ++j) s += kl_elem(p0_mu[i * A + j], p0_sigma[i * A + j], p1_mu[i * A + j], p1_sigma[i * A + j]);
    if (kl_rows) kl_rows[i] = s;
    acc += s;
  }
  RowsPartial t;
  const bool won = block_finish(acc, z0, z1, workspace, &t);
  if (won && threadIdx.x == 0 && kl_mean) *kl_mean = (float)(t.s / (double)n);
}

int rows_grid(int64_t n) {
  int64_t g = (n + kRowThreads - 1) / kRowThreads;
  const int64_t cap = (int64_t)mnr_sm_count() * 4;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" size_t mnr_ppo_loss_workspace_bytes(int64_t mb, int A) {
  if (mb <= 0 || A <= 0) return 0;
  const size_t grid = (size_t)mnr_sm_count() * 4;
  return kHeader + grid * (sizeof(LossPartial) + sizeof(float) * 2 * (size_t)A);
}

extern "C" int mnr_ppo_loss_fwd_bwd(const float* mu_raw, const float* logstd, int logstd_row_stride, const float* value,
                                    const float* actions, const float* old_neglogp, const float* advantages,
                                    const float* old_values, const float* returns, float* old_mu, float* old_sigma,
                                    int write_back, float* g_mu_raw, float* g_logstd, float* g_value, float* metrics_out,
                                    float* mu_mean_out, float* sigma_mean_out, int64_t mb, int A,
                                    const mnr_ppo_loss_cfg_t* cfg, void* workspace, size_t workspace_bytes,
                                    mnr_stream_t stream) {
  MNR_CHECK_PTR(mu_raw); MNR_CHECK_PTR(logstd); MNR_CHECK_PTR(value); MNR_CHECK_PTR(actions); MNR_CHECK_PTR(old_neglogp);
  MNR_CHECK_PTR(advantages); MNR_CHECK_PTR(old_values); MNR_CHECK_PTR(returns); MNR_CHECK_PTR(old_mu); MNR_CHECK_PTR(old_sigma);
  MNR_CHECK_PTR(g_mu_raw); MNR_CHECK_PTR(g_logstd); MNR_CHECK_PTR(g_value); MNR_CHECK_PTR(metrics_out); MNR_CHECK_PTR(cfg);
  MNR_CHECK_PTR(workspace);
  if (mb <= 0 || A <= 0 || A > kMaxA) return MNR_ESHAPE;
  if (logstd_row_stride != 0 && logstd_row_stride != A) return MNR_ESHAPE;
  if (cfg->dist_type != MNR_DIST_NORMAL && cfg->dist_type != MNR_DIST_DREAMERV3_NORMAL) return MNR_EUNSUP;
  if (cfg->bounds_type != MNR_BOUNDS_BOUND && cfg->bounds_type != MNR_BOUNDS_REG) return MNR_EUNSUP;
  if (workspace_bytes < mnr_ppo_loss_workspace_bytes(mb, A)) return MNR_EWORKSPACE;
  if (!mnr_aligned(workspace, 16)) return MNR_EALIGN;
  const bool vec4 = (A % 4 == 0) && mnr_aligned(mu_raw, 16) && mnr_aligned(logstd, 16) && mnr_aligned(actions, 16) &&
                    mnr_aligned(old_mu, 16) && mnr_aligned(old_sigma, 16) && mnr_aligned(g_mu_raw, 16) &&
                    mnr_aligned(g_logstd, 16);
  const int grid = loss_grid(mb);
  if (vec4)
    ppo_loss_kernel<4><<<grid, kThreads, 0, mnr_cs(stream)>>>(
        mu_raw, logstd, logstd_row_stride, value, actions, old_neglogp, advantages, old_values, returns, old_mu,
        old_sigma, write_back, g_mu_raw, g_logstd, g_value, metrics_out, mu_mean_out, sigma_mean_out, mb, A, *cfg,
        (unsigned char*)workspace);
  else
    ppo_loss_kernel<1><<<grid, kThreads, 0, mnr_cs(stream)>>>(
        mu_raw, logstd, logstd_row_stride, value, actions, old_neglogp, advantages, old_values, returns, old_mu,
        old_sigma, write_back, g_mu_raw, g_logstd, g_value, metrics_out, mu_mean_out, sigma_mean_out, mb, A, *cfg,
        (unsigned char*)workspace);
  MNR_LAUNCH_RESULT();
}

extern "C" size_t mnr_rows_reduce_workspace_bytes(int64_t n) {
  (void)n;
  return kHeader + (size_t)mnr_sm_count() * 4 * sizeof(RowsPartial);
}

extern "C" int mnr_actor_loss_rows(const float* old_neglogp, const float* neglogp, const float* advantages, int64_t n,
                                   float e_clip, int use_smooth_clamp, float* loss_out, float* dloss_dneglogp_out,
                                   float* clip_frac_out, void* workspace, size_t workspace_bytes, mnr_stream_t stream) {
  MNR_CHECK_PTR(old_neglogp); MNR_CHECK_PTR(neglogp); MNR_CHECK_PTR(advantages); MNR_CHECK_PTR(loss_out); MNR_CHECK_PTR(workspace);
  if (n <= 0) return MNR_ESHAPE;
  if (workspace_bytes < mnr_rows_reduce_workspace_bytes(n)) return MNR_EWORKSPACE;
  actor_rows_kernel<<<rows_grid(n), kRowThreads, 0, mnr_cs(stream)>>>(old_neglogp, neglogp, advantages, n, e_clip,
                                                                      use_smooth_clamp, loss_out, dloss_dneglogp_out,
                                                                      clip_frac_out, (unsigned char*)workspace);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_critic_loss_rows(const float* old_values, const float* values, const float* returns, int64_t n,
                                    float e_clip, int clip_value_loss, float* loss_out, float* dloss_dvalues_out,
                                    float* explained_var_out, void* workspace, size_t workspace_bytes,
                                    mnr_stream_t stream) {
  MNR_CHECK_PTR(old_values); MNR_CHECK_PTR(values); MNR_CHECK_PTR(returns); MNR_CHECK_PTR(loss_out); MNR_CHECK_PTR(workspace);
  if (n <= 0) return MNR_ESHAPE;
  if (workspace_bytes < mnr_rows_reduce_workspace_bytes(n)) return MNR_EWORKSPACE;
  critic_rows_kernel<<<rows_grid(n), kRowThreads, 0, mnr_cs(stream)>>>(old_values, values, returns, n, e_clip,
                                                                       clip_value_loss, loss_out, dloss_dvalues_out,
                                                                       explained_var_out, (unsigned char*)workspace);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_bounds_loss_rows(const float* mu, int64_t n, int A, int bounds_type, float soft_bound, float* loss_out,
                                    float* dloss_dmu_out, mnr_stream_t stream) {
  MNR_CHECK_PTR(mu); MNR_CHECK_PTR(loss_out);
  if (n <= 0 || A <= 0) return MNR_ESHAPE;
  if (bounds_type != MNR_BOUNDS_BOUND && bounds_type != MNR_BOUNDS_REG) return MNR_EUNSUP;
  bounds_rows_kernel<<<rows_grid(n), kRowThreads, 0, mnr_cs(stream)>>>(mu, n, A, bounds_type, soft_bound, loss_out,
                                                                       dloss_dmu_out);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_policy_kl_rows(const float* p0_mu, const float* p0_sigma, const float* p1_mu, const float* p1_sigma,
                                  int64_t n, int A, float* kl_rows_out, float* kl_mean_out, void* workspace,
                                  size_t workspace_bytes, mnr_stream_t stream) {
  MNR_CHECK_PTR(p0_mu); MNR_CHECK_PTR(p0_sigma); MNR_CHECK_PTR(p1_mu); MNR_CHECK_PTR(p1_sigma); MNR_CHECK_PTR(workspace);
  if (n <= 0 || A <= 0) return MNR_ESHAPE;
  if (workspace_bytes < mnr_rows_reduce_workspace_bytes(n)) return MNR_EWORKSPACE;
  kl_rows_kernel<<<rows_grid(n), kRowThreads, 0, mnr_cs(stream)>>>(p0_mu, p0_sigma, p1_mu, p1_sigma, n, A, kl_rows_out,
                                                                   kl_mean_out, (unsigned char*)workspace);
  MNR_LAUNCH_RESULT();
}

// csrc/heads_fused.cu: the single-kernel path (H in {32, 64, 128}, A <= 32)
int mnr_heads_fused_grid_cap();
bool mnr_heads_fused_supported(int A, int H);
size_t mnr_heads_fused_workspace_bytes(int A, int H);
int mnr_heads_fused_launch(const float* y3a, const float* y3v, const float* w_mu, const float* b_mu, const float* w_v,
                           const float* b_v, const float* w_sig, const float* b_sig, const float* actions,
                           const float* old_neglogp, const float* advantages, const float* old_values, const float* returns,
                           float* old_mu, float* old_sigma, int write_back, float* dy3a, float* dy3v, float* g_mu, float* g_ls,
                           float* g_v, float* dw_mu, float* db_mu, float* dw_v, float* db_v, float* dw_sig, float* db_sig,
                           float* metrics_out, float* mu_mean_out, float* sigma_mean_out, int64_t M, int A, int H,
                           const mnr_ppo_loss_cfg_t* cfg, void* workspace, cudaStream_t stream);

extern "C" int mnr_ppo_heads_is_fused(int A, int H) { return mnr_heads_fused_supported(A, H) ? 1 : 0; }

extern "C" size_t mnr_ppo_heads_workspace_bytes(int64_t M, int A, int H) {
  if (M <= 0 || A <= 0 || H <= 0) return 0;
  if (mnr_heads_fused_supported(A, H)) return mnr_heads_fused_workspace_bytes(A, H);
  const size_t grid = (size_t)mnr_sm_count() * 4;
  const size_t loss = kHeader + grid * (sizeof(LossPartial) + sizeof(float) * 2 * (size_t)A);
  const size_t rows_per_cta = 128;
  const size_t ctas = ((size_t)M + rows_per_cta - 1) / rows_per_cta;
  const size_t grads = 256 + ctas * (size_t)((2 * A + 1) * H + (2 * A + 1)) * sizeof(float);
  return ((loss + 255) / 256) * 256 + grads;
}

extern "C" int mnr_ppo_heads_loss_fwd_bwd(const float* y3a, const float* y3v, const float* w_mu, const float* b_mu,
                                          const float* w_v, const float* b_v, const float* w_sig, const float* b_sig,
                                          const float* actions, const float* old_neglogp, const float* advantages,
                                          const float* old_values, const float* returns, float* old_mu, float* old_sigma,
                                          int write_back, float* dy3a, float* dy3v, float* g_mu, float* g_ls, float* g_v,
                                          float* dw_mu, float* db_mu, float* dw_v, float* db_v, float* dw_sig, float* db_sig,
                                          float* metrics_out, float* mu_mean_out, float* sigma_mean_out, int64_t M, int A,
                                          int H, const mnr_ppo_loss_cfg_t* cfg, void* workspace, size_t workspace_bytes,
                                          mnr_stream_t stream) {
  MNR_CHECK_PTR(y3a); MNR_CHECK_PTR(y3v); MNR_CHECK_PTR(w_mu); MNR_CHECK_PTR(b_mu); MNR_CHECK_PTR(w_v); MNR_CHECK_PTR(b_v);
  MNR_CHECK_PTR(b_sig); MNR_CHECK_PTR(actions); MNR_CHECK_PTR(old_neglogp); MNR_CHECK_PTR(advantages); MNR_CHECK_PTR(old_values);
  MNR_CHECK_PTR(returns); MNR_CHECK_PTR(old_mu); MNR_CHECK_PTR(old_sigma); MNR_CHECK_PTR(dy3a); MNR_CHECK_PTR(dy3v);
  MNR_CHECK_PTR(g_mu); MNR_CHECK_PTR(g_ls); MNR_CHECK_PTR(g_v); MNR_CHECK_PTR(dw_mu); MNR_CHECK_PTR(db_mu); MNR_CHECK_PTR(dw_v);
  MNR_CHECK_PTR(db_v); MNR_CHECK_PTR(db_sig); MNR_CHECK_PTR(metrics_out); MNR_CHECK_PTR(cfg); MNR_CHECK_PTR(workspace);
  if (w_sig && !dw_sig) return MNR_ENULL;
  if (M <= 0 || A <= 0 || A > 32 || H <= 0 || H % kTPR != 0 || H / kTPR > kMaxHS) return MNR_ESHAPE;
  { const int q = H / kTPR; if (q != 1 && q != 2 && q != 4 && q != 8 && q != 16) return MNR_ESHAPE; }
  if (cfg->dist_type != MNR_DIST_NORMAL && cfg->dist_type != MNR_DIST_DREAMERV3_NORMAL) return MNR_EUNSUP;
  if (cfg->bounds_type != MNR_BOUNDS_BOUND && cfg->bounds_type != MNR_BOUNDS_REG) return MNR_EUNSUP;
  if (workspace_bytes < mnr_ppo_heads_workspace_bytes(M, A, H)) return MNR_EWORKSPACE;
  if (!mnr_aligned(workspace, 16)) return MNR_EALIGN;
  if (cfg->dz_act < 0 || cfg->dz_act > 2) return MNR_EUNSUP;
  if (cfg->dz_act != 0 && !mnr_heads_fused_supported(A, H)) return MNR_EUNSUP;   // dZ3 output: single-kernel form only
  if (mnr_heads_fused_supported(A, H))
    return mnr_heads_fused_launch(y3a, y3v, w_mu, b_mu, w_v, b_v, w_sig, b_sig, actions, old_neglogp, advantages, old_values,
                                  returns, old_mu, old_sigma, write_back, dy3a, dy3v, g_mu, g_ls, g_v, dw_mu, db_mu, dw_v, db_v,
                                  dw_sig, db_sig, metrics_out, mu_mean_out, sigma_mean_out, M, A, H, cfg, workspace,
                                  mnr_cs(stream));
  const size_t grid_cap = (size_t)mnr_sm_count() * 4;
  const size_t loss_bytes = ((kHeader + grid_cap * (sizeof(LossPartial) + sizeof(float) * 2 * (size_t)A) + 255) / 256) * 256;

  HeadsArgs ha{y3a, y3v, w_mu, b_mu, w_v, b_v, w_sig, b_sig, actions, old_neglogp, advantages, old_values, returns,
               old_mu, old_sigma, dy3a, dy3v, g_mu, g_ls, g_v, metrics_out, mu_mean_out, sigma_mean_out,
               (unsigned char*)workspace, M, A, H, write_back};
  const size_t smem1 = (size_t)((w_sig ? 2 : 1) * A * H + H) * sizeof(float);
  const int grid1 = heads_grid(M);
#define MNR_HEADS_LAUNCH(AM, HSV) ppo_heads_loss_kernel<AM, HSV><<<grid1, kHeadThreads, smem1, mnr_cs(stream)>>>(ha, *cfg)
#define MNR_HEADS_BY_HS(AM)                                                        \
  switch (H / kTPR) {                                                              \
    case 1: MNR_HEADS_LAUNCH(AM, 1); break;   case 2: MNR_HEADS_LAUNCH(AM, 2); break;   \
    case 4: MNR_HEADS_LAUNCH(AM, 4); break;   case 8: MNR_HEADS_LAUNCH(AM, 8); break;   \
    case 16: MNR_HEADS_LAUNCH(AM, 16); break; default: return MNR_ESHAPE;          \
  }
  if (A <= 8) { MNR_HEADS_BY_HS(8) } else if (A <= 16) { MNR_HEADS_BY_HS(16) } else { MNR_HEADS_BY_HS(32) }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return (int)e;

  const int rows_per_cta = 128;
  const int GW = 2 * A + 1;
  unsigned char* gws = (unsigned char*)workspace + loss_bytes;
  HeadGradArgs ga{y3a, y3v, g_mu, g_ls, g_v, dw_mu, db_mu, dw_sig, db_sig, dw_v, db_v, (float*)(gws + 256),
                  (unsigned int*)gws, M, A, H, rows_per_cta};
  const int grid2 = (int)((M + rows_per_cta - 1) / rows_per_cta);
  // staged g tile [rows_per_cta][GW], reused as the [256/H][GW][H] cross-phase reduction scratch
  const size_t smem2 = (size_t)(rows_per_cta > 256 ? rows_per_cta : 256) * GW * sizeof(float);
  if (A <= 8) head_grads_kernel<8><<<grid2, 256, smem2, mnr_cs(stream)>>>(ga);
  else if (A <= 16) head_grads_kernel<16><<<grid2, 256, smem2, mnr_cs(stream)>>>(ga);
  else {
    e = mnr_allow_max_dyn_smem(head_grads_kernel<32>);
    if (e != cudaSuccess) return (int)e;
    head_grads_kernel<32><<<grid2, 256, smem2, mnr_cs(stream)>>>(ga);
  }
  MNR_LAUNCH_RESULT();
}
