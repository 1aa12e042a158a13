// K4'  policy / value HEADS + the whole PPO objective + its backward + the head parameter gradients in ONE
// persistent kernel (the fast path behind mnr_ppo_heads_loss_fwd_bwd, csrc/ppo_loss.cu):
//
//   mu = Y3a Wmu^T + bmu,  logstd = sigma | Y3a Wsig^T + bsig,  value = Y3v Wv^T + bv   (mineral/agents/ppo/models.py:136-151)
//   Dist / neglogp / entropy / actor, critic, bounds losses / KL / means / total          (nets/dist.py:31-47, ppo.py:271-303,449-498)
//   autograd's backward of all of that:  g_mu, g_logstd, g_value per row,
//   dY3a = g_mu Wmu + g_ls Wsig,  dY3v = g_v Wv                                           (into the towers' backward)
//   dWmu = g_mu^T Y3a, dbmu, dWsig / dsigma, dWv = g_v^T Y3v, dbv                         (head parameter gradients)
//   mu / sigma write-back (mineral/agents/ppo/experience.py:83-87) and their column means (results['mu'|'sigma'])
//
// Layout of the work.  A CTA = 128 threads owns tiles of 128 rows; a tile of Y3a and Y3v is staged ONCE in shared
// memory (coalesced float4 loads, row pitch H+4 floats so that "thread = row" LDS.128 sweeps are conflict-free) and is
// used three times: (1) thread-per-row dot products with the head weights (broadcast LDS.128 from shared memory),
// (2) the weight-gradient outer products with thread = (hidden column u, row group) walking the 128 rows of the
// tile, accumulating in registers ACROSS tiles, (3) it is overwritten in place by dY3 and leaves through coalesced
// float4 stores.  After the tile loop the per-CTA partials go to the workspace, the grid (all CTAs resident) meets at
// a grid barrier and every CTA reduces its share of the outputs over all CTAs' partials in a fixed order
// (deterministic, no atomics); CTA 0 also finalises the ten loss metrics.
#include "common.cuh"
#include "loss_math.cuh"

namespace {
using namespace mnr_loss;

constexpr int kRows = 128;  // rows per tile == threads per CTA

struct FusedHeadsArgs {
  const float *y3a, *y3v;                // (M,H)
  const float *w_mu, *b_mu, *w_v, *b_v;  // (A,H) (A) (H) (1)
  const float *w_sig, *b_sig;            // learned sigma: (A,H) (A); fixed: w_sig = NULL, b_sig = sigma vector (A)
  const float *actions, *old_nlp, *adv, *old_values, *returns;
  float *old_mu, *old_sigma;
  float *dy3a, *dy3v;                    // (M,H)
  float *g_mu, *g_ls, *g_v;              // optional (M,A) (M,A) (M)
  float *dw_mu, *db_mu, *dw_sig, *db_sig, *dw_v, *db_v;
  float *metrics, *mu_mean, *sigma_mean;
  unsigned long long* barrier;           // 2 words
  LossPartial* loss_parts;               // [grid]
  float* partials;                       // [grid][PW]
  int64_t M;
  int A, write_back, n_tiles;
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc), acc = fmaf(a.y, b.y, acc), acc = fmaf(a.z, b.z, acc), acc = fmaf(a.w, b.w, acc);
  return acc;
}

template <int AMAX, int H, bool LEARNED>
__global__ void __launch_bounds__(kRows) heads_fused_kernel(const __grid_constant__ FusedHeadsArgs a,
                                                            const __grid_constant__ mnr_ppo_loss_cfg_t cfg) {
  constexpr int P = H + 4, P4 = P / 4, H4 = H / 4;
  constexpr int G = kRows / H;                               // row groups of the weight-gradient phase
  constexpr int NYMAX = (LEARNED ? 2 : 1) * AMAX;            // gradient rows fed by Y3a
  constexpr int JR = (NYMAX + G - 1) / G;                    // of which one thread owns at most JR
  extern __shared__ __align__(16) float sh[];
  __shared__ double s_red[6][kRows / 32];
  __shared__ mnr::Moments s_mom[2][kRows / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int A = a.A;
  const int GW = 4 * A + 1;  // s_g columns: g_mu | g_ls | g_v | mu | sigma   (odd pitch: conflict-free both ways)
  const int NY = (LEARNED ? 2 : 1) * A;
  float* s_ya = sh;
  float* s_yv = s_ya + kRows * P;
  float* s_wmu = s_yv + kRows * P;                 // (A,H)
  float* s_wsig = s_wmu + AMAX * H;                // (A,H), learned only
  float* s_wv = s_wsig + (LEARNED ? AMAX * H : 0); // (H)
  float* s_g = s_wv + H;                           // [128][GW]

  for (int i = tid; i < A * H; i += kRows) {
    s_wmu[i] = __ldg(a.w_mu + i);
    if (LEARNED) s_wsig[i] = __ldg(a.w_sig + i);
  }
  for (int i = tid; i < H; i += kRows) s_wv[i] = __ldg(a.w_v + i);

  const float w = cfg.grad_scale / (float)a.M;
  const float b_v = __ldg(a.b_v);
  float t_a = 0.f, t_c = 0.f, t_b = 0.f, t_h = 0.f, t_clip = 0.f, t_kl = 0.f;
  mnr::Welford w_d, w_ret;
  w_d.init(), w_ret.init();

  // weight-gradient accumulators (live across tiles): thread = (hidden column u, row group grp)
  const int u = tid % H, grp = tid / H;
  float acc[JR], acc_v = 0.f, acc_col = 0.f;
#pragma unroll
  for (int jj = 0; jj < JR; ++jj) acc[jj] = 0.f;

  const float4* gya = reinterpret_cast<const float4*>(a.y3a);
  const float4* gyv = reinterpret_cast<const float4*>(a.y3v);
  float4* s_ya4 = reinterpret_cast<float4*>(s_ya);
  float4* s_yv4 = reinterpret_cast<float4*>(s_yv);
  const float4* s_wmu4 = reinterpret_cast<const float4*>(s_wmu);
  const float4* s_wsig4 = reinterpret_cast<const float4*>(s_wsig);
  const float4* s_wv4 = reinterpret_cast<const float4*>(s_wv);

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * kRows;
    const int rows_here = (int)min((int64_t)kRows, a.M - row0);
    __syncthreads();  // previous tile's stores out of s_ya / s_yv are done (and the weights are staged)
    // ---- stage the tile: [r][H] -> pitch P ------------------------------------------------------------------
#pragma unroll 8
    for (int i = tid; i < kRows * H4; i += kRows) {
      const int r = i / H4, c = i - r * H4;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vv = va;
      if (r < rows_here) va = __ldg(gya + (row0 + r) * H4 + c), vv = __ldg(gyv + (row0 + r) * H4 + c);
      s_ya4[r * P4 + c] = va, s_yv4[r * P4 + c] = vv;
    }
    __syncthreads();

    // ---- (1) thread = row: head forward --------------------------------------------------------------------
    const int r = tid;
    const bool valid = r < rows_here;
    const int64_t row = valid ? row0 + r : a.M - 1;
    float mu_raw[AMAX], ls[AMAX], value = 0.f;
#pragma unroll
    for (int j = 0; j < AMAX; ++j) mu_raw[j] = 0.f, ls[j] = 0.f;
#pragma unroll 2
    for (int c = 0; c < H4; ++c) {
      const float4 ya = s_ya4[r * P4 + c], yv = s_yv4[r * P4 + c];
      value = dot4(yv, s_wv4[c], value);
#pragma unroll
      for (int j = 0; j < AMAX; ++j) {
        if (j < A) {
          mu_raw[j] = dot4(ya, s_wmu4[j * H4 + c], mu_raw[j]);
          if (LEARNED) ls[j] = dot4(ya, s_wsig4[j * H4 + c], ls[j]);
        }
      }
    }
    value += b_v;

    // ---- per-element loss terms ----------------------------------------------------------------------------------
    float nlp = 0.f, ent = 0.f, bsum = 0.f, kl = 0.f;
    float e_mu[AMAX], e_sg[AMAX], e_dmu[AMAX], e_dsg[AMAX], e_diff[AMAX];
#pragma unroll
    for (int j = 0; j < AMAX; ++j) {
      e_mu[j] = 0.f, e_sg[j] = 1.f, e_dmu[j] = 0.f, e_dsg[j] = 0.f, e_diff[j] = 0.f;
      if (j < A) {
        const float m = mu_raw[j] + __ldg(a.b_mu + j);
        const float l = LEARNED ? ls[j] + __ldg(a.b_sig + j) : __ldg(a.b_sig + j);
        float mu, sg, dmu, dsg, db;
        dist_fwd(cfg, m, l, mu, sg, dmu, dsg);
        const float diff = __ldg(a.actions + (size_t)row * A + j) - mu;
        const float lsg = logf(sg);
        nlp += diff * diff / (2.f * sg * sg) + lsg + kLogSqrt2Pi;
        ent += 0.5f + kLogSqrt2Pi + lsg;
        bsum += bounds_elem(mu, cfg.bounds_type, cfg.soft_bound, &db);
        kl += kl_elem(mu, sg, a.old_mu[(size_t)row * A + j], a.old_sigma[(size_t)row * A + j]);
        e_mu[j] = mu, e_sg[j] = sg, e_dmu[j] = dmu, e_dsg[j] = dsg, e_diff[j] = diff;
      }
    }
    float dl_dnlp, clipped, dl_dv;
    const float ret = __ldg(a.returns + row);
    const float a_loss = actor_row(__ldg(a.old_nlp + row), nlp, __ldg(a.adv + row), cfg.e_clip, cfg.use_smooth_clamp, &dl_dnlp, &clipped);
    const float c_loss = critic_row(__ldg(a.old_values + row), value, ret, cfg.e_clip, cfg.clip_value_loss, &dl_dv);
    const float g_v = valid ? w * 0.5f * cfg.critic_coef * dl_dv : 0.f;
    if (valid) {
      t_a += a_loss, t_c += c_loss, t_b += bsum, t_h += ent, t_clip += clipped, t_kl += kl;
      w_d.push(ret - value), w_ret.push(ret);
      if (a.g_v) a.g_v[row] = g_v;
    }
    const float g_nlp = w * dl_dnlp;
    float* grow = s_g + r * GW;
#pragma unroll
    for (int j = 0; j < AMAX; ++j) {
      if (j < A) {
        const float mu = e_mu[j], sg = e_sg[j], diff = e_diff[j];
        const float inv_var = 1.f / (sg * sg);
        float db;
        bounds_elem(mu, cfg.bounds_type, cfg.soft_bound, &db);
        const float g_m = g_nlp * (-diff * inv_var) + w * cfg.bounds_loss_coef * db;
        const float g_s = g_nlp * (1.f - diff * diff * inv_var) / sg - w * cfg.entropy_coef / sg;
        const float gm = valid ? g_m * e_dmu[j] : 0.f, gl = valid ? g_s * e_dsg[j] : 0.f;
        grow[j] = gm, grow[A + j] = gl;
        grow[2 * A + 1 + j] = valid ? mu : 0.f, grow[3 * A + 1 + j] = valid ? sg : 0.f;
        if (valid) {
          if (a.g_mu) a.g_mu[(size_t)row * A + j] = gm;
          if (a.g_ls) a.g_ls[(size_t)row * A + j] = gl;
          if (a.write_back) a.old_mu[(size_t)row * A + j] = mu, a.old_sigma[(size_t)row * A + j] = sg;
        }
      }
    }
    grow[2 * A] = g_v;
    __syncthreads();

    // ---- (2) thread = (u, grp): weight-gradient outer products over the tile's rows; column sums ------------
#pragma unroll 4
    for (int rr = 0; rr < kRows; ++rr) {
      const float ya = s_ya[rr * P + u];
      const float* g = s_g + rr * GW;
#pragma unroll
      for (int jj = 0; jj < JR; ++jj) {
        const int jrow = grp + jj * G;
        if (jrow < NY) acc[jj] = fmaf(g[jrow], ya, acc[jj]);
      }
      if (grp == G - 1) acc_v = fmaf(g[2 * A], s_yv[rr * P + u], acc_v);
    }
    if (tid < GW) {
      float s = 0.f;
#pragma unroll 8
      for (int rr = 0; rr < kRows; ++rr) s += s_g[rr * GW + tid];
      acc_col += s;
    }
    __syncthreads();

    // ---- (3) thread = row again: dY3 over the tile in place, then coalesced stores ---------------------------
    {
      float gm[AMAX], gl[AMAX];
#pragma unroll
      for (int j = 0; j < AMAX; ++j) {
        gm[j] = 0.f, gl[j] = 0.f;
        if (j < A) gm[j] = grow[j], gl[j] = LEARNED ? grow[A + j] : 0.f;
      }
#pragma unroll 2
      for (int c = 0; c < H4; ++c) {
        float4 da = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < AMAX; ++j) {
          if (j < A) {
            const float4 wm = s_wmu4[j * H4 + c];
            da.x = fmaf(gm[j], wm.x, da.x), da.y = fmaf(gm[j], wm.y, da.y), da.z = fmaf(gm[j], wm.z, da.z), da.w = fmaf(gm[j], wm.w, da.w);
            if (LEARNED) {
              const float4 ws = s_wsig4[j * H4 + c];
              da.x = fmaf(gl[j], ws.x, da.x), da.y = fmaf(gl[j], ws.y, da.y), da.z = fmaf(gl[j], ws.z, da.z), da.w = fmaf(gl[j], ws.w, da.w);
            }
          }
        }
        const float4 wv = s_wv4[c];
        s_ya4[r * P4 + c] = da;
        s_yv4[r * P4 + c] = make_float4(g_v * wv.x, g_v * wv.y, g_v * wv.z, g_v * wv.w);
      }
    }
    __syncthreads();
    float4* oa = reinterpret_cast<float4*>(a.dy3a);
    float4* ov = reinterpret_cast<float4*>(a.dy3v);
#pragma unroll 8
    for (int i = tid; i < kRows * H4; i += kRows) {
      const int rr = i / H4, c = i - rr * H4;
      if (rr < rows_here) oa[(row0 + rr) * H4 + c] = s_ya4[rr * P4 + c], ov[(row0 + rr) * H4 + c] = s_yv4[rr * P4 + c];
    }
  }

  // ---- per-CTA partials -> workspace ----------------------------------------------------------------------------
  const int PW = (2 * A + 1) * H + GW;
  float* part = a.partials + (size_t)blockIdx.x * PW;
#pragma unroll
  for (int jj = 0; jj < JR; ++jj) {
    const int jrow = grp + jj * G;
    if (jrow < NY) part[jrow * H + u] = acc[jj];
  }
  if (!LEARNED)  // sigma rows of the layout stay defined
    for (int i = tid; i < A * H; i += kRows) part[(A + i / H) * H + i % H] = 0.f;
  if (grp == G - 1) part[2 * A * H + u] = acc_v;
  if (tid < GW) part[(2 * A + 1) * H + tid] = acc_col;

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
  if (tid == 0) {
    LossPartial p;
    double ac[6];
    for (int k = 0; k < 6; ++k) {
      ac[k] = 0;
      for (int x = 0; x < kRows / 32; ++x) ac[k] += s_red[k][x];
    }
    p.a = ac[0], p.c = ac[1], p.b = ac[2], p.h = ac[3], p.clip = ac[4], p.kl = ac[5];
    p.d = s_mom[0][0], p.ret = s_mom[1][0];
    for (int x = 1; x < kRows / 32; ++x) p.d = mnr::moments_merge(p.d, s_mom[0][x]), p.ret = mnr::moments_merge(p.ret, s_mom[1][x]);
    a.loss_parts[blockIdx.x] = p;
  }

  mnr::grid_barrier(a.barrier);

  // ---- every CTA reduces its share of the PW outputs over all CTAs (warp per output, fixed order) ----------------
  const int n_cta = gridDim.x;
  for (int o = blockIdx.x * (kRows / 32) + warp; o < PW; o += n_cta * (kRows / 32)) {
    float s = 0.f;
    for (int x = lane; x < n_cta; x += 32) s += __ldcg(a.partials + (size_t)x * PW + o);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) {
      if (o < (2 * A + 1) * H) {
        const int jrow = o / H, uu = o - jrow * H;
        if (jrow < A) a.dw_mu[jrow * H + uu] = s;
        else if (jrow < 2 * A) { if (LEARNED) a.dw_sig[(jrow - A) * H + uu] = s; }
        else a.dw_v[uu] = s;
      } else {
        const int j = o - (2 * A + 1) * H;
        if (j < A) a.db_mu[j] = s;
        else if (j < 2 * A) a.db_sig[j - A] = s;           // learned: bias gradient; fixed: d sigma = colsum g_ls
        else if (j == 2 * A) a.db_v[0] = s;
        else if (j < 3 * A + 1) { if (a.mu_mean) a.mu_mean[j - 2 * A - 1] = s / (float)a.M; }
        else if (a.sigma_mean) a.sigma_mean[j - 3 * A - 1] = s / (float)a.M;
      }
    }
  }
  if (blockIdx.x != 0) return;

  // ---- CTA 0: the ten loss metrics ---------------------------------------------------------------------------------
  double acc2[6] = {0, 0, 0, 0, 0, 0};
  mnr::Moments md2{0, 0, 0}, mr2{0, 0, 0};
  for (int x = tid; x < n_cta; x += kRows) {
    const LossPartial p = a.loss_parts[x];
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
    float m[MNR_PPO_METRICS];
    m[0] = av + 0.5f * c * cfg.critic_coef - h * cfg.entropy_coef + b * cfg.bounds_loss_coef;
    m[1] = av, m[2] = c, m[3] = b, m[4] = h;
    m[5] = (float)(tot[4] / n);
    const float var_d = (float)(td.m2 / (td.n - 1.0)), var_r = (float)(tr.m2 / (tr.n - 1.0));
    m[6] = fminf(fmaxf(1.f - var_d / (var_r + 1e-8f), 0.f), 1.f);
    m[7] = (float)(tot[5] / n);
    m[8] = 0.f, m[9] = 0.f;
    for (int k = 0; k < MNR_PPO_METRICS; ++k) a.metrics[k] = m[k] * cfg.metric_scale;
  }
}

template <int AMAX, int H, bool LEARNED>
size_t fused_smem_bytes(int A) {
  return (size_t)(2 * kRows * (H + 4) + (LEARNED ? 2 : 1) * AMAX * H + H + kRows * (4 * A + 1)) * sizeof(float);
}

template <int AMAX, int H, bool LEARNED>
int launch_fused(FusedHeadsArgs& a, const mnr_ppo_loss_cfg_t& cfg, int grid_cap, cudaStream_t stream) {
  const size_t smem = fused_smem_bytes<AMAX, H, LEARNED>(a.A);
  auto kern = heads_fused_kernel<AMAX, H, LEARNED>;
  static int occ_cached = -1;   // per instantiation; the smem size only varies with A <= AMAX (use the worst case)
  static size_t smem_cached = 0;
  if (occ_cached < 0 || smem > smem_cached) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kRows, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return MNR_EUNSUP;
    occ_cached = occ, smem_cached = smem;
  }
  int grid = mnr_sm_count() * occ_cached;  // every CTA resident: the grid barrier needs it
  if (grid > grid_cap) grid = grid_cap;
  if (grid > a.n_tiles) grid = a.n_tiles;
  kern<<<grid, kRows, smem, stream>>>(a, cfg);
  const cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? MNR_OK : (int)e;
}

template <int AMAX, bool LEARNED>
int launch_by_h(FusedHeadsArgs& a, int H, const mnr_ppo_loss_cfg_t& cfg, int grid_cap, cudaStream_t stream) {
  switch (H) {
    case 32: return launch_fused<AMAX, 32, LEARNED>(a, cfg, grid_cap, stream);
    case 64: return launch_fused<AMAX, 64, LEARNED>(a, cfg, grid_cap, stream);
    case 128: return launch_fused<AMAX, 128, LEARNED>(a, cfg, grid_cap, stream);
    default: return MNR_EUNSUP;
  }
}

}  // namespace

// ---- internal interface used by mnr_ppo_heads_loss_fwd_bwd (csrc/ppo_loss.cu) ----------------------------------------
int mnr_heads_fused_grid_cap() { return mnr_sm_count() * 2; }

bool mnr_heads_fused_supported(int A, int H) { return (H == 32 || H == 64 || H == 128) && A >= 1 && A <= 16; }

size_t mnr_heads_fused_workspace_bytes(int A, int H) {
  const size_t cap = (size_t)mnr_heads_fused_grid_cap();
  const size_t PW = (size_t)(2 * A + 1) * H + 4 * A + 1;
  return 256 + ((cap * sizeof(mnr_loss::LossPartial) + 255) / 256) * 256 + cap * PW * sizeof(float);
}

int mnr_heads_fused_launch(const float* y3a, const float* y3v, const float* w_mu, const float* b_mu, const float* w_v,
                           const float* b_v, const float* w_sig, const float* b_sig, const float* actions,
                           const float* old_neglogp, const float* advantages, const float* old_values, const float* returns,
                           float* old_mu, float* old_sigma, int write_back, float* dy3a, float* dy3v, float* g_mu, float* g_ls,
                           float* g_v, float* dw_mu, float* db_mu, float* dw_v, float* db_v, float* dw_sig, float* db_sig,
                           float* metrics_out, float* mu_mean_out, float* sigma_mean_out, int64_t M, int A, int H,
                           const mnr_ppo_loss_cfg_t* cfg, void* workspace, cudaStream_t stream) {
  if (!mnr_aligned(y3a, 16) || !mnr_aligned(y3v, 16) || !mnr_aligned(dy3a, 16) || !mnr_aligned(dy3v, 16)) return MNR_EALIGN;
  const int cap = mnr_heads_fused_grid_cap();
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  FusedHeadsArgs a;
  a.y3a = y3a, a.y3v = y3v, a.w_mu = w_mu, a.b_mu = b_mu, a.w_v = w_v, a.b_v = b_v, a.w_sig = w_sig, a.b_sig = b_sig;
  a.actions = actions, a.old_nlp = old_neglogp, a.adv = advantages, a.old_values = old_values, a.returns = returns;
  a.old_mu = old_mu, a.old_sigma = old_sigma, a.dy3a = dy3a, a.dy3v = dy3v, a.g_mu = g_mu, a.g_ls = g_ls, a.g_v = g_v;
  a.dw_mu = dw_mu, a.db_mu = db_mu, a.dw_sig = dw_sig, a.db_sig = db_sig, a.dw_v = dw_v, a.db_v = db_v;
  a.metrics = metrics_out, a.mu_mean = mu_mean_out, a.sigma_mean = sigma_mean_out;
  a.barrier = reinterpret_cast<unsigned long long*>(ws);
  a.loss_parts = reinterpret_cast<mnr_loss::LossPartial*>(ws + 256);
  a.partials = reinterpret_cast<float*>(ws + 256 + (((size_t)cap * sizeof(mnr_loss::LossPartial) + 255) / 256) * 256);
  a.M = M, a.A = A, a.write_back = write_back, a.n_tiles = (int)((M + kRows - 1) / kRows);
  const bool learned = w_sig != nullptr;
  if (A <= 8) return learned ? launch_by_h<8, true>(a, H, *cfg, cap, stream) : launch_by_h<8, false>(a, H, *cfg, cap, stream);
  return learned ? launch_by_h<16, true>(a, H, *cfg, cap, stream) : launch_by_h<16, false>(a, H, *cfg, cap, stream);
}
