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
// Layout of the work.  A CTA = 128 threads owns tiles of 128 rows; the Y3a tile (and afterwards, in the same buffer, the
// Y3v tile) is staged ONCE in shared memory (coalesced float4 loads, row pitch H+4 floats so that "thread = row" LDS.128
// sweeps are conflict-free) and is used three times: (1) thread-per-row dot products with the head weights (broadcast
// LDS.128 from shared memory), (2) the weight-gradient outer products with thread = (4 hidden columns, row group)
// walking its rows of the tile, accumulating in registers ACROSS tiles, (3) it is overwritten in place by dY3 and leaves
// through coalesced float4 stores.  After the tile loop the per-CTA partials go to the workspace, the grid (all CTAs resident) meets at
// a grid barrier and every CTA reduces its share of the outputs over all CTAs' partials in a fixed order
// (deterministic, no atomics); CTA 0 also finalises the ten loss metrics.
#include "common.cuh"
#include "loss_math.cuh"

namespace {
using namespace mnr_loss;

constexpr int kRows = 128;     // rows per tile
constexpr int kThreadsHF = 256;  // 2 threads per row: thread (r, h) takes the float4 columns c = h, h + 2, ... of row r

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

// A (the action dimension padded to a multiple of 4; the real one is a.A at run time: pad rows of the head weights are
// zero in shared memory and the rolled loss loops stop at a.A), H compile-time (no guards in the unrolled dot products);
// one shared-memory tile buffer serves the actor tower
// and then the value tower of the same 128 rows, so a CTA needs ~58 KB and 3 CTAs (24 warps) share an SM.  The kernel is
// bound by per-thread instruction latency (every instruction runs once or twice per warp), hence TWO threads per row:
// each takes every other float4 column of the row in the dot-product / dY3 sweeps and every other action column in the
// loss loops; partial sums meet through one __shfl_xor.
template <int A, int H, bool LEARNED>
__global__ void __launch_bounds__(kThreadsHF, 2) heads_fused_kernel(const __grid_constant__ FusedHeadsArgs a,
                                                            const __grid_constant__ mnr_ppo_loss_cfg_t cfg) {
  constexpr int P = H + 8, P4 = P / 4, H4 = H / 4;  // pitch == 8 (mod 32): the 4 rows x 2 halves of an LDS.128 phase never collide
  constexpr int NY = (LEARNED ? 2 : 1) * A;          // gradient rows fed by Y3a: g_mu | g_ls
  constexpr int GP = 4 * A + 4;                      // s_g pitch: g_mu | g_ls | mu | sigma | g_v | pad  (float4-aligned groups,
                                                     // GP mod 32 == 4 or 20: "thread = row" float4 accesses are conflict-free)
  constexpr int GW = 4 * A + 1;                      // columns that get summed over rows
  constexpr int UV = NY * 4 <= 64 ? 4 : (NY * 2 <= 64 ? 2 : 1);  // hidden columns per thread in the weight-gradient phase
  constexpr int HU = H / UV, NRG = kThreadsHF / HU, RPG = kRows / NRG;  // threads per row group, row groups, rows per group
  constexpr int kWarpsHF = kThreadsHF / 32;
  static_assert(A % 4 == 0 && A <= 32 && NRG >= 1 && RPG >= 1 && NY * UV <= 64, "unsupported head shape");
  const int Ar = a.A;  // real action dimension, Ar <= A
  extern __shared__ __align__(16) float sh[];
  __shared__ double s_red[6][kWarpsHF];
  mnr::pdl_trigger();   // PDL: see common.cuh
  mnr::pdl_wait();
  __shared__ mnr::Moments s_mom[2][kWarpsHF];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* s_y = sh;                                   // [128][P]: Y3a tile, then dY3a, then Y3v tile, then dY3v
  float* s_wmu = s_y + kRows * P;                    // (A,H)
  float* s_wsig = s_wmu + A * H;                     // (A,H), learned only
  float* s_wv = s_wsig + (LEARNED ? A * H : 0);      // (H)
  float* s_g = s_wv + H;                             // [128][GP]

  for (int i = tid; i < A * H; i += kThreadsHF) {
    const bool real = i < Ar * H;
    s_wmu[i] = real ? __ldg(a.w_mu + i) : 0.f;
    if (LEARNED) s_wsig[i] = real ? __ldg(a.w_sig + i) : 0.f;
  }
  for (int i = tid; i < H; i += kThreadsHF) s_wv[i] = __ldg(a.w_v + i);
  float b_mu[A], b_sig[A];
#pragma unroll
  for (int j = 0; j < A; ++j) b_mu[j] = j < Ar ? __ldg(a.b_mu + j) : 0.f, b_sig[j] = j < Ar ? __ldg(a.b_sig + j) : 0.f;

  const float w = cfg.grad_scale / (float)a.M;
  const float b_v = __ldg(a.b_v);
  float t_a = 0.f, t_c = 0.f, t_b = 0.f, t_h = 0.f, t_clip = 0.f, t_kl = 0.f;
  mnr::Welford w_d, w_ret;
  w_d.init(), w_ret.init();

  // weight-gradient accumulators, live across tiles: thread = (UV hidden columns from uu, row group rg)
  const int uu = (tid % HU) * UV, rg = tid / HU;
  float accw[NY][UV], accv[UV], acc_col = 0.f;
#pragma unroll
  for (int j = 0; j < NY; ++j)
#pragma unroll
    for (int q = 0; q < UV; ++q) accw[j][q] = 0.f;
#pragma unroll
  for (int q = 0; q < UV; ++q) accv[q] = 0.f;

  float4* s_y4 = reinterpret_cast<float4*>(s_y);
  const float4* s_wmu4 = reinterpret_cast<const float4*>(s_wmu);
  const float4* s_wsig4 = reinterpret_cast<const float4*>(s_wsig);
  const float4* s_wv4 = reinterpret_cast<const float4*>(s_wv);

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * kRows;
    const int rows_here = (int)min((int64_t)kRows, a.M - row0);
    const int r = tid >> 1, h = tid & 1;
    const bool valid = r < rows_here;
    const int64_t row = valid ? row0 + r : a.M - 1;
    float* grow = s_g + r * GP;

    // =========================== actor tower ===========================================================
    __syncthreads();  // previous tile's stores out of s_y are done (and the weights are staged)
    {
      const float4* gy = reinterpret_cast<const float4*>(a.y3a);
#pragma unroll 8
      for (int i = tid; i < kRows * H4; i += kThreadsHF) {
        const int rr = i / H4, c = i - rr * H4;
        s_y4[rr * P4 + c] = rr < rows_here ? __ldg(gy + (row0 + rr) * H4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
    // ---- thread = (row, half): mu / logstd heads ------------------------------------------------------------
    float mu_raw[A], ls[A];
#pragma unroll
    for (int j = 0; j < A; ++j) mu_raw[j] = 0.f, ls[j] = 0.f;
#pragma unroll 2
    for (int c = h; c < H4; c += 2) {
      const float4 ya = s_y4[r * P4 + c];
#pragma unroll
      for (int j = 0; j < A; ++j) {
        mu_raw[j] = dot4(ya, s_wmu4[j * H4 + c], mu_raw[j]);
        if (LEARNED) ls[j] = dot4(ya, s_wsig4[j * H4 + c], ls[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < A; ++j) {
      mu_raw[j] += __shfl_xor_sync(0xffffffffu, mu_raw[j], 1);
      if (LEARNED) ls[j] += __shfl_xor_sync(0xffffffffu, ls[j], 1);
    }
    // ---- per-element loss terms.  ROLLED loops over the action columns (half h takes j = h, h + 2, ..) with the row's
    //      values parked in its own s_g row (slots [0,A) raw mu, [A,2A) raw logstd -> later g_mu, g_ls; [2A,3A) mu;
    //      [3A,4A) sigma): keeps the transcendental-heavy body in the code once instead of A times.
#pragma unroll
    for (int j = 0; j < A; ++j)
      if ((j & 1) == h) {
        grow[j] = mu_raw[j] + b_mu[j], grow[A + j] = LEARNED ? ls[j] + b_sig[j] : b_sig[j];
        if (j >= Ar) grow[2 * A + j] = 0.f, grow[3 * A + j] = 0.f;  // pad columns: defined (and ignored) in the column sums
      }
    float nlp = 0.f, ent = 0.f, bsum = 0.f, kl = 0.f;
    const float* act_row = a.actions + (size_t)row * Ar;
#pragma unroll 1
    for (int j = h; j < Ar; j += 2) {
      float mu, sg, dmu, dsg, db;
      dist_fwd(cfg, grow[j], grow[A + j], mu, sg, dmu, dsg);
      const float diff = __ldg(act_row + j) - mu;
      const float lsg = logf(sg);
      nlp += diff * diff / (2.f * sg * sg) + lsg + kLogSqrt2Pi;
      ent += 0.5f + kLogSqrt2Pi + lsg;
      bsum += bounds_elem(mu, cfg.bounds_type, cfg.soft_bound, &db);
      kl += kl_elem(mu, sg, a.old_mu[(size_t)row * Ar + j], a.old_sigma[(size_t)row * Ar + j]);
      grow[2 * A + j] = mu, grow[3 * A + j] = sg;
    }
    nlp += __shfl_xor_sync(0xffffffffu, nlp, 1), ent += __shfl_xor_sync(0xffffffffu, ent, 1);
    bsum += __shfl_xor_sync(0xffffffffu, bsum, 1), kl += __shfl_xor_sync(0xffffffffu, kl, 1);
    float dl_dnlp, clipped;
    const float a_loss = actor_row(__ldg(a.old_nlp + row), nlp, __ldg(a.adv + row), cfg.e_clip, cfg.use_smooth_clamp, &dl_dnlp, &clipped);
    if (valid && h == 0) t_a += a_loss, t_b += bsum, t_h += ent, t_clip += clipped, t_kl += kl;
    const float g_nlp = w * dl_dnlp;
    const float span = cfg.maxstd - cfg.minstd;
#pragma unroll 1
    for (int j = h; j < Ar; j += 2) {
      const float mu = grow[2 * A + j], sg = grow[3 * A + j];
      const float diff = __ldg(act_row + j) - mu;
      float dmu = 1.f, dsg = sg;                                  // d mu / d mu_raw, d sigma / d logstd (dist.py:32-47)
      if (cfg.dist_type != MNR_DIST_NORMAL) {
        const float sgm = (sg - cfg.minstd) / span;               // the sigmoid value dist_fwd produced
        dmu = 1.f - mu * mu, dsg = span * sgm * (1.f - sgm);
      }
      const float inv_var = 1.f / (sg * sg);
      float db;
      bounds_elem(mu, cfg.bounds_type, cfg.soft_bound, &db);
      const float g_m = g_nlp * (-diff * inv_var) + w * cfg.bounds_loss_coef * db;
      const float g_s = g_nlp * (1.f - diff * diff * inv_var) / sg - w * cfg.entropy_coef / sg;
      const float gm = valid ? g_m * dmu : 0.f, gl = valid ? g_s * dsg : 0.f;
      grow[j] = gm, grow[A + j] = gl;
      if (!valid) grow[2 * A + j] = 0.f, grow[3 * A + j] = 0.f;
    }
    __syncwarp();  // the pair's halves of the row are both in shared memory now
    if (valid && (Ar & 3) == 0) {  // row-contiguous float4 copies of the finished row pieces (half h takes every other piece)
#pragma unroll
      for (int j4 = 0; j4 < A / 4; ++j4) {
        if ((j4 & 1) != h || 4 * j4 >= Ar) continue;
        const float4 gm4 = reinterpret_cast<const float4*>(grow)[j4], gl4 = reinterpret_cast<const float4*>(grow + A)[j4];
        if (a.g_mu) reinterpret_cast<float4*>(a.g_mu + (size_t)row * Ar)[j4] = gm4;
        if (a.g_ls) reinterpret_cast<float4*>(a.g_ls + (size_t)row * Ar)[j4] = gl4;
        if (a.write_back) {
          reinterpret_cast<float4*>(a.old_mu + (size_t)row * Ar)[j4] = reinterpret_cast<const float4*>(grow + 2 * A)[j4];
          reinterpret_cast<float4*>(a.old_sigma + (size_t)row * Ar)[j4] = reinterpret_cast<const float4*>(grow + 3 * A)[j4];
        }
      }
    } else if (valid) {  // action dimension not a multiple of 4: rows are not 16 B aligned, element-wise copies
      for (int j = h; j < Ar; j += 2) {
        if (a.g_mu) a.g_mu[(size_t)row * Ar + j] = grow[j];
        if (a.g_ls) a.g_ls[(size_t)row * Ar + j] = grow[A + j];
        if (a.write_back) a.old_mu[(size_t)row * Ar + j] = grow[2 * A + j], a.old_sigma[(size_t)row * Ar + j] = grow[3 * A + j];
      }
    }
    __syncthreads();
    // ---- thread = (uu, rg): dW_mu / dW_sig outer products over this group's rows -----------------------------------
#pragma unroll 2
    for (int k = 0; k < RPG; ++k) {
      const int rr = rg * RPG + k;
      float yv_[UV];
      if (UV == 4) {
        const float4 t = *reinterpret_cast<const float4*>(s_y + rr * P + uu);
        yv_[0] = t.x, yv_[1] = t.y, yv_[2] = t.z, yv_[3] = t.w;
      } else {
        const float2 t = *reinterpret_cast<const float2*>(s_y + rr * P + uu);
        yv_[0] = t.x, yv_[1] = t.y;
      }
      const float4* g4 = reinterpret_cast<const float4*>(s_g + rr * GP);
#pragma unroll
      for (int j4 = 0; j4 < NY / 4; ++j4) {
        const float4 g = g4[j4];
        const float gq[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int x = 0; x < UV; ++x) accw[4 * j4 + q][x] = fmaf(gq[q], yv_[x], accw[4 * j4 + q][x]);
      }
    }
    __syncthreads();
    // ---- thread = (row, half): dY3a = g_mu Wmu + g_ls Wsig over the tile in place, then coalesced stores --------------
    {
      float gm[A], gl[A];
#pragma unroll
      for (int j = 0; j < A; ++j) gm[j] = grow[j], gl[j] = LEARNED ? grow[A + j] : 0.f;
#pragma unroll 2
      for (int c = h; c < H4; c += 2) {
        float4 da = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < A; ++j) {
          const float4 wm = s_wmu4[j * H4 + c];
          da.x = fmaf(gm[j], wm.x, da.x), da.y = fmaf(gm[j], wm.y, da.y), da.z = fmaf(gm[j], wm.z, da.z), da.w = fmaf(gm[j], wm.w, da.w);
          if (LEARNED) {
            const float4 ws = s_wsig4[j * H4 + c];
            da.x = fmaf(gl[j], ws.x, da.x), da.y = fmaf(gl[j], ws.y, da.y), da.z = fmaf(gl[j], ws.z, da.z), da.w = fmaf(gl[j], ws.w, da.w);
          }
        }
        s_y4[r * P4 + c] = da;
      }
    }
    __syncthreads();
    {
      float4* oa = reinterpret_cast<float4*>(a.dy3a);
      const float4* gy = reinterpret_cast<const float4*>(a.y3v);
#pragma unroll 8
      for (int i = tid; i < kRows * H4; i += kThreadsHF) {  // dY3a out, Y3v in: every thread reuses exactly the slots it just read
        const int rr = i / H4, c = i - rr * H4;
        if (rr < rows_here) oa[(row0 + rr) * H4 + c] = s_y4[rr * P4 + c];
        s_y4[rr * P4 + c] = rr < rows_here ? __ldg(gy + (row0 + rr) * H4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
    // =========================== value tower ===========================================================
    float value = 0.f;
#pragma unroll 4
    for (int c = h; c < H4; c += 2) value = dot4(s_y4[r * P4 + c], s_wv4[c], value);
    value += __shfl_xor_sync(0xffffffffu, value, 1);
    value += b_v;
    float dl_dv;
    const float ret = __ldg(a.returns + row);
    const float c_loss = critic_row(__ldg(a.old_values + row), value, ret, cfg.e_clip, cfg.clip_value_loss, &dl_dv);
    const float g_v = valid ? w * 0.5f * cfg.critic_coef * dl_dv : 0.f;
    if (valid && h == 0) {
      t_c += c_loss;
      w_d.push(ret - value), w_ret.push(ret);
      if (a.g_v) a.g_v[row] = g_v;
    }
    if (h == 0) grow[4 * A] = g_v;
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < RPG; ++k) {
      const int rr = rg * RPG + k;
      const float gv = s_g[rr * GP + 4 * A];
      if (UV == 4) {
        const float4 t = *reinterpret_cast<const float4*>(s_y + rr * P + uu);
        accv[0] = fmaf(gv, t.x, accv[0]), accv[1] = fmaf(gv, t.y, accv[1]), accv[2] = fmaf(gv, t.z, accv[2]), accv[3] = fmaf(gv, t.w, accv[3]);
      } else {
        const float2 t = *reinterpret_cast<const float2*>(s_y + rr * P + uu);
        accv[0] = fmaf(gv, t.x, accv[0]), accv[1] = fmaf(gv, t.y, accv[1]);
      }
    }
    if (tid < GW) {  // column sums of the g tile: db_mu | colsum g_ls | mu, sigma sums | db_v
      float sum = 0.f;
#pragma unroll 8
      for (int rr = 0; rr < kRows; ++rr) sum += s_g[rr * GP + tid];
      acc_col += sum;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = h; c < H4; c += 2) {
      const float4 wv = s_wv4[c];
      s_y4[r * P4 + c] = make_float4(g_v * wv.x, g_v * wv.y, g_v * wv.z, g_v * wv.w);
    }
    __syncthreads();
    {
      float4* ov = reinterpret_cast<float4*>(a.dy3v);
#pragma unroll 8
      for (int i = tid; i < kRows * H4; i += kThreadsHF) {
        const int rr = i / H4, c = i - rr * H4;
        if (rr < rows_here) ov[(row0 + rr) * H4 + c] = s_y4[rr * P4 + c];
      }
    }
  }

  // ---- fold the row groups (through the tile buffer), park the per-CTA partials ------------------------------------------
  const int PW = (2 * A + 1) * H + GW;
  float* part = a.partials + (size_t)blockIdx.x * PW;
  float* s_fold = s_y;  // [NRG][H]
#pragma unroll
  for (int j = 0; j <= NY; ++j) {
    __syncthreads();
#pragma unroll
    for (int x = 0; x < UV; ++x) s_fold[rg * H + uu + x] = j < NY ? accw[j < NY ? j : 0][x] : accv[x];
    __syncthreads();
    if (rg == 0) {
#pragma unroll
      for (int x = 0; x < UV; ++x) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < NRG; ++q) t += s_fold[q * H + uu + x];
        part[(j < NY ? j : 2 * A) * H + uu + x] = t;
      }
    }
  }
  if (!LEARNED)  // sigma rows of the layout stay defined
    for (int i = tid; i < A * H; i += kThreadsHF) part[A * H + i] = 0.f;
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
      for (int x = 0; x < kWarpsHF; ++x) ac[k] += s_red[k][x];
    }
    p.a = ac[0], p.c = ac[1], p.b = ac[2], p.h = ac[3], p.clip = ac[4], p.kl = ac[5];
    p.d = s_mom[0][0], p.ret = s_mom[1][0];
    for (int x = 1; x < kWarpsHF; ++x) p.d = mnr::moments_merge(p.d, s_mom[0][x]), p.ret = mnr::moments_merge(p.ret, s_mom[1][x]);
    a.loss_parts[blockIdx.x] = p;
  }

  mnr::grid_barrier(a.barrier);

  // ---- every CTA reduces its share of the PW outputs over all CTAs (warp per output, fixed order) ----------------
  const int n_cta = gridDim.x;
  for (int o = blockIdx.x * kWarpsHF + warp; o < PW; o += n_cta * kWarpsHF) {
    float sum = 0.f;
    for (int x = lane; x < n_cta; x += 32) sum += __ldcg(a.partials + (size_t)x * PW + o);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) {
      if (o < (2 * A + 1) * H) {
        const int jrow = o / H, u2 = o - jrow * H;
        if (jrow < A) { if (jrow < Ar) a.dw_mu[jrow * H + u2] = sum; }
        else if (jrow < 2 * A) { if (LEARNED && jrow - A < Ar) a.dw_sig[(jrow - A) * H + u2] = sum; }
        else a.dw_v[u2] = sum;
      } else {
        const int j = o - (2 * A + 1) * H;
        if (j < A) { if (j < Ar) a.db_mu[j] = sum; }
        else if (j < 2 * A) { if (j - A < Ar) a.db_sig[j - A] = sum; }  // learned: bias gradient; fixed: d sigma = colsum g_ls
        else if (j < 3 * A) { if (a.mu_mean && j - 2 * A < Ar) a.mu_mean[j - 2 * A] = sum / (float)a.M; }
        else if (j < 4 * A) { if (a.sigma_mean && j - 3 * A < Ar) a.sigma_mean[j - 3 * A] = sum / (float)a.M; }
        else a.db_v[0] = sum;
      }
    }
  }
  if (blockIdx.x != 0) return;

  // ---- CTA 0: the ten loss metrics ---------------------------------------------------------------------------------
  double acc2[6] = {0, 0, 0, 0, 0, 0};
  mnr::Moments md2{0, 0, 0}, mr2{0, 0, 0};
  for (int x = tid; x < n_cta; x += kThreadsHF) {
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

template <int A, int H, bool LEARNED>
size_t fused_smem_bytes() {
  return (size_t)(kRows * (H + 8) + (LEARNED ? 2 : 1) * A * H + H + kRows * (4 * A + 4)) * sizeof(float);
}

template <int A, int H, bool LEARNED>
int launch_fused(FusedHeadsArgs& a, const mnr_ppo_loss_cfg_t& cfg, int grid_cap, cudaStream_t stream) {
  const size_t smem = fused_smem_bytes<A, H, LEARNED>();
  auto kern = heads_fused_kernel<A, H, LEARNED>;
  static int occ_cached = -1;   // per instantiation
  if (occ_cached < 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreadsHF, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return MNR_EUNSUP;
    occ_cached = occ;
  }
  int grid = mnr_sm_count() * occ_cached;  // every CTA resident: the grid barrier needs it
  if (grid > grid_cap) grid = grid_cap;
  if (grid > a.n_tiles) grid = a.n_tiles;
  cudaError_t e = mnr_launch_pdl(kern, dim3(grid), dim3(kThreadsHF), smem, stream, a, cfg);
  if (e == cudaSuccess) e = cudaPeekAtLastError();
  return e == cudaSuccess ? MNR_OK : (int)e;
}

template <int A, bool LEARNED>
int launch_by_h(FusedHeadsArgs& a, int H, const mnr_ppo_loss_cfg_t& cfg, int grid_cap, cudaStream_t stream) {
  switch (H) {
    case 32: return launch_fused<A, 32, LEARNED>(a, cfg, grid_cap, stream);
    case 64: return launch_fused<A, 64, LEARNED>(a, cfg, grid_cap, stream);
    case 128: return launch_fused<A, 128, LEARNED>(a, cfg, grid_cap, stream);
    default: return MNR_EUNSUP;
  }
}

// padded action dimension the kernel is instantiated for: 4, 8, 12, 16, 24, 32
int mnr_heads_fused_apad(int A) {
  const int p = (A + 3) / 4 * 4;
  return p <= 16 ? p : (p <= 24 ? 24 : 32);
}

template <bool LEARNED>
int launch_by_a(FusedHeadsArgs& a, int H, const mnr_ppo_loss_cfg_t& cfg, int grid_cap, cudaStream_t stream) {
  switch (mnr_heads_fused_apad(a.A)) {
    case 4: return launch_by_h<4, LEARNED>(a, H, cfg, grid_cap, stream);
    case 8: return launch_by_h<8, LEARNED>(a, H, cfg, grid_cap, stream);
    case 12: return launch_by_h<12, LEARNED>(a, H, cfg, grid_cap, stream);
    case 16: return launch_by_h<16, LEARNED>(a, H, cfg, grid_cap, stream);
    case 24: return launch_by_h<24, LEARNED>(a, H, cfg, grid_cap, stream);
    case 32: return launch_by_h<32, LEARNED>(a, H, cfg, grid_cap, stream);
    default: return MNR_EUNSUP;
  }
}

}  // namespace

// ---- internal interface used by mnr_ppo_heads_loss_fwd_bwd (csrc/ppo_loss.cu) ----------------------------------------
int mnr_heads_fused_grid_cap() { return mnr_sm_count() * 4; }

bool mnr_heads_fused_supported(int A, int H) { return (H == 32 || H == 64 || H == 128) && A >= 1 && A <= 32; }

size_t mnr_heads_fused_workspace_bytes(int A_real, int H) {
  const size_t cap = (size_t)mnr_heads_fused_grid_cap();
  const size_t A = (size_t)mnr_heads_fused_apad(A_real);
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
  if (A % 4 == 0 && (!mnr_aligned(actions, 16) || !mnr_aligned(old_mu, 16) || !mnr_aligned(old_sigma, 16) ||
                     (g_mu && !mnr_aligned(g_mu, 16)) || (g_ls && !mnr_aligned(g_ls, 16))))
    return MNR_EALIGN;
  return w_sig != nullptr ? launch_by_a<true>(a, H, *cfg, cap, stream) : launch_by_a<false>(a, H, *cfg, cap, stream);
}
