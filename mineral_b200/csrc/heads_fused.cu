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
// Layout of the work.  A CTA = 256 threads owns tiles of 256 / TPR rows (TPR threads per row: 2, 4 or 8, chosen by the host
// from the minibatch size and the head width - see kThreadsHF and launch_fused); the Y3a tile (and the Y3v tile: in its own
// prefetched buffer, or afterwards in the same one) is staged ONCE in shared memory (coalesced float4 loads, row pitch H+8
// floats so that "thread = row piece" LDS.128 sweeps are conflict-free) and is used three times: (1) per-row dot products
// with the head weights (broadcast LDS.128 from shared memory), (2) the weight-gradient outer products with thread =
// (UV hidden columns, row group) walking its rows of the tile, accumulating in registers ACROSS tiles, (3) it is
// overwritten in place by dY3 and leaves through coalesced float4 stores.  After the tile loop the per-CTA partials go to
// the workspace; then EITHER the kernel ends there and the sums over the CTAs run inside the split-reduce launch (deferred
// folds, mnr_fold_defer: the PPO update's path) OR the grid (all CTAs resident) meets at a grid barrier and every CTA
// reduces its share of the outputs over all CTAs' partials in a fixed order, CTA 0 also finalising the ten loss metrics.
// Deterministic either way, no atomics.
#include "common.cuh"
#include "loss_math.cuh"

namespace {
using namespace mnr_loss;

// optional timeline of CTA 0 (debug aid): globaltimer at the phase boundaries of its first tile
__device__ long long g_htrace[32];
__device__ int g_htrace_on = 0;
#define HF_TRACE(i)                                                                            \
  do {                                                                                         \
    if (g_htrace_on && blockIdx.x == 0 && threadIdx.x == 0) {                                  \
      long long t_;                                                                            \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                   \
      g_htrace[i] = t_;                                                                        \
    }                                                                                          \
  } while (0)

constexpr int kThreadsHF = 256;  // TPR threads per row: thread (r, h) takes the float4 columns c = h, h + TPR, ... of row r
// rows per tile = kThreadsHF / TPR: TPR = 2 (128-row tiles) when the minibatch fills the GPU with them; TPR = 8 (32-row tiles)
// for small minibatches (the reference's shipped recipes: 2048 rows = 16 tiles of 128 on 148 SMs), where the kernel is ONE
// tile's dependent chain: 4x the CTAs, a quarter of the instructions per thread

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
  int defer;   // deferred fold (mnr_fold_defer): stop after parking the per-CTA partials
};

// 16 B asynchronous global -> shared copy (LDGSTS), zero fill when !valid
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// CTA-wide sums of NV doubles per thread, all reduced together: warp shuffles, one slot per (value, warp), NV threads add
// the warps in order; every thread receives the totals.  s_buf: NV x NW doubles.  Fixed order => deterministic.
template <int NV, int NW>
__device__ __forceinline__ void block_sum_n(double (&v)[NV], double* s_buf, int tid, double* totals_out = nullptr) {
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = mnr::warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) s_buf[k * NW + warp] = v[k];
  }
  __syncthreads();
  if (tid < NV) {
    double t = 0.0;
    for (int w = 0; w < NW; ++w) t += s_buf[tid * NW + w];
    s_buf[tid * NW] = t;
    if (totals_out) totals_out[tid] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = s_buf[k * NW];
  __syncthreads();
}

// act'(y) from the layer OUTPUT y (cfg.dz_act: 1 ELU with alpha = 1, 2 ReLU) - the same forms as tc_common.cuh's act_bwd_from_y
__device__ __forceinline__ float dz_factor(float y, int act) {
  return y > 0.f ? 1.f : (act == 1 ? y + 1.f : 0.f);
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc), acc = fmaf(a.y, b.y, acc), acc = fmaf(a.z, b.z, acc), acc = fmaf(a.w, b.w, acc);
  return acc;
}

// A (the action dimension padded to a multiple of 4; the real one is a.A at run time: pad rows of the head weights are
// zero in shared memory and the rolled loss loops stop at a.A), H compile-time (no guards in the unrolled dot products);
// two CTAs (16 warps) share an SM.  The kernel is bound by per-thread instruction latency (every instruction runs once or
// twice per warp), hence TPR >= 2 threads per row: thread h of a row takes the float4 columns h, h + TPR, .. in the
// dot-product / dY3 sweeps and the action columns h, h + TPR, .. in the loss loops; partial sums meet through log2(TPR)
// __shfl_xor steps (a row's threads are an aligned group of one warp).
template <int A, int H, bool LEARNED, int TPR>
__global__ void __launch_bounds__(kThreadsHF, 2) heads_fused_kernel(const __grid_constant__ FusedHeadsArgs a,
                                                            const __grid_constant__ mnr_ppo_loss_cfg_t cfg) {
  constexpr int kRows = kThreadsHF / TPR;
  constexpr int P = H + 8, P4 = P / 4, H4 = H / 4;  // pitch == 8 (mod 32): the 4 rows x 2 halves of an LDS.128 phase never collide
  constexpr int NY = (LEARNED ? 2 : 1) * A;          // gradient rows fed by Y3a: g_mu | g_ls
  constexpr int GP = 4 * A + 4;                      // s_g pitch: g_mu | g_ls | mu | sigma | g_v | pad  (float4-aligned groups,
                                                     // GP mod 32 == 4 or 20: "thread = row" float4 accesses are conflict-free)
  constexpr int GW = 4 * A + 1;                      // columns that get summed over rows
  constexpr int UV = NY * 4 <= 64 ? 4 : (NY * 2 <= 64 ? 2 : 1);  // hidden columns per thread in the weight-gradient phase
  constexpr int HU = H / UV, NRG = kThreadsHF / HU, RPG = kRows / NRG;  // threads per row group, row groups, rows per group
  constexpr int kWarpsHF = kThreadsHF / 32;
  constexpr int CS = kThreadsHF / GW, CSR = (kRows + CS - 1) / CS;   // column sums: CS row slices of CSR rows
  static_assert(A % 4 == 0 && A <= 32 && NRG >= 1 && RPG >= 1 && NY * UV <= 64, "unsupported head shape");
  static_assert(TPR >= 2 && TPR <= 32 && (TPR & (TPR - 1)) == 0, "a row's threads are an aligned group inside one warp");
  const int Ar = a.A;  // real action dimension, Ar <= A
  extern __shared__ __align__(16) float sh[];
  __shared__ double s_red[kThreadsHF], s_fin[16];
  mnr::pdl_trigger();   // PDL: see common.cuh
  mnr::pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // PREF_V (H <= 64): the Y3v tile has its own buffer and is requested (cp.async) together with the Y3a tile at the top of
  // the tile, so the DRAM round trip in the middle of the tile is gone; wider towers keep one buffer (2 CTAs per SM).
  constexpr bool PREF_V = H <= 64 || kRows <= 64;
  float* s_y = sh;                                   // [128][P]: Y3a tile, then dY3a (!PREF_V: then Y3v tile, then dY3v)
  float* s_yv = s_y + (PREF_V ? kRows * P : 0);      // [128][P]: Y3v tile, then dY3v
  float* s_wmu = s_yv + kRows * P;                   // (A,H)
  float* s_wsig = s_wmu + A * H;                     // (A,H), learned only
  float* s_wv = s_wsig + (LEARNED ? A * H : 0);      // (H)
  float* s_g = s_wv + H;                             // [128][GP]

  HF_TRACE(0);
  auto stage_weights = [&]() {
    for (int i = tid; i < A * H; i += kThreadsHF) {
      const bool real = i < Ar * H;
      s_wmu[i] = real ? __ldg(a.w_mu + i) : 0.f;
      if (LEARNED) s_wsig[i] = real ? __ldg(a.w_sig + i) : 0.f;
    }
    for (int i = tid; i < H; i += kThreadsHF) s_wv[i] = __ldg(a.w_v + i);
  };
  if (!PREF_V) stage_weights();   // PREF_V: after the first tile's cp.async requests are in flight (below)
  __shared__ float b_mu[A], b_sig[A];   // head biases (fixed sigma: b_sig = the sigma parameter): read once per row and tile
  if (tid < A) b_mu[tid] = tid < Ar ? __ldg(a.b_mu + tid) : 0.f, b_sig[tid] = tid < Ar ? __ldg(a.b_sig + tid) : 0.f;

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
  float4* s_v4 = reinterpret_cast<float4*>(s_yv);   // == s_y4 without PREF_V
  const float4* s_wmu4 = reinterpret_cast<const float4*>(s_wmu);
  const float4* s_wsig4 = reinterpret_cast<const float4*>(s_wsig);
  const float4* s_wv4 = reinterpret_cast<const float4*>(s_wv);

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * kRows;
    const int rows_here = (int)min((int64_t)kRows, a.M - row0);
    const int r = tid / TPR, h = tid % TPR;
    const bool valid = r < rows_here;
    const int64_t row = valid ? row0 + r : a.M - 1;
    float* grow = s_g + r * GP;

    // =========================== actor tower ===========================================================
    __syncthreads();  // previous tile's stores out of s_y are done (and the weights are staged)
    if (PREF_V) {
      const float4* gya = reinterpret_cast<const float4*>(a.y3a);
      const float4* gyv = reinterpret_cast<const float4*>(a.y3v);
#pragma unroll 4
      for (int i = tid; i < kRows * H4; i += kThreadsHF) {
        const int rr = i / H4, c = i - rr * H4;
        const bool ok = rr < rows_here;
        cp_async16(&s_y4[rr * P4 + c], gya + (row0 + (ok ? rr : 0)) * H4 + c, ok);
      }
      cp_async_commit();
#pragma unroll 4
      for (int i = tid; i < kRows * H4; i += kThreadsHF) {
        const int rr = i / H4, c = i - rr * H4;
        const bool ok = rr < rows_here;
        cp_async16(&s_v4[rr * P4 + c], gyv + (row0 + (ok ? rr : 0)) * H4 + c, ok);
      }
      cp_async_commit();
      if (tile == (int)blockIdx.x) stage_weights();
      cp_async_wait<1>();   // the Y3a tile has landed; Y3v keeps flying
    } else {
      const float4* gy = reinterpret_cast<const float4*>(a.y3a);
#pragma unroll 8
      for (int i = tid; i < kRows * H4; i += kThreadsHF) {
        const int rr = i / H4, c = i - rr * H4;
        s_y4[rr * P4 + c] = rr < rows_here ? __ldg(gy + (row0 + rr) * H4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
    HF_TRACE(1);
    // ---- thread = (row, half): mu / logstd heads ------------------------------------------------------------
    float mu_raw[A], ls[A];
#pragma unroll
    for (int j = 0; j < A; ++j) mu_raw[j] = 0.f, ls[j] = 0.f;
#pragma unroll 2
    for (int c = h; c < H4; c += TPR) {
      const float4 ya = s_y4[r * P4 + c];
#pragma unroll
      for (int j = 0; j < A; ++j) {
        mu_raw[j] = dot4(ya, s_wmu4[j * H4 + c], mu_raw[j]);
        if (LEARNED) ls[j] = dot4(ya, s_wsig4[j * H4 + c], ls[j]);
      }
    }
#pragma unroll
    for (int d = 1; d < TPR; d <<= 1)
#pragma unroll
      for (int j = 0; j < A; ++j) {
        mu_raw[j] += __shfl_xor_sync(0xffffffffu, mu_raw[j], d);
        if (LEARNED) ls[j] += __shfl_xor_sync(0xffffffffu, ls[j], d);
      }
    // ---- per-element loss terms.  ROLLED loops over the action columns (half h takes j = h, h + 2, ..) with the row's
    //      values parked in its own s_g row (slots [0,A) raw mu, [A,2A) raw logstd -> later g_mu, g_ls; [2A,3A) mu;
    //      [3A,4A) sigma): keeps the transcendental-heavy body in the code once instead of A times.
#pragma unroll
    for (int j = 0; j < A; ++j)
      if ((j % TPR) == h) {
        grow[j] = mu_raw[j] + b_mu[j], grow[A + j] = LEARNED ? ls[j] + b_sig[j] : b_sig[j];
        if (j >= Ar) grow[2 * A + j] = 0.f, grow[3 * A + j] = 0.f;  // pad columns: defined (and ignored) in the column sums
      }
    HF_TRACE(2);
    float nlp = 0.f, ent = 0.f, bsum = 0.f, kl = 0.f;
    const float* act_row = a.actions + (size_t)row * Ar;
#pragma unroll 1
    for (int j = h; j < Ar; j += TPR) {
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
#pragma unroll
    for (int d = 1; d < TPR; d <<= 1) {
      nlp += __shfl_xor_sync(0xffffffffu, nlp, d), ent += __shfl_xor_sync(0xffffffffu, ent, d);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, d), kl += __shfl_xor_sync(0xffffffffu, kl, d);
    }
    HF_TRACE(3);
    float dl_dnlp, clipped;
    const float a_loss = actor_row(__ldg(a.old_nlp + row), nlp, __ldg(a.adv + row), cfg.e_clip, cfg.use_smooth_clamp, &dl_dnlp, &clipped);
    if (valid && h == 0) t_a += a_loss, t_b += bsum, t_h += ent, t_clip += clipped, t_kl += kl;
    const float g_nlp = w * dl_dnlp;
    const float span = cfg.maxstd - cfg.minstd;
#pragma unroll 1
    for (int j = h; j < Ar; j += TPR) {
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
    HF_TRACE(4);
    __syncwarp();  // the pair's halves of the row are both in shared memory now
    if (valid && (Ar & 3) == 0) {  // row-contiguous float4 copies of the finished row pieces (half h takes every other piece)
#pragma unroll
      for (int j4 = 0; j4 < A / 4; ++j4) {
        if ((j4 % TPR) != h || 4 * j4 >= Ar) continue;
        const float4 gm4 = reinterpret_cast<const float4*>(grow)[j4], gl4 = reinterpret_cast<const float4*>(grow + A)[j4];
        if (a.g_mu) reinterpret_cast<float4*>(a.g_mu + (size_t)row * Ar)[j4] = gm4;
        if (a.g_ls) reinterpret_cast<float4*>(a.g_ls + (size_t)row * Ar)[j4] = gl4;
        if (a.write_back) {
          reinterpret_cast<float4*>(a.old_mu + (size_t)row * Ar)[j4] = reinterpret_cast<const float4*>(grow + 2 * A)[j4];
          reinterpret_cast<float4*>(a.old_sigma + (size_t)row * Ar)[j4] = reinterpret_cast<const float4*>(grow + 3 * A)[j4];
        }
      }
    } else if (valid) {  // action dimension not a multiple of 4: rows are not 16 B aligned, element-wise copies
      for (int j = h; j < Ar; j += TPR) {
        if (a.g_mu) a.g_mu[(size_t)row * Ar + j] = grow[j];
        if (a.g_ls) a.g_ls[(size_t)row * Ar + j] = grow[A + j];
        if (a.write_back) a.old_mu[(size_t)row * Ar + j] = grow[2 * A + j], a.old_sigma[(size_t)row * Ar + j] = grow[3 * A + j];
      }
    }
    __syncthreads();
    HF_TRACE(5);
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
    HF_TRACE(6);
    // ---- thread = (row, half): dY3a = g_mu Wmu + g_ls Wsig over the tile in place, then coalesced stores --------------
    {
      float gm[A], gl[A];
#pragma unroll
      for (int j = 0; j < A; ++j) gm[j] = grow[j], gl[j] = LEARNED ? grow[A + j] : 0.f;
#pragma unroll 2
      for (int c = h; c < H4; c += TPR) {
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
        if (cfg.dz_act) {   // dZ3 = dY3 * act'(Y3): the Y3a value is still in the slot that is about to be overwritten
          const float4 y = s_y4[r * P4 + c];
          da.x *= dz_factor(y.x, cfg.dz_act), da.y *= dz_factor(y.y, cfg.dz_act), da.z *= dz_factor(y.z, cfg.dz_act), da.w *= dz_factor(y.w, cfg.dz_act);
        }
        s_y4[r * P4 + c] = da;
      }
    }
    __syncthreads();
    HF_TRACE(7);
    {
      float4* oa = reinterpret_cast<float4*>(a.dy3a);
      const float4* gy = reinterpret_cast<const float4*>(a.y3v);
#pragma unroll 8
      for (int i = tid; i < kRows * H4; i += kThreadsHF) {  // dY3a out, Y3v in: every thread reuses exactly the slots it just read
        const int rr = i / H4, c = i - rr * H4;
        if (rr < rows_here) oa[(row0 + rr) * H4 + c] = s_y4[rr * P4 + c];
        if (!PREF_V) s_y4[rr * P4 + c] = rr < rows_here ? __ldg(gy + (row0 + rr) * H4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (PREF_V) cp_async_wait<0>();
    __syncthreads();
    HF_TRACE(8);
    // =========================== value tower ===========================================================
    float value = 0.f;
#pragma unroll 4
    for (int c = h; c < H4; c += TPR) value = dot4(s_v4[r * P4 + c], s_wv4[c], value);
#pragma unroll
    for (int d = 1; d < TPR; d <<= 1) value += __shfl_xor_sync(0xffffffffu, value, d);
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
    HF_TRACE(9);
#pragma unroll 4
    for (int k = 0; k < RPG; ++k) {
      const int rr = rg * RPG + k;
      const float gv = s_g[rr * GP + 4 * A];
      if (UV == 4) {
        const float4 t = *reinterpret_cast<const float4*>(s_yv + rr * P + uu);
        accv[0] = fmaf(gv, t.x, accv[0]), accv[1] = fmaf(gv, t.y, accv[1]), accv[2] = fmaf(gv, t.z, accv[2]), accv[3] = fmaf(gv, t.w, accv[3]);
      } else {
        const float2 t = *reinterpret_cast<const float2*>(s_yv + rr * P + uu);
        accv[0] = fmaf(gv, t.x, accv[0]), accv[1] = fmaf(gv, t.y, accv[1]);
      }
    }
    if (tid < CS * GW) {  // column sums of the g tile: db_mu | colsum g_ls | mu, sigma sums | db_v.  thread = (row slice, column)
      const int cs = tid / GW, cc = tid - cs * GW;
      float sum = 0.f;
#pragma unroll 4
      for (int rr = cs * CSR; rr < min(kRows, (cs + 1) * CSR); ++rr) sum += s_g[rr * GP + cc];
      acc_col += sum;
    }
    __syncthreads();
    HF_TRACE(10);
#pragma unroll 4
    for (int c = h; c < H4; c += TPR) {
      const float4 wv = s_wv4[c];
      float4 dv = make_float4(g_v * wv.x, g_v * wv.y, g_v * wv.z, g_v * wv.w);
      if (cfg.dz_act) {
        const float4 y = s_v4[r * P4 + c];
        dv.x *= dz_factor(y.x, cfg.dz_act), dv.y *= dz_factor(y.y, cfg.dz_act), dv.z *= dz_factor(y.z, cfg.dz_act), dv.w *= dz_factor(y.w, cfg.dz_act);
      }
      s_v4[r * P4 + c] = dv;
    }
    __syncthreads();
    {
      float4* ov = reinterpret_cast<float4*>(a.dy3v);
#pragma unroll 8
      for (int i = tid; i < kRows * H4; i += kThreadsHF) {
        const int rr = i / H4, c = i - rr * H4;
        if (rr < rows_here) ov[(row0 + rr) * H4 + c] = s_v4[rr * P4 + c];
      }
    }
  }

  HF_TRACE(11);
  // ---- fold the row groups (through the tile buffer), park the per-CTA partials ------------------------------------------
  const int PW = (2 * A + 1) * H + GW;
  float* part = a.partials + (size_t)blockIdx.x * PW;
  // every thread parks its accumulators of as many gradient rows as the tile buffer holds ([row][row group][H]), ONE barrier,
  // then all threads sum the row groups of their outputs in a fixed order (was: one gradient row per pass, two barriers each)
  float* s_fold = s_y;
  constexpr int FR_MAX = ((PREF_V ? 2 : 1) * kRows * P) / (NRG * H);   // PREF_V: the two tile buffers are contiguous
  constexpr int FR = FR_MAX < NY + 1 ? FR_MAX : NY + 1, NPASS = (NY + 1 + FR - 1) / FR;
  static_assert(FR >= 1, "tile buffer too small for the fold");
#pragma unroll
  for (int pass = 0; pass < NPASS; ++pass) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j <= NY; ++j) {
      if (j / FR != pass) continue;
#pragma unroll
      for (int x = 0; x < UV; ++x) s_fold[((j - pass * FR) * NRG + rg) * H + uu + x] = j < NY ? accw[j < NY ? j : 0][x] : accv[x];
    }
    __syncthreads();
    const int nj = min(FR, NY + 1 - pass * FR);
    for (int o = tid; o < nj * H; o += kThreadsHF) {
      const int jr = o / H, u2 = o - jr * H, j = pass * FR + jr;
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < NRG; ++q) t += s_fold[(jr * NRG + q) * H + u2];
      part[(j < NY ? j : 2 * A) * H + u2] = t;
    }
  }
  if (!LEARNED)  // sigma rows of the layout stay defined
    for (int i = tid; i < A * H; i += kThreadsHF) part[A * H + i] = 0.f;
  __syncthreads();
  if (tid < CS * GW) s_fold[tid] = acc_col;   // [slice][column]
  __syncthreads();
  if (tid < GW) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < CS; ++q) t += s_fold[q * GW + tid];
    part[(2 * A + 1) * H + tid] = t;
  }

  // per-CTA loss partial (12 doubles, layout at the metrics code below): all threads' values reduced together in f64
  HF_TRACE(16);
  {
    const double nd = (double)w_d.n, md_ = (double)w_d.mean, nr = (double)w_ret.n, mr_ = (double)w_ret.mean;
    double v1[12] = {(double)t_a, (double)t_c, (double)t_b, (double)t_h, (double)t_clip, (double)t_kl,
                     nd, nd * md_, (double)w_d.m2 + nd * md_ * md_, nr, nr * mr_, (double)w_ret.m2 + nr * mr_ * mr_};
    block_sum_n<12, kWarpsHF>(v1, s_red, tid, s_fin);
    if (tid < 12) reinterpret_cast<double*>(a.loss_parts)[(size_t)blockIdx.x * 12 + tid] = s_fin[tid];
  }

  HF_TRACE(12);
  if (a.defer) return;   // the sums over the CTAs run inside the split-reduce launch (tc_linear.cu, fold blocks)
  mnr::grid_barrier(a.barrier);
  HF_TRACE(13);

  // ---- every CTA reduces its share of the PW outputs over all CTAs (warp per output, fixed order) ----------------
  const int n_cta = gridDim.x;
  auto put = [&](int o, float sum) {
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
  };
  if (TPR >= 8) {
    // small minibatches, few CTAs (a compile-time choice: each variant carries one form): the warp-per-output form below is
    // a chain of PW / (8 n_cta) dependent L2 round trips per warp (11 us at 16 CTAs).
    // 8 lanes per output (CTA partials x = q, q + 8, ..), the lanes' sums folded by three shuffles: <= 8 loads per lane
    const int q = tid & 7;
    for (int o0 = (blockIdx.x * kThreadsHF + tid) >> 3; o0 - ((tid >> 3) & 3) < PW; o0 += (n_cta * kThreadsHF) >> 3) {
      const int o = o0 < PW ? o0 : PW - 1;   // whole warps stay in the loop (shuffles); the clamped lanes do not write
      float sum = 0.f;
#pragma unroll 8
      for (int x = q; x < n_cta; x += 8) sum += __ldcg(a.partials + (size_t)x * PW + o);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1), sum += __shfl_xor_sync(0xffffffffu, sum, 2), sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      if (q == 0 && o0 < PW) put(o, sum);
    }
  } else {
    for (int o = blockIdx.x * kWarpsHF + warp; o < PW; o += n_cta * kWarpsHF) {
      float sum = 0.f;
      for (int x = lane; x < n_cta; x += 32) sum += __ldcg(a.partials + (size_t)x * PW + o);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
      if (lane == 0) put(o, sum);
    }
  }
  HF_TRACE(14);
  if (blockIdx.x != 0) return;

  // ---- CTA 0: the ten loss metrics ---------------------------------------------------------------------------------
  // loss_parts[x] = 12 doubles {a, c, b, h, clip, kl, n_d, S_d, Q_d, n_r, S_r, Q_r} with S = sum x, Q = sum x^2 (in f64, from
  // the threads' Welford triples): the variances follow from the totals, var = (Q - S^2 / n) / (n - 1), without the chain
  // of Chan merges (one f64 division each) that took 16 us of this kernel's 49 when thread 0 ran it alone.
  // thread = (value k = tid & 15, slice = tid >> 4): 16 slices x n_cta / 16 partials, loads independent, fixed order.
  {
    const int k = tid & 15, sl = tid >> 4;
    double acc = 0.0;
    if (k < 12) {
      const double* lp = reinterpret_cast<const double*>(a.loss_parts);
#pragma unroll 4
      for (int x = sl; x < n_cta; x += kThreadsHF / 16) acc += __ldcg(lp + (size_t)x * 12 + k);
    }
    s_red[sl * 16 + k] = acc;
    __syncthreads();
    if (tid < 12) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < kThreadsHF / 16; ++q) t += s_red[q * 16 + tid];
      s_fin[tid] = t;
    }
    __syncthreads();
  }
  HF_TRACE(17);
  if (tid == 0) {
    const double n = (double)a.M, inv_n = 1.0 / n;
    double tot[6];
    for (int k = 0; k < 6; ++k) tot[k] = s_fin[k];
    const mnr::Moments td{s_fin[6], 0.0, s_fin[8] - s_fin[7] * s_fin[7] / s_fin[6]}, tr{s_fin[9], 0.0, s_fin[11] - s_fin[10] * s_fin[10] / s_fin[9]};
    const float av = (float)(tot[0] * inv_n), c = (float)(tot[1] * inv_n), b = (float)(tot[2] * inv_n), h = (float)(tot[3] * inv_n);
    float m[MNR_PPO_METRICS];
    m[0] = av + 0.5f * c * cfg.critic_coef - h * cfg.entropy_coef + b * cfg.bounds_loss_coef;
    m[1] = av, m[2] = c, m[3] = b, m[4] = h;
    m[5] = (float)(tot[4] * inv_n);
    const float var_d = (float)(td.m2 / (td.n - 1.0)), var_r = (float)(tr.m2 / (tr.n - 1.0));
    m[6] = fminf(fmaxf(1.f - var_d / (var_r + 1e-8f), 0.f), 1.f);
    m[7] = (float)(tot[5] * inv_n);
    m[8] = 0.f, m[9] = 0.f;
    for (int k = 0; k < MNR_PPO_METRICS; ++k) a.metrics[k] = m[k] * cfg.metric_scale;
    HF_TRACE(15);
  }
}

template <int A, int H, bool LEARNED, int TPR>
constexpr size_t fused_smem_bytes() {
  constexpr int kRows = kThreadsHF / TPR;
  return (size_t)((H <= 64 || kRows <= 64 ? 2 : 1) * kRows * (H + 8) + (LEARNED ? 2 : 1) * A * H + H + kRows * (4 * A + 4)) * sizeof(float);
}

// 32-row tiles while those still leave SMs idle (see kThreadsHF), else 128-row tiles
bool small_tiles_for(int64_t M) { return (M + 31) / 32 <= (int64_t)mnr_sm_count() * 2; }

template <int A, int H, bool LEARNED, int TPR>
int launch_fused_t(FusedHeadsArgs& a, const mnr_ppo_loss_cfg_t& cfg, int grid_cap, cudaStream_t stream) {
  const size_t smem = fused_smem_bytes<A, H, LEARNED, TPR>();
  auto kern = heads_fused_kernel<A, H, LEARNED, TPR>;
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
  // deferred fold: one job for the PW-wide partial rows (the layout the in-kernel reduce walks), one for the metrics
  mnr_fold_list_t* fl = mnr_fold_target();
  mnr_fold_job_t* fj = mnr_fold_next();
  a.defer = (fl && fj && !fl->has_metrics) ? 1 : 0;
  if (a.defer) {
    const int Ar = a.A, PW = (2 * A + 1) * H + 4 * A + 1, c0 = (2 * A + 1) * H;
    const float inv_m = 1.f / (float)a.M;
    fj->parts = a.partials, fj->n_parts = grid, fj->width = PW, fj->n_segs = 0;
    auto seg = [&](int begin, int n, float* dst, float scale) {
      if (dst && n > 0) fj->seg[fj->n_segs++] = mnr_fold_seg_t{begin, begin + n, dst, scale, 0};
    };
    seg(0, Ar * H, a.dw_mu, 1.f);
    if (LEARNED) seg(A * H, Ar * H, a.dw_sig, 1.f);
    seg(2 * A * H, H, a.dw_v, 1.f);
    seg(c0, Ar, a.db_mu, 1.f);
    seg(c0 + A, Ar, a.db_sig, 1.f);
    seg(c0 + 2 * A, Ar, a.mu_mean, inv_m);
    seg(c0 + 3 * A, Ar, a.sigma_mean, inv_m);
    seg(c0 + 4 * A, 1, a.db_v, 1.f);
    fl->n_folds++;
    fl->metrics = mnr_heads_metrics_job_t{reinterpret_cast<const double*>(a.loss_parts), a.metrics, a.M, grid, 0,
                                          cfg.critic_coef, cfg.entropy_coef, cfg.bounds_loss_coef, cfg.metric_scale};
    fl->has_metrics = 1;
  }
  cudaError_t e = mnr_launch_pdl(kern, dim3(grid), dim3(kThreadsHF), smem, stream, a, cfg);
  if (e == cudaSuccess) e = cudaPeekAtLastError();
  return e == cudaSuccess ? MNR_OK : (int)e;
}

template <int A, int H, bool LEARNED>
int launch_fused(FusedHeadsArgs& a, const mnr_ppo_loss_cfg_t& cfg, int grid_cap, cudaStream_t stream) {
  // wide heads (AllegroHand: 16 actions, learned sigma, 128 hidden): a 128-row tile with its gradient rows needs > 113 KB of
  // shared memory = ONE CTA of 8 warps per SM and two tiles back to back per CTA; 64-row tiles (4 threads per row) keep two
  // CTAs per SM, prefetch the value tower's tile, and halve the per-thread chains
  constexpr int TPR_BIG = fused_smem_bytes<A, H, LEARNED, 2>() > 113 * 1024 ? 4 : 2;
  const int rows = kThreadsHF / (small_tiles_for(a.M) ? 8 : TPR_BIG);
  a.n_tiles = (int)((a.M + rows - 1) / rows);
  return rows == kThreadsHF / 8 ? launch_fused_t<A, H, LEARNED, 8>(a, cfg, grid_cap, stream)
                                : launch_fused_t<A, H, LEARNED, TPR_BIG>(a, cfg, grid_cap, stream);
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

// debug: enable / read the CTA-0 timeline (globaltimer ns at 16 phase boundaries)
extern "C" int mnr_heads_trace(int enable, long long* host_out) {
  cudaError_t e = cudaMemcpyToSymbol(g_htrace_on, &enable, sizeof(int));
  if (e != cudaSuccess) return (int)e;
  if (host_out) {
    e = cudaMemcpyFromSymbol(host_out, g_htrace, sizeof(long long) * 32);
    if (e != cudaSuccess) return (int)e;
  }
  return MNR_OK;
}

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
  a.defer = 0;
  a.M = M, a.A = A, a.write_back = write_back, a.n_tiles = 0;   // set with the tile height (launch_fused)
  if (A % 4 == 0 && (!mnr_aligned(actions, 16) || !mnr_aligned(old_mu, 16) || !mnr_aligned(old_sigma, 16) ||
                     (g_mu && !mnr_aligned(g_mu, 16)) || (g_ls && !mnr_aligned(g_ls, 16))))
    return MNR_EALIGN;
  return w_sig != nullptr ? launch_by_a<true>(a, H, *cfg, cap, stream) : launch_by_a<false>(a, H, *cfg, cap, stream);
}
