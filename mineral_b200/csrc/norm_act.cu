// a10  LayerNorm + activation between the tower GEMMs, forward and backward, for the reference's shipped PPO
// recipes (mineral/cfgs/agent/DFlexAntPPO.yaml:17-21, AllegroHandPPO.yaml:18-28: Linear -> LayerNorm -> SiLU).
// Replaces nn.LayerNorm + nn.SiLU/ELU/ReLU of mineral/nets/mlp.py:66-71 and autograd's backward through them.
//
//   fwd:  mean, rstd over the N features of a row;  xh = (z - mean) * rstd;  n = gamma * xh + beta;  y = act(n)
//   bwd:  dn = dy * act'(n);  g = dn * gamma;  dz = rstd * (g - mean_N(g) - xh * mean_N(g * xh))
//         dgamma = sum_rows dn * xh,  dbeta = sum_rows dn
//
// HBM-bound element-wise passes (8 B/element forward, 12 B/element backward).  One WARP per row: the row lives in
// registers (float4 pieces strided by the warp so every load instruction covers 512 contiguous bytes), statistics by
// warp shuffles, two-pass variance.  The backward keeps dgamma / dbeta in registers across all rows a warp owns (512-wide
// rows with several rows per warp: in the warp's own shared-memory slab), prefetches the next row while it works on the
// current one, folds the warps of a CTA through shared memory and parks one partial per CTA; then EITHER it ends there and
// the sums over the CTAs run inside the split-reduce launch (deferred folds, mnr_fold_defer: the PPO update's path) OR the
// grid (persistent, all CTAs resident) meets at a grid barrier and every CTA reduces its share of the 2N outputs in a
// fixed order.  Deterministic either way, no atomics.
#include "common.cuh"

namespace {

constexpr int kThreads = 256, kWarps = kThreads / 32;
constexpr int kMaxGroups = 2;

struct LnArgs {
  mnr_ln_group_t g[kMaxGroups];
  int64_t M;
  int N, n_groups;
  float eps;
  float* partials;                // bwd: [groups][grid][2N]
  unsigned long long* barrier;    // bwd
  int defer;                      // bwd: deferred fold (mnr_fold_defer): stop after parking the per-CTA partials
};

template <int ACT>
__device__ __forceinline__ float act_f(float n) {
  if (ACT == 1) return n > 0.f ? n : expm1f(n);
  if (ACT == 2) return fmaxf(n, 0.f);
  if (ACT == 3) return n / (1.f + expf(-n));
  return n;
}
template <int ACT>
__device__ __forceinline__ float act_df(float n) {
  if (ACT == 1) return n > 0.f ? 1.f : expf(n);
  if (ACT == 2) return n > 0.f ? 1.f : 0.f;
  if (ACT == 3) {
    const float s = 1.f / (1.f + expf(-n));
    return s * (1.f + n * (1.f - s));
  }
  return 1.f;
}

__device__ __forceinline__ float warp_allsum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// E = elements per lane (N <= 32 * E).  VEC: N % 128 == 0, lane owns float4 pieces (lane + 32 k); else scalar pieces.
template <int E, bool VEC>
struct Row {
  float v[E];
  __device__ __forceinline__ void load(const float* p, int N, int lane) {
    if (VEC) {
#pragma unroll
      for (int k = 0; k < E / 4; ++k) {
        const int c = 4 * (lane + 32 * k);
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < N) t = __ldg(reinterpret_cast<const float4*>(p + c));
        v[4 * k] = t.x, v[4 * k + 1] = t.y, v[4 * k + 2] = t.z, v[4 * k + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < E; ++k) {
        const int c = lane + 32 * k;
        v[k] = c < N ? __ldg(p + c) : 0.f;
      }
    }
  }
  __device__ __forceinline__ void store(float* p, int N, int lane) const {
    if (VEC) {
#pragma unroll
      for (int k = 0; k < E / 4; ++k) {
        const int c = 4 * (lane + 32 * k);
        if (c < N) *reinterpret_cast<float4*>(p + c) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < E; ++k) {
        const int c = lane + 32 * k;
        if (c < N) p[c] = v[k];
      }
    }
  }
  // slots [4k, 4k + 4) (VEC) or slot k (scalar) of a row that lives in SHARED memory
  __device__ __forceinline__ static void piece(const float* p, int k, int lane, float (&out)[4]) {
    if (VEC) {
      const float4 t = *reinterpret_cast<const float4*>(p + 4 * (lane + 32 * k));
      out[0] = t.x, out[1] = t.y, out[2] = t.z, out[3] = t.w;
    } else {
      out[0] = p[lane + 32 * k];
    }
  }
  __device__ __forceinline__ static void piece_st(float* p, int k, int lane, const float (&v)[4]) {
    if (VEC) *reinterpret_cast<float4*>(p + 4 * (lane + 32 * k)) = make_float4(v[0], v[1], v[2], v[3]);
    else p[lane + 32 * k] = v[0];
  }
  // column index of register slot i
  __device__ __forceinline__ static int col(int i, int lane) { return VEC ? 4 * (lane + 32 * (i / 4)) + (i & 3) : lane + 32 * i; }
};

template <int E, bool VEC, int ACT>
__global__ void __launch_bounds__(kThreads) ln_act_fwd_kernel(const __grid_constant__ LnArgs a) {
  mnr::pdl_trigger();   // PDL: see common.cuh
  mnr::pdl_wait();
  const mnr_ln_group_t& g = a.g[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = a.N;
  const float invN = 1.f / (float)N;
  Row<E, VEC> gam, bet;
  gam.load(g.gamma, N, lane), bet.load(g.beta, N, lane);
  // (a software prefetch of the next row, as in the backward kernel, measured SLOWER here: 59 -> 76 us at 32768 x 512 - the
  // extra registers cost a resident CTA - and unchanged for the narrower layers)
  for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < a.M; row += (int64_t)gridDim.x * kWarps) {
    Row<E, VEC> x;
    x.load(g.z + row * N, N, lane);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) s += x.v[i];
    const float mean = warp_allsum(s) * invN;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const float d = Row<E, VEC>::col(i, lane) < N ? x.v[i] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_allsum(q) * invN + a.eps);
#pragma unroll
    for (int i = 0; i < E; ++i) x.v[i] = act_f<ACT>(fmaf((x.v[i] - mean) * rstd, gam.v[i], bet.v[i]));
    x.store(g.y + row * N, N, lane);
    if (lane == 0) g.mean[row] = mean, g.rstd[row] = rstd;
  }
}

template <int E, bool VEC, int ACT, bool PFW>
__global__ void __launch_bounds__(kThreads, 2) ln_act_bwd_kernel(const __grid_constant__ LnArgs a) {
  extern __shared__ __align__(16) float s_fold[];  // [kWarps][2N]
  mnr::pdl_trigger();   // PDL: see common.cuh
  mnr::pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = a.N;
  const float invN = 1.f / (float)N;
  const int n_groups = a.n_groups;
  // gamma / beta of the group live in shared memory (conflict-free LDS.128 per piece), which frees their 2E registers for a
  // software prefetch: the NEXT row's z / dy loads are issued before this row's arithmetic, so a warp keeps loads in flight
  // through its compute phase (exp, two butterfly sums) instead of alternating between the two (32768 x 256 x 2 towers:
  // 64.4 -> 53.2 us)
  constexpr int PK = VEC ? E / 4 : E, PW_ = VEC ? 4 : 1;   // pieces per lane, slots per piece
  // 16 elements per lane (N = 512): with dgamma / dbeta in registers the prefetch spills (measured slower, also when only z is
  // prefetched), so those rows accumulate dgamma / dbeta in the warp's OWN slab of the fold buffer (a lane re-reads only what it
  // wrote: no synchronisation; 16 LDS.128 + 16 STS.128 per row against 6 KB of DRAM traffic) whenever a warp has rows to overlap
  // PFW (wide rows, chosen by the host): only when a warp owns more than one row
  constexpr bool ACC_SMEM = E > 8 && PFW;
  constexpr bool PF = E <= 8 || PFW;
  const int NP = VEC ? N : 32 * E;                          // padded row length in shared memory (scalar form: zero pad)
  float* s_gam = s_fold + kWarps * 2 * N;
  float* s_bet = s_gam + NP;
  for (int gi = 0; gi < n_groups; ++gi) {   // every CTA serves every group: one grid barrier for all of them
    const mnr_ln_group_t& g = a.g[gi];
    Row<E, VEC> dgam, dbet, gam, bet;   // gam / bet: registers without the prefetch (as before), shared memory with it
#pragma unroll
    for (int i = 0; i < E; ++i) dgam.v[i] = 0.f, dbet.v[i] = 0.f;
    float* slab = s_fold + warp * 2 * N;   // this warp's [dgamma | dbeta] slab
    constexpr bool acc_smem = ACC_SMEM;
    const int64_t stride = (int64_t)gridDim.x * kWarps;
    int64_t row = (int64_t)blockIdx.x * kWarps + warp;
    Row<E, VEC> z, dy;
    float mean = 0.f, rstd = 0.f;
    if (PF) {
      if (row < a.M) {   // the first row's loads go out before gamma / beta are staged: one DRAM round trip for both
        z.load(g.z + row * N, N, lane), dy.load(g.dy + row * N, N, lane);
        mean = __ldg(g.mean + row), rstd = __ldg(g.rstd + row);
      }
      __syncthreads();   // the previous group's rows (s_gam / s_bet) and fold (s_fold) are done
      for (int c = threadIdx.x; c < NP; c += kThreads) s_gam[c] = c < N ? __ldg(g.gamma + c) : 0.f, s_bet[c] = c < N ? __ldg(g.beta + c) : 0.f;
      if (acc_smem) dgam.store(slab, N, lane), dbet.store(slab + N, N, lane);   // zeros
      __syncthreads();
    } else {
      gam.load(g.gamma, N, lane), bet.load(g.beta, N, lane);
    }
    while (row < a.M) {
      const int64_t nxt = row + stride;
      Row<E, VEC> zn, dyn;
      float mean_n = 0.f, rstd_n = 0.f;
      if (!PF) {
        z.load(g.z + row * N, N, lane), dy.load(g.dy + row * N, N, lane);
        mean = __ldg(g.mean + row), rstd = __ldg(g.rstd + row);
      } else if (nxt < a.M) {
        zn.load(g.z + nxt * N, N, lane), dyn.load(g.dy + nxt * N, N, lane);
        mean_n = __ldg(g.mean + nxt), rstd_n = __ldg(g.rstd + nxt);
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < PK; ++k) {
        float gk[4], bk[4];
        if (PF) {
          Row<E, VEC>::piece(s_gam, k, lane, gk), Row<E, VEC>::piece(s_bet, k, lane, bk);
        } else {
#pragma unroll
          for (int e = 0; e < PW_; ++e) gk[e] = gam.v[PW_ * k + e], bk[e] = bet.v[PW_ * k + e];
        }
#pragma unroll
        for (int e = 0; e < PW_; ++e) {
          const int i = PW_ * k + e;
          const bool in = Row<E, VEC>::col(i, lane) < N;
          const float xh = in ? (z.v[i] - mean) * rstd : 0.f;
          const float dn = in ? dy.v[i] * act_df<ACT>(fmaf(xh, gk[e], bk[e])) : 0.f;
          const float gg = dn * gk[e];
          if (acc_smem) gk[e] = dn * xh, bk[e] = dn;   // gamma / beta are consumed: the slots carry this piece's dgamma / dbeta terms
          else dgam.v[i] = fmaf(dn, xh, dgam.v[i]), dbet.v[i] += dn;
          s1 += gg, s2 = fmaf(gg, xh, s2);
          z.v[i] = xh, dy.v[i] = gg;
        }
        if (acc_smem) {
          float dg[4], db[4];
          Row<E, VEC>::piece(slab, k, lane, dg), Row<E, VEC>::piece(slab + N, k, lane, db);
#pragma unroll
          for (int e = 0; e < PW_; ++e) dg[e] += gk[e], db[e] += bk[e];
          Row<E, VEC>::piece_st(slab, k, lane, dg), Row<E, VEC>::piece_st(slab + N, k, lane, db);
        }
      }
      s1 = warp_allsum(s1) * invN, s2 = warp_allsum(s2) * invN;
#pragma unroll
      for (int i = 0; i < E; ++i) dy.v[i] = rstd * (dy.v[i] - s1 - z.v[i] * s2);
      dy.store(g.dz + row * N, N, lane);
      if (PF) z = zn, dy = dyn, mean = mean_n, rstd = rstd_n;
      row = nxt;
    }
    // fold the CTA's warps, park the CTA partial
    __syncthreads();
    if (!acc_smem) dgam.store(slab, N, lane), dbet.store(slab + N, N, lane);
    __syncthreads();
    float* part = a.partials + ((size_t)gi * gridDim.x + blockIdx.x) * 2 * N;
    for (int c = threadIdx.x; c < 2 * N; c += kThreads) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += s_fold[w * 2 * N + c];
      part[c] = t;
    }
  }
  if (a.defer) return;   // dgamma / dbeta are summed over the CTAs inside the split-reduce launch (tc_linear.cu, fold blocks)
  mnr::grid_barrier(a.barrier);
  const int n_cta = gridDim.x;
  for (int o = blockIdx.x * kWarps + warp; o < n_groups * 2 * N; o += n_cta * kWarps) {
    const int gi = o / (2 * N), c = o - gi * 2 * N;
    const float* src = a.partials + (size_t)gi * n_cta * 2 * N + c;
    float t = 0.f;
    for (int x = lane; x < n_cta; x += 32) t += __ldcg(src + (size_t)x * 2 * N);
    t = warp_allsum(t);
    if (lane == 0) {
      if (c < N) a.g[gi].dgamma[c] = t;
      else a.g[gi].dbeta[c - N] = t;
    }
  }
}

int fill(LnArgs& a, const mnr_ln_group_t* groups, int n_groups, int64_t M, int N, bool bwd) {
  if (!groups) return MNR_ENULL;
  if (n_groups <= 0 || n_groups > kMaxGroups || M <= 0 || N <= 0 || N > 512) return MNR_ESHAPE;
  for (int k = 0; k < n_groups; ++k) {
    const mnr_ln_group_t& g = groups[k];
    if (!g.z || !g.gamma || !g.beta || !g.mean || !g.rstd) return MNR_ENULL;
    if (!bwd && !g.y) return MNR_ENULL;
    if (bwd && (!g.dy || !g.dz || !g.dgamma || !g.dbeta)) return MNR_ENULL;
    if (N % 128 == 0 && (!mnr_aligned(g.z, 16) || !mnr_aligned(g.gamma, 16) || !mnr_aligned(g.beta, 16) ||
                         (!bwd && !mnr_aligned(g.y, 16)) || (bwd && (!mnr_aligned(g.dy, 16) || !mnr_aligned(g.dz, 16)))))
      return MNR_EALIGN;
    a.g[k] = g;
  }
  for (int k = n_groups; k < kMaxGroups; ++k) a.g[k] = groups[0];
  a.M = M, a.N = N, a.n_groups = n_groups, a.partials = nullptr, a.barrier = nullptr, a.defer = 0;
  return MNR_OK;
}

// dispatch on (elements per lane, vectorised, activation)
#define MNR_LN_DISPATCH(KERNEL, ...)                                                              \
  do {                                                                                            \
    const bool vec = (N % 128 == 0);                                                              \
    const int epl = vec ? N / 32 : (N + 31) / 32;                                                 \
    if (vec && epl <= 4) { MNR_LN_ACT(KERNEL, 4, true, __VA_ARGS__) }                             \
    else if (vec && epl <= 8) { MNR_LN_ACT(KERNEL, 8, true, __VA_ARGS__) }                        \
    else if (vec) { MNR_LN_ACT(KERNEL, 16, true, __VA_ARGS__) }                                   \
    else if (epl <= 4) { MNR_LN_ACT(KERNEL, 4, false, __VA_ARGS__) }                              \
    else { MNR_LN_ACT(KERNEL, 16, false, __VA_ARGS__) }                                           \
  } while (0)
#define MNR_LN_ACT(KERNEL, EV, VV, ...)                                                           \
  switch (act) {                                                                                  \
    case 0: MNR_LN_GO(KERNEL, EV, VV, 0, __VA_ARGS__) break;                                      \
    case 1: MNR_LN_GO(KERNEL, EV, VV, 1, __VA_ARGS__) break;                                      \
    case 2: MNR_LN_GO(KERNEL, EV, VV, 2, __VA_ARGS__) break;                                      \
    default: MNR_LN_GO(KERNEL, EV, VV, 3, __VA_ARGS__) break;                                     \
  }

int bwd_grid(int64_t M) {
  int64_t want = (M + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)mnr_sm_count() * 2;   // 2 CTAs of 256 threads per SM are always resident (<= 128 regs, <= 32 KB smem)
  if (want > cap) want = cap;
  return (int)(want < 1 ? 1 : want);
}

}  // namespace

extern "C" int mnr_ln_act_fwd(const mnr_ln_group_t* groups, int n_groups, int64_t M, int N, float eps, int act, mnr_stream_t stream) {
  LnArgs a;
  const int rc = fill(a, groups, n_groups, M, N, false);
  if (rc != MNR_OK) return rc;
  if (act < 0 || act > 3) return MNR_EUNSUP;
  a.eps = eps;
  int64_t want = (M + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)mnr_sm_count() * 8;
  dim3 grid((unsigned)(want > cap ? cap : want), n_groups);
#define MNR_LN_GO(KERNEL, EV, VV, AV, ...) mnr_launch_pdl(KERNEL<EV, VV, AV>, grid, dim3(kThreads), 0, mnr_cs(stream), a);
  MNR_LN_DISPATCH(ln_act_fwd_kernel, 0);
#undef MNR_LN_GO
  MNR_LAUNCH_RESULT();
}

extern "C" size_t mnr_ln_act_bwd_workspace_bytes(int n_groups, int64_t M, int N) {
  if (n_groups <= 0 || M <= 0 || N <= 0) return 0;
  return 256 + (size_t)n_groups * bwd_grid(M) * 2 * N * sizeof(float);
}

extern "C" int mnr_ln_act_bwd(const mnr_ln_group_t* groups, int n_groups, int64_t M, int N, int act, void* workspace,
                              size_t workspace_bytes, mnr_stream_t stream) {
  LnArgs a;
  const int rc = fill(a, groups, n_groups, M, N, true);
  if (rc != MNR_OK) return rc;
  if (act < 0 || act > 3) return MNR_EUNSUP;
  MNR_CHECK_PTR(workspace);
  if (workspace_bytes < mnr_ln_act_bwd_workspace_bytes(n_groups, M, N)) return MNR_EWORKSPACE;
  if (!mnr_aligned(workspace, 16)) return MNR_EALIGN;
  a.eps = 0.f;
  a.barrier = reinterpret_cast<unsigned long long*>(workspace);
  a.partials = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + 256);
  const int grid = bwd_grid(M);  // every CTA loops over the groups itself: one grid barrier for all of them
  const size_t smem = ((size_t)kWarps * 2 * N + 2 * (size_t)(N % 128 == 0 ? N : 32 * ((N + 31) / 32 <= 4 ? 4 : 16))) * sizeof(float);
  if (mnr_fold_list_t* fl = mnr_fold_target()) {   // deferred fold: one job per group, [grid][2N] partial rows -> dgamma | dbeta
    if (fl->n_folds + n_groups <= MNR_MAX_FOLDS) {
      a.defer = 1;
      for (int gi = 0; gi < n_groups; ++gi) {
        mnr_fold_job_t& fj = fl->folds[fl->n_folds++];
        fj.parts = a.partials + (size_t)gi * grid * 2 * N, fj.n_parts = grid, fj.width = 2 * N, fj.n_segs = 2;
        fj.seg[0] = mnr_fold_seg_t{0, N, groups[gi].dgamma, 1.f, 0};
        fj.seg[1] = mnr_fold_seg_t{N, 2 * N, groups[gi].dbeta, 1.f, 0};
      }
    }
  }
  const bool pfw = M > (int64_t)grid * kWarps;   // wide rows: prefetch + shared-memory dgamma / dbeta when a warp owns several rows
#define MNR_LN_GO(KERNEL, EV, VV, AV, ...)                                                                     \
  {                                                                                                            \
    auto kern = (EV > 8 && pfw) ? KERNEL<EV, VV, AV, (EV > 8)> : KERNEL<EV, VV, AV, false>;                    \
    cudaError_t e = mnr_allow_max_dyn_smem(kern);                                                              \
    if (e != cudaSuccess) return (int)e;                                                                       \
    mnr_launch_pdl(kern, dim3(grid), dim3(kThreads), smem, mnr_cs(stream), a);                                 \
  }
  MNR_LN_DISPATCH(ln_act_bwd_kernel, 0);
#undef MNR_LN_GO
  MNR_LAUNCH_RESULT();
}
