// K1a + K3a fused post-rollout pass:  GAE scan + (T,N,..)->(N*T,..) env-major pack of every
// rollout field + Welford/Chan partial moments of advantages / values / returns, in ONE pass
// over the rollout (each input byte read once from HBM, each output byte written once).
//
// Replaces ExperienceBuffer.compute_return + transform_op + prepare_training
// (mineral/agents/ppo/experience.py:7-14,96-121) and feeds PPO's value-RMS block
// (mineral/agents/ppo/ppo.py:175-183).
//
// Mapping.  A CTA owns a tile of `nb` consecutive envs over the whole horizon.  For a wide field
// (obs, actions, mu, sigma; `width` floats per (t,n)) the tile's input is T contiguous runs of
// nb*width floats (one per step) and its output is ONE contiguous run of nb*T*width floats.
//   load : T x n_wide 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) into shared
//          memory, issued up front by one warp -> the whole tile is in flight at once (bytes in
//          flight = shared-memory footprint, what an HBM-bound kernel needs; Little's law);
//   scan : while the copies land, one warp runs the GAE recurrence for the tile's envs out of
//          staged scalars and accumulates the moment partials;
//   store: all warps walk the output run linearly (float4, fully coalesced) and pick the
//          matching element out of shared memory (the transpose happens in the smem indexing).
// Tensor-map TMA cannot be used: the (t,n) pitch is width*4 B (148 B for Ant) and tensor maps
// need 16 B-multiple strides; 1-D bulk copies only need 16 B alignment of each run, which holds
// when N % 4 == 0 and nb % 4 == 0.  Otherwise a plain coalesced-load variant is launched.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 128;      // post_rollout_kernel CTA
constexpr int kFinThreads = 256;   // finalize kernels
constexpr int kHeaderBytes = 128;  // [0] n_tiles, [1] ticket ; +64: FinalStats

struct FinalStats {
  float adv_mean, adv_denom, v_mean, v_denom, r_mean, r_denom;
  int normalize_value;
  int _pad;
};

struct TilePartial {
  mnr::Moments adv, val, ret;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One launch, two CTA roles (blockIdx.x < n_scalar: scalar role, else wide role).
//
//  scalar role: 32 consecutive envs, one lane per env.  Every step t the scanning warp reads 32
//    consecutive envs of rewards / values / dones / neglogp (one 128 B line each, all T steps'
//    loads are independent of the recurrence and get batched), runs the GAE recurrence in
//    registers, accumulates Welford moments of advantages / values / returns, and parks the
//    per-env series in shared memory [env][T+1] (odd pitch: conflict-free).  All warps then write
//    the six env-major scalar arrays as linear, fully coalesced runs of 32*T elements.
//  wide role: one (field, env-tile) work item: T bulk copies of nb*width floats into shared memory
//    [t][nb*width], then a linear walk of the item's contiguous env-major output run; lane i reads
//    element i (stride-1 shared-memory reads: no bank conflicts) and stores 4 B (a warp writes one
//    full 128 B line per instruction).
struct Params {
  mnr_wide_field_t f[MNR_MAX_WIDE_FIELDS];
  int nb[MNR_MAX_WIDE_FIELDS];          // envs per tile of field k (multiple of 4)
  int item0[MNR_MAX_WIDE_FIELDS + 1];   // prefix sum of item counts
  int n_wide;
  int n_scalar;                         // scalar-role CTAs = ceil(N / 32)
  const float *rewards, *values, *neglogp, *last_values;
  const uint8_t* dones;
  float *returns_tm, *rewards_em, *values_em, *returns_em, *neglogp_em, *adv_em;
  uint8_t* dones_em;
  int T, N, do_gae;
  float gamma, gamma_tau;
  unsigned char* workspace;
  int debug_skip;  // tuning aid (MNR_POST_ROLLOUT_SKIP): bit 0 skips the scalar role, bit 1 the wide role
};

constexpr int kEnvPerScalarCta = 32;  // the staging loops below use q >> 5 / q & 31

// exact floor(q / d) for 0 <= q < 2^22, 1 <= d < 2^22 with inv = 1.0f / d
__device__ __forceinline__ int fast_div(int q, float inv, int d) {
  int r = __float2int_rz(__fmul_rn((float)q + 0.5f, inv));
  r -= (r * d > q);
  r += ((r + 1) * d <= q);
  return r;
}

__device__ __forceinline__ void scalar_role(const Params& p, float* smem) {
  const int T = p.T, N = p.N;
  const int n0 = blockIdx.x * kEnvPerScalarCta;
  const int nb = min(kEnvPerScalarCta, N - n0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pitch = T + 1;
  float* s_r = smem;
  float* s_v = s_r + kEnvPerScalarCta * pitch;
  float* s_ret = s_v + kEnvPerScalarCta * pitch;
  float* s_adv = s_ret + kEnvPerScalarCta * pitch;
  float* s_nlp = s_adv + kEnvPerScalarCta * pitch;
  float* s_done = s_nlp + kEnvPerScalarCta * pitch;

  // phase A (all threads): stage the tile's scalars, [t][n0 .. n0+nb) is one coalesced line per array and step;
  // every load is independent, so they are all in flight together (one DRAM latency, not T of them)
  for (int q = tid; q < T * kEnvPerScalarCta; q += kThreads) {
    const int t = q >> 5, nl = q & 31;
    if (nl < nb) {
      const size_t g = (size_t)t * N + n0 + nl;
      const int s = nl * pitch + t;
      s_r[s] = __ldg(p.rewards + g), s_v[s] = __ldg(p.values + g), s_nlp[s] = __ldg(p.neglogp + g);
      s_done[s] = (float)__ldg(p.dones + g);
      if (!p.do_gae) s_ret[s] = __ldg(p.returns_tm + g);
    }
  }
  __syncthreads();
  // phase B (one warp, one lane per env): the serial GAE recurrence out of shared memory
  if (warp == 0 && lane < nb) {
    float next_v = p.do_gae ? __ldg(p.last_values + n0 + lane) : 0.f;
    float carry = 0.f;
#pragma unroll 4
    for (int t = T - 1; t >= 0; --t) {
      const int s = lane * pitch + t;
      const float v = s_v[s];
      float ret;
      if (p.do_gae) {  // experience.py:100-108, same association order as scans.cu
        const float alive = __fsub_rn(1.0f, s_done[s]);
        const float delta = __fsub_rn(__fadd_rn(s_r[s], __fmul_rn(__fmul_rn(p.gamma, next_v), alive)), v);
        carry = __fadd_rn(delta, __fmul_rn(__fmul_rn(p.gamma_tau, alive), carry));
        ret = __fadd_rn(carry, v);
        next_v = v;
        s_ret[s] = ret;
      } else {
        ret = s_ret[s];
      }
      s_adv[s] = __fsub_rn(ret, v);  // experience.py:119 (returns - values)
    }
  }
  __syncthreads();
  // time-major returns (storage_dict['returns']), coalesced along the envs
  if (p.do_gae && p.returns_tm) {
    for (int q = tid; q < T * kEnvPerScalarCta; q += kThreads) {
      const int t = q >> 5, nl = q & 31;
      if (nl < nb) p.returns_tm[(size_t)t * N + n0 + nl] = s_ret[nl * pitch + t];
    }
  }
  const size_t base = (size_t)n0 * T;
  const float invT = 1.0f / (float)T;
  // parallel part: coalesced env-major write-out by all warps, moments accumulated on the way
  mnr::Welford wa, wv, wr;
  wa.init(), wv.init(), wr.init();
  for (int q = tid; q < nb * T; q += kThreads) {
    const int nl = fast_div(q, invT, T), t = q - nl * T;
    const int s = nl * pitch + t;
    const float v = s_v[s], ret = s_ret[s], adv = s_adv[s];
    p.rewards_em[base + q] = s_r[s];
    p.values_em[base + q] = v;
    p.returns_em[base + q] = ret;
    p.neglogp_em[base + q] = s_nlp[s];
    p.adv_em[base + q] = adv;
    p.dones_em[base + q] = (uint8_t)s_done[s];
    wa.push(adv), wv.push(v), wr.push(ret);
  }
  const mnr::Moments m_adv = mnr::block_merge(wa.to_moments());
  const mnr::Moments m_val = mnr::block_merge(wv.to_moments());
  const mnr::Moments m_ret = mnr::block_merge(wr.to_moments());
  if (tid == 0) {
    TilePartial* parts = reinterpret_cast<TilePartial*>(p.workspace + kHeaderBytes);
    parts[blockIdx.x] = TilePartial{m_adv, m_val, m_ret};
    if (blockIdx.x == 0) reinterpret_cast<unsigned int*>(p.workspace)[0] = (unsigned int)p.n_scalar;
  }
}

template <bool kBulk>
__device__ __forceinline__ void wide_role(const Params& p, float* smem, uint64_t* bar) {
  const int T = p.T, N = p.N;
  const int item = blockIdx.x - p.n_scalar;
  int k = 0;
#pragma unroll
  for (int j = 1; j < MNR_MAX_WIDE_FIELDS; ++j) k += (j < p.n_wide && item >= p.item0[j]) ? 1 : 0;
  const int NB = p.nb[k];
  const int w = p.f[k].width;
  const int n0 = (item - p.item0[k]) * NB;
  const int nb = min(NB, N - n0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int run = nb * w;  // floats per step in this tile
  const float* src = p.f[k].src;

  if (kBulk) {
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)(T * run) * 4u);
      __syncwarp();
      for (int t = lane; t < T; t += 32) bulk_g2s(smem + (size_t)t * run, src + ((size_t)t * N + n0) * w, (uint32_t)run * 4u, bar);
    }
    mbar_wait(bar, 0);
  } else {
    const float inv_run = 1.0f / (float)run;
    for (int q = tid; q < T * run; q += kThreads) {
      const int t = fast_div(q, inv_run, run), j = q - t * run;
      smem[q] = mnr::ld_stream(src + ((size_t)t * N + n0) * w + j);
    }
    __syncthreads();
  }

  // smem: [t][nl][f] (pitch `run` per step);  out: [nl][t][f].  A thread owns a fixed (nl, f) pair and walks the T
  // steps: per element one LDS, one STG and two pointer bumps, no index divisions.  At a given t consecutive lanes
  // hold consecutive (nl, f) pairs: shared-memory reads are stride-1, global writes are w-float contiguous pieces.
  float* out = p.f[k].dst + (size_t)n0 * T * w;
  const int pairs = nb * w;
  const float inv_w = 1.0f / (float)w;
  for (int pidx = tid; pidx < pairs; pidx += kThreads) {
    const int nl = fast_div(pidx, inv_w, w);
    const float* sp = smem + pidx;
    float* dp = out + (size_t)nl * (T - 1) * w + pidx;  // = out + (nl*T)*w + f
#pragma unroll 8
    for (int t = 0; t < T; ++t) mnr::st_stream(dp + (size_t)t * w, sp[(size_t)t * run]);
  }
}

template <bool kBulk>
__global__ void __launch_bounds__(kThreads, 6) post_rollout_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  float* smem = reinterpret_cast<float*>(smem_raw);
  if ((int)blockIdx.x < p.n_scalar) {
    if (!(p.debug_skip & 1)) scalar_role(p, smem);
  } else if (!(p.debug_skip & 2)) {
    wide_role<kBulk>(p, smem, &bar);
  }
}

// ---- finalize: merge partials (fixed order), derive the 6 normalisation scalars, update value RMS ---
template <typename S>
__device__ void chan_update_state(S* mean, S* var, S* count, float bmean, float bvar, double bcount) {
  S m = *mean, v = *var;
  const S c = *count;
  mnr::chan_update<S>(m, v, c, bmean, bvar, bcount);
  *mean = m, *var = v, *count = mnr::add_rn(c, (S)bcount);
}

__global__ void __launch_bounds__(kFinThreads) finalize_stats_kernel(unsigned char* __restrict__ workspace, void* value_mean,
                                                                  void* value_var, void* value_count, int state_f64,
                                                                  float eps, double* __restrict__ stats_out) {
  const unsigned int n_tiles = reinterpret_cast<const unsigned int*>(workspace)[0];
  const TilePartial* parts = reinterpret_cast<const TilePartial*>(workspace + kHeaderBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  mnr::Moments a{0, 0, 0}, v{0, 0, 0}, r{0, 0, 0};
  for (unsigned int i = tid; i < n_tiles; i += kFinThreads) {
    a = mnr::moments_merge(a, parts[i].adv);
    v = mnr::moments_merge(v, parts[i].val);
    r = mnr::moments_merge(r, parts[i].ret);
  }
  a = mnr::moments_warp_reduce(a), v = mnr::moments_warp_reduce(v), r = mnr::moments_warp_reduce(r);
  __shared__ mnr::Moments sm[3][kFinThreads / 32];
  if (lane == 0) sm[0][warp] = a, sm[1][warp] = v, sm[2][warp] = r;
  __syncthreads();
  if (tid != 0) return;
  for (int w = 1; w < kFinThreads / 32; ++w) {
    a = mnr::moments_merge(a, sm[0][w]);
    v = mnr::moments_merge(v, sm[1][w]);
    r = mnr::moments_merge(r, sm[2][w]);
  }
  FinalStats fs;
  // experience.py:119-120: (adv - mean) / (std_unbiased + 1e-8), all f32
  fs.adv_mean = (float)a.mean;
  const float adv_std = (float)sqrt(a.m2 / (a.n - 1.0));
  fs.adv_denom = __fadd_rn(adv_std, 1e-8f);
  fs.normalize_value = value_mean != nullptr;
  fs.v_mean = fs.r_mean = 0.f, fs.v_denom = fs.r_denom = 1.f;
  const float v_bm = (float)v.mean, v_bv = (float)(v.m2 / (v.n - 1.0));  // x.mean(0), x.var(0, correction=1) in f32
  const float r_bm = (float)r.mean, r_bv = (float)(r.m2 / (r.n - 1.0));
  double s1m = 0, s1v = 0, s2m = 0, s2v = 0;
  if (value_mean) {  // ppo.py:177-181
    if (state_f64) {
      double *m = (double*)value_mean, *vv = (double*)value_var, *c = (double*)value_count;
      chan_update_state<double>(m, vv, c, v_bm, v_bv, v.n);
      fs.v_mean = (float)*m, fs.v_denom = __fsqrt_rn(__fadd_rn((float)*vv, eps));
      s1m = *m, s1v = *vv;
      chan_update_state<double>(m, vv, c, r_bm, r_bv, r.n);
      fs.r_mean = (float)*m, fs.r_denom = __fsqrt_rn(__fadd_rn((float)*vv, eps));
      s2m = *m, s2v = *vv;
    } else {
      float *m = (float*)value_mean, *vv = (float*)value_var, *c = (float*)value_count;
      chan_update_state<float>(m, vv, c, v_bm, v_bv, v.n);
      fs.v_mean = *m, fs.v_denom = __fsqrt_rn(__fadd_rn(*vv, eps));
      s1m = *m, s1v = *vv;
      chan_update_state<float>(m, vv, c, r_bm, r_bv, r.n);
      fs.r_mean = *m, fs.r_denom = __fsqrt_rn(__fadd_rn(*vv, eps));
      s2m = *m, s2v = *vv;
    }
  }
  *reinterpret_cast<FinalStats*>(workspace + 64) = fs;
  if (stats_out) {
    stats_out[0] = a.mean, stats_out[1] = adv_std;
    stats_out[2] = v_bm, stats_out[3] = v_bv, stats_out[4] = r_bm, stats_out[5] = r_bv;
    stats_out[6] = s1m, stats_out[7] = s1v, stats_out[8] = s2m, stats_out[9] = s2v;
  }
}

__device__ __forceinline__ float norm1(float x, float mean, float denom, float lo, float hi) {
  float y = __fdiv_rn(__fsub_rn(x, mean), denom);
  if (lo <= hi) y = fminf(fmaxf(y, lo), hi);  // torch.clamp(min, max)
  return y;
}

__global__ void __launch_bounds__(kFinThreads) finalize_apply_kernel(float* __restrict__ adv, float* __restrict__ values,
                                                                  float* __restrict__ returns, int64_t B, float lo,
                                                                  float hi, const unsigned char* __restrict__ workspace) {
  const FinalStats fs = *reinterpret_cast<const FinalStats*>(workspace + 64);
  const int64_t stride = (int64_t)gridDim.x * kFinThreads;
  const int64_t B4 = B >> 2;
  const bool vec = mnr_aligned_dev(adv) && mnr_aligned_dev(values) && mnr_aligned_dev(returns);
  if (vec) {
    for (int64_t i = (int64_t)blockIdx.x * kFinThreads + threadIdx.x; i < B4; i += stride) {
      float4 a = reinterpret_cast<float4*>(adv)[i];
      a.x = __fdiv_rn(__fsub_rn(a.x, fs.adv_mean), fs.adv_denom), a.y = __fdiv_rn(__fsub_rn(a.y, fs.adv_mean), fs.adv_denom);
      a.z = __fdiv_rn(__fsub_rn(a.z, fs.adv_mean), fs.adv_denom), a.w = __fdiv_rn(__fsub_rn(a.w, fs.adv_mean), fs.adv_denom);
      reinterpret_cast<float4*>(adv)[i] = a;
      if (fs.normalize_value) {
        float4 v = reinterpret_cast<float4*>(values)[i], r = reinterpret_cast<float4*>(returns)[i];
        v.x = norm1(v.x, fs.v_mean, fs.v_denom, lo, hi), v.y = norm1(v.y, fs.v_mean, fs.v_denom, lo, hi);
        v.z = norm1(v.z, fs.v_mean, fs.v_denom, lo, hi), v.w = norm1(v.w, fs.v_mean, fs.v_denom, lo, hi);
        r.x = norm1(r.x, fs.r_mean, fs.r_denom, lo, hi), r.y = norm1(r.y, fs.r_mean, fs.r_denom, lo, hi);
        r.z = norm1(r.z, fs.r_mean, fs.r_denom, lo, hi), r.w = norm1(r.w, fs.r_mean, fs.r_denom, lo, hi);
        reinterpret_cast<float4*>(values)[i] = v;
        reinterpret_cast<float4*>(returns)[i] = r;
      }
    }
  }
  for (int64_t i = (vec ? (B4 << 2) : 0) + (int64_t)blockIdx.x * kFinThreads + threadIdx.x; i < B; i += stride) {
    adv[i] = __fdiv_rn(__fsub_rn(adv[i], fs.adv_mean), fs.adv_denom);
    if (fs.normalize_value) {
      values[i] = norm1(values[i], fs.v_mean, fs.v_denom, lo, hi);
      returns[i] = norm1(returns[i], fs.r_mean, fs.r_denom, lo, hi);
    }
  }
}

struct Plan {
  int nb[MNR_MAX_WIDE_FIELDS];
  int item0[MNR_MAX_WIDE_FIELDS + 1];
  int n_scalar;
  size_t smem;
};

// Per-field tile width: bulk copies of >= ~1 KB (nb*width*4 bytes) but tiles of <= ~32 KB so that
// 6-7 CTAs share an SM (loads of some overlap the stores of others); nb % 4 == 0 keeps every copy
// 16 B aligned.  Returns false if even a 4-env tile of some field exceeds shared memory.
bool make_plan(const mnr_wide_field_t* wide, int n_wide, int T, int N, Plan* pl) {
  // measured on B200: 32 KB tiles (7 CTAs/SM) beat 40 KB at >= 16k envs.  The environment knobs are tuning aids
  // (tools/sweep_post_rollout.py), read once per process.
  static const size_t kTileCap = [] { const char* e = getenv("MNR_POST_ROLLOUT_TILE_KB"); return e ? (size_t)atoi(e) * 1024 : (size_t)32 * 1024; }();
  static const size_t kMinCopy = [] { const char* e = getenv("MNR_POST_ROLLOUT_MIN_COPY"); return e ? (size_t)atoi(e) : (size_t)1024; }();
  const size_t kSmemMax = 227 * 1024 - 2048;
  size_t smem = (size_t)6 * kEnvPerScalarCta * (T + 1) * sizeof(float);
  pl->item0[0] = 0;
  for (int k = 0; k < n_wide; ++k) {
    const size_t row = (size_t)wide[k].width * 4;
    int nb = (int)((kMinCopy + row - 1) / row);
    nb = (nb + 3) / 4 * 4;
    if (nb < 4) nb = 4;
    while (nb > 4 && (size_t)T * nb * row > kTileCap) nb -= 4;
    const int n_pad = (N + 3) / 4 * 4;
    if (nb > n_pad) nb = n_pad;
    const size_t bytes = (size_t)T * nb * row;
    if (bytes > kSmemMax) return false;
    if (bytes > smem) smem = bytes;
    pl->nb[k] = nb;
    pl->item0[k + 1] = pl->item0[k] + (N + nb - 1) / nb;
  }
  for (int k = n_wide; k < MNR_MAX_WIDE_FIELDS; ++k) pl->nb[k] = 4, pl->item0[k + 1] = pl->item0[n_wide];
  if (smem > kSmemMax) return false;
  pl->smem = smem;
  pl->n_scalar = (N + kEnvPerScalarCta - 1) / kEnvPerScalarCta;
  return true;
}

}  // namespace

extern "C" size_t mnr_post_rollout_workspace_bytes(int T, int N) {
  (void)T;
  const size_t max_parts = (size_t)(N + kEnvPerScalarCta - 1) / kEnvPerScalarCta + 1;
  return kHeaderBytes + max_parts * sizeof(TilePartial);
}

extern "C" int mnr_post_rollout_fused(const mnr_wide_field_t* wide, int n_wide, const float* rewards,
                                      const float* values, const uint8_t* dones, const float* neglogp,
                                      const float* last_values, float* returns_tm, float* rewards_em, float* values_em,
                                      float* returns_em, float* neglogp_em, uint8_t* dones_em, float* adv_em, int T,
                                      int N, int do_gae, float gamma, float gamma_tau, void* workspace,
                                      size_t workspace_bytes, mnr_stream_t stream) {
  if (n_wide < 0 || n_wide > MNR_MAX_WIDE_FIELDS) return MNR_ESHAPE;
  if (n_wide > 0) MNR_CHECK_PTR(wide);
  MNR_CHECK_PTR(rewards); MNR_CHECK_PTR(values); MNR_CHECK_PTR(dones); MNR_CHECK_PTR(neglogp);
  MNR_CHECK_PTR(rewards_em); MNR_CHECK_PTR(values_em); MNR_CHECK_PTR(returns_em); MNR_CHECK_PTR(neglogp_em);
  MNR_CHECK_PTR(dones_em); MNR_CHECK_PTR(adv_em); MNR_CHECK_PTR(workspace);
  if (do_gae) MNR_CHECK_PTR(last_values); else MNR_CHECK_PTR(returns_tm);
  if (T <= 0 || N <= 0 || (int64_t)T * N >= (1 << 30)) return MNR_ESHAPE;
  if (workspace_bytes < mnr_post_rollout_workspace_bytes(T, N)) return MNR_EWORKSPACE;
  if (!mnr_aligned(workspace, 16)) return MNR_EALIGN;

  Params p;
  bool bulk_ok = (N % 4 == 0);
  for (int k = 0; k < n_wide; ++k) {
    if (!wide[k].src || !wide[k].dst) return MNR_ENULL;
    if (wide[k].width <= 0) return MNR_ESHAPE;
    p.f[k] = wide[k];
    bulk_ok = bulk_ok && mnr_aligned(wide[k].src, 16);
  }
  Plan pl;
  if (!make_plan(wide, n_wide, T, N, &pl)) return MNR_ESHAPE;  // one 4-env tile does not fit in shared memory
  for (int k = 0; k < MNR_MAX_WIDE_FIELDS; ++k) p.nb[k] = pl.nb[k];
  for (int k = 0; k <= MNR_MAX_WIDE_FIELDS; ++k) p.item0[k] = pl.item0[k];
  p.n_wide = n_wide, p.n_scalar = pl.n_scalar;
  p.rewards = rewards, p.values = values, p.neglogp = neglogp, p.last_values = last_values, p.dones = dones;
  p.returns_tm = returns_tm, p.rewards_em = rewards_em, p.values_em = values_em, p.returns_em = returns_em;
  p.neglogp_em = neglogp_em, p.adv_em = adv_em, p.dones_em = dones_em;
  p.T = T, p.N = N, p.do_gae = do_gae, p.gamma = gamma, p.gamma_tau = gamma_tau;
  p.workspace = (unsigned char*)workspace;
  static const int kDebugSkip = [] { const char* e = getenv("MNR_POST_ROLLOUT_SKIP"); return e ? atoi(e) : 0; }();
  p.debug_skip = kDebugSkip;

  const int grid = pl.n_scalar + pl.item0[n_wide];
  cudaError_t e;
  if (bulk_ok) {
    e = mnr_allow_max_dyn_smem(post_rollout_kernel<true>);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(post_rollout_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    post_rollout_kernel<true><<<grid, kThreads, pl.smem, mnr_cs(stream)>>>(p);
  } else {
    e = mnr_allow_max_dyn_smem(post_rollout_kernel<false>);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(post_rollout_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    post_rollout_kernel<false><<<grid, kThreads, pl.smem, mnr_cs(stream)>>>(p);
  }
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_post_rollout_finalize(float* adv_em, float* values_em, float* returns_em, int64_t B,
                                         void* value_mean, void* value_var, void* value_count, int state_f64, float eps,
                                         float clamp_lo, float clamp_hi, double* stats_out, void* workspace,
                                         size_t workspace_bytes, mnr_stream_t stream) {
  MNR_CHECK_PTR(adv_em); MNR_CHECK_PTR(values_em); MNR_CHECK_PTR(returns_em); MNR_CHECK_PTR(workspace);
  if (B <= 0) return MNR_ESHAPE;
  if (workspace_bytes < kHeaderBytes) return MNR_EWORKSPACE;
  if (value_mean && (!value_var || !value_count)) return MNR_ENULL;
  finalize_stats_kernel<<<1, kFinThreads, 0, mnr_cs(stream)>>>((unsigned char*)workspace, value_mean, value_var,
                                                            value_count, state_f64, eps, stats_out);
  const int64_t want = (B / 4 + kFinThreads - 1) / kFinThreads;
  const int grid = (int)(want < 1 ? 1 : (want > 4 * mnr_sm_count() ? 4 * mnr_sm_count() : want));
  finalize_apply_kernel<<<grid, kFinThreads, 0, mnr_cs(stream)>>>(adv_em, values_em, returns_em, B, clamp_lo, clamp_hi,
                                                               (const unsigned char*)workspace);
  MNR_LAUNCH_RESULT();
}
