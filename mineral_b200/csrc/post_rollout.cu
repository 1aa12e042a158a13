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
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kHeaderBytes = 128;  // [0] n_tiles, [1] ticket ; +64: FinalStats

struct FinalStats {
  float adv_mean, adv_denom, v_mean, v_denom, r_mean, r_denom;
  int normalize_value;
  int _pad;
};

struct TilePartial {
  mnr::Moments adv, val, ret;
};

struct WideArgs {
  mnr_wide_field_t f[MNR_MAX_WIDE_FIELDS];
  int n;
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

// smem carve-up (floats): wide tiles [field][T][nb_pitch*width] | r v ret adv nlp [T][NB] | done [T][NB] u8
template <bool kBulk>
__global__ void __launch_bounds__(kThreads) post_rollout_kernel(
    WideArgs wide, const float* __restrict__ rewards, const float* __restrict__ values,
    const uint8_t* __restrict__ dones, const float* __restrict__ neglogp, const float* __restrict__ last_values,
    float* __restrict__ returns_tm, float* __restrict__ rewards_em, float* __restrict__ values_em,
    float* __restrict__ returns_em, float* __restrict__ neglogp_em, uint8_t* __restrict__ dones_em,
    float* __restrict__ adv_em, int T, int N, int NB, int do_gae, float gamma, float gamma_tau,
    unsigned char* __restrict__ workspace) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;

  const int tile = blockIdx.x;
  const int n0 = tile * NB;
  const int nb = min(NB, N - n0);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  float* s_wide = reinterpret_cast<float*>(smem_raw);
  int wide_floats = 0;
  for (int k = 0; k < wide.n; ++k) wide_floats += T * NB * wide.f[k].width;
  float* s_r = s_wide + wide_floats;
  float* s_v = s_r + T * NB;
  float* s_ret = s_v + T * NB;
  float* s_adv = s_ret + T * NB;
  float* s_nlp = s_adv + T * NB;
  uint8_t* s_done = reinterpret_cast<uint8_t*>(s_nlp + T * NB);

  if (tile == 0 && tid == 0) reinterpret_cast<unsigned int*>(workspace)[0] = gridDim.x;

  // ---- 1. kick off the wide-field loads -------------------------------------------------------
  if (kBulk) {
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) {
        uint32_t total = 0;
        for (int k = 0; k < wide.n; ++k) total += (uint32_t)(T * nb * wide.f[k].width) * 4u;
        mbar_expect_tx(&bar, total);
      }
      __syncwarp();
      float* dst = s_wide;
      for (int k = 0; k < wide.n; ++k) {
        const int w = wide.f[k].width;
        const uint32_t bytes = (uint32_t)(nb * w) * 4u;
        for (int t = lane; t < T; t += 32)
          bulk_g2s(dst + (size_t)t * NB * w, wide.f[k].src + ((size_t)t * N + n0) * w, bytes, &bar);
        dst += (size_t)T * NB * w;
      }
    }
  } else {
    float* dst = s_wide;
    for (int k = 0; k < wide.n; ++k) {
      const int w = wide.f[k].width;
      const int run = nb * w;
      for (int q = tid; q < T * run; q += kThreads) {
        const int t = q / run, j = q - t * run;
        dst[(size_t)t * NB * w + j] = mnr::ld_stream(wide.f[k].src + ((size_t)t * N + n0) * w + j);
      }
      dst += (size_t)T * NB * w;
    }
  }

  // ---- 2. stage the scalar fields ----------------------------------------------------------------
  for (int q = tid; q < T * nb; q += kThreads) {
    const int t = q / nb, nl = q - t * nb;
    const size_t g = (size_t)t * N + n0 + nl;
    const int s = t * NB + nl;
    s_r[s] = __ldg(rewards + g);
    s_v[s] = __ldg(values + g);
    s_nlp[s] = __ldg(neglogp + g);
    s_done[s] = __ldg(dones + g);
    if (!do_gae) s_ret[s] = __ldg(returns_tm + g);
  }
  __syncthreads();

  // ---- 3. GAE + moment partials: the last warp, one lane per env -----------------------------------
  if (warp == kThreads / 32 - 1) {
    mnr::Moments m_adv{0, 0, 0}, m_val{0, 0, 0}, m_ret{0, 0, 0};
    for (int nl = lane; nl < nb; nl += 32) {
      mnr::Welford wa, wv, wr;
      wa.init(), wv.init(), wr.init();
      float next_v = do_gae ? __ldg(last_values + n0 + nl) : 0.f;
      float carry = 0.f;
      for (int t = T - 1; t >= 0; --t) {
        const int s = t * NB + nl;
        const float v = s_v[s];
        float ret;
        if (do_gae) {  // experience.py:100-108, same association order as scans.cu
          const float alive = __fsub_rn(1.0f, (float)s_done[s]);
          const float delta = __fsub_rn(__fadd_rn(s_r[s], __fmul_rn(__fmul_rn(gamma, next_v), alive)), v);
          carry = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_tau, alive), carry));
          ret = __fadd_rn(carry, v);
          s_ret[s] = ret;
          next_v = v;
        } else {
          ret = s_ret[s];
        }
        const float adv = __fsub_rn(ret, v);  // experience.py:119 (returns - values)
        s_adv[s] = adv;
        wa.push(adv), wv.push(v), wr.push(ret);
      }
      m_adv = mnr::moments_merge(m_adv, wa.to_moments());
      m_val = mnr::moments_merge(m_val, wv.to_moments());
      m_ret = mnr::moments_merge(m_ret, wr.to_moments());
    }
    m_adv = mnr::moments_warp_reduce(m_adv);
    m_val = mnr::moments_warp_reduce(m_val);
    m_ret = mnr::moments_warp_reduce(m_ret);
    if (lane == 0) {
      TilePartial* parts = reinterpret_cast<TilePartial*>(workspace + kHeaderBytes);
      parts[tile] = TilePartial{m_adv, m_val, m_ret};
    }
  }

  // ---- 4. wide stores: linear walk of the tile's env-major output run ---------------------------
  if (kBulk) mbar_wait(&bar, 0);
  else __syncthreads();
  {
    const float* src = s_wide;
    for (int k = 0; k < wide.n; ++k) {
      const int w = wide.f[k].width;
      const int per_env = T * w;
      const int total = nb * per_env;
      float* out = wide.f[k].dst + (size_t)n0 * per_env;
      const int pitch = NB * w;
      auto pick = [&](int q) {
        const int nl = q / per_env;
        const int rem = q - nl * per_env;
        const int t = rem / w;
        const int f = rem - t * w;
        return src[t * pitch + nl * w + f];
      };
      if (mnr_aligned_dev(out)) {
        const int total4 = total >> 2;
        for (int q4 = tid; q4 < total4; q4 += kThreads) {
          const int q = q4 << 2;
          float4 o;
          o.x = pick(q), o.y = pick(q + 1), o.z = pick(q + 2), o.w = pick(q + 3);
          mnr::st_stream(reinterpret_cast<float4*>(out) + q4, o);
        }
        for (int q = (total4 << 2) + tid; q < total; q += kThreads) mnr::st_stream(out + q, pick(q));
      } else {
        for (int q = tid; q < total; q += kThreads) mnr::st_stream(out + q, pick(q));
      }
      src += (size_t)T * NB * w;
    }
  }
  __syncthreads();  // s_ret / s_adv complete

  // ---- 5. scalar stores (env-major run of nb*T elements, and time-major returns) --------------------
  {
    const size_t base = (size_t)n0 * T;
    for (int q = tid; q < nb * T; q += kThreads) {
      const int nl = q / T, t = q - nl * T;
      const int s = t * NB + nl;
      rewards_em[base + q] = s_r[s];
      values_em[base + q] = s_v[s];
      returns_em[base + q] = s_ret[s];
      neglogp_em[base + q] = s_nlp[s];
      adv_em[base + q] = s_adv[s];
      dones_em[base + q] = s_done[s];
    }
    if (do_gae && returns_tm) {
      for (int q = tid; q < T * nb; q += kThreads) {
        const int t = q / nb, nl = q - t * nb;
        returns_tm[(size_t)t * N + n0 + nl] = s_ret[t * NB + nl];
      }
    }
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

__global__ void __launch_bounds__(kThreads) finalize_stats_kernel(unsigned char* __restrict__ workspace, void* value_mean,
                                                                  void* value_var, void* value_count, int state_f64,
                                                                  float eps, double* __restrict__ stats_out) {
  const unsigned int n_tiles = reinterpret_cast<const unsigned int*>(workspace)[0];
  const TilePartial* parts = reinterpret_cast<const TilePartial*>(workspace + kHeaderBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  mnr::Moments a{0, 0, 0}, v{0, 0, 0}, r{0, 0, 0};
  for (unsigned int i = tid; i < n_tiles; i += kThreads) {
    a = mnr::moments_merge(a, parts[i].adv);
    v = mnr::moments_merge(v, parts[i].val);
    r = mnr::moments_merge(r, parts[i].ret);
  }
  a = mnr::moments_warp_reduce(a), v = mnr::moments_warp_reduce(v), r = mnr::moments_warp_reduce(r);
  __shared__ mnr::Moments sm[3][kThreads / 32];
  if (lane == 0) sm[0][warp] = a, sm[1][warp] = v, sm[2][warp] = r;
  __syncthreads();
  if (tid != 0) return;
  for (int w = 1; w < kThreads / 32; ++w) {
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

__global__ void __launch_bounds__(kThreads) finalize_apply_kernel(float* __restrict__ adv, float* __restrict__ values,
                                                                  float* __restrict__ returns, int64_t B, float lo,
                                                                  float hi, const unsigned char* __restrict__ workspace) {
  const FinalStats fs = *reinterpret_cast<const FinalStats*>(workspace + 64);
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t B4 = B >> 2;
  const bool vec = mnr_aligned_dev(adv) && mnr_aligned_dev(values) && mnr_aligned_dev(returns);
  if (vec) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < B4; i += stride) {
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
  for (int64_t i = (vec ? (B4 << 2) : 0) + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < B; i += stride) {
    adv[i] = __fdiv_rn(__fsub_rn(adv[i], fs.adv_mean), fs.adv_denom);
    if (fs.normalize_value) {
      values[i] = norm1(values[i], fs.v_mean, fs.v_denom, lo, hi);
      returns[i] = norm1(returns[i], fs.r_mean, fs.r_denom, lo, hi);
    }
  }
}

size_t tile_smem_bytes(const mnr_wide_field_t* wide, int n_wide, int T, int NB) {
  size_t floats = 0;
  for (int k = 0; k < n_wide; ++k) floats += (size_t)T * NB * wide[k].width;
  floats += (size_t)5 * T * NB;
  return floats * 4 + (size_t)T * NB + 16;
}

}  // namespace

extern "C" size_t mnr_post_rollout_workspace_bytes(int T, int N) {
  (void)T;
  const size_t max_tiles = (size_t)(N + 3) / 4 + 1;
  return kHeaderBytes + max_tiles * sizeof(TilePartial);
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
  if (T <= 0 || N <= 0) return MNR_ESHAPE;
  if (workspace_bytes < mnr_post_rollout_workspace_bytes(T, N)) return MNR_EWORKSPACE;
  if (!mnr_aligned(workspace, 16)) return MNR_EALIGN;

  WideArgs args;
  args.n = n_wide;
  bool bulk_ok = (N % 4 == 0);
  for (int k = 0; k < n_wide; ++k) {
    if (!wide[k].src || !wide[k].dst) return MNR_ENULL;
    if (wide[k].width <= 0) return MNR_ESHAPE;
    args.f[k] = wide[k];
    bulk_ok = bulk_ok && mnr_aligned(wide[k].src, 16);
  }

  // Tile width: minimise (waves x nb) subject to the shared-memory budget; nb % 4 == 0 keeps every
  // bulk-copy run 16 B aligned.
  const int sms = mnr_sm_count();
  const size_t smem_cap = 227 * 1024 - 1024;
  int best_nb = 0;
  double best_cost = 1e30;
  for (int nb = 4; nb <= 32; nb += 4) {
    const size_t bytes = tile_smem_bytes(wide, n_wide, T, nb);
    if (bytes > smem_cap) break;
    int per_sm = (int)((228 * 1024) / (bytes + 1024));
    if (per_sm > 8) per_sm = 8;  // 256-thread CTAs: 2048 threads / SM
    if (per_sm < 1) per_sm = 1;
    const int tiles = (N + nb - 1) / nb;
    const int waves = (tiles + sms * per_sm - 1) / (sms * per_sm);
    const double cost = (double)waves * nb + 0.01 * (32 - nb);  // tie-break: wider scalar runs
    if (cost < best_cost) best_cost = cost, best_nb = nb;
  }
  if (best_nb == 0) return MNR_ESHAPE;  // one 4-env tile does not fit in shared memory
  const int NB = best_nb;
  const size_t smem = tile_smem_bytes(wide, n_wide, T, NB);
  const int tiles = (N + NB - 1) / NB;

  cudaError_t e;
  if (bulk_ok) {
    e = cudaFuncSetAttribute(post_rollout_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    post_rollout_kernel<true><<<tiles, kThreads, smem, mnr_cs(stream)>>>(
        args, rewards, values, dones, neglogp, last_values, returns_tm, rewards_em, values_em, returns_em, neglogp_em,
        dones_em, adv_em, T, N, NB, do_gae, gamma, gamma_tau, (unsigned char*)workspace);
  } else {
    e = cudaFuncSetAttribute(post_rollout_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    post_rollout_kernel<false><<<tiles, kThreads, smem, mnr_cs(stream)>>>(
        args, rewards, values, dones, neglogp, last_values, returns_tm, rewards_em, values_em, returns_em, neglogp_em,
        dones_em, adv_em, T, N, NB, do_gae, gamma, gamma_tau, (unsigned char*)workspace);
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
  finalize_stats_kernel<<<1, kThreads, 0, mnr_cs(stream)>>>((unsigned char*)workspace, value_mean, value_var,
                                                            value_count, state_f64, eps, stats_out);
  const int64_t want = (B / 4 + kThreads - 1) / kThreads;
  const int grid = (int)(want < 1 ? 1 : (want > 4 * mnr_sm_count() ? 4 * mnr_sm_count() : want));
  finalize_apply_kernel<<<grid, kThreads, 0, mnr_cs(stream)>>>(adv_em, values_em, returns_em, B, clamp_lo, clamp_hi,
                                                               (const unsigned char*)workspace);
  MNR_LAUNCH_RESULT();
}
