/*
 * mineral_b200 C ABI  —  B200 (sm_100a) kernels for the rollout->update data path
 * of jacr13/mineral.
 *
 * The reference (pure Python, /root/reference) has no FFI: its extension points are
 * Python name look-ups (mineral/scripts/run.py:142, mineral/agents/ppo/ppo.py:47).
 * This header is therefore the boundary *underneath* the Python call surface that
 * `mineral_b200/` mirrors; every entry point cites the reference code it replaces
 * (paths relative to /root/reference).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory unless it says "host";
 *     the library never allocates, frees or retains pointers;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*): no
 *     host synchronisation, no allocation, safe under CUDA-graph capture;
 *   - return 0 on success, >0 = cudaError_t of the launch, <0 = MNR_E* argument error;
 *     there is no CPU fallback: an unsupported argument is an error;
 *   - layouts are the reference's: time-major (T,N,..) rollout storage, env-major
 *     (N*T,..) training rows (row = n*T + t), u8 dones (PPO), f32 done_mask (SHAC),
 *     bool done (replay), int64 indices;
 *   - "state" = RunningMeanStd buffers running_mean[cols], running_var[cols],
 *     running_count[1], all f32 (state_f64 = 0) or all f64 (state_f64 = 1).
 */
#ifndef MINERAL_B200_H
#define MINERAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNR_ABI_VERSION 1

#define MNR_OK 0
#define MNR_ENULL (-1)    /* required pointer is NULL */
#define MNR_ESHAPE (-2)   /* non-positive or unsupported size */
#define MNR_EALIGN (-3)   /* pointer not aligned as the kernel needs */
#define MNR_EUNSUP (-4)   /* unsupported mode / enum value */
#define MNR_EWORKSPACE (-5) /* workspace too small */

typedef void* mnr_stream_t;

int mnr_abi_version(void);
/* Static string for a code returned by any mnr_* call (host). */
const char* mnr_error_string(int code);
/* Number of SMs of the current device (host query, cached). */
int mnr_sm_count(void);

/* ------------------------------------------------------------------------------------
 * K1a  GAE reverse scan.  Replaces ExperienceBuffer.compute_return
 *      (mineral/agents/ppo/experience.py:96-108).  Bit-exact with the reference:
 *      delta = (r + (g*nv)*alive) - V;  A = delta + ((g*tau)*alive)*A';  ret = A + V.
 *      rewards/values/returns/advantages: (T,N) f32 time-major; dones (T,N) u8;
 *      last_values (N).  gamma_tau = (float)(double gamma * double tau).
 *      `advantages` may be NULL.
 * ------------------------------------------------------------------------------------ */
int mnr_gae_f32(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                float* returns, float* advantages, int T, int N, float gamma, float gamma_tau,
                mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K1b  SHAC / SAPO critic targets.  Replaces SHAC.compute_target_values and
 *      compute_target_values_with_entropy (mineral/agents/diffrl/shac.py:877-916).
 *      rew/done_mask/next_values/target: (T,N) f32; entropy (T,N) f32 or NULL.
 *      method: 0 = one-step, 1 = td-lambda.
 *      ent_flags: bit0 negate entropy input (use -logprobs), bit1 offset by |H_t| then *0.5,
 *                 bit2 scale by 1/|H_t|.  abs_target_entropy = |H_t|,
 *      inv_abs_target_entropy = (float)(1.0/|H_t|).  Scalars are the reference's Python
 *      doubles rounded to f32: one_minus_lam = (float)(1.0-lam), lam_gamma = (float)(lam*gamma).
 * ------------------------------------------------------------------------------------ */
int mnr_critic_targets_f32(const float* rew, const float* done_mask, const float* next_values,
                           const float* entropy, float* target, int T, int N, int method,
                           float gamma, float lam, float one_minus_lam, float lam_gamma, float alpha,
                           int ent_flags, float abs_target_entropy, float inv_abs_target_entropy,
                           mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K1a+K3a fused post-rollout pass.  Replaces compute_return + transform_op +
 *      prepare_training (experience.py:7-14,96-121) and produces the batch statistics that
 *      the advantage normalisation (experience.py:119-120) and PPO's value-RMS block
 *      (mineral/agents/ppo/ppo.py:175-183) need, in ONE pass over the rollout.
 *
 *      wide fields: n_wide descriptors (host array) of (T,N,width) f32 -> (N*T,width) f32.
 *      scalars: rewards, values, neglogp (T,N) f32, dones (T,N) u8, last_values (N).
 *      do_gae = 1: run the GAE scan in-kernel and (if returns_tm != NULL) also store the
 *                  time-major returns (storage_dict['returns']); do_gae = 0: read returns_tm.
 *      outputs env-major: rewards_em, values_em, returns_em, neglogp_em, adv_em (B) f32,
 *      dones_em (B) u8, B = N*T.  adv_em = returns - values, NOT yet normalised.
 *      partials: workspace of mnr_post_rollout_workspace_bytes(T,N) bytes holding per-CTA
 *      (count, mean, M2) for adv / values / returns, consumed by mnr_post_rollout_finalize.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  const float* src; /* (T,N,width) time-major */
  float* dst;       /* (N*T,width) env-major  */
  int32_t width;    /* floats per (t,n) */
  int32_t _pad;
} mnr_wide_field_t;

#define MNR_MAX_WIDE_FIELDS 8

size_t mnr_post_rollout_workspace_bytes(int T, int N);

int mnr_post_rollout_fused(const mnr_wide_field_t* wide /*host*/, int n_wide,
                           const float* rewards, const float* values, const uint8_t* dones,
                           const float* neglogp, const float* last_values, float* returns_tm,
                           float* rewards_em, float* values_em, float* returns_em, float* neglogp_em,
                           uint8_t* dones_em, float* adv_em,
                           int T, int N, int do_gae, float gamma, float gamma_tau,
                           void* workspace, size_t workspace_bytes, mnr_stream_t stream);

/* Finalise: merge the per-CTA partials in index order (deterministic), then
 *   adv   <- (adv - mean(adv)) / (std_unbiased(adv) + 1e-8)          (experience.py:119-120)
 *   if value_mean != NULL (normalize_value):                          (ppo.py:177-181)
 *     state.update(values); values <- normalize(values); state.update(returns);
 *     returns <- normalize(returns)   with RunningMeanStd semantics (normalizers.py:44-68)
 *   value state is updated IN PLACE.  clamp_lo > clamp_hi disables the clamp.
 * stats_out (10 doubles, may be NULL): adv mean,std; values batch mean,var; returns batch mean,var;
 *   state mean,var after 1st update; state mean,var after 2nd update. */
int mnr_post_rollout_finalize(float* adv_em, float* values_em, float* returns_em, int64_t B,
                              void* value_mean, void* value_var, void* value_count, int state_f64,
                              float eps, float clamp_lo, float clamp_hi, double* stats_out,
                              void* workspace, size_t workspace_bytes, mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K2  RunningMeanStd.  Replaces mineral/common/normalizers.py:44-78.
 * ------------------------------------------------------------------------------------ */
size_t mnr_rms_moments_workspace_bytes(int64_t rows, int cols, int64_t group_rows);

/* Per-column batch moments of x (rows, cols) f32 for each row group
 * g = [g*group_rows, min((g+1)*group_rows, rows)):  mean_out/var_out (n_groups, cols) f32,
 * var with `correction` (x.mean(0), x.var(0, correction), normalizers.py:58-59).
 * Welford per thread, Chan merges in f64 in a fixed order (deterministic).
 * Every `workspace` in this API must be zero-filled once before its first use; kernels leave the
 * ticket counters they keep there at zero again. */
int mnr_rms_batch_moments_f32(const float* x, int64_t rows, int cols, int64_t group_rows, int correction,
                              float* mean_out, float* var_out, void* workspace, size_t workspace_bytes,
                              mnr_stream_t stream);

/* In-place Chan merge of one batch's moments into the running state, op-for-op the
 * reference's _update_from_moments (normalizers.py:44-55) in the state dtype. */
int mnr_rms_merge(void* mean, void* var, void* count, int state_f64, const float* batch_mean,
                  const float* batch_var, double batch_count, int cols, mnr_stream_t stream);

/* Pre-compute the whole normaliser schedule of one PPO update: for step s = 0..n_steps-1 merge
 * row-group (s % n_groups)'s moments (as produced by mnr_rms_batch_moments_f32 with the same rows /
 * group_rows; the last group may be short) into the state, exactly what obs_rms.update does once
 * per minibatch (ppo.py:258-262), and record mean_tab/denom_tab (n_steps, cols) f32 =
 * running_mean.float(), sqrt(running_var.float() + eps) after that step.  State updated in place. */
int mnr_rms_schedule(void* mean, void* var, void* count, int state_f64, const float* group_mean,
                     const float* group_var, int64_t group_rows, int64_t rows, int n_steps, int cols,
                     float eps, float* mean_tab, float* denom_tab, mnr_stream_t stream);

/* mean_f32/denom_f32 (cols) <- running_mean.float(), sqrt(running_var.float()+eps). */
int mnr_rms_prepare(const void* mean, const void* var, int state_f64, int cols, float eps,
                    float* mean_f32, float* denom_f32, mnr_stream_t stream);

/* y = clamp((x - mean)/denom) (normalize, normalizers.py:65-68) or, inverse = 1,
 * y = denom*clamp(x) + mean (unnormalize, :70-73).  x,y (rows, cols) f32, may alias.
 * clamp_lo > clamp_hi disables the clamp. */
int mnr_rms_apply_f32(const float* x, float* y, int64_t rows, int cols, const float* mean_f32,
                      const float* denom_f32, float clamp_lo, float clamp_hi, int inverse,
                      mnr_stream_t stream);
/* same, writing rows of y with a row stride of y_ld >= cols floats (pad columns untouched) */
int mnr_rms_apply_ld_f32(const float* x, float* y, int64_t rows, int cols, int y_ld, const float* mean_f32,
                         const float* denom_f32, float clamp_lo, float clamp_hi, mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K3  indexed row movement: minibatch gather, replay insert / sample.
 * ------------------------------------------------------------------------------------ */
#define MNR_ROW_F32 0        /* copy f32 rows */
#define MNR_ROW_U8 1         /* copy 1-byte rows (u8 / bool) */
#define MNR_ROW_BOOL_TO_F32 2 /* src bool/u8 -> dst f32 (done.float(), simple_replay.py:84) */
#define MNR_ROW_F32_TO_BOOL 3 /* src f32 -> dst bool (x != 0; dones.bool(), simple_replay.py:39) */
#define MNR_ROW_F32_NORMALIZE 4 /* dst = clamp((src - mean)/denom), fused obs normalisation */

typedef struct {
  const void* src;
  void* dst;
  int32_t width;      /* elements per row */
  int32_t kind;       /* MNR_ROW_* */
  const float* mean;  /* kind 4 only */
  const float* denom; /* kind 4 only */
  float clamp_lo, clamp_hi;
  int32_t dst_ld;     /* row stride of dst in elements; 0 = width (lets a gather write into a column block of a wider buffer) */
  int32_t _pad;
} mnr_row_field_t;

#define MNR_MAX_ROW_FIELDS 12

/* dst[i] = src[idx[i]] for every field, i < n_idx.  idx int64 in [0, src_rows).  Serves the
 * PPO minibatch gather (identity / permutation index), CriticDataset.shuffle
 * (mineral/agents/diffrl/utils.py:55-58) and ReplayBuffer.sample_batch
 * (mineral/buffers/simple_replay.py:73-85).  idx == NULL means rows [row0, row0+n_idx). */
int mnr_gather_rows(const mnr_row_field_t* fields /*host*/, int n_fields, const int64_t* idx,
                    int64_t row0, int64_t n_idx, mnr_stream_t stream);

/* dst[idx[i]] = src[i] (scatter; update_mu_sigma, experience.py:83-87). */
int mnr_scatter_rows(const mnr_row_field_t* fields /*host*/, int n_fields, const int64_t* idx,
                     int64_t row0, int64_t n_idx, mnr_stream_t stream);

/* Ring insert with the reference's wrap rule (simple_replay.py:32-71): n_rows incoming rows at
 * next_p into a ring of `capacity` rows; head [next_p,C) takes the first C-next_p rows, tail
 * [0,p) takes the LAST p rows.  Pointer bookkeeping (next_p, if_full) stays on the host. */
int mnr_replay_insert(const mnr_row_field_t* fields /*host*/, int n_fields, int64_t n_rows,
                      int64_t next_p, int64_t capacity, mnr_stream_t stream);

/* The replay ring as ONE array of 128 B-aligned transition records (C, rec_floats), rec_floats % 32 == 0: field k (obs,
 * next_obs, action, reward, done ...) occupies floats [rec_offsets[k], rec_offsets[k] + width) of every record, offsets
 * multiples of 4; a 1-wide bool field keeps its byte in byte 0 of its slot.  The reference's five arrays
 * (mineral/buffers/simple_replay.py:4-16) become strided views of this array, so a uniform sample reads whole cache lines.
 *   mnr_replay_insert_records: fields[k].src = incoming (n_rows, width) array, kind MNR_ROW_F32 | MNR_ROW_U8 (bool done) |
 *       MNR_ROW_F32_TO_BOOL; ring slots by the reference's wrap rule (simple_replay.py:32-71) like mnr_replay_insert.
 *   mnr_replay_sample_records: out rows i < n_idx <- record idx[i] (simple_replay.py:73-85): fields[k].dst (+ dst_ld) with
 *       kind MNR_ROW_F32 | MNR_ROW_BOOL_TO_F32 (done.float()) | MNR_ROW_F32_NORMALIZE (fused obs normalise + clamp);
 *       records are staged in shared memory by one cp.async.bulk per row (TMA), 32 rows per CTA. */
int mnr_replay_insert_records(const mnr_row_field_t* fields /*host*/, const int* rec_offsets /*host*/, int n_fields,
                              float* records, int rec_floats, int64_t n_rows, int64_t next_p, int64_t capacity,
                              mnr_stream_t stream);
int mnr_replay_sample_records(const mnr_row_field_t* fields /*host*/, const int* rec_offsets /*host*/, int n_fields,
                              const float* records, int rec_floats, const int64_t* idx, int64_t n_idx, mnr_stream_t stream);

/* N-step FIFO push (mineral/buffers/nstep_replay.py:29-94) as a true window, no re-concatenation.
 * in_*: (N,S,..) this call's transitions (obs/next_obs width O, action width A, reward 1, done 1).
 * win_*_in: previous window (N,n,..) oldest first; win_*_out: window after the call (distinct memory).
 * out_*: ((S-skip)*N, ..) step-major rows, skip = max(0, n-1-count_before) warm-up steps.
 * gammas: (n) f32 = (float)(gamma**k).  The reference only works with FLOAT dones (its argmax
 * rejects bool), and then returns `done` as float; both dtypes are accepted here. */
typedef struct {
  const float *obs, *act, *rew, *nobs;
  const void* done;     /* u8/bool, or f32 when done_is_f32 (sac.py:150 keeps dones in a float tensor) */
  int32_t done_is_f32;
  int32_t _pad;
} mnr_nstep_in_t;
typedef struct {
  float *obs, *act, *rew, *nobs;
  void* done;           /* u8/bool, or f32 (0.0 / 1.0) when done_is_f32 */
  int32_t done_is_f32;
  int32_t _pad;
} mnr_nstep_out_t;
int mnr_nstep_push(const mnr_nstep_in_t* in /*host*/, const mnr_nstep_in_t* win_in /*host*/,
                   const mnr_nstep_out_t* win_out /*host*/, const mnr_nstep_out_t* out /*host*/,
                   const float* gammas, int N, int S, int O, int A, int nstep, int64_t count_before,
                   mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K4  fused PPO loss forward + backward.  Replaces, per minibatch, Dist.forward + entropy +
 *     neglogp (mineral/nets/dist.py:31-47, mineral/agents/ppo/models.py:93-108), actor_loss /
 *     critic_loss / bounds_loss / policy_kl (ppo.py:449-498), the means and the total
 *     (ppo.py:283-285), autograd's backward of all of it, and update_mu_sigma
 *     (experience.py:83-87).
 * ------------------------------------------------------------------------------------ */
#define MNR_DIST_NORMAL 0
#define MNR_DIST_DREAMERV3_NORMAL 1
#define MNR_BOUNDS_BOUND 0
#define MNR_BOUNDS_REG 1

typedef struct {
  int32_t dist_type;      /* MNR_DIST_* */
  int32_t bounds_type;    /* MNR_BOUNDS_* */
  int32_t use_smooth_clamp;
  int32_t clip_value_loss;
  float minstd, maxstd;   /* dreamerv3_normal */
  float e_clip;
  float critic_coef, entropy_coef, bounds_loss_coef;
  float soft_bound;       /* 1.1 */
  float grad_scale;       /* d(total)/d(loss): 1.0, or 1/world_size before a SUM gradient all-reduce */
  float metric_scale;     /* metrics_out *= metric_scale (1.0, or 1/world_size before a SUM all-reduce) */
  int32_t dz_act;         /* mnr_ppo_heads_loss_fwd_bwd, single-kernel form only: 0 -> dy3a / dy3v receive d loss / d Y3 (default);
                             1 (ELU) / 2 (ReLU) -> they receive dZ3 = dY3 * act'(Y3), taken from the Y3 tile the kernel has
                             staged anyway, so the top layer's backward GEMMs run with act = 0 and load no saved-output tile */
} mnr_ppo_loss_cfg_t;

#define MNR_PPO_METRICS 10
/* metrics_out (MNR_PPO_METRICS f32): 0 total loss, 1 actor, 2 critic, 3 bounds, 4 entropy,
 * 5 clip_frac, 6 explained_var, 7 kl(new||old), 8 mean(mu) [unused], 9 reserved. */

size_t mnr_ppo_loss_workspace_bytes(int64_t mb, int A);

/* Inputs (mb rows): mu_raw, logstd (mb,A) network outputs BEFORE Dist; logstd_row_stride = A, or 0
 * when one (A) vector is shared by all rows (fixed sigma, models.py:151); value (mb);
 * actions (mb,A); old_neglogp, advantages, old_values, returns (mb).
 * old_mu / old_sigma (mb,A): read for the KL, then OVERWRITTEN with the fresh mu / sigma if
 * write_back != 0.  Outputs: g_mu_raw, g_logstd (mb,A), g_value (mb) = d total / d input (means
 * included); mu_mean_out/sigma_mean_out (A) column means of the fresh mu/sigma (may be NULL). */
int mnr_ppo_loss_fwd_bwd(const float* mu_raw, const float* logstd, int logstd_row_stride,
                         const float* value, const float* actions, const float* old_neglogp,
                         const float* advantages, const float* old_values, const float* returns,
                         float* old_mu, float* old_sigma, int write_back,
                         float* g_mu_raw, float* g_logstd, float* g_value,
                         float* metrics_out, float* mu_mean_out, float* sigma_mean_out,
                         int64_t mb, int A, const mnr_ppo_loss_cfg_t* cfg /*host*/,
                         void* workspace, size_t workspace_bytes, mnr_stream_t stream);

/* Fused policy/value HEADS + loss (two launches): reads the towers' last activations y3a / y3v (M,H), applies
 * the mu / sigma / value heads (mineral/agents/ppo/models.py:136-151: w_mu (A,H), b_mu (A), w_v (H), b_v (1);
 * learned sigma: w_sig (A,H), b_sig (A); fixed sigma: w_sig = NULL and b_sig = the sigma parameter (A)), runs
 * everything mnr_ppo_loss_fwd_bwd does, and returns the gradients the towers' backward needs, dy3a / dy3v (M,H),
 * plus every head parameter gradient (dw_sig = NULL with fixed sigma; db_sig then receives d sigma).
 * g_mu / g_ls (M,A) and g_v (M) are scratch outputs.  H % 8 == 0, H <= 128, A <= 32. */
/* 1 if (A, H) takes the single persistent kernel (csrc/heads_fused.cu: H in {32,64,128}, A <= 32), 0 if the
 * two-kernel form (heads+loss, then head gradients) is used. */
int mnr_ppo_heads_is_fused(int A, int H);
size_t mnr_ppo_heads_workspace_bytes(int64_t M, int A, int H);
int mnr_ppo_heads_loss_fwd_bwd(const float* y3a, const float* y3v, const float* w_mu, const float* b_mu,
                               const float* w_v, const float* b_v, const float* w_sig, const float* b_sig,
                               const float* actions, const float* old_neglogp, const float* advantages,
                               const float* old_values, const float* returns, float* old_mu, float* old_sigma,
                               int write_back, float* dy3a, float* dy3v, float* g_mu, float* g_ls, float* g_v,
                               float* dw_mu, float* db_mu, float* dw_v, float* db_v, float* dw_sig, float* db_sig,
                               float* metrics_out, float* mu_mean_out, float* sigma_mean_out, int64_t M, int A, int H,
                               const mnr_ppo_loss_cfg_t* cfg /*host*/, void* workspace, size_t workspace_bytes,
                               mnr_stream_t stream);

/* Un-reduced module-level twins of the reference's loss functions (ppo.py:453-498): per-row value
 * and per-row derivative.  clip_count_out / ev_out single f32 (clip_frac, explained_var). */
int mnr_actor_loss_rows(const float* old_neglogp, const float* neglogp, const float* advantages,
                        int64_t n, float e_clip, int use_smooth_clamp, float* loss_out,
                        float* dloss_dneglogp_out, float* clip_frac_out, void* workspace,
                        size_t workspace_bytes, mnr_stream_t stream);
int mnr_critic_loss_rows(const float* old_values, const float* values, const float* returns, int64_t n,
                         float e_clip, int clip_value_loss, float* loss_out, float* dloss_dvalues_out,
                         float* explained_var_out, void* workspace, size_t workspace_bytes,
                         mnr_stream_t stream);
int mnr_bounds_loss_rows(const float* mu, int64_t n, int A, int bounds_type, float soft_bound,
                         float* loss_out, float* dloss_dmu_out, mnr_stream_t stream);
int mnr_policy_kl_rows(const float* p0_mu, const float* p0_sigma, const float* p1_mu,
                       const float* p1_sigma, int64_t n, int A, float* kl_rows_out, float* kl_mean_out,
                       void* workspace, size_t workspace_bytes, mnr_stream_t stream);
size_t mnr_rows_reduce_workspace_bytes(int64_t n);

/* ------------------------------------------------------------------------------------
 * G  actor/critic MLP towers (mineral/nets/mlp.py:25-77, mineral/agents/ppo/models.py:136-151):
 *    fused linear-layer kernels that replace nn.Linear + activation + autograd on the update path.
 *    A "group" is one tower's layer; both towers go in one launch.  act: 0 identity, 1 ELU, 2 ReLU.
 *      fwd:        c = Y (M,N) = act(a=X (M,K) . b=W (N,K)^T + bias (N))
 *      bwd_input:  c = dX (M,K) (+)= (a=dY (M,N) * act'(saved=Y)) . b=W (N,K)
 *      bwd_weight: c = dW (N,K) = (a=dY * act'(saved=Y))^T . b=X (M,K);  bias_grad (N) = column sums
 *    saved == NULL means identity (no activation).  Exact fp32 FMA accumulation.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  const float* a;
  const float* b;
  float* c;
  const float* bias;   /* fwd only, may be NULL */
  const float* saved;  /* bwd only: forward output Y, may be NULL */
  float* bias_grad;    /* bwd_weight only, may be NULL */
  const float* b_lo;   /* tensor-core path only: low tf32 part of W (mnr_split_tf32); b is then the high part */
  int32_t N, K;
  int32_t lda;         /* row stride (floats) of X when it is the `a` operand (fwd); 0 = K */
  int32_t ldb;         /* row stride (floats) of the `b` operand (W, or X in bwd_weight); 0 = K */
} mnr_linear_group_t;

#define MNR_MAX_LINEAR_GROUPS 4

int mnr_linear_fwd(const mnr_linear_group_t* groups /*host*/, int n_groups, int M, int act, mnr_stream_t stream);
int mnr_linear_bwd_input(const mnr_linear_group_t* groups /*host*/, int n_groups, int M, int act, int accumulate,
                         mnr_stream_t stream);
size_t mnr_linear_bwd_weight_workspace_bytes(int n_groups, int M, int N, int K);
/* all groups must share N and K */
int mnr_linear_bwd_weight(const mnr_linear_group_t* groups /*host*/, int n_groups, int M, int act, void* workspace,
                          size_t workspace_bytes, mnr_stream_t stream);
/* Tensor-core variants (tcgen05 kind::tf32, 3xTF32 split: fp32-level accuracy, ~1e-6 relative).
 * Same semantics as the three calls above with these extra requirements, else MNR_ESHAPE / MNR_EALIGN:
 *   - at most 2 groups, all with the same N and K; N % 16 == 0 (fwd) / N % 4 == 0 (bwd);
 *   - every matrix base 16 B aligned and every row stride a multiple of 4 floats (TMA); pad K with
 *     lda / ldb (zero-filled pad columns) when K % 4 != 0;
 *   - fwd / bwd_input read W as pre-split b (high tf32 part) and b_lo (remainder) from mnr_split_tf32;
 *   - bwd calls need `saved` (pass act = 0 and saved = a for an identity layer); no accumulate.
 * mnr_tc_supported() is 0 when the driver entry point for tensor maps is unavailable. */
int mnr_tc_supported(void);
int mnr_tc_linear_fwd(const mnr_linear_group_t* groups /*host*/, int n_groups, int M, int act, mnr_stream_t stream);
int mnr_tc_linear_bwd_input(const mnr_linear_group_t* groups /*host*/, int n_groups, int M, int act, mnr_stream_t stream);
size_t mnr_tc_linear_bwd_weight_workspace_bytes(int n_groups, int M, int N, int K);
int mnr_tc_linear_bwd_weight(const mnr_linear_group_t* groups /*host*/, int n_groups, int M, int act, void* workspace,
                             size_t workspace_bytes, mnr_stream_t stream);
/* The same in two steps, so that the split-reduce of several layers runs as ONE launch at the end of the backward
 * pass: _partials runs only the tensor-core kernel (per-split dW / db partials stay in `workspace`, one workspace per
 * layer); mnr_tc_reduce_partials then sums the splits of up to 4 such layers (same groups / M as given to _partials)
 * into their dW (groups[].c) and db (groups[].bias_grad), in a fixed order. */
int mnr_tc_linear_bwd_weight_partials(const mnr_linear_group_t* groups /*host*/, int n_groups, int M, int act, void* workspace,
                                      size_t workspace_bytes, mnr_stream_t stream);
typedef struct {
  const mnr_linear_group_t* groups; /* host */
  int32_t n_groups, M;
  void* workspace;
  size_t workspace_bytes;
  int32_t grid_cap;                 /* the mnr_tc_grid_cap value in force when _partials was launched (0 = none) */
  int32_t _pad;
} mnr_tc_reduce_job_t;
/* CTAs the following mnr_tc_linear_* launches of this host thread may use (0 = one per SM, the default); returns the previous
 * value.  A layer's input-gradient and weight-gradient launches do not depend on each other: capped to complementary shares of
 * the SMs and issued on two streams they run side by side.  mnr_tc_linear_bwd_weight_workspace_bytes depends on the cap in
 * force (the split count follows the CTA count): size the workspace under the same cap as the launch. */
int mnr_tc_grid_cap(int ctas);
/* L2 cache hints for the following mnr_tc_linear_* launches of this host thread (returns the previous value; 0 = none): bit 0 -
 * the result (dX) is read by the very next launches: its TMA stores ask for evict_last; bit 1 - the launch is the LAST reader
 * of its dY / saved-output operands in this optimizer step: their TMA loads ask for evict_first.  One optimizer step of the PPO
 * towers touches ~470 MB of activations and gradients against 126 MB of L2: the hints keep what is about to be re-read. */
int mnr_tc_cache_hints(int flags);
int mnr_tc_reduce_partials(const mnr_tc_reduce_job_t* jobs /*host*/, int n_jobs, mnr_stream_t stream);

/* Deferred folds.  Two kernels of an optimizer step end with "every CTA parks a partial, grid barrier, fixed-order sum over
 * the CTAs": the heads + loss kernel (head parameter gradients, mu / sigma means, the ten metrics; mnr_ppo_heads_loss_fwd_bwd)
 * and LayerNorm's backward (dgamma / dbeta; mnr_ln_act_bwd).  Nothing reads those sums before the optimizer launch
 * (autograd's .grad of mineral/agents/ppo/ppo.py:292-311), so while a fold list is registered with mnr_fold_defer() the two
 * entry points stop after parking their partials - no grid barrier, no tail - and append a description of the outstanding
 * sums to the list; mnr_tc_reduce_partials_folds() then computes them in extra blocks of the ONE split-reduce launch that
 * precedes the optimizer anyway (fixed order: deterministic).  The list is host memory owned by the caller; the partials live
 * in the workspaces given to the deferring calls, which must stay untouched until the reduce launch has run. */
#define MNR_FOLD_MAX_SEGS 8
#ifndef MNR_MAX_FOLDS
#define MNR_MAX_FOLDS 8
#endif
typedef struct {
  int32_t begin, end;   /* columns [begin, end) of the partial rows ... */
  float* dst;           /* ... summed over the rows, times scale, to dst[0 .. end - begin) */
  float scale;
  int32_t _pad;
} mnr_fold_seg_t;
typedef struct {
  const float* parts;   /* [n_parts][width] */
  int32_t n_parts, width, n_segs, _pad;
  mnr_fold_seg_t seg[MNR_FOLD_MAX_SEGS];
} mnr_fold_job_t;
typedef struct {        /* the ten PPO metrics from the per-CTA loss partials of the heads + loss kernel */
  const double* loss_parts;   /* [n_parts][12] */
  float* metrics;             /* (MNR_PPO_METRICS) */
  int64_t M;
  int32_t n_parts, _pad;
  float critic_coef, entropy_coef, bounds_loss_coef, metric_scale;
} mnr_heads_metrics_job_t;
typedef struct {
  mnr_fold_job_t folds[MNR_MAX_FOLDS];
  mnr_heads_metrics_job_t metrics;
  int32_t n_folds, has_metrics;
} mnr_fold_list_t;
/* Register (list != NULL: emptied, then appended to by the deferring calls of THIS host thread) or clear (NULL) the list. */
int mnr_fold_defer(mnr_fold_list_t* list /*host*/);
/* mnr_tc_reduce_partials + the folds of `list` (may be NULL or empty) in the same launch; n_jobs == 0 with a non-empty list
 * launches the folds alone. */
int mnr_tc_reduce_partials_folds(const mnr_tc_reduce_job_t* jobs /*host*/, int n_jobs, const mnr_fold_list_t* list /*host*/,
                                 mnr_stream_t stream);

/* A whole 3-layer tower forward in ONE launch (csrc/tc_tower.cu): y[l] = act(y[l-1] . W_l^T + bias_l), y[-1] = x, for every
 * group (tower), replacing MLP.forward (mineral/nets/mlp.py:25-77; both PPO towers: mineral/agents/ppo/models.py:136-151).
 * Layer outputs go from one GEMM to the next through tensor memory; every y[l] is also written (the backward pass needs it).
 * Same numerics as three mnr_tc_linear_fwd calls (3xTF32).  Shapes: K0 <= 64 (row strides multiples of 4 floats),
 * widths[0] in {128, 256}, widths[1], widths[2] multiples of 32 and <= 128; anything else returns MNR_EUNSUP and
 * mnr_tc_tower_fwd_supported() == 0 (the caller then runs the per-layer kernels). */
typedef struct {
  const float* x;        /* (M, K0), row stride ldx */
  const float* w_hi[3];  /* W_l (N_l, K_l) high tf32 part, row stride ldw[l] (mnr_split_tf32) */
  const float* w_lo[3];  /* remainder */
  const float* bias[3];
  float* y[3];           /* (M, N_l) contiguous */
  int32_t ldx;           /* 0 = K0 */
  int32_t ldw[3];        /* 0 = K_l */
} mnr_tower_group_t;
int mnr_tc_tower_fwd_supported(int K0, const int32_t* widths, int n_layers, int act);
int mnr_tc_tower_fwd(const mnr_tower_group_t* groups /*host*/, int n_groups, int M, int K0, const int32_t* widths /*host*/,
                     int n_layers, int act, mnr_stream_t stream);
int mnr_tc_tower_trace(int enable, long long* host_out);
/* same for the fused heads + loss kernel (csrc/heads_fused.cu): 16 phase boundaries of CTA 0, globaltimer ns (32 int64) */
int mnr_heads_trace(int enable, long long* host_out);

/* Profiling aid: enable != 0 makes CTA 0 of the tensor-core kernels record clock64() at every role's wait / done
 * points; host_out (4 roles x 256 events x 2 tags int64, may be NULL) receives the last timeline. Synchronous. */
int mnr_tc_trace(int enable, long long* host_out);

/* hi = x with the 13 low mantissa bits cleared (a tf32), lo = x - hi, written to (rows, dst_ld) buffers
 * with zero-filled pad columns; all jobs in one launch. */
typedef struct {
  const float* src;
  float* hi;
  float* lo;
  int32_t rows, cols, src_ld, dst_ld;
} mnr_split_job_t;
#define MNR_MAX_SPLIT_JOBS 8
int mnr_split_tf32(const mnr_split_job_t* jobs /*host*/, int n_jobs, mnr_stream_t stream);

/* out (C) = column sums of x (M,C), C <= 256: gradient of a broadcast parameter (fixed sigma, models.py:151) */
size_t mnr_colsum_workspace_bytes(int C);
int mnr_colsum_f32(const float* x, int64_t M, int C, float* out, void* workspace, size_t workspace_bytes,
                   mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * a12  gradient-norm clip + Adam/AdamW with DEVICE-resident learning rate and step count, and
 *      the KL-adaptive LR rule (ppo.py:291-300,338-345; mineral/agents/ppo/utils.py:14-20), so
 *      the optimizer step needs no host round trip and is CUDA-graph capturable.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  double lr;          /* current learning rate (Python double in the reference) */
  double kl_sum;      /* running sum of per-step KL within the mini-epoch */
  int32_t kl_count;
  int32_t step;       /* Adam step count */
  float grad_norm;    /* last total norm */
  float clip_coef;    /* last clip coefficient (<= 1) */
  float step_size;    /* lr / (1 - beta1^step) */
  float bc2_sqrt;     /* sqrt(1 - beta2^step) */
  float lr_f32;
  float _pad;
} mnr_optim_state_t;

size_t mnr_grad_norm_workspace_bytes(int64_t n);

/* total_norm = ||g||_2 (torch clip_grad_norm_), clip_coef = min(1, max_norm/(norm+1e-6)) (1 if
 * max_norm <= 0), st->step += 1 and the Adam scalars for that step.  If kl != NULL its value is
 * added to st->kl_sum (kl_count += 1).  grad_norm_out (f32, may be NULL) receives the norm. */
int mnr_grad_norm_prepare(const float* grad, int64_t n, float max_norm, float beta1, float beta2,
                          mnr_optim_state_t* st, const float* kl, float* grad_norm_out,
                          void* workspace, size_t workspace_bytes, mnr_stream_t stream);

/* torch.optim.Adam (decoupled = 0: grad += wd*p) / AdamW (decoupled = 1: p *= 1 - lr*wd) single
 * step over a flat parameter buffer; grads are scaled by st->clip_coef first; zero_grad != 0
 * clears `grad` afterwards (optim.zero_grad, ppo.py:291). */
int mnr_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float beta1,
                  float beta2, float eps, float weight_decay, int decoupled, int zero_grad,
                  const mnr_optim_state_t* st, mnr_stream_t stream);

/* End of a mini-epoch: avg_kl = kl_sum/kl_count (as f32); lr /= 1.5 if avg_kl > 2*thr (floor
 * min_lr), lr *= 1.5 if avg_kl < thr/2 (cap max_lr); reset the KL accumulator.
 * avg_kl_out / lr_out (f32 / f64 device, may be NULL) record the values for the host. */
int mnr_kl_lr_update(mnr_optim_state_t* st, double kl_threshold, double min_lr, double max_lr,
                     float* avg_kl_out, double* lr_out, mnr_stream_t stream);

/* End of a mini-epoch under lr_schedule 'cos' / 'linear' (mineral/agents/ppo/ppo.py:339-348,
 * ppo/utils.py:23-37): those rules depend on host counters only (mini-epoch index, agent_steps), so the
 * host uploads the value that takes effect after each mini-epoch (*lr_new, f64 device) and this call
 * installs it; avg_kl = kl_sum/kl_count is reported and the accumulator reset as in mnr_kl_lr_update. */
int mnr_lr_table_update(mnr_optim_state_t* st, const double* lr_new, float* avg_kl_out, double* lr_out,
                        mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * a10  LayerNorm + activation between the tower GEMMs (mineral/nets/mlp.py:66-71; the shipped PPO recipes
 *      mineral/cfgs/agent/DFlexAntPPO.yaml:17-21 and AllegroHandPPO.yaml:18-28 use Linear -> LayerNorm -> SiLU).
 *      fwd: y = act(gamma * (z - mean) * rstd + beta), mean / rstd (M) saved;  biased variance, eps inside the sqrt.
 *      bwd: dz from dy (through act' and the normalisation), dgamma / dbeta (N) summed over the M rows
 *           (deterministic).  act: 0 none, 1 ELU, 2 ReLU, 3 SiLU.  N <= 512.  Up to 2 groups per launch.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  const float* z;      /* (M,N) pre-normalisation */
  float* y;            /* fwd out (M,N) */
  const float* gamma;  /* (N) */
  const float* beta;   /* (N) */
  float* mean;         /* (M): fwd out, bwd in */
  float* rstd;         /* (M) */
  const float* dy;     /* bwd in (M,N) */
  float* dz;           /* bwd out (M,N) */
  float* dgamma;       /* bwd out (N) */
  float* dbeta;        /* bwd out (N) */
} mnr_ln_group_t;

int mnr_ln_act_fwd(const mnr_ln_group_t* groups /*host*/, int n_groups, int64_t M, int N, float eps, int act,
                   mnr_stream_t stream);
size_t mnr_ln_act_bwd_workspace_bytes(int n_groups, int64_t M, int N);
int mnr_ln_act_bwd(const mnr_ln_group_t* groups /*host*/, int n_groups, int64_t M, int N, int act, void* workspace,
                   size_t workspace_bytes, mnr_stream_t stream);

/* Single-GPU fused optimizer step: mnr_grad_norm_prepare + mnr_adam_step in ONE launch (grid-wide
 * barrier between the norm and the update; grid <= #SMs so every CTA is resident).  n % 4 == 0, all
 * buffers 16 B aligned.  workspace: mnr_fused_adam_workspace_bytes() bytes, zero-filled once.
 * nan_to_num != 0: non-finite gradient entries count as 0 (SHAC's grad.nan_to_num_(0,0,0),
 * mineral/agents/diffrl/shac.py:841-843); clipped_norm_out (may be NULL) = norm * clip coefficient, the
 * "grad_norm_after_clip" SHAC logs (shac.py:846-856). */
size_t mnr_fused_adam_workspace_bytes(void);
/* Optional: a table of up to 8 weight matrices inside the (zero-initialised) optimizer workspace, at this byte offset:
 *   { int32 count; int32 pad[3]; { int64 off; int32 n, K, ld, pad; float* hi; float* lo; } entry[8]; }
 * off = first element of the matrix in the flat parameter buffer (multiple of 4), n = rows*K, ld = row stride of hi / lo.
 * When count > 0, mnr_fused_clip_adam / mnr_peer_reduce_clip_adam also write the tf32 (hi, lo) split of every new weight
 * (what mnr_split_tf32 computes), so the update chain has no separate split launch per optimizer step. */
size_t mnr_fused_adam_split_table_offset(void);
int mnr_fused_clip_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float max_norm,
                        float beta1, float beta2, float eps, float weight_decay, int decoupled, int zero_grad,
                        mnr_optim_state_t* st, const float* kl, float* grad_norm_out, int nan_to_num,
                        float* clipped_norm_out, void* workspace, size_t workspace_bytes, mnr_stream_t stream);

/* Ride-along job for the NEXT mnr_fused_clip_adam / mnr_peer_reduce_clip_adam[_pp] launch issued by this host thread: extra CTAs
 * of that launch (the optimizer's own cover one float4 per thread and leave most SMs idle for the PPO towers) compute
 * y[r * y_ld + c] = clamp((x[r * cols + c] - mean[c]) / denom[c]) - exactly mnr_rms_apply_ld_f32 (RunningMeanStd.normalize,
 * mineral/common/normalizers.py:57-68) - for the following optimizer step's minibatch, so that launch leaves the step's
 * chain.  The job is consumed (and cleared) by that launch; x / y must not alias anything the optimizer step touches. */
int mnr_optim_queue_normalise(const float* x, float* y, int64_t rows, int cols, int y_ld, const float* mean_f32,
                              const float* denom_f32, float clamp_lo, float clamp_hi);

/* ------------------------------------------------------------------------------------
 * a16  SHAC / SAPO critic regression loss (SHAC.compute_critic_loss, mineral/agents/diffrl/shac.py:869-875):
 *      loss = mean_k mean_rows (pred_k - target)^2 and its gradient g_k = 2 (pred_k - target) / (rows * n_critics).
 *      preds / grads: up to MNR_MAX_CRITICS device pointers (rows) each.  loss_out: 1 float.  Deterministic.
 * ------------------------------------------------------------------------------------ */
#define MNR_MAX_CRITICS 4
typedef struct {
  const float* pred[MNR_MAX_CRITICS];
  float* grad[MNR_MAX_CRITICS];
} mnr_critic_set_t;
size_t mnr_critic_mse_workspace_bytes(int64_t rows);
int mnr_critic_mse_fwd_bwd(const mnr_critic_set_t* set /*host*/, int n_critics, const float* target, int64_t rows,
                           float* loss_out, void* workspace, size_t workspace_bytes, mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * a23  multi-GPU exchange over NVLink / NVSwitch PEER MEMORY (one process per GPU).  Replaces the
 *      DDP gradient all-reduce inside accelerator.backward and the metric gathers
 *      (mineral/agents/ppo/ppo.py:116-123,292-311) and, fused behind it, clip_grad_norm_ +
 *      optimizer.step + zero_grad (ppo.py:291-300).
 *
 *      Every rank owns ONE peer buffer: [signal pad (MNR_PEER_SIGNAL_BYTES, zero) | caller-defined
 *      regions].  mnr_peer_alloc is the one place this library allocates device memory (cudaMalloc +
 *      CUDA IPC handle, because torch's caching allocator cannot export a whole allocation);
 *      the 64-byte handle travels to the other ranks by any host channel (torch.distributed
 *      all_gather_object) and is mapped there with mnr_peer_open.
 * ------------------------------------------------------------------------------------ */
#define MNR_MAX_PEERS 8
#define MNR_PEER_HANDLE_BYTES 64
#define MNR_PEER_SIGNAL_BYTES 16384

typedef struct {
  void* base[MNR_MAX_PEERS]; /* rank r's peer buffer as mapped in THIS process; base[rank] is the local one */
  int32_t rank, world;
} mnr_peer_comm_t;

int mnr_peer_alloc(size_t bytes, void** ptr_out /*host*/, void* handle_out /*host, 64 B*/);
int mnr_peer_open(const void* handle /*host, 64 B*/, void** ptr_out /*host*/);
int mnr_peer_close(void* ptr);
int mnr_peer_free(void* ptr);
/* host read (synchronous) of the local error word: != 0 if a cross-GPU wait timed out (20 s) */
int mnr_peer_error(const mnr_peer_comm_t* comm /*host*/, int* error_out /*host*/);

/* All-gather of a small payload: src (bytes, local) is published through the staging region at
 * `stage_offset` of the local peer buffer; dst (world*bytes, local) receives every rank's payload in
 * rank order.  bytes % 16 == 0.  Same call sequence on every rank. */
int mnr_peer_allgather(const mnr_peer_comm_t* comm /*host*/, size_t stage_offset, const void* src, size_t bytes,
                       void* dst, mnr_stream_t stream);

/* ONE kernel per optimizer step: barrier -> sum over ranks (rank order: bit-identical on every
 * rank) of the n gradient floats at `grad_offset` of every peer buffer (+ n_tail metric floats right
 * behind them -> tail_out) -> global norm -> clip + Adam/AdamW on (param, exp_avg, exp_avg_sq) ->
 * barrier -> zero the local gradient.  `reduced` = n floats of local scratch.  The metric at
 * kl_index (or none if < 0) feeds st->kl_sum as in mnr_grad_norm_prepare. */
int mnr_peer_reduce_clip_adam(const mnr_peer_comm_t* comm /*host*/, size_t grad_offset, int64_t n, int n_tail,
                              float* reduced, float* param, float* exp_avg, float* exp_avg_sq, float max_norm,
                              float beta1, float beta2, float eps, float weight_decay, int decoupled, int zero_grad,
                              mnr_optim_state_t* st, float* tail_out, int kl_index, float* grad_norm_out,
                              void* workspace, size_t workspace_bytes, mnr_stream_t stream);

/* Same with PING-PONG gradient buffers: consecutive optimizer steps write their gradients into two regions of the peer
 * buffer in turn.  A rank that has passed this step's start barrier has left the previous step's kernel, so nobody still
 * reads the previous step's region: it is cleared here (`zero_offset`, 0 = none) and the second cross-GPU barrier of the
 * step is skipped (`end_barrier` = 0; it must be 1 - with zero_grad as wanted - on the last step of an update, whose
 * region the next update's first backward overwrites).  `trace` (may be NULL): 8 x 64 u64 of globaltimer stamps of CTA 0 -
 * kernel start, after the start barrier, after the peer reads, after the grid barrier, after Adam, end; slot = step % 64. */
int mnr_peer_reduce_clip_adam_pp(const mnr_peer_comm_t* comm /*host*/, size_t grad_offset, size_t zero_offset, int end_barrier,
                                 int64_t n, int n_tail, float* reduced, float* param, float* exp_avg, float* exp_avg_sq,
                                 float max_norm, float beta1, float beta2, float eps, float weight_decay, int decoupled,
                                 int zero_grad, mnr_optim_state_t* st, float* tail_out, int kl_index, float* grad_norm_out,
                                 unsigned long long* trace, void* workspace, size_t workspace_bytes, mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * 8(f)-1  rollout-side policy step (between two simulator steps).  mnr_policy_step: policy / value heads on the
 *      towers' last activations y3a / y3v (N,H) (weights as in mnr_ppo_heads_loss_fwd_bwd), Dist
 *      (mineral/nets/dist.py:31-47), actions = mu + sigma * eps (Normal.sample, eps (N,A) standard normal noise
 *      drawn by the caller), neglogp = -sum_j log_prob (ActorCritic.act, mineral/agents/ppo/models.py:117-133),
 *      values un-normalised with the prepared f32 (mean, denom) of the value RunningMeanStd
 *      (PPO.model_act, mineral/agents/ppo/ppo.py:125-131; NULL,NULL: left normalised; clamp_lo > clamp_hi: no clamp)
 *      - all written straight into the time-major storage rows of step t (replaces 5 update_data copies,
 *      ppo.py:146-150) plus the clamped actions the simulator consumes (ppo.py:151, may be NULL).
 *      mnr_post_step: rewards_t = scale * r (+ gamma * values_t * time_out, ppo.py:158-160), dones_t = dones
 *      (u8 / bool in, u8 out; ppo.py:154,161).  A <= 32, H <= 512.
 * ------------------------------------------------------------------------------------ */
int mnr_policy_step(const float* y3a, const float* y3v, const float* w_mu, const float* b_mu, const float* w_v,
                    const float* b_v, const float* w_sig, const float* b_sig, const float* eps,
                    const float* value_mean_f32, const float* value_denom_f32, float clamp_lo, float clamp_hi,
                    float* actions_t, float* neglogp_t, float* values_t, float* mu_t, float* sigma_t,
                    float* actions_clamped, int N, int A, int H, const mnr_ppo_loss_cfg_t* cfg /*host*/,
                    mnr_stream_t stream);
int mnr_post_step(const float* rewards, const void* dones, const void* time_outs, const float* values_t,
                  float reward_scale, float gamma, float* rewards_t, uint8_t* dones_t, int N, mnr_stream_t stream);

/* 8(f)-4  episode accounting on the device (Metrics._update_scores, mineral/common/metrics.py:59-67; Tracker,
 *      mineral/common/tracker.py:9-22): cur_rewards (N) f32 += rewards, cur_lengths (N) i64 += 1; every finished env
 *      (dones u8/bool, in env order) appends (reward, length) to the windows of the last `window` finished episodes
 *      (win_rewards / win_lengths (window) f32, ring slot = episode count mod window, zero-initialised like the
 *      reference's deque) and resets; counters[0] (i64) counts finished episodes.  One launch, no host read. */
int mnr_episode_scores(const float* rewards, const void* dones, float* cur_rewards, int64_t* cur_lengths,
                       float* win_rewards, float* win_lengths, int64_t* counters, int N, int window, mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * 8(f)-2  SAC / DDPG update math on top of the fused replay sample (bit-identical to the reference's CPU torch
 *      evaluation: every op rounded separately, in the reference's order).
 *      mnr_soft_update:      target = target * (1 - tau) + current * tau  over a flat buffer
 *                            (soft_update, mineral/agents/ddpg/utils.py:8-11); one_minus_tau = (float)(1.0 - tau).
 *      mnr_td_target:        out = reward + (1 - done) * gamma_n * (min_k q_k [- alpha * log_prob])  (rows)
 *                            (mineral/agents/sac/sac.py:293-297, ddpg.py:292-293); qs->pred[0..n_q) are the ensemble's
 *                            Q values, log_prob may be NULL (DDPG / backup_entropy off); gamma_n = (float)(gamma ** nstep).
 *      mnr_distl_projection: C51 categorical projection of next_dist (rows, n_atoms) onto the fixed support
 *                            (distl_projection, mineral/agents/ddpg/utils.py:14-31); delta_z = (float)((v_max - v_min) /
 *                            (n_atoms - 1)); n_atoms <= 128.
 * ------------------------------------------------------------------------------------ */
int mnr_soft_update(float* target, const float* current, int64_t n, float one_minus_tau, float tau, mnr_stream_t stream);
int mnr_td_target(const mnr_critic_set_t* qs /*host*/, int n_q, const float* log_prob, const float* reward, const float* done,
                  float alpha, float gamma_n, float* out, int64_t rows, mnr_stream_t stream);
int mnr_distl_projection(const float* next_dist, const float* reward, const float* done, const float* support, float gamma,
                         float v_min, float v_max, float delta_z, int n_atoms, float* out, int64_t rows, mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * 8(f)-2  categorical (C51) critic of DDPG: `DistributionalEnsembleQ` (mineral/agents/ddpg/models.py:117-170) inside
 *      DDPG.update_critic / update_actor (mineral/agents/ddpg/ddpg.py:266-310).  `logits->pred[k]`: (rows, n_atoms) outputs
 *      of critic k's last linear layer, `->grad[k]`: d loss / d logits written by the call; n_atoms <= 128, n_q <= 4.
 *      mnr_distl_target:      target (rows, n_atoms) = min_k distl_projection(softmax(target_logits_k), reward, done,
 *                             gamma_n = gamma^nstep, v_min, v_max, support)   (ddpg.py:275-290, ddpg/utils.py:14-31);
 *      mnr_distl_bce_fwd_bwd: *loss_out = sum_k F.binary_cross_entropy(softmax(logits_k), target) (mean reduction,
 *                             logs clamped at -100) and its gradient through the softmax   (ddpg.py:267-269,295-296);
 *      mnr_distl_actor_seed:  Q = min_k sum_a softmax(logits_k)_a * support_a (get_q_min, models.py:160-163; first
 *                             minimum on ties), *loss_out = -mean(Q) and d loss / d logits_k  (ddpg.py:306-307).
 *      workspace: mnr_distl_workspace_bytes() bytes, zero-initialised once, 8 B aligned.  Deterministic.
 * ------------------------------------------------------------------------------------ */
size_t mnr_distl_workspace_bytes(void);
int mnr_distl_target(const mnr_critic_set_t* target_logits /*host*/, int n_q, const float* reward, const float* done,
                     const float* support, float gamma_n, float v_min, float v_max, float delta_z, int n_atoms,
                     float* target_out, int64_t rows, mnr_stream_t stream);
int mnr_distl_bce_fwd_bwd(const mnr_critic_set_t* logits /*host*/, int n_q, const float* target, int64_t rows, int n_atoms,
                          float* loss_out, void* workspace, size_t workspace_bytes, mnr_stream_t stream);
int mnr_distl_actor_seed(const mnr_critic_set_t* logits /*host*/, int n_q, const float* support, int64_t rows, int n_atoms,
                         float* loss_out, void* workspace, size_t workspace_bytes, mnr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * 8(f)-2 (agents)  row-wise math of the SAC / DDPG agents around the tower GEMMs (mineral/agents/sac/sac.py,
 *      mineral/agents/ddpg/ddpg.py, mineral/agents/ddpg/models.py, mineral/nets/dist.py, mineral/nets/distributions.py).
 *      mnr_copy_cols_f32:    dst[r, 0:cols) = src[r, 0:cols) with row strides (the `torch.cat((state, action), 1)` of
 *                            EnsembleQ.forward, ddpg/models.py:102-106, written into a padded critic-input buffer).
 *      mnr_squashed_sample:  SAC.get_actions (sac.py:119-136): sigma from logstd per Dist.forward (dist.py:37-52;
 *                            logstd (rows,A), or (A) when logstd_broadcast), u = mu + sigma * eps, actions = tanh(u) written
 *                            with row stride act_ld, log_prob (rows) = sum_j [Normal.log_prob(u) - log|d tanh/du|]
 *                            (distributions.py:10-37).  eps NULL: actions = mu (the deterministic branch), log_prob untouched.
 *      mnr_squashed_bwd:     gradient of (alpha/rows) * sum log_prob + <g_actions, actions> w.r.t. (mu, logstd) through the
 *                            reparameterised sample; alpha = exp(*log_alpha) read on the device, or alpha_host when NULL.
 *      mnr_tanh_action(_bwd): DDPG / TD3 actor output a = tanh(mu) [+ clamp(noise)] [clamped] (ddpg/models.py:55-58,
 *                            noise.py:22-38; lo > hi disables a clamp) and d/dmu = g (1 - tanh^2).
 *      mnr_td_target_stats:  mnr_td_target with the temperature on the device + {mean, std (unbiased), max, min} of the
 *                            target (sac.py:306-311) into stats[4] (may be NULL).  One CTA, deterministic.
 *      mnr_actor_seed:       q_min over qs->pred[0..n_q) (first minimum on ties), d actor_loss / d Q_k = -1/rows at the
 *                            argmin into qs->grad[k]; out4 = {actor_loss = mean(alpha*log_prob - q_min), entropy =
 *                            mean(-log_prob), alpha_loss = alpha * mean(-log_prob - target_entropy), alpha};
 *                            *log_alpha_grad = d alpha_loss / d log_alpha (sac.py:313-332; log_prob NULL = DDPG, ddpg.py:299-310).
 *      mnr_critic_mse_sum_fwd_bwd: mnr_critic_mse_fwd_bwd with the SUM over critics (sac.py:303, ddpg.py:296).
 * ------------------------------------------------------------------------------------ */
typedef struct {
  int32_t dist_type;       /* 0 squashed_normal, 1 dreamerv3_squashed_normal */
  int32_t has_minlogstd, has_maxlogstd;
  float minlogstd, maxlogstd; /* squashed_normal: torch.clamp bounds on logstd */
  float minstd, maxstd;       /* dreamerv3_squashed_normal */
  int32_t _pad;
} mnr_squash_cfg_t;
int mnr_copy_cols_f32(const float* src, int src_ld, float* dst, int dst_ld, int64_t rows, int cols, mnr_stream_t stream);
int mnr_squashed_sample(const float* mu, const float* logstd, int logstd_broadcast, const float* eps, int64_t rows, int A,
                        const mnr_squash_cfg_t* cfg /*host*/, float* actions, int act_ld, float* log_prob, mnr_stream_t stream);
int mnr_squashed_bwd(const float* mu, const float* logstd, int logstd_broadcast, const float* eps, const float* g_actions,
                     const float* g_actions2 /* optional second addend, same layout (two critic towers) */, int g_ld,
                     const float* log_alpha, float alpha_host, int64_t rows, int A, const mnr_squash_cfg_t* cfg /*host*/, float* g_mu,
                     float* g_logstd, mnr_stream_t stream);
int mnr_tanh_action(const float* mu, const float* noise, int64_t rows, int A, float noise_lo, float noise_hi, float out_lo,
                    float out_hi, float* actions, int act_ld, mnr_stream_t stream);
int mnr_tanh_action_bwd(const float* mu, const float* g_actions, const float* g_actions2 /* optional */, int g_ld, int64_t rows, int A,
                        float* g_mu, mnr_stream_t stream);
int mnr_td_target_stats(const mnr_critic_set_t* qs /*host*/, int n_q, const float* log_prob, const float* reward, const float* done,
                        const float* log_alpha, float alpha_host, float gamma_n, float* out, int64_t rows, float* stats,
                        mnr_stream_t stream);
int mnr_actor_seed(const mnr_critic_set_t* qs /*host*/, int n_q, const float* log_prob, const float* log_alpha, float alpha_host,
                   float target_entropy, int64_t rows, float* out4, float* log_alpha_grad, mnr_stream_t stream);
/* One env step of SAC/DDPG.explore_env (sac.py:163-177, ddpg.py:176-190) into the (N, T, .) trajectory tensors at step t:
 * obs / next_obs (N,O), actions (N,A), rewards (N) * reward_scale (RewardShaper 'scale'), dones (u8/bool) masked by
 * time_outs (u8/bool, may be NULL; Agent._handle_timeout, agent.py:158-167) stored as f32 0/1. */
int mnr_traj_store_step(const float* obs, const float* next_obs, const float* actions, const float* rewards, const void* dones,
                        const void* time_outs, int N, int O, int A, int T, int t, float reward_scale, float* traj_obs,
                        float* traj_next_obs, float* traj_actions, float* traj_rewards, float* traj_dones, mnr_stream_t stream);
int mnr_critic_mse_sum_fwd_bwd(const mnr_critic_set_t* set /*host*/, int n_critics, const float* target, int64_t rows,
                               float* loss_out, void* workspace, size_t workspace_bytes, mnr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MINERAL_B200_H */
