// 8(f)-2  categorical (C51) critic of DDPG: `DistributionalEnsembleQ` (mineral/agents/ddpg/models.py:117-170) on the update path of
// mineral/agents/ddpg/ddpg.py:266-310.  The critic towers end in an n_atoms-wide linear layer (tower GEMM kernels); what is row-wise
// around them lives here, one thread per batch row, the row's n_atoms (<= 128) values in local arrays:
//
//   mnr_distl_target        target = min_k distl_projection(softmax(target_logits_k), reward, done, gamma^n)    ddpg.py:275-290
//   mnr_distl_bce_fwd_bwd   loss = sum_k mean(BCE(softmax(logits_k), target)) and d loss / d logits_k           ddpg.py:295-297
//   mnr_distl_actor_seed    Q = min_k sum_a softmax(logits_k)_a z_a, actor_loss = -mean(Q), d loss / d logits    models.py:160-163,
//                                                                                                               ddpg.py:299-310
// The projection repeats mnr_distl_projection's arithmetic (csrc/offpolicy.cu: the reference's operation order, separately
// rounded).  Loss sums: per-thread f64, fixed-order block reduce, per-block partials, last-block ticket -> deterministic.
#include "common.cuh"

namespace {

constexpr int kMaxAtoms = 128;
constexpr int kThreadsD = 128;

struct LogitSet {
  const float* x[MNR_MAX_CRITICS];
  float* g[MNR_MAX_CRITICS];
};

// softmax of one row (torch.softmax(dim=1): subtract the row maximum, exp, normalise)
__device__ __forceinline__ void softmax_row(const float* __restrict__ x, int n, float* p) {
  float m = -INFINITY;
  for (int j = 0; j < n; ++j) {
    p[j] = __ldg(x + j);
    m = fmaxf(m, p[j]);
  }
  float s = 0.f;
  for (int j = 0; j < n; ++j) {
    p[j] = expf(p[j] - m);
    s += p[j];
  }
  const float inv = 1.0f / s;
  for (int j = 0; j < n; ++j) p[j] *= inv;
}

__global__ void __launch_bounds__(kThreadsD) distl_target_kernel(LogitSet set, int n_q, const float* __restrict__ reward,
                                                                 const float* __restrict__ done, const float* __restrict__ support, float gamma,
                                                                 float v_min, float v_max, float delta_z, int n, float* __restrict__ out,
                                                                 int64_t rows) {
  const int64_t row = (int64_t)blockIdx.x * kThreadsD + threadIdx.x;
  if (row >= rows) return;
  float p[kMaxAtoms], proj[kMaxAtoms], best[kMaxAtoms];
  const float r = __ldg(reward + row);
  const float g = __fmul_rn(__fsub_rn(1.0f, __ldg(done + row)), gamma);  // (1 - done) * gamma
  for (int k = 0; k < n_q; ++k) {
    softmax_row(set.x[k] + row * n, n, p);
    for (int j = 0; j < n; ++j) proj[j] = 0.f;
    for (int pass = 0; pass < 2; ++pass) {   // the two index_add_ passes of distl_projection (ddpg/utils.py:28-30)
      for (int j = 0; j < n; ++j) {
        float tz = __fadd_rn(r, __fmul_rn(g, __ldg(support + j)));
        tz = fminf(fmaxf(tz, v_min), v_max);
        const float b = __fdiv_rn(__fsub_rn(tz, v_min), delta_z);
        int l = (int)floorf(b), u = (int)ceilf(b);
        if (u > 0 && l == u) l -= 1;
        if (l < n - 1 && l == u) u += 1;
        if (pass == 0) proj[l] = __fadd_rn(proj[l], __fmul_rn(p[j], __fsub_rn((float)u, b)));
        else proj[u] = __fadd_rn(proj[u], __fmul_rn(p[j], __fsub_rn(b, (float)l)));
      }
    }
    for (int j = 0; j < n; ++j) best[j] = k == 0 ? proj[j] : fminf(best[j], proj[j]);   // torch.min(torch.stack(...), dim=0)
  }
  for (int j = 0; j < n; ++j) out[row * n + j] = best[j];
}

__device__ __forceinline__ void finish_sum(double acc, unsigned char* workspace, float* out, double scale) {
  double* parts = reinterpret_cast<double*>(workspace + 256);
  const double tot = mnr::block_sum(acc);
  if (threadIdx.x == 0) parts[blockIdx.x] = tot;
  if (!mnr::last_block_ticket(reinterpret_cast<unsigned int*>(workspace), gridDim.x)) return;
  double mine = 0.0;
  for (unsigned int x = threadIdx.x; x < gridDim.x; x += kThreadsD) mine += parts[x];
  const double all = mnr::block_sum(mine);
  if (threadIdx.x == 0) *out = (float)(all * scale);
}

__global__ void __launch_bounds__(kThreadsD) distl_bce_kernel(LogitSet set, int n_q, const float* __restrict__ target, int64_t rows, int n,
                                                              float* __restrict__ loss_out, unsigned char* __restrict__ workspace) {
  // F.binary_cross_entropy (mean): -(t log p + (1 - t) log(1 - p)), logs clamped at -100; its backward
  // (p - t) / max((1 - p) p, 1e-12) / numel; then the softmax backward p_i (g_i - sum_j g_j p_j)
  const float inv_numel = 1.0f / ((float)rows * (float)n);
  double acc = 0.0;
  float p[kMaxAtoms], gp[kMaxAtoms];
  for (int64_t row = (int64_t)blockIdx.x * kThreadsD + threadIdx.x; row < rows; row += (int64_t)gridDim.x * kThreadsD) {
    const float* t = target + row * n;
    for (int k = 0; k < n_q; ++k) {
      softmax_row(set.x[k] + row * n, n, p);
      float s = 0.f, dot = 0.f;
      for (int j = 0; j < n; ++j) {
        const float tj = __ldg(t + j), pj = p[j];
        const float lp = fmaxf(logf(pj), -100.f), l1p = fmaxf(log1pf(-pj), -100.f);
        s -= tj * lp + (1.f - tj) * l1p;
        gp[j] = (pj - tj) / fmaxf((1.f - pj) * pj, 1e-12f) * inv_numel;
        dot = fmaf(gp[j], pj, dot);
      }
      acc += (double)s;
      float* g = set.g[k] + row * n;
      for (int j = 0; j < n; ++j) g[j] = p[j] * (gp[j] - dot);
    }
  }
  finish_sum(acc, workspace, loss_out, 1.0 / ((double)rows * (double)n));
}

__global__ void __launch_bounds__(kThreadsD) distl_actor_seed_kernel(LogitSet set, int n_q, const float* __restrict__ support, int64_t rows, int n,
                                                                     float* __restrict__ out, unsigned char* __restrict__ workspace) {
  const float w = -1.0f / (float)rows;   // actor_loss = -mean(Q)
  double acc = 0.0;
  float p[kMaxAtoms];
  for (int64_t row = (int64_t)blockIdx.x * kThreadsD + threadIdx.x; row < rows; row += (int64_t)gridDim.x * kThreadsD) {
    float qbest = 0.f;
    int kbest = 0;
    for (int k = 0; k < n_q; ++k) {   // Q_k = sum_a p_a z_a; first minimum wins ties (torch.min over the stacked critics)
      softmax_row(set.x[k] + row * n, n, p);
      float q = 0.f;
      for (int j = 0; j < n; ++j) q = fmaf(p[j], __ldg(support + j), q);
      if (k == 0 || q < qbest) qbest = q, kbest = k;
    }
    acc += (double)qbest;
    for (int k = 0; k < n_q; ++k) {
      float* g = set.g[k] + row * n;
      if (k != kbest) {
        for (int j = 0; j < n; ++j) g[j] = 0.f;
        continue;
      }
      softmax_row(set.x[k] + row * n, n, p);
      for (int j = 0; j < n; ++j) g[j] = w * p[j] * (__ldg(support + j) - qbest);   // d(-Q/B) / d logit_j
    }
  }
  finish_sum(acc, workspace, out, -1.0 / (double)rows);
}

int pack(const mnr_critic_set_t* set, int n_q, bool need_grad, LogitSet* s) {
  if (!set) return MNR_ENULL;
  if (n_q <= 0 || n_q > MNR_MAX_CRITICS) return MNR_ESHAPE;
  for (int k = 0; k < MNR_MAX_CRITICS; ++k) s->x[k] = k < n_q ? set->pred[k] : nullptr, s->g[k] = k < n_q ? set->grad[k] : nullptr;
  for (int k = 0; k < n_q; ++k)
    if (!s->x[k] || (need_grad && !s->g[k])) return MNR_ENULL;
  return MNR_OK;
}

int rows_grid(int64_t rows) {
  int64_t g = (rows + kThreadsD - 1) / kThreadsD;
  const int64_t cap = (int64_t)mnr_sm_count() * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

extern "C" size_t mnr_distl_workspace_bytes(void) { return 256 + (size_t)mnr_sm_count() * 8 * sizeof(double); }

extern "C" int mnr_distl_target(const mnr_critic_set_t* target_logits, int n_q, const float* reward, const float* done, const float* support,
                                float gamma_n, float v_min, float v_max, float delta_z, int n_atoms, float* target_out, int64_t rows,
                                mnr_stream_t stream) {
  MNR_CHECK_PTR(reward); MNR_CHECK_PTR(done); MNR_CHECK_PTR(support); MNR_CHECK_PTR(target_out);
  if (rows <= 0 || n_atoms < 2 || n_atoms > kMaxAtoms) return MNR_ESHAPE;
  LogitSet s;
  const int rc = pack(target_logits, n_q, false, &s);
  if (rc != MNR_OK) return rc;
  distl_target_kernel<<<(unsigned)((rows + kThreadsD - 1) / kThreadsD), kThreadsD, 0, mnr_cs(stream)>>>(s, n_q, reward, done, support, gamma_n, v_min,
                                                                                                      v_max, delta_z, n_atoms, target_out, rows);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_distl_bce_fwd_bwd(const mnr_critic_set_t* set, int n_q, const float* target, int64_t rows, int n_atoms, float* loss_out,
                                     void* workspace, size_t workspace_bytes, mnr_stream_t stream) {
  MNR_CHECK_PTR(target); MNR_CHECK_PTR(loss_out); MNR_CHECK_PTR(workspace);
  if (rows <= 0 || n_atoms < 2 || n_atoms > kMaxAtoms) return MNR_ESHAPE;
  if (workspace_bytes < mnr_distl_workspace_bytes()) return MNR_EWORKSPACE;
  if (!mnr_aligned(workspace, 8)) return MNR_EALIGN;
  LogitSet s;
  const int rc = pack(set, n_q, true, &s);
  if (rc != MNR_OK) return rc;
  distl_bce_kernel<<<rows_grid(rows), kThreadsD, 0, mnr_cs(stream)>>>(s, n_q, target, rows, n_atoms, loss_out, (unsigned char*)workspace);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_distl_actor_seed(const mnr_critic_set_t* set, int n_q, const float* support, int64_t rows, int n_atoms, float* loss_out,
                                    void* workspace, size_t workspace_bytes, mnr_stream_t stream) {
  MNR_CHECK_PTR(support); MNR_CHECK_PTR(loss_out); MNR_CHECK_PTR(workspace);
  if (rows <= 0 || n_atoms < 2 || n_atoms > kMaxAtoms) return MNR_ESHAPE;
  if (workspace_bytes < mnr_distl_workspace_bytes()) return MNR_EWORKSPACE;
  if (!mnr_aligned(workspace, 8)) return MNR_EALIGN;
  LogitSet s;
  const int rc = pack(set, n_q, true, &s);
  if (rc != MNR_OK) return rc;
  distl_actor_seed_kernel<<<rows_grid(rows), kThreadsD, 0, mnr_cs(stream)>>>(s, n_q, support, rows, n_atoms, loss_out, (unsigned char*)workspace);
  MNR_LAUNCH_RESULT();
}
