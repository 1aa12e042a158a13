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
