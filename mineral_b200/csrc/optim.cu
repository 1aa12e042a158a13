// a12  optimizer step with device-resident scalars: global grad-norm clip, Adam / AdamW over a flat
// parameter buffer, and the KL-adaptive learning-rate rule.  Replaces nn.utils.clip_grad_norm_ +
// torch.optim.Adam(W).step + optim.zero_grad (mineral/agents/ppo/ppo.py:291-300) and
// AdaptiveScheduler.update + the avg_kl.item() host sync (ppo.py:335-345,
// mineral/agents/ppo/utils.py:14-20).  Nothing here reads a value on the host, so a whole PPO
// update can be enqueued (or graph-captured) without a single host round trip.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kHeader = 256;

__global__ void __launch_bounds__(kThreads) grad_norm_kernel(const float* __restrict__ g, int64_t n, float max_norm,
                                                             float beta1, float beta2, mnr_optim_state_t* __restrict__ st,
                                                             const float* __restrict__ kl, float* __restrict__ norm_out,
                                                             unsigned char* __restrict__ workspace) {
  __shared__ double s_w[kThreads / 32];
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  float acc = 0.f;
  const int64_t n4 = mnr_aligned_dev(g) ? (n >> 2) : 0;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const float v = __ldg(g + i);
    acc += v * v;
  }
  const double ws = mnr::warp_sum((double)acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = ws;
  __syncthreads();
  double* parts = reinterpret_cast<double*>(workspace + kHeader);
  if (threadIdx.x == 0) {
    double s = 0;
    for (int x = 0; x < kThreads / 32; ++x) s += s_w[x];
    parts[blockIdx.x] = s;
  }
  if (!mnr::last_block_ticket(reinterpret_cast<unsigned int*>(workspace), gridDim.x)) return;
  double mine = 0;
  for (unsigned int x = threadIdx.x; x < gridDim.x; x += kThreads) mine += parts[x];
  const double s = mnr::block_sum(mine);
  if (threadIdx.x != 0) return;
  const float norm = (float)sqrt(s);
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to <= 1
  float coef = 1.0f;
  if (max_norm > 0.f) coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
  st->grad_norm = norm;
  st->clip_coef = coef;
  const int step = st->step + 1;
  st->step = step;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  st->step_size = (float)(st->lr / bc1);
  st->bc2_sqrt = (float)sqrt(bc2);
  st->lr_f32 = (float)st->lr;
  if (kl) {
    st->kl_sum += (double)*kl;
    st->kl_count += 1;
  }
  if (norm_out) *norm_out = norm;
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float clip, float beta1, float beta2, float eps,
                                      float wd, int decoupled, float lr, float step_size, float bc2_sqrt) {
  g *= clip;
  if (wd != 0.f) {
    if (decoupled) p *= 1.0f - lr * wd;  // AdamW: param.mul_(1 - lr * weight_decay)
    else g += wd * p;                    // Adam: grad.add(param, alpha=weight_decay)
  }
  m += (g - m) * (1.0f - beta1);           // exp_avg.lerp_(grad, 1 - beta1)
  v = v * beta2 + (1.0f - beta2) * g * g;  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(kThreads) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, float beta1, float beta2,
                                                        float eps, float wd, int decoupled, int zero_grad,
                                                        const mnr_optim_state_t* __restrict__ st) {
  const float clip = st->clip_coef, lr = st->lr_f32, step_size = st->step_size, bc2 = st->bc2_sqrt;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const bool vec = mnr_aligned_dev(p) && mnr_aligned_dev(g) && mnr_aligned_dev(m) && mnr_aligned_dev(v);
  const int64_t n4 = vec ? (n >> 2) : 0;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    adam1(P.x, G.x, M.x, V.x, clip, beta1, beta2, eps, wd, decoupled, lr, step_size, bc2);
    adam1(P.y, G.y, M.y, V.y, clip, beta1, beta2, eps, wd, decoupled, lr, step_size, bc2);
    adam1(P.z, G.z, M.z, V.z, clip, beta1, beta2, eps, wd, decoupled, lr, step_size, bc2);
    adam1(P.w, G.w, M.w, V.w, clip, beta1, beta2, eps, wd, decoupled, lr, step_size, bc2);
    reinterpret_cast<float4*>(p)[i] = P, reinterpret_cast<float4*>(m)[i] = M, reinterpret_cast<float4*>(v)[i] = V;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    float P = p[i], M = m[i], V = v[i];
    adam1(P, g[i], M, V, clip, beta1, beta2, eps, wd, decoupled, lr, step_size, bc2);
    p[i] = P, m[i] = M, v[i] = V;
    if (zero_grad) g[i] = 0.f;
  }
}

__global__ void kl_lr_kernel(mnr_optim_state_t* st, double thr, double min_lr, double max_lr, float* avg_kl_out,
                             double* lr_out) {
  // avg_kl = torch.mean(torch.stack(ep_kls)) is an f32 tensor; .item() widens it (ppo.py:335-338)
  const float avg = st->kl_count > 0 ? (float)(st->kl_sum / (double)st->kl_count) : 0.f;
  const double kl = (double)avg;
  const double cur = st->lr;
  double lr = cur;
  if (kl > 2.0 * thr) lr = fmax(cur / 1.5, min_lr);  // utils.py:16-17
  if (kl < 0.5 * thr) lr = fmin(cur * 1.5, max_lr);  // utils.py:18-19
  st->lr = lr;
  st->lr_f32 = (float)lr;
  st->kl_sum = 0.0;
  st->kl_count = 0;
  if (avg_kl_out) *avg_kl_out = avg;
  if (lr_out) *lr_out = lr;
}

int flat_grid(int64_t n) {
  int64_t g = (n / 4 + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)mnr_sm_count() * 4;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" size_t mnr_grad_norm_workspace_bytes(int64_t n) {
  (void)n;
  return kHeader + (size_t)mnr_sm_count() * 4 * sizeof(double);
}

extern "C" int mnr_grad_norm_prepare(const float* grad, int64_t n, float max_norm, float beta1, float beta2,
                                     mnr_optim_state_t* st, const float* kl, float* grad_norm_out, void* workspace,
                                     size_t workspace_bytes, mnr_stream_t stream) {
  MNR_CHECK_PTR(grad); MNR_CHECK_PTR(st); MNR_CHECK_PTR(workspace);
  if (n <= 0) return MNR_ESHAPE;
  if (workspace_bytes < mnr_grad_norm_workspace_bytes(n)) return MNR_EWORKSPACE;
  if (!mnr_aligned(st, 8) || !mnr_aligned(workspace, 8)) return MNR_EALIGN;
  grad_norm_kernel<<<flat_grid(n), kThreads, 0, mnr_cs(stream)>>>(grad, n, max_norm, beta1, beta2, st, kl, grad_norm_out,
                                                                  (unsigned char*)workspace);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float beta1,
                             float beta2, float eps, float weight_decay, int decoupled, int zero_grad,
                             const mnr_optim_state_t* st, mnr_stream_t stream) {
  MNR_CHECK_PTR(param); MNR_CHECK_PTR(grad); MNR_CHECK_PTR(exp_avg); MNR_CHECK_PTR(exp_avg_sq); MNR_CHECK_PTR(st);
  if (n <= 0) return MNR_ESHAPE;
  adam_kernel<<<flat_grid(n), kThreads, 0, mnr_cs(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                             weight_decay, decoupled, zero_grad, st);
  MNR_LAUNCH_RESULT();
}

// 'cos' / 'linear' schedules (ppo.py:339-348): the learning rate that takes effect after this mini-epoch is a function of
// host-known counters only (mini-epoch index, agent_steps), so the host uploads one value per mini-epoch before the update
// and the device installs it here - the update stays one CUDA graph with no host read.
__global__ void lr_table_kernel(mnr_optim_state_t* st, const double* lr_new, float* avg_kl_out, double* lr_out) {
  const float avg = st->kl_count > 0 ? (float)(st->kl_sum / (double)st->kl_count) : 0.f;
  const double lr = *lr_new;
  st->lr = lr;
  st->lr_f32 = (float)lr;
  st->kl_sum = 0.0;
  st->kl_count = 0;
  if (avg_kl_out) *avg_kl_out = avg;
  if (lr_out) *lr_out = lr;
}

extern "C" int mnr_lr_table_update(mnr_optim_state_t* st, const double* lr_new, float* avg_kl_out, double* lr_out,
                                   mnr_stream_t stream) {
  MNR_CHECK_PTR(st); MNR_CHECK_PTR(lr_new);
  lr_table_kernel<<<1, 1, 0, mnr_cs(stream)>>>(st, lr_new, avg_kl_out, lr_out);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_kl_lr_update(mnr_optim_state_t* st, double kl_threshold, double min_lr, double max_lr, float* avg_kl_out,
                                double* lr_out, mnr_stream_t stream) {
  MNR_CHECK_PTR(st);
  kl_lr_kernel<<<1, 1, 0, mnr_cs(stream)>>>(st, kl_threshold, min_lr, max_lr, avg_kl_out, lr_out);
  MNR_LAUNCH_RESULT();
}
