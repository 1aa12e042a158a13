// K3  indexed row movement: minibatch gather / scatter, circular replay insert and uniform-sample
// gather (with fused done->f32 and obs normalisation), n-step window push.
// Replaces ExperienceBuffer.__getitem__/update_mu_sigma (mineral/agents/ppo/experience.py:61-87),
// CriticDataset.shuffle (mineral/agents/diffrl/utils.py:55-58), ReplayBuffer.add_to_buffer /
// sample_batch (mineral/buffers/simple_replay.py:32-85) and NStepReplay.add_to_buffer /
// compute_nstep_return (mineral/buffers/nstep_replay.py:29-94).
//
// All fields of a row (obs, action, reward, next_obs, done ...) move in ONE launch: the fields are
// laid side by side as a virtual row of W = sum(width) elements and consecutive threads take
// consecutive elements, so within each field's run a warp touches consecutive addresses (every
// fetched sector fully used except at run edges; rows themselves are random: 148 B Ant rows span
// 5-6 sectors).  Index math is integer-exact.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct RowFields {
  mnr_row_field_t f[MNR_MAX_ROW_FIELDS];
  int start[MNR_MAX_ROW_FIELDS + 1];  // prefix sums of widths
  int n;
};

enum Mode { kGather = 0, kScatter = 1, kRing = 2 };

__device__ __forceinline__ float clamp_nan(float y, float lo, float hi) {
  return (lo <= hi && y == y) ? fminf(fmaxf(y, lo), hi) : y;
}

__device__ __forceinline__ void move_elem(const mnr_row_field_t& fd, int64_t src_row, int64_t dst_row, int e) {
  const int64_t si = src_row * fd.width + e, di = dst_row * (fd.dst_ld > 0 ? fd.dst_ld : fd.width) + e;
  switch (fd.kind) {
    case MNR_ROW_F32:
      static_cast<float*>(fd.dst)[di] = __ldg(static_cast<const float*>(fd.src) + si);
      break;
    case MNR_ROW_U8:
      static_cast<uint8_t*>(fd.dst)[di] = __ldg(static_cast<const uint8_t*>(fd.src) + si);
      break;
    case MNR_ROW_BOOL_TO_F32:
      static_cast<float*>(fd.dst)[di] = __ldg(static_cast<const uint8_t*>(fd.src) + si) ? 1.0f : 0.0f;
      break;
    case MNR_ROW_F32_TO_BOOL:
      static_cast<uint8_t*>(fd.dst)[di] = __ldg(static_cast<const float*>(fd.src) + si) != 0.0f ? 1 : 0;
      break;
    case MNR_ROW_F32_NORMALIZE: {
      const float x = __ldg(static_cast<const float*>(fd.src) + si);
      const float y = __fdiv_rn(__fsub_rn(x, __ldg(fd.mean + e)), __ldg(fd.denom + e));
      static_cast<float*>(fd.dst)[di] = clamp_nan(y, fd.clamp_lo, fd.clamp_hi);
      break;
    }
  }
}

// ring: row i of the incoming chunk -> slot; returns -1 when the slot is later overwritten by the
// tail part of the same insert (simple_replay.py:42-60: head first, then tail)
__device__ __forceinline__ int64_t ring_slot(int64_t i, int64_t n, int64_t next_p, int64_t C) {
  const int64_t end = next_p + n;
  if (end <= C) return next_p + i;
  const int64_t head = C - next_p, p = end - C;
  if (i < head) {
    const int64_t d = next_p + i;
    return d < p ? -1 : d;
  }
  return i - head;  // the LAST p rows land on [0, p)
}

template <int kMode>
__global__ void __launch_bounds__(kThreads) row_move_kernel(RowFields rf, const int64_t* __restrict__ idx, int64_t row0,
                                                            int64_t n_rows, int64_t next_p, int64_t capacity) {
  const int W = rf.start[rf.n];
  const int64_t total = n_rows * W;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < total; q += stride) {
    const int64_t i = q / W;
    const int e = (int)(q - i * W);
    int k = 0;
#pragma unroll
    for (int j = 1; j < MNR_MAX_ROW_FIELDS; ++j) k += (j < rf.n && e >= rf.start[j]) ? 1 : 0;
    const int64_t other = idx ? __ldg(idx + i) : row0 + i;
    int64_t src_row, dst_row;
    if (kMode == kGather) src_row = other, dst_row = i;
    else if (kMode == kScatter) src_row = i, dst_row = other;
    else {
      src_row = i, dst_row = ring_slot(i, n_rows, next_p, capacity);
      if (dst_row < 0) continue;
    }
    move_elem(rf.f[k], src_row, dst_row, e - rf.start[k]);
  }
}

int pack_fields(const mnr_row_field_t* fields, int n_fields, RowFields* out) {
  if (n_fields <= 0 || n_fields > MNR_MAX_ROW_FIELDS) return MNR_ESHAPE;
  if (!fields) return MNR_ENULL;
  out->n = n_fields;
  out->start[0] = 0;
  for (int k = 0; k < n_fields; ++k) {
    if (!fields[k].src || !fields[k].dst) return MNR_ENULL;
    if (fields[k].width <= 0 || (fields[k].dst_ld != 0 && fields[k].dst_ld < fields[k].width)) return MNR_ESHAPE;
    if (fields[k].kind < 0 || fields[k].kind > MNR_ROW_F32_NORMALIZE) return MNR_EUNSUP;
    if (fields[k].kind == MNR_ROW_F32_NORMALIZE && (!fields[k].mean || !fields[k].denom)) return MNR_ENULL;
    out->f[k] = fields[k];
    out->start[k + 1] = out->start[k] + fields[k].width;
  }
  for (int k = n_fields; k < MNR_MAX_ROW_FIELDS; ++k) out->start[k + 1] = out->start[n_fields];
  return MNR_OK;
}

int grid_for(int64_t total) {
  int64_t g = (total + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)mnr_sm_count() * 8;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

// ---- n-step window ---------------------------------------------------------------------------
struct NStepArgs {
  mnr_nstep_in_t in, win_in;
  mnr_nstep_out_t win_out, out;
};

// transition at call-relative step gs (gs < 0: previous window, oldest first)
__device__ __forceinline__ int64_t tr_index(int e, int gs, int S, int n, bool* from_win) {
  *from_win = gs < 0;
  return gs < 0 ? (int64_t)e * n + (n + gs) : (int64_t)e * S + gs;
}

__device__ __forceinline__ uint8_t read_done(const mnr_nstep_in_t& src, int64_t t) {
  return src.done_is_f32 ? (static_cast<const float*>(src.done)[t] != 0.0f) : (static_cast<const uint8_t*>(src.done)[t] != 0);
}

__device__ __forceinline__ void write_done(const mnr_nstep_out_t& dst, int64_t t, uint8_t d) {
  if (dst.done_is_f32) static_cast<float*>(dst.done)[t] = d ? 1.0f : 0.0f;
  else static_cast<uint8_t*>(dst.done)[t] = d;
}

__global__ void __launch_bounds__(kThreads) nstep_kernel(NStepArgs a, const float* __restrict__ gammas, int N, int S, int O,
                                                         int A, int n, int skip) {
  const int W = 2 * O + A + 2;
  const int64_t out_rows = (int64_t)(S - skip) * N;
  const int64_t win_rows = (int64_t)N * n;
  const int64_t total = (out_rows + win_rows) * W;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < total; q += stride) {
    const int64_t row = q / W;
    const int el = (int)(q - row * W);
    bool fw;
    if (row >= out_rows) {  // window after the call: slot j of env e <- step S - n + j
      const int64_t wr = row - out_rows;
      const int e = (int)(wr / n), j = (int)(wr - (int64_t)e * n);
      const int64_t t = tr_index(e, S - n + j, S, n, &fw);
      const mnr_nstep_in_t& src = fw ? a.win_in : a.in;
      if (el < O) a.win_out.obs[wr * O + el] = src.obs[t * O + el];
      else if (el < O + A) a.win_out.act[wr * A + (el - O)] = src.act[t * A + (el - O)];
      else if (el == O + A) a.win_out.rew[wr] = src.rew[t];
      else if (el < 2 * O + A + 1) a.win_out.nobs[wr * O + (el - O - A - 1)] = src.nobs[t * O + (el - O - A - 1)];
      else write_done(a.win_out, wr, read_done(src, t));
      continue;
    }
    const int s = skip + (int)(row / N);
    const int e = (int)(row - (int64_t)(s - skip) * N);
    const int g0 = s - (n - 1);  // oldest step of the window ending at s
    if (el < O + A) {            // oldest obs / action (nstep_replay.py:44-46)
      const int64_t t = tr_index(e, g0, S, n, &fw);
      const mnr_nstep_in_t& src = fw ? a.win_in : a.in;
      if (el < O) a.out.obs[row * O + el] = src.obs[t * O + el];
      else a.out.act[row * A + (el - O)] = src.act[t * A + (el - O)];
      continue;
    }
    // first done inside the window (argmax of the bool row, nstep_replay.py:80)
    int first = -1;
    for (int k = 0; k < n; ++k) {
      const int64_t t = tr_index(e, g0 + k, S, n, &fw);
      const uint8_t d = read_done(fw ? a.win_in : a.in, t);
      if (d && first < 0) first = k;
    }
    if (el == O + A) {  // masked discounted sum, sequential in k (nstep_replay.py:91-93)
      float acc = 0.f;
      for (int k = 0; k < n; ++k) {
        const int64_t t = tr_index(e, g0 + k, S, n, &fw);
        const float r = fw ? a.win_in.rew[t] : a.in.rew[t];
        const float m = (first < 0 || k <= first) ? 1.0f : 0.0f;
        const float term = __fmul_rn(__fmul_rn(r, __ldg(gammas + k)), m);
        acc = (k == 0) ? term : __fadd_rn(acc, term);
      }
      a.out.rew[row] = acc;
    } else if (el < 2 * O + A + 1) {  // next_obs at the first done, else the newest (:85-87)
      const int64_t t = tr_index(e, g0 + (first < 0 ? n - 1 : first), S, n, &fw);
      const int f = el - O - A - 1;
      a.out.nobs[row * O + f] = fw ? a.win_in.nobs[t * O + f] : a.in.nobs[t * O + f];
    } else {
      write_done(a.out, row, first >= 0 ? 1 : 0);  // newest done OR any done (:82-83)
    }
  }
}

}  // namespace

extern "C" int mnr_gather_rows(const mnr_row_field_t* fields, int n_fields, const int64_t* idx, int64_t row0,
                               int64_t n_idx, mnr_stream_t stream) {
  RowFields rf;
  const int rc = pack_fields(fields, n_fields, &rf);
  if (rc != MNR_OK) return rc;
  if (n_idx < 0) return MNR_ESHAPE;
  if (n_idx == 0) return MNR_OK;
  row_move_kernel<kGather><<<grid_for(n_idx * rf.start[rf.n]), kThreads, 0, mnr_cs(stream)>>>(rf, idx, row0, n_idx, 0, 0);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_scatter_rows(const mnr_row_field_t* fields, int n_fields, const int64_t* idx, int64_t row0,
                                int64_t n_idx, mnr_stream_t stream) {
  RowFields rf;
  const int rc = pack_fields(fields, n_fields, &rf);
  if (rc != MNR_OK) return rc;
  if (n_idx < 0) return MNR_ESHAPE;
  if (n_idx == 0) return MNR_OK;
  row_move_kernel<kScatter><<<grid_for(n_idx * rf.start[rf.n]), kThreads, 0, mnr_cs(stream)>>>(rf, idx, row0, n_idx, 0, 0);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_replay_insert(const mnr_row_field_t* fields, int n_fields, int64_t n_rows, int64_t next_p,
                                 int64_t capacity, mnr_stream_t stream) {
  RowFields rf;
  const int rc = pack_fields(fields, n_fields, &rf);
  if (rc != MNR_OK) return rc;
  if (n_rows < 0 || capacity <= 0 || next_p < 0 || next_p > capacity) return MNR_ESHAPE;
  if (next_p + n_rows - capacity > capacity) return MNR_ESHAPE;  // the reference's tail slice would not fit either
  if (n_rows == 0) return MNR_OK;
  row_move_kernel<kRing><<<grid_for(n_rows * rf.start[rf.n]), kThreads, 0, mnr_cs(stream)>>>(rf, nullptr, 0, n_rows,
                                                                                           next_p, capacity);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_nstep_push(const mnr_nstep_in_t* in, const mnr_nstep_in_t* win_in, const mnr_nstep_out_t* win_out,
                              const mnr_nstep_out_t* out, const float* gammas, int N, int S, int O, int A, int nstep,
                              int64_t count_before, mnr_stream_t stream) {
  MNR_CHECK_PTR(in); MNR_CHECK_PTR(win_in); MNR_CHECK_PTR(win_out); MNR_CHECK_PTR(out); MNR_CHECK_PTR(gammas);
  if (N <= 0 || S <= 0 || O <= 0 || A <= 0 || nstep < 2 || count_before < 0) return MNR_ESHAPE;
  NStepArgs a{*in, *win_in, *win_out, *out};
  if (!a.in.obs || !a.in.act || !a.in.rew || !a.in.nobs || !a.in.done) return MNR_ENULL;
  if (!a.win_in.obs || !a.win_in.act || !a.win_in.rew || !a.win_in.nobs || !a.win_in.done) return MNR_ENULL;
  if (!a.win_out.obs || !a.win_out.act || !a.win_out.rew || !a.win_out.nobs || !a.win_out.done) return MNR_ENULL;
  int64_t skip = (int64_t)nstep - 1 - count_before;
  if (skip < 0) skip = 0;
  if (skip > S) skip = S;
  if (skip < S && (!a.out.obs || !a.out.act || !a.out.rew || !a.out.nobs || !a.out.done)) return MNR_ENULL;
  const int64_t total = ((int64_t)(S - skip) * N + (int64_t)N * nstep) * (2 * O + A + 2);
  nstep_kernel<<<grid_for(total), kThreads, 0, mnr_cs(stream)>>>(a, gammas, N, S, O, A, nstep, (int)skip);
  MNR_LAUNCH_RESULT();
}
