// a23 / a12  multi-GPU exchange over NVLink 5 / NVSwitch PEER MEMORY, fused with the optimizer step.
//
// The env-sharded PPO update exchanges only sums over samples: the flat gradient
// and ten loss metrics once per optimizer step, a few KB of statistics once per rollout.  The reference
// hands these to DDP / accelerate (mineral/agents/ppo/ppo.py:116-123,292-311), i.e. NCCL launches with
// a host-visible dependency before clip_grad_norm_ + optimizer.step (ppo.py:291-300).  Here every rank
// maps every other rank's buffer (CUDA IPC over NVLink) and ONE kernel per optimizer step does
//
//   cross-GPU barrier -> P2P loads of all ranks' gradients, summed in rank order (bit-identical result
//   on every rank) -> global L2 norm (grid barrier) -> clip coefficient + Adam scalars -> Adam / AdamW
//   on the reduced gradient -> cross-GPU barrier -> zero the local gradient
//
// so a whole multi-GPU update stays one CUDA graph with no NCCL call inside.  With world == 1 the same
// kernel is the single-launch clip + Adam step.
//
// Barriers follow the flag protocol of one-shot custom all-reduces: block b of rank r owns the counters
// start[b][*] / end[b][*] in its own signal pad and writes its monotonically increasing flag into the
// peers' pads; block b of every rank touches the same elements, so per-block barriers suffice.
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kMaxBlocks = 160;           // >= SM count: one co-resident CTA per SM
constexpr int kMaxPeers = MNR_MAX_PEERS;  // 8
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000 * 1000 * 1000;  // a dead peer must not hang the GPU

// signal pad at the start of every rank's peer buffer (MNR_PEER_SIGNAL_BYTES, zero-initialised)
struct SignalPad {
  uint32_t start[kMaxBlocks][kMaxPeers];
  uint32_t end[kMaxBlocks][kMaxPeers];
  uint32_t flag[kMaxBlocks];  // local only: last flag value block b used
  uint32_t error;             // local only: set when a spin timed out
};
static_assert(sizeof(SignalPad) <= MNR_PEER_SIGNAL_BYTES, "signal pad too small");

struct Comm {
  unsigned char* base[kMaxPeers];
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {  // L1-bypassing load of peer (or local) memory
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_peer_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Cross-GPU barrier of block b with block b of every peer.  Threads 0..world-1 each handle one peer.
// Release / acquire at system scope: everything this CTA wrote before the barrier (and everything earlier
// kernels of this stream wrote) is visible to the peers' block b after it.
template <bool kEnd>
__device__ __forceinline__ void peer_barrier(const Comm& c, uint32_t flag) {
  __syncthreads();
  if ((int)threadIdx.x < c.world) {
    SignalPad* mine = reinterpret_cast<SignalPad*>(c.base[c.rank]);
    SignalPad* theirs = reinterpret_cast<SignalPad*>(c.base[threadIdx.x]);
    uint32_t* dst = kEnd ? &theirs->end[blockIdx.x][c.rank] : &theirs->start[blockIdx.x][c.rank];
    const uint32_t* src = kEnd ? &mine->end[blockIdx.x][threadIdx.x] : &mine->start[blockIdx.x][threadIdx.x];
    st_release_sys(dst, flag);
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(src) != flag) {
      if (globaltimer_ns() - t0 > kSpinTimeoutNs) {
        mine->error = 1u + (uint32_t)threadIdx.x;
        break;
      }
    }
  }
  __syncthreads();
}

constexpr size_t kWsHeader = 256;

// Optional table in the optimizer workspace (after the per-CTA partials; zero = absent): weight matrices whose tf32 (hi, lo)
// copies the tensor-core linear kernels read (csrc/tc_linear.cu).  The Adam phase rewrites them right where it writes the new
// parameter, so the update chain needs no separate split launch per optimizer step.
constexpr int kMaxSplit = 8;
struct SplitEntry {
  long long off;   // first element of the matrix in the flat parameter buffer (multiple of 4)
  int n, K, ld, pad;  // elements (rows * K), row length, row stride of hi / lo
  float* hi;
  float* lo;
};
struct SplitTable {
  int count, pad[3];
  SplitEntry e[kMaxSplit];
};
constexpr size_t kSplitTableOffset = kWsHeader + 160 * sizeof(double);

// Ride-along job of the optimizer launch (mnr_optim_queue_normalise): CTAs behind the optimizer's own (which fill only a third of
// the GPU for the PPO towers: one float4 per thread) normalise + clamp the NEXT optimizer step's input rows into the padded buffer
// the first tower layer reads - mnr_rms_apply_ld_f32's arithmetic, element for element - so that 7 us launch leaves the step.
struct AuxNormalise {
  const float* x;
  float* y;
  int64_t total;   // rows * cols; 0 = no job
  int cols, y_ld;
  const float* mean;
  const float* denom;
  float lo, hi;
  int blocks;
};

struct AdamArgs {
  AuxNormalise aux;
  int main_blocks;   // CTAs [0, main_blocks) run the optimizer step (and its grid barrier), the rest the ride-along job
  float* param;
  float* m;
  float* v;
  float* reduced;  // n floats of local scratch (the summed gradient)
  float* grad;     // world == 1: the gradient itself; world > 1: base[rank] + grad_offset
  int64_t n;       // multiple of 4
  size_t grad_offset;
  int n_tail;      // metrics appended right after the n gradient floats in every rank's buffer
  float* tail_out; // reduced metrics (may be NULL when world == 1)
  int kl_index;    // index of the KL metric in the tail, or -1
  const float* kl; // world == 1 alternative: pointer to the step's KL (may be NULL)
  float max_norm, beta1, beta2, eps, wd;
  int decoupled, zero_grad;
  int nan_to_num;          // world == 1: grad.nan_to_num_(0, 0, 0) before the norm (mineral/agents/diffrl/shac.py:841-843)
  mnr_optim_state_t* st;
  float* norm_out;
  float* clipped_norm_out; // norm * clip coefficient (the "after clip" norm SHAC logs), may be NULL
  double* parts;                // gridDim.x doubles
  unsigned long long* counter;  // grid barrier
  const SplitTable* split;      // in the workspace; count == 0: nothing to split
  // world > 1, ping-pong gradient buffers (mnr_peer_reduce_clip_adam_pp): the buffer of the PREVIOUS step is cleared here
  // (every peer passed this step's start barrier, so it has finished reading it) and the end barrier can be skipped
  float* zero_other;            // n floats (+ tail) of the local buffer used one step earlier, or NULL
  int end_barrier;              // 1: peers-done-reading barrier before this step's buffer is released (needed on the last
                                // step of an update, whose buffer the next update's first backward may overwrite)
  unsigned long long* trace;    // optional: 8 x 64 globaltimer stamps of CTA 0 (start, after start barrier, after the
                                // peer reads, after the grid barrier, after Adam, end), slot = step % 64
};

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float clip, float beta1, float beta2, float eps,
                                      float wd, int decoupled, float lr, float step_size, float bc2_sqrt) {
  g *= clip;
  if (wd != 0.f) {
    if (decoupled) p *= 1.0f - lr * wd;
    else g += wd * p;
  }
  m += (g - m) * (1.0f - beta1);
  v = v * beta2 + (1.0f - beta2) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p -= step_size * (m / denom);
}

__device__ __forceinline__ float4 finite4(float4 v) {
  v.x = isfinite(v.x) ? v.x : 0.f, v.y = isfinite(v.y) ? v.y : 0.f, v.z = isfinite(v.z) ? v.z : 0.f, v.w = isfinite(v.w) ? v.w : 0.f;
  return v;
}

template <int WORLD>  // 0 = runtime world (any), 1 = single GPU (no peers)
__global__ void __launch_bounds__(kThreads, 1) reduce_clip_adam_kernel(const __grid_constant__ Comm c, const __grid_constant__ AdamArgs a) {
  __shared__ double s_w[kThreads / 32];
  __shared__ float s_bcast[4];
  __shared__ float s_tail[32];
  __shared__ SplitTable s_split;
  mnr::pdl_trigger();   // PDL: see common.cuh
  mnr::pdl_wait();
  const int world = WORLD == 1 ? 1 : c.world;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= a.main_blocks) {   // ride-along CTAs: independent of everything below, no barrier
    const AuxNormalise& x = a.aux;
    const int64_t astride = (int64_t)x.blocks * kThreads;
    int64_t i = (int64_t)((int)blockIdx.x - a.main_blocks) * kThreads + tid;
    if (i >= x.total) return;
    // (row, column) of the first element by one division, then stepped: no 64-bit division per element
    int64_t r = i / x.cols;
    int cc = (int)(i - r * x.cols);
    const int64_t dr = astride / x.cols;
    const int dc = (int)(astride - dr * x.cols);
    for (; i < x.total; i += astride) {
      float y = __fdiv_rn(__fsub_rn(mnr::ld_stream(x.x + i), __ldg(x.mean + cc)), __ldg(x.denom + cc));   // normalizers.py:65-68
      if (x.lo <= x.hi) y = fminf(fmaxf(y, x.lo), x.hi);
      x.y[r * x.y_ld + cc] = y;
      r += dr, cc += dc;
      if (cc >= x.cols) cc -= x.cols, r += 1;
    }
    return;
  }
  for (int w = tid; w < (int)(sizeof(SplitTable) / 4); w += kThreads)   // visible after the barriers below
    reinterpret_cast<uint32_t*>(&s_split)[w] = reinterpret_cast<const uint32_t*>(a.split)[w];
  const int64_t n4 = a.n >> 2;
  const int64_t stride = (int64_t)a.main_blocks * kThreads;
  const int64_t first = (int64_t)blockIdx.x * kThreads + tid;

  // optimizer scalars of the PREVIOUS step, read before anyone (block 0, after the grid barrier) updates them
  const int step = a.st->step + 1;
  const double lr = a.st->lr;

  unsigned long long* tr = (WORLD != 1 && a.trace && blockIdx.x == 0 && tid == 0) ? a.trace + (size_t)((step - 1) & 63) * 8 : nullptr;
  if (tr) tr[0] = globaltimer_ns();
  uint32_t flag = 0;
  if (WORLD != 1) {
    SignalPad* mine = reinterpret_cast<SignalPad*>(c.base[c.rank]);
    // A spin that timed out in an EARLIER launch left the sticky error word set: the peers are gone (or diverged), so this
    // and every later step is a no-op - parameters, moments and the optimizer state stay as they were (PPO.train polls
    // PeerComm.error() once per epoch and raises).  Every CTA reads the same word, so all of them leave together.
    if (__ldcg(&mine->error) != 0u) return;
    flag = mine->flag[blockIdx.x] + 1u;
    peer_barrier<false>(c, flag);  // every rank's backward has finished writing its gradient
    if (tr) tr[1] = globaltimer_ns();
  }

  // ---- phase 1: sum over ranks (fixed order), park the result, accumulate the squared norm ----
  float acc = 0.f;
  float4* red4 = reinterpret_cast<float4*>(a.reduced);
  if (WORLD == 1) {
    const float4* g4 = reinterpret_cast<const float4*>(a.grad);
    for (int64_t i = first; i < n4; i += stride) {
      float4 s = g4[i];
      if (a.nan_to_num) s = finite4(s);
      acc += s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;
    }
  } else {
    for (int64_t i = first; i < n4; i += stride) {
      float4 v[kMaxPeers];
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < world) v[r] = ld_peer_f4(reinterpret_cast<const float4*>(c.base[r] + a.grad_offset) + i);
      float4 s = v[0];
#pragma unroll
      for (int r = 1; r < kMaxPeers; ++r)
        if (r < world) s.x += v[r].x, s.y += v[r].y, s.z += v[r].z, s.w += v[r].w;
      red4[i] = s;
      acc += s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;
    }
    if (blockIdx.x == 0 && tid < a.n_tail) {  // loss metrics (pre-scaled by 1/world in the loss kernel)
      float s = 0.f;
      for (int r = 0; r < world; ++r) s += ld_peer_f(reinterpret_cast<const float*>(c.base[r] + a.grad_offset) + a.n + tid);
      s_tail[tid] = s;
      if (a.tail_out) a.tail_out[tid] = s;
    }
  }
  if (tr) tr[2] = globaltimer_ns();
  const double wsum = mnr::warp_sum((double)acc);
  if ((tid & 31) == 0) s_w[tid >> 5] = wsum;
  __syncthreads();
  if (tid == 0) {
    double s = 0;
    for (int x = 0; x < kThreads / 32; ++x) s += s_w[x];
    a.parts[blockIdx.x] = s;
  }
  mnr::grid_barrier(a.counter, (unsigned int)a.main_blocks);
  if (tr) tr[3] = globaltimer_ns();

  // A barrier spin of THIS launch timed out in some CTA (a dead or diverged peer): the sums above hold stale or partial
  // peer data.  The error word was written before the grid barrier, so every CTA sees it here and skips phases 2-4:
  // no optimizer state, parameter or moment is touched and the local gradient is left in place.
  if (WORLD != 1) {
    if (tid == 0) s_bcast[0] = __ldcg(&reinterpret_cast<SignalPad*>(c.base[c.rank])->error) != 0u ? 1.f : 0.f;
    __syncthreads();
    const bool failed = s_bcast[0] != 0.f;
    __syncthreads();
    if (failed) return;
  }

  // ---- phase 2: every CTA derives the same clip coefficient and Adam scalars ----
  if (tid < 32) {
    double mine = 0;
    for (int x = tid; x < a.main_blocks; x += 32) mine += __ldcg(a.parts + x);
    // fixed-order tree over the 32 lanes: identical in every CTA (and on every rank)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if (tid == 0) {
      const float norm = (float)sqrt(mine);
      float coef = 1.0f;
      if (a.max_norm > 0.f) coef = fminf(a.max_norm / (norm + 1e-6f), 1.0f);  // torch clip_grad_norm_
      const double bc1 = 1.0 - pow((double)a.beta1, (double)step);
      const double bc2 = 1.0 - pow((double)a.beta2, (double)step);
      s_bcast[0] = coef, s_bcast[1] = (float)(lr / bc1), s_bcast[2] = (float)sqrt(bc2), s_bcast[3] = (float)lr;
      if (blockIdx.x == 0) {
        mnr_optim_state_t* st = a.st;
        st->grad_norm = norm, st->clip_coef = coef, st->step = step;
        st->step_size = s_bcast[1], st->bc2_sqrt = s_bcast[2], st->lr_f32 = s_bcast[3];
        if (WORLD == 1) {
          if (a.kl) st->kl_sum += (double)*a.kl, st->kl_count += 1;
        } else if (a.kl_index >= 0) {
          st->kl_sum += (double)s_tail[a.kl_index], st->kl_count += 1;
        }
        if (a.norm_out) *a.norm_out = norm;
        if (a.clipped_norm_out) *a.clipped_norm_out = norm * coef;
      }
    }
  }
  __syncthreads();
  const float clip = s_bcast[0], step_size = s_bcast[1], bc2s = s_bcast[2], lr32 = s_bcast[3];

  // ---- phase 3: Adam / AdamW on the elements this thread reduced ----
  float4* p4 = reinterpret_cast<float4*>(a.param);
  float4* m4 = reinterpret_cast<float4*>(a.m);
  float4* v4 = reinterpret_cast<float4*>(a.v);
  const float4* gsrc = WORLD == 1 ? reinterpret_cast<const float4*>(a.grad) : red4;
  for (int64_t i = first; i < n4; i += stride) {
    float4 P = p4[i], G = gsrc[i], M = m4[i], V = v4[i];
    if (WORLD == 1 && a.nan_to_num) G = finite4(G);
    adam1(P.x, G.x, M.x, V.x, clip, a.beta1, a.beta2, a.eps, a.wd, a.decoupled, lr32, step_size, bc2s);
    adam1(P.y, G.y, M.y, V.y, clip, a.beta1, a.beta2, a.eps, a.wd, a.decoupled, lr32, step_size, bc2s);
    adam1(P.z, G.z, M.z, V.z, clip, a.beta1, a.beta2, a.eps, a.wd, a.decoupled, lr32, step_size, bc2s);
    adam1(P.w, G.w, M.w, V.w, clip, a.beta1, a.beta2, a.eps, a.wd, a.decoupled, lr32, step_size, bc2s);
    p4[i] = P, m4[i] = M, v4[i] = V;
    // tf32 (hi, lo) copies of the tensor-core layers' weights (what mnr_split_tf32 would produce from the new parameters)
    const int n_split = s_split.count;
    if (n_split > 0) {
      const long long f = (long long)i * 4;
      for (int q = 0; q < n_split; ++q) {
        const SplitEntry& se = s_split.e[q];
        if (f < se.off || f >= se.off + se.n) continue;
        const int e0 = (int)(f - se.off);
        const float pv[4] = {P.x, P.y, P.z, P.w};
        if (se.ld == se.K) {   // dense rows: the four elements are contiguous in hi / lo too (off and K are multiples of 4)
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(P.x) & 0xffffe000u), h.y = __uint_as_float(__float_as_uint(P.y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(P.z) & 0xffffe000u), h.w = __uint_as_float(__float_as_uint(P.w) & 0xffffe000u);
          l.x = P.x - h.x, l.y = P.y - h.y, l.z = P.z - h.z, l.w = P.w - h.w;
          *reinterpret_cast<float4*>(se.hi + e0) = h;
          *reinterpret_cast<float4*>(se.lo + e0) = l;
        } else {               // padded rows (K not a multiple of 4): element-wise
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = e0 + j;
            if (e >= se.n) break;
            const int r = e / se.K, cc = e - r * se.K;
            const float h = __uint_as_float(__float_as_uint(pv[j]) & 0xffffe000u);
            se.hi[(size_t)r * se.ld + cc] = h;
            se.lo[(size_t)r * se.ld + cc] = pv[j] - h;
          }
        }
        break;
      }
    }
  }

  if (tr) tr[4] = globaltimer_ns();
  // ---- phase 4: peers are done reading my gradient -> it may be cleared / rewritten ----
  if (WORLD != 1) {
    if (a.end_barrier) peer_barrier<true>(c, flag);
    if (tid == 0) reinterpret_cast<SignalPad*>(c.base[c.rank])->flag[blockIdx.x] = flag;
  }
  if (a.zero_grad) {
    float4* g4 = reinterpret_cast<float4*>(a.grad);
    for (int64_t i = first; i < n4; i += stride) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (WORLD != 1 && a.zero_other) {   // the previous step's buffer: every peer left that step before this one's start barrier
    float4* g4 = reinterpret_cast<float4*>(a.zero_other);
    for (int64_t i = first; i < n4; i += stride) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tr) tr[5] = globaltimer_ns();
}

// all-gather of a small payload: block b moves slice b of every rank
__global__ void __launch_bounds__(256) peer_allgather_kernel(const __grid_constant__ Comm c, size_t stage_offset, const uint4* __restrict__ src,
                                                             int64_t n16, uint4* __restrict__ dst) {
  SignalPad* mine = reinterpret_cast<SignalPad*>(c.base[c.rank]);
  const uint32_t flag = mine->flag[blockIdx.x] + 1u;
  const int64_t per = (n16 + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x, hi = min(n16, lo + per);
  uint4* stage = reinterpret_cast<uint4*>(c.base[c.rank] + stage_offset);
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) stage[i] = src[i];
  peer_barrier<false>(c, flag);  // release: my slice is published; acquire: so are the peers'
  for (int r = 0; r < c.world; ++r) {
    const uint4* ps = reinterpret_cast<const uint4*>(c.base[r] + stage_offset);
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) dst[(int64_t)r * n16 + i] = ld_peer_u4(ps + i);
  }
  peer_barrier<true>(c, flag);   // nobody still reads my staging slice when the next call overwrites it
  if (threadIdx.x == 0) mine->flag[blockIdx.x] = flag;
}

int make_comm(const mnr_peer_comm_t* comm, Comm& c) {
  if (comm->world < 1 || comm->world > kMaxPeers || comm->rank < 0 || comm->rank >= comm->world) return MNR_ESHAPE;
  for (int r = 0; r < kMaxPeers; ++r) c.base[r] = nullptr;
  for (int r = 0; r < comm->world; ++r) {
    if (!comm->base[r]) return MNR_ENULL;
    if (!mnr_aligned(comm->base[r], 16)) return MNR_EALIGN;
    c.base[r] = static_cast<unsigned char*>(comm->base[r]);
  }
  c.rank = comm->rank, c.world = comm->world;
  return MNR_OK;
}

// the job queued by mnr_optim_queue_normalise for the next optimizer launch of this host thread
thread_local AuxNormalise t_aux_pending = {};

// grid = optimizer CTAs + ride-along CTAs; fills a.aux / a.main_blocks and clears the pending job
int adam_grid(int64_t n);
dim3 adam_launch_grid(AdamArgs& a, int64_t n) {
  a.main_blocks = adam_grid(n);
  a.aux = t_aux_pending;
  t_aux_pending = AuxNormalise{};
  int extra = 0;
  if (a.aux.total > 0) {
    int64_t want = (a.aux.total + (int64_t)kThreads * 8 - 1) / ((int64_t)kThreads * 8);   // >= 8 elements per thread
    const int cap = mnr_sm_count();
    extra = (int)(want > cap ? cap : want);
    if (extra < 1) extra = 1;
  }
  a.aux.blocks = extra;
  return dim3(a.main_blocks + extra);
}

int adam_grid(int64_t n) {
  int64_t g = (n / 4 + kThreads - 1) / kThreads;
  int cap = mnr_sm_count();  // one CTA per SM: the grid barrier needs every CTA resident
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

// ---- peer buffers (host) ---------------------------------------------------------------------------
extern "C" int mnr_peer_alloc(size_t bytes, void** ptr_out, void* handle_out) {
  MNR_CHECK_PTR(ptr_out); MNR_CHECK_PTR(handle_out);
  if (bytes < MNR_PEER_SIGNAL_BYTES) return MNR_ESHAPE;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == MNR_PEER_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return MNR_OK;
}

extern "C" int mnr_peer_open(const void* handle, void** ptr_out) {
  MNR_CHECK_PTR(handle); MNR_CHECK_PTR(ptr_out);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return (int)e;
  *ptr_out = p;
  return MNR_OK;
}

extern "C" int mnr_peer_close(void* ptr) {
  MNR_CHECK_PTR(ptr);
  const cudaError_t e = cudaIpcCloseMemHandle(ptr);
  return e == cudaSuccess ? MNR_OK : (int)e;
}

extern "C" int mnr_peer_free(void* ptr) {
  MNR_CHECK_PTR(ptr);
  const cudaError_t e = cudaFree(ptr);
  return e == cudaSuccess ? MNR_OK : (int)e;
}

extern "C" int mnr_peer_error(const mnr_peer_comm_t* comm, int* error_out) {
  MNR_CHECK_PTR(comm); MNR_CHECK_PTR(error_out);
  Comm c;
  const int rc = make_comm(comm, c);
  if (rc != MNR_OK) return rc;
  uint32_t v = 0;
  const cudaError_t e = cudaMemcpy(&v, c.base[c.rank] + offsetof(SignalPad, error), sizeof(v), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return (int)e;
  *error_out = (int)v;
  return MNR_OK;
}

extern "C" int mnr_peer_allgather(const mnr_peer_comm_t* comm, size_t stage_offset, const void* src, size_t bytes, void* dst,
                                  mnr_stream_t stream) {
  MNR_CHECK_PTR(comm); MNR_CHECK_PTR(src); MNR_CHECK_PTR(dst);
  if (bytes == 0 || bytes % 16 != 0 || stage_offset % 16 != 0 || stage_offset < MNR_PEER_SIGNAL_BYTES) return MNR_ESHAPE;
  if (!mnr_aligned(src, 16) || !mnr_aligned(dst, 16)) return MNR_EALIGN;
  Comm c;
  const int rc = make_comm(comm, c);
  if (rc != MNR_OK) return rc;
  const int64_t n16 = (int64_t)(bytes / 16);
  int grid = (int)((n16 + 1023) / 1024);
  if (grid > 32) grid = 32;
  if (grid < 1) grid = 1;
  peer_allgather_kernel<<<grid, 256, 0, mnr_cs(stream)>>>(c, stage_offset, (const uint4*)src, n16, (uint4*)dst);
  MNR_LAUNCH_RESULT();
}

// ---- fused optimizer step ------------------------------------------------------------------------------
extern "C" size_t mnr_fused_adam_workspace_bytes(void) { return kSplitTableOffset + sizeof(SplitTable); }
extern "C" size_t mnr_fused_adam_split_table_offset(void) { return kSplitTableOffset; }
static_assert(kSplitTableOffset == kWsHeader + (size_t)kMaxBlocks * sizeof(double), "table sits right after the partials");

namespace {
int fill_adam(AdamArgs& a, float* param, float* exp_avg, float* exp_avg_sq, int64_t n, float max_norm, float beta1, float beta2, float eps,
              float wd, int decoupled, int zero_grad, mnr_optim_state_t* st, float* norm_out, void* workspace, size_t workspace_bytes) {
  if (!param || !exp_avg || !exp_avg_sq || !st || !workspace) return MNR_ENULL;
  if (n <= 0 || n % 4 != 0) return MNR_ESHAPE;
  if (workspace_bytes < mnr_fused_adam_workspace_bytes()) return MNR_EWORKSPACE;
  if (!mnr_aligned(param, 16) || !mnr_aligned(exp_avg, 16) || !mnr_aligned(exp_avg_sq, 16) || !mnr_aligned(st, 8) ||
      !mnr_aligned(workspace, 16))
    return MNR_EALIGN;
  a.param = param, a.m = exp_avg, a.v = exp_avg_sq, a.n = n;
  a.max_norm = max_norm, a.beta1 = beta1, a.beta2 = beta2, a.eps = eps, a.wd = wd, a.decoupled = decoupled, a.zero_grad = zero_grad;
  a.st = st, a.norm_out = norm_out, a.nan_to_num = 0, a.clipped_norm_out = nullptr;
  a.counter = reinterpret_cast<unsigned long long*>(workspace);
  a.parts = reinterpret_cast<double*>(static_cast<unsigned char*>(workspace) + kWsHeader);
  a.split = reinterpret_cast<const SplitTable*>(static_cast<unsigned char*>(workspace) + kSplitTableOffset);
  a.reduced = nullptr, a.grad = nullptr, a.grad_offset = 0, a.n_tail = 0, a.tail_out = nullptr, a.kl_index = -1, a.kl = nullptr;
  a.zero_other = nullptr, a.end_barrier = 1, a.trace = nullptr;
  return MNR_OK;
}
}  // namespace

extern "C" int mnr_fused_clip_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float max_norm,
                                   float beta1, float beta2, float eps, float weight_decay, int decoupled, int zero_grad,
                                   mnr_optim_state_t* st, const float* kl, float* grad_norm_out, int nan_to_num,
                                   float* clipped_norm_out, void* workspace, size_t workspace_bytes, mnr_stream_t stream) {
  AdamArgs a;
  const int rc = fill_adam(a, param, exp_avg, exp_avg_sq, n, max_norm, beta1, beta2, eps, weight_decay, decoupled, zero_grad, st,
                           grad_norm_out, workspace, workspace_bytes);
  if (rc != MNR_OK) return rc;
  MNR_CHECK_PTR(grad);
  if (!mnr_aligned(grad, 16)) return MNR_EALIGN;
  a.grad = grad, a.kl = kl, a.nan_to_num = nan_to_num, a.clipped_norm_out = clipped_norm_out;
  Comm c;
  for (int r = 0; r < kMaxPeers; ++r) c.base[r] = nullptr;
  c.rank = 0, c.world = 1;
  const cudaError_t le = mnr_launch_pdl(reduce_clip_adam_kernel<1>, adam_launch_grid(a, n), dim3(kThreads), 0, mnr_cs(stream), c, a);
  if (le != cudaSuccess) return (int)le;
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_optim_queue_normalise(const float* x, float* y, int64_t rows, int cols, int y_ld, const float* mean_f32,
                                         const float* denom_f32, float clamp_lo, float clamp_hi) {
  MNR_CHECK_PTR(x); MNR_CHECK_PTR(y); MNR_CHECK_PTR(mean_f32); MNR_CHECK_PTR(denom_f32);
  if (rows < 0 || cols <= 0 || y_ld < cols) return MNR_ESHAPE;
  t_aux_pending = AuxNormalise{x, y, rows * cols, cols, y_ld, mean_f32, denom_f32, clamp_lo, clamp_hi, 0};
  return MNR_OK;
}

extern "C" int mnr_peer_reduce_clip_adam_pp(const mnr_peer_comm_t* comm, size_t grad_offset, size_t zero_offset, int end_barrier, int64_t n,
                                            int n_tail, float* reduced, float* param, float* exp_avg, float* exp_avg_sq, float max_norm,
                                            float beta1, float beta2, float eps, float weight_decay, int decoupled, int zero_grad,
                                            mnr_optim_state_t* st, float* tail_out, int kl_index, float* grad_norm_out,
                                            unsigned long long* trace, void* workspace, size_t workspace_bytes, mnr_stream_t stream) {
  MNR_CHECK_PTR(comm); MNR_CHECK_PTR(reduced);
  AdamArgs a;
  int rc = fill_adam(a, param, exp_avg, exp_avg_sq, n, max_norm, beta1, beta2, eps, weight_decay, decoupled, zero_grad, st,
                     grad_norm_out, workspace, workspace_bytes);
  if (rc != MNR_OK) return rc;
  Comm c;
  rc = make_comm(comm, c);
  if (rc != MNR_OK) return rc;
  if (grad_offset % 16 != 0 || grad_offset < MNR_PEER_SIGNAL_BYTES || n_tail < 0 || n_tail > 32 || kl_index >= n_tail) return MNR_ESHAPE;
  if (zero_offset != 0 && (zero_offset % 16 != 0 || zero_offset < MNR_PEER_SIGNAL_BYTES || zero_offset == grad_offset)) return MNR_ESHAPE;
  if (!end_barrier && zero_grad) return MNR_ESHAPE;   // this step's buffer may only be cleared once the peers have read it
  if (!mnr_aligned(reduced, 16) || (trace && !mnr_aligned(trace, 8))) return MNR_EALIGN;
  a.reduced = reduced, a.grad_offset = grad_offset, a.n_tail = n_tail, a.tail_out = tail_out, a.kl_index = kl_index;
  a.grad = reinterpret_cast<float*>(c.base[c.rank] + grad_offset);
  a.zero_other = zero_offset ? reinterpret_cast<float*>(c.base[c.rank] + zero_offset) : nullptr;
  a.end_barrier = end_barrier ? 1 : 0, a.trace = trace;
  const cudaError_t le = mnr_launch_pdl(reduce_clip_adam_kernel<0>, adam_launch_grid(a, n), dim3(kThreads), 0, mnr_cs(stream), c, a);
  if (le != cudaSuccess) return (int)le;
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_peer_reduce_clip_adam(const mnr_peer_comm_t* comm, size_t grad_offset, int64_t n, int n_tail, float* reduced,
                                         float* param, float* exp_avg, float* exp_avg_sq, float max_norm, float beta1, float beta2,
                                         float eps, float weight_decay, int decoupled, int zero_grad, mnr_optim_state_t* st,
                                         float* tail_out, int kl_index, float* grad_norm_out, void* workspace,
                                         size_t workspace_bytes, mnr_stream_t stream) {
  return mnr_peer_reduce_clip_adam_pp(comm, grad_offset, 0, 1, n, n_tail, reduced, param, exp_avg, exp_avg_sq, max_norm, beta1, beta2, eps,
                                      weight_decay, decoupled, zero_grad, st, tail_out, kl_index, grad_norm_out, nullptr, workspace,
                                      workspace_bytes, stream);
}
