// G  actor/critic MLP towers: hand-written fused linear-layer kernels (forward, input-gradient,
// weight-gradient) replacing nn.Linear + activation + autograd (mineral/nets/mlp.py:25-77,
// mineral/agents/ppo/models.py:136-151) on the PPO update path.
//
// One tiled fp32 kernel, three modes, written as C[i][j] = sum_r A(i,r) * B(r,j):
//   FWD  Y[m][n]  = act( sum_k X[m][k] W[n][k] + b[n] )                  i=m  j=n  r=k
//   DX   dX[m][k] = sum_n (dY[m][n] * act'(S[m][n])) W[n][k]   (+= opt.)  i=m  j=k  r=n
//   DW   dW[n][k] = sum_m (dY[m][n] * act'(S[m][n])) X[m][k],            i=n  j=k  r=m (split over CTAs)
//        db[n]    = sum_m  dY[m][n] * act'(S[m][n])
// S is the saved forward output Y (ELU / ReLU / identity: act' is a function of y) so no separate
// activation-backward pass and no dZ tensor ever touches HBM; the bias gradient rides in the DW kernel.
// Both towers (actor, value) go in one launch as "groups" (blockIdx.z).
//
// Exact fp32 FMA arithmetic (this is what the 1e-5 parity bar is checked with); 128x64x16 tiles,
// 256 threads, 8x4 outputs per thread, register-prefetched double buffering.  A tcgen05 3xTF32 core
// for the two big layers is the follow-up.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, kThreads = 256, TM = 8, TN = 4;
constexpr int kMaxGroups = 4;

enum Mode { kFwd = 0, kDx = 1, kDw = 2 };

struct Group {
  const float* a;     // FWD: X (M,K)      DX: dY (M,N)        DW: dY (M,N)
  const float* b;     // FWD: W (N,K)      DX: W (N,K)         DW: X (M,K)
  float* c;           // FWD: Y (M,N)      DX: dX (M,K)        DW: dW (N,K)
  const float* bias;  // FWD: b (N) or null
  const float* s;     // DX/DW: saved forward output for act' (M,N) or null (identity)
  float* bias_grad;   // DW: db (N) or null
  int N, K;           // layer width (out features), input features
  int lda, ldb;       // row strides of X (as `a`) and of the `b` operand
};

struct Groups {
  Group g[kMaxGroups];
  int n;
};

__device__ __forceinline__ float act_fwd(float z, int act) {
  switch (act) {
    case 1: return z > 0.f ? z : expm1f(z);  // ELU(alpha=1)
    case 2: return fmaxf(z, 0.f);            // ReLU
    default: return z;
  }
}
// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float act_bwd_from_y(float y, int act) {
  switch (act) {
    case 1: return y > 0.f ? 1.f : y + 1.f;  // ELU: z<=0 -> exp(z) = y + 1
    case 2: return y > 0.f ? 1.f : 0.f;
    default: return 1.f;
  }
}

template <int kMode>
__global__ void __launch_bounds__(kThreads, 2) linear_kernel(const __grid_constant__ Groups gs, int M, int act,
                                                            int accumulate, int splits, float* __restrict__ partials) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int gi = blockIdx.z / splits, split = blockIdx.z - gi * splits;
  const Group g = gs.g[gi];
  // problem extents in (i, j, r)
  const int I = kMode == kDw ? g.N : M;
  const int J = kMode == kFwd ? g.N : g.K;
  const int R = kMode == kFwd ? g.K : (kMode == kDx ? g.N : M);
  const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
  if (i0 >= I || j0 >= J) return;
  // reduction range of this CTA (DW: split over rows)
  int r_begin = 0, r_end = R;
  if (kMode == kDw) {
    const int per = ((R + splits - 1) / splits + BK - 1) / BK * BK;
    r_begin = split * per;
    r_end = min(R, r_begin + per);
  }
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;

  // element fetchers ---------------------------------------------------------------------------------
  auto fetch_a = [&](int i, int r) -> float {
    if (i >= I || r >= r_end) return 0.f;
    if (kMode == kFwd) return __ldg(g.a + (size_t)i * g.lda + r);          // X[m][k]
    size_t off;
    if (kMode == kDx) off = (size_t)i * g.N + r;                           // dY[m][n], i=m r=n
    else off = (size_t)r * g.N + i;                                        // dY[m][n], i=n r=m
    float v = __ldg(g.a + off);
    if (g.s) v *= act_bwd_from_y(__ldg(g.s + off), act);
    return v;
  };
  auto fetch_b = [&](int r, int j) -> float {
    if (j >= J || r >= r_end) return 0.f;
    if (kMode == kFwd) return __ldg(g.b + (size_t)j * g.ldb + r);          // W[n][k], j=n r=k
    return __ldg(g.b + (size_t)r * g.ldb + j);                               // DX: W[n][k] r=n j=k ; DW: X[m][k] r=m j=k
  };
  // thread -> tile element maps: consecutive threads walk the contiguous global dimension
  constexpr bool kAContigR = (kMode != kDw);   // FWD: k contiguous, DX: n contiguous, DW: i(=n) contiguous
  constexpr bool kBContigR = (kMode == kFwd);  // FWD: k contiguous, else j(=k) contiguous
  float ra[8], rb[4];
  auto load_tiles = [&](int r0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = tid + e * kThreads;
      const int r = kAContigR ? (idx & (BK - 1)) : (idx >> 7);
      const int i = kAContigR ? (idx >> 4) : (idx & (BM - 1));
      ra[e] = fetch_a(i0 + i, r0 + r);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * kThreads;
      const int r = kBContigR ? (idx & (BK - 1)) : (idx >> 6);
      const int j = kBContigR ? (idx >> 4) : (idx & (BN - 1));
      rb[e] = fetch_b(r0 + r, j0 + j);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = tid + e * kThreads;
      const int r = kAContigR ? (idx & (BK - 1)) : (idx >> 7);
      const int i = kAContigR ? (idx >> 4) : (idx & (BM - 1));
      As[buf][r][i] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * kThreads;
      const int r = kBContigR ? (idx & (BK - 1)) : (idx >> 6);
      const int j = kBContigR ? (idx >> 4) : (idx & (BN - 1));
      Bs[buf][r][j] = rb[e];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
  float bsum[TM];  // DW: bias gradient partial (only the tx == 0 column of threads of j-tile 0 uses it)
#pragma unroll
  for (int a = 0; a < TM; ++a) bsum[a] = 0.f;
  const bool do_bias = (kMode == kDw) && g.bias_grad != nullptr && blockIdx.y == 0 && tx == 0;

  int buf = 0;
  if (r_begin < r_end) {
    load_tiles(r_begin);
    store_tiles(0);
  }
  __syncthreads();
  for (int r0 = r_begin; r0 < r_end; r0 += BK) {
    const bool more = r0 + BK < r_end;
    if (more) load_tiles(r0 + BK);  // global loads in flight while the FMAs run
#pragma unroll
    for (int r = 0; r < BK; ++r) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][r][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][r][ty * TM + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][r][tx * TN]);
      const float av[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[TN] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int a = 0; a < TM; ++a) {
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        if (do_bias) bsum[a] += av[a];
      }
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // epilogue -----------------------------------------------------------------------------------------
#pragma unroll
  for (int a = 0; a < TM; ++a) {
    const int i = i0 + ty * TM + a;
    if (i >= I) continue;
#pragma unroll
    for (int b = 0; b < TN; ++b) {
      const int j = j0 + tx * TN + b;
      if (j >= J) continue;
      float v = acc[a][b];
      if (kMode == kFwd) {
        if (g.bias) v += __ldg(g.bias + j);
        g.c[(size_t)i * g.N + j] = act_fwd(v, act);
      } else if (kMode == kDx) {
        float* dst = g.c + (size_t)i * g.K + j;
        *dst = accumulate ? *dst + v : v;
      } else {
        if (splits == 1) g.c[(size_t)i * g.K + j] = v;
        else partials[(((size_t)gi * splits + split) * g.N + i) * g.K + j] = v;
      }
    }
    if (do_bias) {
      if (splits == 1) g.bias_grad[i] = bsum[a];
      else partials[(size_t)gs.n * splits * g.N * g.K + ((size_t)gi * splits + split) * g.N + i] = bsum[a];
    }
  }
}

// dW / db = sum over splits of the partials (fixed order => deterministic)
__global__ void __launch_bounds__(256) reduce_splits_kernel(const __grid_constant__ Groups gs, int splits,
                                                            const float* __restrict__ partials) {
  const int gi = blockIdx.y;
  const Group g = gs.g[gi];
  const int NK = g.N * g.K;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  // four independent accumulation chains, 16 loads in flight (the thin path uses hundreds of splits: a single chain of
  // dependent adds behind L2 latency took 23 us); fixed order => deterministic
  if (idx < NK) {
    const float* p = partials + (size_t)gi * splits * NK + idx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int z = 0;
#pragma unroll 4
    for (; z + 3 < splits; z += 4) {
      s0 += __ldg(p + (size_t)z * NK), s1 += __ldg(p + (size_t)(z + 1) * NK);
      s2 += __ldg(p + (size_t)(z + 2) * NK), s3 += __ldg(p + (size_t)(z + 3) * NK);
    }
    for (; z < splits; ++z) s0 += __ldg(p + (size_t)z * NK);
    g.c[idx] = (s0 + s1) + (s2 + s3);
  } else if (idx < NK + g.N && g.bias_grad) {
    const int n = idx - NK;
    const float* bp = partials + (size_t)gs.n * splits * NK + (size_t)gi * splits * g.N + n;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int z = 0;
#pragma unroll 4
    for (; z + 3 < splits; z += 4) {
      s0 += __ldg(bp + (size_t)z * g.N), s1 += __ldg(bp + (size_t)(z + 1) * g.N);
      s2 += __ldg(bp + (size_t)(z + 2) * g.N), s3 += __ldg(bp + (size_t)(z + 3) * g.N);
    }
    for (; z < splits; ++z) s0 += __ldg(bp + (size_t)z * g.N);
    g.bias_grad[n] = (s0 + s1) + (s2 + s3);
  }
}

// column sums of a (M, C) matrix (gradient of a broadcast parameter, e.g. the fixed sigma vector)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t M, int C, float* __restrict__ out,
                                                     float* __restrict__ partials, unsigned int* __restrict__ ticket) {
  __shared__ float sm[256];
  const int tid = threadIdx.x;
  const int rpar = 256 / C, r_local = tid / C, c = tid - r_local * C;
  float acc = 0.f;
  if (r_local < rpar)
    for (int64_t r = (int64_t)blockIdx.x * rpar + r_local; r < M; r += (int64_t)gridDim.x * rpar) acc += __ldg(x + r * C + c);
  sm[tid] = (r_local < rpar) ? acc : 0.f;
  __syncthreads();
  if (tid < C) {
    float s = 0.f;
    for (int q = 0; q < rpar; ++q) s += sm[q * C + tid];
    partials[(size_t)blockIdx.x * C + tid] = s;
  }
  if (!mnr::last_block_ticket(ticket, gridDim.x)) return;
  if (tid < C) {
    double s = 0;
    for (unsigned int b = 0; b < gridDim.x; ++b) s += (double)partials[(size_t)b * C + tid];
    out[tid] = (float)s;
  }
}

// ---- thin layers: N <= 32 output features (policy / value / Q heads, mineral/agents/ppo/models.py:136-151,
// mineral/agents/ddpg/models.py:29-71,84-111) ----------------------------------------------------------------------------
// The 128x64 tile kernel above leaves 7/8 of a tile idle on an 8-wide head and spends its time in the tile prologue; these
// are row-streaming kernels instead (a head moves M*K*4 bytes and does 2*M*K*N flops: HBM / latency-bound): W (padded to NP
// rows of zeros) sits in shared memory, a warp owns a row at a time, lanes own float4 chunks of the K dimension.
// Identity activation only (heads are plain Linear layers); anything else stays on the tile kernel.
constexpr int kThinThreads = 256, kThinWarps = kThinThreads / 32;

__device__ __forceinline__ float dot4f(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc), acc = fmaf(a.y, b.y, acc), acc = fmaf(a.z, b.z, acc), acc = fmaf(a.w, b.w, acc);
  return acc;
}

template <int NP>
__device__ __forceinline__ void thin_stage_w(const Group& g, float4* s_w) {
  const int K4 = g.K >> 2;
  for (int i = threadIdx.x; i < NP * K4; i += kThinThreads)
    s_w[i] = i < g.N * K4 ? __ldg(reinterpret_cast<const float4*>(g.b) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
}

// Y[m][n] = sum_k X[m][k] W[n][k] + b[n]
template <int NP>
__global__ void __launch_bounds__(kThinThreads) thin_fwd_kernel(const __grid_constant__ Groups gs, int M) {
  extern __shared__ __align__(16) float4 s_w4[];  // [NP][K/4]
  const Group g = gs.g[blockIdx.y];
  thin_stage_w<NP>(g, s_w4);
  const int K4 = g.K >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float bias = (g.bias && lane < g.N) ? __ldg(g.bias + lane) : 0.f;
  for (int m = blockIdx.x * kThinWarps + warp; m < M; m += gridDim.x * kThinWarps) {
    float acc[NP];
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[n] = 0.f;
    const float4* row = reinterpret_cast<const float4*>(g.a + (size_t)m * g.lda);
    for (int c = lane; c < K4; c += 32) {
      const float4 y = __ldg(row + c);
#pragma unroll
      for (int n = 0; n < NP; ++n) acc[n] = dot4f(y, s_w4[n * K4 + c], acc[n]);
    }
    float mine = 0.f;
#pragma unroll
    for (int n = 0; n < NP; ++n) {
      float v = acc[n];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == n) mine = v;
    }
    if (lane < g.N) g.c[(size_t)m * g.N + lane] = mine + bias;
  }
}

// dX[m][k] (+)= sum_n dY[m][n] W[n][k]
template <int NP>
__global__ void __launch_bounds__(kThinThreads) thin_dx_kernel(const __grid_constant__ Groups gs, int M, int accumulate) {
  extern __shared__ __align__(16) float4 s_w4[];
  const Group g = gs.g[blockIdx.y];
  thin_stage_w<NP>(g, s_w4);
  const int K4 = g.K >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = blockIdx.x * kThinWarps + warp; m < M; m += gridDim.x * kThinWarps) {
    const float gl = lane < g.N ? __ldg(g.a + (size_t)m * g.N + lane) : 0.f;
    float4* dst = reinterpret_cast<float4*>(g.c + (size_t)m * g.K);
    for (int c0 = 0; c0 < K4; c0 += 32) {   // uniform trip count: the shuffles below need the whole warp
      const int c = c0 + lane;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        const float gn = __shfl_sync(0xffffffffu, gl, n);
        if (c < K4) {
          const float4 w = s_w4[n * K4 + c];
          acc.x = fmaf(gn, w.x, acc.x), acc.y = fmaf(gn, w.y, acc.y), acc.z = fmaf(gn, w.z, acc.z), acc.w = fmaf(gn, w.w, acc.w);
        }
      }
      if (c < K4) {
        if (accumulate) {
          const float4 o = dst[c];
          acc.x += o.x, acc.y += o.y, acc.z += o.z, acc.w += o.w;
        }
        dst[c] = acc;
      }
    }
  }
}

// dW[n][k] = sum_m dY[m][n] X[m][k], db[n] = sum_m dY[m][n]: CTA (split, group) walks its slab of rows; thread = (float4 chunk
// of K, row phase).  Output: the final tensors when splits == 1, else split partials in linear_kernel<kDw>'s layout
// (summed in fixed order by reduce_splits_kernel).
template <int NP>
__global__ void __launch_bounds__(kThinThreads) thin_dw_kernel(const __grid_constant__ Groups gs, int M, int splits, float* __restrict__ partials) {
  __shared__ __align__(16) float4 s_red[kThinThreads];
  __shared__ float s_b[kThinThreads];
  const int gi = blockIdx.y, split = blockIdx.x;
  const Group g = gs.g[gi];
  const int N = g.N, K = g.K, K4 = K >> 2;
  const int phases = kThinThreads / K4, chunk = threadIdx.x % K4, phase = threadIdx.x / K4;
  const bool active = phase < phases;
  const int per = ((M + splits - 1) / splits + BK - 1) / BK * BK;
  const int r0 = split * per, r1 = min(M, r0 + per);
  float4 acc[NP];
  float bacc[NP];
#pragma unroll
  for (int n = 0; n < NP; ++n) acc[n] = make_float4(0.f, 0.f, 0.f, 0.f), bacc[n] = 0.f;
  if (active)
    for (int m = r0 + phase; m < r1; m += phases) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(g.b + (size_t)m * g.ldb) + chunk);
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        const float gn = n < N ? __ldg(g.a + (size_t)m * N + n) : 0.f;
        acc[n].x = fmaf(gn, x.x, acc[n].x), acc[n].y = fmaf(gn, x.y, acc[n].y), acc[n].z = fmaf(gn, x.z, acc[n].z), acc[n].w = fmaf(gn, x.w, acc[n].w);
        bacc[n] += gn;   // every chunk of a phase sees the same dY values; chunk 0's copy is the one that is kept
      }
    }
  float* w_out = splits == 1 ? g.c : partials + ((size_t)gi * splits + split) * N * K;
  float* b_out = splits == 1 ? g.bias_grad : partials + (size_t)gs.n * splits * N * K + ((size_t)gi * splits + split) * N;
#pragma unroll
  for (int n = 0; n < NP; ++n) {
    if (n >= N) break;
    __syncthreads();
    s_red[threadIdx.x] = acc[n];
    s_b[threadIdx.x] = bacc[n];
    __syncthreads();
    if (threadIdx.x < K4) {
      float4 t = s_red[threadIdx.x];
      for (int q = 1; q < phases; ++q) {
        const float4 o = s_red[q * K4 + threadIdx.x];
        t.x += o.x, t.y += o.y, t.z += o.z, t.w += o.w;
      }
      reinterpret_cast<float4*>(w_out + (size_t)n * K)[threadIdx.x] = t;
    }
    if (threadIdx.x == 0 && (splits > 1 || g.bias_grad)) {
      float t = 0.f;
      for (int q = 0; q < phases; ++q) t += s_b[q * K4];
      b_out[n] = t;
    }
  }
}

int thin_np(int max_n) { return max_n <= 1 ? 1 : (max_n <= 4 ? 4 : (max_n <= 8 ? 8 : (max_n <= 16 ? 16 : 32))); }

// all groups: identity activation, N <= n_cap, K a multiple of 4 with 16 B-aligned rows; `b` (W) dense (N,K)
bool thin_ok(const Groups& gs, int mode, int act, int n_cap) {
  for (int k = 0; k < gs.n; ++k) {
    const Group& g = gs.g[k];
    if (g.N > n_cap || g.K % 4 != 0 || g.K > 1024 || g.s != nullptr) return false;
    if (mode == kDw) {   // a = dY (M,N) scalar loads, b = X rows (ldb), c = dW (N,K)
      if (g.ldb % 4 != 0 || !mnr_aligned(g.b, 16) || !mnr_aligned(g.c, 16) || g.K / 4 > kThinThreads) return false;
    } else {
      if (g.ldb != g.K || !mnr_aligned(g.b, 16)) return false;
      if (mode == kFwd && (g.lda % 4 != 0 || !mnr_aligned(g.a, 16))) return false;
      if (mode == kDx && !mnr_aligned(g.c, 16)) return false;
    }
  }
  (void)act;
  return true;
}

int thin_grid(int M) {
  const int want = (M + kThinWarps - 1) / kThinWarps, cap = mnr_sm_count() * 8;
  return want < cap ? (want < 1 ? 1 : want) : cap;
}

#define MNR_THIN_DISPATCH(NPV, CALL)                 \
  switch (NPV) {                                      \
    case 1: { constexpr int NP = 1; CALL; } break;    \
    case 4: { constexpr int NP = 4; CALL; } break;    \
    case 8: { constexpr int NP = 8; CALL; } break;    \
    case 16: { constexpr int NP = 16; CALL; } break;  \
    default: { constexpr int NP = 32; CALL; } break;  \
  }

int pack_groups(const mnr_linear_group_t* groups, int n_groups, int mode, Groups* out) {
  if (n_groups <= 0 || n_groups > kMaxGroups) return MNR_ESHAPE;
  if (!groups) return MNR_ENULL;
  out->n = n_groups;
  for (int k = 0; k < n_groups; ++k) {
    const mnr_linear_group_t& s = groups[k];
    if (!s.a || !s.b || !s.c) return MNR_ENULL;
    if (s.N <= 0 || s.K <= 0) return MNR_ESHAPE;
    if (mode == kDw && (s.N != groups[0].N || s.K != groups[0].K)) return MNR_ESHAPE;  // shared partial layout
    out->g[k] = Group{s.a, s.b, s.c, s.bias, s.saved, s.bias_grad, s.N, s.K, s.lda > 0 ? s.lda : s.K, s.ldb > 0 ? s.ldb : s.K};
  }
  return MNR_OK;
}

int dw_splits(int M) {
  int s = (M + 511) / 512;  // >= 512 rows per CTA
  const int cap = 4 * mnr_sm_count();
  if (s > cap) s = cap;
  return s < 1 ? 1 : s;
}

// Thin layers (N <= 16: heads): the weight gradient is a row stream whose per-CTA loop is a chain of dependent DRAM round
// trips, so it wants MANY short slabs (64 rows per CTA: 54 -> ~10 us for 8192 x 128 -> 8 measured with 512-row slabs); the
// partials are N*K floats each, so hundreds of splits cost nothing.  Chosen by SHAPE only, so that the workspace size and the
// launch agree whatever the pointers' alignment decides about the kernel.
int dw_splits_for(int M, int N, int K) {
  if (N > 16 || K % 4 != 0 || K > 1024) return dw_splits(M);
  int s = (M + 63) / 64;
  const int cap = 4 * mnr_sm_count();
  if (s > cap) s = cap;
  return s < 1 ? 1 : s;
}

}  // namespace

extern "C" int mnr_linear_fwd(const mnr_linear_group_t* groups, int n_groups, int M, int act, mnr_stream_t stream) {
  Groups gs;
  const int rc = pack_groups(groups, n_groups, kFwd, &gs);
  if (rc != MNR_OK) return rc;
  if (M <= 0) return MNR_ESHAPE;
  if (act < 0 || act > 2) return MNR_EUNSUP;
  int maxN = 0, maxK = 0;
  for (int k = 0; k < n_groups; ++k) maxN = gs.g[k].N > maxN ? gs.g[k].N : maxN, maxK = gs.g[k].K > maxK ? gs.g[k].K : maxK;
  if (act == 0 && thin_ok(gs, kFwd, act, 32) && (size_t)thin_np(maxN) * maxK * 4 <= 48 * 1024) {
    const size_t smem = (size_t)thin_np(maxN) * maxK * 4;
    dim3 tgrid(thin_grid(M), n_groups);
    MNR_THIN_DISPATCH(thin_np(maxN), (thin_fwd_kernel<NP><<<tgrid, kThinThreads, smem, mnr_cs(stream)>>>(gs, M)));
    MNR_LAUNCH_RESULT();
  }
  dim3 grid((M + BM - 1) / BM, (maxN + BN - 1) / BN, n_groups);
  linear_kernel<kFwd><<<grid, kThreads, 0, mnr_cs(stream)>>>(gs, M, act, 0, 1, nullptr);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_linear_bwd_input(const mnr_linear_group_t* groups, int n_groups, int M, int act, int accumulate,
                                    mnr_stream_t stream) {
  Groups gs;
  const int rc = pack_groups(groups, n_groups, kDx, &gs);
  if (rc != MNR_OK) return rc;
  if (M <= 0) return MNR_ESHAPE;
  if (act < 0 || act > 2) return MNR_EUNSUP;
  int maxK = 0, maxN = 0;
  for (int k = 0; k < n_groups; ++k) maxK = gs.g[k].K > maxK ? gs.g[k].K : maxK, maxN = gs.g[k].N > maxN ? gs.g[k].N : maxN;
  if (thin_ok(gs, kDx, act, 32) && (size_t)thin_np(maxN) * maxK * 4 <= 48 * 1024) {
    const size_t smem = (size_t)thin_np(maxN) * maxK * 4;
    dim3 tgrid(thin_grid(M), n_groups);
    MNR_THIN_DISPATCH(thin_np(maxN), (thin_dx_kernel<NP><<<tgrid, kThinThreads, smem, mnr_cs(stream)>>>(gs, M, accumulate)));
    MNR_LAUNCH_RESULT();
  }
  dim3 grid((M + BM - 1) / BM, (maxK + BN - 1) / BN, n_groups);
  linear_kernel<kDx><<<grid, kThreads, 0, mnr_cs(stream)>>>(gs, M, act, accumulate, 1, nullptr);
  MNR_LAUNCH_RESULT();
}

extern "C" size_t mnr_linear_bwd_weight_workspace_bytes(int n_groups, int M, int N, int K) {
  if (n_groups <= 0 || M <= 0 || N <= 0 || K <= 0) return 0;
  return (size_t)n_groups * dw_splits_for(M, N, K) * ((size_t)N * K + N) * sizeof(float);
}

extern "C" int mnr_linear_bwd_weight(const mnr_linear_group_t* groups, int n_groups, int M, int act, void* workspace,
                                     size_t workspace_bytes, mnr_stream_t stream) {
  Groups gs;
  const int rc = pack_groups(groups, n_groups, kDw, &gs);
  if (rc != MNR_OK) return rc;
  if (M <= 0) return MNR_ESHAPE;
  if (act < 0 || act > 2) return MNR_EUNSUP;
  const int N = gs.g[0].N, K = gs.g[0].K;
  const int splits = dw_splits_for(M, N, K);
  if (splits > 1) {
    MNR_CHECK_PTR(workspace);
    if (workspace_bytes < mnr_linear_bwd_weight_workspace_bytes(n_groups, M, N, K)) return MNR_EWORKSPACE;
  }
  if (thin_ok(gs, kDw, act, 16)) {
    dim3 tgrid(splits, n_groups);
    MNR_THIN_DISPATCH(thin_np(N), (thin_dw_kernel<NP><<<tgrid, kThinThreads, 0, mnr_cs(stream)>>>(gs, M, splits, (float*)workspace)));
  } else {
    dim3 grid((N + BM - 1) / BM, (K + BN - 1) / BN, n_groups * splits);
    linear_kernel<kDw><<<grid, kThreads, 0, mnr_cs(stream)>>>(gs, M, act, 0, splits, (float*)workspace);
  }
  if (splits > 1) {
    dim3 rgrid((N * K + N + 255) / 256, n_groups);
    reduce_splits_kernel<<<rgrid, 256, 0, mnr_cs(stream)>>>(gs, splits, (const float*)workspace);
  }
  MNR_LAUNCH_RESULT();
}

extern "C" size_t mnr_colsum_workspace_bytes(int C) { return 256 + (size_t)2 * mnr_sm_count() * (C > 0 ? C : 1) * sizeof(float); }

extern "C" int mnr_colsum_f32(const float* x, int64_t M, int C, float* out, void* workspace, size_t workspace_bytes,
                              mnr_stream_t stream) {
  MNR_CHECK_PTR(x); MNR_CHECK_PTR(out); MNR_CHECK_PTR(workspace);
  if (M <= 0 || C <= 0 || C > 256) return MNR_ESHAPE;
  if (workspace_bytes < mnr_colsum_workspace_bytes(C)) return MNR_EWORKSPACE;
  const int rpar = 256 / C;
  int64_t want = (M + (int64_t)rpar * 64 - 1) / ((int64_t)rpar * 64);
  const int64_t cap = 2 * mnr_sm_count();
  const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  colsum_kernel<<<grid, 256, 0, mnr_cs(stream)>>>(x, M, C, out, (float*)((unsigned char*)workspace + 256),
                                                  (unsigned int*)workspace);
  MNR_LAUNCH_RESULT();
}
