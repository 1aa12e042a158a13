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
  if (idx < NK) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += __ldg(partials + ((size_t)gi * splits + z) * NK + idx);
    g.c[idx] = s;
  } else if (idx < NK + g.N && g.bias_grad) {
    const int n = idx - NK;
    float s = 0.f;
    const float* bp = partials + (size_t)gs.n * splits * NK;
    for (int z = 0; z < splits; ++z) s += __ldg(bp + ((size_t)gi * splits + z) * g.N + n);
    g.bias_grad[n] = s;
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

}  // namespace

extern "C" int mnr_linear_fwd(const mnr_linear_group_t* groups, int n_groups, int M, int act, mnr_stream_t stream) {
  Groups gs;
  const int rc = pack_groups(groups, n_groups, kFwd, &gs);
  if (rc != MNR_OK) return rc;
  if (M <= 0) return MNR_ESHAPE;
  if (act < 0 || act > 2) return MNR_EUNSUP;
  int maxN = 0;
  for (int k = 0; k < n_groups; ++k) maxN = gs.g[k].N > maxN ? gs.g[k].N : maxN;
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
  int maxK = 0;
  for (int k = 0; k < n_groups; ++k) maxK = gs.g[k].K > maxK ? gs.g[k].K : maxK;
  dim3 grid((M + BM - 1) / BM, (maxK + BN - 1) / BN, n_groups);
  linear_kernel<kDx><<<grid, kThreads, 0, mnr_cs(stream)>>>(gs, M, act, accumulate, 1, nullptr);
  MNR_LAUNCH_RESULT();
}

extern "C" size_t mnr_linear_bwd_weight_workspace_bytes(int n_groups, int M, int N, int K) {
  if (n_groups <= 0 || M <= 0 || N <= 0 || K <= 0) return 0;
  return (size_t)n_groups * dw_splits(M) * ((size_t)N * K + N) * sizeof(float);
}

extern "C" int mnr_linear_bwd_weight(const mnr_linear_group_t* groups, int n_groups, int M, int act, void* workspace,
                                     size_t workspace_bytes, mnr_stream_t stream) {
  Groups gs;
  const int rc = pack_groups(groups, n_groups, kDw, &gs);
  if (rc != MNR_OK) return rc;
  if (M <= 0) return MNR_ESHAPE;
  if (act < 0 || act > 2) return MNR_EUNSUP;
  const int N = gs.g[0].N, K = gs.g[0].K;
  const int splits = dw_splits(M);
  if (splits > 1) {
    MNR_CHECK_PTR(workspace);
    if (workspace_bytes < mnr_linear_bwd_weight_workspace_bytes(n_groups, M, N, K)) return MNR_EWORKSPACE;
  }
  dim3 grid((N + BM - 1) / BM, (K + BN - 1) / BN, n_groups * splits);
  linear_kernel<kDw><<<grid, kThreads, 0, mnr_cs(stream)>>>(gs, M, act, 0, splits, (float*)workspace);
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
