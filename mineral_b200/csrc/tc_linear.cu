// G  tensor-core linear layers for the actor/critic towers: tcgen05 (5th-gen tensor cores, TMEM
// accumulators) with a 3xTF32 split so the result keeps fp32-level accuracy (the parity bar is 1e-5).
//
//   x = hi + lo with hi = x & 0xffffe000 (exactly a tf32), lo = x - hi (exact in fp32)
//   a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b     (three kind::tf32 MMAs into the same fp32 accumulator;
//                                                 the dropped lo*lo term is < 2^-20 relative)
//
// One kernel, three modes, C[i][j] = sum_r A(i,r) B(r,j), tile 128 x BN, 32-wide reduction blocks:
//   FWD  Y  = act(X W^T + b)          A = X  (K-major)            B = W pre-split hi/lo (K-major)
//   DX   dX = (dY * act'(Y)) W        A = dY,Y (K-major, fused)   B = W pre-split hi/lo (MN-major: no W^T copy)
//   DW   dW = (dY * act'(Y))^T X      A = dY,Y (MN-major, fused)  B = X (MN-major, split in-kernel); + db
// so neither transposed activations nor a dZ tensor ever exist in HBM.
//
// Warp roles (256 threads, 1 CTA/SM):
//   warp 0    TMA producer: cp.async.bulk.tensor (SWIZZLE_128B boxes) into a 2-4 stage smem ring
//   warps 4-7 converters: turn the fp32 tiles into hi / lo tiles in place (fusing act' and the bias-gradient
//             column sums), fence.proxy.async, arrive; afterwards the same warps run the epilogue
//             (tcgen05.ld from TMEM -> bias + activation -> global)
//   warp 1    MMA issuer: one thread issues 12 tcgen05.mma (4 k-steps x 3 products) per stage and
//             tcgen05.commit's the stage back to the producer
//   warp 2    TMEM allocator
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBM = 128, kBK = 32;
constexpr int kTileA = kBM * kBK * 4;  // 16 KiB
constexpr int kChunk = 32 * 32 * 4;    // 4 KiB: one MN-major chunk (32 reduction rows x 32 floats)
constexpr int kMaxGroups = 2;
constexpr int kMaxStages = 4;

enum Mode { kFwd = 0, kDx = 1, kDw = 2 };

struct alignas(64) TcGroup {
  CUtensorMap a, s, b0, b1;  // b0: B (hi, or raw when converted in-kernel), b1: B lo (pre-split modes)
  float* c;
  const float* bias;
};

struct TcParams {
  TcGroup g[kMaxGroups];
  int I, J, R;  // C rows, C cols, reduction length
  int BN;       // C tile columns = MMA N
  int ldc;
  int act;
  int splits;
  int stages;
  int rows_per_split;  // DW, multiple of 32
  float* partials;     // DW: [groups*splits][I][J] then bias [groups*splits][I]
  int want_bias;
  int n_groups;
};

__device__ __forceinline__ float act_fwd(float z, int act) {
  switch (act) {
    case 1: return z > 0.f ? z : __expf(z) - 1.f;  // ex2.approx path: abs error < 1e-6 on (-1, 0], the epilogue is
                                                   // instruction-bound and expm1f costs ~30 instructions per element
    case 2: return fmaxf(z, 0.f);
    default: return z;
  }
}
__device__ __forceinline__ float act_bwd_from_y(float y, int act) {
  switch (act) {
    case 1: return y > 0.f ? 1.f : y + 1.f;
    case 2: return y > 0.f ? 1.f : 0.f;
    default: return 1.f;
  }
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__device__ __forceinline__ void split4(float4 v, float4& hi, float4& lo) {
  hi.x = tf32_hi(v.x), hi.y = tf32_hi(v.y), hi.z = tf32_hi(v.z), hi.w = tf32_hi(v.w);
  lo.x = v.x - hi.x, lo.y = v.y - hi.y, lo.z = v.z - hi.z, lo.w = v.w - hi.w;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) tc_linear_kernel(const __grid_constant__ TcParams p) {
  constexpr bool A_MN = (MODE == kDw), B_MN = (MODE != kFwd), HAS_S = (MODE != kFwd), CONV_B = (MODE == kDw);
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_tma[kMaxStages], full_conv[kMaxStages], empty_bar[kMaxStages], tmem_full;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gi = blockIdx.z / p.splits, split = blockIdx.z - gi * p.splits;
  const TcGroup& g = p.g[gi];
  const int i0 = blockIdx.x * kBM, j0 = blockIdx.y * p.BN;
  const int BN = p.BN, stages = p.stages;

  // stage layout (1024 B aligned: SWIZZLE_128B atoms are 8 x 128 B)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int offAlo = kTileA, offS = 2 * kTileA;
  const int offB = HAS_S ? 3 * kTileA : 2 * kTileA;
  const int bytesB = BN * kBK * 4;
  const int offBlo = offB + bytesB;
  const int stage_bytes = offBlo + bytesB;

  int r_begin = 0, r_end = p.R;
  if (MODE == kDw) {
    r_begin = split * p.rows_per_split;
    r_end = min(p.R, r_begin + p.rows_per_split);
  }
  const int num_kb = r_end > r_begin ? (r_end - r_begin + kBK - 1) / kBK : 0;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full_tma[s], 1);
      ptx::mbar_init(&full_conv[s], 128);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(&tmem_full, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&g.a);
    ptx::prefetch_tmap(&g.b0);
  }
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < BN) tmem_cols <<= 1;
  if (warp == 2) ptx::tmem_alloc(&tmem_slot, tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer ==========================================================================
    if (lane == 0) {
      const uint32_t tx_bytes = kTileA * (HAS_S ? 2 : 1) + bytesB * (CONV_B ? 1 : 2);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % stages;
        ptx::mbar_wait(&empty_bar[st], ((kb / stages) & 1) ^ 1);
        uint8_t* base = smem + (size_t)st * stage_bytes;
        ptx::mbar_expect_tx(&full_tma[st], tx_bytes);
        const int r0 = r_begin + kb * kBK;
        if (!A_MN) {  // box [128 rows][32 floats along the reduction]
          ptx::tma_load_2d(base, &g.a, r0, i0, &full_tma[st]);
          if (HAS_S) ptx::tma_load_2d(base + offS, &g.s, r0, i0, &full_tma[st]);
        } else {      // 4 chunks of [32 reduction rows][32 floats along i]
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            ptx::tma_load_2d(base + c * kChunk, &g.a, i0 + 32 * c, r0, &full_tma[st]);
            ptx::tma_load_2d(base + offS + c * kChunk, &g.s, i0 + 32 * c, r0, &full_tma[st]);
          }
        }
        if (!B_MN) {  // box [BN rows][32 floats]
          ptx::tma_load_2d(base + offB, &g.b0, r0, j0, &full_tma[st]);
          ptx::tma_load_2d(base + offBlo, &g.b1, r0, j0, &full_tma[st]);
        } else {
          for (int c = 0; c < BN / 32; ++c) {
            ptx::tma_load_2d(base + offB + c * kChunk, &g.b0, j0 + 32 * c, r0, &full_tma[st]);
            if (!CONV_B) ptx::tma_load_2d(base + offBlo + c * kChunk, &g.b1, j0 + 32 * c, r0, &full_tma[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ===============================================================================
    if (lane == 0) {
      const uint32_t idesc = ptx::idesc_tf32(kBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % stages;
        ptx::mbar_wait(&full_conv[st], (kb / stages) & 1);
        ptx::tc_fence_after();
        const uint32_t base = ptx::smem_u32(smem + (size_t)st * stage_bytes);
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          // K-major: step 8 floats (32 B) inside the 128 B swizzle row; MN-major: step 8 reduction rows (1 KiB)
          const uint32_t a_off = A_MN ? k * 1024 : k * 32;
          const uint32_t b_off = B_MN ? k * 1024 : k * 32;
          // K-major: SWIZZLE_128B, 8-row groups 1 KiB apart.  MN-major (32-bit): SWIZZLE_128B_BASE32B, 4-row
          // reduction groups 512 B apart, 32-float chunks along M/N one 4 KiB box apart.
          const uint32_t a_lbo = A_MN ? kChunk : 16, b_lbo = B_MN ? kChunk : 16;
          const uint32_t a_sbo = A_MN ? 512 : 1024, b_sbo = B_MN ? 512 : 1024;
          const uint32_t a_lt = A_MN ? 1 : 2, b_lt = B_MN ? 1 : 2;
          const uint64_t a_hi = ptx::smem_desc(base + a_off, a_lbo, a_sbo, a_lt);
          const uint64_t a_lo = ptx::smem_desc(base + offAlo + a_off, a_lbo, a_sbo, a_lt);
          const uint64_t b_hi = ptx::smem_desc(base + offB + b_off, b_lbo, b_sbo, b_lt);
          const uint64_t b_lo = ptx::smem_desc(base + offBlo + b_off, b_lbo, b_sbo, b_lt);
          ptx::mma_tf32(tmem_base, a_lo, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);  // small terms first
          ptx::mma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
          ptx::mma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
        }
        ptx::tc_commit(&empty_bar[st]);  // stage reusable once these MMAs have read it
      }
      ptx::tc_commit(&tmem_full);
    }
  } else if (warp >= 4) {
    // ===== converters, then epilogue ==================================================================
    const int t = tid - 128;
    const int act = p.act;
    (void)0;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int st = kb % stages;
      ptx::mbar_wait(&full_tma[st], (kb / stages) & 1);
      uint8_t* base = smem + (size_t)st * stage_bytes;
      float4* A4 = reinterpret_cast<float4*>(base);
      float4* Alo4 = reinterpret_cast<float4*>(base + offAlo);
      const float4* S4 = reinterpret_cast<const float4*>(base + offS);
      if (MODE != kDw) {
        // element-wise: the swizzle is irrelevant, walk the tile linearly
#pragma unroll 4
        for (int q = t; q < kTileA / 16; q += 128) {
          float4 a = A4[q];
          if (HAS_S) {
            const float4 s = S4[q];
            a.x *= act_bwd_from_y(s.x, act), a.y *= act_bwd_from_y(s.y, act);
            a.z *= act_bwd_from_y(s.z, act), a.w *= act_bwd_from_y(s.w, act);
          }
          float4 hi, lo;
          split4(a, hi, lo);
          A4[q] = hi, Alo4[q] = lo;
        }
      } else {
        // MN-major A: chunk c = 32 columns of i; thread owns logical 16 B slot j (4 columns) of rows p, p+4, ...
        // SWIZZLE_128B_ATOM_32B stores the logical 32 B chunk (j >> 1) of row r at physical chunk (j >> 1) ^ (r & 3)
        const int c = t >> 5, j = t & 7, ph = (t & 31) >> 3;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int r = ph + 4 * rr;
          const int q = c * (kChunk / 16) + r * 8 + ((((j >> 1) ^ (r & 3)) << 1) | (j & 1));
          float4 a = A4[q];
          const float4 s = S4[q];
          a.x *= act_bwd_from_y(s.x, act), a.y *= act_bwd_from_y(s.y, act);
          a.z *= act_bwd_from_y(s.z, act), a.w *= act_bwd_from_y(s.w, act);
          bsum.x += a.x, bsum.y += a.y, bsum.z += a.z, bsum.w += a.w;
          float4 hi, lo;
          split4(a, hi, lo);
          A4[q] = hi, Alo4[q] = lo;
        }
        float4* B4 = reinterpret_cast<float4*>(base + offB);
        float4* Blo4 = reinterpret_cast<float4*>(base + offBlo);
        for (int q = t; q < bytesB / 16; q += 128) {
          float4 hi, lo;
          split4(B4[q], hi, lo);
          B4[q] = hi, Blo4[q] = lo;
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&full_conv[st]);
    }

    const int q4 = warp - 4;
    if (MODE == kDw && p.want_bias && blockIdx.y == 0) {
      // fold the 4 row phases (lanes j, j+8, j+16, j+24), then lanes 0..7 own 4 columns each of chunk c = warp-4
      bsum.x += __shfl_xor_sync(0xffffffffu, bsum.x, 8), bsum.y += __shfl_xor_sync(0xffffffffu, bsum.y, 8);
      bsum.z += __shfl_xor_sync(0xffffffffu, bsum.z, 8), bsum.w += __shfl_xor_sync(0xffffffffu, bsum.w, 8);
      bsum.x += __shfl_xor_sync(0xffffffffu, bsum.x, 16), bsum.y += __shfl_xor_sync(0xffffffffu, bsum.y, 16);
      bsum.z += __shfl_xor_sync(0xffffffffu, bsum.z, 16), bsum.w += __shfl_xor_sync(0xffffffffu, bsum.w, 16);
      if (lane < 8) {
        float* bp = p.partials + (size_t)p.n_groups * p.splits * p.I * p.J + ((size_t)gi * p.splits + split) * p.I;
        const int n = i0 + 32 * q4 + 4 * lane;
        const float vals[4] = {bsum.x, bsum.y, bsum.z, bsum.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (n + u < p.I) bp[n + u] = vals[u];
      }
    }
  }

  // ===== epilogue, all 8 warps: TMEM -> registers -> shared (row-major staging) -> coalesced global stores ====
  // Once tmem_full has fired every MMA has finished reading the pipeline stages and no TMA is in flight, so the
  // stage memory is reused as a 128 x (BN+4) f32 staging tile.  A warp can only read its own quarter (warp % 4)
  // of the TMEM lanes; the two warps of a quarter take alternate 32-column chunks.  After a CTA barrier the tile
  // is streamed out 16 rows per warp with lanes along the columns: every store instruction writes full lines.
  {
    const int act = p.act;
    if (num_kb > 0) {
      ptx::mbar_wait(&tmem_full, 0);
      ptx::tc_fence_after();
    }
    const int pitch = BN + 4;  // floats; +4 keeps the per-lane float4 row writes bank-conflict free
    float* stage_tile = reinterpret_cast<float*>(smem);
    const int quarter = warp & 3, half = warp >> 2;
    for (int cc = half; cc < (BN + 31) / 32; cc += 2) {
      uint32_t r[32];
      if (num_kb > 0) {
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(cc * 32), r);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) r[q] = 0u;
      }
      float* dstrow = stage_tile + (size_t)(32 * quarter + lane) * pitch + cc * 32;
#pragma unroll
      for (int q = 0; q < 32; q += 4)
        if (cc * 32 + q < BN)
          *reinterpret_cast<float4*>(dstrow + q) = make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]),
                                                               __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
    }
    ptx::tc_fence_before();
    __syncthreads();
    const int ncols = min(BN, p.J - j0);  // valid columns of this tile
    float* cbase;
    size_t ldo;
    if (MODE == kDw) cbase = p.partials + ((size_t)gi * p.splits + split) * p.I * p.J, ldo = p.J;
    else cbase = g.c, ldo = p.ldc;
    const bool vec = (ncols % 4 == 0) && (ldo % 4 == 0) && (j0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(cbase) & 15) == 0);
    for (int rr = 0; rr < 16; ++rr) {
      const int rl = 16 * warp + rr;
      const int row = i0 + rl;
      if (row >= p.I) break;
      const float* src = stage_tile + (size_t)rl * pitch;
      float* dst = cbase + (size_t)row * ldo + j0;
      if (vec) {
        for (int c4 = lane; c4 < ncols / 4; c4 += 32) {
          float4 v = *reinterpret_cast<const float4*>(src + 4 * c4);
          if (MODE == kFwd) {
            if (g.bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + j0) + c4);
              v.x += bb.x, v.y += bb.y, v.z += bb.z, v.w += bb.w;
            }
            v.x = act_fwd(v.x, act), v.y = act_fwd(v.y, act), v.z = act_fwd(v.z, act), v.w = act_fwd(v.w, act);
          }
          *reinterpret_cast<float4*>(dst + 4 * c4) = v;
        }
      } else {
        for (int c = lane; c < ncols; c += 32) {
          float v = src[c];
          if (MODE == kFwd) v = act_fwd(v + (g.bias ? __ldg(g.bias + j0 + c) : 0.f), act);
          dst[c] = v;
        }
      }
    }
  }

  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dW / db = sum over splits (fixed order)
struct ReduceArgs {
  float* w[kMaxGroups];
  float* b[kMaxGroups];
};
__global__ void __launch_bounds__(256) tc_reduce_kernel(ReduceArgs ra, int n_groups, int splits, int I, int J,
                                                        const float* __restrict__ partials) {
  const int gi = blockIdx.y;
  const int NK = I * J;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx < NK) {
    const float* src = partials + (size_t)gi * splits * NK + idx;
    float s = 0.f;
    int z = 0;
    for (; z + 8 <= splits; z += 8) {  // 8 independent loads in flight, summed in a fixed order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (size_t)(z + u) * NK);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; z < splits; ++z) s += __ldg(src + (size_t)z * NK);
    ra.w[gi][idx] = s;
  } else if (idx < NK + I && ra.b[gi]) {
    const int n = idx - NK;
    const float* bp = partials + (size_t)n_groups * splits * NK;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += __ldg(bp + ((size_t)gi * splits + z) * I + n);
    ra.b[gi][n] = s;
  }
}

// W (rows, cols) -> hi / lo tf32 parts in (rows, dst_ld) buffers, zero padded
struct SplitJob {
  const float* src;
  float *hi, *lo;
  int rows, cols, src_ld, dst_ld;
};
struct SplitJobs {
  SplitJob j[MNR_MAX_SPLIT_JOBS];
  int n;
};
__global__ void __launch_bounds__(256) split_kernel(const __grid_constant__ SplitJobs js) {
  const SplitJob& jb = js.j[blockIdx.y];
  const int total = jb.rows * jb.dst_ld;
  for (int idx = blockIdx.x * 256 + threadIdx.x; idx < total; idx += gridDim.x * 256) {
    const int r = idx / jb.dst_ld, c = idx - r * jb.dst_ld;
    const float v = c < jb.cols ? __ldg(jb.src + (size_t)r * jb.src_ld + c) : 0.f;
    const float hi = tf32_hi(v);
    jb.hi[idx] = hi;
    jb.lo[idx] = v - hi;
  }
}

// ---- host ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// f32 matrix (rows, ld) -> map with box [box_rows][32 floats], zero OOB fill.  K-major operand tiles use
// SWIZZLE_128B, MN-major ones SWIZZLE_128B_ATOM_32B (matches UMMA SWIZZLE_128B_BASE32B).
int make_map(CUtensorMap* m, const float* base, int rows, int cols, int ld, int box_rows, bool mn_major = false) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return MNR_EUNSUP;
  if (!mnr_aligned(base, 16) || (ld % 4) != 0 || box_rows > 256) return MNR_EALIGN;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MNR_OK : MNR_EUNSUP;
}

int pick_stages(int mode, int BN) {
  const int stage = (mode == kFwd ? 2 : 3) * kTileA + 2 * BN * kBK * 4;
  int s = (227 * 1024 - 2048) / stage;
  if (s > kMaxStages) s = kMaxStages;
  return s;
}

template <int MODE>
int launch(const TcParams& p, dim3 grid, cudaStream_t stream) {
  const int stage = (MODE == kFwd ? 2 : 3) * kTileA + 2 * p.BN * kBK * 4;
  const size_t smem = (size_t)p.stages * stage + 1024;
  cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  tc_linear_kernel<MODE><<<grid, kThreads, smem, stream>>>(p);
  e = cudaPeekAtLastError();
  return e == cudaSuccess ? MNR_OK : (int)e;
}

int ld_or(int ld, int dflt) { return ld > 0 ? ld : dflt; }

}  // namespace

extern "C" int mnr_tc_supported(void) { return encode_fn() != nullptr ? 1 : 0; }

extern "C" int mnr_tc_linear_fwd(const mnr_linear_group_t* groups, int n_groups, int M, int act, mnr_stream_t stream) {
  if (n_groups <= 0 || n_groups > kMaxGroups) return MNR_ESHAPE;
  MNR_CHECK_PTR(groups);
  if (M <= 0) return MNR_ESHAPE;
  if (act < 0 || act > 2) return MNR_EUNSUP;
  const int N = groups[0].N, K = groups[0].K;
  if (N <= 0 || K <= 0 || N % 16 != 0) return MNR_ESHAPE;
  TcParams p;
  p.BN = N <= 256 ? N : 128;
  if (N > 256 && N % 128 != 0) return MNR_ESHAPE;
  for (int k = 0; k < n_groups; ++k) {
    const mnr_linear_group_t& s = groups[k];
    if (!s.a || !s.b || !s.b_lo || !s.c) return MNR_ENULL;
    if (s.N != N || s.K != K) return MNR_ESHAPE;
    int rc = make_map(&p.g[k].a, s.a, M, K, ld_or(s.lda, K), kBM);
    if (rc == MNR_OK) rc = make_map(&p.g[k].b0, s.b, N, K, ld_or(s.ldb, K), p.BN);
    if (rc == MNR_OK) rc = make_map(&p.g[k].b1, s.b_lo, N, K, ld_or(s.ldb, K), p.BN);
    if (rc != MNR_OK) return rc;
    p.g[k].s = p.g[k].a;
    p.g[k].c = s.c, p.g[k].bias = s.bias;
  }
  p.I = M, p.J = N, p.R = K, p.ldc = N, p.act = act, p.splits = 1, p.rows_per_split = 0, p.partials = nullptr;
  p.want_bias = 0, p.n_groups = n_groups, p.stages = pick_stages(kFwd, p.BN);
  dim3 grid((M + kBM - 1) / kBM, (N + p.BN - 1) / p.BN, n_groups);
  return launch<kFwd>(p, grid, mnr_cs(stream));
}

extern "C" int mnr_tc_linear_bwd_input(const mnr_linear_group_t* groups, int n_groups, int M, int act, mnr_stream_t stream) {
  if (n_groups <= 0 || n_groups > kMaxGroups) return MNR_ESHAPE;
  MNR_CHECK_PTR(groups);
  if (M <= 0) return MNR_ESHAPE;
  if (act < 0 || act > 2) return MNR_EUNSUP;
  const int N = groups[0].N, K = groups[0].K;
  if (N <= 0 || K <= 0 || N % 4 != 0) return MNR_ESHAPE;
  TcParams p;
  const int Kr = (K + 31) / 32 * 32;
  p.BN = Kr <= 256 ? Kr : 256;
  for (int k = 0; k < n_groups; ++k) {
    const mnr_linear_group_t& s = groups[k];
    if (!s.a || !s.b || !s.b_lo || !s.c || !s.saved) return MNR_ENULL;
    if (s.N != N || s.K != K) return MNR_ESHAPE;
    int rc = make_map(&p.g[k].a, s.a, M, N, N, kBM);                                 // dY (M,N), reduction over N
    if (rc == MNR_OK) rc = make_map(&p.g[k].s, s.saved, M, N, N, kBM);               // Y
    if (rc == MNR_OK) rc = make_map(&p.g[k].b0, s.b, N, K, ld_or(s.ldb, K), 32, true);  // W (N,K): rows = reduction
    if (rc == MNR_OK) rc = make_map(&p.g[k].b1, s.b_lo, N, K, ld_or(s.ldb, K), 32, true);
    if (rc != MNR_OK) return rc;
    p.g[k].c = s.c, p.g[k].bias = nullptr;
  }
  p.I = M, p.J = K, p.R = N, p.ldc = K, p.act = act, p.splits = 1, p.rows_per_split = 0, p.partials = nullptr;
  p.want_bias = 0, p.n_groups = n_groups, p.stages = pick_stages(kDx, p.BN);
  dim3 grid((M + kBM - 1) / kBM, (K + p.BN - 1) / p.BN, n_groups);
  return launch<kDx>(p, grid, mnr_cs(stream));
}

namespace {
int tc_dw_splits(int M, int tiles) {
  int s = mnr_sm_count() / tiles;      // one wave: 1 CTA per SM (the kernel's smem footprint allows no more)
  const int max_s = (M + 255) / 256;   // >= 256 rows (8 k-blocks) per CTA
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}
}  // namespace

extern "C" size_t mnr_tc_linear_bwd_weight_workspace_bytes(int n_groups, int M, int N, int K) {
  if (n_groups <= 0 || M <= 0 || N <= 0 || K <= 0) return 0;
  const int Kr = (K + 31) / 32 * 32, BN = Kr <= 256 ? Kr : 256;
  const int tiles = ((N + kBM - 1) / kBM) * ((K + BN - 1) / BN) * n_groups;
  return (size_t)n_groups * tc_dw_splits(M, tiles) * ((size_t)N * K + N) * sizeof(float);
}

extern "C" int mnr_tc_linear_bwd_weight(const mnr_linear_group_t* groups, int n_groups, int M, int act, void* workspace,
                                        size_t workspace_bytes, mnr_stream_t stream) {
  if (n_groups <= 0 || n_groups > kMaxGroups) return MNR_ESHAPE;
  MNR_CHECK_PTR(groups); MNR_CHECK_PTR(workspace);
  if (M <= 0) return MNR_ESHAPE;
  if (act < 0 || act > 2) return MNR_EUNSUP;
  const int N = groups[0].N, K = groups[0].K;
  if (N <= 0 || K <= 0 || N % 4 != 0) return MNR_ESHAPE;
  if (workspace_bytes < mnr_tc_linear_bwd_weight_workspace_bytes(n_groups, M, N, K)) return MNR_EWORKSPACE;
  TcParams p;
  const int Kr = (K + 31) / 32 * 32;
  p.BN = Kr <= 256 ? Kr : 256;
  ReduceArgs ra;
  bool want_bias = false;
  for (int k = 0; k < n_groups; ++k) {
    const mnr_linear_group_t& s = groups[k];
    if (!s.a || !s.b || !s.c || !s.saved) return MNR_ENULL;
    if (s.N != N || s.K != K) return MNR_ESHAPE;
    int rc = make_map(&p.g[k].a, s.a, M, N, N, 32, true);                            // dY (M,N): rows = reduction
    if (rc == MNR_OK) rc = make_map(&p.g[k].s, s.saved, M, N, N, 32, true);
    if (rc == MNR_OK) rc = make_map(&p.g[k].b0, s.b, M, K, ld_or(s.ldb, K), 32, true);  // X (M,K)
    if (rc != MNR_OK) return rc;
    p.g[k].b1 = p.g[k].b0;
    p.g[k].c = s.c, p.g[k].bias = nullptr;
    ra.w[k] = s.c, ra.b[k] = s.bias_grad;
    want_bias = want_bias || s.bias_grad != nullptr;
  }
  for (int k = n_groups; k < kMaxGroups; ++k) ra.w[k] = nullptr, ra.b[k] = nullptr;
  const int ti = (N + kBM - 1) / kBM, tj = (K + p.BN - 1) / p.BN;
  const int splits = tc_dw_splits(M, ti * tj * n_groups);
  p.I = N, p.J = K, p.R = M, p.ldc = K, p.act = act, p.splits = splits;
  p.rows_per_split = ((M + splits - 1) / splits + 31) / 32 * 32;
  p.partials = (float*)workspace, p.want_bias = want_bias ? 1 : 0, p.n_groups = n_groups, p.stages = pick_stages(kDw, p.BN);
  dim3 grid(ti, tj, n_groups * splits);
  const int rc = launch<kDw>(p, grid, mnr_cs(stream));
  if (rc != MNR_OK) return rc;
  dim3 rgrid((N * K + N + 255) / 256, n_groups);
  tc_reduce_kernel<<<rgrid, 256, 0, mnr_cs(stream)>>>(ra, n_groups, splits, N, K, (const float*)workspace);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_split_tf32(const mnr_split_job_t* jobs, int n_jobs, mnr_stream_t stream) {
  if (n_jobs <= 0 || n_jobs > MNR_MAX_SPLIT_JOBS) return MNR_ESHAPE;
  MNR_CHECK_PTR(jobs);
  SplitJobs js;
  js.n = n_jobs;
  int max_total = 0;
  for (int k = 0; k < n_jobs; ++k) {
    if (!jobs[k].src || !jobs[k].hi || !jobs[k].lo) return MNR_ENULL;
    if (jobs[k].rows <= 0 || jobs[k].cols <= 0 || jobs[k].dst_ld < jobs[k].cols) return MNR_ESHAPE;
    js.j[k] = SplitJob{jobs[k].src, jobs[k].hi, jobs[k].lo, jobs[k].rows, jobs[k].cols, jobs[k].src_ld > 0 ? jobs[k].src_ld : jobs[k].cols,
                       jobs[k].dst_ld};
    const int total = jobs[k].rows * jobs[k].dst_ld;
    max_total = total > max_total ? total : max_total;
  }
  dim3 grid((max_total + 255) / 256 > 64 ? 64 : (max_total + 255) / 256, n_jobs);
  split_kernel<<<grid, 256, 0, mnr_cs(stream)>>>(js);
  MNR_LAUNCH_RESULT();
}
