// G  tensor-core linear layers for the actor/critic towers: tcgen05 (5th-gen tensor cores, TMEM
// accumulators) with a 3xTF32 split so the result keeps fp32-level accuracy (the parity bar is 1e-5).
//
//   x = hi + lo with hi = x & 0xffffe000 (exactly a tf32), lo = x - hi (exact in fp32)
//   a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b     (three kind::tf32 MMAs into the same fp32 accumulator;
//                                                 the dropped lo*lo term is < 2^-20 relative)
//
// One kernel, three modes, C[i][j] = sum_r A(i,r) B(r,j), tile 128 x BN, 32-wide reduction blocks:
//   FWD  Y  = act(X W^T + b)          A = X  (K-major tile)             B = W pre-split hi/lo (K-major)
//   DX   dX = (dY * act'(Y)) W        A = dY,Y (K-major tiles, fused)   B = W pre-split hi/lo (MN-major: no W^T copy)
//   DW   dW = (dY * act'(Y))^T X      A = dY,Y (MN-major tiles, fused)  B = X (MN-major, split in shared memory); + db
// so neither transposed activations nor a dZ tensor ever exist in HBM.  The A operand never exists in shared memory in
// split form either: the converter warps write its hi / lo parts straight into TENSOR MEMORY (tcgen05.st) and the
// MMAs read it from there (tcgen05.mma [d_tmem], [a_tmem], b_desc).
//
// Warp roles (512 threads, persistent, 1 CTA/SM):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B boxes) into a 2-4 stage shared-memory ring
//   warp 1      MMA issuer: one thread issues 12 tcgen05.mma (4 k-steps x 3 products) per stage and tcgen05.commit's the
//               stage back to the producer
//   warp 2      TMEM allocator (accumulators + a ring of 64-column A slots, one per stage)
//   warps 4-7   converters: fp32 tile -> hi / lo (act' and the bias-gradient column sums fused) -> TMEM slot;
//               DW: warps 8-11 are a second converter group (each group converts half of every k-block)
//   warps 8-15  epilogue (DW: warps 12-15): tcgen05.ld -> bias + activation -> 32x32 transpose in smem -> full-line stores
#include "tc_common.cuh"

namespace {

constexpr int kThreads = 512;  // 16 warps: producer, MMA, allocator, (idle), 4 converters, 8 epilogue
constexpr int kEpiWarps = 8;
constexpr int kBM = 128, kBK = 32;
constexpr int kTileA = kBM * kBK * 4;  // 16 KiB
constexpr int kChunk = 32 * 32 * 4;    // 4 KiB: one MN-major chunk (32 reduction rows x 32 floats)
constexpr int kMaxGroups = 2;
constexpr int kMaxStages = 6;

enum Mode { kFwd = 0, kDx = 1, kDw = 2 };

// mnr_tc_grid_cap: CTAs the NEXT launches of this host thread may use (0 = one per SM).  Two launches that do not depend on
// each other (a layer's input-gradient and weight-gradient kernels) can then run side by side on disjoint SMs from two streams.
thread_local int t_cache_hints = 0;
thread_local int t_grid_cap = 0;
int eff_sms() {
  const int s = mnr_sm_count();
  return (t_grid_cap > 0 && t_grid_cap < s) ? t_grid_cap : s;
}

// optional timeline of CTA 0 (debug / profiling aid): [role][event][tag] = clock64()
__device__ long long g_trace[4 * 256 * 2];
__device__ int g_trace_on = 0;
#define TC_TRACE(role, idx, tag)                                                      \
  do {                                                                               \
    if (g_trace_on && blockIdx.x == 0 && (idx) < 256) g_trace[((role) * 256 + (idx)) * 2 + (tag)] = clock64(); \
  } while (0)

struct alignas(64) TcGroup {
  CUtensorMap a, s, b0, b1;  // b0: B (hi, or raw when converted in-kernel), b1: B lo (pre-split modes)
  CUtensorMap cmap;          // FWD / DX with tma_store: C as [32 x 32] SWIZZLE_128B boxes
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
  int tiles_i, tiles_j;
  int tma_store;  // FWD / DX: the epilogue leaves through TMA tensor stores (C rows 16 B-multiple apart), else plain stores
  // weight-gradient kernel, second form (tc_dw2_kernel): K_in features per work item (128 or 256 = n_mt lane tiles of 128),
  // depth of the TMEM A-slot ring, 32-feature chunks of X per k-block
  int kt, n_mt, a_ring, nxc;
  // L2 cache hints (mnr_tc_cache_hints): bit 0 - the stored result is read by the next launches: evict_last; bit 1 - this launch is
  // the last reader of its dY / saved-output operands: evict_first
  int hints;
};

// work item -> (group, split, tile row, tile column); consecutive items of one CTA walk the i tiles first
struct TileCoord {
  int gi, split, i0, j0, r_begin, num_kb;
};
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int tile, int mode) {
  TileCoord c;
  const int ti = tile % p.tiles_i;
  int rest = tile / p.tiles_i;
  const int tj = rest % p.tiles_j;
  rest /= p.tiles_j;
  c.split = rest % p.splits;
  c.gi = rest / p.splits;
  c.i0 = ti * kBM, c.j0 = tj * p.BN;
  int r_begin = 0, r_end = p.R;
  if (mode == kDw) {
    r_begin = c.split * p.rows_per_split;
    r_end = min(p.R, r_begin + p.rows_per_split);
  }
  c.r_begin = r_begin;
  c.num_kb = r_end > r_begin ? (r_end - r_begin + kBK - 1) / kBK : 0;
  return c;
}

// PERSISTENT kernel: grid = min(#tiles, #SMs); every role loops over this CTA's tiles.  The fp32 accumulator is
// double-buffered in TMEM (2 x BN columns, BN <= 128) so the epilogue warps drain tile t while TMA / converters / MMA
// already work on tile t+1 (DW with a 256-wide tile keeps a single accumulator).
template <int MODE, int ACT>
__global__ void __launch_bounds__(kThreads, 1) tc_linear_kernel(const __grid_constant__ TcParams p) {
  constexpr bool A_MN = (MODE == kDw), B_MN = (MODE != kFwd), HAS_S = (MODE != kFwd) && (ACT != 0), CONV_B = (MODE == kDw);
  // (ACT == 0 in DX / DW: the incoming gradient is already dZ, e.g. behind a LayerNorm; no saved-output tile is loaded)
  // FWD / DX: the converters write the hi / lo parts of A into TENSOR MEMORY (tcgen05.st) and the MMAs read A from there
  // (tcgen05.mma [d], [a_tmem], b_desc): the split A never exists in shared memory, which halves the shared-memory
  // traffic per k-block (the pipe that bounded this kernel: 12 MMAs x (A + B) operand reads + converter writes).
  // DW: thread = output feature i (TMEM lane), the 32 reduction rows of the k-block become its 32 columns.
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_tma[kMaxStages], full_conv[kMaxStages], empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN, stages = p.stages;
  const int n_tiles = p.tiles_i * p.tiles_j * p.n_groups * p.splits;

  // stage layout (1024 B aligned).  With a saved-output tile S (DX / DW) the low part of A is written over S:
  // the converter thread that reads A[q] and S[q] writes hi -> A[q], lo -> S[q].
  // (offset the extern array itself: integer round-trips make the compiler fall back to generic LD/ST)
  uint8_t* smem = smem_dyn + ((1024u - (ptx::smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int offAlo = kTileA;  // == offS
  const int offB = HAS_S ? 2 * kTileA : kTileA;
  const int bytesB = BN * kBK * 4;
  const int offBlo = offB + bytesB;
  const int stage_bytes = offBlo + bytesB;
  float* epi_stage = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes);  // 8 warps x [32][33] floats, or (TMA
                                                                                     // store) 8 x 4 KiB swizzled [32][32] boxes
  // (TMA store, FWD) behind the 8 boxes: 32 bias floats per epilogue warp, rewritten per chunk
  float* s_bias_all = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(epi_stage) + 8 * 4096);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full_tma[s], 1);
      ptx::mbar_init(&full_conv[s], MODE == kDw ? 256 : 128);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tmem_full[b], 1);
      ptx::mbar_init(&tmem_empty[b], (MODE == kDw ? 4 : kEpiWarps) * 32);
    }
    ptx::fence_mbar_init();
  }
  const uint32_t acc_stride = (uint32_t)((BN + 31) / 32 * 32);  // TMEM columns per accumulator (tcgen05.ld reads 32 at a time)
  // BN > 128 (DW with a 256-wide weight-gradient tile: the dY / Y tiles are then loaded and converted ONCE per k-block
  // instead of once per 128-column half) leaves room for a single accumulator only: the epilogue of a work item is then
  // not overlapped with the next item's main loop, which DW (one or two items per CTA) does not need.
  const bool single_acc = BN > 128;
  const uint32_t n_acc = single_acc ? 1u : 2u;
  const uint32_t a_slots = n_acc * acc_stride;  // first TMEM column of the A ring: `stages` slots of 64 columns (hi 32 | lo 32)
  uint32_t tmem_cols = 32;
  while (tmem_cols < n_acc * acc_stride + (uint32_t)stages * 64u) tmem_cols <<= 1;
  if (warp == 2) ptx::tmem_alloc(&tmem_slot, tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // PDL: everything above touched only shared / tensor memory and overlapped the previous kernel's tail
  mnr::pdl_trigger();
  mnr::pdl_wait();
  if (warp == 3 && lane < p.n_groups) {   // (otherwise idle warp) descriptor fetches off the first loads' / stores' critical path
    const TcGroup& g = p.g[lane];
    ptx::prefetch_tmap(&g.a), ptx::prefetch_tmap(&g.b0);
    if (HAS_S) ptx::prefetch_tmap(&g.s);
    if (!CONV_B) ptx::prefetch_tmap(&g.b1);
    if (MODE != kDw && p.tma_store) ptx::prefetch_tmap(&g.cmap);
  }

  if (warp == 0) {
    // ===== TMA producer ==========================================================================
    if (ptx::elect_one()) {
      const uint32_t tx_bytes = kTileA * (HAS_S ? 2 : 1) + bytesB * (CONV_B ? 1 : 2);
      uint32_t kbg = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile, MODE);
        const TcGroup& g = p.g[tc.gi];
        for (int kb = 0; kb < tc.num_kb; ++kb, ++kbg) {
          const int st = kbg % stages;
          ptx::mbar_wait(&empty_bar[st], ((kbg / stages) & 1) ^ 1);
          TC_TRACE(0, kbg, 0);
          uint8_t* base = smem + (size_t)st * stage_bytes;
          ptx::mbar_expect_tx(&full_tma[st], tx_bytes);
          const int r0 = tc.r_begin + kb * kBK;
          if (p.hints & 2) {   // last reader of dY / Y: their lines may leave L2 first
            const uint64_t pol = ptx::policy_evict_first();
            if (!A_MN) {
              ptx::tma_load_2d_hint(base, &g.a, r0, tc.i0, &full_tma[st], pol);
              if (HAS_S) ptx::tma_load_2d_hint(base + offAlo, &g.s, r0, tc.i0, &full_tma[st], pol);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                ptx::tma_load_2d_hint(base + c * kChunk, &g.a, tc.i0 + 32 * c, r0, &full_tma[st], pol);
                if (HAS_S) ptx::tma_load_2d_hint(base + offAlo + c * kChunk, &g.s, tc.i0 + 32 * c, r0, &full_tma[st], pol);
              }
            }
          } else if (!A_MN) {  // box [128 rows][32 floats along the reduction]
            ptx::tma_load_2d(base, &g.a, r0, tc.i0, &full_tma[st]);
            if (HAS_S) ptx::tma_load_2d(base + offAlo, &g.s, r0, tc.i0, &full_tma[st]);
          } else {      // 4 chunks of [32 reduction rows][32 floats along i]
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              ptx::tma_load_2d(base + c * kChunk, &g.a, tc.i0 + 32 * c, r0, &full_tma[st]);
              if (HAS_S) ptx::tma_load_2d(base + offAlo + c * kChunk, &g.s, tc.i0 + 32 * c, r0, &full_tma[st]);
            }
          }
          if (!B_MN) {  // box [BN rows][32 floats]
            ptx::tma_load_2d(base + offB, &g.b0, r0, tc.j0, &full_tma[st]);
            ptx::tma_load_2d(base + offBlo, &g.b1, r0, tc.j0, &full_tma[st]);
          } else {
            for (int c = 0; c < BN / 32; ++c) {
              ptx::tma_load_2d(base + offB + c * kChunk, &g.b0, tc.j0 + 32 * c, r0, &full_tma[st]);
              if (!CONV_B) ptx::tma_load_2d(base + offBlo + c * kChunk, &g.b1, tc.j0 + 32 * c, r0, &full_tma[st]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ===============================================================================
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_tf32(kBM, BN, 0, B_MN ? 1 : 0);
      // B descriptors.  K-major: SWIZZLE_128B (layout 2), 8-row groups 1 KiB apart (SBO), LBO unused (16 B).
      // MN-major (32-bit): SWIZZLE_128B_BASE32B (layout 1), 4-row reduction groups 512 B apart (SBO), 32-float
      // chunks along N one 4 KiB box apart (LBO).
      const uint64_t b_desc_hi = ptx::smem_desc(0, 0, B_MN ? 512 : 1024, B_MN ? 1 : 2);
      const uint32_t b_desc_lo_const = ((B_MN ? (uint32_t)kChunk : 16u) >> 4) << 16;
      uint32_t kbg = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const TileCoord tc = decode_tile(p, tile, MODE);
        const uint32_t acc = single_acc ? 0u : (it & 1);
        ptx::mbar_wait(&tmem_empty[acc], ((single_acc ? it : (it >> 1)) & 1) ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * acc_stride;
        for (int kb = 0; kb < tc.num_kb; ++kb, ++kbg) {
          const int st = kbg % stages;
          ptx::mbar_wait(&full_conv[st], (kbg / stages) & 1);
          TC_TRACE(2, kbg, 0);
          ptx::tc_fence_after();
          const uint32_t base = ptx::smem_u32(smem + (size_t)st * stage_bytes);
          // descriptor = constant high word | (start address >> 4) in the low word: only the address changes
          const uint32_t b_lo32 = ((base + offB) >> 4) | b_desc_lo_const, blo_lo32 = ((base + offBlo) >> 4) | b_desc_lo_const;
          const uint32_t a_t = tmem_base + a_slots + (uint32_t)st * 64u;  // A ring slot: hi at +0, lo at +32
          // k-steps of 8 that hold real reduction elements: a ragged last k-block (first-layer K = 37 -> 40: one of its four
          // k-steps) does not issue MMAs over TMA's zero fill (9 of the 24 MMAs of a K = 40 tile)
          const int r_left = (MODE == kDw ? min(p.R, tc.r_begin + p.rows_per_split) : p.R) - (tc.r_begin + kb * kBK);
          const int ksteps = r_left >= kBK ? kBK / 8 : (r_left + 7) / 8;
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            if (k >= ksteps) break;
            // B K-major: step 8 floats (32 B) inside the 128 B swizzle row; MN-major: step 8 reduction rows (1 KiB)
            const uint32_t b_step = (B_MN ? k * 1024 : k * 32) >> 4;
            const uint64_t b_hi = b_desc_hi | (uint64_t)(b_lo32 + b_step), b_lo = b_desc_hi | (uint64_t)(blo_lo32 + b_step);
            ptx::mma_tf32_ts(tmem_d, a_t + 32u + 8u * k, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);  // small terms first
            ptx::mma_tf32_ts(tmem_d, a_t + 8u * k, b_lo, idesc, 1u);
            ptx::mma_tf32_ts(tmem_d, a_t + 8u * k, b_hi, idesc, 1u);
          }
          ptx::tc_commit(&empty_bar[st]);  // stage reusable once these MMAs have read it
          TC_TRACE(2, kbg, 1);
        }
        ptx::tc_commit(&tmem_full[acc]);   // accumulator complete (fires at once for an empty reduction range)
      }
    }
  } else if (warp >= 4 && warp < (MODE == kDw ? 12 : 8)) {
    // ===== converters ====================================================================================
    // DW is bound by this role (a [32][128] tile of dY and of Y per k-block, element-wise transposed into TMEM lanes by one
    // warp per scheduler), and its epilogue is small: warps 8-11 form a SECOND converter group there; the two groups share
    // every k-block (half of the reduction rows each) and each keeps its own bias-gradient partial.
    const int cgroup = (warp - 4) >> 2;  // 0: warps 4-7, 1: warps 8-11 (DW only)
    const int t = tid - 128 - 128 * cgroup;
    uint32_t kbg = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile, MODE);
      float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int kb = 0; kb < tc.num_kb; ++kb, ++kbg) {
        const int st = kbg % stages;
        ptx::mbar_wait(&full_tma[st], (kbg / stages) & 1);
        if (t == 0) TC_TRACE(1, kbg, 0);
        uint8_t* base = smem + (size_t)st * stage_bytes;
        float4* A4 = reinterpret_cast<float4*>(base);
        float4* Alo4 = reinterpret_cast<float4*>(base + offAlo);  // holds S on entry when HAS_S
        if (MODE != kDw) {
          // thread = row r of the [128][32] tile (TMEM lane r): SWIZZLE_128B keeps its 16 B chunk c at position c ^ (r & 7),
          // so the 8 lanes of an LDS.128 phase hit 8 different bank groups.  hi -> columns [0,32), lo -> [32,64) of the slot.
          const int r = t;
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int q = r * 8 + (c ^ (r & 7));
            float4 a = A4[q];
            if (HAS_S) {
              const float4 s = Alo4[q];
              a.x *= act_bwd_t<ACT>(s.x), a.y *= act_bwd_t<ACT>(s.y);
              a.z *= act_bwd_t<ACT>(s.z), a.w *= act_bwd_t<ACT>(s.w);
            }
            float4 h, l;
            split4(a, h, l);
            hi[4 * c] = __float_as_uint(h.x), hi[4 * c + 1] = __float_as_uint(h.y), hi[4 * c + 2] = __float_as_uint(h.z), hi[4 * c + 3] = __float_as_uint(h.w);
            lo[4 * c] = __float_as_uint(l.x), lo[4 * c + 1] = __float_as_uint(l.y), lo[4 * c + 2] = __float_as_uint(l.z), lo[4 * c + 3] = __float_as_uint(l.w);
          }
          const uint32_t a_t = tmem_base + a_slots + (uint32_t)st * 64u + ((uint32_t)(32 * (warp & 3)) << 16);
          ptx::tmem_st_32x32(a_t, hi);
          ptx::tmem_st_32x32(a_t + 32u, lo);
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
        } else {
          // MN-major A tile = 4 chunks of [32 reduction rows][32 features]; thread = feature i = 32 * chunk + lane.
          // SWIZZLE_128B_ATOM_32B keeps the 32 B piece (l >> 3) of row r at piece (l >> 3) ^ (r & 3): the 32 lanes of a
          // warp read one whole 128 B row per LDS -> conflict-free.  The bias gradient (column sum of dZ) is a
          // per-thread scalar.  BOTH converter groups work on every k-block - group g takes the reduction rows
          // [16 g, 16 g + 16) of the A tile (16 TMEM columns of hi and of lo) and half of the B tile's split - so the
          // conversion latency on the TMA -> convert -> MMA chain of a stage is half of what one group per k-block
          // took (measured 1.15 us per k-block at N = 128, K = 256 against 1.13 us of MMAs, three 64 KB stages deep).
          const int c = warp & 3, l = lane;
          const float* Af = reinterpret_cast<const float*>(base) + c * (kChunk / 4);
          const float* Sf = reinterpret_cast<const float*>(base + offAlo) + c * (kChunk / 4);
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const int r = 16 * cgroup + rr;
            const int q = r * 32 + ((((l >> 3) ^ (r & 3)) << 3) | (l & 7));
            const float a = HAS_S ? Af[q] * act_bwd_t<ACT>(Sf[q]) : Af[q];
            bsum.x += a;
            const float h = tf32_hi(a);
            hi[rr] = __float_as_uint(h), lo[rr] = __float_as_uint(a - h);
          }
          const uint32_t a_t = tmem_base + a_slots + (uint32_t)st * 64u + ((uint32_t)(32 * c) << 16) + 16u * (uint32_t)cgroup;
          ptx::tmem_st_32x16(a_t, hi);
          ptx::tmem_st_32x16(a_t + 32u, lo);
          float4* B4 = reinterpret_cast<float4*>(base + offB);
          float4* Blo4 = reinterpret_cast<float4*>(base + offBlo);
          for (int q = t + 128 * cgroup; q < bytesB / 16; q += 256) {
            float4 h4, l4;
            split4(B4[q], h4, l4);
            B4[q] = h4, Blo4[q] = l4;
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();
        }
        ptx::mbar_arrive(&full_conv[st]);
        if (t == 0) TC_TRACE(1, kbg, 1);
      }
      if (MODE == kDw && p.want_bias && tc.j0 == 0) {  // bias partials: [group][split][converter group][I]
        float* bp = p.partials + (size_t)p.n_groups * p.splits * p.I * p.J + (((size_t)tc.gi * p.splits + tc.split) * 2 + cgroup) * p.I;
        if (tc.i0 + t < p.I) bp[tc.i0 + t] = bsum.x;
      }
    }
  } else if (warp >= (MODE == kDw ? 12 : 8)) {
    // ===== epilogue: TMEM -> registers -> (bias, act) -> 32x32 transpose in smem -> full 128 B line stores ========
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*quarter, 32*quarter+32)
    constexpr int kEpi = (MODE == kDw) ? 4 : kEpiWarps, kEpi0 = (MODE == kDw) ? 12 : 8;
    float* stg = epi_stage + (size_t)(warp - kEpi0) * 32 * 33;
    const int half = (warp - kEpi0) >> 2;  // the warps of a TMEM quarter take alternate 32-column chunks
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const TileCoord tc = decode_tile(p, tile, MODE);
      const TcGroup& g = p.g[tc.gi];
      const uint32_t acc = single_acc ? 0u : (it & 1);
      // FWD: this warp's bias values (one column per lane, <= 2 chunks since BN <= 128) are requested BEFORE the wait on
      // the accumulator: the global-load latency used to sit on every chunk's critical path (top stall of the kernel)
      float bias_c[2] = {0.f, 0.f};
      if (MODE == kFwd && g.bias) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int cc = half + k * (kEpi / 4);
          const int col = tc.j0 + cc * 32 + lane;
          if (cc * 32 + lane < BN && col < p.J) bias_c[k] = __ldg(g.bias + col);
        }
      }
      ptx::mbar_wait(&tmem_full[acc], (single_acc ? it : (it >> 1)) & 1);
      if (tid == kEpi0 * 32) TC_TRACE(3, it, 0);
      ptx::tc_fence_after();
      float* cbase;
      size_t ldo;
      if (MODE == kDw) cbase = p.partials + ((size_t)tc.gi * p.splits + tc.split) * p.I * p.J, ldo = p.J;
      else cbase = g.c, ldo = p.ldc;
      const int row_base = tc.i0 + 32 * quarter;
      if (MODE != kDw && p.tma_store) {
        // ---- TMA-store epilogue: lane = output row keeps its 32 accumulator columns, applies bias + activation on the
        // registers (bias columns from this warp's shared-memory row, broadcast LDS.128), parks the row as eight 16 B pieces
        // in a [32 rows][128 B] box with the 128 B swizzle (piece c of row r at c ^ (r & 7): conflict-free), and ONE
        // cp.async.bulk.tensor store moves the box to C.  ~215 instructions and 32 LSU wavefronts per 32x32 chunk against
        // ~430 and 96 for the transpose + full-line stores below; the box is reused once the previous store has read it.
        const uint32_t box = ptx::smem_u32(reinterpret_cast<uint8_t*>(epi_stage) + (size_t)(warp - kEpi0) * 4096);
        float* s_bias = s_bias_all + (warp - kEpi0) * 32;
        for (int cc = half; cc < (BN + 31) / 32; cc += kEpi / 4) {
          if (MODE == kFwd) {   // this chunk's bias columns -> the warp's shared-memory row (read back as broadcast LDS.128)
            __syncwarp();
            s_bias[lane] = cc == half ? bias_c[0] : bias_c[1];
            __syncwarp();
          }
          uint32_t r[32];
          if (tc.num_kb > 0) {
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(32 * quarter) << 16) + acc * acc_stride + (uint32_t)(cc * 32), r);
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) r[q] = 0u;
          }
          float v[32];
          if (MODE == kFwd) {
            const float4* b4 = reinterpret_cast<const float4*>(s_bias);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              const float4 b = b4[q4];
              v[4 * q4] = act_fwd_t<ACT>(__uint_as_float(r[4 * q4]) + b.x), v[4 * q4 + 1] = act_fwd_t<ACT>(__uint_as_float(r[4 * q4 + 1]) + b.y);
              v[4 * q4 + 2] = act_fwd_t<ACT>(__uint_as_float(r[4 * q4 + 2]) + b.z), v[4 * q4 + 3] = act_fwd_t<ACT>(__uint_as_float(r[4 * q4 + 3]) + b.w);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]);
          }
          if (ptx::elect_one()) ptx::bulk_wait_read0();   // the previous store of this warp has finished reading the box
          __syncwarp();
          const uint32_t row_addr = box + (uint32_t)lane * 128u;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            ptx::st_shared_v4(row_addr + (uint32_t)((c ^ (lane & 7)) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (ptx::elect_one()) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(epi_stage) + (size_t)(warp - kEpi0) * 4096;
            if (p.hints & 1) ptx::tma_store_2d_hint(&g.cmap, src, tc.j0 + cc * 32, row_base, ptx::policy_evict_last());
            else ptx::tma_store_2d(&g.cmap, src, tc.j0 + cc * 32, row_base);
            ptx::bulk_commit();
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tmem_empty[acc]);
        if (tid == kEpi0 * 32) TC_TRACE(3, it, 1);
        continue;
      }
      for (int cc = half; cc < (BN + 31) / 32; cc += kEpi / 4) {
        uint32_t r[32];
        if (tc.num_kb > 0) {
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(32 * quarter) << 16) + acc * acc_stride + (uint32_t)(cc * 32), r);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) r[q] = 0u;
        }
        const int jb = tc.j0 + cc * 32;
        // lane = row: bias + activation applied here would need a per-column bias per register; instead park the
        // raw row (pitch 33: conflict-free both ways) and finish per column below
#pragma unroll
        for (int q = 0; q < 32; ++q) stg[lane * 33 + q] = __uint_as_float(r[q]);
        __syncwarp();
        // (Round 2 measured the alternative without this transpose - lane = row keeps its 32 columns and writes them as four
        // 256-bit stores, ~6 instead of ~10 instructions per element: 6.05 ms per update against 5.61 ms.  Every such store
        // instruction touches 32 different 128 B lines = 32 LSU wavefronts, 128 per chunk against 96 here: the epilogue is
        // bound by LSU wavefronts, not by issue slots.  profiles/r02_summary.md.)
        // lane = column: every store instruction writes one full 128 B line of one output row.  Batches of 8 rows:
        // 8 independent LDS, then the math, then 8 stores (a guarded one-row-at-a-time loop runs as one long
        // dependent chain: ~100 cycles per row).
        const int col = jb + lane;
        const bool col_ok = col < p.J && cc * 32 + lane < BN;
        const float bias = cc == half ? bias_c[0] : bias_c[1];
        const int rows_valid = min(32, p.I - row_base);
        float* out = cbase + (size_t)row_base * ldo + col;
        if (rows_valid == 32) {
#pragma unroll
          for (int r0 = 0; r0 < 32; r0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = stg[(r0 + u) * 33 + lane];
            if (MODE == kFwd) {
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = act_fwd_t<ACT>(v[u] + bias);
            }
            if (col_ok) {
#pragma unroll
              for (int u = 0; u < 8; ++u) out[(size_t)(r0 + u) * ldo] = v[u];
            }
          }
        } else {
          for (int rr = 0; rr < rows_valid; ++rr) {
            float v = stg[rr * 33 + lane];
            if (MODE == kFwd) v = act_fwd_t<ACT>(v + bias);
            if (col_ok) out[(size_t)rr * ldo] = v;
          }
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[acc]);
      if (tid == kEpi0 * 32) TC_TRACE(3, it, 1);
    }
  }

  if (MODE != kDw && p.tma_store && warp >= 8) {   // this warp's C boxes have landed
    if (ptx::elect_one()) ptx::bulk_wait0();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- weight gradient, second form -------------------------------------------------------------------------------------
// dW (N_out, K_in) = dZ^T X with dZ = dY * act'(Y), computed as  dW^T = X^T dZ  so that the operand roles swap:
//   * the RAW operand X (a [32 rows][K_in] block per k-block) goes to TENSOR MEMORY - the converter warps transpose it on the
//     way (TMEM lane = input feature, the 32 rows of the k-block = 32 columns; hi | lo slots as in the other modes);
//   * the operand that needs the product is converted IN PLACE in shared memory: the hi part of dZ over the dY tile, the lo
//     part over the Y tile, element for element - and that pair is the MN-major B operand of the MMAs (N = N_out).
// A stage therefore holds exactly the bytes TMA brought and nothing else (the first form keeps a second copy of X for its lo
// part: 96 KB per stage at 128 x 256, two stages, a TMA -> convert -> MMA chain of 3.8 us per two k-blocks against 2.3 us of
// MMAs).  Here 128 x 256 runs three 64 KB stages, 64 x 128 six of 32 KB, 256 x 40 three of 72 KB, the transposition handles
// K_in values instead of N_out, and dY / Y are read once per K_in tile of 256 instead of once per 128 output rows.
// The accumulator tile [K_in lanes][N_out columns] leaves transposed - every store instruction writes 32 consecutive floats
// of one dW row - into the same [split][N_out][K_in] partial slabs the first form writes, so tc_reduce_kernel is unchanged.
// Roles: warp 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-11 converters (256 threads), 12-15 epilogue.
struct Dw2Coord {
  int gi, split, i0, j0, r_begin, num_kb, nxc, n_mt;
};
__device__ __forceinline__ Dw2Coord decode_dw2(const TcParams& p, int tile) {
  Dw2Coord c;
  const int ti = tile % p.tiles_i;
  int rest = tile / p.tiles_i;
  const int tj = rest % p.tiles_j;
  rest /= p.tiles_j;
  c.split = rest % p.splits;
  c.gi = rest / p.splits;
  c.i0 = ti * p.kt, c.j0 = tj * p.BN;
  c.r_begin = c.split * p.rows_per_split;
  const int r_end = min(p.R, c.r_begin + p.rows_per_split);
  c.num_kb = r_end > c.r_begin ? (r_end - c.r_begin + kBK - 1) / kBK : 0;
  c.nxc = (min(p.kt, p.J - c.i0) + 31) / 32;
  c.n_mt = (c.nxc + 3) / 4;
  return c;
}

template <int ACT>
__global__ void __launch_bounds__(kThreads, 1) tc_dw2_kernel(const __grid_constant__ TcParams p) {
  constexpr bool HAS_S = ACT != 0;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_tma[kMaxStages], full_conv[kMaxStages], empty_bar[kMaxStages], a_empty[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full, tmem_empty;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN, stages = p.stages, a_ring = p.a_ring;
  const int n_tiles = p.tiles_i * p.tiles_j * p.n_groups * p.splits;
  uint8_t* smem = smem_dyn + ((1024u - (ptx::smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int bytesZ = BN * kBK * 4;             // dY tile (-> dZ hi) and Y tile (-> dZ lo): BN / 32 chunks of 4 KiB each
  const int offS = bytesZ, offX = 2 * bytesZ;
  const int stage_bytes = offX + p.nxc * kChunk;
  float* s_bias = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes);   // [8 converter warps][256]

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full_tma[s], 1);
      ptx::mbar_init(&full_conv[s], 256);
      ptx::mbar_init(&empty_bar[s], 1);
      ptx::mbar_init(&a_empty[s], 1);
    }
    ptx::mbar_init(&tmem_full, 1);
    ptx::mbar_init(&tmem_empty, 4 * 32);
    ptx::fence_mbar_init();
  }
  // tensor memory: n_mt accumulators of BN columns (lane tile mt = features [128 mt, 128 mt + 128)), then the A ring:
  // a_ring k-blocks x n_mt lane tiles x (32 hi | 32 lo) columns
  const uint32_t a_base = (uint32_t)(p.n_mt * BN);
  uint32_t tmem_cols = 32;
  while (tmem_cols < a_base + (uint32_t)(a_ring * p.n_mt) * 64u) tmem_cols <<= 1;
  if (warp == 2) ptx::tmem_alloc(&tmem_slot, tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  mnr::pdl_trigger();
  mnr::pdl_wait();
  if (warp == 3 && lane < p.n_groups) {
    const TcGroup& g = p.g[lane];
    ptx::prefetch_tmap(&g.a), ptx::prefetch_tmap(&g.b0);
    if (HAS_S) ptx::prefetch_tmap(&g.s);
  }

  if (warp == 0) {
    // ===== TMA producer: MN-major boxes [32 reduction rows][32 floats] ================================================
    if (ptx::elect_one()) {
      uint32_t kbg = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const Dw2Coord tc = decode_dw2(p, tile);
        const TcGroup& g = p.g[tc.gi];
        const uint32_t tx_bytes = (uint32_t)bytesZ * (HAS_S ? 2u : 1u) + (uint32_t)tc.nxc * kChunk;
        for (int kb = 0; kb < tc.num_kb; ++kb, ++kbg) {
          const int st = kbg % stages;
          ptx::mbar_wait(&empty_bar[st], ((kbg / stages) & 1) ^ 1);
          TC_TRACE(0, kbg, 0);
          uint8_t* base = smem + (size_t)st * stage_bytes;
          ptx::mbar_expect_tx(&full_tma[st], tx_bytes);
          const int r0 = tc.r_begin + kb * kBK;
          for (int c = 0; c < tc.nxc; ++c) ptx::tma_load_2d(base + offX + c * kChunk, &g.b0, tc.i0 + 32 * c, r0, &full_tma[st]);
          for (int c = 0; c < BN / 32; ++c) {
            ptx::tma_load_2d(base + c * kChunk, &g.a, tc.j0 + 32 * c, r0, &full_tma[st]);
            if (HAS_S) ptx::tma_load_2d(base + offS + c * kChunk, &g.s, tc.j0 + 32 * c, r0, &full_tma[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: D[mt] (+)= A[mt] (TMEM: X^T hi | lo) * B (smem: dZ hi / lo, MN-major) ===========================
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_tf32(kBM, BN, 0, 1);
      const uint64_t b_desc_hi = ptx::smem_desc(0, 0, 512, 1);
      const uint32_t b_desc_lo_const = ((uint32_t)kChunk >> 4) << 16;
      uint32_t kbg = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const Dw2Coord tc = decode_dw2(p, tile);
        ptx::mbar_wait(&tmem_empty, (it & 1) ^ 1);   // the epilogue has drained the accumulators
        ptx::tc_fence_after();
        for (int kb = 0; kb < tc.num_kb; ++kb, ++kbg) {
          const int st = kbg % stages, ar = kbg % a_ring;
          ptx::mbar_wait(&full_conv[st], (kbg / stages) & 1);
          TC_TRACE(2, kbg, 0);
          ptx::tc_fence_after();
          const uint32_t base = ptx::smem_u32(smem + (size_t)st * stage_bytes);
          const uint32_t bh32 = (base >> 4) | b_desc_lo_const, bl32 = ((base + offS) >> 4) | b_desc_lo_const;
          const int r_left = min(p.R, tc.r_begin + p.rows_per_split) - (tc.r_begin + kb * kBK);
          const int ksteps = r_left >= kBK ? kBK / 8 : (r_left + 7) / 8;
          for (int mt = 0; mt < tc.n_mt; ++mt) {
            const uint32_t tmem_d = tmem_base + (uint32_t)(mt * BN);
            const uint32_t a_t = tmem_base + a_base + (uint32_t)(ar * p.n_mt + mt) * 64u;
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
              if (k >= ksteps) break;
              const uint32_t b_step = (uint32_t)(k * 1024) >> 4;   // 8 reduction rows = 1 KiB inside every 4 KiB chunk
              const uint64_t b_hi = b_desc_hi | (uint64_t)(bh32 + b_step), b_lo = b_desc_hi | (uint64_t)(bl32 + b_step);
              ptx::mma_tf32_ts(tmem_d, a_t + 32u + 8u * k, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);  // small terms first
              ptx::mma_tf32_ts(tmem_d, a_t + 8u * k, b_lo, idesc, 1u);
              ptx::mma_tf32_ts(tmem_d, a_t + 8u * k, b_hi, idesc, 1u);
            }
          }
          ptx::tc_commit(&empty_bar[st]);   // the stage's shared memory may be refilled
          ptx::tc_commit(&a_empty[ar]);     // ... and its A slots rewritten
          TC_TRACE(2, kbg, 1);
        }
        ptx::tc_commit(&tmem_full);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===== converters =====================================================================================================
    const int t = tid - 128;                       // 0..255
    const int quarter = warp & 3, grp = (warp - 4) >> 2;
    uint32_t kbg = 0;
    float4 bsum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bsum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const Dw2Coord tc = decode_dw2(p, tile);
      for (int kb = 0; kb < tc.num_kb; ++kb, ++kbg) {
        const int st = kbg % stages, ar = kbg % a_ring;
        ptx::mbar_wait(&full_tma[st], (kbg / stages) & 1);
        ptx::mbar_wait(&a_empty[ar], ((kbg / a_ring) & 1) ^ 1);   // the MMAs that read this A slot have completed
        ptx::tc_fence_after();
        if (t == 0) TC_TRACE(1, kbg, 0);
        uint8_t* base = smem + (size_t)st * stage_bytes;
        // (a) X^T -> tensor memory.  Thread = input feature 32 * chunk + lane; the chunk is [32 rows][32 floats] with
        //     SWIZZLE_128B_ATOM_32B (the 32 B piece (l >> 3) of row r sits at piece (l >> 3) ^ (r & 3)): a warp reads one whole
        //     128 B row per LDS.  Two lane tiles: converter group g owns lane tile g; one: the groups split the 32 rows.
        const uint32_t a_slot = tmem_base + a_base + (uint32_t)(ar * p.n_mt) * 64u + ((uint32_t)(32 * quarter) << 16);
        if (tc.n_mt == 2) {
          const int chunk = grp * 4 + quarter;
          if (chunk < tc.nxc) {
            const float* Xf = reinterpret_cast<const float*>(base + offX) + chunk * (kChunk / 4);
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const float x = Xf[r * 32 + ((((lane >> 3) ^ (r & 3)) << 3) | (lane & 7))];
              const float h = tf32_hi(x);
              hi[r] = __float_as_uint(h), lo[r] = __float_as_uint(x - h);
            }
            ptx::tmem_st_32x32(a_slot + (uint32_t)grp * 64u, hi);
            ptx::tmem_st_32x32(a_slot + (uint32_t)grp * 64u + 32u, lo);
          }
        } else if (quarter < tc.nxc) {
          const float* Xf = reinterpret_cast<const float*>(base + offX) + quarter * (kChunk / 4);
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const int r = 16 * grp + rr;
            const float x = Xf[r * 32 + ((((lane >> 3) ^ (r & 3)) << 3) | (lane & 7))];
            const float h = tf32_hi(x);
            hi[rr] = __float_as_uint(h), lo[rr] = __float_as_uint(x - h);
          }
          ptx::tmem_st_32x16(a_slot + 16u * (uint32_t)grp, hi);
          ptx::tmem_st_32x16(a_slot + 32u + 16u * (uint32_t)grp, lo);
        }
        // (b) dZ = dY * act'(Y) -> (hi over dY, lo over Y), element-wise: the swizzle is the same for both tiles.  A thread
        //     keeps the same 4 logical columns in every chunk and every k-block: its running sums are the bias gradient.
        float4* Z4 = reinterpret_cast<float4*>(base);
        float4* S4 = reinterpret_cast<float4*>(base + offS);
        const int nq = BN * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int q = t + 256 * j;
          if (q < nq) {
            float4 a = Z4[q];
            if (HAS_S) {
              const float4 s = S4[q];
              a.x *= act_bwd_t<ACT>(s.x), a.y *= act_bwd_t<ACT>(s.y), a.z *= act_bwd_t<ACT>(s.z), a.w *= act_bwd_t<ACT>(s.w);
            }
            bsum[j].x += a.x, bsum[j].y += a.y, bsum[j].z += a.z, bsum[j].w += a.w;
            float4 h, l;
            split4(a, h, l);
            Z4[q] = h, S4[q] = l;
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&full_conv[st]);
        if (t == 0) TC_TRACE(1, kbg, 1);
      }
      // bias gradient of this work item: thread (row r = t >> 3, 16 B position t & 7) holds logical columns
      // 8 * ((pos >> 1) ^ (r & 3)) + 4 * (pos & 1) + [0, 4) of every chunk; the 4 rows of a warp meet by two xor shuffles
      // (lane ^ 10, lane ^ 20 keep the logical piece), the 8 warps in shared memory, summed in warp order.
      if (p.want_bias && tc.i0 == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j * 32 < BN) {
            float4 v = bsum[j];
            v.x += __shfl_xor_sync(0xffffffffu, v.x, 10), v.y += __shfl_xor_sync(0xffffffffu, v.y, 10);
            v.z += __shfl_xor_sync(0xffffffffu, v.z, 10), v.w += __shfl_xor_sync(0xffffffffu, v.w, 10);
            v.x += __shfl_xor_sync(0xffffffffu, v.x, 20), v.y += __shfl_xor_sync(0xffffffffu, v.y, 20);
            v.z += __shfl_xor_sync(0xffffffffu, v.z, 20), v.w += __shfl_xor_sync(0xffffffffu, v.w, 20);
            if (lane < 8) *reinterpret_cast<float4*>(s_bias + (warp - 4) * 256 + j * 32 + (lane >> 1) * 8 + (lane & 1) * 4) = v;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (t < BN) {
          float sum = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) sum += s_bias[w * 256 + t];
          float* bp = p.partials + (size_t)p.n_groups * p.splits * p.I * p.J + ((size_t)tc.gi * p.splits + tc.split) * 2 * p.I;
          if (tc.j0 + t < p.I) bp[tc.j0 + t] = sum, bp[p.I + tc.j0 + t] = 0.f;   // (the first form writes two slabs per split)
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) bsum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else if (warp >= 12) {
    // ===== epilogue: accumulator lane = input feature k, columns = output features n -> partial[n][k] ===================
    const int quarter = warp & 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const Dw2Coord tc = decode_dw2(p, tile);
      ptx::mbar_wait(&tmem_full, it & 1);
      if (tid == 12 * 32) TC_TRACE(3, it, 0);
      ptx::tc_fence_after();
      float* slab = p.partials + ((size_t)tc.gi * p.splits + tc.split) * p.I * p.J;
      for (int mt = 0; mt < tc.n_mt; ++mt) {
        const int k_first = tc.i0 + mt * 128 + 32 * quarter;
        if (k_first >= p.J) continue;   // warp-uniform
        const int kf = k_first + lane;
        for (int cc = 0; cc < BN / 32; ++cc) {
          uint32_t r[32];
          if (tc.num_kb > 0) {
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(mt * BN + cc * 32), r);
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) r[q] = 0u;
          }
          if (kf < p.J) {
            float* out = slab + (size_t)(tc.j0 + cc * 32) * p.J + kf;
            const int n_valid = min(32, p.I - (tc.j0 + cc * 32));
            if (n_valid == 32) {
#pragma unroll
              for (int q = 0; q < 32; ++q) out[(size_t)q * p.J] = __uint_as_float(r[q]);
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q)
                if (q < n_valid) out[(size_t)q * p.J] = __uint_as_float(r[q]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty);
      if (tid == 12 * 32) TC_TRACE(3, it, 1);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dW / db = sum over splits (fixed order)
// block = 32 element slots x 8 split slices: every thread sums splits z = slice, slice + 8, .. (independent loads, all in
// flight), the 8 slices meet in shared memory and are added in slice order -> deterministic, one latency round.
// blockIdx.y = (job, group): up to 4 layers' partials are reduced by ONE launch.  blockIdx.x < wblocks: 32 float4 of
// dW;  else: 32 floats of db.
struct ReduceJob {
  float* w[kMaxGroups];
  float* b[kMaxGroups];
  const float* partials;
  int n_groups, splits, I, J, wblocks;
};
constexpr int kMaxReduceJobs = 4;
struct ReduceJobs {
  ReduceJob j[kMaxReduceJobs];
  int fold_rows;   // blockIdx.y < fold_rows: deferred folds (first, so that their blocks start in the grid's first wave)
};

// Deferred folds (include/mineral_b200.h, mnr_fold_defer): out[o] = scale * sum over the n_parts partial rows of column o, for
// the columns a segment names.  Block = 8 columns (one 32 B sector per row) x 32 slices of the rows (row x = slice,
// slice + 32, ..: <= 10 independent loads at 296 CTAs, all in flight), the slices added in order through shared memory.
// blockIdx.y = fold index; the row after the last fold is the heads kernel's metrics job (one block); then the split reduce's rows.
// __noinline__: the split reduce's own blocks must not pay (cold instruction cache) for code they never run - inlined, the
// kernel grew 2.7x and its launch from 12.3 to 14.4 us.
struct FoldJobs {
  mnr_fold_job_t f[MNR_MAX_FOLDS];
  mnr_heads_metrics_job_t m;
  int n, has_m;
};

constexpr int kFoldCols = 8, kFoldSlices = 256 / kFoldCols;
__device__ __noinline__ void fold_block(const FoldJobs& fj, int fi, float4 (*s_part)[32]) {
  if (fi < fj.n && fj.f[fi]._pad == 1) {
    // float4 form (width % 4 == 0, 16 B-aligned partial rows - LayerNorm's [grid][2N] slabs): 8 float4 = one 128 B line per row
    const mnr_fold_job_t& jb = fj.f[fi];
    const int col = threadIdx.x % kFoldCols, slice = threadIdx.x / kFoldCols;
    const int o = (blockIdx.x * kFoldCols + col) * 4;
    if ((int)blockIdx.x * kFoldCols * 4 >= jb.width) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o < jb.width) {
      const float* src = jb.parts + o;
#pragma unroll 10
      for (int x = slice; x < jb.n_parts; x += kFoldSlices) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)x * jb.width));
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
    }
    float4* s = reinterpret_cast<float4*>(s_part);   // [32 slices][8 columns]
    s[slice * kFoldCols + col] = acc;
    __syncthreads();
    if (slice != 0 || o >= jb.width) return;
#pragma unroll 8
    for (int q = 1; q < kFoldSlices; ++q) {
      const float4 v = s[q * kFoldCols + col];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    const float av[4] = {acc.x, acc.y, acc.z, acc.w};
    for (int k = 0; k < jb.n_segs; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (o + e >= jb.seg[k].begin && o + e < jb.seg[k].end) jb.seg[k].dst[o + e - jb.seg[k].begin] = av[e] * jb.seg[k].scale;
    return;
  }
  if (fi < fj.n) {
    const mnr_fold_job_t& jb = fj.f[fi];
    const int col = threadIdx.x % kFoldCols, slice = threadIdx.x / kFoldCols;
    const int o = blockIdx.x * kFoldCols + col;
    if ((int)blockIdx.x * kFoldCols >= jb.width) return;
    float acc = 0.f;
    if (o < jb.width) {
      const float* src = jb.parts + o;
#pragma unroll 10
      for (int x = slice; x < jb.n_parts; x += kFoldSlices) acc += __ldcg(src + (size_t)x * jb.width);
    }
    float* s = reinterpret_cast<float*>(s_part);   // [32 slices][8 columns]
    s[slice * kFoldCols + col] = acc;
    __syncthreads();
    if (slice != 0 || o >= jb.width) return;
#pragma unroll
    for (int q = 1; q < kFoldSlices; ++q) acc += s[q * kFoldCols + col];
    for (int k = 0; k < jb.n_segs; ++k)
      if (o >= jb.seg[k].begin && o < jb.seg[k].end) jb.seg[k].dst[o - jb.seg[k].begin] = acc * jb.seg[k].scale;
    return;
  }
  if (fi != fj.n || !fj.has_m || blockIdx.x != 0) return;
  // the ten loss metrics (csrc/heads_fused.cu, CTA-0 tail): partial rows of 12 doubles {a, c, b, h, clip, kl, n_d, S_d, Q_d,
  // n_r, S_r, Q_r}; thread = (value k = tid & 15, slice = tid >> 4), fixed order
  const mnr_heads_metrics_job_t& m = fj.m;
  double* s_red = reinterpret_cast<double*>(s_part);   // 256 doubles = 2 KB of the 4 KB buffer
  __shared__ double s_fin[16];
  const int k = threadIdx.x & 15, sl = threadIdx.x >> 4;
  double acc = 0.0;
  if (k < 12) {
#pragma unroll 8
    for (int x = sl; x < m.n_parts; x += 16) acc += __ldcg(m.loss_parts + (size_t)x * 12 + k);
  }
  s_red[sl * 16 + k] = acc;
  __syncthreads();
  if (threadIdx.x < 12) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) t += s_red[q * 16 + threadIdx.x];
    s_fin[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double inv_n = 1.0 / (double)m.M;
    const double m2_d = s_fin[8] - s_fin[7] * s_fin[7] / s_fin[6], m2_r = s_fin[11] - s_fin[10] * s_fin[10] / s_fin[9];
    const float av = (float)(s_fin[0] * inv_n), c = (float)(s_fin[1] * inv_n), b = (float)(s_fin[2] * inv_n), h = (float)(s_fin[3] * inv_n);
    float out[MNR_PPO_METRICS];
    out[0] = av + 0.5f * c * m.critic_coef - h * m.entropy_coef + b * m.bounds_loss_coef;
    out[1] = av, out[2] = c, out[3] = b, out[4] = h;
    out[5] = (float)(s_fin[4] * inv_n);
    const float var_d = (float)(m2_d / (s_fin[6] - 1.0)), var_r = (float)(m2_r / (s_fin[9] - 1.0));
    out[6] = fminf(fmaxf(1.f - var_d / (var_r + 1e-8f), 0.f), 1.f);
    out[7] = (float)(s_fin[5] * inv_n);
    out[8] = 0.f, out[9] = 0.f;
    for (int q = 0; q < MNR_PPO_METRICS; ++q) m.metrics[q] = out[q] * m.metric_scale;
  }
}

__global__ void __launch_bounds__(256, 8) tc_reduce_kernel(const __grid_constant__ ReduceJobs jobs, const __grid_constant__ FoldJobs folds) {
  __shared__ float4 s_part[8][32];
  mnr::pdl_trigger();
  mnr::pdl_wait();
  if ((int)blockIdx.y < jobs.fold_rows) {
    fold_block(folds, (int)blockIdx.y, s_part);
    return;
  }
  const int by = (int)blockIdx.y - jobs.fold_rows;
  const ReduceJob& jb = jobs.j[by / kMaxGroups];
  const int gi = by % kMaxGroups;
  if (gi >= jb.n_groups || (int)blockIdx.x >= jb.wblocks + (jb.I + 31) / 32) return;
  const int splits = jb.splits, I = jb.I, wblocks = jb.wblocks;
  const int NK = I * jb.J;
  const float* __restrict__ partials = jb.partials;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((int)blockIdx.x < wblocks) {
    const int e4 = blockIdx.x * 32 + lane;
    if (e4 * 4 < NK) {
      const float4* src = reinterpret_cast<const float4*>(partials + (size_t)gi * splits * NK) + e4;
      const size_t step = (size_t)NK / 4;
#pragma unroll 4
      for (int z = slice; z < splits; z += 8) {
        const float4 v = __ldcs(src + (size_t)z * step);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
    }
  } else {
    const int n = ((int)blockIdx.x - wblocks) * 32 + lane;
    if (n < I && jb.b[gi]) {
      const float* bp = partials + (size_t)jb.n_groups * splits * NK + (size_t)gi * 2 * splits * I + n;  // 2 converter groups
#pragma unroll 4
      for (int z = slice; z < 2 * splits; z += 8) acc.x += __ldcs(bp + (size_t)z * I);
    }
  }
  s_part[slice][lane] = acc;
  __syncthreads();
  if (slice != 0) return;
#pragma unroll
  for (int q = 1; q < 8; ++q) {
    const float4 v = s_part[q][lane];
    acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
  }
  if ((int)blockIdx.x < wblocks) {
    const int e4 = blockIdx.x * 32 + lane;
    if (e4 * 4 < NK) reinterpret_cast<float4*>(jb.w[gi])[e4] = acc;
  } else {
    const int n = ((int)blockIdx.x - wblocks) * 32 + lane;
    if (n < I && jb.b[gi]) jb.b[gi][n] = acc.x;
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
  mnr::pdl_trigger();
  mnr::pdl_wait();
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
constexpr int kEpiStageBytes = 8 * 32 * 33 * 4;   // >= 8 x 4 KiB boxes of the TMA-store epilogue

// MNR_TC_SMALLK_BN64=1: 64-column tiles for short reductions.  MEASURED slower (round 2, whole update at cfg2: 5.76 ms against
// 5.60 ms): twice the tiles re-read the A rows twice as often and the per-tile fixed costs double; kept as an opt-in knob.
int small_k_bn64() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MNR_TC_SMALLK_BN64");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

// MNR_TC_TMA_STORE=0 selects the transpose + full-line-store epilogue (A/B measurements); read once
int tma_store_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MNR_TC_TMA_STORE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}

int stage_bytes_of(int mode, int act, int BN) { return ((mode == kFwd || act == 0) ? 1 : 2) * kTileA + 2 * BN * kBK * 4; }

int pick_stages(int mode, int act, int BN) {
  const int stage = stage_bytes_of(mode, act, BN);
  int s = (227 * 1024 - 2048 - kEpiStageBytes) / stage;
  if (s > kMaxStages) s = kMaxStages;
  // tensor memory: accumulators (two of BN columns, one above 128) + one 64-column A slot per stage <= 512 columns
  const int acc_cols = (BN > 128 ? 1 : 2) * ((BN + 31) / 32 * 32);
  const int tmem_s = (512 - acc_cols) / 64;
  if (s > tmem_s) s = tmem_s;
  return s;
}

template <int MODE>
int launch(TcParams& p, dim3 tiles, cudaStream_t stream) {
  p.tiles_i = (int)tiles.x, p.tiles_j = (int)tiles.y;
  const int n_tiles = (int)(tiles.x * tiles.y * tiles.z);
  const int sms = eff_sms();
  dim3 grid(n_tiles < sms ? n_tiles : sms);  // persistent: one CTA per SM walks the tiles
  const int stage = stage_bytes_of(MODE, p.act, p.BN);
  const size_t smem = (size_t)p.stages * stage + kEpiStageBytes + 1024;
  cudaError_t e = cudaSuccess;
#define MNR_TC_LAUNCH(ACTV)                                                                                              \
  e = mnr_allow_max_dyn_smem(tc_linear_kernel<MODE, ACTV>);                                                                   \
  if (e == cudaSuccess) e = mnr_launch_pdl(tc_linear_kernel<MODE, ACTV>, grid, dim3(kThreads), smem, stream, p);
  switch (p.act) {
    case 1: MNR_TC_LAUNCH(1) break;
    case 2: MNR_TC_LAUNCH(2) break;
    default: MNR_TC_LAUNCH(0) break;
  }
#undef MNR_TC_LAUNCH
  if (e != cudaSuccess) return (int)e;
  e = cudaPeekAtLastError();
  return e == cudaSuccess ? MNR_OK : (int)e;
}

int ld_or(int ld, int dflt) { return ld > 0 ? ld : dflt; }

}  // namespace

extern "C" int mnr_tc_cache_hints(int flags) {
  const int prev = t_cache_hints;
  t_cache_hints = flags;
  return prev;
}

extern "C" int mnr_tc_grid_cap(int ctas) {
  const int prev = t_grid_cap;
  t_grid_cap = ctas > 0 ? ctas : 0;
  return prev;
}

extern "C" int mnr_tc_supported(void) { return encode_fn() != nullptr ? 1 : 0; }

extern "C" int mnr_tc_linear_fwd(const mnr_linear_group_t* groups, int n_groups, int M, int act, mnr_stream_t stream) {
  if (n_groups <= 0 || n_groups > kMaxGroups) return MNR_ESHAPE;
  MNR_CHECK_PTR(groups);
  if (M <= 0) return MNR_ESHAPE;
  if (act < 0 || act > 2) return MNR_EUNSUP;
  const int N = groups[0].N, K = groups[0].K;
  if (N <= 0 || K <= 0 || N % 16 != 0) return MNR_ESHAPE;
  TcParams p;
  int tma_store = tma_store_enabled();
  p.BN = N <= 128 ? N : 128;  // 2 accumulators + the A ring share the 512 TMEM columns
  // short reductions (first layer: K = 37 -> two k-blocks per tile) are bound by the depth of the operand ring, which tensor
  // memory caps at 4 A slots beside two 128-column accumulators: 64-column tiles leave room for 6 stages of 32 KB
  if (small_k_bn64() && K <= 64 && N % 64 == 0 && N >= 128) p.BN = 64;
  if (N > 128 && N % 128 != 0) return MNR_ESHAPE;
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
    if (make_map(&p.g[k].cmap, s.c, M, N, N, 32) != MNR_OK) tma_store = 0;   // C (M,N) as [32 x 32] boxes
  }
  p.tma_store = tma_store;
  p.I = M, p.J = N, p.R = K, p.ldc = N, p.act = act, p.splits = 1, p.rows_per_split = 0, p.partials = nullptr;
  p.want_bias = 0, p.n_groups = n_groups, p.stages = pick_stages(kFwd, act, p.BN), p.hints = t_cache_hints;
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
  int tma_store = tma_store_enabled();
  const int Kr = (K + 31) / 32 * 32;
  p.BN = Kr <= 128 ? Kr : 128;
  if (small_k_bn64() && N <= 64 && Kr % 64 == 0 && Kr >= 128) p.BN = 64;   // reduction over N <= 64: as in the forward launcher
  for (int k = 0; k < n_groups; ++k) {
    const mnr_linear_group_t& s = groups[k];
    if (!s.a || !s.b || !s.b_lo || !s.c || (act != 0 && !s.saved)) return MNR_ENULL;
    if (s.N != N || s.K != K) return MNR_ESHAPE;
    int rc = make_map(&p.g[k].a, s.a, M, N, N, kBM);                                 // dY (M,N), reduction over N
    if (rc == MNR_OK) rc = make_map(&p.g[k].s, act != 0 ? s.saved : s.a, M, N, N, kBM);  // Y
    if (rc == MNR_OK) rc = make_map(&p.g[k].b0, s.b, N, K, ld_or(s.ldb, K), 32, true);  // W (N,K): rows = reduction
    if (rc == MNR_OK) rc = make_map(&p.g[k].b1, s.b_lo, N, K, ld_or(s.ldb, K), 32, true);
    if (rc != MNR_OK) return rc;
    p.g[k].c = s.c, p.g[k].bias = nullptr;
    // dX (M,K): the tensor map needs a 16 B-multiple row stride; otherwise (K % 4 != 0) the plain-store epilogue runs
    if (make_map(&p.g[k].cmap, s.c, M, K, K, 32) != MNR_OK) tma_store = 0;
  }
  p.tma_store = tma_store;
  p.I = M, p.J = K, p.R = N, p.ldc = K, p.act = act, p.splits = 1, p.rows_per_split = 0, p.partials = nullptr;
  p.want_bias = 0, p.n_groups = n_groups, p.stages = pick_stages(kDx, act, p.BN), p.hints = t_cache_hints;
  dim3 grid((M + kBM - 1) / kBM, (K + p.BN - 1) / p.BN, n_groups);
  return launch<kDx>(p, grid, mnr_cs(stream));
}

namespace {
// DW tile width along K_in: up to 256 columns in one tile (single accumulator), 128-column tiles beyond that
int dw_bn(int K) {
  const int Kr = (K + 31) / 32 * 32;
  return Kr <= 256 ? Kr : 128;
}

int tc_dw_splits(int M, int tiles) {
  // Split the reduction (the minibatch rows) so that (a) one accumulation chain in TMEM covers <= 512 rows: the tensor
  // core's fp32 accumulate error grows with the chain length and the parity bar is 1e-5; (b) tiles x splits fills whole
  // waves of the persistent grid (a CTA then walks 2+ work items back to back with double-buffered accumulators).
  const int sms = eff_sms();
  int s_min = (M + 511) / 512;
  if (s_min < 1) s_min = 1;
  const int waves = (tiles * s_min + sms - 1) / sms;
  int s = waves * sms / tiles;
  if (s < s_min) s = s_min;
  const int max_s = (M + 63) / 64;  // >= 64 rows (2 k-blocks) per work item
  if (s > max_s) s = max_s;
  // MNR_TC_DW_SPLIT_CAP (A/B measurements): fewer, longer splits = less partial traffic for the reduce launch, fewer CTAs
  static int cap = -1;
  if (cap < 0) {
    const char* e = getenv("MNR_TC_DW_SPLIT_CAP");
    cap = e ? atoi(e) : 0;
  }
  if (cap > 0 && s > cap) s = cap < s_min ? s_min : cap;
  return s < 1 ? 1 : s;
}

// MNR_TC_DW2=0 keeps the first form of the weight-gradient kernel for every shape (A/B measurements); read once
int dw2_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MNR_TC_DW2");
    v = (e && e[0] == '0') ? 0 : (e && e[0] == '2') ? 2 : 1;   // 2: also for K_in < 64 (tests)
  }
  return v;
}

constexpr int kDw2BiasBytes = 8 * 256 * 4;

// tile / split / pipeline plan of one layer's weight-gradient launch.  v2: tc_dw2_kernel (X^T in tensor memory, dZ converted
// in place as the shared-memory operand) for N_out % 32 == 0; otherwise the first form (dZ^T in tensor memory).
struct DwPlan {
  bool v2;
  int BN, kt, n_mt, nxc, tiles_i, tiles_j, splits, stages, a_ring;
};
DwPlan dw_plan(int n_groups, int M, int N, int K, int act) {
  DwPlan d = {};
  // (K_in < 64, the first layer: measured 35.8 us (first form) against 37.9 us at 32768 x 256 x 37 - both at the HBM
  // bound of the dY / Y stream, the first form with the smaller conversion share)
  if (dw2_enabled() && N % 32 == 0 && (K >= 64 || dw2_enabled() == 2)) {
    d.kt = K > 128 ? 256 : 128, d.n_mt = d.kt / 128;
    d.BN = d.n_mt == 2 ? (N < 128 ? N : 128) : (N < 256 ? N : 256);
    if (N % d.BN == 0) {
      d.nxc = ((K < d.kt ? K : d.kt) + 31) / 32;
      d.tiles_i = (K + d.kt - 1) / d.kt, d.tiles_j = N / d.BN;
      const int stage = 2 * d.BN * kBK * 4 + d.nxc * kChunk;
      d.stages = (227 * 1024 - 2048 - kDw2BiasBytes) / stage;
      if (d.stages > kMaxStages) d.stages = kMaxStages;
      d.a_ring = (512 - d.n_mt * d.BN) / (d.n_mt * 64);
      if (d.a_ring > d.stages) d.a_ring = d.stages;
      d.v2 = d.stages >= 2 && d.a_ring >= 2;
    }
  }
  if (!d.v2) {
    d.BN = dw_bn(K);
    d.tiles_i = (N + kBM - 1) / kBM, d.tiles_j = (K + d.BN - 1) / d.BN;
    d.stages = pick_stages(kDw, act, d.BN);
  }
  d.splits = tc_dw_splits(M, d.tiles_i * d.tiles_j * n_groups);
  return d;
}
}  // namespace

extern "C" size_t mnr_tc_linear_bwd_weight_workspace_bytes(int n_groups, int M, int N, int K) {
  if (n_groups <= 0 || M <= 0 || N <= 0 || K <= 0) return 0;
  const DwPlan d = dw_plan(n_groups, M, N, K, 1);
  return (size_t)n_groups * d.splits * ((size_t)N * K + 2 * (size_t)N) * sizeof(float);  // dW + 2 db slabs per split
}

namespace {
// the DW launch of one layer: tile / split plan + tensor maps; `job` receives what the split reduce needs
int dw_launch(const mnr_linear_group_t* groups, int n_groups, int M, int act, void* workspace, size_t workspace_bytes, bool launch_it,
              ReduceJob* job, cudaStream_t stream) {
  if (n_groups <= 0 || n_groups > kMaxGroups) return MNR_ESHAPE;
  MNR_CHECK_PTR(groups); MNR_CHECK_PTR(workspace);
  if (M <= 0) return MNR_ESHAPE;
  if (act < 0 || act > 2) return MNR_EUNSUP;
  const int N = groups[0].N, K = groups[0].K;
  if (N <= 0 || K <= 0 || N % 4 != 0) return MNR_ESHAPE;
  if (workspace_bytes < mnr_tc_linear_bwd_weight_workspace_bytes(n_groups, M, N, K)) return MNR_EWORKSPACE;
  if (!mnr_aligned(workspace, 16)) return MNR_EALIGN;
  TcParams p;
  const DwPlan d = dw_plan(n_groups, M, N, K, act);
  p.BN = d.BN;
  bool want_bias = false;
  for (int k = 0; k < kMaxGroups; ++k) job->w[k] = nullptr, job->b[k] = nullptr;
  for (int k = 0; k < n_groups; ++k) {
    const mnr_linear_group_t& s = groups[k];
    if (!s.a || !s.b || !s.c || (act != 0 && !s.saved)) return MNR_ENULL;
    if (s.N != N || s.K != K) return MNR_ESHAPE;
    if (!mnr_aligned(s.c, 16)) return MNR_EALIGN;
    if (launch_it) {
      int rc = make_map(&p.g[k].a, s.a, M, N, N, 32, true);                            // dY (M,N): rows = reduction
      if (rc == MNR_OK) rc = make_map(&p.g[k].s, act != 0 ? s.saved : s.a, M, N, N, 32, true);
      if (rc == MNR_OK) rc = make_map(&p.g[k].b0, s.b, M, K, ld_or(s.ldb, K), 32, true);  // X (M,K)
      if (rc != MNR_OK) return rc;
      p.g[k].b1 = p.g[k].b0;
      p.g[k].c = s.c, p.g[k].bias = nullptr;
    }
    job->w[k] = s.c, job->b[k] = s.bias_grad;
    want_bias = want_bias || s.bias_grad != nullptr;
  }
  const int splits = d.splits;
  // (N % 4 == 0 was checked above, so N*K is a multiple of 4 and every dW / partial slab is 16 B aligned)
  job->partials = (const float*)workspace, job->n_groups = n_groups, job->splits = splits, job->I = N, job->J = K;
  job->wblocks = (N * K / 4 + 31) / 32;
  if (!launch_it) return MNR_OK;
  p.I = N, p.J = K, p.R = M, p.ldc = K, p.act = act, p.splits = splits;
  p.rows_per_split = ((M + splits - 1) / splits + 31) / 32 * 32;
  p.partials = (float*)workspace, p.want_bias = want_bias ? 1 : 0, p.n_groups = n_groups, p.stages = d.stages;
  p.tma_store = 0, p.hints = t_cache_hints;
  if (d.v2) {
    p.kt = d.kt, p.n_mt = d.n_mt, p.a_ring = d.a_ring, p.nxc = d.nxc;
    p.tiles_i = d.tiles_i, p.tiles_j = d.tiles_j;
    const int n_tiles = d.tiles_i * d.tiles_j * n_groups * splits;
    const int sms = eff_sms();
    dim3 grid(n_tiles < sms ? n_tiles : sms);
    const size_t smem = (size_t)d.stages * (2 * d.BN * kBK * 4 + d.nxc * kChunk) + kDw2BiasBytes + 1024;
    cudaError_t e = cudaSuccess;
#define MNR_DW2_LAUNCH(ACTV)                                                                                        \
  e = mnr_allow_max_dyn_smem(tc_dw2_kernel<ACTV>);                                                                       \
  if (e == cudaSuccess) e = mnr_launch_pdl(tc_dw2_kernel<ACTV>, grid, dim3(kThreads), smem, stream, p);
    switch (act) {
      case 1: MNR_DW2_LAUNCH(1) break;
      case 2: MNR_DW2_LAUNCH(2) break;
      default: MNR_DW2_LAUNCH(0) break;
    }
#undef MNR_DW2_LAUNCH
    if (e != cudaSuccess) return (int)e;
    e = cudaPeekAtLastError();
    return e == cudaSuccess ? MNR_OK : (int)e;
  }
  dim3 grid(d.tiles_i, d.tiles_j, n_groups * splits);
  return launch<kDw>(p, grid, stream);
}

int reduce_launch(ReduceJobs& jobs, int n_jobs, cudaStream_t stream, const mnr_fold_list_t* list = nullptr) {
  int max_x = 1;
  for (int q = 0; q < n_jobs; ++q) {
    const int x = jobs.j[q].wblocks + (jobs.j[q].I + 31) / 32;
    max_x = x > max_x ? x : max_x;
  }
  FoldJobs folds;
  folds.n = 0, folds.has_m = 0;
  if (list) {
    if (list->n_folds < 0 || list->n_folds > MNR_MAX_FOLDS) return MNR_ESHAPE;
    for (int q = 0; q < list->n_folds; ++q) {
      const mnr_fold_job_t& f = list->folds[q];
      if (!f.parts) return MNR_ENULL;
      if (f.n_parts <= 0 || f.width <= 0 || f.n_segs < 0 || f.n_segs > MNR_FOLD_MAX_SEGS) return MNR_ESHAPE;
      for (int k = 0; k < f.n_segs; ++k)
        if (!f.seg[k].dst || f.seg[k].begin < 0 || f.seg[k].end > f.width || f.seg[k].end < f.seg[k].begin) return MNR_ESHAPE;
      folds.f[q] = f;
      const bool vec = f.width % 4 == 0 && mnr_aligned(f.parts, 16);
      folds.f[q]._pad = vec ? 1 : 0;   // kernel-side flag: float4 form
      const int x = vec ? (f.width + 4 * kFoldCols - 1) / (4 * kFoldCols) : (f.width + kFoldCols - 1) / kFoldCols;
      max_x = x > max_x ? x : max_x;
    }
    folds.n = list->n_folds;
    if (list->has_metrics) {
      if (!list->metrics.loss_parts || !list->metrics.metrics) return MNR_ENULL;
      if (list->metrics.n_parts <= 0 || list->metrics.M <= 0) return MNR_ESHAPE;
      folds.m = list->metrics, folds.has_m = 1;
    }
  }
  const int extra = folds.n + (folds.has_m ? 1 : 0);
  jobs.fold_rows = extra;
  dim3 rgrid(max_x, n_jobs * kMaxGroups + extra);
  mnr_launch_pdl(tc_reduce_kernel, rgrid, dim3(256), 0, stream, jobs, folds);
  MNR_LAUNCH_RESULT();
}
}  // namespace

extern "C" int mnr_tc_linear_bwd_weight(const mnr_linear_group_t* groups, int n_groups, int M, int act, void* workspace,
                                        size_t workspace_bytes, mnr_stream_t stream) {
  ReduceJobs jobs;
  int rc = dw_launch(groups, n_groups, M, act, workspace, workspace_bytes, true, &jobs.j[0], mnr_cs(stream));
  if (rc != MNR_OK) return rc;
  return reduce_launch(jobs, 1, mnr_cs(stream));
}

extern "C" int mnr_tc_linear_bwd_weight_partials(const mnr_linear_group_t* groups, int n_groups, int M, int act, void* workspace,
                                                 size_t workspace_bytes, mnr_stream_t stream) {
  ReduceJob job;
  return dw_launch(groups, n_groups, M, act, workspace, workspace_bytes, true, &job, mnr_cs(stream));
}

extern "C" int mnr_tc_reduce_partials(const mnr_tc_reduce_job_t* jobs_in, int n_jobs, mnr_stream_t stream) {
  return mnr_tc_reduce_partials_folds(jobs_in, n_jobs, nullptr, stream);
}

extern "C" int mnr_tc_reduce_partials_folds(const mnr_tc_reduce_job_t* jobs_in, int n_jobs, const mnr_fold_list_t* list,
                                            mnr_stream_t stream) {
  const bool has_folds = list && (list->n_folds > 0 || list->has_metrics);
  if (n_jobs < 0 || n_jobs > kMaxReduceJobs || (n_jobs == 0 && !has_folds)) return MNR_ESHAPE;   // folds alone are a valid launch
  if (n_jobs > 0) MNR_CHECK_PTR(jobs_in);
  ReduceJobs jobs;
  const int cap_now = t_grid_cap;
  for (int q = 0; q < n_jobs; ++q) {
    const mnr_tc_reduce_job_t& in = jobs_in[q];
    t_grid_cap = in.grid_cap > 0 ? in.grid_cap : 0;   // the split plan of the launch that wrote these partials
    const int rc = dw_launch(in.groups, in.n_groups, in.M, 0, in.workspace, in.workspace_bytes, false, &jobs.j[q], mnr_cs(stream));
    t_grid_cap = cap_now;
    if (rc != MNR_OK) return rc;
  }
  return reduce_launch(jobs, n_jobs, mnr_cs(stream), list);
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
  mnr_launch_pdl(split_kernel, grid, dim3(256), 0, mnr_cs(stream), js);
  MNR_LAUNCH_RESULT();
}

// debug: enable / read the CTA-0 timeline (clock64 ticks), 4 roles x 256 events x 2 tags
extern "C" int mnr_tc_trace(int enable, long long* host_out) {
  cudaError_t e = cudaMemcpyToSymbol(g_trace_on, &enable, sizeof(int));
  if (e != cudaSuccess) return (int)e;
  if (host_out) {
    e = cudaMemcpyFromSymbol(host_out, g_trace, sizeof(long long) * 4 * 256 * 2);
    if (e != cudaSuccess) return (int)e;
  }
  return MNR_OK;
}
