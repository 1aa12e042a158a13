// G  a whole 3-layer tower FORWARD in one tcgen05 kernel: Y1 = act(X W1^T + b1), Y2 = act(Y1 W2^T + b2),
// Y3 = act(Y2 W3^T + b3) for a 128-row tile at a time, with the layer outputs handed from one GEMM to the next THROUGH
// TENSOR MEMORY.  Replaces `MLP.forward` of both PPO towers (mineral/nets/mlp.py:25-77, called from
// mineral/agents/ppo/models.py:136-151) on the update path and in the rollout; same 3xTF32 split as tc_linear.cu, so the
// parity bar (1e-5) holds.
//
// Why: with one launch per layer, layer l+1 re-reads from HBM / L2 what layer l's epilogue just had in registers, and every
// launch pays its own ramp (TMEM allocation, descriptor fetch, first TMA round trip) and its own wave tail.  Here the
// epilogue of layer l applies bias + activation to its accumulator chunk, sends it to HBM (the backward pass needs Y1..Y3)
// AND writes its (hi, lo) tf32 parts into the ring of TMEM slots layer l+1's MMAs read their A operand from - the same
// slots the converter warps of tc_linear.cu fill from shared memory.
//
// Tensor memory (512 columns):   [0,256)   two 128-column buffers: layer-1 half h of the CTA's t-th tile accumulates in buffer
//                                          (h + t) & 1, layer 3 of that tile in buffer t & 1 (N3 <= 128)
//                                [256,384) ACC2: layer 2 (N2 <= 128)
//                                [384,512) A ring: 2 slots x (32 hi | 32 lo) columns = one 32-wide reduction block each
// Shared memory: X tile as (hi, lo) K-major SWIZZLE_128B tiles (split in place by the converter warps; layer 1's A operand
// is the only one read from shared memory), a 3-slot ring of 32 KiB weight blocks (hi | lo of [<=128 rows][32 floats],
// streamed from L2 by TMA for every tile: W2 alone is 256 KiB as a split pair), 16 x 4 KiB epilogue boxes for the TMA stores.
//
// The epilogue warps bound the kernel (8 warps drain 128 x (N1 + N2 + N3) accumulator elements per tile), so the MMA
// issue order keeps them fed: L1(t, half 1), L2(t)'s blocks as their A slots arrive from the epilogue of L1(t), then
// L1(t+1, half 0) - its accumulator buffer is free and its epilogue can start the moment layer 3's is done -, then L3(t).
// Warp roles (512 threads, persistent, 1 CTA/SM): 0 weight producer, 1 MMA issuer, 2 TMEM allocator, 3 X producer,
// 4-7 X converters, 8-15 epilogue (two warps per TMEM lane quarter, alternate 32-column chunks).
#include "tc_common.cuh"

namespace {

constexpr int kBaseThreads = 256;   // warps 0-7: weight producer, MMA issuer, TMEM allocator, X producer, 4 X converters
constexpr int kBM = 128, kBK = 32;
constexpr int kTile = kBM * kBK * 4;   // 16 KiB: [128][32] f32, SWIZZLE_128B
constexpr int kWSlot = 2 * kTile;      // weight ring slot: hi at +0, lo at +kTile
constexpr int kWStages = 3;
constexpr int kASlots = 2;
constexpr int kMaxXkb = 2;             // K0 <= 64
constexpr int kMaxGroups = 2;
constexpr int kLayers = 3;
constexpr uint32_t kAccB = 0, kAcc2 = 256, kARing = 384, kTmemCols = 512;   // two 128-column buffers | ACC2 | A ring
constexpr int kEpi0 = 8;            // epilogue warps 8 .. 8 + EPIW - 1 (EPIW = 8 or 16: two or four warps per TMEM lane quarter)
constexpr int kBoxBytes = 64 * 1024; // TMA store boxes: EPIW warps x (64 KiB / EPIW)

__device__ long long g_ttrace[4 * 256 * 2];
__device__ int g_ttrace_on = 0;
#define TT_TRACE(role, idx, tag)                                                                                        \
  do {                                                                                                                  \
    if (g_ttrace_on && blockIdx.x == 0 && (idx) < 256) g_ttrace[((role) * 256 + (idx)) * 2 + (tag)] = clock64();        \
  } while (0)

struct alignas(64) TowerGroup {
  CUtensorMap x;                             // X (M, K0): boxes [128 rows][32 floats]
  CUtensorMap w_hi[kLayers], w_lo[kLayers];  // W_l (N_l, K_l) pre-split: boxes [min(N_l,128) rows][32 floats]
  CUtensorMap y[kLayers];                    // Y_l (M, N_l): store boxes [32 rows][32 floats]
  const float* bias[kLayers];
  const float* w_ptr[2 * kLayers];           // the same weight buffers as plain pointers (L2 prefetch at kernel start)
  uint32_t w_bytes[kLayers];
};

struct TowerParams {
  TowerGroup g[kMaxGroups];
  int M, K0;
  int N[kLayers];
  int n_groups, tiles_i;
  int flags;   // bit 0: L2 prefetch of the weights at kernel start (MNR_TOWER_FLAGS, A/B measurements)
};

// L2 prefetch of the weight buffers (one 128 B line per thread and step), kept out of line: inlined into the kernel it changed
// the register allocation of the epilogue warps (61 -> 72 us per launch at cfg2)
__device__ __noinline__ void prefetch_weights(const TowerParams& p, uint32_t gt, uint32_t gn) {
  for (int gi = 0; gi < p.n_groups; ++gi)
    for (int q = 0; q < 2 * kLayers; ++q) {
      const char* base = reinterpret_cast<const char*>(p.g[gi].w_ptr[q]);
      const uint32_t lines = (p.g[gi].w_bytes[q >> 1] + 127u) >> 7;
      for (uint32_t l = gt; l < lines; l += gn) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)l * 128));
    }
}

template <int ACT, bool PREFETCH, int EPIW>
__global__ void __launch_bounds__(kBaseThreads + EPIW * 32, 1) tc_tower_fwd_kernel(const __grid_constant__ TowerParams p) {
  constexpr int kEpiWarps = EPIW;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t x_full, x_conv, x_empty;
  __shared__ __align__(8) uint64_t w_full[kWStages], w_empty[kWStages];
  // a_empty[u & 3] completes when the MMAs of A-slot use u have read their slot (= slot u & 1).  FOUR barriers for two slots: a
  // warp that fills use u waits for use u - 2, i.e. on barrier (u - 2) & 3, and the warps of one residue class of u (mod 4) are
  // the only waiters of that barrier and see every one of its phases in order - a parity wait must never be two phases away
  // from the barrier's current one (with one barrier per slot, a warp that handles only every fourth use would be).
  __shared__ __align__(8) uint64_t a_full[kASlots], a_empty[4];
  __shared__ __align__(8) uint64_t b_full[2], b_empty[2], acc2_full;
  __shared__ uint32_t tmem_slot;
  // per-warp bias row (8 epilogue warps; the 16-warp form broadcasts the bias by shuffles: 2 KiB of static shared memory would
  // push the CTA 160 B over the 227 KiB limit)
  __shared__ __align__(16) float s_bias_all[(EPIW == 16 ? 1 : EPIW) * 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N1 = p.N[0], N2 = p.N[1], N3 = p.N[2];
  const int h1 = N1 / 128, nkb1 = (p.K0 + kBK - 1) / kBK, nkb3 = N2 / kBK;
  const int uses_per_tile = 4 * h1 + nkb3;   // A-slot uses: layer 2's 4*h1 reduction blocks, then layer 3's
  const int n_tiles = p.tiles_i * p.n_groups;

  uint8_t* smem = smem_dyn + ((1024u - (ptx::smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint8_t* xs = smem;                                   // [kb][hi 16 KiB | lo 16 KiB]
  uint8_t* ring = smem + kMaxXkb * 2 * kTile;           // kWStages x 32 KiB
  uint8_t* boxes = ring + kWStages * kWSlot;            // 64 KiB: 8 warps x 2 boxes, or 16 warps x 1 box, of 4 KiB

  if (tid == 0) {
    ptx::mbar_init(&x_full, 1), ptx::mbar_init(&x_conv, 128), ptx::mbar_init(&x_empty, 1);
    for (int s = 0; s < kWStages; ++s) ptx::mbar_init(&w_full[s], 1), ptx::mbar_init(&w_empty[s], 1);
    for (int s = 0; s < kASlots; ++s) ptx::mbar_init(&a_full[s], 128);
    for (int s = 0; s < 4; ++s) ptx::mbar_init(&a_empty[s], 1);
    for (int b = 0; b < 2; ++b) ptx::mbar_init(&b_full[b], 1), ptx::mbar_init(&b_empty[b], kEpiWarps * 32);
    ptx::mbar_init(&acc2_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(&tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  mnr::pdl_trigger();
  mnr::pdl_wait();
  // tensor-map descriptors (7 loads + 3 stores per tower) fetched ahead of their first use: the first tile of a CTA otherwise
  // pays one cold descriptor fetch per layer on the epilogue's critical path (CTA-0 timeline: 18.7 us for the first tile
  // against 9.4 us for the following ones)
  if (warp == 2 && lane < kLayers * 3 + 1) {
    for (int gi = 0; gi < p.n_groups; ++gi) {
      const TowerGroup& g = p.g[gi];
      const CUtensorMap* m = lane == 0 ? &g.x : lane <= kLayers ? &g.w_hi[lane - 1] : lane <= 2 * kLayers ? &g.w_lo[lane - 1 - kLayers] : &g.y[lane - 1 - 2 * kLayers];
      ptx::prefetch_tmap(m);
    }
  }
  // Cold start: the split weights (0.4 MB per tower) were written by the optimizer kernel a whole backward pass ago and have
  // left L2; through the 3-slot ring every first-tile block would pay a DRAM round trip.  The converter warps of all CTAs
  // (idle until the first X tile lands) pull distinct 128 B lines of every weight buffer into L2 now.
  if (PREFETCH && warp >= 4 && warp < 8) prefetch_weights(p, blockIdx.x * 128 + (tid - 128), gridDim.x * 128);

  if (warp == 0) {
    // ===== weight producer: the ring items in the order the MMA issuer consumes them ================================
    if (ptx::elect_one()) {
      uint32_t wi = 0;
      auto put = [&](const CUtensorMap* hi, const CUtensorMap* lo, int rows, int c0, int r0) {
        const int st = wi % kWStages;
        ptx::mbar_wait(&w_empty[st], ((wi / kWStages) & 1) ^ 1);
        uint8_t* base = ring + (size_t)st * kWSlot;
        ptx::mbar_expect_tx(&w_full[st], 2u * (uint32_t)rows * 128u);
        ptx::tma_load_2d(base, hi, c0, r0, &w_full[st]);
        ptx::tma_load_2d(base + kTile, lo, c0, r0, &w_full[st]);
        ++wi;
      };
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TowerGroup& g = p.g[tile / p.tiles_i];
        if (tile == (int)blockIdx.x)
          for (int kb = 0; kb < nkb1; ++kb) put(&g.w_hi[0], &g.w_lo[0], 128, kb * kBK, 0);
        if (h1 == 2)
          for (int kb = 0; kb < nkb1; ++kb) put(&g.w_hi[0], &g.w_lo[0], 128, kb * kBK, 128);
        for (int j = 0; j < 4 * h1; ++j) put(&g.w_hi[1], &g.w_lo[1], N2, j * kBK, 0);
        const int nxt = tile + gridDim.x;
        if (nxt < n_tiles) {
          const TowerGroup& gn = p.g[nxt / p.tiles_i];
          for (int kb = 0; kb < nkb1; ++kb) put(&gn.w_hi[0], &gn.w_lo[0], 128, kb * kBK, 0);
        }
        for (int j = 0; j < nkb3; ++j) put(&g.w_hi[2], &g.w_lo[2], N3, j * kBK, 0);
      }
    }
  } else if (warp == 3) {
    // ===== X producer ===============================================================================================
    if (ptx::elect_one()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const TowerGroup& g = p.g[tile / p.tiles_i];
        const int i0 = (tile % p.tiles_i) * kBM;
        ptx::mbar_wait(&x_empty, (it & 1) ^ 1);
        ptx::mbar_expect_tx(&x_full, (uint32_t)nkb1 * kTile);
        for (int kb = 0; kb < nkb1; ++kb) ptx::tma_load_2d(xs + (size_t)kb * 2 * kTile, &g.x, kb * kBK, i0, &x_full);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ===============================================================================================
    // Issue order: L1(0,h=0) | per tile t: L1(t,h=1), L2(t) blocks as their A slots arrive, L1(t+1,h=0), L3(t) blocks.
    // Layer-1 half h of tile t accumulates in buffer (h + t) & 1; layer 3 of tile t in buffer t & 1 (drained by then).
    if (ptx::elect_one()) {
      const uint32_t idesc1 = ptx::idesc_tf32(kBM, 128, 0, 0), idesc2 = ptx::idesc_tf32(kBM, N2, 0, 0), idesc3 = ptx::idesc_tf32(kBM, N3, 0, 0);
      // K-major SWIZZLE_128B operand tiles (A of layer 1 and every B): 8-row groups 1 KiB apart (SBO), LBO unused (16 B)
      const uint64_t desc_hi = ptx::smem_desc(0, 0, 1024, 2);
      const uint32_t desc_lo_const = (16u >> 4) << 16;
      auto kdesc = [&](uint32_t addr, int k) { return desc_hi | (uint64_t)((((addr) >> 4) | desc_lo_const) + (uint32_t)(k * 2)); };
      const uint32_t xs_u = ptx::smem_u32(xs), ring_u = ptx::smem_u32(ring);
      const uint32_t acc2 = tmem_base + kAcc2, aring = tmem_base + kARing;
      uint32_t wi = 0, u = 0, it = 0, bc0 = 0, bc1 = 0;   // ring item, A-slot use, tile, per-buffer use counters
      auto layer1 = [&](uint32_t t, int h) {   // layer 1, half h of this CTA's t-th tile (A = X hi / lo from shared memory)
        const int b = (h + t) & 1;
        ptx::mbar_wait(&b_empty[b], ((b ? bc1 : bc0) & 1) ^ 1);
        b ? ++bc1 : ++bc0;
        if (h == 0) ptx::mbar_wait(&x_conv, t & 1);
        ptx::tc_fence_after();
        TT_TRACE(0, t * 4 + h, 0);
        const uint32_t acc = tmem_base + kAccB + 128u * b;
        for (int kb = 0; kb < nkb1; ++kb, ++wi) {
          const int st = wi % kWStages;
          ptx::mbar_wait(&w_full[st], (wi / kWStages) & 1);
          ptx::tc_fence_after();
          const uint32_t a_u = xs_u + (uint32_t)kb * 2u * kTile, b_u = ring_u + (uint32_t)st * kWSlot;
          const int r_left = p.K0 - kb * kBK;
          const int ksteps = r_left >= kBK ? kBK / 8 : (r_left + 7) / 8;   // no MMAs over TMA's zero fill
          for (int k = 0; k < ksteps; ++k) {
            ptx::mma_tf32(acc, kdesc(a_u + kTile, k), kdesc(b_u, k), idesc1, (kb | k) != 0 ? 1u : 0u);  // small terms first
            ptx::mma_tf32(acc, kdesc(a_u, k), kdesc(b_u + kTile, k), idesc1, 1u);
            ptx::mma_tf32(acc, kdesc(a_u, k), kdesc(b_u, k), idesc1, 1u);
          }
          ptx::tc_commit(&w_empty[st]);
        }
        ptx::tc_commit(&b_full[b]);
        if (h == h1 - 1) ptx::tc_commit(&x_empty);
        TT_TRACE(0, t * 4 + h, 1);
      };
      // A (hi | lo) from TMEM slot, B from the weight ring; `first`: the block that initialises the accumulator
      auto block_ts = [&](uint32_t acc, uint32_t idesc, bool first) {
        const int as = u % kASlots, st = wi % kWStages;
        ptx::mbar_wait(&a_full[as], (u / kASlots) & 1);
        ptx::mbar_wait(&w_full[st], (wi / kWStages) & 1);
        ptx::tc_fence_after();
        const uint32_t a_t = aring + (uint32_t)as * 64u, b_u = ring_u + (uint32_t)st * kWSlot;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ptx::mma_tf32_ts(acc, a_t + 32u + 8u * k, kdesc(b_u, k), idesc, (!first || k != 0) ? 1u : 0u);
          ptx::mma_tf32_ts(acc, a_t + 8u * k, kdesc(b_u + kTile, k), idesc, 1u);
          ptx::mma_tf32_ts(acc, a_t + 8u * k, kdesc(b_u, k), idesc, 1u);
        }
        ptx::tc_commit(&w_empty[st]);
        ptx::tc_commit(&a_empty[u & 3]);
        ++wi, ++u;
      };
      if ((int)blockIdx.x < n_tiles) layer1(0, 0);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        if (h1 == 2) layer1(it, 1);
        // ---- layer 2 -> ACC2 (A = Y1 hi / lo from the TMEM ring).  ACC2 was drained by the previous tile's layer-2 epilogue:
        //      every one of its chunks filled an A slot that the layer-3 blocks issued before this point have waited for ----
        TT_TRACE(1, it, 0);
        for (int j = 0; j < 4 * h1; ++j) block_ts(acc2, idesc2, j == 0);
        ptx::tc_commit(&acc2_full);
        TT_TRACE(1, it, 1);
        if (tile + (int)gridDim.x < n_tiles) layer1(it + 1, 0);
        // ---- layer 3 -> buffer it & 1 (layer 1, half 0 of this tile: drained long ago) ----
        const int b = it & 1;
        ptx::mbar_wait(&b_empty[b], ((b ? bc1 : bc0) & 1) ^ 1);
        b ? ++bc1 : ++bc0;
        ptx::tc_fence_after();
        TT_TRACE(2, it, 0);
        for (int j = 0; j < nkb3; ++j) block_ts(tmem_base + kAccB + 128u * b, idesc3, j == 0);
        ptx::tc_commit(&b_full[b]);
        TT_TRACE(2, it, 1);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== X converters: raw fp32 tile -> hi (in place) / lo, element-wise (the swizzle is the same for both) =======
    const int t = tid - 128;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      ptx::mbar_wait(&x_full, it & 1);
      for (int kb = 0; kb < nkb1; ++kb) {
        float4* H = reinterpret_cast<float4*>(xs + (size_t)kb * 2 * kTile);
        float4* L = reinterpret_cast<float4*>(xs + (size_t)kb * 2 * kTile + kTile);
#pragma unroll
        for (int q = 0; q < kTile / 16 / 128; ++q) {
          float4 h4, l4;
          split4(H[t + 128 * q], h4, l4);
          H[t + 128 * q] = h4, L[t + 128 * q] = l4;
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&x_conv);
    }
  } else if (warp >= kEpi0) {
    // ===== epilogue: TMEM chunk -> bias + activation -> (hi, lo) into the next layer's A slot, and the chunk itself to
    //       HBM through a swizzled [32][32] box and one TMA tensor store (two boxes per warp, used alternately) ==========
    const int quarter = warp & 3, sub = (warp - kEpi0) >> 2, ew = warp - kEpi0;
    constexpr int kPerQ = kEpiWarps / 4;             // warps per TMEM lane quarter: each takes the chunks sub, sub + kPerQ, ... of 4
    constexpr int kBoxes = kBoxBytes / kEpiWarps / 4096;   // store boxes per warp, used in turn (2 with 8 warps, 1 with 16)
    constexpr int W = kEpiWarps == 16 ? 16 : 32;     // accumulator columns per pass (16 epilogue warps live within 80 registers)
    const uint32_t lane_off = (uint32_t)(32 * quarter) << 16;
    uint8_t* box_ptr = boxes + (size_t)ew * (kBoxes * 4096);
    float* s_bias = s_bias_all + (W == 16 ? 0 : ew * 32);
    uint32_t it = 0, nbox = 0, bc0 = 0, bc1 = 0;
    // L2 eviction order of the stored outputs (flags bit 2): the backward pass reads them top layer first, so Y3 and Y2 (50 MB for
    // both towers at 32768 rows) are asked to stay and Y1 (67 MB, needed last, larger than what L2 can keep beside them) to go first
    const bool hints = (p.flags & 4) != 0;
    const uint64_t pol_keep = ptx::policy_evict_last(), pol_go = ptx::policy_evict_first();
    // fine-grained stamps of epilogue warp 0's chunks (trace mode only): role-1 slots 64.., 8 stamps per chunk
    uint32_t fine = 0;
#define TT_FINE()                                                                                                       \
  do {                                                                                                                  \
    if (g_ttrace_on && blockIdx.x == 0 && ew == 0 && lane == 0 && fine < 384) {                                         \
      g_ttrace[(256 + 64 + (fine >> 1)) * 2 + (fine & 1)] = clock64();                                                  \
      ++fine;                                                                                                           \
    }                                                                                                                   \
  } while (0)
    // one 32-column chunk of an accumulator: bias + activation, to HBM through a swizzled [32][32] box and one TMA tensor store;
    // to_a: its (hi, lo) parts also go to the A slot of use `use` (the next layer's operand).
    auto chunk = [&](uint32_t acc_col, float bias_v, const CUtensorMap* ymap, int col, int row, bool to_a, uint32_t use, uint64_t* drained,
                     bool keep = false) {
      TT_FINE();
      __syncwarp();
      if (W == 32) s_bias[lane] = bias_v;
      uint8_t* bp = box_ptr + (nbox % kBoxes) * 4096;
      ++nbox;
      __syncwarp();
      const uint32_t row_addr = ptx::smem_u32(bp) + (uint32_t)lane * 128u;
      const int as = use % kASlots;
      const uint32_t a_t = tmem_base + lane_off + kARing + (uint32_t)as * 64u;
      const float4* b4 = reinterpret_cast<const float4*>(s_bias);
#pragma unroll
      for (int hh = 0; hh < 32 / W; ++hh) {
        uint32_t r[W];
        if (W == 32) ptx::tmem_ld_32x32(tmem_base + lane_off + acc_col, r);
        else ptx::tmem_ld_32x16(tmem_base + lane_off + acc_col + (uint32_t)(hh * W), r);
        ptx::tmem_ld_wait();
        if (hh == 0) TT_FINE();
        if (drained && hh == 32 / W - 1) {   // last read of this accumulator by this warp: the MMA issuer may overwrite it
          ptx::tc_fence_before();
          ptx::mbar_arrive(drained);
        }
        float v[W];
#pragma unroll
        for (int q4 = 0; q4 < W / 4; ++q4) {
          float4 b;
          if (W == 32) {
            b = b4[hh * (W / 4) + q4];
          } else {   // column q's bias lives in lane q
            b.x = __shfl_sync(0xffffffffu, bias_v, hh * W + 4 * q4), b.y = __shfl_sync(0xffffffffu, bias_v, hh * W + 4 * q4 + 1);
            b.z = __shfl_sync(0xffffffffu, bias_v, hh * W + 4 * q4 + 2), b.w = __shfl_sync(0xffffffffu, bias_v, hh * W + 4 * q4 + 3);
          }
          v[4 * q4] = act_fwd_t<ACT>(__uint_as_float(r[4 * q4]) + b.x), v[4 * q4 + 1] = act_fwd_t<ACT>(__uint_as_float(r[4 * q4 + 1]) + b.y);
          v[4 * q4 + 2] = act_fwd_t<ACT>(__uint_as_float(r[4 * q4 + 2]) + b.z), v[4 * q4 + 3] = act_fwd_t<ACT>(__uint_as_float(r[4 * q4 + 3]) + b.w);
        }
        if (hh == 0) TT_FINE();
        if (to_a) {   // the tcgen05.st complete while the box is being written below
          if (hh == 0) {
            if (use >= 2) ptx::mbar_wait(&a_empty[(use - 2) & 3], ((use - 2) >> 2) & 1);   // use - 2 (same slot) has been read
            ptx::tc_fence_after();
            TT_FINE();
          }
#pragma unroll
          for (int q = 0; q < W; ++q) r[q] = __float_as_uint(v[q]) & 0xffffe000u;
          if (W == 32) ptx::tmem_st_32x32(a_t, r);
          else ptx::tmem_st_32x16(a_t + (uint32_t)(hh * W), r);
#pragma unroll
          for (int q = 0; q < W; ++q) r[q] = __float_as_uint(v[q] - __uint_as_float(r[q]));
          if (W == 32) ptx::tmem_st_32x32(a_t + 32u, r);
          else ptx::tmem_st_32x16(a_t + 32u + (uint32_t)(hh * W), r);
        } else if (hh == 0) {
          TT_FINE();
        }
        if (hh == 0) {
          TT_FINE();
          if (ptx::elect_one()) {   // the store last issued from this box has finished reading it
            if (kBoxes == 2) ptx::bulk_wait_read1();
            else ptx::bulk_wait_read0();
          }
          __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < W / 4; ++c) {
          const int piece = hh * (W / 4) + c;
          ptx::st_shared_v4(row_addr + (uint32_t)((piece ^ (lane & 7)) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
      }
      if (to_a) {   // hand the A slot on first: the next layer's MMAs wait for it, the store to HBM does not hurry
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&a_full[as]);
      }
      TT_FINE();
      ptx::fence_proxy_async_smem();
      TT_FINE();
      __syncwarp();
      if (ptx::elect_one()) {
        if (hints) ptx::tma_store_2d_hint(ymap, bp, col, row, keep ? pol_keep : pol_go);
        else ptx::tma_store_2d(ymap, bp, col, row);
        ptx::bulk_commit();
      }
      TT_FINE();
    };
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const TowerGroup& g = p.g[tile / p.tiles_i];
      const int row = (tile % p.tiles_i) * kBM + 32 * quarter;
      const uint32_t u0 = it * (uint32_t)uses_per_tile;
      // ---- layer 1 halves (4 chunks each) ----
      for (int h = 0; h < h1; ++h) {
        const int b = (h + it) & 1;
        float bv[2];
#pragma unroll
        for (int k = 0; k < 4 / kPerQ; ++k) bv[k] = __ldg(g.bias[0] + h * 128 + (sub + k * kPerQ) * 32 + lane);
        ptx::mbar_wait(&b_full[b], (b ? bc1 : bc0) & 1);
        b ? ++bc1 : ++bc0;
        ptx::tc_fence_after();
        if (ew == 0) TT_TRACE(3, it * 4 + h, 0);
        const uint32_t acc = kAccB + 128u * b;
#pragma unroll
        for (int k = 0; k < 4 / kPerQ; ++k) {
          const int c = sub + k * kPerQ;
          chunk(acc + c * 32, bv[k], &g.y[0], h * 128 + c * 32, row, true, u0 + 4 * h + c, k == 4 / kPerQ - 1 ? &b_empty[b] : nullptr);
        }
        if (ew == 0) TT_TRACE(3, it * 4 + h, 1);
      }
      // ---- layer 2 ----
      {
        const int n2c = N2 / 32;
        float bv[2];
#pragma unroll
        for (int k = 0; k < 4 / kPerQ; ++k) {
          const int c = sub + k * kPerQ;
          bv[k] = c < n2c ? __ldg(g.bias[1] + c * 32 + lane) : 0.f;
        }
        ptx::mbar_wait(&acc2_full, it & 1);
        ptx::tc_fence_after();
        if (ew == 0) TT_TRACE(3, it * 4 + 2, 0);
#pragma unroll
        for (int k = 0; k < 4 / kPerQ; ++k) {
          const int c = sub + k * kPerQ;
          if (c < n2c) chunk(kAcc2 + c * 32, bv[k], &g.y[1], c * 32, row, true, u0 + 4 * h1 + c, nullptr, true);
        }
        if (ew == 0) TT_TRACE(3, it * 4 + 2, 1);
      }
      // ---- layer 3 (accumulated in buffer it & 1) ----
      {
        const int b = it & 1;
        const int n3c = N3 / 32;
        float bv[2];
#pragma unroll
        for (int k = 0; k < 4 / kPerQ; ++k) {
          const int c = sub + k * kPerQ;
          bv[k] = c < n3c ? __ldg(g.bias[2] + c * 32 + lane) : 0.f;
        }
        ptx::mbar_wait(&b_full[b], (b ? bc1 : bc0) & 1);
        b ? ++bc1 : ++bc0;
        ptx::tc_fence_after();
        if (ew == 0) TT_TRACE(3, it * 4 + 3, 0);
        const uint32_t acc = kAccB + 128u * b;
        bool arrived = false;
#pragma unroll
        for (int k = 0; k < 4 / kPerQ; ++k) {
          const int c = sub + k * kPerQ;
          if (c < n3c) {
            const bool last = c + kPerQ >= n3c || k == 4 / kPerQ - 1;   // this warp's last chunk of the buffer
            chunk(acc + c * 32, bv[k], &g.y[2], c * 32, row, false, 0u, last ? &b_empty[b] : nullptr, true);
            arrived = arrived || last;
          }
        }
        if (!arrived) {   // a warp without a chunk of this layer still counts on the barrier
          ptx::tc_fence_before();
          ptx::mbar_arrive(&b_empty[b]);
        }
        if (ew == 0) TT_TRACE(3, it * 4 + 3, 1);
      }
    }
    if (ptx::elect_one()) ptx::bulk_wait0();   // this warp's Y boxes have landed
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

constexpr size_t kSmemBytes = 1024 + (size_t)kMaxXkb * 2 * kTile + (size_t)kWStages * kWSlot + kBoxBytes;

bool tower_shape_ok(int K0, const int32_t* widths, int n_layers, int act) {
  if (n_layers != kLayers || act < 0 || act > 2) return false;
  if (K0 <= 0 || K0 > kMaxXkb * kBK) return false;
  const int N1 = widths[0], N2 = widths[1], N3 = widths[2];
  if (N1 != 128 && N1 != 256) return false;
  if (N2 <= 0 || N2 > 128 || N2 % 64 != 0) return false;   // an even number of 32-column chunks (A-slot bookkeeping above)
  if (N3 <= 0 || N3 > 128 || N3 % 32 != 0) return false;
  return true;
}

// MNR_TOWER_EPI=8|16: epilogue warps of the kernel (two or four per TMEM lane quarter); read once
int epi_warps() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MNR_TOWER_EPI");
    v = (e && atoi(e) == 8) ? 8 : 16;   // measured: 4.50 ms per update with 16 against 4.53 with 8 (cfg2)
  }
  return v;
}

// MNR_TC_FUSED_FWD=0 keeps one launch per layer (A/B measurements); read once
int fused_fwd_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MNR_TC_FUSED_FWD");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}

}  // namespace

extern "C" int mnr_tc_tower_fwd_supported(int K0, const int32_t* widths, int n_layers, int act) {
  if (!widths || !encode_fn() || !fused_fwd_enabled()) return 0;
  return tower_shape_ok(K0, widths, n_layers, act) ? 1 : 0;
}

extern "C" int mnr_tc_tower_fwd(const mnr_tower_group_t* groups, int n_groups, int M, int K0, const int32_t* widths, int n_layers, int act,
                                mnr_stream_t stream) {
  if (n_groups <= 0 || n_groups > kMaxGroups) return MNR_ESHAPE;
  MNR_CHECK_PTR(groups); MNR_CHECK_PTR(widths);
  if (M <= 0) return MNR_ESHAPE;
  if (!tower_shape_ok(K0, widths, n_layers, act)) return MNR_EUNSUP;
  TowerParams p;
  p.M = M, p.K0 = K0, p.n_groups = n_groups, p.tiles_i = (M + kBM - 1) / kBM;
  {
    const char* e = getenv("MNR_TOWER_FLAGS");
    p.flags = e ? atoi(e) : 7;   // bit 0: weight prefetch, bit 1: (unused), bit 2: L2 eviction hints on the Y stores
  }
  int kin = K0;
  for (int l = 0; l < kLayers; ++l) p.N[l] = widths[l];
  for (int k = 0; k < n_groups; ++k) {
    const mnr_tower_group_t& s = groups[k];
    if (!s.x) return MNR_ENULL;
    int rc = make_map(&p.g[k].x, s.x, M, K0, s.ldx > 0 ? s.ldx : K0, kBM);
    kin = K0;
    for (int l = 0; l < kLayers && rc == MNR_OK; ++l) {
      if (!s.w_hi[l] || !s.w_lo[l] || !s.bias[l] || !s.y[l]) return MNR_ENULL;
      const int N = widths[l], ldw = s.ldw[l] > 0 ? s.ldw[l] : kin, box = N < 128 ? N : 128;
      rc = make_map(&p.g[k].w_hi[l], s.w_hi[l], N, kin, ldw, box);
      if (rc == MNR_OK) rc = make_map(&p.g[k].w_lo[l], s.w_lo[l], N, kin, ldw, box);
      if (rc == MNR_OK) rc = make_map(&p.g[k].y[l], s.y[l], M, N, N, 32);
      p.g[k].bias[l] = s.bias[l];
      p.g[k].w_ptr[2 * l] = s.w_hi[l], p.g[k].w_ptr[2 * l + 1] = s.w_lo[l], p.g[k].w_bytes[l] = (uint32_t)((size_t)N * ldw * sizeof(float));
      kin = N;
    }
    if (rc != MNR_OK) return rc;
  }
  const int n_tiles = p.tiles_i * n_groups, sms = mnr_sm_count();
  dim3 grid(n_tiles < sms ? n_tiles : sms);
  cudaError_t e = cudaSuccess;
#define MNR_TOWER_GO(ACTV, PF, EW)                                                                                              \
  {                                                                                                                             \
    e = cudaFuncSetAttribute(tc_tower_fwd_kernel<ACTV, PF, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);  \
    if (e == cudaSuccess)                                                                                                       \
      e = mnr_launch_pdl(tc_tower_fwd_kernel<ACTV, PF, EW>, grid, dim3(kBaseThreads + EW * 32), kSmemBytes, mnr_cs(stream), p); \
  }
  // 16 epilogue warps: every warp owns one residue class of the A-slot uses (mod 4), which needs uses per tile % 4 == 0
  const bool epi16 = epi_warps() == 16 && ((4 * (widths[0] / 128) + widths[1] / 32) % 4 == 0);
#define MNR_TOWER_LAUNCH(ACTV)                    \
  if (epi16) {                                    \
    if (p.flags & 1) MNR_TOWER_GO(ACTV, true, 16) \
    else MNR_TOWER_GO(ACTV, false, 16)            \
  } else {                                        \
    if (p.flags & 1) MNR_TOWER_GO(ACTV, true, 8)  \
    else MNR_TOWER_GO(ACTV, false, 8)             \
  }
  switch (act) {
    case 1: MNR_TOWER_LAUNCH(1) break;
    case 2: MNR_TOWER_LAUNCH(2) break;
    default: MNR_TOWER_LAUNCH(0) break;
  }
#undef MNR_TOWER_GO
#undef MNR_TOWER_LAUNCH
  if (e != cudaSuccess) return (int)e;
  MNR_LAUNCH_RESULT();
}

// debug: enable / read the CTA-0 timeline of the fused tower kernel (clock64 ticks), 4 roles x 256 events x 2 tags
extern "C" int mnr_tc_tower_trace(int enable, long long* host_out) {
  cudaError_t e = cudaMemcpyToSymbol(g_ttrace_on, &enable, sizeof(int));
  if (e != cudaSuccess) return (int)e;
  if (host_out) {
    e = cudaMemcpyFromSymbol(host_out, g_ttrace, sizeof(long long) * 4 * 256 * 2);
    if (e != cudaSuccess) return (int)e;
  }
  return MNR_OK;
}
