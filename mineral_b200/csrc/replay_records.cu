// K3c / K3d  circular replay buffer as ONE ring of 128 B-aligned transition RECORDS, staged through shared memory by TMA.
//
// The reference keeps five arrays (obs, action, reward, next_obs, done: mineral/buffers/simple_replay.py:4-16) and gathers a
// uniform sample with five `index` kernels (:73-85): 148 B Ant rows straddle 32 B sectors, so a random row costs 2-3x its
// bytes in DRAM traffic (round 1 measured 7.5 MB read for 2.7 MB of rows) and two dependent round trips per element.
// Here a transition is one record
//
//     [ obs (pad4) | next_obs (pad4) | action (pad4) | reward (4) | done (4: byte 0 = the bool) | pad ]   R floats, R % 32 == 0
//
// (Ant: 40 + 40 + 8 + 4 + 4 = 96 floats = 384 B = exactly three 128 B lines), and `buf_obs / buf_action / buf_reward /
// buf_next_obs / buf_done` are strided VIEWS of the (C, R) record array, so the reference's `(C, *shape)` indexing contract holds.
//
//   sample:  CTA = 32 rows.  32 indices -> shared memory; ONE cp.async.bulk (1-D TMA, R * 4 bytes) per row, issued by the lanes
//            of warp 0, all 32 in flight on one mbarrier (a single DRAM round trip for the CTA's 12 KB); then every thread takes
//            float4 pieces of the staged records and writes them where the agents want them - obs / next_obs normalised and
//            clamped (RunningMeanStd.normalize, mineral/common/normalizers.py:65-68) straight into column blocks of the padded
//            critic-input buffers, done -> f32 (simple_replay.py:84).  Every fetched sector is used: DRAM traffic = record bytes.
//   insert:  thread per float4 of a record: gathers its (up to) four elements from the incoming field arrays and writes whole
//            16 B pieces of the record; ring slot by the reference's wrap rule (simple_replay.py:42-60: head first, then tail).
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kRowsPerCta = 32;
constexpr int kThreadsR = 256;
constexpr int kMaxRecFloats = 1024;   // 4 KiB records at most (128 KiB staged per CTA)

struct RecFields {
  mnr_row_field_t f[MNR_MAX_ROW_FIELDS];
  int off[MNR_MAX_ROW_FIELDS];   // first float of the field inside the record (multiple of 4)
  int n;
};

__device__ __forceinline__ float clamp_nan(float y, float lo, float hi) { return (lo <= hi && y == y) ? fminf(fmaxf(y, lo), hi) : y; }

// field that owns float `e` of the record, or -1 (padding)
__device__ __forceinline__ int field_of(const RecFields& rf, int e) {
  int k = -1;
#pragma unroll
  for (int j = 0; j < MNR_MAX_ROW_FIELDS; ++j)
    if (j < rf.n && e >= rf.off[j] && e < rf.off[j] + ((rf.f[j].width + 3) & ~3)) k = j;
  return k;
}

__global__ void __launch_bounds__(kThreadsR) sample_records_kernel(const __grid_constant__ RecFields rf, const float* __restrict__ records,
                                                                   int rec_floats, const int64_t* __restrict__ idx, int64_t n_idx) {
  extern __shared__ __align__(128) float s_dyn[];
  // [32][rec_floats] staged records | [rec_floats] mean | [rec_floats] denom (per record position; 0 / 1 outside normalised fields)
  float* s_rec = s_dyn;
  float* s_mean = s_dyn + (size_t)kRowsPerCta * rec_floats;
  float* s_denom = s_mean + rec_floats;
  // per 16 B piece of a record: where it goes (built while the TMA copies are in flight)
  __shared__ float* s_dst[kMaxRecFloats / 4];
  __shared__ int s_ld[kMaxRecFloats / 4];
  __shared__ short s_valid[kMaxRecFloats / 4];   // 0: padding, 1..4 elements
  __shared__ unsigned char s_kind[kMaxRecFloats / 4];
  __shared__ float s_clamp[kMaxRecFloats / 4][2];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int64_t s_idx[kRowsPerCta];
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * kRowsPerCta;
  const int rows = (int)min((int64_t)kRowsPerCta, n_idx - row0);
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (tid < rows) s_idx[tid] = __ldg(idx + row0 + tid);
  __syncthreads();
  const uint32_t rec_bytes = (uint32_t)rec_floats * 4u;
  if (tid < 32) {   // warp 0: one bulk copy per row, all in flight on one mbarrier
    if (tid == 0) ptx::mbar_expect_tx(&bar, rec_bytes * (uint32_t)rows);
    __syncwarp();
    if (tid < rows) ptx::bulk_g2s(s_rec + (size_t)tid * rec_floats, records + (size_t)s_idx[tid] * rec_floats, rec_bytes, &bar);
  }
  const int f4_per_row = rec_floats >> 2;
  for (int p = tid; p < f4_per_row; p += kThreadsR) {   // the plan of one record, under the copies' latency
    const int e0 = p << 2;
    const int k = field_of(rf, e0);
    short valid = 0;
    if (k >= 0) {
      const mnr_row_field_t& fd = rf.f[k];
      const int fe = e0 - rf.off[k];
      valid = (short)min(4, fd.width - fe);
      s_dst[p] = static_cast<float*>(fd.dst) + fe;
      s_ld[p] = fd.dst_ld > 0 ? fd.dst_ld : fd.width;
      s_kind[p] = (unsigned char)fd.kind;
      s_clamp[p][0] = fd.clamp_lo, s_clamp[p][1] = fd.clamp_hi;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool nz = fd.kind == MNR_ROW_F32_NORMALIZE && j < valid;
        s_mean[e0 + j] = nz ? __ldg(fd.mean + fe + j) : 0.f;
        s_denom[e0 + j] = nz ? __ldg(fd.denom + fe + j) : 1.f;
      }
    }
    s_valid[p] = valid;
  }
  __syncthreads();
  ptx::mbar_wait(&bar, 0);
  for (int q = tid; q < rows * f4_per_row; q += kThreadsR) {
    const int r = q / f4_per_row, p = q - r * f4_per_row;
    const int valid = s_valid[p];
    if (valid == 0) continue;
    const float4 v = *reinterpret_cast<const float4*>(s_rec + (size_t)r * rec_floats + (p << 2));
    float x[4] = {v.x, v.y, v.z, v.w};
    const int kind = s_kind[p];
    if (kind == MNR_ROW_F32_NORMALIZE) {   // RunningMeanStd.normalize (normalizers.py:65-68), separately rounded ops
      const float4 m = *reinterpret_cast<const float4*>(s_mean + (p << 2)), d = *reinterpret_cast<const float4*>(s_denom + (p << 2));
      const float lo = s_clamp[p][0], hi = s_clamp[p][1];
      x[0] = clamp_nan(__fdiv_rn(__fsub_rn(x[0], m.x), d.x), lo, hi), x[1] = clamp_nan(__fdiv_rn(__fsub_rn(x[1], m.y), d.y), lo, hi);
      x[2] = clamp_nan(__fdiv_rn(__fsub_rn(x[2], m.z), d.z), lo, hi), x[3] = clamp_nan(__fdiv_rn(__fsub_rn(x[3], m.w), d.w), lo, hi);
    } else if (kind == MNR_ROW_BOOL_TO_F32) {
      x[0] = (__float_as_uint(x[0]) & 0xffu) ? 1.0f : 0.0f;   // byte 0 of the slot is the bool (done.float(), simple_replay.py:84)
    }
    float* dst = s_dst[p] + (size_t)(row0 + r) * s_ld[p];
    if (valid == 4 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
      *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < valid) dst[j] = x[j];
    }
  }
}

// ring: row i of the incoming chunk -> slot; -1 when the slot is overwritten by the tail part of the same insert
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

__global__ void __launch_bounds__(kThreadsR) insert_records_kernel(const __grid_constant__ RecFields rf, float* __restrict__ records, int rec_floats,
                                                                   int64_t n_rows, int64_t next_p, int64_t capacity) {
  const int f4_per_row = rec_floats >> 2;
  const int64_t total = n_rows * f4_per_row, stride = (int64_t)gridDim.x * kThreadsR;
  for (int64_t q = (int64_t)blockIdx.x * kThreadsR + threadIdx.x; q < total; q += stride) {
    const int64_t i = q / f4_per_row;
    const int e0 = (int)(q - i * f4_per_row) << 2;
    const int64_t slot = ring_slot(i, n_rows, next_p, capacity);
    if (slot < 0) continue;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = field_of(rf, e0);
    if (k >= 0) {
      const mnr_row_field_t& fd = rf.f[k];
      const int fe = e0 - rf.off[k];
      const int valid = min(4, fd.width - fe);
      if (fd.kind == MNR_ROW_F32) {
        const float* src = static_cast<const float*>(fd.src) + (size_t)i * fd.width + fe;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < valid) x[j] = __ldg(src + j);
      } else if (fd.kind == MNR_ROW_U8) {          // bool / u8 done -> byte 0 of the slot
        x[0] = __uint_as_float(__ldg(static_cast<const uint8_t*>(fd.src) + i) ? 1u : 0u);
      } else if (fd.kind == MNR_ROW_F32_TO_BOOL) {  // dones.bool() (simple_replay.py:39)
        x[0] = __uint_as_float(__ldg(static_cast<const float*>(fd.src) + i) != 0.0f ? 1u : 0u);
      }
    }
    *reinterpret_cast<float4*>(records + (size_t)slot * rec_floats + e0) = make_float4(x[0], x[1], x[2], x[3]);
  }
}

int pack_rec(const mnr_row_field_t* fields, const int* offsets, int n_fields, int rec_floats, bool sample, RecFields* out) {
  if (!fields || !offsets) return MNR_ENULL;
  if (n_fields <= 0 || n_fields > MNR_MAX_ROW_FIELDS) return MNR_ESHAPE;
  if (rec_floats <= 0 || rec_floats % 32 != 0 || rec_floats > kMaxRecFloats) return MNR_ESHAPE;
  out->n = n_fields;
  for (int k = 0; k < MNR_MAX_ROW_FIELDS; ++k) out->off[k] = 0;
  for (int k = 0; k < n_fields; ++k) {
    const mnr_row_field_t& f = fields[k];
    if (f.width <= 0 || offsets[k] < 0 || offsets[k] % 4 != 0 || offsets[k] + ((f.width + 3) & ~3) > rec_floats) return MNR_ESHAPE;
    if (sample) {
      if (!f.dst) return MNR_ENULL;
      if (f.kind != MNR_ROW_F32 && f.kind != MNR_ROW_BOOL_TO_F32 && f.kind != MNR_ROW_F32_NORMALIZE) return MNR_EUNSUP;
      if (f.kind == MNR_ROW_F32_NORMALIZE && (!f.mean || !f.denom)) return MNR_ENULL;
      if (f.dst_ld != 0 && f.dst_ld < f.width) return MNR_ESHAPE;
      if (f.kind == MNR_ROW_BOOL_TO_F32 && f.width != 1) return MNR_ESHAPE;
    } else {
      if (!f.src) return MNR_ENULL;
      if (f.kind != MNR_ROW_F32 && f.kind != MNR_ROW_U8 && f.kind != MNR_ROW_F32_TO_BOOL) return MNR_EUNSUP;
      if (f.kind != MNR_ROW_F32 && f.width != 1) return MNR_ESHAPE;
    }
    out->f[k] = f;
    out->off[k] = offsets[k];
  }
  return MNR_OK;
}

}  // namespace

extern "C" int mnr_replay_insert_records(const mnr_row_field_t* fields, const int* rec_offsets, int n_fields, float* records, int rec_floats,
                                         int64_t n_rows, int64_t next_p, int64_t capacity, mnr_stream_t stream) {
  MNR_CHECK_PTR(records);
  RecFields rf;
  const int rc = pack_rec(fields, rec_offsets, n_fields, rec_floats, false, &rf);
  if (rc != MNR_OK) return rc;
  if (!mnr_aligned(records, 128)) return MNR_EALIGN;
  if (n_rows < 0 || capacity <= 0 || next_p < 0 || next_p > capacity) return MNR_ESHAPE;
  if (next_p + n_rows - capacity > capacity) return MNR_ESHAPE;  // the reference's tail slice would not fit either
  if (n_rows == 0) return MNR_OK;
  const int64_t total = n_rows * (rec_floats / 4);
  int64_t grid = (total + kThreadsR - 1) / kThreadsR;
  const int64_t cap = (int64_t)mnr_sm_count() * 8;
  if (grid > cap) grid = cap;
  insert_records_kernel<<<(unsigned)grid, kThreadsR, 0, mnr_cs(stream)>>>(rf, records, rec_floats, n_rows, next_p, capacity);
  MNR_LAUNCH_RESULT();
}

extern "C" int mnr_replay_sample_records(const mnr_row_field_t* fields, const int* rec_offsets, int n_fields, const float* records,
                                         int rec_floats, const int64_t* idx, int64_t n_idx, mnr_stream_t stream) {
  MNR_CHECK_PTR(records); MNR_CHECK_PTR(idx);
  RecFields rf;
  const int rc = pack_rec(fields, rec_offsets, n_fields, rec_floats, true, &rf);
  if (rc != MNR_OK) return rc;
  if (!mnr_aligned(records, 128)) return MNR_EALIGN;
  if (n_idx < 0) return MNR_ESHAPE;
  if (n_idx == 0) return MNR_OK;
  const size_t smem = (size_t)(kRowsPerCta + 2) * rec_floats * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    const cudaError_t e = mnr_allow_max_dyn_smem(sample_records_kernel);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  const int64_t grid = (n_idx + kRowsPerCta - 1) / kRowsPerCta;
  sample_records_kernel<<<(unsigned)grid, kThreadsR, smem, mnr_cs(stream)>>>(rf, records, rec_floats, idx, n_idx);
  MNR_LAUNCH_RESULT();
}
