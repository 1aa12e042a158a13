// Shared pieces of the tensor-core (tcgen05) kernels: activation forms, the tf32 (hi, lo) split, tensor-map construction.
// Included by tc_linear.cu (one Linear per launch) and tc_tower.cu (a whole tower forward per launch).
#pragma once

#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

// ELU's negative branch, expm1(z) for z <= 0.  Default (variant 0): `__expf(z) - 1` = one MUFU.EX2 + 2 FP32 ops; its ABSOLUTE
// error is <= 1.2e-7 (one ulp of 1.0, the same rounding any f32 value next to 1 carries), i.e. 1e-7 of the activation scale
// against the 1e-5 parity bar - but its RELATIVE error grows as z -> 0- (1e-3 at z = -1e-4), where torch's expm1 stays at
// 1e-7.  Variants 1 / 2 restore the relative accuracy with a Taylor polynomial below |z| = 1/4 (degree 7) / 1/16 (degree 4).
// MEASURED on B200 (profiles/r02_summary.md, same box back to back, whole PPO update at cfg2): variant 0 5.63 ms, variant 2
// 6.36 ms, variant 1 6.50 ms - the epilogue is LSU / issue-bound, 4-8 extra instructions per output element cost 13-15 % of
// the update.  The full-shape parity tests (tests/test_gpu_fullshape.py) pass at 1e-5 with every variant, so the fast form
// is the default; build with -DMNR_ELU_VARIANT=2 for expm1-grade relative accuracy.
#ifndef MNR_ELU_VARIANT
#define MNR_ELU_VARIANT 0
#endif
// exp(z) as ONE multiply + MUFU.EX2 with flush-to-zero: `__expf` compiles to the non-ftz ex2.approx, whose denormal-input
// guard costs an FSETP, a PLOP3 and two predicated FMULs per element (4 of the ~17 instructions per output element the ncu
// source page of the forward epilogue showed).  Flushing is exact for ELU: a denormal z gives exp(z) - 1 = z ~ 0 either way
// and exp(z) underflows only below z = -87, where the result is -1.
__device__ __forceinline__ float exp_ftz(float z) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(z * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float expm1_neg(float z) {
#if MNR_ELU_VARIANT == 0
  return exp_ftz(z) - 1.f;
#elif MNR_ELU_VARIANT == 1
  float p = fmaf(z, 1.f / 5040.f, 1.f / 720.f);
  p = fmaf(p, z, 1.f / 120.f);
  p = fmaf(p, z, 1.f / 24.f);
  p = fmaf(p, z, 1.f / 6.f);
  p = fmaf(p, z, 0.5f);
  p = fmaf(p, z, 1.f);
  return z > -0.25f ? p * z : exp_ftz(z) - 1.f;
#else
  float p = fmaf(z, 1.f / 24.f, 1.f / 6.f);
  p = fmaf(p, z, 0.5f);
  p = fmaf(p, z, 1.f);
  return z > -0.0625f ? p * z : exp_ftz(z) - 1.f;
#endif
}

template <int ACT>
__device__ __forceinline__ float act_fwd_t(float z) {
  if (ACT == 1) return z > 0.f ? z : expm1_neg(z);
  if (ACT == 2) return fmaxf(z, 0.f);
  return z;
}
template <int ACT>
__device__ __forceinline__ float act_bwd_t(float y) {
  if (ACT == 1) return y > 0.f ? 1.f : y + 1.f;
  if (ACT == 2) return y > 0.f ? 1.f : 0.f;
  return 1.f;
}
__device__ __forceinline__ float act_fwd(float z, int act) {
  switch (act) {
    case 1: return z > 0.f ? z : expm1_neg(z);
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

}  // namespace
