// Host-side odds and ends of the C ABI.
#include <stdlib.h>

#include "common.cuh"

bool mnr_pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("MNR_PDL");
    cached = (e && e[0] == '1') ? 1 : 0;   // measured on B200 (profiles/r01_summary.md section 10): 1-5 % SLOWER inside the update graph -> opt-in
  }
  return cached == 1;
}

static thread_local mnr_fold_list_t* t_fold_list = nullptr;
mnr_fold_list_t* mnr_fold_target() { return t_fold_list; }
mnr_fold_job_t* mnr_fold_next() {
  if (!t_fold_list || t_fold_list->n_folds >= MNR_MAX_FOLDS) return nullptr;
  return &t_fold_list->folds[t_fold_list->n_folds];
}
extern "C" int mnr_fold_defer(mnr_fold_list_t* list) {
  t_fold_list = list;
  if (list) list->n_folds = 0, list->has_metrics = 0;
  return MNR_OK;
}

extern "C" int mnr_abi_version(void) { return MNR_ABI_VERSION; }

extern "C" const char* mnr_error_string(int code) {
  switch (code) {
    case MNR_OK: return "ok";
    case MNR_ENULL: return "required pointer is NULL";
    case MNR_ESHAPE: return "non-positive or unsupported size";
    case MNR_EALIGN: return "pointer not aligned as the kernel needs";
    case MNR_EUNSUP: return "unsupported mode";
    case MNR_EWORKSPACE: return "workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown mineral_b200 error";
}

extern "C" int mnr_sm_count(void) {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) {
    cached = sms;
    return sms;
  }
  (void)cudaGetLastError();  // no device (CPU-only build box): report the B200's count, do not cache
  return 148;
}
