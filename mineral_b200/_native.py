"""ctypes binding of the C-ABI library `libmineral_b200.so` (include/mineral_b200.h).

Plumbing only: PyTorch owns device memory and streams; every compute call goes through the
`mnr_*` entry points with raw device pointers.  There is NO fallback: if the library is missing
the import fails, and if a tensor is not a contiguous CUDA tensor of the expected dtype the call
raises.
"""

import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MINERAL_B200_LIB") or os.path.join(_HERE, "lib", "libmineral_b200.so")   # env: A/B builds (tools/)
CSRC_DIR = os.path.join(_HERE, "csrc")

c_f32p = C.c_void_p
c_ptr = C.c_void_p


class MineralB200Error(RuntimeError):
    pass


def build(verbose=False):
    """Compile the CUDA sources for sm_100a into lib/libmineral_b200.so (nvcc cross-compiles on CPU)."""
    out = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise MineralB200Error("building libmineral_b200.so failed")
    return LIB_PATH


# ---- structs (mirror include/mineral_b200.h) ------------------------------------------------
class WideField(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("width", C.c_int32), ("_pad", C.c_int32)]


class RowField(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("width", C.c_int32), ("kind", C.c_int32),
                ("mean", C.c_void_p), ("denom", C.c_void_p), ("clamp_lo", C.c_float), ("clamp_hi", C.c_float),
                ("dst_ld", C.c_int32), ("_pad", C.c_int32)]


class NStepIn(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("act", C.c_void_p), ("rew", C.c_void_p), ("nobs", C.c_void_p),
                ("done", C.c_void_p), ("done_is_f32", C.c_int32), ("_pad", C.c_int32)]


class NStepOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("act", C.c_void_p), ("rew", C.c_void_p), ("nobs", C.c_void_p), ("done", C.c_void_p),
                ("done_is_f32", C.c_int32), ("_pad", C.c_int32)]


class PPOLossCfg(C.Structure):
    _fields_ = [("dist_type", C.c_int32), ("bounds_type", C.c_int32), ("use_smooth_clamp", C.c_int32),
                ("clip_value_loss", C.c_int32), ("minstd", C.c_float), ("maxstd", C.c_float), ("e_clip", C.c_float),
                ("critic_coef", C.c_float), ("entropy_coef", C.c_float), ("bounds_loss_coef", C.c_float),
                ("soft_bound", C.c_float), ("grad_scale", C.c_float), ("metric_scale", C.c_float), ("dz_act", C.c_int32)]


class LinearGroup(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p), ("bias", C.c_void_p), ("saved", C.c_void_p),
                ("bias_grad", C.c_void_p), ("b_lo", C.c_void_p), ("N", C.c_int32), ("K", C.c_int32), ("lda", C.c_int32),
                ("ldb", C.c_int32)]


class SplitJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("hi", C.c_void_p), ("lo", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32),
                ("src_ld", C.c_int32), ("dst_ld", C.c_int32)]


class OptimState(C.Structure):
    _fields_ = [("lr", C.c_double), ("kl_sum", C.c_double), ("kl_count", C.c_int32), ("step", C.c_int32),
                ("grad_norm", C.c_float), ("clip_coef", C.c_float), ("step_size", C.c_float), ("bc2_sqrt", C.c_float),
                ("lr_f32", C.c_float), ("_pad", C.c_float)]


class LnGroup(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("z", "y", "gamma", "beta", "mean", "rstd", "dy", "dz", "dgamma", "dbeta")]


class TcReduceJob(C.Structure):
    _fields_ = [("groups", C.POINTER(LinearGroup)), ("n_groups", C.c_int32), ("M", C.c_int32), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("grid_cap", C.c_int32), ("_pad", C.c_int32)]


class FoldSeg(C.Structure):
    _fields_ = [("begin", C.c_int32), ("end", C.c_int32), ("dst", C.c_void_p), ("scale", C.c_float), ("_pad", C.c_int32)]


class FoldJob(C.Structure):
    _fields_ = [("parts", C.c_void_p), ("n_parts", C.c_int32), ("width", C.c_int32), ("n_segs", C.c_int32), ("_pad", C.c_int32),
                ("seg", FoldSeg * 8)]


class HeadsMetricsJob(C.Structure):
    _fields_ = [("loss_parts", C.c_void_p), ("metrics", C.c_void_p), ("M", C.c_int64), ("n_parts", C.c_int32), ("_pad", C.c_int32),
                ("critic_coef", C.c_float), ("entropy_coef", C.c_float), ("bounds_loss_coef", C.c_float), ("metric_scale", C.c_float)]


class FoldList(C.Structure):
    _fields_ = [("folds", FoldJob * 8), ("metrics", HeadsMetricsJob), ("n_folds", C.c_int32), ("has_metrics", C.c_int32)]


class TowerGroup(C.Structure):
    _fields_ = [("x", C.c_void_p), ("w_hi", C.c_void_p * 3), ("w_lo", C.c_void_p * 3), ("bias", C.c_void_p * 3), ("y", C.c_void_p * 3),
                ("ldx", C.c_int32), ("ldw", C.c_int32 * 3)]


class CriticSet(C.Structure):
    _fields_ = [("pred", C.c_void_p * 4), ("grad", C.c_void_p * 4)]


class SquashCfg(C.Structure):
    _fields_ = [("dist_type", C.c_int32), ("has_minlogstd", C.c_int32), ("has_maxlogstd", C.c_int32), ("minlogstd", C.c_float),
                ("maxlogstd", C.c_float), ("minstd", C.c_float), ("maxstd", C.c_float), ("_pad", C.c_int32)]


class PeerCommStruct(C.Structure):
    _fields_ = [("base", C.c_void_p * 8), ("rank", C.c_int32), ("world", C.c_int32)]


ROW_F32, ROW_U8, ROW_BOOL_TO_F32, ROW_F32_TO_BOOL, ROW_F32_NORMALIZE = 0, 1, 2, 3, 4
DIST_NORMAL, DIST_DREAMERV3_NORMAL = 0, 1
BOUNDS_BOUND, BOUNDS_REG = 0, 1
PPO_METRICS = 10
MAX_WIDE_FIELDS, MAX_ROW_FIELDS = 8, 12

i32, i64, f32, f64, sz, vp = C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t, C.c_void_p

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "mnr_abi_version": (i32, []),
    "mnr_error_string": (C.c_char_p, [i32]),
    "mnr_sm_count": (i32, []),
    "mnr_gae_f32": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, f32, f32, vp]),
    "mnr_critic_targets_f32": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, f32, f32, f32, i32, f32, f32, vp]),
    "mnr_post_rollout_workspace_bytes": (sz, [i32, i32]),
    "mnr_post_rollout_fused": (i32, [C.POINTER(WideField), i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, vp, sz, vp]),
    "mnr_post_rollout_finalize": (i32, [vp, vp, vp, i64, vp, vp, vp, i32, f32, f32, f32, vp, vp, sz, vp]),
    "mnr_rms_moments_workspace_bytes": (sz, [i64, i32, i64]),
    "mnr_rms_batch_moments_f32": (i32, [vp, i64, i32, i64, i32, vp, vp, vp, sz, vp]),
    "mnr_rms_merge": (i32, [vp, vp, vp, i32, vp, vp, f64, i32, vp]),
    "mnr_rms_schedule": (i32, [vp, vp, vp, i32, vp, vp, i64, i64, i32, i32, f32, vp, vp, vp]),
    "mnr_rms_prepare": (i32, [vp, vp, i32, i32, f32, vp, vp, vp]),
    "mnr_rms_apply_f32": (i32, [vp, vp, i64, i32, vp, vp, f32, f32, i32, vp]),
    "mnr_gather_rows": (i32, [C.POINTER(RowField), i32, vp, i64, i64, vp]),
    "mnr_scatter_rows": (i32, [C.POINTER(RowField), i32, vp, i64, i64, vp]),
    "mnr_replay_insert": (i32, [C.POINTER(RowField), i32, i64, i64, i64, vp]),
    "mnr_nstep_push": (i32, [C.POINTER(NStepIn), C.POINTER(NStepIn), C.POINTER(NStepOut), C.POINTER(NStepOut), vp, i32, i32, i32, i32, i32, i64, vp]),
    "mnr_ppo_loss_workspace_bytes": (sz, [i64, i32]),
    "mnr_ppo_loss_fwd_bwd": (i32, [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i64, i32, C.POINTER(PPOLossCfg), vp, sz, vp]),
    "mnr_ppo_heads_workspace_bytes": (sz, [i64, i32, i32]),
    "mnr_ppo_heads_is_fused": (i32, [i32, i32]),
    "mnr_ppo_heads_loss_fwd_bwd": (i32, [vp] * 15 + [i32] + [vp] * 14 + [i64, i32, i32, C.POINTER(PPOLossCfg), vp, sz, vp]),
    "mnr_actor_loss_rows": (i32, [vp, vp, vp, i64, f32, i32, vp, vp, vp, vp, sz, vp]),
    "mnr_critic_loss_rows": (i32, [vp, vp, vp, i64, f32, i32, vp, vp, vp, vp, sz, vp]),
    "mnr_bounds_loss_rows": (i32, [vp, i64, i32, i32, f32, vp, vp, vp]),
    "mnr_policy_kl_rows": (i32, [vp, vp, vp, vp, i64, i32, vp, vp, vp, sz, vp]),
    "mnr_rows_reduce_workspace_bytes": (sz, [i64]),
    "mnr_linear_fwd": (i32, [C.POINTER(LinearGroup), i32, i32, i32, vp]),
    "mnr_linear_bwd_input": (i32, [C.POINTER(LinearGroup), i32, i32, i32, i32, vp]),
    "mnr_linear_bwd_weight_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "mnr_linear_bwd_weight": (i32, [C.POINTER(LinearGroup), i32, i32, i32, vp, sz, vp]),
    "mnr_tc_supported": (i32, []),
    "mnr_tc_linear_fwd": (i32, [C.POINTER(LinearGroup), i32, i32, i32, vp]),
    "mnr_tc_linear_bwd_input": (i32, [C.POINTER(LinearGroup), i32, i32, i32, vp]),
    "mnr_tc_linear_bwd_weight_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "mnr_tc_linear_bwd_weight": (i32, [C.POINTER(LinearGroup), i32, i32, i32, vp, sz, vp]),
    "mnr_tc_linear_bwd_weight_partials": (i32, [C.POINTER(LinearGroup), i32, i32, i32, vp, sz, vp]),
    "mnr_tc_reduce_partials": (i32, [C.POINTER(TcReduceJob), i32, vp]),
    "mnr_fold_defer": (i32, [C.POINTER(FoldList)]),
    "mnr_tc_reduce_partials_folds": (i32, [C.POINTER(TcReduceJob), i32, C.POINTER(FoldList), vp]),
    "mnr_tc_trace": (i32, [i32, vp]),
    "mnr_tc_grid_cap": (i32, [i32]),
    "mnr_tc_cache_hints": (i32, [i32]),
    "mnr_tc_tower_fwd_supported": (i32, [i32, C.POINTER(C.c_int32), i32, i32]),
    "mnr_tc_tower_fwd": (i32, [C.POINTER(TowerGroup), i32, i32, i32, C.POINTER(C.c_int32), i32, i32, vp]),
    "mnr_tc_tower_trace": (i32, [i32, vp]),
    "mnr_heads_trace": (i32, [i32, vp]),
    "mnr_split_tf32": (i32, [C.POINTER(SplitJob), i32, vp]),
    "mnr_rms_apply_ld_f32": (i32, [vp, vp, i64, i32, i32, vp, vp, f32, f32, vp]),
    "mnr_optim_queue_normalise": (i32, [vp, vp, i64, i32, i32, vp, vp, f32, f32]),
    "mnr_colsum_workspace_bytes": (sz, [i32]),
    "mnr_colsum_f32": (i32, [vp, i64, i32, vp, vp, sz, vp]),
    "mnr_grad_norm_workspace_bytes": (sz, [i64]),
    "mnr_grad_norm_prepare": (i32, [vp, i64, f32, f32, f32, vp, vp, vp, vp, sz, vp]),
    "mnr_adam_step": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, i32, vp, vp]),
    "mnr_kl_lr_update": (i32, [vp, f64, f64, f64, vp, vp, vp]),
    "mnr_lr_table_update": (i32, [vp, vp, vp, vp, vp]),
    "mnr_ln_act_fwd": (i32, [C.POINTER(LnGroup), i32, i64, i32, f32, i32, vp]),
    "mnr_ln_act_bwd_workspace_bytes": (sz, [i32, i64, i32]),
    "mnr_ln_act_bwd": (i32, [C.POINTER(LnGroup), i32, i64, i32, i32, vp, sz, vp]),
    "mnr_policy_step": (i32, [vp] * 11 + [f32, f32] + [vp] * 6 + [i32, i32, i32, C.POINTER(PPOLossCfg), vp]),
    "mnr_post_step": (i32, [vp, vp, vp, vp, f32, f32, vp, vp, i32, vp]),
    "mnr_soft_update": (i32, [vp, vp, i64, f32, f32, vp]),
    "mnr_td_target": (i32, [C.POINTER(CriticSet), i32, vp, vp, vp, f32, f32, vp, i64, vp]),
    "mnr_distl_projection": (i32, [vp, vp, vp, vp, f32, f32, f32, f32, i32, vp, i64, vp]),
    "mnr_replay_insert_records": (i32, [vp, vp, i32, vp, i32, i64, i64, i64, vp]),
    "mnr_replay_sample_records": (i32, [vp, vp, i32, vp, i32, vp, i64, vp]),
    "mnr_distl_workspace_bytes": (sz, []),
    "mnr_distl_target": (i32, [C.POINTER(CriticSet), i32, vp, vp, vp, f32, f32, f32, f32, i32, vp, i64, vp]),
    "mnr_distl_bce_fwd_bwd": (i32, [C.POINTER(CriticSet), i32, vp, i64, i32, vp, vp, sz, vp]),
    "mnr_distl_actor_seed": (i32, [C.POINTER(CriticSet), i32, vp, i64, i32, vp, vp, sz, vp]),
    "mnr_episode_scores": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "mnr_fused_adam_workspace_bytes": (sz, []),
    "mnr_fused_adam_split_table_offset": (sz, []),
    "mnr_fused_clip_adam": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, i32, vp, vp, vp, i32, vp, vp, sz, vp]),
    "mnr_critic_mse_workspace_bytes": (sz, [i64]),
    "mnr_critic_mse_fwd_bwd": (i32, [C.POINTER(CriticSet), i32, vp, i64, vp, vp, sz, vp]),
    "mnr_critic_mse_sum_fwd_bwd": (i32, [C.POINTER(CriticSet), i32, vp, i64, vp, vp, sz, vp]),
    "mnr_traj_store_step": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp]),
    "mnr_copy_cols_f32": (i32, [vp, i32, vp, i32, i64, i32, vp]),
    "mnr_squashed_sample": (i32, [vp, vp, i32, vp, i64, i32, C.POINTER(SquashCfg), vp, i32, vp, vp]),
    "mnr_squashed_bwd": (i32, [vp, vp, i32, vp, vp, vp, i32, vp, f32, i64, i32, C.POINTER(SquashCfg), vp, vp, vp]),
    "mnr_tanh_action": (i32, [vp, vp, i64, i32, f32, f32, f32, f32, vp, i32, vp]),
    "mnr_tanh_action_bwd": (i32, [vp, vp, vp, i32, i64, i32, vp, vp]),
    "mnr_td_target_stats": (i32, [C.POINTER(CriticSet), i32, vp, vp, vp, vp, f32, f32, vp, i64, vp, vp]),
    "mnr_actor_seed": (i32, [C.POINTER(CriticSet), i32, vp, vp, f32, f32, i64, vp, vp, vp]),
    "mnr_peer_alloc": (i32, [sz, C.POINTER(vp), vp]),
    "mnr_peer_open": (i32, [vp, C.POINTER(vp)]),
    "mnr_peer_close": (i32, [vp]),
    "mnr_peer_free": (i32, [vp]),
    "mnr_peer_error": (i32, [C.POINTER(PeerCommStruct), C.POINTER(i32)]),
    "mnr_peer_allgather": (i32, [C.POINTER(PeerCommStruct), sz, vp, sz, vp, vp]),
    "mnr_peer_reduce_clip_adam": (i32, [C.POINTER(PeerCommStruct), sz, i64, i32, vp, vp, vp, vp, f32, f32, f32, f32, f32, i32, i32, vp,
                                        vp, i32, vp, vp, sz, vp]),
    "mnr_peer_reduce_clip_adam_pp": (i32, [C.POINTER(PeerCommStruct), sz, sz, i32, i64, i32, vp, vp, vp, vp, f32, f32, f32, f32, f32, i32, i32,
                                           vp, vp, i32, vp, vp, vp, sz, vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: mineral_b200 has no CPU or eager fallback. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C mineral_b200/csrc`)."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.mnr_abi_version() != 1:
        raise ImportError("libmineral_b200.so ABI version mismatch; rebuild it")
    return lib


lib = _load()


LAUNCHES = 0  # kernels enqueued through this binding (each successful mnr_* call launches >= 1 kernel)


_DEBUG_SYNC = bool(os.environ.get("MNR_DEBUG_SYNC"))  # debugging aid: synchronise after every call so a device fault names its launch


def check(rc, what="", kernels=1):
    global LAUNCHES
    if rc != 0:
        raise MineralB200Error(f"{what or 'mnr call'} failed: [{rc}] {lib.mnr_error_string(rc).decode()}")
    LAUNCHES += kernels
    if _DEBUG_SYNC:
        try:
            torch.cuda.synchronize()
        except Exception as e:
            raise MineralB200Error(f"device fault surfaced after {what} (launch #{LAUNCHES})") from e


def stream():
    """cudaStream_t of torch's current stream (so work is ordered with torch ops and graph capture)."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, dtype=None, name="tensor", allow_none=False):
    """Raw device pointer of a contiguous CUDA tensor; loud failure otherwise."""
    if t is None:
        if allow_none:
            return None
        raise MineralB200Error(f"{name} is None")
    if not t.is_cuda:
        raise MineralB200Error(f"{name} must be a CUDA tensor: mineral_b200 has no CPU path (got device {t.device})")
    if not t.is_contiguous():
        raise MineralB200Error(f"{name} must be contiguous")
    if dtype is not None:
        ok = t.dtype in dtype if isinstance(dtype, (tuple, list)) else t.dtype == dtype
        if not ok:
            raise MineralB200Error(f"{name} has dtype {t.dtype}, expected {dtype}")
    return C.c_void_p(t.data_ptr())


def ptr_view(t, name="tensor"):
    """Device pointer of a CUDA view with unit stride in its last dimension (rows may be strided; the caller passes the stride)."""
    if t is None or not t.is_cuda or t.stride(-1) != 1:
        raise MineralB200Error(f"{name} must be a CUDA view with unit column stride")
    return C.c_void_p(t.data_ptr())


def ptr_strided(t, name="tensor"):
    """Device pointer + row stride (floats) of a 2-D f32 CUDA view whose rows are contiguous (e.g. a column slice of a padded buffer)."""
    if t is None or not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise MineralB200Error(f"{name} must be a 2-D f32 CUDA tensor with unit column stride")
    return C.c_void_p(t.data_ptr()), int(t.stride(0))


_WS = {}


def workspace(key, nbytes, device):
    """Zero-initialised persistent scratch (the kernels leave their ticket counters at zero)."""
    k = (key, str(device))
    t = _WS.get(k)
    if t is None or t.numel() < nbytes:
        t = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _WS[k] = t
    return t
