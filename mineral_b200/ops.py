"""Tensor-level wrappers over the C ABI: one Python function per `mnr_*` entry point.

Everything here enqueues a hand-written sm_100a kernel on torch's current CUDA stream; nothing
computes with torch ops and nothing runs on the CPU.
"""

import ctypes as C

import torch

from . import _native as nv
from ._native import check, lib, ptr, stream

F32, F64, U8, I64, BOOL = torch.float32, torch.float64, torch.uint8, torch.int64, torch.bool


# Python-double scalars (gamma * tau, 1.0 - lam, ...) are formed in double precision exactly as the
# reference forms them and are rounded to f32 by ctypes (c_float) when they cross the ABI - the same
# rounding torch applies when a Python scalar meets an f32 tensor.

# ---------------------------------------------------------------------------------------------
# K1  scans
# ---------------------------------------------------------------------------------------------
def gae(rewards, values, dones, last_values, gamma, tau, returns=None, advantages=None):
    """GAE over time-major (T,N[,1]) tensors -> returns (same shape as rewards).
    mineral/agents/ppo/experience.py:96-108."""
    T, N = rewards.shape[0], rewards.shape[1]
    if returns is None:
        returns = torch.empty_like(rewards)
    check(lib.mnr_gae_f32(ptr(rewards, F32, "rewards"), ptr(values, F32, "values"), ptr(dones, (U8, BOOL), "dones"),
                          ptr(last_values, F32, "last_values"), ptr(returns, F32, "returns"),
                          ptr(advantages, F32, "advantages", allow_none=True), T, N, gamma, gamma * tau, stream()), "mnr_gae_f32")
    return returns


def critic_targets(rew, done_mask, next_values, gamma, lam=0.95, method="td-lambda", alpha=None, entropy=None,
                   negate_entropy=False, offset_by_target_entropy=False, scale_by_target_entropy=False,
                   target_entropy=-1.0, out=None):
    """SHAC/SAPO critic targets, (T,N) f32.  mineral/agents/diffrl/shac.py:877-916."""
    T, N = rew.shape
    if out is None:
        out = torch.empty_like(rew)
    m = {"one-step": 0, "td-lambda": 1}.get(method)
    if m is None:
        raise NotImplementedError(method)
    flags = (1 if negate_entropy else 0) | (2 if offset_by_target_entropy else 0) | (4 if scale_by_target_entropy else 0)
    at = abs(target_entropy)
    check(lib.mnr_critic_targets_f32(
        ptr(rew, F32, "rew"), ptr(done_mask, F32, "done_mask", allow_none=(m == 0)), ptr(next_values, F32, "next_values"),
        ptr(entropy, F32, "entropy", allow_none=True) if alpha is not None else None, ptr(out, F32, "target"), T, N, m,
        gamma, lam, 1.0 - lam, lam * gamma, 0.0 if alpha is None else float(alpha), flags, at, (1.0 / at) if at else 0.0,
        stream()), "mnr_critic_targets_f32")
    return out


# ---------------------------------------------------------------------------------------------
# K1a+K3a fused post-rollout
# ---------------------------------------------------------------------------------------------
def post_rollout_workspace(T, N, device):
    return torch.zeros(lib.mnr_post_rollout_workspace_bytes(T, N), dtype=U8, device=device)


def post_rollout_fused(wide, rewards, values, dones, neglogp, last_values, returns_tm, out, T, N, gamma, tau, ws, do_gae=True):
    """wide: list of (src (T,N,w), dst (N*T,w)); out: dict with env-major 'rewards','values','returns','neglogp',
    'dones','advantages'.  One launch: GAE + pack + moment partials."""
    arr = (nv.WideField * max(len(wide), 1))()
    for i, (s, d) in enumerate(wide):
        w = s.numel() // (T * N)
        arr[i] = nv.WideField(ptr(s, F32, "wide src").value, ptr(d, F32, "wide dst").value, w, 0)
    check(lib.mnr_post_rollout_fused(
        arr, len(wide), ptr(rewards, F32, "rewards"), ptr(values, F32, "values"), ptr(dones, (U8, BOOL), "dones"),
        ptr(neglogp, F32, "neglogp"), ptr(last_values, F32, "last_values", allow_none=not do_gae),
        ptr(returns_tm, F32, "returns_tm", allow_none=do_gae), ptr(out["rewards"], F32), ptr(out["values"], F32),
        ptr(out["returns"], F32), ptr(out["neglogp"], F32), ptr(out["dones"], (U8, BOOL)), ptr(out["advantages"], F32),
        T, N, 1 if do_gae else 0, gamma, gamma * tau, ptr(ws, U8), ws.numel(), stream()), "mnr_post_rollout_fused")


def post_rollout_finalize(out, B, value_rms, ws, eps=1e-5, clamp=(-5.0, 5.0), stats_out=None):
    """Normalise advantages; if value_rms (mean,var,count tensors) is given, run PPO's value block in place."""
    if value_rms is None:
        vm = vv = vc = None
        f64 = 0
    else:
        vm, vv, vc = value_rms
        f64 = 1 if vm.dtype == F64 else 0
    lo, hi = clamp if clamp is not None else (1.0, -1.0)
    check(lib.mnr_post_rollout_finalize(
        ptr(out["advantages"], F32), ptr(out["values"], F32), ptr(out["returns"], F32), B,
        ptr(vm, (F32, F64), allow_none=True), ptr(vv, (F32, F64), allow_none=True), ptr(vc, (F32, F64), allow_none=True), f64,
        eps, lo, hi, ptr(stats_out, F64, allow_none=True), ptr(ws, U8), ws.numel(), stream()), "mnr_post_rollout_finalize", kernels=2)


# ---------------------------------------------------------------------------------------------
# K2  RunningMeanStd
# ---------------------------------------------------------------------------------------------
def rms_moments_workspace(rows, cols, group_rows, device):
    return torch.zeros(lib.mnr_rms_moments_workspace_bytes(rows, cols, group_rows), dtype=U8, device=device)


def rms_batch_moments(x2d, group_rows=None, correction=1, ws=None, mean_out=None, var_out=None):
    """x2d (rows, cols) f32 -> per-group column mean/var (n_groups, cols) f32."""
    rows, cols = x2d.shape
    group_rows = rows if group_rows is None else group_rows
    ng = (rows + group_rows - 1) // group_rows
    if mean_out is None:
        mean_out = torch.empty(ng, cols, dtype=F32, device=x2d.device)
    if var_out is None:
        var_out = torch.empty(ng, cols, dtype=F32, device=x2d.device)
    if ws is None:
        ws = nv.workspace("rms_moments", lib.mnr_rms_moments_workspace_bytes(rows, cols, group_rows), x2d.device)
    check(lib.mnr_rms_batch_moments_f32(ptr(x2d, F32, "x"), rows, cols, group_rows, correction, ptr(mean_out, F32),
                                        ptr(var_out, F32), ptr(ws, U8), ws.numel(), stream()), "mnr_rms_batch_moments_f32")
    return mean_out, var_out


def rms_merge(mean, var, count, batch_mean, batch_var, batch_count):
    f64 = 1 if mean.dtype == F64 else 0
    check(lib.mnr_rms_merge(ptr(mean, (F32, F64)), ptr(var, mean.dtype), ptr(count, mean.dtype), f64, ptr(batch_mean, F32),
                            ptr(batch_var, F32), float(batch_count), mean.numel(), stream()), "mnr_rms_merge")


def rms_schedule(mean, var, count, group_mean, group_var, group_rows, rows, n_steps, eps, mean_tab, denom_tab):
    f64 = 1 if mean.dtype == F64 else 0
    check(lib.mnr_rms_schedule(ptr(mean, (F32, F64)), ptr(var, mean.dtype), ptr(count, mean.dtype), f64, ptr(group_mean, F32),
                               ptr(group_var, F32), group_rows, rows, n_steps, mean.numel(), eps, ptr(mean_tab, F32),
                               ptr(denom_tab, F32), stream()), "mnr_rms_schedule")


def rms_prepare(mean, var, eps, mean_f32=None, denom_f32=None):
    cols = mean.numel()
    if mean_f32 is None:
        mean_f32 = torch.empty(cols, dtype=F32, device=mean.device)
    if denom_f32 is None:
        denom_f32 = torch.empty(cols, dtype=F32, device=mean.device)
    check(lib.mnr_rms_prepare(ptr(mean, (F32, F64)), ptr(var, mean.dtype), 1 if mean.dtype == F64 else 0, cols, eps,
                              ptr(mean_f32, F32), ptr(denom_f32, F32), stream()), "mnr_rms_prepare")
    return mean_f32, denom_f32


def rms_apply(x, mean_f32, denom_f32, clamp=None, inverse=False, out=None):
    cols = mean_f32.numel()
    rows = x.numel() // cols
    if out is None:
        out = torch.empty_like(x)
    lo, hi = clamp if clamp is not None else (1.0, -1.0)
    check(lib.mnr_rms_apply_f32(ptr(x, F32, "x"), ptr(out, F32, "out"), rows, cols, ptr(mean_f32, F32), ptr(denom_f32, F32),
                                lo, hi, 1 if inverse else 0, stream()), "mnr_rms_apply_f32")
    return out


# ---------------------------------------------------------------------------------------------
# K3  rows
# ---------------------------------------------------------------------------------------------
def _row_fields(fields):
    """fields: list of (src, dst, width, kind[, mean, denom, lo, hi[, dst_ld]]).  With dst_ld, dst may be a column block of a
    wider row-major buffer (rows dst_ld elements apart)."""
    arr = (nv.RowField * len(fields))()
    for i, f in enumerate(fields):
        src, dst, width, kind = f[:4]
        mean = ptr(f[4], F32).value if len(f) > 4 and f[4] is not None else None
        denom = ptr(f[5], F32).value if len(f) > 5 and f[5] is not None else None
        lo, hi = (f[6], f[7]) if len(f) > 7 and f[6] is not None else (1.0, -1.0)
        dst_ld = int(f[8]) if len(f) > 8 and f[8] else 0
        if dst_ld:
            dptr = nv.ptr_view(dst, "strided row dst").value
        else:
            dptr = ptr(dst, name="row dst").value
        arr[i] = nv.RowField(ptr(src, name="row src").value, dptr, width, kind, mean, denom, lo, hi, dst_ld, 0)
    return arr


def gather_rows(fields, idx=None, row0=0, n=None):
    n = idx.numel() if idx is not None else n
    check(lib.mnr_gather_rows(_row_fields(fields), len(fields), ptr(idx, I64, "idx", allow_none=True), row0, n, stream()),
          "mnr_gather_rows")


def scatter_rows(fields, idx=None, row0=0, n=None):
    n = idx.numel() if idx is not None else n
    check(lib.mnr_scatter_rows(_row_fields(fields), len(fields), ptr(idx, I64, "idx", allow_none=True), row0, n, stream()),
          "mnr_scatter_rows")


def replay_insert(fields, n_rows, next_p, capacity):
    check(lib.mnr_replay_insert(_row_fields(fields), len(fields), n_rows, next_p, capacity, stream()), "mnr_replay_insert")


def replay_insert_records(fields, offsets, records, n_rows, next_p, capacity):
    """fields as in replay_insert (dst ignored: pass the records tensor); offsets: first float of each field in a record."""
    arr = (C.c_int32 * len(offsets))(*offsets)
    check(lib.mnr_replay_insert_records(_row_fields(fields), arr, len(fields), ptr(records, F32, "records"), records.shape[1], n_rows, next_p,
                                        capacity, stream()), "mnr_replay_insert_records")


def replay_sample_records(fields, offsets, records, idx):
    """fields as in gather_rows (src ignored: pass the records tensor); one TMA-staged launch for the whole batch."""
    arr = (C.c_int32 * len(offsets))(*offsets)
    check(lib.mnr_replay_sample_records(_row_fields(fields), arr, len(fields), ptr(records, F32, "records"), records.shape[1],
                                        ptr(idx, I64, "idx"), idx.numel(), stream()), "mnr_replay_sample_records")


def nstep_push(inp, win_in, win_out, out, gammas, N, S, O, A, nstep, count_before):
    def _in(d, f32_done):
        return nv.NStepIn(ptr(d["obs"], F32).value, ptr(d["act"], F32).value, ptr(d["rew"], F32).value,
                          ptr(d["nobs"], F32).value, ptr(d["done"], (F32,) if f32_done else (U8, BOOL)).value,
                          1 if f32_done else 0, 0)

    def _out(d):
        if d is None:
            return nv.NStepOut(None, None, None, None, None, 0, 0)
        return nv.NStepOut(ptr(d["obs"], F32).value, ptr(d["act"], F32).value, ptr(d["rew"], F32).value,
                           ptr(d["nobs"], F32).value, ptr(d["done"], (U8, BOOL, F32)).value, 1 if d["done"].dtype == F32 else 0, 0)

    a, b = _in(inp, inp["done"].dtype == F32), _in(win_in, False)
    c, d = _out(win_out), _out(out)
    check(lib.mnr_nstep_push(C.byref(a), C.byref(b), C.byref(c), C.byref(d), ptr(gammas, F32), N, S, O, A, nstep,
                             count_before, stream()), "mnr_nstep_push")


# ---------------------------------------------------------------------------------------------
# K4  PPO loss
# ---------------------------------------------------------------------------------------------
def ppo_loss_workspace(mb, A, device):
    return torch.zeros(lib.mnr_ppo_loss_workspace_bytes(mb, A), dtype=U8, device=device)


def make_loss_cfg(dist_type="normal", minstd=1.0, maxstd=1.0, e_clip=0.2, use_smooth_clamp=False, clip_value_loss=False,
                  critic_coef=1.0, entropy_coef=0.0, bounds_loss_coef=0.0, bounds_type="bound", soft_bound=1.1, grad_scale=1.0, metric_scale=1.0,
                  dz_act=0):
    dt = {"normal": nv.DIST_NORMAL, "dreamerv3_normal": nv.DIST_DREAMERV3_NORMAL}.get(dist_type)
    if dt is None:
        raise NotImplementedError(f"dist_type {dist_type!r} is not on the accelerated PPO path")
    bt = {"bound": nv.BOUNDS_BOUND, "reg": nv.BOUNDS_REG}.get(bounds_type)
    if bt is None:
        raise NotImplementedError(bounds_type)
    return nv.PPOLossCfg(dt, bt, int(bool(use_smooth_clamp)), int(bool(clip_value_loss)), minstd, maxstd, e_clip,
                         critic_coef, entropy_coef, bounds_loss_coef, soft_bound, grad_scale, metric_scale, int(dz_act))


def ppo_loss_fwd_bwd(mu_raw, logstd, value, actions, old_neglogp, advantages, old_values, returns, old_mu, old_sigma,
                     g_mu_raw, g_logstd, g_value, metrics, cfg, ws, write_back=True, mu_mean=None, sigma_mean=None):
    mb, A = mu_raw.shape
    stride = 0 if logstd.dim() == 1 else A
    check(lib.mnr_ppo_loss_fwd_bwd(
        ptr(mu_raw, F32, "mu_raw"), ptr(logstd, F32, "logstd"), stride, ptr(value, F32, "value"), ptr(actions, F32, "actions"),
        ptr(old_neglogp, F32, "old_neglogp"), ptr(advantages, F32, "advantages"), ptr(old_values, F32, "old_values"),
        ptr(returns, F32, "returns"), ptr(old_mu, F32, "old_mu"), ptr(old_sigma, F32, "old_sigma"), 1 if write_back else 0,
        ptr(g_mu_raw, F32), ptr(g_logstd, F32), ptr(g_value, F32), ptr(metrics, F32), ptr(mu_mean, F32, allow_none=True),
        ptr(sigma_mean, F32, allow_none=True), mb, A, C.byref(cfg), ptr(ws, U8), ws.numel(), stream()), "mnr_ppo_loss_fwd_bwd")


def ppo_heads_workspace(M, A, H, device):
    return torch.zeros(lib.mnr_ppo_heads_workspace_bytes(M, A, H), dtype=U8, device=device)


def ppo_heads_loss_fwd_bwd(y3a, y3v, w_mu, b_mu, w_v, b_v, w_sig, b_sig, actions, old_neglogp, advantages, old_values, returns,
                           old_mu, old_sigma, dy3a, dy3v, g_mu, g_ls, g_v, dw_mu, db_mu, dw_v, db_v, dw_sig, db_sig, metrics, cfg, ws,
                           write_back=True, mu_mean=None, sigma_mean=None):
    M, H = y3a.shape
    A = w_mu.shape[0]
    o = lambda t: ptr(t, F32, allow_none=True)
    check(lib.mnr_ppo_heads_loss_fwd_bwd(
        ptr(y3a, F32, "y3a"), ptr(y3v, F32, "y3v"), ptr(w_mu, F32), ptr(b_mu, F32), ptr(w_v, F32), ptr(b_v, F32), o(w_sig), ptr(b_sig, F32),
        ptr(actions, F32, "actions"), ptr(old_neglogp, F32), ptr(advantages, F32), ptr(old_values, F32), ptr(returns, F32),
        ptr(old_mu, F32), ptr(old_sigma, F32), 1 if write_back else 0, ptr(dy3a, F32), ptr(dy3v, F32), ptr(g_mu, F32), ptr(g_ls, F32),
        ptr(g_v, F32), ptr(dw_mu, F32), ptr(db_mu, F32), ptr(dw_v, F32), ptr(db_v, F32), o(dw_sig), ptr(db_sig, F32), ptr(metrics, F32),
        o(mu_mean), o(sigma_mean), M, A, H, C.byref(cfg), ptr(ws, U8), ws.numel(), stream()), "mnr_ppo_heads_loss_fwd_bwd",
          kernels=1 if lib.mnr_ppo_heads_is_fused(A, H) else 2)


# ---------------------------------------------------------------------------------------------
# G  MLP towers
# ---------------------------------------------------------------------------------------------
ACT_CODES = {None: 0, "Identity": 0, "ELU": 1, "ReLU": 2}


def _linear_groups(groups):
    """groups: list of dict(a=, b=, c=, bias=None, saved=None, bias_grad=None, N=, K=)."""
    arr = (nv.LinearGroup * len(groups))()
    for i, g in enumerate(groups):
        def p(name):
            t = g.get(name)
            return None if t is None else ptr(t, F32, name).value
        arr[i] = nv.LinearGroup(p("a"), p("b"), p("c"), p("bias"), p("saved"), p("bias_grad"), p("b_lo"), g["N"], g["K"],
                                g.get("lda", 0), g.get("ldb", 0))
    return arr


def tc_supported():
    return bool(lib.mnr_tc_supported())


def tc_linear_fwd(groups, M, act):
    check(lib.mnr_tc_linear_fwd(_linear_groups(groups), len(groups), M, act, stream()), "mnr_tc_linear_fwd")


def tc_linear_bwd_input(groups, M, act):
    check(lib.mnr_tc_linear_bwd_input(_linear_groups(groups), len(groups), M, act, stream()), "mnr_tc_linear_bwd_input")


def tc_linear_bwd_weight(groups, M, act, ws):
    check(lib.mnr_tc_linear_bwd_weight(_linear_groups(groups), len(groups), M, act, ptr(ws, U8), ws.numel(), stream()),
          "mnr_tc_linear_bwd_weight", kernels=2)


_TC_GRID_CAP = 0


def tc_grid_cap(ctas):
    """CTAs the following tensor-core linear launches of this thread may use (0 = all SMs); returns the previous value."""
    global _TC_GRID_CAP
    _TC_GRID_CAP = int(ctas) if ctas and ctas > 0 else 0
    return lib.mnr_tc_grid_cap(_TC_GRID_CAP)


def tc_linear_bwd_weight_partials(groups, M, act, ws):
    """Tensor-core dW / db split partials into `ws` (no reduce); returns the job record for tc_reduce_partials."""
    arr = _linear_groups(groups)
    check(lib.mnr_tc_linear_bwd_weight_partials(arr, len(groups), M, act, ptr(ws, U8), ws.numel(), stream()),
          "mnr_tc_linear_bwd_weight_partials")
    return (arr, len(groups), M, ws, _TC_GRID_CAP)


def tc_reduce_partials(jobs, folds=None):
    """One launch: sum the split partials of up to 4 layers (records from tc_linear_bwd_weight_partials) and, with `folds`
    (a FoldList filled between fold_defer(list) and fold_defer(None)), the deferred cross-CTA sums of the heads + loss and
    LayerNorm-backward launches."""
    arr = (nv.TcReduceJob * len(jobs))()
    for i, (groups, n_groups, M, ws, cap) in enumerate(jobs):
        arr[i] = nv.TcReduceJob(groups, n_groups, M, ptr(ws, U8).value, ws.numel(), cap, 0)
    if folds is None:
        check(lib.mnr_tc_reduce_partials(arr, len(jobs), stream()), "mnr_tc_reduce_partials")
    else:
        check(lib.mnr_tc_reduce_partials_folds(arr, len(jobs), C.byref(folds), stream()), "mnr_tc_reduce_partials_folds")


def fold_defer(fold_list):
    """Register (a nv.FoldList) or clear (None) this thread's deferred-fold list (include/mineral_b200.h, mnr_fold_defer)."""
    check(lib.mnr_fold_defer(C.byref(fold_list) if fold_list is not None else None), "mnr_fold_defer")


def tc_tower_fwd_supported(K0, widths, act):
    w = (C.c_int32 * len(widths))(*widths)
    return bool(lib.mnr_tc_tower_fwd_supported(K0, w, len(widths), act))


def tc_tower_fwd(groups, M, K0, widths, act):
    """A whole 3-layer tower forward per launch.  groups: list of dict(x=, ldx=, w_hi=[3], w_lo=[3], ldw=[3], bias=[3], y=[3])."""
    arr = (nv.TowerGroup * len(groups))()
    P3 = C.c_void_p * 3
    for i, g in enumerate(groups):
        arr[i] = nv.TowerGroup(ptr(g["x"], F32, "x").value, P3(*[ptr(t, F32).value for t in g["w_hi"]]), P3(*[ptr(t, F32).value for t in g["w_lo"]]),
                               P3(*[ptr(t, F32).value for t in g["bias"]]), P3(*[ptr(t, F32).value for t in g["y"]]), g.get("ldx", 0),
                               (C.c_int32 * 3)(*g.get("ldw", (0, 0, 0))))
    w = (C.c_int32 * len(widths))(*widths)
    check(lib.mnr_tc_tower_fwd(arr, len(groups), M, K0, w, len(widths), act, stream()), "mnr_tc_tower_fwd")


def split_tf32(jobs):
    """jobs: list of (src (rows, cols) view with row stride src_ld, hi, lo (rows, dst_ld))."""
    arr = (nv.SplitJob * len(jobs))()
    for i, (src, hi, lo, rows, cols, src_ld, dst_ld) in enumerate(jobs):
        arr[i] = nv.SplitJob(ptr(src, F32).value, ptr(hi, F32).value, ptr(lo, F32).value, rows, cols, src_ld, dst_ld)
    check(lib.mnr_split_tf32(arr, len(jobs), stream()), "mnr_split_tf32")


def rms_apply_ld(x, mean_f32, denom_f32, clamp, out, out_ld):
    cols = mean_f32.numel()
    rows = x.numel() // cols
    lo, hi = clamp if clamp is not None else (1.0, -1.0)
    check(lib.mnr_rms_apply_ld_f32(ptr(x, F32, "x"), ptr(out, F32, "out"), rows, cols, out_ld, ptr(mean_f32, F32), ptr(denom_f32, F32),
                                   lo, hi, stream()), "mnr_rms_apply_ld_f32")
    return out


def optim_queue_normalise(x, mean_f32, denom_f32, clamp, out, out_ld):
    """rms_apply_ld as a ride-along job of the NEXT fused optimizer launch of this thread (mnr_optim_queue_normalise)."""
    cols = mean_f32.numel()
    rows = x.numel() // cols
    lo, hi = clamp if clamp is not None else (1.0, -1.0)
    check(lib.mnr_optim_queue_normalise(ptr(x, F32, "x"), ptr(out, F32, "out"), rows, cols, out_ld, ptr(mean_f32, F32), ptr(denom_f32, F32),
                                        lo, hi), "mnr_optim_queue_normalise", kernels=0)


def linear_fwd(groups, M, act):
    check(lib.mnr_linear_fwd(_linear_groups(groups), len(groups), M, act, stream()), "mnr_linear_fwd")


def linear_bwd_input(groups, M, act, accumulate=False):
    check(lib.mnr_linear_bwd_input(_linear_groups(groups), len(groups), M, act, 1 if accumulate else 0, stream()), "mnr_linear_bwd_input")


def linear_bwd_weight(groups, M, act, ws):
    check(lib.mnr_linear_bwd_weight(_linear_groups(groups), len(groups), M, act, ptr(ws, U8), ws.numel(), stream()),
          "mnr_linear_bwd_weight", kernels=2)


def _ln_groups(groups):
    """groups: list of dict(z=, y=, gamma=, beta=, mean=, rstd=, dy=, dz=, dgamma=, dbeta=) (missing -> NULL)."""
    arr = (nv.LnGroup * len(groups))()
    for i, g in enumerate(groups):
        def p(name):
            t = g.get(name)
            return None if t is None else ptr(t, F32, name).value
        arr[i] = nv.LnGroup(p("z"), p("y"), p("gamma"), p("beta"), p("mean"), p("rstd"), p("dy"), p("dz"), p("dgamma"), p("dbeta"))
    return arr


def ln_act_fwd(groups, M, N, eps, act):
    """y = act(LayerNorm(z)) for up to 2 (M,N) matrices in one launch; saves mean / rstd (csrc/norm_act.cu)."""
    check(lib.mnr_ln_act_fwd(_ln_groups(groups), len(groups), M, N, eps, act, stream()), "mnr_ln_act_fwd")


def ln_act_bwd_workspace(n_groups, M, N, device):
    return torch.zeros(lib.mnr_ln_act_bwd_workspace_bytes(n_groups, M, N), dtype=U8, device=device)


def ln_act_bwd(groups, M, N, act, ws):
    """dz, dgamma, dbeta from dy through act' and the normalisation (dz may alias dy)."""
    check(lib.mnr_ln_act_bwd(_ln_groups(groups), len(groups), M, N, act, ptr(ws, U8), ws.numel(), stream()), "mnr_ln_act_bwd")


def colsum(x, out, ws):
    M, Cc = x.shape
    check(lib.mnr_colsum_f32(ptr(x, F32), M, Cc, ptr(out, F32), ptr(ws, U8), ws.numel(), stream()), "mnr_colsum_f32")


# ---------------------------------------------------------------------------------------------
# a12  optimizer
# ---------------------------------------------------------------------------------------------
def optim_state_tensor(lr, device):
    """Device copy of mnr_optim_state_t (48 bytes) initialised with `lr`."""
    st = nv.OptimState(lr=float(lr), kl_sum=0.0, kl_count=0, step=0, grad_norm=0.0, clip_coef=1.0, step_size=0.0,
                       bc2_sqrt=1.0, lr_f32=float(lr), _pad=0.0)
    host = torch.frombuffer(bytearray(bytes(st)), dtype=U8).clone()
    return host.to(device)


def read_optim_state(t):
    return nv.OptimState.from_buffer_copy(bytes(t.cpu().numpy().tobytes()))


def grad_norm_prepare(flat_grad, max_norm, beta1, beta2, st, kl, norm_out, ws):
    check(lib.mnr_grad_norm_prepare(ptr(flat_grad, F32, "grad"), flat_grad.numel(), max_norm, beta1, beta2, ptr(st, U8),
                                    ptr(kl, F32, allow_none=True), ptr(norm_out, F32, allow_none=True), ptr(ws, U8), ws.numel(),
                                    stream()), "mnr_grad_norm_prepare")


def adam_step(param, grad, exp_avg, exp_avg_sq, beta1, beta2, eps, weight_decay, decoupled, zero_grad, st):
    check(lib.mnr_adam_step(ptr(param, F32), ptr(grad, F32), ptr(exp_avg, F32), ptr(exp_avg_sq, F32), param.numel(), beta1,
                            beta2, eps, weight_decay, 1 if decoupled else 0, 1 if zero_grad else 0, ptr(st, U8), stream()),
          "mnr_adam_step")


def fused_clip_adam(param, grad, exp_avg, exp_avg_sq, max_norm, beta1, beta2, eps, weight_decay, decoupled, zero_grad, st, kl,
                    norm_out, ws, nan_to_num=False, clipped_norm_out=None):
    """[grad.nan_to_num_ +] clip_grad_norm_ + Adam(W).step + zero_grad in ONE launch (csrc/peer.cu, world == 1 instance)."""
    check(lib.mnr_fused_clip_adam(ptr(param, F32), ptr(grad, F32), ptr(exp_avg, F32), ptr(exp_avg_sq, F32), param.numel(), max_norm,
                                  beta1, beta2, eps, weight_decay, 1 if decoupled else 0, 1 if zero_grad else 0, ptr(st, U8),
                                  ptr(kl, F32, allow_none=True), ptr(norm_out, F32, allow_none=True), 1 if nan_to_num else 0,
                                  ptr(clipped_norm_out, F32, allow_none=True), ptr(ws, U8), ws.numel(), stream()), "mnr_fused_clip_adam")


def critic_mse_fwd_bwd(preds, target, grads, loss_out, ws):
    """SHAC.compute_critic_loss (shac.py:869-875) + its gradient w.r.t. every critic's prediction, one launch."""
    cs = nv.CriticSet()
    for k, (p_, g_) in enumerate(zip(preds, grads)):
        cs.pred[k], cs.grad[k] = ptr(p_, F32, "pred").value, ptr(g_, F32, "grad").value
    check(lib.mnr_critic_mse_fwd_bwd(C.byref(cs), len(preds), ptr(target, F32, "target"), target.numel(), ptr(loss_out, F32),
                                     ptr(ws, U8), ws.numel(), stream()), "mnr_critic_mse_fwd_bwd")


def lr_table_update(st, lr_new, avg_kl_out=None, lr_out=None):
    """lr_new: (1,) f64 device tensor holding the learning rate that takes effect after this mini-epoch."""
    check(lib.mnr_lr_table_update(ptr(st, U8), ptr(lr_new, F64), ptr(avg_kl_out, F32, allow_none=True),
                                  ptr(lr_out, F64, allow_none=True), stream()), "mnr_lr_table_update")


def kl_lr_update(st, kl_threshold, min_lr, max_lr, avg_kl_out=None, lr_out=None):
    check(lib.mnr_kl_lr_update(ptr(st, U8), kl_threshold, min_lr, max_lr, ptr(avg_kl_out, F32, allow_none=True),
                               ptr(lr_out, F64, allow_none=True), stream()), "mnr_kl_lr_update")


# ---------------------------------------------------------------------------------------------
# 8(f)-1  rollout-side policy step
# ---------------------------------------------------------------------------------------------
def policy_step(y3a, y3v, w_mu, b_mu, w_v, b_v, w_sig, b_sig, eps, value_mean, value_denom, clamp, actions_t, neglogp_t, values_t,
                mu_t, sigma_t, actions_clamped, cfg, rows=None):
    """Heads + Dist + sample + neglogp + value un-normalise into the storage rows of one rollout step (csrc/rollout.cu)."""
    N = eps.shape[0] if rows is None else rows
    A, H = w_mu.shape
    lo, hi = clamp if clamp is not None else (1.0, -1.0)
    o = lambda t: ptr(t, F32, allow_none=True)
    check(lib.mnr_policy_step(ptr(y3a, F32, "y3a"), ptr(y3v, F32, "y3v"), ptr(w_mu, F32), ptr(b_mu, F32), ptr(w_v, F32), ptr(b_v, F32),
                              o(w_sig), ptr(b_sig, F32), ptr(eps, F32, "eps"), o(value_mean), o(value_denom), lo, hi,
                              ptr(actions_t, F32), ptr(neglogp_t, F32), ptr(values_t, F32), ptr(mu_t, F32), ptr(sigma_t, F32),
                              o(actions_clamped), N, A, H, C.byref(cfg), stream()), "mnr_policy_step")


def post_step(rewards, dones, time_outs, values_t, scale, gamma, rewards_t, dones_t):
    """Shaped reward (+ time-out value bootstrap) and dones into the storage rows of one rollout step."""
    N = rewards.numel()
    check(lib.mnr_post_step(ptr(rewards, F32, "rewards"), ptr(dones, (U8, BOOL), "dones"),
                            ptr(time_outs, (U8, BOOL), "time_outs", allow_none=True), ptr(values_t, F32, allow_none=True), scale, gamma,
                            ptr(rewards_t, F32), ptr(dones_t, (U8, BOOL)), N, stream()), "mnr_post_step")


def episode_scores(rewards, dones, cur_rewards, cur_lengths, win_rewards, win_lengths, counters):
    """Metrics._update_scores + Tracker.update on the device, one launch, no host sync (csrc/rollout.cu)."""
    check(lib.mnr_episode_scores(ptr(rewards, F32, "rewards"), ptr(dones, (U8, BOOL), "dones"), ptr(cur_rewards, F32), ptr(cur_lengths, I64),
                                 ptr(win_rewards, F32), ptr(win_lengths, F32), ptr(counters, I64), rewards.numel(), win_rewards.numel(),
                                 stream()), "mnr_episode_scores")


# ---------------------------------------------------------------------------------------------
# 8(f)-2  SAC / DDPG agents: row-wise math around the tower GEMMs (csrc/offpolicy.cu)
# ---------------------------------------------------------------------------------------------
def _critic_set(preds, grads=None):
    cs = nv.CriticSet()
    for k, p_ in enumerate(preds):
        cs.pred[k] = ptr(p_, F32, "pred").value
        if grads is not None:
            cs.grad[k] = ptr(grads[k], F32, "grad").value
    return cs


def critic_mse_sum_fwd_bwd(preds, target, grads, loss_out, ws):
    """sum_k MSE(Q_k, target) (sac.py:303, ddpg.py:296) and its gradient w.r.t. every Q_k, one launch."""
    cs = _critic_set(preds, grads)
    check(lib.mnr_critic_mse_sum_fwd_bwd(C.byref(cs), len(preds), ptr(target, F32, "target"), target.numel(), ptr(loss_out, F32),
                                         ptr(ws, U8), ws.numel(), stream()), "mnr_critic_mse_sum_fwd_bwd")


def distl_workspace(device):
    return torch.zeros(lib.mnr_distl_workspace_bytes(), dtype=U8, device=device)


def distl_target(target_logits, reward, done, support, gamma_n, v_min, v_max, n_atoms, out):
    """min_k distl_projection(softmax(target_logits_k), reward, done, gamma_n) -> out (rows, n_atoms)  (ddpg.py:275-290)."""
    cs = _critic_set(target_logits, None)
    delta_z = (v_max - v_min) / (n_atoms - 1)
    check(lib.mnr_distl_target(C.byref(cs), len(target_logits), ptr(reward, F32, "reward"), ptr(done, F32, "done"), ptr(support, F32, "support"),
                               gamma_n, v_min, v_max, delta_z, n_atoms, ptr(out, F32, "out"), reward.numel(), stream()), "mnr_distl_target")
    return out


def distl_bce_fwd_bwd(logits, target, grads, loss_out, ws, rows, n_atoms):
    """sum_k BCE(softmax(logits_k), target) and d loss / d logits_k  (ddpg.py:267-269,295-296)."""
    cs = _critic_set(logits, grads)
    check(lib.mnr_distl_bce_fwd_bwd(C.byref(cs), len(logits), ptr(target, F32, "target"), rows, n_atoms, ptr(loss_out, F32), ptr(ws, U8),
                                    ws.numel(), stream()), "mnr_distl_bce_fwd_bwd")


def distl_actor_seed(logits, support, grads, loss_out, ws, rows, n_atoms):
    """-mean(min_k E_k[z]) and its gradient w.r.t. the logits of the minimal critic  (models.py:160-163, ddpg.py:306-307)."""
    cs = _critic_set(logits, grads)
    check(lib.mnr_distl_actor_seed(C.byref(cs), len(logits), ptr(support, F32, "support"), rows, n_atoms, ptr(loss_out, F32), ptr(ws, U8),
                                   ws.numel(), stream()), "mnr_distl_actor_seed")


def make_squash_cfg(dist_type="squashed_normal", minlogstd=None, maxlogstd=None, minstd=1.0, maxstd=1.0):
    kinds = {"squashed_normal": 0, "dreamerv3_squashed_normal": 1}
    if dist_type not in kinds:
        raise nv.MineralB200Error(f"dist_type {dist_type!r} has no squashed-sample kernel")
    return nv.SquashCfg(kinds[dist_type], int(minlogstd is not None), int(maxlogstd is not None), float(minlogstd or 0.0), float(maxlogstd or 0.0),
                        float(minstd), float(maxstd), 0)


def copy_cols(src, dst):
    """dst[:, :] = src[:, :] for 2-D f32 views with unit column stride (torch.cat((state, action), 1) into a padded buffer)."""
    sp, sld = nv.ptr_strided(src, "src")
    dp, dld = nv.ptr_strided(dst, "dst")
    assert src.shape == dst.shape
    check(lib.mnr_copy_cols_f32(sp, sld, dp, dld, src.shape[0], src.shape[1], stream()), "mnr_copy_cols_f32")


def squashed_sample(mu, logstd, eps, cfg, actions, log_prob=None, rows=None):
    """SAC.get_actions after the heads: actions = tanh(mu + sigma * eps) into the (strided) `actions` view, log_prob (rows)."""
    rows = mu.shape[0] if rows is None else rows
    A = mu.shape[1]
    ap, ald = nv.ptr_strided(actions, "actions")
    check(lib.mnr_squashed_sample(ptr(mu, F32, "mu"), ptr(logstd, F32, "logstd", allow_none=eps is None), int(logstd is not None and logstd.dim() == 1),
                                  ptr(eps, F32, "eps", allow_none=True), rows, A, C.byref(cfg), ap, ald, ptr(log_prob, F32, allow_none=True),
                                  stream()), "mnr_squashed_sample")


def _second_grad(g_actions2, gld):
    if g_actions2 is None:
        return None
    gp2, gld2 = nv.ptr_strided(g_actions2, "g_actions2")
    assert gld2 == gld, "both action-gradient views must share a row stride"
    return gp2


def squashed_bwd(mu, logstd, eps, g_actions, log_alpha, alpha_host, cfg, g_mu, g_logstd, rows=None, g_actions2=None):
    rows = mu.shape[0] if rows is None else rows
    A = mu.shape[1]
    gp, gld = nv.ptr_strided(g_actions, "g_actions") if g_actions is not None else (None, 0)
    check(lib.mnr_squashed_bwd(ptr(mu, F32, "mu"), ptr(logstd, F32, "logstd"), int(logstd.dim() == 1), ptr(eps, F32, "eps"), gp,
                               _second_grad(g_actions2, gld), gld,
                               ptr(log_alpha, F32, allow_none=True), float(alpha_host), rows, A, C.byref(cfg), ptr(g_mu, F32), ptr(g_logstd, F32),
                               stream()), "mnr_squashed_bwd")


def tanh_action(mu, noise, actions, noise_bounds=None, out_bounds=None, rows=None):
    rows = mu.shape[0] if rows is None else rows
    nlo, nhi = noise_bounds if noise_bounds is not None else (1.0, -1.0)
    olo, ohi = out_bounds if out_bounds is not None else (1.0, -1.0)
    ap, ald = nv.ptr_strided(actions, "actions")
    check(lib.mnr_tanh_action(ptr(mu, F32, "mu"), ptr(noise, F32, "noise", allow_none=True), rows, mu.shape[1], nlo, nhi, olo, ohi, ap, ald,
                              stream()), "mnr_tanh_action")


def tanh_action_bwd(mu, g_actions, g_mu, rows=None, g_actions2=None):
    rows = mu.shape[0] if rows is None else rows
    gp, gld = nv.ptr_strided(g_actions, "g_actions")
    check(lib.mnr_tanh_action_bwd(ptr(mu, F32, "mu"), gp, _second_grad(g_actions2, gld), gld, rows, mu.shape[1], ptr(g_mu, F32), stream()),
          "mnr_tanh_action_bwd")


def td_target_stats(q_values, log_prob, reward, done, log_alpha, alpha_host, gamma_n, out, stats=None):
    cs = _critic_set(q_values)
    check(lib.mnr_td_target_stats(C.byref(cs), len(q_values), ptr(log_prob, F32, allow_none=True), ptr(reward, F32, "reward"), ptr(done, F32, "done"),
                                  ptr(log_alpha, F32, allow_none=True), float(alpha_host), float(gamma_n), ptr(out, F32), reward.numel(),
                                  ptr(stats, F32, allow_none=True), stream()), "mnr_td_target_stats")


def actor_seed(q_values, q_grads, log_prob, log_alpha, alpha_host, target_entropy, out4, log_alpha_grad=None):
    cs = _critic_set(q_values, q_grads)
    check(lib.mnr_actor_seed(C.byref(cs), len(q_values), ptr(log_prob, F32, allow_none=True), ptr(log_alpha, F32, allow_none=True),
                             float(alpha_host), float(target_entropy), q_values[0].numel(), ptr(out4, F32), ptr(log_alpha_grad, F32, allow_none=True),
                             stream()), "mnr_actor_seed")


def traj_store_step(obs, next_obs, actions, rewards, dones, time_outs, t, reward_scale, traj_obs, traj_next_obs, traj_actions,
                    traj_rewards, traj_dones):
    """One explore_env step into the (N, T, .) trajectory tensors (strided rows), dones masked by time-outs, rewards scaled."""
    N, T = traj_rewards.shape[0], traj_rewards.shape[1]
    O, A = obs.shape[1], actions.shape[1]
    check(lib.mnr_traj_store_step(ptr(obs, F32, "obs"), ptr(next_obs, F32, "next_obs"), ptr(actions, F32, "actions"), ptr(rewards, F32, "rewards"),
                                  ptr(dones, (U8, BOOL), "dones"), ptr(time_outs, (U8, BOOL), "time_outs", allow_none=True), N, O, A, T, t,
                                  float(reward_scale), ptr(traj_obs, F32), ptr(traj_next_obs, F32), ptr(traj_actions, F32),
                                  ptr(traj_rewards, F32), ptr(traj_dones, F32), stream()), "mnr_traj_store_step")
