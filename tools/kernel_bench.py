#!/usr/bin/env python
"""Per-kernel roofline table: every hot kernel of SURVEY.md section 8(d) timed alone at its BASELINE config shape,
L2 flushed before every launch, CUDA events on the launching stream, algorithmic bytes / time against the measured
HBM peak (MEASURED_PEAKS.json).  Writes one JSON document (stdout) that profiles/rNN_kernel_rooflines.json keeps.

    python tools/kernel_bench.py > gpurun_out/kernel_rooflines.json
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mineral_b200 import ops  # noqa: E402
from mineral_b200.agents.ppo.experience import ExperienceBuffer  # noqa: E402
from mineral_b200.buffers import NStepReplay, ReplayBuffer  # noqa: E402
from mineral_b200.common.normalizers import RunningMeanStd  # noqa: E402

DEV = "cuda:0"
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    PEAK_SRC = "measured"
except Exception:
    PEAK, PEAK_SRC = 6650.0, "fallback"
FLUSH = torch.zeros(64 << 20, dtype=torch.float32, device=DEV)


def timeit(fn, iters=16):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(iters):
        FLUSH.max()
        torch.cuda._sleep(300_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return statistics.mean(sorted(ms)[: max(1, iters * 3 // 4)])


ROWS = []


def row(name, cfg, nbytes, fn, note=""):
    ms = timeit(fn)
    gbs = nbytes / ms / 1e6
    ROWS.append(dict(kernel=name, config=cfg, algorithmic_bytes=int(nbytes), us=round(ms * 1e3, 2), gbs=round(gbs, 1), frac_of_hbm_peak=round(gbs / PEAK, 3), note=note))
    print(f"{name:34s} {cfg:28s} {nbytes/1e6:9.2f} MB {ms*1e3:8.2f} us {gbs:8.1f} GB/s {gbs/PEAK:6.3f}", file=sys.stderr, flush=True)


def r(*shape):
    return torch.randn(*shape, device=DEV)


def main():
    torch.manual_seed(0)
    # ---- K1a GAE, fused post-rollout, finalize: cfg2 (Ant 4096x32) and cfg5 per GPU (AllegroHand 16384x16) ----
    for tag, (T, N, O, A) in (("cfg2 4096x32 Ant", (32, 4096, 37, 8)), ("cfg5 16384x16 Allegro", (16, 16384, 88, 16)), ("4x cfg2 16384x32 Ant", (32, 16384, 37, 8))):
        B = T * N
        rew, val, lv = r(T, N), r(T, N), r(N)
        dones = (torch.rand(T, N, device=DEV) < 0.02).to(torch.uint8)
        ret, adv = torch.empty(T, N, device=DEV), torch.empty(T, N, device=DEV)
        row("gae_kernel", tag, 17 * B + 4 * N, lambda: ops.gae(rew, val, dones, lv, 0.99, 0.95, returns=ret, advantages=adv))
        buf = ExperienceBuffer(N, T, B, B // 4, {"obs": (O,)}, A, DEV)
        lv1 = torch.zeros(N, 1, device=DEV)
        rowb = 4 * O + 12 * A + 17
        row("post_rollout_kernel", tag, rowb * B + 4 * N + (rowb + 8) * B, lambda: buf._launch_fused_only(lv1, 0.99, 0.95))
        buf._launch_fused_only(lv1, 0.99, 0.95)
        row("finalize_stats+apply", tag, 24 * B, lambda: ops.post_rollout_finalize(buf._em, B, None, buf._ws), "2 launches")
        # K2 obs RMS: batch moments over the whole rollout (once per update) and the per-minibatch normalise
        x = r(B, O)
        mb = B // 4
        ws = ops.rms_moments_workspace(B, O, mb, DEV)
        gm, gv = torch.empty(4, O, device=DEV), torch.empty(4, O, device=DEV)
        row("batch_moments_kernel", tag, 4 * O * B, lambda: ops.rms_batch_moments(x, mb, 1, ws, gm, gv))
        mean, den, out = r(O), torch.rand(O, device=DEV) + 0.5, torch.empty(mb, O, device=DEV)
        row("rms_apply_kernel", tag + f" mb {mb}", 8 * O * mb, lambda: ops.rms_apply(x[:mb], mean, den, (-5.0, 5.0), out=out))
        # K4 loss on pre-computed head outputs (non-fused-towers path)
        mu_raw, ls, v = r(mb, A), r(A), r(mb, 1)
        d = dict(actions=r(mb, A), neglogp=r(mb) + 8, advantages=r(mb), values=r(mb, 1), returns=r(mb, 1), mu=r(mb, A), sigma=torch.rand(mb, A, device=DEV) + 0.1)
        g_mu, g_ls, g_v, met = torch.empty(mb, A, device=DEV), torch.empty(mb, A, device=DEV), torch.empty(mb, 1, device=DEV), torch.zeros(10, device=DEV)
        cfg = ops.make_loss_cfg(dist_type="normal", e_clip=0.2, critic_coef=4, bounds_loss_coef=1e-4, bounds_type="bound")
        lws = ops.ppo_loss_workspace(mb, A, DEV)
        row("ppo_loss_kernel", tag + f" mb {mb}", (12 * A + 20 + 8 * A + 4 + 8 * A) * mb,
            lambda: ops.ppo_loss_fwd_bwd(mu_raw, ls, v, d["actions"], d["neglogp"], d["advantages"], d["values"], d["returns"], d["mu"], d["sigma"],
                                         g_mu, g_ls, g_v, met, cfg, lws, True))
        del buf, x
    # ---- K1b TD-lambda at cfg3 (Humanoid 64x32) and a large shape ----
    for tag, (T, N) in (("cfg3 64x32", (32, 64)), ("65536x32", (32, 65536))):
        rew, dm, nv_, tgt = r(T, N), (torch.rand(T, N, device=DEV) < 0.02).float(), r(T, N), torch.empty(T, N, device=DEV)
        row("critic_targets_kernel", tag, 16 * T * N, lambda: ops.critic_targets(rew, dm, nv_, 0.99, 0.95, out=tgt))
    # ---- LayerNorm + activation passes at the AllegroHand minibatch ----
    M = 32768
    for N in (512, 256, 128):
        z, y, g_, b_ = r(M, N), torch.empty(M, N, device=DEV), torch.rand(N, device=DEV) + 0.5, r(N)
        mean, rstd, dy, dg, db = torch.empty(M, device=DEV), torch.empty(M, device=DEV), r(M, N), torch.empty(N, device=DEV), torch.empty(N, device=DEV)
        grp = [dict(z=z, y=y, gamma=g_, beta=b_, mean=mean, rstd=rstd, dy=dy, dz=dy, dgamma=dg, dbeta=db)]
        row("ln_act_fwd_kernel (SiLU)", f"{M}x{N}", 8 * M * N, lambda: ops.ln_act_fwd(grp, M, N, 1e-5, 3))
        lws = ops.ln_act_bwd_workspace(1, M, N, DEV)
        row("ln_act_bwd_kernel (SiLU)", f"{M}x{N}", 12 * M * N, lambda: ops.ln_act_bwd(grp, M, N, 3, lws))
    # ---- optimizer: clip + Adam over the flat buffer ----
    for n in (102_356, 427_432):
        p, g, m, v = r(n), r(n), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        st = ops.optim_state_tensor(5e-4, DEV)
        fws = torch.zeros(ops.lib.mnr_fused_adam_workspace_bytes(), dtype=torch.uint8, device=DEV)
        row("reduce_clip_adam_kernel<1>", f"{n} params", 32 * n, lambda: ops.fused_clip_adam(p, g, m, v, 1.0, 0.9, 0.999, 1e-5, 0.0, False, True, st, None, None, fws),
            "latency-bound at this size: two grid phases")
    # ---- K3c / K3d / K3e replay at cfg4 (4 M transitions, Ant) ----
    O, A, C = 37, 8, 4_000_000
    rb = ReplayBuffer({"obs": (O,)}, A, C, DEV)
    rb.buf_obs["obs"].normal_(), rb.buf_next_obs["obs"].normal_(), rb.buf_action.normal_(), rb.buf_reward.normal_(), rb.buf_done.zero_()
    rb.cur_capacity, rb.if_full = C, True
    for n in (64 * 32, 4096, 65536):
        traj = ({"obs": r(n, O)}, r(n, A), r(n, 1), {"obs": r(n, O)}, torch.zeros(n, 1, device=DEV))
        rb.next_p = C - n // 2  # every insert crosses the wrap point

        def ins():
            rb.next_p = C - n // 2
            rb.add_to_buffer(traj)
        row("insert_records_kernel (ring insert)", f"cfg4 {n} rows", 2 * (8 * O + 4 * A + 8) * n, ins,
            "algorithmic bytes as SURVEY 8(d) (666 B/row); the ring stores 384 B records")
    rms = torch.nn.ModuleDict({"obs": RunningMeanStd((O,), eps=1e-4, with_clamp=True, initial_count="eps")}).to(DEV)
    rms["obs"].update(rb.buf_obs["obs"][:65536])
    for bs in (8192, 65536):
        idx = torch.randint(C, (bs,), device=DEV)
        row("sample_records_kernel (uniform sample + normalise)", f"cfg4 batch {bs}", 677 * bs, lambda: rb.sample_batch(bs, DEV, indices=idx, obs_rms=rms),
            "random 384 B records (three 128 B lines) out of 1.54 GB, one TMA bulk copy per row; index tensor given (torch.randint excluded); "
            "algorithmic bytes as SURVEY 8(d) (677 B/row)")
    json.dump(dict(peak_gbs=PEAK, peak_source=PEAK_SRC, timing="CUDA events, 256 MiB L2 read-flush before every launch, mean of best 75%", rows=ROWS), sys.stdout, indent=1)


if __name__ == "__main__":
    main()
