#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: the PPO update (GAE + env-major pack + value/obs
normalisation + mini_epochs x minibatches of clipped-surrogate optimizer steps) on a synthetic
Ant-shaped rollout of 4096 envs x 32 steps per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo (B200 kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one full PPO update of one rollout.  `value` = samples/s with the rollout already resident
in HBM; `e2e` = the same through the agent API with the rollout in pinned HOST memory (H2D of the
rollout and D2H of the loss metrics inside the timed region).  Timing: CUDA events per step on the
launching stream, L2 flushed (256 MiB read pass) between steps, max over ranks.  Prints ONE JSON line.
"""

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] at configs[0]'s network / epoch structure (5 epochs x 4 minibatches)
    "ppo_ant_4096x32": dict(T=32, N=4096, O=37, A=8, units=[256, 128, 64], variant="elu_normal_fixed", minibatches=4, mini_epochs=5),
    # the reference's shipped Ant recipe (mineral/cfgs/agent/DFlexAntPPO.yaml): LayerNorm+SiLU towers, mb 2048
    "ppo_ant_4096x32_yaml": dict(T=32, N=4096, O=37, A=8, units=[512, 256, 128], variant="ln_silu_dreamer_learned", minibatches=64, mini_epochs=5),
    # BASELINE.json configs[4] per-GPU shape
    "ppo_allegro_16384x16": dict(T=16, N=16384, O=88, A=16, units=[512, 256, 128], variant="ln_silu_dreamer_learned", minibatches=8, mini_epochs=5),
    # Humanoid-shaped rollout (BASELINE.json configs[2] shape: obs 76, act 21) through the PPO update
    "ppo_humanoid_4096x32": dict(T=32, N=4096, O=76, A=21, units=[256, 128, 64], variant="elu_normal_fixed", minibatches=4, mini_epochs=5),
    "ppo_ant_64x32": dict(T=32, N=64, O=37, A=8, units=[256, 128, 64], variant="elu_normal_fixed", minibatches=4, mini_epochs=5),
}


def workload_config(name, w, world):
    """The `config` object of the JSON line: what defines the workload, identical in both arms (`--impl b200` / `--impl reference`)
    so that the driver can compare them key by key; implementation details (CUDA graph, exchange) go to `impl_config`."""
    B = w["T"] * w["N"]
    return {"workload": name, "envs_per_gpu": w["N"], "horizon": w["T"], "obs": w["O"], "act": w["A"], "mlp": w["units"],
            "variant": w["variant"], "mini_epochs": w["mini_epochs"], "minibatches": w["minibatches"],
            "minibatch_rows": B // w["minibatches"], "optimizer_steps_per_update": w["mini_epochs"] * w["minibatches"],
            "l2": "GPU arm: 256 MiB read-flush between timed steps; CPU arm: not applicable",
            "parallelism": f"env-sharded dp{world}"}


def make_cfg(w, device, multi_gpu, cuda_graph):
    from oracle.reference_shim import to_attr

    mb = w["T"] * w["N"] // w["minibatches"]
    if w["variant"] == "elu_normal_fixed":
        ack = dict(separate_value_mlp=True, mlp_kwargs=dict(units=w["units"], act_type="ELU"), fixed_sigma=True,
                   actor_dist_kwargs=dict(dist_type="normal"))
        optim_type, mgn = "Adam", 1.0
    else:
        ack = dict(separate_value_mlp=True, mlp_kwargs=dict(units=w["units"], norm_type="LayerNorm", act_type="SiLU"), fixed_sigma=False,
                   actor_dist_kwargs=dict(dist_type="dreamerv3_normal", minstd=0.1, maxstd=1.0))
        optim_type, mgn = "AdamW", 0.5
    return to_attr(dict(
        rl_device=device, multi_gpu=multi_gpu, task=dict(env_autoresets=True),
        agent=dict(network=dict(normalize_input=True, actor_critic="ActorCritic", actor_critic_kwargs=ack),
                   ppo=dict(num_actors=w["N"], reward_shaper=dict(fn="scale", scale=1.0), max_agent_steps=1e12, horizon_len=w["T"],
                            minibatch_size=mb, normalize_value=True, value_bootstrap=True, clip_value_loss=False, normalize_advantage=True,
                            gamma=0.99, tau=0.95, entropy_coef=0.0, e_clip=0.2, use_smooth_clamp=False, critic_coef=4,
                            bounds_loss_coef=1e-4, bounds_type="bound", mini_epochs=w["mini_epochs"], optim_type=optim_type,
                            optim_kwargs=dict(lr=5e-4, eps=1e-5), lr_schedule="kl", kl_threshold=0.008, max_grad_norm=mgn,
                            truncate_grads=True, cuda_graph=cuda_graph))))


def oracle_setup(w, seed=0, device="cpu"):
    """Parameters / optimizer / normaliser states of the oracle port for workload w (on the host, or - for the
    'reference op sequence on the GPU' baseline - on `device`)."""
    import torch

    from oracle import mineral_oracle as oc

    norm = w["variant"] != "elu_normal_fixed"
    params = oc.init_actor_critic_params(w["O"], w["A"], w["units"], norm, not norm, seed)
    params = {k: v.to(device).requires_grad_(True) for k, v in params.items()}
    if norm:
        optim = torch.optim.AdamW(list(params.values()), lr=5e-4, eps=1e-5)
        net = dict(n_layers=len(w["units"]), norm=True, act="SiLU", fixed_sigma=False, dist_type="dreamerv3_normal", minstd=0.1, maxstd=1.0)
        mgn = 0.5
    else:
        optim = torch.optim.Adam(list(params.values()), lr=5e-4, eps=1e-5)
        net = dict(n_layers=len(w["units"]), norm=False, act="ELU", fixed_sigma=True, dist_type="normal")
        mgn = 1.0
    hp = dict(gamma=0.99, tau=0.95, e_clip=0.2, critic_coef=4, entropy_coef=0.0, bounds_loss_coef=1e-4, bounds_type="bound",
              clip_value_loss=False, use_smooth_clamp=False, mini_epochs=w["mini_epochs"], minibatch_size=w["T"] * w["N"] // w["minibatches"],
              truncate_grads=True, max_grad_norm=mgn, lr_schedule="kl", kl_threshold=0.008)
    obs_rms = [t.to(device) for t in oc.rms_init((w["O"],), 1e-5, 1, torch.float64)]
    value_rms = [t.to(device) for t in oc.rms_init((1,), 1e-5, 1, torch.float64)]
    return params, optim, net, hp, obs_rms, value_rms


def time_oracle_on_gpu(w, n_updates, warmup, device):
    """The reference's op sequence (eager torch ops + autograd + torch.optim, ~490 ATen ops per optimizer step, a Python loop over
    the horizon for GAE, one .item() per mini-epoch: SURVEY.md section 8a) on the SAME GPU: what the unmodified reference would
    cost on a B200 ('also time reference-on-B200').  Baseline only - nothing of it is on the product path."""
    import torch

    from oracle import mineral_oracle as oc

    params, optim, net, hp, obs_rms, value_rms = oracle_setup(w, device=device)
    st0, last = oc.synth_rollout(w["T"], w["N"], w["O"], w["A"], seed=0)
    st0 = {k: ({kk: vv.to(device) for kk, vv in v.items()} if isinstance(v, dict) else v.to(device)) for k, v in st0.items()}
    last = last.to(device)
    lr, times = 5e-4, []
    for i in range(warmup + n_updates):
        st = {k: (dict(v) if isinstance(v, dict) else v.clone()) for k, v in st0.items()}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, lr = oc.ppo_update(st, last, params, optim, obs_rms, value_rms, hp, net, lr)
        torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def time_oracle(w, n_updates, warmup, budget_s):
    """Time the CPU port (the reference's algorithm with the reference's own torch ops) on host cores."""
    import torch

    from oracle import mineral_oracle as oc

    params, optim, net, hp, obs_rms, value_rms = oracle_setup(w)
    st0, last = oc.synth_rollout(w["T"], w["N"], w["O"], w["A"], seed=0)
    lr, times = 5e-4, []
    t_begin = time.perf_counter()
    for i in range(warmup + n_updates):
        st = {k: (dict(v) if isinstance(v, dict) else v.clone()) for k, v in st0.items()}
        t0 = time.perf_counter()
        _, _, lr = oc.ppo_update(st, last, params, optim, obs_rms, value_rms, hp, net, lr)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    return times


def reference_cfg(w):
    """The workload as the REFERENCE's own config (same keys as oracle/make_golden.py:ppo_cfg, which generated the goldens)."""
    cfg = make_cfg(w, "cpu", False, cuda_graph=False)
    cfg["env_render"], cfg["run"], cfg["seed"] = False, "train", 0
    cfg.agent.update(tracker_len=100, metrics_kwargs={}, print_every=-1, ckpt_every=-1)
    cfg.agent.ppo["multi_gpu"] = False
    cfg.agent.ppo.pop("cuda_graph", None)
    return cfg


def time_reference(w, n_updates, warmup, budget_s):
    """Time the UNMODIFIED reference (mineral.agents.ppo.PPO from oracle/_ref or /root/reference) on host cores: its own
    ExperienceBuffer.compute_return + prepare_training + value-RMS block (ppo.py:172-183, the tail of play_steps, which the
    reference cannot run without a simulator rollout in front of it: those seven lines are driven from here, every call in
    them is the reference's own method) + PPO.train_epoch() on a synthetic rollout written into its storage_dict."""
    import contextlib
    import io
    import tempfile

    import torch

    from oracle import mineral_oracle as oc
    from oracle import reference_shim as rs

    rs.install()
    from mineral.agents.ppo.ppo import PPO as RefPPO

    T, N, O, A = w["T"], w["N"], w["O"], w["A"]
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        agent = RefPPO(reference_cfg(w), logdir=tempfile.mkdtemp(prefix="mineral_ref_"), env=rs.SynthEnv(N, O, A, 1))
    st0, last = oc.synth_rollout(T, N, O, A, seed=0)
    sd = agent.storage.storage_dict
    times, t_begin = [], time.perf_counter()
    for i in range(warmup + n_updates):
        for k, v in st0.items():
            if k == "obses":
                sd["obses"]["obs"].copy_(v["obs"])
            elif k != "returns":
                sd[k].copy_(v)
        t0 = time.perf_counter()
        agent.storage.compute_return(last, agent.gamma, agent.tau)          # ppo.py:172
        agent.storage.prepare_training()                                     # ppo.py:173
        values, returns = agent.storage.data_dict["values"], agent.storage.data_dict["returns"]
        if agent.normalize_value:                                            # ppo.py:175-183
            agent.value_rms.update(values)
            values = agent.value_rms.normalize(values)
            agent.value_rms.update(returns)
            returns = agent.value_rms.normalize(returns)
        agent.storage.data_dict["values"], agent.storage.data_dict["returns"] = values, returns
        agent.set_train()
        agent.train_epoch()                                                  # ppo.py:248-350
        agent.storage.data_dict = None                                       # ppo.py:199
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    return times


def shrink_for_cpu(w, target_s_per_update=4.0):
    """Bounded CPU sample: same workload structure, fewer envs if one full update would take too long."""
    probe = dict(w, N=max(64, w["N"] // 16))
    t = time_oracle(probe, 1, 1, 60)[0]
    est_full = t * w["N"] / probe["N"]
    if est_full <= target_s_per_update:
        return dict(w), est_full
    n = max(64, int(w["N"] * target_s_per_update / est_full) // 64 * 64)
    return dict(w, N=n), est_full * n / w["N"]


def multi_gpu_parity(agent, rank, world, dev):
    """What the driver cannot otherwise see at N > 1 ('N ranks on env shards == 1 rank on the union'):
    (1) after the timed loop every rank holds BIT-identical parameters; (2) a small env-sharded update (T 8 x 64 envs per rank,
    obs 11, act 3, towers [32,16] LayerNorm+SiLU, learned sigma, 2 epochs x 4 minibatches, AdamW, KL schedule) run on `world`
    ranks over the peer-memory kernels equals the same update run by rank 0 alone on the concatenated rollout whose
    minibatch i is the union of the ranks' local minibatch i.  Same construction as tests/test_gpu_multirank.py."""
    import torch
    import torch.distributed as dist

    from mineral_b200.agents import PPO
    from oracle import mineral_oracle as oc
    from oracle.reference_shim import SynthEnv, to_attr

    flat = agent.optim.flat
    every = torch.empty(world, flat.numel(), dtype=flat.dtype, device=dev)
    dist.all_gather_into_tensor(every, flat.contiguous())
    identical = bool(all(torch.equal(every[0], every[r]) for r in range(1, world)))

    T, NL, O, A, NUM_MB, EPOCHS = 8, 64, 11, 3, 4, 2

    def cfg(n_envs, mb, multi):
        return to_attr(dict(
            rl_device=dev, multi_gpu=multi, task=dict(env_autoresets=True),
            agent=dict(network=dict(normalize_input=True, actor_critic="ActorCritic",
                                    actor_critic_kwargs=dict(mlp_kwargs=dict(units=[32, 16], norm_type="LayerNorm", act_type="SiLU"), fixed_sigma=False,
                                                             actor_dist_kwargs=dict(dist_type="dreamerv3_normal", minstd=0.1, maxstd=1.0))),
                       ppo=dict(num_actors=n_envs, reward_shaper=dict(fn="scale", scale=1.0), max_agent_steps=1e9, horizon_len=T, minibatch_size=mb,
                                normalize_value=True, value_bootstrap=True, clip_value_loss=True, normalize_advantage=True, gamma=0.99, tau=0.95,
                                entropy_coef=0.001, e_clip=0.2, use_smooth_clamp=False, critic_coef=4, bounds_loss_coef=1e-4, bounds_type="bound",
                                mini_epochs=EPOCHS, optim_type="AdamW", optim_kwargs=dict(lr=5e-4, eps=1e-5), lr_schedule="kl", kl_threshold=0.008,
                                max_grad_norm=0.5, truncate_grads=True, cuda_graph=False))))

    def take(st, last, envs):
        return {k: ({"obs": v["obs"][:, envs]} if k == "obses" else v[:, envs]) for k, v in st.items()}, last[envs]

    def load(a, st):
        for k, v in st.items():
            if k == "obses":
                a.storage.storage_dict["obses"]["obs"].copy_(v["obs"])
            elif k != "returns":
                a.storage.storage_dict[k].copy_(v)

    def run(a, last):
        a.finish_rollout(last.to(dev))
        a.set_train()
        res = a.train_epoch()
        torch.cuda.synchronize()
        return dict(loss=torch.stack(res["loss/total"]).cpu(), gn=torch.stack(res["grad_norm/all"]).cpu(), kl=torch.stack(res["avg_kl"]).cpu(),
                    params=a.optim.flat.cpu(), obs_mean=a.obs_rms["obs"].running_mean.cpu(), val_var=a.value_rms.running_var.cpu(), lr=a.last_lr)

    st, last = oc.synth_rollout(T, NL * world, O, A, seed=3, done_p=0.05)
    torch.manual_seed(1234)
    shard = PPO(cfg(NL, NL * T // NUM_MB, True), env=SynthEnv(4, O, A, 1))
    dist.broadcast(shard.optim.flat, 0)
    init_flat = shard.optim.flat.clone()
    st_r, last_r = take(st, last, torch.arange(rank * NL, (rank + 1) * NL))
    load(shard, st_r)
    got = run(shard, last_r)
    err = shard.comm.error()
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(got, comm_error=err))
    out = {"ranks_bit_identical_after_timed_loop": identical, "world": world}
    if rank == 0:
        single = PPO(cfg(NL * world, world * NL * T // NUM_MB, False), env=SynthEnv(4, O, A, 1))
        single.optim.flat.copy_(init_flat)
        E = NL // NUM_MB
        order = torch.cat([torch.arange(r * NL + i * E, r * NL + (i + 1) * E) for i in range(NUM_MB) for r in range(world)])
        st_u, last_u = take(st, last, order)
        load(single, st_u)
        ref = run(single, last_u)

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

        worst = {k: max(rel(g[k], ref[k]) for g in gathered) for k in ("loss", "gn", "params", "obs_mean", "val_var")}
        kl_abs = max(float((g["kl"] - ref["kl"]).abs().max()) for g in gathered)
        ok = (worst["loss"] <= 2e-5 and worst["gn"] <= 2e-5 and worst["params"] <= 1e-4 and worst["obs_mean"] <= 1e-6 and worst["val_var"] <= 1e-6
              and kl_abs <= 1e-5 * float(ref["kl"].abs().max()) + 2e-7 and all(g["lr"] == ref["lr"] for g in gathered)
              and all(torch.equal(g["params"], gathered[0]["params"]) for g in gathered) and all(g["comm_error"] == 0 for g in gathered))
        out["union_check"] = {"shape": f"{world} ranks x ({T} steps x {NL} envs), obs {O}, act {A}, [32,16] LayerNorm+SiLU, {EPOCHS} epochs x {NUM_MB} minibatches",
                              "max_rel_diff_vs_single_rank_on_union": worst, "avg_kl_max_abs_diff": kl_abs,
                              "lr_decisions_equal": all(g["lr"] == ref["lr"] for g in gathered),
                              "shard_ranks_bit_identical": all(torch.equal(g["params"], gathered[0]["params"]) for g in gathered),
                              "comm_errors": [g["comm_error"] for g in gathered], "pass": bool(ok)}
        out["pass"] = bool(identical and ok)
    shard.comm.close()
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_ev = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_reference(args, w, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, all threads: the
    UNMODIFIED reference package (oracle/_ref, staged by oracle/build_ref.py; kind "reference") when it is there, else the
    oracle port of its algorithm over the same torch ops (oracle/mineral_oracle.py; kind "port")."""
    if rank != 0:
        return
    import torch

    from oracle import reference_shim as rs

    # all the host threads this process may use - also under torchrun, which exports OMP_NUM_THREADS=1 to every rank
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    cores = torch.get_num_threads()
    ws, est = shrink_for_cpu(w)
    kind = "port"
    if rs.reference_available():
        try:
            times = time_reference(ws, args.steps, args.warmup, budget_s=240)
            kind = "reference"
        except Exception as e:  # an import problem on the box must not take the arm down: fall back to the port and say so
            print(f"reference arm: real reference unavailable ({e!r}); timing the oracle port", file=sys.stderr)
    if kind == "port":
        times = time_oracle(ws, args.steps, args.warmup, budget_s=240)
    per = statistics.mean(times)
    samples = ws["T"] * ws["N"]
    value = samples / per
    what = ("the unmodified reference's PPO agent (mineral.agents.ppo.PPO from oracle/_ref)" if kind == "reference"
            else "the oracle port of the reference's algorithm")
    sample = (f"{len(times)} timed full updates of {ws['N']} envs x {ws['T']} steps ({samples} samples, {ws['mini_epochs']} epochs x "
              f"{ws['minibatches']} minibatches) by {what} on {cores} host threads"
              + ("" if ws["N"] == w["N"] else f"; envs reduced from {w['N']} to bound CPU time"))
    print(json.dumps({
        "impl": "reference", "metric": "ppo_update_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, w, args.gpus),
        "impl_config": {"device": "cpu", "host_threads": cores, "envs_timed": ws["N"]},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=_RESULT_OUT, flush=True)


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ppo_ant_4096x32", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying one CUDA graph per update")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-sweep", action="store_true", help="skip timing the fused post-rollout kernel over rollout sizes")
    ap.add_argument("--no-recipes", action="store_true", help="skip the second entry (the reference's shipped yaml recipe)")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
    # communicator): keep the real stdout aside for the result line and point fd 1 at stderr for everything else.
    global _RESULT_OUT
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    w = WORKLOADS[args.workload]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, w, rank, world)

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: mineral_b200 has no CPU path"
    torch.cuda.set_device(local_rank)
    numa = None
    try:   # pin this rank's host threads (and so its pinned staging buffers) to the CPUs next to its GPU: with N ranks uploading
        import pynvml   # their rollouts at once, a remote-socket staging buffer is what the e2e number pays for
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        numa = len(os.sched_getaffinity(0))
    except Exception:
        pass
    dev = f"cuda:{local_rank}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    from mineral_b200 import _native as nv
    from mineral_b200.agents import PPO
    from oracle import mineral_oracle as oc
    from oracle.reference_shim import SynthEnv

    T, N, O, A = w["T"], w["N"], w["O"], w["A"]
    B = T * N
    torch.manual_seed(0)
    agent = PPO(make_cfg(w, dev, world > 1, cuda_graph=not args.no_graph), env=SynthEnv(4, O, A, 1))
    agent.set_train()
    if world > 1:  # identical initial weights on every rank
        dist.broadcast(agent.optim.flat, 0)

    # synthetic rollout(s): a different seed per rank (each rank owns its own envs)
    st, last = oc.synth_rollout(T, N, O, A, seed=rank)
    host = {k: v for k, v in st.items() if k not in ("returns", "obses")}
    host["obs"] = st["obses"]["obs"]
    host = {k: v.pin_memory() for k, v in host.items()}
    host_last = last.pin_memory()
    dev_last = last.to(dev)
    sd = agent.storage.storage_dict

    def load_rollout(non_blocking=False):
        for k, v in host.items():
            (sd["obses"]["obs"] if k == "obs" else sd[k]).copy_(v, non_blocking=non_blocking)

    load_rollout()
    flush_buf = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    class _Flush:
        """L2 flush between timed iterations: READ 256 MiB (evicts our data and leaves only CLEAN lines
        behind; a write-flush would leave the L2 full of dirty lines that the timed kernels then have to
        write back, charging them for traffic that is not theirs)."""

        @staticmethod
        def zero_():
            flush_buf.max()

    flush = _Flush()

    def one_update():
        agent.finish_rollout(dev_last)
        return agent.train_epoch()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also builds the CUDA graph) ----
    for _ in range(args.warmup):
        one_update()
    barrier()

    # ---- timed: K updates, HBM-resident inputs ----
    nv.LAUNCHES = 0
    replays0 = agent._engine.replays
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s0, s1 in evs:
        flush.zero_()  # L2 flush between timed iterations (not timed)
        s0.record()
        one_update()
        s1.record()
    barrier()
    clocks = sampler.stop()
    launches = nv.LAUNCHES + agent._engine.graph_launches * (agent._engine.replays - replays0)
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = total_ms.item()
    value = B * world * args.steps / (total_ms / 1e3)

    # ---- e2e: rollout in pinned host memory, H2D + update + D2H of the metrics inside the timed region ----
    h2d = sum(v.numel() * v.element_size() for v in host.values()) + host_last.numel() * 4
    e_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    d2h = 0

    # Every step uploads ITS rollout from pinned host memory (ExperienceBuffer.stage_from_host: copy stream -> staging buffers),
    # commits it (device-to-device into the storage the update graph reads), runs the update and reads the metrics back.  The
    # upload of step i+1 is started right after step i's commit, so the copy engines work under step i's update: each H2D
    # starts and ends inside a timed interval (the first one entirely inside step 0's, un-overlapped).
    def e2e_stage():
        agent.storage.stage_from_host(host)

    # The metrics of step i are snapshotted on the compute stream and copied to pinned host memory on a side stream; the host
    # picks them up one step later (a training loop logs asynchronously), so it never blocks on the GPU and launches step i+1
    # while step i is still running.  Every step's result does land on the host inside the timed loop (the last one inside
    # the last interval).
    res_host = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
    res_done = [torch.cuda.Event(), torch.cuda.Event()]
    d2h_stream = torch.cuda.Stream(device=dev)
    snap_ready = torch.cuda.Event()
    seen = []

    def e2e_step(i, stage_next):
        agent.storage.commit_staged()
        dev_last.copy_(host_last, non_blocking=True)
        if stage_next:
            e2e_stage()
        res = one_update()
        snap = torch.stack([torch.stack(res[k]).mean() for k in ("loss/total", "loss/actor", "loss/critic", "avg_kl")])
        snap_ready.record()
        d2h_stream.wait_event(snap_ready)
        with torch.cuda.stream(d2h_stream):
            res_host[i & 1].copy_(snap, non_blocking=True)
            res_done[i & 1].record(d2h_stream)
        snap.record_stream(d2h_stream)

    def e2e_collect(i):
        res_done[i & 1].synchronize()
        seen.append(res_host[i & 1].clone())

    for w_i in range(2):  # untimed: the first pass through the result read-back pays one-off lazy initialisation (~100 ms)
        e2e_stage()
        e2e_step(w_i, False)
        e2e_collect(w_i)
    seen.clear()
    barrier()
    n_steps = len(e_evs)
    for i, (s0, s1) in enumerate(e_evs):
        flush.zero_()
        s0.record()
        if i == 0:
            e2e_stage()
        e2e_step(i, i + 1 < n_steps)
        if i > 0:
            e2e_collect(i - 1)       # the previous step's metrics, now on the host
        if i + 1 == n_steps:
            e2e_collect(i)           # the last step's: wait for it inside its own interval
        s1.record()
    barrier()
    assert len(seen) == n_steps and all(torch.isfinite(t).all() for t in seen), "e2e: a step's metrics did not reach the host"
    d2h = 4 * 4
    e_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in e_evs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / (e_ms.item() / 1e3)

    # ---- N > 1: where the fused exchange kernel spends its time (MNR_PEER_TRACE=1: globaltimer stamps of CTA 0) ----
    peer_breakdown = None
    if world > 1 and getattr(agent.optim, "trace", None) is not None:
        tr = agent.optim.trace.cpu().view(64, 8).double()
        ok = (tr[:, 5] > tr[:, 0]) & (tr[:, 0] > 0)
        if ok.any():
            d = tr[ok]
            peer_breakdown = {"steps_sampled": int(ok.sum()), "unit": "us", "rank": rank,
                              "start_barrier_wait": float((d[:, 1] - d[:, 0]).mean() / 1e3), "peer_reads_and_norm": float((d[:, 2] - d[:, 1]).mean() / 1e3),
                              "grid_barrier": float((d[:, 3] - d[:, 2]).mean() / 1e3), "adam": float((d[:, 4] - d[:, 3]).mean() / 1e3),
                              "end_barrier_and_clear": float((d[:, 5] - d[:, 4]).mean() / 1e3), "whole_kernel": float((d[:, 5] - d[:, 0]).mean() / 1e3),
                              "nvlink_bytes_read_per_step": (world - 1) * (agent.optim.flat.numel() + 32) * 4,
                              "how": "%globaltimer of CTA 0, thread 0 inside reduce_clip_adam_kernel<0> over the last 64 optimizer steps; ncu cannot wrap "
                                     "a multi-rank command (its kernel replay would re-enter the cross-GPU spin barriers)"}

    # ---- N > 1: numerical self-check of the multi-GPU path (untimed) ----
    mgp = None
    if world > 1:
        try:
            mgp = multi_gpu_parity(agent, rank, world, dev)
        except Exception as e:   # reported, never silently dropped
            mgp = {"pass": False, "error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the reference's SHIPPED Ant recipe as a second entry (mineral/cfgs/agent/DFlexAntPPO.yaml:17-21,33-66: towers
    #      [512,256,128] LayerNorm+SiLU, learned sigma, 64 minibatches of 2048 rows x 5 epochs = 320 optimizer steps) ----
    recipes = None
    if world == 1 and not args.no_recipes and args.workload == "ppo_ant_4096x32":
        recipes = []
        for name in ("ppo_ant_4096x32_yaml",):
            try:
                wr = WORKLOADS[name]
                ar = PPO(make_cfg(wr, dev, False, cuda_graph=not args.no_graph), env=SynthEnv(4, wr["O"], wr["A"], 1))
                ar.set_train()
                str_, lastr = oc.synth_rollout(wr["T"], wr["N"], wr["O"], wr["A"], seed=0)
                for k, v in str_.items():
                    if k == "obses":
                        ar.storage.storage_dict["obses"]["obs"].copy_(v["obs"])
                    elif k != "returns":
                        ar.storage.storage_dict[k].copy_(v)
                lastr = lastr.to(dev)
                for _ in range(3):
                    ar.finish_rollout(lastr); ar.train_epoch()
                torch.cuda.synchronize()
                tsr = []
                for _ in range(5):
                    flush.zero_()
                    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ea.record(); ar.finish_rollout(lastr); ar.train_epoch(); eb.record(); torch.cuda.synchronize()
                    tsr.append(ea.elapsed_time(eb))
                msr = statistics.mean(tsr)
                steps_r = wr["mini_epochs"] * wr["minibatches"]
                recipes.append({"workload": name, "value": wr["T"] * wr["N"] / (msr * 1e-3), "unit": "samples/s", "ms_per_update": msr,
                                "optimizer_steps_per_update": steps_r, "us_per_optimizer_step": msr * 1e3 / steps_r,
                                "minibatch_rows": wr["T"] * wr["N"] // wr["minibatches"], "mlp": wr["units"], "variant": wr["variant"],
                                "timing": "5 timed updates after 3 warm-ups, CUDA events, L2 read-flushed between updates"})
                del ar
            except Exception as e:
                recipes.append({"workload": name, "unavailable": repr(e)[:200]})

    # ---- roofline of the dominant HBM-bound kernel: the fused post-rollout pass ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    row = 4 * O + 12 * A + 17
    alg_bytes = row * B + 4 * N + (row + 8) * B  # read everything once, write env-major once (+ returns, adv)

    def time_fused(agent_, n_iter=20):
        ts = []
        for _ in range(n_iter):
            flush.zero_()
            torch.cuda._sleep(400_000)  # ~0.2 ms of GPU-side spin: the CPU enqueues event+kernel+event meanwhile,
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)  # so no launch latency leaks in
            a.record()
            agent_.storage._launch_fused_only(dev_last, agent_.gamma, agent_.tau)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.mean(sorted(ts)[: max(1, len(ts) * 3 // 4)])

    k_ms_single = time_fused(agent)

    def time_fused_rotating(n_buf=4, n_launch=24):
        """Average launch duration over back-to-back launches that rotate over `n_buf` rollouts (n_buf x (in + out) =
        4 x 69 MB at cfg2 > 126 MB L2, so every launch finds its inputs evicted): one event pair around the whole batch,
        so the ~5 us of event-record + launch latency that a single 15 us launch carries is amortised."""
        from mineral_b200.agents.ppo.experience import ExperienceBuffer
        bufs = [agent.storage]
        for _ in range(n_buf - 1):
            b = ExperienceBuffer(N, T, B, B // w["minibatches"], {"obs": (O,)}, A, dev)
            for k, v in agent.storage.storage_dict.items():
                if isinstance(v, dict):
                    for kk, vv in v.items():
                        b.storage_dict[k][kk].copy_(vv)
                else:
                    b.storage_dict[k].copy_(v)
            bufs.append(b)
        for b in bufs:
            b._launch_fused_only(dev_last, agent.gamma, agent.tau)
        ts = []
        for _ in range(5):
            flush.zero_()
            torch.cuda._sleep(4_000_000)  # ~2 ms: the CPU enqueues the whole batch behind this spin
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(n_launch):
                bufs[i % n_buf]._launch_fused_only(dev_last, agent.gamma, agent.tau)
            e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e) / n_launch)
        return statistics.median(ts)

    k_ms = time_fused_rotating()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"kernel": "post_rollout_kernel<true> (GAE + env-major pack + moment partials)", "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes": alg_bytes, "launch_ms": k_ms,
                "timing": "CUDA events on the launching stream around 24 back-to-back launches rotating over 4 rollouts (276 MB > L2: "
                          "every launch reads cold inputs); average per launch",
                "single_launch": {"launch_ms": k_ms_single, "achieved": alg_bytes / (k_ms_single * 1e-3) / 1e9,
                                  "frac": alg_bytes / (k_ms_single * 1e-3) / 1e9 / peak,
                                  "timing": "one launch per event pair, L2 flushed before it: includes ~5 us of event-record and "
                                            "launch latency (an empty launch of the same grid measures 6.2 us this way)"}}
    # traffic of the same kernel from the committed ncu --set full capture (profiles/), per launch
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "post_rollout_traffic.json")))
        if prof.get("workload") == args.workload:
            roofline["traffic"] = prof["dram_bytes_read"] + prof["dram_bytes_write"]
            roofline["traffic_source"] = prof["source"]
    except Exception:
        pass

    # ---- roofline of the DOMINANT kernel: the tcgen05 3xTF32 linear kernel (tc_linear_kernel<FWD|DX|DW>), all the tower
    #      GEMMs of one optimizer step (both towers per launch): L forward + L weight-gradient + (L - 1) input-gradient ----
    roofline_tensor = None
    tw = getattr(agent._engine, "towers", None)
    if tw is not None and tw.use_tc:
        from mineral_b200 import ops
        eng = agent._engine
        x_buf = getattr(eng, f"_n_{eng.norm_keys[0]}")["x"] if eng.norm_keys else eng._raw_x
        x_ld = x_buf.shape[1]
        widths, in_dim, M, L = tw.widths, tw.in_dim, tw.M, tw.L
        ks = [in_dim] + widths[:-1]
        f_fwd = sum(2.0 * 2 * M * k * n for k, n in zip(ks, widths))          # 2 towers
        f_dx = sum(2.0 * 2 * M * k * n for k, n in zip(ks[1:], widths[1:]))
        flops = 2 * f_fwd + f_dx                                               # forward + dW + dX
        fused_fwd = bool(getattr(tw.stack, "fused_fwd", False))   # whole-tower forward kernel (csrc/tc_tower.cu): one launch
        n_launch = (1 if fused_fwd else L) + L + (L - 1)
        # The same launches against the HBM roofline: what every launch must move when each layer is its own kernel (operands read
        # once, result written once, f32): forward X + all Y_l; dW_l reads dY_l, Y_l (act' from the saved output; not for a top layer
        # that receives dZ from the heads kernel, nor behind a LayerNorm) and Y_{l-1}; dX_l reads dY_l, Y_l and writes dY_{l-1}.
        G2 = 2
        top_dz = bool(getattr(tw.stack, "top_is_dz", False))
        by = 4.0 * M * (x_ld + G2 * sum(widths))
        for li in range(L):
            saved = 0 if (tw.normed or (li == L - 1 and top_dz)) else widths[li]
            k_in = x_ld if li == 0 else widths[li - 1]
            by += 4.0 * M * G2 * (widths[li] + saved + k_in) - (4.0 * M * x_ld if li == 0 else 0)   # X is shared by both towers
            if li > 0:
                by += 4.0 * M * G2 * (widths[li] + saved + widths[li - 1])
        # the split-partials reduce (tc_reduce_kernel) is not a tensor-core kernel: it is taken out of the timed sequence; behind a
        # LayerNorm the LN passes (mnr_ln_act_fwd / _bwd) sit between the GEMMs and are timed separately and subtracted
        real_reduce = ops.tc_reduce_partials
        tw.forward_towers(x_buf, x_ld, split=True)

        def gemm_set(with_ln=True):
            tw.forward_towers(x_buf, x_ld, split=False)
            tw.backward_towers(x_buf)

        def timed(fn, n_iter=12):
            ts = []
            for _ in range(n_iter):
                flush.zero_()
                torch.cuda._sleep(2_000_000)   # ~1 ms of GPU-side spin: the CPU enqueues the whole set behind it
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record(); fn(); eb.record(); torch.cuda.synchronize()
                ts.append(ea.elapsed_time(eb))
            return statistics.mean(sorted(ts)[: max(1, len(ts) * 3 // 4)])

        ops.tc_reduce_partials = lambda pending, folds=None: None
        try:
            t_ms = timed(gemm_set)
            t_ln = 0.0
            if tw.normed:   # the LayerNorm + activation passes alone, same order
                st = tw.stack

                def ln_only():
                    G = st.G
                    for i in range(L):
                        ops.ln_act_fwd([dict(z=st.Z[i][g], y=st.Y[i][g], gamma=st.norm[g][i].weight.data, beta=st.norm[g][i].bias.data,
                                             mean=st.mean[i][g], rstd=st.rstd[i][g]) for g in range(G)], M, widths[i], st.norm[0][i].eps, st.act)
                    for i in range(L - 1, -1, -1):
                        ops.ln_act_bwd([dict(z=st.Z[i][g], gamma=st.norm[g][i].weight.data, beta=st.norm[g][i].bias.data, mean=st.mean[i][g],
                                             rstd=st.rstd[i][g], dy=st.dY[i][g], dz=st.dY[i][g], dgamma=st.norm[g][i].weight.grad,
                                             dbeta=st.norm[g][i].bias.grad) for g in range(G)], M, widths[i], st.act, st.ln_ws[i])
                t_ln = timed(ln_only)
        finally:
            ops.tc_reduce_partials = real_reduce
        agent.optim.zero_grad()
        t_gemm = max(t_ms - t_ln, 1e-6)
        tpeak = peaks.get("bf16_tflops", 1590.0)
        tpeak_src = "measured burst (MEASURED_PEAKS.json bf16_tflops)" if "bf16_tflops" in peaks else "fallback (B200_PROFILING.md)"
        roofline_tensor = {
            "kernel": ("tc_tower_fwd_kernel + tc_dw2_kernel / tc_linear_kernel<DX|DW>" if fused_fwd else "tc_linear_kernel<FWD|DX|DW> + tc_dw2_kernel") +
                      f" (tcgen05 kind::tf32, TMEM accumulators, TMA operands): the {n_launch} tower GEMM launches of one "
                      f"optimizer step, 2 towers x ({M} rows, {in_dim}->" + "->".join(str(x) for x in widths) + ")",
            "bound": "tensor", "achieved": flops / (t_gemm * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
            "frac": flops / (t_gemm * 1e-3) / 1e12 / tpeak, "traffic": None, "peak_source": tpeak_src,
            "algorithmic_flops": flops, "launches": n_launch, "set_ms": t_gemm, "launch_ms": t_gemm / n_launch,
            "timing": "CUDA events on the launching stream around the whole forward + backward GEMM sequence of one optimizer step (the split-"
                      "partials reduce taken out" + ("; the LayerNorm passes timed alone in the same order and subtracted" if tw.normed else "") +
                      "), L2 read-flushed before every repetition; mean of the fastest 9 of 12",
            "note": "achieved counts fp32-EQUIVALENT FLOPs (2MNK per GEMM with the real K); the tensor pipe executes 3x that in tf32 (3xTF32 "
                    "split: the 1e-5 parity bar rules out single-pass tf32 / bf16), and tf32 runs at half the bf16 rate used as `peak`, so a "
                    "perfect 3xTF32 kernel reads frac = 1/6 = 0.167",
            "frac_of_3xtf32_ceiling": flops / (t_gemm * 1e-3) / 1e12 / (tpeak / 6.0),
            # the same launches are HBM-bound at this network size (27 fp32 FLOP per byte, ridge of the 3xTF32 pipe ~ 40): their
            # per-layer compulsory traffic over the same time, against the measured copy bandwidth
            "hbm_view": {"bound": "hbm", "algorithmic_bytes": by, "achieved": by / (t_gemm * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": by / (t_gemm * 1e-3) / 1e9 / peak,
                         "what": "sum over the launches of operands read once + result written once (f32); layer outputs are saved for "
                                 "the backward pass, so every layer boundary crosses HBM unless L2 still holds it"}}

        # DRAM traffic of the same six launches from the committed ncu --set full capture (profiles/), per step like `achieved`
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r02_gemm_set_traffic.json")))
            if prof.get("workload") == args.workload and fused_fwd:
                roofline_tensor["traffic"] = prof["dram_bytes_read"] + prof["dram_bytes_write"]
                roofline_tensor["traffic_source"] = prof["source"]
        except Exception:
            pass

    # ---- share of the update each kernel takes, from the committed ncu launch list of this command (profiles/) ----
    shares = None
    try:
        sh = json.load(open(os.path.join(ROOT, "profiles", "r02_shares.json")))
        if sh.get("workload") == args.workload:
            shares = sh
    except Exception:
        pass

    sweep = None
    if not args.no_kernel_sweep:
        from mineral_b200.agents.ppo.experience import ExperienceBuffer
        sweep = []
        for n_env in (4096, 16384, 65536):
            buf = ExperienceBuffer(n_env, T, n_env * T, n_env * T // 4, {"obs": (O,)}, A, dev)
            lv = torch.zeros(n_env, 1, device=dev)

            class _A:  # minimal holder for time_fused
                storage, gamma, tau = buf, 0.99, 0.95
            torch.cuda.synchronize()
            ms = []
            for _ in range(12):
                flush.zero_()
                torch.cuda._sleep(400_000)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); buf._launch_fused_only(lv, 0.99, 0.95); b.record(); torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            m = statistics.mean(sorted(ms)[:9])
            byts = row * n_env * T + 4 * n_env + (row + 8) * n_env * T
            sweep.append({"envs": n_env, "ms": m, "GB/s": byts / m / 1e6, "frac": byts / m / 1e6 / peak})

    # ---- the reference's op sequence as eager torch on this GPU (reported baseline, not the target) ----
    torch_gpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            tg = time_oracle_on_gpu(w, 3, 2, dev)
            per_g = statistics.mean(tg)
            torch_gpu = {"value": B / per_g, "unit": "samples/s", "ms_per_step": per_g * 1e3, "kind": "port",
                         "what": "oracle/mineral_oracle.py:ppo_update with every tensor on cuda:0 - the reference's eager torch op "
                                 "sequence (autograd, torch.optim, Python GAE loop) on the same B200; 3 timed updates after 2 warm-ups"}
        except Exception as e:  # a baseline must never take the bench line down
            torch_gpu = {"unavailable": repr(e)[:200]}

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N = 1 only (the reference arm covers every N)
        cores = torch.get_num_threads()
        ws, _ = shrink_for_cpu(w)
        times = time_oracle(ws, 3, 1, budget_s=45)
        per = statistics.mean(times)
        cpu = {"value": ws["T"] * ws["N"] / per, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{len(times)} full updates of {ws['N']} envs x {ws['T']} steps after 1 warm-up, {cores} torch threads, {per:.2f} s each"}

    line = {
        "metric": "ppo_update_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, w, world),
        "impl_config": {"cuda_graph": agent._engine.graph is not None,
                        "exchange": "none (1 GPU)" if world == 1 else "fused peer-memory all-reduce + clip + Adam kernel over NVLink (csrc/peer.cu)"},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "pipeline": "every step uploads its own rollout from pinned host memory (ExperienceBuffer.stage_from_host -> commit_staged) and "
                            "copies its metrics to pinned host memory; the upload of step i+1 runs on the copy engines under step i's update and "
                            "the host collects step i's metrics one step later, so it never blocks on the GPU"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # `roofline` = the dominant kernel of the step by time (the tcgen05 tower GEMMs, ~70 % of an update: see `shares`);
        # `roofline_hbm` = the dominant HBM-bound kernel (the fused post-rollout pass that BASELINE.json's second metric names)
        "roofline": roofline_tensor if roofline_tensor is not None else roofline,
        "roofline_hbm": roofline,
        "shares": shares,
        "cpu_baseline": cpu,
        "torch_ops_on_gpu": torch_gpu,
        "gae_normalise_hbm_gbs": achieved,
    }
    if sweep:
        line["post_rollout_sweep"] = sweep
    if recipes:
        line["recipes"] = recipes
    if mgp is not None:
        line["multi_gpu_parity"] = mgp
    if peer_breakdown is not None:
        line["peer_kernel_breakdown"] = peer_breakdown
    if world > 1:
        line["impl_config"]["exchange_barriers_per_step"] = 2 if os.environ.get("MNR_PEER_END_BARRIER", "0") == "1" else 1
        line["impl_config"]["host_threads_pinned_near_gpu"] = numa
    print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
